#!/bin/bash
timeout 900 python -m pytest tests/test_heightfield_gpu.py -q -x 2>&1 | tail -3
python bench.py --workload C4 --dist blocky --heightfield --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/hf.json 2> gpurun_out/hf.err; python -c "
import json; d=json.load(open('gpurun_out/hf.json')); print(d['value'], d['ms_per_step'], d['consumer'])"
