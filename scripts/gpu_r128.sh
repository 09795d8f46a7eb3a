#!/bin/bash
# parity tests + the radius-128 column of the config-5 sweep + default lines. Usage: gpu_r128.sh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -6
for v in "--workload C3 --dist uniform --radius 128 --biomes 16 --chunks 32" "--workload C3 --dist uniform --radius 128 --biomes 64 --chunks 32" "--workload C3 --dist blocky --radius 128 --biomes 64 --chunks 32" "--workload C3 --dist blocky --radius 128 --biomes 1024 --chunks 32" "--workload C3 --dist uniform --radius 128 --biomes 256 --chunks 8" "--workload C3 --dist uniform" "--workload C3 --dist blocky" "--workload C1 --dist uniform"; do
  timeout 600 python bench.py $v --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("$v | %.0f Mpx/s step %.3f ms emit %.3f events %.3f vscan %.3f bins/px %.2f plan %s" % (d["value"], d["ms_per_step"], d["phases_ms"]["emit"], d["phases_ms"]["events"], d["phases_ms"]["remap_vscan"], d["config"]["bins_per_pixel"], d["config"]["plan"]))
except Exception as e:
    print("$v | failed", e); print(open("gpurun_out/ab.err").read()[-400:])
PY
done
