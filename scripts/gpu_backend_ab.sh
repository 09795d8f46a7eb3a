#!/bin/bash
# barrier/max-reduce plumbing over gloo vs nccl on N GPUs, and the same bench single-process on GPU 0
N=${1:-2}
mkdir -p gpurun_out
for be in gloo nccl; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --backend $be > gpurun_out/ab_$be.json 2> gpurun_out/ab_$be.err
  echo "$be rc=$?"; tail -2 gpurun_out/ab_$be.err | cut -c1-300
  python -c "
import json
d=json.load(open('gpurun_out/ab_$be.json')); print('$be', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms'].items()})"
done
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab_single.json 2> gpurun_out/ab_single.err
python -c "
import json
d=json.load(open('gpurun_out/ab_single.json')); print('single', round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms'].items()})"
