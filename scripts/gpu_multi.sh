#!/bin/bash
# bench.py under torchrun on N GPUs of one box (+ the same bench on one GPU of that box). Usage: gpu_multi.sh <N> <tag>
n=${1:-2}; tag=${2:-multi}
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu --no-e2e > gpurun_out/bench_${tag}_1of${n}.json 2> gpurun_out/bench_${tag}_1of${n}.err; echo "1 gpu rc=$?"
cat gpurun_out/bench_${tag}_1of${n}.json | cut -c1-200
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --no-cpu > gpurun_out/bench_${tag}_${n}gpu.json 2> gpurun_out/bench_${tag}_${n}gpu.err; echo "$n gpu rc=$?"
tail -3 gpurun_out/bench_${tag}_${n}gpu.err | cut -c1-300
cat gpurun_out/bench_${tag}_${n}gpu.json | cut -c1-400
