#!/bin/bash
# full ncu capture of selected kernels on the uniform and the blocky C3 batch. Usage: gpu_ncu_k.sh <tag> <kernel regex> [count]
tag=${1:-run}; kr=${2:-events_kernel}; cnt=${3:-1}
mkdir -p gpurun_out
for d in uniform blocky; do
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$kr" -s $cnt -c $cnt -f -o gpurun_out/prof_${tag}_${d} python bench.py --chunks 64 --dist $d --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_${tag}_${d}.log 2>&1; echo "ncu $d rc=$?"
done
