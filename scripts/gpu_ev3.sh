#!/bin/bash
# gpu_ev2 + one full ncu capture of selected kernels on the 64-chunk uniform batch. Usage: gpu_ev3.sh <tag> <kernel regex> <count>
tag=${1:-ev}; kr=${2:-vscan_kernel}; cnt=${3:-1}
bash scripts/gpu_ev2.sh $tag
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$kr" -s $cnt -c $cnt -f -o gpurun_out/prof_${tag} python bench.py --chunks 64 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_${tag}.log 2>&1; echo "ncu rc=$?"
