#!/bin/bash
# full ncu capture of the small kernels. Usage: gpu_ncu2.sh <tag> [bench args]
tag=${1:-run}; shift
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:vscan_kernel|rowcount_kernel|remap_kernel|presence_kernel" -s 4 -c 4 -f -o gpurun_out/prof_small_${tag} python bench.py --chunks 64 --steps 1 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/ncu_small_${tag}.log 2>&1; echo "ncu small rc=$?"
tail -2 gpurun_out/ncu_small_${tag}.log | cut -c1-300
