#!/bin/bash
# full ncu capture of emit_kernel / events_kernel / vscan. Usage: gpu_ncu_ev.sh <tag> [bench args]
tag=${1:-run}; shift
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:emit_kernel|events_kernel|vscan_kernel" -s 3 -c 3 -f -o gpurun_out/prof_ev_${tag} python bench.py --chunks 64 --steps 1 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/ncu_ev_${tag}.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_ev_${tag}.log | cut -c1-300
