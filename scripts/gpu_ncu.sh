#!/bin/bash
# ncu: launch list of our kernels + one full capture of the emitting march. Usage: gpu_ncu.sh <tag> [bench args]
tag=${1:-run}; shift
mkdir -p gpurun_out
KR='regex:march_kernel|vscan_kernel|presence_kernel|remap_kernel|rowscan_kernel|dict_prefix_kernel|winmask|rowcount'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 40 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --chunks 32 --steps 2 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/ncu_list_${tag}.log 2>&1; echo "ncu list rc=$?"
cut -d, -f5,14- gpurun_out/launches_${tag}.csv | cut -c1-200 | tail -24
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 1 -c 1 -f -o gpurun_out/prof_emit_${tag} python bench.py --chunks 32 --steps 1 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/ncu_full_${tag}.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_${tag}.log
