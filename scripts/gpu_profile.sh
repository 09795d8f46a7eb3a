#!/bin/bash
# ncu evidence for the bench command: launch list of our kernels, one full capture of the emitting march, one of the
# small kernels. Usage: gpu_profile.sh <tag> [bench args]
tag=${1:-run}; shift
mkdir -p gpurun_out
KR='regex:march_kernel|vscan_kernel|presence_kernel|remap_kernel|rowscan_kernel|dict_prefix_kernel|rowcount_kernel|generic'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 64 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/ncu_list_${tag}.log 2>&1; echo "ncu list rc=$?"
cut -d, -f5,14- gpurun_out/launches_${tag}.csv | cut -c1-160 | tail -24
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 1 -c 1 -f -o gpurun_out/prof_march_${tag} python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/ncu_march_${tag}.log 2>&1; echo "ncu march rc=$?"
tail -2 gpurun_out/ncu_march_${tag}.log | cut -c1-300
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:vscan_kernel|rowcount_kernel|remap_kernel|presence_kernel" -s 4 -c 4 -f -o gpurun_out/prof_small_${tag} python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/ncu_small_${tag}.log 2>&1; echo "ncu small rc=$?"
tail -2 gpurun_out/ncu_small_${tag}.log | cut -c1-300
