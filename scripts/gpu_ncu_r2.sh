#!/bin/bash
# one `ncu --set full` capture of every kernel of a step (uniform + blocky C3, 64 chunks). Usage: gpu_ncu_r2.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
KR='regex:emit_kernel|events_kernel|vscan_kernel|presence_kernel|bases_kernel'
timeout 900 ncu --set full --clock-control none --import-source on -k "$KR" -s 6 -c 6 -f -o gpurun_out/prof_${tag}_uniform python bench.py --chunks 64 --steps 1 --warmup 1 --no-cpu --no-e2e --no-consumer --parity-chunks 0 > gpurun_out/ncu_${tag}_uniform.log 2>&1; echo "uniform rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$KR" -s 6 -c 6 -f -o gpurun_out/prof_${tag}_blocky python bench.py --chunks 64 --dist blocky --steps 1 --warmup 1 --no-cpu --no-e2e --no-consumer --parity-chunks 0 > gpurun_out/ncu_${tag}_blocky.log 2>&1; echo "blocky rc=$?"
ls -la gpurun_out/prof_${tag}_*.ncu-rep
