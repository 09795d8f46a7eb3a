#!/bin/bash
# 8-GPU evidence: strong-scaling bench at 1/2/4/8 GPUs (with e2e + D2H ceiling), the standalone D2H ceiling, and BASELINE
# config 5 (radius x biome sweep) at 8 GPUs on a reduced grid. Usage: gpu_final8.sh <tag>    (gpurun --gpus 8)
tag=${1:-r2h}
bash scripts/gpu_multi.sh 8 ${tag}
timeout 300 python scripts/d2h_ceiling.py --gpus 8 --gib 2 > gpurun_out/d2h_ceiling_${tag}.json 2>/dev/null; cat gpurun_out/d2h_ceiling_${tag}.json | cut -c1-600
SWEEP_R="8 64 128" SWEEP_B="4 64 1024" bash scripts/gpu_sweep_multi.sh 8 ${tag}_8gpu | tail -20
