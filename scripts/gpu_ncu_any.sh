#!/bin/bash
# full ncu capture of one kernel for arbitrary bench arguments. Usage: gpu_ncu_any.sh <tag> <kernel regex> <skip> <bench args...>
tag=$1; kr=$2; skip=$3; shift 3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$kr" -s $skip -c 1 -f -o gpurun_out/prof_${tag} python bench.py "$@" --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_${tag}.log 2>&1; echo "ncu rc=$?"
