#!/bin/bash
# tests + bench + ncu launch list on one B200. Usage: gpu_round.sh <tag> [bench args...]
tag=${1:-run}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_${tag}.log
tail -25 gpurun_out/pytest_${tag}.log
timeout 900 python bench.py --chunks 32 --steps 3 --warmup 2 --no-cpu "$@" > gpurun_out/bench_small_${tag}.json 2> gpurun_out/bench_small_${tag}.err; echo "bench small rc=$?"
tail -3 gpurun_out/bench_small_${tag}.err; cat gpurun_out/bench_small_${tag}.json
timeout 1500 python bench.py "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${tag}.err; cat gpurun_out/bench_${tag}.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --chunks 32 --steps 2 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/ncu_bench_${tag}.log 2>&1; echo "ncu rc=$?"
tail -12 gpurun_out/launches_${tag}.csv | cut -c1-300
