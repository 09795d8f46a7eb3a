#!/bin/bash
# First-contact GPU check: smoke, parity tests, memcheck of the smoke case. Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py --smoke > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/memcheck.log
tail -15 gpurun_out/memcheck.log
