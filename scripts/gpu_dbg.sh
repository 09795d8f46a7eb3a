#!/bin/bash
# cycle counters of the march kernel (SHF_DEBUG_FLAGS bit 1). Usage: gpu_dbg.sh "<bench args>" [flags]
SHF_DEBUG_FLAGS=${2:-2} timeout 600 python bench.py $1 --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 >/dev/null | grep "shf dbg" | tail -1
