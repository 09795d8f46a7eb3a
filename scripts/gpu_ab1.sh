#!/bin/bash
# tests + a few bench lines of the current build. Usage: gpu_ab1.sh <tag>
tag=${1:-ab}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -4
for v in "--workload C3 --dist uniform" "--workload C3 --dist blocky" "--workload C3 --dist blocky --biomes 256 --chunks 64" "--workload C3 --dist uniform --biomes 256 --chunks 32" "--workload C3 --dist blocky --biomes 16" "--workload C2 --dist blocky" "--workload C1 --dist uniform"; do
  timeout 600 python bench.py $v --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("$v | %.0f Mpx/s step %.3f ms emit %.3f events %.3f vscan %.3f dict %.3f bins/px %.2f plan %s" % (d["value"], d["ms_per_step"], d["phases_ms"]["emit"], d["phases_ms"]["events"], d["phases_ms"]["remap_vscan"], d["phases_ms"]["dictionary"], d["config"]["bins_per_pixel"], d["config"]["plan"]))
except Exception as e:
    print("$v | failed", e); print(open("gpurun_out/ab.err").read()[-400:])
PY
done
