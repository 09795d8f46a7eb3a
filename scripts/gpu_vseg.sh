#!/bin/bash
mkdir -p gpurun_out
for S in 4 6 8; do
for v in "--workload C1 --dist uniform" "--workload C2 --dist uniform" "--workload C2 --dist blocky" "--workload C4 --dist blocky" "--workload C4 --dist uniform"; do
  SHF_DEBUG_VSEG=$S timeout 600 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("S<=$S | $v | %.0f Mpx/s step %.3f ms | vscan %.3f emit %.3f" % (d["value"], d["ms_per_step"], d["phases_ms"]["remap_vscan"], d["phases_ms"]["emit"]))
except Exception as e:
    print("$v | failed", e); print(open("gpurun_out/ab.err").read()[-400:])
PY
done
done
