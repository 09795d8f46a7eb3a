#!/bin/bash
# rows-per-CTA A/B on the single-chunk configurations. Usage: gpu_ty.sh
mkdir -p gpurun_out
for ty in 16 8; do
for v in "--workload C1 --dist uniform" "--workload C1 --dist blocky" "--workload C2 --dist uniform" "--workload C2 --dist blocky" "--workload C4 --dist blocky"; do
  SHF_DEBUG_TY=$ty timeout 600 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("TY<=$ty | $v | %.0f Mpx/s step %.3f ms emit %.3f events %.3f vscan %.3f dict %.3f plan %s" % (d["value"], d["ms_per_step"], d["phases_ms"]["emit"], d["phases_ms"]["events"], d["phases_ms"]["remap_vscan"], d["phases_ms"]["dictionary"], d["config"]["plan"]["rows_per_cta"]))
except Exception as e:
    print("$v | failed", e); print(open("gpurun_out/ab.err").read()[-400:])
PY
done
done
