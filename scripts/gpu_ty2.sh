#!/bin/bash
mkdir -p gpurun_out
run() { # ty, bench args
  ty=$1; shift
  SHF_DEBUG_TY=$ty timeout 600 python bench.py "$@" --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("TY<=$ty | $* | %.0f Mpx/s step %.3f ms emit %.3f vscan %.3f plan %s" % (d["value"], d["ms_per_step"], d["phases_ms"]["emit"], d["phases_ms"]["remap_vscan"], d["config"]["plan"]["rows_per_cta"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/ab.err").read()[-300:])
PY
}
for ty in 16 14; do run $ty --workload C4 --dist uniform; run $ty --workload C4 --dist blocky; done
for ty in 8 7; do run $ty --workload C2 --dist uniform; run $ty --workload C2 --dist blocky; done
for ty in 8 4; do run $ty --workload C1 --dist uniform; done
