#!/bin/bash
# parity tests, then bench lines (new path only). Usage: gpu_ev2.sh <tag> [notest]
tag=${1:-ev}
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then timeout 1200 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -8; fi
for v in "--workload C3 --dist uniform" "--workload C3 --dist blocky" "--workload C2 --dist blocky" "--workload C4 --dist blocky" "--workload C1 --dist uniform"; do
  name=$(echo $v | tr -d ' -' )
  timeout 600 python bench.py $v --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/q_${tag}_${name}.json 2> gpurun_out/q_${tag}_${name}.err; echo "$v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/q_${tag}_${name}.json"))
    print(" value %.0f Mpx/s ms/step %.3f emit_frac %.3f step_frac %.3f bins/px %.2f plan %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["config"]["bins_per_pixel"], d["config"]["plan"]))
    print(" phases", {k: round(v,3) for k,v in d["phases_ms"].items()})
except Exception as e:
    print(" failed", e); print(open("gpurun_out/q_${tag}_${name}.err").read()[-600:])
PY
done
