#!/bin/bash
# tests + fuzz + single-chunk lines with and without the run-ahead path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -4
python tests/fuzz_gpu.py 300 11 | tail -1
for spec in 1 0; do
for v in "--workload C1 --dist uniform" "--workload C2 --dist uniform" "--workload C2 --dist blocky" "--workload C4 --dist blocky" "--workload C3 --dist uniform" "--workload C3 --dist blocky"; do
  if [ $spec = 0 ]; then export SHF_NO_SPECULATION=1; else unset SHF_NO_SPECULATION; fi
  timeout 600 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("run-ahead=$spec | $v | %.0f Mpx/s step %.3f ms | phases %s" % (d["value"], d["ms_per_step"], {k: round(x,3) for k,x in d["phases_ms"].items()}))
except Exception as e:
    print("$v | failed", e); print(open("gpurun_out/ab.err").read()[-400:])
PY
done
done
