#!/bin/bash
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("   %.0f Mpx/s step %.4f ms | plan %s |" % (d["value"], d["ms_per_step"], d["plan"]), {k: round(v,4) for k,v in d["phases_ms"].items()})
except Exception as e:
    print("   failed", e, open("gpurun_out/ab.err").read()[-300:])
PY
}
for v in "--chunks 32" "--chunks 256" "--chunks 32 --dist blocky" "--dist blocky"; do
 for mode in "X=1" "SHF_DEBUG_TY=8 SHF_DEBUG_NP=2 SHF_DEBUG_EXTRA=1" "SHF_DEBUG_TY=8 SHF_DEBUG_NP=2 SHF_DEBUG_EXTRA=2" "SHF_DEBUG_TY=8 SHF_DEBUG_NP=1 SHF_DEBUG_EXTRA=2"; do
  echo "== $v  [$mode]"
  env $mode timeout 240 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e --no-consumer --parity-chunks 1 > gpurun_out/ab.json 2> gpurun_out/ab.err
  line gpurun_out/ab.json
 done
done
