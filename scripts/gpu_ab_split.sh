#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x --timeout 200 --timeout-method=thread -k "c3_full or segments or async or persistent" 2>&1 | tail -3
line() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("   %.0f Mpx/s step %.4f ms | parity %s |" % (d["value"], d["ms_per_step"], (d.get("parity_checked") or {}).get("ok")), {k: round(v,4) for k,v in d["phases_ms"].items()})
except Exception as e:
    print("   failed", e, open("gpurun_out/ab.err").read()[-300:])
PY
}
for v in "--chunks 32" "--chunks 256" "--chunks 32 --dist blocky" "--dist blocky" "--chunks 64"; do
 for mode in "X=1" "SHF_NO_SPLIT=1"; do
  echo "== $v  [$mode]"
  env $mode timeout 240 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e --no-consumer --parity-chunks 2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  line gpurun_out/ab.json
 done
done
