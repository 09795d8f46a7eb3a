#!/bin/bash
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("   %.0f Mpx/s step %.4f ms | " % (d["value"], d["ms_per_step"]), {k: round(v,4) for k,v in d["phases_ms"].items()})
except Exception as e:
    print("   failed", e)
PY
}
for v in "--chunks 256" "--chunks 32" "--workload C2 --min-seconds 0.2"; do
 for mode in "persist+cut" "SHF_DEBUG_NOCUT=1" "SHF_NO_PERSIST=1"; do
  echo "== $v  [$mode]"
  if [ "$mode" = "persist+cut" ]; then env="X=1"; else env="$mode"; fi
  env $env timeout 240 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e --no-consumer --parity-chunks 0 > gpurun_out/ab.json 2> gpurun_out/ab.err
  line gpurun_out/ab.json
 done
done
