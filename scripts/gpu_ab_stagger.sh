#!/bin/bash
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("   step %.4f ms | emit %.4f | remap_vscan %.4f events %.4f" % (d["ms_per_step"], d["phases_ms"]["emit"], d["phases_ms"]["remap_vscan"], d["phases_ms"]["events"]))
except Exception as e:
    print("   failed", e, open("gpurun_out/ab.err").read()[-300:])
PY
}
for c in 32 256; do
for mode in "X=1" "SHF_DEBUG_STAGGER=10" "SHF_DEBUG_STAGGER=30" "SHF_DEBUG_STAGGER=60" "SHF_DEBUG_STAGGER=100" "SHF_DEBUG_STAGGER=30 SHF_NO_SPLIT=1" "SHF_NO_SPLIT=1"; do
  echo "== $c chunks [$mode]"
  env $mode timeout 240 python bench.py --chunks $c --steps 30 --warmup 3 --no-cpu --no-e2e --no-consumer --parity-chunks 0 > gpurun_out/ab.json 2> gpurun_out/ab.err
  line gpurun_out/ab.json
done
done
