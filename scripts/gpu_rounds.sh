#!/bin/bash
# emit time against the number of tile rounds (chunks * 32 tiles / 148 SMs): fixed cost per launch vs cost per round
mkdir -p gpurun_out
for c in 9 18 28 32 37 74 148 256; do
  timeout 240 python bench.py --chunks $c --steps 20 --warmup 3 --no-cpu --no-e2e --no-consumer --parity-chunks 0 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - $c <<'PY'
import json,sys
c=int(sys.argv[1])
try:
    d=json.load(open("gpurun_out/ab.json"))
    p=d["phases_ms"]
    print("chunks %3d rounds %.2f | step %.4f emit %.4f (%.1f us/round) dict %.4f remap_vscan %.4f events %.4f" % (c, c*32/148, d["ms_per_step"], p["emit"], p["emit"]*1e3/(c*32/148), p["dictionary"], p["remap_vscan"], p["events"]))
except Exception as e:
    print("failed", e)
PY
done
