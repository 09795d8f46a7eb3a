#!/bin/bash
mkdir -p gpurun_out
for tf in tests/test_parity_gpu.py tests/test_fuzz_gpu.py tests/test_biome_gpu.py; do
  timeout 900 python -m pytest $tf -q -m gpu -x --timeout 200 --timeout-method=thread > gpurun_out/pytest_chk_$(basename $tf .py).log 2>&1; echo "pytest $tf rc=$?"
  tail -2 gpurun_out/pytest_chk_$(basename $tf .py).log | cut -c1-300
done
line() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("   %.0f Mpx/s step %.4f ms (with events %.4f) | parity %s rep %s |" % (d["value"], d["ms_per_step"], d["ms_per_step_with_phase_events"] or 0, (d.get("parity_checked") or {}).get("ok"), d["calls_repeated_on_checked_path"]), {k: round(v,4) for k,v in d["phases_ms"].items()})
except Exception as e:
    print("   failed", e, open("gpurun_out/ab.err").read()[-400:])
PY
}
for v in "--chunks 32" "--chunks 256" "--chunks 32 --dist blocky" "--dist blocky" "--workload C2 --min-seconds 0.2"; do
  echo "== $v"
  timeout 240 python bench.py $v --steps 30 --warmup 3 --no-cpu --no-e2e --no-consumer --parity-chunks 2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  line gpurun_out/ab.json
done
