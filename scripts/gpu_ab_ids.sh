#!/bin/bash
mkdir -p gpurun_out
for tf in tests/test_parity_gpu.py tests/test_biome_gpu.py tests/test_fuzz_gpu.py tests/test_cpp_dropin.py tests/test_heightfield_gpu.py; do
  timeout 900 python -m pytest $tf -q -m gpu -x --timeout 200 --timeout-method=thread > gpurun_out/pytest_ids_$(basename $tf .py).log 2>&1; echo "pytest $tf rc=$?"
  tail -3 gpurun_out/pytest_ids_$(basename $tf .py).log | cut -c1-300
done
line() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("   %.0f Mpx/s step %.4f ms | parity %s rep %s |" % (d["value"], d["ms_per_step"], (d.get("parity_checked") or {}).get("ok"), d["calls_repeated_on_checked_path"]), {k: round(v,4) for k,v in d["phases_ms"].items()})
except Exception as e:
    print("   failed", e, open("gpurun_out/ab.err").read()[-300:])
PY
}
for v in "--chunks 32" "--chunks 256" "--chunks 32 --dist blocky" "--dist blocky" "--workload C2 --min-seconds 0.2" "--workload C4 --min-seconds 0.2" "--workload C1 --min-seconds 0.2"; do
 for mode in "X=1" "SHF_NO_SMALL_IDS=1"; do
  echo "== $v  [$mode]"
  env $mode timeout 240 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e --no-consumer --parity-chunks 2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  line gpurun_out/ab.json
 done
done
