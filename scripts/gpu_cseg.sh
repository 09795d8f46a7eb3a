#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -4
python tests/fuzz_gpu.py 300 31 | tail -1
for S in 1 2 4; do
for v in "--workload C1 --dist uniform" "--workload C2 --dist uniform" "--workload C2 --dist blocky" "--workload C4 --dist blocky" "--workload C4 --dist uniform"; do
  SHF_DEBUG_CSEG=$S timeout 600 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("cseg<=$S | $v | %.0f Mpx/s step %.3f ms | vscan %.3f emit %.3f" % (d["value"], d["ms_per_step"], d["phases_ms"]["remap_vscan"], d["phases_ms"]["emit"]))
except Exception as e:
    print("$v | failed", e); print(open("gpurun_out/ab.err").read()[-400:])
PY
done
done
for v in "--workload C1 --dist uniform" "--workload C2 --dist uniform" "--workload C4 --dist blocky" "--workload C3 --dist uniform"; do
  timeout 600 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('planned | $v | %.0f Mpx/s step %.3f ms emit %.3f' % (d['value'], d['ms_per_step'], d['phases_ms']['emit']))"
done
