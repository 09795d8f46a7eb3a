#!/bin/bash
# Parity tests + the device-resident lines that matter for strong scaling (256 / 32 chunks, uniform / blocky) and the
# single-call configs. Usage: gpu_r2_quick.sh <tag> [pytest-filter]
tag=${1:-q}
mkdir -p gpurun_out
for tf in tests/test_parity_gpu.py tests/test_heightfield_gpu.py tests/test_biome_gpu.py tests/test_fuzz_gpu.py tests/test_cpp_dropin.py; do
  timeout 900 python -m pytest $tf -q -m gpu -x --timeout 200 --timeout-method=thread ${2:+-k "$2"} > gpurun_out/pytest_${tag}_$(basename $tf .py).log 2>&1; echo "pytest $tf rc=$?"
  tail -4 gpurun_out/pytest_${tag}_$(basename $tf .py).log | cut -c1-400
done
summ() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(" value %.0f Mpx/s ms/step %.4f emit_frac %.3f step_frac %.3f bins/px %.2f plan %s repeated %s parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["bins_per_pixel"], d["plan"], d["calls_repeated_on_checked_path"], (d.get("parity_checked") or {}).get("ok")))
    print(" phases", {k: round(v,4) for k,v in d["phases_ms"].items()})
except Exception as e:
    print(" failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
}
for v in "--chunks 256" "--chunks 32" "--chunks 64" "--chunks 32 --dist blocky" "--dist blocky" "--workload C2 --min-seconds 0.3" "--workload C2 --dist blocky --min-seconds 0.3" "--workload C4 --min-seconds 0.3" "--workload C1 --min-seconds 0.3"; do
  name=$(echo $v | tr -d ' -.' )
  timeout 240 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e --no-consumer > gpurun_out/q_${tag}_${name}.json 2> gpurun_out/q_${tag}_${name}.err; echo "$v rc=$?"
  summ gpurun_out/q_${tag}_${name}.json
done
KR='regex:emit_kernel|events_kernel|vscan_kernel|presence_kernel|remap_kernel|rowscan_kernel|vpatch|generic'
for v in "--chunks 32" "--chunks 256"; do
  name=$(echo $v | tr -d ' -.' )
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 30 --csv --log-file gpurun_out/launches_${tag}_${name}.csv python bench.py $v --steps 2 --warmup 2 --no-cpu --no-e2e --no-consumer --parity-chunks 0 > gpurun_out/ncu_list_${tag}_${name}.log 2>&1; echo "ncu list $v rc=$?"
  python - gpurun_out/launches_${tag}_${name}.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
last={}; order=[]
for r in rows[1:]:
    k=r[ki].split('(')[0][:60]
    if k not in last: order.append(k)
    last[k]=float(r[vi].replace(',',''))
print({k: round(last[k]/1e3,1) for k in order}, "us (last launch of each kernel)")
PY
done
