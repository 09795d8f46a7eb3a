#!/bin/bash
# Round status on one B200: parity tests, the default bench line (with e2e + cpu baseline), the reference arm, variant
# lines, the ncu launch list of the bench command and one full capture of emit/events/vscan. Usage: gpu_status.sh <tag>
tag=${1:-status}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_${tag}.log
timeout 900 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${tag}.err; cat gpurun_out/bench_${tag}.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_${tag}.json | cut -c1-600
for v in "--workload C3 --dist blocky" "--workload C2 --dist uniform" "--workload C2 --dist blocky" "--workload C4 --dist blocky" "--workload C4 --dist uniform" "--workload C1 --dist uniform"; do
  name=$(echo $v | tr -d ' -' )
  timeout 600 python bench.py $v --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/q_${tag}_${name}.json 2> gpurun_out/q_${tag}_${name}.err; echo "$v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/q_${tag}_${name}.json"))
    print(" value %.0f Mpx/s ms/step %.3f emit_frac %.3f step_frac %.3f bins/px %.2f plan %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["config"]["bins_per_pixel"], d["config"]["plan"]))
    print(" phases", {k: round(v,3) for k,v in d["phases_ms"].items()})
except Exception as e:
    print(" failed", e); print(open("gpurun_out/q_${tag}_${name}.err").read()[-600:])
PY
done
KR='regex:emit_kernel|events_kernel|vscan_kernel|presence_kernel|remap_kernel|rowscan_kernel|dict_prefix_kernel|generic'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 64 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_list_${tag}.log 2>&1; echo "ncu list rc=$?"
cut -d, -f5,14- gpurun_out/launches_${tag}.csv | cut -c1-160 | tail -24
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:emit_kernel|events_kernel|vscan_kernel|remap_kernel|presence_kernel" -s 5 -c 5 -f -o gpurun_out/prof_${tag} python bench.py --chunks 64 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_${tag}.log 2>&1; echo "ncu full rc=$?"
tail -2 gpurun_out/ncu_full_${tag}.log | cut -c1-300
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:emit_kernel|events_kernel|vscan_kernel" -s 3 -c 3 -f -o gpurun_out/prof_blocky_${tag} python bench.py --chunks 64 --dist blocky --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_fullb_${tag}.log 2>&1; echo "ncu full blocky rc=$?"
