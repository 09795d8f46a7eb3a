#!/bin/bash
# Round-2 evidence on one B200: smoke, parity tests, the default bench line (value, roofline, parity_checked, consumer,
# config-4 chain, e2e + D2H ceiling, cpu baseline), the reference arm, variant lines with clock records, the ncu launch
# list of the bench command and full captures (uniform, blocky). Usage: gpu_status_r2.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 300 --timeout-method=thread > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_${tag}.log
timeout 900 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${tag}.err; cat gpurun_out/bench_${tag}.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "ref rc=$?"
cut -c1-500 gpurun_out/bench_ref_${tag}.json
for v in "--workload C3 --dist blocky" "--workload C2 --dist uniform" "--workload C2 --dist blocky" "--workload C4 --dist blocky" "--workload C4 --dist uniform" "--workload C1 --dist uniform" "--workload C3 --chunks 32" "--workload C3 --chunks 32 --dist blocky"; do
  name=$(echo $v | tr -d ' -' )
  timeout 600 python bench.py $v --steps 20 --warmup 3 --min-seconds 0.6 --no-cpu --no-e2e --no-consumer > gpurun_out/q_${tag}_${name}.json 2> gpurun_out/q_${tag}_${name}.err; echo "$v rc=$?"
  python - gpurun_out/q_${tag}_${name}.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(" value %.0f Mpx/s ms/step %.4f emit_frac %.3f step_frac %.3f bins/px %.2f plan %s parity %s clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["bins_per_pixel"], d["plan"], (d.get("parity_checked") or {}).get("ok"), {k: d["clocks"][k] for k in ("sm_mhz","samples_in_timed_region","reasons")}))
    print(" phases", {k: round(v,4) for k,v in d["phases_ms"].items()})
except Exception as e:
    print(" failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
KR='regex:emit_kernel|events_kernel|vscan_kernel|presence_kernel|remap_kernel|rowscan_kernel|bases_kernel|generic|heightfield|biome'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 64 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-consumer --parity-chunks 0 > gpurun_out/ncu_list_${tag}.log 2>&1; echo "ncu list rc=$?"
cut -d, -f5,14- gpurun_out/launches_${tag}.csv | cut -c1-150 | tail -16
KF='regex:emit_kernel|events_kernel|vscan_kernel|presence_kernel|bases_kernel'
timeout 900 ncu --set full --clock-control none --import-source on -k "$KF" -s 6 -c 6 -f -o gpurun_out/prof_${tag}_uniform python bench.py --chunks 64 --steps 1 --warmup 1 --no-cpu --no-e2e --no-consumer --parity-chunks 0 > gpurun_out/ncu_full_${tag}_uniform.log 2>&1; echo "ncu full uniform rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$KF" -s 6 -c 6 -f -o gpurun_out/prof_${tag}_blocky python bench.py --chunks 64 --dist blocky --steps 1 --warmup 1 --no-cpu --no-e2e --no-consumer --parity-chunks 0 > gpurun_out/ncu_full_${tag}_blocky.log 2>&1; echo "ncu full blocky rc=$?"
