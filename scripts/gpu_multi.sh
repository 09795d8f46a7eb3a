#!/bin/bash
# bench.py at 1, 2, 4, ... N GPUs of one box, the way the driver launches it (torchrun, strong scaling: one 256-chunk batch
# sharded). Usage: gpu_multi.sh <N> <tag> [extra bench args]     (gpurun --gpus N)
n=${1:-2}; tag=${2:-multi}; shift 2
mkdir -p gpurun_out
k=1
while [ $k -le $n ]; do
  if [ $k -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu "$@" > gpurun_out/bench_${tag}_1gpu.json 2> gpurun_out/bench_${tag}_1gpu.err; echo "1 gpu rc=$?"
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port $((29500+k)) bench.py --gpus $k --no-cpu "$@" > gpurun_out/bench_${tag}_${k}gpu.json 2> gpurun_out/bench_${tag}_${k}gpu.err; echo "$k gpu rc=$?"
  fi
  python - gpurun_out/bench_${tag}_${k}gpu.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    e=d.get("e2e") or {}
    print(" N=%d value %.0f Mpx/s step %.4f ms scaling %s | chunks/gpu %s | e2e %.0f Mpx/s (%.1f GB/s D2H, ceiling %.1f, frac %.2f) | parity %s clocks %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["scaling"], d["config"]["chunks_per_gpu"], e.get("value",0), e.get("d2h_gb_per_s",0), e.get("d2h_ceiling_gb_per_s",0), e.get("frac_of_d2h_ceiling",0), (d.get("parity_checked") or {}).get("ok"), (d.get("clocks") or {}).get("sm_mhz")))
except Exception as ex:
    print(" failed", ex); print(open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
  k=$((k*2))
done
