#!/bin/bash
# BASELINE.json config 5 at N GPUs (strong scaling: one 256-chunk batch sharded): radius 8..128 x biome count 4..1024
# on 512x512 chunks, uniform (dense) and blocky (sparse) ids. Usage: gpu_sweep_multi.sh <N> <tag>   (gpurun --gpus N)
n=${1:-8}; tag=${2:-sweep8}
mkdir -p gpurun_out
out=gpurun_out/sweep_${tag}.txt
: > $out
port=29600
for dist in uniform blocky; do
for r in ${SWEEP_R:-8 16 32 64 128}; do
for b in ${SWEEP_B:-4 16 64 256 1024}; do
  chunks=256; if [ "$dist" = uniform ] && [ $b -ge 256 ]; then chunks=64; fi
  port=$((port+1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload C3 --dist $dist --radius $r --biomes $b --chunks $chunks --steps 5 --warmup 3 --no-cpu --no-e2e --no-consumer --parity-chunks 1 > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - >> $out <<PY
import json
try:
    d=json.load(open("gpurun_out/sw.json"))
    print("$dist r=$r B=$b chunks=$chunks gpus=$n | %.0f Mpx/s | step %.3f ms | emit %.3f | bins/px %.2f | K=%s TY=%s | step_frac(rank0) %.3f | parity %s" % (d["value"], d["ms_per_step"], d["phases_ms"]["emit"], d["bins_per_pixel"], d["plan"]["k_sets"], d["plan"]["rows_per_cta"], d["roofline"]["whole_step_frac"], (d.get("parity_checked") or {}).get("ok")))
except Exception as e:
    print("$dist r=$r B=$b | failed", e, open("gpurun_out/sw.err").read()[-300:].replace("\n"," "))
PY
done; done; done
cat $out
