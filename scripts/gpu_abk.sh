#!/bin/bash
# A/B of two builds of the library (SHF_LIBRARY) over a few workloads. Usage: gpu_abk.sh <tag>
tag=${1:-ab}
mkdir -p gpurun_out
for lib in superterrainplus_b200/libshf_b200_dense.so superterrainplus_b200/libshf_b200.so; do
for v in "--workload C3 --dist uniform" "--workload C3 --dist blocky" "--workload C3 --dist blocky --biomes 256 --chunks 64" "--workload C3 --dist uniform --biomes 256 --chunks 32" "--workload C3 --dist blocky --biomes 128 --chunks 64" "--workload C3 --dist blocky --biomes 16" "--workload C3 --dist rare --chunks 64"; do
  SHF_LIBRARY=$PWD/$lib timeout 600 python bench.py $v --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("$(basename $lib) | $v | %.0f Mpx/s step %.3f ms emit %.3f events %.3f vscan %.3f bins/px %.2f K=%s" % (d["value"], d["ms_per_step"], d["phases_ms"]["emit"], d["phases_ms"]["events"], d["phases_ms"]["remap_vscan"], d["config"]["bins_per_pixel"], d["config"]["plan"]["k_sets"]))
except Exception as e:
    print("$(basename $lib) | $v | failed", e); print(open("gpurun_out/ab.err").read()[-400:])
PY
done
done
