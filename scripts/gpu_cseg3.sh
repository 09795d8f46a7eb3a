#!/bin/bash
for cfg in "16 1" "8 1" "8 2" "8 3" "16 2"; do set -- $cfg
for v in "--workload C4 --dist blocky" "--workload C4 --dist uniform"; do
  SHF_DEBUG_TY=$1 SHF_DEBUG_CSEG=$2 timeout 600 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('TY=$1 cseg=$2 | $v | %.0f Mpx/s step %.3f ms emit %.3f vscan %.3f' % (d['value'], d['ms_per_step'], d['phases_ms']['emit'], d['phases_ms']['remap_vscan']))"
done; done
