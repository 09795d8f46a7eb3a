#!/bin/bash
# A/B of a debug flag on the device-resident bench. Usage: gpu_ab.sh "<bench args>" flagsA flagsB ...
args=$1; shift
for f in "$@"; do
  SHF_DEBUG_FLAGS=$f timeout 600 python bench.py $args --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('flags=$f $args: %.0f Mpx/s step %.3f ms emit_frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']), {k: round(v,3) for k,v in d['phases_ms'].items()})"
done
