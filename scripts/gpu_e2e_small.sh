#!/bin/bash
for v in "--workload C1 --dist uniform" "--workload C1 --dist blocky" "--workload C2 --dist blocky" "--workload C4 --dist blocky"; do
  timeout 600 python bench.py $v --steps 20 --warmup 3 --e2e-steps 10 --cpu-seconds 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; c=d['cpu_baseline']
print('$v | device-resident %.0f Mpx/s (%.3f ms) | e2e %.0f Mpx/s (%.3f ms, d2h %.1f MB) | reference as shipped %.1f Mpx/s (%d threads)' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['d2h_bytes_per_step']/1e6, c['value'], c['cores']))"
done
