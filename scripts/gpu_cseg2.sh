#!/bin/bash
for v in "--workload C1 --dist uniform" "--workload C1 --dist blocky" "--workload C2 --dist uniform" "--workload C2 --dist blocky" "--workload C4 --dist blocky" "--workload C4 --dist uniform" "--workload C3 --dist uniform"; do
  timeout 600 python bench.py $v --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('planned | $v | %.0f Mpx/s step %.3f ms' % (d['value'], d['ms_per_step']), {k: round(x,3) for k,x in d['phases_ms'].items()})"
done
