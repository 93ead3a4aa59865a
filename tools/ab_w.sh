#!/bin/bash
# usage: tools/ab_w.sh WORKLOAD "ENV=a" "ENV=b" ... -- one bench.py line per environment setting for a workload
W=$1; shift
for cfg in "$@"; do
  env $cfg timeout 400 python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
L=sys.stdin.readlines()
try:
    d=json.loads(L[-1]); print('[$W $cfg]  value %.4g  ms/step %.3f  frac %.3f  e2e %.4g' % (d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']))
except Exception as e:
    print('[$cfg] FAILED', ''.join(L[-5:]))
"
done
