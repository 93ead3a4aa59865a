#!/bin/bash
# usage: tools/ab_wl.sh WORKLOAD "ENV1=a ENV2=b" "ENV1=c" ... -- like ab.sh for another bench workload
wl=$1; shift
for cfg in "$@"; do
  env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl 2>&1 | python -c "
import sys,json
L=sys.stdin.readlines()
try:
    d=json.loads(L[-1]); print('[$wl $cfg]  value %.4g  ms/step %.3f  frac %.3f  e2e %.4g' % (d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']))
except Exception as e:
    print('[$cfg] FAILED', ''.join(L[-5:]))
"
done
