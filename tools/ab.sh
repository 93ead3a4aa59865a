#!/bin/bash
# usage: tools/ab.sh "ENV1=a ENV2=b" "ENV1=c" ... -- one bench.py line per environment setting
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
L=sys.stdin.readlines()
try:
    d=json.loads(L[-1]); print('[$cfg]  value %.4g  ms/step %.3f  frac %.3f  e2e %.4g' % (d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']))
except Exception as e:
    print('[$cfg] FAILED', ''.join(L[-5:]))
"
done
