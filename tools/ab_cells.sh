#!/bin/bash
# usage: tools/ab_cells.sh CELLS "ENV1=a ENV2=b" "ENV1=c" ... -- like ab.sh on a mesh of CELLS^3 cubes per rank
cells=$1; shift
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --cells $cells 2>&1 | python -c "
import sys,json
L=sys.stdin.readlines()
try:
    d=json.loads(L[-1]); print('[cells $cells $cfg]  value %.4g  ms/step %.3f  frac %.3f  e2e %.4g' % (d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']))
except Exception as e:
    print('[$cfg] FAILED', ''.join(L[-5:]))
"
done
