#!/bin/bash
# usage: tools/bench_variants.sh "TILE:MINB ..."  -- runs bench.py for each kernel variant
for v in $1; do
  t=${v%%:*}; m=${v##*:}
  PDES_TILE=$t PDES_MINB=$m python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('tile $t minb $m  value %.4g  ms/step %.3f  frac %.3f  e2e %.4g' % (d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']))"
done
