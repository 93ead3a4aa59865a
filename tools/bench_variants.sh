#!/bin/bash
# usage: tools/bench_variants.sh "0 1 2 ..." [extra bench args] -- runs bench.py for each kernel variant (PDES_VARIANT)
for v in $1; do
  PDES_VARIANT=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline $2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('variant $v  value %.4g  ms/step %.3f  frac %.3f  e2e %.4g' % (d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']))"
done
