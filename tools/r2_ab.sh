#!/bin/bash
# A/B runs of bench.py under different environment settings: r2_ab.sh "VAR=val VAR2=val" "..." ...
mkdir -p gpurun_out
for cfg in "$@"; do
  out=$(env $cfg python bench.py --no-cpu-baseline --no-parity --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['ms_per_step_blocks']['median'], d['roofline']['frac'])")
  echo "AB [$cfg] $out"
done
