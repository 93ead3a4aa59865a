#!/bin/bash
# quick: bench (no parity) + ncu of one kernel.  usage: r2_run_q.sh <tag> <kernel-regex> [skip]
mkdir -p gpurun_out
T=$1; K=$2; S=${3:-9}
python bench.py --no-cpu-baseline --no-parity > gpurun_out/r2_${T}_bench.json 2> gpurun_out/r2_${T}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_${T}_bench.json")); print("BENCH", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["value"])
except Exception as e: print("failed", e)
PY
tail -3 gpurun_out/r2_${T}_bench.err
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/r2_${T}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
