#!/bin/bash
# tests + bench + ncu of the element kernel
mkdir -p gpurun_out
T=${1:-x}
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_${T}_pytest.log 2>&1
tail -4 gpurun_out/r2_${T}_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r2_${T}_bench.json 2> gpurun_out/r2_${T}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_${T}_bench.json")); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d.get("parity"), d["e2e"]["value"])
except Exception as e: print("failed", e)
PY
tail -3 gpurun_out/r2_${T}_bench.err
ncu --set full --clock-control none --import-source on -k regex:${2:-k_element_tma} -s 9 -c 1 -o gpurun_out/r2_${T}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
