#!/bin/bash
# round-2 GPU run B: new element kernel -- tests, A/B bench, launch list
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_b_pytest.log 2>&1
tail -4 gpurun_out/r2_b_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r2_b_bench_tma.json 2> gpurun_out/r2_b_bench_tma.err
PDES_ELEM_TMA=0 python bench.py --no-cpu-baseline --no-parity > gpurun_out/r2_b_bench_tile.json 2>/dev/null
python - <<'PY'
import json
for n in ("tma","tile"):
    try:
        d=json.load(open(f"gpurun_out/r2_b_bench_{n}.json")); print(n, d["ms_per_step"], d["value"], d["roofline"]["frac"], d.get("parity"))
    except Exception as e: print(n, "failed", e)
PY
tail -3 gpurun_out/r2_b_bench_tma.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r2_b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
grep -E "k_element|k_face" gpurun_out/r2_b_launches.csv | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | head -24
