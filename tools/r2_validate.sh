#!/bin/bash
# full GPU validation of the tree as it is: tests, smoke, the C1 (small mesh) and C3 bench lines
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_last_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --workload c1_2d_p1_roe --no-cpu-baseline --steps 400 --warmup 40 2>/dev/null | grep "^{" > gpurun_out/r2_last_bench_c1.json
python bench.py --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | grep "^{" > gpurun_out/r2_last_bench_c3.json
python - <<'PY'
import json
for f in ("c1", "c3"):
    d = json.loads(open("gpurun_out/r2_last_bench_%s.json" % f).read())
    print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "parity", (d.get("parity") or {}).get("ok"), d["gpu_launches"])
PY
