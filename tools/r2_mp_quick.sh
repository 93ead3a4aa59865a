#!/bin/bash
# N = 1 on GPU 0 and N ranks (weak scaling, with the parity block) on the same box
NG=${1:-2}
P='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "parity", (d.get("parity") or {}).get("ok"))'
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu-baseline --no-parity --steps 40 --warmup 5 2>/dev/null | python -c "$P" gpu0-alone
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --no-cpu-baseline --steps 40 --warmup 5 2>/dev/null | tee gpurun_out/r2_last_bench_n${NG}_weak.json | python -c "$P" n$NG-weak
