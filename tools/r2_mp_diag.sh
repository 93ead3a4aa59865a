#!/bin/bash
# where does the multi-GPU step time go?  N=1 on each GPU alone, both at once (independent), N=2 with / without the norm
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-parity --steps 40 --warmup 5"
P='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], d["ms_per_step"], d["ms_per_step_blocks"]["median"], d["clocks"])'
CUDA_VISIBLE_DEVICES=0 $B 2>/dev/null | python -c "$P" gpu0-alone
CUDA_VISIBLE_DEVICES=1 $B 2>/dev/null | python -c "$P" gpu1-alone
(CUDA_VISIBLE_DEVICES=0 $B 2>/dev/null | python -c "$P" gpu0-concurrent) &
CUDA_VISIBLE_DEVICES=1 $B 2>/dev/null | python -c "$P" gpu1-concurrent
wait
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-cpu-baseline --no-parity --steps 40 --warmup 5"
$T 2>/dev/null | python -c "$P" n2-default
PDES_STEPS_NO_NORM=1 $T 2>/dev/null | python -c "$P" n2-nonorm
PDES_GRAPH_MP=0 PDES_STEPS_NO_NORM=1 $T 2>/dev/null | python -c "$P" n2-nonorm-nograph
PDES_HALO_FUSED=0 $T 2>/dev/null | python -c "$P" n2-unfused
