#!/bin/bash
# multi-GPU check: transport tests, then N=1 on GPU 0 and N=NG (default / NCCL norm) on the same box
NG=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_multi_process.py -m gpu -x -q 2>&1 | tail -5
B="python bench.py --no-cpu-baseline --no-parity --steps 40 --warmup 5"
P='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], d["ms_per_step"], d["ms_per_step_blocks"]["median"], d["value"], d["e2e"]["value"], d["gpu_launches"])'
CUDA_VISIBLE_DEVICES=0 $B 2>/dev/null | python -c "$P" gpu0-alone
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --no-cpu-baseline --steps 40 --warmup 5"
$T 2>gpurun_out/mp_diag2.err | tee gpurun_out/mp_diag2_n$NG.json | python -c "$P" n$NG-default
tail -3 gpurun_out/mp_diag2.err
PDES_NORM_NCCL=1 $T --no-parity 2>/dev/null | python -c "$P" n$NG-ncclnorm
$T --no-parity --scaling strong 2>/dev/null | tee gpurun_out/mp_diag2_strong_n$NG.json | python -c "$P" n$NG-strong54
