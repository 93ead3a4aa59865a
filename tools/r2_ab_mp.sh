#!/bin/bash
# A/B of library builds at N=1 and N=2 on the same box: r2_ab_mp.sh lib1.so lib2.so ...
P='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], d["ms_per_step"], d["ms_per_step_blocks"]["median"])'
for lib in "$@"; do
  export PDES_LIB=$PWD/$lib
  CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu-baseline --no-parity --steps 40 --warmup 5 2>/dev/null | python -c "$P" "$lib n1"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --no-cpu-baseline --no-parity --steps 40 --warmup 5 2>/dev/null | python -c "$P" "$lib n2"
done
