#!/bin/bash
# final multi-GPU evidence at N ranks: (N == 2: the transport tests), weak-scaling bench with the parity block, strong-scaling
# bench (54^3 mesh cut N ways), the N = 1 rate of GPU 0 on the same box, e2e without the NUMA binding
NG=${1:-2}
mkdir -p gpurun_out
if [ "$NG" = "2" ]; then python -m pytest tests/test_multi_process.py -m gpu -x -q 2>&1 | tail -3; fi
P='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "parity", (d.get("parity") or {}).get("ok"), d["e2e"].get("host_numa_binding"))'
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu-baseline --no-parity --steps 40 --warmup 5 2>/dev/null | tee gpurun_out/r2_final_n1_on_n${NG}_box.json | python -c "$P" gpu0-alone
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --no-cpu-baseline"
$T 2>gpurun_out/r2_final_mp.err | tee gpurun_out/r2_final_bench_n${NG}_weak.json | python -c "$P" n$NG-weak
tail -2 gpurun_out/r2_final_mp.err
$T --no-parity --scaling strong 2>/dev/null | tee gpurun_out/r2_final_bench_n${NG}_strong54.json | python -c "$P" n$NG-strong54
if [ "$NG" = "8" ]; then $T --no-parity --no-numa-bind --steps 5 2>/dev/null | python -c "$P" n$NG-weak-nonuma; fi
