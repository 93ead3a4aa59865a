#!/bin/bash
# round-2 evidence: launch list (duration + DRAM bytes per launch) and ONE full ncu capture (gpurun returns <= 64 MiB per call)
# usage: r2_profile.sh <tag> <kernel-regex> [bench]
mkdir -p gpurun_out
T=${1:-final}; K=${2:-k_element_tma}
if [ "$3" = "bench" ]; then
  python bench.py > gpurun_out/r2_${T}_bench.json 2> gpurun_out/r2_${T}_bench.err
  tail -2 gpurun_out/r2_${T}_bench.err
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
      --log-file gpurun_out/r2_${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
fi
ncu --set full --clock-control none --import-source on -k regex:$K -s 9 -c 1 -o gpurun_out/r2_${T}_${K} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
ls -la gpurun_out/r2_${T}_*
