#!/bin/bash
# usage: tools/skel.sh lib1.so lib2.so ... -- per-kernel ncu durations of bench.py (2 RK4 steps) for each library build
for L in "$@"; do
  out=gpurun_out/ncu_$(basename $L .so).csv
  PDES_LIB=$L PDES_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"k_face_flux|k_element_rk|k_fused" -c 16 --csv --log-file $out python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_times.py $out
done
