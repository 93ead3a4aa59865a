#!/bin/bash
# C1 (60 k DOF, latency-bound): persistent face kernel at every size (PDES_FACE_SMALL=0) against the small-launch switch
run() {
  out=$(env $1 python bench.py --workload c1_2d_p1_roe --no-cpu-baseline --no-parity --steps 400 --warmup 40 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['ms_per_step_blocks']['median'], d['gpu_launches'])")
  echo "AB [$1] $out"
}
run "PDES_FACE_SMALL=0"
run "PDES_FACE_SMALL=1"
run "PDES_FACE_SMALL=0"
run "PDES_FACE_SMALL=1"
