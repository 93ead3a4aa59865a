#!/bin/bash
# A/B of library builds: r2_ab_lib.sh lib1.so lib2.so ...   (PDES_LIB selects the library; one small parity check each)
for lib in "$@"; do
  export PDES_LIB=$PWD/$lib
  par=$(python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, oracle, pdesolver_jl_b200 as pd
from common import CASES, perturbed, rel_l2
dim,p,ic,opts=CASES["c3_3d_p2_roe_src"]; opts=dict(opts)
op=pd.build_operator(dim,p); mesh=pd.structured_mesh(op,7,shuffle_seed=3)
orc=oracle.Problem(mesh,op,opts); q0=perturbed(orc.exact_state(ic))
eqn=pd.EulerData(mesh,op,opts); eqn.q[...]=q0
pd.evalResidual(mesh,op,eqn,opts); e1=rel_l2(eqn.res, orc.eval_residual(q0))
opts["use_itermax"]=False; eqn.q[...]=q0
pd.rk4(pd.evalResidual,5e-5,3*5e-5,mesh,op,eqn,opts); _,qr,_=orc.rk4(q0,5e-5,3*5e-5)
print(f"{e1:.1e} {rel_l2(eqn.q,qr):.1e}")
PY
)
  out=$(python bench.py --no-cpu-baseline --no-parity --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['ms_per_step_blocks']['median'], d['roofline']['frac'])")
  echo "ABLIB [$lib] parity $par bench $out"
done
