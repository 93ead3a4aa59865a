"""Residual parity at a larger size (oracle takes seconds): usage check_large.py <case> <n>"""
import sys, time, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle, pdesolver_jl_b200 as pd
from common import CASES, KIND, perturbed, rel_l2
case, n = sys.argv[1], int(sys.argv[2])
dim, p, ic, opts = CASES[case]
op = pd.build_operator(dim, p, KIND.get(case, "omega"))
mesh = pd.structured_mesh(op, n)
orc = oracle.Problem(mesh, op, opts)
q0 = perturbed(orc.exact_state(ic), amp=1e-2 if case in KIND else 1e-3)
eqn = pd.EulerData(mesh, op, opts)
eqn.q[...] = q0
pd.evalResidual(mesh, op, eqn, opts)
t = time.time(); ref = orc.eval_residual(q0, omp=True); dt = time.time() - t
print(case, n, "nE", mesh.numEl, "rel-L2", rel_l2(eqn.res, ref), "oracle s", round(dt, 2))
