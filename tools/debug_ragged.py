import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle, pdesolver_jl_b200 as pd
from common import CASES, perturbed
dim, p, ic, opts = CASES["c3_3d_p2_roe_src"]
op = pd.build_operator(dim, p)
for n, seed in [(5, 5), (6, 6), (7, 7), (7, None), (7, 3), (8, 3)]:
    mesh = pd.structured_mesh(op, n, shuffle_seed=seed)
    orc = oracle.Problem(mesh, op, opts)
    q0 = perturbed(orc.exact_state(ic))
    eqn = pd.EulerData(mesh, op, opts)
    eqn.q[...] = q0
    pd.evalResidual(mesh, op, eqn, opts)
    ref = orc.eval_residual(q0)
    d = np.abs(eqn.res - ref).max(axis=(0, 1)) / np.abs(ref).max()
    bad = np.nonzero(d > 1e-10)[0]
    print(n, seed, "nE", mesh.numEl, "bad elements", len(bad), bad[:20], bad[-5:] if len(bad) else "")
    if len(bad):
        bset = set(mesh.bndryfaces["element"].tolist())
        print("  bad on boundary:", sum(int(b) in bset for b in bad), "tile idx", sorted(set((bad // 16).tolist()))[:20])
        # which nodes/vars
        e = bad[0]
        print("  diff el", e, np.abs(eqn.res[:, :, e] - ref[:, :, e]).max(axis=0))
        bf = mesh.bndryfaces[mesh.bndryfaces["element"] == e]
        print("  bfaces", bf, "idx", np.nonzero(mesh.bndryfaces["element"] == e)[0])
