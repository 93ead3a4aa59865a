#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/euler_oracle.c, serial build).

The reference (Julia 0.6 + un-vendored packages) cannot run in the build container, so these are NOT outputs of
the reference itself: they freeze the oracle -- which is pinned against the reference's own known-answer vectors in
tests/test_oracle_golden.py and tests/golden/reference_known_answers.json -- on small seeded problems, so that
(a) a drifting oracle is caught on the CPU and (b) the CUDA path is also compared with committed numbers.

    python tests/golden/make_fixtures.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
import pdesolver_jl_b200 as pd  # noqa: E402
from common import CASES, KIND, perturbed  # noqa: E402

FIXTURES = {"c1_2d_p1_roe": (4, 1e-3), "c3_3d_p2_roe_src": (2, 5e-5), "c2_2d_p2_es": (3, 1e-3), "2d_p2_roe": (3, 1e-3),
            "3d_p1_roe_src": (3, 5e-5)}


def build(case):
    n, h = FIXTURES[case]
    dim, p, ic, opts = CASES[case]
    op = pd.build_operator(dim, p, KIND.get(case, "omega"))
    mesh = pd.structured_mesh(op, n, shuffle_seed=11)
    orc = oracle.Problem(mesh, op, dict(opts))
    q0 = perturbed(orc.exact_state(ic), amp=1e-2 if case in KIND else 1e-3)
    return op, mesh, dict(opts), orc, q0, h


def main():
    for case in FIXTURES:
        op, mesh, opts, orc, q0, h = build(case)
        res = orc.eval_residual(q0)
        t, q5, norms = orc.rk4(q0, h, 5 * h)
        t2, ql, norms_l = orc.lserk54(q0, h, 5 * h)
        np.savez_compressed(os.path.join(HERE, case + ".npz"), q0=q0, res=res, q_rk4=q5, norms_rk4=norms, t_rk4=t,
                            q_lserk=ql, norms_lserk=norms_l, h=h)
        print(case, q0.shape, float(np.linalg.norm(res)))


if __name__ == "__main__":
    main()
