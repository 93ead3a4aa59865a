#!/usr/bin/env python
"""Workload for `ncu` captures of the kernels added for SURVEY.md §8(f) rows N2 / N4 (run under ncu with a kernel filter):
one type-2 entropy-stable residual on 24.6 k p=2 tets, then a short GMRES solve on the C1 mesh."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pdesolver_jl_b200 as pd          # noqa: E402
from pdesolver_jl_b200 import ic        # noqa: E402
from common import perturbed            # noqa: E402

op = pd.build_operator(3, 2)
mesh = pd.structured_mesh(op, 16)
opts = {"Flux_name": "IRFlux", "Volume_flux_name": "IRFlux", "volume_integral_type": 2, "face_integral_type": 2,
        "FaceElementIntegral_name": "ESLFFaceIntegral", "BC1_name": "ExpBC"}
eqn = pd.EulerData(mesh, op, opts)
eqn.q[...] = perturbed(ic.ICDict["ICExp"](mesh.coords, pd.ParamType(opts)))
for _ in range(3):
    pd.evalResidual(mesh, op, eqn, opts)
eqn.close()

op = pd.build_operator(2, 1)
mesh = pd.structured_mesh(op, 50, diagonal="\\")
opts = {"Flux_name": "RoeFlux", "BC1_name": "isentropicVortexBC", "krylov_reltol": 1e-30, "krylov_itermax": 40,
        "krylov_restart": 30}
eqn = pd.EulerData(mesh, op, opts)
eqn.q[...] = ic.ICDict["ICIsentropicVortex"](mesh.coords, pd.ParamType(opts))
pd.linearSolve(mesh, op, eqn, opts, np.asfortranarray(np.random.RandomState(0).standard_normal(eqn.q.shape)))
eqn.close()
