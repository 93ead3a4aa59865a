#!/bin/bash
for lib in "$@"; do
PDES_LIB=$PWD/$lib python - <<'PY'
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, pdesolver_jl_b200 as pd
from pdesolver_jl_b200 import ic
from common import perturbed
for dim, p, n in ((2, 2, 200), (3, 2, 16), (3, 1, 24), (2, 1, 300)):
    op = pd.build_operator(dim, p)
    mesh = pd.structured_mesh(op, n)
    opts = {"Flux_name": "IRFlux", "Volume_flux_name": "IRFlux", "volume_integral_type": 2, "face_integral_type": 2,
            "FaceElementIntegral_name": "ESLFFaceIntegral", "use_itermax": False,
            "BC1_name": "isentropicVortexBC" if dim == 2 else "ExpBC"}
    eqn = pd.EulerData(mesh, op, opts)
    eqn.q[...] = perturbed(ic.ICDict["ICIsentropicVortex" if dim == 2 else "ICExp"](mesh.coords, pd.ParamType(opts)))
    h, S = 1e-5, 20
    pd.rk4(pd.evalResidual, h, S * h, mesh, op, eqn, opts)
    t0 = time.perf_counter()
    for _ in range(3):
        pd.rk4(pd.evalResidual, h, S * h, mesh, op, eqn, opts)
    dt = (time.perf_counter() - t0) / 3
    print(os.environ["PDES_LIB"].split("/")[-1], dim, p, mesh.numEl, "DOF-evals/s %.3e" % (mesh.numDof * 4 * S / dt))
PY
done
