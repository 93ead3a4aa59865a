#!/usr/bin/env python
"""Secondary measurements (not the headline bench): throughput of the kernels behind SURVEY.md §8(f) rows N2 / N4 and of the
Jacobian-vector product, on one GPU, wall clock around synchronous public-API calls (host arrays in and out unless noted).
One JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    import pdesolver_jl_b200 as pd
    from pdesolver_jl_b200 import ic
    from common import perturbed
    out = []
    # N2: face_integral_type 2 (ESLF) + split-form volume integrals, device-resident RK4 steps
    for dim, p, n in ((2, 2, 200), (3, 2, 16)):
        op = pd.build_operator(dim, p)
        mesh = pd.structured_mesh(op, n)
        opts = {"Flux_name": "IRFlux", "Volume_flux_name": "IRFlux", "volume_integral_type": 2, "face_integral_type": 2,
                "FaceElementIntegral_name": "ESLFFaceIntegral", "use_itermax": False,
                "BC1_name": "isentropicVortexBC" if dim == 2 else "ExpBC"}
        eqn = pd.EulerData(mesh, op, opts)
        eqn.q[...] = perturbed(ic.ICDict["ICIsentropicVortex" if dim == 2 else "ICExp"](mesh.coords, pd.ParamType(opts)))
        h = 1e-5
        S = 20
        dt = timed(lambda: pd.rk4(pd.evalResidual, h, S * h, mesh, op, eqn, opts), 3)
        out.append(dict(what="N2 ESLF face-element integrals + split form, rk4() call of %d steps" % S, dim=dim, degree=p,
                        elements=int(mesh.numEl), dof=int(mesh.numDof), s_per_call=dt,
                        dof_evals_per_s=mesh.numDof * 4 * S / dt))
        eqn.close()
    # config 5: Jacobian-vector products, GMRES iterations, Newton on the C1 mesh (matrix-free)
    op = pd.build_operator(2, 1)
    mesh = pd.structured_mesh(op, 50, diagonal="\\")
    opts = {"Flux_name": "RoeFlux", "BC1_name": "isentropicVortexBC", "jac_type": 4}
    eqn = pd.EulerData(mesh, op, opts)
    q0 = ic.ICDict["ICIsentropicVortex"](mesh.coords, pd.ParamType(opts))
    eqn.q[...] = q0
    v = np.asfortranarray(np.random.RandomState(0).standard_normal(q0.shape))
    dt = timed(lambda: pd.evaldRdqProduct(mesh, op, eqn, opts, v), 20)
    out.append(dict(what="evaldRdqProduct (host v in, J v out)", elements=int(mesh.numEl), dof=int(mesh.numDof), s_per_call=dt,
                    dof_products_per_s=mesh.numDof / dt))
    kopts = dict(opts, krylov_reltol=1e-30, krylov_itermax=300, krylov_restart=30)
    t0 = time.perf_counter()
    pd.linearSolve(mesh, op, eqn, kopts, v)
    dt = time.perf_counter() - t0
    out.append(dict(what="GMRES(30), 300 iterations on the device (C1 mesh)", dof=int(mesh.numDof), s_total=dt,
                    s_per_iteration=dt / eqn.krylov_info["iterations"], iterations=eqn.krylov_info["iterations"]))
    nopts = dict(opts, itermax=10, res_abstol=1e-9, res_reltol=1e-30, krylov_reltol=1e-3, krylov_itermax=3000, krylov_restart=100)
    eqn.q[...] = q0
    t0 = time.perf_counter()
    pd.newton(pd.evalResidual, mesh, op, eqn, nopts)
    dt = time.perf_counter() - t0
    out.append(dict(what="newton() on the C1 mesh (perf/input_vals_2d_newton.jl problem, matrix-free; reference: 51.55 s with an "
                         "explicit Jacobian + sparse direct solves, perf/perf_history_2d_newton.txt:5)", s_total=dt,
                    residual_norms=[float(x) for x in eqn.convergence], **eqn.newton_info))
    # the same Newton solve with the element-block Jacobi right preconditioner (the reference: -pc_type bjacobi, right side)
    eqn.q[...] = q0
    t0 = time.perf_counter()
    pd.newton(pd.evalResidual, mesh, op, eqn, dict(nopts, krylov_pc="element_block_jacobi"))
    dt = time.perf_counter() - t0
    out.append(dict(what="newton() on the C1 mesh with the element-block Jacobi right preconditioner", s_total=dt,
                    residual_norms=[float(x) for x in eqn.convergence], **eqn.newton_info))
    dt = timed(lambda: pd.diagnostics(mesh, op, eqn, opts), 20)
    out.append(dict(what="diagnostics() (host q in, residual + 5 functionals)", dof=int(mesh.numDof), s_per_call=dt))
    eqn.close()
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
