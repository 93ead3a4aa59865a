"""Pins the oracle's face-element integrals (face_integral_type = 2 on SBP-Omega operators) on
the identities the reference's own tests assert for them (test/euler/test_ESS.jl):

  :9-88, :384-441    calcECFaceIntegral == the matrix form  (P_gamma^T R^T N_d W R P_nu) o F  of calcECFaceIntegralTest
  :445-483           lemma 3: the entropy-variable contraction of the EC integral equals the jump of the potential flux
  :583-705           the penalty is conservative (sum resL = -sum resR) and entropy dissipative (w^T res < eps)
  :923-930           a uniform state gives a zero residual
"""
import dataclasses

import numpy as np
import pytest

import oracle
import pdesolver_jl_b200 as pd
from common import perturbed

G, G1 = 1.4, 0.4


def es_opts(name):
    return {"Flux_name": "IRFlux", "Volume_flux_name": "IRFlux", "volume_integral_type": 2, "face_integral_type": 2,
            "FaceElementIntegral_name": name, "BC1_name": "isentropicVortexBC"}


def one_interface(mesh, i):
    """The same mesh with interface i only (per-interface resL / resR of getFaceElementIntegral)."""
    return dataclasses.replace(mesh, interfaces=mesh.interfaces[i:i + 1].copy(),
                               nrm_face=np.asfortranarray(mesh.nrm_face[:, :, i:i + 1]))


def ir_flux(qL, qR, n):
    """bc_solvers.jl:725-874 calcEulerFlux_IR through the oracle's node-level function"""
    dim = len(n)
    F = np.zeros(dim + 2)
    L = oracle.lib()
    L.orc_ir_flux(dim, G, oracle._ptr(np.ascontiguousarray(qL)), oracle._ptr(np.ascontiguousarray(qR)),
                  oracle._ptr(np.ascontiguousarray(n, dtype=np.float64)), 1, oracle._ptr(F))
    return F


def entropy_vars(q):
    """convertToEntropy_ / gamma_1 (conversion.jl:50-89; test_ESS.jl:190-210) for q[nd, nn]"""
    dim = q.shape[0] - 2
    rho, mom, E = q[0], q[1:1 + dim], q[dim + 1]
    rho_int = E - 0.5 * (mom ** 2).sum(axis=0) / rho
    s = np.log(G1 * rho_int / rho ** G)
    return np.concatenate([((rho_int * (G + 1 - s) - E) / rho_int)[None], mom / rho_int, (-rho / rho_int)[None]]) / G1


CASES = [(2, 1, 3), (2, 2, 2), (3, 1, 2), (3, 2, 1)]


def setup(dim, p, n, name):
    op = pd.build_operator(dim, p)
    mesh = pd.structured_mesh(op, n, shuffle_seed=6)
    P = oracle.Problem(mesh, op, es_opts(name))
    q0 = perturbed(P.exact_state("ICIsentropicVortex" if dim == 2 else "ICExp"), amp=2e-2)
    return op, mesh, q0


@pytest.mark.parametrize("dim,p,n", CASES)
def test_ec_integral_matches_matrix_form(dim, p, n):
    op, mesh, q0 = setup(dim, p, n, "ECFaceIntegral")
    f = op.face
    nn, nd = op.numnodes, dim + 2
    # R' (test_ESS.jl: params.Rprime): face node k <- stencil node j, as a [nfn, nn] matrix in stencil order
    Rp = f.interp.T
    for i in range(0, mesh.numInterfaces, max(1, mesh.numInterfaces // 7)):
        m1 = one_interface(mesh, i)
        res = oracle.Problem(m1, op, es_opts("ECFaceIntegral")).face_integrals(q0)
        it = mesh.interfaces[i]
        eL, eR, fL, fR, o = int(it["elementL"]), int(it["elementR"]), int(it["faceL"]), int(it["faceR"]), int(it["orient"])
        qL, qR = q0[:, :, eL], q0[:, :, eR]
        refL, refR = np.zeros((nd, nn)), np.zeros((nd, nn))
        for d in range(dim):
            nrm = np.zeros(dim)
            nrm[d] = 1.0
            A = (f.wface * mesh.nrm_face[d, :, i])[:, None] * Rp[f.nbrperm[:, o], :]      # N_d W R' (rows through nbrperm)
            B = Rp.T @ A                                                                   # stencil x stencil
            for a in range(nn):
                for b in range(nn):
                    F = ir_flux(qL[:, f.perm[a, fL]], qR[:, f.perm[b, fR]], nrm)
                    refL[:, f.perm[a, fL]] -= B[a, b] * F
                    refR[:, f.perm[b, fR]] += B[a, b] * F
        assert np.linalg.norm(res[:, :, eL] - refL) < 1e-12 * refL.size
        assert np.linalg.norm(res[:, :, eR] - refR) < 1e-12 * refR.size
        others = np.delete(res, [eL, eR], axis=2)
        assert not others.any()


@pytest.mark.parametrize("dim,p,n", CASES)
def test_ec_integral_potential_flux_identity(dim, p, n):
    """lemma 3 (test_ESS.jl:445-483): -(wL.resL + wR.resR) = sum_d sum_k wface_k nrm[d,k] (psi_d(qL) - psi_d(qR)) at the face
    nodes, psi_d = momentum_d for the IR entropy variables (test_ESS.jl:91-101)"""
    op, mesh, q0 = setup(dim, p, n, "ECFaceIntegral")
    f = op.face
    for i in range(0, mesh.numInterfaces, max(1, mesh.numInterfaces // 9)):
        res = oracle.Problem(one_interface(mesh, i), op, es_opts("ECFaceIntegral")).face_integrals(q0)
        it = mesh.interfaces[i]
        eL, eR, fL, fR, o = int(it["elementL"]), int(it["elementR"]), int(it["faceL"]), int(it["faceR"]), int(it["orient"])
        lhs = (entropy_vars(q0[:, :, eL]) * res[:, :, eL]).sum() + (entropy_vars(q0[:, :, eR]) * res[:, :, eR]).sum()
        rhs = 0.0
        for d in range(dim):
            psiL = f.interp.T @ q0[1 + d, f.perm[:, fL], eL]                      # at the face nodes of elementL
            psiR = (f.interp.T @ q0[1 + d, f.perm[:, fR], eR])[f.nbrperm[:, o]]   # elementR's node nbrperm[k] = L's node k
            rhs += (f.wface * mesh.nrm_face[d, :, i] * (psiL - psiR)).sum()
        assert abs(-lhs - rhs) < 1e-12


@pytest.mark.parametrize("dim,p,n", CASES)
def test_penalty_is_conservative_and_dissipative(dim, p, n):
    op, mesh, q0 = setup(dim, p, n, "ELFPenaltyFaceIntegral")
    for i in range(0, mesh.numInterfaces, max(1, mesh.numInterfaces // 9)):
        res = oracle.Problem(one_interface(mesh, i), op, es_opts("ELFPenaltyFaceIntegral")).face_integrals(q0)
        it = mesh.interfaces[i]
        eL, eR = int(it["elementL"]), int(it["elementR"])
        assert np.allclose(res[:, :, eL].sum(axis=1), -res[:, :, eR].sum(axis=1), rtol=0, atol=1e-13)   # :638
        ds = (entropy_vars(q0[:, :, eL]) * res[:, :, eL]).sum() + (entropy_vars(q0[:, :, eR]) * res[:, :, eR]).sum()
        assert ds < np.finfo(float).eps                                                                  # :693
        assert ds < -1e-12            # the perturbed state has a jump at every interface: strictly dissipative
    # ESLF = EC + penalty (calcESFaceIntegral, faceElementIntegrals.jl:160-176)
    P = {k: oracle.Problem(mesh, op, es_opts(k)) for k in ("ECFaceIntegral", "ELFPenaltyFaceIntegral", "ESLFFaceIntegral")}
    tot = P["ECFaceIntegral"].face_integrals(q0) + P["ELFPenaltyFaceIntegral"].face_integrals(q0)
    assert np.allclose(P["ESLFFaceIntegral"].face_integrals(q0), tot, rtol=1e-13, atol=1e-14)
    # zero penalty for a continuous state (test_ESS.jl:640-648: zero_penalty)
    qc = np.asfortranarray(np.broadcast_to(q0[:, :1, :1], q0.shape))
    assert np.abs(P["ELFPenaltyFaceIntegral"].face_integrals(qc)).max() < 1e-13


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (3, 1), (3, 2)])
@pytest.mark.parametrize("name", ["ECFaceIntegral", "ESLFFaceIntegral"])
def test_uniform_state_zero_residual(dim, p, name):
    """factRes0 (test_ESS.jl:923-930)"""
    op = pd.build_operator(dim, p)
    mesh = pd.structured_mesh(op, 2, shuffle_seed=1)
    opts = dict(es_opts(name), BC1_name="FreeStreamBC", Ma=0.4, aoa=10.0)
    P = oracle.Problem(mesh, op, opts)
    assert np.abs(P.eval_residual(P.exact_state("ICFreeStream"))).max() < 1e-13


def test_entropy_conservation_of_the_ec_scheme():
    """Interior of the EC scheme (split-form volume + EC face integrals): w^T (volume + face) equals the boundary potential
    flux only -- the interior interfaces cancel (test_ESS.jl:487-572).  With the penalty the same sum decreases."""
    op = pd.build_operator(2, 2)
    mesh = pd.structured_mesh(op, 3, shuffle_seed=3)
    P = oracle.Problem(mesh, op, es_opts("ECFaceIntegral"))
    q0 = perturbed(P.exact_state("ICIsentropicVortex"), amp=2e-2)
    w = np.stack([entropy_vars(q0[:, :, e]) for e in range(mesh.numEl)], axis=2)
    vol = P.volume_integrals(q0, precompute=False)
    ec = P.face_integrals(q0)
    # boundary potential flux: sum over boundary faces of wface nrm . psi at the face nodes
    f = op.face
    bpsi = 0.0
    for b, bf in enumerate(mesh.bndryfaces):
        e, fc = int(bf["element"]), int(bf["face"])
        for d in range(2):
            bpsi += (f.wface * mesh.nrm_bndry[d, :, b] * (f.interp.T @ q0[1 + d, f.perm[:, fc], e])).sum()
    total = (w * (vol + ec)).sum()
    assert abs(total - bpsi) < 1e-11
    pen = oracle.Problem(mesh, op, es_opts("ELFPenaltyFaceIntegral")).face_integrals(q0)
    assert (w * pen).sum() < -1e-10


# --- Lax-Wendroff entropy kernel (ELW2PenaltyFaceIntegral / ESLW2FaceIntegral) ------------------------------------------
def _jac_x(q, gamma=1.4, h=1e-7):
    """dF_x/dq by central differences of the oracle's Euler flux (the reference uses the complex step, test_3d.jl:78-104)."""
    import ctypes as C
    L = oracle.lib()
    nd = len(q)
    dim = nd - 2
    dirx = np.zeros(dim)
    dirx[0] = 1.0
    A = np.zeros((nd, nd))
    for j in range(nd):
        fp, fm = np.zeros(nd), np.zeros(nd)
        dq = np.zeros(nd)
        dq[j] = h
        qp, qm = q + dq, q - dq
        L.orc_euler_flux(dim, gamma, qp.ctypes.data_as(C.c_void_p), dirx.ctypes.data_as(C.c_void_p), fp.ctypes.data_as(C.c_void_p))
        L.orc_euler_flux(dim, gamma, qm.ctypes.data_as(C.c_void_p), dirx.ctypes.data_as(C.c_void_p), fm.ctypes.data_as(C.c_void_p))
        A[:, j] = (fp - fm) / (2 * h)
    return A


@pytest.mark.parametrize("q", [[1.0, 0.3, -0.2, 2.0], [1.2, -0.5, 0.1, 3.0], [1.1, 0.3, -0.2, 0.4, 2.5], [0.9, -0.1, 0.6, -0.3, 2.2]])
def test_eigensystem_identities(q):
    """test_3d.jl:107-148 / test_lowlevel.jl:440-470: Y Lambda Y^-1 is the flux Jacobian and Y S2 Y^T = A0 = dq/dw."""
    import ctypes as C
    L = oracle.lib()
    q = np.array(q)
    nd = len(q)
    dim = nd - 2
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    Y = np.zeros((nd, nd), order="F")
    lam, S2 = np.zeros(nd), np.zeros(nd)
    A0 = np.zeros((nd, nd), order="F")
    L.orc_evecs_x(dim, 1.4, ptr(q), ptr(Y))
    L.orc_evals_x(dim, 1.4, ptr(q), ptr(lam))
    L.orc_escaling_x(dim, 1.4, ptr(q), ptr(S2))
    L.orc_ira0(dim, 1.4, ptr(q), ptr(A0))
    assert np.allclose(Y @ np.diag(lam) @ np.linalg.inv(Y), _jac_x(q), atol=2e-7)
    assert np.allclose(Y @ np.diag(S2) @ Y.T, A0, atol=1e-12)


@pytest.mark.parametrize("dim", [2, 3])
def test_lw2_kernel_is_rotated_eigen_dissipation(dim):
    """applyEntropyKernel(LW2Kernel) (faceElementIntegrals.jl:393-440) = |n| Y_n |Lambda_n| S2 Y_n^T dw with the eigensystem
    of the flux Jacobian in direction n: symmetric positive semi-definite (dw^T flux >= 0), homogeneous of degree 1 in n,
    bounded by the Lax-Friedrichs kernel (|lambda| <= lambda_max) and equal to the x-direction formula for n = e_x."""
    import ctypes as C
    L = oracle.lib()
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.RandomState(5)
    nd = dim + 2
    q = np.array([1.0, 0.3, -0.2, 2.0] if dim == 2 else [1.1, 0.3, -0.2, 0.4, 2.5])
    for trial in range(6):
        n = rng.standard_normal(dim) * 0.7
        if trial == 0:
            n = np.eye(dim)[0] * 0.5
        M = np.zeros((nd, nd))
        for j in range(nd):
            e = np.zeros(nd)
            e[j] = 1.0
            f = np.zeros(nd)
            L.orc_lw2_entropy_kernel(dim, 1.4, ptr(q), ptr(e), ptr(n), ptr(f))
            M[:, j] = f
        assert np.allclose(M, M.T, atol=1e-12)
        ev = np.linalg.eigvalsh(0.5 * (M + M.T))
        assert ev.min() > -1e-12
        f2 = np.zeros(nd)
        dw = rng.standard_normal(nd)
        n2 = 3.0 * n
        L.orc_lw2_entropy_kernel(dim, 1.4, ptr(q), ptr(dw), ptr(n2), ptr(f2))
        assert np.allclose(f2, 3.0 * (M @ dw), rtol=1e-12, atol=1e-13)
        # Lax-Friedrichs bound: lambda_max A0 - M is positive semi-definite
        A0 = np.zeros((nd, nd), order="F")
        L.orc_ira0(dim, 1.4, ptr(q), ptr(A0))
        L.orc_lambda_max.restype = C.c_double
        lmax = L.orc_lambda_max(dim, 1.4, ptr(q), ptr(n))
        assert np.linalg.eigvalsh(lmax * A0 - M).min() > -1e-10
        if trial == 0:
            Y = np.zeros((nd, nd), order="F")
            lam, S2 = np.zeros(nd), np.zeros(nd)
            L.orc_evecs_x(dim, 1.4, ptr(q), ptr(Y))
            L.orc_evals_x(dim, 1.4, ptr(q), ptr(lam))
            L.orc_escaling_x(dim, 1.4, ptr(q), ptr(S2))
            assert np.allclose(M, 0.5 * Y @ np.diag(np.abs(lam) * S2) @ Y.T, atol=1e-12)


@pytest.mark.parametrize("dim,p,n", [(2, 1, 3), (2, 2, 2), (3, 1, 2), (3, 2, 1)])
def test_lw2_penalty_is_conservative_and_dissipative(dim, p, n):
    """runESTest(..., penalty_lw2) (test_ESS.jl:583-705, 993): the Lax-Wendroff penalty conserves, dissipates entropy and
    vanishes for a continuous state; ESLW2 = EC + penalty."""
    op, mesh, q0 = setup(dim, p, n, "ELW2PenaltyFaceIntegral")
    for i in range(0, mesh.numInterfaces, max(1, mesh.numInterfaces // 9)):
        res = oracle.Problem(one_interface(mesh, i), op, es_opts("ELW2PenaltyFaceIntegral")).face_integrals(q0)
        it = mesh.interfaces[i]
        eL, eR = int(it["elementL"]), int(it["elementR"])
        assert np.allclose(res[:, :, eL].sum(axis=1), -res[:, :, eR].sum(axis=1), rtol=0, atol=1e-13)
        ds = (entropy_vars(q0[:, :, eL]) * res[:, :, eL]).sum() + (entropy_vars(q0[:, :, eR]) * res[:, :, eR]).sum()
        assert ds < -1e-12
    P = {k: oracle.Problem(mesh, op, es_opts(k)) for k in ("ECFaceIntegral", "ELW2PenaltyFaceIntegral", "ESLW2FaceIntegral")}
    tot = P["ECFaceIntegral"].face_integrals(q0) + P["ELW2PenaltyFaceIntegral"].face_integrals(q0)
    assert np.allclose(P["ESLW2FaceIntegral"].face_integrals(q0), tot, rtol=1e-13, atol=1e-14)
    qc = np.asfortranarray(np.broadcast_to(q0[:, :1, :1], q0.shape))
    assert np.abs(P["ELW2PenaltyFaceIntegral"].face_integrals(qc)).max() < 1e-13


def test_projection_matrix_properties():
    """test/euler/Utils.jl:236-325 (test_utils_projection): for normals swept over the circle / sphere the projection to
    normal-tangential coordinates is orthogonal, keeps density and energy, keeps the momentum magnitude, its second row is
    the normal and the tangent is a unit vector orthogonal to it; projectToXY inverts projectToNT."""
    import ctypes as C
    L = oracle.lib()
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    q2, q3 = np.array([1.0, 2.0, 3.0, 7.0]), np.array([1.0, 2.0, 3.0, 4.0, 15.0])
    normals = [(2, np.array([np.cos(t), np.sin(t)])) for t in np.arange(0.0, 2 * np.pi, 0.1)]
    normals += [(3, np.array([np.sin(t) * np.cos(f), np.sin(t) * np.sin(f), np.cos(t)]))
                for t in np.arange(0.0, np.pi, 0.1) for f in np.arange(0.0, 2 * np.pi, 0.1)]
    for dim, n in normals:
        nd = dim + 2
        P = np.zeros((nd, nd), order="F")
        L.orc_projection_matrix(dim, ptr(n), ptr(P))
        assert np.allclose(P @ P.T, np.eye(nd), atol=1e-12)
        assert np.allclose(P[1, 1:1 + dim], n, atol=1e-15)
        t = P[2, 1:1 + dim]
        assert abs(np.linalg.norm(t) - 1.0) < 1e-12 and abs(t @ n) < 1e-13
        q = q2 if dim == 2 else q3
        qp = P @ q
        assert qp[0] == q[0] and qp[-1] == q[-1]
        assert abs(np.sum(qp[1:-1] ** 2) - np.sum(q[1:-1] ** 2)) < 1e-12
        assert np.allclose(P.T @ qp, q, atol=1e-12)
