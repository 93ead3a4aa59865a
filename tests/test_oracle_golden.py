"""Pins the CPU oracle against every known-answer vector and identity the
reference's own tests hold for the hot path."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import _ptr
from pdesolver_jl_b200 import mesh as pmesh
from pdesolver_jl_b200 import sbp

G = 1.4
L = oracle.lib()


def pressure(q):
    q = np.asarray(q, float)
    return L.orc_calc_pressure(len(q) - 2, G, _ptr(q))


def euler_flux(q, n):
    q, n = np.asarray(q, float), np.asarray(n, float)
    F = np.zeros(len(q))
    L.orc_euler_flux(len(q) - 2, G, _ptr(q), _ptr(n), _ptr(F))
    return F


def roe(qL, qR, n):
    qL, qR, n = (np.asarray(a, float) for a in (qL, qR, n))
    F = np.zeros(len(qL))
    L.orc_roe_solver(len(qL) - 2, G, _ptr(qL), _ptr(qR), _ptr(n), _ptr(F))
    return F


def ir(qL, qR, n):
    qL, qR = np.asarray(qL, float), np.asarray(qR, float)
    n = np.asfortranarray(np.asarray(n, float))
    dim = len(qL) - 2
    ndir = 1 if n.ndim == 1 else n.shape[1]
    F = np.zeros((len(qL), ndir), order="F")
    L.orc_ir_flux(dim, G, _ptr(qL), _ptr(qR), _ptr(n), ndir, _ptr(F))
    return F[:, 0] if n.ndim == 1 else F


def irslf(qL, qR, n):
    qL, qR, n = (np.asarray(a, float) for a in (qL, qR, n))
    F = np.zeros(len(qL))
    L.orc_irslf_flux(len(qL) - 2, G, _ptr(qL), _ptr(qR), _ptr(n), _ptr(F))
    return F


# --- test/euler/test_lowlevel.jl:491-521 -------------------------------------
def test_pressure_and_flux_2d():
    q = [1.0, 2.0, 3.0, 7.0]
    assert abs(pressure(q) - 0.2) < 1e-14
    assert np.allclose(euler_flux(q, [1.0, 0.0]), [2.0, 4.2, 6, 14.4], atol=1e-14, rtol=0)


# --- test/euler/test_3d.jl:163-201 -------------------------------------------
def test_pressure_and_flux_3d():
    q = [1.0, 2, 3, 4, 15]
    assert abs(pressure(q) - 0.2) < 1e-12
    assert np.allclose(euler_flux(q, [1, 0, 0]), [2, 4.2, 6, 8, 30.4], atol=1e-12, rtol=0)
    assert np.allclose(euler_flux(q, [0, 1, 0]), [3, 6, 9.2, 12, 45.6], atol=1e-12, rtol=0)
    assert np.allclose(euler_flux(q, [0, 0, 1]), [4, 8, 12, 16.2, 60.8], atol=1e-12, rtol=0)


# --- test/euler/test_lowlevel.jl:626-650 --------------------------------------
def test_isentropic_vortex_point():
    s = np.zeros(4)
    L.orc_isentropic_vortex(2, G, 287.058, _ptr(np.array([1.0, 0.0])), _ptr(s))
    assert np.allclose(s, [2.0, 0.0, -1.3435, 2.236960], atol=1e-4, rtol=0)


# --- test/euler/test_lowlevel.jl:94-146 ---------------------------------------
def test_ir_variables():
    q = np.array([1.0, 2.0, 3.0, 7.0])
    v_analytic = np.array([-2 * 4.99528104378295, 4.0, 6, -2 * 1])
    w = np.zeros(4)
    L.orc_convert_to_ir(2, G, _ptr(q), _ptr(w))
    assert np.linalg.norm(w - v_analytic / (G - 1)) < 1e-12
    # in-place operation is allowed by the reference
    q2 = q.copy()
    L.orc_convert_to_ir(2, G, _ptr(q2), _ptr(q2))
    assert np.linalg.norm(q2 - v_analytic / (G - 1)) < 1e-12


def test_ira0_is_symmetric_dq_dw():
    """getIRA0 = dq/dw (IR_stab.jl:15-110): check against finite differences of
    the inverse map w(q)."""
    for q in ([1.0, 0.3, -0.2, 2.0], [1.1, 0.3, -0.2, 0.4, 2.5]):
        q = np.array(q)
        nd = len(q)
        A0 = np.zeros((nd, nd), order="F")
        L.orc_ira0(nd - 2, G, _ptr(q), _ptr(A0))
        assert np.allclose(A0, A0.T, atol=1e-13)
        J = np.zeros((nd, nd))      # dw/dq
        for j in range(nd):
            dq = np.zeros(nd)
            dq[j] = 1e-6
            wp, wm = np.zeros(nd), np.zeros(nd)
            L.orc_convert_to_ir(nd - 2, G, _ptr(q + dq), _ptr(wp))
            L.orc_convert_to_ir(nd - 2, G, _ptr(q - dq), _ptr(wm))
            J[:, j] = (wp - wm) / 2e-6
        assert np.allclose(A0 @ J, np.eye(nd), atol=1e-7)


# --- test/euler/test_lowlevel.jl:529-617, test_dg.jl:11-62 --------------------
def test_roe_consistency_and_bcs():
    q = np.array([1.0, 2.0, 3.0, 7.0])
    nrm = np.array([1.0, 0.0])          # dxidx^T*[1,0] for element 1 of tri2l
    assert np.allclose(roe(q, q, nrm), euler_flux(q, nrm), atol=1e-13)
    op = sbp.build_operator(2, 1)
    m = pmesh.two_element_mesh(op)
    P = oracle.Problem(m, op, {"BC1_name": "isentropicVortexBC"})
    coords = np.array([1.0, 0.0])
    qv = np.zeros(4)
    L.orc_isentropic_vortex(2, G, 287.058, _ptr(coords), _ptr(qv))
    F = np.zeros(4)
    L.orc_bc_flux(P.ref(), oracle.BC_IDS["isentropicVortexBC"], _ptr(qv), _ptr(coords), _ptr(nrm), _ptr(F))
    assert np.allclose(F, euler_flux(qv, nrm), atol=1e-13)
    qw = qv.copy()
    qw[2] = 0          # flow parallel to the wall x = const
    L.orc_bc_flux(P.ref(), oracle.BC_IDS["noPenetrationBC"], _ptr(qw), _ptr(coords), _ptr(nrm), _ptr(F))
    qproj = qw.copy()
    qproj[1] = 0.0
    assert np.allclose(F, euler_flux(qproj, nrm), atol=1e-13)
    q3 = np.array([1.0, 2, 3, 4, 15])
    n3 = np.array([0.3, -0.2, 0.9])
    assert np.allclose(roe(q3, q3, n3), euler_flux(q3, n3), atol=1e-13)


# --- test/euler/test_flux.jl:314-367 (testRoe) --------------------------------
@pytest.mark.parametrize("dim", [2, 3])
def test_roe_upwinding(dim):
    if dim == 2:
        qL, nrm = np.array([1.0, 2.0, 3.0, 7.0]), np.array([1.0, 0.0])
        qI = np.array([1.0, 1.0, 1.0, 7.0])
    else:
        qL, nrm = np.array([1.0, 2, 3, 4, 15]), np.array([1.0, 0.0, 0.0])
        qI = np.array([1.0, 1, 1, 1, 15])
    qR = qL + 1
    assert np.linalg.norm(roe(qL, qR, nrm) - euler_flux(qL, nrm)) < 1e-13
    f = roe(qR, qL, nrm)
    assert np.linalg.norm(f - euler_flux(qR, nrm)) < 1e-13
    assert np.linalg.norm(f + roe(qL, qR, -nrm)) < 1e-13
    assert np.linalg.norm(roe(qI, qI + 1, nrm) + roe(qI + 1, qI, -nrm)) < 1e-13


# --- test/euler/test_flux.jl:4-16,87-111: independent slow IR flux -------------
def _logmean(aL, aR):
    xi = aL / aR
    f = (xi - 1) / (xi + 1)
    u = f * f
    if u < 1e-2:
        F = 1.0 + u / 3.0 + u * u / 5.0 + u * u * u / 7.0
    else:
        F = np.log(xi) / 2.0 / f
    return (aL + aR) / (2 * F)


def _ir_flux_slow(qL, qR, nrm):
    pL, pR = pressure(qL), pressure(qR)
    z5_ln = _logmean(np.sqrt(qL[0] * pL), np.sqrt(qR[0] * pR))
    z1L, z1R = np.sqrt(qL[0] / pL), np.sqrt(qR[0] / pR)
    rho_hat = 0.5 * (z1L + z1R) * z5_ln
    z1_avg = 0.5 * (z1L + z1R)
    u_hat = 0.5 * (z1L * qL[1] / qL[0] + z1R * qR[1] / qR[0]) / z1_avg
    v_hat = 0.5 * (z1L * qL[2] / qL[0] + z1R * qR[2] / qR[0]) / z1_avg
    p1_hat = 0.5 * (np.sqrt(qL[0] * pL) + np.sqrt(qR[0] * pR)) / z1_avg
    z1_ln = _logmean(z1L, z1R)
    p2_hat = (G + 1) * z5_ln / (2 * G * z1_ln) + (G - 1) * 0.5 * (np.sqrt(qL[0] * pL) + np.sqrt(qR[0] * pR)) / (2 * G * z1_avg)
    h_hat = G * p2_hat / (rho_hat * (G - 1)) + 0.5 * (u_hat * u_hat + v_hat * v_hat)
    fx = np.array([rho_hat * u_hat, rho_hat * u_hat * u_hat + p1_hat, rho_hat * u_hat * v_hat, rho_hat * u_hat * h_hat])
    fy = np.array([rho_hat * v_hat, rho_hat * u_hat * v_hat, rho_hat * v_hat * v_hat + p1_hat, rho_hat * v_hat * h_hat])
    return fx * nrm[0] + fy * nrm[1]


def test_ir_flux_against_slow_version():
    qL = np.array([1.0, 2.0, 3.0, 7.0])
    qR = qL + 1
    nrm = np.array([1.0, 2.0])
    assert np.allclose(ir(qL, qR, nrm), _ir_flux_slow(qL, qR, nrm), atol=1e-12, rtol=0)
    # logavg: both branches of the series switch agree
    for a, b in [(1.0, 1.0 + 1e-9), (1.0, 1.06), (1.0, 1.07), (2.0, 0.3)]:
        exact = (a - b) / np.log(a / b)
        assert abs(L.orc_logavg(a, b) - exact) < 1e-12 * max(1.0, abs(exact)) + 1e-7 * (abs(a - b) < 1e-8)


# --- test/euler/test_flux.jl:26-79: symmetry / antisymmetry / consistency / multi-D
@pytest.mark.parametrize("dim", [2, 3])
def test_two_point_flux_properties(dim):
    rng = np.random.default_rng(5)
    if dim == 2:
        qL, nrm = np.array([1.0, 2.0, 3.0, 7.0]), np.array([1.0, 1.0])
    else:
        qL, nrm = np.array([1.0, 2, 3, 4, 15]), np.array([1.0, 1, 1])
    qR = qL + 1
    assert np.allclose(ir(qL, qR, nrm), ir(qR, qL, nrm), atol=1e-12)
    assert np.allclose(ir(qL, qR, nrm), -ir(qR, qL, -nrm), atol=1e-12)
    assert np.allclose(ir(qL, qL, nrm), euler_flux(qL, nrm), atol=1e-12)
    nD = np.asfortranarray(rng.random((dim, dim)))
    FD = ir(qL, qR, nD)
    for i in range(dim):
        assert np.allclose(FD[:, i], ir(qL, qR, nD[:, i].copy()), atol=1e-13)
    # IRSLF: consistent, conservative under (swap, -n), and dissipative in entropy
    assert np.allclose(irslf(qL, qL, nrm), euler_flux(qL, nrm), atol=1e-12)
    assert np.allclose(irslf(qL, qR, nrm), -irslf(qR, qL, -nrm), atol=1e-11)
    wL, wR = np.zeros(dim + 2), np.zeros(dim + 2)
    L.orc_convert_to_ir(dim, G, _ptr(qL), _ptr(wL))
    L.orc_convert_to_ir(dim, G, _ptr(qR), _ptr(wR))
    assert (wL - wR) @ (irslf(qL, qR, nrm) - ir(qL, qR, nrm)) >= 0


# --- test/euler/test_lowlevel.jl:685-827: integral-level goldens (tri2l mesh) --
def _tri2l(kind):
    op = sbp.build_operator(2, 1, kind)
    m = pmesh.two_element_mesh(op)
    opts = {"Flux_name": "RoeFlux", "BC1_name": "FreeStreamBC", "Ma": 0.5, "aoa": 45.0}
    P = oracle.Problem(m, op, opts)
    q = np.zeros(P.shape, order="F")
    q[:] = np.array([1.0, 0.35355, 0.35355, 2.0])[:, None, None]   # ICRho1E2U3
    return op, m, P, q


def test_golden_flux_parametric_and_volume_blocks():
    op, m, P, q = _tri2l("gamma")
    fp = P.euler_flux_parametric(q)
    # reference element 2 is our element index 1, element 1 is index 0
    for i in range(3):
        assert np.allclose(fp[:, i, 0, 1], [0.0, -0.750001, 0.750001, 0.0], atol=1e-5)
        assert np.allclose(fp[:, i, 1, 1], [0.35355, 0.12499, 0.874999, 0.972263], atol=1e-5)
        assert np.allclose(fp[:, i, 0, 0], [0.35355, 0.874999, 0.124998, .972263], atol=1e-5)
        assert np.allclose(fp[:, i, 1, 0], [0.0, 0.750001, -0.750001, 0.0], atol=1e-5)
    res = P.volume_integrals(q)
    el1_res = np.array([[-0.35355, 0, 0.35355], [-0.874999, 0.750001, 0.124998],
                        [-0.124998, -0.750001, 0.874999], [-0.972263, 0, 0.972263]])
    el2_res = np.array([[-0.35355, 0.35355, 0], [-0.124998, 0.874999, -0.75001],
                        [-0.874999, 0.124998, 0.75001], [-0.972263, 0.972263, 0]])
    assert np.allclose(res[:, :, 1], el1_res, atol=1e-4)
    assert np.allclose(res[:, :, 0], el2_res, atol=1e-4)
    assert np.allclose(P.volume_integrals(q, precompute=False), res, atol=1e-14)


def test_golden_boundary_flux_and_blocks():
    op, m, P, q = _tri2l("gamma")
    # the golden bndryflux of a uniform state is the Euler flux along the outward normal
    P2 = oracle.Problem(m, op, {"BC1_name": "noPenetrationBC"})
    F = np.zeros(4)
    gold = {0: [-0.35355, -0.874999, -0.124998, -0.972263], 1: [-0.35355, -0.124998, -0.874999, -0.972263],
            2: [0.35355, 0.124998, 0.874999, 0.972263], 3: [0.35355, 0.874999, 0.124998, 0.972263]}
    for b in range(4):
        n = np.ascontiguousarray(m.nrm_bndry[:, 0, b])
        assert np.allclose(euler_flux(q[:, 0, 0], n), gold[b], atol=1e-5)
    # boundary integral of that (constant) flux: -R^T W f, independent of the face rule
    bf = np.zeros((4, op.face.numnodes, 4), order="F")
    for b in range(4):
        bf[:, :, b] = np.array(gold[b])[:, None]
    res = np.zeros(P.shape, order="F")
    R = sbp.face_matrices(op)
    for b in range(4):
        e, f = int(m.bndryfaces[b]["element"]), int(m.bndryfaces[b]["face"])
        res[:, :, e] -= bf[:, :, b] @ np.diag(op.face.wface) @ R[f]
    el1_res = np.array([[0.35355, 0, -0.35355], [0.124998, -0.750001, -0.874999],
                        [0.874999, 0.750001, -0.124998], [0.972263, 0, -0.972263]])
    el2_res = np.array([[0.35355, -0.35355, 0], [0.874999, -0.124998, 0.750001],
                        [0.124998, -0.874999, -0.750001], [0.972263, -0.972263, 0]])
    assert np.allclose(res[:, :, 1], el1_res, atol=1e-5)
    assert np.allclose(res[:, :, 0], el2_res, atol=1e-5)
    # and the oracle's own interpolate -> BC flux -> integrate pass agrees when
    # the BC state equals the interior state (FreeStream, Ma=0.5, aoa=45deg)
    qfs = P.exact_state("ICFreeStream")
    resb, bflux = P.boundary_integrals(qfs)
    for b in range(4):
        n = np.ascontiguousarray(m.nrm_bndry[:, 0, b])
        assert np.allclose(bflux[:, 0, b], euler_flux(qfs[:, 0, 0], n), atol=1e-13)


# --- test_lowlevel.jl:813-821, test_dg.jl:71-128: uniform flow => zero residual --
@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_uniform_flow_zero_residual(dim, p):
    op = sbp.build_operator(dim, p)
    m = pmesh.structured_mesh(op, 3, shuffle_seed=1)
    opts = {"Flux_name": "RoeFlux", "BC1_name": "FreeStreamBC", "Ma": 0.4, "aoa": 10.0}
    P = oracle.Problem(m, op, opts)
    q = P.exact_state("ICFreeStream")
    qb = P.interpolate_boundary(q)
    assert np.abs(qb - q[:, :1, :1]).max() < 1e-13        # exact interpolation (test_dg.jl:79-84)
    assert np.abs(P.eval_residual(q)).max() < 1e-13
    assert np.abs(P.eval_residual(q, precompute=False)).max() < 1e-13


# --- test/euler/test_flux.jl:255-296: precompute == no-precompute ---------------
@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_precompute_equals_nopre(dim, p):
    op = sbp.build_operator(dim, p)
    m = pmesh.structured_mesh(op, 3, shuffle_seed=2, domain=(0.0, 1.0))
    opts = {"Flux_name": "RoeFlux", "BC1_name": "ExpBC", "SRCname": "SRCExp"}
    P = oracle.Problem(m, op, opts)
    q = P.exact_state("ICExp")
    assert np.linalg.norm(P.volume_integrals(q) - P.volume_integrals(q, False)) < 1e-13
    assert np.linalg.norm(P.face_integrals(q) - P.face_integrals(q, False)) < 1e-13
    assert np.linalg.norm(P.eval_residual(q) - P.eval_residual(q, False)) < 1e-13


# --- test/euler/test_flux.jl:167-234: split form identities ----------------------
def test_split_form_identities():
    op = sbp.build_operator(2, 2, "diage")
    m = pmesh.structured_mesh(op, 3, domain=(0.0, 1.0))
    opts = {"Flux_name": "IRSLFFlux", "Volume_flux_name": "IRFlux", "volume_integral_type": 2,
            "BC1_name": "FreeStreamBC", "Ma": 0.4, "aoa": 5.0}
    P = oracle.Problem(m, op, opts)
    q = P.exact_state("ICExp")
    res = P.volume_integrals(q)
    # S skew, F* symmetric => 1^T (S o F*) 1 = 0 per element and equation
    assert np.abs(res.sum(axis=1)).max() < 1e-13
    # constant field => zero residual for the entropy-stable scheme (test_flux.jl:236-253)
    qf = P.exact_state("ICFreeStream")
    assert np.abs(P.eval_residual(qf)).max() < 1e-12


# --- test/euler/test_ESS.jl:722-740: IR flux + diag-E type-1 face integrals conserve entropy
def test_entropy_conservation_diagE():
    op = sbp.build_operator(2, 2, "diage")
    m = pmesh.structured_mesh(op, 3)
    opts = {"Flux_name": "IRFlux", "Volume_flux_name": "IRFlux", "volume_integral_type": 2,
            "BC1_name": "FreeStreamBC", "Ma": 0.4}
    P = oracle.Problem(m, op, opts)
    q = P.exact_state("ICIsentropicVortex")
    q *= 1 + 1e-2 * np.sin(np.arange(q.size).reshape(q.shape, order="F"))
    w = np.zeros_like(q)
    for e in range(m.numEl):
        for j in range(op.numnodes):
            qq = np.ascontiguousarray(q[:, j, e])
            ww = np.zeros(4)
            L.orc_convert_to_ir(2, G, _ptr(qq), _ptr(ww))
            w[:, j, e] = ww
    res = P.volume_integrals(q) + P.face_integrals(q, precompute=False)
    # w^T R = - sum over boundary faces of psi_n; evaluate the boundary term explicitly
    # psi = rho*u (IR potential flux), so w^T R + sum_bndry wface * psi.n == 0
    total = np.sum(w * res)
    R = sbp.face_matrices(op)
    bterm = 0.0
    for b in range(m.numBoundaryFaces):
        e, f = int(m.bndryfaces[b]["element"]), int(m.bndryfaces[b]["face"])
        qb = q[:, :, e] @ R[f].T
        for i in range(op.face.numnodes):
            bterm += op.face.wface[i] * (qb[1:3, i] @ m.nrm_bndry[:, i, b])
    assert abs(total - bterm) < 1e-12 * max(1.0, abs(bterm))


# --- test/euler/test_rk4.jl:23-43: rk4 integrates a quartic exactly ---------------
def test_rk4_quartic():
    RHS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double)
    POST = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double), C.c_int)

    def f(ctx, q, r, t):
        r[0] = 4 * t ** 3 + 3 * t ** 2 + 2 * t + 1
        return 0
    q = np.array([1.0])
    r = np.zeros(1)
    ns, st = C.c_int64(0), C.c_int(0)
    h, t_max = 0.1, 1.0
    L.orc_rk4.argtypes = [RHS, POST, C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_void_p,
                          C.c_int64, C.c_double, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    t = L.orc_rk4(RHS(f), POST(lambda c, r, n: 1.0), None, h, t_max, 1, _ptr(q), _ptr(r), -1, -1.0, 1,
                  None, 0, C.byref(ns), C.byref(st))
    exact = t_max ** 4 + t_max ** 3 + t_max ** 2 + t_max + 1
    assert abs(q[0] - exact) < 1e-14 * 10
    assert abs(t - t_max) < 1e-14
    assert ns.value == 10


# --- mesh/operator consistency (test/euler/test_curvilinear.jl:3-118) -------------
@pytest.mark.parametrize("dim,p,kind", [(2, 1, "omega"), (2, 2, "omega"), (3, 1, "omega"),
                                        (3, 2, "omega"), (2, 2, "diage")])
def test_freestream_preservation_operator(dim, p, kind):
    """(S_x + 0.5 E_x) 1 = 0 with E_x assembled from nrm_face / nrm_bndry."""
    op = sbp.build_operator(dim, p, kind)
    m = pmesh.structured_mesh(op, 2, shuffle_seed=3)
    R = sbp.face_matrices(op)
    nn = op.numnodes
    Ex = np.zeros((m.numEl, dim, nn, nn))
    for k, I in enumerate(m.interfaces):
        eL, eR, fL, fR, o = (int(I[n]) for n in ("elementL", "elementR", "faceL", "faceR", "orient"))
        pr = op.face.nbrperm[:, o]
        for d in range(dim):
            nL = m.nrm_face[d, :, k]
            Ex[eL, d] += R[fL].T @ np.diag(nL * op.face.wface) @ R[fL]
            RR = R[fR][pr, :]
            Ex[eR, d] += RR.T @ np.diag(-nL * op.face.wface) @ RR
    for b, B in enumerate(m.bndryfaces):
        e, f = int(B["element"]), int(B["face"])
        for d in range(dim):
            Ex[e, d] += R[f].T @ np.diag(m.nrm_bndry[d, :, b] * op.face.wface) @ R[f]
    one = np.ones(nn)
    for e in range(m.numEl):
        for d in range(dim):
            Qx = sum(op.Q[:, :, k] * m.dxidx[k, d, 0, e] for k in range(dim))
            Sx = 0.5 * (Qx - Qx.T)
            assert abs(one @ Ex[e, d] @ one) < 1e-12
            assert np.abs(Ex[e, d] - Ex[e, d].T).max() < 1e-12
            assert np.abs((Sx + 0.5 * Ex[e, d]) @ one).max() < 1e-12


# --- test/euler/test_rk4.jl:45-64: lserk54 integrates a quartic in t exactly ---------------------
def test_lserk54_quartic():
    RHS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double)
    POST = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_void_p, C.c_int)

    def rhs(ctx, q, res, t):
        C.cast(res, C.POINTER(C.c_double))[0] = 4 * t ** 3 + 3 * t ** 2 + 2 * t + 1
        return 0

    def post(ctx, res, calc_norm):
        return 1.0
    q = np.array([1.0])
    res = np.zeros(1)
    ns, st = C.c_int64(0), C.c_int(0)
    L.orc_lserk54.argtypes = [RHS, POST, C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_void_p,
                              C.c_int64, C.c_double, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    t = L.orc_lserk54(RHS(rhs), POST(post), None, 0.01, 1.0, 1, _ptr(q), _ptr(res), -1, -1.0, 1, None, 0,
                      C.byref(ns), C.byref(st))
    exact = t ** 4 + t ** 3 + t ** 2 + t + 1
    assert abs(t - 1.0) < 1e-12 and abs(q[0] - exact) < 1e-12


# --- committed fixtures: the oracle must keep reproducing them (tests/golden/make_fixtures.py) ------------
@pytest.mark.parametrize("case", ["c1_2d_p1_roe", "c3_3d_p2_roe_src", "c2_2d_p2_es", "2d_p2_roe", "3d_p1_roe_src"])
def test_oracle_reproduces_committed_fixtures(case):
    import os
    import sys
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold)
    import make_fixtures
    op, mesh, opts, orc, q0, h = make_fixtures.build(case)
    fx = np.load(os.path.join(gold, case + ".npz"))
    assert np.array_equal(q0, fx["q0"])
    res = orc.eval_residual(q0)
    assert np.linalg.norm(res - fx["res"]) <= 1e-14 * np.linalg.norm(fx["res"])
    t, q5, norms = orc.rk4(q0, h, 5 * h)
    assert t == float(fx["t_rk4"]) and np.linalg.norm(q5 - fx["q_rk4"]) <= 1e-14 * np.linalg.norm(q5)


def test_reference_known_answers_file():
    import json
    import os
    ka = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_known_answers.json")))
    a = ka["calcPressure_2d"]
    assert abs(pressure(a["q"]) - a["p"]) < a["tol"]
    a = ka["calcEulerFlux_2d"]
    assert np.allclose(euler_flux(a["q"], a["dir"]), a["F"], atol=a["tol"], rtol=0)
    a = ka["calcEulerFlux_3d"]
    for d, F in a["F"].items():
        assert np.allclose(euler_flux(a["q"], [float(c) for c in d]), F, atol=a["tol"], rtol=0)
    a = ka["calcIsentropicVortex"]
    sol = np.zeros(4)
    L.orc_isentropic_vortex(2, G, 287.058, _ptr(np.array(a["coords"])), _ptr(sol))
    assert np.allclose(sol, a["q"], atol=a["tol"])


# --- test/euler/convergence/p1/conservative_dg/runtests.jl:1-42: the reference's integration-level golden --------
def _steady_vortex_error(runner, name, h):
    """Pseudo-time RK4 to the steady state of the isentropic vortex on the reference's own mesh fixture, then the
    M-weighted L2 error against the exact solution (startup_func.jl:188-201: calcNorm(|q - q_IC|))."""
    import os
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    op = sbp.build_operator(2, 1)
    mesh = pmesh.simplex_mesh(op, fx["vertex_coords"], fx["triangles"])
    opts = {"Flux_name": "RoeFlux", "BC1_name": "isentropicVortexBC", "use_itermax": False}
    P = oracle.Problem(mesh, op, opts)
    q0 = P.exact_state("ICIsentropicVortex")
    q = runner(mesh, op, opts, P, q0, h)
    M = 1.0 / P.mass_matrix_inverse()
    return float(np.sqrt(np.sum(M * (q - q0) ** 2)))


def _oracle_runner(mesh, op, opts, P, q0, h):
    q = q0
    for _ in range(20):
        _, q, norms = P.rk4(q, h, 2000 * h)
        if norms[-1] < 1e-12:
            return q
    raise AssertionError("pseudo-time iteration did not converge")


def test_reference_convergence_golden():
    """The reference asserts err[1] = 0.01200 (x/ 1.25) on squarevortex_small and a convergence slope of 2.00 +- 0.1
    between squarevortex_small and squarevortex_large (p=1 SBP-Omega DG, Roe flux, isentropicVortexBC, steady state
    reached there by Newton).  The oracle, marched to the same steady state, must land inside those bands."""
    e1 = _steady_vortex_error(_oracle_runner, "squarevortex_small", 0.02)
    e2 = _steady_vortex_error(_oracle_runner, "squarevortex_large", 0.01)
    assert 0.01200 / 1.25 < e1 < 0.01200 * 1.25
    assert abs(e1 - 0.01200) < 2e-5           # measured 0.0120031: the golden to its printed precision
    slope = np.log(e1 / e2) / np.log(2.0)
    assert 1.9 < slope < 2.1


# --- boundary functors beyond the four of the named configurations (bc.jl:767-793, 1454-1537, 1702-1722, 2140-2152) ------
def _bc_flux(P, bc_id, q, nrm, coords=None):
    dim = len(nrm)
    flux = np.zeros(dim + 2)
    x = np.zeros(dim) if coords is None else np.ascontiguousarray(coords, dtype=np.float64)
    oracle.lib().orc_bc_flux(P.ref(), int(bc_id), _ptr(np.ascontiguousarray(q)), _ptr(x),
                             _ptr(np.ascontiguousarray(nrm, dtype=np.float64)), _ptr(flux))
    return flux


def _random_states(dim, n, seed, vscale=0.6):
    rng = np.random.RandomState(seed)
    rho = 0.5 + rng.rand(n)
    vel = rng.standard_normal((n, dim)) * vscale
    p = 0.5 + rng.rand(n)
    E = p / 0.4 + 0.5 * rho * (vel ** 2).sum(axis=1)
    return np.column_stack([rho, rho[:, None] * vel, E])


@pytest.mark.parametrize("dim", [2, 3])
def test_no_penetration_es_bc_is_entropy_stable(dim):
    """test_ESSBC (test/euler/test_ESS.jl:748-779): psi_n - w^T f <= 1e-12 with the IR entropy variables, psi = momentum,
    on subsonic states like the vortex the reference evaluates it on (the Lax-Friedrichs speed is taken at the average of
    q and its reflection, whose normal velocity is zero, so the bound needs a subsonic wall-normal velocity)."""
    op = sbp.build_operator(dim, 1)
    P = oracle.Problem(pmesh.structured_mesh(op, 1), op, {"BC1_name": "noPenetrationESBC"})
    rng = np.random.RandomState(11)
    L = oracle.lib()
    for q in _random_states(dim, 200, 5, vscale=0.25):
        n = rng.standard_normal(dim)
        n /= np.linalg.norm(n)
        f = _bc_flux(P, oracle.BC_IDS["noPenetrationESBC"], q, n)
        w = np.zeros(dim + 2)
        L.orc_convert_to_ir(dim, 1.4, _ptr(np.ascontiguousarray(q)), _ptr(w))
        assert (q[1:1 + dim] * n).sum() - w @ f <= 1e-12
        # for a state that already satisfies the wall condition the flux is the pressure force only (qg == q: the
        # Lax-Friedrichs dissipation vanishes)
        qt = q.copy()
        qt[1:1 + dim] -= (qt[1:1 + dim] @ n) * n
        ft = _bc_flux(P, oracle.BC_IDS["noPenetrationESBC"], qt, n)
        pt = 0.4 * (qt[dim + 1] - 0.5 * (qt[1:1 + dim] ** 2).sum() / qt[0])
        assert np.allclose(ft, np.concatenate([[0.0], pt * n, [0.0]]), rtol=0, atol=1e-13)


@pytest.mark.parametrize("dim", [2, 3])
def test_constant_state_and_zero_flux_bcs(dim):
    op = sbp.build_operator(dim, 1)
    P = oracle.Problem(pmesh.structured_mesh(op, 1), op, {"BC1_name": "Rho1E2U3BC"})
    L = oracle.lib()
    rng = np.random.RandomState(2)
    for q in _random_states(dim, 20, 9):
        n = rng.standard_normal(dim)
        for name, qg in (("Rho1E2U3BC", np.array([1.0] + [0.35355] * dim + [2.0])), ("allOnesBC", np.ones(dim + 2))):
            ref = np.zeros(dim + 2)
            L.orc_roe_solver(dim, 1.4, _ptr(np.ascontiguousarray(q)), _ptr(qg), _ptr(n), _ptr(ref))
            assert np.array_equal(_bc_flux(P, oracle.BC_IDS[name], q, n), ref)
        assert not _bc_flux(P, oracle.BC_IDS["ZeroFluxBC"], q, n).any()


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_rho1e2u3_uniform_state_zero_residual(dim, p):
    """test/euler/test_dg.jl:70-84: the uniform Rho1E2U3 state with Rho1E2U3BC on every boundary gives a zero residual."""
    op = sbp.build_operator(dim, p)
    mesh = pmesh.structured_mesh(op, 3, shuffle_seed=4)
    P = oracle.Problem(mesh, op, {"Flux_name": "RoeFlux", "BC1_name": "Rho1E2U3BC"})
    q = np.zeros(P.shape, order="F")
    q[0], q[dim + 1] = 1.0, 2.0
    q[1:1 + dim] = 0.35355
    assert np.abs(P.eval_residual(q)).max() < 1e-13


def test_calc_norm_matches_reference_definition():
    """test/euler/Utils.jl:111-135 (test_utils_misc): calcNorm(eqn, v) = sqrt(sum v M v) (Utils.jl:427-449), the strong
    form with Minv in place of M."""
    import ctypes as C
    rng = np.random.RandomState(11)
    M = rng.rand(10) + 0.1
    data = rng.rand(10)
    L.orc_calc_norm.restype = C.c_double
    L.orc_calc_norm.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
    assert np.isclose(L.orc_calc_norm(10, _ptr(M), _ptr(data)), np.sqrt(np.sum(data * M * data)), rtol=1e-15)
    Minv = 1.0 / M
    assert np.isclose(L.orc_calc_norm(10, _ptr(Minv), _ptr(data)), np.sqrt(np.sum(data * Minv * data)), rtol=1e-15)


def test_complex_step_oracle_against_central_differences():
    """oracle/euler_oracle_cs.c (the reference's J*v: complex step through the Complex128 residual) vs central differences
    of the real oracle, and its linearity in v -- the checker of the device's dual-number product."""
    import pdesolver_jl_b200 as pd
    from common import CASES, perturbed, rel_l2
    for case, n, bcs in [("c1_2d_p1_roe", 4, ("isentropicVortexBC", "noPenetrationBC")),
                         ("c3_3d_p2_roe_src", 2, ("ExpBC", "noPenetrationESBC")), ("2d_p2_roe", 3, ("FreeStreamBC", "allOnesBC"))]:
        dim, p, ic, opts = CASES[case]
        op = pd.build_operator(dim, p)
        sides = [0, 1, 0, 1] if dim == 2 else [0, 1, 0, 1, 0, 1]
        opts = dict(opts, BC1_name=bcs[0], BC2_name=bcs[1], Ma=0.5, aoa=5.0)
        mesh = pd.structured_mesh(op, n, shuffle_seed=6, bc_sides=sides)
        orc = oracle.Problem(mesh, op, opts)
        q0 = perturbed(orc.exact_state(ic))
        rng = np.random.RandomState(0)
        v = np.asfortranarray(rng.standard_normal(q0.shape) * np.abs(q0))
        w = np.asfortranarray(rng.standard_normal(q0.shape) * np.abs(q0))
        Jv, Jw = orc.eval_jvp_complex_step(q0, v), orc.eval_jvp_complex_step(q0, w)
        eps = 1e-6
        fd = (orc.eval_residual(np.asfortranarray(q0 + eps * v)) - orc.eval_residual(np.asfortranarray(q0 - eps * v))) / (2 * eps)
        assert rel_l2(Jv, fd) < 1e-8
        assert rel_l2(orc.eval_jvp_complex_step(q0, 2.0 * v - 3.0 * w), 2.0 * Jv - 3.0 * Jw) < 1e-13
