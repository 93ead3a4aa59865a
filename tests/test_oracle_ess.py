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
