"""Shared helpers for the parity tests: deterministic states."""
import numpy as np


def perturbed(q, amp=1e-3):
    """q * (1 + amp*sin(k + 7j + 13e)): deterministic, no RNG (the reference's perturb_ic uses an
    unseeded rand(), solver/euler/startup_func.jl:227-233), makes the face jumps non-zero."""
    nd, nn, nE = q.shape
    k = np.arange(nd)[:, None, None]
    j = np.arange(nn)[None, :, None]
    e = np.arange(nE)[None, None, :]
    return np.asfortranarray(q * (1.0 + amp * np.sin(k + 7.0 * j + 13.0 * e)))


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


ES_OPTS = {"Flux_name": "IRSLFFlux", "Volume_flux_name": "IRFlux", "volume_integral_type": 2,
           "BC1_name": "isentropicVortexBC"}
KIND = {"c2_2d_p2_es": "diage", "2d_p2_es_ir": "diage", "2d_p2_es_roe": "diage",
        # operators whose node counts have no tuned kernel instantiation (size-generic kernels)
        "2d_p2_alt_roe": "omega_alt", "3d_p2_alt_roe_src": "omega_alt", "3d_p1_es": "diage"}

CASES = {
    # name: (dim, degree, IC, opts)
    "c2_2d_p2_es": (2, 2, "ICIsentropicVortex", dict(ES_OPTS)),
    "2d_p2_es_ir": (2, 2, "ICIsentropicVortex", dict(ES_OPTS, Flux_name="IRFlux")),
    "2d_p2_es_roe": (2, 2, "ICIsentropicVortex", dict(ES_OPTS, Flux_name="RoeFlux")),
    "c1_2d_p1_roe": (2, 1, "ICIsentropicVortex",
                     {"Flux_name": "RoeFlux", "BC1_name": "isentropicVortexBC"}),
    "2d_p2_roe": (2, 2, "ICIsentropicVortex",
                  {"Flux_name": "RoeFlux", "BC1_name": "isentropicVortexBC"}),
    "3d_p1_roe_src": (3, 1, "ICExp",
                      {"Flux_name": "RoeFlux", "BC1_name": "ExpBC", "SRCname": "SRCExp"}),
    "c3_3d_p2_roe_src": (3, 2, "ICExp",
                         {"Flux_name": "RoeFlux", "BC1_name": "ExpBC", "SRCname": "SRCExp"}),
    "2d_p2_alt_roe": (2, 2, "ICIsentropicVortex", {"Flux_name": "RoeFlux", "BC1_name": "isentropicVortexBC"}),
    "3d_p2_alt_roe_src": (3, 2, "ICExp", {"Flux_name": "RoeFlux", "BC1_name": "ExpBC", "SRCname": "SRCExp"}),
    "3d_p1_es": (3, 1, "ICExp", dict(ES_OPTS, BC1_name="ExpBC")),
}
