"""ctypes binding of the C ABI in ``include/pdes_euler_b200.h`` (libpdes_euler_b200.so).

This is the same binding a Julia host makes with ``ccall``;
nothing here computes: every entry point forwards to the CUDA library, and
loading fails loudly when the library has not been built -- there is no CPU
fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PDES_LIB") or os.path.join(_HERE, "libpdes_euler_b200.so")   # PDES_LIB: kernel A/B builds

PDES_OK = 0
PDES_ERR_NEG_DENSITY = 1
PDES_ERR_NEG_PRESSURE = 2
PDES_ERR_USAGE = -1
PDES_ERR_CUDA = -2
PDES_ERR_UNSUPPORTED = -3
PDES_ERR_COMM = -4

FLUX_IDS = {"RoeFlux": 1, "IRFlux": 2, "IRSLFFlux": 3, "StandardFlux": 4}
BC_IDS = {"isentropicVortexBC": 1, "ExpBC": 2, "FreeStreamBC": 3, "noPenetrationBC": 4, "Rho1E2U3BC": 5,
          "allOnesBC": 6, "ZeroFluxBC": 7, "noPenetrationESBC": 8}
SRC_IDS = {"SRC0": 0, "SRCExp": 1}
FEI_IDS = {"ECFaceIntegral": 1, "ELFPenaltyFaceIntegral": 2, "ESLFFaceIntegral": 3, "ELW2PenaltyFaceIntegral": 4,
           "ESLW2FaceIntegral": 5}

EXPORTS = [
    "pdes_create", "pdes_destroy", "pdes_last_error", "pdes_last_error_location",
    "pdes_set_operator", "pdes_set_mesh", "pdes_set_peer", "pdes_get_unique_id", "pdes_set_comm",
    "pdes_pack_send", "pdes_inject_recv", "pdes_set_q", "pdes_get_q", "pdes_get_res", "pdes_q_dev",
    "pdes_res_dev", "pdes_eval_residual", "pdes_eval_residual_async", "pdes_sync", "pdes_rk4",
    "pdes_rk4_steps_async", "pdes_get_minv", "pdes_get_timings", "pdes_kernel_launch_count",
    "pdes_set_q_dev", "pdes_stream", "pdes_pin_host", "pdes_unpin_host", "pdes_eval_jvp", "pdes_lserk54",
    "pdes_newton_krylov", "pdes_gmres", "pdes_diagnostics",
    "pdes_set_peer_elements", "pdes_pack_send_elements", "pdes_inject_recv_elements", "pdes_set_krylov_pc",
    "pdes_eval_residual_host",
]


class PdesConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim", "nn", "nfn", "ss", "norient", "sparse_face",
                                         "index_base", "device")] + \
               [(n, C.c_int64) for n in ("nE", "nF", "nB")] + \
               [(n, C.c_int32) for n in ("numBC", "npeers", "volume_integral_type",
                                         "face_integral_type", "flux_id", "volume_flux_id", "src_id",
                                         "check_density", "check_pressure", "face_element_id")] + \
               [(n, C.c_double) for n in ("gamma", "R", "Ma", "aoa", "rho_free", "E_free")]


class PdesTimings(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("t_send", "t_dataprep", "t_volume", "t_face", "t_sharedface",
                                          "t_source", "t_func", "t_timemarch", "t_wait", "t_allreduce")] + \
               [("n_residual_evals", C.c_int64), ("n_kernel_launches", C.c_int64)]


class PdesNewtonOpts(C.Structure):
    _fields_ = [("itermax", C.c_int64)] + \
               [(n, C.c_double) for n in ("res_abstol", "res_reltol", "step_tol", "step_fac", "krylov_reltol",
                                          "krylov_abstol", "krylov_dtol")] + \
               [("krylov_itermax", C.c_int64), ("krylov_restart", C.c_int32)]


class PdesNewtonResult(C.Structure):
    _fields_ = [("converged", C.c_int32), ("krylov_reason", C.c_int32)] + \
               [(n, C.c_int64) for n in ("newton_iters", "krylov_iters", "residual_evals")] + \
               [(n, C.c_double) for n in ("res_norm", "res_norm_rel", "step_norm")]


_lib = None


def lib():
    """The loaded CUDA library.  Raises if it is missing: the product path never
    falls back to a CPU implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        p, i32, i64, d = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.pdes_create.argtypes = [C.POINTER(PdesConfig), C.POINTER(p)]
        L.pdes_destroy.argtypes = [p]
        L.pdes_destroy.restype = None
        L.pdes_last_error.argtypes = [p]
        L.pdes_last_error.restype = C.c_char_p
        L.pdes_last_error_location.argtypes = [p, C.POINTER(i64), C.POINTER(i64)]
        L.pdes_set_operator.argtypes = [p] * 7
        L.pdes_set_mesh.argtypes = [p] * 11
        L.pdes_set_peer.argtypes = [p, i32, i32, i64, p, p, p]
        L.pdes_get_unique_id.argtypes = [p]
        L.pdes_set_comm.argtypes = [p, p, i32, i32]
        L.pdes_set_krylov_pc.argtypes = [p, i32]
        L.pdes_eval_residual_host.argtypes = [p, p, p, C.c_double]
        L.pdes_set_peer_elements.argtypes = [p, i32, i64, p, i64, i64]
        L.pdes_pack_send_elements.argtypes = [p, i32, p]
        L.pdes_inject_recv_elements.argtypes = [p, i32, p]
        L.pdes_pack_send.argtypes = [p, i32, p]
        L.pdes_inject_recv.argtypes = [p, i32, p]
        for n in ("pdes_set_q", "pdes_get_q", "pdes_get_res", "pdes_get_minv", "pdes_set_q_dev", "pdes_diagnostics"):
            getattr(L, n).argtypes = [p, p]
        L.pdes_q_dev.argtypes = [p]
        L.pdes_q_dev.restype = p
        L.pdes_res_dev.argtypes = [p]
        L.pdes_res_dev.restype = p
        L.pdes_eval_jvp.argtypes = [p, p, p]
        L.pdes_pin_host.argtypes = [p, i64]
        L.pdes_unpin_host.argtypes = [p]
        L.pdes_stream.argtypes = [p]
        L.pdes_stream.restype = p
        L.pdes_eval_residual.argtypes = [p, d]
        L.pdes_eval_residual_async.argtypes = [p, d]
        L.pdes_sync.argtypes = [p]
        L.pdes_rk4.argtypes = [p, d, d, i64, d, i32, C.POINTER(d), p, i64, C.POINTER(i64)]
        L.pdes_lserk54.argtypes = [p, d, d, i64, d, i32, C.POINTER(d), p, i64, C.POINTER(i64)]
        L.pdes_rk4_steps_async.argtypes = [p, d, i64]
        L.pdes_get_timings.argtypes = [p, C.POINTER(PdesTimings)]
        L.pdes_newton_krylov.argtypes = [p, C.POINTER(PdesNewtonOpts), p, p, C.POINTER(PdesNewtonResult)]
        L.pdes_gmres.argtypes = [p, p, p, d, d, d, i64, i32, C.POINTER(i64), C.POINTER(d), C.POINTER(i32)]
        L.pdes_kernel_launch_count.argtypes = [p]
        L.pdes_kernel_launch_count.restype = i64
        _lib = L
    return _lib
