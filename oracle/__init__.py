"""ctypes loader for the CPU oracle (oracle/euler_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / reference arm.  Nothing under pdesolver.jl_b200/
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FLUX_IDS = {"RoeFlux": 1, "IRFlux": 2, "IRSLFFlux": 3, "StandardFlux": 4}
BC_IDS = {"isentropicVortexBC": 1, "ExpBC": 2, "FreeStreamBC": 3, "noPenetrationBC": 4, "Rho1E2U3BC": 5,
          "allOnesBC": 6, "ZeroFluxBC": 7, "noPenetrationESBC": 8}
SRC_IDS = {"SRC0": 0, "SRCExp": 1}
FEI_IDS = {"ECFaceIntegral": 1, "ELFPenaltyFaceIntegral": 2, "ESLFFaceIntegral": 3, "ELW2PenaltyFaceIntegral": 4,
           "ESLW2FaceIntegral": 5}


def build(force=False):
    """Compile liborc.so / liborc_omp.so / liborc_cs.so with the committed Makefile."""
    if force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ("liborc.so", "liborc_omp.so", "liborc_cs.so")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)


class OrcProblem(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("dim", "nd", "nn", "nfn", "ss", "nfaces", "norient", "sparse_face")] + \
               [(n, C.c_int64) for n in ("nE", "nF", "nB")] + \
               [(n, C.c_int32) for n in ("numBC", "flux_id", "volume_flux_id",
                                         "volume_integral_type", "src_id", "check_density",
                                         "check_pressure", "face_element_id")] + \
               [(n, C.c_double) for n in ("gamma", "R", "Ma", "aoa", "rho_free", "E_free")] + \
               [(n, C.c_void_p) for n in ("Q", "w", "interp", "wface", "perm", "nbrperm", "dxidx",
                                          "jac", "coords", "nrm_face", "nrm_bndry", "coords_bndry",
                                          "ifaces", "bfaces", "bndry_offsets", "bc_ids")]


class OrcPeer(C.Structure):
    _fields_ = [("nfaces", C.c_int64), ("bndries_local", C.c_void_p), ("interfaces", C.c_void_p),
                ("nrm_sharedface", C.c_void_p), ("q_send", C.c_void_p), ("q_recv", C.c_void_p),
                ("q_recv_el", C.c_void_p), ("el_offset", C.c_int64)]


_libs = {}


def lib_cs():
    """liborc_cs.so: the complex-step J*v restatement (euler_oracle_cs.c)."""
    if "cs" not in _libs:
        build()
        L = C.CDLL(os.path.join(_HERE, "liborc_cs.so"))
        L.orc_eval_jvp_complex_step.restype = C.c_int
        L.orc_eval_jvp_complex_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        _libs["cs"] = L
    return _libs["cs"]


def lib(omp=False):
    name = "liborc_omp.so" if omp else "liborc.so"
    if name not in _libs:
        build()
        L = C.CDLL(os.path.join(_HERE, name))
        d, i, p = C.c_double, C.c_int, C.c_void_p
        L.orc_calc_pressure.restype = d
        L.orc_calc_pressure.argtypes = [i, d, p]
        L.orc_euler_flux.argtypes = [i, d, p, p, p]
        L.orc_roe_solver.argtypes = [i, d, p, p, p, p]
        L.orc_logavg.restype = d
        L.orc_logavg.argtypes = [d, d]
        L.orc_ir_flux.argtypes = [i, d, p, p, p, i, p]
        L.orc_convert_to_ir.argtypes = [i, d, p, p]
        L.orc_ira0.argtypes = [i, d, p, p]
        L.orc_evals_x.argtypes = [i, d, p, p]
        L.orc_evecs_x.argtypes = [i, d, p, p]
        L.orc_escaling_x.argtypes = [i, d, p, p]
        L.orc_lw2_entropy_kernel.argtypes = [i, d, p, p, p, p]
        L.orc_projection_matrix.argtypes = [i, p, p]
        L.orc_lambda_max.restype = d
        L.orc_lambda_max.argtypes = [i, d, p, p]
        L.orc_irslf_flux.argtypes = [i, d, p, p, p, p]
        L.orc_isentropic_vortex.argtypes = [i, d, d, p, p]
        L.orc_calc_exp.argtypes = [i, d, p, p]
        L.orc_free_stream.argtypes = [i, d, d, d, d, p]
        L.orc_src_exp.argtypes = [i, d, p, p]
        L.orc_bc_flux.argtypes = [p, i, p, p, p, p]
        L.orc_eval_residual.restype = i
        L.orc_eval_residual.argtypes = [p, p, p, d, i, i, p, p]
        L.orc_volume_integrals.argtypes = [p, p, p, i]
        L.orc_euler_flux_parametric.argtypes = [p, p, p]
        L.orc_face_integrals.argtypes = [p, p, p, i]
        L.orc_boundary_integrals.argtypes = [p, p, p, p]
        L.orc_interpolate_boundary.argtypes = [p, p, p]
        L.orc_get_send_data_face.argtypes = [p, p, p]
        L.orc_mass_matrix_inverse.argtypes = [p, p]
        L.orc_calc_norm.restype = d
        L.orc_calc_norm.argtypes = [C.c_int64, p, p]
        L.orc_rk4.restype = d
        L.orc_fill_exact.argtypes = [p, i, p, C.c_int64, p]
        L.orc_rk4_euler.restype = d
        L.orc_rk4_euler.argtypes = [p, p, d, d, C.c_int64, d, i, i, p, C.c_int64, p, p]
        L.orc_lserk54.restype = d
        L.orc_lserk54_euler.restype = d
        L.orc_lserk54_euler.argtypes = [p, p, d, d, C.c_int64, d, i, i, p, C.c_int64, p, p]
        _libs[name] = L
    return _libs[name]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _vec(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())


class Problem:
    """Owns the numpy arrays an OrcProblem points at."""

    def __init__(self, mesh, sbp, opts):
        f = sbp.face
        self.mesh, self.sbp, self.opts = mesh, sbp, opts
        self.nd = mesh.numDofPerNode
        keep = self._keep = {}

        def F(name, arr, dtype=np.float64):
            a = np.asfortranarray(np.asarray(arr, dtype=dtype))
            keep[name] = a
            return _ptr(a)
        P = OrcProblem()
        P.dim, P.nd, P.nn, P.nfn = mesh.dim, self.nd, sbp.numnodes, f.numnodes
        P.ss, P.nfaces, P.norient = f.stencilsize, mesh.dim + 1, f.nbrperm.shape[1]
        P.sparse_face = int(f.sparse)
        P.nE, P.nF, P.nB = mesh.numEl, mesh.numInterfaces, mesh.numBoundaryFaces
        P.numBC = mesh.numBC
        P.flux_id = FLUX_IDS[opts.get("Flux_name", "RoeFlux")]
        P.volume_flux_id = FLUX_IDS[opts.get("Volume_flux_name", "StandardFlux")]
        P.volume_integral_type = opts.get("volume_integral_type", 1)
        P.src_id = SRC_IDS[opts.get("SRCname", "SRC0")]
        P.check_density = int(opts.get("check_density", True))
        P.check_pressure = int(opts.get("check_pressure", True))
        P.face_element_id = (FEI_IDS[opts.get("FaceElementIntegral_name", "ESLFFaceIntegral")]
                             if int(opts.get("face_integral_type", 1)) == 2 else 0)
        g = opts.get("gamma", 1.4)
        P.gamma, P.R = g, opts.get("R", 287.058)
        Ma = opts.get("Ma", -1.0)
        P.Ma, P.aoa = Ma, opts.get("aoa", 0.0) * np.pi / 180
        p_free = opts.get("p_free", 1 / g)
        P.rho_free, P.E_free = 1.0, p_free / (g - 1) + 0.5 * Ma * Ma
        P.Q, P.w, P.interp, P.wface = F("Q", sbp.Q), F("w", sbp.w), F("interp", f.interp), F("wface", f.wface)
        P.perm, P.nbrperm = F("perm", f.perm, np.int64), F("nbrperm", f.nbrperm, np.int64)
        P.dxidx, P.jac, P.coords = F("dxidx", mesh.dxidx), F("jac", mesh.jac), F("coords", mesh.coords)
        P.nrm_face, P.nrm_bndry = F("nrm_face", mesh.nrm_face), F("nrm_bndry", mesh.nrm_bndry)
        P.coords_bndry = F("coords_bndry", mesh.coords_bndry)
        keep["ifaces"] = np.ascontiguousarray(mesh.interfaces)
        keep["bfaces"] = np.ascontiguousarray(mesh.bndryfaces)
        P.ifaces, P.bfaces = _ptr(keep["ifaces"]), _ptr(keep["bfaces"])
        P.bndry_offsets = F("bo", mesh.bndry_offsets, np.int64)
        bc = [BC_IDS[opts.get(f"BC{i + 1}_name", "isentropicVortexBC")] for i in range(mesh.numBC)]
        P.bc_ids = F("bc", np.array(bc), np.int32)
        self.P = P
        # peers
        self.peers = (OrcPeer * max(mesh.npeers, 1))()
        self.q_send, self.q_recv = [], []
        for p in range(mesh.npeers):
            n = len(mesh.bndries_local[p])
            keep[f"bl{p}"] = np.ascontiguousarray(mesh.bndries_local[p])
            keep[f"si{p}"] = np.ascontiguousarray(mesh.shared_interfaces[p])
            keep[f"ns{p}"] = np.asfortranarray(mesh.nrm_sharedface[p])
            self.q_send.append(np.zeros((self.nd, f.numnodes, n), order="F"))
            self.q_recv.append(np.zeros((self.nd, f.numnodes, n), order="F"))
            pr = self.peers[p]
            pr.nfaces = n
            pr.bndries_local, pr.interfaces = _ptr(keep[f"bl{p}"]), _ptr(keep[f"si{p}"])
            pr.nrm_sharedface = _ptr(keep[f"ns{p}"])
            pr.q_send, pr.q_recv = _ptr(self.q_send[p]), _ptr(self.q_recv[p])
            pr.q_recv_el, pr.el_offset = None, 0
        self.q_recv_el = [None] * mesh.npeers

    def set_recv_elements(self, p, q_el):
        """parallel_data = element: whole remote elements [nd, nn, n_remote] of peer index p, in the order of the remote
        element numbers (first one = mesh.shared_element_offsets[p]); the receive side of getSendDataElement."""
        self.q_recv_el[p] = np.asfortranarray(q_el, dtype=np.float64)
        self.peers[p].q_recv_el = _ptr(self.q_recv_el[p])
        self.peers[p].el_offset = int(self.mesh.shared_element_offsets[p])

    @staticmethod
    def get_send_data_element(q, local_elements):
        """getSendDataElement (Utils/parallel.jl:276-293): send_buff[:, :, j] = q[:, :, local_element_lists[idx][j]]"""
        return np.asfortranarray(q[:, :, np.asarray(local_elements, dtype=np.int64)])

    @property
    def shape(self):
        return (self.nd, self.sbp.numnodes, self.mesh.numEl)

    def ref(self):
        return C.byref(self.P)

    # -- passes ---------------------------------------------------------------
    def start_exchange(self, q, omp=False):
        for p in range(self.mesh.npeers):
            lib(omp).orc_get_send_data_face(self.ref(), _ptr(q), C.byref(self.peers[p]))

    def eval_residual(self, q, t=0.0, precompute=True, omp=False):
        q = np.asfortranarray(q, dtype=np.float64)
        res = np.zeros(self.shape, order="F")
        err = np.zeros(2, dtype=np.int64)
        st = lib(omp).orc_eval_residual(self.ref(), _ptr(q), _ptr(res), float(t), int(precompute),
                                        self.mesh.npeers, self.peers, _ptr(err))
        if st == 1:
            raise FloatingPointError(f"Negative density detected at element {err[0]} node {err[1]}")
        if st == 2:
            raise FloatingPointError(f"Negative pressure detected at element {err[0]} node {err[1]}")
        return res

    def eval_jvp_complex_step(self, q, v, eps=1e-20):
        """The reference's own J*v: imag(R(q + i eps v))/eps with the residual in complex arithmetic (liborc_cs.so,
        euler_oracle_cs.c; newton_setup.jl:632-662).  Dense-face Roe configurations."""
        q = np.asfortranarray(q, dtype=np.float64)
        v = np.asfortranarray(v, dtype=np.float64)
        out = np.zeros(self.shape, order="F")
        L = lib_cs()
        if L.orc_eval_jvp_complex_step(self.ref(), _ptr(q), _ptr(v), float(eps), _ptr(out)) != 0:
            raise NotImplementedError("complex-step oracle: dense-face Roe path only")
        return out

    def volume_integrals(self, q, precompute=True):
        res = np.zeros(self.shape, order="F")
        lib().orc_volume_integrals(self.ref(), _ptr(np.asfortranarray(q)), _ptr(res), int(precompute))
        return res

    def face_integrals(self, q, precompute=True):
        res = np.zeros(self.shape, order="F")
        lib().orc_face_integrals(self.ref(), _ptr(np.asfortranarray(q)), _ptr(res), int(precompute))
        return res

    def boundary_integrals(self, q):
        res = np.zeros(self.shape, order="F")
        bf = np.zeros((self.nd, self.sbp.face.numnodes, max(self.mesh.numBoundaryFaces, 1)), order="F")
        lib().orc_boundary_integrals(self.ref(), _ptr(np.asfortranarray(q)), _ptr(res), _ptr(bf))
        return res, bf

    def euler_flux_parametric(self, q):
        fp = np.zeros(self.shape + (self.mesh.dim,), order="F")
        lib().orc_euler_flux_parametric(self.ref(), _ptr(np.asfortranarray(q)), _ptr(fp))
        return fp

    def interpolate_boundary(self, q):
        qb = np.zeros((self.nd, self.sbp.face.numnodes, self.mesh.numBoundaryFaces), order="F")
        lib().orc_interpolate_boundary(self.ref(), _ptr(np.asfortranarray(q)), _ptr(qb))
        return qb

    def mass_matrix_inverse(self):
        Minv = np.zeros(self.shape, order="F")
        lib().orc_mass_matrix_inverse(self.ref(), _ptr(Minv))
        return Minv

    def lserk54(self, q, h, t_max, itermax=-1, res_tol=-1.0, real_time=False, precompute=True, omp=False):
        """lserk54 (NonlinearSolvers/lserk.jl).  Returns (t, q_final, norms)."""
        return self.rk4(q, h, t_max, itermax, res_tol, real_time, precompute, omp, _fn="orc_lserk54_euler")

    def rk4(self, q, h, t_max, itermax=-1, res_tol=-1.0, real_time=False, precompute=True, omp=False,
            _fn="orc_rk4_euler"):
        """Returns (t, q_final, norms).  q is not modified."""
        qv = np.asfortranarray(q, dtype=np.float64).copy(order="F")
        cap = int(round(t_max / h)) + 2
        norms = np.zeros(cap)
        ns = C.c_int64(0)
        st = C.c_int(0)
        t = getattr(lib(omp), _fn)(self.ref(), _ptr(qv), float(h), float(t_max), int(itermax),
                                   float(res_tol), int(real_time), int(precompute), _ptr(norms), cap,
                                   C.byref(ns), C.byref(st))
        if st.value:
            raise FloatingPointError("negative density/pressure in oracle rk4")
        return t, qv, norms[:ns.value]

    # -- states ---------------------------------------------------------------
    def exact_state(self, name, coords=None):
        """ICIsentropicVortex / ICExp / ICFreeStream evaluated at the nodes
        (ic.jl macro-generated ICs call calc<Name> per node)."""
        coords = self.mesh.coords if coords is None else coords
        dim = self.mesh.dim
        cs = np.asfortranarray(np.asarray(coords, dtype=np.float64)).reshape(dim, -1, order="F")
        cs = np.asfortranarray(cs)
        out = np.zeros((self.nd, cs.shape[1]), order="F")
        kind = {"ICIsentropicVortex": 1, "ICExp": 2, "ICFreeStream": 3}[name]
        lib().orc_fill_exact(self.ref(), kind, _ptr(cs), cs.shape[1], _ptr(out))
        return np.asfortranarray(out.reshape((self.nd,) + tuple(np.shape(coords)[1:]), order="F"))


def exchange(problems, qs, omp=False):
    """In-process stand-in for the MPI Isend/Irecv of Utils/parallel.jl:82-141:
    fill every rank's q_send, then copy it into the matching peer's q_recv."""
    for pr, q in zip(problems, qs):
        pr.start_exchange(np.asfortranarray(q), omp)
    for r, pr in enumerate(problems):
        for p, peer_rank in enumerate(pr.mesh.peer_parts):
            other = problems[peer_rank]
            po = other.mesh.peer_parts.index(r)
            pr.q_recv[p][...] = other.q_send[po]
