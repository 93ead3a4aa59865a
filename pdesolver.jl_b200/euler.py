"""Host-side mirror of the reference's physics-module API for the Euler hot path.

Same names, argument meaning and error behaviour as the reference:

* ``evalResidual(mesh, sbp, eqn, opts, t=0.0)`` -- ``src/solver/euler/euler.jl:111-175``:
  reads only ``eqn.q``, overwrites ``eqn.res`` (no Minv applied).
* ``rk4(f, h, t_max, mesh, sbp, eqn, opts, res_tol=-1.0, real_time=False)`` --
  ``src/NonlinearSolvers/rk4.jl:404-410``: advances ``eqn.q_vec`` and returns ``t``.
* ``EulerData`` -- the fields of ``EulerData_`` the path touches
  (``src/solver/euler/types.jl:401-721``): ``q, res, q_vec, res_vec, M, Minv, params``.

Everything numerical happens in ``libpdes_euler_b200.so`` (hand-written sm_100a
kernels) through the C ABI of ``include/pdes_euler_b200.h``; this module only
marshals the host arrays, exactly as the Julia shim in INTEGRATION.md does with
``ccall``.  There is no CPU fallback: without the library or without a B200 the
calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import BC_IDS, FEI_IDS, FLUX_IDS, SRC_IDS


class PDESolverError(RuntimeError):
    """ErrorException of the reference (unsupported option combinations, usage errors)."""


class PhysicsError(FloatingPointError):
    """``error("Negative density detected")`` / ``error("Negative pressure detected")``
    (euler.jl:552-556, 598-603); carries the offending element and node."""

    def __init__(self, msg, element, node):
        super().__init__(msg)
        self.element, self.node = element, node


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


class ParamType:
    """``ParamType`` (types.jl:57-173, 241-256): the scalars the kernels need."""

    def __init__(self, opts):
        g = float(opts.get("gamma", 1.4))
        self.gamma, self.gamma_1 = g, g - 1
        self.R = float(opts.get("R", 287.058))
        self.cv = self.R / (g - 1)
        self.Ma = float(opts.get("Ma", -1.0))
        self.aoa = float(opts.get("aoa", 0.0)) * np.pi / 180
        self.rho_free = 1.0
        self.p_free = float(opts.get("p_free", 1 / g))
        self.E_free = self.p_free / (g - 1) + 0.5 * self.Ma * self.Ma
        self.t = 0.0


class EulerData:
    """Solution data object of the Euler physics module, backed by a device context.

    ``q``/``res`` are Fortran-ordered ``[numDofPerNode, numNodesPerElement, numEl]``
    host arrays; ``q_vec``/``res_vec`` alias them (DG: types.jl:600-602).
    """

    def __init__(self, mesh, sbp, opts, device=0, comm=None):
        self.mesh, self.sbp, self.opts = mesh, sbp, opts
        self.params = ParamType(opts)
        nd, nn, nE = mesh.numDofPerNode, sbp.numnodes, mesh.numEl
        self.q = np.zeros((nd, nn, nE), order="F")
        self.res = np.zeros((nd, nn, nE), order="F")
        self.q_vec = self.q.reshape(-1, order="F")     # views: share memory with q / res
        self.res_vec = self.res.reshape(-1, order="F")
        assert np.shares_memory(self.q, self.q_vec)
        self._ctx = C.c_void_p(None)
        self._keep = []
        self._L = _cabi.lib()
        self._create(device)
        Minv = np.zeros((nd, nn, nE), order="F")
        self._check(self._L.pdes_get_minv(self._ctx, _ptr(Minv)))
        self.Minv = Minv.reshape(-1, order="F")
        self.M = 1.0 / self.Minv
        # page-lock eqn.q / eqn.res: they cross PCIe on every evalResidual / rk4 call
        self._pinned = [a for a in (self.q, self.res) if self._L.pdes_pin_host(_ptr(a), a.nbytes) == 0]
        if comm is not None:
            self.set_comm(*comm)

    # -- context ---------------------------------------------------------------------------------
    def _check(self, rc):
        if rc == 0:
            return
        msg = self._L.pdes_last_error(self._ctx if self._ctx else None)
        msg = msg.decode() if msg else f"pdes error {rc}"
        if rc > 0:
            e, n = C.c_int64(-1), C.c_int64(-1)
            self._L.pdes_last_error_location(self._ctx, C.byref(e), C.byref(n))
            raise PhysicsError(msg, e.value, n.value)
        raise PDESolverError(msg)

    def _create(self, device):
        mesh, sbp, opts, f = self.mesh, self.sbp, self.opts, self.sbp.face
        cfg = _cabi.PdesConfig()
        cfg.dim, cfg.nn, cfg.nfn, cfg.ss = mesh.dim, sbp.numnodes, f.numnodes, f.stencilsize
        cfg.norient, cfg.sparse_face = f.nbrperm.shape[1], int(f.sparse)
        cfg.index_base, cfg.device = 0, device
        cfg.nE, cfg.nF, cfg.nB = mesh.numEl, mesh.numInterfaces, mesh.numBoundaryFaces
        cfg.numBC, cfg.npeers = mesh.numBC, mesh.npeers
        cfg.volume_integral_type = int(opts.get("volume_integral_type", 1))
        cfg.face_integral_type = int(opts.get("face_integral_type", 1))
        try:
            cfg.flux_id = FLUX_IDS[opts.get("Flux_name", "RoeFlux")]
            cfg.volume_flux_id = FLUX_IDS[opts.get("Volume_flux_name", "StandardFlux")]
            cfg.src_id = SRC_IDS[opts.get("SRCname", "SRC0")]
            if cfg.face_integral_type == 2:      # default functor: input/read_input.jl:185
                cfg.face_element_id = FEI_IDS[opts.get("FaceElementIntegral_name", "ESLFFaceIntegral")]
            bc = [BC_IDS[opts.get(f"BC{i + 1}_name", "isentropicVortexBC")] for i in range(mesh.numBC)]
        except KeyError as e:
            raise PDESolverError(f"unsupported functor name {e}") from None
        cfg.check_density = int(opts.get("check_density", True))
        cfg.check_pressure = int(opts.get("check_pressure", True))
        p = self.params
        cfg.gamma, cfg.R, cfg.Ma, cfg.aoa = p.gamma, p.R, p.Ma, p.aoa
        cfg.rho_free, cfg.E_free = p.rho_free, p.E_free
        rc = self._L.pdes_create(C.byref(cfg), C.byref(self._ctx))
        if rc:
            self._ctx = C.c_void_p(None)
            self._check(rc)
        L, ctx = self._L, self._ctx
        perm = np.asfortranarray(np.asarray(f.perm, dtype=np.int64))
        nbr = np.asfortranarray(np.asarray(f.nbrperm, dtype=np.int64))
        self._check(L.pdes_set_operator(ctx, _ptr(_f64(sbp.Q)), _ptr(_f64(sbp.w)), _ptr(_f64(f.interp)),
                                        _ptr(perm), _ptr(nbr), _ptr(_f64(f.wface))))
        ifaces = np.ascontiguousarray(mesh.interfaces)
        bfaces = np.ascontiguousarray(mesh.bndryfaces)
        bo = np.ascontiguousarray(mesh.bndry_offsets, dtype=np.int64)
        bcs = np.ascontiguousarray(bc, dtype=np.int32)
        self._check(L.pdes_set_mesh(ctx, _ptr(_f64(mesh.dxidx)), _ptr(_f64(mesh.jac)), _ptr(_f64(mesh.coords)),
                                    _ptr(_f64(mesh.nrm_face)), _ptr(_f64(mesh.nrm_bndry)),
                                    _ptr(_f64(mesh.coords_bndry)), _ptr(ifaces), _ptr(bfaces), _ptr(bo), _ptr(bcs)))
        for pi in range(mesh.npeers):
            bl = np.ascontiguousarray(mesh.bndries_local[pi])
            si = np.ascontiguousarray(mesh.shared_interfaces[pi])
            ns = _f64(mesh.nrm_sharedface[pi])
            self._check(L.pdes_set_peer(ctx, pi, int(mesh.peer_parts[pi]), len(bl), _ptr(bl), _ptr(si), _ptr(ns)))
            if cfg.face_integral_type == 2:
                # parallel_data = element (input/read_input.jl:250-258): the element-data halo
                le = np.ascontiguousarray(mesh.local_element_lists[pi], dtype=np.int64)
                self._check(L.pdes_set_peer_elements(ctx, pi, len(le), _ptr(le), len(mesh.remote_global_elnum[pi]),
                                                     int(mesh.shared_element_offsets[pi])))

    def set_comm(self, unique_id: bytes, rank: int, nranks: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._L.pdes_set_comm(self._ctx, buf, rank, nranks))

    @staticmethod
    def get_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = _cabi.lib().pdes_get_unique_id(buf)
        if rc:
            raise PDESolverError(_cabi.lib().pdes_last_error(None).decode())
        return bytes(buf)

    def close(self):
        for a in getattr(self, "_pinned", []):
            self._L.pdes_unpin_host(_ptr(a))
        self._pinned = []
        if getattr(self, "_ctx", None):
            self._L.pdes_destroy(self._ctx)
            self._ctx = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- test hooks for the halo exchange without a second GPU ----------------------------------------
    def pack_send(self, peer_idx):
        n = len(self.mesh.bndries_local[peer_idx])
        out = np.zeros((self.mesh.numDofPerNode, self.sbp.face.numnodes, n), order="F")
        self._check(self._L.pdes_set_q(self._ctx, _ptr(self.q)))
        self._check(self._L.pdes_pack_send(self._ctx, peer_idx, _ptr(out)))
        return out

    def pack_send_elements(self, peer_idx):
        n = len(self.mesh.local_element_lists[peer_idx])
        out = np.zeros((self.mesh.numDofPerNode, self.sbp.numnodes, n), order="F")
        self._check(self._L.pdes_set_q(self._ctx, _ptr(self.q)))
        self._check(self._L.pdes_pack_send_elements(self._ctx, peer_idx, _ptr(out)))
        return out

    def inject_recv_elements(self, peer_idx, q_recv):
        self._check(self._L.pdes_inject_recv_elements(self._ctx, peer_idx, _ptr(_f64(q_recv))))

    def inject_recv(self, peer_idx, q_recv):
        self._check(self._L.pdes_inject_recv(self._ctx, peer_idx, _ptr(_f64(q_recv))))

    def timings(self):
        t = _cabi.PdesTimings()
        self._check(self._L.pdes_get_timings(self._ctx, C.byref(t)))
        return {n: getattr(t, n) for n, _ in t._fields_}

    def kernel_launch_count(self):
        return int(self._L.pdes_kernel_launch_count(self._ctx))


def evalResidual(mesh, sbp, eqn: EulerData, opts, t=0.0):
    """``evalResidual(mesh, sbp, eqn, opts, t)`` of the Euler module (euler.jl:111-175)."""
    eqn.params.t = t
    L, ctx = eqn._L, eqn._ctx
    # one call: eqn.q up, R(q), eqn.res down -- pipelined in chunks with the copies where the mesh allows it
    eqn._check(L.pdes_eval_residual_host(ctx, _ptr(eqn.q), _ptr(eqn.res), float(t)))
    return None


def evaldRdqProduct(mesh, sbp, eqn: EulerData, opts, v, out=None):
    """``evaldRdqProduct(mesh, sbp, eqn, opts, input_array, output_array)`` (src/interface2.jl:454-498):
    ``out = dR/dq(eqn.q) * v`` -- the matrix-free operator of the Jacobian-free Newton-Krylov path
    (NonlinearSolvers/newton_setup.jl:632-662).  The reference perturbs q by ``i*eps*v`` and takes
    ``imag(res)/eps``; the device evaluates the same residual on dual numbers."""
    v = np.asfortranarray(np.asarray(v, dtype=np.float64).reshape(eqn.q.shape, order="F"))
    if out is None:
        out = np.zeros_like(eqn.q, order="F")
    L, ctx = eqn._L, eqn._ctx
    eqn._check(L.pdes_set_q(ctx, _ptr(eqn.q)))
    eqn._check(L.pdes_eval_jvp(ctx, _ptr(v), _ptr(out)))
    return out


def lserk54(f, h, t_max, mesh, sbp, eqn: EulerData, opts, res_tol=-1.0, real_time=False):
    """``lserk54(f, h, t_max, mesh, sbp, eqn, opts; res_tol, real_time)`` (NonlinearSolvers/lserk.jl): the
    five-stage 2N-storage scheme behind ``run_type = 30``; same contract as ``rk4``."""
    return rk4(f, h, t_max, mesh, sbp, eqn, opts, res_tol=res_tol, real_time=real_time, _entry="pdes_lserk54")


def rk4(f, h, t_max, mesh, sbp, eqn: EulerData, opts, res_tol=-1.0, real_time=False, _entry="pdes_rk4"):
    """``rk4(f, h, t_max, mesh, sbp, eqn, opts; res_tol, real_time)`` (rk4.jl:404-410).

    ``f`` must be this module's ``evalResidual`` (the fused stage kernels evaluate
    it on the device).  Advances ``eqn.q_vec`` in place and returns ``t``; the
    per-step residual norms (the ``convergence.dat`` column) are left in
    ``eqn.convergence``.
    """
    if f is not evalResidual:
        raise PDESolverError("rk4: f must be pdesolver_jl_b200.evalResidual (device-resident right-hand side)")
    use_itermax = bool(opts.get("use_itermax", "itermax" in opts))
    itermax = int(opts["itermax"]) if use_itermax else -1
    L, ctx = eqn._L, eqn._ctx
    t_steps = int(round(t_max / h))
    cap = max(min(t_steps, max(itermax, 1)) if use_itermax else t_steps, 1)
    norms = np.zeros(cap)
    t_out, ns = C.c_double(0.0), C.c_int64(0)
    eqn._check(L.pdes_set_q(ctx, _ptr(eqn.q)))
    rc = getattr(L, _entry)(ctx, float(h), float(t_max), itermax, float(res_tol), int(bool(real_time)),
                    C.byref(t_out), _ptr(norms), cap, C.byref(ns))
    eqn._check(rc)
    eqn._check(L.pdes_get_q(ctx, _ptr(eqn.q)))
    eqn.convergence = norms[:ns.value].copy()
    return t_out.value


def diagnostics(mesh, sbp, eqn: EulerData, opts):
    """The functionals ``majorIterationCallback`` logs (solver/euler/euler.jl:330-407) for ``eqn.q``, evaluated on the
    device in one reduction pass after one residual evaluation: ``calcEntropyIntegral``, ``contractResEntropyVars``
    (``w^T R``), ``calcKineticEnergy``, ``calcKineticEnergydt`` (solver/euler/entropy_flux.jl:141-186, 414-485),
    ``volume`` and ``integrateQ`` (:231-247).  ``eqn.res`` is left holding R(q) on the device side only."""
    nd = eqn.q.shape[0]
    out = np.zeros(6 + nd)
    L, ctx = eqn._L, eqn._ctx
    eqn._check(L.pdes_set_q(ctx, _ptr(eqn.q)))
    eqn._check(L.pdes_diagnostics(ctx, _ptr(out)))
    return {"entropy_integral": out[0], "wT_res": out[1], "kinetic_energy": out[2], "kinetic_energy_dt": out[3],
            "volume": out[4], "integral_q": out[5:5 + nd].copy(), "enstrophy": out[5 + nd]}


def calcEnstrophy(mesh, sbp, eqn: EulerData, opts, q_arr=None):
    """solver/euler/entropy_flux.jl:322-355 (3D; ``write_enstrophy`` in majorIterationCallback, euler.jl:370-376)"""
    if mesh.dim != 3:
        raise PDESolverError("calcEnstrophy: 3D only (entropy_flux.jl:326)")
    return diagnostics(mesh, sbp, eqn, opts)["enstrophy"]


def calcEntropyIntegral(mesh, sbp, eqn: EulerData, opts, q_vec=None):
    """solver/euler/entropy_flux.jl:141-157"""
    return diagnostics(mesh, sbp, eqn, opts)["entropy_integral"]


def contractResEntropyVars(mesh, sbp, eqn: EulerData, opts, q_vec=None, res_vec=None):
    """solver/euler/entropy_flux.jl:164-186 with res_vec = the weak residual of eqn.q"""
    return diagnostics(mesh, sbp, eqn, opts)["wT_res"]


def integrateQ(mesh, sbp, eqn: EulerData, opts, q_vec=None):
    """solver/euler/entropy_flux.jl:231-247"""
    return diagnostics(mesh, sbp, eqn, opts)["integral_q"]


def calcKineticEnergy(mesh, sbp, eqn: EulerData, opts, q_vec=None):
    """solver/euler/entropy_flux.jl:414-440"""
    return diagnostics(mesh, sbp, eqn, opts)["kinetic_energy"]


def calcKineticEnergydt(mesh, sbp, eqn: EulerData, opts, q_vec=None, res_vec=None):
    """solver/euler/entropy_flux.jl:456-485 with res_vec = Minv R(eqn.q)"""
    return diagnostics(mesh, sbp, eqn, opts)["kinetic_energy_dt"]


def _krylov_opts(opts):
    """Linear-solver defaults of the input system (input/read_input.jl:493-496, 560-570: GMRES restart 30)."""
    return dict(reltol=float(opts.get("krylov_reltol", 1e-2)), abstol=float(opts.get("krylov_abstol", 1e-50)),
                dtol=float(opts.get("krylov_dtol", 1e5)), itermax=int(opts.get("krylov_itermax", 1000)),
                restart=int(opts.get("krylov_restart", 30)))


PC_IDS = {"none": 0, "PCNone": 0, "element_block_jacobi": 1, "bjacobi": 1}


def _set_pc(eqn, opts):
    """``opts["krylov_pc"]``: "none" (default) or "element_block_jacobi" -- the right preconditioner standing in for the
    reference's PETSc ``-pc_type bjacobi -ksp_pc_side right`` (input/read_input.jl:560-570)."""
    name = opts.get("krylov_pc", "none")
    if name not in PC_IDS:
        raise PDESolverError(f"unsupported preconditioner {name!r}")
    eqn._check(eqn._L.pdes_set_krylov_pc(eqn._ctx, PC_IDS[name]))


def linearSolve(mesh, sbp, eqn: EulerData, opts, b, x=None):
    """``linearSolve(ls, b, x)`` (linearsolvers) for the matrix-free operator of jac_type 4: solves
    ``dR/dq(eqn.q) x = b`` by restarted GMRES on the device (no preconditioner), zero initial guess.
    Returns ``x``; iteration count, final residual norm and PETSc-style reason are left in ``eqn.krylov_info``."""
    b = np.asfortranarray(np.asarray(b, dtype=np.float64).reshape(eqn.q.shape, order="F"))
    if x is None:
        x = np.zeros_like(eqn.q, order="F")
    k = _krylov_opts(opts)
    L, ctx = eqn._L, eqn._ctx
    its, rn, reason = C.c_int64(0), C.c_double(0.0), C.c_int32(0)
    _set_pc(eqn, opts)
    eqn._check(L.pdes_set_q(ctx, _ptr(eqn.q)))
    eqn._check(L.pdes_gmres(ctx, _ptr(b), _ptr(x), k["reltol"], k["abstol"], k["dtol"], k["itermax"], k["restart"],
                            C.byref(its), C.byref(rn), C.byref(reason)))
    eqn.krylov_info = {"iterations": its.value, "rnorm": rn.value, "reason": reason.value}
    return x


def newton(func, mesh, sbp, eqn: EulerData, opts, pmesh=None, t=0.0):
    """``newton(func, mesh, sbp, eqn, opts, pmesh=mesh, t=0.0)`` (NonlinearSolvers/newton.jl:54-74) for the
    matrix-free configuration (``jac_type = 4``: Jacobian-vector products of the residual, Krylov linear solves;
    BASELINE.json configuration 5).  Newton's method on ``R(q) = 0`` starting from ``eqn.q``: ``dR/dq dq = -R(q)`` by
    GMRES on the device, ``q += dq``, until ``res_abstol`` / ``res_reltol`` (relative to the first residual) /
    ``step_tol`` or ``itermax`` (newtonInner, newton.jl:137-304; checkConvergence :402-445).  The residual-norm and
    step-norm histories are left in ``eqn.convergence`` / ``eqn.step_norms``; ``eqn.newton_info`` holds the counts.
    Returns None like the reference."""
    if func is not evalResidual:
        raise PDESolverError("newton: func must be pdesolver_jl_b200.evalResidual (device-resident residual)")
    jac_type = int(opts.get("jac_type", 4))
    if jac_type != 4:
        raise PDESolverError("newton: only the matrix-free Jacobian (jac_type = 4) is implemented on the device")
    from ._cabi import PdesNewtonOpts, PdesNewtonResult
    k = _krylov_opts(opts)
    o = PdesNewtonOpts(itermax=int(opts.get("itermax", 10)), res_abstol=float(opts.get("res_abstol", 1e-6)),
                       res_reltol=float(opts.get("res_reltol", 1e-6)), step_tol=float(opts.get("step_tol", 0.0)),
                       step_fac=1.0, krylov_reltol=k["reltol"], krylov_abstol=k["abstol"], krylov_dtol=k["dtol"],
                       krylov_itermax=k["itermax"], krylov_restart=k["restart"])
    res_norms = np.zeros(o.itermax + 1)
    step_norms = np.zeros(max(o.itermax, 1))
    r = PdesNewtonResult()
    L, ctx = eqn._L, eqn._ctx
    eqn.params.t = t
    _set_pc(eqn, opts)
    eqn._check(L.pdes_set_q(ctx, _ptr(eqn.q)))
    eqn._check(L.pdes_newton_krylov(ctx, C.byref(o), _ptr(res_norms), _ptr(step_norms), C.byref(r)))
    eqn._check(L.pdes_get_q(ctx, _ptr(eqn.q)))
    eqn._check(L.pdes_get_res(ctx, _ptr(eqn.res)))
    eqn.convergence = res_norms[:r.newton_iters + 1].copy()
    eqn.step_norms = step_norms[:r.newton_iters].copy()
    eqn.newton_info = {"converged": bool(r.converged), "newton_iters": r.newton_iters, "krylov_iters": r.krylov_iters,
                       "residual_evals": r.residual_evals, "krylov_reason": r.krylov_reason,
                       "res_norm": r.res_norm, "res_norm_rel": r.res_norm_rel, "step_norm": r.step_norm}
    return None


def createObjects(mesh, sbp, opts, device=0):
    """Synthetic-mesh analogue of ``createObjects`` (solver/euler/startup_func.jl:45-66):
    returns ``(mesh, sbp, eqn, opts)`` with the device context initialised."""
    return mesh, sbp, EulerData(mesh, sbp, opts, device=device), opts
