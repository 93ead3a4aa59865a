/*
 * pdes_euler_b200.h -- C ABI of libpdes_euler_b200.so
 *
 * Drop-in boundary for PDESolver.jl's Euler hot path on one B200 (sm_100a):
 *
 *   evalResidual(mesh, sbp, eqn::EulerData, opts, t)   src/solver/euler/euler.jl:111-175
 *   rk4(f, h, t_max, mesh, sbp, eqn, opts; res_tol, real_time)
 *                                                       src/NonlinearSolvers/rk4.jl:144-344,404-410
 *   startSolutionExchange / finishExchangeData          src/Utils/parallel.jl:29-49,178-208
 *
 * The Julia host keeps the PumiInterface mesh, the SummationByParts operators
 * and the options dictionary; it uploads operator matrices, metrics and
 * connectivity once (pdes_set_operator / pdes_set_mesh / pdes_set_peer), then
 * drives pdes_eval_residual or pdes_rk4.  All array arguments are HOST pointers
 * in the reference's (Julia, column-major) layouts unless the name ends in
 * _dev.  No torch / CUDA types appear in any signature.  There is no CPU
 * fallback: every entry point fails with PDES_ERR_CUDA if no sm_100 device is
 * usable.
 *
 * Return convention (reference: Julia exceptions, euler.jl:552-556,598-603,653):
 *    0  success
 *   >0  physics error: PDES_ERR_NEG_DENSITY / PDES_ERR_NEG_PRESSURE; element and
 *       node (index_base applied) via pdes_last_error_location
 *   <0  usage / CUDA error; text via pdes_last_error
 */
#ifndef PDES_EULER_B200_H
#define PDES_EULER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct PdesCtx PdesCtx;

enum {
  PDES_OK = 0,
  PDES_ERR_NEG_DENSITY = 1,   /* euler.jl:543-570 checkDensity  */
  PDES_ERR_NEG_PRESSURE = 2,  /* euler.jl:586-611 checkPressure */
  PDES_ERR_USAGE = -1,
  PDES_ERR_CUDA = -2,
  PDES_ERR_UNSUPPORTED = -3,  /* euler.jl:653,796,855,863 ErrorException for option combos */
  PDES_ERR_COMM = -4
};
/* faceElementIntegrals.jl:735-741 FaceElementDict (face_integral_type 2): EC, Lax-Friedrichs and Lax-Wendroff entropy penalties */
enum { PDES_FEI_EC = 1, PDES_FEI_ELF_PENALTY = 2, PDES_FEI_ESLF = 3, PDES_FEI_ELW2_PENALTY = 4, PDES_FEI_ESLW2 = 5 };


/* FluxDict names (src/solver/euler/flux.jl) -> ids */
enum { PDES_FLUX_ROE = 1, PDES_FLUX_IR = 2, PDES_FLUX_IRSLF = 3, PDES_FLUX_STANDARD = 4 };
/* BCDict names (src/solver/euler/bc.jl:2342-2369) -> ids */
enum { PDES_BC_ISENTROPIC_VORTEX = 1, PDES_BC_EXP = 2, PDES_BC_FREESTREAM = 3, PDES_BC_NOPENETRATION = 4,
       PDES_BC_RHO1E2U3 = 5, PDES_BC_ALLONES = 6, PDES_BC_ZEROFLUX = 7, PDES_BC_NOPENETRATION_ES = 8 };
/* SRCDict names (src/solver/euler/source.jl) -> ids */
enum { PDES_SRC_NONE = 0, PDES_SRC_EXP = 1 };

/* ODLCommonTools Interface / Boundary records as Julia lays them out
 * (constructor order: src/solver/euler/shock_capturing_mesh.jl:248). */
typedef struct { uint32_t elementL, elementR; uint8_t faceL, faceR, orient, pad; } PdesInterface; /* 12 B */
typedef struct { uint32_t element; uint8_t face, pad[3]; } PdesBoundary;                          /* 8 B  */

typedef struct {
  int32_t dim;                  /* mesh.dim: 2 | 3                                  */
  int32_t nn;                   /* mesh.numNodesPerElement = sbp.numnodes           */
  int32_t nfn;                  /* mesh.numNodesPerFace = sbpface.numnodes          */
  int32_t ss;                   /* sbpface.stencilsize (dense faces); 1 for sparse  */
  int32_t norient;              /* size(sbpface.nbrperm, 2)                         */
  int32_t sparse_face;          /* 1: SparseFace of SBPDiagonalE operators          */
  int32_t index_base;           /* 1 when called from Julia, 0 from C/Python        */
  int32_t device;               /* CUDA device ordinal                              */
  int64_t nE, nF, nB;           /* numEl, numInterfaces, numBoundaryFaces           */
  int32_t numBC;                /* mesh.numBC                                       */
  int32_t npeers;               /* mesh.npeers                                      */
  int32_t volume_integral_type; /* opts["volume_integral_type"]: 1 | 2              */
  int32_t face_integral_type;   /* opts["face_integral_type"]: 1 | 2 (face-element)  */
  int32_t flux_id;              /* opts["Flux_name"]                                */
  int32_t volume_flux_id;       /* opts["Volume_flux_name"] (type 2 only)           */
  int32_t src_id;               /* opts["SRCname"]                                  */
  int32_t check_density;        /* opts["check_density"]                            */
  int32_t check_pressure;       /* opts["check_pressure"]                           */
  int32_t face_element_id;      /* opts["FaceElementIntegral_name"] (type 2 only): PDES_FEI_* */
  double gamma, R;              /* params.gamma, params.R   (types.jl:241-244)      */
  double Ma, aoa;               /* params.Ma, params.aoa [rad] (types.jl:246-247)   */
  double rho_free, E_free;      /* params.rho_free, params.E_free (types.jl:249-252)*/
} PdesConfig;

/* Timings: same field meaning as the reference's Timings struct
 * (src/Utils/Utils.jl:567-613), seconds accumulated since pdes_create, measured
 * with CUDA events on the library's streams. */
typedef struct {
  double t_send, t_dataprep, t_volume, t_face, t_sharedface, t_source, t_func, t_timemarch,
         t_wait, t_allreduce;
  int64_t n_residual_evals, n_kernel_launches;
} PdesTimings;

int pdes_create(const PdesConfig *cfg, PdesCtx **out);
void pdes_destroy(PdesCtx *ctx);
const char *pdes_last_error(const PdesCtx *ctx);       /* ctx may be NULL: last global error */
int pdes_last_error_location(const PdesCtx *ctx, int64_t *element, int64_t *node);

/* sbp.Q[nn,nn,dim], sbp.w[nn], sbpface.interp[ss,nfn], sbpface.perm[ss,dim+1]
 * (sparse faces: [nfn,dim+1]), sbpface.nbrperm[nfn,norient], sbpface.wface[nfn] */
int pdes_set_operator(PdesCtx *ctx, const double *Q, const double *w, const double *interp,
                      const int64_t *perm, const int64_t *nbrperm, const double *wface);

/* mesh.dxidx[dim,dim,nn,nE], jac[nn,nE], coords[dim,nn,nE], nrm_face[dim,nfn,nF],
 * nrm_bndry[dim,nfn,nB], coords_bndry[dim,nfn,nB], interfaces[nF], bndryfaces[nB],
 * bndry_offsets[numBC+1] (BC i owns [offsets[i], offsets[i+1])), bc_ids[numBC].
 * Valid until the metrics change (reference: updateMetricDependents, types.jl:898-915):
 * call again to invalidate. */
int pdes_set_mesh(PdesCtx *ctx, const double *dxidx, const double *jac, const double *coords,
                  const double *nrm_face, const double *nrm_bndry, const double *coords_bndry,
                  const PdesInterface *interfaces, const PdesBoundary *bndryfaces,
                  const int64_t *bndry_offsets, const int32_t *bc_ids);

/* One SharedFaceData (Utils/parallel_types.jl:70-183): peer rank, its
 * bndries_local[nfaces], shared interfaces[nfaces] (elementL local) and
 * mesh.nrm_sharedface[peer][dim,nfn,nfaces]. */
int pdes_set_peer(PdesCtx *ctx, int32_t peer_idx, int32_t peer_rank, int64_t nfaces,
                  const PdesBoundary *bndries_local, const PdesInterface *shared_interfaces,
                  const double *nrm_sharedface);

/* face_integral_type = 2 on a partitioned mesh (parallel_data = element, input/read_input.jl:250-258): the element-data
 * halo.  local_elements[nsend] = mesh.local_element_lists[peer] (this rank's elements the peer needs, in the peer's
 * remote-element order; getSendDataElement, Utils/parallel.jl:276-293); nrecv = number of the peer's elements in this
 * rank's halo; shared_element_offset = mesh.shared_element_offsets[peer], the element number shared_interfaces[j].elementR
 * starts from (calcSharedFaceElementIntegrals_element_inner, flux.jl:442-496).  index_base applies to both. */
int pdes_set_peer_elements(PdesCtx *ctx, int32_t peer_idx, int64_t nsend, const int64_t *local_elements, int64_t nrecv,
                           int64_t shared_element_offset);
/* test hooks of the element-data halo when no second GPU exists ([nd,nn,nsend] / [nd,nn,nrecv], host pointers) */
int pdes_pack_send_elements(PdesCtx *ctx, int32_t peer_idx, double *q_send_out);
int pdes_inject_recv_elements(PdesCtx *ctx, int32_t peer_idx, const double *q_recv);

/* Multi-GPU: NCCL communicator from a 128-byte ncclUniqueId the host broadcast
 * (replaces mesh.comm / MPI.Isend/Irecv!, parallel_types.jl:620-684).
 * The communicator carries the set-up, the halos of J*v and of the type-2 face
 * integrals, and the Krylov inner products.  The per-evaluation halo
 * (startSolutionExchange / finishExchangeData, Utils/parallel.jl:29-208) and the
 * stage-1 norm all-reduce go peer to peer, from inside the face / norm kernels: at the FIRST evaluation after this call -- which is
 * therefore collective over all ranks -- every rank exports its receive buffer
 * through CUDA IPC (one process per GPU, same node) and maps its neighbours';
 * if any mapping fails all ranks fall back to ncclSend/ncclRecv together.  Every
 * rank must then issue the same sequence of evaluations, as with MPI. */
int pdes_get_unique_id(uint8_t id_out[128]);
int pdes_set_comm(PdesCtx *ctx, const uint8_t id[128], int32_t rank, int32_t nranks);
/* Test hook when no second GPU exists: expose the packed send buffer and
 * inject the receive buffer by hand (host pointers, [nd,nfn,nfaces]). */
int pdes_pack_send(PdesCtx *ctx, int32_t peer_idx, double *q_send_out);
int pdes_inject_recv(PdesCtx *ctx, int32_t peer_idx, const double *q_recv);

/* eqn.q / eqn.res [nd,nn,nE] (alias eqn.q_vec / eqn.res_vec for DG). */
int pdes_set_q(PdesCtx *ctx, const double *q);
int pdes_get_q(PdesCtx *ctx, double *q);
int pdes_get_res(PdesCtx *ctx, double *res);
double *pdes_q_dev(PdesCtx *ctx);      /* device views, for callers that keep state in HBM */
double *pdes_res_dev(PdesCtx *ctx);
/* copy eqn.q from a DEVICE pointer (same layout), asynchronously on the library's compute stream */
int pdes_set_q_dev(PdesCtx *ctx, const double *q_dev);
/* page-lock / unlock a host array the caller will pass to pdes_set_q / pdes_get_q / pdes_get_res repeatedly
 * (eqn.q, eqn.res), so that the copies run at full PCIe rate */
int pdes_pin_host(void *ptr, int64_t bytes);
int pdes_unpin_host(void *ptr);
/* the library's compute stream (a cudaStream_t) so that a caller can bracket launches with its own events */
void *pdes_stream(PdesCtx *ctx);

/* evalResidual: q -> res (no Minv), synchronous.  euler.jl:111-175.  `t` is accepted for the signature's sake: every
 * restated source term and boundary functor (SRCExp, the eight BCs) is time independent, as in the named configurations. */
int pdes_eval_residual(PdesCtx *ctx, double t);
/* evalResidual with host arrays in one call: q[nd,nn,nE] in, res[nd,nn,nE] out (page-locked memory: pdes_pin_host).
 * The evaluation is pipelined in chunks with the upload of q and the download of res (full-duplex PCIe), for the
 * host-driven integrators that call evalResidual every stage; bit-identical to pdes_set_q + pdes_eval_residual +
 * pdes_get_res, which is also what partitioned or small meshes fall back to. */
int pdes_eval_residual_host(PdesCtx *ctx, const double *q, double *res, double t);
/* same, but returns after enqueueing; pdes_sync reports errors (bench / overlap) */
int pdes_eval_residual_async(PdesCtx *ctx, double t);
int pdes_sync(PdesCtx *ctx);

/* Jacobian-vector product out = dR/dq(q) * v at the resident q (no Minv), v/out [nd,nn,nE] host arrays: the
 * product the reference forms as imag(R(q + i*eps*v))/eps with eps = 1e-20 (evaldRdqProduct interface2.jl:454-498,
 * applyLinearOperator NonlinearSolvers/newton_setup.jl:632-662); evaluated here on dual numbers, exact to round-off
 * (pinned against a C99-complex restatement of the complex step, oracle/euler_oracle_cs.c).  Roe and entropy-stable
 * (IR / IRSLF, diagonal-E) configurations, any operator size; on a partitioned mesh (needs pdes_set_comm) the shared-face
 * states and directions are exchanged first -- collective.  PDES_ERR_UNSUPPORTED for face_integral_type 2. */
int pdes_eval_jvp(PdesCtx *ctx, const double *v, double *out);

/* Matrix-free Newton-Krylov on the resident q (configuration 5): the reference's
 * newton()/newtonInner (NonlinearSolvers/newton.jl:54-304) with jac_type=4 -- dR/dq * delta_q = -R(q) solved by
 * restarted GMRES on the complex-step product (here: pdes_eval_jvp's dual-number product), update
 * q += step_fac*delta_q, convergence tests of checkConvergence (newton.jl:402-445) on the strong-residual norm
 * sqrt(sum Minv res^2) (physicsRhs, jacobian/residual_evaluation.jl:64-88).  Linear-solver defaults:
 * krylov_reltol 1e-2, abstol 1e-50, dtol 1e5, itermax 1000 (read_input.jl:493-496), GMRES restart 30
 * (read_input.jl:569), preconditioner: pdes_set_krylov_pc (default none).  The Krylov basis and all reductions stay on the device. */
typedef struct PdesNewtonOpts {
  int64_t itermax;          /* Newton iterations (opts["itermax"]) */
  double res_abstol, res_reltol, step_tol, step_fac;
  double krylov_reltol, krylov_abstol, krylov_dtol;
  int64_t krylov_itermax;
  int32_t krylov_restart;
} PdesNewtonOpts;
typedef struct PdesNewtonResult {
  int32_t converged;        /* 1: a checkConvergence test passed */
  int32_t krylov_reason;    /* last linear solve: 1 rtol, 2 abstol, 3 breakdown (exact), -1 itermax, -2 dtol */
  int64_t newton_iters, krylov_iters, residual_evals;
  double res_norm, res_norm_rel, step_norm;
} PdesNewtonResult;
/* res_norms_out[itermax+1]: recordResNorm history (entry 0 = initial residual); step_norms_out[itermax] */
int pdes_newton_krylov(PdesCtx *ctx, const PdesNewtonOpts *o, double *res_norms_out, double *step_norms_out,
                       PdesNewtonResult *result);
/* Right preconditioner of the Krylov solves (the reference: PETSc options -pc_type bjacobi -ksp_pc_side right,
 * input/read_input.jl:560-570).  PDES_PC_ELEMENT_BLOCK_JACOBI: the element-diagonal blocks dR_e/dq_e of the DG Jacobian,
 * built matrix-free from coloured Jacobian-vector products at every Newton iterate and inverted on the device. */
#define PDES_PC_NONE 0
#define PDES_PC_ELEMENT_BLOCK_JACOBI 1
int pdes_set_krylov_pc(PdesCtx *ctx, int32_t pc_type);
/* One linear solve dR/dq(q) x = b with the same GMRES (x0 = 0); b, x host arrays [nd,nn,nE]. */
int pdes_gmres(PdesCtx *ctx, const double *b, double *x, double reltol, double abstol, double dtol,
               int64_t itermax, int32_t restart, int64_t *iters_out, double *rnorm_out, int32_t *reason_out);

/* rk4 (rk4.jl:144-344) on the resident q.  itermax < 0: use_itermax=false.
 * norms_out[norms_cap] receives the stage-1 norm of every executed step (the
 * convergence.dat column); nsteps_out the number of executed step heads;
 * t_out the value rk4 returns.  With itermax, q is left at x_old + (h/2) k1
 * exactly as the reference leaves it. */
int pdes_rk4(PdesCtx *ctx, double h, double t_max, int64_t itermax, double res_tol,
             int32_t real_time, double *t_out, double *norms_out, int64_t norms_cap,
             int64_t *nsteps_out);
/* lserk54 (NonlinearSolvers/lserk.jl:39-248; run_type 30, solver/common.jl:603-605): same arguments as pdes_rk4.
 * The res_tol / itermax tests precede the stage-1 update there, so those exits leave q untouched. */
int pdes_lserk54(PdesCtx *ctx, double h, double t_max, int64_t itermax, double res_tol,
                 int32_t real_time, double *t_out, double *norms_out, int64_t norms_cap,
                 int64_t *nsteps_out);
/* n plain RK4 steps, no host sync inside (bench inner loop; CUDA-graph replay) */
int pdes_rk4_steps_async(PdesCtx *ctx, double h, int64_t nsteps);

/* Functionals of majorIterationCallback (solver/euler/euler.jl:330-407) for the resident q, reduced on the device after
 * one residual evaluation: out[0] calcEntropyIntegral, out[1] contractResEntropyVars (w^T R), out[2] calcKineticEnergy,
 * out[3] calcKineticEnergydt (solver/euler/entropy_flux.jl:141-186, 414-485), out[4] mesh.volume (sum of M),
 * out[5..5+nd) integrateQ (entropy_flux.jl:231-247), out[5+nd] calcEnstrophy (entropy_flux.jl:322-355 with calcVorticity
 * euler_funcs.jl:1095-1155; 3D, 0 in 2D): out holds 6+nd doubles.  Per mesh part: the Allreduce over ranks stays with the host. */
int pdes_diagnostics(PdesCtx *ctx, double *out);

/* eqn.Minv[nd,nn,nE] as the reference computes it (mass_matrix.jl:20-44) */
int pdes_get_minv(PdesCtx *ctx, double *Minv);
int pdes_get_timings(PdesCtx *ctx, PdesTimings *out);
/* number of CUDA kernels this library has launched on ctx so far */
int64_t pdes_kernel_launch_count(const PdesCtx *ctx);

#ifdef __cplusplus
}
#endif
#endif
