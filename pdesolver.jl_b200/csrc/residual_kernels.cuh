// Euler residual + fused RK4 stage for sm_100a (dense-face SBP-Omega operators, Roe flux), fp64.
//
// One residual evaluation is two launches, both atomic-free and deterministic:
//
//   k_face_flux      one numerical flux per face node, evaluated ONCE per interface (the element-centric
//                    variant that recomputed it on both sides was FP64-pipe bound: profiles/r1_v2_*):
//                      interpolateFace / interiorFaceInterpolate!        flux.jl:613-641, 79-125
//                      calcFaceFlux + RoeSolver + calcSAT                flux.jl:37-64, bc_solvers.jl:29-420
//                      interpolateBoundary + getBCFluxes (BC functors)   bc.jl:49-80, 162-175, 251-284
//                      calcSharedFaceIntegrals_nopre_inner (flux part)   flux.jl:264-308
//                    output: wface[i] * flux[:, i] per face, 8*nd*nfn bytes per face in HBM/L2.
//   k_element_rk     everything that is owned by one element, every res entry produced by one thread:
//                      dataPrep checks, getEulerFlux, weakdifferentiate! euler.jl:441-519, 543-611, 628-658
//                      interiorfaceintegrate! / boundaryintegrate! / boundaryFaceIntegrate!  (gather form)
//                      applySourceTerm (tabulated), pde_post_func, RK4 axpy rk4.jl:244-319, 446-457
//
// The reference materialises aux_vars / flux_parametric / q_face / flux_face / q_bndry / bndryflux and makes
// >= 12 sweeps (euler.jl:441-519); here the only intermediate is the face flux.
//
// Thread roles inside a CTA exchange data through shared memory:
//   "variable threads"  (tile item, variable k): all operator applications (Q^T F, R q, R^T W f) are small dense
//                       products whose coefficients are compile-time-indexed entries of the kernel-parameter
//                       operator table (uniform constant loads feeding DFMA); the row lives in registers.
//   "node threads"      (tile item, node): pointwise nonlinear work (Euler flux, Roe flux, BC functors).
#pragma once
#include <stdint.h>
#include "euler_device.cuh"

#ifndef PDES_OPT
#define PDES_OPT 0      // all three measured slower on C3 (1.101 / 1.141 / 1.129 vs 1.093 ms per RK4 step); bit 3: face normals requested at the top of the tile. bit 0: pipelined metrics loads in S1; bit 1: face records two faces ahead; bit 2: staged Minv
#endif

#ifndef PDES_SKEL
#define PDES_SKEL 0     // measurement builds only (results are meaningless): bit 0 drops the operator / Roe arithmetic,
#endif                  // 1: no face-record loads (element), 2: no volume-flux tile (element), 3: no L2 prefetches,
                        // 4: no element gathers (face), 5: no record stores (face), 6: no epilogue streams (element)

namespace pdes {

// host-side bookkeeping per (element, local face): which face covers it (every element face must be claimed by
// exactly one interface, boundary face or shared face)
struct EFace {
  int32_t gface;
  uint8_t right, orient, pad[2];
};

// per face: what k_face_flux gathers
struct __align__(16) FaceRec {
  int32_t elL, elR;  // elR: right element (interior) | index into mesh.bndryfaces (boundary) | unused (shared)
  uint8_t fL, fR, orient, kind;
  int32_t aux;       // boundary: BC functor id; shared: index into the receive buffer
};
enum FaceKind : uint8_t { FK_INTERIOR = 0, FK_BOUNDARY = 2, FK_SHARED = 3 };

struct Ctl {               // device-resident control block
  int32_t stop;            // kernels return immediately when set (physics error or res_tol reached)
  int32_t err_code;        // != 0: physics error, decoded on the host from err_loc
  unsigned long long err_loc;  // ((code-1) << 62) | (element << 8) | node of the lowest offending location
  int32_t converged_step;  // step head at which norm < res_tol (or -1)
  int32_t norm_count;      // step heads whose norm has been committed (the slot the next commit writes)
  int32_t kry_done;        // GMRES: the convergence test passed on the device; the Krylov / J*v kernels enqueued behind it are no-ops
};

template <int DIM, int NN, int NFN>
struct OpTab {
  static constexpr int NF = DIM + 1;
  static constexpr int NOR = (DIM == 2) ? 1 : 3;
  double Qt[DIM * NN][NN];    // Qt[d*NN+j][i] = sbp.Q[j,i,d]   (res_i += Q[j,i,d] F_j : weakdifferentiate!, trans=true)
  double RfN[NF * NFN][NN];   // RfN[f*NFN+i][node] = sum_j interp[j,i] [perm[j,f]==node]   (face integration)
  double interp[NN][NFN];     // sbpface.interp[j,i] (stencil order: face interpolation after a perm-ordered gather)
  double wface[NFN];
  int32_t perm[NF][NN];       // sbpface.perm[j,f] (0-based)
  int32_t nbrperm[NOR][NFN];  // sbpface.nbrperm[i,orient] (0-based)
};

struct PhysPar {
  double gamma, R, Ma, aoa, rho_free, E_free;
  int32_t check_density, check_pressure;
};

enum EpiMode { EPI_RES = 0, EPI_RK = 1 };

// Chunk pipeline (PDES_PIPE, pdes_api.cu: enqueue_residual_pipe): the face list and the element range are cut into
// chunks and launched alternately, F0 F1 E0 F2 E1 ..., every launch with programmatic stream serialization, so that
// the CTAs of a launch start as soon as the LAST WAVE of the previous launch is resident (no drain between launches).
// Data dependencies are then carried by counters instead of kernel boundaries: the last CTA of chunk c of a family
// publishes done[c] = epoch, in chunk order.
struct PipeArgs {
  unsigned* arrive;          // [chunk] CTAs of this launch that have finished (re-armed by the last one)
  unsigned* done_self;       // [chunk] epoch of the last complete launch of this family on chunk c
  const unsigned* done_dep;  // the other family's done[]: what this launch consumes
  int32_t on, chunk, ncta, dep_chunk;   // wait for done_dep[dep_chunk] >= dep_epoch (publication is in chunk order)
  uint32_t epoch, dep_epoch;
};

// Halo exchange fused into the face kernel (k_face_tma; pdes_api.cu: start_exchange).  Replaces startSolutionExchange /
// finishExchangeData (Utils/parallel.jl:29-49, 178-208) AND the launches around them: the warps of the ONE face launch of an
// evaluation first interpolate the shared faces of this rank and store the states straight into the neighbours' receive
// buffers (peer memory over NVLink), the last of them publishes the evaluation number in every neighbour's flag slot; the
// shared faces are the last tiles of the same launch and poll the local flags before they read the received states.
struct HaloArgs {
  int32_t on, npeers;
  int64_t nS, s0;                  // shared faces: count, index of the first one in the face list
  double* const* face_dst;         // [2][nS] (by evaluation parity) slot of the face in the neighbour's receive buffer
  unsigned* const* peer_flags;     // [npeers] this rank's flag slot in the neighbour's buffer (+32: abort slot)
  const unsigned* flags;           // [32] evaluation number per neighbour | [32] abort | [2] {epoch, pack counter}
  unsigned* ctr;                   // -> {epoch of the last complete evaluation, pack tiles done} (local)
  const double* recv_base;         // [2][nsend] local receive buffers
  int64_t nsend;
};

struct FaceArgs {
  const double* q;             // [ND,NN,nE]
  const FaceRec* faces;        // [nF + nB + nS]
  const double* nrm;           // [DIM,NFN,nF+nB+nS]   nrm_face | nrm_bndry | nrm_sharedface  (or [DIM,nG] when every
                               // face has node-independent normals: straight-sided meshes, detected at upload)
  int32_t nrm_face_stride, nrm_node_stride;   // in doubles: (NFN*DIM, DIM) or (DIM, 0)
  const double* coords_bndry;  // [DIM,NFN,nB]
  const double* q_recv;        // [ND,NFN,nS] (peer's own face-node order)
  const double* v_recv;        // J*v on a partitioned mesh: the direction v at the neighbours' shared-face nodes, same layout
  double* fluxe;               // [ND,NFN,dim+1,nE]: per (element, local face) the contribution the element integrates,
                               // -wface_i f*(:,i) for elementL, +wface_i f*(:,i) in elementR's own node order
  int64_t g0, ng;              // face range of this launch
  int32_t prefetch_ahead;      // tiles between this CTA and the one whose gathers it prefetches into L2
  int32_t ext_bc;              // some boundary face uses a functor of bc_flux_ext (bc >= 5)
  const Ctl* ctl;
  PhysPar ph;
  PipeArgs pipe;
  const int32_t* tab_dev;      // device copy of OpTab::perm | OpTab::nbrperm (or nullptr)
  const double* optab_dev;     // device copy of OpTab::interp | OpTab::wface (k_face_element)
  HaloArgs halo;
};

struct ElemArgs {
  const double* q;             // [ND,NN,nE]
  const double* dxidx;         // [DIM,DIM,NN,nE]  (or [DIM,DIM,nE] when node-independent: straight-sided elements)
  int32_t dx_el_stride, dx_node_stride;       // in doubles: (NN*DIM*DIM, DIM*DIM) or (DIM*DIM, 0)
  const double* fluxe;         // [ND,NFN,dim+1,nE] from k_face_flux
  const double* srcw;          // [ND,NN,nE] (w_j/jac_j) * S(x_j), or nullptr   (EPI_RES)
  const double* srcm;          // [ND,NN,nE] Minv_j * srcw                     (EPI_RK)
  double* res;                 // EPI_RES: [ND,NN,nE]
  // EPI_RK (rk4.jl:244-319): k = Minv*res; q_next = x_old + ah*k; ksum updated; last stage: x_new
  const double* minv;          // [NN,nE]  1/(w_j/jac_j)   (mass_matrix.jl:20-44)
  const double* mass;          // [NN,nE]  w_j/jac_j = eqn.M: the weight of calcNorm (Utils.jl:427-449)
  const double* x_old;
  double* ksum;
  double* q_next;
  double* norm_partials;       // [gridDim.x] sum_j M_j k_j^2 per CTA (stage 1 only)
  double ah;                   // a_s * h
  double h6;                   // h/6 (last stage)
  double hh;                   // LSERK54: delta_t  (then ah = a_s, h6 = b_s, ksum = dq_vec, x_old = q)
  int32_t scheme;              // 0: rk4 (rk4.jl:244-319), 1: lserk54 (lserk.jl:183-205), 2: rk4 without the running sum
  int32_t stage;               // 1..4 (rk4) | 1..5 (lserk54)
  int32_t prefetch_ahead;      // tiles between this CTA and the one whose inputs it prefetches into L2
  int32_t discard_records;     // drop the consumed face records from L2 (discard.global.L2): no write-back
  int32_t reverse;             // k_element_rk sweeps its tiles from the last to the first (see pdes_api.cu: L2 reuse)
  int32_t stagger_ns;          // k_element_tma: warp w starts (w mod 4) * stagger_ns later (de-synchronises the tile phases)
  unsigned* halo_epoch;        // fused halo (HaloArgs): the evaluation is complete -> ++*halo_epoch (block 0), or nullptr
  unsigned* tile_ctr;          // k_element_tma, dynamic tile deal: {tiles handed out beyond the first round, warps finished}
  int64_t e_begin, nE;         // element range [e_begin, nE) of this launch (e_begin a multiple of the tile size)
  Ctl* ctl;
  PhysPar ph;
  PipeArgs pipe;
  const double* s2_dev;        // split-form kernels: device copy of OpTabS::S2 (staged by cp.async) ...
  const int32_t* inv_dev;      // ... and of OpTabS::inv
};

// ---- chunk pipeline primitives -------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_relaxed_u32(const void* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// the launch that follows this one in the stream may start (its CTAs fill the slots our last wave leaves free)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// returns when the launch that precedes this one in the stream has completed and flushed
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// one thread: wait until *p >= target (epochs wrap: signed difference).  Deadlock-free: the launch that publishes *p is
// earlier in the stream and every one of its CTAs is resident or done before ours may start (launch_dependents is
// the first thing a CTA does).  The wait gives up when another CTA raised `stop` and after ~1 s (err_code 3) instead
// of hanging the device.
__device__ __forceinline__ bool pipe_spin(const unsigned* p, unsigned target, Ctl* ctl) {
  for (unsigned it = 0;; ++it) {
    if ((int)(ld_relaxed_u32(p) - target) >= 0) return true;
    if ((it & 31u) == 31u && ld_relaxed_u32(&ctl->stop)) return false;
    if (it > (1u << 22)) {
      atomicExch(&ctl->err_code, 3);
      atomicExch(&ctl->stop, 1);
      return false;
    }
    __nanosleep(100);
  }
}
// consumer side, called by ONE thread before the block barrier that precedes the first dependent access
__device__ __forceinline__ void pipe_acquire(const PipeArgs& pp, Ctl* ctl) {
  if (pp.dep_chunk >= 0) pipe_spin(pp.done_dep + pp.dep_chunk, pp.dep_epoch, ctl);
  if (!(pp.on & 2)) asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// producer side, called by ONE thread after a block barrier that follows the CTA's last store
__device__ __forceinline__ void pipe_release(const PipeArgs& pp, Ctl* ctl) {
  if (!(pp.on & 4)) __threadfence();
  const unsigned old = atomicAdd(pp.arrive + pp.chunk, 1u);
  if (old + 1u == (unsigned)pp.ncta) {
    pp.arrive[pp.chunk] = 0u;                                  // re-armed for the next evaluation
    // publication in chunk order: done[c] >= e then implies done[c'] >= e for every c' < c
    if (pp.chunk > 0) pipe_spin(pp.done_self + pp.chunk - 1, pp.epoch, ctl);
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(pp.done_self + pp.chunk), "r"(pp.epoch) : "memory");
  }
}

// Last launch of a pipelined evaluation (one thread).  A launch that does not wait for its predecessor may COMPLETE
// before it, so an ordinary stream operation that follows the evaluation (norm reduction, copy, event) must not rely
// on the completion of the last chunk alone: this kernel completes only when every element chunk -- and with them every
// face chunk -- of the evaluation has published its results.
__global__ void k_pipe_join(const unsigned* done_elem_last, unsigned epoch, Ctl* ctl) {
  pdl_launch_dependents();
  if (ld_relaxed_u32(&ctl->stop)) return;
  pipe_spin(done_elem_last, epoch, ctl);
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  if (!(PDES_SKEL & 8)) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Ampere-style asynchronous global->shared copies (LDGSTS): issued at the top of a tile, consumed stages later
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
// streaming variant: the line is marked evict-first in L2 (data touched once per launch must not push out q / the face
// records that the next kernel re-reads)
__device__ __forceinline__ void cp_async16_stream(void* smem, const void* gmem, unsigned long long pol) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// contiguous tile of n doubles, both sides 16-byte aligned (tile bases are: even tile sizes, 256-B aligned arrays)
__device__ __forceinline__ void async_tile(double* dst, const double* src, int n, int tid, int T) {
  const int n2 = n >> 1;
  for (int i = tid; i < n2; i += T) cp_async16(dst + 2 * i, src + 2 * i);
  if ((n & 1) && tid == 0) cp_async8(dst + n - 1, src + n - 1);
}

__device__ __forceinline__ void async_tile_stream(double* dst, const double* src, int n, int tid, int T,
                                                  unsigned long long pol) {
  const int n2 = n >> 1;
  for (int i = tid; i < n2; i += T) cp_async16_stream(dst + 2 * i, src + 2 * i, pol);
  if ((n & 1) && tid == 0) cp_async8(dst + n - 1, src + n - 1);
}

__host__ __device__ constexpr int pad_stride(int n, int nd) {
  // smallest m >= n with m % 16 == nd % 16: (item, variable)-indexed fp64 accesses of a half-warp
  // then fall into distinct banks
  int m = n;
  while (m % 16 != nd % 16) ++m;
  return m;
}

// noPenetrationESBC (bc.jl:767-793): qg = q with the normal momentum negated (getDirichletState :860-918), flux =
// calcLFFlux(q, qg) = (F(q) + F(qg) - lambda_max (qg - q)) / 2 with lambda_max = getLambdaMax at the average state
// (bc_solvers.jl:428-446, IR_stab.jl:310-325, euler_funcs.jl:1887-1913 with absvalue3).  T = double | Dual.
template <int DIM, typename T>
__device__ __forceinline__ void noslip_es_flux(const T* q, const double* n, double gamma, T* flux) {
  constexpr int ND = DIM + 2;
  double nn2 = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) nn2 += n[d] * n[d];
  const double dA = ::sqrt(nn2), fac = 1.0 / dA;
  double nh[DIM];
  T Unrm = T(0.0);
#pragma unroll
  for (int d = 0; d < DIM; ++d) { nh[d] = n[d] * fac; Unrm += T(nh[d]) * q[1 + d]; }
  T qg[ND], fL[ND], fR[ND], qa[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) qg[i] = q[i];
#pragma unroll
  for (int d = 0; d < DIM; ++d) qg[1 + d] = T(-2.0 * nh[d]) * Unrm + q[1 + d];
  euler_flux<DIM, T>(q, n, gamma - 1.0, fL);
  euler_flux<DIM, T>(qg, n, gamma - 1.0, fR);
#pragma unroll
  for (int i = 0; i < ND; ++i) qa[i] = T(0.5) * (q[i] + qg[i]);
  const T p = calc_pressure<DIM, T>(qa, gamma - 1.0);
  const T rinv = T(1.0) / qa[0];
  T Un = T(0.0);
#pragma unroll
  for (int d = 0; d < DIM; ++d) Un += T(n[d]) * qa[1 + d] * rinv;
  const T a2 = T(gamma) * p * rinv;
  const T a = a2 * fast_rsqrt(a2);                       // sqrt(a2)
  const T lam = absvalue3(Un) + T(dA) * a;
#pragma unroll
  for (int i = 0; i < ND; ++i) flux[i] = T(0.5) * (fL[i] + fR[i] - lam * (qg[i] - q[i]));
}

// boundary-condition functors (bc.jl:554-567, 1756-1768, 1573-1587, 717-765): Dirichlet state + Roe, or
// Euler flux of the wall-projected state
template <int DIM>
__device__ __noinline__ void bc_flux(int bc, const double* q, const double* x, const double* n, const PhysPar& ph,
                                     double* flux) {
  constexpr int ND = DIM + 2;
  double qg[ND];
  if (bc == 4) {  // noPenetrationBC
    double nn2 = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) nn2 += n[d] * n[d];
    double fac = 1.0 / sqrt(nn2), Unrm = 0.0, nh[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { nh[d] = n[d] * fac; Unrm += nh[d] * q[1 + d]; }
#pragma unroll
    for (int i = 0; i < ND; ++i) qg[i] = q[i];
#pragma unroll
    for (int d = 0; d < DIM; ++d) qg[1 + d] -= nh[d] * Unrm;
    euler_flux<DIM>(qg, n, ph.gamma - 1.0, flux);
    return;
  }
  if (bc == 1) isentropic_vortex<DIM>(x, ph.gamma, ph.R, qg);
  else if (bc == 2) calc_exp<DIM>(x, ph.gamma, qg);
  else free_stream<DIM>(ph.rho_free, ph.E_free, ph.Ma, ph.aoa, qg);
  roe_flux<DIM>(q, qg, n, ph.gamma, flux);
}

// The functors beyond the four of the named configurations live in their own out-of-line function, called by the
// kernels for bc >= 5: folded into bc_flux they enlarge its frame and register footprint, which the kernels pay for on
// the interior path too (measured: +3 % on C3).
template <int DIM>
__device__ __noinline__ void bc_flux_ext(int bc, const double* q, const double* n, const PhysPar& ph, double* flux) {
  constexpr int ND = DIM + 2;
  if (bc == 7) {  // ZeroFluxBC (bc.jl:2140-2152)
#pragma unroll
    for (int i = 0; i < ND; ++i) flux[i] = 0.0;
    return;
  }
  if (bc == 8) {  // noPenetrationESBC (bc.jl:767-793, 860-918): reflected state + calcLFFlux (bc_solvers.jl:428-446)
    noslip_es_flux<DIM, double>(q, n, ph.gamma, flux);
    return;
  }
  double qg[ND];
  if (bc == 5) {  // Rho1E2U3BC (bc.jl:1454-1537, calcRho1Energy2U3 common_funcs.jl:754-779)
    qg[0] = 1.0; qg[DIM + 1] = 2.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) qg[1 + d] = 0.35355;
  } else {        // allOnesBC (bc.jl:1702-1722)
#pragma unroll
    for (int i = 0; i < ND; ++i) qg[i] = 1.0;
  }
  roe_flux<DIM>(q, qg, n, ph.gamma, flux);
}

// dispatch used by every face kernel
template <int DIM>
__device__ __forceinline__ void bc_flux_any(int bc, const double* q, const double* x, const double* n, const PhysPar& ph,
                                            double* flux) {
  if (bc >= 5) bc_flux_ext<DIM>(bc, q, n, ph, flux);
  else bc_flux<DIM>(bc, q, x, n, ph, flux);
}

// ------------------------------------------------------------------------------------------------------
// k_face_flux: FT faces per CTA
// ------------------------------------------------------------------------------------------------------
template <int DIM, int NN, int NFN, int FT>
struct FaceCfg {
  static constexpr int ND = DIM + 2;
  static constexpr int PER = ND > NFN ? ND : NFN;             // threads per face
  static constexpr int T = ((FT * PER + 31) / 32) * 32;
  static constexpr int FS = (pad_stride(NFN * ND, ND) + 1) & ~1;   // per-face stride of a face-state tile (even: LDS.128)
};

// shared-memory working set of one face tile (aliased onto the element tile's storage by the fused kernel)
template <int DIM, int NN, int NFN, int FT>
struct FaceTileSmem {
  static constexpr int FS = FaceCfg<DIM, NN, NFN, FT>::FS;
  double sL[FT * FS];
  double sR[FT * FS];
  FaceRec sRec[FT];
  int s_dst[2 * FT];            // (element*NF + face) of the record each side of a face writes, or -1
  int s_perm[DIM + 1][NN];      // BYTE offset of volume node perm[j,f] inside an element block (8 * ND * perm)
  int s_nbrperm[OpTab<DIM, NN, NFN>::NOR][NFN];
};

template <int DIM, int NN, int NFN, int FT, int TB>
__device__ __forceinline__ void face_tables(const OpTab<DIM, NN, NFN>& op, FaceTileSmem<DIM, NN, NFN, FT>& sm, int tid,
                                            const int32_t* dev = nullptr) {
  constexpr int NF = DIM + 1, NOR = OpTab<DIM, NN, NFN>::NOR;
  if (dev) {
    // device copy (perm | nbrperm): one coalesced load per warp; the parameter bank is read with a per-thread index
    // otherwise, which the hardware serialises address by address (4 % of k_face_flux's stall samples)
    for (int idx = tid; idx < NF * NN; idx += TB) (&sm.s_perm[0][0])[idx] = __ldg(dev + idx) * ((DIM + 2) * 8);
    for (int idx = tid; idx < NOR * NFN; idx += TB) (&sm.s_nbrperm[0][0])[idx] = __ldg(dev + NF * NN + idx);
    return;
  }
  for (int idx = tid; idx < NF * NN; idx += TB) sm.s_perm[idx / NN][idx % NN] = op.perm[idx / NN][idx % NN] * ((DIM + 2) * 8);
  for (int idx = tid; idx < NOR * NFN; idx += TB)
    sm.s_nbrperm[idx / NFN][idx % NFN] = op.nbrperm[idx / NFN][idx % NFN];
}

// One tile of nf <= FT faces starting at face g0, executed by a CTA of TB threads.  ga = first face of the tile whose
// gathers are prefetched into L2 (or < 0).  The caller has filled sm.s_perm / sm.s_nbrperm (face_tables) and puts a
// block barrier between consecutive tiles.
// barrier of the threads that share a face tile: the CTA, or one warp (k_face_flux_w: TB = 32)
template <bool WARPSYNC>
__device__ __forceinline__ void tile_sync() {
  if (WARPSYNC) __syncwarp();
  else __syncthreads();
}

template <int DIM, int NN, int NFN, int FT, int TB, bool WARPSYNC = false, bool PRELOADED = false, bool EXTBC = false>
__device__ __forceinline__ void face_tile(const OpTab<DIM, NN, NFN>& op, const FaceArgs& a,
                                          FaceTileSmem<DIM, NN, NFN, FT>& sm, int64_t g0, int nf, int64_t ga, int tid,
                                          FaceRec* carry = nullptr) {
  using Cfg = FaceCfg<DIM, NN, NFN, FT>;
  constexpr int ND = Cfg::ND, FS = Cfg::FS, NF = DIM + 1, EL = NN * ND;
  static_assert(TB >= Cfg::T, "block too small for the face tile");
  double* sL = sm.sL;
  double* sR = sm.sR;
  FaceRec* sRec = sm.sRec;
  const int64_t gend = a.g0 + a.ng;
  if (tid < nf) {
    // PRELOADED (persistent kernel): this tile's record was fetched while the previous tile ran
    const FaceRec r = PRELOADED ? *carry : a.faces[g0 + tid];
    sRec[tid] = r;
    sm.s_dst[2 * tid] = r.elL * NF + r.fL;
    sm.s_dst[2 * tid + 1] = r.kind == FK_INTERIOR ? r.elR * NF + r.fR : -1;
  }
  // the tile that will run on this SM slot next: fetch its records now, prefetch what they point at when done
  FaceRec nxt;
  nxt.kind = 255;
  if (ga >= 0 && tid < FT && ga + tid < gend) nxt = a.faces[ga + tid];
#if (PDES_OPT & 8)
  // the normal of this thread's face node depends on the face number only: requested here, consumed in stage B
  double nrm_early[DIM];
  if (tid < nf * NFN) {
    const double* np_ = a.nrm + (g0 + tid / NFN) * a.nrm_face_stride + (tid % NFN) * a.nrm_node_stride;
#pragma unroll
    for (int d = 0; d < DIM; ++d) nrm_early[d] = __ldg(np_ + d);
  }
#endif
  tile_sync<WARPSYNC>();

  // ---- A: interpolate both sides to the face nodes (variable threads) ----------------------------------
  if (tid < nf * ND) {
    const int fi = tid / ND, k = tid - fi * ND;
    const FaceRec r = sRec[fi];
    double ql[NN], qr[NN];
    {
      const double* b = a.q + (int64_t)r.elL * EL + k;
#pragma unroll
      for (int j = 0; j < NN; ++j)      // (one 64-bit add per address: the index arithmetic was 6 of 7 instructions per load)
        ql[j] = (PDES_SKEL & 16) ? 1.0 + j
                                 : __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(b) + (unsigned)sm.s_perm[r.fL][j]));
    }
    if (r.kind == FK_INTERIOR) {
      const double* b = a.q + (int64_t)r.elR * EL + k;
#pragma unroll
      for (int j = 0; j < NN; ++j)
        qr[j] = (PDES_SKEL & 16) ? 2.0 + j
                                 : __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(b) + (unsigned)sm.s_perm[r.fR][j]));
    } else {
#pragma unroll
      for (int j = 0; j < NN; ++j) qr[j] = 0.0;
    }
    // both sides share every interpolation coefficient (one uniform load feeds two DFMAs)
#pragma unroll
    for (int i = 0; i < NFN; ++i) {
      double s = 0.0, t = 0.0;
#if (PDES_SKEL & 1)
#pragma unroll
      for (int j = i; j < NN; j += NFN) { s += ql[j]; t += qr[j]; }
#else
#pragma unroll
      for (int j = 0; j < NN; ++j) {
        const double c = op.interp[j][i];
        s = fma(c, ql[j], s);
        t = fma(c, qr[j], t);
      }
#endif
      sL[fi * FS + i * ND + k] = s;
      // elementR's face node i coincides with elementL's face node nbrperm[i,orient] (involution)
      if (r.kind == FK_INTERIOR) sR[fi * FS + sm.s_nbrperm[r.orient][i] * ND + k] = t;
    }
    if (r.kind == FK_SHARED) {
      // permuteinterface! (Utils/parallel.jl:198-201): received node i of the peer is own node nbrperm[i,orient]
      const double* b = a.q_recv + (int64_t)r.aux * (NFN * ND) + k;
#pragma unroll
      for (int i = 0; i < NFN; ++i) sR[fi * FS + sm.s_nbrperm[r.orient][i] * ND + k] = b[i * ND];
    }
  }
  tile_sync<WARPSYNC>();

  // ---- B: numerical flux at every face node (node threads) ------------------------------------------------
  // results overwrite the face-state tiles: sL <- -w f* in elementL's node order, sR <- +w f* in elementR's
  // node order (the contributions interiorfaceintegrate! hands to the two elements)
  {
    const bool nact = tid < nf * NFN;
    const int fi = nact ? tid / NFN : 0, i = tid - fi * NFN;
    const FaceRec r = sRec[fi];
    const int64_t g = g0 + fi;
    double nrm[DIM], qL[ND], qR[ND], flux[ND];
    if (nact) {
#if (PDES_OPT & 8)
#pragma unroll
      for (int d = 0; d < DIM; ++d) nrm[d] = nrm_early[d];
#else
      const double* np_ = a.nrm + g * a.nrm_face_stride + i * a.nrm_node_stride;
#pragma unroll
      for (int d = 0; d < DIM; ++d) nrm[d] = __ldg(np_ + d);
#endif
#pragma unroll
      for (int k = 0; k < ND; ++k) { qL[k] = sL[fi * FS + i * ND + k]; qR[k] = sR[fi * FS + i * ND + k]; }
    }
    tile_sync<WARPSYNC>();       // every node thread holds its inputs: the tiles may be overwritten
    if (nact) {
      if (r.kind == FK_BOUNDARY) {
        // separate copies so that only this (rare) path touches local memory
        const double* xp = a.coords_bndry + ((int64_t)r.elR * NFN + i) * DIM;   // elR of a boundary face: its index in bndryfaces
        double xb[DIM], nb_[DIM], qb[ND], fb[ND];
#pragma unroll
        for (int d = 0; d < DIM; ++d) { xb[d] = xp[d]; nb_[d] = nrm[d]; }
#pragma unroll
        for (int k = 0; k < ND; ++k) qb[k] = qL[k];
        if (EXTBC) bc_flux_any<DIM>(r.aux, qb, xb, nb_, a.ph, fb);
        else bc_flux<DIM>(r.aux, qb, xb, nb_, a.ph, fb);
#pragma unroll
        for (int k = 0; k < ND; ++k) flux[k] = fb[k];
      } else {
#if (PDES_SKEL & 1)
#pragma unroll
        for (int k = 0; k < ND; ++k) flux[k] = qL[k] + qR[k] * nrm[k % DIM];
#else
        roe_flux<DIM>(qL, qR, nrm, a.ph.gamma, flux);
#endif
      }
      const double w = op.wface[i];
      const int ir = (r.kind == FK_INTERIOR) ? sm.s_nbrperm[r.orient][i] : i;
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        const double wf = w * flux[k];
        sL[fi * FS + i * ND + k] = -wf;
        sR[fi * FS + ir * ND + k] = wf;
      }
    }
  }
  tile_sync<WARPSYNC>();

  // ---- C: store one record per (element, local face): 8*ND*NFN contiguous bytes each, half a warp per record ---
  constexpr int FL = NFN * ND;
  {
    const int half = tid >> 4, hl = tid & 15;
    for (int rec = half; rec < ((PDES_SKEL & 32) ? (int)(sL[0] == 12345.678) : 2 * nf); rec += TB / 16) {
      const int di = sm.s_dst[rec];
      if (di < 0) continue;
      const double* src = ((rec & 1) ? sR : sL) + (rec >> 1) * FS;
      double* dst = a.fluxe + (int64_t)di * FL;
      if (FL % 2 == 0) {       // records are 16-byte aligned: move them as double2
        for (int c2 = hl; c2 < FL / 2; c2 += 16)
          reinterpret_cast<double2*>(dst)[c2] = reinterpret_cast<const double2*>(src)[c2];
      } else {
        for (int c1 = hl; c1 < FL; c1 += 16) dst[c1] = src[c1];
      }
    }
  }
  if (nxt.kind != 255) {
    const char* pq = reinterpret_cast<const char*>(a.q + (int64_t)nxt.elL * EL);
#pragma unroll
    for (int o = 0; o < EL * 8 + 127; o += 128) prefetch_l2(pq + o);
    if (nxt.kind == FK_INTERIOR) {
      const char* pr = reinterpret_cast<const char*>(a.q + (int64_t)nxt.elR * EL);
#pragma unroll
      for (int o = 0; o < EL * 8 + 127; o += 128) prefetch_l2(pr + o);
    }
    prefetch_l2(a.nrm + (ga + tid) * a.nrm_face_stride);
  }
  if (PRELOADED) *carry = nxt;
}

// EXTBC: boundary faces may carry one of the functors of bc_flux_ext (a separate instantiation, launched only when the
// mesh uses one, so that the kernel of the named configurations is unchanged by them)
template <int DIM, int NN, int NFN, int FT, int MINB, bool EXTBC = false, bool PIPE = false>
__global__ void __launch_bounds__((FaceCfg<DIM, NN, NFN, FT>::T), MINB)
k_face_flux(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ FaceArgs a) {
  // one CTA per tile: a persistent loop over tiles was measured (no gain, and under the 80-register cap the loop
  // state spills: 1.134 vs 1.093 ms per RK4 step on C3)
  constexpr int T = FaceCfg<DIM, NN, NFN, FT>::T;
  __shared__ FaceTileSmem<DIM, NN, NFN, FT> sm;
  const int tid = threadIdx.x;
  if (PIPE) pdl_launch_dependents();
  if (a.ctl->stop) return;
  // PIPE: q (and the record slots this tile overwrites) belong to the element chunks of the previous evaluation; one
  // thread waits for them while the others fetch the tile's (static) face records -- face_tile's first barrier joins
  if (PIPE && tid == T - 1) pipe_acquire(a.pipe, const_cast<Ctl*>(a.ctl));
  const int64_t g0 = a.g0 + (int64_t)blockIdx.x * FT;
  const int64_t rem = a.g0 + a.ng - g0;
  const int nf = (int)(rem < FT ? rem : FT);
  const int64_t ga = a.prefetch_ahead > 0 ? g0 + (int64_t)a.prefetch_ahead * FT : -1;
  if (a.prefetch_ahead > 0 && tid == 0 && ga + (int64_t)a.prefetch_ahead * FT < a.g0 + a.ng) {
    const char* pr = reinterpret_cast<const char*>(a.faces + ga + (int64_t)a.prefetch_ahead * FT);
#pragma unroll
    for (int o = 0; o < FT * 16; o += 128) prefetch_l2(pr + o);
  }
  face_tables<DIM, NN, NFN, FT, T>(op, sm, tid, a.tab_dev);
  face_tile<DIM, NN, NFN, FT, T, false, false, EXTBC>(op, a, sm, g0, nf, ga, tid);
  if (PIPE) {
    __syncthreads();
    if (tid == 0) pipe_release(a.pipe, const_cast<Ctl*>(a.ctl));
  }
}

// k_face_flux_p (PDES_FACE_P=1): persistent form of k_face_flux.  One wave of CTAs strides over the tiles; the
// records of a CTA's NEXT tile are fetched at the top of the current tile and carried in registers, so the dependent
// latency at the head of every tile (FaceRec fetch -> barrier -> gathers: 27 % of k_face_flux's stall samples,
// profiles/r1_end_face_flux_c3.txt region [0,159)) is paid once per CTA instead of once per tile, and the same
// records drive the L2 prefetch of the next tile's gathers.
template <int DIM, int NN, int NFN, int FT, int MINB>
__global__ void __launch_bounds__((FaceCfg<DIM, NN, NFN, FT>::T), MINB)
k_face_flux_p(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ FaceArgs a) {
  constexpr int T = FaceCfg<DIM, NN, NFN, FT>::T;
  __shared__ FaceTileSmem<DIM, NN, NFN, FT> sm;
  if (a.ctl->stop) return;
  const int tid = threadIdx.x;
  const int ntiles = (int)((a.ng + FT - 1) / FT);
  int tile = blockIdx.x;
  if (tile >= ntiles) return;
  face_tables<DIM, NN, NFN, FT, T>(op, sm, tid);
  FaceRec carry;
  carry.kind = 255;
  if (tid < FT && (int64_t)tile * FT + tid < a.ng) carry = a.faces[a.g0 + (int64_t)tile * FT + tid];
  const int stride = gridDim.x;
#pragma unroll 1
  for (; tile < ntiles; tile += stride) {
    const int64_t g0 = a.g0 + (int64_t)tile * FT;
    const int64_t rem = a.g0 + a.ng - g0;
    const int nf = (int)(rem < FT ? rem : FT);
    const int64_t ga = tile + stride < ntiles ? g0 + (int64_t)stride * FT : -1;
    face_tile<DIM, NN, NFN, FT, T, false, true>(op, a, sm, g0, nf, ga, tid, &carry);
    __syncthreads();
  }
}

// k_face_flux_w (PDES_FACE_W=1): warp-autonomous form of k_face_flux.  Every warp owns FT = 5 faces (25 variable lanes,
// 30 node lanes) with its own shared-memory tile and only warp-level barriers, so the warps of a CTA drift apart and
// one warp's gathers overlap another's arithmetic (block barriers are ~25 % of k_face_flux's stall samples).
template <int DIM, int NN, int NFN, int FT, int WPC, int MINB>
__global__ void __launch_bounds__(32 * WPC, MINB)
k_face_flux_w(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ FaceArgs a) {
  static_assert(FaceCfg<DIM, NN, NFN, FT>::T == 32, "one warp per tile");
  __shared__ FaceTileSmem<DIM, NN, NFN, FT> sm[WPC];
  if (a.ctl->stop) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t g0 = a.g0 + ((int64_t)blockIdx.x * WPC + warp) * FT;
  const int64_t rem = a.g0 + a.ng - g0;
  if (rem <= 0) return;
  const int nf = (int)(rem < FT ? rem : FT);
  const int64_t ga = a.prefetch_ahead > 0 ? g0 + (int64_t)a.prefetch_ahead * FT : -1;
  face_tables<DIM, NN, NFN, FT, 32>(op, sm[warp], lane);
  face_tile<DIM, NN, NFN, FT, 32, true>(op, a, sm[warp], g0, nf, ga, lane);
}

// ------------------------------------------------------------------------------------------------------
// k_face_flux_tma: the same computation as k_face_flux with the element gathers done by the bulk-copy (TMA)
// engine: persistent CTAs, two shared-memory stages; while tile t is interpolated and its fluxes evaluated, the
// 2*FT element blocks of tile t+gridDim are landing in the other stage (cp.async.bulk + mbarrier complete_tx).
// The variable threads then read their columns from shared memory instead of issuing 2*NN 40-byte-strided global
// loads each (the LSU data pipe was 70 % busy in k_face_flux: profiles/r1_final_face_flux_c3.txt).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int DIM, int NN, int NFN, int FT>
struct FaceTmaCfg {
  static constexpr int ND = DIM + 2, EL = NN * ND, ELB = EL * 8;
  static constexpr bool ALIGNED = (ELB % 16) == 0;             // element blocks 16-byte aligned for every element
  static constexpr int CPY = ALIGNED ? ELB : ELB + 8;          // bytes per bulk copy (from the 16-byte floor)
  static constexpr int SLOTD = CPY / 8 + 2;                    // doubles per staged element (16-byte multiple)
  static constexpr int PER = ND > NFN ? ND : NFN;
  static constexpr int T = ((FT * PER + 31) / 32) * 32;
  static constexpr int FS = pad_stride(NFN * ND, ND);
  static constexpr size_t static_smem = sizeof(double) * (4 * FT * SLOTD + 2 * FT * FS) + 32 * FT + 1024;
  static constexpr bool FITS = static_smem <= 47 * 1024;       // static shared memory limit
};

template <int DIM, int NN, int NFN, int FT, int MINB>
__global__ void __launch_bounds__((FaceTmaCfg<DIM, NN, NFN, FT>::T), MINB)
k_face_flux_tma(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ FaceArgs a) {
  using Cfg = FaceTmaCfg<DIM, NN, NFN, FT>;
  constexpr int ND = Cfg::ND, T = Cfg::T, FS = Cfg::FS, NF = DIM + 1, SLOTD = Cfg::SLOTD;
  constexpr int FL = NFN * ND;
  __shared__ __align__(16) double sQ[2][2 * FT * SLOTD];
  __shared__ double sL[FT * FS];
  __shared__ double sR[FT * FS];
  __shared__ FaceRec sRec[2][FT];
  __shared__ int s_dst[2 * FT];
  __shared__ int s_perm[NF][NN];
  __shared__ int s_nbrperm[OpTab<DIM, NN, NFN>::NOR][NFN];
  __shared__ __align__(8) unsigned long long bar[2];

  if (a.ctl->stop) return;
  const int tid = threadIdx.x;
  const int64_t ntiles = (a.ng + FT - 1) / FT;
  for (int idx = tid; idx < NF * NN; idx += T) s_perm[idx / NN][idx % NN] = op.perm[idx / NN][idx % NN];
  for (int idx = tid; idx < OpTab<DIM, NN, NFN>::NOR * NFN; idx += T)
    s_nbrperm[idx / NFN][idx % NFN] = op.nbrperm[idx / NFN][idx % NFN];
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  auto tile_nf = [&](int64_t tile) {
    const int64_t rem = a.ng - tile * FT;
    return (int)(rem < FT ? rem : FT);
  };
  auto load_recs = [&](int64_t tile, int st) {
    if (tid < tile_nf(tile)) sRec[st][tid] = a.faces[a.g0 + tile * FT + tid];
  };
  // warp 0: one bulk copy per staged element (the 16-byte aligned block that contains it)
  auto issue = [&](int64_t tile, int st) {
    const int nf = tile_nf(tile);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    unsigned total = 0;
    for (int slot = tid; slot < 2 * FT; slot += 32) {
      int el = -1;
      if ((slot >> 1) < nf) {
        const FaceRec r = sRec[st][slot >> 1];
        el = (slot & 1) ? (r.kind == FK_INTERIOR ? r.elR : -1) : r.elL;
      }
      if (el >= 0) {
        const char* src = reinterpret_cast<const char*>(a.q) + (int64_t)el * Cfg::ELB;
        src -= (reinterpret_cast<uintptr_t>(src) & 15);
        bulk_g2s(&sQ[st][slot * SLOTD], src, Cfg::CPY, &bar[st]);
        total += Cfg::CPY;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (tid == 0) mbar_expect_tx(&bar[st], total);
  };

  int64_t tile = blockIdx.x;
  if (tile >= ntiles) return;
  load_recs(tile, 0);
  __syncthreads();
  if (tid < 32) issue(tile, 0);
  int st = 0;
  unsigned parity[2] = {0, 0};
  for (; tile < ntiles; tile += gridDim.x) {
    const int64_t next = tile + gridDim.x;
    const int nf = tile_nf(tile);
    const int64_t g0 = a.g0 + tile * FT;
    if (next < ntiles) load_recs(next, st ^ 1);
    if (tid < nf) {
      const FaceRec r = sRec[st][tid];
      s_dst[2 * tid] = r.elL * NF + r.fL;
      s_dst[2 * tid + 1] = r.kind == FK_INTERIOR ? r.elR * NF + r.fR : -1;
    }
    __syncthreads();       // records of the next tile visible; the previous tile's stores have read sL / sR
    if (tid < 32 && next < ntiles) issue(next, st ^ 1);
    mbar_wait(&bar[st], parity[st]);
    parity[st] ^= 1;

    // ---- A: interpolate both sides to the face nodes (variable threads), columns read from the staged elements
    if (tid < nf * ND) {
      const int fi = tid / ND, k = tid - fi * ND;
      const FaceRec r = sRec[st][fi];
      double ql[NN], qr[NN];
      {
        const int off = Cfg::ALIGNED ? 0 : (r.elL & 1);
        const double* b = &sQ[st][(2 * fi) * SLOTD + off + k];
#pragma unroll
        for (int j = 0; j < NN; ++j) ql[j] = b[s_perm[r.fL][j] * ND];
      }
      if (r.kind == FK_INTERIOR) {
        const int off = Cfg::ALIGNED ? 0 : (r.elR & 1);
        const double* b = &sQ[st][(2 * fi + 1) * SLOTD + off + k];
#pragma unroll
        for (int j = 0; j < NN; ++j) qr[j] = b[s_perm[r.fR][j] * ND];
      } else {
#pragma unroll
        for (int j = 0; j < NN; ++j) qr[j] = 0.0;
      }
#pragma unroll
      for (int i = 0; i < NFN; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NN; ++j) s = fma(op.interp[j][i], ql[j], s);
        sL[fi * FS + i * ND + k] = s;
      }
      if (r.kind == FK_INTERIOR) {
#pragma unroll
        for (int i = 0; i < NFN; ++i) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < NN; ++j) s = fma(op.interp[j][i], qr[j], s);
          sR[fi * FS + s_nbrperm[r.orient][i] * ND + k] = s;
        }
      } else if (r.kind == FK_SHARED) {
        const double* b = a.q_recv + (int64_t)r.aux * (NFN * ND) + k;
#pragma unroll
        for (int i = 0; i < NFN; ++i) sR[fi * FS + s_nbrperm[r.orient][i] * ND + k] = b[i * ND];
      }
    }
    __syncthreads();

    // ---- B: numerical flux at every face node (node threads) ---------------------------------------------
    {
      const bool nact = tid < nf * NFN;
      const int fi = nact ? tid / NFN : 0, i = tid - fi * NFN;
      const FaceRec r = sRec[st][fi];
      const int64_t g = g0 + fi;
      double nrm[DIM], qL[ND], qR[ND], flux[ND];
      if (nact) {
        const double* np_ = a.nrm + g * a.nrm_face_stride + i * a.nrm_node_stride;
#pragma unroll
        for (int d = 0; d < DIM; ++d) nrm[d] = __ldg(np_ + d);
#pragma unroll
        for (int k = 0; k < ND; ++k) { qL[k] = sL[fi * FS + i * ND + k]; qR[k] = sR[fi * FS + i * ND + k]; }
      }
      __syncthreads();
      if (nact) {
        if (r.kind == FK_BOUNDARY) {
          const double* xp = a.coords_bndry + ((int64_t)r.elR * NFN + i) * DIM;
          double xb[DIM], nb_[DIM], qb[ND], fb[ND];
#pragma unroll
          for (int d = 0; d < DIM; ++d) { xb[d] = xp[d]; nb_[d] = nrm[d]; }
#pragma unroll
          for (int k = 0; k < ND; ++k) qb[k] = qL[k];
          bc_flux<DIM>(r.aux, qb, xb, nb_, a.ph, fb);
#pragma unroll
          for (int k = 0; k < ND; ++k) flux[k] = fb[k];
        } else {
          roe_flux<DIM>(qL, qR, nrm, a.ph.gamma, flux);
        }
        const double w = op.wface[i];
        const int ir = (r.kind == FK_INTERIOR) ? s_nbrperm[r.orient][i] : i;
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          const double wf = w * flux[k];
          sL[fi * FS + i * ND + k] = -wf;
          sR[fi * FS + ir * ND + k] = wf;
        }
      }
    }
    __syncthreads();

    // ---- C: store one record per (element, local face), half a warp per record ----------------------------
    {
      const int half = tid >> 4, hl = tid & 15;
      for (int rec = half; rec < 2 * nf; rec += T / 16) {
        const int di = s_dst[rec];
        if (di < 0) continue;
        const double* src = ((rec & 1) ? sR : sL) + (rec >> 1) * FS;
        double* dst = a.fluxe + (int64_t)di * FL;
        for (int c1 = hl; c1 < FL; c1 += 16) dst[c1] = src[c1];
      }
    }
    st ^= 1;
  }
}

// Coalesced epilogue shared by the element kernels: the tile of staged values (EPI_RES: acc; EPI_RK: Minv*acc) is
// combined with the tabulated source and either stored as res or pushed through the fused RK4 stage
// (rk4.jl:244-319); two dofs per access, CH accesses in flight per thread.
// STAGED: the source / x_old / ksum tiles were copied to shared memory (sStr = [srcm | x_old | ksum], E*EL each).
// WARP: the tile belongs to one warp (T = 32, tid = lane): no block barrier, the warp writes its own norm partial.
template <int NN, int ND, int E, int T, int MODE, bool STAGED = false, bool WARP = false, bool MINV = false>
__device__ __forceinline__ void epilogue_tile(const ElemArgs& a, const double* sq, int ne, int64_t e0, int tid,
                                              double* s_red, const double* sStr = nullptr) {
  constexpr int EL = NN * ND;
  double nrm2 = 0.0;
  {
    constexpr int CH = 3;
    const int ntile = ne * EL;
    const int npair = (ntile + 1) / 2;
    const int64_t base = e0 * EL;            // even: 16-byte aligned in every array
    const double2* sq2 = reinterpret_cast<const double2*>(sq);
    const double* psrc = MODE == EPI_RES ? a.srcw : a.srcm;
    for (int i0 = 0; i0 < npair; i0 += CH * T) {
      double2 v[CH], sv[CH], xo[CH], ks[CH], mw[CH];
      bool ok[CH], two[CH];
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int i2 = i0 + u * T + tid;
        ok[u] = i2 < npair;
        two[u] = ok[u] && (2 * i2 + 1 < ntile);
        const int64_t dof = base + 2 * i2;
        sv[u] = xo[u] = ks[u] = mw[u] = make_double2(0.0, 0.0);
        double2 mi = make_double2(1.0, 1.0);
        if (MINV && ok[u]) {
          // pde_post_func (res_vec *= Minv) applied here, to the staged raw rows: the loads join the batch instead of
          // standing between the face products and the staging stores (4.7 % of the kernel's stall samples)
          const double* it = a.minv + e0 * NN;
          mi.x = __ldg(it + (2 * i2) / ND);
          if (two[u]) mi.y = __ldg(it + (2 * i2 + 1) / ND);
        }
        if (!(PDES_OPT & 16) && MODE == EPI_RK && a.stage == 1 && ok[u]) {
          // calcNorm weights M = w_j/jac_j of the two dofs (the tile's nodes are contiguous in M: node = dof / ND);
          // requested with the other loads of the batch: a load next to its use cost 25 us per step (stall profile)
          const double* mt = a.mass + e0 * NN;
          mw[u].x = __ldg(mt + (2 * i2) / ND);
          if (two[u]) mw[u].y = __ldg(mt + (2 * i2 + 1) / ND);
        }
        if (STAGED && ok[u]) {
          const double2* s2 = reinterpret_cast<const double2*>(sStr);
          // scheme 2 (rk4 without the running sum): slot 1 = x_old (stage 4: q4), slot 2 = q2 (stage 2) | w' (stage 4)
          const bool has_src = psrc && !(a.scheme == 2 && a.stage == 4);
          const bool has_ks = a.scheme == 2 ? (a.stage == 2 || a.stage == 4) : a.stage > 1;
          if (two[u]) {
            v[u] = sq2[i2];
            if (has_src) sv[u] = s2[i2];
            xo[u] = s2[(E * EL) / 2 + i2];
            if (has_ks) ks[u] = s2[E * EL + i2];
          } else {
            v[u] = make_double2(sq[2 * i2], 0.0);
            if (has_src) sv[u].x = sStr[2 * i2];
            xo[u].x = sStr[E * EL + 2 * i2];
            if (has_ks) ks[u].x = sStr[2 * E * EL + 2 * i2];
          }
        } else if (two[u]) {
          v[u] = sq2[i2];
          if (psrc) sv[u] = __ldg(reinterpret_cast<const double2*>(psrc + dof));
          if (MODE == EPI_RK) {
            xo[u] = __ldg(reinterpret_cast<const double2*>(a.x_old + dof));
            if (a.stage > 1) ks[u] = *reinterpret_cast<const double2*>(a.ksum + dof);
          }
        } else if (ok[u]) {
          v[u] = make_double2(sq[2 * i2], 0.0);
          if (psrc) sv[u].x = __ldg(psrc + dof);
          if (MODE == EPI_RK) {
            xo[u].x = __ldg(a.x_old + dof);
            if (a.stage > 1) ks[u].x = a.ksum[dof];
          }
        }
        if (MINV && ok[u]) { v[u].x *= mi.x; v[u].y *= mi.y; }
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        if (!ok[u]) continue;
        const int i2 = i0 + u * T + tid;
        const int64_t dof = base + 2 * i2;
        const double2 k = make_double2(v[u].x + sv[u].x, v[u].y + sv[u].y);
        double2 o1 = k, o2 = k;
        if (MODE == EPI_RK && a.stage == 1) {
          // calcNorm: sum res*M*res (Utils.jl:427-449); the tile's nodes are contiguous in M: node = dof / ND
          // (a multiplication by the stored M as in the reference: an FP64 division per dof cost 25 us per step)
          if (PDES_OPT & 16) {
            const double* mt = a.mass + e0 * NN;
            mw[u].x = __ldg(mt + (2 * i2) / ND);
            if (two[u]) mw[u].y = __ldg(mt + (2 * i2 + 1) / ND);
          }
          nrm2 = fma(k.x * mw[u].x, k.x, nrm2);
          if (two[u]) nrm2 = fma(k.y * mw[u].y, k.y, nrm2);
        }
        if (MODE == EPI_RK && a.scheme == 1) {
          // lserk54: dq = a_s*dq + delta_t*res ; q += b_s*dq
          if (a.stage == 1) o1 = make_double2(a.hh * k.x, a.hh * k.y);
          else o1 = make_double2(a.ah * ks[u].x + a.hh * k.x, a.ah * ks[u].y + a.hh * k.y);
          o2 = make_double2(xo[u].x + a.h6 * o1.x, xo[u].y + a.h6 * o1.y);
        } else if (MODE == EPI_RK && a.scheme == 2) {
          // classical RK4 without the running sum k1 + 2 k2 + 2 k3 (4 of the 17 state-vector passes of a step):
          //   q2 = x + h/2 k1, q3 = x + h/2 k2, q4 = x + h k3  =>  h/6 (k1 + 2 k2 + 2 k3) = (q2 + 2 q3 + q4)/3 - 4x/3
          //   stage 2 stores  w' = q2 + 2 q3 - x + (h/2) srcm        (the only extra vector)
          //   stage 4         x_new = (w' + q4)/3 + (h/6) Minv R(q4)  [(h/6) srcm of k4 is the srcm term of w']
          // identical in exact arithmetic to rk4.jl:244-319; the rounding differs by a few ulp of |q| per step
          if (a.stage == 4) {
            const double third = 1.0 / 3.0;
            o2 = make_double2(fma(a.h6, v[u].x, (ks[u].x + xo[u].x) * third), fma(a.h6, v[u].y, (ks[u].y + xo[u].y) * third));
          } else {
            o2 = make_double2(xo[u].x + a.ah * k.x, xo[u].y + a.ah * k.y);
            if (a.stage == 2)
              o1 = make_double2(ks[u].x + 2.0 * o2.x - xo[u].x + a.ah * sv[u].x, ks[u].y + 2.0 * o2.y - xo[u].y + a.ah * sv[u].y);
          }
        } else if (MODE == EPI_RK) {
          if (a.stage == 1) {
            o1 = k;                                                                   // ksum
            o2 = make_double2(xo[u].x + a.ah * k.x, xo[u].y + a.ah * k.y);            // q_next
          } else if (a.stage < 4) {
            o1 = make_double2(ks[u].x + 2.0 * k.x, ks[u].y + 2.0 * k.y);
            o2 = make_double2(xo[u].x + a.ah * k.x, xo[u].y + a.ah * k.y);
          } else {
            o2 = make_double2(xo[u].x + a.h6 * (ks[u].x + k.x), xo[u].y + a.h6 * (ks[u].y + k.y));
          }
        }
        if (MODE == EPI_RES) {
          if (two[u]) *reinterpret_cast<double2*>(a.res + dof) = k; else a.res[dof] = k.x;
        } else {
          if (a.scheme == 1 || (a.scheme == 0 && a.stage < 4) || (a.scheme == 2 && a.stage == 2)) {
            if (two[u]) __stcs(reinterpret_cast<double2*>(a.ksum + dof), o1); else __stcs(a.ksum + dof, o1.x);
          }
          if (two[u]) *reinterpret_cast<double2*>(a.q_next + dof) = o2; else a.q_next[dof] = o2.x;
        }
      }
    }
  }
  if (MODE == EPI_RK && a.stage == 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
    if (WARP) {
      if (tid == 0) a.norm_partials[e0 / E] = nrm2;
    } else {
      if ((tid & 31) == 0) s_red[tid >> 5] = nrm2;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < T / 32; ++w) t += s_red[w];
        a.norm_partials[e0 / E] = t;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// k_element_rk: E elements per CTA; every variable thread owns TWO (element, variable) rows -- elements p and
// p + E/2 -- so that each uniform coefficient load feeds two DFMAs
// ------------------------------------------------------------------------------------------------------
template <int DIM, int NN, int NFN, int E>
struct TileCfg {
  static constexpr int ND = DIM + 2, NF = DIM + 1;
  static constexpr int HP = E / 2;                              // row pairs
  static constexpr int VT = HP * ND;                            // variable threads
  static constexpr int T = ((VT + 31) / 32) * 32;
  static constexpr int SQ = NN * ND;                            // per-element stride of the q tile (contiguous: cp.async)
  static constexpr int SF = ND * DIM * NN;                      // per-element stride of the volume-flux tile
  static constexpr int SU = SF > 3 * SQ ? SF : 3 * SQ;         // the tile is reused for the three epilogue streams
  static constexpr size_t smem_bytes = sizeof(double) * (size_t)E * (SQ + SU);
  static_assert(E % 2 == 0, "tile bases must stay 16-byte aligned");
};

// One tile of ne <= E elements starting at element e0, executed by a CTA of TB threads.  ea = first element of the
// tile whose inputs are prefetched into L2 (or < 0); na its size.  COHERENT: the face records were written by other
// CTAs of the SAME launch (fused kernel): plain loads instead of the read-only path, and the consumed records are
// dropped from L2 without a write-back (discard.global.L2), since nothing reads them again.
template <int DIM, int NN, int NFN, int E, int MODE, int TB, bool COHERENT, bool MMA = false>
__device__ __forceinline__ void element_tile(const OpTab<DIM, NN, NFN>& op, const ElemArgs& a, unsigned char* smem_raw,
                                             double* s_red, int64_t e0, int ne, int64_t ea, int na, int tid) {
  using Cfg = TileCfg<DIM, NN, NFN, E>;
  constexpr int ND = Cfg::ND, NF = Cfg::NF, T = TB, SQ = Cfg::SQ, HP = Cfg::HP;
  static_assert(TB >= Cfg::T, "block too small for the element tile");
  constexpr int EL = NN * ND;                       // doubles per element
  constexpr int FL = NFN * ND;                      // doubles per face
  double* sq = reinterpret_cast<double*>(smem_raw);             // [E][EL]
  double* sF = sq + E * SQ;                                     // [E][ND][DIM][NN]
  auto ldrec = [](const double* p) { return (PDES_SKEL & 2) ? 1.0 : (COHERENT ? *p : __ldg(p)); };

  const double gami = a.ph.gamma - 1.0;

  // ---- S0: q tile (asynchronous copy) + L2 prefetch of the streams of the tile this SM slot runs next --------
  async_tile(sq, a.q + e0 * EL, ne * EL, tid, T);
  cp_async_commit();
  {
    if (ea >= 0) {
      const int64_t b0 = ea * EL * 8, nb = (int64_t)na * EL * 8;
      for (int64_t o = (int64_t)tid * 128; o < nb; o += (int64_t)T * 128) {
        prefetch_l2(reinterpret_cast<const char*>(a.q) + b0 + o);
        if (MODE == EPI_RES && a.srcw) prefetch_l2(reinterpret_cast<const char*>(a.srcw) + b0 + o);
        if (MODE == EPI_RK) {
          const bool s2 = a.scheme == 2;
          if (a.srcm && !(s2 && a.stage == 4)) prefetch_l2(reinterpret_cast<const char*>(a.srcm) + b0 + o);
          if (!(s2 && a.stage == 4)) prefetch_l2(reinterpret_cast<const char*>(a.x_old) + b0 + o);
          if (s2 ? a.stage == 4 : a.stage > 1) prefetch_l2(reinterpret_cast<const char*>(a.ksum) + b0 + o);
        }
      }
      if (!COHERENT)     // (fused kernel: the records are written into L2 shortly before they are consumed)
        for (int64_t o = (int64_t)tid * 128; o < (int64_t)na * NF * FL * 8; o += (int64_t)T * 128)
          prefetch_l2(reinterpret_cast<const char*>(a.fluxe) + ea * NF * FL * 8 + o);
      for (int64_t o = (int64_t)tid * 128; o < (int64_t)na * a.dx_el_stride * 8; o += (int64_t)T * 128)
        prefetch_l2(reinterpret_cast<const char*>(a.dxidx) + ea * a.dx_el_stride * 8 + o);
    }
  }
  // this tile's own epilogue streams and face contributions: in L2 by the time S3 / S4 ask for them
  {
    const int64_t b0 = e0 * EL * 8, nb = (int64_t)ne * EL * 8;
    for (int64_t o = (int64_t)tid * 128; o < nb; o += (int64_t)T * 128) {
      if (MODE == EPI_RES && a.srcw) prefetch_l2(reinterpret_cast<const char*>(a.srcw) + b0 + o);
      if (MODE == EPI_RK) {
        const bool s2 = a.scheme == 2;
        if (a.srcm && !(s2 && a.stage == 4)) prefetch_l2(reinterpret_cast<const char*>(a.srcm) + b0 + o);
        if (!(s2 && a.stage == 4)) prefetch_l2(reinterpret_cast<const char*>(a.x_old) + b0 + o);
        if (s2 ? a.stage == 4 : a.stage > 1) prefetch_l2(reinterpret_cast<const char*>(a.ksum) + b0 + o);
      }
    }
    if (!COHERENT)
      for (int64_t o = (int64_t)tid * 128; o < (int64_t)ne * NF * FL * 8; o += (int64_t)T * 128)
        prefetch_l2(reinterpret_cast<const char*>(a.fluxe) + e0 * NF * FL * 8 + o);
    if ((PDES_OPT & 64) && MODE == EPI_RK)
      for (int64_t o = (int64_t)tid * 128; o < (int64_t)ne * NN * 8; o += (int64_t)T * 128)
        prefetch_l2(reinterpret_cast<const char*>(a.minv + e0 * NN) + o);
    if (!(PDES_OPT & 32) && MODE == EPI_RK && a.stage == 1)
      for (int64_t o = (int64_t)tid * 128; o < (int64_t)ne * NN * 8; o += (int64_t)T * 128)
        prefetch_l2(reinterpret_cast<const char*>(a.mass + e0 * NN) + o);
  }
  // the metrics of this thread's first node are requested before the wait for the q tile, those of the next node
  // before the current one is processed (they are L2 hits after the prefetch, but still ~1000 cycles away)
  double dxn[DIM * DIM];
  if ((PDES_OPT & 1) && tid < ne * NN) {
    const int s = tid / NN, j = tid - s * NN;
    const double* dx = a.dxidx + (e0 + s) * a.dx_el_stride + j * a.dx_node_stride;
#pragma unroll
    for (int m = 0; m < DIM * DIM; ++m) dxn[m] = __ldg(dx + m);
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- S1: Euler flux in the parametric directions at every node (getEulerFlux) ---------------
  for (int it = tid; it < ne * NN; it += T) {
    const int s = it / NN, j = it - s * NN;
    double qn[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) qn[k] = sq[s * SQ + j * ND + k];
    double dxl[DIM * DIM];
    if (!(PDES_OPT & 1)) {
      const double* dx = a.dxidx + (e0 + s) * a.dx_el_stride + j * a.dx_node_stride;
#pragma unroll
      for (int m = 0; m < DIM * DIM; ++m) dxn[m] = __ldg(dx + m);
    }
#pragma unroll
    for (int m = 0; m < DIM * DIM; ++m) dxl[m] = dxn[m];
    if ((PDES_OPT & 1) && it + T < ne * NN) {
      const int s2 = (it + T) / NN, j2 = (it + T) - s2 * NN;
      const double* dx = a.dxidx + (e0 + s2) * a.dx_el_stride + j2 * a.dx_node_stride;
#pragma unroll
      for (int m = 0; m < DIM * DIM; ++m) dxn[m] = __ldg(dx + m);
    }
    const double press = calc_pressure<DIM>(qn, gami);
    if (!PDES_SKEL && ((a.ph.check_density && !(qn[0] > 0.0)) || (a.ph.check_pressure && !(press > 0.0)))) {
      const int code = (a.ph.check_density && !(qn[0] > 0.0)) ? 1 : 2;
      const unsigned long long loc = ((unsigned long long)(e0 + s) << 8) | (unsigned)j;
      // density errors win over pressure errors (checkDensity runs first), lowest location wins
      atomicMin(&a.ctl->err_loc, ((unsigned long long)(code - 1) << 62) | loc);
      atomicExch(&a.ctl->err_code, 1);
      atomicExch(&a.ctl->stop, 1);
    }
    const double rinv = 1.0 / qn[0];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double U = 0.0;
#pragma unroll
      for (int p = 0; p < DIM; ++p) U += qn[1 + p] * dxl[d + DIM * p];
      U *= rinv;
      double* Fo = sF + ((s * ND) * DIM + d) * NN + j;
      Fo[0] = qn[0] * U;
#pragma unroll
      for (int p = 0; p < DIM; ++p) Fo[(1 + p) * DIM * NN] = qn[1 + p] * U + dxl[d + DIM * p] * press;
      Fo[(DIM + 1) * DIM * NN] = (qn[DIM + 1] + press) * U;
    }
  }
  __syncthreads();       // sF complete; nobody reads the q tile any more (it becomes the output staging tile)
  if (MODE == EPI_RK && (PDES_OPT & 4)) {
    // Minv[node] replicated over the ND slots of the node in the dead q tile: the owner of a row later reads and
    // then overwrites exactly its own slots, so the multiply by Minv costs LDS instead of exposed global loads
    const double* mb = a.minv + e0 * NN;
    for (int idx = tid; idx < ne * EL; idx += T) cp_async8(sq + idx, mb + idx / ND);
    cp_async_commit();
  }

  constexpr bool STAGED = (MODE == EPI_RK);
  if constexpr (MMA) {
    // ---- S2 + S3 on the FP64 tensor-core path (mma.sync.m8n8k4.f64): both operator products are small dense GEMMs
    //   out[(s,k), i] = sum_(d,j) F[(s,k), (d,j)] Qt[(d,j), i]  +  sum_(f,n) rec[(s,k), (f,n)] RfN[(f,n), i]
    // with M = E*ND rows (8 per tile), N = NN (two 8-column tiles), K = DIM*NN (+ padding) and NF*NFN.  As DFMA they are
    // 39 % of this kernel's warp instructions (one LDCU per two DFMA); one DMMA does the work of eight DFMA issues.
    // A fragments come straight from the volume-flux tile / the face records, B fragments (the operator) are held in
    // registers, the accumulators go to the staging tile in fragment layout.
    constexpr int K2 = DIM * NN, KS2 = (K2 + 3) / 4, K3 = NF * NFN, KS3 = (K3 + 3) / 4, NT = (NN + 7) / 8;
    constexpr int MT = (E * ND + 7) / 8, WARPS = T / 32, MPW = (MT + WARPS - 1) / WARPS;
    const int warp = tid >> 5, lane = tid & 31, gr = lane >> 2, gc = lane & 3;
    const double* tabQ = a.s2_dev;                 // device copy of Qt [K2][NN] | RfN [K3][NN]
    const double* tabR = a.s2_dev + K2 * NN;
    double cfr[MPW][NT][2];
#pragma unroll
    for (int m = 0; m < MPW; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n) { cfr[m][n][0] = 0.0; cfr[m][n][1] = 0.0; }
    {
      double bq[KS2][NT];
#pragma unroll
      for (int t = 0; t < KS2; ++t)
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const int kr = 4 * t + gc, col = 8 * n + gr;
          bq[t][n] = (kr < K2 && col < NN) ? __ldg(tabQ + kr * NN + col) : 0.0;
        }
#pragma unroll
      for (int m = 0; m < MPW; ++m) {
        const int mt = warp + m * WARPS;
        if (mt < MT) {
          const int r = mt * 8 + gr;
          const double* Arow = sF + (r < E * ND ? r : 0) * K2;
#pragma unroll
          for (int t = 0; t < KS2; ++t) {
            const int kc = 4 * t + gc;
            const double av = kc < K2 ? Arow[kc] : 0.0;
#pragma unroll
            for (int n = 0; n < NT; ++n)
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(cfr[m][n][0]), "+d"(cfr[m][n][1]) : "d"(av), "d"(bq[t][n]));
          }
        }
      }
    }
    if (STAGED) {
      // the volume-flux tile is dead: its storage receives the epilogue's streams (srcm | x_old | ksum), which are
      // in flight while the face products run.  The Minv tile (requested before S2) has landed.
      cp_async_wait<0>();
      __syncthreads();
      const unsigned long long pol = policy_evict_first();
      if (!(PDES_SKEL & 64)) {
      if (a.scheme == 2) {
        // rk4 without the running sum (see epilogue_tile): stage 2 re-reads its own input rows (q2: L2 hits, the tile was
        // loaded by this CTA microseconds ago), stage 4 reads w' and its own input rows (q4) and neither x_old nor srcm
        if (a.stage < 4) {
          if (a.srcm) async_tile_stream(sF, a.srcm + e0 * EL, ne * EL, tid, T, pol);
          if (a.stage == 1) async_tile(sF + E * EL, a.x_old + e0 * EL, ne * EL, tid, T);
          else async_tile_stream(sF + E * EL, a.x_old + e0 * EL, ne * EL, tid, T, pol);
          if (a.stage == 2) async_tile(sF + 2 * E * EL, a.q + e0 * EL, ne * EL, tid, T);
        } else {
          async_tile(sF + E * EL, a.q + e0 * EL, ne * EL, tid, T);
          async_tile_stream(sF + 2 * E * EL, a.ksum + e0 * EL, ne * EL, tid, T, pol);
        }
      } else {
      if (a.srcm) async_tile_stream(sF, a.srcm + e0 * EL, ne * EL, tid, T, pol);
      if (a.stage == 1) async_tile(sF + E * EL, a.x_old + e0 * EL, ne * EL, tid, T);   // x_old == q of this stage
      else async_tile_stream(sF + E * EL, a.x_old + e0 * EL, ne * EL, tid, T, pol);
      if (a.stage > 1) async_tile_stream(sF + 2 * E * EL, a.ksum + e0 * EL, ne * EL, tid, T, pol);
      }
      }
      cp_async_commit();
    }
    {
      double br[KS3][NT];
#pragma unroll
      for (int t = 0; t < KS3; ++t)
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const int kr = 4 * t + gc, col = 8 * n + gr;
          br[t][n] = (kr < K3 && col < NN) ? __ldg(tabR + kr * NN + col) : 0.0;
        }
#pragma unroll
      for (int m = 0; m < MPW; ++m) {
        const int mt = warp + m * WARPS;
        if (mt < MT) {
          const int r = mt * 8 + gr;
          const int sr = r / ND, kr_ = r - sr * ND;
          const bool rok = sr < ne;
          const double* G = a.fluxe + (e0 + (rok ? sr : 0)) * (NF * FL) + kr_;
          double av[KS3];
#pragma unroll
          for (int t = 0; t < KS3; ++t) {
            const int kc = 4 * t + gc;
            av[t] = (rok && kc < K3) ? ldrec(G + kc * ND) : 0.0;
          }
#pragma unroll
          for (int t = 0; t < KS3; ++t)
#pragma unroll
            for (int n = 0; n < NT; ++n)
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(cfr[m][n][0]), "+d"(cfr[m][n][1]) : "d"(av[t]), "d"(br[t][n]));
          // accumulator fragment: row gr of the tile, columns 2 gc, 2 gc + 1 of each 8-column tile -> staging tile
          if (rok) {
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int i = 8 * n + 2 * gc + h;
                if (i < NN) {
                  double v = cfr[m][n][h];
                  if (MODE == EPI_RK) v *= __ldg(a.minv + (e0 + sr) * NN + i);      // pde_post_func: res_vec *= Minv
                  sq[sr * SQ + i * ND + kr_] = v;
                }
              }
          }
        }
      }
    }
  } else {
  // ---- S2 + S3: variable threads -------------------------------------------------------------------
  const int vp = tid / ND, vk = tid - vp * ND;
  const int s0 = vp, s1 = vp + HP;
  const bool act0 = tid < Cfg::VT && s0 < ne, act1 = tid < Cfg::VT && s1 < ne;
  const int s1c = act1 ? s1 : s0;          // the second row of a ragged tile recomputes the first (never stored)
  const int s0c = act0 ? s0 : 0;
  double acc0[NN], acc1[NN];
#pragma unroll
  for (int u = 0; u < NN; ++u) { acc0[u] = 0.0; acc1[u] = 0.0; }
  // face contributions of the two rows: signed, in the element's node order (k_face_flux).  The records of face
  // f+1 are requested before the products of face f are issued (software pipeline; they are L2 hits).
  const double* G0 = a.fluxe + (e0 + s0c) * (NF * FL) + vk;
  const double* G1 = a.fluxe + (e0 + (act0 ? s1c : 0)) * (NF * FL) + vk;
  // records of faces 0 and 1 are requested before the volume products, face f+2 before the products of face f
  double g0v[NFN], g1v[NFN], h0v[NFN], h1v[NFN];
  if (act0) {
#pragma unroll
    for (int i = 0; i < NFN; ++i) { g0v[i] = ldrec(G0 + i * ND); g1v[i] = ldrec(G1 + i * ND); }
    if (NF > 1 && (PDES_OPT & 2)) {
#pragma unroll
      for (int i = 0; i < NFN; ++i) { h0v[i] = ldrec(G0 + (NFN + i) * ND); h1v[i] = ldrec(G1 + (NFN + i) * ND); }
    }
    // S2: volume integral  res[k,i] = sum_d sum_j Q[j,i,d] F_d[k,j]   (weakdifferentiate!, trans=true)
    // (the loop over directions stays rolled: fully unrolled operator products overflow the instruction cache)
    const double* F0 = sF + (s0 * ND + vk) * DIM * NN;
    const double* F1 = sF + (s1c * ND + vk) * DIM * NN;
#pragma unroll 1
    for (int d = 0; d < DIM; ++d) {
#pragma unroll
      for (int j = 0; j < NN; ++j) {
        const double f0 = (PDES_SKEL & 4) ? 1.0 : F0[d * NN + j], f1 = (PDES_SKEL & 4) ? 1.0 : F1[d * NN + j];
#if (PDES_SKEL & 1)
        acc0[j] += f0; acc1[j] += f1;
#else
#pragma unroll
        for (int u = 0; u < NN; ++u) {
          const double c = op.Qt[d * NN + j][u];
          acc0[u] = fma(c, f0, acc0[u]);
          acc1[u] = fma(c, f1, acc1[u]);
        }
#endif
      }
    }
  }
  if (STAGED) {
    // the volume-flux tile is dead: its storage receives the epilogue's streams (srcm | x_old | ksum), which are
    // in flight while the face products run.  The Minv tile (requested before S2) has landed.
    cp_async_wait<0>();
    __syncthreads();
    const unsigned long long pol = policy_evict_first();
    if (!(PDES_SKEL & 64)) {
    if (a.scheme == 2) {
      // rk4 without the running sum (see epilogue_tile): stage 2 re-reads its own input rows (q2: L2 hits, the tile was
      // loaded by this CTA microseconds ago), stage 4 reads w' and its own input rows (q4) and neither x_old nor srcm
      if (a.stage < 4) {
        if (a.srcm) async_tile_stream(sF, a.srcm + e0 * EL, ne * EL, tid, T, pol);
        if (a.stage == 1) async_tile(sF + E * EL, a.x_old + e0 * EL, ne * EL, tid, T);
        else async_tile_stream(sF + E * EL, a.x_old + e0 * EL, ne * EL, tid, T, pol);
        if (a.stage == 2) async_tile(sF + 2 * E * EL, a.q + e0 * EL, ne * EL, tid, T);
      } else {
        async_tile(sF + E * EL, a.q + e0 * EL, ne * EL, tid, T);
        async_tile_stream(sF + 2 * E * EL, a.ksum + e0 * EL, ne * EL, tid, T, pol);
      }
    } else {
    if (a.srcm) async_tile_stream(sF, a.srcm + e0 * EL, ne * EL, tid, T, pol);
    if (a.stage == 1) async_tile(sF + E * EL, a.x_old + e0 * EL, ne * EL, tid, T);   // x_old == q of this stage
    else async_tile_stream(sF + E * EL, a.x_old + e0 * EL, ne * EL, tid, T, pol);
    if (a.stage > 1) async_tile_stream(sF + 2 * E * EL, a.ksum + e0 * EL, ne * EL, tid, T, pol);
    }
    }
    cp_async_commit();
  }
  if (act0) {
    // S3: face integration  res[k,node] += sum_f sum_i Rf[f][i][node] * (-+ w_i f*[k,i])
    // (interiorfaceintegrate!, boundaryintegrate!, boundaryFaceIntegrate!)
#pragma unroll 1
    for (int f = 0; f < NF; ++f) {
      double n0v[NFN], n1v[NFN];
      constexpr int AH = (PDES_OPT & 2) ? 2 : 1;
      if (f + AH < NF) {
#pragma unroll
        for (int i = 0; i < NFN; ++i) {
          n0v[i] = ldrec(G0 + ((f + AH) * NFN + i) * ND);
          n1v[i] = ldrec(G1 + ((f + AH) * NFN + i) * ND);
        }
      }
#if (PDES_SKEL & 1)
#pragma unroll
      for (int i = 0; i < NFN; ++i) { acc0[i] += g0v[i]; acc1[i] += g1v[i]; }
#else
#pragma unroll
      for (int i = 0; i < NFN; ++i)
#pragma unroll
        for (int u = 0; u < NN; ++u) {
          const double c = op.RfN[f * NFN + i][u];
          acc0[u] = fma(c, g0v[i], acc0[u]);
          acc1[u] = fma(c, g1v[i], acc1[u]);
        }
#endif
#pragma unroll
      for (int i = 0; i < NFN; ++i) {
        if (PDES_OPT & 2) { g0v[i] = h0v[i]; g1v[i] = h1v[i]; h0v[i] = n0v[i]; h1v[i] = n1v[i]; }
        else { g0v[i] = n0v[i]; g1v[i] = n1v[i]; }
      }
    }
    // pde_post_func: res_vec *= Minv (EPI_RK); staged for the coalesced epilogue
    if (MODE == EPI_RK && !(PDES_OPT & 64)) {
#pragma unroll
      for (int u = 0; u < NN; ++u) {
        if (PDES_OPT & 4) {
          acc0[u] *= sq[s0 * SQ + u * ND + vk];
          acc1[u] *= sq[s1c * SQ + u * ND + vk];
        } else {
          acc0[u] *= __ldg(a.minv + (e0 + s0) * NN + u);
          acc1[u] *= __ldg(a.minv + (e0 + s1c) * NN + u);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < NN; ++u) sq[s0 * SQ + u * ND + vk] = acc0[u];
    if (act1) {
#pragma unroll
      for (int u = 0; u < NN; ++u) sq[s1 * SQ + u * ND + vk] = acc1[u];
    }
  }
  }
  if (STAGED) cp_async_wait<0>();
  __syncthreads();
  if (a.discard_records) {
    // every thread has consumed its records: drop the tile's full 128-byte lines (tiles are line-aligned whenever
    // E*NF*FL*8 is a multiple of 128; partial lines at the ends are left alone)
    const uintptr_t b = reinterpret_cast<uintptr_t>(a.fluxe + e0 * (NF * FL));
    const uintptr_t lo = (b + 127) & ~(uintptr_t)127, hi = (b + (uintptr_t)ne * NF * FL * 8) & ~(uintptr_t)127;
    for (uintptr_t p = lo + (uintptr_t)tid * 128; p < hi; p += (uintptr_t)T * 128)
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
  }

  // ---- S4: coalesced epilogue (source, res | fused RK4 stage, stage-1 norm partial) ----------------------------
  epilogue_tile<NN, ND, E, T, MODE, STAGED, false, (MODE == EPI_RK) && (PDES_OPT & 64) != 0>(a, sq, ne, e0, tid, s_red, sF);
}

template <int DIM, int NN, int NFN, int E, int MODE, int MINB, bool PIPE = false, bool MMA = false>
__global__ void __launch_bounds__((TileCfg<DIM, NN, NFN, E>::T), MINB)
k_element_rk(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ ElemArgs a) {
  using Cfg = TileCfg<DIM, NN, NFN, E>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[Cfg::T / 32];
  const int tid = threadIdx.x;
  if (PIPE) pdl_launch_dependents();
  if (a.ctl->stop) return;
  if (PIPE) {
    // the face chunks 0..chunk of THIS evaluation (which in turn waited for the previous evaluation's element chunks)
    if (tid == 0) pipe_acquire(a.pipe, a.ctl);
    __syncthreads();
  }
  // reverse: CTAs are dispatched in blockIdx order, so the sweep starts with the elements whose face records (and q
  // gathers) k_face_flux touched LAST and which are therefore still in L2; "ahead" then means lower tiles
  const int64_t bid = a.reverse ? (int64_t)(gridDim.x - 1 - blockIdx.x) : (int64_t)blockIdx.x;
  const int64_t e0 = a.e_begin + bid * E;
  const int ne = (int)((a.nE - e0) < E ? (a.nE - e0) : E);
  int64_t ea = a.reverse ? e0 - (int64_t)a.prefetch_ahead * E : e0 + (int64_t)a.prefetch_ahead * E;
  int na = 0;
  if (a.prefetch_ahead > 0 && ea < a.nE && ea >= a.e_begin) na = (int)((a.nE - ea) < E ? (a.nE - ea) : E);
  else ea = -1;
  element_tile<DIM, NN, NFN, E, MODE, Cfg::T, false, MMA>(op, a, smem_raw, s_red, e0, ne, ea, na, tid);
  if (PIPE) {
    __syncthreads();
    if (tid == 0) pipe_release(a.pipe, a.ctl);
  }
}

// ------------------------------------------------------------------------------------------------------
// k_fused: one residual evaluation (+ RK stage) in ONE launch.
//
// k_face_flux is instruction-issue bound and k_element_rk is HBM-latency bound; run back to back, their times add and
// the face records make a round trip through DRAM (2 x 171 MB of the 917 MB a C3 evaluation moves).  Here every CTA
// draws a ticket T (atomic counter, so tickets follow the order in which CTAs really start) and runs
//     face group T            NSUB tiles of FT faces of the face list, which is sorted by the lowest element it touches
//     element tile T - lag    whose faces are then a PREFIX of that list: groups [0, need[T-lag]) with need[u] <= u + lag
// so that the records an element tile integrates were written ~lag tickets earlier and are still in L2; the tile drops
// them from L2 when it is done (discard.global.L2: no write-back).  CTAs in their face phase and CTAs in their
// element phase share every SM, which overlaps the issue-bound and the memory-bound halves of the evaluation.
//
// Synchronisation: a face group publishes flags[g] = epoch (release) and advances the watermark W over the
// leading complete groups; an element tile waits for W >= need (acquire).  Deadlock-free by construction: the
// groups a ticket waits for have smaller tickets, i.e. they are running or done, and a face phase never waits.
// The wait polls ctl->stop (physics error raised by another CTA) and gives up after a bounded number of polls
// (err_code 3) instead of hanging the device.  The last CTA to finish resets the ticket / watermark counters and
// bumps the epoch, so the launch can be replayed from a CUDA graph.
// ------------------------------------------------------------------------------------------------------
struct Sched {
  unsigned ticket, watermark, finished, epoch;
};

struct FusedArgs {
  FaceArgs f;                  // f.g0, f.ng: the face range of this launch (groups count from f.g0)
  ElemArgs e;
  const int32_t* tile_list;    // element tile -> tile number (first element / E); nullptr: identity
  const int32_t* need;         // per element tile: number of leading face groups that must be complete
  int32_t n_tiles, n_groups, lag;
  int32_t prefetch_ahead_groups;
  int32_t acquire_fence;       // element tiles issue fence.acq_rel.gpu after their wait (invalidates the SM's L1)
  Sched* sched;
  unsigned* flags;             // [n_groups]
};

template <int DIM, int NN, int NFN, int E, int FT, int NSUB>
struct FusedCfg {
  using TC = TileCfg<DIM, NN, NFN, E>;
  using FC = FaceCfg<DIM, NN, NFN, FT>;
  static constexpr int T = TC::T > FC::T ? TC::T : FC::T;
  static constexpr size_t face_bytes = sizeof(FaceTileSmem<DIM, NN, NFN, FT>);
  static constexpr size_t smem_bytes = TC::smem_bytes > face_bytes ? TC::smem_bytes : face_bytes;
};

// polls use relaxed loads: ld.acquire compiles to LDG.STRONG.GPU + CCTL.IVALL, and an L1 invalidation per poll starves
// the face gathers of every CTA on the SM (measured: 10x slower).  The acquire is one fence after the poll succeeds.
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_sc_gpu() { asm volatile("fence.sc.gpu;" ::: "memory"); }

// One warp scans the flags of the 32 groups that follow the watermark and publishes the longest complete prefix
// (atomicMax: every published value is a valid prefix, so concurrent scans commute).  A serial one-group-per-step
// advance costs two dependent L2 round trips per group (~0.5 us x 5672 groups: measured 2.7 ms per evaluation).
// Returns the watermark this warp knows of (warp-uniform).
__device__ __forceinline__ unsigned advance_watermark(Sched* sched, const unsigned* flags, int n_groups, unsigned epoch,
                                                     int lane) {
  unsigned w = __shfl_sync(0xffffffffu, lane == 0 ? ld_relaxed_u32(&sched->watermark) : 0u, 0);
  while (w < (unsigned)n_groups) {
    const bool set = w + lane < (unsigned)n_groups && ld_relaxed_u32(flags + w + lane) == epoch;
    const unsigned m = __ballot_sync(0xffffffffu, set);
    const int lead = m == 0xffffffffu ? 32 : __ffs(~m) - 1;
    if (lead == 0) break;
    if (lane == 0) {
      fence_acq_rel_gpu();       // relay: the flag reads synchronise with the groups' releases, the max republishes them
      atomicMax(&sched->watermark, w + lead);
    }
    w += lead;
    if (lead < 32) break;
  }
  return w;
}

template <int DIM, int NN, int NFN, int E, int FT, int NSUB, int MODE, int MINB>
__global__ void __launch_bounds__((FusedCfg<DIM, NN, NFN, E, FT, NSUB>::T), MINB)
k_fused(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ FusedArgs a) {
  using Cfg = FusedCfg<DIM, NN, NFN, E, FT, NSUB>;
  constexpr int T = Cfg::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[T / 32];
  __shared__ unsigned s_ctl[3];      // ticket, epoch, stop
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_ctl[2] = (unsigned)*reinterpret_cast<const volatile int32_t*>(&a.e.ctl->stop);
    s_ctl[1] = *reinterpret_cast<const volatile unsigned*>(&a.sched->epoch);
    s_ctl[0] = atomicAdd(&a.sched->ticket, 1u);
  }
  __syncthreads();
  const int ticket = (int)s_ctl[0];
  const unsigned epoch = s_ctl[1];
  bool live = s_ctl[2] == 0;

  // ---- face group ---------------------------------------------------------------------------------------
  if (live && ticket < a.n_groups) {
    auto& sm = *reinterpret_cast<FaceTileSmem<DIM, NN, NFN, FT>*>(smem_raw);
    face_tables<DIM, NN, NFN, FT, T>(op, sm, tid);
    const int64_t gend = a.f.g0 + a.f.ng;
    const int64_t gg = a.f.g0 + (int64_t)ticket * (NSUB * FT);
#pragma unroll 1
    for (int sub = 0; sub < NSUB; ++sub) {
      const int64_t g0 = gg + sub * FT;
      if (g0 >= gend) break;
      const int nf = (int)((gend - g0) < FT ? (gend - g0) : FT);
      const int64_t ga = a.prefetch_ahead_groups > 0 ? g0 + (int64_t)a.prefetch_ahead_groups * (NSUB * FT) : -1;
      face_tile<DIM, NN, NFN, FT, T>(op, a.f, sm, g0, nf, ga, tid);
      __syncthreads();
    }
    // every thread's records are written (barrier above): publish the group, then advance the watermark over the
    // leading complete groups (fence.sc between the flag store and the scan: of two groups that complete
    // concurrently at least one sees the other's flag; waiting tiles run the same scan, so progress never
    // depends on who wins)
    if (tid < 32) {
      if (tid == 0) {
        st_release_u32(a.flags + ticket, epoch);
        fence_sc_gpu();
      }
      __syncwarp();
      advance_watermark(a.sched, a.flags, a.n_groups, epoch, tid);
    }
  }

  // ---- element tile ---------------------------------------------------------------------------------------
  const int u = ticket - a.lag;
  if (live && u >= 0 && u < a.n_tiles) {
    if (tid < 32) {
      const unsigned need = (unsigned)a.need[u];
      unsigned ok = 1, spins = 0;
      unsigned w = __shfl_sync(0xffffffffu, tid == 0 ? ld_relaxed_u32(&a.sched->watermark) : 0u, 0);
      while (w < need) {
        w = advance_watermark(a.sched, a.flags, a.n_groups, epoch, tid);
        if (w >= need) break;
        __nanosleep(128);
        int stop = 0;
        if (tid == 0) stop = *reinterpret_cast<const volatile int32_t*>(&a.e.ctl->stop);
        stop = __shfl_sync(0xffffffffu, stop, 0);
        if (stop) { ok = 0; break; }
        if (++spins > (1u << 22)) {          // ~1 s: never reached unless the schedule is broken
          if (tid == 0) {
            atomicExch(&a.e.ctl->err_code, 3);
            atomicExch(&a.e.ctl->stop, 1);
          }
          ok = 0;
          break;
        }
      }
      if (tid == 0) {
        if (a.acquire_fence) fence_acq_rel_gpu();
        s_ctl[2] = ok ? 0u : 1u;
      }
    }
    __syncthreads();       // also separates the face phase's shared-memory use from the element tile's
    live = s_ctl[2] == 0;
    if (live) {
      const int64_t tile = a.tile_list ? a.tile_list[u] : u;
      const int64_t e0 = tile * E;
      const int ne = (int)((a.e.nE - e0) < E ? (a.e.nE - e0) : E);
      int64_t ea = -1;
      int na = 0;
      const int ua = u + a.e.prefetch_ahead;
      if (a.e.prefetch_ahead > 0 && ua < a.n_tiles) {
        ea = (a.tile_list ? (int64_t)a.tile_list[ua] : (int64_t)ua) * E;
        na = (int)((a.e.nE - ea) < E ? (a.e.nE - ea) : E);
      }
      element_tile<DIM, NN, NFN, E, MODE, T, true>(op, a.e, smem_raw, s_red, e0, ne, ea, na, tid);
    }
  }

  // ---- the last CTA re-arms the scheduler ---------------------------------------------------------------------
  if (tid == 0) {
    __threadfence();
    const unsigned done = atomicAdd(&a.sched->finished, 1u);
    if (done == gridDim.x - 1) {
      a.sched->ticket = 0;
      a.sched->watermark = 0;
      a.sched->finished = 0;
      __threadfence();
      a.sched->epoch = epoch + 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// k_element_w: warp-autonomous element kernel for node-independent metrics (straight-sided elements).
// Every warp owns G consecutive elements and never meets a block barrier, so warps drift apart and one warp's
// loads overlap another's FMA blocks.  The volume-flux tile of k_element_rk (15 doubles per node) is replaced by
// 4 doubles per node (U_d = dxidx[d,:].u, p): the variable threads rebuild F_d[k,j] = (q[k,j] + [k=E] p_j) U_dj +
// dxidx[d,k-1] p_j on the fly (+3 FP64 per 11), which cuts shared memory from 1.76 KB to 0.8 KB per element and
// lets ~2x the warps be resident.  Same arithmetic otherwise (two rows per variable thread, pipelined face records).
// ------------------------------------------------------------------------------------------------------
template <int DIM, int NN, int NFN, int G, int WPC>
struct WarpCfg {
  static constexpr int ND = DIM + 2, NF = DIM + 1, EL = NN * ND;
  static constexpr int HP = G / 2;
  static constexpr int QW = G * EL;                    // q tile, later the staged rows
  static constexpr int UW = G * NN * (DIM + 1);        // U_d (DIM) and p per node
  static constexpr int MW = G * NN;                    // Minv tile
  static constexpr int WS = ((QW + UW + MW + 1) / 2) * 2;   // doubles per warp (16-byte multiple)
  static constexpr int T = 32 * WPC;
  static constexpr size_t smem_bytes = sizeof(double) * (size_t)WS * WPC;
  static_assert(G % 2 == 0 && HP * ND <= 32, "one warp: G/2 row pairs x ND variables");
};

template <int DIM, int NN, int NFN, int G, int WPC, int MODE, int MINB>
__global__ void __launch_bounds__((32 * WPC), MINB)
k_element_w(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ ElemArgs a) {
  using Cfg = WarpCfg<DIM, NN, NFN, G, WPC>;
  constexpr int ND = Cfg::ND, NF = Cfg::NF, EL = Cfg::EL, HP = Cfg::HP;
  constexpr int FL = NFN * ND;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (a.ctl->stop) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sq = reinterpret_cast<double*>(smem_raw) + warp * Cfg::WS;    // [G][EL]
  double* sU = sq + Cfg::QW;                                            // [DIM+1][G*NN]: U_0..U_{DIM-1}, p
  double* sM = sU + Cfg::UW;                                            // [G][NN]
  const int64_t e0 = a.e_begin + ((int64_t)blockIdx.x * WPC + warp) * G;
  if (e0 >= a.nE) return;
  const int ne = (int)((a.nE - e0) < G ? (a.nE - e0) : G);
  const double gami = a.ph.gamma - 1.0;

  // ---- S0: the warp's q and Minv tiles (asynchronous copies), L2 prefetch of its later streams ---------------
  async_tile(sq, a.q + e0 * EL, ne * EL, lane, 32);
  if (MODE == EPI_RK) async_tile(sM, a.minv + e0 * NN, ne * NN, lane, 32);
  cp_async_commit();
  {
    const int64_t b0 = e0 * EL * 8, nb = (int64_t)ne * EL * 8;
    for (int64_t o = (int64_t)lane * 128; o < nb; o += 32 * 128) {
      if (MODE == EPI_RES && a.srcw) prefetch_l2(reinterpret_cast<const char*>(a.srcw) + b0 + o);
      if (MODE == EPI_RK) {
        if (a.srcm) prefetch_l2(reinterpret_cast<const char*>(a.srcm) + b0 + o);
        if (a.stage > 1) {
          prefetch_l2(reinterpret_cast<const char*>(a.x_old) + b0 + o);
          prefetch_l2(reinterpret_cast<const char*>(a.ksum) + b0 + o);
        }
      }
    }
    for (int64_t o = (int64_t)lane * 128; o < (int64_t)ne * NF * FL * 8; o += 32 * 128)
      prefetch_l2(reinterpret_cast<const char*>(a.fluxe) + e0 * NF * FL * 8 + o);
  }
  cp_async_wait<0>();
  __syncwarp();

  // ---- S1: node items: checks, pressure and the contravariant velocities U_d = dxidx[d,:].u ----------------
  for (int it = lane; it < ne * NN; it += 32) {
    const int s = it / NN, j = it - s * NN;
    double qn[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) qn[k] = sq[it * ND + k];
    const double* dx = a.dxidx + (e0 + s) * a.dx_el_stride;
    const double press = calc_pressure<DIM>(qn, gami);
    if ((a.ph.check_density && !(qn[0] > 0.0)) || (a.ph.check_pressure && !(press > 0.0))) {
      const int code = (a.ph.check_density && !(qn[0] > 0.0)) ? 1 : 2;
      const unsigned long long loc = ((unsigned long long)(e0 + s) << 8) | (unsigned)j;
      atomicMin(&a.ctl->err_loc, ((unsigned long long)(code - 1) << 62) | loc);
      atomicExch(&a.ctl->err_code, 1);
      atomicExch(&a.ctl->stop, 1);
    }
    const double rinv = 1.0 / qn[0];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double U = 0.0;
#pragma unroll
      for (int p = 0; p < DIM; ++p) U += qn[1 + p] * __ldg(dx + d + DIM * p);
      sU[d * (G * NN) + it] = U * rinv;
    }
    sU[DIM * (G * NN) + it] = press;
  }
  __syncwarp();

  // ---- S2 + S3: variable lanes, two rows each ----------------------------------------------------------------
  const int vp = lane / ND, vk = lane - vp * ND;
  const int s0 = vp, s1 = vp + HP;
  const bool act0 = lane < HP * ND && s0 < ne, act1 = lane < HP * ND && s1 < ne;
  const int s0c = act0 ? s0 : 0, s1c = act1 ? s1 : s0c;
  double acc0[NN], acc1[NN];
#pragma unroll
  for (int u = 0; u < NN; ++u) { acc0[u] = 0.0; acc1[u] = 0.0; }
  const double* G0 = a.fluxe + (e0 + s0c) * (NF * FL) + vk;
  const double* G1 = a.fluxe + (e0 + s1c) * (NF * FL) + vk;
  double g0v[NFN], g1v[NFN];
  if (act0) {
#pragma unroll
    for (int i = 0; i < NFN; ++i) { g0v[i] = __ldg(G0 + i * ND); g1v[i] = __ldg(G1 + i * ND); }
    // F_d[k,j] = (q[k,j] + ek p_j) U_dj + ck_d p_j with ek = [k is the energy], ck_d = dxidx[d,k-1] for momentum rows
    const double ek = vk == ND - 1 ? 1.0 : 0.0;
    double c0[DIM], c1[DIM];
    {
      const double* dx0 = a.dxidx + (e0 + s0c) * a.dx_el_stride;
      const double* dx1 = a.dxidx + (e0 + s1c) * a.dx_el_stride;
      const bool mom = vk >= 1 && vk <= DIM;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        c0[d] = mom ? __ldg(dx0 + d + DIM * (vk - 1)) : 0.0;
        c1[d] = mom ? __ldg(dx1 + d + DIM * (vk - 1)) : 0.0;
      }
    }
    const double* q0 = sq + s0c * EL + vk;
    const double* q1 = sq + s1c * EL + vk;
    const double* P0 = sU + DIM * (G * NN) + s0c * NN;
    const double* P1 = sU + DIM * (G * NN) + s1c * NN;
#pragma unroll 1
    for (int d = 0; d < DIM; ++d) {
      const double* U0 = sU + d * (G * NN) + s0c * NN;
      const double* U1 = sU + d * (G * NN) + s1c * NN;
      const double cd0 = c0[0] * (d == 0) + (DIM > 1 ? c0[DIM > 1 ? 1 : 0] * (d == 1) : 0.0) + (DIM > 2 ? c0[DIM > 2 ? 2 : 0] * (d == 2) : 0.0);
      const double cd1 = c1[0] * (d == 0) + (DIM > 1 ? c1[DIM > 1 ? 1 : 0] * (d == 1) : 0.0) + (DIM > 2 ? c1[DIM > 2 ? 2 : 0] * (d == 2) : 0.0);
#pragma unroll 1
      for (int j = 0; j < NN; ++j) {
        const double p0 = P0[j], p1 = P1[j];
        const double f0 = fma(fma(ek, p0, q0[j * ND]), U0[j], cd0 * p0);
        const double f1 = fma(fma(ek, p1, q1[j * ND]), U1[j], cd1 * p1);
#pragma unroll
        for (int u = 0; u < NN; ++u) {
          const double c = op.Qt[d * NN + j][u];
          acc0[u] = fma(c, f0, acc0[u]);
          acc1[u] = fma(c, f1, acc1[u]);
        }
      }
    }
    // S3: face integration (interiorfaceintegrate!, boundaryintegrate!, boundaryFaceIntegrate!)
#pragma unroll 1
    for (int f = 0; f < NF; ++f) {
      double n0v[NFN], n1v[NFN];
      const int fn = f + 1 < NF ? f + 1 : f;
#pragma unroll
      for (int i = 0; i < NFN; ++i) { n0v[i] = __ldg(G0 + (fn * NFN + i) * ND); n1v[i] = __ldg(G1 + (fn * NFN + i) * ND); }
#pragma unroll
      for (int i = 0; i < NFN; ++i)
#pragma unroll
        for (int u = 0; u < NN; ++u) {
          const double c = op.RfN[f * NFN + i][u];
          acc0[u] = fma(c, g0v[i], acc0[u]);
          acc1[u] = fma(c, g1v[i], acc1[u]);
        }
#pragma unroll
      for (int i = 0; i < NFN; ++i) { g0v[i] = n0v[i]; g1v[i] = n1v[i]; }
    }
    if (MODE == EPI_RK) {
#pragma unroll
      for (int u = 0; u < NN; ++u) { acc0[u] *= sM[s0c * NN + u]; acc1[u] *= sM[s1c * NN + u]; }
    }
  }
  __syncwarp();            // every lane is done reading the q tile: it becomes the staging tile
  if (act0) {
#pragma unroll
    for (int u = 0; u < NN; ++u) sq[s0 * EL + u * ND + vk] = acc0[u];
    if (act1) {
#pragma unroll
      for (int u = 0; u < NN; ++u) sq[s1 * EL + u * ND + vk] = acc1[u];
    }
  }
  __syncwarp();
  epilogue_tile<NN, ND, G, 32, MODE, false, true>(a, sq, ne, e0, lane, nullptr);
}

// getSendDataFace (Utils/parallel.jl:249-258): q_send[:, i, j] = R q on the shared faces, one thread per
// (shared face, face node, variable)
template <int DIM, int NN, int NFN>
__global__ void k_pack_send(const __grid_constant__ OpTab<DIM, NN, NFN> op, const double* __restrict__ q,
                            const int32_t* __restrict__ sh_el, const uint8_t* __restrict__ sh_face, int64_t nS,
                            double* __restrict__ q_send, double* const* __restrict__ face_dst, const Ctl* ctl) {
  constexpr int ND = DIM + 2;
  if (ctl->stop) return;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nS * NFN * ND) return;
  int k = (int)(t % ND);
  int i = (int)((t / ND) % NFN);
  int64_t j = t / (ND * NFN);
  const double* b = q + (int64_t)sh_el[j] * (NN * ND) + k;
  int f = sh_face[j];
  double s = 0.0;
  for (int n = 0; n < NN; ++n) s = fma(op.interp[n][i], b[op.perm[f][n] * ND], s);
  // face_dst (peer-to-peer halo): the face's slot in the NEIGHBOUR's receive buffer -- the interpolated states are
  // stored straight into peer memory over NVLink, no send buffer and no copy in between
  if (face_dst) face_dst[j][i * ND + k] = s;
  else q_send[t] = s;
}

// second pass of the stage-1 norm: deterministic sum of the per-CTA partials of this rank
// ---- peer-to-peer halo exchange (pdes_api.cu: start_exchange) -------------------------------------------------------
// after the copies into the neighbours' receive buffers (same stream): publish "my data of evaluation `epoch` has landed"
__global__ void k_halo_signal(unsigned* const* peer_flags, int npeers, unsigned epoch, const Ctl* ctl) {
  if (ctl->stop) return;
  const int p = threadIdx.x;
  if (p >= npeers) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[p]), "r"(epoch) : "memory");
}
// before the shared-face fluxes: every neighbour's data of evaluation `epoch` is in the local receive buffer.  The wait
// ends when this rank has raised `stop`, and after two minutes (err_code 4) instead of hanging the device for ever.
__global__ void k_halo_wait(const unsigned* flags, int npeers, unsigned epoch, Ctl* ctl) {
  const int p = threadIdx.x;
  if (p >= npeers) return;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (unsigned it = 0;; ++it) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + p) : "memory");
    if ((int)(v - epoch) >= 0) break;
    if ((it & 63u) == 63u) {
      if (ld_relaxed_u32(&ctl->stop)) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 120000000000ull) {
        atomicExch(&ctl->err_code, 4);
        atomicExch(&ctl->stop, 1);
        break;
      }
    }
    __nanosleep(200);
  }
}

// calcNorm's MPI.Allreduce (Utils.jl:443-448) through peer memory: every rank's receive buffer carries a ring of
// NORM_RING x 32 slots {value, tag}; k_norm_reduce stores this rank's partial sum and the step tag straight into slot
// [step % NORM_RING][rank] of EVERY rank (remote stores over NVLink, release at system scope), k_norm_commit -- enqueued where the
// values have long arrived -- waits for the tags and adds the slots in rank order (the same sum, bit for bit, on all ranks).
// Ring depth: a rank is at most one evaluation ahead of a neighbour, i.e. < NORM_RING steps ahead of anybody for <= 32 ranks.
constexpr int NORM_RING = 16;
struct NormX {
  int32_t on, rank, nranks;
  double* const* slots;        // [nranks] slot ring of every rank (own one included)
  unsigned* nctr;              // local: number of committed norms (the step tag)
};

// what k_norm_commit stores (also the tail of k_norm_reduce on one GPU: one launch less per step)
struct NormOut {
  double quirk_scale;          // reproduces the reference's double reduction in parallel runs (rk4.jl:451-453); 1 in serial
  double* norms;
  int64_t norms_cap;
  double res_tol;
  int32_t pseudo_time, fuse;   // fuse: k_norm_reduce commits the norm itself (no other rank contributes)
};
__device__ __forceinline__ void norm_commit(double sum, const NormOut& o, Ctl* ctl) {
  const double nv = sqrt(sum * o.quirk_scale);
  const int slot = ctl->norm_count;
  if (slot < o.norms_cap) o.norms[slot] = nv;
  ctl->norm_count = slot + 1;
  if (o.pseudo_time && nv < o.res_tol) { ctl->converged_step = slot; ctl->stop = 1; }
}

// second pass of the stage-1 norm: deterministic sum of the per-tile partials of this rank (fixed assignment of the
// partials to 1024 threads x 4 independent accumulators, fixed tree)
__global__ void __launch_bounds__(1024, 1)
k_norm_reduce(const double* __restrict__ partials, int n1, double* norm_sq_out, Ctl* ctl, NormX nx, NormOut no) {
  __shared__ double sh[1024];
  if (ctl->stop) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = threadIdx.x;
  for (; i + 3 * 1024 < n1; i += 4 * 1024) {
    s0 += partials[i]; s1 += partials[i + 1024]; s2 += partials[i + 2048]; s3 += partials[i + 3072];
  }
  for (; i < n1; i += 1024) s0 += partials[i];
  sh[threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *norm_sq_out = sh[0];
    if (no.fuse) norm_commit(sh[0], no, ctl);
  }
  if (nx.on && (int)threadIdx.x < nx.nranks) {
    const unsigned step = ld_relaxed_u32(nx.nctr);
    double* slot = nx.slots[threadIdx.x] + ((size_t)(step % NORM_RING) * 32 + nx.rank) * 2;
    slot[0] = sh[0];
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot + 1), "l"((unsigned long long)step + 1ull) : "memory");
  }
}

// norm_sq is the (all-reduced) sum over ranks.  quirk_scale reproduces the reference's double reduction in
// parallel runs (Utils.jl:443-448 then rk4.jl:451-453: the logged norm is sqrt(P) too large); 1 in serial.
// The slot is a device-side counter so that a captured CUDA graph of one RK4 step can be replayed unchanged.
__global__ void k_norm_commit(const double* norm_sq, NormOut no, Ctl* ctl, NormX nx) {
  if (ctl->stop) return;
  double sum = *norm_sq;
  if (nx.on) {
    const unsigned step = *nx.nctr;
    const double* ring = nx.slots[nx.rank] + (size_t)(step % NORM_RING) * 32 * 2;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    sum = 0.0;
    for (int r = 0; r < nx.nranks; ++r) {
      for (unsigned it = 0;; ++it) {
        unsigned long long tag;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(tag) : "l"(ring + 2 * r + 1) : "memory");
        if (tag == (unsigned long long)step + 1ull) break;
        if ((it & 63u) == 63u) {
          if (ld_relaxed_u32(&ctl->stop)) return;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t1 - t0 > 120000000000ull) { atomicCAS(&ctl->err_code, 0, 4); atomicExch(&ctl->stop, 1); return; }
        }
        __nanosleep(100);
      }
      sum += *reinterpret_cast<const volatile double*>(ring + 2 * r);
    }
    *nx.nctr = step + 1u;
  }
  norm_commit(sum, no, ctl);
}

// applySourceTerm tabulation (source.jl:27-47): srcw[:,j,e] = (w_j / jac[j,e]) * SRCExp(coords[:,j,e])
template <int DIM>
__global__ void k_tabulate_source(const double* __restrict__ coords, const double* __restrict__ jac,
                                  const double* __restrict__ w, int nn, int64_t nE, double gamma,
                                  double* __restrict__ srcw) {
  constexpr int ND = DIM + 2;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nE * nn) return;
  int j = (int)(t % nn);
  double x[DIM], S[ND];
#pragma unroll
  for (int d = 0; d < DIM; ++d) x[d] = coords[t * DIM + d];
  src_exp<DIM>(x, gamma, S);
  double fac = w[j] / jac[t];
#pragma unroll
  for (int k = 0; k < ND; ++k) srcw[t * ND + k] = fac * S[k];
}

// srcm = Minv * srcw (EPI_RK epilogue)
__global__ void k_srcm(const double* __restrict__ srcw, const double* __restrict__ minv, int nd, int64_t n_nodes,
                       double* __restrict__ srcm) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes * nd) return;
  srcm[t] = minv[t / nd] * srcw[t];
}

__global__ void k_minv(const double* __restrict__ jac, const double* __restrict__ w, int nn, int64_t nE,
                       double* __restrict__ minv, double* __restrict__ mass) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nE * nn) return;
  const double m = w[t % nn] / jac[t];
  mass[t] = m;
  minv[t] = 1.0 / m;
}

}  // namespace pdes
