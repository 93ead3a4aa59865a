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

namespace pdes {

// per (element, local face): where the element finds its face flux and how to read it
struct __align__(8) EFace {
  int32_t gface;     // index into the face-flux array (interfaces, then boundary faces, then shared faces)
  uint8_t right;     // 1: this element is elementR of the interface (+flux, node order permuted by nbrperm)
  uint8_t orient;
  uint8_t pad[2];
};

// per face: what k_face_flux gathers
struct __align__(16) FaceRec {
  int32_t elL, elR;  // elR: right element (interior) | unused
  uint8_t fL, fR, orient, kind;
  int32_t aux;       // boundary: BC functor id; shared: index into the receive buffer
};
enum FaceKind : uint8_t { FK_INTERIOR = 0, FK_BOUNDARY = 2, FK_SHARED = 3 };

struct Ctl {               // device-resident control block
  int32_t stop;            // kernels return immediately when set (physics error or res_tol reached)
  int32_t err_code;        // != 0: physics error, decoded on the host from err_loc
  unsigned long long err_loc;  // ((code-1) << 62) | (element << 8) | node of the lowest offending location
  int32_t converged_step;  // step head at which norm < res_tol (or -1)
  int32_t pad;
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

struct FaceArgs {
  const double* q;             // [ND,NN,nE]
  const FaceRec* faces;        // [nF + nB + nS]
  const double* nrm;           // [DIM,NFN,nF+nB+nS]   nrm_face | nrm_bndry | nrm_sharedface  (or [DIM,nG] when every
                               // face has node-independent normals: straight-sided meshes, detected at upload)
  int32_t nrm_face_stride, nrm_node_stride;   // in doubles: (NFN*DIM, DIM) or (DIM, 0)
  const double* coords_bndry;  // [DIM,NFN,nB]
  const double* q_recv;        // [ND,NFN,nS] (peer's own face-node order)
  double* fluxw;               // [ND,NFN,nF+nB+nS]    wface[i] * flux[:,i]
  int64_t g0, ng;              // face range of this launch
  int64_t nF;                  // first boundary face
  const Ctl* ctl;
  PhysPar ph;
};

struct ElemArgs {
  const double* q;             // [ND,NN,nE]
  const double* dxidx;         // [DIM,DIM,NN,nE]  (or [DIM,DIM,nE] when node-independent: straight-sided elements)
  int32_t dx_el_stride, dx_node_stride;       // in doubles: (NN*DIM*DIM, DIM*DIM) or (DIM*DIM, 0)
  const EFace* efaces;         // [nE][NF]
  const double* fluxw;         // from k_face_flux
  const double* srcw;          // [ND,NN,nE] (w_j/jac_j) * S(x_j), or nullptr
  double* res;                 // EPI_RES: [ND,NN,nE]
  // EPI_RK (rk4.jl:244-319): k = Minv*res; q_next = x_old + ah*k; ksum updated; last stage: x_new
  const double* minv;          // [NN,nE]  1/(w_j/jac_j)   (mass_matrix.jl:20-44)
  const double* x_old;
  double* ksum;
  double* q_next;
  double* norm_partials;       // [gridDim.x] sum_j M_j k_j^2 per CTA (stage 1 only)
  double ah;                   // a_s * h
  double h6;                   // h/6 (last stage)
  int32_t stage;               // 1..4
  int32_t prefetch_ahead;      // tiles between this CTA and the one whose inputs it prefetches into L2
  int64_t nE;
  Ctl* ctl;
  PhysPar ph;
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__host__ __device__ constexpr int pad_stride(int n, int nd) {
  // smallest m >= n with m % 16 == nd % 16: (item, variable)-indexed fp64 accesses of a half-warp
  // then fall into distinct banks
  int m = n;
  while (m % 16 != nd % 16) ++m;
  return m;
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

// ------------------------------------------------------------------------------------------------------
// k_face_flux: FT faces per CTA
// ------------------------------------------------------------------------------------------------------
template <int DIM, int NN, int NFN, int FT>
struct FaceCfg {
  static constexpr int ND = DIM + 2;
  static constexpr int PER = ND > NFN ? ND : NFN;             // threads per face
  static constexpr int T = ((FT * PER + 31) / 32) * 32;
  static constexpr int FS = pad_stride(NFN * ND, ND);          // per-face stride of a face-state tile
};

template <int DIM, int NN, int NFN, int FT, int MINB>
__global__ void __launch_bounds__((FaceCfg<DIM, NN, NFN, FT>::T), MINB)
k_face_flux(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ FaceArgs a) {
  using Cfg = FaceCfg<DIM, NN, NFN, FT>;
  constexpr int ND = Cfg::ND, T = Cfg::T, FS = Cfg::FS, NF = DIM + 1, EL = NN * ND;
  __shared__ double sL[FT * FS];
  __shared__ double sR[FT * FS];
  __shared__ FaceRec sRec[FT];
  __shared__ int s_perm[NF][NN];
  __shared__ int s_nbrperm[OpTab<DIM, NN, NFN>::NOR][NFN];

  if (a.ctl->stop) return;
  const int tid = threadIdx.x;
  const int64_t g0 = a.g0 + (int64_t)blockIdx.x * FT;
  const int64_t rem = a.g0 + a.ng - g0;
  const int nf = (int)(rem < FT ? rem : FT);
  if (tid < nf) sRec[tid] = a.faces[g0 + tid];
  for (int idx = tid; idx < NF * NN; idx += T) s_perm[idx / NN][idx % NN] = op.perm[idx / NN][idx % NN];
  for (int idx = tid; idx < OpTab<DIM, NN, NFN>::NOR * NFN; idx += T)
    s_nbrperm[idx / NFN][idx % NFN] = op.nbrperm[idx / NFN][idx % NFN];
  __syncthreads();

  // ---- A: interpolate both sides to the face nodes (variable threads) ----------------------------------
  if (tid < nf * ND) {
    const int fi = tid / ND, k = tid - fi * ND;
    const FaceRec r = sRec[fi];
    double ql[NN], qr[NN];
    {
      const double* b = a.q + (int64_t)r.elL * EL + k;
#pragma unroll
      for (int j = 0; j < NN; ++j) ql[j] = __ldg(b + s_perm[r.fL][j] * ND);
    }
    if (r.kind == FK_INTERIOR) {
      const double* b = a.q + (int64_t)r.elR * EL + k;
#pragma unroll
      for (int j = 0; j < NN; ++j) qr[j] = __ldg(b + s_perm[r.fR][j] * ND);
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
        // elementR's face node i coincides with elementL's face node nbrperm[i,orient] (involution)
        sR[fi * FS + s_nbrperm[r.orient][i] * ND + k] = s;
      }
    } else if (r.kind == FK_SHARED) {
      // permuteinterface! (Utils/parallel.jl:198-201): received node i of the peer is own node nbrperm[i,orient]
      const double* b = a.q_recv + (int64_t)r.aux * (NFN * ND) + k;
#pragma unroll
      for (int i = 0; i < NFN; ++i) sR[fi * FS + s_nbrperm[r.orient][i] * ND + k] = b[i * ND];
    }
  }
  __syncthreads();

  // ---- B: numerical flux at every face node (node threads), scaled by wface ------------------------------
  if (tid < nf * NFN) {
    const int fi = tid / NFN, i = tid - fi * NFN;
    const FaceRec r = sRec[fi];
    const int64_t g = g0 + fi;
    const double* np_ = a.nrm + g * a.nrm_face_stride + i * a.nrm_node_stride;
    double nrm[DIM], qL[ND], flux[ND];
#pragma unroll
    for (int d = 0; d < DIM; ++d) nrm[d] = __ldg(np_ + d);
    double* po = sL + fi * FS + i * ND;
#pragma unroll
    for (int k = 0; k < ND; ++k) qL[k] = po[k];
    if (r.kind == FK_BOUNDARY) {
      // separate copies so that only this (rare) path touches local memory
      const double* xp = a.coords_bndry + ((g - a.nF) * NFN + i) * DIM;
      double xb[DIM], nb_[DIM], qb[ND], fb[ND];
#pragma unroll
      for (int d = 0; d < DIM; ++d) { xb[d] = xp[d]; nb_[d] = nrm[d]; }
#pragma unroll
      for (int k = 0; k < ND; ++k) qb[k] = qL[k];
      bc_flux<DIM>(r.aux, qb, xb, nb_, a.ph, fb);
#pragma unroll
      for (int k = 0; k < ND; ++k) flux[k] = fb[k];
    } else {
      double qR[ND];
      const double* pn = sR + fi * FS + i * ND;
#pragma unroll
      for (int k = 0; k < ND; ++k) qR[k] = pn[k];
      roe_flux<DIM>(qL, qR, nrm, a.ph.gamma, flux);
    }
    const double w = op.wface[i];
#pragma unroll
    for (int k = 0; k < ND; ++k) po[k] = w * flux[k];
  }
  __syncthreads();

  // ---- C: coalesced store of the tile ----------------------------------------------------------------------
  double* dst = a.fluxw + g0 * (NFN * ND);
  for (int idx = tid; idx < nf * NFN * ND; idx += T) {
    const int fi = idx / (NFN * ND), r = idx - fi * (NFN * ND);
    dst[idx] = sL[fi * FS + r];
  }
}

// ------------------------------------------------------------------------------------------------------
// k_element_rk: E elements per CTA
// ------------------------------------------------------------------------------------------------------
template <int DIM, int NN, int NFN, int E>
struct TileCfg {
  static constexpr int ND = DIM + 2, NF = DIM + 1;
  static constexpr int VT = E * ND;                             // variable threads
  static constexpr int T = ((VT + 31) / 32) * 32;
  static constexpr int SQ = pad_stride(NN * ND, ND);           // per-element stride of the q tile
  static constexpr int SF = ND * DIM * NN;                      // per-element stride of the volume-flux tile
  static constexpr size_t smem_bytes = sizeof(double) * (size_t)E * (SQ + SF);
};

template <int DIM, int NN, int NFN, int E, int MODE, int MINB>
__global__ void __launch_bounds__((TileCfg<DIM, NN, NFN, E>::T), MINB)
k_element_rk(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ ElemArgs a) {
  using Cfg = TileCfg<DIM, NN, NFN, E>;
  constexpr int ND = Cfg::ND, NF = Cfg::NF, T = Cfg::T, SQ = Cfg::SQ;
  constexpr int EL = NN * ND;                       // doubles per element
  constexpr int FL = NFN * ND;                      // doubles per face
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sq = reinterpret_cast<double*>(smem_raw);             // [E][SQ]
  double* sF = sq + E * SQ;                                     // [E][ND][DIM][NN]
  __shared__ int s_nbrperm[OpTab<DIM, NN, NFN>::NOR][NFN];
  __shared__ double s_red[T / 32];

  if (a.ctl->stop) return;
  const int tid = threadIdx.x;
  const int64_t e0 = (int64_t)blockIdx.x * E;
  const int ne = (int)((a.nE - e0) < E ? (a.nE - e0) : E);
  const double gami = a.ph.gamma - 1.0;

  // ---- S0: tile load; variable threads fetch their element's face records and prefetch the face fluxes ------
  for (int idx = tid; idx < OpTab<DIM, NN, NFN>::NOR * NFN; idx += T)
    s_nbrperm[idx / NFN][idx % NFN] = op.nbrperm[idx / NFN][idx % NFN];
  {
    const double* src = a.q + e0 * EL;
    for (int idx = tid; idx < ne * EL; idx += T) {
      const int s = idx / EL, r = idx - s * EL;
      sq[s * SQ + r] = src[idx];
    }
  }
  const int v = tid;
  const int vs = v / ND, vk = v - vs * ND;
  const bool v_active = v < ne * ND;
  EFace ef[NF];
  if (v_active) {
    const EFace* pe = a.efaces + (e0 + vs) * NF;
#pragma unroll
    for (int f = 0; f < NF; ++f) ef[f] = pe[f];
    if (vk == 0) {
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const char* pf = reinterpret_cast<const char*>(a.fluxw + (int64_t)ef[f].gface * FL);
        prefetch_l2(pf);
        if (FL * 8 > 128) prefetch_l2(pf + FL * 8 - 8);
      }
    }
  }
  // L2 prefetch of the contiguous streams of the tile that will run on this SM slot next (CTAs are dispatched
  // in order, so tile blockIdx + gridDim-resident runs when this one retires): turns its HBM latency into L2 latency
  {
    const int64_t ea = e0 + (int64_t)a.prefetch_ahead * E;
    if (ea < a.nE) {
      const int na = (int)((a.nE - ea) < E ? (a.nE - ea) : E);
      const int64_t b0 = ea * EL * 8, nb = (int64_t)na * EL * 8;
      for (int64_t o = (int64_t)tid * 128; o < nb; o += (int64_t)T * 128) {
        prefetch_l2(reinterpret_cast<const char*>(a.q) + b0 + o);
        if (a.srcw) prefetch_l2(reinterpret_cast<const char*>(a.srcw) + b0 + o);
        if (MODE == EPI_RK && a.stage > 1) {
          prefetch_l2(reinterpret_cast<const char*>(a.x_old) + b0 + o);
          prefetch_l2(reinterpret_cast<const char*>(a.ksum) + b0 + o);
        }
      }
      for (int64_t o = (int64_t)tid * 128; o < (int64_t)na * a.dx_el_stride * 8; o += (int64_t)T * 128)
        prefetch_l2(reinterpret_cast<const char*>(a.dxidx) + ea * a.dx_el_stride * 8 + o);
      for (int64_t o = (int64_t)tid * 128; o < (int64_t)na * NF * 8; o += (int64_t)T * 128)
        prefetch_l2(reinterpret_cast<const char*>(a.efaces) + ea * NF * 8 + o);
      if (MODE == EPI_RK)
        for (int64_t o = (int64_t)tid * 128; o < (int64_t)na * NN * 8; o += (int64_t)T * 128)
          prefetch_l2(reinterpret_cast<const char*>(a.minv) + ea * NN * 8 + o);
    }
  }
  __syncthreads();

  // ---- S1: Euler flux in the parametric directions at every node (getEulerFlux) ---------------
  for (int it = tid; it < ne * NN; it += T) {
    const int s = it / NN, j = it - s * NN;
    double qn[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) qn[k] = sq[s * SQ + j * ND + k];
    const double* dx = a.dxidx + (e0 + s) * a.dx_el_stride + j * a.dx_node_stride;
    double dxl[DIM * DIM];
#pragma unroll
    for (int m = 0; m < DIM * DIM; ++m) dxl[m] = __ldg(dx + m);
    const double press = calc_pressure<DIM>(qn, gami);
    if ((a.ph.check_density && !(qn[0] > 0.0)) || (a.ph.check_pressure && !(press > 0.0))) {
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
  __syncthreads();

  double acc[NN];
#pragma unroll
  for (int u = 0; u < NN; ++u) acc[u] = 0.0;
  if (v_active) {
    // face fluxes of this (element, variable) row: issued now (L2 hits after the prefetch), consumed after S2.
    // interiorfaceintegrate!: elementL subtracts, elementR adds and reads node nbrperm[i,orient]
    double fl[NF][NFN];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const double* b = a.fluxw + (int64_t)ef[f].gface * FL + vk;
#pragma unroll
      for (int i = 0; i < NFN; ++i) {
        const int ii = ef[f].right ? s_nbrperm[ef[f].orient][i] : i;
        fl[f][i] = __ldg(b + ii * ND);
      }
    }
    // ---- S2: volume integral  res[k,i] = sum_d sum_j Q[j,i,d] F_d[k,j] ---------------------------
    // (the loop over directions stays rolled: fully unrolled operator products overflow the instruction cache)
    const double* Fv = sF + (vs * ND + vk) * DIM * NN;
#pragma unroll 1
    for (int d = 0; d < DIM; ++d) {
      double Fj[NN];
#pragma unroll
      for (int j = 0; j < NN; ++j) Fj[j] = Fv[d * NN + j];
#pragma unroll
      for (int j = 0; j < NN; ++j)
#pragma unroll
        for (int u = 0; u < NN; ++u) acc[u] = fma(op.Qt[d * NN + j][u], Fj[j], acc[u]);
    }
    // ---- S3: face integration  res[k,node] -+= sum_f sum_i Rf[f][i][node] * w_i f*[k,i] -----------
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const double sgn = ef[f].right ? 1.0 : -1.0;
#pragma unroll
      for (int i = 0; i < NFN; ++i) {
        const double x = sgn * fl[f][i];
#pragma unroll
        for (int u = 0; u < NN; ++u) acc[u] = fma(op.RfN[f * NFN + i][u], x, acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < NN; ++u) sq[vs * SQ + u * ND + vk] = acc[u];   // only this thread reads/writes its row
  }
  __syncthreads();

  // ---- S4: epilogue (coalesced): source, then either res or the fused RK4 stage ----------------
  // loads of a chunk of CH dofs per thread are issued together before any of them is consumed
  double nrm2 = 0.0;
  {
    constexpr int CH = 4;
    const int ntile = ne * EL;
    for (int base = 0; base < ntile; base += CH * T) {
      double val[CH], sv[CH], mi[CH], xo[CH], ks[CH];
      int64_t dof[CH];
      bool ok[CH];
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int idx = base + u * T + tid;
        ok[u] = idx < ntile;
        int s = 0, r = 0;
        if (ok[u]) { s = idx / EL; r = idx - s * EL; }
        dof[u] = ok[u] ? (e0 + s) * EL + r : 0;
        val[u] = ok[u] ? sq[s * SQ + r] : 0.0;
        sv[u] = (ok[u] && a.srcw) ? a.srcw[dof[u]] : 0.0;
        if (MODE == EPI_RK) {
          mi[u] = ok[u] ? a.minv[(e0 + s) * NN + r / ND] : 1.0;
          xo[u] = ok[u] ? a.x_old[dof[u]] : 0.0;
          ks[u] = (ok[u] && a.stage > 1) ? a.ksum[dof[u]] : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        if (!ok[u]) continue;
        const double vv = val[u] + sv[u];
        if (MODE == EPI_RES) {
          a.res[dof[u]] = vv;
        } else {
          const double k = mi[u] * vv;             // pde_post_func: res_vec *= Minv
          if (a.stage == 1) {
            nrm2 += k * k / mi[u];                 // calcNorm: sum res*M*res (Utils.jl:427-449)
            a.ksum[dof[u]] = k;
            a.q_next[dof[u]] = xo[u] + a.ah * k;
          } else if (a.stage < 4) {
            a.ksum[dof[u]] = ks[u] + 2.0 * k;
            a.q_next[dof[u]] = xo[u] + a.ah * k;
          } else {
            a.q_next[dof[u]] = xo[u] + a.h6 * (ks[u] + k);
          }
        }
      }
    }
  }
  if (MODE == EPI_RK && a.stage == 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = nrm2;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < T / 32; ++w) t += s_red[w];
      a.norm_partials[blockIdx.x] = t;
    }
  }
}

// getSendDataFace (Utils/parallel.jl:249-258): q_send[:, i, j] = R q on the shared faces, one thread per
// (shared face, face node, variable)
template <int DIM, int NN, int NFN>
__global__ void k_pack_send(const __grid_constant__ OpTab<DIM, NN, NFN> op, const double* __restrict__ q,
                            const int32_t* __restrict__ sh_el, const uint8_t* __restrict__ sh_face, int64_t nS,
                            double* __restrict__ q_send, const Ctl* ctl) {
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
  q_send[t] = s;
}

// second pass of the stage-1 norm: deterministic sum of the per-CTA partials of this rank
__global__ void k_norm_reduce(const double* __restrict__ partials, int n1, double* norm_sq_out, const Ctl* ctl) {
  __shared__ double sh[256];
  if (ctl->stop) return;
  double s = 0.0;
  for (int i = threadIdx.x; i < n1; i += 256) s += partials[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *norm_sq_out = sh[0];
}

// norm_sq is the (all-reduced) sum over ranks.  quirk_scale reproduces the reference's double reduction in
// parallel runs (Utils.jl:443-448 then rk4.jl:451-453: the logged norm is sqrt(P) too large); 1 in serial.
__global__ void k_norm_commit(const double* norm_sq, double quirk_scale, double* norms, int64_t slot, double res_tol,
                              int pseudo_time, Ctl* ctl) {
  if (ctl->stop) return;
  double nv = sqrt(*norm_sq * quirk_scale);
  norms[slot] = nv;
  if (pseudo_time && nv < res_tol) { ctl->converged_step = (int32_t)slot; ctl->stop = 1; }
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

__global__ void k_minv(const double* __restrict__ jac, const double* __restrict__ w, int nn, int64_t nE,
                       double* __restrict__ minv) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nE * nn) return;
  minv[t] = 1.0 / (w[t % nn] / jac[t]);
}

}  // namespace pdes
