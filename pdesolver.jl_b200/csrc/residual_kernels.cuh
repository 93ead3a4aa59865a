// Fused Euler residual kernel for sm_100a (dense-face SBP-Omega operators, Roe flux).
//
// One CTA evaluates the complete residual of a tile of E elements: every entry
// of res is produced by exactly one thread, there are no atomics and no
// materialised intermediates in HBM (the reference makes >= 12 mesh sweeps
// through aux_vars / flux_parametric / q_face / flux_face / q_bndry / bndryflux,
// src/solver/euler/euler.jl:441-519).  What one launch replaces:
//
//   dataPrep + checkDensity/checkPressure     euler.jl:441-519, 543-611
//   getEulerFlux + weakdifferentiate!(trans)  euler_funcs.jl:23-58, euler.jl:628-658
//   interpolateFace + calcFaceFlux (Roe)      flux.jl:613-641, 37-64, bc_solvers.jl:29-420
//   interiorfaceintegrate!                    euler.jl:770-802
//   interpolateBoundary + getBCFluxes + boundaryintegrate!   bc.jl:49-80,162-175,251-284, euler.jl:669-690
//   calcSharedFaceIntegrals_nopre_inner       flux.jl:264-308 (receive buffer filled by the halo exchange)
//   applySourceTerm                           source.jl:27-47 (time-independent source tabulated at upload)
//   pde_post_func (Minv, stage-1 norm) + the RK4 axpy loops   rk4.jl:244-319, 446-457
//
// Work decomposition inside a CTA (T threads, E elements, ND = DIM+2 variables):
//   * "variable threads": thread v < E*ND owns (element s = v/ND, variable k = v%ND).  All operator
//     applications (Q^T F, face interpolation R q, face integration R^T W f) are small dense products
//     whose coefficients are compile-time-indexed entries of the kernel-parameter operator table
//     (constant bank operands of DFMA); the thread keeps q[k,:] and the residual row res[k,:] in registers.
//   * "node items" (E*NN) evaluate the Euler flux in the DIM parametric directions.
//   * "face-node items" (E*NF*NFN) evaluate the numerical (Roe / boundary-condition) flux.
// Roles exchange data through shared memory.  The face flux is evaluated on both sides of an interior
// face (element-centric gather) so that no scatter is needed.
#pragma once
#include <stdint.h>
#include "euler_device.cuh"

namespace pdes {

// per (element, local face) connectivity record built at upload time from mesh.interfaces /
// mesh.bndryfaces / mesh.shared_interfaces
struct __align__(16) EFace {
  int32_t nbr;       // neighbour element (interior faces)
  int32_t idx;       // interface / boundary-face / shared-face index (normals, coordinates, receive buffer)
  uint8_t kind;      // FaceKind
  uint8_t fnbr;      // neighbour's local face
  uint8_t orient;    // interface orientation
  uint8_t bc;        // BC functor id (boundary faces)
  uint32_t pad;
};
enum FaceKind : uint8_t { FK_INTERIOR_L = 0, FK_INTERIOR_R = 1, FK_BOUNDARY = 2, FK_SHARED = 3 };

struct Ctl {               // device-resident control block
  int32_t stop;            // kernels return immediately when set (physics error or res_tol reached)
  int32_t err_code;        // 0 | PDES_ERR_NEG_DENSITY | PDES_ERR_NEG_PRESSURE
  unsigned long long err_loc;  // (element << 8) | node of the lowest offending location
  int32_t converged_step;  // step head at which norm < res_tol (or -1)
  int32_t pad;
};

template <int DIM, int NN, int NFN>
struct OpTab {
  static constexpr int NF = DIM + 1;
  static constexpr int NOR = (DIM == 2) ? 1 : 3;
  static constexpr int NNP = NN, NFP = NF * NFN;
  double Qt[DIM * NN][NNP];   // Qt[d*NN+j][i] = sbp.Q[j,i,d]   (res_i += Q[j,i,d] F_j : weakdifferentiate!, trans=true)
  double RfN[NF * NFN][NNP];  // RfN[f*NFN+i][node] = sum_j interp[j,i] [perm[j,f]==node]   (face integration)
  double RfT[NN][NFP];        // RfT[node][f*NFN+i] = the same matrix, transposed               (face interpolation)
  double interp[NN][NFN];     // sbpface.interp[j,i] (stencil order, used for the neighbour side)
  double wface[NFN];
  int32_t perm[NF][NN];       // sbpface.perm[j,f] (0-based)
  int32_t nbrperm[NOR][NFN];  // sbpface.nbrperm[i,orient] (0-based)
};

struct PhysPar {
  double gamma, R, Ma, aoa, rho_free, E_free;
  int32_t check_density, check_pressure;
};

enum EpiMode { EPI_RES = 0, EPI_RK = 1 };

struct ResArgs {
  const double* q;             // [ND,NN,nE]
  const double* dxidx;         // [DIM,DIM,NN,nE]
  const EFace* efaces;         // [nE][NF]
  const double* nrm_face;      // [DIM,NFN,nF]
  const double* nrm_bndry;     // [DIM,NFN,nB]
  const double* coords_bndry;  // [DIM,NFN,nB]
  const double* nrm_shared;    // [DIM,NFN,nS]  all peers concatenated
  const double* q_recv;        // [ND,NFN,nS]   all peers concatenated (peer's own face-node order)
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
  // tiling
  int64_t nE;
  const int32_t* elist;        // optional compacted element list (launch over surface elements)
  int64_t nlist;
  int32_t skip_shared;         // 1: elements that own a shared face are left to the elist launch
  Ctl* ctl;
  PhysPar ph;
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__host__ __device__ constexpr int pad_stride(int n, int nd) {
  // smallest m >= n with m % 16 == nd % 16: (element, variable)-indexed fp64 accesses of a half-warp
  // then fall into distinct banks
  int m = n;
  while (m % 16 != nd % 16) ++m;
  return m;
}

template <int DIM, int NN, int NFN, int E>
struct TileCfg {
  static constexpr int ND = DIM + 2, NF = DIM + 1;
  static constexpr int H = 1;                                  // threads per (element, variable) row
  static constexpr int VT = E * ND * H;                        // variable threads
  static constexpr int T = ((VT + 31) / 32) * 32;
  static constexpr int SQ = pad_stride(NN * ND, ND);           // per-element stride of the q tile
  static constexpr int FS = pad_stride(NF * NFN * ND, ND);     // per-element stride of face-state tiles
  static constexpr int SF = ND * DIM * NN;                      // per-element stride of the volume-flux tile
  static constexpr int UNION = (2 * FS > SF) ? 2 * FS : SF;
  static constexpr size_t smem_bytes = sizeof(double) * (size_t)E * (SQ + UNION) + sizeof(EFace) * E * NF;
};

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

template <int DIM, int NN, int NFN, int E, int MODE, int MINB>
__global__ void __launch_bounds__((TileCfg<DIM, NN, NFN, E>::T), MINB)
k_residual_roe(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ ResArgs a) {
  using Cfg = TileCfg<DIM, NN, NFN, E>;
  constexpr int ND = Cfg::ND, NF = Cfg::NF, T = Cfg::T, SQ = Cfg::SQ, FS = Cfg::FS;
  constexpr int EL = NN * ND;                       // doubles per element
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sq = reinterpret_cast<double*>(smem_raw);             // [E][SQ]
  double* sU = sq + E * SQ;                                     // union
  double* sF = sU;                                              // [E][ND][DIM][NN]
  double* sOwn = sU;                                            // [E][FS]
  double* sNbr = sU + E * FS;                                   // [E][FS]
  EFace* sEf = reinterpret_cast<EFace*>(sU + E * Cfg::UNION);   // [E][NF]
  __shared__ int s_el[E];
  __shared__ int s_skip[E];
  __shared__ int s_perm[NF][NN];
  __shared__ int s_nbrperm[OpTab<DIM, NN, NFN>::NOR][NFN];
  __shared__ double s_red[T / 32];

  if (a.ctl->stop) return;
  const int tid = threadIdx.x;
  const int64_t ntot = a.elist ? a.nlist : a.nE;
  const int64_t e0 = (int64_t)blockIdx.x * E;
  const int ne = (int)((ntot - e0) < E ? (ntot - e0) : E);
  const double gami = a.ph.gamma - 1.0;

  // ---- S0: tile load --------------------------------------------------------------------------
  if (tid < E) {
    int el = -1;
    if (tid < ne) el = a.elist ? a.elist[e0 + tid] : (int)(e0 + tid);
    s_el[tid] = el;
  }
  for (int idx = tid; idx < NF * NN; idx += T) s_perm[idx / NN][idx % NN] = op.perm[idx / NN][idx % NN];
  for (int idx = tid; idx < OpTab<DIM, NN, NFN>::NOR * NFN; idx += T) s_nbrperm[idx / NFN][idx % NFN] = op.nbrperm[idx / NFN][idx % NFN];
  __syncthreads();
  for (int idx = tid; idx < ne * NF; idx += T) {
    int s = idx / NF, f = idx - s * NF;
    sEf[idx] = a.efaces[(int64_t)s_el[s] * NF + f];
  }
  if (a.elist == nullptr) {
    const double* src = a.q + e0 * EL;
    for (int idx = tid; idx < ne * EL; idx += T) {
      int s = idx / EL, r = idx - s * EL;
      sq[s * SQ + r] = src[idx];
    }
  } else {
    for (int idx = tid; idx < ne * EL; idx += T) {
      int s = idx / EL, r = idx - s * EL;
      sq[s * SQ + r] = a.q[(int64_t)s_el[s] * EL + r];
    }
  }
  __syncthreads();

  // elements this launch must not touch (they own a shared face and are done by the elist launch)
  if (tid < E) {
    int sk = 0;
    if (tid < ne && a.skip_shared) {
#pragma unroll
      for (int f = 0; f < NF; ++f) sk |= (sEf[tid * NF + f].kind == FK_SHARED);
    }
    s_skip[tid] = sk;
  }
  __syncthreads();
  // L2 prefetch of everything the later stages gather: neighbour elements, face normals, and the
  // epilogue's streams (the loads themselves are issued much later; this converts their HBM latency into L2 latency)
  for (int idx = tid; idx < ne * NF; idx += T) {
    const EFace ef = sEf[idx];
    if (ef.kind <= FK_INTERIOR_R) {
      const char* pq = reinterpret_cast<const char*>(a.q + (int64_t)ef.nbr * EL);
#pragma unroll
      for (int o = 0; o < EL * 8 + 127; o += 128) prefetch_l2(pq + o);
      const char* pn = reinterpret_cast<const char*>(a.nrm_face + (int64_t)ef.idx * NFN * DIM);
      prefetch_l2(pn);
      prefetch_l2(pn + NFN * DIM * 8 - 8);
    } else if (ef.kind == FK_BOUNDARY) {
      prefetch_l2(a.nrm_bndry + (int64_t)ef.idx * NFN * DIM);
      prefetch_l2(a.coords_bndry + (int64_t)ef.idx * NFN * DIM);
    }
  }
  if (a.elist == nullptr) {
    const int64_t b0 = e0 * EL * 8, nb = (int64_t)ne * EL * 8;
    for (int64_t o = (int64_t)tid * 128; o < nb; o += (int64_t)T * 128) {
      if (a.srcw) prefetch_l2(reinterpret_cast<const char*>(a.srcw) + b0 + o);
      if (MODE == EPI_RK) {
        if (a.stage > 1) {
          prefetch_l2(reinterpret_cast<const char*>(a.x_old) + b0 + o);
          prefetch_l2(reinterpret_cast<const char*>(a.ksum) + b0 + o);
        }
      }
    }
    if (MODE == EPI_RK)
      for (int64_t o = (int64_t)tid * 128; o < (int64_t)ne * NN * 8; o += (int64_t)T * 128)
        prefetch_l2(reinterpret_cast<const char*>(a.minv) + e0 * NN * 8 + o);
  }
  // variable threads: v -> (element vs, variable vk)
  const int v = tid;
  const int vs = v / ND, vk = v - vs * ND;
  const bool v_active = (v < ne * ND) && !s_skip[vs < E ? vs : 0];

  // ---- S1: Euler flux in the parametric directions at every node (getEulerFlux) ---------------
  for (int it = tid; it < ne * NN; it += T) {
    int s = it / NN, j = it - s * NN;
    double qn[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) qn[k] = sq[s * SQ + j * ND + k];
    const double* dx = a.dxidx + ((int64_t)s_el[s] * NN + j) * (DIM * DIM);
    double dxl[DIM * DIM];
#pragma unroll
    for (int m = 0; m < DIM * DIM; ++m) dxl[m] = __ldg(dx + m);
    double press = calc_pressure<DIM>(qn, gami);
    if ((a.ph.check_density && !(qn[0] > 0.0)) || (a.ph.check_pressure && !(press > 0.0))) {
      int code = (a.ph.check_density && !(qn[0] > 0.0)) ? 1 : 2;
      unsigned long long loc = ((unsigned long long)s_el[s] << 8) | (unsigned)j;
      // density errors win over pressure errors (checkDensity runs first), lowest location wins
      unsigned long long key = ((unsigned long long)(code - 1) << 62) | loc;
      atomicMin(&a.ctl->err_loc, key);
      atomicExch(&a.ctl->err_code, 1);  // decoded on the host from err_loc
      atomicExch(&a.ctl->stop, 1);
    }
    double rinv = 1.0 / qn[0];
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

  // ---- S2: volume integral  res[k,i] = sum_d sum_j Q[j,i,d] F_d[k,j] ---------------------------
  // (loops over directions / faces stay rolled: the fully unrolled operator products overflow the instruction
  // cache; one direction or one face at a time keeps >= NN independent DFMA chains in flight)
  double acc[NN];
#pragma unroll
  for (int u = 0; u < NN; ++u) acc[u] = 0.0;
  if (v_active) {
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
  }
  __syncthreads();   // sF is dead; the face-state tiles reuse its storage

  // ---- S3/S4: face interpolation of the own and of the neighbour state, one face at a time -------
  if (v_active) {
    double qk[NN];
#pragma unroll
    for (int j = 0; j < NN; ++j) qk[j] = sq[vs * SQ + j * ND + vk];
#pragma unroll 1
    for (int f = 0; f < NF; ++f) {
      const EFace ef = sEf[vs * NF + f];
      // neighbour values in the neighbour's stencil order: q[k, perm[j,fnbr], nbr] (issued first: L2 latency)
      double qn[NN];
      const bool interior = ef.kind <= FK_INTERIOR_R;
      if (interior) {
        const int64_t loc = (int64_t)ef.nbr - e0;
        if (a.elist == nullptr && loc >= 0 && loc < ne) {
          const double* b = sq + (int)loc * SQ + vk;
#pragma unroll
          for (int j = 0; j < NN; ++j) qn[j] = b[s_perm[ef.fnbr][j] * ND];
        } else {
          const double* b = a.q + (int64_t)ef.nbr * EL + vk;
#pragma unroll
          for (int j = 0; j < NN; ++j) qn[j] = __ldg(b + s_perm[ef.fnbr][j] * ND);
        }
      }
      // own state at the NFN nodes of face f
#pragma unroll
      for (int i = 0; i < NFN; ++i) {
        double sacc = 0.0;
#pragma unroll
        for (int n = 0; n < NN; ++n) sacc = fma(op.RfN[f * NFN + i][n], qk[n], sacc);
        sOwn[vs * FS + (f * NFN + i) * ND + vk] = sacc;
      }
      if (interior) {
#pragma unroll
        for (int i = 0; i < NFN; ++i) {
          double sacc = 0.0;
#pragma unroll
          for (int j = 0; j < NN; ++j) sacc = fma(op.interp[j][i], qn[j], sacc);
          // the neighbour's face node i coincides with own face node nbrperm[i,orient] (involution)
          const int io = s_nbrperm[ef.orient][i];
          sNbr[vs * FS + (f * NFN + io) * ND + vk] = sacc;
        }
      } else if (ef.kind == FK_SHARED) {
        // permuteinterface! (Utils/parallel.jl:198-201): received face-node i of the peer is own node nbrperm[i]
        const double* b = a.q_recv + (int64_t)ef.idx * (NFN * ND) + vk;
#pragma unroll
        for (int i = 0; i < NFN; ++i) {
          const int io = s_nbrperm[ef.orient][i];
          sNbr[vs * FS + (f * NFN + io) * ND + vk] = b[i * ND];
        }
      }
    }
  }
  __syncthreads();

  // ---- S5: numerical flux at every face node, scaled by -/+ wface (in place in sOwn) ---------
  // one instance of the Roe solver (rolled loop: the unrolled operator products already fill the
  // instruction cache); the left/right roles are selected on the data.  Normals were prefetched to L2.
#pragma unroll 1
  for (int it = tid; it < ne * NF * NFN; it += T) {
    const int s = it / (NF * NFN), r = it - s * (NF * NFN);
    if (s_skip[s]) continue;
    const int f = r / NFN, i = r - f * NFN;
    const EFace ef = sEf[s * NF + f];
    const bool right = ef.kind == FK_INTERIOR_R;
    const int ii = right ? s_nbrperm[ef.orient][i] : i;   // node index in the left element's ordering
    const double* base = a.nrm_face;
    if (ef.kind == FK_BOUNDARY) base = a.nrm_bndry;
    else if (ef.kind == FK_SHARED) base = a.nrm_shared;
    const double* np_ = base + ((int64_t)ef.idx * NFN + ii) * DIM;
    double nrm[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) nrm[d] = __ldg(np_ + d);
    double* po = sOwn + s * FS + r * ND;
    const double* pn = sNbr + s * FS + r * ND;
    double flux[ND];
    double scale = right ? op.wface[ii] : -op.wface[ii];
    if (ef.kind == FK_BOUNDARY) {
      // separate copies so that only this (rare) path touches local memory
      const double* xp = a.coords_bndry + ((int64_t)ef.idx * NFN + i) * DIM;
      double xb[DIM], nb_[DIM], qb[ND], fb[ND];
#pragma unroll
      for (int d = 0; d < DIM; ++d) { xb[d] = xp[d]; nb_[d] = nrm[d]; }
#pragma unroll
      for (int k = 0; k < ND; ++k) qb[k] = po[k];
      bc_flux<DIM>(ef.bc, qb, xb, nb_, a.ph, fb);
#pragma unroll
      for (int k = 0; k < ND; ++k) flux[k] = fb[k];
    } else {
      double qa[ND], qb[ND];
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        const double o = po[k], nb = pn[k];
        qa[k] = right ? nb : o;
        qb[k] = right ? o : nb;
      }
      roe_flux<DIM>(qa, qb, nrm, a.ph.gamma, flux);
    }
#pragma unroll
    for (int k = 0; k < ND; ++k) po[k] = scale * flux[k];
  }
  __syncthreads();

  // ---- S6: face integration  res[k,node] += sum_f sum_i Rf[f][i][node] * (+-w f*)[k,i] ---------
  if (v_active) {
    const double* fv = sOwn + vs * FS + vk;
#pragma unroll 1
    for (int f = 0; f < NF; ++f) {
      double fl[NFN];
#pragma unroll
      for (int i = 0; i < NFN; ++i) fl[i] = fv[(f * NFN + i) * ND];
#pragma unroll
      for (int i = 0; i < NFN; ++i)
#pragma unroll
        for (int u = 0; u < NN; ++u) acc[u] = fma(op.RfN[f * NFN + i][u], fl[i], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < NN; ++u) sq[vs * SQ + u * ND + vk] = acc[u];
  }
  __syncthreads();

  // ---- S7: epilogue (coalesced): source, then either res or the fused RK4 stage ----------------
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
        if (ok[u]) { s = idx / EL; r = idx - s * EL; ok[u] = !s_skip[s]; }
        dof[u] = ok[u] ? (int64_t)s_el[s] * EL + r : 0;
        val[u] = ok[u] ? sq[s * SQ + r] : 0.0;
        sv[u] = (ok[u] && a.srcw) ? a.srcw[dof[u]] : 0.0;
        if (MODE == EPI_RK) {
          mi[u] = ok[u] ? a.minv[(int64_t)s_el[s] * NN + r / ND] : 1.0;
          xo[u] = ok[u] ? a.x_old[dof[u]] : 0.0;
          ks[u] = (ok[u] && a.stage > 1) ? a.ksum[dof[u]] : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        if (!ok[u]) continue;
        const double v = val[u] + sv[u];
        if (MODE == EPI_RES) {
          a.res[dof[u]] = v;
        } else {
          const double k = mi[u] * v;              // pde_post_func: res_vec *= Minv
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
__global__ void k_norm_reduce(const double* __restrict__ partials, int n1, const double* __restrict__ partials2, int n2,
                              double* norm_sq_out, const Ctl* ctl) {
  __shared__ double sh[256];
  if (ctl->stop) return;
  double s = 0.0;
  for (int i = threadIdx.x; i < n1; i += 256) s += partials[i];
  for (int i = threadIdx.x; i < n2; i += 256) s += partials2[i];
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
