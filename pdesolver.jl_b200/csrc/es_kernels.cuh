// Entropy-stable path (BASELINE config 2): SBPDiagonalE operators (sparse faces: face node i IS volume node
// perm[i,f]), split-form volume integrals with the Ismail-Roe two-point flux, IRSLF / IR / Roe interface flux.
//
//   k_face_flux_sparse   calcFaceIntegral_nopre with a SparseFace (flux.jl:79-125; index rule
//                        sbp_sat_reduced_sc.jl:969-972: iL = perm[i,faceL], iR = perm[nbrperm[i,orient],faceR]),
//                        IRSLFFlux (flux.jl:964-976 -> bc_solvers.jl:898-909), boundary functors (bc.jl:251-284)
//   k_element_split      calcVolumeIntegralsSplitFormLinear (euler_funcs.jl:240-288, S = (Q - Q^T)/2
//                        flux_types.jl:969-971) + face records + source + fused RK4 stage
//
// The volume kernel evaluates each of the nn(nn-1)/2 two-point fluxes of an element ONCE (pair threads), with
// the square roots and logarithms of the Ismail-Roe parameter vector tabulated per node, then node threads gather
// -2 S[i,m,d] F_d over their nn-1 partners.  The kernel is FP64-pipe bound (~150 FP64 instructions per pair).
#pragma once
#include "residual_kernels.cuh"

#ifndef PDES_SPLITN_UNROLL
#define PDES_SPLITN_UNROLL 2      // C2: 17.88 ms per RK4 step; 1: 18.16, 4: 19.3, 11: 22.8 (instruction cache)
#endif

namespace pdes {

template <int DIM, int NN, int NFN>
struct OpTabS {
  static constexpr int NF = DIM + 1;
  static constexpr int NOR = (DIM == 2) ? 1 : 3;
  double S2[DIM][NN][NN];     // 2*S[i][m][d] = Q[i,m,d] - Q[m,i,d]
  double wface[NFN];
  int32_t perm[NF][NFN];      // sparse face: volume node of face node i on face f
  int32_t nbrperm[NOR][NFN];
  int32_t inv[NN][DIM];       // face-node slots (f*NFN+i) that coincide with volume node n, or -1
};

enum FluxId { FLUX_ROE = 1, FLUX_IR = 2, FLUX_IRSLF = 3 };

template <int DIM, typename T>
__device__ __forceinline__ void numerical_flux(int flux_id, const T* qL, const T* qR, const double* n,
                                               double gamma, T* F) {
  if (flux_id == FLUX_IRSLF) irslf_flux<DIM>(qL, qR, n, gamma, F);
  else if (flux_id == FLUX_IR) ir_flux_single<DIM>(qL, qR, n, gamma, F);
  else roe_flux<DIM>(qL, qR, n, gamma, F);
}

// one thread per (face, face node): no interpolation, states are read at the coinciding volume nodes
template <int DIM, int NN, int NFN>
#ifndef PDES_SPARSE_MINB
#define PDES_SPARSE_MINB 8      // 64 registers, 8 CTAs per SM: C2 12.27 ms per RK4 step (1: 12.57, 6: 12.42, 10: 12.56)
#endif
__global__ void __launch_bounds__(128, PDES_SPARSE_MINB)
k_face_flux_sparse(const __grid_constant__ OpTabS<DIM, NN, NFN> op, const __grid_constant__ FaceArgs a, int flux_id) {
  constexpr int ND = DIM + 2, NF = DIM + 1, EL = NN * ND, FL = NFN * ND;
  if (a.ctl->stop) return;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.ng * NFN) return;
  const int64_t g = a.g0 + t / NFN;
  const int i = (int)(t % NFN);
  const FaceRec r = a.faces[g];
  double qL[ND], qR[ND], nrm[DIM], flux[ND];
  {
    const double* b = a.q + (int64_t)r.elL * EL + op.perm[r.fL][i] * ND;
#pragma unroll
    for (int k = 0; k < ND; ++k) qL[k] = __ldg(b + k);
    const double* np_ = a.nrm + g * a.nrm_face_stride + i * a.nrm_node_stride;
#pragma unroll
    for (int d = 0; d < DIM; ++d) nrm[d] = __ldg(np_ + d);
  }
  int iR = i;
  if (r.kind == FK_BOUNDARY) {
    const double* xp = a.coords_bndry + ((int64_t)r.elR * NFN + i) * DIM;
    double xb[DIM], nb_[DIM], qb[ND], fb[ND];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { xb[d] = xp[d]; nb_[d] = nrm[d]; }
#pragma unroll
    for (int k = 0; k < ND; ++k) qb[k] = qL[k];
    bc_flux_any<DIM>(r.aux, qb, xb, nb_, a.ph, fb);
#pragma unroll
    for (int k = 0; k < ND; ++k) flux[k] = fb[k];
  } else {
    iR = op.nbrperm[r.orient][i];
    const double* b = r.kind == FK_INTERIOR ? a.q + (int64_t)r.elR * EL + op.perm[r.fR][iR] * ND
                                            : a.q_recv + ((int64_t)r.aux * NFN + iR) * ND;
#pragma unroll
    for (int k = 0; k < ND; ++k) qR[k] = __ldg(b + k);
    numerical_flux<DIM>(flux_id, qL, qR, nrm, a.ph.gamma, flux);
  }
  const double w = op.wface[i];
  double* dl = a.fluxe + ((int64_t)r.elL * NF + r.fL) * FL + i * ND;
#pragma unroll
  for (int k = 0; k < ND; ++k) dl[k] = -w * flux[k];
  if (r.kind == FK_INTERIOR) {
    double* dr = a.fluxe + ((int64_t)r.elR * NF + r.fR) * FL + iR * ND;
#pragma unroll
    for (int k = 0; k < ND; ++k) dr[k] = w * flux[k];
  }
}

// getSendDataFace for sparse faces: q_send[:, i, j] = q[:, perm[i, face], element]
template <int DIM, int NN, int NFN>
__global__ void k_pack_send_sparse(const __grid_constant__ OpTabS<DIM, NN, NFN> op, const double* __restrict__ q,
                                   const int32_t* __restrict__ sh_el, const uint8_t* __restrict__ sh_face, int64_t nS,
                                   double* __restrict__ q_send, double* const* __restrict__ face_dst, const Ctl* ctl) {
  constexpr int ND = DIM + 2;
  if (ctl->stop) return;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nS * NFN * ND) return;
  int k = (int)(t % ND);
  int i = (int)((t / ND) % NFN);
  int64_t j = t / (ND * NFN);
  const double v = q[((int64_t)sh_el[j] * NN + op.perm[sh_face[j]][i]) * ND + k];
  if (face_dst) face_dst[j][i * ND + k] = v;      // peer-to-peer halo: straight into the neighbour's receive buffer
  else q_send[t] = v;
}


// getSendDataElement (Utils/parallel.jl:276-293): send_buff[:, :, j] = q[:, :, local_element_lists[peer][j]] (all peers'
// lists concatenated)
__global__ void k_pack_send_element(const double* __restrict__ q, const int32_t* __restrict__ el_list, int64_t nel, int el_len,
                                    double* __restrict__ out, const Ctl* ctl) {
  if (ctl->stop) return;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nel * el_len) return;
  const int64_t j = t / el_len;
  out[t] = q[(int64_t)el_list[j] * el_len + (t - j * el_len)];
}

// ------------------------------------------------------------------------------------------------------
// k_face_element: face_integral_type = 2 for dense-face (SBP-Omega) operators:
// getFaceElementIntegral (flux.jl:132-160) with the functors ECFaceIntegral / ELFPenaltyFaceIntegral /
// ESLFFaceIntegral (faceElementIntegrals.jl:586-655):
//   calcECFaceIntegral (:58-117)           nn x nn two-point Ismail-Roe fluxes in the Cartesian directions between the
//                                          stencil nodes of the two elements, weighted by
//                                          E_ij^d = sum_k interp[i,k] interp[j,nbrperm[k]] wface[k] nrm[d,k]
//   calcEntropyPenaltyIntegral (:209-290)  entropy variables interpolated to the face, LFKernel (:455-468), interpolated back
// Boundary faces keep the standard boundary integral (Dirichlet state + Roe, bc.jl:251-284; boundaryintegrate!).
// One CTA per face.  The result is one record per (element, local face) holding the contribution to EVERY volume node
// of the element ([nd, nn], element node order), which k_element_split<..., DENSEREC> adds: atomic-free, deterministic.
// ------------------------------------------------------------------------------------------------------
enum FaceElementId { FEI_EC = 1, FEI_ELF_PENALTY = 2, FEI_ESLF = 3, FEI_ELW2_PENALTY = 4, FEI_ESLW2 = 5 };

template <int DIM, int NN, int NFN>
__global__ void __launch_bounds__(128)
k_face_element(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ FaceArgs a, int fei) {
  constexpr int ND = DIM + 2, NF = DIM + 1, EL = NN * ND, T = 128;
  __shared__ double sq[2][NN][ND];          // states at the stencil nodes (stencil order) of elementL / elementR
  __shared__ IRNode<DIM> sZ[2][NN];
  __shared__ double sw[2][NN][ND];          // IR entropy variables
  __shared__ double sG[NN * NN][ND];        // sum_d E_ij^d F_d(q_i, q_j)
  __shared__ double sc[DIM][NFN];           // wface[k] * nrm[d,k]
  __shared__ double spen[NFN][ND];          // wface-weighted flux at the face nodes
  __shared__ double srec[2][NN][ND];
  // operator coefficients indexed by a per-thread (i, j): from shared memory (a constant-bank load with a divergent
  // index is replayed once per distinct address)
  __shared__ double sA[NN][NFN];            // interp[i][k]
  __shared__ double sB[NN][DIM][NFN];       // interp[j][nbrperm[k]] * wface[k] * nrm[d,k]
  __shared__ double swf[NFN];               // wface
  __shared__ int spL[NN], spR[NN], snbr[NFN];   // perm[:, fL], perm[:, fR], nbrperm[:, orient]
  __shared__ double sWf[2][NFN][ND], sQf[2][NFN][ND];   // face-interpolated entropy variables / their conservative states
  if (a.ctl->stop) return;
  const int tid = threadIdx.x;
  const int64_t g = a.g0 + blockIdx.x;
  const FaceRec r = a.faces[g];
  const bool interior = r.kind == FK_INTERIOR;
  // shared face (calcSharedFaceElementIntegrals_element_inner, flux.jl:442-496): elementR is a whole element of the
  // neighbour, received into q_recv (element-data halo, getSendDataElement Utils/parallel.jl:276-293; r.elR = its slot);
  // only elementL's record is kept
  const bool shared = r.kind == FK_SHARED;
  const bool two = interior || shared;
  // operator tables from their device copies (tab_dev: perm | nbrperm, optab_dev: interp | wface): a read of the
  // kernel-parameter bank with a per-thread index is replayed once per distinct address
  if (tid < NN) {
    spL[tid] = __ldg(a.tab_dev + r.fL * NN + tid);
    spR[tid] = two ? __ldg(a.tab_dev + r.fR * NN + tid) : 0;
  }
  if (tid >= 32 && tid < 32 + NFN) {
    const int k = tid - 32;
    snbr[k] = __ldg(a.tab_dev + NF * NN + (two ? r.orient : 0) * NFN + k);
    swf[k] = __ldg(a.optab_dev + NN * NFN + k);
  }
  for (int idx = tid; idx < NN * NFN; idx += T) sA[idx / NFN][idx % NFN] = __ldg(a.optab_dev + idx);
  __syncthreads();
  for (int idx = tid; idx < NN * ND; idx += T) {
    const int j = idx / ND, k = idx - j * ND;
    sq[0][j][k] = __ldg(a.q + (int64_t)r.elL * EL + spL[j] * ND + k);
    sq[1][j][k] = two ? __ldg((shared ? a.q_recv : a.q) + (int64_t)r.elR * EL + spR[j] * ND + k) : 0.0;
    srec[0][j][k] = 0.0;
    srec[1][j][k] = 0.0;
  }
  for (int idx = tid; idx < DIM * NFN; idx += T) {
    const int d = idx / NFN, k = idx - d * NFN;
    sc[d][k] = swf[k] * __ldg(a.nrm + g * a.nrm_face_stride + k * a.nrm_node_stride + d);
  }
  __syncthreads();
  const int* nbr = snbr;
  if (two) {
    for (int idx = tid; idx < NN * DIM * NFN; idx += T) {
      const int j = idx / (DIM * NFN), d = (idx / NFN) % DIM, k = idx % NFN;
      sB[j][d][k] = sA[j][nbr[k]] * sc[d][k];
    }
  }

  if (!two) {
    // interpolateBoundary + BC functor + boundaryintegrate! (bc.jl:162-175, 251-284; euler.jl:669-690)
    if (tid < NFN) {
      const int k = tid;
      double qb[ND], xb[DIM], nb_[DIM], fb[ND];
#pragma unroll
      for (int p = 0; p < ND; ++p) qb[p] = 0.0;
      for (int j = 0; j < NN; ++j) {
        const double c = sA[j][k];
#pragma unroll
        for (int p = 0; p < ND; ++p) qb[p] = fma(c, sq[0][j][p], qb[p]);
      }
      const double* xp = a.coords_bndry + ((int64_t)r.elR * NFN + k) * DIM;
#pragma unroll
      for (int d = 0; d < DIM; ++d) { xb[d] = xp[d]; nb_[d] = __ldg(a.nrm + g * a.nrm_face_stride + k * a.nrm_node_stride + d); }
      bc_flux_any<DIM>(r.aux, qb, xb, nb_, a.ph, fb);
#pragma unroll
      for (int p = 0; p < ND; ++p) spen[k][p] = swf[k] * fb[p];
    }
    __syncthreads();
    for (int idx = tid; idx < NN * ND; idx += T) {
      const int j = idx / ND, p = idx - j * ND;
      double s = 0.0;
      for (int k = 0; k < NFN; ++k) s = fma(sA[j][k], spen[k][p], s);
      srec[0][j][p] = -s;
    }
    __syncthreads();
  } else {
    if (fei == FEI_EC || fei == FEI_ESLF || fei == FEI_ESLW2) {
      for (int idx = tid; idx < 2 * NN; idx += T) sZ[idx / NN][idx % NN] = ir_node<DIM>(sq[idx / NN][idx % NN], a.ph.gamma - 1.0);
      __syncthreads();
      for (int pr = tid; pr < NN * NN; pr += T) {
        const int i = pr / NN, j = pr - i * NN;
        double dirs[DIM][DIM], F[DIM][ND];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
          for (int e = 0; e < DIM; ++e) dirs[d][e] = d == e ? 1.0 : 0.0;
        ir_flux<DIM, DIM>(sZ[0][i], sZ[1][j], dirs, a.ph.gamma, F);
        double gsum[ND];
#pragma unroll
        for (int p = 0; p < ND; ++p) gsum[p] = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          double Eij = 0.0;
#pragma unroll
          for (int k = 0; k < NFN; ++k) Eij = fma(sA[i][k], sB[j][d][k], Eij);
#pragma unroll
          for (int p = 0; p < ND; ++p) gsum[p] = fma(Eij, F[d][p], gsum[p]);
        }
#pragma unroll
        for (int p = 0; p < ND; ++p) sG[pr][p] = gsum[p];
      }
      __syncthreads();
      for (int idx = tid; idx < NN * ND; idx += T) {
        const int i = idx / ND, p = idx - i * ND;
        double sl = 0.0, sr = 0.0;
        for (int j = 0; j < NN; ++j) { sl += sG[i * NN + j][p]; sr += sG[j * NN + i][p]; }
        srec[0][i][p] = -sl;
        srec[1][i][p] = sr;
      }
      __syncthreads();
    }
    if (fei != FEI_EC) {      // ELF / ELW2 penalty, alone or on top of the entropy-conservative integral
      for (int idx = tid; idx < 2 * NN; idx += T) convert_to_ir<DIM>(sq[idx / NN][idx % NN], a.ph.gamma, sw[idx / NN][idx % NN]);
      __syncthreads();
      // entropy variables interpolated to the face nodes: one thread per (side, face node, variable) ...
      for (int idx = tid; idx < 2 * NFN * ND; idx += T) {
        const int sd = idx / (NFN * ND), k = (idx / ND) % NFN, p = idx % ND;
        const int kk = sd == 0 ? k : nbr[k];
        double w = 0.0;
        for (int j = 0; j < NN; ++j) w += sA[j][kk] * sw[sd][j][p];
        sWf[sd][k][p] = w;
      }
      __syncthreads();
      // ... their conservative states: one thread per (side, face node) ...
      if (tid < 2 * NFN) {
        const int sd = tid / NFN, k = tid - sd * NFN;
        double wv[ND], qv[ND];
#pragma unroll
        for (int p = 0; p < ND; ++p) wv[p] = sWf[sd][k][p];
        convert_from_ir<DIM>(wv, a.ph.gamma, qv);
#pragma unroll
        for (int p = 0; p < ND; ++p) sQf[sd][k][p] = qv[p];
      }
      __syncthreads();
      // ... and the Lax-Friedrichs entropy kernel: one thread per face node
      if (tid < NFN) {
        const int k = tid;
        double qa[ND], dw[ND], fl[ND], nrm[DIM];
#pragma unroll
        for (int p = 0; p < ND; ++p) { qa[p] = 0.5 * (sQf[0][k][p] + sQf[1][k][p]); dw[p] = sWf[0][k][p] - sWf[1][k][p]; }
#pragma unroll
        for (int d = 0; d < DIM; ++d) nrm[d] = __ldg(a.nrm + g * a.nrm_face_stride + k * a.nrm_node_stride + d);
        if (fei == FEI_ELW2_PENALTY || fei == FEI_ESLW2) lw2_entropy_kernel<DIM>(qa, dw, nrm, a.ph.gamma, fl);
        else lf_entropy_kernel<DIM>(qa, dw, nrm, a.ph.gamma, fl);
#pragma unroll
        for (int p = 0; p < ND; ++p) spen[k][p] = fl[p] * swf[k];
      }
      __syncthreads();
      for (int idx = tid; idx < NN * ND; idx += T) {
        const int j = idx / ND, p = idx - j * ND;
        double sl = srec[0][j][p], sr = srec[1][j][p];
        for (int k = 0; k < NFN; ++k) {
          sl -= sA[j][k] * spen[k][p];
          sr += sA[j][nbr[k]] * spen[k][p];
        }
        srec[0][j][p] = sl;
        srec[1][j][p] = sr;
      }
      __syncthreads();
    }
  }
  // records in element node order: stencil node j of face f is volume node perm[f][j]
  for (int idx = tid; idx < NN * ND; idx += T) {
    const int j = idx / ND, p = idx - j * ND;
    a.fluxe[((int64_t)r.elL * NF + r.fL) * EL + spL[j] * ND + p] = srec[0][j][p];
    if (interior) a.fluxe[((int64_t)r.elR * NF + r.fR) * EL + spR[j] * ND + p] = srec[1][j][p];
  }
}

// k_face_element_b: the same integrals, FB faces per CTA.  One face per CTA (k_face_element) leaves most of the 128 threads
// idle outside the pair stage and meets twelve block barriers per face (profiles/r1_n2_face_element_3d_p2.txt: FP64 pipe 26 %,
// barrier the top stall, 5 CTAs per SM by registers); here every stage loops over (face, item) pairs of the FB faces, so the
// barriers are paid once per FB faces and every stage has FB times the parallel work.  Same arithmetic, same summation order.
template <int DIM, int NN, int NFN, int FB>
struct FaceElemBCfg {
  static constexpr int ND = DIM + 2;
  // doubles per face: sq 2*NN*ND | sZ 2*NN*(DIM+4) | sw 2*NN*ND | sG NN*NN*ND | sc DIM*NFN | spen NFN*ND | srec 2*NN*ND |
  //                   sB NN*DIM*NFN | sWf, sQf 2 * 2*NFN*ND
  static constexpr int PER = 2 * NN * ND + 2 * NN * (DIM + 4) + 2 * NN * ND + NN * NN * ND + DIM * NFN + NFN * ND + 2 * NN * ND +
                             NN * DIM * NFN + 4 * NFN * ND;
  static constexpr size_t smem_bytes = sizeof(double) * ((size_t)FB * PER + NN * NFN + NFN) + sizeof(int) * (size_t)FB * (2 * NN + NFN + 4);
};

template <int DIM, int NN, int NFN, int FB>
__global__ void __launch_bounds__(128)
k_face_element_b(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ FaceArgs a, int fei) {
  using Cfg = FaceElemBCfg<DIM, NN, NFN, FB>;
  constexpr int ND = DIM + 2, NF = DIM + 1, EL = NN * ND, T = 128, NZ = DIM + 4;
  extern __shared__ __align__(16) unsigned char smem_feb[];
  double* base = reinterpret_cast<double*>(smem_feb);
  // per-face tiles (face-major), then the operator tables shared by all faces
  auto Fq = [&](int f) { return base + (size_t)f * Cfg::PER; };                    // [2][NN][ND]
  auto Fz = [&](int f) { return Fq(f) + 2 * NN * ND; };                              // [2][NN][NZ]
  auto Fw = [&](int f) { return Fz(f) + 2 * NN * NZ; };                              // [2][NN][ND]
  auto Fg = [&](int f) { return Fw(f) + 2 * NN * ND; };                              // [NN*NN][ND]
  auto Fc = [&](int f) { return Fg(f) + NN * NN * ND; };                             // [DIM][NFN]
  auto Fp = [&](int f) { return Fc(f) + DIM * NFN; };                                // [NFN][ND]
  auto Fr = [&](int f) { return Fp(f) + NFN * ND; };                                 // [2][NN][ND]
  auto Fb = [&](int f) { return Fr(f) + 2 * NN * ND; };                              // [NN][DIM][NFN]
  auto Fwf = [&](int f) { return Fb(f) + NN * DIM * NFN; };                          // [2][NFN][ND]
  auto Fqf = [&](int f) { return Fwf(f) + 2 * NFN * ND; };                           // [2][NFN][ND]
  double* sA = base + (size_t)FB * Cfg::PER;                                         // [NN][NFN] interp
  double* swf = sA + NN * NFN;                                                       // [NFN]     wface
  int* si = reinterpret_cast<int*>(swf + NFN);
  auto IpL = [&](int f) { return si + f * (2 * NN + NFN + 4); };                      // perm[:, fL]
  auto IpR = [&](int f) { return IpL(f) + NN; };                                      // perm[:, fR]
  auto Inb = [&](int f) { return IpR(f) + NN; };                                      // nbrperm[:, orient]
  auto Ikd = [&](int f) { return Inb(f) + NFN; };                                     // kind | -1: no face
  if (a.ctl->stop) return;
  const int tid = threadIdx.x;
  const int64_t g0 = a.g0 + (int64_t)blockIdx.x * FB;
  const int nfc = (int)((a.g0 + a.ng - g0) < FB ? (a.g0 + a.ng - g0) : FB);
  const double gami = a.ph.gamma - 1.0;

  for (int idx = tid; idx < NN * NFN + NFN; idx += T) sA[idx] = __ldg(a.optab_dev + idx);      // interp | wface (contiguous)
  for (int idx = tid; idx < nfc * (2 * NN + NFN); idx += T) {
    const int f = idx / (2 * NN + NFN), x = idx - f * (2 * NN + NFN);
    const FaceRec r = a.faces[g0 + f];
    const bool two = r.kind != FK_BOUNDARY;
    int v;
    if (x < NN) v = __ldg(a.tab_dev + r.fL * NN + x);
    else if (x < 2 * NN) v = two ? __ldg(a.tab_dev + r.fR * NN + (x - NN)) : 0;
    else v = __ldg(a.tab_dev + NF * NN + (two ? r.orient : 0) * NFN + (x - 2 * NN));
    IpL(f)[x] = v;
    if (x == 0) Ikd(f)[0] = r.kind;
  }
  __syncthreads();
  for (int idx = tid; idx < nfc * NN * ND; idx += T) {
    const int f = idx / (NN * ND), x = idx - f * (NN * ND), j = x / ND, k = x - j * ND;
    const FaceRec r = a.faces[g0 + f];
    const bool shared = r.kind == FK_SHARED, two = r.kind != FK_BOUNDARY;
    Fq(f)[x] = __ldg(a.q + (int64_t)r.elL * EL + IpL(f)[j] * ND + k);
    Fq(f)[NN * ND + x] = two ? __ldg((shared ? a.q_recv : a.q) + (int64_t)r.elR * EL + IpR(f)[j] * ND + k) : 0.0;
    Fr(f)[x] = 0.0;
    Fr(f)[NN * ND + x] = 0.0;
  }
  for (int idx = tid; idx < nfc * DIM * NFN; idx += T) {
    const int f = idx / (DIM * NFN), x = idx - f * (DIM * NFN), d = x / NFN, k = x - d * NFN;
    Fc(f)[x] = swf[k] * __ldg(a.nrm + (g0 + f) * a.nrm_face_stride + k * a.nrm_node_stride + d);
  }
  __syncthreads();
  for (int idx = tid; idx < nfc * NN * DIM * NFN; idx += T) {
    const int f = idx / (NN * DIM * NFN), x = idx - f * (NN * DIM * NFN);
    const int j = x / (DIM * NFN), d = (x / NFN) % DIM, k = x % NFN;
    Fb(f)[x] = sA[j * NFN + Inb(f)[k]] * Fc(f)[d * NFN + k];
  }
  // ---- boundary faces: interpolateBoundary + BC functor + boundaryintegrate! ------------------------------------------
  for (int idx = tid; idx < nfc * NFN; idx += T) {
    const int f = idx / NFN, k = idx - f * NFN;
    if (Ikd(f)[0] != FK_BOUNDARY) continue;
    const FaceRec r = a.faces[g0 + f];
    double qb[ND], xb[DIM], nb_[DIM], fb[ND];
#pragma unroll
    for (int p = 0; p < ND; ++p) qb[p] = 0.0;
    for (int j = 0; j < NN; ++j) {
      const double c = sA[j * NFN + k];
#pragma unroll
      for (int p = 0; p < ND; ++p) qb[p] = fma(c, Fq(f)[j * ND + p], qb[p]);
    }
    const double* xp = a.coords_bndry + ((int64_t)r.elR * NFN + k) * DIM;
#pragma unroll
    for (int d = 0; d < DIM; ++d) { xb[d] = xp[d]; nb_[d] = __ldg(a.nrm + (g0 + f) * a.nrm_face_stride + k * a.nrm_node_stride + d); }
    bc_flux_any<DIM>(r.aux, qb, xb, nb_, a.ph, fb);
#pragma unroll
    for (int p = 0; p < ND; ++p) Fp(f)[k * ND + p] = swf[k] * fb[p];
  }
  __syncthreads();
  for (int idx = tid; idx < nfc * NN * ND; idx += T) {
    const int f = idx / (NN * ND), x = idx - f * (NN * ND), j = x / ND, p = x - j * ND;
    if (Ikd(f)[0] != FK_BOUNDARY) continue;
    double s_ = 0.0;
    for (int k = 0; k < NFN; ++k) s_ = fma(sA[j * NFN + k], Fp(f)[k * ND + p], s_);
    Fr(f)[x] = -s_;
  }
  // ---- two-sided faces -------------------------------------------------------------------------------------------------
  if (fei == FEI_EC || fei == FEI_ESLF || fei == FEI_ESLW2) {
    for (int idx = tid; idx < nfc * 2 * NN; idx += T) {
      const int f = idx / (2 * NN), x = idx - f * (2 * NN);
      if (Ikd(f)[0] == FK_BOUNDARY) continue;
      const IRNode<DIM> z = ir_node<DIM>(Fq(f) + x * ND, gami);
      double* zz = Fz(f) + x * NZ;
      zz[0] = z.z1;
#pragma unroll
      for (int d = 0; d < DIM; ++d) zz[1 + d] = z.zv[d];
      zz[DIM + 1] = z.z5; zz[DIM + 2] = z.l1; zz[DIM + 3] = z.l5;
    }
    __syncthreads();
    for (int idx = tid; idx < nfc * NN * NN; idx += T) {
      const int f = idx / (NN * NN), pr = idx - f * (NN * NN), i = pr / NN, j = pr - i * NN;
      if (Ikd(f)[0] == FK_BOUNDARY) continue;
      IRNode<DIM> zi, zj;
      const double* za = Fz(f) + i * NZ;
      const double* zb = Fz(f) + (NN + j) * NZ;
      zi.z1 = za[0]; zj.z1 = zb[0];
#pragma unroll
      for (int d = 0; d < DIM; ++d) { zi.zv[d] = za[1 + d]; zj.zv[d] = zb[1 + d]; }
      zi.z5 = za[DIM + 1]; zi.l1 = za[DIM + 2]; zi.l5 = za[DIM + 3];
      zj.z5 = zb[DIM + 1]; zj.l1 = zb[DIM + 2]; zj.l5 = zb[DIM + 3];
      double dirs[DIM][DIM], F[DIM][ND];
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int e = 0; e < DIM; ++e) dirs[d][e] = d == e ? 1.0 : 0.0;
      ir_flux<DIM, DIM>(zi, zj, dirs, a.ph.gamma, F);
      double gsum[ND];
#pragma unroll
      for (int p = 0; p < ND; ++p) gsum[p] = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double Eij = 0.0;
#pragma unroll
        for (int k = 0; k < NFN; ++k) Eij = fma(sA[i * NFN + k], Fb(f)[(j * DIM + d) * NFN + k], Eij);
#pragma unroll
        for (int p = 0; p < ND; ++p) gsum[p] = fma(Eij, F[d][p], gsum[p]);
      }
#pragma unroll
      for (int p = 0; p < ND; ++p) Fg(f)[pr * ND + p] = gsum[p];
    }
    __syncthreads();
    for (int idx = tid; idx < nfc * NN * ND; idx += T) {
      const int f = idx / (NN * ND), x = idx - f * (NN * ND), i = x / ND, p = x - i * ND;
      if (Ikd(f)[0] == FK_BOUNDARY) continue;
      double sl = 0.0, sr = 0.0;
      for (int j = 0; j < NN; ++j) { sl += Fg(f)[(i * NN + j) * ND + p]; sr += Fg(f)[(j * NN + i) * ND + p]; }
      Fr(f)[x] = -sl;
      Fr(f)[NN * ND + x] = sr;
    }
    __syncthreads();
  }
  if (fei != FEI_EC) {      // ELF / ELW2 penalty, alone or on top of the entropy-conservative integral
    for (int idx = tid; idx < nfc * 2 * NN; idx += T) {
      const int f = idx / (2 * NN), x = idx - f * (2 * NN);
      if (Ikd(f)[0] == FK_BOUNDARY) continue;
      convert_to_ir<DIM>(Fq(f) + x * ND, a.ph.gamma, Fw(f) + x * ND);
    }
    __syncthreads();
    for (int idx = tid; idx < nfc * 2 * NFN * ND; idx += T) {
      const int f = idx / (2 * NFN * ND), x = idx - f * (2 * NFN * ND);
      if (Ikd(f)[0] == FK_BOUNDARY) continue;
      const int sd = x / (NFN * ND), k = (x / ND) % NFN, p = x % ND;
      const int kk = sd == 0 ? k : Inb(f)[k];
      double w = 0.0;
      for (int j = 0; j < NN; ++j) w += sA[j * NFN + kk] * Fw(f)[(sd * NN + j) * ND + p];
      Fwf(f)[x] = w;
    }
    __syncthreads();
    for (int idx = tid; idx < nfc * 2 * NFN; idx += T) {
      const int f = idx / (2 * NFN), x = idx - f * (2 * NFN);
      if (Ikd(f)[0] == FK_BOUNDARY) continue;
      double wv[ND], qv[ND];
#pragma unroll
      for (int p = 0; p < ND; ++p) wv[p] = Fwf(f)[x * ND + p];
      convert_from_ir<DIM>(wv, a.ph.gamma, qv);
#pragma unroll
      for (int p = 0; p < ND; ++p) Fqf(f)[x * ND + p] = qv[p];
    }
    __syncthreads();
    for (int idx = tid; idx < nfc * NFN; idx += T) {
      const int f = idx / NFN, k = idx - f * NFN;
      if (Ikd(f)[0] == FK_BOUNDARY) continue;
      double qa[ND], dw[ND], fl[ND], nrm[DIM];
#pragma unroll
      for (int p = 0; p < ND; ++p) {
        qa[p] = 0.5 * (Fqf(f)[k * ND + p] + Fqf(f)[(NFN + k) * ND + p]);
        dw[p] = Fwf(f)[k * ND + p] - Fwf(f)[(NFN + k) * ND + p];
      }
#pragma unroll
      for (int d = 0; d < DIM; ++d) nrm[d] = __ldg(a.nrm + (g0 + f) * a.nrm_face_stride + k * a.nrm_node_stride + d);
      if (fei == FEI_ELW2_PENALTY || fei == FEI_ESLW2) lw2_entropy_kernel<DIM>(qa, dw, nrm, a.ph.gamma, fl);
      else lf_entropy_kernel<DIM>(qa, dw, nrm, a.ph.gamma, fl);
#pragma unroll
      for (int p = 0; p < ND; ++p) Fp(f)[k * ND + p] = fl[p] * swf[k];
    }
    __syncthreads();
    for (int idx = tid; idx < nfc * NN * ND; idx += T) {
      const int f = idx / (NN * ND), x = idx - f * (NN * ND), j = x / ND, p = x - j * ND;
      if (Ikd(f)[0] == FK_BOUNDARY) continue;
      double sl = Fr(f)[x], sr = Fr(f)[NN * ND + x];
      for (int k = 0; k < NFN; ++k) {
        sl -= sA[j * NFN + k] * Fp(f)[k * ND + p];
        sr += sA[j * NFN + Inb(f)[k]] * Fp(f)[k * ND + p];
      }
      Fr(f)[x] = sl;
      Fr(f)[NN * ND + x] = sr;
    }
  }
  __syncthreads();
  // records in element node order: stencil node j of face f is volume node perm[f][j]
  for (int idx = tid; idx < nfc * NN * ND; idx += T) {
    const int f = idx / (NN * ND), x = idx - f * (NN * ND), j = x / ND, p = x - j * ND;
    const FaceRec r = a.faces[g0 + f];
    a.fluxe[((int64_t)r.elL * NF + r.fL) * EL + IpL(f)[j] * ND + p] = Fr(f)[x];
    if (r.kind == FK_INTERIOR) a.fluxe[((int64_t)r.elR * NF + r.fR) * EL + IpR(f)[j] * ND + p] = Fr(f)[NN * ND + x];
  }
}

template <int DIM, int NN, int NFN, int E>
struct SplitCfg {
  static constexpr int ND = DIM + 2, NF = DIM + 1;
  static constexpr int NP = NN * (NN - 1) / 2;                   // node pairs
  static constexpr int NZ = DIM + 4;                             // z1, zv[DIM], z5, log z1, log z5
  static constexpr int NC = DIM * ND;                            // flux components per pair
  // the pair stage dominates: T divides E*NP (E=16, nn=12: 1056 = 3 * 352) so that every thread evaluates the same
  // number of two-point fluxes
  static constexpr int T = (E * NP) % 352 == 0 ? 352 : 192;
  static constexpr int PS = E * NP + 1;                          // component stride of the pair-flux tile (odd)
  static constexpr int ZS = E * NN + 1;                          // component stride of the node tile (odd)
  static constexpr size_t smem_bytes =
      sizeof(double) * ((size_t)E * NN * ND + (size_t)NZ * ZS + (size_t)NC * PS + DIM * NN * NN);
  static_assert(E % 2 == 0, "tile bases must stay 16-byte aligned");
};

template <int DIM, int NN, int NFN, int E, int MODE, bool DENSEREC = false>
__global__ void __launch_bounds__((SplitCfg<DIM, NN, NFN, E>::T), 2)
k_element_split(const __grid_constant__ OpTabS<DIM, NN, NFN> op, const __grid_constant__ ElemArgs a) {
  using Cfg = SplitCfg<DIM, NN, NFN, E>;
  constexpr int ND = Cfg::ND, NF = Cfg::NF, NP = Cfg::NP, NZ = Cfg::NZ, NC = Cfg::NC, T = Cfg::T;
  constexpr int PS = Cfg::PS, ZS = Cfg::ZS;
  constexpr int EL = NN * ND, FL = NFN * ND;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sq = reinterpret_cast<double*>(smem_raw);     // [E][EL]       q tile, later the staged output
  double* sZ = sq + E * EL;                             // [NZ][E*NN]    component-major (conflict-free)
  double* sFp = sZ + NZ * ZS;                           // [NC][E*NP]    component-major
  double* sS2 = sFp + NC * PS;                          // [DIM][NN][NN]
  __shared__ double s_red[T / 32];
  __shared__ unsigned char s_pj[NP], s_pk[NP];          // pair -> (j, k), j > k
  __shared__ unsigned char s_pidx[NN][NN];              // (i, m) -> pair index

  if (a.ctl->stop) return;
  const int tid = threadIdx.x;
  const int64_t e0 = a.e_begin + (int64_t)blockIdx.x * E;
  const int ne = (int)((a.nE - e0) < E ? (a.nE - e0) : E);
  const double gami = a.ph.gamma - 1.0;

  async_tile(sq, a.q + e0 * EL, ne * EL, tid, T);
  cp_async_commit();
  for (int idx = tid; idx < DIM * NN * NN; idx += T) sS2[idx] = (&op.S2[0][0][0])[idx];
  for (int idx = tid; idx < NN * NN; idx += T) {
    const int i = idx / NN, m = idx - i * NN;
    const int jj = m > i ? m : i, kk = m > i ? i : m;
    const int pr = jj * (jj - 1) / 2 + kk;
    s_pidx[i][m] = (unsigned char)(i == m ? 0 : pr);
    if (i > m) { s_pj[pr] = (unsigned char)i; s_pk[pr] = (unsigned char)m; }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- node threads: checks + Ismail-Roe parameter vector and its logarithms ---------------------------------
  for (int it = tid; it < ne * NN; it += T) {
    const int s = it / NN, j = it - s * NN;
    double qn[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) qn[k] = sq[it * ND + k];
    const double press = calc_pressure<DIM>(qn, gami);
    if ((a.ph.check_density && !(qn[0] > 0.0)) || (a.ph.check_pressure && !(press > 0.0))) {
      const int code = (a.ph.check_density && !(qn[0] > 0.0)) ? 1 : 2;
      const unsigned long long loc = ((unsigned long long)(e0 + s) << 8) | (unsigned)j;
      atomicMin(&a.ctl->err_loc, ((unsigned long long)(code - 1) << 62) | loc);
      atomicExch(&a.ctl->err_code, 1);
      atomicExch(&a.ctl->stop, 1);
      qn[0] = 1.0; qn[DIM + 1] = 1.0;       // keep the arithmetic finite; the result is discarded
#pragma unroll
      for (int d = 0; d < DIM; ++d) qn[1 + d] = 0.0;
    }
    const IRNode<DIM> z = ir_node<DIM>(qn, gami);
    sZ[0 * ZS + it] = z.z1;
#pragma unroll
    for (int d = 0; d < DIM; ++d) sZ[(1 + d) * ZS + it] = z.zv[d];
    sZ[(DIM + 1) * ZS + it] = z.z5; sZ[(DIM + 2) * ZS + it] = z.l1; sZ[(DIM + 3) * ZS + it] = z.l5;
  }
  __syncthreads();

  // ---- pair threads: F_d(q_j, q_k) for k < j in the DIM parametric directions of node j ------------------------
  for (int it = tid; it < ne * NP; it += T) {
    const int s = it / NP, pr = it - s * NP;
    const int j = s_pj[pr], k = s_pk[pr];            // pr = j(j-1)/2 + k, j > k
    IRNode<DIM> zj, zk;
    const int nj = s * NN + j, nk = s * NN + k;
    zj.z1 = sZ[nj]; zk.z1 = sZ[nk];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { zj.zv[d] = sZ[(1 + d) * ZS + nj]; zk.zv[d] = sZ[(1 + d) * ZS + nk]; }
    zj.z5 = sZ[(DIM + 1) * ZS + nj]; zj.l1 = sZ[(DIM + 2) * ZS + nj]; zj.l5 = sZ[(DIM + 3) * ZS + nj];
    zk.z5 = sZ[(DIM + 1) * ZS + nk]; zk.l1 = sZ[(DIM + 2) * ZS + nk]; zk.l5 = sZ[(DIM + 3) * ZS + nk];
    const double* dx = a.dxidx + (e0 + s) * a.dx_el_stride + j * a.dx_node_stride;
    double dirs[DIM][DIM], F[DIM][ND];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int p = 0; p < DIM; ++p) dirs[d][p] = __ldg(dx + d + DIM * p);
    ir_flux<DIM, DIM>(zj, zk, dirs, a.ph.gamma, F);
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int c = 0; c < ND; ++c) sFp[(d * ND + c) * PS + it] = F[d][c];
  }
  __syncthreads();

  // ---- (element, node, variable) threads: res[c,i] = -sum_m 2 S[i,m,d] F_d[c](pair(i,m)) + face records, x Minv ----
  for (int it = tid; it < ne * EL; it += T) {
    const int s = it / EL, r = it - s * EL;
    const int i = r / ND, c = r - i * ND;
    // the face records and Minv are requested before the pair-flux gather (they were the top stall of this kernel:
    // profiles/r1_es_element_split.txt, long scoreboard at the add that consumed them)
    double grec[DENSEREC ? NF : DIM];
    if (DENSEREC) {
      // face-element integrals (k_face_element): one [nd, nn] record per local face
      const double* G = a.fluxe + (e0 + s) * (NF * EL) + r;
#pragma unroll
      for (int f = 0; f < NF; ++f) grec[f] = __ldg(G + f * EL);
    } else {
      const double* G = a.fluxe + (e0 + s) * (NF * FL);
#pragma unroll
      for (int u = 0; u < DIM; ++u) {
        const int slot = op.inv[i][u];
        grec[u] = slot >= 0 ? __ldg(G + slot * ND + c) : 0.0;
      }
    }
    const double mv = MODE == EPI_RK ? __ldg(a.minv + (e0 + s) * NN + i) : 1.0;
    double acc = 0.0;
    const double* Fc = sFp + c * PS + s * NP;
    const double* Sc = sS2 + i * NN;
#pragma unroll
    for (int m = 0; m < NN; ++m) {
      // S2[i][i] = 0, so the diagonal term (which reads pair 0) contributes nothing
      const int pi = s_pidx[i][m];
#pragma unroll
      for (int d = 0; d < DIM; ++d) acc = fma(-Sc[d * NN * NN + m], Fc[d * ND * PS + pi], acc);
    }
#pragma unroll
    for (int u = 0; u < (DENSEREC ? NF : DIM); ++u) acc += grec[u];
    if (MODE == EPI_RK) acc *= mv;
    sq[it] = acc;
  }
  __syncthreads();
  epilogue_tile<NN, ND, E, T, MODE>(a, sq, ne, e0, tid, s_red);
}

// ------------------------------------------------------------------------------------------------------
// k_element_split_n: node-centric form of k_element_split.  One thread per (element, node i) evaluates the nn-1
// two-point fluxes F(q_i, q_m) of ITS node and accumulates -2 S[i,m,d] F_d in registers.  Every flux is evaluated
// twice (once per end point), but nothing is exchanged between threads: no pair-flux tile (68 of the 85 KB of
// k_element_split), no gather stage (60 shared-memory loads per residual entry), one block barrier less; the kernel's
// instructions become mostly the FP64 work itself (k_element_split: 21 % FP64-pipe utilisation, barrier and
// shared-memory stalls on top, profiles/r1_es_element_split.txt).  The Ismail-Roe flux is bitwise symmetric in its two
// states and the sum over m runs in the same order, so the result equals k_element_split's bit for bit.
// ------------------------------------------------------------------------------------------------------
template <int DIM, int NN, int NFN, int E>
struct SplitNCfg {
  static constexpr int ND = DIM + 2, NF = DIM + 1;
  static constexpr int NZ = DIM + 4;                             // z1, zv[DIM], z5, log z1, log z5
  static constexpr int T = ((E * NN + 31) / 32) * 32;
  static constexpr int ZS = E * NN + 1;                          // component stride of the node tile (odd)
  static constexpr size_t smem_bytes = sizeof(double) * ((size_t)E * NN * ND + (size_t)NZ * ZS + DIM * NN * NN);
  static_assert(E % 2 == 0, "tile bases must stay 16-byte aligned");
};

template <int DIM, int NN, int NFN, int E, int MODE, bool DENSEREC, int MINB>
__global__ void __launch_bounds__((SplitNCfg<DIM, NN, NFN, E>::T), MINB)
k_element_split_n(const __grid_constant__ OpTabS<DIM, NN, NFN> op, const __grid_constant__ ElemArgs a) {
  using Cfg = SplitNCfg<DIM, NN, NFN, E>;
  constexpr int ND = Cfg::ND, NF = Cfg::NF, NZ = Cfg::NZ, T = Cfg::T, ZS = Cfg::ZS;
  constexpr int EL = NN * ND, FL = NFN * ND;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sq = reinterpret_cast<double*>(smem_raw);     // [E][EL]       q tile, later the staged output
  double* sZ = sq + E * EL;                             // [NZ][E*NN]    component-major (conflict-free)
  double* sS2 = sZ + NZ * ZS;                           // [DIM][NN][NN]
  __shared__ double s_red[T / 32];

  if (a.ctl->stop) return;
  const int tid = threadIdx.x;
  const int64_t e0 = a.e_begin + (int64_t)blockIdx.x * E;
  const int ne = (int)((a.nE - e0) < E ? (a.nE - e0) : E);
  const double gami = a.ph.gamma - 1.0;

  async_tile(sq, a.q + e0 * EL, ne * EL, tid, T);
  cp_async_commit();
  for (int idx = tid; idx < DIM * NN * NN; idx += T) sS2[idx] = (&op.S2[0][0][0])[idx];
  const bool act = tid < ne * NN;
  const int s = tid / NN, i = tid - s * NN;
  // face records, Minv and (node-independent) metrics of this thread's node: in flight during the node stage
  double grec[(DENSEREC ? NF : DIM) * ND], mv = 1.0, dx0[DIM][DIM];
  if (act) {
    if (DENSEREC) {
      const double* G = a.fluxe + (e0 + s) * (NF * EL) + i * ND;
#pragma unroll
      for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int c = 0; c < ND; ++c) grec[f * ND + c] = __ldg(G + f * EL + c);
    } else {
      const double* G = a.fluxe + (e0 + s) * (NF * FL);
#pragma unroll
      for (int u = 0; u < DIM; ++u) {
        const int slot = op.inv[i][u];
#pragma unroll
        for (int c = 0; c < ND; ++c) grec[u * ND + c] = slot >= 0 ? __ldg(G + slot * ND + c) : 0.0;
      }
    }
    if (MODE == EPI_RK) mv = __ldg(a.minv + (e0 + s) * NN + i);
    const double* dx = a.dxidx + (e0 + s) * a.dx_el_stride + i * a.dx_node_stride;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int p = 0; p < DIM; ++p) dx0[d][p] = __ldg(dx + d + DIM * p);
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- node stage: checks + Ismail-Roe parameter vector and its logarithms (kept in registers and published) ----
  IRNode<DIM> zi;
  if (act) {
    double qn[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) qn[k] = sq[tid * ND + k];
    const double press = calc_pressure<DIM>(qn, gami);
    if ((a.ph.check_density && !(qn[0] > 0.0)) || (a.ph.check_pressure && !(press > 0.0))) {
      const int code = (a.ph.check_density && !(qn[0] > 0.0)) ? 1 : 2;
      const unsigned long long loc = ((unsigned long long)(e0 + s) << 8) | (unsigned)i;
      atomicMin(&a.ctl->err_loc, ((unsigned long long)(code - 1) << 62) | loc);
      atomicExch(&a.ctl->err_code, 1);
      atomicExch(&a.ctl->stop, 1);
      qn[0] = 1.0; qn[DIM + 1] = 1.0;       // keep the arithmetic finite; the result is discarded
#pragma unroll
      for (int d = 0; d < DIM; ++d) qn[1 + d] = 0.0;
    }
    zi = ir_node<DIM>(qn, gami);
    sZ[0 * ZS + tid] = zi.z1;
#pragma unroll
    for (int d = 0; d < DIM; ++d) sZ[(1 + d) * ZS + tid] = zi.zv[d];
    sZ[(DIM + 1) * ZS + tid] = zi.z5; sZ[(DIM + 2) * ZS + tid] = zi.l1; sZ[(DIM + 3) * ZS + tid] = zi.l5;
  }
  __syncthreads();       // node tile complete; nobody reads the q tile any more (it becomes the output staging tile)

  // ---- res[:, i] = -sum_m 2 S[i,m,d] F_d(q_max(i,m), q_min(i,m)) with the directions of node max(i,m) -----------
  if (act) {
    double acc[ND];
#pragma unroll
    for (int c = 0; c < ND; ++c) acc[c] = 0.0;
    const double* Sc = sS2 + i * NN;
    const int n0 = s * NN;
    // (partners enumerated without the node itself -- S[i][i] = 0 -- so that the loop is branch-free and can be
    // unrolled: independent two-point fluxes interleave, the flux itself being one long dependent chain)
    constexpr int UNR = PDES_SPLITN_UNROLL;
#pragma unroll UNR
    for (int mm = 0; mm < NN - 1; ++mm) {
      const int m = mm + (mm >= i ? 1 : 0);
      IRNode<DIM> zm;
      zm.z1 = sZ[n0 + m];
#pragma unroll
      for (int d = 0; d < DIM; ++d) zm.zv[d] = sZ[(1 + d) * ZS + n0 + m];
      zm.z5 = sZ[(DIM + 1) * ZS + n0 + m]; zm.l1 = sZ[(DIM + 2) * ZS + n0 + m]; zm.l5 = sZ[(DIM + 3) * ZS + n0 + m];
      double dirs[DIM][DIM], F[DIM][ND];
      if (a.dx_node_stride != 0 && m > i) {
        // curved elements: the two-point flux of the pair uses the metrics of its higher-numbered node
        const double* dx = a.dxidx + (e0 + s) * a.dx_el_stride + m * a.dx_node_stride;
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
          for (int p = 0; p < DIM; ++p) dirs[d][p] = __ldg(dx + d + DIM * p);
      } else {
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
          for (int p = 0; p < DIM; ++p) dirs[d][p] = dx0[d][p];
      }
      if (m > i) ir_flux<DIM, DIM>(zm, zi, dirs, a.ph.gamma, F);
      else ir_flux<DIM, DIM>(zi, zm, dirs, a.ph.gamma, F);
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const double sc = -Sc[d * NN * NN + m];
#pragma unroll
        for (int c = 0; c < ND; ++c) acc[c] = fma(sc, F[d][c], acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < ND; ++c) {
#pragma unroll
      for (int u = 0; u < (DENSEREC ? NF : DIM); ++u) acc[c] += grec[u * ND + c];
      if (MODE == EPI_RK) acc[c] *= mv;
      sq[tid * ND + c] = acc[c];
    }
  }
  __syncthreads();
  epilogue_tile<NN, ND, E, T, MODE>(a, sq, ne, e0, tid, s_red);
}

// ------------------------------------------------------------------------------------------------------
// k_element_split_r: k_element_split_n with every two-point flux evaluated ONCE.  The node-centric kernel is FP64-pipe
// bound (68 % busy) and evaluates each flux at both of its end points; here the nn(nn-1)/2 pairs of an element are
// scheduled in nn/2 rounds of disjoint pairs (round k: {i, i+k mod nn}, a round-robin tournament): the thread of node i
// evaluates the flux, keeps its share and passes the flux to node i+k through a double-buffered shared-memory tile
// (one barrier per round, a regular access pattern: no index tables, no separate gather stage).  The sum over the partner
// nodes runs in round order instead of node order: results differ from k_element_split(_n) by rounding only.
// ------------------------------------------------------------------------------------------------------
template <int DIM, int NN, int NFN, int E>
struct SplitRCfg {
  static constexpr int ND = DIM + 2, NF = DIM + 1;
  static constexpr int NZ = DIM + 4;                             // z1, zv[DIM], z5, log z1, log z5
  static constexpr int T = ((E * NN + 31) / 32) * 32;
  static constexpr int ZS = E * NN + 1;                          // component stride of the node tile (odd)
  static constexpr int NC = DIM * ND;                            // flux components per pair
  static constexpr int XS = E * NN + 1;                          // component stride of the exchange tile (odd)
  static constexpr size_t smem_bytes =
      sizeof(double) * ((size_t)E * NN * ND + (size_t)NZ * ZS + DIM * NN * NN + 2 * (size_t)NC * XS);
  static_assert(E % 2 == 0, "tile bases must stay 16-byte aligned");
};

template <int DIM, int NN, int NFN, int E, int MODE, bool DENSEREC, int MINB>
__global__ void __launch_bounds__((SplitRCfg<DIM, NN, NFN, E>::T), MINB)
k_element_split_r(const __grid_constant__ OpTabS<DIM, NN, NFN> op, const __grid_constant__ ElemArgs a) {
  using Cfg = SplitRCfg<DIM, NN, NFN, E>;
  constexpr int ND = Cfg::ND, NF = Cfg::NF, NZ = Cfg::NZ, T = Cfg::T, ZS = Cfg::ZS, NC = Cfg::NC, XS = Cfg::XS;
  constexpr int EL = NN * ND, FL = NFN * ND;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sq = reinterpret_cast<double*>(smem_raw);     // [E][EL]       q tile, later the staged output
  double* sZ = sq + E * EL;                             // [NZ][E*NN]    component-major (conflict-free)
  double* sS2 = sZ + NZ * ZS;                           // [DIM][NN][NN]
  double* sX = sS2 + DIM * NN * NN;                     // [2][NC][E*NN]  pair fluxes on their way to the other end point
  __shared__ double s_red[T / 32];

  if (a.ctl->stop) return;
  const int tid = threadIdx.x;
  const int64_t e0 = a.e_begin + (int64_t)blockIdx.x * E;
  const int ne = (int)((a.nE - e0) < E ? (a.nE - e0) : E);
  const double gami = a.ph.gamma - 1.0;

  async_tile(sq, a.q + e0 * EL, ne * EL, tid, T);
  cp_async_commit();
  // (the operator table comes from its device copy: an indexed read of the kernel-parameter bank per entry held 8 % of
  // the kernel's stall samples once a CTA had only 8 elements to amortise it over)
  for (int idx = tid; idx < DIM * NN * NN; idx += T) cp_async8(sS2 + idx, a.s2_dev + idx);
  cp_async_commit();
  const bool act = tid < ne * NN;
  const int s = tid / NN, i = tid - s * NN;
  // face records, Minv and (node-independent) metrics of this thread's node: in flight during the node stage
  double grec[(DENSEREC ? NF : DIM) * ND], mv = 1.0, dx0[DIM][DIM];
  if (act) {
    if (DENSEREC) {
      const double* G = a.fluxe + (e0 + s) * (NF * EL) + i * ND;
#pragma unroll
      for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int c = 0; c < ND; ++c) grec[f * ND + c] = __ldg(G + f * EL + c);
    } else {
      const double* G = a.fluxe + (e0 + s) * (NF * FL);
#pragma unroll
      for (int u = 0; u < DIM; ++u) {
        const int slot = __ldg(a.inv_dev + i * DIM + u);
#pragma unroll
        for (int c = 0; c < ND; ++c) grec[u * ND + c] = slot >= 0 ? __ldg(G + slot * ND + c) : 0.0;
      }
    }
    if (MODE == EPI_RK) mv = __ldg(a.minv + (e0 + s) * NN + i);
    const double* dx = a.dxidx + (e0 + s) * a.dx_el_stride + i * a.dx_node_stride;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int p = 0; p < DIM; ++p) dx0[d][p] = __ldg(dx + d + DIM * p);
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- node stage: checks + Ismail-Roe parameter vector and its logarithms (kept in registers and published) ----
  IRNode<DIM> zi;
  if (act) {
    double qn[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) qn[k] = sq[tid * ND + k];
    const double press = calc_pressure<DIM>(qn, gami);
    if ((a.ph.check_density && !(qn[0] > 0.0)) || (a.ph.check_pressure && !(press > 0.0))) {
      const int code = (a.ph.check_density && !(qn[0] > 0.0)) ? 1 : 2;
      const unsigned long long loc = ((unsigned long long)(e0 + s) << 8) | (unsigned)i;
      atomicMin(&a.ctl->err_loc, ((unsigned long long)(code - 1) << 62) | loc);
      atomicExch(&a.ctl->err_code, 1);
      atomicExch(&a.ctl->stop, 1);
      qn[0] = 1.0; qn[DIM + 1] = 1.0;       // keep the arithmetic finite; the result is discarded
#pragma unroll
      for (int d = 0; d < DIM; ++d) qn[1 + d] = 0.0;
    }
    zi = ir_node<DIM>(qn, gami);
    sZ[0 * ZS + tid] = zi.z1;
#pragma unroll
    for (int d = 0; d < DIM; ++d) sZ[(1 + d) * ZS + tid] = zi.zv[d];
    sZ[(DIM + 1) * ZS + tid] = zi.z5; sZ[(DIM + 2) * ZS + tid] = zi.l1; sZ[(DIM + 3) * ZS + tid] = zi.l5;
  }
  __syncthreads();       // node tile complete; nobody reads the q tile any more (it becomes the output staging tile)

  // ---- rounds k = 1 .. nn/2: thread (s, i) evaluates the pair {i, i+k mod nn} ONCE, adds its own share and hands the flux
  // to the other end point through shared memory (double-buffered: one barrier per round).  For even nn the last round
  // holds each pair twice (i+k+k = i): only the threads i < nn/2 evaluate it.
  double acc[ND];
#pragma unroll
  for (int c = 0; c < ND; ++c) acc[c] = 0.0;
  {
    constexpr int NR = NN / 2;
    const double* Sc = sS2 + i * NN;
    const int n0 = s * NN;
#pragma unroll 1
    for (int k = 1; k <= NR; ++k) {
      double* xb = sX + (k & 1) * (NC * XS);
      const bool half = (NN % 2 == 0) && (k == NR);
      int m = i + k;
      if (m >= NN) m -= NN;
      if (act && (!half || i < NR)) {
        IRNode<DIM> zm;
        zm.z1 = sZ[n0 + m];
#pragma unroll
        for (int d = 0; d < DIM; ++d) zm.zv[d] = sZ[(1 + d) * ZS + n0 + m];
        zm.z5 = sZ[(DIM + 1) * ZS + n0 + m]; zm.l1 = sZ[(DIM + 2) * ZS + n0 + m]; zm.l5 = sZ[(DIM + 3) * ZS + n0 + m];
        double dirs[DIM][DIM], F[DIM][ND];
        if (a.dx_node_stride != 0 && m > i) {
          // curved elements: the two-point flux of the pair uses the metrics of its higher-numbered node
          const double* dx = a.dxidx + (e0 + s) * a.dx_el_stride + m * a.dx_node_stride;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int p = 0; p < DIM; ++p) dirs[d][p] = __ldg(dx + d + DIM * p);
        } else {
#pragma unroll
          for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int p = 0; p < DIM; ++p) dirs[d][p] = dx0[d][p];
        }
        if (m > i) ir_flux<DIM, DIM>(zm, zi, dirs, a.ph.gamma, F);
        else ir_flux<DIM, DIM>(zi, zm, dirs, a.ph.gamma, F);
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          const double sc = -Sc[d * NN * NN + m];
#pragma unroll
          for (int c = 0; c < ND; ++c) {
            acc[c] = fma(sc, F[d][c], acc[c]);
            xb[(d * ND + c) * XS + tid] = F[d][c];
          }
        }
      }
      __syncthreads();
      int ip = i - k;
      if (ip < 0) ip += NN;
      if (act && (!half || ip < NR)) {
        const double* xr = xb + n0 + ip;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          const double sc = -Sc[d * NN * NN + ip];
#pragma unroll
          for (int c = 0; c < ND; ++c) acc[c] = fma(sc, xr[(d * ND + c) * XS], acc[c]);
        }
      }
    }
  }
  if (act) {
#pragma unroll
    for (int c = 0; c < ND; ++c) {
#pragma unroll
      for (int u = 0; u < (DENSEREC ? NF : DIM); ++u) acc[c] += grec[u * ND + c];
      if (MODE == EPI_RK) acc[c] *= mv;
      sq[tid * ND + c] = acc[c];
    }
  }
  __syncthreads();
  epilogue_tile<NN, ND, E, T, MODE>(a, sq, ne, e0, tid, s_red);
}

}  // namespace pdes
