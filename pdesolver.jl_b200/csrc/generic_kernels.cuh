// Operator-size generic kernels: the catch-all behind the tuned instantiations.
//
// createSBPOperator (src/solver/common.jl:276-390) can hand the physics module an operator of ANY degree and family
// (getTriSBPOmega0 / getTetSBPOmega / getTriSBPGamma / getTetSBPGamma: dense face interpolation with a stencil of
// `stencilsize` volume nodes; getTriSBPDiagE / getTetSBPDiagE: sparse faces).  The tuned kernels are compiled for the node
// counts of the named configurations; every other (dim, numnodes, numfacenodes, stencilsize) is served here, with the
// operator tables in global memory and all sizes run-time values.  Same decomposition as the tuned path -- one flux per face
// node written as per-(element, local face) records, then one thread per (element, node) -- same node-level functions
// (euler_device.cuh), same epilogue semantics (RK4 with the running sum, LSERK54, stage-1 norm partials), no atomics.
// Written for generality, not speed (the volume kernel re-evaluates the Euler flux of every node of the element per thread).
//
//   dense faces, Roe flux     k_gen_face<DIM, T>  + k_gen_element<DIM, MODE>        (T = Dual: the tangent records of J*v)
//                                                   k_gen_jvp_element<DIM>
//   sparse faces, split form  k_gen_face_sparse<DIM, T> + k_gen_element_split<DIM, MODE>, k_gen_jvp_element_split<DIM>
//                             (calcVolumeIntegralsSplitFormLinear euler_funcs.jl:240-288, IR / IRSLF / Roe interface flux)
//   halo                      k_gen_pack (getSendDataFace, Utils/parallel.jl:249-258) on any vector (q, or the J*v direction)
#pragma once
#include <type_traits>
#include "residual_kernels.cuh"
#include "es_kernels.cuh"
#include "jvp_kernels.cuh"

namespace pdes {

struct GenTab {
  int32_t nn, nfn, ss, nor, sparse;
  const double* Qt;        // [dim*nn][nn]      Qt[(d*nn+j)*nn + i] = sbp.Q[j,i,d]
  const double* RfN;       // [(dim+1)*nfn][nn] RfN[(f*nfn+i)*nn + node] = sum_j interp[j,i] [perm[j,f] == node]   (dense faces)
  const double* interp;    // [ss][nfn]         interp[j*nfn + i] = sbpface.interp[j,i]                            (dense faces)
  const double* wface;     // [nfn]
  const int32_t* perm;     // dense: [dim+1][ss] perm[f*ss + j]; sparse: [dim+1][nfn] perm[f*nfn + i]   (0-based volume nodes)
  const int32_t* nbrperm;  // [nor][nfn]
  const double* S2;        // [dim][nn][nn]     2 S[i,m,d] = Q[i,m,d] - Q[m,i,d]                                   (split form)
  const int32_t* inv;      // [nn][dim+1]       face-node slots (f*nfn + i) that coincide with volume node n, or -1 (sparse faces)
};

template <typename T> struct GenScal;
template <> struct GenScal<double> { static __device__ __forceinline__ double out(double x) { return x; } };
template <> struct GenScal<Dual> { static __device__ __forceinline__ double out(const Dual& x) { return x.d; } };

// one thread per (face, face node).  T = double: records of the residual; T = Dual: records of its directional derivative
// along v (v_recv: the direction on the neighbours' side of the shared faces).
template <int DIM, typename T>
__global__ void __launch_bounds__(128)
k_gen_face(const GenTab op, const __grid_constant__ FaceArgs a, const double* __restrict__ v, const double* __restrict__ v_recv) {
  constexpr int ND = DIM + 2, NF = DIM + 1;
  constexpr bool DUAL = !std::is_same<T, double>::value;
  if (a.ctl->stop || a.ctl->kry_done) return;
  const int nn = op.nn, nfn = op.nfn, ss = op.ss, EL = nn * ND, FL = nfn * ND;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.ng * nfn) return;
  const int64_t g = a.g0 + t / nfn;
  const int i = (int)(t % nfn);
  const FaceRec r = a.faces[g];
  T qL[ND], qR[ND], flux[ND];
  double nrm[DIM];
  {
    const double* np_ = a.nrm + g * a.nrm_face_stride + i * a.nrm_node_stride;
#pragma unroll
    for (int d = 0; d < DIM; ++d) nrm[d] = np_[d];
#pragma unroll
    for (int k = 0; k < ND; ++k) qL[k] = T(0.0);
    const int64_t b = (int64_t)r.elL * EL;
    for (int j = 0; j < ss; ++j) {
      const double c = op.interp[j * nfn + i];
      const int64_t o = b + (int64_t)op.perm[r.fL * ss + j] * ND;
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        if constexpr (DUAL) { qL[k].v = fma(c, a.q[o + k], qL[k].v); qL[k].d = fma(c, v[o + k], qL[k].d); }
        else qL[k] = fma(c, a.q[o + k], qL[k]);
      }
    }
  }
  int iR = i;
  if (r.kind == FK_BOUNDARY) {
    const double* xp = a.coords_bndry + ((int64_t)r.elR * nfn + i) * DIM;
    double xb[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) xb[d] = xp[d];
    if constexpr (DUAL) bc_flux_dual<DIM>(r.aux, qL, xb, nrm, a.ph, flux);
    else bc_flux_any<DIM>(r.aux, qL, xb, nrm, a.ph, flux);
  } else {
    iR = op.nbrperm[r.orient * nfn + i];
#pragma unroll
    for (int k = 0; k < ND; ++k) qR[k] = T(0.0);
    if (r.kind == FK_INTERIOR) {
      const int64_t b = (int64_t)r.elR * EL;
      for (int j = 0; j < ss; ++j) {
        const double c = op.interp[j * nfn + iR];
        const int64_t o = b + (int64_t)op.perm[r.fR * ss + j] * ND;
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          if constexpr (DUAL) { qR[k].v = fma(c, a.q[o + k], qR[k].v); qR[k].d = fma(c, v[o + k], qR[k].d); }
          else qR[k] = fma(c, a.q[o + k], qR[k]);
        }
      }
    } else {
      // shared face: the neighbour's interpolated states in ITS face-node order (permuteinterface!, Utils/parallel.jl:198-201)
      const int64_t o = ((int64_t)r.aux * nfn + iR) * ND;
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        if constexpr (DUAL) qR[k] = Dual(a.q_recv[o + k], v_recv[o + k]);
        else qR[k] = a.q_recv[o + k];
      }
    }
    roe_flux<DIM, T>(qL, qR, nrm, a.ph.gamma, flux);
  }
  const double w = op.wface[i];
  double* dl = a.fluxe + ((int64_t)r.elL * NF + r.fL) * FL + i * ND;
#pragma unroll
  for (int k = 0; k < ND; ++k) dl[k] = -w * GenScal<T>::out(flux[k]);
  if (r.kind == FK_INTERIOR) {
    double* dr = a.fluxe + ((int64_t)r.elR * NF + r.fR) * FL + iR * ND;
#pragma unroll
    for (int k = 0; k < ND; ++k) dr[k] = w * GenScal<T>::out(flux[k]);
  }
}

// sparse faces: face node i of face f IS volume node perm[i,f] (sbp_sat_reduced_sc.jl:969-972).  T = Dual: tangent records.
template <int DIM, typename T>
__global__ void __launch_bounds__(128)
k_gen_face_sparse(const GenTab op, const __grid_constant__ FaceArgs a, int flux_id, const double* __restrict__ v,
                  const double* __restrict__ v_recv) {
  constexpr int ND = DIM + 2, NF = DIM + 1;
  constexpr bool DUAL = !std::is_same<T, double>::value;
  if (a.ctl->stop || a.ctl->kry_done) return;
  const int nn = op.nn, nfn = op.nfn, EL = nn * ND, FL = nfn * ND;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.ng * nfn) return;
  const int64_t g = a.g0 + t / nfn;
  const int i = (int)(t % nfn);
  const FaceRec r = a.faces[g];
  T qL[ND], qR[ND], flux[ND];
  double nrm[DIM];
  {
    const int64_t o = (int64_t)r.elL * EL + (int64_t)op.perm[r.fL * nfn + i] * ND;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      if constexpr (DUAL) qL[k] = Dual(a.q[o + k], v[o + k]);
      else qL[k] = a.q[o + k];
    }
    const double* np_ = a.nrm + g * a.nrm_face_stride + i * a.nrm_node_stride;
#pragma unroll
    for (int d = 0; d < DIM; ++d) nrm[d] = np_[d];
  }
  int iR = i;
  if (r.kind == FK_BOUNDARY) {
    const double* xp = a.coords_bndry + ((int64_t)r.elR * nfn + i) * DIM;
    double xb[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) xb[d] = xp[d];
    if constexpr (DUAL) bc_flux_dual<DIM>(r.aux, qL, xb, nrm, a.ph, flux);
    else bc_flux_any<DIM>(r.aux, qL, xb, nrm, a.ph, flux);
  } else {
    iR = op.nbrperm[r.orient * nfn + i];
    const bool interior = r.kind == FK_INTERIOR;
    const int64_t o = interior ? (int64_t)r.elR * EL + (int64_t)op.perm[r.fR * nfn + iR] * ND : ((int64_t)r.aux * nfn + iR) * ND;
    const double* bq = interior ? a.q : a.q_recv;
    const double* bv = interior ? v : v_recv;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      if constexpr (DUAL) qR[k] = Dual(bq[o + k], bv[o + k]);
      else qR[k] = bq[o + k];
    }
    numerical_flux<DIM, T>(flux_id, qL, qR, nrm, a.ph.gamma, flux);
  }
  const double w = op.wface[i];
  double* dl = a.fluxe + ((int64_t)r.elL * NF + r.fL) * FL + i * ND;
#pragma unroll
  for (int k = 0; k < ND; ++k) dl[k] = -w * GenScal<T>::out(flux[k]);
  if (r.kind == FK_INTERIOR) {
    double* dr = a.fluxe + ((int64_t)r.elR * NF + r.fR) * FL + iR * ND;
#pragma unroll
    for (int k = 0; k < ND; ++k) dr[k] = w * GenScal<T>::out(flux[k]);
  }
}

// dataPrep checks of one node (euler.jl:543-611): the first offending (element, node), density before pressure
template <int DIM>
__device__ __forceinline__ bool gen_check_node(const ElemArgs& a, const double* qn, int64_t e, int j) {
  const double press = calc_pressure<DIM>(qn, a.ph.gamma - 1.0);
  if ((a.ph.check_density && !(qn[0] > 0.0)) || (a.ph.check_pressure && !(press > 0.0))) {
    const int code = (a.ph.check_density && !(qn[0] > 0.0)) ? 1 : 2;
    const unsigned long long loc = ((unsigned long long)e << 8) | (unsigned)j;
    atomicMin(&a.ctl->err_loc, ((unsigned long long)(code - 1) << 62) | loc);
    atomicExch(&a.ctl->err_code, 1);
    atomicExch(&a.ctl->stop, 1);
    return false;
  }
  return true;
}

// the epilogue of one node: + source, then `res` (EPI_RES) or pde_post_func + the RK4 / LSERK54 stage update and the
// stage-1 norm partial (rk4.jl:244-319, 446-457; lserk.jl:183-205).  Schemes 0 and 1 (the sum-free RK4 form needs the
// staged epilogue of the tuned kernels).  Returns this node's sum_k M k^2.
template <int DIM, int MODE>
__device__ __forceinline__ double gen_epilogue(const ElemArgs& a, int64_t node, const double* acc) {
  constexpr int ND = DIM + 2;
  const int64_t dof0 = node * ND;
  double nrm2 = 0.0;
  if (MODE == EPI_RES) {
#pragma unroll
    for (int c = 0; c < ND; ++c) a.res[dof0 + c] = acc[c] + (a.srcw ? a.srcw[dof0 + c] : 0.0);
    return 0.0;
  }
  const double mv = a.minv[node], mw = a.stage == 1 ? a.mass[node] : 0.0;
#pragma unroll
  for (int c = 0; c < ND; ++c) {
    const int64_t dof = dof0 + c;
    const double k = acc[c] * mv + (a.srcm ? a.srcm[dof] : 0.0);
    const double xo = a.x_old[dof];
    if (a.stage == 1) nrm2 = fma(k * mw, k, nrm2);
    if (a.scheme == 1) {
      const double dq = a.stage == 1 ? a.hh * k : a.ah * a.ksum[dof] + a.hh * k;
      a.ksum[dof] = dq;
      a.q_next[dof] = xo + a.h6 * dq;
    } else if (a.stage == 1) {
      a.ksum[dof] = k;
      a.q_next[dof] = xo + a.ah * k;
    } else if (a.stage < 4) {
      a.ksum[dof] = a.ksum[dof] + 2.0 * k;
      a.q_next[dof] = xo + a.ah * k;
    } else {
      a.q_next[dof] = xo + a.h6 * (a.ksum[dof] + k);
    }
  }
  return nrm2;
}

// deterministic per-CTA sum of the norm partials (128 threads)
__device__ __forceinline__ void gen_block_norm(const ElemArgs& a, double nrm2) {
  __shared__ double s_red[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = nrm2;
  __syncthreads();
  if (threadIdx.x == 0) a.norm_partials[blockIdx.x] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
}

// weakdifferentiate! (trans = true) of the Euler flux: acc[:] = sum_d sum_j Q[j,i,d] F_d(q_j)     (euler.jl:628-658)
template <int DIM, typename T>
__device__ __forceinline__ void gen_volume(const GenTab& op, const ElemArgs& a, const double* __restrict__ v, int64_t e, int i,
                                           double* acc, bool check) {
  constexpr int ND = DIM + 2;
  constexpr bool DUAL = !std::is_same<T, double>::value;
  const int nn = op.nn, EL = nn * ND;
  for (int j = 0; j < nn; ++j) {
    T qn[ND], F[ND];
    double qv[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      qv[k] = a.q[e * EL + j * ND + k];
      if constexpr (DUAL) qn[k] = Dual(qv[k], v[e * EL + j * ND + k]);
      else qn[k] = qv[k];
    }
    if (check && i == 0 && !gen_check_node<DIM>(a, qv, e, j)) {
      // keep the arithmetic finite; the result is discarded
#pragma unroll
      for (int k = 0; k < ND; ++k) qn[k] = T(k == 0 || k == ND - 1 ? 1.0 : 0.0);
    }
    const double* dx = a.dxidx + e * a.dx_el_stride + j * a.dx_node_stride;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double dir[DIM];
#pragma unroll
      for (int p = 0; p < DIM; ++p) dir[p] = dx[d + DIM * p];
      euler_flux<DIM, T>(qn, dir, a.ph.gamma - 1.0, F);
      const double c = op.Qt[((int64_t)d * nn + j) * nn + i];
#pragma unroll
      for (int k = 0; k < ND; ++k) acc[k] = fma(c, GenScal<T>::out(F[k]), acc[k]);
    }
  }
}

// interiorfaceintegrate! / boundaryintegrate! in gather form: acc[:] += sum_f sum_i' RfN[f,i'][i] * record[f][i']
template <int DIM>
__device__ __forceinline__ void gen_face_gather(const GenTab& op, const ElemArgs& a, int64_t e, int i, double* acc) {
  constexpr int ND = DIM + 2, NF = DIM + 1;
  const int nn = op.nn, nfn = op.nfn, FL = nfn * ND;
  const double* G = a.fluxe + e * (NF * FL);
  for (int m = 0; m < NF * nfn; ++m) {
    const double c = op.RfN[(int64_t)m * nn + i];
    if (c == 0.0) continue;          // (Gamma-type operators: the stencil of a face is a subset of the volume nodes)
#pragma unroll
    for (int k = 0; k < ND; ++k) acc[k] = fma(c, G[m * ND + k], acc[k]);
  }
}

// one thread per (element, node i)
template <int DIM, int MODE>
__global__ void __launch_bounds__(128)
k_gen_element(const GenTab op, const __grid_constant__ ElemArgs a) {
  constexpr int ND = DIM + 2;
  if (a.ctl->stop || a.ctl->kry_done) return;
  const int nn = op.nn;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = t < (a.nE - a.e_begin) * nn;
  double nrm2 = 0.0;
  if (act) {
    const int64_t e = a.e_begin + t / nn;
    const int i = (int)(t % nn);
    double acc[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) acc[k] = 0.0;
    gen_volume<DIM, double>(op, a, nullptr, e, i, acc, true);
    gen_face_gather<DIM>(op, a, e, i, acc);
    nrm2 = gen_epilogue<DIM, MODE>(a, e * nn + i, acc);
  }
  if (MODE == EPI_RK && a.stage == 1) gen_block_norm(a, nrm2);
}

// out = dR/dq * v for dense faces (the tangent records come from k_gen_face<DIM, Dual>)
template <int DIM>
__global__ void __launch_bounds__(128)
k_gen_jvp_element(const GenTab op, const __grid_constant__ ElemArgs a, const double* __restrict__ v, double* __restrict__ out) {
  constexpr int ND = DIM + 2;
  if (a.ctl->kry_done) return;
  const int nn = op.nn;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nE * nn) return;
  const int64_t e = t / nn;
  const int i = (int)(t % nn);
  double acc[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) acc[k] = 0.0;
  gen_volume<DIM, Dual>(op, a, v, e, i, acc, false);
  gen_face_gather<DIM>(op, a, e, i, acc);
#pragma unroll
  for (int k = 0; k < ND; ++k) out[(e * nn + i) * ND + k] = acc[k];
}

// split form with the Ismail-Roe flux (euler_funcs.jl:240-288): acc[:] = -sum_m 2 S[i,m,d] F_d(q_hi, q_lo), the pair's
// flux taken with the metrics of its higher-numbered node, as the reference's (i, j < i) loop does
template <int DIM, typename T>
__device__ __forceinline__ void gen_split_volume(const GenTab& op, const ElemArgs& a, const double* __restrict__ v, int64_t e,
                                                 int i, double* acc, bool check) {
  constexpr int ND = DIM + 2;
  constexpr bool DUAL = !std::is_same<T, double>::value;
  const int nn = op.nn, EL = nn * ND;
  const double gami = a.ph.gamma - 1.0;
  double qv[ND];
  T qi[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) qv[k] = a.q[e * EL + i * ND + k];
  bool ok = true;
  if (check) ok = gen_check_node<DIM>(a, qv, e, i);
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    if (!ok) qv[k] = (k == 0 || k == ND - 1) ? 1.0 : 0.0;
    if constexpr (DUAL) qi[k] = Dual(qv[k], v[e * EL + i * ND + k]);
    else qi[k] = qv[k];
  }
  const IRNode<DIM, T> zi = ir_node<DIM>(qi, gami);
  for (int m = 0; m < nn; ++m) {
    if (m == i) continue;
    double qmv[ND];
    T qm[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      qmv[k] = a.q[e * EL + m * ND + k];
      if constexpr (DUAL) qm[k] = Dual(qmv[k], v[e * EL + m * ND + k]);
      else qm[k] = qmv[k];
    }
    const double pm = calc_pressure<DIM>(qmv, gami);
    if (!(qmv[0] > 0.0) || !(pm > 0.0)) continue;        // reported by the thread of node m; keeps the logarithms finite
    const IRNode<DIM, T> zm = ir_node<DIM>(qm, gami);
    const int hi = m > i ? m : i;
    const double* dx = a.dxidx + e * a.dx_el_stride + hi * a.dx_node_stride;
    double dirs[DIM][DIM];
    T F[DIM][ND];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int p = 0; p < DIM; ++p) dirs[d][p] = dx[d + DIM * p];
    if (m > i) ir_flux<DIM, DIM>(zm, zi, dirs, a.ph.gamma, F);
    else ir_flux<DIM, DIM>(zi, zm, dirs, a.ph.gamma, F);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const double sc = -op.S2[((int64_t)d * nn + i) * nn + m];
#pragma unroll
      for (int c = 0; c < ND; ++c) acc[c] = fma(sc, GenScal<T>::out(F[d][c]), acc[c]);
    }
  }
}

// sparse-face records: the face-node slots that coincide with volume node i
template <int DIM>
__device__ __forceinline__ void gen_sparse_gather(const GenTab& op, const ElemArgs& a, int64_t e, int i, double* acc) {
  constexpr int ND = DIM + 2, NF = DIM + 1;
  const double* G = a.fluxe + e * (NF * op.nfn * ND);
  for (int u = 0; u < NF; ++u) {
    const int slot = op.inv[i * NF + u];
    if (slot < 0) break;
#pragma unroll
    for (int c = 0; c < ND; ++c) acc[c] += G[slot * ND + c];
  }
}

template <int DIM, int MODE>
__global__ void __launch_bounds__(128)
k_gen_element_split(const GenTab op, const __grid_constant__ ElemArgs a) {
  constexpr int ND = DIM + 2;
  if (a.ctl->stop || a.ctl->kry_done) return;
  const int nn = op.nn;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = t < (a.nE - a.e_begin) * nn;
  double nrm2 = 0.0;
  if (act) {
    const int64_t e = a.e_begin + t / nn;
    const int i = (int)(t % nn);
    double acc[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) acc[k] = 0.0;
    gen_split_volume<DIM, double>(op, a, nullptr, e, i, acc, true);
    gen_sparse_gather<DIM>(op, a, e, i, acc);
    nrm2 = gen_epilogue<DIM, MODE>(a, e * nn + i, acc);
  }
  if (MODE == EPI_RK && a.stage == 1) gen_block_norm(a, nrm2);
}

// out = dR/dq * v for the split form (tangent records from k_gen_face_sparse<DIM, Dual>)
template <int DIM>
__global__ void __launch_bounds__(128)
k_gen_jvp_element_split(const GenTab op, const __grid_constant__ ElemArgs a, const double* __restrict__ v,
                        double* __restrict__ out) {
  constexpr int ND = DIM + 2;
  if (a.ctl->kry_done) return;
  const int nn = op.nn;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nE * nn) return;
  const int64_t e = t / nn;
  const int i = (int)(t % nn);
  double acc[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) acc[k] = 0.0;
  gen_split_volume<DIM, Dual>(op, a, v, e, i, acc, false);
  gen_sparse_gather<DIM>(op, a, e, i, acc);
#pragma unroll
  for (int k = 0; k < ND; ++k) out[(e * nn + i) * ND + k] = acc[k];
}

// getSendDataFace (Utils/parallel.jl:249-258) of any element vector (the state, or the direction of a J*v product):
// one thread per (shared face, face node, variable)
template <int DIM>
__global__ void k_gen_pack(const GenTab op, const double* __restrict__ x, const int32_t* __restrict__ sh_el,
                           const uint8_t* __restrict__ sh_face, int64_t nS, double* __restrict__ x_send,
                           double* const* __restrict__ face_dst, const Ctl* ctl) {
  constexpr int ND = DIM + 2;
  if (ctl->stop) return;
  const int nn = op.nn, nfn = op.nfn, ss = op.ss;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nS * nfn * ND) return;
  const int k = (int)(t % ND);
  const int i = (int)((t / ND) % nfn);
  const int64_t j = t / ((int64_t)ND * nfn);
  const double* b = x + (int64_t)sh_el[j] * (nn * ND) + k;
  const int f = sh_face[j];
  double s = 0.0;
  if (op.sparse) s = b[(int64_t)op.perm[f * nfn + i] * ND];
  else
    for (int n = 0; n < ss; ++n) s = fma(op.interp[n * nfn + i], b[(int64_t)op.perm[f * ss + n] * ND], s);
  if (face_dst) face_dst[j][i * ND + k] = s;
  else x_send[t] = s;
}

}  // namespace pdes
