// Jacobian-vector products of the Euler residual: J*v = d/d(eps) R(q + eps*v).
//
// The reference obtains it by running evalResidual in Complex128 with q_vec += i*eps*v, eps = 1e-20, and taking
// imag(res)/eps (NonlinearSolvers/newton_setup.jl:632-662, jacobian/residual_evaluation.jl:64-88,
// interface2.jl:454-498).  Here the same residual is evaluated on dual numbers {value, tangent}; the tangent is
// J*v exactly (no eps).  Dense-face operators + Roe flux (configuration 5 = the configuration-1 mesh).
// These kernels serve the Krylov loop, not the RK4 hot loop: they are straightforward (one thread per face node /
// per element node) and share every node-level function with the tuned kernels through the scalar template.
#pragma once
#include "residual_kernels.cuh"

namespace pdes {

template <int DIM>
__device__ inline void bc_flux_dual(int bc, const Dual* q, const double* x, const double* n, const PhysPar& ph, Dual* flux) {
  constexpr int ND = DIM + 2;
  Dual qg[ND];
  if (bc == 4) {  // noPenetrationBC: Euler flux of the wall-projected state (bc.jl:717-765)
    double nn2 = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) nn2 += n[d] * n[d];
    const double fac = 1.0 / ::sqrt(nn2);
    double nh[DIM];
    Dual Unrm = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) { nh[d] = n[d] * fac; Unrm += Dual(nh[d]) * q[1 + d]; }
#pragma unroll
    for (int i = 0; i < ND; ++i) qg[i] = q[i];
#pragma unroll
    for (int d = 0; d < DIM; ++d) qg[1 + d] -= Dual(nh[d]) * Unrm;
    euler_flux<DIM, Dual>(qg, n, ph.gamma - 1.0, flux);
    return;
  }
  if (bc == 7) {  // ZeroFluxBC
#pragma unroll
    for (int i = 0; i < ND; ++i) flux[i] = Dual(0.0);
    return;
  }
  if (bc == 8) {  // noPenetrationESBC
    noslip_es_flux<DIM, Dual>(q, n, ph.gamma, flux);
    return;
  }
  double qd[ND];
  if (bc == 1) isentropic_vortex<DIM>(x, ph.gamma, ph.R, qd);
  else if (bc == 2) calc_exp<DIM>(x, ph.gamma, qd);
  else if (bc == 5) {
    qd[0] = 1.0; qd[DIM + 1] = 2.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) qd[1 + d] = 0.35355;
  } else if (bc == 6) {
#pragma unroll
    for (int i = 0; i < ND; ++i) qd[i] = 1.0;
  } else free_stream<DIM>(ph.rho_free, ph.E_free, ph.Ma, ph.aoa, qd);
#pragma unroll
  for (int i = 0; i < ND; ++i) qg[i] = Dual(qd[i]);      // the Dirichlet state does not depend on q
  roe_flux<DIM, Dual>(q, qg, n, ph.gamma, flux);
}

// one thread per (face, face node): tangent of -w f* / +w f* into the per-(element, face) records
template <int DIM, int NN, int NFN>
__global__ void __launch_bounds__(128)
k_jvp_face(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ FaceArgs a, const double* __restrict__ v) {
  constexpr int ND = DIM + 2, NF = DIM + 1, EL = NN * ND, FL = NFN * ND;
  if (a.ctl->kry_done) return;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.ng * NFN) return;
  const int64_t g = a.g0 + t / NFN;
  const int i = (int)(t % NFN);
  const FaceRec r = a.faces[g];
  Dual qL[ND], qR[ND], flux[ND];
  double nrm[DIM];
  {
    const double* np_ = a.nrm + g * a.nrm_face_stride + i * a.nrm_node_stride;
#pragma unroll
    for (int d = 0; d < DIM; ++d) nrm[d] = np_[d];
    const int64_t b = (int64_t)r.elL * EL;
    for (int j = 0; j < NN; ++j) {
      const double c = op.interp[j][i];
      const int64_t o = b + op.perm[r.fL][j] * ND;
#pragma unroll
      for (int k = 0; k < ND; ++k) { qL[k].v = fma(c, a.q[o + k], qL[k].v); qL[k].d = fma(c, v[o + k], qL[k].d); }
    }
  }
  int iR = i;
  if (r.kind == FK_BOUNDARY) {
    const double* xp = a.coords_bndry + ((int64_t)r.elR * NFN + i) * DIM;
    double xb[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) xb[d] = xp[d];
    bc_flux_dual<DIM>(r.aux, qL, xb, nrm, a.ph, flux);
  } else {
    iR = op.nbrperm[r.orient][i];
    if (r.kind == FK_SHARED) {
      // the neighbour's interpolated state and direction, in ITS face-node order (permuteinterface!, Utils/parallel.jl:198-201)
      const int64_t o = ((int64_t)r.aux * NFN + iR) * ND;
#pragma unroll
      for (int k = 0; k < ND; ++k) qR[k] = Dual(a.q_recv[o + k], a.v_recv[o + k]);
    } else {
      const int64_t b = (int64_t)r.elR * EL;
      for (int j = 0; j < NN; ++j) {
        const double c = op.interp[j][iR];
        const int64_t o = b + op.perm[r.fR][j] * ND;
#pragma unroll
        for (int k = 0; k < ND; ++k) { qR[k].v = fma(c, a.q[o + k], qR[k].v); qR[k].d = fma(c, v[o + k], qR[k].d); }
      }
    }
    roe_flux<DIM, Dual>(qL, qR, nrm, a.ph.gamma, flux);
  }
  const double w = op.wface[i];
  double* dl = a.fluxe + ((int64_t)r.elL * NF + r.fL) * FL + i * ND;
#pragma unroll
  for (int k = 0; k < ND; ++k) dl[k] = -w * flux[k].d;
  if (r.kind == FK_INTERIOR) {
    double* dr = a.fluxe + ((int64_t)r.elR * NF + r.fR) * FL + iR * ND;
#pragma unroll
    for (int k = 0; k < ND; ++k) dr[k] = w * flux[k].d;
  }
}

// one thread per (element, node i): out[:, i, e] = tangent of (Q^T F)[:, i] + face integration of the tangent records
template <int DIM, int NN, int NFN>
__global__ void __launch_bounds__(128)
k_jvp_element(const __grid_constant__ OpTab<DIM, NN, NFN> op, const __grid_constant__ ElemArgs a,
              const double* __restrict__ v, double* __restrict__ out) {
  constexpr int ND = DIM + 2, NF = DIM + 1, EL = NN * ND, FL = NFN * ND;
  if (a.ctl->kry_done) return;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nE * NN) return;
  const int64_t e = t / NN;
  const int i = (int)(t % NN);
  double acc[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) acc[k] = 0.0;
  for (int j = 0; j < NN; ++j) {
    Dual qn[ND], F[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) qn[k] = Dual(a.q[e * EL + j * ND + k], v[e * EL + j * ND + k]);
    const double* dx = a.dxidx + e * a.dx_el_stride + j * a.dx_node_stride;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double dir[DIM];
#pragma unroll
      for (int p = 0; p < DIM; ++p) dir[p] = dx[d + DIM * p];
      euler_flux<DIM, Dual>(qn, dir, a.ph.gamma - 1.0, F);
      const double c = op.Qt[d * NN + j][i];
#pragma unroll
      for (int k = 0; k < ND; ++k) acc[k] = fma(c, F[k].d, acc[k]);
    }
  }
  const double* G = a.fluxe + e * (NF * FL);
  for (int m = 0; m < NF * NFN; ++m) {
    const double c = op.RfN[m][i];
#pragma unroll
    for (int k = 0; k < ND; ++k) acc[k] = fma(c, G[m * ND + k], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < ND; ++k) out[e * EL + i * ND + k] = acc[k];
}

}  // namespace pdes
