// Device-side diagnostics of majorIterationCallback (solver/euler/euler.jl:330-407): the
// functionals the reference logs during a run, reduced over the resident q and the weak residual R(q) in ONE pass
// instead of five host sweeps:
//   [0] calcEntropyIntegral      sum_j M_j U(q_j),  U = -rho (log p - gamma log rho)/(gamma-1)   entropy_flux.jl:141-157,
//                                                                                                  euler_funcs.jl:1059-1071
//   [1] contractResEntropyVars   sum_j w(q_j) . R_j,  w = convertToEntropy(q)/(gamma-1)            entropy_flux.jl:164-186,
//                                                                                                  conversion.jl:50-89
//   [2] calcKineticEnergy        0.5 sum_j M_j rho |v|^2 / volume                                  entropy_flux.jl:414-440
//   [3] calcKineticEnergydt      sum_j M_j v.(dq_mom/dt - v drho/dt) / volume, dq/dt = Minv R      entropy_flux.jl:456-485
//   [4] volume = sum_j M_j       (mesh.volume)
//   [5..5+nd) integrateQ         sum_j M_j q_j[k]                                                  entropy_flux.jl:231-247
//   [5+nd]   calcEnstrophy       0.5 sum_j M_j rho |curl v|^2 / volume (3D)                        entropy_flux.jl:322-355
// Two-pass deterministic reduction (CTA partials, then k_reduce_rows).
#pragma once
#include <stdint.h>

namespace pdes {

constexpr int DIAG_T = 256;

template <int DIM>
__global__ void __launch_bounds__(DIAG_T)
k_diag_partials(const double* __restrict__ q, const double* __restrict__ res, const double* __restrict__ mass,
                int64_t n_nodes, double gamma, double* __restrict__ partials) {
  constexpr int ND = DIM + 2, NV = 5 + ND;
  const double gami = gamma - 1.0;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  for (int64_t j = (int64_t)blockIdx.x * DIAG_T + threadIdx.x; j < n_nodes; j += (int64_t)gridDim.x * DIAG_T) {
    double qn[ND], rn[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) { qn[k] = q[j * ND + k]; rn[k] = res[j * ND + k]; }
    const double M = mass[j], rho = qn[0];
    double m2 = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) m2 += qn[1 + d] * qn[1 + d];
    const double k1 = 0.5 * m2 / rho;
    const double rho_int = qn[DIM + 1] - k1;          // rho * internal energy
    const double p = gami * rho_int;
    acc[0] += (-rho * (log(p) - gamma * log(rho)) / gami) * M;
    // entropy variables (conversion.jl:50-89) scaled by 1/(gamma-1)
    const double s = log(gami * rho_int / pow(rho, gamma));
    const double fac = 1.0 / rho_int;
    double w[ND];
    w[0] = (rho_int * (gamma + 1.0 - s) - qn[DIM + 1]) * fac;
#pragma unroll
    for (int d = 0; d < DIM; ++d) w[1 + d] = qn[1 + d] * fac;
    w[DIM + 1] = -rho * fac;
    double wr = 0.0;
#pragma unroll
    for (int k = 0; k < ND; ++k) wr += (w[k] / gami) * rn[k];
    acc[1] += wr;
    // kinetic energy and its rate; dq/dt = Minv R = R / M
    double vv = 0.0, term = 0.0;
    const double drhodt = rn[0] / M;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const double v = qn[1 + d] / rho;
      vv += v * v;
      term += v * (rn[1 + d] / M - drhodt * v);
    }
    acc[2] += M * rho * vv;
    acc[3] += M * term;
    acc[4] += M;
#pragma unroll
    for (int k = 0; k < ND; ++k) acc[5 + k] += M * qn[k];
  }
  __shared__ double sh[NV][DIAG_T / 32];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double s = acc[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[v][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int wp = 0; wp < DIAG_T / 32; ++wp) s += sh[threadIdx.x][wp];
    partials[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
}

// calcEnstrophy (entropy_flux.jl:322-355) = 1/(2V) sum_j (w_j/jac_j) rho_j |omega_j|^2, omega = curl v by calcVorticity
// (euler_funcs.jl:1095-1155): velocities differentiated in the parametric directions with D_d = H^-1 Q_d
// (differentiateElement!), mapped to x-y-z with the UNSCALED metrics of the element's FIRST node (the reference indexes the
// [3,3,nn] array dxidx_unscaled with two subscripts, i.e. node 1 for every node -- identical for straight-sided elements),
// 3D only.  One thread per (element, node), any operator size; partial sums per CTA.
__global__ void __launch_bounds__(DIAG_T)
k_enstrophy_partials(const double* __restrict__ q, const double* __restrict__ Q, const double* __restrict__ w,
                     const double* __restrict__ dxidx, int64_t dx_el_stride, const double* __restrict__ mass, int nn,
                     int64_t nE, double* __restrict__ partials) {
  constexpr int ND = 5;
  double acc = 0.0;
  for (int64_t t = (int64_t)blockIdx.x * DIAG_T + threadIdx.x; t < nE * nn; t += (int64_t)gridDim.x * DIAG_T) {
    const int64_t e = t / nn;
    const int i = (int)(t - e * nn);
    const double* qe = q + e * nn * ND;
    double dv[3][3];      // [velocity component][parametric direction]
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
      for (int d = 0; d < 3; ++d) dv[v][d] = 0.0;
    for (int j = 0; j < nn; ++j) {
      const double rinv = 1.0 / qe[j * ND];
      const double vel[3] = {qe[j * ND + 1] * rinv, qe[j * ND + 2] * rinv, qe[j * ND + 3] * rinv};
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double c = Q[i + (int64_t)nn * (j + (int64_t)nn * d)];
#pragma unroll
        for (int v = 0; v < 3; ++v) dv[v][d] = fma(c, vel[v], dv[v][d]);
      }
    }
    const double wi = 1.0 / w[i];
    const double jac0 = w[0] / mass[e * nn];
    const double* dx = dxidx + e * dx_el_stride;            // first node of the element
    double xy[3][3];      // [velocity component][Cartesian direction]
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) s += (dv[v][d] * wi) * (dx[d + 3 * c] * jac0);
        xy[v][c] = s;
      }
    const double o0 = xy[2][1] - xy[1][2], o1 = -xy[2][0] + xy[0][2], o2 = xy[1][0] - xy[0][1];
    acc += mass[e * nn + i] * qe[i * ND] * (o0 * o0 + o1 * o1 + o2 * o2);
  }
  __shared__ double sh[DIAG_T / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int wp = 0; wp < DIAG_T / 32; ++wp) s += sh[wp];
    partials[blockIdx.x] = s;
  }
}

}  // namespace pdes
