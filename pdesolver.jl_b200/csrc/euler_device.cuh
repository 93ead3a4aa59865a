// Node-level Euler physics as device functions (fp64).  Each function states
// the reference routine whose arithmetic it reproduces (paths relative to
// /root/reference/src/solver/euler).  Templated on the spatial dimension and a
// scalar type T so the same code serves the real residual and, later, the
// dual-number (Jacobian-vector) residual.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pdes {

template <int DIM> struct Dims { static constexpr int ND = DIM + 2; };

// Reciprocal / reciprocal square root for normal-range arguments (densities, enthalpies, face areas): the
// MUFU seed (2^-20 relative) followed by two Newton steps (-> 2^-80, i.e. <= 1 ulp after rounding), without the
// denormal / special-value slow path of the CUDA math library that costs a branch and ~2x the instructions.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  double e = fma(-(h * y), y, 0.5);
  y = fma(y, e, y);
  e = fma(-(h * y), y, 0.5);
  return fma(y, e, y);
}

__device__ __forceinline__ double absv(double x) { return fabs(x); }   // Utils/complexify.jl:25-44 absvalue
__device__ __forceinline__ double maxv(double a, double b) { return fmax(a, b); }

// ---------------------------------------------------------------------------------------------------------
// Dual numbers {value, tangent}: the device-side counterpart of the reference's complex-step residual
// (Complex128 + Utils/complexify.jl): J*v = tangent of R(q + eps*v), exact to round-off.  absvalue flips the
// sign by the real part, max/isless compare real parts (complexify.jl:25-44, 207-217).
// ---------------------------------------------------------------------------------------------------------
struct Dual {
  double v, d;
  __host__ __device__ Dual() : v(0.0), d(0.0) {}
  __host__ __device__ Dual(double a) : v(a), d(0.0) {}
  __host__ __device__ Dual(double a, double b) : v(a), d(b) {}
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.v * b.d + a.d * b.v); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  const double r = 1.0 / b.v, q = a.v * r;
  return Dual(q, (a.d - q * b.d) * r);
}
__device__ __forceinline__ Dual& operator+=(Dual& a, Dual b) { a.v += b.v; a.d += b.d; return a; }
__device__ __forceinline__ Dual& operator-=(Dual& a, Dual b) { a.v -= b.v; a.d -= b.d; return a; }
__device__ __forceinline__ Dual fast_rcp(Dual x) { const double r = fast_rcp(x.v); return Dual(r, -r * r * x.d); }
__device__ __forceinline__ Dual fast_rsqrt(Dual x) {
  const double y = fast_rsqrt(x.v);
  return Dual(y, -0.5 * y * y * y * x.d);
}
__device__ __forceinline__ Dual absv(Dual x) { return x.v < 0.0 ? -x : x; }
__device__ __forceinline__ Dual maxv(Dual a, Dual b) { return a.v > b.v ? a : b; }
// Utils/complexify.jl:157-172 absvalue3: smooth |x| (delta = 1e-7), on the value and its tangent
__device__ __forceinline__ double absvalue3(double x) {
  const double delta = 1e-7, v1 = fabs(x);
  return v1 > delta ? v1 : ((x * x) / delta + delta) / 2;
}
__device__ __forceinline__ Dual absvalue3(Dual x) {
  const double delta = 1e-7, v1 = fabs(x.v);
  if (v1 > delta) return x.v < 0.0 ? -x : x;
  return Dual(((x.v * x.v) / delta + delta) / 2, x.v * x.d / delta);
}

// elementary functions on dual numbers (entropy-stable fluxes in J*v); val(): the real part (comparisons, branch selection)
using ::sqrt;
using ::log;
__device__ __forceinline__ double val(double x) { return x; }
__device__ __forceinline__ double val(const Dual& x) { return x.v; }
__device__ __forceinline__ Dual sqrt(Dual x) { const double s = ::sqrt(x.v); return Dual(s, 0.5 * x.d / s); }
__device__ __forceinline__ Dual log(Dual x) { return Dual(::log(x.v), x.d / x.v); }

// euler_funcs.jl:856-863 / 897-903 calcPressure
template <int DIM, typename T>
__device__ __forceinline__ T calc_pressure(const T* q, double gami) {
  T ke = q[1] * q[1];
#pragma unroll
  for (int d = 1; d < DIM; ++d) ke += q[1 + d] * q[1 + d];
  return gami * (q[DIM + 1] - 0.5 * ke / q[0]);
}

// euler_funcs.jl:512-536 / 749-774 calcEulerFlux
template <int DIM, typename T>
__device__ __forceinline__ void euler_flux(const T* q, const double* n, double gami, T* F) {
  T press = calc_pressure<DIM>(q, gami);
  T U = q[1] * n[0];
#pragma unroll
  for (int d = 1; d < DIM; ++d) U += q[1 + d] * n[d];
  U = U / q[0];
  F[0] = q[0] * U;
#pragma unroll
  for (int d = 0; d < DIM; ++d) F[1 + d] = q[1 + d] * U + n[d] * press;
  F[DIM + 1] = (q[DIM + 1] + press) * U;
}

// bc_solvers.jl:29-187 RoeSolver + :207-420 calcSAT:
//   flux = 0.5*(|A_hat| - A_hat)(q - qg) + F_euler(q, n), eigenvalue floors 0.025*rhoA
// Same formulas, with the ten divisions / square roots of the reference regrouped into three rsqrt and three
// reciprocals (1/rho = rsqrt(rho)^2, sqrt(rho) = rho*rsqrt(rho), dA*a = x*rsqrt(x) with x = dA^2 a^2,
// gami/a^2 = 1/(H - phi)); every regrouping is exact up to a few ulp (parity budget: 1e-12).
template <int DIM, typename T>
__device__ __forceinline__ void roe_flux(const T* q, const T* qg, const double* n, double gamma, T* flux) {
  constexpr int ND = DIM + 2;
  const double gami = gamma - 1.0;
  const double sat_Vn = 0.025, sat_Vl = 0.025;
  const T rL = fast_rsqrt(q[0]), rR = fast_rsqrt(qg[0]);
  const T sqL = q[0] * rL, sqR = qg[0] * rR;
  const T invL = rL * rL, invR = rR * rR;
  T vL[DIM], vR[DIM];
  T phiL = 0.0, phiR = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    vL[d] = q[1 + d] * invL; phiL += vL[d] * vL[d];
    vR[d] = qg[1 + d] * invR; phiR += vR[d] * vR[d];
  }
  phiL = 0.5 * phiL; phiR = 0.5 * phiR;
  const T HL = gamma * q[DIM + 1] * invL - gami * phiL;
  const T HR = gamma * qg[DIM + 1] * invR - gami * phiR;
  const T pressL = gami * (q[DIM + 1] - q[0] * phiL);     // p of the left state, reused by its Euler flux
  const T fac = fast_rcp(sqL + sqR);
  T v[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) v[d] = (sqL * vL[d] + sqR * vR[d]) * fac;
  const T H = (sqL * HL + sqR * HR) * fac;
  T dq[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) dq[i] = q[i] - qg[i];

  // ---- calcSAT ----
  double dA2 = 0.0;
  T Un = 0.0, phi = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) { dA2 += n[d] * n[d]; phi += v[d] * v[d]; Un += v[d] * n[d]; }
  phi = 0.5 * phi;
  const T Hm = H - phi;                 // a^2 = gami*(H - phi)
  const T x = dA2 * (gami * Hm);        // (dA a)^2
  const T rx = fast_rsqrt(x);
  const T dAa = x * rx;                 // dA * a
  T l1 = Un + dAa, l2 = Un - dAa, l3 = Un;
  const T rhoA = absv(Un) + dAa;
  l1 = 0.5 * (maxv(absv(l1), sat_Vn * rhoA) - l1);
  l2 = 0.5 * (maxv(absv(l2), sat_Vn * rhoA) - l2);
  l3 = 0.5 * (maxv(absv(l3), sat_Vl * rhoA) - l3);
  T e1 = phi * dq[0];
#pragma unroll
  for (int d = 0; d < DIM; ++d) e1 -= v[d] * dq[1 + d];
  e1 += dq[DIM + 1];
  T e2 = -Un * dq[0];
#pragma unroll
  for (int d = 0; d < DIM; ++d) e2 += n[d] * dq[1 + d];
  const T tmp1 = 0.5 * (l1 + l2) - l3;
  const T tmp2 = fast_rcp(Hm);          // gami / a^2
  const T tmp3 = fast_rcp(dA2);
  const T tmp4 = 0.5 * (l1 - l2) * rx;  // 0.5*(l1-l2)/(dA a)
  // sat = l3*dq + tmp1*(tmp2*E1dq + tmp3*E2dq) + tmp4*(E3dq + gami*E4dq)
  const T c1 = tmp1 * tmp2 * e1 + tmp4 * e2;          // multiplies [1, v, H]
  const T c2 = tmp1 * tmp3 * e2 + tmp4 * gami * e1;   // multiplies [0, n, Un]

  // ---- Euler flux of the left state ----
  T U = vL[0] * n[0];
#pragma unroll
  for (int d = 1; d < DIM; ++d) U += vL[d] * n[d];
  flux[0] = l3 * dq[0] + c1 + q[0] * U;
#pragma unroll
  for (int d = 0; d < DIM; ++d) flux[1 + d] = l3 * dq[1 + d] + c1 * v[d] + c2 * n[d] + (q[1 + d] * U + n[d] * pressL);
  flux[DIM + 1] = l3 * dq[DIM + 1] + c1 * H + c2 * Un + (q[DIM + 1] + pressL) * U;
}

// ---------------------------------------------------------------------------------------------------------
// entropy-stable two-point fluxes (config 2)
// ---------------------------------------------------------------------------------------------------------

// bc_solvers.jl:942-956 logavg(aL, aR) = (aL + aR) / (2 F) with log(aL/aR) supplied as lL - lR (logs tabulated per node:
// two per node instead of two per node pair); f = (xi-1)/(xi+1) is evaluated as (aL-aR)/(aL+aR).  Returns 2 F.
template <typename T>
__device__ __forceinline__ T logavg_2F(T aL, T aR, T lL, T lR, T inv_sum) {
  const T f = (aL - aR) * inv_sum;
  const T u = f * f;
  T F;
  if (val(u) < 1e-3) F = 1.0 + u * (1.0 / 3.0 + u * (1.0 / 5.0 + u * (1.0 / 7.0 + u * (1.0 / 9.0))));
  else F = 0.5 * (lL - lR) * fast_rcp(f);
  return 2.0 * F;
}
// per-node quantities of the Ismail-Roe flux: z1 = sqrt(rho/p), z_{1+d} = z1*u_d, z5 = sqrt(rho*p), log z1, log z5
template <int DIM, typename T = double>
struct IRNode {
  T z1, zv[DIM], z5, l1, l5;
};

template <int DIM, typename T>
__device__ __forceinline__ IRNode<DIM, T> ir_node(const T* q, double gami) {
  IRNode<DIM, T> z;
  const T p = calc_pressure<DIM>(q, gami);
  const T rinv = fast_rcp(q[0]);
  z.z1 = sqrt(q[0] * fast_rcp(p));
  z.z5 = sqrt(q[0] * p);
#pragma unroll
  for (int d = 0; d < DIM; ++d) z.zv[d] = z.z1 * q[1 + d] * rinv;
  z.l1 = log(z.z1);
  z.l5 = log(z.z5);
  return z;
}

// bc_solvers.jl:776-805 (2D) / 842-874 (3D) calcEulerFlux_IR for NDIR directions (dir[d][:]), F[d][:]
template <int DIM, int NDIR, typename T>
__device__ __forceinline__ void ir_flux(const IRNode<DIM, T>& L, const IRNode<DIM, T>& R, const double (*dir)[DIM],
                                        double gamma, T (*F)[DIM + 2]) {
  const double gamma_1 = gamma - 1.0;
  const T s1 = L.z1 + R.z1, s5 = L.z5 + R.z5;
  const T is1 = fast_rcp(s1), is5 = fast_rcp(s5);
  // with logavg(a) = s_a / (2 F_a) the quotients of the reference formulas need three reciprocals instead of six
  // (the two-point flux is FP64-pipe bound: the reciprocals were 21 % of k_element_split_n's samples):
  //   z5_ln / z1_ln = p1_hat * (2 F_1) / (2 F_5),   1 / rho_hat = 2 (2 F_5) / (s1 s5)
  const T tF5 = logavg_2F(L.z5, R.z5, L.l5, R.l5, is5);
  const T tF1 = logavg_2F(L.z1, R.z1, L.l1, R.l1, is1);
  const T itF5 = fast_rcp(tF5);
  const T la5 = s5 * itF5;                       // z5_ln
  const T rho_hat = 0.5 * s1 * la5;
  T vh[DIM], vv = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) { vh[d] = (L.zv[d] + R.zv[d]) * is1; vv += vh[d] * vh[d]; }
  const T p1_hat = s5 * is1;
  const T p2_hat = ((gamma + 1) / (2 * gamma)) * (p1_hat * tF1 * itF5) + (gamma_1 / (2 * gamma)) * p1_hat;
  const T h_hat = (gamma / gamma_1) * p2_hat * (2.0 * tF5 * is1 * is5) + 0.5 * vv;
#pragma unroll
  for (int i = 0; i < NDIR; ++i) {
    T un = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) un += dir[i][d] * vh[d];
    const T mv_n = rho_hat * un;
    F[i][0] = mv_n;
#pragma unroll
    for (int d = 0; d < DIM; ++d) F[i][1 + d] = mv_n * vh[d] + dir[i][d] * p1_hat;
    F[i][DIM + 1] = mv_n * h_hat;
  }
}

// bc_solvers.jl:898-909 calcEulerFlux_IRSLF = IR flux + applyEntropyKernel_diagE with the LFKernel
// (faceElementIntegrals.jl:455-468, 510-575): F += lambda_max(q_avg) * A0(q_avg) * (w(qL) - w(qR));
// convertToIR_ conversion.jl:160-204, getIRA0 IR_stab.jl:15-110, getLambdaMax euler_funcs.jl:1887-1913 with
// absvalue3 (Utils/complexify.jl:157-172)
template <int DIM>
__device__ __forceinline__ void convert_to_ir(const double* qc, double gamma, double* qe) {
  const double gamma_1 = gamma - 1.0, gamma_1i = 1.0 / gamma_1;
  double k1 = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) k1 += qc[1 + d] * qc[1 + d];
  k1 = 0.5 * k1 / qc[0];
  const double rho_int = qc[DIM + 1] - k1;
  const double s = log(gamma_1 * rho_int / pow(qc[0], gamma));
  const double fac = 1.0 / rho_int;
  qe[0] = ((rho_int * (gamma + 1 - s) - qc[DIM + 1]) * fac) * gamma_1i;
#pragma unroll
  for (int d = 0; d < DIM; ++d) qe[1 + d] = qc[1 + d] * fac * gamma_1i;
  qe[DIM + 1] = -qc[0] * fac * gamma_1i;
}

// convertToIR_ with the logarithms of the Ismail-Roe parameter vector: z1 = sqrt(rho/p), z5 = sqrt(rho p) give
// log rho = l1 + l5, log p = l5 - l1, and gamma_1 rho_int = p, so the physical entropy s = log(p / rho^gamma) needs neither
// the pow nor the log of conversion.jl:170-175 (they were ~40 % of k_face_flux_sparse's instructions)
template <int DIM, typename T>
__device__ __forceinline__ void convert_to_ir_z(const T* qc, const IRNode<DIM, T>& z, double gamma, T* qe) {
  const double gamma_1 = gamma - 1.0, gamma_1i = 1.0 / gamma_1;
  T k1 = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) k1 += qc[1 + d] * qc[1 + d];
  k1 = 0.5 * k1 * fast_rcp(qc[0]);
  const T rho_int = qc[DIM + 1] - k1;
  const T s = (z.l5 - z.l1) - gamma * (z.l1 + z.l5);
  const T fac = fast_rcp(rho_int);
  qe[0] = ((rho_int * (gamma + 1 - s) - qc[DIM + 1]) * fac) * gamma_1i;
#pragma unroll
  for (int d = 0; d < DIM; ++d) qe[1 + d] = qc[1 + d] * fac * gamma_1i;
  qe[DIM + 1] = -qc[0] * fac * gamma_1i;
}

template <int DIM, typename T>
__device__ __forceinline__ void irslf_flux(const T* qL, const T* qR, const double* n, double gamma, T* F) {
  constexpr int ND = DIM + 2;
  const double gami = gamma - 1.0;
  const IRNode<DIM, T> zL = ir_node<DIM>(qL, gami), zR = ir_node<DIM>(qR, gami);
  double dirs[1][DIM];
  T Fi[1][ND];
#pragma unroll
  for (int d = 0; d < DIM; ++d) dirs[0][d] = n[d];
  ir_flux<DIM, 1>(zL, zR, dirs, gamma, Fi);
  T qa[ND], vL[ND], vR[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) qa[i] = 0.5 * (qL[i] + qR[i]);
  convert_to_ir_z<DIM>(qL, zL, gamma, vL);
  convert_to_ir_z<DIM>(qR, zR, gamma, vR);
#pragma unroll
  for (int i = 0; i < ND; ++i) vL[i] -= vR[i];
  // A0 = dq/dw at q_avg (symmetric), applied to delta w
  const T p = calc_pressure<DIM>(qa, gami);
  const T rho = qa[0], rhoe = qa[DIM + 1], rhoinv = fast_rcp(rho);
  const T h = (rhoe + p) * rhoinv, a2 = gamma * p * rhoinv;
  T out[ND];
  out[0] = rho * vL[0] + rhoe * vL[DIM + 1];
  out[DIM + 1] = rhoe * vL[0] + (rho * h * h - a2 * p / gami) * vL[DIM + 1];
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    out[0] += qa[1 + c] * vL[1 + c];
    out[DIM + 1] += qa[1 + c] * h * vL[1 + c];
    T r = qa[1 + c] * vL[0] + h * qa[1 + c] * vL[DIM + 1];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      T a = qa[1 + d] * qa[1 + c] * rhoinv;
      if (d == c) a += p;
      r += a * vL[1 + d];
    }
    out[1 + c] = r;
  }
  // lambda_max = absvalue3(Un) + dA * a
  T Un = 0.0;
  double dA = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) { Un += n[d] * qa[1 + d] * rhoinv; dA += n[d] * n[d]; }
  dA = ::sqrt(dA);
  const T lambda_max = absvalue3(Un) + dA * sqrt(a2);
#pragma unroll
  for (int i = 0; i < ND; ++i) F[i] = Fi[0][i] + out[i] * lambda_max;
}


// conversion.jl:225-259 convertToConservativeFromIR_
template <int DIM>
__device__ __forceinline__ void convert_from_ir(const double* qe, double gamma, double* qc) {
  const double gamma_1 = gamma - 1.0;
  double k1 = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) k1 += qe[1 + d] * qe[1 + d];
  k1 = 0.5 * gamma_1 * k1 / qe[DIM + 1];
  const double s = gamma - gamma_1 * qe[0] + k1;
  // conversion.jl:240-243: rho_int = gamma_1 exp(-s/gamma_1) (gamma_1 / (-gamma_1 w_last)^gamma)^(1/gamma_1), evaluated
  // as ONE exponential of the collected exponent (one log + one exp instead of two pow + one exp)
  const double rho_int = gamma_1 * exp((log(gamma_1) - s - gamma * log(-gamma_1 * qe[DIM + 1])) / gamma_1);
  qc[0] = -qe[DIM + 1] * rho_int;
#pragma unroll
  for (int d = 0; d < DIM; ++d) qc[1 + d] = qe[1 + d] * rho_int;
  qc[DIM + 1] = (1.0 - k1) * rho_int / gamma_1;
}

// applyEntropyKernel(LFKernel) (faceElementIntegrals.jl:455-468): out = lambda_max(q_avg, n) * A0(q_avg) * delta_w
// (getIRA0 IR_stab.jl:15-110, getLambdaMax euler_funcs.jl:1887-1913 with absvalue3)
template <int DIM>
__device__ __forceinline__ void lf_entropy_kernel(const double* qa, const double* dw, const double* n, double gamma,
                                                  double* out) {
  const double gami = gamma - 1.0;
  const double p = calc_pressure<DIM>(qa, gami);
  const double rho = qa[0], rhoe = qa[DIM + 1], rhoinv = 1.0 / rho;
  const double h = (rhoe + p) * rhoinv, a2 = gamma * p * rhoinv;
  out[0] = rho * dw[0] + rhoe * dw[DIM + 1];
  out[DIM + 1] = rhoe * dw[0] + (rho * h * h - a2 * p / gami) * dw[DIM + 1];
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    out[0] += qa[1 + c] * dw[1 + c];
    out[DIM + 1] += qa[1 + c] * h * dw[1 + c];
    double r = qa[1 + c] * dw[0] + h * qa[1 + c] * dw[DIM + 1];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double a = qa[1 + d] * qa[1 + c] * rhoinv;
      if (d == c) a += p;
      r += a * dw[1 + d];
    }
    out[1 + c] = r;
  }
  double Un = 0.0, dA = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) { Un += n[d] * qa[1 + d] * rhoinv; dA += n[d] * n[d]; }
  dA = sqrt(dA);
  const double delta = 1e-7;
  const double v1 = fabs(Un);
  const double aUn = v1 > delta ? v1 : ((Un * Un) / delta + delta) / 2;
  const double lambda_max = aUn + dA * sqrt(a2);
#pragma unroll
  for (int i = 0; i < DIM + 2; ++i) out[i] *= lambda_max;
}

// applyEntropyKernel(LW2Kernel) (faceElementIntegrals.jl:393-440): |n| P^T Y |Lambda| S2 Y^T P delta_w with the
// eigensystem of the x-direction flux Jacobian at q_avg rotated into normal-tangential coordinates: getProjectionMatrix /
// getOrthogonalVector / getBinormalVector / projectToNT / projectToXY (Utils/projections.jl:25-375), calcEvecsx / calcEvalsx
// / calcEScalingx (eigensystem.jl:301-364, 479-615, 853-909).  Column c of Y is applied on the fly (no matrix stored).
template <int DIM>
__device__ __forceinline__ void lw2_entropy_kernel(const double* qa, const double* dw, const double* nrm_in, double gamma,
                                                   double* out) {
  constexpr int ND = DIM + 2;
  const double gami = gamma - 1.0;
  double len_fac = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) len_fac += nrm_in[d] * nrm_in[d];
  len_fac = sqrt(len_fac);
  double n[3] = {0.0, 0.0, 0.0}, Pm[DIM][DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) n[d] = nrm_in[d] / len_fac;
  const double add_fac = 1e-50;
  if (DIM == 2) {
    const double v1 = 1.0, v2 = -n[0] / (n[1] + add_fac), w1 = -n[1] / (n[0] + add_fac), w2 = 1.0;
    const double fac = rint(fabs(n[1]));
    double t1 = fac * v1 + (1 - fac) * w1, t2 = fac * v2 + (1 - fac) * w2;
    const double len = sqrt(t1 * t1 + t2 * t2);
    Pm[0][0] = n[0]; Pm[0][1] = n[1];
    Pm[1][0] = t1 / len; Pm[1][1] = t2 / len;
  } else {
    const double n1 = n[0], n2 = n[1], n3 = n[2];
    const double v3 = -(n1 + n2) / (n3 + add_fac), w2 = -(n1 + n3) / (n2 + add_fac), x1 = -(n2 + n3) / (n1 + add_fac);
    double fac = rint(fabs(n3));
    const double z1 = fac * x1 + (1 - fac) * 1.0, z2 = fac * 1.0 + (1 - fac) * w2, z3 = fac * 1.0 + (1 - fac) * 1.0;
    fac = rint(fabs(n1));
    double t1 = fac * 1.0 + (1 - fac) * z1, t2 = fac * 1.0 + (1 - fac) * z2, t3 = fac * v3 + (1 - fac) * z3;
    const double len = sqrt(t1 * t1 + t2 * t2 + t3 * t3);
    t1 /= len; t2 /= len; t3 /= len;
    Pm[0][0] = n1; Pm[0][1] = n2; Pm[0][DIM - 1] = n3;
    Pm[1][0] = t1; Pm[1][1] = t2; Pm[1][DIM - 1] = t3;
    Pm[DIM - 1][0] = n2 * t3 - n3 * t2; Pm[DIM - 1][1] = -(n1 * t3 - n3 * t1); Pm[DIM - 1][DIM - 1] = n1 * t2 - n2 * t1;
  }
  // projectToNT
  double q[ND], t1v[ND];
  q[0] = qa[0]; q[DIM + 1] = qa[DIM + 1];
  t1v[0] = dw[0]; t1v[DIM + 1] = dw[DIM + 1];
#pragma unroll
  for (int r = 0; r < DIM; ++r) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int c = 0; c < DIM; ++c) { a += Pm[r][c] * qa[1 + c]; b += Pm[r][c] * dw[1 + c]; }
    q[1 + r] = a; t1v[1 + r] = b;
  }
  // eigensystem in the (rotated) x direction
  const double q1 = q[0], t2 = 1.0 / q1;
  double ke = 0.0, vsq = 0.0, m2 = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) { ke += q[1 + d] * q[1 + d] * 0.5; vsq += q[1 + d] * q[1 + d] * t2 * t2; m2 += q[1 + d] * q[1 + d]; }
  const double a2 = gami * t2 * gamma * (q[DIM + 1] - t2 * ke), a = sqrt(a2), ia = 1.0 / a;
  const double r2 = sqrt(2.0) * 0.5, c1 = q1 * r2 * ia, u = q[1] * t2;
  const double H = (1.0 / gami) * (a2 + gami * vsq * 0.5), ua = q[1] * t2 * a;
  double Y[ND][ND];      // Y[r][c]
#pragma unroll
  for (int r = 0; r < ND; ++r)
#pragma unroll
    for (int c = 0; c < ND; ++c) Y[r][c] = 0.0;
  Y[0][0] = 1.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) Y[1 + d][0] = q[1 + d] * t2;
  Y[DIM + 1][0] = 0.5 * vsq;
  if (DIM == 2) {
    Y[2][1] = -q1; Y[3][1] = -q[2];
  } else {
    Y[3][1] = q1; Y[DIM + 1][1] = q[DIM];
    Y[2][DIM - 1] = -q1; Y[DIM + 1][DIM - 1] = -q[2];
  }
#pragma unroll
  for (int sg = 0; sg < 2; ++sg) {
    const int c = DIM + sg;
    const double s = sg == 0 ? 1.0 : -1.0;
    Y[0][c] = c1;
    Y[1][c] = c1 * (u + s * a);
#pragma unroll
    for (int d = 1; d < DIM; ++d) Y[1 + d][c] = q[1 + d] * r2 * ia;
    Y[DIM + 1][c] = c1 * (H + s * ua);
  }
  double lam[ND], S2[ND];
#pragma unroll
  for (int i = 0; i < DIM; ++i) lam[i] = u;
  lam[DIM] = u + a; lam[DIM + 1] = u - a;
  S2[0] = (gami * q1) / gamma;
  const double sc = -gami * (1.0 / (q1 * q1 * q1)) * (m2 - q1 * q[DIM + 1] * 2.0) * 0.5;
#pragma unroll
  for (int i = 1; i < ND; ++i) S2[i] = sc;
  double t2v[ND];
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < ND; ++r) s += Y[r][j] * t1v[r];
    t2v[j] = s * (len_fac * fabs(lam[j]) * S2[j]);
  }
#pragma unroll
  for (int r = 0; r < ND; ++r) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < ND; ++j) s += Y[r][j] * t2v[j];
    t1v[r] = s;
  }
  // projectToXY
  out[0] = t1v[0]; out[DIM + 1] = t1v[DIM + 1];
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < DIM; ++r) s += Pm[r][c] * t1v[1 + r];
    out[1 + c] = s;
  }
}

template <int DIM, typename T>
__device__ __forceinline__ void ir_flux_single(const T* qL, const T* qR, const double* n, double gamma, T* F) {
  constexpr int ND = DIM + 2;
  const IRNode<DIM, T> zL = ir_node<DIM>(qL, gamma - 1.0), zR = ir_node<DIM>(qR, gamma - 1.0);
  double dirs[1][DIM];
  T Fi[1][ND];
#pragma unroll
  for (int d = 0; d < DIM; ++d) dirs[0][d] = n[d];
  ir_flux<DIM, 1>(zL, zR, dirs, gamma, Fi);
#pragma unroll
  for (int i = 0; i < ND; ++i) F[i] = Fi[0][i];
}

// common_funcs.jl:25-78 calcIsentropicVortex (2D), :204-283 (3D)
template <int DIM>
__device__ inline void isentropic_vortex(const double* c, double gamma, double R, double* sol) {
  const double cv = R / (gamma - 1.0);
  const double r_in = 1, rho_in = 2, M_in = 0.95, p_in = 1 / gamma;
  double x = c[0], y = c[1];
  double theta, theta3 = 0.0;
  if (DIM == 3) {
    double z = c[2];
    const double phi_z = M_PI / 4;
    double theta1 = atan2(z, x);
    double phi2 = 0.5 * M_PI - theta1;
    double r_xz = sqrt(x * x + z * z);
    x = r_xz * sin(phi_z + phi2);
    theta3 = theta1 + phi_z + phi2 - 0.5 * M_PI;
    theta = atan2(x, y);
  } else {
    theta = atan2(y, x);
  }
  double r = sqrt(x * x + y * y);
  double tmp1 = ((gamma - 1) / 2) * M_in * M_in;
  double rho_r = rho_in * pow(1 + tmp1 * (1 - (r_in * r_in) / (r * r)), 1 / (gamma - 1));
  double p_r = p_in * pow(rho_r / rho_in, gamma);
  double a_r = sqrt(gamma * p_r / rho_r);
  double M_r = sqrt((2 / (gamma - 1)) * (pow(rho_in / rho_r, gamma - 1)) * (1 + tmp1) - 2 / (gamma - 1));
  double U_r = M_r * a_r;
  double e_r = cv * p_r / (rho_r * R);
  double E_r = rho_r * e_r + 0.5 * rho_r * U_r * U_r;
  sol[0] = rho_r;
  if (DIM == 2) {
    sol[1] = rho_r * (U_r * sin(theta));
    sol[2] = rho_r * (-U_r * cos(theta));
  } else {
    double v_r = U_r * sin(theta), u_r = -U_r * cos(theta);
    double w_r = u_r * sin(theta3);
    u_r = u_r * cos(theta3);
    sol[1] = rho_r * u_r;
    sol[2] = rho_r * v_r;
    sol[3] = rho_r * w_r;
  }
  sol[DIM + 1] = E_r;
}

// source.jl:85-96 MMSExp constants; common_funcs.jl:841-857 / 899-936 calcExp
struct MMSExp {
  static constexpr double a = 1.0 / 500, b = 0.01, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 20,
                          d1 = 1, d2 = 0.05, d3 = 0.15, d4 = 0.25, d5 = 1;
};

template <int DIM>
__device__ inline void calc_exp(const double* c, double gamma, double* q) {
  const double gamma_1 = gamma - 1.0;
  if (DIM == 2) {
    double x = c[0], y = c[1];
    const double af = 1.0 / 5, b = 0.01;
    q[0] = exp(af * x * y + b);
    q[1] = exp(af * 2 * x * y + b);
    q[2] = exp(af * 3 * x * y + b);
    q[3] = (1 / gamma_1 + 0.5) * exp(af * 5 * x * y + b) + 0.5 * exp(af * 3 * x * y + b);
  } else {
    typedef MMSExp M;
    double xyz = c[0] * c[1] * c[2];
    double t2 = exp(M::b);
    double t3 = M::a * M::c1 * xyz;
    q[0] = M::d1 * t2 * exp(t3);
    q[1] = M::d2 * t2 * exp(M::a * M::c2 * xyz);
    q[2] = M::d3 * t2 * exp(M::a * M::c3 * xyz);
    q[3] = M::d4 * t2 * exp(M::a * M::c4 * xyz);
    q[4] = (t2 * exp(-t3) * ((M::d2 * M::d2) * exp(M::a * M::c2 * xyz * 2.0) + (M::d3 * M::d3) * exp(M::a * M::c3 * xyz * 2.0)
            + (M::d4 * M::d4) * exp(M::a * M::c4 * xyz * 2.0)) * (1.0 / 2.0)) / M::d1 + (M::d5 * t2 * exp(M::a * M::c5 * xyz)) / gamma_1;
  }
}

// common_funcs.jl:312-351 calcFreeStream
template <int DIM>
__device__ inline void free_stream(double rho_free, double E_free, double Ma, double aoa, double* sol) {
  sol[0] = rho_free;
  sol[DIM + 1] = E_free;
  sol[1] = rho_free * Ma * cos(aoa);
  if (DIM == 2) {
    sol[2] = rho_free * Ma * sin(aoa);
  } else {
    sol[2] = 0.0;
    sol[3] = -rho_free * Ma * sin(aoa);
  }
}

// source.jl:65-81 (2D) / 98-177 (3D) SRCExp = div F(q_exact).  Written here as
// the analytic divergence of the flux of the calcExp fields (all of the form
// C*exp(k*phi), phi = x*y (2D) or x*y*z (3D), grad g = k*g*grad phi) instead of
// the reference's machine-generated expression; tests compare the two.
template <int DIM>
__device__ inline void src_exp(const double* c, double gamma, double* S) {
  const double gamma_1 = gamma - 1.0;
  double k[5], A[5];       // rho, m_1..m_DIM, p as A*exp(k*phi)
  double phi, gphi[3];
  if (DIM == 2) {
    const double af = 1.0 / 5, eb = exp(0.01);
    phi = c[0] * c[1];
    gphi[0] = c[1]; gphi[1] = c[0]; gphi[2] = 0.0;
    k[0] = af; k[1] = 2 * af; k[2] = 3 * af; k[3] = 0.0; k[4] = 5 * af;
    A[0] = eb; A[1] = eb; A[2] = eb; A[3] = 0.0; A[4] = eb;   // p = exp(5 af xy + b)
  } else {
    typedef MMSExp M;
    const double eb = exp(M::b);
    phi = c[0] * c[1] * c[2];
    gphi[0] = c[1] * c[2]; gphi[1] = c[0] * c[2]; gphi[2] = c[0] * c[1];
    k[0] = M::a * M::c1; k[1] = M::a * M::c2; k[2] = M::a * M::c3; k[3] = M::a * M::c4; k[4] = M::a * M::c5;
    A[0] = M::d1 * eb; A[1] = M::d2 * eb; A[2] = M::d3 * eb; A[3] = M::d4 * eb; A[4] = M::d5 * eb;
  }
  double rho = A[0] * exp(k[0] * phi);
  double m[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < DIM; ++d) m[d] = A[1 + d] * exp(k[1 + d] * phi);
  double p = A[4] * exp(k[4] * phi);
  // E = p/gamma_1 + 0.5*sum m_d^2/rho ; each term is an exponential of phi
  double ke[3], kke[3];
  double E = p / gamma_1;
#pragma unroll
  for (int d = 0; d < DIM; ++d) { ke[d] = 0.5 * m[d] * m[d] / rho; kke[d] = 2 * k[1 + d] - k[0]; E += ke[d]; }
  // dE/dphi
  double dE = k[4] * p / gamma_1;
#pragma unroll
  for (int d = 0; d < DIM; ++d) dE += kke[d] * ke[d];
  double S0 = 0.0, SE = 0.0, Sm[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    // flux in direction d: [m_d, m_i m_d/rho + delta p, (E+p) m_d/rho]; d/dx_d = (d/dphi) * gphi[d]
    double g = gphi[d];
    S0 += k[1 + d] * m[d] * g;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      double mm = m[i] * m[d] / rho;
      Sm[i] += (k[1 + i] + k[1 + d] - k[0]) * mm * g;
    }
    Sm[d] += k[4] * p * g;
    double ud = m[d] / rho;
    SE += ((dE + k[4] * p) * ud + (E + p) * (k[1 + d] - k[0]) * ud) * g;
  }
  S[0] = S0;
#pragma unroll
  for (int i = 0; i < DIM; ++i) S[1 + i] = Sm[i];
  S[DIM + 1] = SE;
}

}  // namespace pdes
