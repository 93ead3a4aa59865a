/*
 * euler_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of PDESolver.jl's Euler `evalResidual` + `rk4` hot
 * path, in the reference's own pass structure (separate dataPrep passes with
 * materialised aux_vars / flux_parametric / q_face / flux_face / q_bndry /
 * bndryflux, then volume, boundary, face, shared-face and source integrals).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library; the product (pdesolver.jl_b200/csrc) never does.
 *
 * Parity status: the reference is Julia 0.6 + un-vendored packages
 * (SummationByParts.jl, PumiInterface.jl, ODLCommonTools.jl) and cannot run in
 * the build container, so there is no oracle/_ref.  The node-level functions
 * below are pinned against every known-answer vector the reference's tests
 * hold (tests/test_oracle_golden.py); the SBP operator *application* loops are
 * pinned through the reference's integral-level goldens and identities; the
 * SBP operator *values* are inputs ("parity unpinned").
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src unless noted).  Arrays use the Julia column-major layout
 * of the reference; indices are 0-based.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdio.h>

#define ORC_MAXD 5   /* max numDofPerNode */
#define ORC_MAXFN 16 /* max nodes per face */

typedef struct { uint32_t elementL, elementR; uint8_t faceL, faceR, orient, pad; } OrcInterface;
typedef struct { uint32_t element; uint8_t face, pad[3]; } OrcBoundary;

enum { ORC_FLUX_ROE = 1, ORC_FLUX_IR = 2, ORC_FLUX_IRSLF = 3, ORC_FLUX_STANDARD = 4 };
enum { ORC_BC_ISENTROPIC_VORTEX = 1, ORC_BC_EXP = 2, ORC_BC_FREESTREAM = 3,
       ORC_BC_NOPENETRATION = 4, ORC_BC_RHO1E2U3 = 5, ORC_BC_ALLONES = 6, ORC_BC_ZEROFLUX = 7,
       ORC_BC_NOPENETRATION_ES = 8 };
enum { ORC_SRC_NONE = 0, ORC_SRC_EXP = 1 };
/* faceElementIntegrals.jl:735-741 FaceElementDict */
enum { ORC_FEI_EC = 1, ORC_FEI_ELF_PENALTY = 2, ORC_FEI_ESLF = 3, ORC_FEI_ELW2_PENALTY = 4, ORC_FEI_ESLW2 = 5 };

typedef struct {
  int32_t dim, nd, nn, nfn, ss, nfaces, norient, sparse_face;
  int64_t nE, nF, nB;
  int32_t numBC, flux_id, volume_flux_id, volume_integral_type, src_id;
  int32_t check_density, check_pressure;
  int32_t face_element_id;   /* 0: face_integral_type 1; else face_integral_type 2 with this FaceElementIntegralType */
  double gamma, R, Ma, aoa, rho_free, E_free;
  const double *Q, *w, *interp, *wface;
  const int64_t *perm, *nbrperm;
  const double *dxidx, *jac, *coords, *nrm_face, *nrm_bndry, *coords_bndry;
  const OrcInterface *ifaces;
  const OrcBoundary *bfaces;
  const int64_t *bndry_offsets;
  const int32_t *bc_ids;
} OrcProblem;

typedef struct {               /* one SharedFaceData (Utils/parallel_types.jl:70-183) */
  int64_t nfaces;
  const OrcBoundary *bndries_local;
  const OrcInterface *interfaces;
  const double *nrm_sharedface; /* [dim,nfn,nfaces] */
  double *q_send, *q_recv;      /* [nd,nfn,nfaces] */
  /* parallel_data = element (face_integral_type 2): the peer's elements adjacent to this part, whole elements
   * [nd,nn,n_remote] in the order of the remote element numbers, and the number of the first one (shared_element_offsets) */
  const double *q_recv_el;
  int64_t el_offset;
} OrcPeer;

/* ------------------------------------------------------------------------ */
/* node-level physics                                                        */
/* ------------------------------------------------------------------------ */

/* euler_funcs.jl:856-863 (2D), 897-903 (3D): calcPressure, conservative vars */
double orc_calc_pressure(int dim, double gamma, const double *q) {
  double ke = 0.0;
  for (int d = 0; d < dim; ++d) ke += q[1 + d] * q[1 + d];
  return (gamma - 1.0) * (q[dim + 1] - 0.5 * ke / q[0]);
}

/* euler_funcs.jl:512-536 (2D), 749-774 (3D): calcEulerFlux in direction dir */
void orc_euler_flux(int dim, double gamma, const double *q, const double *dir, double *F) {
  double press = orc_calc_pressure(dim, gamma, q);
  double U = 0.0;
  for (int d = 0; d < dim; ++d) U += q[1 + d] * dir[d];
  U /= q[0];
  F[0] = q[0] * U;
  for (int d = 0; d < dim; ++d) F[1 + d] = q[1 + d] * U + dir[d] * press;
  F[dim + 1] = (q[dim + 1] + press) * U;
}

/* bc_solvers.jl:207-310 (2D), 313-420 (3D): calcSAT = 0.5*(|A_hat| - A)*dq   */
static void calc_sat(int dim, double gamma, const double *vel, double H, const double *dq,
                     const double *nrm, double *sat) {
  const double sat_Vn = 0.025, sat_Vl = 0.025, tau = 1.0;
  double gami = gamma - 1.0;
  int nd = dim + 2;
  double dA = 0.0, Un = 0.0, phi = 0.0;
  for (int d = 0; d < dim; ++d) { dA += nrm[d] * nrm[d]; phi += vel[d] * vel[d]; Un += vel[d] * nrm[d]; }
  dA = sqrt(dA);
  phi *= 0.5;
  double a = sqrt(gami * (H - phi));
  double lambda1 = Un + dA * a, lambda2 = Un - dA * a, lambda3 = Un;
  double rhoA = fabs(Un) + dA * a;
  lambda1 = 0.5 * (tau * fmax(fabs(lambda1), sat_Vn * rhoA) - lambda1);
  lambda2 = 0.5 * (tau * fmax(fabs(lambda2), sat_Vn * rhoA) - lambda2);
  lambda3 = 0.5 * (tau * fmax(fabs(lambda3), sat_Vl * rhoA) - lambda3);
  double E1dq[ORC_MAXD], E2dq[ORC_MAXD];
  for (int i = 0; i < nd; ++i) sat[i] = lambda3 * dq[i];
  /* E1*dq */
  double e1 = phi * dq[0];
  for (int d = 0; d < dim; ++d) e1 -= vel[d] * dq[1 + d];
  e1 += dq[dim + 1];
  E1dq[0] = e1;
  for (int d = 0; d < dim; ++d) E1dq[1 + d] = e1 * vel[d];
  E1dq[dim + 1] = e1 * H;
  /* E2*dq */
  double e2 = -Un * dq[0];
  for (int d = 0; d < dim; ++d) e2 += nrm[d] * dq[1 + d];
  E2dq[0] = 0.0;
  for (int d = 0; d < dim; ++d) E2dq[1 + d] = e2 * nrm[d];
  E2dq[dim + 1] = e2 * Un;
  double tmp1 = 0.5 * (lambda1 + lambda2) - lambda3;
  double tmp2 = gami / (a * a);
  double tmp3 = 1.0 / (dA * dA);
  for (int i = 0; i < nd; ++i) sat[i] = sat[i] + tmp1 * (tmp2 * E1dq[i] + tmp3 * E2dq[i]);
  /* E3*dq, E4*dq */
  E1dq[0] = e2;
  for (int d = 0; d < dim; ++d) E1dq[1 + d] = e2 * vel[d];
  E1dq[dim + 1] = e2 * H;
  E2dq[0] = 0.0;
  for (int d = 0; d < dim; ++d) E2dq[1 + d] = e1 * nrm[d];
  E2dq[dim + 1] = e1 * Un;
  tmp1 = 0.5 * (lambda1 - lambda2) / (dA * a);
  for (int i = 0; i < nd; ++i) sat[i] = sat[i] + tmp1 * (E1dq[i] + gami * E2dq[i]);
}

/* bc_solvers.jl:29-103 (2D), 111-187 (3D): RoeSolver(params,q,qg,aux,nrm,flux) */
void orc_roe_solver(int dim, double gamma, const double *q, const double *qg,
                    const double *nrm, double *flux) {
  int nd = dim + 2;
  double gami = gamma - 1.0;
  double fac = 1.0 / q[0], velL[3], velR[3], phi = 0.0;
  for (int d = 0; d < dim; ++d) { velL[d] = q[1 + d] * fac; phi += velL[d] * velL[d]; }
  phi *= 0.5;
  double HL = gamma * q[dim + 1] * fac - gami * phi;
  fac = 1.0 / qg[0];
  phi = 0.0;
  for (int d = 0; d < dim; ++d) { velR[d] = qg[1 + d] * fac; phi += velR[d] * velR[d]; }
  phi *= 0.5;
  double HR = gamma * qg[dim + 1] * fac - gami * phi;
  double sqL = sqrt(q[0]), sqR = sqrt(qg[0]);
  fac = 1.0 / (sqL + sqR);
  double vel[3];
  for (int d = 0; d < dim; ++d) vel[d] = (sqL * velL[d] + sqR * velR[d]) * fac;
  double H = (sqL * HL + sqR * HR) * fac;
  double dq[ORC_MAXD], sat[ORC_MAXD], ef[ORC_MAXD];
  for (int i = 0; i < nd; ++i) dq[i] = q[i] - qg[i];
  calc_sat(dim, gamma, vel, H, dq, nrm, sat);
  orc_euler_flux(dim, gamma, q, nrm, ef);
  for (int i = 0; i < nd; ++i) flux[i] = sat[i] + ef[i];
}

/* bc_solvers.jl:942-956: logavg */
double orc_logavg(double aL, double aR) {
  double xi = aL / aR;
  double f = (xi - 1) / (xi + 1);
  double u = f * f;
  double F;
  if (u < 1e-3) {
    /* @evalpoly(u, 1, 1/3, 1/5, 1/7, 1/9): Horner */
    F = 1.0 + u * (1.0 / 3.0 + u * (1.0 / 5.0 + u * (1.0 / 7.0 + u * (1.0 / 9.0))));
  } else {
    F = (log(xi) / 2.0) / f;
  }
  return (aL + aR) / (2 * F);
}

/* bc_solvers.jl:776-805 (2D), 842-874 (3D): calcEulerFlux_IR, ndir directions;
 * dir[dim x ndir] column-major, F[nd x ndir].  ndir=1 is the single-direction
 * method (:725-752, :808-838). */
void orc_ir_flux(int dim, double gamma, const double *qL, const double *qR,
                 const double *dir, int ndir, double *F) {
  double gamma_1 = gamma - 1.0;
  double pL = orc_calc_pressure(dim, gamma, qL), pR = orc_calc_pressure(dim, gamma, qR);
  double z1L = sqrt(qL[0] / pL), z1R = sqrt(qR[0] / pR);
  double zvL[3], zvR[3];
  for (int d = 0; d < dim; ++d) { zvL[d] = z1L * qL[1 + d] / qL[0]; zvR[d] = z1R * qR[1 + d] / qR[0]; }
  double z5L = sqrt(qL[0] * pL), z5R = sqrt(qR[0] * pR);
  double rho_hat = 0.5 * (z1L + z1R) * orc_logavg(z5L, z5R);
  double vh[3], vv = 0.0;
  for (int d = 0; d < dim; ++d) { vh[d] = (zvL[d] + zvR[d]) / (z1L + z1R); vv += vh[d] * vh[d]; }
  double p1_hat = (z5L + z5R) / (z1L + z1R);
  double p2_hat = ((gamma + 1) / (2 * gamma)) * orc_logavg(z5L, z5R) / orc_logavg(z1L, z1R)
                + (gamma_1 / (2 * gamma)) * (z5L + z5R) / (z1L + z1R);
  double h_hat = gamma * p2_hat / (rho_hat * gamma_1) + 0.5 * vv;
  int nd = dim + 2;
  for (int i = 0; i < ndir; ++i) {
    const double *n = dir + dim * i;
    double un = 0.0;
    for (int d = 0; d < dim; ++d) un += n[d] * vh[d];
    double mv_n = rho_hat * un;
    F[nd * i + 0] = mv_n;
    for (int d = 0; d < dim; ++d) F[nd * i + 1 + d] = mv_n * vh[d] + n[d] * p1_hat;
    F[nd * i + dim + 1] = mv_n * h_hat;
  }
}

/* conversion.jl:160-204: convertToIR_ (entropy variables scaled by 1/gamma_1) */
void orc_convert_to_ir(int dim, double gamma, const double *qc, double *qe) {
  double gamma_1 = gamma - 1.0, gamma_1i = 1 / gamma_1;
  double k1 = 0.0;
  for (int d = 0; d < dim; ++d) k1 += qc[1 + d] * qc[1 + d];
  k1 = 0.5 * k1 / qc[0];
  double rho_int = qc[dim + 1] - k1;
  double s = log(gamma_1 * rho_int / pow(qc[0], gamma));
  double fac = 1.0 / rho_int;
  double tmp1 = -qc[0] * fac * gamma_1i;
  double e = qc[dim + 1];
  qe[0] = ((rho_int * (gamma + 1 - s) - e) * fac) * gamma_1i;
  for (int d = 0; d < dim; ++d) qe[1 + d] = qc[1 + d] * fac * gamma_1i;
  qe[dim + 1] = tmp1;
}

/* IR_stab.jl:15-110: getIRA0 = dq/dw, A0[nd x nd] column-major */
void orc_ira0(int dim, double gamma, const double *q, double *A0) {
  int nd = dim + 2;
  double gamma_1 = gamma - 1.0;
  double p = orc_calc_pressure(dim, gamma, q);
  double rho = q[0], rhoe = q[dim + 1], rhoinv = 1 / rho;
  double h = (rhoe + p) * rhoinv, a2 = gamma * p * rhoinv;
#define A(i, j) A0[(i) + nd * (j)]
  A(0, 0) = rho;
  for (int d = 0; d < dim; ++d) { A(1 + d, 0) = q[1 + d]; A(0, 1 + d) = q[1 + d]; }
  A(dim + 1, 0) = rhoe; A(0, dim + 1) = rhoe;
  for (int c = 0; c < dim; ++c) {
    for (int r = 0; r < dim; ++r) {
      double v = q[1 + c] * q[1 + r] * rhoinv;
      if (r == c) v += p;
      A(1 + r, 1 + c) = v;
    }
    A(dim + 1, 1 + c) = q[1 + c] * h;
    A(1 + c, dim + 1) = h * q[1 + c];
  }
  A(dim + 1, dim + 1) = rho * h * h - a2 * p / gamma_1;
#undef A
}

/* Utils/complexify.jl:157-172: absvalue3 (Harten-type smooth abs, delta=1e-7) */
static double absvalue3(double val) {
  const double delta = 1e-7;
  double v1 = fabs(val);
  if (v1 > delta) return v1;
  return ((val * val) / delta + delta) / 2;
}

/* euler_funcs.jl:1887-1913: getLambdaMax(params, qL, dir) */
double orc_lambda_max(int dim, double gamma, const double *qL, const double *dir) {
  double Un = 0.0, dA = 0.0, rhoLinv = 1 / qL[0];
  double pL = orc_calc_pressure(dim, gamma, qL);
  double aL = sqrt(gamma * pL * rhoLinv);
  for (int i = 0; i < dim; ++i) { Un += dir[i] * qL[i + 1] * rhoLinv; dA += dir[i] * dir[i]; }
  dA = sqrt(dA);
  return absvalue3(Un) + dA * aL;
}

/* bc_solvers.jl:898-909 calcEulerFlux_IRSLF = IR flux +
 * faceElementIntegrals.jl:510-575 applyEntropyKernel_diagE with the
 * LFKernel (:455-468): F += lambda_max(q_avg) * A0(q_avg) * (w(qL) - w(qR)) */
void orc_irslf_flux(int dim, double gamma, const double *qL, const double *qR,
                    const double *dir, double *F) {
  int nd = dim + 2;
  orc_ir_flux(dim, gamma, qL, qR, dir, 1, F);
  double q_avg[ORC_MAXD], vL[ORC_MAXD], vR[ORC_MAXD], A0[ORC_MAXD * ORC_MAXD], Ft[ORC_MAXD];
  for (int i = 0; i < nd; ++i) q_avg[i] = 0.5 * (qL[i] + qR[i]);
  orc_convert_to_ir(dim, gamma, qL, vL);
  orc_convert_to_ir(dim, gamma, qR, vR);
  for (int i = 0; i < nd; ++i) vL[i] = vL[i] - vR[i];
  orc_ira0(dim, gamma, q_avg, A0);
  double lambda_max = orc_lambda_max(dim, gamma, q_avg, dir);
  for (int i = 0; i < nd; ++i) {          /* smallmatvec!(A0, delta_w, flux) */
    double s = 0.0;
    for (int j = 0; j < nd; ++j) s += A0[i + nd * j] * vL[j];
    Ft[i] = s * lambda_max;
  }
  for (int i = 0; i < nd; ++i) F[i] += Ft[i];
}

/* flux.jl:783-795 (RoeFlux), :964-976 (IRSLFFlux), IRFlux: functor dispatch */
static void face_flux_functor(const OrcProblem *P, int flux_id, const double *qL, const double *qR,
                              const double *nrm, double *F) {
  switch (flux_id) {
    case ORC_FLUX_ROE: orc_roe_solver(P->dim, P->gamma, qL, qR, nrm, F); break;
    case ORC_FLUX_IR: orc_ir_flux(P->dim, P->gamma, qL, qR, nrm, 1, F); break;
    case ORC_FLUX_IRSLF: orc_irslf_flux(P->dim, P->gamma, qL, qR, nrm, F); break;
    default: fprintf(stderr, "oracle: unsupported flux id %d\n", flux_id); abort();
  }
}

/* common_funcs.jl:25-78: calcIsentropicVortex (2D) */
static void isentropic_vortex_2d(double gamma, double R, double cv, const double *coords, double *sol) {
  double x = coords[0], y = coords[1];
  double r_in = 1, rho_in = 2, M_in = 0.95, p_in = 1 / gamma;
  double r = sqrt(x * x + y * y);
  double theta = atan2(y, x);
  double tmp1 = ((gamma - 1) / 2) * M_in * M_in;
  double rho_r = rho_in * pow(1 + tmp1 * (1 - (r_in * r_in) / (r * r)), 1 / (gamma - 1));
  double p_r = p_in * pow(rho_r / rho_in, gamma);
  double a_r = sqrt(gamma * p_r / rho_r);
  double M_r = sqrt((2 / (gamma - 1)) * (pow(rho_in / rho_r, gamma - 1)) * (1 + tmp1) - 2 / (gamma - 1));
  double U_r = M_r * a_r;
  double u_r = U_r * sin(theta), v_r = -U_r * cos(theta);
  double e_r = cv * p_r / (rho_r * R);
  double E_r = rho_r * e_r + 0.5 * rho_r * U_r * U_r;
  sol[0] = rho_r; sol[1] = rho_r * u_r; sol[2] = rho_r * v_r; sol[3] = E_r;
}

/* common_funcs.jl:204-283: calcIsentropicVortex (3D, axis rotated by pi/4) */
static void isentropic_vortex_3d(double gamma, double R, double cv, const double *coords, double *sol) {
  double x = coords[0], y = coords[1], z = coords[2];
  double phi_z = M_PI / 4;
  double theta1 = atan2(z, x);
  double phi2 = 0.5 * M_PI - theta1;
  double r_xz = sqrt(x * x + z * z);
  x = r_xz * sin(phi_z + phi2);
  double theta3 = theta1 + phi_z + phi2 - 0.5 * M_PI;
  double r_in = 1, rho_in = 2, M_in = 0.95, p_in = 1 / gamma;
  double r = sqrt(x * x + y * y);
  double theta = atan2(x, y);            /* phi_z > 0 branch */
  double tmp1 = ((gamma - 1) / 2) * M_in * M_in;
  double rho_r = rho_in * pow(1 + tmp1 * (1 - (r_in * r_in) / (r * r)), 1 / (gamma - 1));
  double p_r = p_in * pow(rho_r / rho_in, gamma);
  double a_r = sqrt(gamma * p_r / rho_r);
  double M_r = sqrt((2 / (gamma - 1)) * (pow(rho_in / rho_r, gamma - 1)) * (1 + tmp1) - 2 / (gamma - 1));
  double U_r = M_r * a_r;
  double v_r = U_r * sin(theta), u_r = -U_r * cos(theta);
  double e_r = cv * p_r / (rho_r * R);
  double E_r = rho_r * e_r + 0.5 * rho_r * U_r * U_r;
  double w_r = u_r * sin(theta3);
  u_r = u_r * cos(theta3);
  sol[0] = rho_r; sol[1] = rho_r * u_r; sol[2] = rho_r * v_r; sol[3] = rho_r * w_r; sol[4] = E_r;
}

void orc_isentropic_vortex(int dim, double gamma, double R, const double *coords, double *sol) {
  double cv = R / (gamma - 1);           /* types.jl:241-244 */
  if (dim == 2) isentropic_vortex_2d(gamma, R, cv, coords, sol);
  else isentropic_vortex_3d(gamma, R, cv, coords, sol);
}

/* source.jl:85-96: MMSExp constants */
static const double MMSExp_a = 1.0 / 500, MMSExp_b = 0.01, MMSExp_c1 = 1, MMSExp_c2 = 2, MMSExp_c3 = 3,
                    MMSExp_c4 = 4, MMSExp_c5 = 20, MMSExp_d1 = 1, MMSExp_d2 = 0.05, MMSExp_d3 = 0.15,
                    MMSExp_d4 = 0.25, MMSExp_d5 = 1;

/* common_funcs.jl:841-857 (2D), 899-936 (3D): calcExp */
void orc_calc_exp(int dim, double gamma, const double *coords, double *q) {
  double gamma_1 = gamma - 1.0;
  if (dim == 2) {
    double x = coords[0], y = coords[1];
    double af = 1.0 / 5, b = 0.01;
    q[0] = exp(af * x * y + b);
    q[1] = exp(af * 2 * x * y + b);
    q[2] = exp(af * 3 * x * y + b);
    q[3] = (1 / gamma_1 + 0.5) * exp(af * 5 * x * y + b) + 0.5 * exp(af * 3 * x * y + b);
  } else {
    double x = coords[0], y = coords[1], z = coords[2];
    double a = MMSExp_a, b = MMSExp_b, c1 = MMSExp_c1, c2 = MMSExp_c2, c3 = MMSExp_c3, c4 = MMSExp_c4,
           c5 = MMSExp_c5, d1 = MMSExp_d1, d2 = MMSExp_d2, d3 = MMSExp_d3, d4 = MMSExp_d4, d5 = MMSExp_d5;
    double t2 = exp(b);
    double t3 = a * c1 * x * y * z;
    q[0] = d1 * t2 * exp(t3);
    q[1] = d2 * t2 * exp(a * c2 * x * y * z);
    q[2] = d3 * t2 * exp(a * c3 * x * y * z);
    q[3] = d4 * t2 * exp(a * c4 * x * y * z);
    q[4] = (t2 * exp(-t3) * ((d2 * d2) * exp(a * c2 * x * y * z * 2.0) + (d3 * d3) * exp(a * c3 * x * y * z * 2.0)
            + (d4 * d4) * exp(a * c4 * x * y * z * 2.0)) * (1.0 / 2.0)) / d1 + (d5 * t2 * exp(a * c5 * x * y * z)) / gamma_1;
  }
}

/* common_funcs.jl:312-351: calcFreeStream (aoa in radians, types.jl:247) */
void orc_free_stream(int dim, double rho_free, double E_free, double Ma, double aoa, double *sol) {
  double rho = rho_free;
  sol[0] = rho;
  sol[dim + 1] = E_free;
  if (dim == 2) {
    sol[1] = rho * Ma * cos(aoa);
    sol[2] = rho * Ma * sin(aoa);
  } else {
    sol[1] = rho * Ma * cos(aoa);
    sol[2] = 0.0;
    sol[3] = -rho * Ma * sin(aoa);
  }
}

/* source.jl:65-81 (2D), 98-177 (3D): SRCExp functor (time independent) */
void orc_src_exp(int dim, double gamma, const double *coords, double *q) {
  double gamma_1 = gamma - 1.0;
  if (dim == 2) {
    double x = coords[0], y = coords[1];
    double af = 1.0 / 5, b = 0.01;
    q[0] = 2 * y * af * exp(2 * x * y * af + b) + 3 * x * af * exp(3 * x * y * af + b);
    q[1] = 3 * y * af * exp(3 * x * y * af + b) + 5 * y * af * exp(5 * x * y * af + b) + 4 * x * af * exp(4 * x * y * af + b);
    q[2] = 4 * y * af * exp(4 * x * y * af + b) + 10 * x * af * exp(5 * x * y * af + b);
    q[3] = 6 * y * af * (1 / gamma_1 + 1.5) * exp(6 * x * y * af + b) + 2 * y * af * exp(4 * x * y * af + b)
         + 7 * x * af * (1 / gamma_1 + 1.5) * exp(7 * x * y * af + b) + 2.5 * x * af * exp(5 * x * y * af + b);
    return;
  }
  double x = coords[0], y = coords[1], z = coords[2];
  double a = MMSExp_a, b = MMSExp_b, c1 = MMSExp_c1, c2 = MMSExp_c2, c3 = MMSExp_c3, c4 = MMSExp_c4,
         c5 = MMSExp_c5, d1 = MMSExp_d1, d2 = MMSExp_d2, d3 = MMSExp_d3, d4 = MMSExp_d4, d5 = MMSExp_d5;
  double t2 = exp(b), t3 = a * c2 * x * y * z, t4 = exp(t3), t5 = a * c4 * x * y * z, t6 = exp(t5);
  double t7 = c4 * d4 * t6 * x * y, t8 = a * c3 * x * y * z, t9 = exp(t8), t10 = c3 * d3 * t9 * x * z;
  double t11 = a * c5 * x * y * z, t12 = exp(t11), t13 = 1.0 / d1, t16 = a * c1 * x * y * z, t14 = exp(-t16);
  double t15 = c2 * d2 * t4 * y * z, t17 = b - t16, t18 = exp(t17), t19 = d2 * d2, t20 = a * c2 * x * y * z * 2.0;
  double t21 = exp(t20), t22 = d3 * d3, t23 = a * c3 * x * y * z * 2.0, t24 = exp(t23), t25 = d4 * d4;
  double t26 = a * c4 * x * y * z * 2.0, t27 = exp(t26), t28 = b + t11, t29 = exp(t28);
  double t30 = c2 * t19 * t21, t31 = c3 * t22 * t24, t32 = c4 * t25 * t27, t33 = t30 + t31 + t32;
  double t34 = t19 * t21, t35 = t22 * t24, t36 = t25 * t27, t37 = t34 + t35 + t36, t38 = 1.0 / gamma_1;
  double t39 = c1 - c4, t40 = exp(-a * t39 * x * y * z), t41 = c1 - c3, t42 = exp(-a * t41 * x * y * z);
  double t43 = d5 * t29, t44 = d5 * t29 * t38, t45 = t13 * t18 * t37 * (1.0 / 2.0), t46 = t43 + t44 + t45;
  double t47 = c1 - c2, t48 = exp(-a * t47 * x * y * z);
  q[0] = a * t2 * (t7 + t10 + t15);
  q[1] = a * c5 * d5 * t2 * t12 * y * z + a * d2 * t2 * t4 * t13 * t14 * (t7 + t10 - c1 * d4 * t6 * x * y + c2 * d4 * t6 * x * y
         - c1 * d3 * t9 * x * z + c2 * d3 * t9 * x * z - c1 * d2 * t4 * y * z + c2 * d2 * t4 * y * z * 2.0);
  q[2] = a * c5 * d5 * t2 * t12 * x * z + a * d3 * t2 * t9 * t13 * t14 * ((t7 + t15 - c1 * d4 * t6 * x * y + c3 * d4 * t6 * x * y
         - c1 * d3 * t9 * x * z) + (c3 * d3 * t9 * x * z * 2.0 - c1 * d2 * t4 * y * z + c3 * d2 * t4 * y * z));
  q[3] = a * c5 * d5 * t2 * t12 * x * y + a * d4 * t2 * t6 * t13 * t14 * ((t10 + t15 - c1 * d4 * t6 * x * y + c4 * d4 * t6 * x * y * 2.0
         - c1 * d3 * t9 * x * z) + (c4 * d3 * t9 * x * z - c1 * d2 * t4 * y * z + c4 * d2 * t4 * y * z));
  q[4] = d4 * t13 * t40 * (a * t13 * t18 * t33 * x * y + a * c5 * d5 * t29 * x * y + a * c5 * d5 * t29 * t38 * x * y
           - a * c1 * t13 * t18 * t37 * x * y * (1.0 / 2.0))
       + d3 * t13 * t42 * (a * t13 * t18 * t33 * x * z + a * c5 * d5 * t29 * x * z + a * c5 * d5 * t29 * t38 * x * z
           - a * c1 * t13 * t18 * t37 * x * z * (1.0 / 2.0))
       + d2 * t13 * t48 * (a * t13 * t18 * t33 * y * z + a * c5 * d5 * t29 * y * z + a * c5 * d5 * t29 * t38 * y * z
           - a * c1 * t13 * t18 * t37 * y * z * (1.0 / 2.0))
       - a * d4 * t13 * t39 * t40 * t46 * x * y - a * d3 * t13 * t41 * t42 * t46 * x * z - a * d2 * t13 * t46 * t47 * t48 * y * z;
}

/* bc.jl:554-567 isentropicVortexBC, :1756-1768 ExpBC, :1573-1587 FreeStreamBC
 * (Dirichlet state then RoeSolver); :717-765 + :1082-1150 noPenetrationBC
 * (Euler flux of the wall-projected state, Roe call commented out there). */
void orc_bc_flux(const OrcProblem *P, int bc_id, const double *q, const double *coords,
                 const double *nrm, double *flux) {
  int dim = P->dim, nd = dim + 2;
  double qg[ORC_MAXD];
  switch (bc_id) {
    case ORC_BC_ISENTROPIC_VORTEX:
      orc_isentropic_vortex(dim, P->gamma, P->R, coords, qg);
      orc_roe_solver(dim, P->gamma, q, qg, nrm, flux);
      break;
    case ORC_BC_EXP:
      orc_calc_exp(dim, P->gamma, coords, qg);
      orc_roe_solver(dim, P->gamma, q, qg, nrm, flux);
      break;
    case ORC_BC_FREESTREAM:
      orc_free_stream(dim, P->rho_free, P->E_free, P->Ma, P->aoa, qg);
      orc_roe_solver(dim, P->gamma, q, qg, nrm, flux);
      break;
    case ORC_BC_NOPENETRATION: {
      double n[3], nn2 = 0.0, Unrm = 0.0;
      for (int d = 0; d < dim; ++d) nn2 += nrm[d] * nrm[d];
      double fac = 1.0 / sqrt(nn2);
      for (int d = 0; d < dim; ++d) { n[d] = nrm[d] * fac; Unrm += n[d] * q[1 + d]; }
      for (int i = 0; i < nd; ++i) qg[i] = q[i];
      for (int d = 0; d < dim; ++d) qg[1 + d] -= n[d] * Unrm;
      orc_euler_flux(dim, P->gamma, qg, nrm, flux);
      break;
    }
    case ORC_BC_RHO1E2U3:      /* bc.jl:1454-1537, calcRho1Energy2U3 common_funcs.jl:754-779 */
      qg[0] = 1.0;
      for (int d = 0; d < dim; ++d) qg[1 + d] = 0.35355;
      qg[dim + 1] = 2.0;
      orc_roe_solver(dim, P->gamma, q, qg, nrm, flux);
      break;
    case ORC_BC_ALLONES:       /* bc.jl:1702-1722, calcOnes common_funcs.jl:647-654 */
      for (int i = 0; i < nd; ++i) qg[i] = 1.0;
      orc_roe_solver(dim, P->gamma, q, qg, nrm, flux);
      break;
    case ORC_BC_ZEROFLUX:      /* bc.jl:2140-2152 */
      for (int i = 0; i < nd; ++i) flux[i] = 0.0;
      break;
    case ORC_BC_NOPENETRATION_ES: {
      /* bc.jl:767-793: reflected state (getDirichletState :860-918) + calcLFFlux (bc_solvers.jl:428-446) with
       * getLambdaMaxSimple (IR_stab.jl:310-325) */
      double n[3], nn2 = 0.0, Unrm = 0.0, fluxL[ORC_MAXD], fluxR[ORC_MAXD], q_avg[ORC_MAXD];
      for (int d = 0; d < dim; ++d) nn2 += nrm[d] * nrm[d];
      double fac = 1.0 / sqrt(nn2);
      for (int d = 0; d < dim; ++d) { n[d] = nrm[d] * fac; Unrm += n[d] * q[1 + d]; }
      for (int i = 0; i < nd; ++i) qg[i] = q[i];
      for (int d = 0; d < dim; ++d) qg[1 + d] = -2 * Unrm * n[d] + q[1 + d];
      orc_euler_flux(dim, P->gamma, q, nrm, fluxL);
      orc_euler_flux(dim, P->gamma, qg, nrm, fluxR);
      for (int i = 0; i < nd; ++i) q_avg[i] = 0.5 * (q[i] + qg[i]);
      double lambda_max = orc_lambda_max(dim, P->gamma, q_avg, nrm);
      for (int i = 0; i < nd; ++i) flux[i] = 0.5 * (fluxL[i] + fluxR[i] - lambda_max * (qg[i] - q[i]));
      break;
    }
    default: fprintf(stderr, "oracle: unsupported BC id %d\n", bc_id); abort();
  }
}

/* ------------------------------------------------------------------------ */
/* SummationByParts.jl operator application (external package; semantics     */
/* confirmed inside the reference: jacobian/jacobian.jl:1015-1125,           */
/* solver/euler/faceElementIntegrals.jl:81-103,241-285,                      */
/* solver/euler/sbp_sat_reduced_sc.jl:969-972, test/euler/test_curvilinear.jl:30-48) */
/* ------------------------------------------------------------------------ */
#define IDX3(k, j, e, n1, n2) ((k) + (int64_t)(n1) * ((j) + (int64_t)(n2) * (e)))

/* face interpolation of one element: uface[:,i] = sum_j interp[j,col(i)] u[:,perm[j,face]]
 * col(i) = i (left / boundary) or nbrperm[i,orient] (right).  Sparse faces:
 * uface[:,i] = u[:, perm[col(i), face]]. */
static void face_interp_el(const OrcProblem *P, const double *u_el, int face, int orient_or_neg,
                           double *uface) {
  int nd = P->nd, nfn = P->nfn, ss = P->ss;
  for (int i = 0; i < nfn; ++i) {
    int col = orient_or_neg < 0 ? i : (int)P->nbrperm[i + nfn * orient_or_neg];
    double *uf = uface + nd * i;
    if (P->sparse_face) {
      int64_t node = P->perm[col + (int64_t)nfn * face];
      for (int k = 0; k < nd; ++k) uf[k] = u_el[k + nd * node];
    } else {
      for (int k = 0; k < nd; ++k) uf[k] = 0.0;
      for (int j = 0; j < ss; ++j) {
        double c = P->interp[j + ss * col];
        int64_t node = P->perm[j + (int64_t)ss * face];
        for (int k = 0; k < nd; ++k) uf[k] += c * u_el[k + nd * node];
      }
    }
  }
}

/* face integration into one element: res[:,perm[j,face]] += sgn*interp[j,col(i)]*wface[i]*flux[:,i] */
static void face_integrate_el(const OrcProblem *P, const double *flux, int face, int orient_or_neg,
                              double sgn, double *res_el) {
  int nd = P->nd, nfn = P->nfn, ss = P->ss;
  for (int i = 0; i < nfn; ++i) {
    int col = orient_or_neg < 0 ? i : (int)P->nbrperm[i + nfn * orient_or_neg];
    const double *f = flux + nd * i;
    if (P->sparse_face) {
      int64_t node = P->perm[col + (int64_t)nfn * face];
      for (int k = 0; k < nd; ++k) {
        double v = sgn * P->wface[i] * f[k];
#pragma omp atomic
        res_el[k + nd * node] += v;
      }
    } else {
      for (int j = 0; j < ss; ++j) {
        double c = P->interp[j + ss * col] * P->wface[i];
        int64_t node = P->perm[j + (int64_t)ss * face];
        for (int k = 0; k < nd; ++k) {
          double v = sgn * c * f[k];
#pragma omp atomic
          res_el[k + nd * node] += v;
        }
      }
    }
  }
}

/* weakdifferentiate!(sbp, d, flux, res, trans=true): res[:,i,e] += Q[j,i,d]*flux[:,j,e]
 * (call site euler.jl:637-640) */
static void weakdifferentiate_trans(const OrcProblem *P, int d, const double *flux, double *res) {
  int nd = P->nd, nn = P->nn;
  const double *Qd = P->Q + (int64_t)nn * nn * d;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < P->nE; ++e)
    for (int i = 0; i < nn; ++i)
      for (int j = 0; j < nn; ++j) {
        double c = Qd[j + nn * i];
        for (int k = 0; k < nd; ++k)
          res[IDX3(k, i, e, nd, nn)] += c * flux[IDX3(k, j, e, nd, nn)];
      }
}

/* ------------------------------------------------------------------------ */
/* evalResidual passes                                                       */
/* ------------------------------------------------------------------------ */
typedef struct {
  double *aux_vars, *flux_parametric, *q_face, *flux_face, *q_bndry, *bndryflux;
} OrcWork;

/* the intermediates are as large as the state: keep one set per process (the reference allocates them once in the
 * EulerData constructor, types.jl:584-646) instead of paying malloc + first-touch on every evaluation */
static OrcWork g_work;
static size_t g_work_key[4];
static OrcWork work_alloc_raw(const OrcProblem *P);
static void work_free_raw(OrcWork *W);
static OrcWork work_alloc(const OrcProblem *P) {
  size_t key[4] = {(size_t)P->nd * P->nn * P->nE * P->dim, (size_t)P->nF, (size_t)P->nB, (size_t)P->nfn};
  if (memcmp(key, g_work_key, sizeof(key)) != 0) {
    if (g_work_key[0]) work_free_raw(&g_work);
    g_work = work_alloc_raw(P);
    memcpy(g_work_key, key, sizeof(key));
  }
  return g_work;
}
static void work_free(OrcWork *W) { (void)W; }
static OrcWork work_alloc_raw(const OrcProblem *P) {
  OrcWork W;
  size_t nd = P->nd, nn = P->nn, nfn = P->nfn;
  W.aux_vars = (double *)malloc(sizeof(double) * nn * P->nE);
  W.flux_parametric = (double *)malloc(sizeof(double) * nd * nn * P->nE * P->dim);
  W.q_face = (double *)malloc(sizeof(double) * nd * 2 * nfn * (P->nF ? P->nF : 1));
  W.flux_face = (double *)malloc(sizeof(double) * nd * nfn * (P->nF ? P->nF : 1));
  W.q_bndry = (double *)malloc(sizeof(double) * nd * nfn * (P->nB ? P->nB : 1));
  W.bndryflux = (double *)malloc(sizeof(double) * nd * nfn * (P->nB ? P->nB : 1));
  return W;
}
static void work_free_raw(OrcWork *W) {
  free(W->aux_vars); free(W->flux_parametric); free(W->q_face); free(W->flux_face);
  free(W->q_bndry); free(W->bndryflux);
}

/* euler_funcs.jl:819-834 getAuxVars + euler.jl:543-611 checkDensity/checkPressure.
 * returns 0, or 1 (negative density) / 2 (negative pressure) with the first
 * offending element/node in err_loc (reference throws). */
static int aux_and_checks(const OrcProblem *P, const double *q, double *aux, int64_t *err_loc) {
  int nd = P->nd, nn = P->nn;
  int status = 0;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < P->nE; ++e)
    for (int j = 0; j < nn; ++j)
      aux[j + nn * e] = orc_calc_pressure(P->dim, P->gamma, q + IDX3(0, j, e, nd, nn));
  if (P->check_density)
    for (int64_t e = 0; e < P->nE && !status; ++e)
      for (int j = 0; j < nn; ++j)
        if (q[IDX3(0, j, e, nd, nn)] <= 0.0) { status = 1; err_loc[0] = e; err_loc[1] = j; break; }
  if (P->check_pressure && !status)
    for (int64_t e = 0; e < P->nE && !status; ++e)
      for (int j = 0; j < nn; ++j)
        if (aux[j + nn * e] <= 0.0) { status = 2; err_loc[0] = e; err_loc[1] = j; break; }
  return status;
}

/* euler_funcs.jl:23-58 getEulerFlux -> flux_parametric[nd,nn,nE,dim] */
static void get_euler_flux(const OrcProblem *P, const double *q, double *fp) {
  int nd = P->nd, nn = P->nn, dim = P->dim;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < P->nE; ++e)
    for (int j = 0; j < nn; ++j)
      for (int k = 0; k < dim; ++k) {
        double nrm[3];
        for (int p = 0; p < dim; ++p)
          nrm[p] = P->dxidx[k + dim * (p + dim * (j + (int64_t)nn * e))];
        orc_euler_flux(dim, P->gamma, q + IDX3(0, j, e, nd, nn), nrm,
                       fp + IDX3(0, j, e, nd, nn) + (int64_t)nd * nn * P->nE * k);
      }
}

/* flux.jl:613-641 interpolateFace (SBP interiorfaceinterpolate!) -> q_face[nd,2,nfn,nF] */
static void interpolate_face(const OrcProblem *P, const double *q, double *q_face) {
  int nd = P->nd, nn = P->nn, nfn = P->nfn;
#pragma omp parallel for schedule(static)
  for (int64_t f = 0; f < P->nF; ++f) {
    OrcInterface I = P->ifaces[f];
    double uL[ORC_MAXD * ORC_MAXFN], uR[ORC_MAXD * ORC_MAXFN];
    face_interp_el(P, q + (int64_t)nd * nn * I.elementL, I.faceL, -1, uL);
    face_interp_el(P, q + (int64_t)nd * nn * I.elementR, I.faceR, I.orient, uR);
    for (int i = 0; i < nfn; ++i)
      for (int k = 0; k < nd; ++k) {
        q_face[k + nd * (0 + 2 * (i + (int64_t)nfn * f))] = uL[k + nd * i];
        q_face[k + nd * (1 + 2 * (i + (int64_t)nfn * f))] = uR[k + nd * i];
      }
  }
}

/* flux.jl:37-64 calcFaceFlux */
static void calc_face_flux(const OrcProblem *P, const double *q_face, double *flux_face) {
  int nd = P->nd, nfn = P->nfn, dim = P->dim;
#pragma omp parallel for schedule(static)
  for (int64_t f = 0; f < P->nF; ++f)
    for (int j = 0; j < nfn; ++j)
      face_flux_functor(P, P->flux_id, q_face + nd * (0 + 2 * (j + (int64_t)nfn * f)),
                        q_face + nd * (1 + 2 * (j + (int64_t)nfn * f)),
                        P->nrm_face + dim * (j + (int64_t)nfn * f),
                        flux_face + nd * (j + (int64_t)nfn * f));
}

/* euler.jl:777-778 interiorfaceintegrate!(sbpface, interfaces, flux_face, res, Subtract) */
static void interior_face_integrate(const OrcProblem *P, const double *flux_face, double *res) {
  int nd = P->nd, nn = P->nn, nfn = P->nfn;
#pragma omp parallel for schedule(static)
  for (int64_t f = 0; f < P->nF; ++f) {
    OrcInterface I = P->ifaces[f];
    const double *fl = flux_face + (int64_t)nd * nfn * f;
    face_integrate_el(P, fl, I.faceL, -1, -1.0, res + (int64_t)nd * nn * I.elementL);
    face_integrate_el(P, fl, I.faceR, I.orient, +1.0, res + (int64_t)nd * nn * I.elementR);
  }
}

/* flux.jl:79-125 calcFaceIntegral_nopre */
static void calc_face_integral_nopre(const OrcProblem *P, const double *q, double *res) {
  int nd = P->nd, nn = P->nn, nfn = P->nfn, dim = P->dim;
#pragma omp parallel for schedule(static)
  for (int64_t f = 0; f < P->nF; ++f) {
    OrcInterface I = P->ifaces[f];
    double uL[ORC_MAXD * ORC_MAXFN], uR[ORC_MAXD * ORC_MAXFN], fl[ORC_MAXD * ORC_MAXFN];
    face_interp_el(P, q + (int64_t)nd * nn * I.elementL, I.faceL, -1, uL);
    face_interp_el(P, q + (int64_t)nd * nn * I.elementR, I.faceR, I.orient, uR);
    for (int j = 0; j < nfn; ++j)
      face_flux_functor(P, P->flux_id, uL + nd * j, uR + nd * j,
                        P->nrm_face + dim * (j + (int64_t)nfn * f), fl + nd * j);
    face_integrate_el(P, fl, I.faceL, -1, -1.0, res + (int64_t)nd * nn * I.elementL);
    face_integrate_el(P, fl, I.faceR, I.orient, +1.0, res + (int64_t)nd * nn * I.elementR);
  }
}

/* bc.jl:162-175 interpolateBoundary (SBP boundaryinterpolate!) */
static void interpolate_boundary(const OrcProblem *P, const double *q, double *q_bndry) {
  int nd = P->nd, nn = P->nn, nfn = P->nfn;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < P->nB; ++b)
    face_interp_el(P, q + (int64_t)nd * nn * P->bfaces[b].element, P->bfaces[b].face, -1,
                   q_bndry + (int64_t)nd * nfn * b);
}

/* bc.jl:49-80 getBCFluxes + :251-284 calcBoundaryFlux (DG) */
static void get_bc_fluxes(const OrcProblem *P, const double *q_bndry, double *bndryflux) {
  int nd = P->nd, nfn = P->nfn, dim = P->dim;
  for (int i = 0; i < P->numBC; ++i)
#pragma omp parallel for schedule(static)
    for (int64_t b = P->bndry_offsets[i]; b < P->bndry_offsets[i + 1]; ++b)
      for (int j = 0; j < nfn; ++j)
        orc_bc_flux(P, P->bc_ids[i], q_bndry + nd * (j + (int64_t)nfn * b),
                    P->coords_bndry + dim * (j + (int64_t)nfn * b),
                    P->nrm_bndry + dim * (j + (int64_t)nfn * b),
                    bndryflux + nd * (j + (int64_t)nfn * b));
}

/* euler.jl:675 boundaryintegrate!(sbpface, bndryfaces, bndryflux, res, Subtract) */
static void boundary_integrate(const OrcProblem *P, const double *bndryflux, double *res) {
  int nd = P->nd, nn = P->nn, nfn = P->nfn;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < P->nB; ++b)
    face_integrate_el(P, bndryflux + (int64_t)nd * nfn * b, P->bfaces[b].face, -1, -1.0,
                      res + (int64_t)nd * nn * P->bfaces[b].element);
}

/* euler_funcs.jl:158-196 calcVolumeIntegrals_nopre */
static void calc_volume_integrals_nopre(const OrcProblem *P, const double *q, double *res) {
  int nd = P->nd, nn = P->nn, dim = P->dim;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < P->nE; ++e) {
    double flux_el[ORC_MAXD * 32 * 3];
    for (int j = 0; j < nn; ++j)
      for (int k = 0; k < dim; ++k) {
        double nrm[3];
        for (int p = 0; p < dim; ++p)
          nrm[p] = P->dxidx[k + dim * (p + dim * (j + (int64_t)nn * e))];
        orc_euler_flux(dim, P->gamma, q + IDX3(0, j, e, nd, nn), nrm, flux_el + nd * (j + nn * k));
      }
    for (int k = 0; k < dim; ++k) {
      const double *Qd = P->Q + (int64_t)nn * nn * k;
      for (int i = 0; i < nn; ++i)
        for (int j = 0; j < nn; ++j)
          for (int c = 0; c < nd; ++c)
            res[IDX3(c, i, e, nd, nn)] += Qd[j + nn * i] * flux_el[c + nd * (j + nn * k)];
    }
  }
}

/* euler_funcs.jl:240-288 calcVolumeIntegralsSplitFormLinear, S = 0.5(Q - Q^T)
 * (flux_types.jl:969-971) with the IR multi-direction flux */
static void calc_volume_integrals_split_linear(const OrcProblem *P, const double *q, double *res) {
  int nd = P->nd, nn = P->nn, dim = P->dim;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < P->nE; ++e)
    for (int j = 0; j < nn; ++j)
      for (int k = 0; k < j; ++k) {
        double nrmD[9], F_d[ORC_MAXD * 3];
        for (int d = 0; d < dim; ++d)
          for (int p = 0; p < dim; ++p)
            nrmD[p + dim * d] = P->dxidx[d + dim * (p + dim * (j + (int64_t)nn * e))];
        if (P->volume_flux_id != ORC_FLUX_IR) { fprintf(stderr, "oracle: volume flux must be IR\n"); abort(); }
        orc_ir_flux(dim, P->gamma, q + IDX3(0, j, e, nd, nn), q + IDX3(0, k, e, nd, nn), nrmD, dim, F_d);
        for (int d = 0; d < dim; ++d) {
          const double *Qd = P->Q + (int64_t)nn * nn * d;
          double S = 0.5 * (Qd[j + nn * k] - Qd[k + nn * j]);
          for (int p = 0; p < nd; ++p) {
            res[IDX3(p, j, e, nd, nn)] -= 2 * S * F_d[p + nd * d];
            res[IDX3(p, k, e, nd, nn)] += 2 * S * F_d[p + nd * d];
          }
        }
      }
}


/* conversion.jl:225-259 convertToConservativeFromIR_ */
void orc_convert_from_ir(int dim, double gamma, const double *qe, double *qc) {
  double gamma_1 = gamma - 1.0, k1 = 0.0;
  for (int d = 0; d < dim; ++d) k1 += qe[1 + d] * qe[1 + d];
  k1 = 0.5 * gamma_1 * k1 / qe[dim + 1];
  double s = gamma - gamma_1 * qe[0] + k1;
  double rho_int = exp(-s / gamma_1) * pow(gamma_1 / pow(-gamma_1 * qe[dim + 1], gamma), 1 / gamma_1);
  rho_int *= gamma_1;
  qc[0] = -qe[dim + 1] * rho_int;
  for (int d = 0; d < dim; ++d) qc[1 + d] = qe[1 + d] * rho_int;
  qc[dim + 1] = (1.0 - k1) * rho_int / gamma_1;
}

/* faceElementIntegrals.jl:58-117 calcECFaceIntegral (DenseFace): two-point fluxes between every stencil node of
 * elementL and every stencil node of elementR in the Cartesian directions (nrmD = I), weighted by
 * E_ij^d = sum_k interp[i,k] interp[j,nbrperm[k]] wface[k] nrm[d,k].  resL/resR: [nd, nn] element blocks. */
static void calc_ec_face_integral(const OrcProblem *P, const OrcInterface *f, const double *qL, const double *qR,
                                  const double *nrm_xy, double *resL, double *resR) {
  int nd = P->nd, dim = P->dim, ss = P->ss, nfn = P->nfn;
  double nrmD[9] = {0}, fluxD[ORC_MAXD * 3];
  for (int d = 0; d < dim; ++d) nrmD[d + dim * d] = 1.0;
  if (P->flux_id != ORC_FLUX_IR) { fprintf(stderr, "oracle: the face-element integrals use IRFlux\n"); abort(); }
  for (int i = 0; i < ss; ++i) {
    int p_i = (int)P->perm[i + ss * f->faceL];
    for (int j = 0; j < ss; ++j) {
      int p_j = (int)P->perm[j + ss * f->faceR];
      orc_ir_flux(dim, P->gamma, qL + nd * p_i, qR + nd * p_j, nrmD, dim, fluxD);
      for (int d = 0; d < dim; ++d) {
        double Eij = 0.0;
        for (int k = 0; k < nfn; ++k) {
          int kR = (int)P->nbrperm[k + nfn * f->orient];
          Eij += P->interp[i + ss * k] * P->interp[j + ss * kR] * P->wface[k] * nrm_xy[d + dim * k];
        }
        for (int p = 0; p < nd; ++p) {
          resL[p + nd * p_i] -= Eij * fluxD[p + nd * d];
          resR[p + nd * p_j] += Eij * fluxD[p + nd * d];
        }
      }
    }
  }
}

/* faceElementIntegrals.jl:209-290 calcEntropyPenaltyIntegral (DenseFace) with the LFKernel (:455-468):
 * the entropy variables are interpolated to the face, the penalty lambda_max A0(q_avg) (wL - wR) wface is
 * interpolated back */
/* ---- eigensystem of the x-direction flux Jacobian and the Lax-Wendroff entropy kernel ---------------------------------
 * calcEvalsx (eigensystem.jl:301-364), calcEvecsx (:479-615, columns of Y, column-major Y[r + nd*c]), calcEScalingx
 * (:853-909: A0 = Y diag(S2) Y^T, Merriam's scaling); pinned by test_3d.jl:107-148 / test_lowlevel.jl:440-470
 * (Y Lambda Y^-1 = dF/dq, Y S2 Y^T = A0, 1e-12). */
void orc_evals_x(int dim, double gamma, const double *q, double *Lambda) {
  double gami = gamma - 1.0, t2 = 1.0 / q[0], u = q[1] * t2, ke = 0.0;
  for (int d = 0; d < dim; ++d) ke += q[1 + d] * q[1 + d] * 0.5;
  double a = sqrt(gami * t2 * gamma * (q[dim + 1] - t2 * ke));
  for (int i = 0; i < dim; ++i) Lambda[i] = u;
  Lambda[dim] = u + a;
  Lambda[dim + 1] = u - a;
}

void orc_evecs_x(int dim, double gamma, const double *q, double *Y) {
  int nd = dim + 2;
  double gami = gamma - 1.0, q1 = q[0], t2 = 1.0 / q1, ke = 0.0, vsq = 0.0;
  for (int d = 0; d < dim; ++d) { ke += q[1 + d] * q[1 + d] * 0.5; vsq += q[1 + d] * q[1 + d] * t2 * t2; }
  double a2 = gami * t2 * gamma * (q[dim + 1] - t2 * ke), a = sqrt(a2), ia = 1.0 / a;
  double r2 = sqrt(2.0) * 0.5;
  double c1 = q1 * r2 * ia;                       /* t13 / t15: rho / (sqrt(2) a) */
  double u = q[1] * t2;
  double H = (1.0 / gami) * (a2 + gami * vsq * 0.5);   /* t25 / t29 */
  double ua = q[1] * t2 * a;                      /* t26 / t30 */
  for (int i = 0; i < nd * nd; ++i) Y[i] = 0.0;
  /* column 1: entropy wave */
  Y[0] = 1.0;
  for (int d = 0; d < dim; ++d) Y[1 + d] = q[1 + d] * t2;
  Y[dim + 1] = 0.5 * vsq;
  if (dim == 2) {
    /* column 2: shear wave */
    Y[2 + nd * 1] = -q1;
    Y[3 + nd * 1] = -q[2];
  } else {
    Y[3 + nd * 1] = q1;                           /* R[4,2] = q1, R[5,2] = q4 */
    Y[4 + nd * 1] = q[3];
    Y[2 + nd * 2] = -q1;                          /* R[3,3] = -q1, R[5,3] = -q3 */
    Y[4 + nd * 2] = -q[2];
  }
  /* acoustic waves */
  for (int sgn = 0; sgn < 2; ++sgn) {
    int c = dim + sgn;
    double s = sgn == 0 ? 1.0 : -1.0;
    Y[0 + nd * c] = c1;
    Y[1 + nd * c] = c1 * (u + s * a);
    for (int d = 1; d < dim; ++d) Y[1 + d + nd * c] = q[1 + d] * r2 * ia;
    Y[dim + 1 + nd * c] = c1 * (H + s * ua);
  }
}

void orc_escaling_x(int dim, double gamma, const double *q, double *S) {
  double gami = gamma - 1.0, q1 = q[0], t2 = 1.0 / (q1 * q1 * q1), m2 = 0.0;
  for (int d = 0; d < dim; ++d) m2 += q[1 + d] * q[1 + d];
  double t = -gami * t2 * (m2 - q1 * q[dim + 1] * 2.0) * 0.5;
  S[0] = (gami * q1) / gamma;
  for (int i = 1; i < dim + 2; ++i) S[i] = t;
}

/* getOrthogonalVector / getBinormalVector / getProjectionMatrix (Utils/projections.jl:25-206): rows 2..dim+1 of P */
static void projection_rows(int dim, const double *n, double (*Pm)[3]) {
  const double add_fac = 1e-50;
  if (dim == 2) {
    double v1 = 1.0, v2 = -n[0] / (n[1] + add_fac), w1 = -n[1] / (n[0] + add_fac), w2 = 1.0;
    double fac = rint(fabs(n[1]));   /* Julia round(): ties to even */
    double t1 = fac * v1 + (1 - fac) * w1, t2 = fac * v2 + (1 - fac) * w2;
    double len = sqrt(t1 * t1 + t2 * t2);
    Pm[0][0] = n[0]; Pm[0][1] = n[1]; Pm[0][2] = 0.0;
    Pm[1][0] = t1 / len; Pm[1][1] = t2 / len; Pm[1][2] = 0.0;
    return;
  }
  double n1 = n[0], n2 = n[1], n3 = n[2];
  double v1 = 1.0, v2 = 1.0, v3 = -(n1 + n2) / (n3 + add_fac);
  double w1 = 1.0, w2 = -(n1 + n3) / (n2 + add_fac), w3 = 1.0;
  double x1 = -(n2 + n3) / (n1 + add_fac), x2 = 1.0, x3 = 1.0;
  double fac = rint(fabs(n3));
  double z1 = fac * x1 + (1 - fac) * w1, z2 = fac * x2 + (1 - fac) * w2, z3 = fac * x3 + (1 - fac) * w3;
  fac = rint(fabs(n1));
  double t1 = fac * v1 + (1 - fac) * z1, t2 = fac * v2 + (1 - fac) * z2, t3 = fac * v3 + (1 - fac) * z3;
  double len = sqrt(t1 * t1 + t2 * t2 + t3 * t3);
  t1 /= len; t2 /= len; t3 /= len;
  Pm[0][0] = n1; Pm[0][1] = n2; Pm[0][2] = n3;
  Pm[1][0] = t1; Pm[1][1] = t2; Pm[1][2] = t3;
  Pm[2][0] = n2 * t3 - n3 * t2; Pm[2][1] = -(n1 * t3 - n3 * t1); Pm[2][2] = n1 * t2 - n2 * t1;
}

/* getProjectionMatrix (Utils/projections.jl:25-69) as the full [nd x nd] matrix, column-major; pinned by
 * test/euler/Utils.jl:236-325 (P^T = P^-1, unit tangent orthogonal to the normal, round trip) */
void orc_projection_matrix(int dim, const double *nrm, double *Pout) {
  int nd = dim + 2;
  double Pm[3][3];
  projection_rows(dim, nrm, Pm);
  for (int i = 0; i < nd * nd; ++i) Pout[i] = 0.0;
  Pout[0] = 1.0;
  Pout[(nd - 1) + nd * (nd - 1)] = 1.0;
  for (int r = 0; r < dim; ++r)
    for (int c = 0; c < dim; ++c) Pout[(1 + r) + nd * (1 + c)] = Pm[r][c];
}

/* applyEntropyKernel(LW2Kernel) (faceElementIntegrals.jl:393-440): P^T Y |Lambda| S2 Y^T P delta_w * |nrm|, the eigensystem
 * taken in the face-normal direction by rotating q_avg into normal-tangential coordinates */
void orc_lw2_entropy_kernel(int dim, double gamma, const double *q_avg, const double *delta_w, const double *nrm_in,
                            double *flux) {
  int nd = dim + 2;
  double len_fac = 0.0, n[3] = {0, 0, 0}, Pm[3][3], qp[ORC_MAXD], t1[ORC_MAXD], t2v[ORC_MAXD];
  double Y[ORC_MAXD * ORC_MAXD], Lambda[ORC_MAXD], S2[ORC_MAXD];
  for (int d = 0; d < dim; ++d) len_fac += nrm_in[d] * nrm_in[d];
  len_fac = sqrt(len_fac);
  for (int d = 0; d < dim; ++d) n[d] = nrm_in[d] / len_fac;
  projection_rows(dim, n, Pm);
  /* projectToNT */
  qp[0] = q_avg[0]; qp[dim + 1] = q_avg[dim + 1];
  t1[0] = delta_w[0]; t1[dim + 1] = delta_w[dim + 1];
  for (int r = 0; r < dim; ++r) {
    double a = 0.0, b = 0.0;
    for (int c = 0; c < dim; ++c) { a += Pm[r][c] * q_avg[1 + c]; b += Pm[r][c] * delta_w[1 + c]; }
    qp[1 + r] = a; t1[1 + r] = b;
  }
  orc_evecs_x(dim, gamma, qp, Y);
  orc_evals_x(dim, gamma, qp, Lambda);
  orc_escaling_x(dim, gamma, qp, S2);
  for (int j = 0; j < nd; ++j) {                    /* smallmatTvec!: Y^T t1 */
    double s = 0.0;
    for (int r = 0; r < nd; ++r) s += Y[r + nd * j] * t1[r];
    t2v[j] = s * (len_fac * fabs(Lambda[j]) * S2[j]);
  }
  for (int r = 0; r < nd; ++r) {                    /* smallmatvec!: Y t2 */
    double s = 0.0;
    for (int j = 0; j < nd; ++j) s += Y[r + nd * j] * t2v[j];
    t1[r] = s;
  }
  /* projectToXY */
  flux[0] = t1[0]; flux[dim + 1] = t1[dim + 1];
  for (int c = 0; c < dim; ++c) {
    double a = 0.0;
    for (int r = 0; r < dim; ++r) a += Pm[r][c] * t1[1 + r];
    flux[1 + c] = a;
  }
}

static void calc_entropy_penalty_integral(const OrcProblem *P, const OrcInterface *f, const double *qL,
                                          const double *qR, const double *nrm_face, double *resL, double *resR) {
  int nd = P->nd, dim = P->dim, ss = P->ss, nfn = P->nfn;
  double wL[ORC_MAXD * 32], wR[ORC_MAXD * 32];
  for (int i = 0; i < ss; ++i) {
    orc_convert_to_ir(dim, P->gamma, qL + nd * (int)P->perm[i + ss * f->faceL], wL + nd * i);
    orc_convert_to_ir(dim, P->gamma, qR + nd * (int)P->perm[i + ss * f->faceR], wR + nd * i);
  }
  for (int i = 0; i < nfn; ++i) {
    int ni = (int)P->nbrperm[i + nfn * f->orient];
    const double *dir = nrm_face + dim * i;
    double wL_i[ORC_MAXD] = {0}, wR_i[ORC_MAXD] = {0}, qL_i[ORC_MAXD], qR_i[ORC_MAXD], q_avg[ORC_MAXD],
           delta_w[ORC_MAXD], flux[ORC_MAXD], A0[ORC_MAXD * ORC_MAXD];
    for (int j = 0; j < ss; ++j) {
      double interpL = P->interp[j + ss * i], interpR = P->interp[j + ss * ni];
      for (int k = 0; k < nd; ++k) { wL_i[k] += interpL * wL[k + nd * j]; wR_i[k] += interpR * wR[k + nd * j]; }
    }
    orc_convert_from_ir(dim, P->gamma, wL_i, qL_i);
    orc_convert_from_ir(dim, P->gamma, wR_i, qR_i);
    for (int j = 0; j < nd; ++j) { q_avg[j] = 0.5 * (qL_i[j] + qR_i[j]); delta_w[j] = wL_i[j] - wR_i[j]; }
    if (P->face_element_id == ORC_FEI_ELW2_PENALTY || P->face_element_id == ORC_FEI_ESLW2) {
      orc_lw2_entropy_kernel(dim, P->gamma, q_avg, delta_w, dir, flux);
    } else {
      /* applyEntropyKernel(LFKernel): lambda_max * A0 * delta_w */
      orc_ira0(dim, P->gamma, q_avg, A0);
      double lambda_max = orc_lambda_max(dim, P->gamma, q_avg, dir);
      for (int r = 0; r < nd; ++r) {
        double s = 0.0;
        for (int c = 0; c < nd; ++c) s += A0[r + nd * c] * delta_w[c];
        flux[r] = s * lambda_max;
      }
    }
    for (int j = 0; j < nd; ++j) flux[j] *= P->wface[i];
    for (int j = 0; j < ss; ++j) {
      int j_pL = (int)P->perm[j + ss * f->faceL], j_pR = (int)P->perm[j + ss * f->faceR];
      for (int p = 0; p < nd; ++p) {
        resL[p + nd * j_pL] -= P->interp[j + ss * i] * flux[p];
        resR[p + nd * j_pR] += P->interp[j + ss * ni] * flux[p];
      }
    }
  }
}

/* flux.jl:132-160 getFaceElementIntegral + the functors of faceElementIntegrals.jl:586-655 */
/* calcSharedFaceElementIntegrals_element_inner (flux.jl:442-496): the face-element integral of a shared face needs every
 * volume node of the remote element, received whole (getSendDataElement, Utils/parallel.jl:276-293); only elementL is
 * updated, the remote contribution goes to a throw-away array */
static void shared_face_element_integrals(const OrcProblem *P, const OrcPeer *peer, const double *q, double *res) {
  int nd = P->nd, nn = P->nn, dim = P->dim, nfn = P->nfn;
  double resR[ORC_MAXD * 32];
  for (int64_t j = 0; j < peer->nfaces; ++j) {
    const OrcInterface *f = &peer->interfaces[j];
    const double *qL = q + (int64_t)nd * nn * f->elementL;
    const double *qR = peer->q_recv_el + (int64_t)nd * nn * (f->elementR - peer->el_offset);
    double *resL = res + (int64_t)nd * nn * f->elementL;
    const double *nrm = peer->nrm_sharedface + (int64_t)dim * nfn * j;
    for (int k = 0; k < nd * nn; ++k) resR[k] = 0.0;
    if (P->face_element_id == ORC_FEI_EC || P->face_element_id == ORC_FEI_ESLF || P->face_element_id == ORC_FEI_ESLW2)
      calc_ec_face_integral(P, f, qL, qR, nrm, resL, resR);
    if (P->face_element_id == ORC_FEI_ELF_PENALTY || P->face_element_id == ORC_FEI_ESLF ||
        P->face_element_id == ORC_FEI_ELW2_PENALTY || P->face_element_id == ORC_FEI_ESLW2)
      calc_entropy_penalty_integral(P, f, qL, qR, nrm, resL, resR);
  }
}

static void face_element_integrals(const OrcProblem *P, const double *q, double *res) {
  int nd = P->nd, nn = P->nn, dim = P->dim, nfn = P->nfn;
  if (P->ss > 32) { fprintf(stderr, "oracle: stencil too large\n"); abort(); }
  for (int64_t i = 0; i < P->nF; ++i) {
    const OrcInterface *f = &P->ifaces[i];
    const double *qL = q + (int64_t)nd * nn * f->elementL, *qR = q + (int64_t)nd * nn * f->elementR;
    double *resL = res + (int64_t)nd * nn * f->elementL, *resR = res + (int64_t)nd * nn * f->elementR;
    const double *nrm = P->nrm_face + (int64_t)dim * nfn * i;
    if (P->face_element_id == ORC_FEI_EC || P->face_element_id == ORC_FEI_ESLF || P->face_element_id == ORC_FEI_ESLW2)
      calc_ec_face_integral(P, f, qL, qR, nrm, resL, resR);
    if (P->face_element_id == ORC_FEI_ELF_PENALTY || P->face_element_id == ORC_FEI_ESLF ||
        P->face_element_id == ORC_FEI_ELW2_PENALTY || P->face_element_id == ORC_FEI_ESLW2)
      calc_entropy_penalty_integral(P, f, qL, qR, nrm, resL, resR);
  }
}

/* source.jl:27-47 applySourceTerm */
static void apply_source_term(const OrcProblem *P, double *res) {
  int nd = P->nd, nn = P->nn, dim = P->dim;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < P->nE; ++e)
    for (int j = 0; j < nn; ++j) {
      double qv[ORC_MAXD];
      orc_src_exp(dim, P->gamma, P->coords + dim * (j + (int64_t)nn * e), qv);
      double fac = P->w[j] / P->jac[j + (int64_t)nn * e];
      for (int k = 0; k < nd; ++k) res[IDX3(k, j, e, nd, nn)] += fac * qv[k];
    }
}

/* Utils/parallel.jl:249-258 getSendDataFace (SBP boundaryinterpolate! into q_send) */
void orc_get_send_data_face(const OrcProblem *P, const double *q, OrcPeer *peer) {
  int nd = P->nd, nn = P->nn, nfn = P->nfn;
  for (int64_t j = 0; j < peer->nfaces; ++j)
    face_interp_el(P, q + (int64_t)nd * nn * peer->bndries_local[j].element,
                   peer->bndries_local[j].face, -1, peer->q_send + (int64_t)nd * nfn * j);
}

/* Utils/parallel.jl:198-201 permuteinterface! then flux.jl:264-308
 * calcSharedFaceIntegrals_nopre_inner (boundaryFaceIntegrate! into the local element) */
static void shared_face_integrals(const OrcProblem *P, OrcPeer *peer, double *res) {
  int nd = P->nd, nn = P->nn, nfn = P->nfn, dim = P->dim;
  for (int64_t j = 0; j < peer->nfaces; ++j) {
    OrcInterface I = peer->interfaces[j];
    double qR[ORC_MAXD * ORC_MAXFN], fl[ORC_MAXD * ORC_MAXFN];
    const double *recv = peer->q_recv + (int64_t)nd * nfn * j;
    for (int i = 0; i < nfn; ++i) {          /* permuteinterface!: node i <- nbrperm[i,orient] */
      int src = (int)P->nbrperm[i + nfn * I.orient];
      for (int k = 0; k < nd; ++k) qR[k + nd * i] = recv[k + nd * src];
    }
    const double *qL = peer->q_send + (int64_t)nd * nfn * j;
    for (int k = 0; k < nfn; ++k)
      face_flux_functor(P, P->flux_id, qL + nd * k, qR + nd * k,
                        peer->nrm_sharedface + dim * (k + (int64_t)nfn * j), fl + nd * k);
    face_integrate_el(P, fl, I.faceL, -1, -1.0, res + (int64_t)nd * nn * I.elementL);
  }
}

/* euler.jl:111-175 evalResidual.  precompute != 0 follows the default
 * precompute_* = true path of dataPrep (euler.jl:441-519); precompute == 0 the
 * fused *_nopre forms (euler_funcs.jl:158-196, flux.jl:79-125, bc.jl:290-328).
 * The caller must have filled peers[p].q_send (orc_get_send_data_face, the
 * reference's startSolutionExchange) and exchanged q_send -> the peer's q_recv.
 * Returns 0, 1 (negative density) or 2 (negative pressure). */
int orc_eval_residual(const OrcProblem *P, const double *q, double *res, double t, int precompute,
                      int npeers, OrcPeer *peers, int64_t *err_loc) {
  (void)t;  /* scoped BCs and SRCExp are time independent */
  int nd = P->nd, nn = P->nn, nfn = P->nfn;
  OrcWork W = work_alloc(P);
  memset(res, 0, sizeof(double) * nd * nn * P->nE);
  int status = aux_and_checks(P, q, W.aux_vars, err_loc);
  if (status) { work_free(&W); return status; }
  if (precompute) {
    get_euler_flux(P, q, W.flux_parametric);            /* runs even for split form (Appendix E.3) */
    if (!P->face_element_id) {
      interpolate_face(P, q, W.q_face);
      calc_face_flux(P, W.q_face, W.flux_face);
    }
  }
  interpolate_boundary(P, q, W.q_bndry);
  get_bc_fluxes(P, W.q_bndry, W.bndryflux);
  /* evalVolumeIntegrals euler.jl:628-658 */
  if (P->volume_integral_type == 1) {
    if (precompute)
      for (int d = 0; d < P->dim; ++d)
        weakdifferentiate_trans(P, d, W.flux_parametric + (int64_t)nd * nn * P->nE * d, res);
    else
      calc_volume_integrals_nopre(P, q, res);
  } else {
    calc_volume_integrals_split_linear(P, q, res);
  }
  /* evalBoundaryIntegrals euler.jl:669-690 */
  boundary_integrate(P, W.bndryflux, res);
  /* evalFaceIntegrals euler.jl:770-802 */
  if (P->face_element_id) {                              /* face_integral_type == 2 (euler.jl:783-793) */
    face_element_integrals(P, q, res);
  } else if (precompute) interior_face_integrate(P, W.flux_face, res);
  else calc_face_integral_nopre(P, q, res);
  /* evalSharedFaceIntegrals euler.jl:843-867 */
  for (int p = 0; p < npeers; ++p) {
    if (P->face_element_id) {
      if (!peers[p].q_recv_el) { fprintf(stderr, "oracle: face_integral_type 2 needs parallel_data = element\n"); abort(); }
      shared_face_element_integrals(P, &peers[p], q, res);
    } else {
      shared_face_integrals(P, &peers[p], res);
    }
  }
  /* evalSourceTerm euler.jl:889-901 */
  if (P->src_id == ORC_SRC_EXP) apply_source_term(P, res);
  (void)nfn;
  work_free(&W);
  return status;
}

/* individual passes exported for the integral-level golden tests */
void orc_volume_integrals(const OrcProblem *P, const double *q, double *res, int precompute) {
  OrcWork W = work_alloc(P);
  int nd = P->nd, nn = P->nn;
  if (P->volume_integral_type == 2) calc_volume_integrals_split_linear(P, q, res);
  else if (precompute) {
    get_euler_flux(P, q, W.flux_parametric);
    for (int d = 0; d < P->dim; ++d)
      weakdifferentiate_trans(P, d, W.flux_parametric + (int64_t)nd * nn * P->nE * d, res);
  } else calc_volume_integrals_nopre(P, q, res);
  work_free(&W);
}
void orc_euler_flux_parametric(const OrcProblem *P, const double *q, double *fp) { get_euler_flux(P, q, fp); }
void orc_face_integrals(const OrcProblem *P, const double *q, double *res, int precompute) {
  OrcWork W = work_alloc(P);
  if (P->face_element_id) face_element_integrals(P, q, res);
  else if (precompute) {
    interpolate_face(P, q, W.q_face);
    calc_face_flux(P, W.q_face, W.flux_face);
    interior_face_integrate(P, W.flux_face, res);
  } else calc_face_integral_nopre(P, q, res);
  work_free(&W);
}
void orc_boundary_integrals(const OrcProblem *P, const double *q, double *res, double *bndryflux_out) {
  OrcWork W = work_alloc(P);
  interpolate_boundary(P, q, W.q_bndry);
  get_bc_fluxes(P, W.q_bndry, W.bndryflux);
  boundary_integrate(P, W.bndryflux, res);
  if (bndryflux_out) memcpy(bndryflux_out, W.bndryflux, sizeof(double) * P->nd * P->nfn * P->nB);
  work_free(&W);
}
void orc_interpolate_boundary(const OrcProblem *P, const double *q, double *q_bndry) {
  interpolate_boundary(P, q, q_bndry);
}

/* Utils/mass_matrix.jl:20-44 calcMassMatrixInverse (DG dofs are consecutive) */
void orc_mass_matrix_inverse(const OrcProblem *P, double *Minv) {
  int nd = P->nd, nn = P->nn;
  for (int64_t e = 0; e < P->nE; ++e)
    for (int j = 0; j < nn; ++j)
      for (int k = 0; k < nd; ++k)
        Minv[IDX3(k, j, e, nd, nn)] = 1 / (P->w[j] / P->jac[j + (int64_t)nn * e]);
}

/* Utils/Utils.jl:427-449 calcNorm: sqrt(sum res*M*res), M = 1/Minv entrywise as eqn.M */
double orc_calc_norm(int64_t n, const double *M, const double *res_vec) {
  double val = 0.0;
  for (int64_t i = 0; i < n; ++i) val += res_vec[i] * M[i] * res_vec[i];
  return sqrt(val);
}

/* ------------------------------------------------------------------------ */
/* NonlinearSolvers/rk4.jl:144-344: rk4(f, h, t_max, q_vec, res_vec, pre_func, */
/* post_func, ctx, opts; res_tol, real_time)                                 */
/* ------------------------------------------------------------------------ */
typedef int (*orc_rhs_fn)(void *ctx, const double *q_vec, double *res_vec, double t);
/* post_func returns the norm when calc_norm != 0 (rk4.jl:446-457) */
typedef double (*orc_post_fn)(void *ctx, double *res_vec, int calc_norm);

/* Returns t as the reference does (rk4.jl:323-343).  norms_out (nullable)
 * receives sol_norm of each executed step (the convergence.dat column).
 * itermax < 0 disables the itermax test (use_itermax=false).  status_out gets
 * the first non-zero status of f (physics error), after which the loop stops. */
double orc_rk4(orc_rhs_fn f, orc_post_fn post, void *ctx, double h, double t_max, int64_t m,
               double *q_vec, double *res_vec, int64_t itermax, double res_tol, int real_time,
               double *norms_out, int64_t norms_cap, int64_t *nsteps_out, int *status_out) {
  double t = 0.0, treal = 0.0;
  int64_t t_steps = (int64_t)llround(t_max / h);
  double *x_old = (double *)malloc(sizeof(double) * m), *k1 = (double *)calloc(m, sizeof(double)),
         *k2 = (double *)calloc(m, sizeof(double)), *k3 = (double *)calloc(m, sizeof(double)),
         *k4 = (double *)calloc(m, sizeof(double));
  memcpy(x_old, q_vec, sizeof(double) * m);
  int64_t nsteps = 0;
  int status = 0;
  for (int64_t i = 2; i <= t_steps + 1; ++i) {
    t = (i - 2) * h;
    /* stage 1 */
    if (real_time) treal = t;
    if ((status = f(ctx, q_vec, res_vec, treal))) break;
    double sol_norm = post(ctx, res_vec, 1);
    for (int64_t j = 0; j < m; ++j) { k1[j] = res_vec[j]; q_vec[j] = x_old[j] + (h / 2) * k1[j]; }
    if (norms_out && nsteps < norms_cap) norms_out[nsteps] = sol_norm;
    ++nsteps;
    if ((sol_norm < res_tol) && !real_time) break;
    if (itermax >= 0 && i > itermax) break;
    /* stage 2 */
    if (real_time) treal = t + h / 2;
    if ((status = f(ctx, q_vec, res_vec, treal))) break;
    post(ctx, res_vec, 0);
    for (int64_t j = 0; j < m; ++j) { k2[j] = res_vec[j]; q_vec[j] = x_old[j] + (h / 2) * k2[j]; }
    /* stage 3 */
    if (real_time) treal = t + h / 2;
    if ((status = f(ctx, q_vec, res_vec, treal))) break;
    post(ctx, res_vec, 0);
    for (int64_t j = 0; j < m; ++j) { k3[j] = res_vec[j]; q_vec[j] = x_old[j] + h * k3[j]; }
    /* stage 4 */
    if (real_time) treal = t + h;
    if ((status = f(ctx, q_vec, res_vec, treal))) break;
    post(ctx, res_vec, 0);
    for (int64_t j = 0; j < m; ++j) k4[j] = res_vec[j];
    for (int64_t j = 0; j < m; ++j) {
      x_old[j] = x_old[j] + (h / 6) * (k1[j] + 2 * k2[j] + 2 * k3[j] + k4[j]);
      q_vec[j] = x_old[j];
    }
  }
  t += h;
  free(x_old); free(k1); free(k2); free(k3); free(k4);
  if (nsteps_out) *nsteps_out = nsteps;
  if (status_out) *status_out = status;
  return t;
}

/* rk4 driving the Euler residual of one (serial) problem: pde_pre_func is a
 * no-op for DG (Utils.jl:161-176), pde_post_func = res_vec *= Minv (+ norm). */
typedef struct { const OrcProblem *P; double *Minv, *M; int precompute; int64_t err_loc[2]; } EulerRkCtx;
static int euler_rhs(void *c, const double *q, double *res, double t) {
  EulerRkCtx *C = (EulerRkCtx *)c;
  return orc_eval_residual(C->P, q, res, t, C->precompute, 0, NULL, C->err_loc);
}
static double euler_post(void *c, double *res_vec, int calc_norm) {
  EulerRkCtx *C = (EulerRkCtx *)c;
  int64_t m = (int64_t)C->P->nd * C->P->nn * C->P->nE;
  for (int64_t j = 0; j < m; ++j) res_vec[j] = C->Minv[j] * res_vec[j];
  return calc_norm ? orc_calc_norm(m, C->M, res_vec) : 0.0;
}
double orc_rk4_euler(const OrcProblem *P, double *q_vec, double h, double t_max, int64_t itermax,
                     double res_tol, int real_time, int precompute, double *norms_out,
                     int64_t norms_cap, int64_t *nsteps_out, int *status_out) {
  int64_t m = (int64_t)P->nd * P->nn * P->nE;
  EulerRkCtx C;
  C.P = P; C.precompute = precompute;
  C.Minv = (double *)malloc(sizeof(double) * m);
  C.M = (double *)malloc(sizeof(double) * m);
  orc_mass_matrix_inverse(P, C.Minv);
  for (int64_t e = 0; e < P->nE; ++e)            /* mass_matrix.jl:62-82 calcMassMatrix */
    for (int j = 0; j < P->nn; ++j)
      for (int k = 0; k < P->nd; ++k)
        C.M[IDX3(k, j, e, P->nd, P->nn)] = P->w[j] / P->jac[j + (int64_t)P->nn * e];
  double *res_vec = (double *)malloc(sizeof(double) * m);
  double t = orc_rk4(euler_rhs, euler_post, &C, h, t_max, m, q_vec, res_vec, itermax, res_tol, real_time,
                     norms_out, norms_cap, nsteps_out, status_out);
  free(res_vec); free(C.Minv); free(C.M);
  return t;
}

/* ------------------------------------------------------------------------ */
/* NonlinearSolvers/lserk.jl:39-248: lserk54 (Carpenter-Kennedy 5-stage 2N-storage RK).  Note the stopping   */
/* tests come BEFORE the stage-1 update (lserk.jl:161-181), unlike rk4.                                      */
/* ------------------------------------------------------------------------ */
double orc_lserk54(orc_rhs_fn f, orc_post_fn post, void *ctx, double delta_t, double t_max, int64_t m,
                   double *q_vec, double *res_vec, int64_t itermax, double res_tol, int real_time,
                   double *norms_out, int64_t norms_cap, int64_t *nsteps_out, int *status_out) {
  const double a_coeffs[5] = {0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                              -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
  const double b_coeffs[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                              1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                              2277821191437.0 / 14882151754819.0};
  const double c_coeffs[5] = {0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                              2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0};
  double t = 0.0, treal = 0.0;
  int64_t t_steps = (int64_t)llround(t_max / delta_t);
  double *dq_vec = (double *)calloc(m, sizeof(double));
  int64_t nsteps = 0;
  int status = 0;
  for (int64_t i = 2; i <= t_steps + 1; ++i) {
    t = (i - 2) * delta_t;
    if (real_time) treal = t;
    if ((status = f(ctx, q_vec, res_vec, treal))) break;
    double sol_norm = post(ctx, res_vec, 1);
    if (norms_out && nsteps < norms_cap) norms_out[nsteps] = sol_norm;
    ++nsteps;
    if ((sol_norm < res_tol) && !real_time) break;
    if (itermax >= 0 && i > itermax) break;
    double fac = b_coeffs[0];
    for (int64_t j = 0; j < m; ++j) { dq_vec[j] = delta_t * res_vec[j]; q_vec[j] += fac * dq_vec[j]; }
    for (int stage = 1; stage < 5; ++stage) {
      if (real_time) treal = t + c_coeffs[stage] * delta_t;
      if ((status = f(ctx, q_vec, res_vec, treal))) break;
      post(ctx, res_vec, 0);
      fac = a_coeffs[stage];
      double fac2 = b_coeffs[stage];
      for (int64_t j = 0; j < m; ++j) { dq_vec[j] = fac * dq_vec[j] + delta_t * res_vec[j]; q_vec[j] += fac2 * dq_vec[j]; }
    }
    if (status) break;
  }
  t += delta_t;
  free(dq_vec);
  if (nsteps_out) *nsteps_out = nsteps;
  if (status_out) *status_out = status;
  return t;
}

double orc_lserk54_euler(const OrcProblem *P, double *q_vec, double h, double t_max, int64_t itermax,
                         double res_tol, int real_time, int precompute, double *norms_out,
                         int64_t norms_cap, int64_t *nsteps_out, int *status_out) {
  int64_t m = (int64_t)P->nd * P->nn * P->nE;
  EulerRkCtx C;
  C.P = P; C.precompute = precompute;
  C.Minv = (double *)malloc(sizeof(double) * m);
  C.M = (double *)malloc(sizeof(double) * m);
  orc_mass_matrix_inverse(P, C.Minv);
  for (int64_t e = 0; e < P->nE; ++e)
    for (int j = 0; j < P->nn; ++j)
      for (int k = 0; k < P->nd; ++k)
        C.M[IDX3(k, j, e, P->nd, P->nn)] = P->w[j] / P->jac[j + (int64_t)P->nn * e];
  double *res_vec = (double *)malloc(sizeof(double) * m);
  double t = orc_lserk54(euler_rhs, euler_post, &C, h, t_max, m, q_vec, res_vec, itermax, res_tol, real_time,
                         norms_out, norms_cap, nsteps_out, status_out);
  free(res_vec); free(C.Minv); free(C.M);
  return t;
}

/* ic.jl (macro-generated ICs): evaluate calc<Name>(params, coords_j, sol) at n nodes.
 * kind: 1 ICIsentropicVortex, 2 ICExp, 3 ICFreeStream */
void orc_fill_exact(const OrcProblem *P, int kind, const double *coords, int64_t n, double *out) {
  int dim = P->dim, nd = P->nd;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    if (kind == 1) orc_isentropic_vortex(dim, P->gamma, P->R, coords + dim * i, out + nd * i);
    else if (kind == 2) orc_calc_exp(dim, P->gamma, coords + dim * i, out + nd * i);
    else orc_free_stream(dim, P->rho_free, P->E_free, P->Ma, P->aoa, out + nd * i);
  }
}
