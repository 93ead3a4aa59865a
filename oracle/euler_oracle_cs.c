/*
 * euler_oracle_cs.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The reference's Jacobian-vector product restated as the reference computes it: the residual evaluated in COMPLEX
 * arithmetic at q + i*eps*v, eps = 1e-20, and J*v = imag(R)/eps (applyLinearOperator NonlinearSolvers/newton_setup.jl:632-662,
 * evaldRdqProduct interface2.jl:454-498, physicsRhs jacobian/residual_evaluation.jl:64-88; evalResidual is generic in Tsol,
 * types Tsol = Complex128, Tmsh = Float64 from getDataTypes solver/common.jl).  Complex-step semantics of Utils/complexify.jl:
 * absvalue flips the sign by the real part (:25-31), max / isless compare real parts (:207-217), absvalue3 (:157-172).
 *
 * Scope: the dense-face Roe path (volume_integral_type 1, face_integral_type 1, RoeFlux, the eight boundary functors);
 * the source term is state independent and drops out.  This file #includes euler_oracle.c (structs, real-valued helpers)
 * and is built as its own library, liborc_cs.so.  It pins the CUDA path's dual-number J*v to round-off
 * (tests/test_gpu_parity.py::test_jvp_matches_complex_step_oracle) where central differences only reach 1e-8.
 */
#include "euler_oracle.c"      /* (before <complex.h>: the restatement uses `I` as a variable name) */
#include <complex.h>

typedef double complex cplx;

static cplx c_absvalue(cplx x) { return creal(x) < 0.0 ? -x : x; }                /* complexify.jl:25-31 */
static cplx c_max(cplx a, cplx b) { return creal(a) < creal(b) ? b : a; }         /* isless on real parts, :207-217 */
static cplx c_absvalue3(cplx val) {                                               /* :157-172 */
  const double delta = 1e-7;
  cplx v1 = c_absvalue(val);
  if (creal(v1) > delta) return v1;
  return ((val * val) / delta + delta) / 2;
}

/* euler_funcs.jl:856-863, 897-903 */
static cplx c_pressure(int dim, double gamma, const cplx *q) {
  cplx ke = 0.0;
  for (int d = 0; d < dim; ++d) ke += q[1 + d] * q[1 + d];
  return (gamma - 1.0) * (q[dim + 1] - 0.5 * ke / q[0]);
}

/* euler_funcs.jl:512-536, 749-774 */
static void c_euler_flux(int dim, double gamma, const cplx *q, const double *dir, cplx *F) {
  cplx press = c_pressure(dim, gamma, q);
  cplx U = 0.0;
  for (int d = 0; d < dim; ++d) U += q[1 + d] * dir[d];
  U /= q[0];
  F[0] = q[0] * U;
  for (int d = 0; d < dim; ++d) F[1 + d] = q[1 + d] * U + dir[d] * press;
  F[dim + 1] = (q[dim + 1] + press) * U;
}

/* bc_solvers.jl:207-310, 313-420 calcSAT */
static void c_calc_sat(int dim, double gamma, const cplx *vel, cplx H, const cplx *dq, const double *nrm, cplx *sat) {
  const double sat_Vn = 0.025, sat_Vl = 0.025, tau = 1.0;
  double gami = gamma - 1.0, dA = 0.0;
  int nd = dim + 2;
  cplx Un = 0.0, phi = 0.0;
  for (int d = 0; d < dim; ++d) { dA += nrm[d] * nrm[d]; phi += vel[d] * vel[d]; Un += vel[d] * nrm[d]; }
  dA = sqrt(dA);
  phi *= 0.5;
  cplx a = csqrt(gami * (H - phi));
  cplx lambda1 = Un + dA * a, lambda2 = Un - dA * a, lambda3 = Un;
  cplx rhoA = c_absvalue(Un) + dA * a;
  lambda1 = 0.5 * (tau * c_max(c_absvalue(lambda1), sat_Vn * rhoA) - lambda1);
  lambda2 = 0.5 * (tau * c_max(c_absvalue(lambda2), sat_Vn * rhoA) - lambda2);
  lambda3 = 0.5 * (tau * c_max(c_absvalue(lambda3), sat_Vl * rhoA) - lambda3);
  cplx E1dq[ORC_MAXD], E2dq[ORC_MAXD];
  for (int i = 0; i < nd; ++i) sat[i] = lambda3 * dq[i];
  cplx e1 = phi * dq[0];
  for (int d = 0; d < dim; ++d) e1 -= vel[d] * dq[1 + d];
  e1 += dq[dim + 1];
  E1dq[0] = e1;
  for (int d = 0; d < dim; ++d) E1dq[1 + d] = e1 * vel[d];
  E1dq[dim + 1] = e1 * H;
  cplx e2 = -Un * dq[0];
  for (int d = 0; d < dim; ++d) e2 += nrm[d] * dq[1 + d];
  E2dq[0] = 0.0;
  for (int d = 0; d < dim; ++d) E2dq[1 + d] = e2 * nrm[d];
  E2dq[dim + 1] = e2 * Un;
  cplx tmp1 = 0.5 * (lambda1 + lambda2) - lambda3;
  cplx tmp2 = gami / (a * a);
  double tmp3 = 1.0 / (dA * dA);
  for (int i = 0; i < nd; ++i) sat[i] = sat[i] + tmp1 * (tmp2 * E1dq[i] + tmp3 * E2dq[i]);
  E1dq[0] = e2;
  for (int d = 0; d < dim; ++d) E1dq[1 + d] = e2 * vel[d];
  E1dq[dim + 1] = e2 * H;
  E2dq[0] = 0.0;
  for (int d = 0; d < dim; ++d) E2dq[1 + d] = e1 * nrm[d];
  E2dq[dim + 1] = e1 * Un;
  tmp1 = 0.5 * (lambda1 - lambda2) / (dA * a);
  for (int i = 0; i < nd; ++i) sat[i] = sat[i] + tmp1 * (E1dq[i] + gami * E2dq[i]);
}

/* bc_solvers.jl:29-103, 111-187 RoeSolver */
static void c_roe_solver(int dim, double gamma, const cplx *q, const cplx *qg, const double *nrm, cplx *flux) {
  int nd = dim + 2;
  double gami = gamma - 1.0;
  cplx fac = 1.0 / q[0], velL[3], velR[3], phi = 0.0;
  for (int d = 0; d < dim; ++d) { velL[d] = q[1 + d] * fac; phi += velL[d] * velL[d]; }
  phi *= 0.5;
  cplx HL = gamma * q[dim + 1] * fac - gami * phi;
  fac = 1.0 / qg[0];
  phi = 0.0;
  for (int d = 0; d < dim; ++d) { velR[d] = qg[1 + d] * fac; phi += velR[d] * velR[d]; }
  phi *= 0.5;
  cplx HR = gamma * qg[dim + 1] * fac - gami * phi;
  cplx sqL = csqrt(q[0]), sqR = csqrt(qg[0]);
  fac = 1.0 / (sqL + sqR);
  cplx vel[3];
  for (int d = 0; d < dim; ++d) vel[d] = (sqL * velL[d] + sqR * velR[d]) * fac;
  cplx H = (sqL * HL + sqR * HR) * fac;
  cplx dq[ORC_MAXD], sat[ORC_MAXD], ef[ORC_MAXD];
  for (int i = 0; i < nd; ++i) dq[i] = q[i] - qg[i];
  c_calc_sat(dim, gamma, vel, H, dq, nrm, sat);
  c_euler_flux(dim, gamma, q, nrm, ef);
  for (int i = 0; i < nd; ++i) flux[i] = sat[i] + ef[i];
}

/* euler_funcs.jl:1887-1913 getLambdaMax */
static cplx c_lambda_max(int dim, double gamma, const cplx *qL, const double *dir) {
  cplx Un = 0.0, rhoLinv = 1 / qL[0];
  double dA = 0.0;
  cplx pL = c_pressure(dim, gamma, qL);
  cplx aL = csqrt(gamma * pL * rhoLinv);
  for (int i = 0; i < dim; ++i) { Un += dir[i] * qL[i + 1] * rhoLinv; dA += dir[i] * dir[i]; }
  return c_absvalue3(Un) + sqrt(dA) * aL;
}

/* the boundary functors (bc.jl:554-567, 1756-1768, 1573-1587, 717-765, 1454-1537, 1702-1722, 2140-2152, 767-793): the
 * Dirichlet states are real (they do not depend on q) */
static void c_bc_flux(const OrcProblem *P, int bc_id, const cplx *q, const double *coords, const double *nrm, cplx *flux) {
  int dim = P->dim, nd = dim + 2;
  double qd[ORC_MAXD];
  cplx qg[ORC_MAXD];
  int dirichlet = 1;
  switch (bc_id) {
    case ORC_BC_ISENTROPIC_VORTEX: orc_isentropic_vortex(dim, P->gamma, P->R, coords, qd); break;
    case ORC_BC_EXP: orc_calc_exp(dim, P->gamma, coords, qd); break;
    case ORC_BC_FREESTREAM: orc_free_stream(dim, P->rho_free, P->E_free, P->Ma, P->aoa, qd); break;
    case ORC_BC_RHO1E2U3:
      qd[0] = 1.0;
      for (int d = 0; d < dim; ++d) qd[1 + d] = 0.35355;
      qd[dim + 1] = 2.0;
      break;
    case ORC_BC_ALLONES:
      for (int i = 0; i < nd; ++i) qd[i] = 1.0;
      break;
    default: dirichlet = 0;
  }
  if (dirichlet) {
    for (int i = 0; i < nd; ++i) qg[i] = qd[i];
    c_roe_solver(dim, P->gamma, q, qg, nrm, flux);
    return;
  }
  if (bc_id == ORC_BC_ZEROFLUX) {
    for (int i = 0; i < nd; ++i) flux[i] = 0.0;
    return;
  }
  double n[3], nn2 = 0.0;
  cplx Unrm = 0.0;
  for (int d = 0; d < dim; ++d) nn2 += nrm[d] * nrm[d];
  double fac = 1.0 / sqrt(nn2);
  for (int d = 0; d < dim; ++d) { n[d] = nrm[d] * fac; Unrm += n[d] * q[1 + d]; }
  for (int i = 0; i < nd; ++i) qg[i] = q[i];
  if (bc_id == ORC_BC_NOPENETRATION) {
    for (int d = 0; d < dim; ++d) qg[1 + d] -= n[d] * Unrm;
    c_euler_flux(dim, P->gamma, qg, nrm, flux);
  } else if (bc_id == ORC_BC_NOPENETRATION_ES) {
    cplx fluxL[ORC_MAXD], fluxR[ORC_MAXD], q_avg[ORC_MAXD];
    for (int d = 0; d < dim; ++d) qg[1 + d] = -2 * Unrm * n[d] + q[1 + d];
    c_euler_flux(dim, P->gamma, q, nrm, fluxL);
    c_euler_flux(dim, P->gamma, qg, nrm, fluxR);
    for (int i = 0; i < nd; ++i) q_avg[i] = 0.5 * (q[i] + qg[i]);
    cplx lambda_max = c_lambda_max(dim, P->gamma, q_avg, nrm);
    for (int i = 0; i < nd; ++i) flux[i] = 0.5 * (fluxL[i] + fluxR[i] - lambda_max * (qg[i] - q[i]));
  } else {
    fprintf(stderr, "oracle_cs: unsupported BC id %d\n", bc_id);
    abort();
  }
}

/* interiorFaceInterpolate! / boundaryinterpolate! (dense faces): uface[:,i] = sum_j interp[j,col(i)] u[:,perm[j,face]] */
static void c_face_interp_el(const OrcProblem *P, const cplx *u_el, int face, int orient_or_neg, cplx *uface) {
  int nd = P->nd, nfn = P->nfn, ss = P->ss;
  for (int i = 0; i < nfn; ++i) {
    int col = orient_or_neg < 0 ? i : (int)P->nbrperm[i + nfn * orient_or_neg];
    cplx *uf = uface + nd * i;
    for (int k = 0; k < nd; ++k) uf[k] = 0.0;
    for (int j = 0; j < ss; ++j) {
      double c = P->interp[j + ss * col];
      int64_t node = P->perm[j + (int64_t)ss * face];
      for (int k = 0; k < nd; ++k) uf[k] += c * u_el[k + nd * node];
    }
  }
}
static void c_face_integrate_el(const OrcProblem *P, const cplx *flux, int face, int orient_or_neg, double sgn, cplx *res_el) {
  int nd = P->nd, nfn = P->nfn, ss = P->ss;
  for (int i = 0; i < nfn; ++i) {
    int col = orient_or_neg < 0 ? i : (int)P->nbrperm[i + nfn * orient_or_neg];
    const cplx *f = flux + nd * i;
    for (int j = 0; j < ss; ++j) {
      double c = P->interp[j + ss * col] * P->wface[i];
      int64_t node = P->perm[j + (int64_t)ss * face];
      for (int k = 0; k < nd; ++k) res_el[k + nd * node] += sgn * c * f[k];
    }
  }
}

/* evalResidual (euler.jl:111-175) in Complex128 at q + i*eps*v; out = imag(res)/eps.  Returns 0, or -1 for an unsupported
 * configuration. */
int orc_eval_jvp_complex_step(const OrcProblem *P, const double *q, const double *v, double eps, double *out) {
  if (P->sparse_face || P->volume_integral_type != 1 || P->face_element_id != 0 || P->flux_id != ORC_FLUX_ROE) return -1;
  const int dim = P->dim, nd = P->nd, nn = P->nn, nfn = P->nfn;
  const int64_t el = (int64_t)nd * nn, n = el * P->nE;
  cplx *qc = (cplx *)malloc(sizeof(cplx) * n), *res = (cplx *)calloc(n, sizeof(cplx));
  for (int64_t i = 0; i < n; ++i) qc[i] = q[i] + I * (eps * v[i]);
  /* volume integrals: getEulerFlux (euler_funcs.jl:23-58) + weakdifferentiate!(trans = true) (euler.jl:628-658) */
  for (int64_t e = 0; e < P->nE; ++e)
    for (int j = 0; j < nn; ++j) {
      const double *dx = P->dxidx + (int64_t)dim * dim * (j + (int64_t)nn * e);
      for (int d = 0; d < dim; ++d) {
        double dir[3];
        cplx F[ORC_MAXD];
        for (int p = 0; p < dim; ++p) dir[p] = dx[d + dim * p];
        c_euler_flux(dim, P->gamma, qc + el * e + nd * j, dir, F);
        const double *Qd = P->Q + (int64_t)nn * nn * d;
        for (int i = 0; i < nn; ++i) {
          double c = Qd[j + nn * i];
          for (int k = 0; k < nd; ++k) res[el * e + nd * i + k] += c * F[k];
        }
      }
    }
  /* boundary integrals (bc.jl:251-284, euler.jl:669-690): res -= R^T W flux */
  for (int b = 0; b < P->numBC; ++b)
    for (int64_t i = P->bndry_offsets[b]; i < P->bndry_offsets[b + 1]; ++i) {
      const OrcBoundary *bf = P->bfaces + i;
      cplx qb[ORC_MAXD * ORC_MAXFN], fb[ORC_MAXD * ORC_MAXFN];
      c_face_interp_el(P, qc + el * bf->element, bf->face, -1, qb);
      for (int k = 0; k < nfn; ++k)
        c_bc_flux(P, P->bc_ids[b], qb + nd * k, P->coords_bndry + (int64_t)dim * (k + (int64_t)nfn * i),
                  P->nrm_bndry + (int64_t)dim * (k + (int64_t)nfn * i), fb + nd * k);
      c_face_integrate_el(P, fb, bf->face, -1, -1.0, res + el * bf->element);
    }
  /* interior faces (flux.jl:79-125): Roe flux at every face node, integrated into both elements */
  for (int64_t f = 0; f < P->nF; ++f) {
    const OrcInterface *it = P->ifaces + f;
    cplx qL[ORC_MAXD * ORC_MAXFN], qR[ORC_MAXD * ORC_MAXFN], fl[ORC_MAXD * ORC_MAXFN];
    c_face_interp_el(P, qc + el * it->elementL, it->faceL, -1, qL);
    c_face_interp_el(P, qc + el * it->elementR, it->faceR, it->orient, qR);
    for (int k = 0; k < nfn; ++k)
      c_roe_solver(dim, P->gamma, qL + nd * k, qR + nd * k, P->nrm_face + (int64_t)dim * (k + (int64_t)nfn * f), fl + nd * k);
    c_face_integrate_el(P, fl, it->faceL, -1, -1.0, res + el * it->elementL);
    c_face_integrate_el(P, fl, it->faceR, it->orient, 1.0, res + el * it->elementR);
  }
  for (int64_t i = 0; i < n; ++i) out[i] = cimag(res[i]) / eps;
  free(qc);
  free(res);
  return 0;
}
