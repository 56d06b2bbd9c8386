/*
 * pnec_oracle.c — CPU restatement of the reference's PNEC/NEC refinement.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pnec_b200/ may link, import or call
 * this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs can check and time the CUDA path
 * against the reference's algorithm.
 *
 * PARITY STATUS
 *   - residual functors: PINNED.  Checked against the reference's own numpy
 *     energies (scripts/pnec/common.py:13-58) through tests/golden/.
 *   - Levenberg-Marquardt loop, numeric differentiation, quaternion manifold:
 *     PARITY UNPINNED.  That arithmetic lives in ceres-solver (>= 2.1, not
 *     vendored by the reference, not installed here; Dockerfile:17 pins a stale
 *     1.13.0).  It is restated below from upstream Ceres 2.x behaviour
 *     (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc,
 *     numeric_diff.h, manifold.cc) as documented in SURVEY.md section 8c.  The
 *     reference holds no golden vectors for it.
 *
 *   - front stages of PNEC::Solve (opengv's rotation eigensolver, weighted eigensolver,
 *     orchestration): pnec_oracle_frame.c, included at the end of this file; its header states
 *     what is pinned (the LM against MINPACK) and what is not (that it is what opengv executes).
 *
 * What is restated, with the reference lines it follows:
 *   functor_*            include/optimization/pnec_residual.h:50-150,
 *                        include/optimization/nec_residual.h:47-69
 *   evaluate_numeric     ceres::NumericDiffCostFunction<F, CENTRAL, 1,1,1,4>
 *                        as instantiated at src/optimization/pnec_ceres.cc:83-87,
 *                        94-98, 124-128 and src/optimization/nec_ceres.cc:82-85,
 *                        composed with ceres::EigenQuaternionManifold
 *                        (pnec_ceres.cc:103-106)
 *   oracle_solve         ceres::Solve(options_ [defaults], ...) at
 *                        pnec_ceres.cc:110,167 / nec_ceres.cc:97
 *   angles_from_vec      src/common/common.cc:103-116
 *   result pose          PNECCeres::Result, pnec_ceres.cc:201-206
 *   metrics              src/common/common.cc:210-259
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { V_NEC = 0, V_TARGET = 1, V_HOST = 2, V_SYMMETRIC = 3 };
enum { JAC_NUMERIC_CENTRAL = 0, JAC_ANALYTIC = 1 };

enum {
  ST_CONVERGED_FUNCTION = 0,
  ST_CONVERGED_PARAMETER = 1,
  ST_CONVERGED_GRADIENT = 2,
  ST_CONVERGED_RADIUS = 3,
  ST_MAX_ITERATIONS = 4,
  ST_FAILURE = 5,
  ST_NONFINITE = 6,
  ST_EMPTY = 7
};

/* Mirrors pnec_solver_opts (include/pnec_b200.h) field for field, plus the
 * oracle-only jacobian_mode. */
typedef struct oracle_opts {
  int32_t variant;
  int32_t max_num_iterations;
  int32_t max_num_consecutive_invalid_steps;
  int32_t jacobi_scaling;
  double regularization;
  double function_tolerance;
  double gradient_tolerance;
  double parameter_tolerance;
  double initial_trust_region_radius;
  double max_trust_region_radius;
  double min_trust_region_radius;
  double min_relative_decrease;
  double min_lm_diagonal;
  double max_lm_diagonal;
  int32_t jacobian_mode;
  int32_t reserved;
} oracle_opts;

void oracle_opts_default(oracle_opts *o) {
  o->variant = V_TARGET;
  o->max_num_iterations = 50;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->regularization = 1.0e-13;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->jacobian_mode = JAC_NUMERIC_CENTRAL;
  o->reserved = 0;
}

/* ---------------------------------------------------------------- helpers */

/* Eigen::Quaternion::toRotationMatrix() evaluated exactly as Eigen writes it,
 * which matters because Ceres' numeric differentiation feeds it non-unit
 * quaternions.  q = (x, y, z, w) (Eigen coeffs() order).  R row-major. */
static void quat_to_rot(const double q[4], double R[3][3]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0][0] = 1.0 - (tyy + tzz);
  R[0][1] = txy - twz;
  R[0][2] = txz + twy;
  R[1][0] = txy + twz;
  R[1][1] = 1.0 - (txx + tzz);
  R[1][2] = tyz - twx;
  R[2][0] = txz - twy;
  R[2][1] = tyz + twx;
  R[2][2] = 1.0 - (txx + tyy);
}

/* pnec::common::SkewFromVector, src/common/common.cc:96-101. */
static void skew(const double v[3], double S[3][3]) {
  S[0][0] = 0.0;   S[0][1] = -v[2]; S[0][2] = v[1];
  S[1][0] = v[2];  S[1][1] = 0.0;   S[1][2] = -v[0];
  S[2][0] = -v[1]; S[2][1] = v[0];  S[2][2] = 0.0;
}

static void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static void matvec(const double M[3][3], const double v[3], double o[3]) {
  for (int i = 0; i < 3; ++i) o[i] = M[i][0] * v[0] + M[i][1] * v[1] + M[i][2] * v[2];
}
/* row-vector times matrix: o = v^T M */
static void vecmat(const double v[3], const double M[3][3], double o[3]) {
  for (int j = 0; j < 3; ++j) o[j] = v[0] * M[0][j] + v[1] * M[1][j] + v[2] * M[2][j];
}
static void matmul(const double A[3][3], const double B[3][3], double C[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
}
static void transpose(const double A[3][3], double T[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i][j] = A[j][i];
}
/* column-major double[9] (Eigen::Matrix3d storage) -> [row][col] */
static void load_cov(const double *c, double S[3][3]) {
  for (int col = 0; col < 3; ++col)
    for (int row = 0; row < 3; ++row) S[row][col] = c[col * 3 + row];
}

static void translation_from_angles(double theta, double phi, double t[3]) {
  t[0] = sin(theta) * cos(phi);
  t[1] = sin(theta) * sin(phi);
  t[2] = cos(theta);
}

/* pnec::common::AnglesFromVec, src/common/common.cc:103-116. */
void oracle_angles_from_vec(const double v[3], double *theta, double *phi) {
  const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (n == 0.0) {
    *theta = 0.0;
    *phi = 0.0;
    return;
  }
  const double nv[3] = {v[0] / n, v[1] / n, v[2] / n};
  *theta = acos(nv[2]);
  if (fabs(*theta) < 1e-10) {
    *phi = 0.0;
  } else {
    *phi = atan2(nv[1], nv[0]);
  }
}

/* --------------------------------------------------------------- functors */

/* One call of the residual functor, in the literal evaluation order of the
 * Eigen expressions (row vector pushed through the matrix chain left to
 * right; for PNECSymmetrical the two 3x3 terms are formed and added first). */
static double functor(int variant, double theta, double phi, const double q[4],
                      const double f1[3], const double f2[3], const double *cov_t,
                      const double *cov_h, double reg) {
  double t[3], R[3][3], Rf2[3], n[3];
  translation_from_angles(theta, phi, t);
  quat_to_rot(q, R);
  matvec(R, f2, Rf2);
  cross3(f1, Rf2, n);
  const double num = dot3(t, n);
  if (variant == V_NEC) return num; /* nec_residual.h:60-61 */

  double S[3][3], Rt[3][3], F[3][3], Ft[3][3], v[3], u[3];
  transpose(R, Rt);
  if (variant == V_TARGET) { /* pnec_residual.h:94-102 */
    load_cov(cov_t, S);
    skew(f1, F);
    transpose(F, Ft);
    vecmat(t, F, v);
    vecmat(v, R, u);
    vecmat(u, S, v);
    vecmat(v, Rt, u);
    vecmat(u, Ft, v);
    return num / sqrt(dot3(v, t) + reg);
  }
  if (variant == V_HOST) { /* pnec_residual.h:63-70: skew of the ROTATED f1 */
    double Rf1[3];
    load_cov(cov_t, S);
    matvec(R, f1, Rf1);
    skew(Rf1, F);
    transpose(F, Ft);
    vecmat(t, F, v);
    vecmat(v, S, u);
    vecmat(u, Ft, v);
    return num / sqrt(dot3(v, t) + reg);
  }
  /* V_SYMMETRIC, pnec_residual.h:128-140 */
  double S1[3][3], S2[3][3], G[3][3], Gt[3][3], A[3][3], B[3][3], T1[3][3], T2[3][3];
  load_cov(cov_h, S1);
  load_cov(cov_t, S2);
  skew(f1, F);
  transpose(F, Ft);
  skew(Rf2, G);
  transpose(G, Gt);
  matmul(F, R, T1);
  matmul(T1, S2, T2);
  matmul(T2, Rt, T1);
  matmul(T1, Ft, A);
  matmul(G, S1, T1);
  matmul(T1, Gt, B);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] += B[i][j];
  vecmat(t, A, v);
  return num / sqrt(dot3(v, t) + reg);
}

/* ---------------------------------------------------- manifold (Ceres 2.x) */

/* ceres::EigenQuaternionManifold::Plus: q+ = [sin|d| d/|d|, cos|d|] (x) q,
 * storage (x, y, z, w). */
static void quat_plus(const double x[4], const double d[3], double out[4]) {
  const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd == 0.0) {
    memcpy(out, x, 4 * sizeof(double));
    return;
  }
  const double s = sin(nd) / nd;
  const double dw = cos(nd), dx = s * d[0], dy = s * d[1], dz = s * d[2];
  const double xx = x[0], xy = x[1], xz = x[2], xw = x[3];
  out[3] = dw * xw - dx * xx - dy * xy - dz * xz;
  out[0] = dw * xx + dx * xw + dy * xz - dz * xy;
  out[1] = dw * xy - dx * xz + dy * xw + dz * xx;
  out[2] = dw * xz + dx * xy - dy * xx + dz * xw;
}

/* ceres::EigenQuaternionManifold::PlusJacobian, 4x3, rows in storage order
 * (x, y, z, w). */
static void quat_plus_jacobian(const double x[4], double J[4][3]) {
  const double qx = x[0], qy = x[1], qz = x[2], qw = x[3];
  J[3][0] = -qx; J[3][1] = -qy; J[3][2] = -qz;
  J[0][0] = qw;  J[0][1] = qz;  J[0][2] = -qy;
  J[1][0] = -qz; J[1][1] = qw;  J[1][2] = qx;
  J[2][0] = qy;  J[2][1] = -qx; J[2][2] = qw;
}

/* Program::Plus over the three parameter blocks: x = (theta, phi, qx qy qz qw),
 * delta = (dtheta, dphi, d1 d2 d3). */
static void state_plus(const double x[6], const double delta[5], double out[6]) {
  out[0] = x[0] + delta[0];
  out[1] = x[1] + delta[1];
  quat_plus(x + 2, delta + 2, out + 2);
}

/* ------------------------------------------------------------- evaluation */

typedef struct problem_view {
  int variant;
  int64_t n;
  const double *f1, *f2, *cov_t, *cov_h;
  double reg;
} problem_view;

static double residual_at(const problem_view *p, int64_t i, const double x[6]) {
  return functor(p->variant, x[0], x[1], x + 2, p->f1 + 3 * i, p->f2 + 3 * i,
                 p->cov_t ? p->cov_t + 9 * i : NULL, p->cov_h ? p->cov_h + 9 * i : NULL,
                 p->reg);
}

/* cost-only evaluation: 1/2 sum r^2, accumulated block by block like
 * ProgramEvaluator. */
static double evaluate_cost(const problem_view *p, const double x[6], double *residuals) {
  double cost = 0.0;
  for (int64_t i = 0; i < p->n; ++i) {
    const double r = residual_at(p, i, x);
    if (residuals) residuals[i] = r;
    cost += 0.5 * r * r;
  }
  return cost;
}

/* ceres::internal::NumericDiff<..., CENTRAL>::EvaluateJacobianForParameterBlock
 * for one scalar residual: per ambient coordinate j,
 *   h = max(|x_j| * relative_step_size(1e-6), sqrt(DBL_EPSILON)),
 *   d r / d x_j = (r(x + h e_j) - r(x - h e_j)) * (1/h / 2). */
static void numeric_row(const problem_view *p, int64_t i, const double x[6], double amb[6]) {
  const double min_step = sqrt(DBL_EPSILON);
  double xp[6];
  memcpy(xp, x, sizeof(xp));
  /* each parameter block (theta | phi | q) is differenced on its own, but the
   * other blocks are held at x, so one shared scratch is equivalent */
  for (int j = 0; j < 6; ++j) {
    double h = fabs(x[j]) * 1e-6;
    if (h < min_step) h = min_step;
    xp[j] = x[j] + h;
    double r = residual_at(p, i, xp);
    double one_over_delta = 1.0 / h;
    xp[j] = x[j] - h;
    const double rm = residual_at(p, i, xp);
    r -= rm;
    one_over_delta /= 2;
    xp[j] = x[j];
    amb[j] = r * one_over_delta;
  }
}

/* Analytic local Jacobian row — not what the reference
 * does (it differentiates numerically); kept as an independent cross-check of
 * the CUDA kernels' closed form at tighter tolerance than central differences
 * allow. */
static void analytic_row(const problem_view *p, int64_t i, const double x[6], double *r_out,
                         double row[5]) {
  const double *f1 = p->f1 + 3 * i, *f2 = p->f2 + 3 * i;
  const double st = sin(x[0]), ct = cos(x[0]), sp = sin(x[1]), cp = cos(x[1]);
  const double t[3] = {st * cp, st * sp, ct};
  const double t_th[3] = {ct * cp, ct * sp, -st};
  const double t_ph[3] = {-st * sp, st * cp, 0.0};
  double R[3][3], Rt[3][3], g[3], a[3], de_dt[3], de_dw[3];
  quat_to_rot(x + 2, R);
  transpose(R, Rt);
  matvec(R, f2, g);
  cross3(t, f1, a);
  const double e = dot3(a, g);
  cross3(f1, g, de_dt);
  cross3(g, a, de_dw);
  double ds_dt[3] = {0, 0, 0}, ds_dw[3] = {0, 0, 0}; /* derivative of s^2 */
  double s2 = 1.0;
  if (p->variant != V_NEC) {
    s2 = p->reg;
    double S[3][3];
    if (p->variant == V_TARGET || p->variant == V_SYMMETRIC) {
      double b[3], Sb[3], RSb[3], tmp[3];
      load_cov(p->cov_t + 9 * i, S);
      matvec(Rt, a, b);
      matvec(S, b, tmp);
      /* symmetric part only: x^T S x == x^T sym(S) x */
      double St[3][3], tmp2[3];
      transpose(S, St);
      matvec(St, b, tmp2);
      for (int k = 0; k < 3; ++k) Sb[k] = 0.5 * (tmp[k] + tmp2[k]);
      s2 += dot3(b, Sb);
      matvec(R, Sb, RSb);
      cross3(f1, RSb, tmp);
      for (int k = 0; k < 3; ++k) ds_dt[k] += 2.0 * tmp[k];
      cross3(RSb, a, tmp);
      for (int k = 0; k < 3; ++k) ds_dw[k] += 2.0 * tmp[k];
    }
    if (p->variant == V_HOST || p->variant == V_SYMMETRIC) {
      /* c = t x h with h = R f1 (HOST, covariance cov_t) or h = R f2
       * (SYMMETRIC second term, covariance cov_h) */
      double h[3], c[3], Sc[3], tmp[3], tmp2[3], St[3][3];
      if (p->variant == V_HOST) {
        load_cov(p->cov_t + 9 * i, S);
        matvec(R, f1, h);
      } else {
        load_cov(p->cov_h + 9 * i, S);
        memcpy(h, g, sizeof(h));
      }
      cross3(t, h, c);
      matvec(S, c, tmp);
      transpose(S, St);
      matvec(St, c, tmp2);
      for (int k = 0; k < 3; ++k) Sc[k] = 0.5 * (tmp[k] + tmp2[k]);
      s2 += dot3(c, Sc);
      cross3(h, Sc, tmp);
      for (int k = 0; k < 3; ++k) ds_dt[k] += 2.0 * tmp[k];
      const double th = dot3(t, h), sh = dot3(Sc, h);
      for (int k = 0; k < 3; ++k) ds_dw[k] += 2.0 * (th * Sc[k] - sh * t[k]);
    }
  }
  const double s = sqrt(s2);
  const double r = e / s;
  const double k2 = e / (2.0 * s2 * s);
  double dr_dt[3], dr_dw[3];
  for (int k = 0; k < 3; ++k) {
    dr_dt[k] = de_dt[k] / s - k2 * ds_dt[k];
    dr_dw[k] = de_dw[k] / s - k2 * ds_dw[k];
  }
  *r_out = r;
  row[0] = dot3(dr_dt, t_th);
  row[1] = dot3(dr_dt, t_ph);
  /* delta is a half-angle left perturbation: R <- Exp(2 delta) R */
  row[2] = 2.0 * dr_dw[0];
  row[3] = 2.0 * dr_dw[1];
  row[4] = 2.0 * dr_dw[2];
}

/* residuals + local (tangent) Jacobian N x 5 (row-major) + cost. */
static double evaluate_full(const problem_view *p, const double x[6], int jac_mode,
                            double *residuals, double *J) {
  double cost = 0.0;
  double PJ[4][3];
  quat_plus_jacobian(x + 2, PJ);
  for (int64_t i = 0; i < p->n; ++i) {
    double r, row[5];
    if (jac_mode == JAC_ANALYTIC) {
      analytic_row(p, i, x, &r, row);
      r = residual_at(p, i, x); /* residual always from the literal functor */
    } else {
      double amb[6];
      r = residual_at(p, i, x);
      numeric_row(p, i, x, amb);
      row[0] = amb[0];
      row[1] = amb[1];
      /* ResidualBlock::Evaluate: local = global(1x4) * PlusJacobian(4x3) */
      for (int c = 0; c < 3; ++c)
        row[2 + c] = amb[2] * PJ[0][c] + amb[3] * PJ[1][c] + amb[4] * PJ[2][c] + amb[5] * PJ[3][c];
    }
    residuals[i] = r;
    memcpy(J + 5 * i, row, sizeof(row));
    cost += 0.5 * r * r;
  }
  return cost;
}

/* One evaluation: cost, gradient J^T r (5), upper triangle of J^T J (15). */
int oracle_eval(int variant, int jac_mode, int64_t n, const double *f1, const double *f2,
                const double *cov_t, const double *cov_h, double reg, const double pose7[7],
                double *out_cost, double *out_grad, double *out_jtj) {
  problem_view p = {variant, n, f1, f2, cov_t, cov_h, reg};
  double x[6];
  oracle_angles_from_vec(pose7 + 4, &x[0], &x[1]);
  memcpy(x + 2, pose7, 4 * sizeof(double));
  double *res = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  double *J = (double *)malloc(sizeof(double) * 5 * (size_t)(n > 0 ? n : 1));
  if (!res || !J) return -1;
  const double cost = evaluate_full(&p, x, jac_mode, res, J);
  double g[5] = {0}, H[15] = {0};
  for (int64_t i = 0; i < n; ++i) {
    const double *row = J + 5 * i;
    int k = 0;
    for (int a = 0; a < 5; ++a) {
      g[a] += row[a] * res[i];
      for (int b = a; b < 5; ++b) H[k++] += row[a] * row[b];
    }
  }
  if (out_cost) *out_cost = cost;
  if (out_grad) memcpy(out_grad, g, sizeof(g));
  if (out_jtj) memcpy(out_jtj, H, sizeof(H));
  free(res);
  free(J);
  return 0;
}

/* ------------------------------------------------------------- LM (Ceres) */

/* Cholesky solve of the 5x5 SPD system A y = b (the normal equations
 * J^T J + D^2, what SPARSE_NORMAL_CHOLESKY — the Ceres default when a sparse
 * backend is built in, as in the reference's Dockerfile — factorises).
 * Returns 0 on success, -1 if A is not positive definite / not finite. */
static int chol_solve5(const double A_in[5][5], const double b[5], double y[5]) {
  double L[5][5];
  memset(L, 0, sizeof(L));
  for (int j = 0; j < 5; ++j) {
    double d = A_in[j][j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (!(d > 0.0) || !isfinite(d)) return -1;
    L[j][j] = sqrt(d);
    for (int i = j + 1; i < 5; ++i) {
      double s = A_in[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = s / L[j][j];
    }
  }
  double z[5];
  for (int i = 0; i < 5; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i][k] * z[k];
    z[i] = s / L[i][i];
  }
  for (int i = 4; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < 5; ++k) s -= L[k][i] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = 0; i < 5; ++i)
    if (!isfinite(y[i])) return -1;
  return 0;
}

static double norm_n(const double *v, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s);
}

typedef struct oracle_info {
  int32_t status;
  int32_t iterations;
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  double initial_cost;
  double final_cost;
  double final_radius;
  double final_gradient_max_norm;
} oracle_info;

/* ceres::TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy,
 * monotonic steps, no inner iterations, no bounds, Jacobi scaling.
 * Parameter blocks in AddResidualBlock order: theta | phi | quaternion. */
int oracle_solve(const oracle_opts *o, int64_t n, const double *f1, const double *f2,
                 const double *cov_t, const double *cov_h, const double init_pose7[7],
                 double out_pose7[7], oracle_info *info) {
  problem_view p = {o->variant, n, f1, f2, cov_t, cov_h, o->regularization};
  oracle_info inf;
  memset(&inf, 0, sizeof(inf));

  /* PNECCeres::InitValues(orientation, translation), pnec_ceres.cc:188-192 */
  double x[6];
  oracle_angles_from_vec(init_pose7 + 4, &x[0], &x[1]);
  memcpy(x + 2, init_pose7, 4 * sizeof(double));

  int status = ST_MAX_ITERATIONS;
  if (n <= 0) {
    status = ST_EMPTY;
    goto done;
  }
  {
    double *residuals = (double *)malloc(sizeof(double) * (size_t)n);
    double *J = (double *)malloc(sizeof(double) * 5 * (size_t)n);
    double *model_res = (double *)malloc(sizeof(double) * (size_t)n);
    if (!residuals || !J || !model_res) return -1;

    double scale[5] = {1, 1, 1, 1, 1};
    double grad[5], diagonal[5] = {0, 0, 0, 0, 0};
    double radius = o->initial_trust_region_radius;
    double decrease_factor = 2.0;
    int reuse_diagonal = 0;
    int num_invalid = 0;
    int iteration = 0;
    int step_is_successful = 1; /* IterationZero marks itself successful */
    double x_norm = norm_n(x, 6);
    double gradient_max_norm = 0.0;

    /* IterationZero -> EvaluateGradientAndJacobian */
    double x_cost = evaluate_full(&p, x, o->jacobian_mode, residuals, J);
    inf.initial_cost = x_cost;
    if (!isfinite(x_cost)) {
      status = ST_NONFINITE;
      free(residuals); free(J); free(model_res);
      goto done;
    }
#define GRADIENT_AND_SCALE()                                                        \
  do {                                                                              \
    for (int a_ = 0; a_ < 5; ++a_) grad[a_] = 0.0;                                  \
    for (int64_t i_ = 0; i_ < n; ++i_)                                              \
      for (int a_ = 0; a_ < 5; ++a_) grad[a_] += J[5 * i_ + a_] * residuals[i_];    \
    if (o->jacobi_scaling) {                                                        \
      if (iteration == 0) {                                                         \
        double sq_[5] = {0, 0, 0, 0, 0};                                            \
        for (int64_t i_ = 0; i_ < n; ++i_)                                          \
          for (int a_ = 0; a_ < 5; ++a_) sq_[a_] += J[5 * i_ + a_] * J[5 * i_ + a_]; \
        for (int a_ = 0; a_ < 5; ++a_) scale[a_] = 1.0 / (1.0 + sqrt(sq_[a_]));     \
      }                                                                             \
      for (int64_t i_ = 0; i_ < n; ++i_)                                            \
        for (int a_ = 0; a_ < 5; ++a_) J[5 * i_ + a_] *= scale[a_];                 \
    }                                                                               \
    /* gradient_max_norm = || x - Plus(x, -g) ||_inf in the ambient space */        \
    double ng_[5], xs_[6];                                                          \
    for (int a_ = 0; a_ < 5; ++a_) ng_[a_] = -grad[a_];                             \
    state_plus(x, ng_, xs_);                                                        \
    gradient_max_norm = 0.0;                                                        \
    for (int a_ = 0; a_ < 6; ++a_) {                                                \
      const double d_ = fabs(x[a_] - xs_[a_]);                                      \
      if (d_ > gradient_max_norm) gradient_max_norm = d_;                           \
    }                                                                               \
  } while (0)
    GRADIENT_AND_SCALE();

    for (;;) {
      /* FinalizeIterationAndCheckIfMinimizerCanContinue */
      if (step_is_successful) inf.num_successful_steps++;
      else inf.num_unsuccessful_steps++;
      if (iteration >= o->max_num_iterations) { status = ST_MAX_ITERATIONS; break; }
      if (step_is_successful && gradient_max_norm <= o->gradient_tolerance) {
        status = ST_CONVERGED_GRADIENT; break;
      }
      if (radius <= o->min_trust_region_radius) { status = ST_CONVERGED_RADIUS; break; }
      iteration++;

      /* ComputeTrustRegionStep -> LevenbergMarquardtStrategy::ComputeStep */
      if (!reuse_diagonal) {
        for (int a = 0; a < 5; ++a) {
          double sq = 0.0;
          for (int64_t i = 0; i < n; ++i) sq += J[5 * i + a] * J[5 * i + a];
          diagonal[a] = fmin(fmax(sq, o->min_lm_diagonal), o->max_lm_diagonal);
        }
      }
      double A[5][5], rhs[5], y[5], step[5];
      memset(A, 0, sizeof(A));
      memset(rhs, 0, sizeof(rhs));
      for (int64_t i = 0; i < n; ++i) {
        const double *row = J + 5 * i;
        for (int a = 0; a < 5; ++a) {
          rhs[a] += row[a] * residuals[i];
          for (int b = a; b < 5; ++b) A[a][b] += row[a] * row[b];
        }
      }
      for (int a = 0; a < 5; ++a) {
        for (int b = 0; b < a; ++b) A[a][b] = A[b][a];
        /* lm_diagonal = sqrt(diagonal / radius) appended as rows: adds D^2 */
        const double lm = sqrt(diagonal[a] / radius);
        A[a][a] += lm * lm;
      }
      const int solve_ok = (chol_solve5(A, rhs, y) == 0);
      reuse_diagonal = 1;
      int step_is_valid = 0;
      double model_cost_change = 0.0;
      if (solve_ok) {
        for (int a = 0; a < 5; ++a) step[a] = -y[a];
        /* model_cost_change = -(J s)^T (f + J s / 2) */
        for (int64_t i = 0; i < n; ++i) {
          double m = 0.0;
          for (int a = 0; a < 5; ++a) m += J[5 * i + a] * step[a];
          model_res[i] = m;
        }
        for (int64_t i = 0; i < n; ++i)
          model_cost_change += model_res[i] * (residuals[i] + model_res[i] / 2.0);
        model_cost_change = -model_cost_change;
        step_is_valid = (model_cost_change > 0.0);
      }
      if (!step_is_valid) {
        /* HandleInvalidStep */
        if (++num_invalid >= o->max_num_consecutive_invalid_steps) {
          status = ST_FAILURE;
          break;
        }
        radius *= 0.5; /* StepIsInvalid */
        reuse_diagonal = 1;
        step_is_successful = 0;
        continue;
      }
      num_invalid = 0;
      double delta[5], cand[6];
      for (int a = 0; a < 5; ++a) delta[a] = step[a] * scale[a];

      /* ComputeCandidatePointAndEvaluateCost */
      state_plus(x, delta, cand);
      double cand_cost = evaluate_cost(&p, cand, NULL);
      if (!isfinite(cand_cost)) cand_cost = DBL_MAX;

      /* ParameterToleranceReached (ambient step norm) — checked BEFORE the
       * step is accepted, so the candidate is not applied on convergence */
      double diff[6];
      for (int a = 0; a < 6; ++a) diff[a] = x[a] - cand[a];
      const double step_norm = norm_n(diff, 6);
      if (step_norm <= o->parameter_tolerance * (x_norm + o->parameter_tolerance)) {
        status = ST_CONVERGED_PARAMETER;
        break;
      }
      /* FunctionToleranceReached */
      const double cost_change = x_cost - cand_cost;
      if (fabs(cost_change) <= o->function_tolerance * x_cost) {
        status = ST_CONVERGED_FUNCTION;
        break;
      }
      /* IsStepSuccessful: TrustRegionStepEvaluator::StepQuality (monotonic) */
      const double rho =
          (cand_cost >= DBL_MAX) ? -DBL_MAX : cost_change / model_cost_change;
      if (rho > o->min_relative_decrease) {
        /* HandleSuccessfulStep */
        memcpy(x, cand, sizeof(cand));
        x_norm = norm_n(x, 6);
        x_cost = evaluate_full(&p, x, o->jacobian_mode, residuals, J);
        GRADIENT_AND_SCALE();
        step_is_successful = 1;
        /* LevenbergMarquardtStrategy::StepAccepted */
        radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3));
        radius = fmin(o->max_trust_region_radius, radius);
        decrease_factor = 2.0;
        reuse_diagonal = 0;
      } else {
        step_is_successful = 0;
        /* LevenbergMarquardtStrategy::StepRejected */
        radius = radius / decrease_factor;
        decrease_factor *= 2.0;
        reuse_diagonal = 1;
      }
    }
#undef GRADIENT_AND_SCALE
    inf.iterations = iteration;
    inf.final_cost = x_cost;
    inf.final_radius = radius;
    inf.final_gradient_max_norm = gradient_max_norm;
    free(residuals);
    free(J);
    free(model_res);
  }
done:
  inf.status = status;
  /* PNECCeres::Result(): q.normalized(), t(theta, phi) */
  {
    const double qn = norm_n(x + 2, 4);
    for (int k = 0; k < 4; ++k) out_pose7[k] = qn > 0.0 ? x[2 + k] / qn : x[2 + k];
    translation_from_angles(x[0], x[1], out_pose7 + 4);
  }
  if (info) *info = inf;
  return 0;
}

/* Batched driver: problems distributed over `num_threads` host threads the way
 * run_simulation.sh:58-65 fans processes out.  offsets == NULL => uniform N. */
int oracle_solve_batch(const oracle_opts *o, int64_t num_problems, int64_t n_per_problem,
                       const int64_t *offsets, const double *f1, const double *f2,
                       const double *cov_t, const double *cov_h, const double *init_poses,
                       double *out_poses, oracle_info *infos, int num_threads) {
  int rc = 0;
#ifdef _OPENMP
  if (num_threads < 1) num_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(num_threads)
#endif
  for (int64_t b = 0; b < num_problems; ++b) {
    const int64_t s = offsets ? offsets[b] : b * n_per_problem;
    const int64_t e = offsets ? offsets[b + 1] : (b + 1) * n_per_problem;
    oracle_info inf;
    const int r = oracle_solve(o, e - s, f1 + 3 * s, f2 + 3 * s, cov_t ? cov_t + 9 * s : NULL,
                               cov_h ? cov_h + 9 * s : NULL, init_poses + 7 * b,
                               out_poses + 7 * b, &inf);
    if (infos) infos[b] = inf;
    if (r != 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
      rc = r;
    }
  }
  return rc;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ----------------------------------------------------------------- metrics */

static void pose_rot(const double pose7[7], double R[3][3]) {
  double q[4];
  const double qn = norm_n(pose7, 4);
  for (int k = 0; k < 4; ++k) q[k] = pose7[k] / qn;
  quat_to_rot(q, R);
}

/* pnec::common::RotationalDifference, src/common/common.cc:210-214:
 * |log(R1^T R2)| in DEGREES. */
double oracle_rotational_difference_deg(const double pose_a[7], const double pose_b[7]) {
  /* relative quaternion qa^-1 * qb, angle = 2 atan2(|v|, |w|) (Sophus logAndTheta) */
  double a[4], b[4];
  const double na = norm_n(pose_a, 4), nb = norm_n(pose_b, 4);
  for (int k = 0; k < 4; ++k) { a[k] = pose_a[k] / na; b[k] = pose_b[k] / nb; }
  const double ax = -a[0], ay = -a[1], az = -a[2], aw = a[3];
  const double w = aw * b[3] - ax * b[0] - ay * b[1] - az * b[2];
  const double x = aw * b[0] + ax * b[3] + ay * b[2] - az * b[1];
  const double y = aw * b[1] - ax * b[2] + ay * b[3] + az * b[0];
  const double z = aw * b[2] + ax * b[1] - ay * b[0] + az * b[3];
  const double vn = sqrt(x * x + y * y + z * z);
  const double theta = 2.0 * atan2(vn, fabs(w));
  return fabs(theta) * 180.0 / M_PI;
}

/* pnec::common::TranslationalDifference, src/common/common.cc:216-235, DEGREES. */
double oracle_translational_difference_deg(const double t1[3], const double t2[3],
                                           int both_directions) {
  const double n1 = norm_n(t1, 3), n2 = norm_n(t2, 3);
  double error;
  if (n1 < 1e-10 || n1 < 1e-10) { /* sic: the reference tests translation_1 twice */
    error = M_PI / 2;
  } else {
    const double c = dot3(t1, t2) / (n1 * n2);
    if (both_directions) {
      const double e1 = acos(c), e2 = acos(-c);
      error = e1 < e2 ? e1 : e2;
    } else {
      error = acos(c);
    }
  }
  return error * 180.0 / M_PI;
}

/* pnec::common::CostFunction, src/common/common.cc:237-259: mean PNEC energy
 * WITHOUT regularisation. */
double oracle_cost_function(int64_t n, const double *f1, const double *f2, const double *cov,
                            const double pose7[7]) {
  double R[3][3], cost = 0.0;
  pose_rot(pose7, R);
  const double *t = pose7 + 4;
  for (int64_t i = 0; i < n; ++i) {
    double F[3][3], S[3][3], v[3], tt[3], Rf2[3], nrm[3], Stt[3];
    skew(f1 + 3 * i, F);
    load_cov(cov + 9 * i, S);
    vecmat(t, F, v);
    vecmat(v, R, tt);
    matvec(R, f2 + 3 * i, Rf2);
    cross3(f1 + 3 * i, Rf2, nrm);
    const double num = dot3(t, nrm);
    matvec(S, tt, Stt);
    cost += num * num / dot3(tt, Stt);
  }
  return cost / (double)n;
}

/* Sum of squared residuals (energy WITH regularisation), for pinning the
 * functors against scripts/pnec/common.py:13-58. */
double oracle_energy(int variant, int64_t n, const double *f1, const double *f2,
                     const double *cov_t, const double *cov_h, double reg, const double q[4],
                     const double t[3]) {
  /* the python energies take t directly, so bypass the (theta, phi) chart */
  double theta, phi;
  oracle_angles_from_vec(t, &theta, &phi);
  double e = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double r = functor(variant, theta, phi, q, f1 + 3 * i, f2 + 3 * i,
                             cov_t ? cov_t + 9 * i : NULL, cov_h ? cov_h + 9 * i : NULL, reg);
    e += r * r;
  }
  return e;
}

/* ------------------------------------------------------- unscented transform */

/* pnec::common::RotationBetweenPoints, src/common/common.cc:118-124 */
static void rotation_between_points(const double p1[3], const double p2[3], double R[3][3]) {
  double v[3], V[3][3], V2[3][3];
  cross3(p1, p2, v);
  const double c = dot3(p1, p2);
  skew(v, V);
  matmul(V, V, V2);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = (i == j ? 1.0 : 0.0) + V[i][j] + V2[i][j] / (1 + c);
}

/* pnec::common::UnscentedTransform, src/common/common.cc:467-525.  camera_model:
 * 0 = Omnidirectional, 1 = Pinhole (enum order of include/common/common.h:61).
 * mu[3], cov[9] and K_inv[9] column-major; out[9] column-major. */
void oracle_unscented_transform(const double mu[3], const double *cov, const double *K_inv,
                                double kappa, int camera_model, double *out) {
  const int n = 2, m = 2 * n + 1;
  double S[3][3], K[3][3], C[3][3];
  load_cov(cov, S);
  load_cov(K_inv, K);
  memset(C, 0, sizeof(C));
  if (camera_model == 0) {
    const double z[3] = {0.0, 0.0, 1.0};
    const double nm = norm_n(mu, 3);
    const double mn[3] = {mu[0] / nm, mu[1] / nm, mu[2] / nm};
    double R[3][3], Rt[3][3], T1[3][3], T2[3][3], L[3][3];
    rotation_between_points(z, mn, R);
    transpose(R, Rt);
    matmul(Rt, S, T1);
    matmul(T1, R, T2);
    memset(L, 0, sizeof(L));
    L[0][0] = sqrt(T2[0][0]);
    L[1][0] = T2[1][0] / L[0][0];
    L[1][1] = sqrt(T2[1][1] - L[1][0] * L[1][0]);
    matmul(R, L, C);
  } else {
    C[0][0] = sqrt(S[0][0]);
    C[1][0] = S[1][0] / C[0][0];
    C[1][1] = sqrt(S[1][1] - C[1][0] * C[1][0]);
  }
  double points[5][3], weights[5], tp[5][3], mean[3] = {0, 0, 0};
  memcpy(points[0], mu, 3 * sizeof(double));
  weights[0] = kappa / ((float)n + kappa);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) points[1 + i][k] = mu[k] + C[k][i];
    weights[1 + i] = 0.5 / ((float)n + kappa);
  }
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) points[1 + n + i][k] = mu[k] - C[k][i];
    weights[1 + n + i] = 0.5 / ((float)n + kappa);
  }
  for (int i = 0; i < m; ++i) {
    double q[3];
    if (camera_model == 0) memcpy(q, points[i], sizeof(q));
    else matvec(K, points[i], q);
    const double nq = norm_n(q, 3);
    for (int k = 0; k < 3; ++k) {
      tp[i][k] = q[k] / nq;
      mean[k] = mean[k] + weights[i] * tp[i][k];
    }
  }
  double sigma[3][3];
  memset(sigma, 0, sizeof(sigma));
  for (int i = 0; i < m; ++i)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        sigma[r][c] = sigma[r][c] + weights[i] * (tp[i][r] - mean[r]) * (tp[i][c] - mean[c]);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) out[c * 3 + r] = sigma[r][c];
}

void oracle_unscented_transform_batch(int64_t n, const double *mus, const double *covs,
                                      const double *K_inv, double kappa, int camera_model,
                                      double *out) {
  for (int64_t i = 0; i < n; ++i)
    oracle_unscented_transform(mus + 3 * i, covs + 9 * i, K_inv, kappa, camera_model, out + 9 * i);
}

/* KeyPoint::Unproject, src/frames/keypoints.cc:49-62, for n keypoints.
 * points[n][2], covs2[n][4] column-major 2x2, K_inv[9] column-major. */
void oracle_keypoints_unproject_batch(int64_t n, const double *points, const double *covs2,
                                      const double *K_inv, double *out_bvs, double *out_covs) {
  double K[3][3];
  load_cov(K_inv, K);
  for (int64_t i = 0; i < n; ++i) {
    const double mu[3] = {points[2 * i], points[2 * i + 1], 1.0};
    double v[3];
    matvec(K, mu, v); /* pnec::common::Unproject, common.cc:460-465 */
    const double nv = norm_n(v, 3);
    for (int k = 0; k < 3; ++k) out_bvs[3 * i + k] = v[k] / nv;
    double cov[9] = {covs2[4 * i], covs2[4 * i + 1], 0.0, covs2[4 * i + 2], covs2[4 * i + 3], 0.0, 0.0, 0.0, 0.0};
    oracle_unscented_transform(mu, cov, K_inv, 1.0, 1, out_covs + 9 * i);
  }
}

/* ---------------------------------------------- translation given rotation */

/* Smallest-eigenvalue eigenvector of a symmetric 3x3 by cyclic Jacobi rotations.  The
 * reference uses Eigen::SelfAdjointEigenSolver (scf.cc:135) / Eigen::EigenSolver
 * (common.cc:158); any converged symmetric eigen-solver gives the same eigenvector up to
 * rounding and SIGN (Eigen's sign is an implementation detail: compare modulo sign). */
static void sym3_smallest_eigvec(const double M[3][3], double v[3], double *lambda) {
  double a[3][3], q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  memcpy(a, M, sizeof(a));
  for (int sweep = 0; sweep < 50; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int r = p + 1; r < 3; ++r) {
        if (a[p][r] == 0.0) continue;
        const double theta = (a[r][r] - a[p][p]) / (2.0 * a[p][r]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akr = a[k][r];
          a[k][p] = c * akp - sn * akr;
          a[k][r] = sn * akp + c * akr;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], ark = a[r][k];
          a[p][k] = c * apk - sn * ark;
          a[r][k] = sn * apk + c * ark;
        }
        for (int k = 0; k < 3; ++k) {
          const double qkp = q[k][p], qkr = q[k][r];
          q[k][p] = c * qkp - sn * qkr;
          q[k][r] = sn * qkp + c * qkr;
        }
      }
  }
  int j = 0;
  if (a[1][1] < a[j][j]) j = 1;
  if (a[2][2] < a[j][j]) j = 2;
  *lambda = a[j][j];
  const double nv = sqrt(q[0][j] * q[0][j] + q[1][j] * q[1][j] + q[2][j] * q[2][j]);
  for (int k = 0; k < 3; ++k) v[k] = q[k][j] / nv;
}

/* pnec::optimization::fibonacci_sphere, src/optimization/scf.cc:53-72 (float casts included). */
void oracle_fibonacci_sphere(int samples, double *points) {
  const double phi = M_PI * (3.0 - sqrt(5.0));
  for (int i = 0; i < samples; ++i) {
    const double y = 1.0 - ((float)i / (float)(samples - 1)) * 2.0;
    const double radius = sqrt(1 - y * y);
    const double theta = phi * (float)i;
    points[3 * i] = cos(theta) * radius;
    points[3 * i + 1] = y;
    points[3 * i + 2] = sin(theta) * radius;
  }
}

/* pnec::optimization::obj_fun, scf.cc:43-51 */
static double scf_obj_fun(const double t[3], int64_t n, const double (*A)[3][3], const double (*B)[3][3]) {
  double cost = 0;
  for (int64_t i = 0; i < n; ++i) {
    double At[3], Bt[3];
    matvec(A[i], t, At);
    matvec(B[i], t, Bt);
    cost += dot3(t, At) / dot3(t, Bt);
  }
  return cost;
}

double oracle_scf_objective(int64_t n, const double *A, const double *B, const double t[3]) {
  return scf_obj_fun(t, n, (const double(*)[3][3])A, (const double(*)[3][3])B);
}

/* The SCF stage of PNEC::WeightedEigensolver for one frame pair:
 * A_i / B_i (pnec.cc:317-328), Fibonacci scan (pnec.cc:330-340), scf (scf.cc:128-147 with
 * alt_construct_E, scf.cc:109-126, whose `frac` vector is all zeros for i < n). */
int oracle_scf_translation(int64_t n, const double *f1, const double *f2, const double *cov,
                           const double pose7[7], double reg, int samples, int steps,
                           double out_t[3], double *out_cost) {
  double R[3][3], Rt[3][3];
  pose_rot(pose7, R);
  transpose(R, Rt);
  double(*A)[3][3] = malloc(sizeof(double[3][3]) * ((size_t)n + 1));
  double(*Bm)[3][3] = malloc(sizeof(double[3][3]) * ((size_t)n + 1));
  double *pts = malloc(sizeof(double) * 3 * (size_t)(samples > 0 ? samples : 1));
  if (!A || !Bm || !pts) return -1;
  for (int64_t i = 0; i < n; ++i) {
    double F[3][3], Ft[3][3], S[3][3], g[3], v[3], T1[3][3], T2[3][3];
    skew(f1 + 3 * i, F);
    transpose(F, Ft);
    load_cov(cov + 9 * i, S);
    matvec(R, f2 + 3 * i, g);
    matvec(F, g, v); /* bv1_skew * rotation * bvs2[i] */
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) A[i][r][c] = v[r] * v[c];
    matmul(F, R, T1);
    matmul(T1, S, T2);
    matmul(T2, Rt, T1);
    matmul(T1, Ft, Bm[i]);
    for (int r = 0; r < 3; ++r) Bm[i][r][r] += reg;
  }
  oracle_fibonacci_sphere(samples, pts);
  double best[3] = {pose7[4], pose7[5], pose7[6]};
  double best_cost = scf_obj_fun(best, n, A, Bm);
  for (int k = 0; k < samples; ++k) {
    const double c = scf_obj_fun(pts + 3 * k, n, A, Bm);
    if (c < best_cost) {
      best_cost = c;
      memcpy(best, pts + 3 * k, sizeof(best));
    }
  }
  double t[3] = {best[0], best[1], best[2]};
  for (int it = 0; it < steps; ++it) {
    double E[3][3];
    memset(E, 0, sizeof(E));
    for (int64_t i = 0; i < n; ++i) {
      double Bt[3];
      matvec(Bm[i], t, Bt);
      const double phi_B = dot3(t, Bt);
      const double frac = 0.0; /* sic: frac.resize(n) followed by push_back */
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) E[r][c] += (1.0 / phi_B) * (A[i][r][c] - frac * Bm[i][r][c]);
    }
    double lam;
    sym3_smallest_eigvec(E, t, &lam);
  }
  memcpy(out_t, t, sizeof(t));
  if (out_cost) *out_cost = scf_obj_fun(t, n, A, Bm);
  free(A);
  free(Bm);
  free(pts);
  return 0;
}

/* TranslationFromM(ComposeM(bvs_1, bvs_2, R)), common.cc:127-181.  out_M: xx xy xz yy yz zz. */
void oracle_nec_translation(int64_t n, const double *f1, const double *f2, const double pose7[7],
                            double out_t[3], double *out_M) {
  double R[3][3], M[3][3];
  pose_rot(pose7, R);
  memset(M, 0, sizeof(M));
  for (int64_t i = 1; i < n; ++i) { /* sic: ComposeM starts at i = 1 */
    double g[3], nrm[3];
    matvec(R, f2 + 3 * i, g);
    cross3(f1 + 3 * i, g, nrm);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) M[r][c] += nrm[r] * nrm[c];
  }
  double lam;
  sym3_smallest_eigvec(M, out_t, &lam);
  if (out_M) {
    out_M[0] = M[0][0]; out_M[1] = M[0][1]; out_M[2] = M[0][2];
    out_M[3] = M[1][1]; out_M[4] = M[1][2]; out_M[5] = M[2][2];
  }
}

/* front stages of PNEC::Solve (NEC eigensolver, weighted eigensolver, orchestration) */
#include "pnec_oracle_frame.c"
