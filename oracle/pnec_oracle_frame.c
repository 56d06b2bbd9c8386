/*
 * pnec_oracle_frame.c — CPU restatement of the stages of PNEC::Solve in FRONT of the Ceres
 * refinement, textually included by pnec_oracle.c.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as pnec_oracle.c).
 *
 * What is restated, with the reference call sites it follows:
 *   oracle_eigensolver          opengv::relative_pose::eigensolver(adapter), called at
 *                               src/rel_pose_estimation/pnec.cc:274 and :313
 *   oracle_weighted_eigensolver PNEC::WeightedEigensolver, pnec.cc:283-348
 *   oracle_frame_solve          PNEC::Solve (use_ransac_ == false), pnec.cc:77-124
 *
 * PARITY STATUS
 *   opengv is a third-party dependency that is NOT in the reference tree (it arrives inside
 *   basalt through `git clone --recursive`, unpinned: Dockerfile:21, README.md:62-65) and is
 *   not installed here.  Its eigensolver is restated from its published algorithm (Kneip &
 *   Lynen, "Direct optimization of frame-to-frame rotation", ICCV 2013) and from opengv's
 *   public sources as recalled (src/relative_pose/methods.cpp `eigensolver`,
 *   src/relative_pose/modules/main.cpp `eigensolver_main`,
 *   src/relative_pose/modules/eigensolver/modules.cpp `getSmallestEVwithJacobian`,
 *   include/opengv/OptimizationFunctor.hpp + Eigen's unsupported NonLinearOptimization
 *   `LevenbergMarquardt<NumericalDiff<Eigensolver_step>>`):
 *     1. the six 3x3 moment matrices  xxF = sum_i w_i f1x f1x f2 f2^T, yyF, zzF, xyF, yzF, zxF;
 *     2. M(c) = sum_i n_i n_i^T, n_i = f1_i x R'(c) f2_i with R'(c) the Cayley rotation WITHOUT
 *        its 1/(1+|c|^2) factor (`cayley2rot_reduced`), assembled from the moments;
 *     3. lambda_min(M) by the closed-form trigonometric root of the characteristic cubic and
 *        its analytic derivative with respect to c;
 *     4. a MINPACK-style Levenberg-Marquardt (Eigen's port of lmdif: forward differences,
 *        epsfcn = 0, factor = 100, ftol = 5e-5, xtol = 10 eps, gtol = 0, maxfev = 100) on the
 *        3 residuals "d lambda_min / d c = 0", started at rot2cayley(R_init);
 *     5. the result rotation is cayley2rot(c) (normalised).
 *   PINNED: step 4 against the Fortran MINPACK that scipy.optimize.leastsq wraps, on this
 *   file's own residual function (tests/test_oracle_golden.py); step 2/3 against numpy's
 *   eigvalsh and finite differences.  UNPINNED: that steps 1-5 are what opengv executes —
 *   no opengv build or golden vector is available in this image or in the reference tree.
 */

/* ------------------------------------------------------------------ moments */

/* G[u][v] = sum_i w_i f1_u f1_v f2 f2^T  (opengv's xxF = G[0][0], xyF = G[0][1], zxF = G[2][0] ...) */
typedef struct es_moments {
  double G[3][3][3][3];
} es_moments;

static void es_accumulate(es_moments *m, int64_t n, const double *f1, const double *f2,
                          const double *weights) {
  memset(m, 0, sizeof(*m));
  for (int64_t i = 0; i < n; ++i) {
    const double *a = f1 + 3 * i, *b = f2 + 3 * i;
    const double w = weights ? weights[i] : 1.0;
    double F[3][3];
    /* F = f2 f2^T; with weights the adapter holds f2 * sqrt(w) (pnec.cc:302-306) */
    double bw[3] = {b[0], b[1], b[2]};
    if (weights) {
      const double sw = sqrt(w);
      for (int k = 0; k < 3; ++k) bw[k] = b[k] * sw;
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) F[r][c] = bw[r] * bw[c];
    for (int u = 0; u < 3; ++u)
      for (int v = 0; v < 3; ++v) {
        const double s = a[u] * a[v];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) m->G[u][v][r][c] += s * F[r][c];
      }
  }
}

/* opengv::math::cayley2rot_reduced and its three derivatives */
static void cayley_reduced(const double c[3], double R[3][3], double dR[3][3][3]) {
  const double x = c[0], y = c[1], z = c[2];
  R[0][0] = 1 + x * x - y * y - z * z; R[0][1] = 2 * (x * y - z);           R[0][2] = 2 * (x * z + y);
  R[1][0] = 2 * (x * y + z);           R[1][1] = 1 - x * x + y * y - z * z; R[1][2] = 2 * (y * z - x);
  R[2][0] = 2 * (x * z - y);           R[2][1] = 2 * (y * z + x);           R[2][2] = 1 - x * x - y * y + z * z;
  if (!dR) return;
  const double dx[3][3] = {{2 * x, 2 * y, 2 * z}, {2 * y, -2 * x, -2}, {2 * z, 2, -2 * x}};
  const double dy[3][3] = {{-2 * y, 2 * x, 2}, {2 * x, 2 * y, 2 * z}, {-2, 2 * z, -2 * y}};
  const double dz[3][3] = {{-2 * z, -2, 2 * x}, {2, -2 * z, 2 * y}, {2 * x, 2 * y, 2 * z}};
  memcpy(dR[0], dx, sizeof(dx));
  memcpy(dR[1], dy, sizeof(dy));
  memcpy(dR[2], dz, sizeof(dz));
}

static double bilinear(const double a[3], const double G[3][3], const double b[3]) {
  double s = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) s += a[r] * G[r][c] * b[c];
  return s;
}

/* n_a = sum over the two (u, p, sign) terms of f1_u * (R f2)_p  (cross product f1 x R f2) */
static const int kCrossU[3][2] = {{1, 2}, {2, 0}, {0, 1}};
static const int kCrossP[3][2] = {{2, 1}, {0, 2}, {1, 0}};
static const double kCrossS[2] = {1.0, -1.0};

/* opengv eigensolver::composeMwithJacobians: M = sum n n^T from the moments, and dM/dc_k */
static void es_compose_m(const es_moments *m, const double c[3], double M[3][3], double dM[3][3][3]) {
  double R[3][3], dR[3][3][3];
  cayley_reduced(c, R, dM ? dR : NULL);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double acc = 0, dacc[3] = {0, 0, 0};
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
          const int u = kCrossU[a][i], p = kCrossP[a][i], v = kCrossU[b][j], s = kCrossP[b][j];
          const double sg = kCrossS[i] * kCrossS[j];
          acc += sg * bilinear(R[p], m->G[u][v], R[s]);
          if (dM)
            for (int k = 0; k < 3; ++k)
              dacc[k] += sg * (bilinear(dR[k][p], m->G[u][v], R[s]) + bilinear(R[p], m->G[u][v], dR[k][s]));
        }
      M[a][b] = acc;
      if (dM)
        for (int k = 0; k < 3; ++k) dM[k][a][b] = dacc[k];
    }
}

/* opengv eigensolver::getSmallestEVwithJacobian: closed-form smallest root of the
 * characteristic cubic of M and its derivative by the chain rule. */
static double es_smallest_ev(const es_moments *m, const double cay[3], double jac[3]) {
  double M[3][3], dM[3][3][3];
  es_compose_m(m, cay, M, dM);
  const double b = -M[0][0] - M[1][1] - M[2][2];
  const double c = -pow(M[0][2], 2) - pow(M[1][2], 2) - pow(M[0][1], 2) + M[0][0] * M[1][1] +
                   M[0][0] * M[2][2] + M[1][1] * M[2][2];
  const double d = M[1][1] * pow(M[0][2], 2) + M[0][0] * pow(M[1][2], 2) + M[2][2] * pow(M[0][1], 2) -
                   M[0][0] * M[1][1] * M[2][2] - 2 * M[0][1] * M[1][2] * M[0][2];
  const double s = 2 * pow(b, 3) - 9 * b * c + 27 * d;
  const double t = 4 * pow(pow(b, 2) - 3 * c, 3);
  const double alpha = acos(s / sqrt(t));
  const double beta = alpha / 3;
  const double y = cos(beta);
  const double r = 0.5 * sqrt(t);
  const double w = pow(r, 1.0 / 3.0);
  const double k = w * y;
  const double ev = (-b - 2 * k) / 3;
  if (jac)
    for (int q = 0; q < 3; ++q) {
      const double(*J)[3] = dM[q];
      const double bj = -J[0][0] - J[1][1] - J[2][2];
      const double cj = -2.0 * M[0][2] * J[0][2] - 2.0 * M[1][2] * J[1][2] - 2.0 * M[0][1] * J[0][1] +
                        J[0][0] * M[1][1] + M[0][0] * J[1][1] + J[0][0] * M[2][2] + M[0][0] * J[2][2] +
                        J[1][1] * M[2][2] + M[1][1] * J[2][2];
      const double dj = J[1][1] * pow(M[0][2], 2) + M[1][1] * 2 * M[0][2] * J[0][2] +
                        J[0][0] * pow(M[1][2], 2) + M[0][0] * 2 * M[1][2] * J[1][2] +
                        J[2][2] * pow(M[0][1], 2) + M[2][2] * 2 * M[0][1] * J[0][1] -
                        J[0][0] * M[1][1] * M[2][2] - M[0][0] * J[1][1] * M[2][2] - M[0][0] * M[1][1] * J[2][2] -
                        2 * (J[0][1] * M[1][2] * M[0][2] + M[0][1] * J[1][2] * M[0][2] + M[0][1] * M[1][2] * J[0][2]);
      const double sj = 2 * 3 * pow(b, 2) * bj - 9 * bj * c - 9 * b * cj + 27 * dj;
      const double tj = 4 * 3 * pow(pow(b, 2) - 3 * c, 2) * (2 * b * bj - 3 * cj);
      const double alphaj = -1 / sqrt(1 - (pow(s, 2) / t)) * (sj * sqrt(t) - s * 0.5 * pow(t, -0.5) * tj) / t;
      const double betaj = alphaj / 3;
      const double yj = -sin(beta) * betaj;
      const double rj = 0.25 * pow(t, -0.5) * tj;
      const double wj = (1.0 / 3.0) * pow(r, -2.0 / 3.0) * rj;
      const double kj = wj * y + w * yj;
      jac[q] = (-bj - 2 * kj) / 3;
    }
  return ev;
}

/* exported for the tests: lambda_min and gradient at a Cayley vector, from raw correspondences */
double oracle_es_smallest_ev(int64_t n, const double *f1, const double *f2, const double *weights,
                             const double cayley[3], double jac[3], double M_out[9]) {
  es_moments m;
  es_accumulate(&m, n, f1, f2, weights);
  if (M_out) {
    double M[3][3];
    es_compose_m(&m, cayley, M, NULL);
    memcpy(M_out, M, sizeof(M));
  }
  return es_smallest_ev(&m, cayley, jac);
}

/* ---------------------------------------------- MINPACK lmdif, n = m = 3 */

enum { LMN = 3 };

typedef void (*lm_fcn)(const void *ctx, const double x[LMN], double fvec[LMN]);

typedef struct lm_params {
  double ftol, xtol, gtol, factor, epsfcn;
  int maxfev;
  int fev_per_jacobian; /* Eigen's NumericalDiff (Forward) re-evaluates f(x): n + 1; MINPACK fdjac2: n */
} lm_params;

typedef struct lm_result {
  int info; /* MINPACK info code 0..8 */
  int nfev;
  int iterations; /* successful iterations + 1, MINPACK's `iter` */
  double fnorm;
} lm_result;

static double enorm3(const double *v, int n) {
  /* MINPACK enorm guards against over/underflow only; the values here are O(1..1e6) */
  double s = 0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s);
}

/* qrfac with column pivoting; a is m x n, row-major a[i][j]; on exit the strict upper triangle
 * holds R, rdiag its diagonal, the lower trapezoid the Householder vectors. */
static void lm_qrfac(double a[LMN][LMN], int ipvt[LMN], double rdiag[LMN], double acnorm[LMN]) {
  double wa[LMN];
  for (int j = 0; j < LMN; ++j) {
    double col[LMN];
    for (int i = 0; i < LMN; ++i) col[i] = a[i][j];
    acnorm[j] = enorm3(col, LMN);
    rdiag[j] = acnorm[j];
    wa[j] = rdiag[j];
    ipvt[j] = j;
  }
  for (int j = 0; j < LMN; ++j) {
    int kmax = j;
    for (int k = j; k < LMN; ++k)
      if (rdiag[k] > rdiag[kmax]) kmax = k;
    if (kmax != j) {
      for (int i = 0; i < LMN; ++i) {
        const double t = a[i][j];
        a[i][j] = a[i][kmax];
        a[i][kmax] = t;
      }
      rdiag[kmax] = rdiag[j];
      wa[kmax] = wa[j];
      const int k = ipvt[j];
      ipvt[j] = ipvt[kmax];
      ipvt[kmax] = k;
    }
    double col[LMN];
    for (int i = j; i < LMN; ++i) col[i - j] = a[i][j];
    double ajnorm = enorm3(col, LMN - j);
    if (ajnorm != 0.0) {
      if (a[j][j] < 0.0) ajnorm = -ajnorm;
      for (int i = j; i < LMN; ++i) a[i][j] /= ajnorm;
      a[j][j] += 1.0;
      for (int k = j + 1; k < LMN; ++k) {
        double sum = 0;
        for (int i = j; i < LMN; ++i) sum += a[i][j] * a[i][k];
        const double temp = sum / a[j][j];
        for (int i = j; i < LMN; ++i) a[i][k] -= temp * a[i][j];
        if (rdiag[k] != 0.0) {
          const double t = a[j][k] / rdiag[k];
          const double u = 1.0 - t * t;
          rdiag[k] *= sqrt(u > 0.0 ? u : 0.0);
          const double q = rdiag[k] / wa[k];
          if (0.05 * q * q <= DBL_EPSILON) {
            double rest[LMN];
            for (int i = j + 1; i < LMN; ++i) rest[i - j - 1] = a[i][k];
            rdiag[k] = enorm3(rest, LMN - j - 1);
            wa[k] = rdiag[k];
          }
        }
      }
    }
    rdiag[j] = -ajnorm;
  }
}

static void lm_qrsolv(double r[LMN][LMN], const int ipvt[LMN], const double diag[LMN],
                      const double qtb[LMN], double x[LMN], double sdiag[LMN]) {
  double wa[LMN];
  for (int j = 0; j < LMN; ++j) {
    for (int i = j; i < LMN; ++i) r[i][j] = r[j][i];
    x[j] = r[j][j];
    wa[j] = qtb[j];
  }
  for (int j = 0; j < LMN; ++j) {
    const int l = ipvt[j];
    if (diag[l] != 0.0) {
      for (int k = j; k < LMN; ++k) sdiag[k] = 0.0;
      sdiag[j] = diag[l];
      double qtbpj = 0.0;
      for (int k = j; k < LMN; ++k) {
        if (sdiag[k] == 0.0) continue;
        double sn, cs;
        if (fabs(r[k][k]) < fabs(sdiag[k])) {
          const double cotan = r[k][k] / sdiag[k];
          sn = 0.5 / sqrt(0.25 + 0.25 * cotan * cotan);
          cs = sn * cotan;
        } else {
          const double tn = sdiag[k] / r[k][k];
          cs = 0.5 / sqrt(0.25 + 0.25 * tn * tn);
          sn = cs * tn;
        }
        r[k][k] = cs * r[k][k] + sn * sdiag[k];
        const double temp = cs * wa[k] + sn * qtbpj;
        qtbpj = -sn * wa[k] + cs * qtbpj;
        wa[k] = temp;
        for (int i = k + 1; i < LMN; ++i) {
          const double t2 = cs * r[i][k] + sn * sdiag[i];
          sdiag[i] = -sn * r[i][k] + cs * sdiag[i];
          r[i][k] = t2;
        }
      }
    }
    sdiag[j] = r[j][j];
    r[j][j] = x[j];
  }
  int nsing = LMN;
  for (int j = 0; j < LMN; ++j) {
    if (sdiag[j] == 0.0 && nsing == LMN) nsing = j;
    if (nsing < LMN) wa[j] = 0.0;
  }
  for (int k = 1; k <= nsing; ++k) {
    const int j = nsing - k;
    double sum = 0;
    for (int i = j + 1; i < nsing; ++i) sum += r[i][j] * wa[i];
    wa[j] = (wa[j] - sum) / sdiag[j];
  }
  for (int j = 0; j < LMN; ++j) x[ipvt[j]] = wa[j];
}

static void lm_lmpar(double r[LMN][LMN], const int ipvt[LMN], const double diag[LMN],
                     const double qtb[LMN], double delta, double *par, double x[LMN]) {
  const double dwarf = DBL_MIN;
  double wa1[LMN], wa2[LMN], sdiag[LMN];
  int nsing = LMN;
  for (int j = 0; j < LMN; ++j) {
    wa1[j] = qtb[j];
    if (r[j][j] == 0.0 && nsing == LMN) nsing = j;
    if (nsing < LMN) wa1[j] = 0.0;
  }
  for (int k = 1; k <= nsing; ++k) {
    const int j = nsing - k;
    wa1[j] /= r[j][j];
    const double temp = wa1[j];
    for (int i = 0; i < j; ++i) wa1[i] -= r[i][j] * temp;
  }
  for (int j = 0; j < LMN; ++j) x[ipvt[j]] = wa1[j];

  int iter = 0;
  for (int j = 0; j < LMN; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = enorm3(wa2, LMN);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) {
    *par = 0.0;
    return;
  }
  double parl = 0.0;
  if (nsing >= LMN) {
    for (int j = 0; j < LMN; ++j) {
      const int l = ipvt[j];
      wa1[j] = diag[l] * (wa2[l] / dxnorm);
    }
    for (int j = 0; j < LMN; ++j) {
      double sum = 0;
      for (int i = 0; i < j; ++i) sum += r[i][j] * wa1[i];
      wa1[j] = (wa1[j] - sum) / r[j][j];
    }
    const double temp = enorm3(wa1, LMN);
    parl = ((fp / delta) / temp) / temp;
  }
  for (int j = 0; j < LMN; ++j) {
    double sum = 0;
    for (int i = 0; i <= j; ++i) sum += r[i][j] * qtb[i];
    wa1[j] = sum / diag[ipvt[j]];
  }
  const double gnorm = enorm3(wa1, LMN);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = dwarf / (delta < 0.1 ? delta : 0.1);
  if (*par < parl) *par = parl;
  if (*par > paru) *par = paru;
  if (*par == 0.0) *par = gnorm / dxnorm;
  for (;;) {
    ++iter;
    if (*par == 0.0) *par = (dwarf > 0.001 * paru) ? dwarf : 0.001 * paru;
    double temp = sqrt(*par);
    for (int j = 0; j < LMN; ++j) wa1[j] = temp * diag[j];
    lm_qrsolv(r, ipvt, wa1, qtb, x, sdiag);
    for (int j = 0; j < LMN; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = enorm3(wa2, LMN);
    temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    for (int j = 0; j < LMN; ++j) {
      const int l = ipvt[j];
      wa1[j] = diag[l] * (wa2[l] / dxnorm);
    }
    for (int j = 0; j < LMN; ++j) {
      wa1[j] /= sdiag[j];
      const double t2 = wa1[j];
      for (int i = j + 1; i < LMN; ++i) wa1[i] -= r[i][j] * t2;
    }
    temp = enorm3(wa1, LMN);
    const double parc = ((fp / delta) / temp) / temp;
    if (fp > 0.0 && *par > parl) parl = *par;
    if (fp < 0.0 && *par < paru) paru = *par;
    *par = (parl > *par + parc) ? parl : *par + parc;
  }
}

/* lmdif as Eigen's LevenbergMarquardt<NumericalDiff<F>>::minimize runs it (minimizeInit +
 * minimizeOneStep loop) */
static void lm_minimize(lm_fcn fcn, const void *ctx, const lm_params *p, double x[LMN], lm_result *res) {
  double fvec[LMN], fjac[LMN][LMN], diag[LMN], qtf[LMN], wa1[LMN], wa2[LMN], wa3[LMN], wa4[LMN];
  int ipvt[LMN];
  const double epsmch = DBL_EPSILON;
  int info = 0, nfev = 1, iter = 1;
  double par = 0.0, delta = 0.0, xnorm = 0.0;
  fcn(ctx, x, fvec);
  double fnorm = enorm3(fvec, LMN);
  const double eps = sqrt(p->epsfcn > epsmch ? p->epsfcn : epsmch);
  for (;;) {
    /* forward-difference Jacobian (fdjac2 / NumericalDiff Forward) */
    for (int j = 0; j < LMN; ++j) {
      const double temp = x[j];
      double h = eps * fabs(temp);
      if (h == 0.0) h = eps;
      x[j] = temp + h;
      fcn(ctx, x, wa4);
      x[j] = temp;
      for (int i = 0; i < LMN; ++i) fjac[i][j] = (wa4[i] - fvec[i]) / h;
    }
    nfev += p->fev_per_jacobian;
    lm_qrfac(fjac, ipvt, wa1, wa2);
    if (iter == 1) {
      for (int j = 0; j < LMN; ++j) diag[j] = (wa2[j] == 0.0) ? 1.0 : wa2[j];
      for (int j = 0; j < LMN; ++j) wa3[j] = diag[j] * x[j];
      xnorm = enorm3(wa3, LMN);
      delta = p->factor * xnorm;
      if (delta == 0.0) delta = p->factor;
    }
    for (int i = 0; i < LMN; ++i) wa4[i] = fvec[i];
    for (int j = 0; j < LMN; ++j) {
      if (fjac[j][j] != 0.0) {
        double sum = 0;
        for (int i = j; i < LMN; ++i) sum += fjac[i][j] * wa4[i];
        const double temp = -sum / fjac[j][j];
        for (int i = j; i < LMN; ++i) wa4[i] += fjac[i][j] * temp;
      }
      fjac[j][j] = wa1[j];
      qtf[j] = wa4[j];
    }
    double gnorm = 0.0;
    if (fnorm != 0.0)
      for (int j = 0; j < LMN; ++j) {
        const int l = ipvt[j];
        if (wa2[l] != 0.0) {
          double sum = 0;
          for (int i = 0; i <= j; ++i) sum += fjac[i][j] * (qtf[i] / fnorm);
          const double g = fabs(sum / wa2[l]);
          if (g > gnorm) gnorm = g;
        }
      }
    if (gnorm <= p->gtol) {
      info = 4;
      break;
    }
    for (int j = 0; j < LMN; ++j)
      if (wa2[j] > diag[j]) diag[j] = wa2[j];
    double ratio;
    do {
      lm_lmpar(fjac, ipvt, diag, qtf, delta, &par, wa1);
      for (int j = 0; j < LMN; ++j) {
        wa1[j] = -wa1[j];
        wa2[j] = x[j] + wa1[j];
        wa3[j] = diag[j] * wa1[j];
      }
      const double pnorm = enorm3(wa3, LMN);
      if (iter == 1 && pnorm < delta) delta = pnorm;
      fcn(ctx, wa2, wa4);
      ++nfev;
      const double fnorm1 = enorm3(wa4, LMN);
      double actred = -1.0;
      if (0.1 * fnorm1 < fnorm) actred = 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
      for (int j = 0; j < LMN; ++j) {
        wa3[j] = 0.0;
        const double temp = wa1[ipvt[j]];
        for (int i = 0; i <= j; ++i) wa3[i] += fjac[i][j] * temp;
      }
      const double t1 = enorm3(wa3, LMN) / fnorm, t2 = sqrt(par) * pnorm / fnorm;
      const double temp1 = t1 * t1, temp2 = t2 * t2;
      const double prered = temp1 + temp2 / 0.5;
      const double dirder = -(temp1 + temp2);
      ratio = 0.0;
      if (prered != 0.0) ratio = actred / prered;
      if (ratio <= 0.25) {
        double temp = 0.5;
        if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
        delta = temp * (delta < pnorm / 0.1 ? delta : pnorm / 0.1);
        par /= temp;
      } else if (!(par != 0.0 && ratio < 0.75)) {
        delta = pnorm / 0.5;
        par = 0.5 * par;
      }
      if (ratio >= 1e-4) {
        for (int j = 0; j < LMN; ++j) {
          x[j] = wa2[j];
          wa2[j] = diag[j] * x[j];
        }
        for (int i = 0; i < LMN; ++i) fvec[i] = wa4[i];
        xnorm = enorm3(wa2, LMN);
        fnorm = fnorm1;
        ++iter;
      }
      const int small_red = fabs(actred) <= p->ftol && prered <= p->ftol && 0.5 * ratio <= 1.0;
      if (small_red) info = 1;
      if (delta <= p->xtol * xnorm) info = 2;
      if (small_red && info == 2) info = 3;
      if (info != 0) break;
      if (nfev >= p->maxfev) info = 5;
      if (fabs(actred) <= epsmch && prered <= epsmch && 0.5 * ratio <= 1.0) info = 6;
      if (delta <= epsmch * xnorm) info = 7;
      if (gnorm <= epsmch) info = 8;
      if (info != 0) break;
    } while (ratio < 1e-4);
    if (info != 0) break;
  }
  res->info = info;
  res->nfev = nfev;
  res->iterations = iter;
  res->fnorm = fnorm;
}

/* Eigensolver_step (opengv OptimizationFunctor): the residuals are the three components of
 * d lambda_min / d cayley */
static void es_step_fcn(const void *ctx, const double x[3], double fvec[3]) {
  es_smallest_ev((const es_moments *)ctx, x, fvec);
}

typedef struct oracle_es_info {
  int32_t lm_info, nfev, iterations, reserved;
  double smallest_ev; /* lambda_min of the reduced M at the result */
  double cayley[3];
} oracle_es_info;

static void es_default_params(lm_params *p) {
  /* opengv eigensolver_main: lm.resetParameters(); ftol = 0.00005; xtol = 1.E1 * eps; maxfev = 100.
   * resetParameters(): factor = 100, gtol = 0, epsfcn = 0. */
  p->ftol = 0.00005;
  p->xtol = 1.0e1 * DBL_EPSILON;
  p->gtol = 0.0;
  p->factor = 100.0;
  p->epsfcn = 0.0;
  p->maxfev = 100;
  p->fev_per_jacobian = LMN + 1;
}

/* opengv::math::rot2cayley of the rotation of a pose: [c]x = (R - I)(R + I)^-1, i.e. q_xyz / q_w */
static void rot2cayley_pose(const double pose7[7], double c[3]) {
  for (int k = 0; k < 3; ++k) c[k] = pose7[k] / pose7[3];
}

static void cayley2quat(const double c[3], double q[4]) {
  const double s = 1.0 / sqrt(1.0 + c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  q[0] = c[0] * s; q[1] = c[1] * s; q[2] = c[2] * s; q[3] = s;
}

static void es_run(const es_moments *m, const double init_pose7[7], double out_quat[4], oracle_es_info *info) {
  lm_params p;
  es_default_params(&p);
  double x[3];
  rot2cayley_pose(init_pose7, x);
  lm_result res;
  lm_minimize(es_step_fcn, m, &p, x, &res);
  cayley2quat(x, out_quat);
  if (info) {
    info->lm_info = res.info;
    info->nfev = res.nfev;
    info->iterations = res.iterations;
    info->reserved = 0;
    info->smallest_ev = es_smallest_ev(m, x, NULL);
    memcpy(info->cayley, x, sizeof(x));
  }
}

/* rotation_t opengv::relative_pose::eigensolver(adapter [bvs1, bvs2(*sqrt w), R12 = init rotation]).
 * out_quat: unit quaternion (x, y, z, w) of cayley2rot(result). */
int oracle_eigensolver(int64_t n, const double *f1, const double *f2, const double *weights,
                       const double init_pose7[7], double out_quat[4], oracle_es_info *info) {
  es_moments m;
  es_accumulate(&m, n, f1, f2, weights);
  es_run(&m, init_pose7, out_quat, info);
  return 0;
}

/* the LM alone on a caller-supplied functor configuration: used to pin lm_minimize against
 * scipy's MINPACK (tests).  fev_per_jacobian = 3 reproduces MINPACK's own nfev accounting. */
int oracle_es_lm(int64_t n, const double *f1, const double *f2, const double *weights,
                 const double x0[3], double ftol, double xtol, double gtol, double factor, int maxfev,
                 int fev_per_jacobian, double x_out[3], int32_t *info_out, int32_t *nfev_out) {
  es_moments m;
  es_accumulate(&m, n, f1, f2, weights);
  lm_params p = {ftol, xtol, gtol, factor, 0.0, maxfev, fev_per_jacobian};
  lm_result res;
  memcpy(x_out, x0, 3 * sizeof(double));
  lm_minimize(es_step_fcn, &m, &p, x_out, &res);
  if (info_out) *info_out = res.info;
  if (nfev_out) *nfev_out = res.nfev;
  return 0;
}

/* ------------------------------------------------------ PNEC::Eigensolver */

/* PNEC::Eigensolver, use_ransac_ == false (pnec.cc:273-279): opengv rotation, then
 * TranslationFromM(ComposeM(bvs1, bvs2, rotation)). */
int oracle_nec_eigensolver_pose(int64_t n, const double *f1, const double *f2, const double init_pose7[7],
                                double out_pose7[7], oracle_es_info *info) {
  oracle_eigensolver(n, f1, f2, NULL, init_pose7, out_pose7, info);
  out_pose7[4] = out_pose7[5] = out_pose7[6] = 0.0;
  oracle_nec_translation(n, f1, f2, out_pose7, out_pose7 + 4, NULL);
  return 0;
}

/* pnec::common::Weight(..., host_frame = false) (common.cc:183-208) times the 1e-8 of pnec.cc:300 */
void oracle_weights(int64_t n, const double *f1, const double *cov, const double pose7[7], double reg,
                    double *weights) {
  double R[3][3];
  pose_rot(pose7, R);
  for (int64_t i = 0; i < n; ++i) {
    double F[3][3], S[3][3], tF[3], tt[3], St[3];
    skew(f1 + 3 * i, F);
    load_cov(cov + 9 * i, S);
    vecmat(pose7 + 4, F, tF);  /* t^T [f1]x */
    vecmat(tF, R, tt);         /* t^T [f1]x R */
    matvec(S, tt, St);
    weights[i] = 1.0 / (dot3(tt, St) + reg) * 1.0e-8;
  }
}

/* PNEC::WeightedEigensolver, pnec.cc:283-348.  `initial_pose` is the pose the WEIGHTS are
 * computed from in every iteration (pnec.cc:296-300) and the first iteration's start. */
int oracle_weighted_eigensolver(int64_t n, const double *f1, const double *f2, const double *cov,
                                const double initial_pose7[7], double reg, int weighted_iterations,
                                int fibonacci_samples, int scf_steps, double out_pose7[7]) {
  double rel[7];
  memcpy(rel, initial_pose7, sizeof(rel));
  /* Sophus::SE3d stores a unit quaternion */
  {
    const double qn = norm_n(rel, 4);
    for (int k = 0; k < 4; ++k) rel[k] /= qn;
  }
  double *weights = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (!weights) return -1;
  for (int it = 0; it + 1 < weighted_iterations; ++it) {
    oracle_weights(n, f1, cov, initial_pose7, reg, weights);
    double q[4];
    oracle_eigensolver(n, f1, f2, weights, rel, q, NULL);
    double pose[7] = {q[0], q[1], q[2], q[3], rel[4], rel[5], rel[6]};
    double t[3];
    oracle_scf_translation(n, f1, f2, cov, pose, reg, fibonacci_samples, scf_steps, t, NULL);
    memcpy(rel, q, sizeof(q));
    memcpy(rel + 4, t, sizeof(t));
  }
  free(weights);
  memcpy(out_pose7, rel, sizeof(rel));
  return 0;
}

/* ------------------------------------------------------------------ RANSAC
 *
 * PNEC::Eigensolver with use_ransac_ == true (src/rel_pose_estimation/pnec.cc:239-272), i.e.
 *   opengv::sac::Ransac<opengv::sac_problems::relative_pose::EigensolverSacProblem>
 * with threshold_ = 1e-6, max_iterations_ = Options::max_ransac_iterations_ (5000), sample size
 * Options::ransac_sample_size_ (10), probability_ 0.99 (opengv's default), followed by
 * optimizeModelCoefficients on the inliers and TranslationFromM(ComposeM(inlier bvs, rotation)).
 *
 * PARITY UNPINNED, and only ever statistically pinnable: opengv (not in the reference tree) draws
 * its samples from a time-seeded std::mt19937 (SampleConsensusProblem(randomSeed = true)) and its
 * start perturbations from rand().  Restated from opengv's public sources as recalled
 * (sac/implementation/Ransac.hpp `computeModel`, sac/implementation/SampleConsensusProblem.hpp
 * `getSamples` / `drawIndexSample`, sac_problems/relative_pose/EigensolverSacProblem.cpp,
 * relative_pose/methods.cpp `eigensolver(adapter, indices, output)`, triangulation/methods.cpp
 * `triangulate2`):
 *   computeModel: k = 1, best = -INT_MAX; while (iterations < k):
 *     getSamples -> drawIndexSample: for i < sample_size: swap(shuffled[i], shuffled[i + rnd() % (n - i)]);
 *       the sample is shuffled[0 .. sample_size).  `shuffled` starts as 0..n-1 and PERSISTS across
 *       iterations.  Fewer correspondences than the sample size: no model (empty inlier set).
 *     computeModelCoefficients: start = rot2cayley(adapter.getR12()) + U(-1, 1) * maxVariation per
 *       component (maxVariation = 0.1 in the source as recalled, next to a comment reading 0.01; the
 *       value is an option here), then eigensolver(adapter, sample, model): the six moment sums over the
 *       sample, eigensolver_main (Levenberg-Marquardt as above), rotation = cayley2rot(x),
 *       translation = eigenvector of the smallest eigenvalue of M(x) (its magnitude does not enter the
 *       score), sign: along the optical flow f1 - R f2 of the sample's FIRST correspondence.
 *     countWithinDistance -> getSelectedDistancesToModel: adapter.sett12 / setR12(model) -- which is
 *       why the NEXT iteration's getR12() returns THIS model's rotation: the start rotation random-walks
 *       along the chain of hypotheses --, p = triangulate2 (midpoint), score = (1 - f1 . p/|p|) +
 *       (1 - f2 . p'/|p'|), p' = R^T (p - t); inlier iff score < threshold (strict).
 *     count > best: keep the model; k = log(1 - probability) / log(1 - w^sample_size), w = count / n,
 *       with 1 - w^s clamped to [eps, 1 - eps].
 *     ++iterations; iterations > max_iterations: stop.
 *   inliers = selectWithinDistance(best model).
 * pnec.cc:253-272 then runs optimizeModelCoefficients (eigensolver over the inliers started at the best
 * model's rotation, no perturbation) and takes the translation from ComposeM over the inlier arrays
 * (which skips their first element, common.cc:131).
 *
 * TWO MODES.  `sequential = 1` keeps opengv's sequential state as described (persistent shuffle, start
 * rotation chained through the adapter).  `sequential = 0` (default; what the CUDA path implements)
 * makes every hypothesis a function of (seed, pair, iteration) alone -- a fresh partial Fisher-Yates
 * from 0..n-1 and a start at the INITIAL rotation --, which is what allows hypotheses to be evaluated
 * in parallel and then replayed through the bookkeeping above in order: same distribution of samples,
 * starts scattered around the initial rotation instead of around the previous hypothesis.  The two
 * modes are compared statistically in tests/test_oracle_frame.py.  The random stream is counter based
 * (splitmix64 of seed, pair, iteration, draw) in both; the draws keep opengv's forms
 * (`r % (n - i)` with r in [0, 2^31), `r / RAND_MAX`).
 *
 * Where the reference has undefined behaviour (fewer correspondences than the sample size, or a best
 * model without inliers: optimizeModelCoefficients then indexes an empty vector) this restatement
 * returns the start pose (unit quaternion) and an empty inlier set.
 */

static uint64_t rs_mix(uint64_t z) { /* splitmix64 finaliser */
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
/* 31 random bits for (seed, pair, iteration, draw): the range of rand() / of opengv's rnd() */
static uint32_t rs_u31(uint64_t seed, uint64_t pair, uint64_t iteration, uint64_t draw) {
  const uint64_t h = rs_mix(rs_mix(rs_mix(seed ^ 0x51ed270b7a2f3c15ULL) + pair) + (iteration << 8) + draw);
  return (uint32_t)(h >> 33);
}

typedef struct oracle_ransac_opts {
  int32_t max_iterations; /* 5000  Options::max_ransac_iterations_ */
  int32_t sample_size;    /* 10    Options::ransac_sample_size_    */
  int32_t sequential;     /* 0     1: opengv's sequential state (see above) */
  int32_t reserved;
  double threshold;       /* 1e-6  pnec.cc:250 */
  double probability;     /* 0.99  opengv default */
  double max_variation;   /* 0.1   EigensolverSacProblem::computeModelCoefficients */
  uint64_t seed;
} oracle_ransac_opts;

void oracle_ransac_opts_default(oracle_ransac_opts *o) {
  o->max_iterations = 5000;
  o->sample_size = 10;
  o->sequential = 0;
  o->reserved = 0;
  o->threshold = 1.0e-6;
  o->probability = 0.99;
  o->max_variation = 0.1;
  o->seed = 1;
}

/* opengv::triangulation::triangulate2 (midpoint) + the bearing-vector reprojection score of
 * EigensolverSacProblem::getSelectedDistancesToModel */
static double ransac_score(const double R[3][3], const double t[3], const double f1[3], const double f2[3]) {
  double g[3];
  matvec(R, f2, g); /* f2 in frame 1 */
  const double b0 = dot3(t, f1), b1 = dot3(t, g);
  const double a00 = dot3(f1, f1), a10 = dot3(f1, g), a01 = -a10, a11 = -dot3(g, g);
  const double det = a00 * a11 - a01 * a10;
  const double l0 = (a11 * b0 - a01 * b1) / det, l1 = (-a10 * b0 + a00 * b1) / det;
  double p[3], pp[3], d[3];
  for (int k = 0; k < 3; ++k) p[k] = 0.5 * (l0 * f1[k] + t[k] + l1 * g[k]);
  for (int k = 0; k < 3; ++k) d[k] = p[k] - t[k];
  for (int k = 0; k < 3; ++k) pp[k] = R[0][k] * d[0] + R[1][k] * d[1] + R[2][k] * d[2]; /* R^T (p - t) */
  const double np = norm_n(p, 3), npp = norm_n(pp, 3);
  return (1.0 - dot3(f1, p) / np) + (1.0 - dot3(f2, pp) / npp);
}
double oracle_ransac_score(const double pose7[7], const double f1[3], const double f2[3]) {
  double R[3][3];
  pose_rot(pose7, R);
  return ransac_score(R, pose7 + 4, f1, f2);
}

/* eigensolver on a subset: Cayley parameters, rotation (quaternion) + translation direction with
 * opengv's sign rule.  x: in = start, out = result of the minimisation. */
static void ransac_model(int64_t m, const int64_t *idx, const double *f1, const double *f2, double x[3],
                         double R[3][3], double t[3], double quat[4]) {
  es_moments mom;
  memset(&mom, 0, sizeof(mom));
  for (int64_t s = 0; s < m; ++s) {
    const double *a = f1 + 3 * idx[s], *b = f2 + 3 * idx[s];
    for (int u = 0; u < 3; ++u)
      for (int v = 0; v < 3; ++v)
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) mom.G[u][v][r][c] += a[u] * a[v] * b[r] * b[c];
  }
  lm_params p;
  es_default_params(&p);
  lm_result res;
  lm_minimize(es_step_fcn, &mom, &p, x, &res);
  cayley2quat(x, quat);
  quat_to_rot(quat, R);
  double M[3][3], lam;
  es_compose_m(&mom, x, M, NULL);
  sym3_smallest_eigvec(M, t, &lam);
  /* sign: along the optical flow f1 - R f2 of the first correspondence of the subset */
  double g[3];
  matvec(R, f2 + 3 * idx[0], g);
  const double *a0 = f1 + 3 * idx[0];
  const double flow = (a0[0] - g[0]) * t[0] + (a0[1] - g[1]) * t[1] + (a0[2] - g[2]) * t[2];
  if (flow < 0.0)
    for (int k = 0; k < 3; ++k) t[k] = -t[k];
}

/* opengv::sac::Ransac<EigensolverSacProblem>::computeModel + selectWithinDistance for one frame pair.
 *   inlier_mask[n] (0/1) or NULL; best_model7: the winning hypothesis (unit quaternion + signed unit
 *   translation); returns 0, or -2 when n < sample_size (no model: zero inliers, best_model7 = start). */
int oracle_ransac_compute_model(const oracle_ransac_opts *o, int64_t pair_index, int64_t n, const double *f1,
                                const double *f2, const double init_pose7[7], double best_model7[7],
                                uint8_t *inlier_mask, int32_t *num_inliers, int32_t *iterations) {
  const int ns = o->sample_size;
  double q0[4];
  {
    const double qn = norm_n(init_pose7, 4);
    for (int k = 0; k < 4; ++k) q0[k] = init_pose7[k] / qn;
  }
  if (num_inliers) *num_inliers = 0;
  if (iterations) *iterations = 0;
  if (inlier_mask) memset(inlier_mask, 0, (size_t)(n > 0 ? n : 0));
  memcpy(best_model7, q0, sizeof(q0));
  memcpy(best_model7 + 4, init_pose7 + 4, 3 * sizeof(double));
  if (ns < 1 || n < ns) return -2;
  int64_t *perm = malloc(sizeof(int64_t) * (size_t)n), *sample = malloc(sizeof(int64_t) * (size_t)ns);
  if (!perm || !sample) return -1;
  double c0[3], ccur[3];
  rot2cayley_pose(init_pose7, c0);
  memcpy(ccur, c0, sizeof(c0));
  int best = -1, iters = 0;
  double bestR[3][3], bestt[3] = {0, 0, 0}, bestq[4] = {0, 0, 0, 1}, k = 1.0;
  memset(bestR, 0, sizeof(bestR));
  for (int64_t i = 0; i < n; ++i) perm[i] = i;
  while ((double)iters < k) {
    if (!o->sequential)
      for (int64_t i = 0; i < n; ++i) perm[i] = i; /* every hypothesis from the identity permutation */
    for (int s = 0; s < ns; ++s) { /* drawIndexSample */
      const int64_t j = s + (int64_t)(rs_u31(o->seed, (uint64_t)pair_index, (uint64_t)iters, (uint64_t)s) % (uint64_t)(n - s));
      const int64_t tmp = perm[s];
      perm[s] = perm[j];
      perm[j] = tmp;
    }
    for (int s = 0; s < ns; ++s) sample[s] = perm[s];
    double x[3], R[3][3], t[3], q[4];
    const double *cs = o->sequential ? ccur : c0;
    for (int d = 0; d < 3; ++d) {
      const double u = (double)rs_u31(o->seed, (uint64_t)pair_index, (uint64_t)iters, (uint64_t)(ns + d)) / 2147483647.0;
      x[d] = cs[d] + (u - 0.5) * 2.0 * o->max_variation;
    }
    ransac_model(ns, sample, f1, f2, x, R, t, q);
    memcpy(ccur, x, sizeof(x)); /* adapter.setR12(model.rotation) inside the scoring */
    int count = 0;
    for (int64_t i = 0; i < n; ++i)
      if (ransac_score(R, t, f1 + 3 * i, f2 + 3 * i) < o->threshold) ++count;
    if (count > best) {
      best = count;
      memcpy(bestR, R, sizeof(bestR));
      memcpy(bestt, t, sizeof(bestt));
      memcpy(bestq, q, sizeof(bestq));
      const double w = (double)count / (double)n;
      double p_no = 1.0 - pow(w, (double)ns);
      if (p_no < DBL_EPSILON) p_no = DBL_EPSILON;
      if (p_no > 1.0 - DBL_EPSILON) p_no = 1.0 - DBL_EPSILON;
      k = log(1.0 - o->probability) / log(p_no);
    }
    ++iters;
    if (iters > o->max_iterations) break;
  }
  int32_t ni = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int in = ransac_score(bestR, bestt, f1 + 3 * i, f2 + 3 * i) < o->threshold;
    if (inlier_mask) inlier_mask[i] = (uint8_t)in;
    ni += in;
  }
  if (num_inliers) *num_inliers = ni;
  if (iterations) *iterations = iters;
  memcpy(best_model7, bestq, sizeof(bestq));
  memcpy(best_model7 + 4, bestt, sizeof(bestt));
  free(perm);
  free(sample);
  return 0;
}

/* PNEC::Eigensolver with RANSAC for one frame pair (pnec.cc:239-272): computeModel, then
 * optimizeModelCoefficients = eigensolver over the inliers from the best model's rotation, then
 * TranslationFromM(ComposeM(in_bvs1, in_bvs2, rotation)) -- ComposeM skips the first inlier. */
int oracle_ransac_eigensolver(const oracle_ransac_opts *o, int64_t pair_index, int64_t n, const double *f1,
                              const double *f2, const double init_pose7[7], double out_pose7[7],
                              uint8_t *inlier_mask, int32_t *num_inliers, int32_t *iterations) {
  uint8_t *mask = inlier_mask ? inlier_mask : malloc((size_t)(n > 0 ? n : 1));
  if (!mask) return -1;
  double best[7];
  int32_t ni = 0;
  const int rc = oracle_ransac_compute_model(o, pair_index, n, f1, f2, init_pose7, best, mask, &ni, iterations);
  if (num_inliers) *num_inliers = ni;
  if (rc != 0 || ni == 0) {
    /* no model / no inlier: undefined in the reference; the start pose (unit quaternion) here */
    const double qn = norm_n(init_pose7, 4);
    for (int k = 0; k < 4; ++k) out_pose7[k] = init_pose7[k] / qn;
    memcpy(out_pose7 + 4, init_pose7 + 4, 3 * sizeof(double));
    if (!inlier_mask) free(mask);
    return rc == -1 ? -1 : 0;
  }
  double *g1 = malloc(sizeof(double) * 3 * (size_t)ni), *g2 = malloc(sizeof(double) * 3 * (size_t)ni);
  if (!g1 || !g2) return -1;
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i)
    if (mask[i]) {
      memcpy(g1 + 3 * m, f1 + 3 * i, 3 * sizeof(double));
      memcpy(g2 + 3 * m, f2 + 3 * i, 3 * sizeof(double));
      ++m;
    }
  /* the same two calls PNEC::Eigensolver makes without RANSAC, on the inlier arrays, started at the
   * best model's rotation */
  oracle_nec_eigensolver_pose(ni, g1, g2, best, out_pose7, NULL);
  free(g1);
  free(g2);
  if (!inlier_mask) free(mask);
  return 0;
}

/* PNEC::Solve, pnec.cc:77-124.  Mirrors pnec::rel_pose_estimation::Options (pnec_config.h:46-65):
 * use_nec, use_ceres, weighted_iterations, regularization, use_ransac (+ its settings). */
typedef struct oracle_frame_opts {
  int32_t use_nec, use_ceres, weighted_iterations, fibonacci_samples, scf_steps, use_ransac;
  oracle_opts ceres;
  oracle_ransac_opts ransac;
} oracle_frame_opts;

void oracle_frame_opts_default(oracle_frame_opts *o) {
  o->use_nec = 0;
  o->use_ceres = 1;
  o->weighted_iterations = 10;
  o->fibonacci_samples = 500;
  o->scf_steps = 10;
  o->use_ransac = 1; /* pnec_config.h:58 */
  oracle_opts_default(&o->ceres);
  oracle_ransac_opts_default(&o->ransac);
}

/* pair_index keys the random stream of the RANSAC stage.  inlier_mask[n] / num_inliers /
 * ransac_iterations may be NULL; without RANSAC the mask is all ones and num_inliers = 0 (the
 * reference clears `inliers`, pnec.cc:277). */
int oracle_frame_solve(const oracle_frame_opts *o, int64_t pair_index, int64_t n, const double *f1, const double *f2,
                       const double *cov, const double init_pose7[7], double out_pose7[7],
                       double es_pose7[7], uint8_t *inlier_mask, int32_t *num_inliers, int32_t *ransac_iterations) {
  double es[7];
  double *g1 = NULL, *g2 = NULL, *gc = NULL;
  int rc = 0;
  if (num_inliers) *num_inliers = 0;
  if (ransac_iterations) *ransac_iterations = 0;
  if (o->use_ransac) {
    uint8_t *mask = inlier_mask ? inlier_mask : malloc((size_t)(n > 0 ? n : 1));
    if (!mask) return -1;
    int32_t ni = 0;
    rc = oracle_ransac_eigensolver(&o->ransac, pair_index, n, f1, f2, init_pose7, es, mask, &ni, ransac_iterations);
    if (rc) return rc;
    if (num_inliers) *num_inliers = ni;
    /* InlierExtraction, pnec.cc:210-229 */
    g1 = malloc(sizeof(double) * 3 * (size_t)(ni > 0 ? ni : 1));
    g2 = malloc(sizeof(double) * 3 * (size_t)(ni > 0 ? ni : 1));
    gc = cov ? malloc(sizeof(double) * 9 * (size_t)(ni > 0 ? ni : 1)) : NULL;
    if (!g1 || !g2 || (cov && !gc)) return -1;
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i)
      if (mask[i]) {
        memcpy(g1 + 3 * m, f1 + 3 * i, 3 * sizeof(double));
        memcpy(g2 + 3 * m, f2 + 3 * i, 3 * sizeof(double));
        if (cov) memcpy(gc + 9 * m, cov + 9 * i, 9 * sizeof(double));
        ++m;
      }
    if (!inlier_mask) free(mask);
    f1 = g1; f2 = g2; cov = gc; n = ni;
  } else {
    oracle_nec_eigensolver_pose(n, f1, f2, init_pose7, es, NULL);
    if (inlier_mask) memset(inlier_mask, 1, (size_t)(n > 0 ? n : 0));
  }
  if (es_pose7) memcpy(es_pose7, es, sizeof(es));
  oracle_opts c = o->ceres;
  if (o->use_nec) {
    if (!o->use_ceres) {
      memcpy(out_pose7, es, sizeof(es));
    } else {
      c.variant = V_NEC;
      rc = oracle_solve(&c, n, f1, f2, NULL, NULL, es, out_pose7, NULL);
    }
    goto done;
  }
  double init[7];
  if (o->weighted_iterations > 1) {
    rc = oracle_weighted_eigensolver(n, f1, f2, cov, es, c.regularization, o->weighted_iterations,
                                     o->fibonacci_samples, o->scf_steps, init);
    if (rc) goto done;
  } else if (o->weighted_iterations == 1) {
    memcpy(init, es, sizeof(es));
  } else {
    memcpy(init, init_pose7, sizeof(init));
    const double qn = norm_n(init, 4);
    for (int k = 0; k < 4; ++k) init[k] /= qn;
  }
  if (!o->use_ceres) {
    memcpy(out_pose7, init, sizeof(init));
    goto done;
  }
  c.variant = V_TARGET; /* PNEC::CeresSolver -> Optimize(bvs1, bvs2, covs, reg) default noise frame */
  rc = oracle_solve(&c, n, f1, f2, cov, NULL, init, out_pose7, NULL);
done:
  free(g1);
  free(g2);
  free(gc);
  return rc;
}

/* inlier_mask [total] / num_inliers [B] / ransac_iterations [B] may be NULL; pair b's random stream
 * is keyed by pair_index_base + b. */
int oracle_frame_solve_batch(const oracle_frame_opts *o, int64_t num_problems, int64_t n_per_problem,
                             const int64_t *offsets, const double *f1, const double *f2, const double *cov,
                             const double *init_poses, double *out_poses, double *es_poses, int num_threads,
                             int64_t pair_index_base, uint8_t *inlier_mask, int32_t *num_inliers,
                             int32_t *ransac_iterations) {
  int rc = 0;
#ifdef _OPENMP
  if (num_threads < 1) num_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads)
#endif
  for (int64_t b = 0; b < num_problems; ++b) {
    const int64_t s = offsets ? offsets[b] : b * n_per_problem;
    const int64_t e = offsets ? offsets[b + 1] : (b + 1) * n_per_problem;
    const int r = oracle_frame_solve(o, pair_index_base + b, e - s, f1 + 3 * s, f2 + 3 * s, cov ? cov + 9 * s : NULL,
                                     init_poses + 7 * b, out_poses + 7 * b, es_poses ? es_poses + 7 * b : NULL,
                                     inlier_mask ? inlier_mask + s : NULL, num_inliers ? num_inliers + b : NULL,
                                     ransac_iterations ? ransac_iterations + b : NULL);
    if (r != 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
      rc = r;
    }
  }
  return rc;
}
