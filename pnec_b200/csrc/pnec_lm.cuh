// pnec_lm.cuh — the Levenberg-Marquardt update of one frame pair, Ceres semantics
// (ceres::TrustRegionMinimizer + LevenbergMarquardtStrategy as the reference runs them,
// src/optimization/pnec_ceres.cc:110 with default ceres::Solver::Options).
#pragma once

#include "pnec_device.cuh"

namespace pnec {

// ------------------------------------------------------------------------ LM
//
// State of ceres::TrustRegionMinimizer + LevenbergMarquardtStrategy for one
// problem.  Lives in shared memory between evaluations so that it costs no
// registers while the CTA streams correspondences.
//
// The per-iteration update is the serial part of a solve (one warp, the other
// warps of the CTA wait), so it is written as straight-line, branch-free code:
// Newton-refined MUFU reciprocals instead of IEEE division, LDL^T instead of
// Cholesky (no square roots), small-angle polynomials + angle addition instead of
// sin/cos calls.  All of it stays within a few ulp of the IEEE forms.
struct LMState {
  // Points and totals are double-buffered so that accepting a step flips an index
  // instead of copying: pts[xi] is the accepted point x, pts[xi ^ 1] the candidate;
  // tot[ti] holds (J^T J, J^T r, cost) at x, tot[ti ^ 1] receives the next evaluation.
  double pts[2][6];  // (theta, phi, qx, qy, qz, qw)
  double scs[2][4];  // sin/cos of theta, phi of the same points: s_th, c_th, s_ph, c_ph
  double tot[2][kAccPad];
  double scale[5];   // Jacobi column scaling, fixed at iteration 0
  double inv_scale2[5];  // 1 / scale^2
  double diag[5];    // LM diagonal of the scaled Jacobian (reused on rejection)
  double x_cost, inv_radius, decrease_factor, inv_model_cost_change, initial_cost;
  int xi, ti;
  int iteration, num_invalid, reuse_diagonal, step_successful, grad_converged, status, done;
  int pass_mode;     // what the CTA evaluates next: kPassFull or kPassCost
};

constexpr int kPassFull = 0;  // residual + Jacobian + JtJ/Jtr at the candidate
constexpr int kPassCost = 1;  // cost only (the step is predicted to hit function_tolerance)

__host__ __device__ constexpr int tri(int a, int b) {  // a <= b
  return a * 5 - (a * (a - 1)) / 2 + (b - a);
}

// 1/x: MUFU.RCP64H seed + Newton steps (each squares the relative error); no
// special-case branches.  x must be a finite, non-zero normal number.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);        // e + e^2
  y = fma(y, e, y);        // y (1 + e + e^2): relative error eps^3
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);        // eps^6
  return y;
}
// sqrt(x) for x >= 0 through rsqrt (no slow-path branch)
__device__ __forceinline__ double fast_sqrt(double x) { return x > 0.0 ? x * rsqrt(x) : 0.0; }

// sin(a)/a and cos(a) as polynomials in z = a^2, |a| <= 0.25 (truncation < 1e-19).
// Estrin evaluation: the dependent chain is 4 deep instead of 8 (fp64 latency is what the
// serial LM phase is made of).
__device__ __forceinline__ void small_sinc_cos(double z, double &sinc, double &cs) {
  const double z2 = z * z, z4 = z2 * z2;
  // sinc = sum_k (-1)^k z^k / (2k+1)!,  k = 0..7
  const double s01 = fma(z, -1.0 / 6.0, 1.0);
  const double s23 = fma(z, -1.0 / 5040.0, 1.0 / 120.0);
  const double s45 = fma(z, -1.0 / 39916800.0, 1.0 / 362880.0);
  const double s67 = fma(z, -1.0 / 1307674368000.0, 1.0 / 6227020800.0);
  sinc = fma(z4, fma(z2, s67, s45), fma(z2, s23, s01));
  // cos = sum_k (-1)^k z^k / (2k)!,  k = 0..8
  const double c01 = fma(z, -0.5, 1.0);
  const double c23 = fma(z, -1.0 / 720.0, 1.0 / 24.0);
  const double c45 = fma(z, -1.0 / 3628800.0, 1.0 / 40320.0);
  const double c67 = fma(z, -1.0 / 87178291200.0, 1.0 / 479001600.0);
  const double c8 = 1.0 / 20922789888000.0;
  cs = fma(z4, fma(z4, c8, fma(z2, c67, c45)), fma(z2, c23, c01));
}
constexpr double kSmallAngle2 = 0.0625;  // (0.25 rad)^2

// (sin, cos)(a + d) from (sin, cos)(a); exact sincos when the step is large.
__device__ __forceinline__ void advance_sincos(double a_new, double d, double s, double c,
                                               double &s_new, double &c_new) {
  const double z = d * d;
  if (z <= kSmallAngle2) {
    double sinc, cd;
    small_sinc_cos(z, sinc, cd);
    const double sd = d * sinc;
    s_new = fma(s, cd, c * sd);
    c_new = fma(c, cd, -s * sd);
  } else {
    sincos(a_new, &s_new, &c_new);
  }
}

// ceres::EigenQuaternionManifold::Plus with the small-angle forms (no sqrt, no
// division, no sin/cos call while |d| <= 0.25 rad).
__device__ __forceinline__ void quat_plus_fast(const double x[4], const double d[3],
                                               double out[4]) {
  const double z = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double s, dw;
  if (z <= kSmallAngle2) {
    small_sinc_cos(z, s, dw);
  } else {
    const double nd = sqrt(z);
    double sn;
    sincos(nd, &sn, &dw);
    s = sn / nd;
  }
  const double dx = s * d[0], dy = s * d[1], dz = s * d[2];
  const double xx = x[0], xy = x[1], xz = x[2], xw = x[3];
  out[3] = dw * xw - dx * xx - dy * xy - dz * xz;
  out[0] = dw * xx + dx * xw + dy * xz - dz * xy;
  out[1] = dw * xy - dx * xz + dy * xw + dz * xx;
  out[2] = dw * xz + dx * xy - dy * xx + dz * xw;
}

// Pose constants from sin/cos of (theta, phi) and the quaternion.
__device__ __forceinline__ void make_pose_const_sc(const double sc[4], const double q[4],
                                                   PoseConst &pc) {
  const double st = sc[0], ct = sc[1], sp = sc[2], cp = sc[3];
  pc.t[0] = st * cp;  pc.t[1] = st * sp;  pc.t[2] = ct;
  pc.tth[0] = ct * cp; pc.tth[1] = ct * sp; pc.tth[2] = -st;
  pc.tph[0] = -st * sp; pc.tph[1] = st * cp; pc.tph[2] = 0.0;
  const double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
  const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  pc.R[0] = 1.0 - (tyy + tzz); pc.R[1] = txy - twz;         pc.R[2] = txz + twy;
  pc.R[3] = txy + twz;         pc.R[4] = 1.0 - (txx + tzz); pc.R[5] = tyz - twx;
  pc.R[6] = txz - twy;         pc.R[7] = tyz + twx;         pc.R[8] = 1.0 - (txx + tyy);
}

// Rare path of lm_step (a step above 0.25 rad): exact sin/cos.  Kept out of line so that the
// Payne-Hanek slow paths of sincos() do not bloat the registers of the common path.
__device__ __noinline__ void lm_candidate_large_step(const double *x, const double *sc,
                                                     const double *delta, double *cand,
                                                     double *scc) {
  advance_sincos(cand[0], delta[0], sc[0], sc[1], scc[0], scc[1]);
  advance_sincos(cand[1], delta[1], sc[2], sc[3], scc[2], scc[3]);
  quat_plus_fast(x + 2, delta + 2, cand + 2);
}

// Rare path of gradient_converged: the quaternion part of || x - Plus(x, -g) ||_inf.
__device__ __noinline__ double quaternion_chord_max(const double *q, const double *g) {
  const double ng[3] = {-g[0], -g[1], -g[2]};
  double qs[4];
  quat_plus_fast(q, ng, qs);
  double m = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) m = fmax(m, fabs(q[i] - qs[i]));
  return m;
}

// gradient_max_norm <= tol, where gradient_max_norm = || x - Plus(x, -g) ||_inf
// (TrustRegionMinimizer::EvaluateGradientAndJacobian).  The theta/phi part of
// that norm is |g0|, |g1|; the quaternion part is only evaluated when those pass.
__device__ __forceinline__ bool gradient_converged(const double x[6], const double g[5],
                                                   double tol) {
  if (!(fmax(fabs(g[0]), fabs(g[1])) <= tol)) return false;
  const double nd2 = g[2] * g[2] + g[3] * g[3] + g[4] * g[4];
  if (nd2 > kSmallAngle2) return false;  // chord 2 sin(|g|/2) is far above any tolerance
  return quaternion_chord_max(x + 2, g + 2) <= tol;
}

// 5x5 SPD solve A d = b by block elimination over the (theta, phi | rotation) split with
// closed-form 2x2 / 3x3 adjugate inverses: 2 reciprocals and a ~30-deep dependent chain
// (LDL^T + substitutions is ~65 deep).  Returns false if A is not positive definite
// (Sylvester: all leading minors of P and of the Schur complement positive) or the
// solution is not finite: LINEAR_SOLVER_FAILURE in Ceres terms, an invalid step.
__device__ __forceinline__ bool spd_solve5(const double A[5][5], const double b[5], double d[5]) {
  // P = A[0:2,0:2]
  const double p00 = A[0][0], p01 = A[0][1], p11 = A[1][1];
  const double detP = fma(p00, p11, -p01 * p01);
  const double iP = fast_rcp(detP);
  const double i00 = p11 * iP, i01 = -p01 * iP, i11 = p00 * iP;  // P^-1
  // W = P^-1 Q (2x3), u = P^-1 b1
  double W0[3], W1[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    W0[j] = fma(i00, A[0][2 + j], i01 * A[1][2 + j]);
    W1[j] = fma(i01, A[0][2 + j], i11 * A[1][2 + j]);
  }
  const double u0 = fma(i00, b[0], i01 * b[1]), u1 = fma(i01, b[0], i11 * b[1]);
  // Schur complement S = R - Q^T W (symmetric 3x3), c = b2 - Q^T u
  double S[3][3], c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = i; j < 3; ++j) {
      S[i][j] = A[2 + i][2 + j] - fma(A[0][2 + i], W0[j], A[1][2 + i] * W1[j]);
      S[j][i] = S[i][j];
    }
    c[i] = b[2 + i] - fma(A[0][2 + i], u0, A[1][2 + i] * u1);
  }
  // adjugate of S
  const double c00 = fma(S[1][1], S[2][2], -S[1][2] * S[1][2]);
  const double c01 = fma(S[0][2], S[1][2], -S[0][1] * S[2][2]);
  const double c02 = fma(S[0][1], S[1][2], -S[0][2] * S[1][1]);
  const double c11 = fma(S[0][0], S[2][2], -S[0][2] * S[0][2]);
  const double c12 = fma(S[0][1], S[0][2], -S[0][0] * S[1][2]);
  const double c22 = fma(S[0][0], S[1][1], -S[0][1] * S[0][1]);
  const double detS = fma(S[0][0], c00, fma(S[0][1], c01, S[0][2] * c02));
  const double iS = fast_rcp(detS);
  const double n0 = fma(c00, c[0], fma(c01, c[1], c02 * c[2]));
  const double n1 = fma(c01, c[0], fma(c11, c[1], c12 * c[2]));
  const double n2 = fma(c02, c[0], fma(c12, c[1], c22 * c[2]));
  d[2] = n0 * iS;
  d[3] = n1 * iS;
  d[4] = n2 * iS;
  d[0] = u0 - fma(W0[0], d[2], fma(W0[1], d[3], W0[2] * d[4]));
  d[1] = u1 - fma(W1[0], d[2], fma(W1[1], d[3], W1[2] * d[4]));
  bool ok = (p00 > 0.0) && (detP > 0.0) && (S[0][0] > 0.0) && (c22 > 0.0) && (detS > 0.0);
#pragma unroll
  for (int i = 0; i < 5; ++i) ok = ok && (fabs(d[i]) < CUDART_INF);
  return ok;
}

// One trust-region solve attempt at the accepted point.  Ceres solves, in Jacobi-scaled
// coordinates, (S H S + D) y = S g with D = diag / radius and takes delta = -S y.  With
// d = S y this is (H + S^-1 D S^-1) d = g: the same system without the 45 scaling products,
// and model_cost_change = (y.Sg + y^T D y) / 2 = sum d_a (g_a + L_a d_a) / 2, L = D / s^2.
// Returns false for an invalid step (LINEAR_SOLVER_FAILURE or model_cost_change <= 0).
__device__ __forceinline__ bool lm_trust_region_step(const double *Hg, const double scale[5],
                                                     const double inv_scale2[5], double diag[5],
                                                     bool reuse_diagonal, double inv_radius,
                                                     const pnec_solver_opts &o, double delta[5],
                                                     double &mcc) {
  double A[5][5], lam[5], d[5];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
#pragma unroll
    for (int b = a; b < 5; ++b) {
      A[a][b] = Hg[tri(a, b)];
      A[b][a] = A[a][b];
    }
  }
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    if (!reuse_diagonal)  // squared column norm of the scaled Jacobian, clamped
      diag[a] = fmin(fmax(A[a][a] * scale[a] * scale[a], o.min_lm_diagonal), o.max_lm_diagonal);
    lam[a] = diag[a] * inv_radius * inv_scale2[a];
    A[a][a] += lam[a];
  }
  bool valid = spd_solve5(A, Hg + 15, d);
  mcc = 0.0;
#pragma unroll
  for (int a = 0; a < 5; ++a) mcc = fma(d[a], fma(lam[a], d[a], Hg[15 + a]), mcc);
  mcc *= 0.5;
#pragma unroll
  for (int a = 0; a < 5; ++a) delta[a] = -d[a];
  return valid && (mcc > 0.0);
}

// The whole LM update between two evaluations, straight-line:
//   (FIRST)  IterationZero bookkeeping,
//   (!FIRST) FunctionToleranceReached, IsStepSuccessful, StepAccepted / StepRejected,
//   then FinalizeIterationAndCheckIfMinimizerCanContinue, ComputeTrustRegionStep (invalid
//   steps retried in place: they need no evaluation), candidate = Plus(x, delta),
//   ParameterToleranceReached, choice of the next pass, pose constants of the candidate.
// `st` is the shared-memory state; every lane of the calling warp executes this with
// identical values, lane 0 stores.  New totals are in st.tot[st.ti ^ 1].
#ifdef PNEC_PHASE_TIMING
__device__ unsigned long long g_lm_probe[8];
#define LM_PROBE(k) do { const long long t_now = clock64(); if (lane == 0) atomicAdd(&g_lm_probe[k], (unsigned long long)(t_now - t_prev)); t_prev = t_now; } while (0)
#else
#define LM_PROBE(k) do { } while (0)
#endif

// The LM update runs on lane 0 of the LM warp only (measured 4 % faster than all 32 lanes
// redundantly, and it keeps the fp64 pipe free for the co-resident CTAs' evaluations).
// PNEC_LM_FULL_WARP=1 restores the redundant variant.
#ifndef PNEC_LM_FULL_WARP
#define PNEC_LM_FULL_WARP 0
#endif
constexpr bool kLmFullWarp = PNEC_LM_FULL_WARP != 0;

template <bool FIRST>
__device__ __forceinline__ void lm_step(LMState &st, const pnec_solver_opts &o, int lane,
                                        PoseConst &s_pc) {
#ifdef PNEC_PHASE_TIMING
  long long t_prev = clock64();
#endif
  const int xi = st.xi, ti = st.ti;
  const double *newtot = st.tot[ti ^ 1];
  // The trust-region radius is carried as its inverse: every use is a division by it
  // (D = diag / radius, radius /= factor), so the update chain needs no reciprocal.
  double inv_radius = st.inv_radius, decrease_factor = st.decrease_factor, x_cost = st.x_cost;
  int reuse_diagonal = st.reuse_diagonal;
  bool accept = true;
  double cand_cost = newtot[20];
  if (!(fabs(cand_cost) < CUDART_INF)) cand_cost = DBL_MAX;
  if (FIRST) {
    if (cand_cost >= DBL_MAX) {
      if (lane == 0) { st.status = PNEC_STATUS_NONFINITE; st.done = 1; st.initial_cost = newtot[20]; st.x_cost = newtot[20]; }
      return;
    }
    x_cost = cand_cost;
  } else {
    const double cost_change = x_cost - cand_cost;
    if (fabs(cost_change) <= o.function_tolerance * x_cost) {  // before acceptance, like Ceres
      if (lane == 0) { st.status = PNEC_STATUS_CONVERGED_FUNCTION; st.done = 1; }
      return;
    }
    const double rho = (cand_cost >= DBL_MAX) ? -DBL_MAX : cost_change * st.inv_model_cost_change;
    accept = rho > o.min_relative_decrease;
    const double c = 2.0 * rho - 1.0;
    // StepAccepted: radius /= max(1/3, 1 - (2 rho - 1)^3), capped; StepRejected: radius /= factor
    inv_radius *= accept ? fmax(1.0 / 3.0, 1.0 - c * c * c) : decrease_factor;
    if (accept) inv_radius = fmax(inv_radius, 1.0 / o.max_trust_region_radius);
    decrease_factor = accept ? 2.0 : 2.0 * decrease_factor;
    reuse_diagonal = accept ? 0 : 1;
    x_cost = accept ? cand_cost : x_cost;
  }
  LM_PROBE(0);  // judge
  const int nxi = accept ? (FIRST ? xi : xi ^ 1) : xi;  // index of the accepted point
  const int nti = accept ? ti ^ 1 : ti;                 // index of its totals
  const double *Hg = st.tot[nti];
  const double *x = st.pts[nxi];
  const double *sc = st.scs[nxi];

  double scale[5], inv_scale2[5], diag[5];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    if (FIRST) {
      const double one_plus = 1.0 + fast_sqrt(Hg[tri(a, a)]);  // scale = 1 / (1 + |J col|)
      scale[a] = o.jacobi_scaling ? fast_rcp(one_plus) : 1.0;
      inv_scale2[a] = o.jacobi_scaling ? one_plus * one_plus : 1.0;
    } else {
      scale[a] = st.scale[a];
      inv_scale2[a] = st.inv_scale2[a];
    }
    diag[a] = FIRST ? 0.0 : st.diag[a];
  }
  // the gradient test of a newly accepted point (cheap unless |g| is already tiny)
  int grad_conv = st.grad_converged;
  if (accept) grad_conv = gradient_converged(x, Hg + 15, o.gradient_tolerance) ? 1 : 0;
  LM_PROBE(1);  // scale, gradient test
  int iteration = st.iteration, num_invalid = st.num_invalid, step_successful = accept ? 1 : 0;
  int status = st.status, done = 0;
  double delta[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, mcc = 0.0;
  for (;;) {  // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iteration >= o.max_num_iterations) { status = PNEC_STATUS_MAX_ITERATIONS; done = 1; break; }
    if (step_successful && grad_conv) { status = PNEC_STATUS_CONVERGED_GRADIENT; done = 1; break; }
    if (inv_radius * o.min_trust_region_radius >= 1.0) { status = PNEC_STATUS_CONVERGED_RADIUS; done = 1; break; }
    ++iteration;
    const bool valid = lm_trust_region_step(Hg, scale, inv_scale2, diag, reuse_diagonal != 0, inv_radius, o, delta, mcc);
    reuse_diagonal = 1;
    if (valid) { num_invalid = 0; break; }
    if (++num_invalid >= o.max_num_consecutive_invalid_steps) { status = PNEC_STATUS_FAILURE; done = 1; break; }
    inv_radius *= 2.0;  // StepIsInvalid: radius *= 0.5
    step_successful = 0;
  }

  LM_PROBE(2);  // trust-region step
  double cand[6], scc[4];
  int pass_mode = kPassFull;
  if (!done) {
    cand[0] = x[0] + delta[0];
    cand[1] = x[1] + delta[1];
    const double zt = delta[0] * delta[0], zp = delta[1] * delta[1];
    const double zq = fma(delta[2], delta[2], fma(delta[3], delta[3], delta[4] * delta[4]));
    if (fmax(fmax(zt, zp), zq) <= kSmallAngle2) {
      // common case, one branch: three independent small-angle evaluations + angle addition
      double sct, cdt, scp, cdp, sq, cq;
      small_sinc_cos(zt, sct, cdt);
      small_sinc_cos(zp, scp, cdp);
      small_sinc_cos(zq, sq, cq);
      const double sdt = delta[0] * sct, sdp = delta[1] * scp;
      scc[0] = fma(sc[0], cdt, sc[1] * sdt);
      scc[1] = fma(sc[1], cdt, -sc[0] * sdt);
      scc[2] = fma(sc[2], cdp, sc[3] * sdp);
      scc[3] = fma(sc[3], cdp, -sc[2] * sdp);
      const double dx = sq * delta[2], dy = sq * delta[3], dz = sq * delta[4];
      const double xx = x[2], xy = x[3], xz = x[4], xw = x[5];
      cand[5] = cq * xw - dx * xx - dy * xy - dz * xz;
      cand[2] = cq * xx + dx * xw + dy * xz - dz * xy;
      cand[3] = cq * xy - dx * xz + dy * xw + dz * xx;
      cand[4] = cq * xz + dx * xy - dy * xx + dz * xw;
    } else {
      lm_candidate_large_step(x, sc, delta, cand, scc);
    }
    // ParameterToleranceReached depends on the step only.  Ceres evaluates the candidate's
    // cost first and then returns without applying the step, so the evaluation cannot change
    // the outcome: decide here and skip it.
    const double e0 = x[0] - cand[0], e1 = x[1] - cand[1], e2 = x[2] - cand[2];
    const double e3 = x[3] - cand[3], e4 = x[4] - cand[4], e5 = x[5] - cand[5];
    const double sn = fma(e0, e0, e1 * e1) + fma(e2, e2, e3 * e3) + fma(e4, e4, e5 * e5);
    // |x| <= |theta| + |phi| + |q| gives a cheap upper bound of the tolerance; the exact
    // norm is only formed when the step is small enough for the test to possibly pass.
    const double ptol_hi = o.parameter_tolerance * (fabs(x[0]) + fabs(x[1]) + 2.0 + o.parameter_tolerance);
    if (sn <= ptol_hi * ptol_hi) {
      double xn = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) xn = fma(x[i], x[i], xn);
      const double ptol = o.parameter_tolerance * (sqrt(xn) + o.parameter_tolerance);
      if (sn <= ptol * ptol) {  // step_norm <= tolerance, compared squared
        status = PNEC_STATUS_CONVERGED_PARAMETER;
        done = 1;
      }
    }
    // If the quadratic model already predicts |cost change| <= function_tolerance * cost the
    // iteration will almost surely terminate there: evaluate the cost alone first.
    pass_mode = (mcc <= 1.05 * o.function_tolerance * x_cost) ? kPassCost : kPassFull;
  }
  LM_PROBE(3);  // candidate
  if (kLmFullWarp) __syncwarp();  // every lane has read the state it needs; lane 0 may now overwrite it
  if (lane == 0) {
    if (!done) {
      PoseConst pcn;
      make_pose_const_sc(scc, cand + 2, pcn);
      s_pc = pcn;
#pragma unroll
      for (int i = 0; i < 6; ++i) st.pts[nxi ^ 1][i] = cand[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) st.scs[nxi ^ 1][i] = scc[i];
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      if (FIRST) {
        st.scale[a] = scale[a];
        st.inv_scale2[a] = inv_scale2[a];
      }
      st.diag[a] = diag[a];
    }
    st.xi = nxi;
    st.ti = nti;
    st.x_cost = x_cost;
    if (FIRST) st.initial_cost = x_cost;
    st.inv_radius = inv_radius;
    st.decrease_factor = decrease_factor;
    st.inv_model_cost_change = (mcc > 0.0) ? fast_rcp(mcc) : 0.0;  // off the critical path
    st.iteration = iteration;
    st.num_invalid = num_invalid;
    st.reuse_diagonal = reuse_diagonal;
    st.step_successful = step_successful;
    st.grad_converged = grad_conv;
    st.status = status;
    st.done = done;
    st.pass_mode = pass_mode;
  }
  LM_PROBE(4);  // pose constants + state store
}

// After a cost-only pass: FunctionToleranceReached, or fall through to a full pass at the
// same candidate (nothing else is decided from a cost-only pass).
__device__ __forceinline__ void lm_after_cost_pass(LMState &st, double cand_cost_in,
                                                   const pnec_solver_opts &o, int lane) {
  double cand_cost = cand_cost_in;
  if (!(fabs(cand_cost) < CUDART_INF)) cand_cost = DBL_MAX;
  const bool conv = fabs(st.x_cost - cand_cost) <= o.function_tolerance * st.x_cost;
  if (lane == 0) {
    if (conv) {
      st.status = PNEC_STATUS_CONVERGED_FUNCTION;
      st.done = 1;
    } else {
      st.pass_mode = kPassFull;
    }
  }
}

}  // namespace pnec
