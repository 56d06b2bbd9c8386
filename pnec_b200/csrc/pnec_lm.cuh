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
// The per-iteration update is the serial part of a solve, so it is written as straight-line code
// shared out over the lanes of one warp (see "warp-cooperative update" below): Newton-refined MUFU
// reciprocals instead of IEEE division, adjugate block elimination instead of Cholesky (no square
// roots), small-angle polynomials + angle addition instead of sin/cos calls.  All of it stays within
// a few ulp of the IEEE forms.
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
  double xch[40];    // exchange between the lanes of the warp that runs the update (lm_trust_region_step)
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
  // explicit roundings (no contraction into fma): pose_const_lanes computes the same bits
  const double twx = __dmul_rn(tx, qw), twy = __dmul_rn(ty, qw), twz = __dmul_rn(tz, qw);
  const double txx = __dmul_rn(tx, qx), txy = __dmul_rn(ty, qx), txz = __dmul_rn(tz, qx);
  const double tyy = __dmul_rn(ty, qy), tyz = __dmul_rn(tz, qy), tzz = __dmul_rn(tz, qz);
  pc.R[0] = __dsub_rn(1.0, __dadd_rn(tyy, tzz)); pc.R[1] = __dsub_rn(txy, twz); pc.R[2] = __dadd_rn(txz, twy);
  pc.R[3] = __dadd_rn(txy, twz); pc.R[4] = __dsub_rn(1.0, __dadd_rn(txx, tzz)); pc.R[5] = __dsub_rn(tyz, twx);
  pc.R[6] = __dsub_rn(txz, twy); pc.R[7] = __dadd_rn(tyz, twx); pc.R[8] = __dsub_rn(1.0, __dadd_rn(txx, tyy));
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

// ---------------------------------------------------------------- warp-cooperative update
//
// The update between two evaluations is the serial part of a solve: one warp runs it while the pair's
// evaluation waits, so what counts is its dependent chain (measured on B200: DFMA 8 cycles, a
// shared-memory load 29, a store + __syncwarp + load exchange between lanes 43).  The warp works on it
// together: each stage computes its independent outputs (the five damped diagonal entries, the 2x3 block
// W and u, the six entries of the Schur complement and its right-hand side, the six cofactors, the three
// small-angle evaluations, the four quaternion components, the eighteen pose constants) one per lane, and
// the lanes exchange them through a 40-double scratch in shared memory.  Everything a lane needs from
// the state is read in ONE batch of loads right after the accept / reject decision, before the first
// scratch store (the compiler cannot move a shared-memory load across a store it cannot tell apart),
// so the chain sees one load latency, seven exchanges and about 55 dependent fp64 operations.
// The arithmetic of every output is what a single lane would compute.

// scratch layout (LMState::xch)
constexpr int kXDiag = 0;   // 5: H_aa + lambda_a
constexpr int kXLam = 5;    // 5: lambda_a
constexpr int kXW0 = 10;    // 4: W0[0..2], u0        (P^-1 Q, P^-1 b1: first rows)
constexpr int kXW1 = 14;    // 4: W1[0..2], u1
constexpr int kXS = 18;     // 6: S00 S01 S02 S11 S12 S22 (Schur complement), then
constexpr int kXC = 24;     // 3: c = b2 - Q^T u
constexpr int kXAdj = 27;   // 6: c00 c01 c02 c11 c12 c22 (cofactors of S)
constexpr int kXD = 33;     // 5: d

// value `lane` of a table of 4-bit entries
__device__ __forceinline__ int nib(unsigned long long table, int lane) {
  return static_cast<int>((table >> (4 * lane)) & 0xFull);
}

// Pose constants from a quaternion q (x, y, z, w) and (sin theta, cos theta, sin phi, cos phi) in shared
// memory, one per lane: R (row-major) on lanes 0..8, t, dt/dtheta, dt/dphi on lanes 9..17.
__device__ __forceinline__ void pose_const_lanes(const double *q, const double *scn, PoseConst &s_pc, int lane) {
  if (lane >= 18) return;
  double v;
  if (lane < 9) {
    //   R0 = 1 - (tyy + tzz)   R1 = txy - twz         R2 = txz + twy          with tab = 2 q_a q_b
    //   R3 = txy + twz         R4 = 1 - (txx + tzz)   R5 = tyz - twx
    //   R6 = txz - twy         R7 = tyz + twx         R8 = 1 - (txx + tyy)
    // (explicit roundings: make_pose_const_sc must give the same bits)
    const double p = __dmul_rn(2.0 * q[nib(0x022201211ull, lane)], q[nib(0x010100001ull, lane)]);
    const double r = __dmul_rn(2.0 * q[nib(0x101022122ull, lane)], q[nib(0x133323332ull, lane)]);
    const bool diagonal = (lane == 0) || (lane == 4) || (lane == 8);
    const bool minus = (lane == 1) || (lane == 5) || (lane == 6);
    v = diagonal ? __dsub_rn(1.0, __dadd_rn(p, r)) : __dadd_rn(p, minus ? -r : r);
  } else {
    //   t = (st cp, st sp, ct)   dt/dtheta = (ct cp, ct sp, -st)   dt/dphi = (-st sp, st cp, 0)
    const int m = lane - 9;
    const double u = scn[nib(0x000011100ull, m)];                      // st or ct
    const double su = (m == 5 || m == 6) ? -u : u;
    const int vi = nib(0x032423423ull, m);                             // 2: sp, 3: cp, 4: the constant 1
    const double w = (vi == 4) ? 1.0 : scn[vi & 3];
    v = (m == 8) ? 0.0 : su * w;
  }
  reinterpret_cast<double *>(&s_pc)[lane] = v;
}

// What a lane reads of the totals (J^T J upper triangle, J^T r) at the accepted point.
struct LmLaneTotals {
  double haa;      // H[a][a], a = min(lane, 4)
  double h01;      // H[0][1]
  double wx, wy;   // lanes 0..2: H[0][2+j], H[1][2+j]; lane 3: g0, g1
  double sp, sq;   // lanes 0..8: H[0][2+i], H[1][2+i] of the lane's Schur entry
  double sbase;    // off-diagonal H[2+i][2+j] (lanes 1, 2, 4) or g[2+i] (lanes 6..8)
  double g[5];     // J^T r
};

// One trust-region solve attempt at the accepted point, all 32 lanes.  Ceres solves, in Jacobi-scaled
// coordinates, (S H S + D) y = S g with D = diag / radius and takes delta = -S y.  With d = S y this is
// (H + S^-1 D S^-1) d = g: the same system without the 45 scaling products, and
// model_cost_change = (y.Sg + y^T D y) / 2 = sum d_a (g_a + L_a d_a) / 2, L = D / s^2.
// The 5x5 SPD system is solved by block elimination over the (theta, phi | rotation) split with
// closed-form 2x2 / 3x3 adjugate inverses (2 reciprocals).  Lane a < 5 owns parameter a: diag_l in/out;
// d[0..4] out on every lane.  Returns false for an invalid step: LINEAR_SOLVER_FAILURE in Ceres terms
// (not positive definite by Sylvester's criterion, or a non-finite solution) or model_cost_change <= 0.
__device__ __forceinline__ bool lm_trust_region_step(const LmLaneTotals &t, double *xch, int lane,
                                                     double scale_l, double inv_scale2_l, double &diag_l,
                                                     bool reuse_diagonal, double inv_radius,
                                                     const pnec_solver_opts &o, double d[5], double &mcc) {
  if (!reuse_diagonal)  // squared column norm of the scaled Jacobian, clamped
    diag_l = fmin(fmax(t.haa * scale_l * scale_l, o.min_lm_diagonal), o.max_lm_diagonal);
  const double lam_l = diag_l * inv_radius * inv_scale2_l;
  if (lane < 5) {
    xch[kXDiag + lane] = t.haa + lam_l;
    xch[kXLam + lane] = lam_l;
  }
  __syncwarp();
  // P = A[0:2,0:2] and its inverse: every lane
  const double p00 = xch[kXDiag], p11 = xch[kXDiag + 1];
  const int ls = min(lane, 8);                                          // Schur stage: entry of this lane
  const int si = nib(0x210211000ull, ls), sk = nib(0x333221210ull, ls);  // row i; column j (3: right-hand side)
  const bool sdiag = (ls == 0) || (ls == 3) || (ls == 5);
  const double sdiag_v = xch[kXDiag + 2 + si];
  double lam[5];
#pragma unroll
  for (int a = 0; a < 5; ++a) lam[a] = xch[kXLam + a];
  const double detP = fma(p00, p11, -t.h01 * t.h01);
  const double iP = fast_rcp(detP);
  const double i00 = p11 * iP, i01 = -t.h01 * iP, i11 = p00 * iP;
  if (lane < 4) {  // W = P^-1 Q (lanes 0..2: column j), u = P^-1 b1 (lane 3)
    xch[kXW0 + lane] = fma(i00, t.wx, i01 * t.wy);
    xch[kXW1 + lane] = fma(i01, t.wx, i11 * t.wy);
  }
  __syncwarp();
  {  // Schur complement S = R - Q^T W (lanes 0..5: (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)), c = b2 - Q^T u (lanes 6..8)
    const double base = sdiag ? sdiag_v : t.sbase;
    const double sv = base - fma(t.sp, xch[kXW0 + sk], t.sq * xch[kXW1 + sk]);
    if (lane < 9) xch[kXS + lane] = sv;
  }
  __syncwarp();
  const double *S = xch + kXS;
  {  // cofactors of S: lanes 0..5
    const int l = min(lane, 5);
    const double v = fma(S[nib(0x010123ull, l)], S[nib(0x325445ull, l)], -S[nib(0x102214ull, l)] * S[nib(0x142354ull, l)]);
    if (lane < 6) xch[kXAdj + lane] = v;
  }
  const double S00 = S[0], S01 = S[1], S02 = S[2];
  const double c0 = xch[kXC], c1 = xch[kXC + 1], c2 = xch[kXC + 2];
  __syncwarp();
  const double *C = xch + kXAdj;
  const double c22 = C[5];
  const double detS = fma(S00, C[0], fma(S01, C[1], S02 * C[2]));
  const double iS = fast_rcp(detS);
  {  // rotation part: lanes 0..2, row r of adj(S) . c / det
    const int r = min(lane, 2);
    const int e1 = nib(0x431ull, r), e2 = nib(0x542ull, r);  // rows (0 1 2), (1 3 4), (2 4 5) of the packed cofactors
    const double n = fma(C[r], c0, fma(C[e1], c1, C[e2] * c2));
    if (lane < 3) xch[kXD + 2 + lane] = n * iS;
  }
  const double *W = xch + ((lane & 1) ? kXW1 : kXW0);
  const double w0 = W[0], w1 = W[1], w2 = W[2], wu = W[3];
  __syncwarp();
  {  // translation part: lanes 0..1
    const double v = wu - fma(w0, xch[kXD + 2], fma(w1, xch[kXD + 3], w2 * xch[kXD + 4]));
    if (lane < 2) xch[kXD + lane] = v;
  }
  __syncwarp();
  bool ok = (p00 > 0.0) && (detP > 0.0) && (S00 > 0.0) && (c22 > 0.0) && (detS > 0.0);
  double term[5];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    d[a] = xch[kXD + a];
    ok = ok && (fabs(d[a]) < CUDART_INF);
    term[a] = fma(lam[a], d[a], t.g[a]);
  }
  mcc = 0.0;
#pragma unroll
  for (int a = 0; a < 5; ++a) mcc = fma(d[a], term[a], mcc);
  mcc *= 0.5;
  return ok && (mcc > 0.0);
}

// The whole LM update between two evaluations, executed by all 32 lanes of one warp (converged):
//   (FIRST)  IterationZero bookkeeping,
//   (!FIRST) FunctionToleranceReached, IsStepSuccessful, StepAccepted / StepRejected,
//   then FinalizeIterationAndCheckIfMinimizerCanContinue, ComputeTrustRegionStep (invalid
//   steps retried in place: they need no evaluation), candidate = Plus(x, delta),
//   ParameterToleranceReached, choice of the next pass, pose constants of the candidate.
// `st` is the shared-memory state; the new totals are in st.tot[st.ti ^ 1].  The caller synchronises
// the warp before reading what this wrote.
#ifdef PNEC_PHASE_TIMING
__device__ unsigned long long g_lm_probe[8];
#define LM_PROBE(k) do { const long long t_now = clock64(); if (lane == 0) atomicAdd(&g_lm_probe[k], (unsigned long long)(t_now - t_prev)); t_prev = t_now; } while (0)
#else
#define LM_PROBE(k) do { } while (0)
#endif

// (`first` is a run-time flag, not a template parameter: two instantiations doubled the update's code, and the
// update is bound by instruction fetch -- the L1.5 instruction cache delivers ~4 lines per cycle to the whole
// GPU, and the solve kernels used 66-73 % of that with the update's code streamed in for every update.)
// `post(pass_mode)` is called by the whole warp as soon as the candidate's pose constants are stored.
struct LmNoPost {
  __device__ __forceinline__ void operator()(int) const {}
};
template <class Post = LmNoPost>
__device__ __forceinline__ void lm_step(const bool FIRST, LMState &st, const pnec_solver_opts &o, int lane,
                                        PoseConst &s_pc, Post post = Post()) {
#ifdef PNEC_PHASE_TIMING
  long long t_prev = clock64();
#endif
  const int xi = st.xi, ti = st.ti;
  // The trust-region radius is carried as its inverse: every use is a division by it
  // (D = diag / radius, radius /= factor), so the update chain needs no reciprocal.
  double inv_radius = st.inv_radius, decrease_factor = st.decrease_factor, x_cost = st.x_cost;
  int reuse_diagonal = st.reuse_diagonal;
  const double inv_mcc_prev = st.inv_model_cost_change;
  int grad_conv = st.grad_converged;
  int iteration = st.iteration, num_invalid = st.num_invalid, status = st.status;
  bool accept = true;
  const double new_cost = st.tot[ti ^ 1][20];
  double cand_cost = new_cost;
  if (!(fabs(cand_cost) < CUDART_INF)) cand_cost = DBL_MAX;
  if (FIRST) {
    if (cand_cost >= DBL_MAX) {
      if (lane == 0) { st.status = PNEC_STATUS_NONFINITE; st.done = 1; st.initial_cost = new_cost; st.x_cost = new_cost; }
      return;
    }
    x_cost = cand_cost;
  } else {
    const double cost_change = x_cost - cand_cost;
    if (fabs(cost_change) <= o.function_tolerance * x_cost) {  // before acceptance, like Ceres
      if (lane == 0) { st.status = PNEC_STATUS_CONVERGED_FUNCTION; st.done = 1; }
      return;
    }
    const double rho = (cand_cost >= DBL_MAX) ? -DBL_MAX : cost_change * inv_mcc_prev;
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
  double *cnd = st.pts[nxi ^ 1], *scn = st.scs[nxi ^ 1];  // the candidate's slots

  // ---- the one batch of state loads (lane a < 5 owns parameter a, the others shadow parameter 4)
  const int a5 = min(lane, 4);
  LmLaneTotals t;
  double x[6], sc_s, sc_c, q4[4];
  double scale_l = 1.0, inv_scale2_l = 1.0, diag_l = 0.0;
  {
    const double *Hg = st.tot[nti];
    const double *xs = st.pts[nxi];
    const double *scs = st.scs[nxi];
    const int lw = min(lane, 3), ls = min(lane, 8);
    const int si = nib(0x210211000ull, ls);
    t.haa = Hg[a5 * 5 - (a5 * (a5 - 1)) / 2];  // tri(a, a)
    t.h01 = Hg[1];
    t.wx = Hg[lw < 3 ? 2 + lw : 15];
    t.wy = Hg[lw < 3 ? 6 + lw : 16];
    t.sp = Hg[2 + si];
    t.sq = Hg[6 + si];
    t.sbase = Hg[ls < 6 ? 9 + ls : 17 + si];
#pragma unroll
    for (int a = 0; a < 5; ++a) t.g[a] = Hg[15 + a];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = xs[i];
    sc_s = scs[2 * (lane & 1)];
    sc_c = scs[2 * (lane & 1) + 1];
    const int k = (lane - 2) & 3;  // quaternion component of lanes 2..5
    q4[0] = xs[2 + k];
    q4[1] = xs[2 + nib(0x0123ull, k)];
    q4[2] = xs[2 + nib(0x1032ull, k)];
    q4[3] = xs[2 + nib(0x2301ull, k)];
    if (!FIRST) {
      scale_l = st.scale[a5];
      inv_scale2_l = st.inv_scale2[a5];
      diag_l = st.diag[a5];
    }
  }
  if (FIRST) {
    const double one_plus = 1.0 + fast_sqrt(t.haa);  // scale = 1 / (1 + |J col|)
    scale_l = o.jacobi_scaling ? fast_rcp(one_plus) : 1.0;
    inv_scale2_l = o.jacobi_scaling ? one_plus * one_plus : 1.0;
    diag_l = 0.0;
  }
  // the gradient test of a newly accepted point (cheap unless |g| is already tiny)
  if (accept) grad_conv = gradient_converged(x, t.g, o.gradient_tolerance) ? 1 : 0;
  LM_PROBE(1);  // loads, scale, gradient test
  int step_successful = accept ? 1 : 0, done = 0;
  double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, mcc = 0.0;
  for (;;) {  // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iteration >= o.max_num_iterations) { status = PNEC_STATUS_MAX_ITERATIONS; done = 1; break; }
    if (step_successful && grad_conv) { status = PNEC_STATUS_CONVERGED_GRADIENT; done = 1; break; }
    if (inv_radius * o.min_trust_region_radius >= 1.0) { status = PNEC_STATUS_CONVERGED_RADIUS; done = 1; break; }
    ++iteration;
    const bool valid = lm_trust_region_step(t, st.xch, lane, scale_l, inv_scale2_l, diag_l, reuse_diagonal != 0,
                                            inv_radius, o, d, mcc);
    reuse_diagonal = 1;
    if (valid) { num_invalid = 0; break; }
    if (++num_invalid >= o.max_num_consecutive_invalid_steps) { status = PNEC_STATUS_FAILURE; done = 1; break; }
    inv_radius *= 2.0;  // StepIsInvalid: radius *= 0.5
    step_successful = 0;
  }

  LM_PROBE(2);  // trust-region step
  int pass_mode = kPassFull;
  if (!done) {
    const double delta[5] = {-d[0], -d[1], -d[2], -d[3], -d[4]};
    const double zt = delta[0] * delta[0], zp = delta[1] * delta[1];
    const double zq = fma(delta[2], delta[2], fma(delta[3], delta[3], delta[4] * delta[4]));
    if (fmax(fmax(zt, zp), zq) <= kSmallAngle2) {
      // common case: the three small-angle evaluations on lanes 0 (theta), 1 (phi), 2.. (rotation)
      double sinc, cs;
      small_sinc_cos(lane == 0 ? zt : lane == 1 ? zp : zq, sinc, cs);
      if (lane < 2) {  // angle addition
        const double dl = lane ? delta[1] : delta[0], sd = dl * sinc;
        cnd[lane] = (lane ? x[1] : x[0]) + dl;
        scn[2 * lane] = fma(sc_s, cs, sc_c * sd);
        scn[2 * lane + 1] = fma(sc_c, cs, -sc_s * sd);
      } else if (lane < 6) {  // EigenQuaternionManifold::Plus, component k of (x, y, z, w)
        const int k = lane - 2;
        //   x' = cq x + dx w + dy z - dz y     y' = cq y - dx z + dy w + dz x
        //   z' = cq z + dx y - dy x + dz w     w' = cq w - dx x - dy y - dz z
        const double dx = sinc * delta[2], dy = sinc * delta[3], dz = sinc * delta[4];
        const double sx = (k == 1 || k == 3) ? -dx : dx, sy = (k == 2 || k == 3) ? -dy : dy, sz = (k == 0 || k == 3) ? -dz : dz;
        cnd[2 + k] = fma(sz, q4[3], fma(sy, q4[2], fma(sx, q4[1], cs * q4[0])));
      }
    } else if (lane == 0) {
      double cand[6], scc[4];
      cand[0] = x[0] + delta[0];
      cand[1] = x[1] + delta[1];
      lm_candidate_large_step(st.pts[nxi], st.scs[nxi], delta, cand, scc);
#pragma unroll
      for (int i = 0; i < 6; ++i) cnd[i] = cand[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) scn[i] = scc[i];
    }
    __syncwarp();
    // If the quadratic model already predicts |cost change| <= function_tolerance * cost the
    // iteration will almost surely terminate there: evaluate the cost alone first.
    pass_mode = (mcc <= 1.05 * o.function_tolerance * x_cost) ? kPassCost : kPassFull;
    // The candidate is complete: pose constants, and the caller may hand it to the evaluation already
    // (solve_slots_kernel does; what follows is off the pair's critical path).
    pose_const_lanes(cnd + 2, scn, s_pc, lane);
    post(pass_mode);
    LM_PROBE(3);  // candidate
    // ParameterToleranceReached depends on the step only.  Ceres evaluates the candidate's
    // cost first and then returns without applying the step, so the evaluation cannot change
    // the outcome: decide here and skip it (a caller that posted the candidate above discards its sums).
    const double e0 = x[0] - cnd[0], e1 = x[1] - cnd[1], e2 = x[2] - cnd[2];
    const double e3 = x[3] - cnd[3], e4 = x[4] - cnd[4], e5 = x[5] - cnd[5];
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
  }
  if (lane < 5) {
    if (FIRST) {
      st.scale[lane] = scale_l;
      st.inv_scale2[lane] = inv_scale2_l;
    }
    st.diag[lane] = diag_l;
  }
  const double imcc = (mcc > 0.0) ? fast_rcp(mcc) : 0.0;
  if (lane == 0) {
    st.xi = nxi;
    st.ti = nti;
    st.x_cost = x_cost;
    if (FIRST) st.initial_cost = x_cost;
    st.inv_radius = inv_radius;
    st.decrease_factor = decrease_factor;
    st.inv_model_cost_change = imcc;
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
