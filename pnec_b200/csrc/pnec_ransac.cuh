// pnec_ransac.cuh — RANSAC over the NEC eigensolver:
//   PNEC::Eigensolver with use_ransac_ == true, src/rel_pose_estimation/pnec.cc:239-272, i.e.
//   opengv::sac::Ransac<opengv::sac_problems::relative_pose::EigensolverSacProblem>::computeModel
//   (threshold 1e-6, sample size Options::ransac_sample_size_, at most Options::max_ransac_iterations_)
//   and the inlier extraction that follows it (pnec.cc:210-229).
// opengv is not part of the reference tree; the algorithm is restated in oracle/pnec_oracle_frame.c
// (RANSAC section), which this file follows in its `sequential = 0` mode: every hypothesis is a
// function of (seed, pair, iteration) alone, so a CTA evaluates a ROUND of hypotheses in parallel —
// sample, 36 moment sums, Levenberg-Marquardt on four lanes (es_lm_group), model, reprojection score
// of every correspondence — and then replays opengv's sequential bookkeeping (first strict maximum of
// the inlier count, adaptive k = log(1 - p) / log(1 - w^s), stop at iterations >= k) over the round in
// order.  The outcome equals the sequential loop's; hypotheses of a round beyond the stopping point
// are wasted work, which is why rounds start small.
//
// Pass 1 (ransac_kernel): one CTA per frame pair, NW warps, 8 NW hypotheses per full round; clean data
// stops after one or two rounds.  A pair that is not finished after `defer_after` iterations (many
// outliers, or a degenerate pair that never collects inliers and runs to max_iterations) is handed
// to pass 2, which spreads the remaining hypotheses of such pairs over the WHOLE device in
// super-rounds (ransac_hyp_kernel: blocks of 32 hypotheses as work items of a persistent grid, as many
// as the current k asks for; ransac_plan_kernel: the bookkeeping replayed in order, and the plan of
// the next super-round), so a 5000-iteration pair costs a few super-rounds instead of 5000 / 8 rounds
// of one CTA.  ransac_final_kernel then rebuilds the winning hypothesis and extracts the inliers.
// Everything a hypothesis computes lives in ONE out-of-line function (ransac_hypothesis) and the score
// is spelled out with explicit roundings, so all kernels produce the same bits for the same hypothesis.
//
// After the loop the winning model selects the inliers, which are written out compacted (bearing
// vectors, covariances, indices) at the pair's own offset: the stages that follow run on those arrays
// with a per-pair count.
#pragma once

#include "pnec_eigensolver.cuh"

namespace pnec {

constexpr int kRansacMaxSample = 32;
constexpr int kRansacSuper = 2048;   // hypotheses per pair and super-round of pass 2 (1024: 8 % slower; 4096: no faster)
constexpr int kRansacBlock = 8;      // hypotheses per work item of pass 2 (one warp: eight 4-lane groups)

// what pass 1 hands to pass 2 for an unfinished pair, and what pass 2 updates
struct RansacPairState {
  double best[16];     // winning hypothesis so far: R (9, row-major), t (3), q (4) -- valid if best_in_state
  double k;            // opengv's k
  int best_count, iters, done, best_h, best_in_state, pad;
};

struct RansacArgs {
  BatchView bv;              // f1, f2 (+ ct when out_ct), poses: start rotation (R12 of the adapter)
  double *best_poses;        // [B][7] winning hypothesis: unit quaternion + signed unit translation
                             //        (the start pose when there is no model or no inlier)
  int *num_inliers;          // [B]
  int *iterations;           // [B] opengv's iterations_
  int *inlier_index;         // [total] or nullptr: ascending indices (within the pair) of the inliers
  double *out_f1, *out_f2;   // [total][3] inliers of pair b at its own offset (InlierExtraction)
  double *out_ct;            // [total][9] or nullptr
  int max_iterations, sample_size;
  double threshold, probability, max_variation;
  unsigned long long seed;
  long long pair_index_base; // pair b draws from the stream of pair_index_base + b
  EsLmParams lm;
  // two passes
  int defer_after;           // pass 1 hands over pairs unfinished after this many iterations (0: never)
  RansacPairState *state;    // [B]
  int *defer;                // [0] number of deferred pairs, [1] work cursor, [2] work items of the current
                             // super-round, [3] cursor of ransac_final_kernel; the list follows at defer + 4 [B]
  int *blk_prefix;           // [B + 1] exclusive prefix of the work items per deferred slot
  int *hyp_count;            // [B][kRansacSuper] inlier counts of the current super-round
  // pass 1 as three kernels per round (ransac_pre_kernel / ransac_lm_kernel / ransac_post_kernel)
  double *sp_mom;            // [B][8][36] moment sums of the round's samples
  double *sp_x;              // [B][8][3]  Cayley start of the round's hypotheses, overwritten by the LM result
  int *sp_i0;                // [B][8]     first correspondence of the sample (sign of the translation)
  int *sp_active;            // [B][8]     hypothesis takes part in this round
  int *sp_live;              // [B]        pair still in pass 1
  int sp_round;              // round r evaluates hypotheses 8 r .. 8 r + 7
};

__device__ __forceinline__ unsigned long long rs_mix(unsigned long long z) {  // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
// 31 random bits for (seed, pair, iteration, draw), as in the oracle
__device__ __forceinline__ unsigned rs_u31(unsigned long long seed, unsigned long long pair, unsigned long long iteration,
                                           unsigned long long draw) {
  const unsigned long long h = rs_mix(rs_mix(rs_mix(seed ^ 0x51ed270b7a2f3c15ULL) + pair) + (iteration << 8) + draw);
  return static_cast<unsigned>(h >> 33);
}

// M(c) = sum n n^T of opengv's composeM from the 36 moments (every lane the whole matrix)
__device__ __forceinline__ void es_compose_m(const double *mom, int stride, const double c[3], double M[6]) {
  const double x = c[0], y = c[1], z = c[2];
  const double R[3][3] = {{1 + x * x - y * y - z * z, 2 * (x * y - z), 2 * (x * z + y)},
                          {2 * (x * y + z), 1 - x * x + y * y - z * z, 2 * (y * z - x)},
                          {2 * (x * z - y), 2 * (y * z + x), 1 - x * x - y * y + z * z}};
  double A[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[i][j] = 0.5 * R[i][j];
#pragma unroll
  for (int k = 0; k < 6; ++k) M[k] = 0.0;
  es_term<0>(mom, stride, R, A, M);
  es_term<1>(mom, stride, R, A, M);
  es_term<2>(mom, stride, R, A, M);
  es_term<3>(mom, stride, R, A, M);
  es_term<4>(mom, stride, R, A, M);
  es_term<5>(mom, stride, R, A, M);
}

// rotation matrix (row-major) of a unit quaternion (x, y, z, w)
__device__ __forceinline__ void quat_rotation(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1.0 - 2.0 * (y * y + z * z); R[1] = 2.0 * (x * y - z * w);       R[2] = 2.0 * (x * z + y * w);
  R[3] = 2.0 * (x * y + z * w);       R[4] = 1.0 - 2.0 * (x * x + z * z); R[5] = 2.0 * (y * z - x * w);
  R[6] = 2.0 * (x * z - y * w);       R[7] = 2.0 * (y * z + x * w);       R[8] = 1.0 - 2.0 * (x * x + y * y);
}

// a . b with the rounding sequence spelled out (see ransac_score)
__device__ __forceinline__ double rdot(double a0, double a1, double a2, double b0, double b1, double b2) {
  return fma(a2, b2, fma(a1, b1, __dmul_rn(a0, b0)));
}

// EigensolverSacProblem::getSelectedDistancesToModel for one correspondence: midpoint triangulation
// (opengv::triangulation::triangulate2) and the two 1 - cos reprojection errors.  m: R (9) then t (3).
// Explicit fma / __dmul_rn / __dadd_rn throughout: the inlier decision of a correspondence must not
// depend on which kernel (pass 1, pass 2, final selection) evaluates it.
__device__ __forceinline__ double ransac_score(const double *m, const double f1[3], const double f2[3]) {
  const double g0 = rdot(m[0], m[1], m[2], f2[0], f2[1], f2[2]);
  const double g1 = rdot(m[3], m[4], m[5], f2[0], f2[1], f2[2]);
  const double g2 = rdot(m[6], m[7], m[8], f2[0], f2[1], f2[2]);
  const double t0 = m[9], t1 = m[10], t2 = m[11];
  const double b0 = rdot(t0, t1, t2, f1[0], f1[1], f1[2]), b1 = rdot(t0, t1, t2, g0, g1, g2);
  const double a00 = rdot(f1[0], f1[1], f1[2], f1[0], f1[1], f1[2]);
  const double a10 = rdot(f1[0], f1[1], f1[2], g0, g1, g2);
  const double a11 = -rdot(g0, g1, g2, g0, g1, g2);
  // A = [[a00, -a10], [a10, a11]], lambda = A^-1 b
  const double idet = fast_rcp(fma(a00, a11, __dmul_rn(a10, a10)));
  const double l0 = __dmul_rn(fma(a11, b0, __dmul_rn(a10, b1)), idet);
  const double l1 = __dmul_rn(fma(a00, b1, -__dmul_rn(a10, b0)), idet);
  const double p0 = __dmul_rn(0.5, fma(l1, g0, fma(l0, f1[0], t0)));
  const double p1 = __dmul_rn(0.5, fma(l1, g1, fma(l0, f1[1], t1)));
  const double p2 = __dmul_rn(0.5, fma(l1, g2, fma(l0, f1[2], t2)));
  const double d0 = __dadd_rn(p0, -t0), d1 = __dadd_rn(p1, -t1), d2 = __dadd_rn(p2, -t2);
  const double q0 = rdot(m[0], m[3], m[6], d0, d1, d2);  // R^T (p - t)
  const double q1 = rdot(m[1], m[4], m[7], d0, d1, d2);
  const double q2 = rdot(m[2], m[5], m[8], d0, d1, d2);
  const double e1 = fma(-rdot(f1[0], f1[1], f1[2], p0, p1, p2), rsqrt(rdot(p0, p1, p2, p0, p1, p2)), 1.0);
  const double e2 = fma(-rdot(f2[0], f2[1], f2[2], q0, q1, q2), rsqrt(rdot(q0, q1, q2, q0, q1, q2)), 1.0);
  return __dadd_rn(e1, e2);
}

// Everything one hypothesis computes (getSamples + computeModelCoefficients of iteration h), by the
// 4-lane group `g` of its warp; ALL 32 lanes of a warp call it together (the Levenberg-Marquardt
// loop votes across the warp), groups without work pass active = false.  Out of line: ONE body for
// every kernel, so that a hypothesis is the same bits wherever it is evaluated.  Shared-memory
// scratch of the calling CTA: s_mom [36][NG], s_front / s_bpos / s_bval [kRansacMaxSample][NG];
// the model goes to model16 = R (9), t (3), q (4) (written by lane 0 of the group).
// WIDE: es_lm_group's wide turns (same bits either way) -- on where a warp is left with one or two
// hypotheses and latency counts (one CTA per pair, the rebuild of the winner), off in the
// device-filling pass 2, which is bound by instruction fetch.
template <bool WIDE>
__device__ __noinline__ void ransac_hypothesis(const double *f1, const double *f2, int n, int ns, double c0x, double c0y,
                                               double c0z, unsigned long long seed, unsigned long long pair, int h,
                                               double max_variation, const EsLmParams *lm, bool active, int g,
                                               int NG, double *s_mom, int *s_front, int *s_bpos, int *s_bval,
                                               double *model16) {
  const int lane = threadIdx.x & 31, sub = lane & 3;
  double x[3] = {c0x, c0y, c0z};
  if (active) {
    if (sub == 0) {
      // drawIndexSample from the identity permutation: front = positions 0 .. ns-1, the touched
      // positions beyond them in a short list
      int nb = 0;
      for (int i = 0; i < ns; ++i) s_front[i * NG + g] = i;
      for (int i = 0; i < ns; ++i) {
        const int j = i + static_cast<int>(rs_u31(seed, pair, h, i) % static_cast<unsigned>(n - i));
        const int vi = s_front[i * NG + g];
        if (j < ns) {
          s_front[i * NG + g] = s_front[j * NG + g];
          s_front[j * NG + g] = vi;
        } else {
          int k = 0;
          while (k < nb && s_bpos[k * NG + g] != j) ++k;
          if (k == nb) { s_bpos[k * NG + g] = j; s_bval[k * NG + g] = j; ++nb; }
          s_front[i * NG + g] = s_bval[k * NG + g];
          s_bval[k * NG + g] = vi;
        }
      }
    }
    __syncwarp(0xfu << (lane & ~3));
    // the 36 moment sums over the sample, in sample order: lane `sub` owns the rows p = sub and
    // p = sub + 4 of sym(f1 f1^T) (x) sym(f2 f2^T)
    double acc0[6] = {0, 0, 0, 0, 0, 0}, acc1[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < ns; ++i) {
      const int idx = s_front[i * NG + g];
      const double a[3] = {f1[3 * idx], f1[3 * idx + 1], f1[3 * idx + 2]};
      const double c[3] = {f2[3 * idx], f2[3 * idx + 1], f2[3 * idx + 2]};
      const double A[6] = {a[0] * a[0], a[0] * a[1], a[0] * a[2], a[1] * a[1], a[1] * a[2], a[2] * a[2]};
      const double F[6] = {c[0] * c[0], c[0] * c[1], c[0] * c[2], c[1] * c[1], c[1] * c[2], c[2] * c[2]};
      const double A0 = sub == 0 ? A[0] : sub == 1 ? A[1] : sub == 2 ? A[2] : A[3];
      const double A1 = sub == 0 ? A[4] : A[5];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        acc0[q] = fma(A0, F[q], acc0[q]);
        acc1[q] = fma(A1, F[q], acc1[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      s_mom[(6 * sub + q) * NG + g] = acc0[q];
      if (sub < 2) s_mom[(6 * (sub + 4) + q) * NG + g] = acc1[q];
    }
    // computeModelCoefficients: "randomize the starting point a bit"
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double u = static_cast<double>(rs_u31(seed, pair, h, ns + d)) / 2147483647.0;
      x[d] += (u - 0.5) * 2.0 * max_variation;
    }
  }
  __syncwarp();
  {
    int info, nfev;
    es_lm_group<WIDE>(s_mom + g, NG, 1, *lm, active, sub, x, info, nfev);
  }
  if (active) {
    // eigensolver_main's tail: rotation = cayley2rot(x), translation along the eigenvector of the
    // smallest eigenvalue of M(x), towards the optical flow of the sample's first correspondence
    double M[6], t[3], lam, q[4], R[9];
    es_compose_m(s_mom + g, NG, x, M);
    sym3_smallest_eigvec(M, t, lam);
    const double sc = 1.0 / sqrt(1.0 + x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    q[0] = x[0] * sc; q[1] = x[1] * sc; q[2] = x[2] * sc; q[3] = sc;
    quat_rotation(q, R);
    const int i0 = s_front[g];
    const double a[3] = {f1[3 * i0], f1[3 * i0 + 1], f1[3 * i0 + 2]};
    const double c[3] = {f2[3 * i0], f2[3 * i0 + 1], f2[3 * i0 + 2]};
    double gg[3];
    rot(R, c, gg);
    const double flow = (a[0] - gg[0]) * t[0] + (a[1] - gg[1]) * t[1] + (a[2] - gg[2]) * t[2];
    if (flow < 0.0) { t[0] = -t[0]; t[1] = -t[1]; t[2] = -t[2]; }
    if (sub == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) model16[k] = R[k];
      model16[9] = t[0]; model16[10] = t[1]; model16[11] = t[2];
      model16[12] = q[0]; model16[13] = q[1]; model16[14] = q[2]; model16[15] = q[3];
    }
  }
}

// opengv's k after a new best model with `count` inliers of n
__device__ __forceinline__ double ransac_k(int count, int n, int ns, double probability) {
  const double w = static_cast<double>(count) / static_cast<double>(n);
  double p_no = 1.0 - pow(w, static_cast<double>(ns));
  p_no = fmax(DBL_EPSILON, p_no);
  p_no = fmin(1.0 - DBL_EPSILON, p_no);
  return log(1.0 - probability) / log(p_no);
}

// countWithinDistance of NH models (shared memory, 16 doubles apart) over the pair's correspondences
// by the whole CTA; per-warp counts to s_count[warp * NH + k] (ballots make them warp-uniform).
template <int NW, int NH>
__device__ __forceinline__ void ransac_count(const double *f1, const double *f2, int n, const double *s_model, int nh,
                                             double threshold, int *s_count, int warp) {
  constexpr int NT = NW * 32;
  const int lane = threadIdx.x & 31;
  int cnt[NH];
#pragma unroll
  for (int k = 0; k < NH; ++k) cnt[k] = 0;
  for (int i0 = warp * 32; i0 < n; i0 += NT) {
    const int i = i0 + lane;
    const bool valid = i < n;
    const int ii = valid ? i : n - 1;
    const double a[3] = {f1[3 * ii], f1[3 * ii + 1], f1[3 * ii + 2]};
    const double c[3] = {f2[3 * ii], f2[3 * ii + 1], f2[3 * ii + 2]};
#pragma unroll
    for (int k = 0; k < NH; ++k) {
      if (k < nh) {  // warp-uniform
        const bool in = valid && ransac_score(s_model + 16 * k, a, c) < threshold;
        cnt[k] += __popc(__ballot_sync(0xffffffffu, in));
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NH; ++k) s_count[warp * NH + k] = cnt[k];
  }
}

// selectWithinDistance(best model) + InlierExtraction (pnec.cc:210-229) by the whole CTA; writes the
// pair's outputs.  s_best: the winning model in shared memory.
template <int NW>
__device__ __forceinline__ void ransac_select(const RansacArgs &args, long long b, long long s, int n, const double *f1,
                                              const double *f2, const double *s_best, int iterations, int *s_wcnt) {
  constexpr int NT = NW * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int running = 0;
  for (int i0 = 0; i0 < n; i0 += NT) {
    const int i = i0 + tid;
    bool in = false;
    double a[3], c[3];
    if (i < n) {
      a[0] = f1[3 * i]; a[1] = f1[3 * i + 1]; a[2] = f1[3 * i + 2];
      c[0] = f2[3 * i]; c[1] = f2[3 * i + 1]; c[2] = f2[3 * i + 2];
      in = ransac_score(s_best, a, c) < args.threshold;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int pos = running + __popc(bal & ((1u << lane) - 1u));
    int tile = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      if (w < warp) pos += s_wcnt[w];
      tile += s_wcnt[w];
    }
    if (in) {
      const long long o = s + pos;
#pragma unroll
      for (int k = 0; k < 3; ++k) { args.out_f1[3 * o + k] = a[k]; args.out_f2[3 * o + k] = c[k]; }
      if (args.out_ct) {
#pragma unroll
        for (int k = 0; k < 9; ++k) args.out_ct[9 * o + k] = args.bv.ct[9 * (s + i) + k];
      }
      if (args.inlier_index) args.inlier_index[o] = i;
    }
    running += tile;
    __syncthreads();
  }
  if (tid == 0) {
    const double *pose = args.bv.poses + 7 * b;
    double *bp = args.best_poses + 7 * b;
    if (running > 0) {
      bp[0] = s_best[12]; bp[1] = s_best[13]; bp[2] = s_best[14]; bp[3] = s_best[15];
      bp[4] = s_best[9]; bp[5] = s_best[10]; bp[6] = s_best[11];
    } else {  // a best model without inliers: undefined in the reference; the start pose here (as the oracle)
      const double qn = 1.0 / sqrt(pose[0] * pose[0] + pose[1] * pose[1] + pose[2] * pose[2] + pose[3] * pose[3]);
      bp[0] = pose[0] * qn; bp[1] = pose[1] * qn; bp[2] = pose[2] * qn; bp[3] = pose[3] * qn;
      bp[4] = pose[4]; bp[5] = pose[5]; bp[6] = pose[6];
    }
    args.num_inliers[b] = running;
    args.iterations[b] = iterations;
  }
}

// ------------------------------------------------------------------ pass 1: one CTA per pair

template <int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) ransac_kernel(const __grid_constant__ RansacArgs args) {
  constexpr int NG = NW * 8;  // 4-lane groups = hypotheses per full round
  __shared__ double s_mom[kEsMom * NG];     // [k][g]
  __shared__ double s_model[NG * 16];       // R (9), t (3), q (4)
  __shared__ double s_best[16];
  __shared__ int s_front[kRansacMaxSample * NG], s_bpos[kRansacMaxSample * NG], s_bval[kRansacMaxSample * NG];
  __shared__ int s_count[NW * NG];
  __shared__ int s_state[5];                // best count, iterations, done, hypotheses of the next round, best h
  __shared__ int s_wcnt[NW];
  __shared__ double s_k;
  const int tid = threadIdx.x, g = tid >> 2;
  const long long b = blockIdx.x;
  long long s, e;
  problem_range(args.bv, b, s, e);
  const int n = static_cast<int>(e - s);
  const int ns = args.sample_size;
  const double *f1 = args.bv.f1 + 3 * s, *f2 = args.bv.f2 + 3 * s;
  const double *pose = args.bv.poses + 7 * b;
  const unsigned long long pair = static_cast<unsigned long long>(args.pair_index_base + b);
  if (n < ns || ns < 1) {  // getSamples: "Can not select %zu unique points out of %zu": no model
    if (tid == 0) {
      // Sophus::SE3d holds a unit quaternion
      const double qn = 1.0 / sqrt(pose[0] * pose[0] + pose[1] * pose[1] + pose[2] * pose[2] + pose[3] * pose[3]);
      double *bp = args.best_poses + 7 * b;
      bp[0] = pose[0] * qn; bp[1] = pose[1] * qn; bp[2] = pose[2] * qn; bp[3] = pose[3] * qn;
      bp[4] = pose[4]; bp[5] = pose[5]; bp[6] = pose[6];
      args.num_inliers[b] = 0;
      args.iterations[b] = 0;
    }
    return;
  }
  // opengv::math::rot2cayley of the start rotation
  const double c0[3] = {pose[0] / pose[3], pose[1] / pose[3], pose[2] / pose[3]};
  if (tid == 0) {
    s_state[0] = -1;
    s_state[1] = 0;
    s_state[2] = 0;
    s_state[3] = NG < 8 ? NG : 8;  // first round: 8 hypotheses (clean data stops after a handful)
    s_state[4] = -1;
    s_k = 1.0;
  }
  __syncthreads();

  for (int base = 0;;) {
    const int round = s_state[3];
    // hypothesis j of the round goes to warp j % NW: a short round leaves every warp with one or two
    // hypotheses, which es_lm_group then runs with wide turns (the first round of 8 over 4 warps)
    const int j = (g & 7) * NW + (g >> 3);
    const int h = base + j;  // == opengv's iterations_ for this hypothesis
    const bool active = j < round && h <= args.max_iterations;
    ransac_hypothesis<true>(f1, f2, n, ns, c0[0], c0[1], c0[2], args.seed, pair, h, args.max_variation, &args.lm, active, g,
                            NG, s_mom, s_front, s_bpos, s_bval, s_model + 16 * j);
    __syncthreads();
    ransac_count<NW, NG>(f1, f2, n, s_model, min(round, args.max_iterations - base + 1), args.threshold, s_count,
                         tid >> 5);
    __syncthreads();
    // computeModel's bookkeeping, in order
    if (tid == 0) {
      int best = s_state[0], iters = s_state[1], done = 0;
      double k = s_k;
      for (int j = 0; j < round; ++j) {
        if (!(static_cast<double>(iters) < k)) { done = 1; break; }
        int count = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) count += s_count[w * NG + j];
        if (count > best) {
          best = count;
          s_state[4] = base + j;
#pragma unroll
          for (int m = 0; m < 16; ++m) s_best[m] = s_model[16 * j + m];
          k = ransac_k(count, n, ns, args.probability);
        }
        ++iters;
        if (iters > args.max_iterations) { done = 1; break; }
      }
      if (!done && !(static_cast<double>(iters) < k)) done = 1;
      s_state[0] = best; s_state[1] = iters; s_state[2] = done;
      // the next round: what the stopping rule still asks for (k only falls from here, so more would be
      // wasted, and a short round runs with wide turns), a full round while k is far away
      const double rem = k - static_cast<double>(iters);
      s_state[3] = !(rem < static_cast<double>(NG)) ? NG : max(1, static_cast<int>(ceil(rem)));
      s_k = k;
    }
    __syncthreads();
    if (s_state[2]) break;
    base += round;
    if (args.defer_after > 0 && s_state[1] >= args.defer_after) {
      // unfinished: the rest of this pair's hypotheses are spread over the device by pass 2
      if (tid == 0) {
        RansacPairState &st = args.state[b];
#pragma unroll
        for (int m = 0; m < 16; ++m) st.best[m] = s_best[m];
        st.k = s_k;
        st.best_count = s_state[0];
        st.iters = s_state[1];
        st.done = 0;
        st.best_h = s_state[4];
        st.best_in_state = 1;
        args.defer[4 + atomicAdd(args.defer, 1)] = static_cast<int>(b);
      }
      return;
    }
  }
  ransac_select<NW>(args, b, s, n, f1, f2, s_best, s_state[1], s_wcnt);
}

// ------------------------------------------------------------------ pass 1, split
//
// ransac_kernel spends 65 % of its stall samples waiting for instructions (ncu: `no_inst`, L1.5 instruction
// bandwidth 96 % used, ICC hit rate 53 %): sampling, the unrolled Levenberg-Marquardt, the model and the
// scoring loop are one 126 KB body, and every warp of an SM is somewhere else in it.  es_lm_kernel -- the
// same LM, alone in its kernel -- does not have the problem (89 % ICC hits, 40 % of the bandwidth).  For
// large batches a round of pass 1 therefore runs as three kernels over all pairs still in it:
//   ransac_pre_kernel   sample, 36 moment sums, perturbed start          (one warp per pair, 8 hypotheses)
//   ransac_lm_kernel    es_lm_group on B x 8 virtual pairs               (as es_lm_kernel)
//   ransac_post_kernel  model, scores, bookkeeping, inlier extraction    (one warp per pair)
// with the pair's state (best model, count, k, iterations) in RansacPairState between the kernels.  The
// same device functions in the same order as ransac_hypothesis / ransac_kernel<1, .>; a pair unfinished
// after `defer_after` iterations goes to pass 2 exactly as before (its best model travels in the state,
// so no hypothesis of pass 1 is ever rebuilt by another kernel).

__global__ void __launch_bounds__(128) ransac_pre_kernel(const __grid_constant__ RansacArgs args) {
  constexpr int NG = 8;
  __shared__ int s_front[4][kRansacMaxSample * NG], s_bpos[4][kRansacMaxSample * NG], s_bval[4][kRansacMaxSample * NG];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, sub = lane & 3;
  const long long b = static_cast<long long>(blockIdx.x) * 4 + warp;
  if (b >= args.bv.num_problems) return;
  long long s, e;
  problem_range(args.bv, b, s, e);
  const int n = static_cast<int>(e - s);
  const int ns = args.sample_size;
  const double *f1 = args.bv.f1 + 3 * s, *f2 = args.bv.f2 + 3 * s;
  const double *pose = args.bv.poses + 7 * b;
  int live;
  if (args.sp_round == 0) {
    live = !(n < ns || ns < 1);
    if (lane == 0) {
      RansacPairState &st = args.state[b];
      st.k = 1.0;
      st.best_count = -1;
      st.iters = 0;
      st.done = 0;
      st.best_h = -1;
      st.best_in_state = 1;
      args.sp_live[b] = live;
      if (!live) {  // getSamples: "Can not select %zu unique points out of %zu": no model
        const double qn = 1.0 / sqrt(pose[0] * pose[0] + pose[1] * pose[1] + pose[2] * pose[2] + pose[3] * pose[3]);
        double *bp = args.best_poses + 7 * b;
        bp[0] = pose[0] * qn; bp[1] = pose[1] * qn; bp[2] = pose[2] * qn; bp[3] = pose[3] * qn;
        bp[4] = pose[4]; bp[5] = pose[5]; bp[6] = pose[6];
        args.num_inliers[b] = 0;
        args.iterations[b] = 0;
      }
    }
  } else {
    live = args.sp_live[b];
  }
  const int h = args.sp_round * NG + g;  // == opengv's iterations_ for this hypothesis
  const bool active = live && h <= args.max_iterations;
  const long long vb = b * NG + g;
  if (sub == 0) args.sp_active[vb] = active ? 1 : 0;
  if (!active) return;
  const unsigned long long pair = static_cast<unsigned long long>(args.pair_index_base + b);
  int *front = s_front[warp], *bpos = s_bpos[warp], *bval = s_bval[warp];
  if (sub == 0) {
    // drawIndexSample from the identity permutation (as in ransac_hypothesis)
    int nb = 0;
    for (int i = 0; i < ns; ++i) front[i * NG + g] = i;
    for (int i = 0; i < ns; ++i) {
      const int j = i + static_cast<int>(rs_u31(args.seed, pair, h, i) % static_cast<unsigned>(n - i));
      const int vi = front[i * NG + g];
      if (j < ns) {
        front[i * NG + g] = front[j * NG + g];
        front[j * NG + g] = vi;
      } else {
        int k = 0;
        while (k < nb && bpos[k * NG + g] != j) ++k;
        if (k == nb) { bpos[k * NG + g] = j; bval[k * NG + g] = j; ++nb; }
        front[i * NG + g] = bval[k * NG + g];
        bval[k * NG + g] = vi;
      }
    }
  }
  __syncwarp(0xfu << (lane & ~3));
  double acc0[6] = {0, 0, 0, 0, 0, 0}, acc1[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < ns; ++i) {
    const int idx = front[i * NG + g];
    const double a[3] = {f1[3 * idx], f1[3 * idx + 1], f1[3 * idx + 2]};
    const double c[3] = {f2[3 * idx], f2[3 * idx + 1], f2[3 * idx + 2]};
    const double A[6] = {a[0] * a[0], a[0] * a[1], a[0] * a[2], a[1] * a[1], a[1] * a[2], a[2] * a[2]};
    const double F[6] = {c[0] * c[0], c[0] * c[1], c[0] * c[2], c[1] * c[1], c[1] * c[2], c[2] * c[2]};
    const double A0 = sub == 0 ? A[0] : sub == 1 ? A[1] : sub == 2 ? A[2] : A[3];
    const double A1 = sub == 0 ? A[4] : A[5];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      acc0[q] = fma(A0, F[q], acc0[q]);
      acc1[q] = fma(A1, F[q], acc1[q]);
    }
  }
  double *mom = args.sp_mom + vb * kEsMom;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    mom[6 * sub + q] = acc0[q];
    if (sub < 2) mom[6 * (sub + 4) + q] = acc1[q];
  }
  if (sub == 0) {
    // opengv::math::rot2cayley of the start rotation, then computeModelCoefficients: "randomize the starting point a bit"
    double x[3] = {pose[0] / pose[3], pose[1] / pose[3], pose[2] / pose[3]};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double u = static_cast<double>(rs_u31(args.seed, pair, h, ns + d)) / 2147483647.0;
      x[d] += (u - 0.5) * 2.0 * args.max_variation;
      args.sp_x[3 * vb + d] = x[d];
    }
    args.sp_i0[vb] = front[g];
  }
}

// es_lm_kernel on the round's hypotheses: Cayley start in, Cayley result out (sp_x), moments at sp_mom.
__global__ void __launch_bounds__(kEsLmThreads) ransac_lm_kernel(const __grid_constant__ RansacArgs args) {
  __shared__ double s_mom[kEsMom * kEsLmPairs];
  const int tid = threadIdx.x, sub = tid & 3, slot = tid >> 2;
  const long long nvirt = args.bv.num_problems * 8;
  const long long first = static_cast<long long>(blockIdx.x) * kEsLmPairs;
  const long long vb = first + slot;
  const bool active = vb < nvirt && args.sp_active[vb] != 0;
  if (!__syncthreads_or(active)) return;
  {
    const long long cnt = min(static_cast<long long>(kEsLmPairs), nvirt - first);
    for (long long i = tid; i < kEsLmPairs * kEsMom; i += kEsLmThreads) {
      const int p = static_cast<int>(i / kEsMom), k = static_cast<int>(i % kEsMom);
      s_mom[k * kEsLmPairs + p] = i < cnt * kEsMom ? args.sp_mom[first * kEsMom + i] : 0.0;
    }
    __syncthreads();
  }
  const long long vv = active ? vb : first;
  double x[3] = {args.sp_x[3 * vv], args.sp_x[3 * vv + 1], args.sp_x[3 * vv + 2]};
  int info = 0, nfev = 0;
  es_lm_group<false>(s_mom + slot, kEsLmPairs, 1, args.lm, active, sub, x, info, nfev);
  if (active && sub == 0) {
    args.sp_x[3 * vb] = x[0];
    args.sp_x[3 * vb + 1] = x[1];
    args.sp_x[3 * vb + 2] = x[2];
  }
}

__global__ void __launch_bounds__(32, 16) ransac_post_kernel(const __grid_constant__ RansacArgs args) {
  constexpr int NG = 8;
  __shared__ double s_model[NG * 16];
  __shared__ double s_best[16];
  __shared__ int s_count[NG];
  __shared__ int s_wcnt[1];
  __shared__ int s_flag, s_iters;
  const int lane = threadIdx.x, g = lane >> 2, sub = lane & 3;
  const long long b = blockIdx.x;
  if (!args.sp_live[b]) return;
  long long s, e;
  problem_range(args.bv, b, s, e);
  const int n = static_cast<int>(e - s);
  const int ns = args.sample_size;
  const double *f1 = args.bv.f1 + 3 * s, *f2 = args.bv.f2 + 3 * s;
  const long long vb = b * NG + g;
  const int base = args.sp_round * NG;
  if (args.sp_active[vb]) {
    // eigensolver_main's tail (as in ransac_hypothesis): rotation = cayley2rot(x), translation along the eigenvector
    // of the smallest eigenvalue of M(x), towards the optical flow of the sample's first correspondence
    const double x[3] = {args.sp_x[3 * vb], args.sp_x[3 * vb + 1], args.sp_x[3 * vb + 2]};
    double M[6], t[3], lam, q[4], R[9];
    es_compose_m(args.sp_mom + vb * kEsMom, 1, x, M);
    sym3_smallest_eigvec(M, t, lam);
    const double sc = 1.0 / sqrt(1.0 + x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    q[0] = x[0] * sc; q[1] = x[1] * sc; q[2] = x[2] * sc; q[3] = sc;
    quat_rotation(q, R);
    const int i0 = args.sp_i0[vb];
    const double a[3] = {f1[3 * i0], f1[3 * i0 + 1], f1[3 * i0 + 2]};
    const double c[3] = {f2[3 * i0], f2[3 * i0 + 1], f2[3 * i0 + 2]};
    double gg[3];
    rot(R, c, gg);
    const double flow = (a[0] - gg[0]) * t[0] + (a[1] - gg[1]) * t[1] + (a[2] - gg[2]) * t[2];
    if (flow < 0.0) { t[0] = -t[0]; t[1] = -t[1]; t[2] = -t[2]; }
    if (sub == 0) {
      double *m = s_model + 16 * g;
#pragma unroll
      for (int k = 0; k < 9; ++k) m[k] = R[k];
      m[9] = t[0]; m[10] = t[1]; m[11] = t[2];
      m[12] = q[0]; m[13] = q[1]; m[14] = q[2]; m[15] = q[3];
    }
  }
  __syncthreads();
  const int round = NG;
  ransac_count<1, NG>(f1, f2, n, s_model, min(round, args.max_iterations - base + 1), args.threshold, s_count, 0);
  __syncthreads();
  RansacPairState &st = args.state[b];
  if (lane == 0) {
    // computeModel's bookkeeping, in order (as in ransac_kernel)
    int best = st.best_count, iters = st.iters, done = 0;
    double k = st.k;
    for (int j = 0; j < round; ++j) {
      if (!(static_cast<double>(iters) < k)) { done = 1; break; }
      const int count = s_count[j];
      if (count > best) {
        best = count;
        st.best_h = base + j;
#pragma unroll
        for (int m = 0; m < 16; ++m) st.best[m] = s_model[16 * j + m];
        k = ransac_k(count, n, ns, args.probability);
      }
      ++iters;
      if (iters > args.max_iterations) { done = 1; break; }
    }
    if (!done && !(static_cast<double>(iters) < k)) done = 1;
    st.best_count = best;
    st.iters = iters;
    st.k = k;
    s_iters = iters;
    int flag = 0;
    if (done) {
      flag = 1;
    } else if (args.defer_after > 0 && iters >= args.defer_after) {
      // unfinished: the rest of this pair's hypotheses are spread over the device by pass 2
      st.done = 0;
      st.best_in_state = 1;
      args.defer[4 + atomicAdd(args.defer, 1)] = static_cast<int>(b);
      flag = 2;
    }
    if (flag) args.sp_live[b] = 0;
    s_flag = flag;
  }
  __syncthreads();
  if (s_flag == 1) {
    if (lane < 16) s_best[lane] = st.best[lane];
    __syncthreads();
    ransac_select<1>(args, b, s, n, f1, f2, s_best, s_iters, s_wcnt);
  }
}

// ------------------------------------------------------------------ pass 2

// Hypotheses a deferred pair gets in the next super-round, as work items of kRansacBlock (0 when done):
// what the current k still asks for, but never more than it has consumed so far (and at most
// kRansacSuper).  k is a poor predictor early on -- with outliers the first good model typically drops
// it from thousands to below a hundred -- so the speculation doubles instead of trusting it: the
// hypotheses computed stay within twice those the sequential loop consumes.
// (Granting twice or three times what a pair has consumed, always or only while k exceeds max_iterations,
// was measured: within noise on clean data, 4 % slower with 25 % outliers.)
__host__ __device__ __forceinline__ int ransac_grant(int iters, int max_iterations) {
  const int grow = iters > 4 * kRansacBlock ? iters : 4 * kRansacBlock;
  const int left = max_iterations + 1 - iters;
  const int cap = grow < kRansacSuper ? grow : kRansacSuper;
  return left < cap ? left : cap;
}
__device__ __forceinline__ int ransac_blocks_wanted(const RansacPairState &st, int max_iterations) {
  if (st.done) return 0;
  const double want_d = ceil(st.k) - static_cast<double>(st.iters);
  const int cap = ransac_grant(st.iters, max_iterations);
  int want = want_d > static_cast<double>(cap) ? cap : static_cast<int>(want_d);
  want = max(1, min(want, cap));
  return (want + kRansacBlock - 1) / kRansacBlock;
}

// After every ransac_hyp_kernel: one WARP per deferred pair replays computeModel's bookkeeping over the counts
// of the super-round that just ran, 32 counts per step.  The sequential loop
//     for j: if !(iters < k) stop;  if count[j] > best: new best, new k;  ++iters;  if iters > max stop
// only changes state at a strict new maximum, so a step finds those (prefix maximum over the lanes), walks the
// few of them in order, and consumes the hypotheses in between in one go (they are consumed while
// iters < ceil(k) and iters <= max): the same outcome as the loop, in 64 coalesced loads instead of 2048
// dependent ones (the one-thread-per-pair form took 0.13-0.24 ms per super-round on C2).
__global__ void __launch_bounds__(128) ransac_replay_kernel(const __grid_constant__ RansacArgs args) {
  const int lane = threadIdx.x & 31;
  const int slot = static_cast<int>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (slot >= args.defer[0]) return;
  const long long b = (args.defer + 4)[slot];
  RansacPairState &st = args.state[b];
  if (st.done) return;
  long long s, e;
  problem_range(args.bv, b, s, e);
  const int n = static_cast<int>(e - s);
  const int nh = (args.blk_prefix[slot + 1] - args.blk_prefix[slot]) * kRansacBlock;
  const int *cnt = args.hyp_count + static_cast<long long>(slot) * kRansacSuper;
  int best = st.best_count, iters = st.iters, done = 0, best_h = st.best_h, in_state = st.best_in_state;
  const int first = iters;
  double k = st.k;
  // hypotheses that may still be consumed from `iters` on under the current k and the iteration cap
  auto allowed = [&]() -> int {
    const double by_k = ceil(k) - static_cast<double>(iters);  // iters < k  <=>  iters < ceil(k)
    const int by_max = args.max_iterations + 1 - iters;          // the one that makes iters > max is still consumed
    if (!(by_k > 0.0)) return 0;
    return by_k < static_cast<double>(by_max) ? static_cast<int>(by_k) : by_max;
  };
  for (int j0 = 0; j0 < nh && !done; j0 += 32) {
    const int nvalid = min(32, nh - j0);
    const int c = lane < nvalid ? cnt[j0 + lane] : INT_MIN;
    int m = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, m, d);
      if (lane >= d) m = max(m, t);
    }
    int prev = __shfl_up_sync(0xffffffffu, m, 1);
    if (lane == 0) prev = INT_MIN;
    const unsigned events = __ballot_sync(0xffffffffu, lane < nvalid && c > max(prev, best));
    int pos = 0;
    while (pos < nvalid) {
      const unsigned rest = pos < 32 ? events >> pos : 0u;
      const int ev = rest ? pos + __ffs(rest) - 1 : nvalid;  // next strict maximum, or the end of the step
      const int want = ev - pos, can = allowed();
      if (can < want) {  // the loop stops inside the run of non-maxima
        iters += can;
        done = 1;
        break;
      }
      iters += want;
      if (iters > args.max_iterations) { done = 1; break; }
      pos = ev;
      if (pos >= nvalid) break;
      if (!(static_cast<double>(iters) < k)) { done = 1; break; }
      best = __shfl_sync(0xffffffffu, c, pos);
      best_h = first + j0 + pos;
      in_state = 0;
      k = ransac_k(best, n, args.sample_size, args.probability);
      ++iters;
      ++pos;
      if (iters > args.max_iterations) { done = 1; break; }
    }
  }
  if (!done && !(static_cast<double>(iters) < k)) done = 1;
  if (lane == 0) {
    st.best_count = best; st.iters = iters; st.done = done; st.best_h = best_h; st.best_in_state = in_state;
    st.k = k;
  }
}

// One CTA, after pass 1 and after every replay: plans the next super-round (work items per slot, their
// exclusive prefix, the total) and rewinds the work cursor.
__global__ void __launch_bounds__(1024) ransac_plan_kernel(const __grid_constant__ RansacArgs args) {
  __shared__ int s_scan[1024];
  __shared__ int s_running;
  const int tid = threadIdx.x;
  const int count = args.defer[0];
  const int *list = args.defer + 4;
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < count; base += 1024) {
    const int slot = base + tid;
    int want = 0;
    if (slot < count) want = ransac_blocks_wanted(args.state[list[slot]], args.max_iterations);
    // exclusive scan of `want` over this chunk of slots
    s_scan[tid] = want;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int v = tid >= o ? s_scan[tid - o] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    if (slot < count) args.blk_prefix[slot] = s_running + s_scan[tid] - want;
    __syncthreads();
    if (tid == 1023) s_running += s_scan[1023];
    __syncthreads();
  }
  if (tid == 0) {
    args.blk_prefix[count] = s_running;
    args.defer[1] = 0;
    args.defer[2] = s_running;
    args.defer[3] = 0;
  }
}

// Persistent grid of INDEPENDENT warps: a work item is one block of 8 hypotheses of one deferred pair
// (the eight 4-lane groups of a warp, as in pass 1 with one warp per pair), scored against all of the
// pair's correspondences by the same warp.  No CTA-wide barrier: a warp whose Levenberg-Marquardt runs
// long holds up nobody else.
template <int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB) ransac_hyp_kernel(const __grid_constant__ RansacArgs args) {
  constexpr int NG = 8;
  __shared__ double s_mom[WPC][kEsMom * NG];
  __shared__ double s_model[WPC][NG * 16];
  __shared__ int s_front[WPC][kRansacMaxSample * NG], s_bpos[WPC][kRansacMaxSample * NG], s_bval[WPC][kRansacMaxSample * NG];
  __shared__ int s_count[WPC][NG];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2;
  const int total = args.defer[2], count = args.defer[0];
  const int *list = args.defer + 4;
  // The first item of a warp is its own index, later ones come from the cursor: a super-round with fewer
  // items than the grid has warps (clean data: a few hundred) occupies the first CTAs only and the others
  // leave at once, instead of every CTA of the machine holding its registers and shared memory for one
  // busy warp (stage on C2 15.0 -> 14.4 ms).
  // (Tried on top: splitting a block of 8 hypotheses over four warps of two, run with es_lm_group's wide
  // turns, for super-rounds that do not fill the machine -- stage 14.3 -> 13.9 ms, but the default frame
  // solve 19.2 -> 20.5 ms: a second 100 KB instantiation of the hypothesis competes for the instruction
  // caches with every other kernel in flight.  What a small super-round costs is the cold pass of each
  // warp through that code, not the slowest Levenberg-Marquardt run.)
  const int nwarps = static_cast<int>(gridDim.x) * WPC;
  bool first_item = true;
  for (;;) {
    int w = static_cast<int>(blockIdx.x) * WPC + warp;
    if (!first_item) {
      if (lane == 0) w = nwarps + atomicAdd(args.defer + 1, 1);
      w = __shfl_sync(0xffffffffu, w, 0);
    }
    first_item = false;
    if (w >= total) return;
    // slot with blk_prefix[slot] <= w < blk_prefix[slot + 1]
    int lo = 0, hi = count;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (args.blk_prefix[mid] <= w) lo = mid; else hi = mid;
    }
    const int slot = lo, blk = w - args.blk_prefix[slot];
    const long long b = list[slot];
    long long s, e;
    problem_range(args.bv, b, s, e);
    const int n = static_cast<int>(e - s);
    const double *f1 = args.bv.f1 + 3 * s, *f2 = args.bv.f2 + 3 * s;
    const double *pose = args.bv.poses + 7 * b;
    const int first = args.state[b].iters + blk * kRansacBlock;
    const int h = first + g;
    const bool active = h <= args.max_iterations;
    ransac_hypothesis<false>(f1, f2, n, args.sample_size, pose[0] / pose[3], pose[1] / pose[3], pose[2] / pose[3], args.seed,
                      static_cast<unsigned long long>(args.pair_index_base + b), h, args.max_variation, &args.lm, active,
                      g, NG, s_mom[warp], s_front[warp], s_bpos[warp], s_bval[warp], s_model[warp] + 16 * g);
    __syncwarp();
    const int nh = min(NG, args.max_iterations - first + 1);
    ransac_count<1, NG>(f1, f2, n, s_model[warp], nh, args.threshold, s_count[warp], 0);
    __syncwarp();
    if (lane < NG)
      args.hyp_count[static_cast<long long>(slot) * kRansacSuper + blk * kRansacBlock + lane] = lane < nh ? s_count[warp][lane] : 0;
    __syncwarp();
  }
}

// Persistent grid over the deferred pairs: rebuild the winning hypothesis when pass 2 found it (the
// same out-of-line function: the same bits the count was made with), then the inlier extraction.
__global__ void __launch_bounds__(128) ransac_final_kernel(const __grid_constant__ RansacArgs args) {
  constexpr int NW = 4, NG = 32;
  __shared__ double s_mom[kEsMom * NG];
  __shared__ double s_best[16];
  __shared__ int s_front[kRansacMaxSample * NG], s_bpos[kRansacMaxSample * NG], s_bval[kRansacMaxSample * NG];
  __shared__ int s_wcnt[NW];
  __shared__ int s_work;
  const int tid = threadIdx.x, g = tid >> 2;
  const int count = args.defer[0];
  const int *list = args.defer + 4;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_work = atomicAdd(args.defer + 3, 1);
    __syncthreads();
    const int slot = s_work;
    if (slot >= count) return;
    const long long b = list[slot];
    const RansacPairState &st = args.state[b];
    long long s, e;
    problem_range(args.bv, b, s, e);
    const int n = static_cast<int>(e - s);
    const double *f1 = args.bv.f1 + 3 * s, *f2 = args.bv.f2 + 3 * s;
    const double *pose = args.bv.poses + 7 * b;
    if (st.best_in_state) {
      if (tid < 16) s_best[tid] = st.best[tid];
    } else if (tid < 32) {  // warp 0, group 0 recomputes hypothesis best_h
      ransac_hypothesis<true>(f1, f2, n, args.sample_size, pose[0] / pose[3], pose[1] / pose[3], pose[2] / pose[3], args.seed,
                        static_cast<unsigned long long>(args.pair_index_base + b), st.best_h, args.max_variation,
                        &args.lm, g == 0, g, NG, s_mom, s_front, s_bpos, s_bval, s_best);
    }
    __syncthreads();
    ransac_select<NW>(args, b, s, n, f1, f2, s_best, st.iters, s_wcnt);
  }
}

}  // namespace pnec
