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
// One CTA per frame pair, NW warps, 8 NW hypotheses per full round.  After the loop the winning model
// selects the inliers, which are written out compacted (bearing vectors, covariances, indices) at the
// pair's own offset: the stages that follow run on those arrays with a per-pair count.
#pragma once

#include "pnec_eigensolver.cuh"

namespace pnec {

constexpr int kRansacMaxSample = 32;

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

// EigensolverSacProblem::getSelectedDistancesToModel for one correspondence: midpoint triangulation
// (opengv::triangulation::triangulate2) and the two 1 - cos reprojection errors.  m: R (9) then t (3).
__device__ __forceinline__ double ransac_score(const double *m, const double f1[3], const double f2[3]) {
  const double g[3] = {m[0] * f2[0] + m[1] * f2[1] + m[2] * f2[2], m[3] * f2[0] + m[4] * f2[1] + m[5] * f2[2],
                       m[6] * f2[0] + m[7] * f2[1] + m[8] * f2[2]};
  const double t[3] = {m[9], m[10], m[11]};
  const double b0 = dot3(t, f1), b1 = dot3(t, g);
  const double a00 = dot3(f1, f1), a10 = dot3(f1, g), a01 = -a10, a11 = -dot3(g, g);
  const double idet = fast_rcp(a00 * a11 - a01 * a10);
  const double l0 = (a11 * b0 - a01 * b1) * idet, l1 = (-a10 * b0 + a00 * b1) * idet;
  double p[3], d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    p[k] = 0.5 * (l0 * f1[k] + t[k] + l1 * g[k]);
    d[k] = p[k] - t[k];
  }
  const double pp[3] = {m[0] * d[0] + m[3] * d[1] + m[6] * d[2], m[1] * d[0] + m[4] * d[1] + m[7] * d[2],
                        m[2] * d[0] + m[5] * d[1] + m[8] * d[2]};  // R^T (p - t)
  return (1.0 - dot3(f1, p) * rsqrt(dot3(p, p))) + (1.0 - dot3(f2, pp) * rsqrt(dot3(pp, pp)));
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) ransac_kernel(const __grid_constant__ RansacArgs args) {
  constexpr int NT = NW * 32, NG = NW * 8;  // threads, 4-lane groups = hypotheses per full round
  __shared__ double s_mom[kEsMom * NG];     // [k][g]
  __shared__ double s_model[NG][16];        // R (9), t (3), q (4)
  __shared__ double s_best[16];
  __shared__ int s_front[kRansacMaxSample][NG], s_bpos[kRansacMaxSample][NG], s_bval[kRansacMaxSample][NG];
  __shared__ int s_count[NW][NG];
  __shared__ int s_state[4];                // best count, iterations, done, hypotheses of the next round
  __shared__ double s_k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = tid & 3, g = tid >> 2;
  const long long b = blockIdx.x;
  long long s, e;
  problem_range(args.bv, b, s, e);
  const int n = static_cast<int>(e - s);
  const int ns = args.sample_size;
  const double *f1 = args.bv.f1 + 3 * s, *f2 = args.bv.f2 + 3 * s;
  const double *pose = args.bv.poses + 7 * b;
  const unsigned long long pair = static_cast<unsigned long long>(args.pair_index_base + b);
  // Sophus::SE3d holds a unit quaternion
  const double qn = 1.0 / sqrt(pose[0] * pose[0] + pose[1] * pose[1] + pose[2] * pose[2] + pose[3] * pose[3]);
  if (n < ns || ns < 1) {  // getSamples: "Can not select %zu unique points out of %zu": no model
    if (tid == 0) {
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
    s_k = 1.0;
  }
  __syncthreads();

  for (int base = 0;;) {
    const int round = s_state[3];
    // ---------------------------------------------------------------- hypotheses of this round
    const int h = base + g;                                  // == opengv's iterations_ for this hypothesis
    const bool active = g < round && h <= args.max_iterations;
    double x[3] = {c0[0], c0[1], c0[2]};
    if (active) {
      if (sub == 0) {
        // drawIndexSample from the identity permutation: front = positions 0 .. ns-1, the touched
        // positions beyond them in a short list
        int nb = 0;
        for (int i = 0; i < ns; ++i) s_front[i][g] = i;
        for (int i = 0; i < ns; ++i) {
          const int j = i + static_cast<int>(rs_u31(args.seed, pair, h, i) % static_cast<unsigned>(n - i));
          const int vi = s_front[i][g];
          if (j < ns) {
            s_front[i][g] = s_front[j][g];
            s_front[j][g] = vi;
          } else {
            int k = 0;
            while (k < nb && s_bpos[k][g] != j) ++k;
            if (k == nb) { s_bpos[k][g] = j; s_bval[k][g] = j; ++nb; }
            s_front[i][g] = s_bval[k][g];
            s_bval[k][g] = vi;
          }
        }
      }
      __syncwarp(0xfu << (lane & ~3));
      // the 36 moment sums over the sample, in sample order: lane `sub` owns the rows p = sub and
      // p = sub + 4 of sym(f1 f1^T) (x) sym(f2 f2^T)
      double acc0[6] = {0, 0, 0, 0, 0, 0}, acc1[6] = {0, 0, 0, 0, 0, 0};
      for (int i = 0; i < ns; ++i) {
        const int idx = s_front[i][g];
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
        const double u = static_cast<double>(rs_u31(args.seed, pair, h, ns + d)) / 2147483647.0;
        x[d] = c0[d] + (u - 0.5) * 2.0 * args.max_variation;
      }
    }
    __syncwarp();
    {
      int info, nfev;
      es_lm_group(s_mom + g, NG, args.lm, active, sub, x, info, nfev);
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
      const int i0 = s_front[0][g];
      const double a[3] = {f1[3 * i0], f1[3 * i0 + 1], f1[3 * i0 + 2]};
      const double c[3] = {f2[3 * i0], f2[3 * i0 + 1], f2[3 * i0 + 2]};
      double gg[3];
      rot(R, c, gg);
      const double flow = (a[0] - gg[0]) * t[0] + (a[1] - gg[1]) * t[1] + (a[2] - gg[2]) * t[2];
      if (flow < 0.0) { t[0] = -t[0]; t[1] = -t[1]; t[2] = -t[2]; }
      if (sub == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) s_model[g][k] = R[k];
        s_model[g][9] = t[0]; s_model[g][10] = t[1]; s_model[g][11] = t[2];
        s_model[g][12] = q[0]; s_model[g][13] = q[1]; s_model[g][14] = q[2]; s_model[g][15] = q[3];
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- countWithinDistance
    // every thread scores its correspondences against all hypotheses of the round; ballots make
    // the counts warp-uniform
    {
      int cnt[NG];
#pragma unroll
      for (int k = 0; k < NG; ++k) cnt[k] = 0;
      const int nh = min(round, args.max_iterations - base + 1);  // active hypotheses: 0 .. nh-1
      for (int i0 = warp * 32; i0 < n; i0 += NT) {
        const int i = i0 + lane;
        const bool valid = i < n;
        const int ii = valid ? i : n - 1;
        const double a[3] = {f1[3 * ii], f1[3 * ii + 1], f1[3 * ii + 2]};
        const double c[3] = {f2[3 * ii], f2[3 * ii + 1], f2[3 * ii + 2]};
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          if (k < nh) {  // warp-uniform
            const bool in = valid && ransac_score(s_model[k], a, c) < args.threshold;
            cnt[k] += __popc(__ballot_sync(0xffffffffu, in));
          }
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NG; ++k) s_count[warp][k] = cnt[k];
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- computeModel's bookkeeping, in order
    if (tid == 0) {
      int best = s_state[0], iters = s_state[1], done = 0;
      double k = s_k;
      for (int j = 0; j < round; ++j) {
        if (!(static_cast<double>(iters) < k)) { done = 1; break; }
        int count = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) count += s_count[w][j];
        if (count > best) {
          best = count;
#pragma unroll
          for (int m = 0; m < 16; ++m) s_best[m] = s_model[j][m];
          const double wfrac = static_cast<double>(count) / static_cast<double>(n);
          double p_no = 1.0 - pow(wfrac, static_cast<double>(ns));
          p_no = fmax(DBL_EPSILON, p_no);
          p_no = fmin(1.0 - DBL_EPSILON, p_no);
          k = log(1.0 - args.probability) / log(p_no);
        }
        ++iters;
        if (iters > args.max_iterations) { done = 1; break; }
      }
      if (!done && !(static_cast<double>(iters) < k)) done = 1;
      s_state[0] = best; s_state[1] = iters; s_state[2] = done;
      s_state[3] = min(NG, 2 * round);
      s_k = k;
    }
    __syncthreads();
    if (s_state[2]) break;
    base += round;
  }

  // -------------------------------------------------------------------- selectWithinDistance + InlierExtraction
  __shared__ int s_wcnt[NW];
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
    double *bp = args.best_poses + 7 * b;
    if (running > 0) {
      bp[0] = s_best[12]; bp[1] = s_best[13]; bp[2] = s_best[14]; bp[3] = s_best[15];
      bp[4] = s_best[9]; bp[5] = s_best[10]; bp[6] = s_best[11];
    } else {  // a best model without inliers: undefined in the reference; the start pose here (as the oracle)
      bp[0] = pose[0] * qn; bp[1] = pose[1] * qn; bp[2] = pose[2] * qn; bp[3] = pose[3] * qn;
      bp[4] = pose[4]; bp[5] = pose[5]; bp[6] = pose[6];
    }
    args.num_inliers[b] = running;
    args.iterations[b] = s_state[1];
  }
}

}  // namespace pnec
