// pnec_eigensolver.cuh — rotation from the NEC eigenvalue minimisation:
//   es_moments_kernel   the six 3x3 moment matrices of opengv::relative_pose::eigensolver
//                       (xxF .. zxF = sum_i w_i f1_u f1_v f2 f2^T), optionally weighted with
//                       pnec::common::Weight * 1e-8 (src/common/common.cc:183-208,
//                       src/rel_pose_estimation/pnec.cc:294-306)
//   es_lm_kernel        the minimisation itself: lambda_min(M(c)) over the Cayley parameters by
//                       Levenberg-Marquardt on its gradient (Eigen's port of MINPACK lmdif as opengv
//                       configures it), i.e. opengv::relative_pose::eigensolver(adapter) as called at
//                       pnec.cc:274 (PNEC::Eigensolver) and pnec.cc:313 (PNEC::WeightedEigensolver)
// opengv is not part of the reference tree (it arrives un-pinned inside basalt); the algorithm is
// restated from its publication and public sources, see oracle/pnec_oracle_frame.c.
//
// The streaming part is one pass over f1, f2 (and the covariances for the weights); the
// minimisation touches only the 36 moments of a frame pair, so it runs FOUR LANES per pair (M and
// its three derivatives are assembled one per lane) with the moments in shared memory and a single
// call site for the function evaluation (a small state machine), so that the lanes of a warp stay
// converged through the expensive part.
#pragma once

#include "pnec_translation.cuh"

namespace pnec {

constexpr int kEsMom = 36;  // sym(f1 f1^T) [6] x sym(f2 f2^T) [6], both in xx xy xz yy yz zz order

// ------------------------------------------------------------------ moments

struct EsMomentArgs {
  BatchView bv;   // poses: the pose the weights are computed from (weighted only)
  double reg;
  double *out;    // [B][36]
};

// 48 -> 2 values per lane: after the five exchange steps lane L holds the warp-wide sums of the
// value indices base(L) + {0, 1}, base = 24 b4 + 12 b3 + 6 b2 + 3 b1 + 2 b0 (the second one is
// padding when b0 = 1).  Fixed order => deterministic.
__device__ __forceinline__ void warp_transpose_reduce48(double v[48], int lane) {
  exchange_step<48, 16>(v, lane);
  exchange_step<24, 8>(v, lane);
  exchange_step<12, 4>(v, lane);
  exchange_step<6, 2>(v, lane);
  v[3] = 0.0;
  exchange_step<4, 1>(v, lane);
}

// one correspondence into the 36 sums (acc[6 p + q] += sym(f1 f1^T)[p] * sym(f2 f2^T)[q])
template <bool WEIGHTED>
__device__ __forceinline__ void es_accumulate(const double R[9], const double t[3], double reg, const double f1[3],
                                              double f2[3], const double s6[6], double acc[48]) {
  if (WEIGHTED) {
    // Weight(bv1, bv2, t, R, cov, reg, false) * 1e-8: 1 / (b^T S b + reg), b = R^T (t x f1);
    // the adapter then holds f2 * sqrt(weight) (pnec.cc:302-306)
    double a[3], bb[3], Sb[3];
    cross3(t, f1, a);
    rot_t(R, a, bb);
    sym_mul6(s6, bb, Sb);
    const double w = 1.0e-8 / (dot3(bb, Sb) + reg);
    const double sw = sqrt(w);
#pragma unroll
    for (int k = 0; k < 3; ++k) f2[k] *= sw;
  }
  const double A[6] = {f1[0] * f1[0], f1[0] * f1[1], f1[0] * f1[2], f1[1] * f1[1], f1[1] * f1[2], f1[2] * f1[2]};
  const double F[6] = {f2[0] * f2[0], f2[0] * f2[1], f2[0] * f2[2], f2[1] * f2[1], f2[1] * f2[2], f2[2] * f2[2]};
#pragma unroll
  for (int p = 0; p < 6; ++p)
#pragma unroll
    for (int q = 0; q < 6; ++q) acc[6 * p + q] = fma(A[p], F[q], acc[6 * p + q]);
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(128) es_moments_kernel(const __grid_constant__ EsMomentArgs args) {
  __shared__ double s_part[4][kEsMom];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = blockIdx.x;
  long long s, e;
  problem_range(args.bv, b, s, e);
  double R[9], t[3] = {0, 0, 0};
  if (WEIGHTED) {
    const double *pose = args.bv.poses + 7 * b;
    pose_rotation(pose, R);
    t[0] = pose[4]; t[1] = pose[5]; t[2] = pose[6];
  }
  double acc[48];
#pragma unroll
  for (int k = 0; k < 48; ++k) acc[k] = 0.0;
  for (long long i = s + tid; i < e; i += 128) {
    double f1[3], f2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { f1[k] = args.bv.f1[3 * i + k]; f2[k] = args.bv.f2[3 * i + k]; }
    double s6[6] = {0, 0, 0, 0, 0, 0};
    if (WEIGHTED) {
      double c9[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) c9[k] = args.bv.ct[9 * i + k];
      pack_sym(c9, s6);
    }
    es_accumulate<WEIGHTED>(R, t, args.reg, f1, f2, s6, acc);
  }
  warp_transpose_reduce48(acc, lane);
  const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1, b0 = lane & 1;
  const int base = 24 * b4 + 12 * b3 + 6 * b2 + 3 * b1 + 2 * b0;
  if (base < kEsMom) s_part[warp][base] = acc[0];
  if (b0 == 0 && base + 1 < kEsMom) s_part[warp][base + 1] = acc[1];
  __syncthreads();
  if (tid < kEsMom) args.out[kEsMom * b + tid] = (s_part[0][tid] + s_part[1][tid]) + (s_part[2][tid] + s_part[3][tid]);
}

// The same sums as a persistent, barrier-free stream (the K1 scheme): every warp owns whole frame
// pairs and a private S-stage ring of 32-correspondence tiles filled by bulk async copies (TMA) that
// runs across pair boundaries.  Needs 16-byte aligned arrays; es_moments_kernel is the fallback.
template <bool WEIGHTED, int WPC, int S, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB)
es_moments_warp_kernel(const __grid_constant__ EsMomentArgs args) {
  constexpr int V = WEIGHTED ? PNEC_VARIANT_TARGET : PNEC_VARIANT_NEC;
  constexpr int T = 32;
  constexpr int kStageDoubles = T * VariantTraits<V>::kDoubles;
  __shared__ __align__(8) uint64_t s_full[WPC][S];
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
  const long long W = static_cast<long long>(gridDim.x) * WPC;
  const long long gw = static_cast<long long>(blockIdx.x) * WPC + warp;
  const long long B = args.bv.num_problems;
  const int nmine = (B > gw) ? static_cast<int>((B - gw + W - 1) / W) : 0;
  double *ring = dyn_smem + static_cast<size_t>(warp) * S * kStageDoubles;
  uint64_t *full = s_full[warp];
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < S; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncwarp();
  WarpTileProducer<V, S, T> prod(args.bv, ring, full, gw, W, nmine, lane);
  prod.open();
#pragma unroll 1
  for (int i = 0; i < S; ++i) prod.issue();

  int c_stage = 0;
  uint32_t c_parity = 0;
  for (int j = 0; j < nmine; ++j) {
    const long long prob = gw + j * W;
    long long s, e;
    problem_range(args.bv, prob, s, e);
    const int head = static_cast<int>(s & 1LL);
    int c_left = (e > s) ? static_cast<int>(e - s) + head : 0;  // elements left, head included
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
    if (WEIGHTED) {
      const double *pose = args.bv.poses + 7 * prob;
      pose_rotation(pose, R);
      t[0] = pose[4]; t[1] = pose[5]; t[2] = pose[6];
    }
    double acc[48];
#pragma unroll
    for (int k = 0; k < 48; ++k) acc[k] = 0.0;
    int lo = head;  // first valid element of the tile: head on the first tile, 0 afterwards
    while (c_left > 0) {
      mbar_wait(&full[c_stage], c_parity);
      const double *base = ring + c_stage * kStageDoubles;
      const bool valid = (lane >= lo) && (lane < c_left);
      double a1[3], a2[3], c1[6], c2[6];
      if (valid) load_corr<V>(base, base + 3 * T, base + 6 * T, base + 15 * T, lane, a1, a2, c1, c2);
      __syncwarp();  // every lane holds its correspondence: the stage may be refilled
      prod.issue();
      if (valid) es_accumulate<WEIGHTED>(R, t, args.reg, a1, a2, c1, acc);
      if (++c_stage == S) {
        c_stage = 0;
        c_parity ^= 1u;
      }
      c_left -= T;
      lo = 0;
    }
    warp_transpose_reduce48(acc, lane);
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1, b0 = lane & 1;
    const int idx = 24 * b4 + 12 * b3 + 6 * b2 + 3 * b1 + 2 * b0;
    if (idx < kEsMom) args.out[kEsMom * prob + idx] = acc[0];
    if (b0 == 0 && idx + 1 < kEsMom) args.out[kEsMom * prob + idx + 1] = acc[1];
  }
}

// ------------------------------------------------- lambda_min(M(c)) and gradient

// scatter of W_uv = R' G_uv R'^T (sym 6) into M = sum n n^T (sym 6), n = f1 x R' f2:
//   M_ab = sum eps_aup eps_bvs W_uv[p][s]
template <int UV>
__device__ __forceinline__ void es_scatter(const double W[6], double M[6]) {
  // indices: 0 = 00, 1 = 01, 2 = 02, 3 = 11, 4 = 12, 5 = 22
  if (UV == 0) { M[3] += W[5]; M[5] += W[3]; M[4] -= W[4]; }
  if (UV == 1) { M[5] -= 2.0 * W[1]; M[1] -= W[5]; M[2] += W[4]; M[4] += W[2]; }
  if (UV == 2) { M[3] -= 2.0 * W[2]; M[1] += W[4]; M[2] -= W[3]; M[4] += W[1]; }
  if (UV == 3) { M[0] += W[5]; M[5] += W[0]; M[2] -= W[2]; }
  if (UV == 4) { M[0] -= 2.0 * W[4]; M[1] += W[2]; M[2] += W[1]; M[4] -= W[0]; }
  if (UV == 5) { M[0] += W[3]; M[3] += W[0]; M[1] -= W[1]; }
}

// One of the six (u, v) terms: X_uv = A P + (A P)^T with P = G_uv R'^T, scattered into Y.
// With A = R'/2 that is W_uv = R' G_uv R'^T (Y = M); with A = dR'/dc_k it is dW_uv/dc_k (Y = dM/dc_k).
template <int UV>
__device__ __forceinline__ void es_term(const double *mom, int stride, const double R[3][3],
                                        const double A[3][3], double Y[6]) {
  double G[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) G[q] = mom[(6 * UV + q) * stride];
  const double Gf[3][3] = {{G[0], G[1], G[2]}, {G[1], G[3], G[4]}, {G[2], G[4], G[5]}};
  double P[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int sI = 0; sI < 3; ++sI) P[r][sI] = Gf[r][0] * R[sI][0] + Gf[r][1] * R[sI][1] + Gf[r][2] * R[sI][2];
  double X[6];
  int k = 0;
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int sI = p; sI < 3; ++sI)
      X[k++] = (A[p][0] * P[0][sI] + A[p][1] * P[1][sI] + A[p][2] * P[2][sI]) +
               (A[sI][0] * P[0][p] + A[sI][1] * P[1][p] + A[sI][2] * P[2][p]);
  es_scatter<UV>(X, Y);
}

// opengv eigensolver::getSmallestEVwithJacobian: closed-form smallest root of the characteristic
// cubic of M(c) and its derivative with respect to the Cayley parameters, computed by a GROUP OF
// FOUR LANES (sub = lane & 3): lane 0 assembles M, lanes 1..3 one dM/dc_k each; M is broadcast and
// every lane then differentiates the closed form along its own dM.  All four lanes return the same
// ev and grad.  `mom` points at this pair's first moment; consecutive moments are `stride` apart.
__device__ __forceinline__ double es_smallest_ev(const double *mom, int stride, const double c[3], int sub,
                                                 double grad[3]) {
  const double x = c[0], y = c[1], z = c[2];
  // opengv::math::cayley2rot_reduced: (1 + |c|^2) * rotation
  const double R[3][3] = {{1 + x * x - y * y - z * z, 2 * (x * y - z), 2 * (x * z + y)},
                          {2 * (x * y + z), 1 - x * x + y * y - z * z, 2 * (y * z - x)},
                          {2 * (x * z - y), 2 * (y * z + x), 1 - x * x - y * y + z * z}};
  // A = R'/2 (sub 0) or dR'/dc_(sub-1), selected by value so that the four lanes stay converged
  const double hx = sub == 1 ? 1.0 : 0.0, hy = sub == 2 ? 1.0 : 0.0, hz = sub == 3 ? 1.0 : 0.0, h0 = sub == 0 ? 0.5 : 0.0;
  // dR'/dx = 2 [[x, y, z], [y, -x, -1], [z, 1, -x]], dR'/dy = 2 [[-y, x, 1], [x, y, z], [-1, z, -y]],
  // dR'/dz = 2 [[-z, -1, x], [1, -z, y], [x, y, z]]
  double A[3][3];
  A[0][0] = h0 * R[0][0] + 2 * (hx * x - hy * y - hz * z);
  A[0][1] = h0 * R[0][1] + 2 * (hx * y + hy * x - hz);
  A[0][2] = h0 * R[0][2] + 2 * (hx * z + hy + hz * x);
  A[1][0] = h0 * R[1][0] + 2 * (hx * y + hy * x + hz);
  A[1][1] = h0 * R[1][1] + 2 * (-hx * x + hy * y - hz * z);
  A[1][2] = h0 * R[1][2] + 2 * (-hx + hy * z + hz * y);
  A[2][0] = h0 * R[2][0] + 2 * (hx * z - hy + hz * x);
  A[2][1] = h0 * R[2][1] + 2 * (hx + hy * z + hz * y);
  A[2][2] = h0 * R[2][2] + 2 * (-hx * x - hy * y + hz * z);
  double Y[6] = {0, 0, 0, 0, 0, 0};
  es_term<0>(mom, stride, R, A, Y);
  es_term<1>(mom, stride, R, A, Y);
  es_term<2>(mom, stride, R, A, Y);
  es_term<3>(mom, stride, R, A, Y);
  es_term<4>(mom, stride, R, A, Y);
  es_term<5>(mom, stride, R, A, Y);
  // M from lane 0 of the group; this lane's derivative direction J = Y (meaningless on lane 0)
  double M[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) M[k] = __shfl_sync(0xffffffffu, Y[k], 0, 4);
  // M: 0 = 00, 1 = 01, 2 = 02, 3 = 11, 4 = 12, 5 = 22
  const double m00 = M[0], m01 = M[1], m02 = M[2], m11 = M[3], m12 = M[4], m22 = M[5];
  const double b = -m00 - m11 - m22;
  const double cc = -m02 * m02 - m12 * m12 - m01 * m01 + m00 * m11 + m00 * m22 + m11 * m22;
  const double d = m11 * m02 * m02 + m00 * m12 * m12 + m22 * m01 * m01 - m00 * m11 * m22 - 2 * m01 * m12 * m02;
  const double s = 2 * b * b * b - 9 * b * cc + 27 * d;
  // t = 4 q^3, sqrt(t) = 2 q^(3/2), w = (sqrt(t)/2)^(1/3) = sqrt(q); everything below is opengv's chain
  // rule with the divisions turned into reciprocals
  const double q = b * b - 3 * cc;
  const double irq = rsqrt(q);          // 1 / w
  const double w = q * irq;
  const double ist = 0.5 * irq * irq * irq;  // 1 / sqrt(t)
  const double ratio = s * ist;         // cos(alpha)
  // y = cos(alpha / 3): the largest root of 4 y^3 - 3 y = cos(alpha), in [1/2, 1] where the cubic is
  // increasing and convex, so Newton from y = 1 descends to it monotonically (instead of acos + cos).
  // (A single-precision acos / cos seed was tried: no faster, and it lands on the other side of the
  // double root at y = 1/2 for near-degenerate pairs.)
  double cb = 1.0;
#pragma unroll 1
  for (int it = 0; it < 48; ++it) {
    const double f = fma(fma(4.0 * cb, cb, -3.0), cb, -ratio);
    const double fp = fma(12.0 * cb, cb, -3.0);
    const double dy = f * fast_rcp(fp);
    cb -= dy;
    if (!(fabs(dy) > 4e-16)) break;
  }
  const double sb = sqrt(fmax(0.0, fma(-cb, cb, 1.0)));  // sin(alpha / 3) >= 0
  const double ev = (-b - 2.0 * (w * cb)) * (1.0 / 3.0);
  const double inv_sin_alpha = -rsqrt(fma(-ratio, ratio, 1.0));
  double g;
  {
    const double j00 = Y[0], j01 = Y[1], j02 = Y[2], j11 = Y[3], j12 = Y[4], j22 = Y[5];
    const double bj = -j00 - j11 - j22;
    const double cj = -2.0 * m02 * j02 - 2.0 * m12 * j12 - 2.0 * m01 * j01 + j00 * m11 + m00 * j11 + j00 * m22 +
                      m00 * j22 + j11 * m22 + m11 * j22;
    const double dj = j11 * m02 * m02 + 2.0 * m11 * m02 * j02 + j00 * m12 * m12 + 2.0 * m00 * m12 * j12 +
                      j22 * m01 * m01 + 2.0 * m22 * m01 * j01 - j00 * m11 * m22 - m00 * j11 * m22 - m00 * m11 * j22 -
                      2.0 * (j01 * m12 * m02 + m01 * j12 * m02 + m01 * m12 * j02);
    const double sj = 6.0 * b * b * bj - 9.0 * bj * cc - 9.0 * b * cj + 27.0 * dj;
    const double qj = 2.0 * b * bj - 3.0 * cj;
    const double tj = 12.0 * q * q * qj;
    // d alpha = -(1 / sin alpha) d(s / sqrt t) = -(1 / sin alpha) (sj / sqrt t - s tj / (2 t sqrt t))
    const double alphaj = inv_sin_alpha * (sj * ist - 0.5 * s * tj * (ist * ist * ist));
    const double yj = -sb * (alphaj * (1.0 / 3.0));
    const double wj = 0.5 * qj * irq;
    const double kj = wj * cb + w * yj;
    g = (-bj - 2.0 * kj) * (1.0 / 3.0);
  }
  grad[0] = __shfl_sync(0xffffffffu, g, 1, 4);
  grad[1] = __shfl_sync(0xffffffffu, g, 2, 4);
  grad[2] = __shfl_sync(0xffffffffu, g, 3, 4);
  return ev;
}

// ------------------------------------------------------- MINPACK lmdif, n = m = 3

__device__ __forceinline__ double pick3(const double v[3], int k) { return k == 0 ? v[0] : (k == 1 ? v[1] : v[2]); }
__device__ __forceinline__ void put3(double v[3], int k, double x) {
  if (k == 0) v[0] = x; else if (k == 1) v[1] = x; else v[2] = x;
}
// The MINPACK logic runs on every lane of a 4-lane group and the groups of a warp sit in different
// states, so its cost is paid by the whole warp in almost every turn of the loop: divisions and square
// roots (each a ~40-instruction sequence with a slow-path branch in IEEE form) go through the
// branch-free reciprocal / reciprocal-square-root forms (<= 2 ulp).  Operands that can be zero are
// guarded by MINPACK's own tests before they are used as divisors.
__device__ __forceinline__ double es_div(double a, double b) { return a * fast_rcp(b); }
__device__ __forceinline__ double es_sqrt(double x) { return fast_sqrt(x); }
__device__ __forceinline__ double enorm3(const double v[3]) { return es_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// qrfac with column pivoting (a[i][j], row i, column j).  On exit: strict upper triangle = R,
// rdiag = diagonal of R, lower trapezoid = Householder vectors, acnorm = input column norms.
__device__ __forceinline__ void es_qrfac(double a[3][3], int ipvt[3], double rdiag[3], double acnorm[3]) {
  double wa[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    acnorm[j] = es_sqrt(a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j]);
    rdiag[j] = acnorm[j];
    wa[j] = acnorm[j];
    ipvt[j] = j;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    int kmax = j;
    double best = rdiag[j];
#pragma unroll
    for (int k = j + 1; k < 3; ++k)
      if (rdiag[k] > best) { best = rdiag[k]; kmax = k; }
#pragma unroll
    for (int k = j + 1; k < 3; ++k)
      if (kmax == k) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double tmp = a[i][j]; a[i][j] = a[i][k]; a[i][k] = tmp; }
        rdiag[k] = rdiag[j];
        wa[k] = wa[j];
        const int ti = ipvt[j]; ipvt[j] = ipvt[k]; ipvt[k] = ti;
      }
    double ss = 0.0;
#pragma unroll
    for (int i = j; i < 3; ++i) ss += a[i][j] * a[i][j];
    double ajnorm = es_sqrt(ss);
    if (ajnorm != 0.0) {
      if (a[j][j] < 0.0) ajnorm = -ajnorm;
      {
        const double inv = fast_rcp(ajnorm);
#pragma unroll
        for (int i = j; i < 3; ++i) a[i][j] *= inv;
      }
      a[j][j] += 1.0;
#pragma unroll
      for (int k = j + 1; k < 3; ++k) {
        double sum = 0.0;
#pragma unroll
        for (int i = j; i < 3; ++i) sum += a[i][j] * a[i][k];
        const double temp = es_div(sum, a[j][j]);  // a[j][j] in [1, 2]
#pragma unroll
        for (int i = j; i < 3; ++i) a[i][k] -= temp * a[i][j];
        if (rdiag[k] != 0.0) {
          const double tq = es_div(a[j][k], rdiag[k]);
          rdiag[k] *= es_sqrt(fmax(0.0, 1.0 - tq * tq));
          const double qq = es_div(rdiag[k], wa[k]);
          if (0.05 * qq * qq <= DBL_EPSILON) {
            double rs = 0.0;
#pragma unroll
            for (int i = j + 1; i < 3; ++i) rs += a[i][k] * a[i][k];
            rdiag[k] = es_sqrt(rs);
            wa[k] = rdiag[k];
          }
        }
      }
    }
    rdiag[j] = -ajnorm;
  }
}

// qrsolv: r upper triangle in, lower triangle receives the transposed S; `dp[j]` = diag[ipvt[j]].
// xp receives the solution in PIVOTED order (xp[j] belongs to variable ipvt[j]).
__device__ __forceinline__ void es_qrsolv(double r[3][3], const double dp[3], const double qtb[3], double xp[3],
                                          double sdiag[3]) {
  double wa[3], save[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int i = j; i < 3; ++i) r[i][j] = r[j][i];
    save[j] = r[j][j];
    wa[j] = qtb[j];
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (dp[j] != 0.0) {
#pragma unroll
      for (int k = j; k < 3; ++k) sdiag[k] = 0.0;
      sdiag[j] = dp[j];
      double qtbpj = 0.0;
#pragma unroll
      for (int k = j; k < 3; ++k) {
        if (sdiag[k] != 0.0) {
          double sn, cs;
          if (fabs(r[k][k]) < fabs(sdiag[k])) {
            const double cotan = es_div(r[k][k], sdiag[k]);
            sn = 0.5 * rsqrt(0.25 + 0.25 * cotan * cotan);
            cs = sn * cotan;
          } else {
            const double tn = es_div(sdiag[k], r[k][k]);
            cs = 0.5 * rsqrt(0.25 + 0.25 * tn * tn);
            sn = cs * tn;
          }
          r[k][k] = cs * r[k][k] + sn * sdiag[k];
          const double temp = cs * wa[k] + sn * qtbpj;
          qtbpj = -sn * wa[k] + cs * qtbpj;
          wa[k] = temp;
#pragma unroll
          for (int i = k + 1; i < 3; ++i) {
            const double t2 = cs * r[i][k] + sn * sdiag[i];
            sdiag[i] = -sn * r[i][k] + cs * sdiag[i];
            r[i][k] = t2;
          }
        }
      }
    }
    sdiag[j] = r[j][j];
    r[j][j] = save[j];
  }
  int nsing = 3;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (sdiag[j] == 0.0 && nsing == 3) nsing = j;
    if (nsing < 3) wa[j] = 0.0;
  }
#pragma unroll
  for (int j = 2; j >= 0; --j) {
    if (j < nsing) {
      double sum = 0.0;
#pragma unroll
      for (int i = j + 1; i < 3; ++i)
        if (i < nsing) sum += r[i][j] * wa[i];
      wa[j] = es_div(wa[j] - sum, sdiag[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) xp[j] = wa[j];
}

// lmpar in pivoted coordinates: dp[j] = diag[ipvt[j]]; xp[j] = step component of variable ipvt[j].
__device__ __forceinline__ void es_lmpar(double r[3][3], const double dp[3], const double qtb[3], double delta,
                                         double &par, double xp[3]) {
  const double dwarf = DBL_MIN;
  double wa1[3], wa2[3], sdiag[3];
  int nsing = 3;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    wa1[j] = qtb[j];
    if (r[j][j] == 0.0 && nsing == 3) nsing = j;
    if (nsing < 3) wa1[j] = 0.0;
  }
#pragma unroll
  for (int j = 2; j >= 0; --j) {
    if (j < nsing) {
      wa1[j] = es_div(wa1[j], r[j][j]);
      const double temp = wa1[j];
#pragma unroll
      for (int i = 0; i < j; ++i) wa1[i] -= r[i][j] * temp;
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) { xp[j] = wa1[j]; wa2[j] = dp[j] * xp[j]; }
  double dxnorm = enorm3(wa2);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) {
    par = 0.0;
    return;
  }
  double parl = 0.0;
  if (nsing >= 3) {
    {
      const double inv = fast_rcp(dxnorm);
#pragma unroll
      for (int j = 0; j < 3; ++j) wa1[j] = dp[j] * (wa2[j] * inv);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double sum = 0.0;
#pragma unroll
      for (int i = 0; i < j; ++i) sum += r[i][j] * wa1[i];
      wa1[j] = es_div(wa1[j] - sum, r[j][j]);
    }
    const double temp = enorm3(wa1);
    parl = ((fp / delta) / temp) / temp;  // (once per call: IEEE, temp may vanish)
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i <= j; ++i) sum += r[i][j] * qtb[i];
    wa1[j] = es_div(sum, dp[j]);
  }
  const double gnorm = enorm3(wa1);
  double paru = es_div(gnorm, delta);
  if (paru == 0.0) paru = dwarf / fmin(delta, 0.1);
  par = fmin(fmax(par, parl), paru);
  if (par == 0.0) par = es_div(gnorm, dxnorm);
  for (int iter = 1;; ++iter) {
    if (par == 0.0) par = fmax(dwarf, 0.001 * paru);
    double temp = es_sqrt(par);
#pragma unroll
    for (int j = 0; j < 3; ++j) wa1[j] = temp * dp[j];
    es_qrsolv(r, wa1, qtb, xp, sdiag);
#pragma unroll
    for (int j = 0; j < 3; ++j) wa2[j] = dp[j] * xp[j];
    dxnorm = enorm3(wa2);
    temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    {
      const double inv = fast_rcp(dxnorm);
#pragma unroll
      for (int j = 0; j < 3; ++j) wa1[j] = dp[j] * (wa2[j] * inv);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      wa1[j] = es_div(wa1[j], sdiag[j]);
      const double t2 = wa1[j];
#pragma unroll
      for (int i = j + 1; i < 3; ++i) wa1[i] -= r[i][j] * t2;
    }
    temp = enorm3(wa1);
    const double it = fast_rcp(temp);
    const double parc = (fp * fast_rcp(delta)) * it * it;
    if (fp > 0.0) parl = fmax(parl, par);
    if (fp < 0.0) paru = fmin(paru, par);
    par = fmax(parl, par + parc);
  }
}

struct EsLmArgs {
  const double *moments;   // [B][36]
  const double *poses_in;  // [B][7]: start rotation (R12 of the adapter); translation is passed through
  double *poses_out;       // [B][7]: unit quaternion of cayley2rot(result) + the input translation
  int *out_info;           // [B] MINPACK info code, or nullptr
  int *out_nfev;           // [B] function evaluations (Eigen's accounting), or nullptr
  double *out_ev;          // [B] lambda_min of the reduced M at the result, or nullptr
  const int *fixed;        // [B] or nullptr: pairs to pass through unchanged (fixed point reached)
  int *q_same;             // [B] or nullptr: out, result quaternion == input quaternion bit for bit
                           // (may alias `fixed`: a rotation that repeats once repeats forever, because the
                           // next call would start from the same bits)
  int copy_translation;    // pass poses_in[4..6] through to poses_out[4..6]
  long long num_problems;
  double ftol, xtol, gtol, factor;
  int maxfev;
};

constexpr int kEsLmThreads = 128;
constexpr int kEsLmPairs = kEsLmThreads / 4;  // four lanes per frame pair

// The Levenberg-Marquardt run of one 4-lane group (opengv's eigensolver_main on the 36 moments at
// `mom`, consecutive moments `stride` doubles apart).  Must be called by all 32 lanes of a warp with
// group = lane / 4, sub = lane % 4; groups without work pass active = false.  The moments of group g'
// must be readable at mom + (g' - g) * gpitch by every lane of the warp (see the wide turns below).
// The four lanes share the function evaluation (see es_smallest_ev) and run the scalar logic
// redundantly, so a group never diverges.
//
// The eight groups of a warp move through lmdif's phases TOGETHER (the phase is warp-uniform):
// f(x0); then per outer iteration the three forward-difference columns, the QR factorisation /
// gradient test / first step, and trial points until every group of the warp has either accepted
// one (ratio >= 1e-4) or terminated -- groups that are through wait for the others.  A turn of the
// loop therefore executes ONE kind of post-processing for the whole warp (with unsynchronised
// groups nearly every turn paid for the factorisation, the step computation and the step
// assessment, each for a fraction of the lanes), and there is a single call site of the evaluation.
// Each group performs exactly lmdif's sequence of operations; only the interleaving differs.
//
// WIDE TURNS.  A warp's duration is its slowest group's, and a kernel's is that of the few pairs that
// run into maxfev: most of the time one or two groups of a warp are still at work and 24+ lanes idle.
// As soon as at most two groups are left, every turn evaluates FOUR points per group at once: the
// group's own lanes take its base point (the pending trial point, or x when a Jacobian is due) and
// three finished groups take base + h_k e_k, k = 0..2 -- the forward-difference columns lmdif would
// ask for next if the trial point is accepted.  It nearly always is, so an outer iteration costs one
// evaluation turn instead of four; a rejected trial point discards the three columns.  The values
// are the ones the narrow turns would compute (same points, same arithmetic, nfev counted as lmdif
// counts it): results are bit-identical, only the latency of the tail changes.  A warp with a single
// pair (frame_rounds_kernel, one pair per call) runs wide from the first turn.
struct EsLmParams {
  double ftol, xtol, gtol, factor;
  int maxfev;
};

//
// WIDE = false compiles the wide turns out: the RANSAC kernels are bound by instruction fetch, their
// eight hypotheses per warp finish within a few turns of each other, and the extra code costs them
// more than the shorter tail returns (measured: stage on C2 15.2 -> 16.8 ms with wide turns).
template <bool WIDE>
__device__ __forceinline__ void es_lm_group(const double *mom, int stride, int gpitch, const EsLmParams &args,
                                            bool active, int sub, double x[3], int &info_out, int &nfev_out) {
  const unsigned kFull = 0xffffffffu;
  const double epsmch = DBL_EPSILON;
  const double eps = es_sqrt(epsmch);  // epsfcn = 0
  const int lane = static_cast<int>(threadIdx.x) & 31, grp = lane >> 2;
  double fvec[3] = {0, 0, 0}, r[3][3], diag[3] = {1, 1, 1}, dp[3] = {1, 1, 1}, qtf[3] = {0, 0, 0}, wa1[3] = {0, 0, 0};
  double wa2[3] = {x[0], x[1], x[2]}, acn[3] = {0, 0, 0};
  int ipvt[3] = {0, 1, 2};
  double fnorm = 0.0, par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0, h = 0.0, pnorm = 0.0;
  int info = 0, nfev = 0, iter = 1;
  int phase = 0;        // warp-uniform: 0 = f(x0), 1..3 = forward-difference column phase - 1, 4 = trial point
  bool done = !active;  // this group has terminated
  bool trying = false;  // this group has a trial point pending (phase 4)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r[i][j] = 0.0;
  if (__all_sync(kFull, done)) {
    info_out = 0;
    nfev_out = 0;
    return;
  }

  for (;;) {
    // ---- wide or narrow turn (warp-uniform; the set of unfinished groups only changes after phases 3 and 4)
    const unsigned act_lanes = __ballot_sync(kFull, !done);
    const bool wide = WIDE && __popc(act_lanes) <= 8 && (phase == 0 || phase == 1 || phase == 4);
    const bool was_trying = trying;
    // roles of a wide turn: an unfinished group evaluates its own base point (col = -1); the finished
    // groups, in lane order, serve the unfinished ones three by three (col = 0..2)
    int col = -1, serve = grp;
    int src_lane[3] = {lane, lane, lane};  // an unfinished group's helpers (first lane of each)
    if (wide) {
      const unsigned a0 = __ffs(act_lanes) - 1;                        // first lane of the first unfinished group
      const unsigned rest = act_lanes & ~(0xfu << a0);
      const int a1 = rest ? __ffs(rest) - 1 : -1;                      // ... of the second, if any
      const unsigned idle = ~act_lanes;
      if (done) {
        const int rank = __popc(idle & ((1u << (lane & ~3)) - 1u)) >> 2;  // finished groups below this one
        const int tgt = rank / 3;
        col = rank - 3 * tgt;
        serve = tgt == 0 ? static_cast<int>(a0 >> 2) : (tgt == 1 && a1 >= 0 ? a1 >> 2 : -1);
        if (serve < 0) { serve = grp; col = -1; }  // nothing to do this turn
      } else {
        const int first = (a1 >= 0 && lane >= a1) ? 3 : 0;  // rank of this group's first helper
#pragma unroll
        for (int k = 0; k < 3; ++k) src_lane[k] = static_cast<int>(__fns(idle, 0, 4 * (first + k) + 1));
      }
    }
    // ---- the point to evaluate
    double xe[3] = {x[0], x[1], x[2]};
    if (trying) { xe[0] = wa2[0]; xe[1] = wa2[1]; xe[2] = wa2[2]; }
    const double *mom_e = mom;
    if (wide) {
      const int src = 4 * serve;
#pragma unroll
      for (int j = 0; j < 3; ++j) xe[j] = __shfl_sync(kFull, xe[j], src);
      mom_e = mom + (serve - grp) * gpitch;
      if (col >= 0) {
        const double xj = pick3(xe, col);
        double hh = eps * fabs(xj);
        if (hh == 0.0) hh = eps;
        put3(xe, col, xj + hh);
      }
    } else if (phase >= 1 && phase <= 3) {
      const double xj = pick3(x, phase - 1);
      h = eps * fabs(xj);
      if (h == 0.0) h = eps;
      put3(xe, phase - 1, xj + h);
    }
    double fe[3];
    es_smallest_ev(mom_e, stride, xe, sub, fe);
    // a wide turn's columns: f(base + h_k e_k) from the helpers
    double fd[3][3];
    if (wide) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < 3; ++i) fd[k][i] = __shfl_sync(kFull, fe[i], src_lane[k]);
    }

    bool need_step = false;
    bool new_jac = false;  // r holds fresh forward differences at x
    if (phase == 0) {
      if (!done) {
        fvec[0] = fe[0]; fvec[1] = fe[1]; fvec[2] = fe[2];
        fnorm = enorm3(fvec);
        nfev = 1;
      }
      if (!wide) {
        phase = 1;
        continue;
      }
    }
    if (!wide && phase <= 3) {
      if (!done) {
        const int j = phase - 1;
        const double inv_h = fast_rcp(h);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double v = (fe[i] - fvec[i]) * inv_h;
          if (j == 0) r[i][0] = v; else if (j == 1) r[i][1] = v; else r[i][2] = v;
        }
      }
      if (phase < 3) {
        phase += 1;
        continue;
      }
      new_jac = !done;
    }
    if (phase == 4 && was_trying) {
      // ---- trial point evaluated: fe = f(x + p)
      ++nfev;
      const double fnorm1 = enorm3(fe);
      double actred = -1.0;
      const double inv_fnorm = fast_rcp(fnorm);  // fnorm > 0: a zero residual stops at the gradient test
      if (0.1 * fnorm1 < fnorm) actred = 1.0 - (fnorm1 * inv_fnorm) * (fnorm1 * inv_fnorm);
      // wa3 = R * P^T p  (wa1 holds p in pivoted order)
      double w3[3] = {0, 0, 0};
#pragma unroll
      for (int j2 = 0; j2 < 3; ++j2)
#pragma unroll
        for (int i = 0; i <= j2; ++i) w3[i] += r[i][j2] * wa1[j2];
      const double t1 = enorm3(w3) * inv_fnorm, t2 = es_sqrt(par) * pnorm * inv_fnorm;
      const double temp1 = t1 * t1, temp2 = t2 * t2;
      const double prered = temp1 + temp2 * 2.0;
      const double dirder = -(temp1 + temp2);
      double ratio = 0.0;
      if (prered != 0.0) ratio = es_div(actred, prered);
      if (ratio <= 0.25) {
        double temp = 0.5;
        if (actred < 0.0) temp = es_div(0.5 * dirder, dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
        delta = temp * fmin(delta, pnorm * 10.0);
        par = es_div(par, temp);
      } else if (!(par != 0.0 && ratio < 0.75)) {
        delta = pnorm * 2.0;
        par = 0.5 * par;
      }
      if (ratio >= 1e-4) {
#pragma unroll
        for (int j2 = 0; j2 < 3; ++j2) { x[j2] = wa2[j2]; fvec[j2] = fe[j2]; }
        const double dx[3] = {diag[0] * x[0], diag[1] * x[1], diag[2] * x[2]};
        xnorm = enorm3(dx);
        fnorm = fnorm1;
        ++iter;
      }
      const bool small_red = fabs(actred) <= args.ftol && prered <= args.ftol && 0.5 * ratio <= 1.0;
      if (small_red) info = 1;
      if (delta <= args.xtol * xnorm) info = 2;
      if (small_red && info == 2) info = 3;
      if (info == 0) {
        if (nfev >= args.maxfev) info = 5;
        if (fabs(actred) <= epsmch && prered <= epsmch && 0.5 * ratio <= 1.0) info = 6;
        if (delta <= epsmch * xnorm) info = 7;
        if (gnorm <= epsmch) info = 8;
      }
      trying = false;
      if (info != 0) {
        done = true;
      } else if (ratio < 1e-4) {
        need_step = true;  // inner loop of lmdif: same Jacobian, smaller region
      }
    }
    if (wide && !done && !need_step) {
      // the helpers' points were x + h_k e_k (x = the base point: just accepted, or unchanged)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double hh = eps * fabs(x[k]);
        if (hh == 0.0) hh = eps;
        const double inv_h = fast_rcp(hh);
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i][k] = (fd[k][i] - fvec[i]) * inv_h;
      }
      new_jac = true;
    }
    if (new_jac) {
      nfev += 4;  // Eigen's NumericalDiff (Forward) evaluates f(x) again: n + 1 calls per Jacobian
      double rdiag[3];
      es_qrfac(r, ipvt, rdiag, acn);
      if (iter == 1) {
#pragma unroll
        for (int j2 = 0; j2 < 3; ++j2) diag[j2] = (acn[j2] == 0.0) ? 1.0 : acn[j2];
        const double dx[3] = {diag[0] * x[0], diag[1] * x[1], diag[2] * x[2]};
        xnorm = enorm3(dx);
        delta = args.factor * xnorm;
        if (delta == 0.0) delta = args.factor;
      }
      // qtf = first n components of Q^T fvec; store R's diagonal
      double w4[3] = {fvec[0], fvec[1], fvec[2]};
#pragma unroll
      for (int j2 = 0; j2 < 3; ++j2) {
        if (r[j2][j2] != 0.0) {
          double sum = 0.0;
#pragma unroll
          for (int i = j2; i < 3; ++i) sum += r[i][j2] * w4[i];
          const double temp = -es_div(sum, r[j2][j2]);
#pragma unroll
          for (int i = j2; i < 3; ++i) w4[i] += r[i][j2] * temp;
        }
        r[j2][j2] = rdiag[j2];
        qtf[j2] = w4[j2];
      }
      gnorm = 0.0;
      if (fnorm != 0.0) {
        const double inv_fnorm = fast_rcp(fnorm);
#pragma unroll
        for (int j2 = 0; j2 < 3; ++j2) {
          const double cn = pick3(acn, ipvt[j2]);
          if (cn != 0.0) {
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i <= j2; ++i) sum += r[i][j2] * (qtf[i] * inv_fnorm);
            gnorm = fmax(gnorm, fabs(es_div(sum, cn)));
          }
        }
      }
      if (gnorm <= args.gtol) {
        info = 4;
        done = true;
      } else {
#pragma unroll
        for (int j2 = 0; j2 < 3; ++j2) diag[j2] = fmax(diag[j2], acn[j2]);
#pragma unroll
        for (int j2 = 0; j2 < 3; ++j2) dp[j2] = pick3(diag, ipvt[j2]);
        need_step = true;
      }
    }
    if (need_step) {
      // wa1 <- step in pivoted order, wa2 <- x + p
      es_lmpar(r, dp, qtf, delta, par, wa1);
      double dpn[3];
#pragma unroll
      for (int j2 = 0; j2 < 3; ++j2) {
        wa1[j2] = -wa1[j2];
        dpn[j2] = dp[j2] * wa1[j2];
      }
      wa2[0] = x[0]; wa2[1] = x[1]; wa2[2] = x[2];
#pragma unroll
      for (int j2 = 0; j2 < 3; ++j2) put3(wa2, ipvt[j2], pick3(x, ipvt[j2]) + wa1[j2]);
      pnorm = enorm3(dpn);
      if (iter == 1) delta = fmin(delta, pnorm);
      trying = true;
    }
    // ---- next phase (warp-uniform): more trial points while any group has one pending, else a new
    // outer iteration for the groups that are not done, else out
    if (__any_sync(kFull, trying)) phase = 4;
    else if (__all_sync(kFull, done)) break;
    else phase = 1;
  }
  info_out = info;
  nfev_out = nfev;
}

// Four lanes per frame pair (es_lm_group), moments staged in shared memory.  GPW = pairs per warp:
// 8 fills every 4-lane group (32 pairs per CTA; es_lm_group goes wide for the warp's tail), 2 leaves
// six groups of each warp free from the start, so that EVERY turn is a wide one (8 pairs per CTA):
// per pair more instructions but a quarter of the dependent turns -- for batches that do not fill
// the machine with narrow turns (the launcher decides, run_es_lm).
template <int GPW>
__global__ void __launch_bounds__(kEsLmThreads) es_lm_kernel(const __grid_constant__ EsLmArgs args) {
  constexpr int kPairs = (kEsLmThreads / 32) * GPW;  // pairs per CTA
  __shared__ double s_mom[kEsMom * kEsLmPairs];      // [k][slot], slot = 4-lane group of the CTA
  const int tid = threadIdx.x, sub = tid & 3, slot = tid >> 2, warp = tid >> 5, grp = slot & 7;
  const long long first = static_cast<long long>(blockIdx.x) * kPairs;
  const long long b = first + warp * GPW + grp;
  const bool in_range = grp < GPW && b < args.num_problems;
  const long long bb = in_range ? b : args.num_problems - 1;
  const bool passthrough = in_range && args.fixed && args.fixed[bb];
  const bool active = in_range && !passthrough;
  // coalesced staging of this CTA's moments, transposed to [k][slot]
  {
    const long long cnt = min(static_cast<long long>(kPairs), args.num_problems - first);
    if (GPW < 8)
      for (int i = tid; i < kEsLmPairs * kEsMom; i += kEsLmThreads) s_mom[i] = 0.0;
    if (GPW < 8) __syncthreads();
    for (long long i = tid; i < kPairs * kEsMom; i += kEsLmThreads) {
      const int p = static_cast<int>(i / kEsMom), k = static_cast<int>(i % kEsMom);
      const int sl = (p / GPW) * 8 + (p % GPW);
      if (GPW < 8) {
        if (i < cnt * kEsMom) s_mom[k * kEsLmPairs + sl] = args.moments[first * kEsMom + i];
      } else {
        s_mom[k * kEsLmPairs + sl] = i < cnt * kEsMom ? args.moments[first * kEsMom + i] : 0.0;
      }
    }
    __syncthreads();
  }
  const double *mom = s_mom + slot;  // every lane its own slot: es_lm_group's wide turns address the others from it
  const double *pin = args.poses_in + 7 * bb;
  // opengv::math::rot2cayley: [c]x = (R - I)(R + I)^-1, i.e. q_xyz / q_w
  double x[3] = {pin[0] / pin[3], pin[1] / pin[3], pin[2] / pin[3]};

  const EsLmParams params{args.ftol, args.xtol, args.gtol, args.factor, args.maxfev};
  int info = 0, nfev = 0;
  es_lm_group<true>(mom, kEsLmPairs, 1, params, active, sub, x, info, nfev);
  if (passthrough && sub == 0) {
    double *po = args.poses_out + 7 * b;
#pragma unroll
    for (int k = 0; k < 4; ++k) po[k] = pin[k];
    if (args.copy_translation) { po[4] = pin[4]; po[5] = pin[5]; po[6] = pin[6]; }
    if (args.q_same) args.q_same[b] = 1;
  }
  double ev_final = 0.0;
  if (args.out_ev) {  // all four lanes of every group take part in the evaluation
    double g[3];
    ev_final = es_smallest_ev(mom, kEsLmPairs, x, sub, g);
  }
  if (active && sub == 0) {
    double *po = args.poses_out + 7 * b;
    const double sc = 1.0 / sqrt(1.0 + x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    po[0] = x[0] * sc; po[1] = x[1] * sc; po[2] = x[2] * sc; po[3] = sc;
    if (args.copy_translation) { po[4] = pin[4]; po[5] = pin[5]; po[6] = pin[6]; }
    if (args.q_same)
      args.q_same[b] = __double_as_longlong(po[0]) == __double_as_longlong(pin[0]) &&
                       __double_as_longlong(po[1]) == __double_as_longlong(pin[1]) &&
                       __double_as_longlong(po[2]) == __double_as_longlong(pin[2]) &&
                       __double_as_longlong(po[3]) == __double_as_longlong(pin[3]);
    if (args.out_info) args.out_info[b] = info;
    if (args.out_nfev) args.out_nfev[b] = nfev;
    if (args.out_ev) args.out_ev[b] = ev_final;
  }
}

// ------------------------------------------------------------ pipeline helpers

// Sophus::SE3d holds a unit quaternion: q <- q / |q| (PNEC::Solve with weighted_iterations_ == 0
// hands `initial_pose` to the refinement, pnec.cc:107-109).
__global__ void normalize_poses_kernel(const double *in, double *out, long long B) {
  const long long b = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double *p = in + 7 * b;
  const double inv = 1.0 / sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
  double *o = out + 7 * b;
  o[0] = p[0] * inv; o[1] = p[1] * inv; o[2] = p[2] * inv; o[3] = p[3] * inv;
  o[4] = p[4]; o[5] = p[5]; o[6] = p[6];
}

}  // namespace pnec
