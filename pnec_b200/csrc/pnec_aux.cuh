// pnec_aux.cuh — parity metric (pnec::common::CostFunction) and unscented transform kernels.
#pragma once

#include "pnec_batch.cuh"

namespace pnec {

// ---------------------------------------------------------------- cost kernel
// pnec::common::CostFunction, src/common/common.cc:237-259: mean of
// (t^T (f1 x R f2))^2 / (b^T S b), no regularisation.  pose: unit quaternion taken
// from the normalised stored quaternion, translation as stored.
__global__ void __launch_bounds__(128) cost_kernel(BatchView bv, double *out) {
  __shared__ double s_red[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = blockIdx.x;
  long long s, e;
  problem_range(bv, b, s, e);
  const double *p = bv.poses + 7 * b;
  const double qn = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
  double x[6] = {0.0, 0.0, p[0] / qn, p[1] / qn, p[2] / qn, p[3] / qn};
  PoseConst pc;
  make_pose_const(x, pc);
  const double t[3] = {p[4], p[5], p[6]};
  double sum = 0.0;
  for (long long i = s + tid; i < e; i += 128) {
    double f1[3], f2[3], c[9], g[3], a[3], bb[3], Sb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { f1[k] = bv.f1[3 * i + k]; f2[k] = bv.f2[3 * i + k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = bv.ct[9 * i + k];
    rot(pc.R, f2, g);
    cross3(t, f1, a);
    const double num = dot3(a, g);
    rot_t(pc.R, a, bb);
    sym_mul(c, bb, Sb);
    sum += num * num / dot3(bb, Sb);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  if (tid == 0) out[b] = (s_red[0] + s_red[1] + s_red[2] + s_red[3]) / static_cast<double>(e - s);
}

// ------------------------------------------------------- unscented transform
// pnec::common::UnscentedTransform, src/common/common.cc:467-525, one thread per point.
// 96 B in + 72 B out per point, ~250 flops: HBM-bound.
struct UtArgs {
  const double *mus, *covs;
  double *out;
  long long n;
  double Kinv[9];  // column-major
  double kappa;
  int camera_model;
  int use_bulk;  // mus / covs / out 16-byte aligned
};

// The transform of one point: mu[3], S[9] (column-major) -> out[9] (column-major).
__device__ __forceinline__ void unscented_point(const UtArgs &a, const double mu[3], const double S[9],
                                                double out[9]) {
  const bool omni = (a.camera_model == PNEC_CAMERA_OMNIDIRECTIONAL);
  // C = [c0 c1 0]: image-plane Cholesky columns (rotated for omnidirectional cameras)
  double c0[3], c1[3];
  if (omni) {
    // RotationBetweenPoints((0,0,1), mu.normalized()), common.cc:118-124
    const double inv_n = 1.0 / sqrt(mu[0] * mu[0] + mu[1] * mu[1] + mu[2] * mu[2]);
    const double m[3] = {mu[0] * inv_n, mu[1] * inv_n, mu[2] * inv_n};
    const double v[3] = {-m[1], m[0], 0.0};  // z x m
    const double k = 1.0 / (1.0 + m[2]);
    // R = I + [v]x + [v]x^2 k, row-major
    double R[9];
    R[0] = 1.0 + (-v[1] * v[1]) * k; R[1] = (v[0] * v[1]) * k;           R[2] = v[1];
    R[3] = (v[0] * v[1]) * k;        R[4] = 1.0 + (-v[0] * v[0]) * k;    R[5] = -v[0];
    R[6] = -v[1];                    R[7] = v[0];                        R[8] = 1.0 + (-(v[0] * v[0] + v[1] * v[1])) * k;
    // local = (R^T S R) top-left 2x2:  local[p][q] = sum_rc R[r][p] S[r][c] R[c][q]
    double SR[3][2];  // (S R)[:, 0:2]
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 2; ++q)
        SR[r][q] = S[0 * 3 + r] * R[0 * 3 + q] + S[1 * 3 + r] * R[1 * 3 + q] + S[2 * 3 + r] * R[2 * 3 + q];
    const double l_00 = R[0] * SR[0][0] + R[3] * SR[1][0] + R[6] * SR[2][0];
    const double l_10 = R[1] * SR[0][0] + R[4] * SR[1][0] + R[7] * SR[2][0];
    const double l_11 = R[1] * SR[0][1] + R[4] * SR[1][1] + R[7] * SR[2][1];
    const double L00 = sqrt(l_00), L10 = l_10 / L00, L11 = sqrt(l_11 - L10 * L10);
    // C = R * [[L00,0,0],[L10,L11,0],[0,0,0]]
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      c0[r] = R[r * 3 + 0] * L00 + R[r * 3 + 1] * L10;
      c1[r] = R[r * 3 + 1] * L11;
    }
  } else {
    const double L00 = sqrt(S[0]), L10 = S[1] / L00, L11 = sqrt(S[4] - L10 * L10);
    c0[0] = L00; c0[1] = L10; c0[2] = 0.0;
    c1[0] = 0.0; c1[1] = L11; c1[2] = 0.0;
  }
  const double nk = static_cast<double>(2.0f) + a.kappa;  // (float)n + kappa, common.cc:495
  const double w0 = a.kappa / nk, wi = 0.5 / nk;
  double tp[5][3], mean[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int p = 0; p < 5; ++p) {
    const double sg = (p == 0) ? 0.0 : ((p <= 2) ? 1.0 : -1.0);
    const double *c = (p == 1 || p == 3) ? c0 : c1;
    double q[3] = {mu[0] + sg * c[0], mu[1] + sg * c[1], mu[2] + sg * c[2]};
    if (!omni) {
      const double x = a.Kinv[0] * q[0] + a.Kinv[3] * q[1] + a.Kinv[6] * q[2];
      const double y = a.Kinv[1] * q[0] + a.Kinv[4] * q[1] + a.Kinv[7] * q[2];
      const double z = a.Kinv[2] * q[0] + a.Kinv[5] * q[1] + a.Kinv[8] * q[2];
      q[0] = x; q[1] = y; q[2] = z;
    }
    const double inv = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const double w = (p == 0) ? w0 : wi;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      tp[p][k] = q[k] * inv;
      mean[k] += w * tp[p][k];
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) out[k] = 0.0;
#pragma unroll
  for (int p = 0; p < 5; ++p) {
    const double w = (p == 0) ? w0 : wi;
    const double d[3] = {tp[p][0] - mean[0], tp[p][1] - mean[1], tp[p][2] - mean[2]};
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) out[c * 3 + r] += w * d[r] * d[c];
  }
}

// One CTA per tile of 128 points.  Full, 16-byte aligned tiles go through the TMA engine both
// ways: two bulk loads (3 KB of points, 9 KB of covariances) into shared memory, each thread
// transforms its point, results are staged over the covariance tile and leave as one 9 KB
// bulk store.  The last partial tile (and unaligned arrays) use plain per-thread accesses.
__global__ void __launch_bounds__(128) unscented_kernel(const __grid_constant__ UtArgs a) {
  constexpr int T = 128;
  __shared__ __align__(16) double s_mu[3 * T];
  __shared__ __align__(16) double s_cov[9 * T];
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x;
  const long long first = static_cast<long long>(blockIdx.x) * T;
  const int cnt = static_cast<int>(min(static_cast<long long>(T), a.n - first));
  double mu[3], S[9], out[9];
  if (cnt == T && a.use_bulk) {
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      fence_mbar_init();
      mbar_arrive_expect_tx(&s_bar, T * 96u);
      bulk_g2s(s_mu, a.mus + 3 * first, T * 24u, &s_bar);
      bulk_g2s(s_cov, a.covs + 9 * first, T * 72u, &s_bar);
    }
    __syncthreads();
    mbar_wait(&s_bar, 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) mu[k] = s_mu[3 * tid + k];
#pragma unroll
    for (int k = 0; k < 9; ++k) S[k] = s_cov[9 * tid + k];
    unscented_point(a, mu, S, out);
    __syncthreads();  // every thread has read its covariance: the tile can be overwritten
#pragma unroll
    for (int k = 0; k < 9; ++k) s_cov[9 * tid + k] = out[k];
    fence_proxy_async();  // generic-proxy writes -> visible to the bulk (async-proxy) store
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(a.out + 9 * first, s_cov, T * 72u);
      bulk_commit_group();
      bulk_wait_group_read0();  // shared memory must outlive the read
    }
  } else if (tid < cnt) {
    const long long i = first + tid;
#pragma unroll
    for (int k = 0; k < 3; ++k) mu[k] = a.mus[3 * i + k];
#pragma unroll
    for (int k = 0; k < 9; ++k) S[k] = a.covs[9 * i + k];
    unscented_point(a, mu, S, out);
#pragma unroll
    for (int k = 0; k < 9; ++k) a.out[9 * i + k] = out[k];
  }
}

// KeyPoint::Unproject (src/frames/keypoints.cc:49-62) for n keypoints: one thread per keypoint,
// 48 B in, 96 B out; full tiles of 128 keypoints move through the TMA engine both ways.
struct KpArgs {
  const double *points, *covs2;
  double *out_bvs, *out_covs;
  long long n;
  double Kinv[9];  // column-major
  int use_bulk;
};

// Out of line: ONE body for keypoint_kernel and keypoint_assemble_kernel, so that solver inputs
// built from keypoints on the way in are bit-identical to those of pnec_keypoints_unproject_batch.
__device__ __noinline__ void keypoint_unproject_kinv(const double *Kinv, double px, double py, double c0, double c1,
                                                     double c2v, double c3, double *bv, double *out) {
  // pnec::common::Unproject: (K_inv (x, y, 1)).normalized()
  const double x = Kinv[0] * px + Kinv[3] * py + Kinv[6];
  const double y = Kinv[1] * px + Kinv[4] * py + Kinv[7];
  const double z = Kinv[2] * px + Kinv[5] * py + Kinv[8];
  const double inv = 1.0 / sqrt(x * x + y * y + z * z);
  bv[0] = x * inv; bv[1] = y * inv; bv[2] = z * inv;
  if (!out) return;
  UtArgs u{};
#pragma unroll
  for (int k = 0; k < 9; ++k) u.Kinv[k] = Kinv[k];
  u.kappa = 1.0;
  u.camera_model = PNEC_CAMERA_PINHOLE;
  const double mu[3] = {px, py, 1.0};
  const double S[9] = {c0, c1, 0.0, c2v, c3, 0.0, 0.0, 0.0, 0.0};  // column-major
  double o9[9];
  unscented_point(u, mu, S, o9);
#pragma unroll
  for (int k = 0; k < 9; ++k) out[k] = o9[k];
}

__device__ __forceinline__ void keypoint_unproject(const KpArgs &a, const double pt[2],
                                                   const double c2[4], double bv[3], double out[9]) {
  keypoint_unproject_kinv(a.Kinv, pt[0], pt[1], c2[0], c2[1], c2[2], c2[3], bv, out);
}

__global__ void __launch_bounds__(128) keypoint_kernel(const __grid_constant__ KpArgs a) {
  constexpr int T = 128;
  __shared__ __align__(16) double s_in[6 * T];    // points [2T] | covs2 [4T]
  __shared__ __align__(16) double s_out[12 * T];  // bvs [3T] | covs [9T]
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x;
  const long long first = static_cast<long long>(blockIdx.x) * T;
  const int cnt = static_cast<int>(min(static_cast<long long>(T), a.n - first));
  double pt[2], c2[4], bv[3], out[9];
  if (cnt == T && a.use_bulk) {
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      fence_mbar_init();
      mbar_arrive_expect_tx(&s_bar, T * 48u);
      bulk_g2s(s_in, a.points + 2 * first, T * 16u, &s_bar);
      bulk_g2s(s_in + 2 * T, a.covs2 + 4 * first, T * 32u, &s_bar);
    }
    __syncthreads();
    mbar_wait(&s_bar, 0);
    pt[0] = s_in[2 * tid]; pt[1] = s_in[2 * tid + 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) c2[k] = s_in[2 * T + 4 * tid + k];
    keypoint_unproject(a, pt, c2, bv, out);
#pragma unroll
    for (int k = 0; k < 3; ++k) s_out[3 * tid + k] = bv[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) s_out[3 * T + 9 * tid + k] = out[k];
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(a.out_bvs + 3 * first, s_out, T * 24u);
      bulk_s2g(a.out_covs + 9 * first, s_out + 3 * T, T * 72u);
      bulk_commit_group();
      bulk_wait_group_read0();
    }
  } else if (tid < cnt) {
    const long long i = first + tid;
    pt[0] = a.points[2 * i]; pt[1] = a.points[2 * i + 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) c2[k] = a.covs2[4 * i + k];
    keypoint_unproject(a, pt, c2, bv, out);
#pragma unroll
    for (int k = 0; k < 3; ++k) a.out_bvs[3 * i + k] = bv[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) a.out_covs[9 * i + k] = out[k];
  }
}

// PNEC_COV_PACKED -> the 3x3 layout the kernels stream (pack_sym of the result returns the input bits)
__global__ void __launch_bounds__(256) expand_covs_kernel(const double *packed, double *full, long long first,
                                                          long long count) {
  const long long i = first + static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= first + count) return;
  const double *s = packed + 6 * i;
  double *c = full + 9 * i;
  const double xx = s[0], xy = s[1], xz = s[2], yy = s[3], yz = s[4], zz = s[5];
  c[0] = xx; c[1] = xy; c[2] = xz; c[3] = xy; c[4] = yy; c[5] = yz; c[6] = xz; c[7] = yz; c[8] = zz;
}

// Frame2Frame::GetFeatures (src/rel_pose_estimation/frame2frame.cc:359-392) on the device: the solver's
// inputs of `total` correspondences from the keypoints of the two frames -- pixel position and 2x2
// image covariance, the fields a pnec::features::KeyPoint is constructed from (keypoints.cc:42-62) --
// selected by the matches' query / train indices.  64 B per correspondence cross the host link
// instead of 120.
struct KpAssembleArgs {
  const double *host_points, *target_points;   // [Kh][2], [Kt][2]
  const double *host_covs2, *target_covs2;     // [K][4] column-major 2x2, or [K][3] (xx, xy, yy) when packed
  const int *host_index, *target_index;        // [total] rows of the two tables, or nullptr: row i
  double *f1, *f2, *ct, *ch;                   // [total][3], [total][3], [total][9], [total][9] (ct / ch may be null)
  long long first, count;                      // correspondences [first, first + count)
  double Kinv[9];
  int packed;
};

__global__ void __launch_bounds__(128) keypoint_assemble_kernel(const __grid_constant__ KpAssembleArgs a) {
  const long long i = a.first + static_cast<long long>(blockIdx.x) * 128 + threadIdx.x;
  if (i >= a.first + a.count) return;
  const long long hi = a.host_index ? a.host_index[i] : i, ti = a.target_index ? a.target_index[i] : i;
  auto cov = [&](const double *tab, long long r, double c[4]) {
    if (a.packed) { c[0] = tab[3 * r]; c[1] = tab[3 * r + 1]; c[2] = c[1]; c[3] = tab[3 * r + 2]; }
    else { c[0] = tab[4 * r]; c[1] = tab[4 * r + 1]; c[2] = tab[4 * r + 2]; c[3] = tab[4 * r + 3]; }
  };
  double bv[3], o9[9], c[4] = {0, 0, 0, 0};
  if (a.ch) cov(a.host_covs2, hi, c);
  keypoint_unproject_kinv(a.Kinv, a.host_points[2 * hi], a.host_points[2 * hi + 1], c[0], c[1], c[2], c[3], bv,
                          a.ch ? o9 : nullptr);
#pragma unroll
  for (int k = 0; k < 3; ++k) a.f1[3 * i + k] = bv[k];
  if (a.ch) {
#pragma unroll
    for (int k = 0; k < 9; ++k) a.ch[9 * i + k] = o9[k];
  }
  if (a.ct) cov(a.target_covs2, ti, c);
  keypoint_unproject_kinv(a.Kinv, a.target_points[2 * ti], a.target_points[2 * ti + 1], c[0], c[1], c[2], c[3], bv,
                          a.ct ? o9 : nullptr);
#pragma unroll
  for (int k = 0; k < 3; ++k) a.f2[3 * i + k] = bv[k];
  if (a.ct) {
#pragma unroll
    for (int k = 0; k < 9; ++k) a.ct[9 * i + k] = o9[k];
  }
}

}  // namespace pnec
