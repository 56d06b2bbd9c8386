// pnec_batch.cuh — batch views, correspondence loads, block reductions and bulk-copy (TMA) tile issue shared by the kernels.
#pragma once

#include "pnec_device.cuh"

namespace pnec {

// the one dynamic shared-memory window of every kernel (tile rings / resident correspondences)
extern __shared__ __align__(16) double dyn_smem[];


struct BatchView {
  const double *f1, *f2, *ct, *ch;
  const long long *offsets;  // device, B+1, or nullptr for uniform
  long long n_uniform, num_problems, total;
  const double *poses;  // [B][7]
  const int *counts;    // device, [B], or nullptr: problem b owns only the first counts[b] elements of its
                        // range (inliers compacted to the front of the pair's slot after RANSAC)
};

__device__ __forceinline__ void problem_range(const BatchView &bv, long long b, long long &s,
                                              long long &e) {
  if (bv.offsets) {
    s = bv.offsets[b];
    e = bv.offsets[b + 1];
  } else {
    s = b * bv.n_uniform;
    e = s + bv.n_uniform;
  }
  if (bv.counts) e = s + bv.counts[b];
}

template <int V>
struct VariantTraits {
  static constexpr bool kHasCt = (V != PNEC_VARIANT_NEC);
  static constexpr bool kHasCh = (V == PNEC_VARIANT_SYMMETRIC);
  // doubles per correspondence as laid out in HBM (the algorithmic bytes / 8)
  static constexpr int kDoubles = 6 + (kHasCt ? 9 : 0) + (kHasCh ? 9 : 0);
};

// One correspondence from raw (f1, f2, ct, ch) arrays at element index i; the 3x3 covariances
// are reduced to their packed symmetric part on the way in.
template <int V>
__device__ __forceinline__ void load_corr(const double *f1, const double *f2, const double *ct,
                                          const double *ch, long long i, double a1[3],
                                          double a2[3], double s1[6], double s2[6]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) a1[k] = f1[3 * i + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) a2[k] = f2[3 * i + k];
  if (VariantTraits<V>::kHasCt) {
    double c[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = ct[9 * i + k];
    pack_sym(c, s1);
  }
  if (VariantTraits<V>::kHasCh) {
    double c[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = ch[9 * i + k];
    pack_sym(c, s2);
  }
}

template <int V, int NT>
__device__ __forceinline__ void eval_pass(const PoseConst &pc, double reg, const double *f1,
                                          const double *f2, const double *ct, const double *ch,
                                          int begin, int end, int tid, double acc[kNumAcc]) {
  for (int i = begin + tid; i < end; i += NT) {
    double a1[3], a2[3], c1[6], c2[6], r, row[5];
    load_corr<V>(f1, f2, ct, ch, i, a1, a2, c1, c2);
    residual_row<V>(pc, reg, a1, a2, c1, c2, r, row);
    accumulate(acc, r, row);
  }
}

template <int V, int NT>
__device__ __forceinline__ double eval_pass_cost(const PoseConst &pc, double reg, const double *f1,
                                                 const double *f2, const double *ct,
                                                 const double *ch, int begin, int end, int tid) {
  double sum = 0.0;
  for (int i = begin + tid; i < end; i += NT) {
    double a1[3], a2[3], c1[6], c2[6];
    load_corr<V>(f1, f2, ct, ch, i, a1, a2, c1, c2);
    const double r = residual_only<V>(pc, reg, a1, a2, c1, c2);
    sum = fma(r, r, sum);
  }
  return sum;
}

// Block reduction of one value (cost passes); result valid in warp 0.  One __syncthreads().
template <int NW>
__device__ __forceinline__ double block_reduce_scalar(double v, double (*s_part)[kAccPad], int warp,
                                                      int lane, int owner = 0) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) s_part[warp][kNumAcc - 1] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == owner) {
#pragma unroll
    for (int w = 0; w < NW; ++w) t += s_part[w][kNumAcc - 1];
    t *= 0.5;
  }
  return t;
}

// Block reduction of the 21 partial sums into dst[0..20] (shared memory), scaled; valid for
// warp 0 after the call (other warps must pass a barrier first).  One __syncthreads().
template <int NW>
__device__ __forceinline__ void block_reduce(const double acc[kNumAcc],
                                             double (*s_part)[kAccPad], int warp, int lane,
                                             double *dst, int owner = 0) {
  const double v = warp_transpose_reduce(acc, lane);
  const int idx = warp_reduce_owner_index(lane);
  if (idx >= 0 && idx < kNumAcc) s_part[warp][idx] = v;
  __syncthreads();
  if (warp == owner) {
    if (lane < kNumAcc) {
      double mine = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) mine += s_part[w][lane];
      dst[lane] = mine * acc_scale(lane);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void load_pose_const(const PoseConst &src, PoseConst &dst) {
#pragma unroll
  for (int i = 0; i < 9; ++i) dst.R[i] = src.R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dst.t[i] = src.t[i];
    dst.tth[i] = src.tth[i];
    dst.tph[i] = src.tph[i];
  }
}

// Issues the bulk copies of `cnt` correspondences starting at global element
// `first` (even) into the arrays at sf1/sf2/sct/sch, completing on `bar`.
// Called by ONE thread.  Sizes are rounded up to an even element count (16-byte
// granularity of cp.async.bulk); if that would run past the end of the batch the
// last element is copied with plain loads instead.
template <int V>
__device__ __forceinline__ void issue_bulk(const BatchView &bv, long long first, int cnt,
                                           double *sf1, double *sf2, double *sct, double *sch,
                                           uint64_t *bar) {
  int cb = cnt + (cnt & 1);
  if (first + cb > bv.total) {
    cb = cnt - 1;  // cnt is odd here
    const long long g = first + cb;
#pragma unroll
    for (int k = 0; k < 3; ++k) sf1[3 * cb + k] = bv.f1[3 * g + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) sf2[3 * cb + k] = bv.f2[3 * g + k];
    if (VariantTraits<V>::kHasCt) {
#pragma unroll
      for (int k = 0; k < 9; ++k) sct[9 * cb + k] = bv.ct[9 * g + k];
    }
    if (VariantTraits<V>::kHasCh) {
#pragma unroll
      for (int k = 0; k < 9; ++k) sch[9 * cb + k] = bv.ch[9 * g + k];
    }
  }
  const uint32_t bytes = static_cast<uint32_t>(cb) * 8u * VariantTraits<V>::kDoubles;
  mbar_arrive_expect_tx(bar, bytes);
  if (cb > 0) {
    bulk_g2s(sf1, bv.f1 + 3 * first, cb * 24u, bar);
    bulk_g2s(sf2, bv.f2 + 3 * first, cb * 24u, bar);
    if (VariantTraits<V>::kHasCt) bulk_g2s(sct, bv.ct + 9 * first, cb * 72u, bar);
    if (VariantTraits<V>::kHasCh) bulk_g2s(sch, bv.ch + 9 * first, cb * 72u, bar);
  }
}

// Producer side of a warp-private ring of S tiles of T correspondences (shared by K1 and the moments
// kernel): the warp owns problems gw, gw + W, gw + 2W, ... and streams their correspondences tile by
// tile with bulk async copies; the ring runs ACROSS problem boundaries, so HBM requests never drain
// between problems.  All members are warp-uniform; lane 0 issues.  Tiles start at even global
// correspondence indices so that every byte range is 16-byte aligned for any ragged offset.
template <int V, int S, int T>
struct WarpTileProducer {
  static constexpr int kStageDoubles = T * VariantTraits<V>::kDoubles;
  static constexpr bool kCt = VariantTraits<V>::kHasCt, kCh = VariantTraits<V>::kHasCh;
  const BatchView &bv;
  double *ring;
  uint64_t *full;
  long long gw, W;
  int nmine, lane;
  int pj = 0, p_left = 0, p_stage = 0;
  const double *p_f1 = nullptr, *p_f2 = nullptr, *p_ct = nullptr, *p_ch = nullptr;

  __device__ __forceinline__ WarpTileProducer(const BatchView &bv_, double *ring_, uint64_t *full_, long long gw_,
                                              long long W_, int nmine_, int lane_)
      : bv(bv_), ring(ring_), full(full_), gw(gw_), W(W_), nmine(nmine_), lane(lane_) {}

  // position on the first tile of the next non-empty problem
  __device__ __forceinline__ void open() {
    while (pj < nmine) {
      long long s, e;
      problem_range(bv, gw + pj * W, s, e);
      if (e > s) {
        const long long g0 = s & ~1LL;  // even => 16-byte aligned in every array
        p_left = static_cast<int>(e - g0);
        p_f1 = bv.f1 + 3 * g0;
        p_f2 = bv.f2 + 3 * g0;
        if (kCt) p_ct = bv.ct + 9 * g0;
        if (kCh) p_ch = bv.ch + 9 * g0;
        return;
      }
      ++pj;
    }
  }
  // issue the current tile, then advance
  __device__ __forceinline__ void issue() {
    if (pj >= nmine) return;
    const int cnt = min(T, p_left);
    if (lane == 0) {
      double *base = ring + p_stage * kStageDoubles;
      int cb = cnt + (cnt & 1);  // bulk copies move 16-byte units: round up to an even count ...
      if ((cnt & 1) && p_f1 + 3 * cb > bv.f1 + 3 * bv.total) {
        cb = cnt - 1;  // ... unless that runs past the end of the batch: last element by hand
#pragma unroll
        for (int k = 0; k < 3; ++k) base[3 * cb + k] = p_f1[3 * cb + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) base[3 * T + 3 * cb + k] = p_f2[3 * cb + k];
        if (kCt) {
#pragma unroll
          for (int k = 0; k < 9; ++k) base[6 * T + 9 * cb + k] = p_ct[9 * cb + k];
        }
        if (kCh) {
#pragma unroll
          for (int k = 0; k < 9; ++k) base[15 * T + 9 * cb + k] = p_ch[9 * cb + k];
        }
      }
      mbar_arrive_expect_tx(&full[p_stage], static_cast<uint32_t>(cb) * 8u * VariantTraits<V>::kDoubles);
      if (cb > 0) {
        bulk_g2s(base, p_f1, cb * 24u, &full[p_stage]);
        bulk_g2s(base + 3 * T, p_f2, cb * 24u, &full[p_stage]);
        if (kCt) bulk_g2s(base + 6 * T, p_ct, cb * 72u, &full[p_stage]);
        if (kCh) bulk_g2s(base + 15 * T, p_ch, cb * 72u, &full[p_stage]);
      }
    }
    p_stage = (p_stage + 1 == S) ? 0 : p_stage + 1;
    p_left -= T;
    if (p_left > 0) {
      p_f1 += 3 * T;
      p_f2 += 3 * T;
      if (kCt) p_ct += 9 * T;
      if (kCh) p_ch += 9 * T;
    } else {
      ++pj;
      open();
    }
  }
};

// Same region, plain cooperative loads (unaligned base pointers).
template <int V, int NT>
__device__ __forceinline__ void copy_plain(const BatchView &bv, long long first, int cnt,
                                           double *sf1, double *sf2, double *sct, double *sch,
                                           int tid) {
  for (int j = tid; j < cnt * 3; j += NT) {
    sf1[j] = bv.f1[3 * first + j];
    sf2[j] = bv.f2[3 * first + j];
  }
  if (VariantTraits<V>::kHasCt)
    for (int j = tid; j < cnt * 9; j += NT) sct[j] = bv.ct[9 * first + j];
  if (VariantTraits<V>::kHasCh)
    for (int j = tid; j < cnt * 9; j += NT) sch[j] = bv.ch[9 * first + j];
}

}  // namespace pnec
