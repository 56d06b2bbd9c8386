// pnec_eval.cuh — K1, the fused residual + analytic Jacobian + J^T J / J^T r / cost pass (the HBM-roofline kernel).
// Replaces one ceres::Problem evaluation = N x NumericDiffCostFunction<F, CENTRAL, 1,1,1,4>::Evaluate
// (src/optimization/pnec_ceres.cc:92-101).
#pragma once

#include "pnec_batch.cuh"

namespace pnec {

// ---------------------------------------------------------------- eval kernel

struct EvalArgs {
  BatchView bv;
  double reg;
  double *out_cost, *out_grad, *out_jtj;
  int use_bulk;
};

template <int V, int NW, int S, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) eval_kernel(const __grid_constant__ EvalArgs args) {
  constexpr int NT = NW * 32;
  constexpr int T = NT;  // correspondences per tile: one per thread
  constexpr int kStageDoubles = T * VariantTraits<V>::kDoubles;
  __shared__ __align__(8) uint64_t s_full[S];
  __shared__ PoseConst s_pc;
  __shared__ double s_part[NW][kAccPad];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = blockIdx.x;
  long long s, e;
  problem_range(args.bv, b, s, e);
  const int n = static_cast<int>(e - s);
  const long long g0 = s & ~1LL;
  const int head = static_cast<int>(s - g0);
  const int span = n + head;
  const int ntiles = (n > 0) ? (span + T - 1) / T : 0;

  auto stage_f1 = [&](int st) { return dyn_smem + st * kStageDoubles; };
  auto stage_f2 = [&](int st) { return dyn_smem + st * kStageDoubles + 3 * T; };
  auto stage_ct = [&](int st) { return dyn_smem + st * kStageDoubles + 6 * T; };
  auto stage_ch = [&](int st) { return dyn_smem + st * kStageDoubles + 15 * T; };
  auto issue_tile = [&](int k) {
    const int st = k % S;
    const int cnt = min(T, span - k * T);
    issue_bulk<V>(args.bv, g0 + static_cast<long long>(k) * T, cnt, stage_f1(st), stage_f2(st),
                  stage_ct(st), stage_ch(st), &s_full[st]);
  };

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < S; ++i) mbar_init(&s_full[i], 1);
    fence_mbar_init();
    const double *p = args.bv.poses + 7 * b;
    double x[6];
    angles_from_vec(p + 4, x[0], x[1]);
    x[2] = p[0]; x[3] = p[1]; x[4] = p[2]; x[5] = p[3];
    PoseConst pc0;
    make_pose_const(x, pc0);
    s_pc = pc0;
  }
  __syncthreads();
  if (tid == 0 && args.use_bulk) {
    for (int k = 0; k < min(S, ntiles); ++k) issue_tile(k);
  }
  PoseConst pc;
  load_pose_const(s_pc, pc);
  double acc[kNumAcc];
#pragma unroll
  for (int i = 0; i < kNumAcc; ++i) acc[i] = 0.0;

  for (int k = 0; k < ntiles; ++k) {
    const int st = k % S;
    if (args.use_bulk) {
      mbar_wait(&s_full[st], (k / S) & 1);
    } else {
      copy_plain<V, NT>(args.bv, g0 + static_cast<long long>(k) * T, min(T, span - k * T),
                        stage_f1(st), stage_f2(st), stage_ct(st), stage_ch(st), tid);
      __syncthreads();
    }
    const int i = k * T + tid;
    const bool valid = (i >= head) && (i < span);
    double a1[3], a2[3], c1[6], c2[6];
    if (valid) load_corr<V>(stage_f1(st), stage_f2(st), stage_ct(st), stage_ch(st), tid, a1, a2, c1, c2);
    __syncthreads();  // every thread holds its correspondence: the stage may be refilled
    if (tid == 0 && args.use_bulk && k + S < ntiles) issue_tile(k + S);
    if (valid) {
      double r, row[5];
      residual_row<V>(pc, args.reg, a1, a2, c1, c2, r, row);
      accumulate(acc, r, row);
    }
  }
  __shared__ double s_tot[kAccPad];
  block_reduce<NW>(acc, s_part, warp, lane, s_tot);
  if (warp == 0 && lane < kNumAcc) {
    const double mine = s_tot[lane];  // lane j writes value j
    if (lane < 15) {
      if (args.out_jtj) args.out_jtj[15 * b + lane] = mine;
    } else if (lane < 20) {
      if (args.out_grad) args.out_grad[5 * b + (lane - 15)] = mine;
    } else {
      if (args.out_cost) args.out_cost[b] = mine;
    }
  }
}

// ----------------------------------------------------- eval kernel, warp-private
//
// K1 as a persistent, barrier-free stream.  Every warp owns its frame pairs
// (problem gw, gw + W, gw + 2W, ...), its own S-stage ring of 32-correspondence
// tiles in shared memory and its own mbarriers; lane 0 keeps the ring S tiles ahead
// with bulk async copies and the ring runs across problem boundaries, so HBM
// requests never drain between problems.  No __syncthreads anywhere: the only
// cross-lane traffic is the transposing reduction once per problem.  Pose
// constants (acos/atan2/sincos, the expensive scalar part) are prepared
// lane-parallel for CHUNK problems at a time.
template <int V, int WPC, int S, int CHUNK, int MINB, int T>
__global__ void __launch_bounds__(WPC * 32, MINB)
eval_warp_kernel(const __grid_constant__ EvalArgs args) {
  static_assert(T % 32 == 0, "tiles are whole warps of correspondences");
  constexpr int kStageDoubles = T * VariantTraits<V>::kDoubles;
  __shared__ __align__(8) uint64_t s_full[WPC][S];
  __shared__ PoseConst s_pcs[WPC][CHUNK];

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

  // ---- producer (uniform across the warp; lane 0 issues): see WarpTileProducer
  WarpTileProducer<V, S, T> prod(args.bv, ring, full, gw, W, nmine, lane);
  prod.open();
#pragma unroll 1
  for (int i = 0; i < S; ++i) prod.issue();

  int c_stage = 0;
  uint32_t c_parity = 0;
  for (int j = 0; j < nmine; ++j) {
    if (j % CHUNK == 0) {
      __syncwarp();
      if (lane < CHUNK && j + lane < nmine) {
        const double *p = args.bv.poses + 7 * (gw + (j + lane) * W);
        double x[6];
        angles_from_vec(p + 4, x[0], x[1]);
        x[2] = p[0]; x[3] = p[1]; x[4] = p[2]; x[5] = p[3];
        PoseConst pc0;
        make_pose_const(x, pc0);
        s_pcs[warp][lane] = pc0;
      }
      __syncwarp();
    }
    const long long prob = gw + j * W;
    long long s, e;
    problem_range(args.bv, prob, s, e);
    const int head = static_cast<int>(s & 1LL);
    int c_left = (e > s) ? static_cast<int>(e - s) + head : 0;  // elements left, head included
    PoseConst pc;
    load_pose_const(s_pcs[warp][j % CHUNK], pc);
    double acc[kNumAcc];
#pragma unroll
    for (int i = 0; i < kNumAcc; ++i) acc[i] = 0.0;
    int lo = head;  // first valid element of the tile: head on the first tile, 0 afterwards
    while (c_left > 0) {
      mbar_wait(&full[c_stage], c_parity);
      const double *base = ring + c_stage * kStageDoubles;
      if (T == 32) {
        const bool valid = (lane >= lo) && (lane < c_left);
        double a1[3], a2[3], c1[6], c2[6];
        if (valid) load_corr<V>(base, base + 3 * T, base + 6 * T, base + 15 * T, lane, a1, a2, c1, c2);
        __syncwarp();  // every lane holds its correspondence: the stage may be refilled
        prod.issue();
        if (valid) {
          double r, row[5];
          residual_row<V>(pc, args.reg, a1, a2, c1, c2, r, row);
          accumulate(acc, r, row);
        }
      } else {
        // wider tiles (larger bulk copies): the warp walks the tile 32 correspondences at a time
#pragma unroll 1
        for (int sub = 0; sub < T; sub += 32) {
          const int i = sub + lane;
          if ((i >= lo) && (i < c_left)) {
            double a1[3], a2[3], c1[6], c2[6], r, row[5];
            load_corr<V>(base, base + 3 * T, base + 6 * T, base + 15 * T, i, a1, a2, c1, c2);
            residual_row<V>(pc, args.reg, a1, a2, c1, c2, r, row);
            accumulate(acc, r, row);
          }
        }
        __syncwarp();
        prod.issue();
      }
      if (++c_stage == S) {
        c_stage = 0;
        c_parity ^= 1u;
      }
      c_left -= T;
      lo = 0;
    }
    const double v = warp_transpose_reduce(acc, lane);
    const int idx = warp_reduce_owner_index(lane);
    if (idx >= 0 && idx < kNumAcc) {
      const double out = v * acc_scale(idx);
      if (idx < 15) {
        if (args.out_jtj) args.out_jtj[15 * prob + idx] = out;
      } else if (idx < 20) {
        if (args.out_grad) args.out_grad[5 * prob + (idx - 15)] = out;
      } else {
        if (args.out_cost) args.out_cost[prob] = out;
      }
    }
  }
}

}  // namespace pnec
