// pnec_solve_slots.cuh — the LM solve of PNECCeres::Optimize (src/optimization/pnec_ceres.cc:70-168) for
// large batches of frame pairs that fit shared memory: evaluation and LM update decoupled.
//
// solve_kernel gives every frame pair a CTA whose warps evaluate together and then wait at a barrier
// while one lane runs the serial LM update: half of a pair's life its evaluation warps are parked, and
// shared memory (not warps) caps the pairs resident per SM.  Here a CTA owns P pair SLOTS in shared
// memory, one OWNER warp per slot and one group of evaluation warps for all of them:
//
//   owner warp of slot s   claims a pair (atomic work counter, one pair ahead), bulk-copies it into the slot,
//                          repacks the 3x3 covariances to their symmetric part IN PLACE (96 instead of 120 B
//                          per correspondence: four slots per SM at N = 512 instead of three), then loops
//                          { wait for the 4 partial sums -> LM update, which posts the next candidate's pose
//                          constants as soon as they are stored } and writes the result;
//   evaluation warps       take the posted candidates in order, evaluate their quarter of the slot (residual +
//                          Jacobian row + J^T J / J^T r, or cost only) and hand the partial sums to the owner.
//                          They never wait for an LM update: while slot A is in its update they evaluate
//                          slot B.
//
// Hand-over is by mbarriers, no __syncthreads after start-up and no polling: an owner posts a candidate
// as a ticket in a small ring (one mbarrier per ring entry, tickets numbered by an atomic counter in
// shared memory), the evaluation warps sleep in try_wait on the next ticket and so all work through the
// posted candidates in the same order; the owner sleeps on its slot's "partial sums ready" barrier
// (4 arrivals).  A pair's correspondences are assigned to lanes, and its partial sums
// added, exactly as solve_kernel<V, 4, .> does: results are bit-identical to that kernel.
#pragma once

#include "pnec_solve.cuh"

namespace pnec {

constexpr int kSlotEvalWarps = 4;  // evaluation warps per CTA (== the cross-warp split of solve_kernel<V, 4, .>)
constexpr int kSlotTickets = 8;    // ring entries; a slot has at most one ticket outstanding, so P <= 8 never laps a reader

// The hand-over from the owners to the evaluation warps.
struct SlotTickets {
  unsigned long long bar[kSlotTickets];  // mbarriers, 1 arrival: ticket posted
  int slot[kSlotTickets];        // -1: no more work (posted once per evaluation group by the last owner to leave)
  unsigned int tail;
  unsigned int owners_gone;
};
// Called by ONE lane of an owner: everything it wrote before (candidate, mode) is released to the readers.
__device__ __forceinline__ void slot_post(SlotTickets &tk, int slot) {
  const unsigned int t = atomicAdd(&tk.tail, 1u);
  tk.slot[t & (kSlotTickets - 1)] = slot;
  mbar_arrive(reinterpret_cast<uint64_t *>(&tk.bar[t & (kSlotTickets - 1)]));
}

// Slot layout, in doubles, for `cap` correspondences (cap a multiple of 32):
//   [K packed covariance areas, 6 cap each][f1: 3 cap][f2: 3 cap]
// A packed covariance area is a sequence of 32-correspondence chunks [32 x (xx, xy, xz)][32 x (yy, yz, zz)]:
// stride-3 accesses, free of bank conflicts (a plain stride of 6 doubles would be 2-way conflicted).
template <int V>
struct SlotLayout {
  static constexpr int kCov = (VariantTraits<V>::kHasCt ? 1 : 0) + (VariantTraits<V>::kHasCh ? 1 : 0);
  static constexpr int kDoubles = 6 + 6 * kCov;
};

#ifdef PNEC_SLOT_TIMING
// debug build: cycles summed over the batch: 0 load trip 1, 1 repack, 2 load trip 2, 3 owner waits for the
// evaluation, 4 LM update, 5 pair total, 6 evaluation warps busy, 7 evaluation warps polling, 8 passes, 9 init
__device__ unsigned long long g_slot_probe[16];
// (accumulated in registers, flushed once per warp: an atomic per probe costs hundreds of cycles)
#define SLOT_PROBE(k) do { const long long t_now = clock64(); probe_acc[k] += t_now - t_prev; t_prev = t_now; } while (0)
#define SLOT_PROBE_START() t_prev = clock64()
#define SLOT_PROBE_DECL() long long probe_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long t_prev = 0; const long long t_birth = clock64()
#define SLOT_PROBE_FLUSH() do { if (lane == 0) { probe_acc[12 + (warp_role_)] += clock64() - t_birth; for (int k_ = 0; k_ < 16; ++k_) atomicAdd(&g_slot_probe[k_], (unsigned long long)probe_acc[k_]); } } while (0)
#else
#define SLOT_PROBE(k) do { } while (0)
#define SLOT_PROBE_START() do { } while (0)
#define SLOT_PROBE_DECL() do { } while (0)
#define SLOT_PROBE_FLUSH() do { } while (0)
#endif

struct SlotCtl {
  LMState lm;
  PoseConst pc;
  double part[kSlotEvalWarps][kAccPad];
  int head, span, mode, pad;
};

// L2 prefetch of a byte range (the next pair of a slot, while the current one is being solved)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

template <int V>
__device__ __forceinline__ void slot_load_corr(const double *base, int cap, int j, double a1[3], double a2[3],
                                               double c1[6], double c2[6]) {
  constexpr int K = SlotLayout<V>::kCov;
  const double *f1 = base + 6 * K * cap + 3 * j;
  const double *f2 = f1 + 3 * cap;
#pragma unroll
  for (int k = 0; k < 3; ++k) a1[k] = f1[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) a2[k] = f2[k];
  const int off = 6 * (j & ~31) + 3 * (j & 31);
  if (K >= 1) {
    const double *p = base + off;
    c1[0] = p[0]; c1[1] = p[1]; c1[2] = p[2]; c1[3] = p[96]; c1[4] = p[97]; c1[5] = p[98];
  }
  if (K == 2) {
    const double *p = base + 6 * cap + off;
    c2[0] = p[0]; c2[1] = p[1]; c2[2] = p[2]; c2[3] = p[96]; c2[4] = p[97]; c2[5] = p[98];
  }
}

// One array of a slot load: `cb` (even) correspondences of `dpc` doubles each by bulk copy, plus the odd
// last element of the whole batch by hand (cp.async.bulk moves 16-byte units and must not run past the end).
__device__ __forceinline__ void slot_copy_array(double *dst, const double *src, int cb, bool tail, int dpc,
                                                uint64_t *bar) {
  if (tail) {
    for (int k = 0; k < dpc; ++k) dst[dpc * cb + k] = src[dpc * cb + k];
  }
  if (cb > 0) bulk_g2s(dst, src, static_cast<uint32_t>(cb) * dpc * 8u, bar);
}

// Trip 1 of a slot load: the raw 3x3 covariances land at the front of the slot (72 B per correspondence
// over what will be the packed covariances and f1), f2 goes to its final place when that is free.
// Trip 2 (after the in-place repack): what trip 1 could not place.  Called by ONE lane.
template <int V>
__device__ __forceinline__ void slot_issue(const BatchView &bv, long long g0, int span, int cap, double *base,
                                           uint64_t *bar, int trip) {
  constexpr int K = SlotLayout<V>::kCov;
  int cb = span + (span & 1);
  bool tail = false;
  if (g0 + cb > bv.total) {
    cb = span - 1;
    tail = true;
  }
  double *sf1 = base + 6 * K * cap, *sf2 = sf1 + 3 * cap;
  const bool f1_now = (K == 0) ? (trip == 1) : (trip == 2);
  const bool f2_now = (K == 1) ? (trip == 1) : f1_now;
  const bool cov_now = (K > 0) && trip == 1;
  const uint32_t bytes = static_cast<uint32_t>(cb) * 8u * ((f1_now ? 3 : 0) + (f2_now ? 3 : 0) + (cov_now ? 9 * K : 0));
  mbar_arrive_expect_tx(bar, bytes);
  if (cov_now) {
    slot_copy_array(base, bv.ct + 9 * g0, cb, tail, 9, bar);
    if (K == 2) slot_copy_array(base + 9 * cap, bv.ch + 9 * g0, cb, tail, 9, bar);
  }
  if (f1_now) slot_copy_array(sf1, bv.f1 + 3 * g0, cb, tail, 3, bar);
  if (f2_now) slot_copy_array(sf2, bv.f2 + 3 * g0, cb, tail, 3, bar);
}

// The next pair of a slot on its way into L2 while the current one is being solved.
template <int V>
__device__ __forceinline__ void slot_prefetch(const BatchView &bv, long long b) {
  long long s, e;
  problem_range(bv, b, s, e);
  const long long g0 = s & ~1LL;
  const long long cnt = (e - g0) & ~1LL;  // whole 16-byte units inside the batch
  if (cnt <= 0) return;
  bulk_prefetch_l2(bv.f1 + 3 * g0, static_cast<uint32_t>(cnt) * 24u);
  bulk_prefetch_l2(bv.f2 + 3 * g0, static_cast<uint32_t>(cnt) * 24u);
  if (VariantTraits<V>::kHasCt) bulk_prefetch_l2(bv.ct + 9 * g0, static_cast<uint32_t>(cnt) * 72u);
  if (VariantTraits<V>::kHasCh) bulk_prefetch_l2(bv.ch + 9 * g0, static_cast<uint32_t>(cnt) * 72u);
}

// In-place repack of one raw covariance area (9 doubles per correspondence from `src`) to the chunked
// symmetric layout at `dst` <= src, by one warp, two 32-correspondence waves per turn.  Wave k writes
// [192 k, 192 (k + 1)) and later waves read from 288 (k + 1) on: a wave never overwrites unread input.
__device__ __forceinline__ void slot_repack(double *dst, const double *src, int span, int lane) {
  for (int j0 = 0; j0 < span; j0 += 64) {
    double p[2][6];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (j0 + 32 * u < span) {
        const double *r = src + 9 * (j0 + 32 * u + lane);
        double c[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) c[i] = r[i];
        pack_sym(c, p[u]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (j0 + 32 * u < span) {
        double *d = dst + 6 * (j0 + 32 * u) + 3 * lane;
        d[0] = p[u][0]; d[1] = p[u][1]; d[2] = p[u][2];
        d[96] = p[u][3]; d[97] = p[u][4]; d[98] = p[u][5];
      }
    }
    __syncwarp();
  }
}

// Start points of all pairs, one thread each (acos, atan2 and two sincos per pair: a few microseconds for
// the batch at full SIMT width, instead of 1200 instructions of a lone lane in front of every pair).
__global__ void __launch_bounds__(128) solve_prep_kernel(const double *__restrict__ poses, long long num_problems,
                                                         double *__restrict__ start_state) {
  const long long b = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= num_problems) return;
  double x[6], sc[4];
  solve_start_point(poses + 7 * b, x, sc);
  double *o = start_state + 10 * b;
#pragma unroll
  for (int i = 0; i < 6; ++i) o[i] = x[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) o[6 + i] = sc[i];
}

// ---- the order the pairs are started in
//
// A pair costs ~7 k cycles per LM iteration whatever the machine does around it, so one that runs into
// max_num_iterations (50: 355 k cycles = 0.18 ms; three of the 10 000 pairs of C2) ends 0.18 ms after
// it was STARTED, and in index order the last of them starts anywhere in a 0.26 ms kernel.  Which pairs
// those are shows at the start point: they are the ones whose normal equations are badly conditioned.
// The ratio of the smallest to the largest diagonal entry of J^T J over the pair's first 32
// correspondences -- one residual row per lane, 6 % of the data -- puts every pair of C2 with 20 or more
// iterations among the worst-conditioned 9 % (measured with the oracle, two seeds; cost, gradient norm
// or Jacobi-scaled measures do not separate them).  The pairs are therefore started in ascending order of that
// ratio (counting sort over quarter-octave bins): longest expected run first, the three-iteration bulk
// last.  Per-pair results do not depend on the order.
constexpr int kOrderBins = 264;  // quarter octaves from 2^-64 (and below) to 1

template <int V>
__global__ void __launch_bounds__(128) solve_score_kernel(BatchView bv, double reg, const double *__restrict__ start_state,
                                                          int *__restrict__ bins, int *__restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const long long b = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (b >= bv.num_problems) return;
  long long s, e;
  problem_range(bv, b, s, e);
  const int n = static_cast<int>(min(e - s, 32LL));
  double d[5] = {0, 0, 0, 0, 0};
  if (lane < n) {
    // pose constants from the start point and the sines / cosines solve_prep_kernel stored with it
    // (make_pose_const without its two sincos)
    const double *x = start_state + 10 * b;
    const double st = x[6], ct_ = x[7], sp = x[8], cp = x[9];
    PoseConst pc;
    pc.t[0] = st * cp;  pc.t[1] = st * sp;  pc.t[2] = ct_;
    pc.tth[0] = ct_ * cp; pc.tth[1] = ct_ * sp; pc.tth[2] = -st;
    pc.tph[0] = -st * sp; pc.tph[1] = st * cp; pc.tph[2] = 0.0;
    {
      const double qx = x[2], qy = x[3], qz = x[4], qw = x[5];
      const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
      const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
      const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
      const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
      pc.R[0] = 1.0 - (tyy + tzz); pc.R[1] = txy - twz;         pc.R[2] = txz + twy;
      pc.R[3] = txy + twz;         pc.R[4] = 1.0 - (txx + tzz); pc.R[5] = tyz - twx;
      pc.R[6] = txz - twy;         pc.R[7] = tyz + twx;         pc.R[8] = 1.0 - (txx + tyy);
    }
    const long long i = s + lane;
    double f1[3], f2[3], c9[9], ct[6] = {0, 0, 0, 0, 0, 0}, ch[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) { f1[k] = bv.f1[3 * i + k]; f2[k] = bv.f2[3 * i + k]; }
    if (VariantTraits<V>::kHasCt) {
#pragma unroll
      for (int k = 0; k < 9; ++k) c9[k] = bv.ct[9 * i + k];
      pack_sym(c9, ct);
    }
    if (VariantTraits<V>::kHasCh) {
#pragma unroll
      for (int k = 0; k < 9; ++k) c9[k] = bv.ch[9 * i + k];
      pack_sym(c9, ch);
    }
    double r, row[5];
    residual_row<V>(pc, reg, f1, f2, ct, ch, r, row);
    d[0] = row[0] * row[0]; d[1] = row[1] * row[1];
    d[2] = 4.0 * row[2] * row[2]; d[3] = 4.0 * row[3] * row[3]; d[4] = 4.0 * row[4] * row[4];  // tangent columns are 2 dr/dw
  }
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
  if (lane == 0) {
    const double lo = fmin(fmin(d[0], d[1]), fmin(d[2], fmin(d[3], d[4])));
    const double hi = fmax(fmax(d[0], d[1]), fmax(d[2], fmax(d[3], d[4])));
    int bin = 0;  // empty pairs, zero or non-finite rows: first
    if (hi > 0.0 && lo > 0.0 && hi < DBL_MAX) {
      const double ratio = lo / hi;  // in (0, 1]
      const long long q = (__double_as_longlong(ratio) >> 50) - ((1023LL - 64LL) << 2);  // quarter octaves above 2^-64
      bin = static_cast<int>(q < 1 ? 1 : (q > kOrderBins - 1 ? kOrderBins - 1 : q));
    }
    bins[b] = bin;
    atomicAdd(hist + bin, 1);
  }
}

// Counting sort by bin: every CTA forms the exclusive prefix of the histogram for itself (264 entries) and
// its threads take their places from per-bin cursors (`cursor` zeroed with the histogram before the call).
__global__ void __launch_bounds__(256) solve_order_scatter_kernel(const int *__restrict__ bins, const int *__restrict__ hist,
                                                                  int *__restrict__ cursor, int *__restrict__ order,
                                                                  long long num_problems) {
  __shared__ int s_off[kOrderBins];
  __shared__ int s_wsum[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // exclusive scan of hist[0 .. 263]: 256 threads take one entry each, the last eight entries follow serially
  const int v = hist[tid];
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  int base = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) base += w < warp ? s_wsum[w] : 0;
  s_off[tid] = base + incl - v;
  if (tid == 255) {
    int run = base + incl;
    for (int k = 256; k < kOrderBins; ++k) { s_off[k] = run; run += hist[k]; }
  }
  __syncthreads();
  const long long b = static_cast<long long>(blockIdx.x) * blockDim.x + tid;
  if (b >= num_problems) return;
  const int bin = bins[b];
  order[s_off[bin] + atomicAdd(cursor + bin, 1)] = static_cast<int>(b);
}

// The owner warp of one slot: claim, load, repack, LM loop, result; until the batch is exhausted.
template <int V, int P, int G>
__device__ __forceinline__ void slot_owner_warp(const SolveArgs &args, int slot, SlotCtl &ctl, double *base,
                                                SlotTickets &tk, uint64_t *bar_lm, uint64_t *bar_load,
                                                int lane) {
  constexpr int K = SlotLayout<V>::kCov;
  const int cap = args.cap_elems;
  const pnec_solver_opts &o = args.o;
  uint32_t par_lm = 0, par_load = 0;
  // Pairs are claimed one ahead: the atomic's round trip, the load of the next start point and the L2
  // prefetch of the next pair's correspondences run under the current pair's solve.
  const long long stride = static_cast<long long>(gridDim.x) * P;
  long long b = static_cast<long long>(blockIdx.x) * P + slot;
  if (args.work_counter) {
    unsigned int v = 0;
    if (lane == 0) v = atomicAdd(args.work_counter, 1u);
    b = __shfl_sync(0xffffffffu, v, 0);
    if (args.work_order && b < args.bv.num_problems) b = args.work_order[b];
  }
  double start_v = (lane < 10 && b < args.bv.num_problems) ? args.start_state[10 * b + lane] : 0.0;
  SLOT_PROBE_DECL();
  for (;;) {
    if (b >= args.bv.num_problems) break;
    long long s, e;
    problem_range(args.bv, b, s, e);
    const int n = static_cast<int>(e - s);
    const long long g0 = s & ~1LL;  // even => 16-byte aligned in every array
    const int head = static_cast<int>(s - g0);
    const int span = n + head;
    SLOT_PROBE_START();
#ifdef PNEC_SLOT_TIMING
    const long long t_pair = t_prev;
#endif
    unsigned int claimed = 0;  // lane 0: the next pair; read after the first candidate is posted
    if (lane == 0) {
      // (write-after-read against the previous pair's evaluation: every read of the slot returned its
      // value before the reader arrived on bar_lm, so no proxy fence is needed in front of the copies)
      if (n > 0) slot_issue<V>(args.bv, g0, span, cap, base, bar_load, 1);
      if (args.work_counter) claimed = atomicAdd(args.work_counter, 1u);
    }
    SLOT_PROBE(10);
    {  // start state: the point from solve_prep_kernel, pose constants one per lane
      if (lane < 6) ctl.lm.pts[0][lane] = start_v;
      else if (lane < 10) ctl.lm.scs[0][lane - 6] = start_v;
      if (lane == 0) solve_reset_state(args, n, ctl.lm);
      __syncwarp();
      SLOT_PROBE(11);
      pose_const_lanes(ctl.lm.pts[0] + 2, ctl.lm.scs[0], ctl.pc, lane);
    }
    long long b_next = b + stride;
    __syncwarp();
    SLOT_PROBE(9);
    if (n > 0) {
      if (K > 0) {
        mbar_wait(bar_load, par_load);
        par_load ^= 1u;
        SLOT_PROBE(0);
        slot_repack(base, base, span, lane);
        if (K == 2) slot_repack(base + 6 * cap, base + 9 * cap, span, lane);
        __syncwarp();
        SLOT_PROBE(1);
        if (lane == 0) slot_issue<V>(args.bv, g0, span, cap, base, bar_load, 2);
      }
      mbar_wait(bar_load, par_load);
      par_load ^= 1u;
      __syncwarp();
      SLOT_PROBE(2);
      if (lane == 0) {
        ctl.head = head;
        ctl.span = span;
        ctl.mode = kPassFull;
        slot_post(tk, slot);
      }
      if (args.work_counter) {
        b_next = __shfl_sync(0xffffffffu, claimed, 0);
        if (args.work_order && b_next < args.bv.num_problems) b_next = args.work_order[b_next];
      }
      if (b_next < args.bv.num_problems) {
        if (lane == 0) slot_prefetch<V>(args.bv, b_next);
        if (lane < 10) start_v = args.start_state[10 * b_next + lane];
      }
      bool first = true;
      for (;;) {
        mbar_wait(bar_lm, par_lm);
        par_lm ^= 1u;
        SLOT_PROBE(3);
        bool posted = false;
        if (ctl.mode == kPassCost) {
          double t = 0.0;
#pragma unroll
          for (int w = 0; w < kSlotEvalWarps; ++w) t += ctl.part[w][kNumAcc - 1];
          t *= 0.5;
          lm_after_cost_pass(ctl.lm, t, o, lane);
        } else {
          if (lane < kNumAcc) {
            double mine = 0.0;
#pragma unroll
            for (int w = 0; w < kSlotEvalWarps; ++w) mine += ctl.part[w][lane];
            ctl.lm.tot[ctl.lm.ti ^ 1][lane] = mine * acc_scale(lane);
          }
          __syncwarp();
          // the candidate is posted from inside the update, before its bookkeeping
          lm_step(first, ctl.lm, o, lane, ctl.pc, [&](int pass_mode) {
            __syncwarp();
            if (lane == 0) {
              ctl.mode = pass_mode;
              slot_post(tk, slot);
            }
            posted = true;
          });
        }
        __syncwarp();
        SLOT_PROBE(4);
        if (ctl.lm.done) {
          if (posted) {  // converged by parameter tolerance after the candidate went out: discard its sums
            mbar_wait(bar_lm, par_lm);
            par_lm ^= 1u;
          }
          break;
        }
        first = false;
        if (!posted && lane == 0) {  // after a cost-only pass: the same candidate again, in full
          ctl.mode = ctl.lm.pass_mode;
          slot_post(tk, slot);
        }
      }
    }
    if (lane == 0) solve_write_result(args, b, ctl.lm);
    __syncwarp();
#ifdef PNEC_SLOT_TIMING
    probe_acc[5] += clock64() - t_pair;
#endif
    if (n <= 0) {  // (nothing was posted: the claim and the next start point are still to be read)
      if (args.work_counter) {
        b_next = __shfl_sync(0xffffffffu, claimed, 0);
        if (args.work_order && b_next < args.bv.num_problems) b_next = args.work_order[b_next];
      }
      if (lane < 10 && b_next < args.bv.num_problems) start_v = args.start_state[10 * b_next + lane];
    }
    b = b_next;
  }
  if (lane == 0 && atomicAdd(&tk.owners_gone, 1u) == P - 1) {
    for (int g = 0; g < G; ++g) slot_post(tk, -1);  // consecutive tickets: one for every evaluation group
  }
#ifdef PNEC_SLOT_TIMING
  const int warp_role_ = 0;
#endif
  SLOT_PROBE_FLUSH();
}

// An evaluation warp: quarter `warp` of the posted candidates that fall to its group (ticket t goes to group
// t mod G), in ticket order.
template <int V, int P, int G>
__device__ __forceinline__ void slot_eval_warp(const SolveArgs &args, SlotCtl *ctl, SlotTickets &tk,
                                               uint64_t *bar_lm, int group, int warp, int lane) {
  constexpr int kD = SlotLayout<V>::kDoubles;
  const int cap = args.cap_elems;
  const double reg = args.o.regularization;
  SLOT_PROBE_DECL();
  SLOT_PROBE_START();
  for (unsigned int h = group;; h += G) {
    mbar_wait(reinterpret_cast<uint64_t *>(&tk.bar[h & (kSlotTickets - 1)]), (h / kSlotTickets) & 1u);
    const int s = tk.slot[h & (kSlotTickets - 1)];
    if (s < 0) break;
    SlotCtl &c = ctl[s];
    const int mode = c.mode;
    SLOT_PROBE(7);
    const double *base = dyn_smem + static_cast<size_t>(s) * kD * cap;
    PoseConst pc;
    load_pose_const(c.pc, pc);
    const int hi = c.span;
    if (mode == kPassCost) {
      double sum = 0.0;
      for (int j = c.head + warp * 32 + lane; j < hi; j += kSlotEvalWarps * 32) {
        double a1[3], a2[3], c1[6], c2[6];
        slot_load_corr<V>(base, cap, j, a1, a2, c1, c2);
        const double r = residual_only<V>(pc, reg, a1, a2, c1, c2);
        sum = fma(r, r, sum);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) c.part[warp][kNumAcc - 1] = sum;
    } else {
      double acc[kNumAcc];
#pragma unroll
      for (int i = 0; i < kNumAcc; ++i) acc[i] = 0.0;
      for (int j = c.head + warp * 32 + lane; j < hi; j += kSlotEvalWarps * 32) {
        double a1[3], a2[3], c1[6], c2[6], r, row[5];
        slot_load_corr<V>(base, cap, j, a1, a2, c1, c2);
        residual_row<V>(pc, reg, a1, a2, c1, c2, r, row);
        accumulate(acc, r, row);
      }
      const double v = warp_transpose_reduce(acc, lane);
      const int idx = warp_reduce_owner_index(lane);
      if (idx >= 0 && idx < kNumAcc) c.part[warp][idx] = v;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_lm[s]);
    SLOT_PROBE(6);
#ifdef PNEC_SLOT_TIMING
    if (warp == 0) probe_acc[8] += 1;
#endif
  }
#ifdef PNEC_SLOT_TIMING
  const int warp_role_ = 1;
#endif
  SLOT_PROBE_FLUSH();
}

// P slots (= owner warps) and G groups of four evaluation warps per CTA: <2, 1> twice per SM, or <4, 2> once
// (the two groups then serve all four slots of the SM).
template <int V, int P, int G>
__global__ void __launch_bounds__((kSlotEvalWarps * G + P) * 32, G == 1 ? 2 : 1)
solve_slots_kernel(const __grid_constant__ SolveArgs args) {
  static_assert(P <= kSlotTickets, "a reader must never be lapped");
  __shared__ __align__(8) uint64_t s_bar_lm[P], s_bar_load[P];
  __shared__ __align__(8) SlotTickets s_tk;
  __shared__ SlotCtl s_ctl[P];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < P; ++s) {
      mbar_init(&s_bar_lm[s], kSlotEvalWarps);
      mbar_init(&s_bar_load[s], 1);
    }
#pragma unroll
    for (int r = 0; r < kSlotTickets; ++r) mbar_init(reinterpret_cast<uint64_t *>(&s_tk.bar[r]), 1);
    s_tk.tail = 0u;
    s_tk.owners_gone = 0u;
    fence_mbar_init();
  }
  __syncthreads();
  if (warp >= kSlotEvalWarps * G) {
    const int slot = warp - kSlotEvalWarps * G;
    double *base = dyn_smem + static_cast<size_t>(slot) * SlotLayout<V>::kDoubles * args.cap_elems;
    slot_owner_warp<V, P, G>(args, slot, s_ctl[slot], base, s_tk, &s_bar_lm[slot], &s_bar_load[slot], lane);
  } else {
    slot_eval_warp<V, P, G>(args, s_ctl, s_tk, s_bar_lm, warp / kSlotEvalWarps, warp % kSlotEvalWarps, lane);
  }
  __syncthreads();
  // the work counter serves the next launch on this stream: the last CTA to leave rewinds it
  if (tid == 0 && args.work_counter) {
    __threadfence();
    if (atomicAdd(args.work_counter + 1, 1u) == gridDim.x - 1) {
      args.work_counter[0] = 0u;
      args.work_counter[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace pnec
