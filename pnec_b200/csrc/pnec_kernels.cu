// pnec_kernels.cu — sm_100a kernels and the C-ABI of the PNEC frame-pair solver.
//
//   solve_kernel   one CTA per frame pair.  The pair's correspondences are pulled
//                  from HBM ONCE into shared memory with 1-D bulk async copies (TMA
//                  engine, mbarrier complete_tx), then every LM iteration — fused
//                  residual + analytic Jacobian + J^T J / J^T r reduction, 5x5
//                  damped Cholesky, SO(3) x S^2 retraction, accept/reject — runs
//                  out of shared memory with no host round trip.
//   eval_kernel    the fused residual + Jacobian + J^T J pass alone (K1), streaming
//                  tiles through an S-stage bulk-copy ring: the HBM-roofline kernel.
//   cost_kernel    pnec::common::CostFunction (parity metric).
//
// Replaces (reference, file:line): PNECCeres::Optimize src/optimization/pnec_ceres.cc:70-168,
// NECCeres::Optimize src/optimization/nec_ceres.cc:73-101, and under them
// ceres::Solve + N x NumericDiffCostFunction<F, CENTRAL, 1,1,1,4>.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "pnec_device.cuh"

namespace pnec {

struct BatchView {
  const double *f1, *f2, *ct, *ch;
  const long long *offsets;  // device, B+1, or nullptr for uniform
  long long n_uniform, num_problems, total;
  const double *poses;  // [B][7]
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

// --------------------------------------------------------------- solve kernel

struct SolveArgs {
  BatchView bv;
  pnec_solver_opts o;
  double *out_poses;
  int *out_status;
  int *out_iters;
  double *out_cost;
  double *out_init_cost;
  int cap_elems;  // resident capacity of the dynamic smem, in correspondences (even)
  int use_bulk;   // all base pointers 16-byte aligned
  long long *dbg; // PNEC_PHASE_TIMING builds only: per-CTA cycle counters
};

// PNECCeres::InitValues(orientation, translation) (pnec_ceres.cc:188-192) + the start state of
// the minimiser; called by one thread.
__device__ __forceinline__ void solve_init_state(const SolveArgs &args, long long b, int n,
                                                 LMState &st, PoseConst &s_pc) {
  const double *p = args.bv.poses + 7 * b;
  double *x = st.pts[0], *sc = st.scs[0];
  angles_from_vec(p + 4, x[0], x[1]);
  x[2] = p[0]; x[3] = p[1]; x[4] = p[2]; x[5] = p[3];
  sincos(x[0], &sc[0], &sc[1]);
  sincos(x[1], &sc[2], &sc[3]);
  st.inv_radius = 1.0 / args.o.initial_trust_region_radius;
  st.decrease_factor = 2.0;
  st.inv_model_cost_change = 0.0;
  st.x_cost = 0.0;
  st.initial_cost = 0.0;
  st.xi = 0;
  st.ti = 0;
  st.iteration = 0;
  st.num_invalid = 0;
  st.reuse_diagonal = 0;
  st.step_successful = 1;
  st.grad_converged = 0;
  st.status = (n <= 0) ? PNEC_STATUS_EMPTY : PNEC_STATUS_MAX_ITERATIONS;
  st.done = (n <= 0) ? 1 : 0;
  st.pass_mode = kPassFull;
  PoseConst pc0;
  make_pose_const_sc(sc, x + 2, pc0);
  s_pc = pc0;
}

// PNECCeres::Result(): q.normalized(), t(theta, phi) (pnec_ceres.cc:201-206) + the summary.
__device__ __forceinline__ void solve_write_result(const SolveArgs &args, long long b,
                                                   const LMState &st) {
  const double *x = st.pts[st.xi], *sc = st.scs[st.xi];
  const double qn = sqrt(x[2] * x[2] + x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);
  const double iq = qn > 0.0 ? 1.0 / qn : 1.0;
  double *op = args.out_poses + 7 * b;
  op[0] = x[2] * iq; op[1] = x[3] * iq; op[2] = x[4] * iq; op[3] = x[5] * iq;
  op[4] = sc[0] * sc[3]; op[5] = sc[0] * sc[2]; op[6] = sc[1];
  if (args.out_status) args.out_status[b] = st.status;
  if (args.out_iters) args.out_iters[b] = st.iteration;
  if (args.out_cost) args.out_cost[b] = st.x_cost;
  if (args.out_init_cost) args.out_init_cost[b] = st.initial_cost;
}

extern __shared__ __align__(16) double dyn_smem[];

template <int V, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) solve_kernel(const __grid_constant__ SolveArgs args) {
  constexpr int NT = NW * 32;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ PoseConst s_pc;
  __shared__ LMState s_lm;
  __shared__ double s_part[NW][kAccPad];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = blockIdx.x;
  long long s, e;
  problem_range(args.bv, b, s, e);
  const int n = static_cast<int>(e - s);
  const long long g0 = s & ~1LL;  // even => 16-byte aligned in every array
  const int head = static_cast<int>(s - g0);
  const int span = n + head;
  const bool resident = (span + (span & 1)) <= args.cap_elems;
  const pnec_solver_opts &o = args.o;
  // The warp that runs the serial part (set-up, LM update).  Warp w of every CTA sits on SM
  // sub-partition w % 4, so a fixed choice would pile the serial fp64 work of all co-resident
  // problems onto one sub-partition; rotate it per CTA instead.
  const int lmw = (NW == 1) ? 0 : static_cast<int>((blockIdx.x + blockIdx.x / 148u) % NW);
  const int lmt = lmw * 32;  // its lane 0

  double *sf1 = dyn_smem;
  double *sf2 = sf1 + 3 * args.cap_elems;
  double *sct = sf2 + 3 * args.cap_elems;
  double *sch = sct + (VariantTraits<V>::kHasCt ? 9 * args.cap_elems : 0);

  if (tid == lmt) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
    // get the HBM -> shared-memory copies going before the scalar set-up below
    if (n > 0 && resident && args.use_bulk)
      issue_bulk<V>(args.bv, g0, span, sf1, sf2, sct, sch, &s_bar);
    solve_init_state(args, b, n, s_lm, s_pc);
  }
  __syncthreads();

  if (n > 0) {
    if (resident) {
      if (args.use_bulk) {
        mbar_wait(&s_bar, 0);
      } else {
        copy_plain<V, NT>(args.bv, g0, span, sf1, sf2, sct, sch, tid);
        __syncthreads();
      }
    }
#ifdef PNEC_PHASE_TIMING
    long long t_e = 0, t_l = 0, t_w = 0, n_pass = 0, t_start = clock64();
#endif
    bool first = true;
    for (;;) {
#ifdef PNEC_PHASE_TIMING
      const long long t0 = clock64();
#endif
      PoseConst pc;
      load_pose_const(s_pc, pc);
      const int mode = first ? kPassFull : s_lm.pass_mode;
      const double *pf1 = resident ? sf1 : args.bv.f1 + 3 * s;
      const double *pf2 = resident ? sf2 : args.bv.f2 + 3 * s;
      const double *pct = resident ? sct : (VariantTraits<V>::kHasCt ? args.bv.ct + 9 * s : nullptr);
      const double *pch = resident ? sch : (VariantTraits<V>::kHasCh ? args.bv.ch + 9 * s : nullptr);
      const int lo = resident ? head : 0, hi = resident ? span : n;
      if (mode == kPassCost) {
        double sum;
        if (resident) sum = eval_pass_cost<V, NT>(pc, o.regularization, sf1, sf2, sct, sch, lo, hi, tid);
        else sum = eval_pass_cost<V, NT>(pc, o.regularization, pf1, pf2, pct, pch, lo, hi, tid);
        const double cand_cost = block_reduce_scalar<NW>(sum, s_part, warp, lane, lmw);
        if (warp == lmw) lm_after_cost_pass(s_lm, cand_cost, o, lane);
      } else {
        double acc[kNumAcc];
#pragma unroll
        for (int i = 0; i < kNumAcc; ++i) acc[i] = 0.0;
        // two call sites so the resident one compiles to shared-memory loads
        if (resident) eval_pass<V, NT>(pc, o.regularization, sf1, sf2, sct, sch, lo, hi, tid, acc);
        else eval_pass<V, NT>(pc, o.regularization, pf1, pf2, pct, pch, lo, hi, tid, acc);
        block_reduce<NW>(acc, s_part, warp, lane, s_lm.tot[s_lm.ti ^ 1], lmw);
#ifdef PNEC_PHASE_TIMING
        const long long t1 = clock64();
        t_e += t1 - t0;
#endif
        if (warp == lmw && (kLmFullWarp || lane == 0)) {
          if (first) lm_step<true>(s_lm, o, lane, s_pc);
          else lm_step<false>(s_lm, o, lane, s_pc);
        }
#ifdef PNEC_PHASE_TIMING
        const long long t2 = clock64();
        t_l += t2 - t1;
        ++n_pass;
#endif
      }
#ifdef PNEC_PHASE_TIMING
      const long long t3 = clock64();
#endif
      __syncthreads();
#ifdef PNEC_PHASE_TIMING
      t_w += clock64() - t3;
#endif
      if (s_lm.done) break;
      first = false;
    }
#ifdef PNEC_PHASE_TIMING
    if (args.dbg && (tid == lmt || tid == ((lmw + 1) % NW) * 32)) {
      long long *d = args.dbg + 12 * b + (tid != lmt ? 6 : 0);
      d[0] = t_e; d[1] = t_l; d[2] = t_w; d[3] = n_pass; d[4] = clock64() - t_start; d[5] = t_start;
    }
#endif
  }

  if (tid == lmt) solve_write_result(args, b, s_lm);
}

// ------------------------------------------------------ solve kernel, streaming
//
// Same LM solve for frame pairs too large to keep 2+ of them resident per SM (N > 896):
// nothing is resident.  Warp w owns tiles w, w + NW, ... of the pair and a private S-stage
// ring; lane 0 keeps the ring S tiles ahead with bulk async copies and simply wraps around
// at the end of a pass, so the first tiles of the NEXT pass are already in flight while the
// LM update runs.  Pass 1 comes from HBM; later passes hit L2 (the working set of the
// resident CTAs is a few tens of MB).  On exit every warp drains its outstanding copies.
template <int V, int NW, int S, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
solve_stream_kernel(const __grid_constant__ SolveArgs args) {
  constexpr int T = 32;
  constexpr int kStageDoubles = T * VariantTraits<V>::kDoubles;
  __shared__ __align__(8) uint64_t s_full[NW][S];
  __shared__ PoseConst s_pc;
  __shared__ LMState s_lm;
  __shared__ double s_part[NW][kAccPad];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
  const long long b = blockIdx.x;
  long long s, e;
  problem_range(args.bv, b, s, e);
  const int n = static_cast<int>(e - s);
  const long long g0 = s & ~1LL;
  const int head = static_cast<int>(s - g0);
  const int span = n + head;
  const int ntiles = (n > 0) ? (span + T - 1) / T : 0;
  const int nt_w = (ntiles > warp) ? (ntiles - warp + NW - 1) / NW : 0;  // tiles of this warp
  const pnec_solver_opts &o = args.o;
  const int lmw = (NW == 1) ? 0 : static_cast<int>((blockIdx.x + blockIdx.x / 148u) % NW);
  const int lmt = lmw * 32;
  double *ring = dyn_smem + static_cast<size_t>(warp) * S * kStageDoubles;
  uint64_t *full = s_full[warp];

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < S; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncwarp();

  // producer: tile p_i of this warp's sequence goes to stage p_stage; wraps at nt_w
  int p_i = 0, p_stage = 0, in_flight = 0;
  auto producer_issue = [&]() {
    if (lane == 0) {
      const int k = warp + p_i * NW;
      double *base = ring + p_stage * kStageDoubles;
      issue_bulk<V>(args.bv, g0 + static_cast<long long>(k) * T, min(T, span - k * T), base,
                    base + 3 * T, base + 6 * T, base + 15 * T, &full[p_stage]);
    }
    p_i = (p_i + 1 == nt_w) ? 0 : p_i + 1;
    p_stage = (p_stage + 1 == S) ? 0 : p_stage + 1;
    ++in_flight;
  };
  if (nt_w > 0) {
#pragma unroll 1
    for (int i = 0; i < S; ++i) producer_issue();
  }

  if (tid == lmt) {
    solve_init_state(args, b, n, s_lm, s_pc);
  }
  __syncthreads();

  int c_stage = 0;
  uint32_t c_parity = 0;
  if (n > 0) {
    bool first = true;
    for (;;) {
      PoseConst pc;
      load_pose_const(s_pc, pc);
      const int mode = first ? kPassFull : s_lm.pass_mode;
      double acc[kNumAcc];
#pragma unroll
      for (int i = 0; i < kNumAcc; ++i) acc[i] = 0.0;
      for (int i = 0; i < nt_w; ++i) {
        mbar_wait(&full[c_stage], c_parity);
        const double *base = ring + c_stage * kStageDoubles;
        const int idx = (warp + i * NW) * T + lane;
        const bool valid = (idx >= head) && (idx < span);
        double a1[3], a2[3], c1[6], c2[6];
        if (valid) load_corr<V>(base, base + 3 * T, base + 6 * T, base + 15 * T, lane, a1, a2, c1, c2);
        __syncwarp();  // the stage may be refilled
        --in_flight;
        producer_issue();
        if (++c_stage == S) {
          c_stage = 0;
          c_parity ^= 1u;
        }
        if (valid) {
          if (mode == kPassCost) {
            const double r = residual_only<V>(pc, o.regularization, a1, a2, c1, c2);
            acc[kNumAcc - 1] = fma(r, r, acc[kNumAcc - 1]);
          } else {
            double r, row[5];
            residual_row<V>(pc, o.regularization, a1, a2, c1, c2, r, row);
            accumulate(acc, r, row);
          }
        }
      }
      if (mode == kPassCost) {
        const double cand_cost = block_reduce_scalar<NW>(acc[kNumAcc - 1], s_part, warp, lane, lmw);
        if (warp == lmw) lm_after_cost_pass(s_lm, cand_cost, o, lane);
      } else {
        block_reduce<NW>(acc, s_part, warp, lane, s_lm.tot[s_lm.ti ^ 1], lmw);
        if (warp == lmw && (kLmFullWarp || lane == 0)) {
          if (first) lm_step<true>(s_lm, o, lane, s_pc);
          else lm_step<false>(s_lm, o, lane, s_pc);
        }
      }
      __syncthreads();
      if (s_lm.done) break;
      first = false;
    }
  }
  // drain the copies that were issued ahead: the CTA must not exit with bulk copies in flight
  while (in_flight > 0) {
    mbar_wait(&full[c_stage], c_parity);
    if (++c_stage == S) {
      c_stage = 0;
      c_parity ^= 1u;
    }
    --in_flight;
  }

  if (tid == lmt) solve_write_result(args, b, s_lm);
}

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
  constexpr bool kCt = VariantTraits<V>::kHasCt, kCh = VariantTraits<V>::kHasCh;
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

  // ---- producer cursor (uniform across the warp; lane 0 issues).  Running pointers to the
  // next tile of the current producer problem, elements left in it, next ring stage.
  int pj = 0, p_left = 0, p_stage = 0;
  const double *p_f1 = nullptr, *p_f2 = nullptr, *p_ct = nullptr, *p_ch = nullptr;
  auto producer_open = [&]() {  // position on the first tile of the next non-empty problem
    while (pj < nmine) {
      long long s, e;
      problem_range(args.bv, gw + pj * W, s, e);
      if (e > s) {
        const long long g0 = s & ~1LL;  // even => 16-byte aligned in every array
        p_left = static_cast<int>(e - g0);
        p_f1 = args.bv.f1 + 3 * g0;
        p_f2 = args.bv.f2 + 3 * g0;
        if (kCt) p_ct = args.bv.ct + 9 * g0;
        if (kCh) p_ch = args.bv.ch + 9 * g0;
        return;
      }
      ++pj;
    }
  };
  auto producer_issue = [&]() {  // issue the current producer tile, then advance
    if (pj >= nmine) return;
    const int cnt = min(T, p_left);
    if (lane == 0) {
      double *base = ring + p_stage * kStageDoubles;
      int cb = cnt + (cnt & 1);  // bulk copies move 16-byte units: round up to an even count ...
      if ((cnt & 1) && p_f1 + 3 * cb > args.bv.f1 + 3 * args.bv.total) {
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
      producer_open();
    }
  };
  producer_open();
#pragma unroll 1
  for (int i = 0; i < S; ++i) producer_issue();

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
        producer_issue();
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
        producer_issue();
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

}  // namespace pnec

// =================================================================== host side

using namespace pnec;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

#define PNEC_CUDA(call)                                                                      \
  do {                                                                                       \
    cudaError_t err__ = (call);                                                              \
    if (err__ != cudaSuccess)                                                                \
      return fail(PNEC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));     \
  } while (0)

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

}  // namespace

struct pnec_handle {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  int64_t launches = 0;
  // staging for HOST-memspace calls and for the device copy of offsets
  DevBuf d_f1, d_f2, d_ct, d_ch, d_off, d_poses;
  DevBuf d_out_poses, d_out_status, d_out_iters, d_out_cost, d_out_init, d_out_grad, d_out_jtj;
  DevBuf d_ut_mu, d_ut_cov, d_ut_out;
  std::mutex mu;
};

namespace {

struct Staged {
  BatchView bv;
  long long max_n = 0;
};

int validate_batch(const pnec_batch *b, int variant, bool need_poses) {
  if (!b) return fail(PNEC_ERR_INVALID_ARGUMENT, "batch is NULL");
  if (b->num_problems < 0) return fail(PNEC_ERR_INVALID_ARGUMENT, "num_problems < 0");
  if (!b->offsets && b->n_per_problem < 0)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "n_per_problem < 0");
  if (b->memspace != PNEC_MEM_HOST && b->memspace != PNEC_MEM_DEVICE)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown memspace");
  if (variant < PNEC_VARIANT_NEC || variant > PNEC_VARIANT_SYMMETRIC)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown residual variant");
  long long total = 0;
  if (b->offsets) {
    if (b->offsets[0] < 0) return fail(PNEC_ERR_INVALID_ARGUMENT, "offsets[0] < 0");
    for (int64_t i = 0; i < b->num_problems; ++i)
      if (b->offsets[i + 1] < b->offsets[i])
        return fail(PNEC_ERR_INVALID_ARGUMENT, "offsets must be non-decreasing");
    total = b->offsets[b->num_problems];
  } else {
    total = b->num_problems * b->n_per_problem;
  }
  if (total > 0) {
    if (!b->bvs_host || !b->bvs_target)
      return fail(PNEC_ERR_INVALID_ARGUMENT, "bearing vector arrays are NULL");
    if (variant != PNEC_VARIANT_NEC && !b->covs_target)
      return fail(PNEC_ERR_INVALID_ARGUMENT, "covs_target is NULL for a PNEC variant");
    if (variant == PNEC_VARIANT_SYMMETRIC && !b->covs_host)
      return fail(PNEC_ERR_INVALID_ARGUMENT, "covs_host is NULL for the SYMMETRIC variant");
  }
  if (need_poses && b->num_problems > 0 && !b->poses)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "poses is NULL");
  return PNEC_OK;
}

// Builds the device view of a batch: copies H2D for HOST batches, always copies
// the (host) offsets.  Everything is enqueued on `stream`.
int stage_batch(pnec_handle *h, const pnec_batch *b, int variant, cudaStream_t stream,
                Staged *out) {
  const long long B = b->num_problems;
  long long total, max_n = 0;
  if (b->offsets) {
    total = b->offsets[B];
    for (long long i = 0; i < B; ++i) max_n = std::max<long long>(max_n, b->offsets[i + 1] - b->offsets[i]);
  } else {
    total = B * b->n_per_problem;
    max_n = b->n_per_problem;
  }
  BatchView bv{};
  bv.num_problems = B;
  bv.total = b->offsets ? b->offsets[B] : total;
  bv.n_uniform = b->offsets ? 0 : b->n_per_problem;
  const long long base = b->offsets ? b->offsets[0] : 0;
  (void)base;
  if (b->offsets) {
    PNEC_CUDA(h->d_off.ensure(sizeof(long long) * (B + 1)));
    PNEC_CUDA(cudaMemcpyAsync(h->d_off.p, b->offsets, sizeof(long long) * (B + 1),
                              cudaMemcpyHostToDevice, stream));
    bv.offsets = static_cast<const long long *>(h->d_off.p);
  }
  const bool need_ct = variant != PNEC_VARIANT_NEC;
  const bool need_ch = variant == PNEC_VARIANT_SYMMETRIC;
  if (b->memspace == PNEC_MEM_HOST) {
    const size_t nel = static_cast<size_t>(bv.total);
    PNEC_CUDA(h->d_f1.ensure(nel * 24));
    PNEC_CUDA(h->d_f2.ensure(nel * 24));
    PNEC_CUDA(h->d_poses.ensure(static_cast<size_t>(B) * 56));
    if (nel) {
      PNEC_CUDA(cudaMemcpyAsync(h->d_f1.p, b->bvs_host, nel * 24, cudaMemcpyHostToDevice, stream));
      PNEC_CUDA(cudaMemcpyAsync(h->d_f2.p, b->bvs_target, nel * 24, cudaMemcpyHostToDevice, stream));
    }
    if (need_ct) {
      PNEC_CUDA(h->d_ct.ensure(nel * 72));
      if (nel)
        PNEC_CUDA(cudaMemcpyAsync(h->d_ct.p, b->covs_target, nel * 72, cudaMemcpyHostToDevice, stream));
    }
    if (need_ch) {
      PNEC_CUDA(h->d_ch.ensure(nel * 72));
      if (nel)
        PNEC_CUDA(cudaMemcpyAsync(h->d_ch.p, b->covs_host, nel * 72, cudaMemcpyHostToDevice, stream));
    }
    if (B && b->poses)
      PNEC_CUDA(cudaMemcpyAsync(h->d_poses.p, b->poses, static_cast<size_t>(B) * 56,
                                cudaMemcpyHostToDevice, stream));
    bv.f1 = static_cast<const double *>(h->d_f1.p);
    bv.f2 = static_cast<const double *>(h->d_f2.p);
    bv.ct = need_ct ? static_cast<const double *>(h->d_ct.p) : nullptr;
    bv.ch = need_ch ? static_cast<const double *>(h->d_ch.p) : nullptr;
    bv.poses = static_cast<const double *>(h->d_poses.p);
  } else {
    bv.f1 = b->bvs_host;
    bv.f2 = b->bvs_target;
    bv.ct = need_ct ? b->covs_target : nullptr;
    bv.ch = need_ch ? b->covs_host : nullptr;
    bv.poses = b->poses;
  }
  out->bv = bv;
  out->max_n = max_n;
  return PNEC_OK;
}

bool bulk_ok(const BatchView &bv) {
  return aligned16(bv.f1) && aligned16(bv.f2) && (!bv.ct || aligned16(bv.ct)) &&
         (!bv.ch || aligned16(bv.ch));
}

int bytes_per_corr(int variant) {
  switch (variant) {
    case PNEC_VARIANT_NEC: return 48;
    case PNEC_VARIANT_SYMMETRIC: return 192;
    default: return 120;
  }
}

// ------------------------------------------------------------ kernel launchers

constexpr size_t kStaticSmemReserve = 3072;  // static __shared__ of the kernels + slack

template <int V, int NW, int MINB>
int launch_solve_t(pnec_handle *h, const SolveArgs &a, size_t dyn, cudaStream_t stream) {
  auto kern = solve_kernel<V, NW, MINB>;
  PNEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(dyn)));
  kern<<<static_cast<unsigned>(a.bv.num_problems), NW * 32, dyn, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  return PNEC_OK;
}

template <int V>
int launch_solve_v(pnec_handle *h, const SolveArgs &a, int nw, size_t dyn, cudaStream_t stream) {
  switch (nw) {
    case 1: return launch_solve_t<V, 1, 8>(h, a, dyn, stream);
    case 2: return launch_solve_t<V, 2, 4>(h, a, dyn, stream);
    case 4: return launch_solve_t<V, 4, 3>(h, a, dyn, stream);
    case 8: return launch_solve_t<V, 8, 1>(h, a, dyn, stream);
    default: return fail(PNEC_ERR_INVALID_ARGUMENT, "unsupported warps per problem");
  }
}

#ifdef PNEC_PHASE_TIMING
void dump_phase_timing(const SolveArgs &a) {
  cudaDeviceSynchronize();
  const long long B = a.bv.num_problems;
  std::vector<long long> hbuf(12 * B);
  cudaMemcpy(hbuf.data(), a.dbg, sizeof(long long) * 12 * B, cudaMemcpyDeviceToHost);
  double s[12] = {0};
  for (long long b = 0; b < B; ++b)
    for (int k = 0; k < 12; ++k) s[k] += hbuf[12 * b + k];
  unsigned long long probe[8];
  cudaMemcpyFromSymbol(probe, g_lm_probe, sizeof(probe));
  std::fprintf(stderr, "[lm_step probes, cycles per CTA] judge %.0f bookkeeping %.0f tr-step %.0f candidate %.0f store %.0f\n",
               (double)probe[0] / B, (double)probe[1] / B, (double)probe[2] / B, (double)probe[3] / B, (double)probe[4] / B);
  unsigned long long zero[8] = {0};
  cudaMemcpyToSymbol(g_lm_probe, zero, sizeof(zero));
  std::fprintf(stderr, "[phase timing, mean cycles per CTA] warp0: eval %.0f lm %.0f barrier %.0f full-passes %.2f "
               "total %.0f | warp1: eval %.0f lm(idle) %.0f barrier %.0f\n",
               s[0] / B, s[1] / B, s[2] / B, s[3] / B, s[4] / B, s[6] / B, s[7] / B, s[8] / B);
}
#endif

template <int V, int NW, int S, int MINB>
int launch_solve_stream_t(pnec_handle *h, const SolveArgs &a, cudaStream_t stream) {
  auto kern = solve_stream_kernel<V, NW, S, MINB>;
  const size_t dyn = static_cast<size_t>(NW) * S * 32 * VariantTraits<V>::kDoubles * 8;
  PNEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(dyn)));
  kern<<<static_cast<unsigned>(a.bv.num_problems), NW * 32, dyn, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  return PNEC_OK;
}

template <int V>
int launch_solve_stream_v(pnec_handle *h, const SolveArgs &a, cudaStream_t stream) {
  switch (env_int("PNEC_B200_STREAM_CFG", 3)) {
    case 1: return launch_solve_stream_t<V, 6, 4, 2>(h, a, stream);   // 2 CTAs x 6 warps per SM
    case 2: return launch_solve_stream_t<V, 12, 4, 1>(h, a, stream);  // 1 CTA x 12 warps
    case 3: return launch_solve_stream_t<V, 4, 4, 3>(h, a, stream);   // 3 CTAs x 4 warps
    case 4: return launch_solve_stream_t<V, 6, 3, 2>(h, a, stream);
    case 5: return launch_solve_stream_t<V, 1, 4, 12>(h, a, stream);  // one warp per pair, 12 pairs per SM
    case 6: return launch_solve_stream_t<V, 2, 4, 6>(h, a, stream);
    case 7: return launch_solve_stream_t<V, 1, 3, 12>(h, a, stream);
    case 8: return launch_solve_stream_t<V, 1, 6, 8>(h, a, stream);
    default: return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown PNEC_B200_STREAM_CFG");
  }
}

int launch_solve(pnec_handle *h, const SolveArgs &a0, int variant, long long max_n,
                 cudaStream_t stream) {
  SolveArgs a = a0;
  if (a.bv.num_problems == 0) return PNEC_OK;
  // Large frame pairs: stream every pass (bulk-copy rings) instead of keeping one pair per SM
  // resident.  Needs 16-byte aligned arrays; SYMMETRIC (192 B / correspondence) stays resident-first.
  const long long stream_min_n = env_int("PNEC_B200_STREAM_MIN_N", 896);
  if (bulk_ok(a.bv) && !env_int("PNEC_B200_NO_BULK", 0) && max_n > stream_min_n) {
    a.use_bulk = 1;
    a.cap_elems = 0;
    a.dbg = nullptr;
    switch (variant) {
      case PNEC_VARIANT_NEC: return launch_solve_stream_v<PNEC_VARIANT_NEC>(h, a, stream);
      case PNEC_VARIANT_TARGET: return launch_solve_stream_v<PNEC_VARIANT_TARGET>(h, a, stream);
      case PNEC_VARIANT_HOST: return launch_solve_stream_v<PNEC_VARIANT_HOST>(h, a, stream);
      default: return launch_solve_stream_v<PNEC_VARIANT_SYMMETRIC>(h, a, stream);
    }
  }
  int nw = env_int("PNEC_B200_SOLVE_WARPS", 0);
  // Warps per frame pair (measured on B200, tools/nw_sweep.py): the solve is latency-bound, so
  // what pays is the number of pairs resident per SM, not the width of one pair.  One warp per
  // pair wins while >= 7 pairs fit in shared memory; 4 warps once only 3 fit.
  if (nw == 0) nw = max_n <= 320 ? 1 : max_n <= 448 ? 2 : max_n <= 1024 ? 4 : 8;
  const int bpc = bytes_per_corr(variant);
  const size_t cap_bytes = h->smem_optin - kStaticSmemReserve;
  const long long want_elems = ((max_n + 1) + 1) & ~1LL;  // head element + round up to even
  long long cap_elems = std::min<long long>(want_elems, static_cast<long long>(cap_bytes / bpc) & ~1LL);
  if (cap_elems < 2) cap_elems = 2;
  a.cap_elems = static_cast<int>(cap_elems);
  a.use_bulk = bulk_ok(a.bv) ? 1 : 0;
  if (env_int("PNEC_B200_NO_BULK", 0)) a.use_bulk = 0;
  a.dbg = nullptr;
#ifdef PNEC_PHASE_TIMING
  static long long *dbg_buf = nullptr;
  if (!dbg_buf) cudaMalloc(&dbg_buf, sizeof(long long) * 12 * 1000000);
  a.dbg = dbg_buf;
#endif
  const size_t dyn = static_cast<size_t>(cap_elems) * bpc;
  int rc;
  switch (variant) {
    case PNEC_VARIANT_NEC: rc = launch_solve_v<PNEC_VARIANT_NEC>(h, a, nw, dyn, stream); break;
    case PNEC_VARIANT_TARGET: rc = launch_solve_v<PNEC_VARIANT_TARGET>(h, a, nw, dyn, stream); break;
    case PNEC_VARIANT_HOST: rc = launch_solve_v<PNEC_VARIANT_HOST>(h, a, nw, dyn, stream); break;
    default: rc = launch_solve_v<PNEC_VARIANT_SYMMETRIC>(h, a, nw, dyn, stream); break;
  }
#ifdef PNEC_PHASE_TIMING
  if (rc == PNEC_OK && env_int("PNEC_B200_DUMP_TIMING", 0)) dump_phase_timing(a);
#endif
  return rc;
}

template <int V, int NW, int S, int MINB>
int launch_eval_t(pnec_handle *h, const EvalArgs &a, cudaStream_t stream) {
  auto kern = eval_kernel<V, NW, S, MINB>;
  const size_t dyn = static_cast<size_t>(S) * NW * 32 * VariantTraits<V>::kDoubles * 8;
  PNEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(dyn)));
  kern<<<static_cast<unsigned>(a.bv.num_problems), NW * 32, dyn, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  return PNEC_OK;
}

template <int V, int WPC, int S, int CHUNK, int MINB, int T = 32>
int launch_eval_warp_t(pnec_handle *h, const EvalArgs &a, cudaStream_t stream) {
  auto kern = eval_warp_kernel<V, WPC, S, CHUNK, MINB, T>;
  const size_t dyn = static_cast<size_t>(WPC) * S * T * VariantTraits<V>::kDoubles * 8;
  PNEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(dyn)));
  const long long want = (a.bv.num_problems + WPC - 1) / WPC;
  const long long cap = static_cast<long long>(h->sm_count) * MINB;
  const unsigned grid = static_cast<unsigned>(std::max<long long>(1, std::min(want, cap)));
  kern<<<grid, WPC * 32, dyn, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  return PNEC_OK;
}

template <int V>
int launch_eval_v(pnec_handle *h, const EvalArgs &a, long long max_n, cudaStream_t stream) {
  (void)max_n;
  int cfg = env_int("PNEC_B200_EVAL_CFG", 0);
  if (cfg == 0) cfg = a.use_bulk ? 11 : 2;  // 3 stages of 32 correspondences per warp: measured best
  switch (cfg) {
    case 2: return launch_eval_t<V, 4, 4, 3>(h, a, stream);    // CTA per problem, 128-wide tiles
    case 10: return launch_eval_warp_t<V, 4, 4, 8, 3>(h, a, stream);  // warp-private, 4 stages
    case 11: return launch_eval_warp_t<V, 4, 3, 8, 3>(h, a, stream);  // warp-private, 3 stages (default)
    case 12: return launch_eval_warp_t<V, 4, 6, 8, 2>(h, a, stream);
    case 13: return launch_eval_warp_t<V, 8, 3, 8, 2>(h, a, stream);
    case 14: return launch_eval_warp_t<V, 4, 2, 8, 3, 64>(h, a, stream);   // 64-wide tiles, 2 stages
    case 15: return launch_eval_warp_t<V, 4, 3, 8, 3, 64>(h, a, stream);   // 64-wide tiles, 3 stages (smem: 2 CTAs)
    case 16: return launch_eval_warp_t<V, 4, 2, 8, 3, 128>(h, a, stream);  // 128-wide tiles
    case 17: return launch_eval_warp_t<V, 4, 2, 8, 3>(h, a, stream);       // 32-wide tiles, 2 stages
    default: return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown PNEC_B200_EVAL_CFG");
  }
}

int launch_eval(pnec_handle *h, const EvalArgs &a0, int variant, long long max_n,
                cudaStream_t stream) {
  EvalArgs a = a0;
  if (a.bv.num_problems == 0) return PNEC_OK;
  a.use_bulk = bulk_ok(a.bv) ? 1 : 0;
  if (env_int("PNEC_B200_NO_BULK", 0)) a.use_bulk = 0;
  switch (variant) {
    case PNEC_VARIANT_NEC: return launch_eval_v<PNEC_VARIANT_NEC>(h, a, max_n, stream);
    case PNEC_VARIANT_TARGET: return launch_eval_v<PNEC_VARIANT_TARGET>(h, a, max_n, stream);
    case PNEC_VARIANT_HOST: return launch_eval_v<PNEC_VARIANT_HOST>(h, a, max_n, stream);
    default: return launch_eval_v<PNEC_VARIANT_SYMMETRIC>(h, a, max_n, stream);
  }
}

}  // namespace

// ===================================================================== C-ABI

extern "C" {

int pnec_version(void) { return PNEC_B200_VERSION_MAJOR * 1000 + PNEC_B200_VERSION_MINOR; }

const char *pnec_last_error(void) { return g_last_error.c_str(); }

const char *pnec_status_string(int32_t status) {
  switch (status) {
    case PNEC_STATUS_CONVERGED_FUNCTION: return "converged: function tolerance";
    case PNEC_STATUS_CONVERGED_PARAMETER: return "converged: parameter tolerance";
    case PNEC_STATUS_CONVERGED_GRADIENT: return "converged: gradient tolerance";
    case PNEC_STATUS_CONVERGED_RADIUS: return "converged: minimum trust region radius";
    case PNEC_STATUS_MAX_ITERATIONS: return "no convergence: maximum iterations";
    case PNEC_STATUS_FAILURE: return "failure: consecutive invalid steps";
    case PNEC_STATUS_NONFINITE: return "failure: non-finite cost at the start point";
    case PNEC_STATUS_EMPTY: return "empty problem";
    default: return "unknown";
  }
}

void pnec_solver_opts_default(pnec_solver_opts *o) {
  if (!o) return;
  o->variant = PNEC_VARIANT_TARGET;
  o->max_num_iterations = 50;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->regularization = 1.0e-13;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
}

int pnec_create(int device, pnec_handle **out) {
  if (!out) return fail(PNEC_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(PNEC_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(PNEC_ERR_INVALID_ARGUMENT, "bad device index");
  PNEC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop{};
  PNEC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(PNEC_ERR_UNSUPPORTED, "pnec_b200 is built for sm_100a (Blackwell) only");
  pnec_handle *h = new (std::nothrow) pnec_handle();
  if (!h) return fail(PNEC_ERR_ALLOC, "out of host memory");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  *out = h;
  return PNEC_OK;
}

void pnec_destroy(pnec_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  DevBuf *bufs[] = {&h->d_f1, &h->d_f2, &h->d_ct, &h->d_ch, &h->d_off, &h->d_poses,
                    &h->d_out_poses, &h->d_out_status, &h->d_out_iters, &h->d_out_cost,
                    &h->d_out_init, &h->d_out_grad, &h->d_out_jtj, &h->d_ut_mu, &h->d_ut_cov,
                    &h->d_ut_out};
  for (DevBuf *b : bufs) b->release();
  delete h;
}

int64_t pnec_launch_count(const pnec_handle *h) { return h ? h->launches : 0; }

int pnec_solve_batch(pnec_handle *h, const pnec_batch *batch, const pnec_solver_opts *opts,
                     const pnec_solve_out *out, void *cuda_stream) {
  if (!h || !opts || !out) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  int rc = validate_batch(batch, opts->variant, true);
  if (rc != PNEC_OK) return rc;
  const long long B = batch->num_problems;
  if (B > 0 && !out->poses) return fail(PNEC_ERR_INVALID_ARGUMENT, "out->poses is NULL");
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  Staged st;
  rc = stage_batch(h, batch, opts->variant, stream, &st);
  if (rc != PNEC_OK) return rc;
  SolveArgs a{};
  a.bv = st.bv;
  a.o = *opts;
  const bool host = batch->memspace == PNEC_MEM_HOST;
  if (host) {
    PNEC_CUDA(h->d_out_poses.ensure(static_cast<size_t>(B) * 56));
    PNEC_CUDA(h->d_out_status.ensure(static_cast<size_t>(B) * 4));
    PNEC_CUDA(h->d_out_iters.ensure(static_cast<size_t>(B) * 4));
    PNEC_CUDA(h->d_out_cost.ensure(static_cast<size_t>(B) * 8));
    PNEC_CUDA(h->d_out_init.ensure(static_cast<size_t>(B) * 8));
    a.out_poses = static_cast<double *>(h->d_out_poses.p);
    a.out_status = out->status ? static_cast<int *>(h->d_out_status.p) : nullptr;
    a.out_iters = out->iterations ? static_cast<int *>(h->d_out_iters.p) : nullptr;
    a.out_cost = out->cost ? static_cast<double *>(h->d_out_cost.p) : nullptr;
    a.out_init_cost = out->initial_cost ? static_cast<double *>(h->d_out_init.p) : nullptr;
  } else {
    a.out_poses = out->poses;
    a.out_status = out->status;
    a.out_iters = out->iterations;
    a.out_cost = out->cost;
    a.out_init_cost = out->initial_cost;
  }
  rc = launch_solve(h, a, opts->variant, st.max_n, stream);
  if (rc != PNEC_OK) return rc;
  if (host) {
    PNEC_CUDA(cudaMemcpyAsync(out->poses, a.out_poses, static_cast<size_t>(B) * 56,
                              cudaMemcpyDeviceToHost, stream));
    if (out->status)
      PNEC_CUDA(cudaMemcpyAsync(out->status, a.out_status, static_cast<size_t>(B) * 4,
                                cudaMemcpyDeviceToHost, stream));
    if (out->iterations)
      PNEC_CUDA(cudaMemcpyAsync(out->iterations, a.out_iters, static_cast<size_t>(B) * 4,
                                cudaMemcpyDeviceToHost, stream));
    if (out->cost)
      PNEC_CUDA(cudaMemcpyAsync(out->cost, a.out_cost, static_cast<size_t>(B) * 8,
                                cudaMemcpyDeviceToHost, stream));
    if (out->initial_cost)
      PNEC_CUDA(cudaMemcpyAsync(out->initial_cost, a.out_init_cost, static_cast<size_t>(B) * 8,
                                cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
  }
  return PNEC_OK;
}

int pnec_eval_batch(pnec_handle *h, const pnec_batch *batch, int32_t variant,
                    double regularization, const pnec_eval_out *out, void *cuda_stream) {
  if (!h || !out) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  int rc = validate_batch(batch, variant, true);
  if (rc != PNEC_OK) return rc;
  const long long B = batch->num_problems;
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  Staged st;
  rc = stage_batch(h, batch, variant, stream, &st);
  if (rc != PNEC_OK) return rc;
  EvalArgs a{};
  a.bv = st.bv;
  a.reg = regularization;
  const bool host = batch->memspace == PNEC_MEM_HOST;
  if (host) {
    PNEC_CUDA(h->d_out_cost.ensure(static_cast<size_t>(B) * 8));
    PNEC_CUDA(h->d_out_grad.ensure(static_cast<size_t>(B) * 40));
    PNEC_CUDA(h->d_out_jtj.ensure(static_cast<size_t>(B) * 120));
    a.out_cost = out->cost ? static_cast<double *>(h->d_out_cost.p) : nullptr;
    a.out_grad = out->gradient ? static_cast<double *>(h->d_out_grad.p) : nullptr;
    a.out_jtj = out->jtj ? static_cast<double *>(h->d_out_jtj.p) : nullptr;
  } else {
    a.out_cost = out->cost;
    a.out_grad = out->gradient;
    a.out_jtj = out->jtj;
  }
  rc = launch_eval(h, a, variant, st.max_n, stream);
  if (rc != PNEC_OK) return rc;
  if (host) {
    if (out->cost)
      PNEC_CUDA(cudaMemcpyAsync(out->cost, a.out_cost, static_cast<size_t>(B) * 8,
                                cudaMemcpyDeviceToHost, stream));
    if (out->gradient)
      PNEC_CUDA(cudaMemcpyAsync(out->gradient, a.out_grad, static_cast<size_t>(B) * 40,
                                cudaMemcpyDeviceToHost, stream));
    if (out->jtj)
      PNEC_CUDA(cudaMemcpyAsync(out->jtj, a.out_jtj, static_cast<size_t>(B) * 120,
                                cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
  }
  return PNEC_OK;
}

int pnec_cost_function_batch(pnec_handle *h, const pnec_batch *batch, double *out_mean_energy,
                             void *cuda_stream) {
  if (!h || !out_mean_energy) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  int rc = validate_batch(batch, PNEC_VARIANT_TARGET, true);
  if (rc != PNEC_OK) return rc;
  const long long B = batch->num_problems;
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  Staged st;
  rc = stage_batch(h, batch, PNEC_VARIANT_TARGET, stream, &st);
  if (rc != PNEC_OK) return rc;
  const bool host = batch->memspace == PNEC_MEM_HOST;
  double *d_out = out_mean_energy;
  if (host) {
    PNEC_CUDA(h->d_out_cost.ensure(static_cast<size_t>(B) * 8));
    d_out = static_cast<double *>(h->d_out_cost.p);
  }
  cost_kernel<<<static_cast<unsigned>(B), 128, 0, stream>>>(st.bv, d_out);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  if (host) {
    PNEC_CUDA(cudaMemcpyAsync(out_mean_energy, d_out, static_cast<size_t>(B) * 8,
                              cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
  }
  return PNEC_OK;
}

int pnec_unscented_transform_batch(pnec_handle *h, int64_t n, int32_t memspace, const double *mus,
                                   const double *covs, const double *K_inv, double kappa,
                                   int32_t camera_model, double *out_covs, void *cuda_stream) {
  if (!h) return fail(PNEC_ERR_INVALID_ARGUMENT, "handle is NULL");
  if (n < 0) return fail(PNEC_ERR_INVALID_ARGUMENT, "n < 0");
  if (n == 0) return PNEC_OK;
  if (!mus || !covs || !out_covs) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL array");
  if (camera_model != PNEC_CAMERA_OMNIDIRECTIONAL && camera_model != PNEC_CAMERA_PINHOLE)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown camera model");
  if (camera_model == PNEC_CAMERA_PINHOLE && !K_inv)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "K_inv is NULL for a pinhole camera");
  if (memspace != PNEC_MEM_HOST && memspace != PNEC_MEM_DEVICE)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown memspace");
  std::lock_guard<std::mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  UtArgs a{};
  a.n = n;
  a.kappa = kappa;
  a.camera_model = camera_model;
  for (int k = 0; k < 9; ++k) a.Kinv[k] = K_inv ? K_inv[k] : ((k % 4 == 0) ? 1.0 : 0.0);
  const size_t nn = static_cast<size_t>(n);
  if (memspace == PNEC_MEM_HOST) {
    PNEC_CUDA(h->d_ut_mu.ensure(nn * 24));
    PNEC_CUDA(h->d_ut_cov.ensure(nn * 72));
    PNEC_CUDA(h->d_ut_out.ensure(nn * 72));
    PNEC_CUDA(cudaMemcpyAsync(h->d_ut_mu.p, mus, nn * 24, cudaMemcpyHostToDevice, stream));
    PNEC_CUDA(cudaMemcpyAsync(h->d_ut_cov.p, covs, nn * 72, cudaMemcpyHostToDevice, stream));
    a.mus = static_cast<const double *>(h->d_ut_mu.p);
    a.covs = static_cast<const double *>(h->d_ut_cov.p);
    a.out = static_cast<double *>(h->d_ut_out.p);
  } else {
    a.mus = mus;
    a.covs = covs;
    a.out = out_covs;
  }
  a.use_bulk = (aligned16(a.mus) && aligned16(a.covs) && aligned16(a.out) && !env_int("PNEC_B200_NO_BULK", 0)) ? 1 : 0;
  unscented_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  if (memspace == PNEC_MEM_HOST) {
    PNEC_CUDA(cudaMemcpyAsync(out_covs, a.out, nn * 72, cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
  }
  return PNEC_OK;
}

}  // extern "C"
