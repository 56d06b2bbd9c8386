// pnec_solve.cuh — whole-LM-solve kernels: shared-memory resident (solve_kernel) and streaming (solve_stream_kernel).
// Replace PNECCeres::Optimize (src/optimization/pnec_ceres.cc:70-168), NECCeres::Optimize
// (src/optimization/nec_ceres.cc:73-101) and ceres::Solve under them.
#pragma once

#include "pnec_batch.cuh"
#include "pnec_lm.cuh"

namespace pnec {

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
  unsigned int *work_counter;  // solve_slots_kernel: {next pair, CTAs that have left}; nullptr = static partition
  double *start_state;         // solve_slots_kernel: [B][10] (theta, phi, q, sin/cos of theta and phi), solve_prep_kernel
  const int *work_order;       // solve_slots_kernel: ticket k of the work counter is pair work_order[k] (nullptr: pair k)
};

// PNECCeres::InitValues(orientation, translation) (pnec_ceres.cc:188-192): the start point
// x = (theta, phi, q) of a pair and the sines / cosines of its angles.
__device__ __forceinline__ void solve_start_point(const double *pose, double x[6], double sc[4]) {
  angles_from_vec(pose + 4, x[0], x[1]);
  x[2] = pose[0]; x[3] = pose[1]; x[4] = pose[2]; x[5] = pose[3];
  sincos(x[0], &sc[0], &sc[1]);
  sincos(x[1], &sc[2], &sc[3]);
}

// The start state of the minimiser apart from the start point; called by one thread.
__device__ __forceinline__ void solve_reset_state(const SolveArgs &args, int n, LMState &st) {
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
}

// Start point + start state + pose constants; called by one thread.
__device__ __forceinline__ void solve_init_state(const SolveArgs &args, long long b, int n,
                                                 LMState &st, PoseConst &s_pc) {
  double *x = st.pts[0], *sc = st.scs[0];
  solve_start_point(args.bv.poses + 7 * b, x, sc);
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


template <int V, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) solve_kernel(const __grid_constant__ SolveArgs args) {
  constexpr int NT = NW * 32;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ PoseConst s_pc;
  __shared__ LMState s_lm;
  __shared__ double s_part[NW][kAccPad];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = args.work_order ? args.work_order[blockIdx.x] : blockIdx.x;  // (pnec_solve_slots.cuh: the order)
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
        if (warp == lmw) {
          lm_step(first, s_lm, o, lane, s_pc);
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
  const long long b = args.work_order ? args.work_order[blockIdx.x] : blockIdx.x;  // (pnec_solve_slots.cuh: the order)
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
        if (warp == lmw) {
          lm_step(first, s_lm, o, lane, s_pc);
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

}  // namespace pnec
