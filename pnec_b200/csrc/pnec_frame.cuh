// pnec_frame.cuh — the weighted rounds of PNEC::WeightedEigensolver (src/rel_pose_estimation/pnec.cc:283-348)
// for ONE frame pair per CTA, all rounds in one launch: rotation by the eigensolver LM (es_lm_group,
// four lanes of warp 0), translation by scan + SCF (scf_pair, the whole CTA), leaving the loop at the
// bitwise fixed point of the round map.  Same device functions and the same order of operations as
// the per-round kernels (pnec_eigensolver.cuh, pnec_translation.cuh) — results agree to the last bit
// or two (a different kernel, so the compiler may fuse a different product of a sum of products);
// used for small batches, where the ~27 per-round launches of that path cost more than the work.
#pragma once

#include "pnec_eigensolver.cuh"

namespace pnec {

struct FrameRoundsArgs {
  ScfArgs scf;             // bv / sphere / reg / samples / steps / cap_elems / spill / cache / fixed as for scf_kernel;
                           // poses, prev_poses and out_t are set per round by the kernel
  EsLmParams lm;
  const double *moments;   // [B][36] weighted moments
  const double *es_poses;  // [B][7]  round 0: the PNEC::Eigensolver pose
  double *rounds;          // [R][B][7] poses of rounds 1..R (scratch, as in the per-round path)
  double *final_poses;     // [B][7] pose after the last executed round
  long long num_problems;
  int num_rounds;
};

template <int NW>
__global__ void __launch_bounds__(NW * 32) frame_rounds_kernel(const __grid_constant__ FrameRoundsArgs args) {
  __shared__ double s_mom[kEsMom];
  __shared__ int s_rot_same;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = blockIdx.x, B = args.num_problems;
  if (tid < kEsMom) s_mom[tid] = args.moments[kEsMom * b + tid];
  if (tid == 0) s_rot_same = 0;
  __syncthreads();
  const double *prev = args.es_poses + 7 * b;
  for (int k = 1; k <= args.num_rounds; ++k) {
    double *cur = args.rounds + 7 * (static_cast<long long>(k - 1) * B + b);
    // ---- rotation: a rotation that repeated once repeats forever (the next solve would start from
    // the same bits), so it is solved for only until then
    if (warp == 0) {
      if (s_rot_same) {
        if (lane < 4) cur[lane] = prev[lane];
      } else {
        double x[3] = {prev[0] / prev[3], prev[1] / prev[3], prev[2] / prev[3]};  // rot2cayley
        int info, nfev;
        es_lm_group<true>(s_mom, 1, 0, args.lm, lane < 4, lane & 3, x, info, nfev);
        if (lane == 0) {
          const double sc = 1.0 / sqrt(1.0 + x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
          const double q[4] = {x[0] * sc, x[1] * sc, x[2] * sc, sc};
          bool same = args.scf.fixed != nullptr;  // shortcuts switched off: solve every round
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            same = same && __double_as_longlong(q[i]) == __double_as_longlong(prev[i]);
            cur[i] = q[i];
          }
          s_rot_same = same ? 1 : 0;
        }
      }
    }
    __syncthreads();
    // ---- translation
    ScfArgs a = args.scf;
    a.bv.poses = args.rounds + 7 * static_cast<long long>(k - 1) * B;
    a.prev_poses = (k == 1) ? args.es_poses : args.rounds + 7 * static_cast<long long>(k - 2) * B;
    a.out_t = const_cast<double *>(a.bv.poses) + 4;
    a.out_stride = 7;
    scf_pair<NW>(a, b);
    __syncthreads();
    prev = cur;
    if (args.scf.fixed && args.scf.fixed[b]) break;  // (R, t) -> itself: the remaining rounds repeat it
  }
  if (tid < 7) args.final_poses[7 * b + tid] = prev[tid];
}

}  // namespace pnec
