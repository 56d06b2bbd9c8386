// pnec_capi.cu — host side and C-ABI (include/pnec_b200.h) of the B200 PNEC frame-pair solver.
// The only translation unit: the sm_100a kernels live in the headers it includes.
//
//   pnec_solve.cuh   solve_kernel / solve_stream_kernel: whole LM solve per frame pair on device
//   pnec_eval.cuh    eval_warp_kernel (K1): fused residual + Jacobian + J^T J, the roofline kernel
//   pnec_aux.cuh     cost_kernel (parity metric), unscented_kernel (covariance propagation)
//   pnec_translation.cuh  scf_kernel / nec_translation_kernel: translation given rotation
//   pnec_eigensolver.cuh  es_moments_kernel / es_lm_kernel: rotation by NEC eigenvalue minimisation
//   pnec_frame.cuh   frame_rounds_kernel: all weighted rounds of one pair in one launch (small batches)
//   pnec_ransac.cuh  ransac_kernel: RANSAC over the eigensolver + inlier extraction
//   pnec_lm.cuh      Ceres-semantics Levenberg-Marquardt update; pnec_device.cuh: the math
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "pnec_aux.cuh"
#include "pnec_eigensolver.cuh"
#include "pnec_eval.cuh"
#include "pnec_frame.cuh"
#include "pnec_ransac.cuh"
#include "pnec_solve.cuh"
#include "pnec_solve_slots.cuh"
#include "pnec_translation.cuh"

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
      return fail(PNEC_ERR_CUDA, std::string(#call) + " (pnec_capi.cu:" + std::to_string(__LINE__) + \
                                 "): " + cudaGetErrorString(err__));                         \
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

// Tuning / debugging switches (PNEC_B200_* environment variables), read ONCE when a handle is
// created: no getenv on the call path (a single-pair solve is a 0.1 ms affair).
struct Config {
  int dump_timing = 0;
  int frame_chunks = 0;
  int fused_rounds_max_pairs = 512;
  int h2d_chunks = 0;
  int no_bulk = 0;
  int no_frame_shortcuts = 0;
  int no_lm_ahead = 0;
  int no_stager = 0;
  int ransac_warps = 0;
  int solve_order = 1;         // refinement: start the worst-conditioned pairs first
  int solve_order_min_pairs = -1;  // ... in the kernels with a CTA per pair from this many pairs (-1: 32 per SM)
  int es_wide_max_pairs = -1;  // rotation LM: two pairs per warp up to this many pairs (-1: by SM count)
  int ransac_defer = -1;  // PNEC_B200_RANSAC_DEFER: iterations after which pass 1 hands a pair to pass 2 (0: never)
  int ransac_split = 1;   // PNEC_B200_RANSAC_SPLIT: pass 1 of large batches as three kernels per round (0: one kernel)
  int scf_debug = 0;
  int scf_defer = 48;
  int scf_warps = 0;
  int solve_warps = 0;
  int solve_slots = 1;  // PNEC_B200_SOLVE_SLOTS: 0 never, 1 large batches (default), 2 whenever the pairs fit
  int stream_min_n = 896;
  int copy_threads = 0;  // PNEC_B200_COPY_THREADS: worker threads of the pageable-input stager (0 = auto)
  void load() {
    dump_timing = env_int("PNEC_B200_DUMP_TIMING", 0);
    es_wide_max_pairs = env_int("PNEC_B200_ES_WIDE_MAX_PAIRS", -1);
    frame_chunks = env_int("PNEC_B200_FRAME_CHUNKS", 0);
    fused_rounds_max_pairs = env_int("PNEC_B200_FUSED_ROUNDS_MAX_PAIRS", 512);
    h2d_chunks = env_int("PNEC_B200_H2D_CHUNKS", 0);
    no_bulk = env_int("PNEC_B200_NO_BULK", 0);
    no_frame_shortcuts = env_int("PNEC_B200_NO_FRAME_SHORTCUTS", 0);
    no_lm_ahead = env_int("PNEC_B200_NO_LM_AHEAD", 0);
    no_stager = env_int("PNEC_B200_NO_STAGER", 0);
    ransac_warps = env_int("PNEC_B200_RANSAC_WARPS", 0);
    ransac_defer = env_int("PNEC_B200_RANSAC_DEFER", -1);
    ransac_split = env_int("PNEC_B200_RANSAC_SPLIT", 1);
    scf_debug = env_int("PNEC_B200_SCF_DEBUG", 0);
    scf_defer = env_int("PNEC_B200_SCF_DEFER", 48);
    scf_warps = env_int("PNEC_B200_SCF_WARPS", 0);
    solve_warps = env_int("PNEC_B200_SOLVE_WARPS", 0);
    solve_slots = env_int("PNEC_B200_SOLVE_SLOTS", 1);
    solve_order = env_int("PNEC_B200_SOLVE_ORDER", 1);
    solve_order_min_pairs = env_int("PNEC_B200_SOLVE_ORDER_MIN_PAIRS", -1);
    stream_min_n = env_int("PNEC_B200_STREAM_MIN_N", 896);
    copy_threads = env_int("PNEC_B200_COPY_THREADS", 0);
  }
};

// ---------------------------------------------------------- pageable host inputs
//
// The reference hands over std::vector storage, i.e. pageable memory, which cudaMemcpyAsync moves
// through the driver's own bounce buffer at ~11 GB/s.  HostStager copies such inputs into a ring of
// pinned slots with a few worker threads (a parallel memcpy runs at several times that) and sends
// every slot with its own async copy, so the DMA of one slot overlaps the memcpy of the next.
class CopyPool {
 public:
  explicit CopyPool(int threads) : n_(std::max(1, threads)) {
    for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
  // memcpy(dst, src, bytes) split over the pool; returns when done
  void copy(void *dst, const void *src, size_t bytes) {
    if (n_ == 1 || bytes < (1u << 20)) {
      std::memcpy(dst, src, bytes);
      return;
    }
    {
      std::lock_guard<std::mutex> l(mu_);
      dst_ = static_cast<char *>(dst);
      src_ = static_cast<const char *>(src);
      bytes_ = bytes;
      pending_ = n_ - 1;
      ++gen_;
    }
    cv_.notify_all();
    part(0);
    std::unique_lock<std::mutex> l(mu_);
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void part(int i) {
    const size_t per = ((bytes_ + n_ - 1) / n_ + 4095) & ~size_t(4095);
    const size_t b0 = std::min(bytes_, per * i), b1 = std::min(bytes_, per * (i + 1));
    if (b1 > b0) std::memcpy(dst_ + b0, src_ + b0, b1 - b0);
  }
  void loop(int i) {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      part(i);
      {
        std::lock_guard<std::mutex> l(mu_);
        --pending_;
      }
      done_.notify_one();
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  unsigned long long gen_ = 0;
  int pending_ = 0;
  bool stop_ = false;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t bytes_ = 0;
};

class HostStager {
 public:
  static constexpr int kSlots = 4;
  static constexpr size_t kSlotBytes = 8u << 20;
  int copy_threads = 0;  // 0 = auto
  ~HostStager() {
    for (int i = 0; i < kSlots; ++i) {
      if (ev_[i]) cudaEventDestroy(ev_[i]);
    }
    if (ring_) cudaFreeHost(ring_);
  }
  static bool pageable(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
  }
  // dst (device) <- src (pageable host), enqueued on `stream`; returns when the last slot is queued
  cudaError_t h2d(void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    if (!ring_) {
      cudaError_t e = cudaHostAlloc(&ring_, kSlots * kSlotBytes, cudaHostAllocDefault);
      if (e != cudaSuccess) return e;
      for (int i = 0; i < kSlots; ++i) {
        e = cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
      }
      int t = static_cast<int>(std::thread::hardware_concurrency());
      if (copy_threads > 0) t = 2 * copy_threads;
      pool_.reset(new CopyPool(std::min(8, std::max(1, t / 2))));
    }
    for (size_t off = 0; off < bytes; off += kSlotBytes) {
      const size_t n = std::min(kSlotBytes, bytes - off);
      char *slot = static_cast<char *>(ring_) + static_cast<size_t>(next_) * kSlotBytes;
      if (used_[next_]) {
        cudaError_t e = cudaEventSynchronize(ev_[next_]);  // the slot's previous DMA has drained
        if (e != cudaSuccess) return e;
      }
      pool_->copy(slot, static_cast<const char *>(src) + off, n);
      cudaError_t e = cudaMemcpyAsync(static_cast<char *>(dst) + off, slot, n, cudaMemcpyHostToDevice, stream);
      if (e != cudaSuccess) return e;
      e = cudaEventRecord(ev_[next_], stream);
      if (e != cudaSuccess) return e;
      used_[next_] = true;
      next_ = (next_ + 1) % kSlots;
    }
    return cudaSuccess;
  }

 private:
  void *ring_ = nullptr;
  cudaEvent_t ev_[kSlots] = {};
  bool used_[kSlots] = {};
  int next_ = 0;
  std::unique_ptr<CopyPool> pool_;
};

}  // namespace

struct pnec_handle {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  size_t smem_per_sm = 0;
  int64_t launches = 0;
  Config cfg;
  // staging for HOST-memspace calls and for the device copy of offsets
  DevBuf d_f1, d_f2, d_ct, d_ch, d_off, d_poses;
  DevBuf d_out_poses, d_out_status, d_out_iters, d_out_cost, d_out_init, d_out_grad, d_out_jtj;
  DevBuf d_ut_mu, d_ut_cov, d_ut_out, d_kp_bv, d_sphere, d_tr_out, d_tr_aux;
  DevBuf d_es_mom, d_es_w, d_es_info, d_es_ev, d_fr_es, d_fr_a;  // eigensolver / frame pipeline
  DevBuf d_fr_cache, d_fr_flags;  // ScfScanCache[B]; int q_same[B], fixed[B]
  DevBuf d_scf_defer;             // int count, cursor, pad[2], list[B]  (standalone SCF calls)
  DevBuf d_fr_defer;              // the same per chunk of a frame solve: (4 + B) ints
  DevBuf d_scf_spill;             // [total][9] SCF terms of pairs that do not fit shared memory
  DevBuf d_rs_f1, d_rs_f2, d_rs_ct;  // RANSAC: inliers compacted to the front of every pair's slot
  DevBuf d_rs_best, d_rs_cnt, d_rs_iters, d_rs_idx;  // winning models [B][7], counts, iterations, indices
  DevBuf d_rs_state, d_rs_defer, d_rs_prefix, d_rs_hyp;  // pass 2 of the RANSAC stage (pnec_ransac.cuh)
  DevBuf d_rs_sp_mom, d_rs_sp_x, d_rs_sp_int;            // pass 1 as three kernels per round
  DevBuf d_kp_hp, d_kp_tp, d_kp_hc, d_kp_tc, d_kp_hi, d_kp_ti;  // keypoint tables and match indices (HOST callers)
  DevBuf d_pk_ct, d_pk_ch;  // PNEC_COV_PACKED covariances of HOST callers before expansion
  DevBuf d_slot_ctr;        // work counters of solve_slots_kernel, one {next pair, CTAs gone} per stream seen
  std::vector<cudaStream_t> slot_ctr_streams;
  std::vector<DevBuf> d_slot_start;  // start points [B][10] of solve_prep_kernel, one buffer per stream seen
  DevBuf d_kp_out[8];  // device outputs of the from-keypoints entry points for HOST callers
  static constexpr int kMaxChunks = 8;
  static constexpr int kMaxRounds = 64;
  cudaStream_t side[kMaxChunks] = {}, lm_side[kMaxChunks] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxChunks] = {}, ev_es[kMaxChunks] = {};
  cudaEvent_t ev_round[kMaxChunks][kMaxRounds] = {};
  cudaEvent_t ev_stage[4] = {};  // stage timings of a frame solve (out->stage_ms)
  cudaEvent_t ev_last = nullptr; // end of the previous call's device work (see CallScope)
  cudaStream_t last_stream = nullptr;
  DevBuf d_fr_rounds;             // poses of every weighted round: [rounds][B][7]
  HostStager stager;              // pinned ring + copy threads for pageable host inputs
  int sphere_samples = -1;
  std::recursive_mutex mu;  // the keypoint entry points call the batch entry points on the same handle
};

namespace {

// Every call of a handle uses the same scratch buffers, and DEVICE-memspace calls return before their
// kernels have run: a later call on ANOTHER stream must not touch the scratch while they are in
// flight.  The end of every call is marked with an event on its stream and a call on a different
// stream waits for it first (same stream: stream order already serialises).  If a call fails after it
// has forked work onto the handle's side streams, the device is synchronised before returning, so that
// no copy into or out of the caller's buffers is still in flight.
struct CallScope {
  pnec_handle *h;
  cudaStream_t stream;
  bool forked = false, ok = false;
  CallScope(pnec_handle *h_, cudaStream_t s) : h(h_), stream(s) {
    if (h->ev_last && h->last_stream != stream) cudaStreamWaitEvent(stream, h->ev_last, 0);
  }
  ~CallScope() {
    if (forked && !ok) cudaDeviceSynchronize();
    if (!h->ev_last && cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming) != cudaSuccess) {
      h->ev_last = nullptr;
      cudaGetLastError();
      return;
    }
    cudaEventRecord(h->ev_last, stream);
    h->last_stream = stream;
  }
};

struct Staged {
  BatchView bv;
  long long max_n = 0;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-kernel maximum that belongs to the device
// context, not to a handle: the bookkeeping is process-wide (per device and kernel), and the limit is
// only ever raised -- a second handle asking for less must not lower it under a launch of the first.
template <class Kernel>
cudaError_t ensure_dyn_smem(pnec_handle *h, Kernel kern, size_t dyn) {
  static std::mutex mu;
  static std::unordered_map<const void *, size_t> granted[64];
  const void *key = reinterpret_cast<const void *>(kern);
  std::lock_guard<std::mutex> lock(mu);
  auto &map = granted[h->device & 63];
  auto it = map.find(key);
  if (it != map.end() && it->second >= dyn) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn));
  if (e == cudaSuccess) map[key] = dyn;
  return e;
}

int validate_batch(const pnec_batch *b, int variant, bool need_poses) {
  if (!b) return fail(PNEC_ERR_INVALID_ARGUMENT, "batch is NULL");
  if (b->num_problems < 0) return fail(PNEC_ERR_INVALID_ARGUMENT, "num_problems < 0");
  if (!b->offsets && b->n_per_problem < 0)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "n_per_problem < 0");
  if (b->memspace != PNEC_MEM_HOST && b->memspace != PNEC_MEM_DEVICE)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown memspace");
  if (b->cov_layout != PNEC_COV_FULL && b->cov_layout != PNEC_COV_PACKED)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown cov_layout");
  if (variant < PNEC_VARIANT_NEC || variant > PNEC_VARIANT_SYMMETRIC)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown residual variant");
  // grids are one CTA per frame pair and the kernels index a pair's correspondences with int
  if (b->num_problems > 0x7fffffffLL) return fail(PNEC_ERR_INVALID_ARGUMENT, "more than 2^31 - 1 frame pairs");
  long long total = 0;
  if (b->offsets) {
    if (b->offsets[0] < 0) return fail(PNEC_ERR_INVALID_ARGUMENT, "offsets[0] < 0");
    for (int64_t i = 0; i < b->num_problems; ++i) {
      if (b->offsets[i + 1] < b->offsets[i])
        return fail(PNEC_ERR_INVALID_ARGUMENT, "offsets must be non-decreasing");
      if (b->offsets[i + 1] - b->offsets[i] > 0x7ffffff0LL)
        return fail(PNEC_ERR_INVALID_ARGUMENT, "a frame pair has more than 2^31 - 16 correspondences");
    }
    total = b->offsets[b->num_problems];
  } else {
    if (b->n_per_problem > 0x7ffffff0LL)
      return fail(PNEC_ERR_INVALID_ARGUMENT, "n_per_problem exceeds 2^31 - 16");
    if (b->num_problems > 0 && b->n_per_problem > (0x7fffffffffffffffLL / 72) / b->num_problems)
      return fail(PNEC_ERR_INVALID_ARGUMENT, "num_problems * n_per_problem overflows");
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

// Device view of a batch.  HOST batches: device buffers are (re)allocated and the view points into
// them, but nothing is copied yet (stage_copy).  The (host) offsets are always copied.
int stage_alloc(pnec_handle *h, const pnec_batch *b, int variant, cudaStream_t stream, Staged *out) {
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
  bv.total = total;
  bv.n_uniform = b->offsets ? 0 : b->n_per_problem;
  if (b->offsets) {
    PNEC_CUDA(h->d_off.ensure(sizeof(long long) * (B + 1)));
    PNEC_CUDA(cudaMemcpyAsync(h->d_off.p, b->offsets, sizeof(long long) * (B + 1),
                              cudaMemcpyHostToDevice, stream));
    bv.offsets = static_cast<const long long *>(h->d_off.p);
  }
  const bool need_ct = variant != PNEC_VARIANT_NEC;
  const bool need_ch = variant == PNEC_VARIANT_SYMMETRIC;
  const bool packed = b->cov_layout == PNEC_COV_PACKED;
  if (b->memspace == PNEC_MEM_HOST) {
    const size_t nel = static_cast<size_t>(total);
    PNEC_CUDA(h->d_f1.ensure(nel * 24));
    PNEC_CUDA(h->d_f2.ensure(nel * 24));
    PNEC_CUDA(h->d_poses.ensure(static_cast<size_t>(B) * 56));
    if (need_ct) PNEC_CUDA(h->d_ct.ensure(nel * 72));
    if (need_ch) PNEC_CUDA(h->d_ch.ensure(nel * 72));
    if (packed && need_ct) PNEC_CUDA(h->d_pk_ct.ensure(nel * 48));
    if (packed && need_ch) PNEC_CUDA(h->d_pk_ch.ensure(nel * 48));
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
    if (packed) {  // expanded into the handle's buffers by stage_copy
      const size_t nel = static_cast<size_t>(std::max<long long>(total, 1));
      if (need_ct) { PNEC_CUDA(h->d_ct.ensure(nel * 72)); bv.ct = static_cast<const double *>(h->d_ct.p); }
      if (need_ch) { PNEC_CUDA(h->d_ch.ensure(nel * 72)); bv.ch = static_cast<const double *>(h->d_ch.p); }
    }
  }
  out->bv = bv;
  out->max_n = max_n;
  return PNEC_OK;
}

// H2D of pairs [p0, p1) of a HOST batch into the buffers of stage_alloc; PNEC_COV_PACKED covariances
// (HOST or DEVICE) are expanded to the 3x3 layout the kernels stream.  Otherwise a no-op for DEVICE batches.
int stage_copy(pnec_handle *h, const pnec_batch *b, int variant, long long p0, long long p1,
               cudaStream_t stream) {
  const bool host = b->memspace == PNEC_MEM_HOST, packed = b->cov_layout == PNEC_COV_PACKED;
  if ((!host && !packed) || p1 <= p0) return PNEC_OK;
  const long long e0 = b->offsets ? b->offsets[p0] : p0 * b->n_per_problem;
  const long long e1 = b->offsets ? b->offsets[p1] : p1 * b->n_per_problem;
  const size_t nel = static_cast<size_t>(e1 - e0);
  auto at = [](void *base, long long bytes) { return static_cast<char *>(base) + bytes; };
  // pinned sources go straight to the copy engine; large pageable ones through the pinned ring
  auto h2d = [&](void *dst, const double *src, size_t bytes) -> cudaError_t {
    if (bytes >= (4u << 20) && !h->cfg.no_stager && HostStager::pageable(src))
      return h->stager.h2d(dst, src, bytes, stream);
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
  };
  auto covs = [&](const double *src, DevBuf &pk, DevBuf &full) -> int {
    if (!packed) {
      PNEC_CUDA(h2d(at(full.p, e0 * 72), src + 9 * e0, nel * 72));
      return PNEC_OK;
    }
    const double *dsrc = src;
    if (host) {
      PNEC_CUDA(h2d(at(pk.p, e0 * 48), src + 6 * e0, nel * 48));
      dsrc = static_cast<const double *>(pk.p);
    }
    expand_covs_kernel<<<static_cast<unsigned>((nel + 255) / 256), 256, 0, stream>>>(dsrc, static_cast<double *>(full.p), e0,
                                                                                  static_cast<long long>(nel));
    PNEC_CUDA(cudaGetLastError());
    h->launches++;
    return PNEC_OK;
  };
  if (nel) {
    if (host) {
      PNEC_CUDA(h2d(at(h->d_f1.p, e0 * 24), b->bvs_host + 3 * e0, nel * 24));
      PNEC_CUDA(h2d(at(h->d_f2.p, e0 * 24), b->bvs_target + 3 * e0, nel * 24));
    }
    int rc;
    if (variant != PNEC_VARIANT_NEC && (rc = covs(b->covs_target, h->d_pk_ct, h->d_ct)) != PNEC_OK) return rc;
    if (variant == PNEC_VARIANT_SYMMETRIC && (rc = covs(b->covs_host, h->d_pk_ch, h->d_ch)) != PNEC_OK) return rc;
  }
  if (host && b->poses)
    PNEC_CUDA(cudaMemcpyAsync(at(h->d_poses.p, p0 * 56), b->poses + 7 * p0, static_cast<size_t>(p1 - p0) * 56,
                              cudaMemcpyHostToDevice, stream));
  return PNEC_OK;
}

// Builds the device view of a batch: copies H2D for HOST batches, always copies
// the (host) offsets.  Everything is enqueued on `stream`.
int stage_batch(pnec_handle *h, const pnec_batch *b, int variant, cudaStream_t stream,
                Staged *out) {
  int rc = stage_alloc(h, b, variant, stream, out);
  if (rc != PNEC_OK) return rc;
  return stage_copy(h, b, variant, 0, b->num_problems, stream);
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
  PNEC_CUDA(ensure_dyn_smem(h, kern, dyn));
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
  PNEC_CUDA(ensure_dyn_smem(h, kern, dyn));
  kern<<<static_cast<unsigned>(a.bv.num_problems), NW * 32, dyn, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  return PNEC_OK;
}

template <int V>
int launch_solve_stream_v(pnec_handle *h, const SolveArgs &a, cudaStream_t stream) {
  // 3 CTAs x 4 warps per SM, 4-stage rings (measured best of the round-1 sweep over 1 .. 12 warps per CTA)
  return launch_solve_stream_t<V, 4, 4, 3>(h, a, stream);
}

// solve_slots_kernel draws its pairs from a device counter that the kernel itself rewinds when its last
// CTA leaves.  Launches on one stream run one after the other and can share a counter; every stream the
// handle sees gets its own (64 of them; beyond that the kernel falls back to a static partition).
constexpr int kSlotCounters = 64;
int slot_counter(pnec_handle *h, cudaStream_t stream, unsigned int **out, DevBuf **start) {
  *out = nullptr;
  *start = nullptr;
  if (h->d_slot_start.empty()) h->d_slot_start.resize(kSlotCounters + 1);
  if (!h->d_slot_ctr.p) {
    PNEC_CUDA(h->d_slot_ctr.ensure(kSlotCounters * 2 * sizeof(unsigned int)));
    PNEC_CUDA(cudaMemset(h->d_slot_ctr.p, 0, kSlotCounters * 2 * sizeof(unsigned int)));
    PNEC_CUDA(cudaDeviceSynchronize());
  }
  size_t i = 0;
  for (; i < h->slot_ctr_streams.size(); ++i)
    if (h->slot_ctr_streams[i] == stream) break;
  if (i == h->slot_ctr_streams.size()) {
    if (i == kSlotCounters) return PNEC_OK;  // (the caller takes the one-CTA-per-pair kernel)
    h->slot_ctr_streams.push_back(stream);
  }
  *out = static_cast<unsigned int *>(h->d_slot_ctr.p) + 2 * i;
  *start = &h->d_slot_start[i];
  return PNEC_OK;
}

// Two slots per CTA, two CTAs per SM: the capacity of a slot in correspondences (a multiple of 32), 0 if
// the kernel cannot be resident twice.
template <int V>
int slots_capacity(pnec_handle *h, int *static_smem) {
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, solve_slots_kernel<V, 2, 1>) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  *static_smem = static_cast<int>(fa.sharedSizeBytes);
  const long long per_cta = static_cast<long long>(h->smem_per_sm) / 2 - 1024 - static_cast<long long>(fa.sharedSizeBytes);
  const long long cap = per_cta / (2LL * SlotLayout<V>::kDoubles * 8);
  return static_cast<int>(std::max<long long>(0, cap & ~31LL));
}

// Start points (solve_prep_kernel) and the order the pairs are started in (worst-conditioned first,
// pnec_solve_slots.cuh) in the stream's scratch: start points [B][10], then order [B], bins [B],
// histogram and cursors (ints).  a.work_order stays nullptr when the order is switched off.
template <int V>
int prepare_start_and_order(pnec_handle *h, SolveArgs &a, DevBuf *start, cudaStream_t stream) {
  const size_t nbp = static_cast<size_t>(a.bv.num_problems);
  PNEC_CUDA(start->ensure(nbp * 80 + nbp * 8 + 2 * kOrderBins * sizeof(int)));
  a.start_state = static_cast<double *>(start->p);
  solve_prep_kernel<<<static_cast<unsigned>((a.bv.num_problems + 127) / 128), 128, 0, stream>>>(a.bv.poses, a.bv.num_problems, a.start_state);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  a.work_order = nullptr;
  if (h->cfg.solve_order && a.bv.num_problems < 0x7fffffffLL) {
    int *order = reinterpret_cast<int *>(a.start_state + 10 * nbp), *bins = order + nbp, *hist = bins + nbp, *cursor = hist + kOrderBins;
    PNEC_CUDA(cudaMemsetAsync(hist, 0, 2 * kOrderBins * sizeof(int), stream));
    solve_score_kernel<V><<<static_cast<unsigned>((a.bv.num_problems + 3) / 4), 128, 0, stream>>>(a.bv, a.o.regularization, a.start_state, bins, hist);
    solve_order_scatter_kernel<<<static_cast<unsigned>((a.bv.num_problems + 255) / 256), 256, 0, stream>>>(bins, hist, cursor, order, a.bv.num_problems);
    PNEC_CUDA(cudaGetLastError());
    h->launches += 2;
    a.work_order = order;
  }
  return PNEC_OK;
}

template <int V>
int launch_solve_slots_v(pnec_handle *h, SolveArgs a, long long need, cudaStream_t stream, bool *done) {
  static int cap_max = -1, static_smem = 0;  // per process: the same kernel image on every device of a box
  if (cap_max < 0) cap_max = slots_capacity<V>(h, &static_smem);
  const long long cap = (need + 31) & ~31LL;
  if (cap > cap_max) return PNEC_OK;
  // (one CTA per SM with four slots and two evaluation groups serving all of them -- solve_slots_kernel<V, 4, 2>
  // -- was measured too: the same 0.485 ms on C2, so the queueing for an evaluation group is not what binds)
  constexpr int P = 2;
  const size_t dyn = static_cast<size_t>(P) * SlotLayout<V>::kDoubles * 8 * static_cast<size_t>(cap);
  PNEC_CUDA(ensure_dyn_smem(h, solve_slots_kernel<V, 2, 1>, dyn));
  DevBuf *start = nullptr;
  int rc = slot_counter(h, stream, &a.work_counter, &start);
  if (rc != PNEC_OK) return rc;
  if (!start) return PNEC_OK;
  rc = prepare_start_and_order<V>(h, a, start, stream);
  if (rc != PNEC_OK) return rc;
  a.cap_elems = static_cast<int>(cap);
  a.use_bulk = 1;
  a.dbg = nullptr;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(2LL * h->sm_count, (a.bv.num_problems + P - 1) / P));
  solve_slots_kernel<V, 2, 1><<<grid, (kSlotEvalWarps + 2) * 32, dyn, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
#ifdef PNEC_SLOT_TIMING
  if (h->cfg.dump_timing) {
    cudaDeviceSynchronize();
    unsigned long long pr[16], zero[16] = {0};
    cudaMemcpyFromSymbol(pr, g_slot_probe, sizeof(pr));
    cudaMemcpyToSymbol(g_slot_probe, zero, sizeof(zero));
    const double B = static_cast<double>(a.bv.num_problems);
    std::fprintf(stderr, "[slots init] issue+claim %.0f state %.0f | warp lifetimes, cycles: owner %.0f eval %.0f (grid %u)\n", pr[10] / B, pr[11] / B,
                 (double)pr[12] / (2.0 * grid), (double)pr[13] / (4.0 * grid), grid);
    std::fprintf(stderr, "[slots, cycles per pair] init %.0f load1 %.0f repack %.0f load2 %.0f wait-eval %.0f lm %.0f total %.0f | "
                 "eval warps (per warp, per pair): busy %.0f poll %.0f | passes %.2f\n",
                 pr[9] / B, pr[0] / B, pr[1] / B, pr[2] / B, pr[3] / B, pr[4] / B, pr[5] / B,
                 pr[6] / B / kSlotEvalWarps, pr[7] / B / kSlotEvalWarps, pr[8] / B);
#ifdef PNEC_PHASE_TIMING
    unsigned long long probe[8], zero8[8] = {0};
    cudaMemcpyFromSymbol(probe, g_lm_probe, sizeof(probe));
    cudaMemcpyToSymbol(g_lm_probe, zero8, sizeof(zero8));
    std::fprintf(stderr, "[lm_step probes, cycles per pair] judge %.0f bookkeeping %.0f tr-step %.0f candidate %.0f store %.0f\n",
                 (double)probe[0] / B, (double)probe[1] / B, (double)probe[2] / B, (double)probe[3] / B, (double)probe[4] / B);
#endif
  }
#endif
  *done = true;
  return PNEC_OK;
}

// The kernels with a CTA per pair take their pair from the same order (the hardware starts CTAs in
// blockIdx order) once the batch is several waves deep; below that the three extra launches cost more
// than a late straggler.
int order_for_cta_kernels(pnec_handle *h, SolveArgs &a, int variant, cudaStream_t stream) {
  a.work_order = nullptr;
  const long long min_pairs = h->cfg.solve_order_min_pairs >= 0 ? h->cfg.solve_order_min_pairs : 32LL * h->sm_count;
  if (!h->cfg.solve_order || a.bv.num_problems < min_pairs) return PNEC_OK;
  unsigned int *ctr = nullptr;
  DevBuf *start = nullptr;
  int rc = slot_counter(h, stream, &ctr, &start);
  if (rc != PNEC_OK || !start) return rc;
  switch (variant) {
    case PNEC_VARIANT_NEC: return prepare_start_and_order<PNEC_VARIANT_NEC>(h, a, start, stream);
    case PNEC_VARIANT_TARGET: return prepare_start_and_order<PNEC_VARIANT_TARGET>(h, a, start, stream);
    case PNEC_VARIANT_HOST: return prepare_start_and_order<PNEC_VARIANT_HOST>(h, a, start, stream);
    default: return prepare_start_and_order<PNEC_VARIANT_SYMMETRIC>(h, a, start, stream);
  }
}

int launch_solve(pnec_handle *h, const SolveArgs &a0, int variant, long long max_n,
                 cudaStream_t stream) {
  SolveArgs a = a0;
  if (a.bv.num_problems == 0) return PNEC_OK;
  // Batches that oversubscribe the device many times over, of pairs for which solve_kernel has room for
  // three per SM only (4 warps per pair, N > 448) but four slots fit: evaluation and LM update decoupled
  // (pnec_solve_slots.cuh; measured on B200: 10000 x 512 0.449 vs 0.473 ms).  Smaller pairs are better
  // off with one warp per pair and seven or more pairs per SM, smaller batches with a CTA per pair.
  const int slots_mode = h->cfg.solve_slots;
  if (slots_mode != 0 && bulk_ok(a.bv) && !h->cfg.no_bulk && max_n > 0 &&
      (slots_mode == 2 || (max_n > 448 && a.bv.num_problems >= 32LL * h->sm_count))) {
    // a pair starts at an even correspondence index of its arrays (16-byte alignment of the bulk copies):
    // one more element when its range starts at an odd one
    const bool even_starts = !a.bv.offsets && (a.bv.n_uniform % 2 == 0);
    const long long need = max_n + (even_starts ? 0 : 1);
    bool done = false;
    int rc;
    switch (variant) {
      case PNEC_VARIANT_NEC: rc = launch_solve_slots_v<PNEC_VARIANT_NEC>(h, a, need, stream, &done); break;
      case PNEC_VARIANT_TARGET: rc = launch_solve_slots_v<PNEC_VARIANT_TARGET>(h, a, need, stream, &done); break;
      case PNEC_VARIANT_HOST: rc = launch_solve_slots_v<PNEC_VARIANT_HOST>(h, a, need, stream, &done); break;
      default: rc = launch_solve_slots_v<PNEC_VARIANT_SYMMETRIC>(h, a, need, stream, &done); break;
    }
    if (rc != PNEC_OK || done) return rc;
  }
  // Large frame pairs: stream every pass (bulk-copy rings) instead of keeping one pair per SM
  // resident.  Needs 16-byte aligned arrays; SYMMETRIC (192 B / correspondence) stays resident-first.
  const long long stream_min_n = h->cfg.stream_min_n;
  if (bulk_ok(a.bv) && !h->cfg.no_bulk && max_n > stream_min_n) {
    a.use_bulk = 1;
    a.cap_elems = 0;
    a.dbg = nullptr;
    const int orc = order_for_cta_kernels(h, a, variant, stream);
    if (orc != PNEC_OK) return orc;
    switch (variant) {
      case PNEC_VARIANT_NEC: return launch_solve_stream_v<PNEC_VARIANT_NEC>(h, a, stream);
      case PNEC_VARIANT_TARGET: return launch_solve_stream_v<PNEC_VARIANT_TARGET>(h, a, stream);
      case PNEC_VARIANT_HOST: return launch_solve_stream_v<PNEC_VARIANT_HOST>(h, a, stream);
      default: return launch_solve_stream_v<PNEC_VARIANT_SYMMETRIC>(h, a, stream);
    }
  }
  int nw = h->cfg.solve_warps;
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
  if (h->cfg.no_bulk) a.use_bulk = 0;
  a.dbg = nullptr;
#ifdef PNEC_PHASE_TIMING
  static long long *dbg_buf = nullptr;
  if (!dbg_buf) cudaMalloc(&dbg_buf, sizeof(long long) * 12 * 1000000);
  a.dbg = dbg_buf;
#endif
  const size_t dyn = static_cast<size_t>(cap_elems) * bpc;
  int rc;
  rc = order_for_cta_kernels(h, a, variant, stream);
  if (rc != PNEC_OK) return rc;
  switch (variant) {
    case PNEC_VARIANT_NEC: rc = launch_solve_v<PNEC_VARIANT_NEC>(h, a, nw, dyn, stream); break;
    case PNEC_VARIANT_TARGET: rc = launch_solve_v<PNEC_VARIANT_TARGET>(h, a, nw, dyn, stream); break;
    case PNEC_VARIANT_HOST: rc = launch_solve_v<PNEC_VARIANT_HOST>(h, a, nw, dyn, stream); break;
    default: rc = launch_solve_v<PNEC_VARIANT_SYMMETRIC>(h, a, nw, dyn, stream); break;
  }
#ifdef PNEC_PHASE_TIMING
  if (rc == PNEC_OK && h->cfg.dump_timing) dump_phase_timing(a);
#endif
  return rc;
}

template <int V, int NW, int S, int MINB>
int launch_eval_t(pnec_handle *h, const EvalArgs &a, cudaStream_t stream) {
  auto kern = eval_kernel<V, NW, S, MINB>;
  const size_t dyn = static_cast<size_t>(S) * NW * 32 * VariantTraits<V>::kDoubles * 8;
  PNEC_CUDA(ensure_dyn_smem(h, kern, dyn));
  kern<<<static_cast<unsigned>(a.bv.num_problems), NW * 32, dyn, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  return PNEC_OK;
}

template <int V, int WPC, int S, int CHUNK, int MINB, int T = 32>
int launch_eval_warp_t(pnec_handle *h, const EvalArgs &a, cudaStream_t stream) {
  auto kern = eval_warp_kernel<V, WPC, S, CHUNK, MINB, T>;
  const size_t dyn = static_cast<size_t>(WPC) * S * T * VariantTraits<V>::kDoubles * 8;
  PNEC_CUDA(ensure_dyn_smem(h, kern, dyn));
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
  // warp-private rings of 3 stages x 32 correspondences, 4 warps per CTA, 3 CTAs per SM (measured best of
  // the round-1 sweep: 2 / 4 / 6 stages, 64- and 128-wide tiles, 8 warps); CTA-per-pair cooperative
  // loads for arrays that are not 16-byte aligned
  if (a.use_bulk) return launch_eval_warp_t<V, 4, 3, 8, 3>(h, a, stream);
  return launch_eval_t<V, 4, 4, 3>(h, a, stream);
}

int launch_eval(pnec_handle *h, const EvalArgs &a0, int variant, long long max_n,
                cudaStream_t stream) {
  EvalArgs a = a0;
  if (a.bv.num_problems == 0) return PNEC_OK;
  a.use_bulk = bulk_ok(a.bv) ? 1 : 0;
  if (h->cfg.no_bulk) a.use_bulk = 0;
  switch (variant) {
    case PNEC_VARIANT_NEC: return launch_eval_v<PNEC_VARIANT_NEC>(h, a, max_n, stream);
    case PNEC_VARIANT_TARGET: return launch_eval_v<PNEC_VARIANT_TARGET>(h, a, max_n, stream);
    case PNEC_VARIANT_HOST: return launch_eval_v<PNEC_VARIANT_HOST>(h, a, max_n, stream);
    default: return launch_eval_v<PNEC_VARIANT_SYMMETRIC>(h, a, max_n, stream);
  }
}


// Pairs [start, start + cnt) of a device view.  Ragged batches index their arrays absolutely
// through `offsets`, so only the offsets pointer moves; uniform batches move the array bases.
BatchView sub_view(const BatchView &bv, long long start, long long cnt) {
  BatchView v = bv;
  v.num_problems = cnt;
  if (bv.offsets) {
    v.offsets = bv.offsets + start;
  } else {
    const long long e0 = start * bv.n_uniform;
    v.f1 = bv.f1 + 3 * e0;
    v.f2 = bv.f2 + 3 * e0;
    if (bv.ct) v.ct = bv.ct + 9 * e0;
    if (bv.ch) v.ch = bv.ch + 9 * e0;
    v.total = cnt * bv.n_uniform;
  }
  if (bv.poses) v.poses = bv.poses + 7 * start;
  if (bv.counts) v.counts = bv.counts + start;
  return v;
}

// ------------------------------------------------- stage launchers (device views only)

// The refinement with its outputs: device arrays when `host` is false, else staged through the
// handle and copied to the caller's host arrays (the caller synchronises).
int run_solve(pnec_handle *h, const BatchView &bv, long long max_n, const pnec_solver_opts &o, bool host,
              double *poses, int32_t *status, int32_t *iterations, double *cost, double *initial_cost,
              cudaStream_t stream) {
  const long long B = bv.num_problems;
  SolveArgs a{};
  a.bv = bv;
  a.o = o;
  if (host) {
    PNEC_CUDA(h->d_out_poses.ensure(static_cast<size_t>(B) * 56));
    PNEC_CUDA(h->d_out_status.ensure(static_cast<size_t>(B) * 4));
    PNEC_CUDA(h->d_out_iters.ensure(static_cast<size_t>(B) * 4));
    PNEC_CUDA(h->d_out_cost.ensure(static_cast<size_t>(B) * 8));
    PNEC_CUDA(h->d_out_init.ensure(static_cast<size_t>(B) * 8));
    a.out_poses = static_cast<double *>(h->d_out_poses.p);
    a.out_status = status ? static_cast<int *>(h->d_out_status.p) : nullptr;
    a.out_iters = iterations ? static_cast<int *>(h->d_out_iters.p) : nullptr;
    a.out_cost = cost ? static_cast<double *>(h->d_out_cost.p) : nullptr;
    a.out_init_cost = initial_cost ? static_cast<double *>(h->d_out_init.p) : nullptr;
  } else {
    a.out_poses = poses;
    a.out_status = status;
    a.out_iters = iterations;
    a.out_cost = cost;
    a.out_init_cost = initial_cost;
  }
  int rc = launch_solve(h, a, o.variant, max_n, stream);
  if (rc != PNEC_OK) return rc;
  if (host) {
    PNEC_CUDA(cudaMemcpyAsync(poses, a.out_poses, static_cast<size_t>(B) * 56, cudaMemcpyDeviceToHost, stream));
    if (status)
      PNEC_CUDA(cudaMemcpyAsync(status, a.out_status, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToHost, stream));
    if (iterations)
      PNEC_CUDA(cudaMemcpyAsync(iterations, a.out_iters, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToHost, stream));
    if (cost)
      PNEC_CUDA(cudaMemcpyAsync(cost, a.out_cost, static_cast<size_t>(B) * 8, cudaMemcpyDeviceToHost, stream));
    if (initial_cost)
      PNEC_CUDA(cudaMemcpyAsync(initial_cost, a.out_init_cost, static_cast<size_t>(B) * 8, cudaMemcpyDeviceToHost, stream));
  }
  return PNEC_OK;
}

int ensure_sphere(pnec_handle *h, int fibonacci_samples) {
  // fibonacci_sphere(samples), scf.cc:53-72, with its float casts
  if (h->sphere_samples == fibonacci_samples) return PNEC_OK;
  std::vector<double> pts(static_cast<size_t>(std::max(fibonacci_samples, 1)) * 3);
  const double phi = M_PI * (3.0 - std::sqrt(5.0));
  for (int i = 0; i < fibonacci_samples; ++i) {
    const double y = 1.0 - ((float)i / (float)(fibonacci_samples - 1)) * 2.0;
    const double radius = std::sqrt(1 - y * y);
    const double theta = phi * (float)i;
    pts[3 * i] = std::cos(theta) * radius;
    pts[3 * i + 1] = y;
    pts[3 * i + 2] = std::sin(theta) * radius;
  }
  PNEC_CUDA(h->d_sphere.ensure(pts.size() * 8));
  PNEC_CUDA(cudaMemcpy(h->d_sphere.p, pts.data(), pts.size() * 8, cudaMemcpyHostToDevice));
  h->sphere_samples = fibonacci_samples;
  return PNEC_OK;
}

// dynamic shared memory of the SCF kernels: the largest pair, capped at what the device offers
size_t scf_smem_bytes(const pnec_handle *h, long long max_n) {
  const size_t want = static_cast<size_t>(std::max<long long>(max_n, 1)) * 72;
  const size_t cap = ((h->smem_optin - 8192) / 72) * 72;  // static shared memory of scf_list_kernel + slack
  return std::min(want, cap);
}

// Pairs above the shared-memory capacity keep their terms in HBM: allocate before launching.
int ensure_scf_spill(pnec_handle *h, const BatchView &bv, long long max_n) {
  if (scf_smem_bytes(h, max_n) >= static_cast<size_t>(std::max<long long>(max_n, 1)) * 72) return PNEC_OK;
  PNEC_CUDA(h->d_scf_spill.ensure(static_cast<size_t>(bv.total) * 72));
  return PNEC_OK;
}

// scf_kernel: rotation + start translation from bv.poses, result at out_t + out_stride * b (may be the
// translation slot of bv.poses itself: every read of the pose precedes the final write).  `bv` may be a
// sub-view; `spill_base` is the spill array of the WHOLE batch for ragged views (absolute indices) and
// of this sub-view's first correspondence for uniform ones.
int run_scf(pnec_handle *h, const BatchView &bv, long long max_n, double reg, int samples, int steps,
            double *out_t, int out_stride, double *out_cost, cudaStream_t stream,
            ScfScanCache *cache = nullptr, const int *q_same = nullptr, int *fixed = nullptr,
            int *defer_buf = nullptr /* 4 + B ints, else the handle's */, double *spill = nullptr,
            const double *prev_poses = nullptr) {
  const size_t dyn = scf_smem_bytes(h, max_n);
  const bool fits = dyn >= static_cast<size_t>(std::max<long long>(max_n, 1)) * 72;
  if (!fits && !spill) return fail(PNEC_ERR_INVALID_ARGUMENT, "internal: SCF spill array missing");
  int rc = ensure_sphere(h, samples);
  if (rc != PNEC_OK) return rc;
  ScfArgs a{};
  a.bv = bv;
  a.sphere = static_cast<const double *>(h->d_sphere.p);
  a.reg = reg;
  a.samples = samples;
  a.steps = steps;
  a.cap_elems = static_cast<int>(dyn / 72);
  a.spill = fits ? nullptr : spill;
  a.out_t = out_t;
  a.out_stride = out_stride;
  a.out_cost = out_cost;
  a.cache = cache;
  a.q_same = q_same;
  a.fixed = fixed;
  a.prev_poses = prev_poses;
  a.dbg = nullptr;
  static long long *dbg_buf = nullptr;
  const bool debug = h->cfg.scf_debug != 0;
  if (debug) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, sizeof(long long) * 4 * 1000000);
    if (bv.num_problems <= 1000000) a.dbg = dbg_buf;
  }
  // warps per pair in pass 1: 4 while several pairs fit an SM; large pairs leave room for one or
  // two CTAs per SM only, which then get more warps (the scan parallelises over them; the sums
  // that depend on the thread split stay on 4 warps, so the result does not change)
  int nw = h->cfg.scf_warps;
  if (nw == 0) nw = dyn > 110 * 1024 ? 16 : dyn > 72 * 1024 ? 8 : 4;
  if (nw != 1 && nw != 2 && nw != 8 && nw != 16) nw = 4;
  const int defer = h->cfg.scf_defer;  // survivors above which a pair goes to pass 2; 0 = one pass
  int *d_defer = defer_buf;
  if (defer > 0) {
    if (!d_defer) {
      PNEC_CUDA(h->d_scf_defer.ensure(sizeof(int) * (4 + static_cast<size_t>(bv.num_problems))));
      d_defer = static_cast<int *>(h->d_scf_defer.p);
    }
    PNEC_CUDA(cudaMemsetAsync(d_defer, 0, sizeof(int) * 4, stream));
    a.defer_count = d_defer;
    a.defer_list = d_defer + 4;
    a.defer_threshold = defer;
  }
  void (*kern)(ScfArgs) = nw == 1 ? scf_kernel<1> : nw == 2 ? scf_kernel<2> : nw == 8 ? scf_kernel<8>
                          : nw == 16 ? scf_kernel<16> : scf_kernel<4>;
  PNEC_CUDA(ensure_dyn_smem(h, kern, dyn));
  kern<<<static_cast<unsigned>(bv.num_problems), nw * 32, dyn, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  if (defer > 0) {
    ScfArgs a2 = a;
    a2.defer_list = nullptr;
    a2.defer_count = nullptr;
    a2.work_list = d_defer + 4;
    a2.work_count = d_defer;
    a2.work_cursor = d_defer + 1;
    auto kern2 = scf_list_kernel<16>;
    PNEC_CUDA(ensure_dyn_smem(h, kern2, dyn));
    const unsigned grid2 = static_cast<unsigned>(std::min<long long>(bv.num_problems, 2LL * h->sm_count));
    kern2<<<grid2, 16 * 32, dyn, stream>>>(a2);
    PNEC_CUDA(cudaGetLastError());
    h->launches++;
  }
  if (a.dbg) {  // debug only: synchronises
    cudaStreamSynchronize(stream);
    const long long B = bv.num_problems;
    std::vector<long long> hb(4 * B);
    cudaMemcpy(hb.data(), a.dbg, sizeof(long long) * 4 * B, cudaMemcpyDeviceToHost);
    long long cnt[3] = {0, 0, 0}, surv = 0, max_surv = 0, max_cyc = 0, over64 = 0;
    double cyc[3] = {0, 0, 0}, scan_cyc = 0;
    for (long long b = 0; b < B; ++b) {
      const int path = static_cast<int>(hb[4 * b]);
      cnt[path]++;
      cyc[path] += hb[4 * b + 3];
      if (path == 2) { surv += hb[4 * b + 1]; scan_cyc += hb[4 * b + 2]; if (hb[4 * b + 1] > 64) over64++; }
      max_surv = std::max(max_surv, hb[4 * b + 1]);
      max_cyc = std::max(max_cyc, hb[4 * b + 3]);
    }
    std::fprintf(stderr, "[scf] fixed %lld | reused %lld (mean %.0f cyc) | scanned %lld (mean %.0f cyc, scan part %.0f, mean survivors %.1f, max %lld, >64: %lld) | max CTA %lld cyc\n",
                 cnt[0], cnt[1], cnt[1] ? cyc[1] / cnt[1] : 0.0, cnt[2], cnt[2] ? cyc[2] / cnt[2] : 0.0,
                 cnt[2] ? scan_cyc / cnt[2] : 0.0, cnt[2] ? double(surv) / cnt[2] : 0.0, max_surv, over64, max_cyc);
  }
  return PNEC_OK;
}

int run_nec_translation(pnec_handle *h, const BatchView &bv, double *out_t, int out_stride, double *out_M,
                        cudaStream_t stream) {
  // large batches: a warp per pair (same bits, pnec_translation.cuh); few pairs: a CTA each for the latency
  if (bv.num_problems >= 8LL * h->sm_count)
    nec_translation_warp_kernel<<<static_cast<unsigned>((bv.num_problems + 3) / 4), 128, 0, stream>>>(bv, out_t, out_stride, out_M);
  else
    nec_translation_kernel<<<static_cast<unsigned>(bv.num_problems), 128, 0, stream>>>(bv, out_t, out_stride, out_M);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  return PNEC_OK;
}

// moments of the eigensolver; weighted: bv.poses = the poses the weights are computed from
int run_es_moments(pnec_handle *h, const BatchView &bv, bool weighted, double reg, double *d_mom,
                   cudaStream_t stream) {
  EsMomentArgs a{};
  a.bv = bv;
  a.reg = reg;
  a.out = d_mom;
  if (bulk_ok(bv) && !h->cfg.no_bulk) {
    // persistent warp-private rings (TMA), 4 warps x 3 stages, 3 CTAs per SM
    constexpr int WPC = 4, S = 3, MINB = 3;
    const long long want = (bv.num_problems + WPC - 1) / WPC;
    const unsigned grid = static_cast<unsigned>(std::max<long long>(1, std::min<long long>(want, 1LL * h->sm_count * MINB)));
    if (weighted) {
      auto kern = es_moments_warp_kernel<true, WPC, S, MINB>;
      const size_t dyn = static_cast<size_t>(WPC) * S * 32 * VariantTraits<PNEC_VARIANT_TARGET>::kDoubles * 8;
      PNEC_CUDA(ensure_dyn_smem(h, kern, dyn));
      kern<<<grid, WPC * 32, dyn, stream>>>(a);
    } else {
      auto kern = es_moments_warp_kernel<false, WPC, S, MINB>;
      const size_t dyn = static_cast<size_t>(WPC) * S * 32 * VariantTraits<PNEC_VARIANT_NEC>::kDoubles * 8;
      PNEC_CUDA(ensure_dyn_smem(h, kern, dyn));
      kern<<<grid, WPC * 32, dyn, stream>>>(a);
    }
  } else if (weighted) {
    es_moments_kernel<true><<<static_cast<unsigned>(bv.num_problems), 128, 0, stream>>>(a);
  } else {
    es_moments_kernel<false><<<static_cast<unsigned>(bv.num_problems), 128, 0, stream>>>(a);
  }
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  return PNEC_OK;
}

// opengv eigensolver_main with the parameters opengv sets (ftol 5e-5, xtol 10 eps, maxfev 100;
// resetParameters(): factor 100, gtol 0, epsfcn 0)
int run_es_lm(pnec_handle *h, long long B, const double *d_mom, const double *d_poses_in, double *d_poses_out,
              int *d_info, double *d_ev, cudaStream_t stream, const int *fixed = nullptr, int *q_same = nullptr,
              bool copy_translation = true) {
  EsLmArgs a{};
  a.moments = d_mom;
  a.poses_in = d_poses_in;
  a.poses_out = d_poses_out;
  a.out_info = d_info;
  a.out_nfev = nullptr;
  a.out_ev = d_ev;
  a.fixed = fixed;
  a.q_same = q_same;
  a.copy_translation = copy_translation ? 1 : 0;
  a.num_problems = B;
  a.ftol = 0.00005;
  a.xtol = 1.0e1 * DBL_EPSILON;
  a.gtol = 0.0;
  a.factor = 100.0;
  a.maxfev = 100;
  // Two pairs per warp (every turn a wide one, es_lm_group) while that fits the machine in one wave
  // (196 registers: two CTAs of 8 pairs per SM); eight per warp beyond.  Measured on the C2 shape:
  // 1000 pairs 0.257 -> 0.235 ms, 2500 pairs 0.267 -> 0.251 ms, 5000 pairs 0.293 -> 0.317 ms.
  const long long wide_max = h->cfg.es_wide_max_pairs >= 0 ? h->cfg.es_wide_max_pairs : 16LL * h->sm_count;
  if (B <= wide_max) {
    es_lm_kernel<2><<<static_cast<unsigned>((B + 7) / 8), kEsLmThreads, 0, stream>>>(a);
  } else {
    es_lm_kernel<8><<<static_cast<unsigned>((B + kEsLmPairs - 1) / kEsLmPairs), kEsLmThreads, 0, stream>>>(a);
  }
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  return PNEC_OK;
}

// opengv's Ransac<EigensolverSacProblem>::computeModel + selectWithinDistance + InlierExtraction
// (pnec.cc:239-251, 210-229).  `bv` may be a sub-view starting at pair `pair0` / correspondence `elem0`
// of the batch the scratch arrays were sized for.
struct RansacOut {
  double *best;     // [B][7]
  int *count;       // [B]
  int *iters;       // [B]
  int *index;       // [total] or nullptr
  double *f1, *f2;  // [total][3]
  double *ct;       // [total][9] or nullptr
};

// scratch of the two-pass scheme for a batch of B pairs cut into at most kMaxChunks chunks (each chunk
// owns the slice [c0, c0 + cnt) of every array and its own 4-int header)
int ensure_ransac_scratch(pnec_handle *h, long long B) {
  const size_t nb = static_cast<size_t>(B);
  PNEC_CUDA(h->d_rs_state.ensure(nb * sizeof(RansacPairState)));
  PNEC_CUDA(h->d_rs_defer.ensure(sizeof(int) * (4 * pnec_handle::kMaxChunks + nb)));
  PNEC_CUDA(h->d_rs_prefix.ensure(sizeof(int) * (pnec_handle::kMaxChunks + nb)));
  PNEC_CUDA(h->d_rs_hyp.ensure(sizeof(int) * nb * kRansacSuper));
  if (h->cfg.ransac_split) {
    PNEC_CUDA(h->d_rs_sp_mom.ensure(nb * 8 * kEsMom * sizeof(double)));
    PNEC_CUDA(h->d_rs_sp_x.ensure(nb * 8 * 3 * sizeof(double)));
    PNEC_CUDA(h->d_rs_sp_int.ensure(nb * (8 + 8 + 1) * sizeof(int)));  // i0 [B][8], active [B][8], live [B]
  }
  return PNEC_OK;
}

// `bv`: pairs [pair0, pair0 + bv.num_problems) of the batch the scratch was sized for, chunk `chunk`.
// `may_sync`: the caller is a small HOST-memspace call (it synchronises anyway): look at the number of pairs
// pass 1 deferred and skip the ~20 launches of pass 2 when there is none (the usual case for one clean pair).
int run_ransac(pnec_handle *h, const BatchView &bv, const pnec_frame_opts &o, long long pair_index_base,
               const RansacOut &out, long long pair0, int chunk, cudaStream_t stream, bool may_sync = false) {
  RansacArgs a{};
  a.bv = bv;
  a.best_poses = out.best;
  a.num_inliers = out.count;
  a.iterations = out.iters;
  a.inlier_index = out.index;
  a.out_f1 = out.f1;
  a.out_f2 = out.f2;
  a.out_ct = out.ct;
  a.max_iterations = o.max_ransac_iterations;
  a.sample_size = o.ransac_sample_size;
  a.threshold = o.ransac_threshold;
  a.probability = o.ransac_probability;
  a.max_variation = o.ransac_max_variation;
  a.seed = o.ransac_seed;
  a.pair_index_base = pair_index_base;
  a.lm = EsLmParams{0.00005, 1.0e1 * DBL_EPSILON, 0.0, 100.0, 100};
  // Pass 1, one CTA per pair.  Many pairs: one warp each (8 hypotheses per round; the machine is filled
  // by the pairs).  Few pairs: more warps, i.e. more hypotheses per round and a faster scoring pass.
  int nw = h->cfg.ransac_warps;
  if (nw == 0) nw = bv.num_problems >= 4LL * h->sm_count ? 1 : 4;
  // Pairs unfinished after a few rounds go to pass 2 (their hypotheses spread over the whole device).
  int defer_after = h->cfg.ransac_defer;
  if (defer_after < 0) defer_after = nw == 1 ? 24 : 56;
  if (o.max_ransac_iterations + 1 <= defer_after) defer_after = 0;
  a.defer_after = defer_after;
  a.state = static_cast<RansacPairState *>(h->d_rs_state.p) + pair0;
  a.defer = static_cast<int *>(h->d_rs_defer.p) + (4 * chunk + pair0);
  a.blk_prefix = static_cast<int *>(h->d_rs_prefix.p) + (chunk + pair0);
  a.hyp_count = static_cast<int *>(h->d_rs_hyp.p) + pair0 * kRansacSuper;
  if (defer_after > 0) PNEC_CUDA(cudaMemsetAsync(a.defer, 0, 4 * sizeof(int), stream));
  const unsigned grid = static_cast<unsigned>(bv.num_problems);
  // compiled for 16 warps per SM (128 registers): the hypotheses are latency bound (255 / 168 registers,
  // i.e. 8 / 12 warps per SM, measured 15 % / 3 % slower)
  if (nw == 1 && defer_after > 0 && h->cfg.ransac_split && h->d_rs_sp_mom.p) {
    // large batches: every round of pass 1 as three kernels over the pairs still in it (pnec_ransac.cuh:
    // the one-kernel form is bound by instruction fetch)
    const long long B0 = bv.num_problems;
    a.sp_mom = static_cast<double *>(h->d_rs_sp_mom.p) + pair0 * 8 * kEsMom;
    a.sp_x = static_cast<double *>(h->d_rs_sp_x.p) + pair0 * 8 * 3;
    int *ints = static_cast<int *>(h->d_rs_sp_int.p);
    const long long nb_all = static_cast<long long>(h->d_rs_sp_int.cap / (17 * sizeof(int)));  // layout by capacity
    a.sp_i0 = ints + pair0 * 8;
    a.sp_active = ints + nb_all * 8 + pair0 * 8;
    a.sp_live = ints + nb_all * 16 + pair0;
    const int rounds1 = (std::min(defer_after, o.max_ransac_iterations + 1) + 7) / 8;
    for (int r = 0; r < rounds1; ++r) {
      a.sp_round = r;
      ransac_pre_kernel<<<static_cast<unsigned>((B0 + 3) / 4), 128, 0, stream>>>(a);
      ransac_lm_kernel<<<static_cast<unsigned>((B0 * 8 + kEsLmPairs - 1) / kEsLmPairs), kEsLmThreads, 0, stream>>>(a);
      ransac_post_kernel<<<grid, 32, 0, stream>>>(a);
    }
    PNEC_CUDA(cudaGetLastError());
    h->launches += 3 * rounds1;
  } else {
    if (nw == 1) ransac_kernel<1, 16><<<grid, 32, 0, stream>>>(a);
    else if (nw == 2) ransac_kernel<2, 8><<<grid, 64, 0, stream>>>(a);
    else ransac_kernel<4, 2><<<grid, 128, 0, stream>>>(a);  // few pairs: all the registers, no spills in the LM
    PNEC_CUDA(cudaGetLastError());
    h->launches++;
  }
  if (defer_after > 0 && may_sync) {
    int deferred = -1;
    PNEC_CUDA(cudaMemcpyAsync(&deferred, a.defer, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
    if (deferred == 0) return PNEC_OK;
  }
  if (defer_after > 0) {
    // Pass 2: every super-round finishes a deferred pair or consumes its whole grant (ransac_grant), so
    // this many super-rounds cover the worst case; with nothing (left) to do the kernels return at once.
    int rounds = 0;
    for (int it = defer_after; it <= o.max_ransac_iterations; ++rounds) it += ransac_grant(it, o.max_ransac_iterations);
    const unsigned pgrid = static_cast<unsigned>(std::min<long long>(  // 4 CTAs of 4 independent warps per SM
        4LL * h->sm_count, (bv.num_problems * (kRansacSuper / kRansacBlock) + 3) / 4));
    const unsigned rgrid = static_cast<unsigned>((bv.num_problems + 3) / 4);  // a warp per pair that may have been deferred
    ransac_plan_kernel<<<1, 1024, 0, stream>>>(a);
    for (int r = 0; r < rounds; ++r) {
      ransac_hyp_kernel<4, 4><<<pgrid, 128, 0, stream>>>(a);
      ransac_replay_kernel<<<rgrid, 128, 0, stream>>>(a);
      ransac_plan_kernel<<<1, 1024, 0, stream>>>(a);
    }
    ransac_final_kernel<<<static_cast<unsigned>(std::min<long long>(2LL * h->sm_count, bv.num_problems)), 128, 0, stream>>>(a);
    PNEC_CUDA(cudaGetLastError());
    h->launches += 2 + 3 * rounds;
  }
  return PNEC_OK;
}

int validate_ransac_opts(const pnec_frame_opts *o) {
  if (o->ransac_sample_size < 1 || o->ransac_sample_size > kRansacMaxSample)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "ransac_sample_size must be in [1, 32]");
  if (o->max_ransac_iterations < 0) return fail(PNEC_ERR_INVALID_ARGUMENT, "max_ransac_iterations < 0");
  if (!(o->ransac_threshold > 0.0) || !(o->ransac_probability > 0.0) || !(o->ransac_probability < 1.0) ||
      !(o->ransac_max_variation >= 0.0))
    return fail(PNEC_ERR_INVALID_ARGUMENT, "bad RANSAC threshold / probability / max_variation");
  return PNEC_OK;
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
  h->smem_per_sm = prop.sharedMemPerMultiprocessor;
  h->cfg.load();
  h->stager.copy_threads = h->cfg.copy_threads;
  *out = h;
  return PNEC_OK;
}

void pnec_destroy(pnec_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  DevBuf *bufs[] = {&h->d_f1, &h->d_f2, &h->d_ct, &h->d_ch, &h->d_off, &h->d_poses,
                    &h->d_out_poses, &h->d_out_status, &h->d_out_iters, &h->d_out_cost,
                    &h->d_out_init, &h->d_out_grad, &h->d_out_jtj, &h->d_ut_mu, &h->d_ut_cov,
                    &h->d_ut_out, &h->d_kp_bv, &h->d_sphere, &h->d_tr_out,
                    &h->d_tr_aux, &h->d_es_mom, &h->d_es_w, &h->d_es_info, &h->d_es_ev,
                    &h->d_fr_es, &h->d_fr_a, &h->d_fr_cache, &h->d_fr_flags, &h->d_scf_defer, &h->d_fr_defer, &h->d_scf_spill,
                    &h->d_rs_f1, &h->d_rs_f2, &h->d_rs_ct, &h->d_rs_best, &h->d_rs_cnt, &h->d_rs_iters, &h->d_rs_idx,
                    &h->d_rs_state, &h->d_rs_defer, &h->d_rs_prefix, &h->d_rs_hyp,
                    &h->d_rs_sp_mom, &h->d_rs_sp_x, &h->d_rs_sp_int,
                    &h->d_kp_hp, &h->d_kp_tp, &h->d_kp_hc, &h->d_kp_tc, &h->d_kp_hi, &h->d_kp_ti};
  for (DevBuf *b : bufs) b->release();
  for (DevBuf &b : h->d_kp_out) b.release();
  h->d_pk_ct.release();
  h->d_pk_ch.release();
  h->d_slot_ctr.release();
  for (DevBuf &b : h->d_slot_start) b.release();
  for (int i = 0; i < pnec_handle::kMaxChunks; ++i) {
    if (h->side[i]) cudaStreamDestroy(h->side[i]);
    if (h->lm_side[i]) cudaStreamDestroy(h->lm_side[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    if (h->ev_es[i]) cudaEventDestroy(h->ev_es[i]);
    for (int r = 0; r < pnec_handle::kMaxRounds; ++r)
      if (h->ev_round[i][r]) cudaEventDestroy(h->ev_round[i][r]);
  }
  h->d_fr_rounds.release();
  for (cudaEvent_t ev : h->ev_stage)
    if (ev) cudaEventDestroy(ev);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_last) cudaEventDestroy(h->ev_last);
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
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
  Staged st;
  const bool host = batch->memspace == PNEC_MEM_HOST;
  // HOST batches of some size are cut into chunks, each on its own stream: H2D of chunk k + 1 runs
  // under the solve of chunk k and the D2H of its results (the copies are the bulk of such a call).
  int chunks = 1;
  if (host) {
    const long long total = batch->offsets ? batch->offsets[B] : B * batch->n_per_problem;
    chunks = h->cfg.h2d_chunks;
    if (chunks <= 0) chunks = (total * bytes_per_corr(opts->variant) >= (64ll << 20) && B >= 64) ? 4 : 1;
    chunks = static_cast<int>(std::min<long long>(std::min(chunks, pnec_handle::kMaxChunks), B));
  }
  if (chunks <= 1) {
    rc = stage_batch(h, batch, opts->variant, stream, &st);
    if (rc != PNEC_OK) return rc;
    rc = run_solve(h, st.bv, st.max_n, *opts, host, out->poses, out->status, out->iterations, out->cost,
                   out->initial_cost, stream);
    if (rc != PNEC_OK) return rc;
    if (host) PNEC_CUDA(cudaStreamSynchronize(stream));
    return PNEC_OK;
  }
  rc = stage_alloc(h, batch, opts->variant, stream, &st);
  if (rc != PNEC_OK) return rc;
  const size_t nb = static_cast<size_t>(B);
  PNEC_CUDA(h->d_out_poses.ensure(nb * 56));
  PNEC_CUDA(h->d_out_status.ensure(nb * 4));
  PNEC_CUDA(h->d_out_iters.ensure(nb * 4));
  PNEC_CUDA(h->d_out_cost.ensure(nb * 8));
  PNEC_CUDA(h->d_out_init.ensure(nb * 8));
  double *d_poses = static_cast<double *>(h->d_out_poses.p);
  int32_t *d_status = out->status ? static_cast<int32_t *>(h->d_out_status.p) : nullptr;
  int32_t *d_iters = out->iterations ? static_cast<int32_t *>(h->d_out_iters.p) : nullptr;
  double *d_cost = out->cost ? static_cast<double *>(h->d_out_cost.p) : nullptr;
  double *d_init = out->initial_cost ? static_cast<double *>(h->d_out_init.p) : nullptr;
  if (!h->ev_fork) PNEC_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  PNEC_CUDA(cudaEventRecord(h->ev_fork, stream));  // after the copy of the offsets
  scope.forked = true;
  for (int c = 0; c < chunks; ++c) {
    if (!h->side[c]) PNEC_CUDA(cudaStreamCreateWithFlags(&h->side[c], cudaStreamNonBlocking));
    if (!h->ev_join[c]) PNEC_CUDA(cudaEventCreateWithFlags(&h->ev_join[c], cudaEventDisableTiming));
    cudaStream_t cs = h->side[c];
    PNEC_CUDA(cudaStreamWaitEvent(cs, h->ev_fork, 0));
    // chunk boundaries balance correspondences for ragged batches
    long long p0, p1;
    if (batch->offsets) {
      const long long total = batch->offsets[B];
      auto first_at = [&](long long target) {
        return std::lower_bound(batch->offsets, batch->offsets + B, target) - batch->offsets;
      };
      p0 = c == 0 ? 0 : first_at(total * c / chunks);
      p1 = c + 1 == chunks ? B : first_at(total * (c + 1) / chunks);
    } else {
      p0 = B * c / chunks;
      p1 = B * (c + 1) / chunks;
    }
    if (p1 > p0) {
      rc = stage_copy(h, batch, opts->variant, p0, p1, cs);
      if (rc != PNEC_OK) return rc;
      const BatchView bv = sub_view(st.bv, p0, p1 - p0);
      rc = run_solve(h, bv, st.max_n, *opts, false, d_poses + 7 * p0, d_status ? d_status + p0 : nullptr,
                     d_iters ? d_iters + p0 : nullptr, d_cost ? d_cost + p0 : nullptr,
                     d_init ? d_init + p0 : nullptr, cs);
      if (rc != PNEC_OK) return rc;
      const size_t cnt = static_cast<size_t>(p1 - p0);
      PNEC_CUDA(cudaMemcpyAsync(out->poses + 7 * p0, d_poses + 7 * p0, cnt * 56, cudaMemcpyDeviceToHost, cs));
      if (d_status) PNEC_CUDA(cudaMemcpyAsync(out->status + p0, d_status + p0, cnt * 4, cudaMemcpyDeviceToHost, cs));
      if (d_iters) PNEC_CUDA(cudaMemcpyAsync(out->iterations + p0, d_iters + p0, cnt * 4, cudaMemcpyDeviceToHost, cs));
      if (d_cost) PNEC_CUDA(cudaMemcpyAsync(out->cost + p0, d_cost + p0, cnt * 8, cudaMemcpyDeviceToHost, cs));
      if (d_init) PNEC_CUDA(cudaMemcpyAsync(out->initial_cost + p0, d_init + p0, cnt * 8, cudaMemcpyDeviceToHost, cs));
    }
    PNEC_CUDA(cudaEventRecord(h->ev_join[c], cs));
  }
  for (int c = 0; c < chunks; ++c) PNEC_CUDA(cudaStreamWaitEvent(stream, h->ev_join[c], 0));
  PNEC_CUDA(cudaStreamSynchronize(stream));
  scope.ok = true;
  return PNEC_OK;
}

int pnec_eval_batch(pnec_handle *h, const pnec_batch *batch, int32_t variant,
                    double regularization, const pnec_eval_out *out, void *cuda_stream) {
  if (!h || !out) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  int rc = validate_batch(batch, variant, true);
  if (rc != PNEC_OK) return rc;
  const long long B = batch->num_problems;
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
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
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
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
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
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
  a.use_bulk = (aligned16(a.mus) && aligned16(a.covs) && aligned16(a.out) && !h->cfg.no_bulk) ? 1 : 0;
  unscented_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  if (memspace == PNEC_MEM_HOST) {
    PNEC_CUDA(cudaMemcpyAsync(out_covs, a.out, nn * 72, cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
  }
  return PNEC_OK;
}

int pnec_keypoints_unproject_batch(pnec_handle *h, int64_t n, int32_t memspace, const double *points,
                                   const double *covs2, const double *K_inv, double *out_bvs,
                                   double *out_covs, void *cuda_stream) {
  if (!h) return fail(PNEC_ERR_INVALID_ARGUMENT, "handle is NULL");
  if (n < 0) return fail(PNEC_ERR_INVALID_ARGUMENT, "n < 0");
  if (n == 0) return PNEC_OK;
  if (!points || !covs2 || !K_inv || !out_bvs || !out_covs)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL array");
  if (memspace != PNEC_MEM_HOST && memspace != PNEC_MEM_DEVICE)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown memspace");
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
  KpArgs a{};
  a.n = n;
  for (int k = 0; k < 9; ++k) a.Kinv[k] = K_inv[k];
  const size_t nn = static_cast<size_t>(n);
  if (memspace == PNEC_MEM_HOST) {
    PNEC_CUDA(h->d_ut_mu.ensure(nn * 16));
    PNEC_CUDA(h->d_ut_cov.ensure(nn * 32));
    PNEC_CUDA(h->d_kp_bv.ensure(nn * 24));
    PNEC_CUDA(h->d_ut_out.ensure(nn * 72));
    PNEC_CUDA(cudaMemcpyAsync(h->d_ut_mu.p, points, nn * 16, cudaMemcpyHostToDevice, stream));
    PNEC_CUDA(cudaMemcpyAsync(h->d_ut_cov.p, covs2, nn * 32, cudaMemcpyHostToDevice, stream));
    a.points = static_cast<const double *>(h->d_ut_mu.p);
    a.covs2 = static_cast<const double *>(h->d_ut_cov.p);
    a.out_bvs = static_cast<double *>(h->d_kp_bv.p);
    a.out_covs = static_cast<double *>(h->d_ut_out.p);
  } else {
    a.points = points;
    a.covs2 = covs2;
    a.out_bvs = out_bvs;
    a.out_covs = out_covs;
  }
  a.use_bulk = (aligned16(a.points) && aligned16(a.covs2) && aligned16(a.out_bvs) &&
                aligned16(a.out_covs) && !h->cfg.no_bulk) ? 1 : 0;
  keypoint_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(a);
  PNEC_CUDA(cudaGetLastError());
  h->launches++;
  if (memspace == PNEC_MEM_HOST) {
    PNEC_CUDA(cudaMemcpyAsync(out_bvs, a.out_bvs, nn * 24, cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaMemcpyAsync(out_covs, a.out_covs, nn * 72, cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
  }
  return PNEC_OK;
}

int pnec_scf_translation_batch(pnec_handle *h, const pnec_batch *batch, double regularization,
                               int32_t fibonacci_samples, int32_t scf_steps, double *out_translations,
                               double *out_cost, void *cuda_stream) {
  if (!h || !out_translations) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  if (fibonacci_samples < 0 || scf_steps < 0)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "negative sample / step count");
  int rc = validate_batch(batch, PNEC_VARIANT_TARGET, true);
  if (rc != PNEC_OK) return rc;
  const long long B = batch->num_problems;
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
  Staged st;
  rc = stage_batch(h, batch, PNEC_VARIANT_TARGET, stream, &st);
  if (rc != PNEC_OK) return rc;
  const bool host = batch->memspace == PNEC_MEM_HOST;
  double *d_t = out_translations, *d_c = out_cost;
  if (host) {
    PNEC_CUDA(h->d_tr_out.ensure(static_cast<size_t>(B) * 24));
    PNEC_CUDA(h->d_tr_aux.ensure(static_cast<size_t>(B) * 8));
    d_t = static_cast<double *>(h->d_tr_out.p);
    d_c = out_cost ? static_cast<double *>(h->d_tr_aux.p) : nullptr;
  }
  rc = ensure_scf_spill(h, st.bv, st.max_n);
  if (rc != PNEC_OK) return rc;
  rc = run_scf(h, st.bv, st.max_n, regularization, fibonacci_samples, scf_steps, d_t, 3, d_c, stream, nullptr,
               nullptr, nullptr, nullptr, static_cast<double *>(h->d_scf_spill.p));
  if (rc != PNEC_OK) return rc;
  if (host) {
    PNEC_CUDA(cudaMemcpyAsync(out_translations, d_t, static_cast<size_t>(B) * 24, cudaMemcpyDeviceToHost, stream));
    if (out_cost)
      PNEC_CUDA(cudaMemcpyAsync(out_cost, d_c, static_cast<size_t>(B) * 8, cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
  }
  return PNEC_OK;
}

int pnec_nec_translation_batch(pnec_handle *h, const pnec_batch *batch, double *out_translations,
                               double *out_M, void *cuda_stream) {
  if (!h || !out_translations) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  int rc = validate_batch(batch, PNEC_VARIANT_NEC, true);
  if (rc != PNEC_OK) return rc;
  const long long B = batch->num_problems;
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
  Staged st;
  rc = stage_batch(h, batch, PNEC_VARIANT_NEC, stream, &st);
  if (rc != PNEC_OK) return rc;
  const bool host = batch->memspace == PNEC_MEM_HOST;
  double *d_t = out_translations, *d_M = out_M;
  if (host) {
    PNEC_CUDA(h->d_tr_out.ensure(static_cast<size_t>(B) * 24));
    PNEC_CUDA(h->d_tr_aux.ensure(static_cast<size_t>(B) * 48));
    d_t = static_cast<double *>(h->d_tr_out.p);
    d_M = out_M ? static_cast<double *>(h->d_tr_aux.p) : nullptr;
  }
  rc = run_nec_translation(h, st.bv, d_t, 3, d_M, stream);
  if (rc != PNEC_OK) return rc;
  if (host) {
    PNEC_CUDA(cudaMemcpyAsync(out_translations, d_t, static_cast<size_t>(B) * 24, cudaMemcpyDeviceToHost, stream));
    if (out_M)
      PNEC_CUDA(cudaMemcpyAsync(out_M, d_M, static_cast<size_t>(B) * 48, cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
  }
  return PNEC_OK;
}

int pnec_eigensolver_batch(pnec_handle *h, const pnec_batch *batch, const double *weight_poses,
                           double regularization, double *out_poses, int32_t *out_lm_info,
                           double *out_smallest_ev, void *cuda_stream) {
  if (!h || !out_poses) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  const bool weighted = weight_poses != nullptr;
  const int variant = weighted ? PNEC_VARIANT_TARGET : PNEC_VARIANT_NEC;
  int rc = validate_batch(batch, variant, true);
  if (rc != PNEC_OK) return rc;
  const long long B = batch->num_problems;
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
  Staged st;
  rc = stage_batch(h, batch, variant, stream, &st);
  if (rc != PNEC_OK) return rc;
  const bool host = batch->memspace == PNEC_MEM_HOST;
  const size_t nb = static_cast<size_t>(B);
  PNEC_CUDA(h->d_es_mom.ensure(nb * kEsMom * 8));
  double *d_mom = static_cast<double *>(h->d_es_mom.p);
  double *d_out = out_poses, *d_ev = out_smallest_ev;
  int *d_info = out_lm_info;
  const double *d_wp = weight_poses;
  if (host) {
    PNEC_CUDA(h->d_fr_a.ensure(nb * 56));
    PNEC_CUDA(h->d_es_info.ensure(nb * 4));
    PNEC_CUDA(h->d_es_ev.ensure(nb * 8));
    d_out = static_cast<double *>(h->d_fr_a.p);
    d_info = out_lm_info ? static_cast<int *>(h->d_es_info.p) : nullptr;
    d_ev = out_smallest_ev ? static_cast<double *>(h->d_es_ev.p) : nullptr;
    if (weighted) {
      PNEC_CUDA(h->d_es_w.ensure(nb * 56));
      PNEC_CUDA(cudaMemcpyAsync(h->d_es_w.p, weight_poses, nb * 56, cudaMemcpyHostToDevice, stream));
      d_wp = static_cast<const double *>(h->d_es_w.p);
    }
  }
  BatchView mv = st.bv;
  if (weighted) mv.poses = d_wp;
  rc = run_es_moments(h, mv, weighted, regularization, d_mom, stream);
  if (rc != PNEC_OK) return rc;
  rc = run_es_lm(h, B, d_mom, st.bv.poses, d_out, d_info, d_ev, stream);
  if (rc != PNEC_OK) return rc;
  if (host) {
    PNEC_CUDA(cudaMemcpyAsync(out_poses, d_out, nb * 56, cudaMemcpyDeviceToHost, stream));
    if (out_lm_info)
      PNEC_CUDA(cudaMemcpyAsync(out_lm_info, d_info, nb * 4, cudaMemcpyDeviceToHost, stream));
    if (out_smallest_ev)
      PNEC_CUDA(cudaMemcpyAsync(out_smallest_ev, d_ev, nb * 8, cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
  }
  return PNEC_OK;
}

void pnec_frame_opts_default(pnec_frame_opts *o) {
  if (!o) return;
  o->use_nec = 0;
  o->use_ceres = 1;
  o->weighted_iterations = 10;
  o->use_ransac = 1;
  o->fibonacci_samples = 500;
  o->scf_steps = 10;
  pnec_solver_opts_default(&o->ceres);
  o->max_ransac_iterations = 5000;
  o->ransac_sample_size = 10;
  o->ransac_threshold = 1.0e-6;
  o->ransac_probability = 0.99;
  o->ransac_max_variation = 0.1;
  o->ransac_seed = 1;
  o->ransac_pair_index_base = 0;
}

int pnec_ransac_batch(pnec_handle *h, const pnec_batch *batch, const pnec_frame_opts *opts, int64_t pair_index_base,
                      double *out_models, int32_t *out_num_inliers, int32_t *out_iterations,
                      int32_t *out_inlier_index, void *cuda_stream) {
  if (!h || !opts || !out_models || !out_num_inliers) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  int rc = validate_ransac_opts(opts);
  if (rc != PNEC_OK) return rc;
  rc = validate_batch(batch, PNEC_VARIANT_NEC, true);
  if (rc != PNEC_OK) return rc;
  const long long B = batch->num_problems;
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
  Staged st;
  rc = stage_batch(h, batch, PNEC_VARIANT_NEC, stream, &st);
  if (rc != PNEC_OK) return rc;
  const bool host = batch->memspace == PNEC_MEM_HOST;
  const size_t nb = static_cast<size_t>(B), nel = static_cast<size_t>(std::max<long long>(st.bv.total, 1));
  PNEC_CUDA(h->d_rs_f1.ensure(nel * 24));
  PNEC_CUDA(h->d_rs_f2.ensure(nel * 24));
  PNEC_CUDA(h->d_rs_iters.ensure(nb * 4));
  RansacOut ro{};
  ro.best = out_models;
  ro.count = out_num_inliers;
  ro.iters = out_iterations ? out_iterations : static_cast<int *>(h->d_rs_iters.p);
  ro.index = out_inlier_index;
  ro.f1 = static_cast<double *>(h->d_rs_f1.p);
  ro.f2 = static_cast<double *>(h->d_rs_f2.p);
  if (host) {
    PNEC_CUDA(h->d_rs_best.ensure(nb * 56));
    PNEC_CUDA(h->d_rs_cnt.ensure(nb * 4));
    if (out_inlier_index) PNEC_CUDA(h->d_rs_idx.ensure(nel * 4));
    ro.best = static_cast<double *>(h->d_rs_best.p);
    ro.count = static_cast<int *>(h->d_rs_cnt.p);
    ro.iters = static_cast<int *>(h->d_rs_iters.p);
    ro.index = out_inlier_index ? static_cast<int *>(h->d_rs_idx.p) : nullptr;
  }
  rc = ensure_ransac_scratch(h, B);
  if (rc != PNEC_OK) return rc;
  rc = run_ransac(h, st.bv, *opts, pair_index_base, ro, 0, 0, stream, host && B <= 64);
  if (rc != PNEC_OK) return rc;
  if (host) {
    PNEC_CUDA(cudaMemcpyAsync(out_models, ro.best, nb * 56, cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaMemcpyAsync(out_num_inliers, ro.count, nb * 4, cudaMemcpyDeviceToHost, stream));
    if (out_iterations) PNEC_CUDA(cudaMemcpyAsync(out_iterations, ro.iters, nb * 4, cudaMemcpyDeviceToHost, stream));
    if (out_inlier_index)
      PNEC_CUDA(cudaMemcpyAsync(out_inlier_index, ro.index, static_cast<size_t>(st.bv.total) * 4, cudaMemcpyDeviceToHost, stream));
    PNEC_CUDA(cudaStreamSynchronize(stream));
  }
  return PNEC_OK;
}

int pnec_frame_solve_batch(pnec_handle *h, const pnec_batch *batch, const pnec_frame_opts *opts,
                           const pnec_frame_out *out, void *cuda_stream) {
  if (!h || !opts || !out) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  const bool ransac = opts->use_ransac != 0;
  if (ransac) {
    const int rrc = validate_ransac_opts(opts);
    if (rrc != PNEC_OK) return rrc;
  }
  if (opts->weighted_iterations < 0 || opts->fibonacci_samples < 0 || opts->scf_steps < 0)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "negative iteration / sample / step count");
  const bool nec = opts->use_nec != 0;
  const bool weighted = !nec && opts->weighted_iterations > 1;
  const int variant = nec ? PNEC_VARIANT_NEC : PNEC_VARIANT_TARGET;
  int rc = validate_batch(batch, variant, true);
  if (rc != PNEC_OK) return rc;
  const long long B = batch->num_problems;
  if (B > 0 && !out->poses) return fail(PNEC_ERR_INVALID_ARGUMENT, "out->poses is NULL");
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
  Staged st;
  rc = stage_batch(h, batch, variant, stream, &st);
  if (rc != PNEC_OK) return rc;
  const bool host = batch->memspace == PNEC_MEM_HOST;
  const size_t nb = static_cast<size_t>(B);
  // ---- scratch for the whole batch (allocated before anything is forked: cudaMalloc synchronises)
  PNEC_CUDA(h->d_es_mom.ensure(nb * kEsMom * 8));
  PNEC_CUDA(h->d_fr_es.ensure(nb * 56));
  PNEC_CUDA(h->d_fr_a.ensure(nb * 56));
  PNEC_CUDA(h->d_fr_cache.ensure(nb * sizeof(ScfScanCache)));
  PNEC_CUDA(h->d_fr_flags.ensure(nb * 2 * sizeof(int)));
  PNEC_CUDA(h->d_fr_defer.ensure(sizeof(int) * (4 * pnec_handle::kMaxChunks + nb)));
  const int weighted_rounds = weighted ? opts->weighted_iterations - 1 : 0;
  if (weighted) PNEC_CUDA(h->d_fr_rounds.ensure(nb * 56 * static_cast<size_t>(weighted_rounds)));
  double *const d_rounds = static_cast<double *>(h->d_fr_rounds.p);
  const bool lm_ahead = weighted_rounds <= pnec_handle::kMaxRounds && !h->cfg.no_lm_ahead;
  // small batches: one launch for all weighted rounds of a pair (the per-round launches would dominate)
  const bool fused_rounds = weighted && B <= h->cfg.fused_rounds_max_pairs;
  if (weighted) {
    rc = ensure_sphere(h, opts->fibonacci_samples);
    if (rc != PNEC_OK) return rc;
    rc = ensure_scf_spill(h, st.bv, st.max_n);
    if (rc != PNEC_OK) return rc;
  }
  double *const d_spill = static_cast<double *>(h->d_scf_spill.p);
  double *o_poses = out->poses, *o_cost = out->cost;
  int32_t *o_status = out->status, *o_iters = out->iterations;
  // RANSAC scratch: the inliers of every pair, compacted to the front of the pair's own slot
  RansacOut ro{};
  if (ransac) {
    const size_t nel = static_cast<size_t>(std::max<long long>(st.bv.total, 1));
    PNEC_CUDA(h->d_rs_f1.ensure(nel * 24));
    PNEC_CUDA(h->d_rs_f2.ensure(nel * 24));
    if (!nec) PNEC_CUDA(h->d_rs_ct.ensure(nel * 72));
    PNEC_CUDA(h->d_rs_best.ensure(nb * 56));
    PNEC_CUDA(h->d_rs_cnt.ensure(nb * 4));
    PNEC_CUDA(h->d_rs_iters.ensure(nb * 4));
    if (host && out->inlier_index) PNEC_CUDA(h->d_rs_idx.ensure(nel * 4));
    {
      const int src = ensure_ransac_scratch(h, B);
      if (src != PNEC_OK) return src;
    }
    ro.best = static_cast<double *>(h->d_rs_best.p);
    ro.count = (!host && out->num_inliers) ? out->num_inliers : static_cast<int *>(h->d_rs_cnt.p);
    ro.iters = (!host && out->ransac_iterations) ? out->ransac_iterations : static_cast<int *>(h->d_rs_iters.p);
    ro.index = out->inlier_index ? (host ? static_cast<int *>(h->d_rs_idx.p) : out->inlier_index) : nullptr;
    ro.f1 = static_cast<double *>(h->d_rs_f1.p);
    ro.f2 = static_cast<double *>(h->d_rs_f2.p);
    ro.ct = nec ? nullptr : static_cast<double *>(h->d_rs_ct.p);
  } else if (!host && out->num_inliers) {
    PNEC_CUDA(cudaMemsetAsync(out->num_inliers, 0, nb * 4, stream));  // inliers.clear(), pnec.cc:277
  }
  if (host) {
    PNEC_CUDA(h->d_out_poses.ensure(nb * 56));
    PNEC_CUDA(h->d_out_status.ensure(nb * 4));
    PNEC_CUDA(h->d_out_iters.ensure(nb * 4));
    PNEC_CUDA(h->d_out_cost.ensure(nb * 8));
    o_poses = static_cast<double *>(h->d_out_poses.p);
    o_status = out->status ? static_cast<int32_t *>(h->d_out_status.p) : nullptr;
    o_iters = out->iterations ? static_cast<int32_t *>(h->d_out_iters.p) : nullptr;
    o_cost = out->cost ? static_cast<double *>(h->d_out_cost.p) : nullptr;
  }
  double *const d_mom = static_cast<double *>(h->d_es_mom.p);
  double *const d_es = (!host && out->es_poses) ? out->es_poses : static_cast<double *>(h->d_fr_es.p);
  double *const d_a = static_cast<double *>(h->d_fr_a.p);
  ScfScanCache *const d_cache = static_cast<ScfScanCache *>(h->d_fr_cache.p);
  int *const d_qsame = static_cast<int *>(h->d_fr_flags.p), *const d_fixed = d_qsame + B;
  int *const d_defer = static_cast<int *>(h->d_fr_defer.p);
  const bool shortcuts = !h->cfg.no_frame_shortcuts;
  const bool timed = out->stage_ms != nullptr;
  scope.forked = true;  // chunks and the rotation chain run on the handle's side streams
  if (timed)
    for (cudaEvent_t &ev : h->ev_stage)
      if (!ev) PNEC_CUDA(cudaEventCreate(&ev));

  // The stages of one chunk of frame pairs, back to back on one stream.
  auto run_chunk = [&](long long c0, long long cnt, int chunk, cudaStream_t cs) -> int {
    BatchView bv = sub_view(st.bv, c0, cnt);
    double *mom = d_mom + kEsMom * c0, *es = d_es + 7 * c0, *pa = d_a + 7 * c0;
    int rcc;
    auto mark = [&](int k) -> cudaError_t { return timed ? cudaEventRecord(h->ev_stage[k], cs) : cudaSuccess; };
    PNEC_CUDA(mark(0));
    const double *es_start = bv.poses;
    const double *caller_poses = bv.poses;
    if (ransac) {
      // 0. RANSAC + InlierExtraction: from here on the pair IS its inliers (compacted arrays at the
      // same offsets, per-pair counts); the eigensolver below then is optimizeModelCoefficients,
      // started at the winning hypothesis (pnec.cc:253-256)
      const long long e0 = st.bv.offsets ? 0 : c0 * st.bv.n_uniform;  // ragged views index absolutely
      RansacOut r = ro;
      r.best += 7 * c0; r.count += c0; r.iters += c0;
      r.f1 += 3 * e0; r.f2 += 3 * e0;
      if (r.ct) r.ct += 9 * e0;
      if (r.index) r.index += e0;
      if ((rcc = run_ransac(h, bv, *opts, opts->ransac_pair_index_base + c0, r, c0, chunk, cs, host && B <= 64)) != PNEC_OK)
        return rcc;
      bv.f1 = r.f1; bv.f2 = r.f2;
      if (r.ct) bv.ct = r.ct;
      bv.counts = r.count;
      es_start = r.best;
    }
    // 1. PNEC::Eigensolver: rotation, then TranslationFromM(ComposeM(bvs1, bvs2, rotation))
    if ((rcc = run_es_moments(h, bv, false, 0.0, mom, cs)) != PNEC_OK) return rcc;
    if ((rcc = run_es_lm(h, cnt, mom, es_start, es, nullptr, nullptr, cs)) != PNEC_OK) return rcc;
    BatchView ev = bv;
    ev.poses = es;
    if ((rcc = run_nec_translation(h, ev, es + 4, 7, nullptr, cs)) != PNEC_OK) return rcc;
    PNEC_CUDA(mark(1));  // FrameTiming::nec_es_: Eigensolver + InlierExtraction (pnec.cc:149-161)
    // 2./3. the start pose of the refinement
    const double *init = es;
    if (weighted) {
      // weights from ES_solution in every round (pnec.cc:296-300): one set of weighted moments
      if ((rcc = run_es_moments(h, ev, true, opts->ceres.regularization, mom, cs)) != PNEC_OK) return rcc;
      // The round (R, t) -> (R', t') is a deterministic map with everything else held constant, so
      //  * a rotation that repeats bit for bit reuses its sphere scan (ScfScanCache), and
      //  * a pair whose whole pose repeats has reached a fixed point: later rounds are skipped.
      // Both give exactly what recomputing would.
      //
      // The rotation of round k depends on the rotation of round k - 1 only (the moments are
      // constant), not on any translation, so the chain of rotation solves runs ahead on its own
      // stream and the SCF rounds follow it: round k's SCF overlaps round k + 1's rotation solve.
      ScfScanCache *cache = d_cache + c0;
      int *rot_same = d_qsame + c0, *fixed = d_fixed + c0;
      int *defer = d_defer + (4 * chunk + c0);  // chunk k: 4 header ints, then its list (disjoint regions)
      PNEC_CUDA(cudaMemsetAsync(cache, 0, static_cast<size_t>(cnt) * sizeof(ScfScanCache), cs));
      PNEC_CUDA(cudaMemsetAsync(rot_same, 0, static_cast<size_t>(cnt) * sizeof(int), cs));
      PNEC_CUDA(cudaMemsetAsync(fixed, 0, static_cast<size_t>(cnt) * sizeof(int), cs));
      const int rounds = opts->weighted_iterations - 1;
      if (fused_rounds) {
        // small batch: all rounds of a pair in one CTA and one launch (pnec_frame.cuh)
        const size_t dyn = scf_smem_bytes(h, st.max_n);
        const bool fits = dyn >= static_cast<size_t>(std::max<long long>(st.max_n, 1)) * 72;
        FrameRoundsArgs fa{};
        fa.scf.bv = bv;
        fa.scf.sphere = static_cast<const double *>(h->d_sphere.p);
        fa.scf.reg = opts->ceres.regularization;
        fa.scf.samples = opts->fibonacci_samples;
        fa.scf.steps = opts->scf_steps;
        fa.scf.cap_elems = static_cast<int>(dyn / 72);
        fa.scf.spill = fits ? nullptr : (st.bv.offsets ? d_spill : d_spill + 9 * c0 * st.bv.n_uniform);
        fa.scf.cache = shortcuts ? cache : nullptr;
        fa.scf.fixed = shortcuts ? fixed : nullptr;
        fa.lm = EsLmParams{0.00005, 1.0e1 * DBL_EPSILON, 0.0, 100.0, 100};
        fa.moments = mom;
        fa.es_poses = es;
        fa.rounds = d_rounds + 7 * c0;  // round k of pair b at rounds + 7 ((k - 1) B' + b), B' = this chunk's pairs
        fa.final_poses = pa;
        fa.num_problems = cnt;
        fa.num_rounds = rounds;
        auto kern = frame_rounds_kernel<4>;
        PNEC_CUDA(ensure_dyn_smem(h, kern, dyn));
        kern<<<static_cast<unsigned>(cnt), 128, dyn, cs>>>(fa);
        PNEC_CUDA(cudaGetLastError());
        h->launches++;
        init = pa;
      } else {
      auto round_poses = [&](int k) {  // k = 0: the eigensolver pose; k >= 1: round k
        return k == 0 ? es : d_rounds + 7 * (static_cast<long long>(k - 1) * B + c0);
      };
      cudaStream_t ls = cs;
      if (lm_ahead) {
        if (!h->lm_side[chunk]) {
          // small kernels on the critical path: dispatch their blocks ahead of queued SCF blocks
          int lo = 0, hi = 0;
          PNEC_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
          PNEC_CUDA(cudaStreamCreateWithPriority(&h->lm_side[chunk], cudaStreamNonBlocking, hi));
        }
        if (!h->ev_es[chunk]) PNEC_CUDA(cudaEventCreateWithFlags(&h->ev_es[chunk], cudaEventDisableTiming));
        ls = h->lm_side[chunk];
        PNEC_CUDA(cudaEventRecord(h->ev_es[chunk], cs));
        PNEC_CUDA(cudaStreamWaitEvent(ls, h->ev_es[chunk], 0));
      }
      auto run_lm_round = [&](int k) -> int {
        // a rotation that repeats once repeats forever: rot_same is both the skip flag and the result
        return run_es_lm(h, cnt, mom, round_poses(k - 1), round_poses(k), nullptr, nullptr, ls,
                         shortcuts ? rot_same : nullptr, shortcuts ? rot_same : nullptr, false);
      };
      auto run_scf_round = [&](int k) -> int {
        BatchView sv = bv;
        sv.poses = round_poses(k);
        return run_scf(h, sv, st.max_n, opts->ceres.regularization, opts->fibonacci_samples, opts->scf_steps,
                       round_poses(k) + 4, 7, nullptr, cs, shortcuts ? cache : nullptr, nullptr,
                       shortcuts ? fixed : nullptr, defer,
                       d_spill ? (st.bv.offsets ? d_spill : d_spill + 9 * c0 * st.bv.n_uniform) : nullptr,
                       round_poses(k - 1));
      };
      if (lm_ahead) {
        for (int k = 1; k <= rounds; ++k) {
          if ((rcc = run_lm_round(k)) != PNEC_OK) return rcc;
          cudaEvent_t &evk = h->ev_round[chunk][k - 1];
          if (!evk) PNEC_CUDA(cudaEventCreateWithFlags(&evk, cudaEventDisableTiming));
          PNEC_CUDA(cudaEventRecord(evk, ls));
        }
        for (int k = 1; k <= rounds; ++k) {
          PNEC_CUDA(cudaStreamWaitEvent(cs, h->ev_round[chunk][k - 1], 0));
          if ((rcc = run_scf_round(k)) != PNEC_OK) return rcc;
        }
      } else {
        for (int k = 1; k <= rounds; ++k) {
          if ((rcc = run_lm_round(k)) != PNEC_OK) return rcc;
          if ((rcc = run_scf_round(k)) != PNEC_OK) return rcc;
        }
      }
      init = round_poses(rounds);
      }
    } else if (!nec && opts->weighted_iterations == 0) {
      normalize_poses_kernel<<<static_cast<unsigned>((cnt + 127) / 128), 128, 0, cs>>>(caller_poses, pa, cnt);
      PNEC_CUDA(cudaGetLastError());
      h->launches++;
      init = pa;
    }
    PNEC_CUDA(mark(2));  // FrameTiming::it_es_: WeightedEigensolver (pnec.cc:180-188)
    // 4. refinement
    if (opts->use_ceres) {
      pnec_solver_opts so = opts->ceres;
      so.variant = variant;
      BatchView rv = bv;
      rv.poses = init;
      rcc = run_solve(h, rv, st.max_n, so, false, o_poses + 7 * c0, o_status ? o_status + c0 : nullptr,
                      o_iters ? o_iters + c0 : nullptr, o_cost ? o_cost + c0 : nullptr, nullptr, cs);
      if (rcc != PNEC_OK) return rcc;
    } else {
      PNEC_CUDA(cudaMemcpyAsync(o_poses + 7 * c0, init, static_cast<size_t>(cnt) * 56, cudaMemcpyDeviceToDevice, cs));
    }
    PNEC_CUDA(mark(3));  // FrameTiming::ceres_ (pnec.cc:165-169, 196-201)
    return PNEC_OK;
  };

  // Frame pairs are independent and every stage ends in a tail (a few pairs need 100 function
  // evaluations, or a sphere scan that cannot prune): the batch is cut into chunks that run their
  // stages on separate streams, so one chunk's tail is filled with another chunk's work.
  int chunks = h->cfg.frame_chunks;
  if (chunks <= 0) chunks = B >= 4096 ? 4 : (B >= 1024 ? 2 : 1);
  chunks = std::min<long long>(std::min(chunks, pnec_handle::kMaxChunks), B);
  if (fused_rounds) chunks = 1;  // the fused kernel lays its round poses out for one chunk
  if (timed) chunks = 1;         // the stage events are recorded on one stream
  if (chunks == 1) {
    rc = run_chunk(0, B, 0, stream);
    if (rc != PNEC_OK) return rc;
  } else {
    if (!h->ev_fork) PNEC_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    PNEC_CUDA(cudaEventRecord(h->ev_fork, stream));
    for (int c = 0; c < chunks; ++c) {
      if (!h->side[c]) PNEC_CUDA(cudaStreamCreateWithFlags(&h->side[c], cudaStreamNonBlocking));
      if (!h->ev_join[c]) PNEC_CUDA(cudaEventCreateWithFlags(&h->ev_join[c], cudaEventDisableTiming));
      PNEC_CUDA(cudaStreamWaitEvent(h->side[c], h->ev_fork, 0));
      const long long c0 = B * c / chunks, c1 = B * (c + 1) / chunks;
      rc = run_chunk(c0, c1 - c0, c, h->side[c]);
      if (rc != PNEC_OK) return rc;
      PNEC_CUDA(cudaEventRecord(h->ev_join[c], h->side[c]));
    }
    for (int c = 0; c < chunks; ++c) PNEC_CUDA(cudaStreamWaitEvent(stream, h->ev_join[c], 0));
  }
  if (host) {
    PNEC_CUDA(cudaMemcpyAsync(out->poses, o_poses, nb * 56, cudaMemcpyDeviceToHost, stream));
    if (out->status && opts->use_ceres)
      PNEC_CUDA(cudaMemcpyAsync(out->status, o_status, nb * 4, cudaMemcpyDeviceToHost, stream));
    if (out->iterations && opts->use_ceres)
      PNEC_CUDA(cudaMemcpyAsync(out->iterations, o_iters, nb * 4, cudaMemcpyDeviceToHost, stream));
    if (out->cost && opts->use_ceres)
      PNEC_CUDA(cudaMemcpyAsync(out->cost, o_cost, nb * 8, cudaMemcpyDeviceToHost, stream));
    if (out->es_poses)
      PNEC_CUDA(cudaMemcpyAsync(out->es_poses, d_es, nb * 56, cudaMemcpyDeviceToHost, stream));
    if (ransac) {
      if (out->num_inliers) PNEC_CUDA(cudaMemcpyAsync(out->num_inliers, ro.count, nb * 4, cudaMemcpyDeviceToHost, stream));
      if (out->ransac_iterations)
        PNEC_CUDA(cudaMemcpyAsync(out->ransac_iterations, ro.iters, nb * 4, cudaMemcpyDeviceToHost, stream));
      if (out->inlier_index)
        PNEC_CUDA(cudaMemcpyAsync(out->inlier_index, ro.index, static_cast<size_t>(st.bv.total) * 4,
                                  cudaMemcpyDeviceToHost, stream));
    } else {
      if (out->num_inliers) std::memset(out->num_inliers, 0, nb * 4);
      if (out->ransac_iterations) std::memset(out->ransac_iterations, 0, nb * 4);
    }
    PNEC_CUDA(cudaStreamSynchronize(stream));
  } else if (!ransac && out->ransac_iterations) {
    PNEC_CUDA(cudaMemsetAsync(out->ransac_iterations, 0, nb * 4, stream));
  }
  if (timed) {
    PNEC_CUDA(cudaEventSynchronize(h->ev_stage[3]));
    for (int k = 0; k < 3; ++k) PNEC_CUDA(cudaEventElapsedTime(&out->stage_ms[k], h->ev_stage[k], h->ev_stage[k + 1]));
  }
  scope.ok = true;
  return PNEC_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ keypoints -> batch

namespace {

struct KpStaged {
  KpAssembleArgs a;       // device pointers of tables / indices / outputs
  long long total = 0;
  int cov_doubles = 4;
  bool host = false, indexed = false;
};

int validate_keypoints(const pnec_keypoint_batch *kb, bool need_poses, long long *total_out) {
  if (!kb) return fail(PNEC_ERR_INVALID_ARGUMENT, "keypoint batch is NULL");
  if (kb->num_problems < 0 || kb->num_problems > 0x7fffffffLL) return fail(PNEC_ERR_INVALID_ARGUMENT, "bad num_problems");
  if (kb->memspace != PNEC_MEM_HOST && kb->memspace != PNEC_MEM_DEVICE)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "unknown memspace");
  long long total;
  if (kb->offsets) {
    if (kb->offsets[0] < 0) return fail(PNEC_ERR_INVALID_ARGUMENT, "offsets[0] < 0");
    for (int64_t i = 0; i < kb->num_problems; ++i)
      if (kb->offsets[i + 1] < kb->offsets[i] || kb->offsets[i + 1] - kb->offsets[i] > 0x7ffffff0LL)
        return fail(PNEC_ERR_INVALID_ARGUMENT, "offsets must be non-decreasing (and a pair below 2^31 correspondences)");
    total = kb->offsets[kb->num_problems];
  } else {
    if (kb->n_per_problem < 0 || kb->n_per_problem > 0x7ffffff0LL) return fail(PNEC_ERR_INVALID_ARGUMENT, "bad n_per_problem");
    if (kb->num_problems > 0 && kb->n_per_problem > (0x7fffffffffffffffLL / 72) / kb->num_problems)
      return fail(PNEC_ERR_INVALID_ARGUMENT, "num_problems * n_per_problem overflows");
    total = kb->num_problems * kb->n_per_problem;
  }
  if (total > 0) {
    if (!kb->host_points || !kb->target_points || !kb->K_inv)
      return fail(PNEC_ERR_INVALID_ARGUMENT, "keypoint tables / K_inv are NULL");
    if (kb->num_host_keypoints < 0 || kb->num_target_keypoints < 0 || kb->num_host_keypoints > 0x7fffffffLL ||
        kb->num_target_keypoints > 0x7fffffffLL)
      return fail(PNEC_ERR_INVALID_ARGUMENT, "bad keypoint table size");
    if (!kb->host_index && kb->num_host_keypoints < total)
      return fail(PNEC_ERR_INVALID_ARGUMENT, "host keypoint table shorter than the batch (no host_index given)");
    if (!kb->target_index && kb->num_target_keypoints < total)
      return fail(PNEC_ERR_INVALID_ARGUMENT, "target keypoint table shorter than the batch (no target_index given)");
    if (kb->memspace == PNEC_MEM_HOST) {  // host indices can be checked; device ones are the caller's contract
      if (kb->host_index)
        for (long long i = 0; i < total; ++i)
          if (kb->host_index[i] < 0 || kb->host_index[i] >= kb->num_host_keypoints)
            return fail(PNEC_ERR_INVALID_ARGUMENT, "host_index out of range");
      if (kb->target_index)
        for (long long i = 0; i < total; ++i)
          if (kb->target_index[i] < 0 || kb->target_index[i] >= kb->num_target_keypoints)
            return fail(PNEC_ERR_INVALID_ARGUMENT, "target_index out of range");
    }
  }
  if (need_poses && kb->num_problems > 0 && !kb->poses) return fail(PNEC_ERR_INVALID_ARGUMENT, "poses is NULL");
  *total_out = total;
  return PNEC_OK;
}

// Device buffers for the assembled batch (and, for HOST callers, for the tables); nothing is copied yet.
int kp_alloc(pnec_handle *h, const pnec_keypoint_batch *kb, long long total, KpStaged *ks) {
  const size_t nel = static_cast<size_t>(std::max<long long>(total, 1));
  const bool need_ct = kb->target_covs2 != nullptr, need_ch = kb->host_covs2 != nullptr;
  PNEC_CUDA(h->d_f1.ensure(nel * 24));
  PNEC_CUDA(h->d_f2.ensure(nel * 24));
  if (need_ct) PNEC_CUDA(h->d_ct.ensure(nel * 72));
  if (need_ch) PNEC_CUDA(h->d_ch.ensure(nel * 72));
  PNEC_CUDA(h->d_poses.ensure(static_cast<size_t>(std::max<long long>(kb->num_problems, 1)) * 56));
  ks->total = total;
  ks->cov_doubles = kb->packed_covs ? 3 : 4;
  ks->host = kb->memspace == PNEC_MEM_HOST;
  ks->indexed = kb->host_index || kb->target_index;
  KpAssembleArgs &a = ks->a;
  a = KpAssembleArgs{};
  a.f1 = static_cast<double *>(h->d_f1.p);
  a.f2 = static_cast<double *>(h->d_f2.p);
  a.ct = need_ct ? static_cast<double *>(h->d_ct.p) : nullptr;
  a.ch = need_ch ? static_cast<double *>(h->d_ch.p) : nullptr;
  a.packed = kb->packed_covs ? 1 : 0;
  for (int k = 0; k < 9; ++k) a.Kinv[k] = kb->K_inv[k];
  if (ks->host) {
    const size_t kh = static_cast<size_t>(std::max<long long>(kb->num_host_keypoints, 1));
    const size_t kt = static_cast<size_t>(std::max<long long>(kb->num_target_keypoints, 1));
    PNEC_CUDA(h->d_kp_hp.ensure(kh * 16));
    PNEC_CUDA(h->d_kp_tp.ensure(kt * 16));
    if (need_ch) PNEC_CUDA(h->d_kp_hc.ensure(kh * 8 * ks->cov_doubles));
    if (need_ct) PNEC_CUDA(h->d_kp_tc.ensure(kt * 8 * ks->cov_doubles));
    if (kb->host_index) PNEC_CUDA(h->d_kp_hi.ensure(nel * 4));
    if (kb->target_index) PNEC_CUDA(h->d_kp_ti.ensure(nel * 4));
    a.host_points = static_cast<const double *>(h->d_kp_hp.p);
    a.target_points = static_cast<const double *>(h->d_kp_tp.p);
    a.host_covs2 = need_ch ? static_cast<const double *>(h->d_kp_hc.p) : nullptr;
    a.target_covs2 = need_ct ? static_cast<const double *>(h->d_kp_tc.p) : nullptr;
    a.host_index = kb->host_index ? static_cast<const int *>(h->d_kp_hi.p) : nullptr;
    a.target_index = kb->target_index ? static_cast<const int *>(h->d_kp_ti.p) : nullptr;
  } else {
    a.host_points = kb->host_points;
    a.target_points = kb->target_points;
    a.host_covs2 = kb->host_covs2;
    a.target_covs2 = kb->target_covs2;
    a.host_index = kb->host_index;
    a.target_index = kb->target_index;
  }
  return PNEC_OK;
}

// H2D of what correspondences [e0, e1) need (HOST callers): with index arrays the whole tables (once,
// e0 == 0), else the matching rows.  Then the assembly kernel for [e0, e1).  Poses [p0, p1) ride along.
int kp_stage_range(pnec_handle *h, const pnec_keypoint_batch *kb, const KpStaged &ks, long long e0, long long e1,
                   long long p0, long long p1, cudaStream_t stream) {
  auto h2d = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
    if (bytes == 0) return cudaSuccess;
    if (bytes >= (4u << 20) && !h->cfg.no_stager && HostStager::pageable(src)) return h->stager.h2d(dst, src, bytes, stream);
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
  };
  auto at = [](const void *base, long long bytes) { return const_cast<char *>(static_cast<const char *>(base)) + bytes; };
  const size_t cd = static_cast<size_t>(ks.cov_doubles) * 8;
  if (ks.host) {
    if (ks.indexed) {
      if (e0 == 0) {
        PNEC_CUDA(h2d(h->d_kp_hp.p, kb->host_points, static_cast<size_t>(kb->num_host_keypoints) * 16));
        PNEC_CUDA(h2d(h->d_kp_tp.p, kb->target_points, static_cast<size_t>(kb->num_target_keypoints) * 16));
        if (ks.a.ch) PNEC_CUDA(h2d(h->d_kp_hc.p, kb->host_covs2, static_cast<size_t>(kb->num_host_keypoints) * cd));
        if (ks.a.ct) PNEC_CUDA(h2d(h->d_kp_tc.p, kb->target_covs2, static_cast<size_t>(kb->num_target_keypoints) * cd));
        if (kb->host_index) PNEC_CUDA(h2d(h->d_kp_hi.p, kb->host_index, static_cast<size_t>(ks.total) * 4));
        if (kb->target_index) PNEC_CUDA(h2d(h->d_kp_ti.p, kb->target_index, static_cast<size_t>(ks.total) * 4));
      }
    } else {
      const size_t n = static_cast<size_t>(e1 - e0);
      PNEC_CUDA(h2d(at(h->d_kp_hp.p, e0 * 16), at(kb->host_points, e0 * 16), n * 16));
      PNEC_CUDA(h2d(at(h->d_kp_tp.p, e0 * 16), at(kb->target_points, e0 * 16), n * 16));
      if (ks.a.ch) PNEC_CUDA(h2d(at(h->d_kp_hc.p, e0 * cd), at(kb->host_covs2, e0 * cd), n * cd));
      if (ks.a.ct) PNEC_CUDA(h2d(at(h->d_kp_tc.p, e0 * cd), at(kb->target_covs2, e0 * cd), n * cd));
    }
    if (kb->poses && p1 > p0)
      PNEC_CUDA(cudaMemcpyAsync(at(h->d_poses.p, p0 * 56), kb->poses + 7 * p0, static_cast<size_t>(p1 - p0) * 56,
                                cudaMemcpyHostToDevice, stream));
  }
  if (e1 > e0) {
    KpAssembleArgs a = ks.a;
    a.first = e0;
    a.count = e1 - e0;
    keypoint_assemble_kernel<<<static_cast<unsigned>((a.count + 127) / 128), 128, 0, stream>>>(a);
    PNEC_CUDA(cudaGetLastError());
    h->launches++;
  }
  return PNEC_OK;
}

pnec_batch kp_device_batch(pnec_handle *h, const pnec_keypoint_batch *kb, const KpStaged &ks) {
  pnec_batch b{};
  b.num_problems = kb->num_problems;
  b.n_per_problem = kb->n_per_problem;
  b.offsets = kb->offsets;
  b.memspace = PNEC_MEM_DEVICE;
  b.bvs_host = ks.a.f1;
  b.bvs_target = ks.a.f2;
  b.covs_target = ks.a.ct;
  b.covs_host = ks.a.ch;
  b.poses = ks.host ? static_cast<const double *>(h->d_poses.p) : kb->poses;
  return b;
}

}  // namespace

extern "C" {

int pnec_keypoints_to_batch(pnec_handle *h, const pnec_keypoint_batch *kb, pnec_batch *out_batch, void *cuda_stream) {
  if (!h || !out_batch) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  long long total = 0;
  int rc = validate_keypoints(kb, false, &total);
  if (rc != PNEC_OK) return rc;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
  KpStaged ks;
  rc = kp_alloc(h, kb, total, &ks);
  if (rc != PNEC_OK) return rc;
  rc = kp_stage_range(h, kb, ks, 0, total, 0, kb->num_problems, stream);
  if (rc != PNEC_OK) return rc;
  *out_batch = kp_device_batch(h, kb, ks);
  return PNEC_OK;
}

int pnec_solve_from_keypoints_batch(pnec_handle *h, const pnec_keypoint_batch *kb, const pnec_solver_opts *opts,
                                    const pnec_solve_out *out, void *cuda_stream) {
  if (!h || !opts || !out) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  long long total = 0;
  int rc = validate_keypoints(kb, true, &total);
  if (rc != PNEC_OK) return rc;
  const long long B = kb->num_problems;
  if (B > 0 && !out->poses) return fail(PNEC_ERR_INVALID_ARGUMENT, "out->poses is NULL");
  if (total > 0 && opts->variant != PNEC_VARIANT_NEC && !kb->target_covs2)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "target_covs2 is NULL for a PNEC variant");
  if (total > 0 && opts->variant == PNEC_VARIANT_SYMMETRIC && !kb->host_covs2)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "host_covs2 is NULL for the SYMMETRIC variant");
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
  KpStaged ks;
  rc = kp_alloc(h, kb, total, &ks);
  if (rc != PNEC_OK) return rc;
  pnec_batch db = kp_device_batch(h, kb, ks);
  if (!ks.host) {
    rc = kp_stage_range(h, kb, ks, 0, total, 0, B, stream);
    if (rc != PNEC_OK) return rc;
    return pnec_solve_batch(h, &db, opts, out, cuda_stream);
  }
  // HOST: device outputs, then chunks (copy + assembly of chunk k + 1 under the solve of chunk k)
  const size_t nb = static_cast<size_t>(B);
  const size_t sizes[5] = {56, 4, 4, 8, 8};
  void *host_ptr[5] = {out->poses, out->status, out->iterations, out->cost, out->initial_cost};
  void *dev_ptr[5] = {};
  for (int k = 0; k < 5; ++k) {
    if (!host_ptr[k]) continue;
    PNEC_CUDA(h->d_kp_out[k].ensure(nb * sizes[k]));
    dev_ptr[k] = h->d_kp_out[k].p;
  }
  int chunks = h->cfg.h2d_chunks;
  if (chunks <= 0) chunks = (total * 64 >= (32ll << 20) && B >= 64 && !ks.indexed) ? 4 : 1;
  if (ks.indexed) chunks = 1;
  chunks = static_cast<int>(std::min<long long>(std::min(chunks, pnec_handle::kMaxChunks), B));
  if (!h->ev_fork) PNEC_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  PNEC_CUDA(cudaEventRecord(h->ev_fork, stream));
  scope.forked = true;
  for (int c = 0; c < chunks; ++c) {
    if (!h->side[c]) PNEC_CUDA(cudaStreamCreateWithFlags(&h->side[c], cudaStreamNonBlocking));
    if (!h->ev_join[c]) PNEC_CUDA(cudaEventCreateWithFlags(&h->ev_join[c], cudaEventDisableTiming));
    cudaStream_t cs = h->side[c];
    PNEC_CUDA(cudaStreamWaitEvent(cs, h->ev_fork, 0));
    long long p0, p1;
    if (kb->offsets) {
      auto first_at = [&](long long target) { return std::lower_bound(kb->offsets, kb->offsets + B, target) - kb->offsets; };
      p0 = c == 0 ? 0 : first_at(total * c / chunks);
      p1 = c + 1 == chunks ? B : first_at(total * (c + 1) / chunks);
    } else {
      p0 = B * c / chunks;
      p1 = B * (c + 1) / chunks;
    }
    if (p1 > p0) {
      const long long e0 = kb->offsets ? kb->offsets[p0] : p0 * kb->n_per_problem;
      const long long e1 = kb->offsets ? kb->offsets[p1] : p1 * kb->n_per_problem;
      rc = kp_stage_range(h, kb, ks, e0, e1, p0, p1, cs);
      if (rc != PNEC_OK) return rc;
      pnec_batch sb = db;  // pairs [p0, p1): ragged views index absolutely, uniform ones move their bases
      sb.num_problems = p1 - p0;
      sb.poses = db.poses + 7 * p0;
      if (kb->offsets) {
        sb.offsets = kb->offsets + p0;
      } else {
        sb.bvs_host += 3 * e0;
        sb.bvs_target += 3 * e0;
        if (sb.covs_target) sb.covs_target += 9 * e0;
        if (sb.covs_host) sb.covs_host += 9 * e0;
      }
      pnec_solve_out so{};
      so.poses = static_cast<double *>(dev_ptr[0]) + 7 * p0;
      so.status = dev_ptr[1] ? static_cast<int32_t *>(dev_ptr[1]) + p0 : nullptr;
      so.iterations = dev_ptr[2] ? static_cast<int32_t *>(dev_ptr[2]) + p0 : nullptr;
      so.cost = dev_ptr[3] ? static_cast<double *>(dev_ptr[3]) + p0 : nullptr;
      so.initial_cost = dev_ptr[4] ? static_cast<double *>(dev_ptr[4]) + p0 : nullptr;
      rc = pnec_solve_batch(h, &sb, opts, &so, cs);
      if (rc != PNEC_OK) return rc;
      for (int k = 0; k < 5; ++k)
        if (host_ptr[k])
          PNEC_CUDA(cudaMemcpyAsync(static_cast<char *>(host_ptr[k]) + sizes[k] * p0,
                                    static_cast<char *>(dev_ptr[k]) + sizes[k] * p0, sizes[k] * (p1 - p0),
                                    cudaMemcpyDeviceToHost, cs));
    }
    PNEC_CUDA(cudaEventRecord(h->ev_join[c], cs));
  }
  for (int c = 0; c < chunks; ++c) PNEC_CUDA(cudaStreamWaitEvent(stream, h->ev_join[c], 0));
  PNEC_CUDA(cudaStreamSynchronize(stream));
  scope.ok = true;
  return PNEC_OK;
}

int pnec_frame_solve_from_keypoints_batch(pnec_handle *h, const pnec_keypoint_batch *kb, const pnec_frame_opts *opts,
                                          const pnec_frame_out *out, void *cuda_stream) {
  if (!h || !opts || !out) return fail(PNEC_ERR_INVALID_ARGUMENT, "NULL argument");
  long long total = 0;
  int rc = validate_keypoints(kb, true, &total);
  if (rc != PNEC_OK) return rc;
  const long long B = kb->num_problems;
  if (B > 0 && !out->poses) return fail(PNEC_ERR_INVALID_ARGUMENT, "out->poses is NULL");
  if (total > 0 && !opts->use_nec && !kb->target_covs2)
    return fail(PNEC_ERR_INVALID_ARGUMENT, "target_covs2 is NULL for the PNEC pipeline");
  if (B == 0) return PNEC_OK;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  PNEC_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  CallScope scope(h, stream);
  KpStaged ks;
  rc = kp_alloc(h, kb, total, &ks);
  if (rc != PNEC_OK) return rc;
  rc = kp_stage_range(h, kb, ks, 0, total, 0, B, stream);
  if (rc != PNEC_OK) return rc;
  pnec_batch db = kp_device_batch(h, kb, ks);
  if (!ks.host) return pnec_frame_solve_batch(h, &db, opts, out, cuda_stream);
  // HOST: results through device buffers of the handle
  const size_t nb = static_cast<size_t>(B), nel = static_cast<size_t>(std::max<long long>(total, 1));
  const size_t sizes[8] = {56, 56, 4, 4, 8, 4, 4, 4}, counts[8] = {nb, nb, nb, nb, nb, nb, nel, nb};
  void *host_ptr[8] = {out->poses, out->es_poses, out->status, out->iterations, out->cost, out->num_inliers,
                       out->inlier_index, out->ransac_iterations};
  void *dev_ptr[8] = {};
  for (int k = 0; k < 8; ++k) {
    if (!host_ptr[k]) continue;
    PNEC_CUDA(h->d_kp_out[k].ensure(counts[k] * sizes[k]));
    dev_ptr[k] = h->d_kp_out[k].p;
  }
  pnec_frame_out fo{};
  fo.poses = static_cast<double *>(dev_ptr[0]);
  fo.es_poses = static_cast<double *>(dev_ptr[1]);
  fo.status = static_cast<int32_t *>(dev_ptr[2]);
  fo.iterations = static_cast<int32_t *>(dev_ptr[3]);
  fo.cost = static_cast<double *>(dev_ptr[4]);
  fo.num_inliers = static_cast<int32_t *>(dev_ptr[5]);
  fo.inlier_index = static_cast<int32_t *>(dev_ptr[6]);
  fo.ransac_iterations = static_cast<int32_t *>(dev_ptr[7]);
  fo.stage_ms = out->stage_ms;
  rc = pnec_frame_solve_batch(h, &db, opts, &fo, cuda_stream);
  if (rc != PNEC_OK) return rc;
  for (int k = 0; k < 8; ++k) {
    if (!host_ptr[k]) continue;
    if ((k == 2 || k == 3 || k == 4) && !opts->use_ceres) continue;  // untouched without the refinement
    if (k == 6 && !opts->use_ransac) continue;                      // no inlier list without RANSAC
    PNEC_CUDA(cudaMemcpyAsync(host_ptr[k], dev_ptr[k], counts[k] * sizes[k], cudaMemcpyDeviceToHost, stream));
  }
  PNEC_CUDA(cudaStreamSynchronize(stream));
  return PNEC_OK;
}

}  // extern "C"
