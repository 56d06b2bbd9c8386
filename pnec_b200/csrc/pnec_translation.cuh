// pnec_translation.cuh — translation given rotation:
//   scf_kernel   the SCF stage of PNEC::WeightedEigensolver: A_i / B_i construction
//                (src/rel_pose_estimation/pnec.cc:317-328), Fibonacci-sphere scan of the
//                sum of Rayleigh quotients (pnec.cc:330-340, src/optimization/scf.cc:43-72)
//                and the self-consistent-field iteration (scf.cc:109-147)
//   nec_translation_kernel   TranslationFromM(ComposeM(bvs_1, bvs_2, R)),
//                src/common/common.cc:127-181 (the NEC translation of PNEC::Eigensolver,
//                pnec.cc:270-278)
// One CTA per frame pair.  The scan is the heavy part: (samples + 1) x N Rayleigh quotients
// per pair, evaluated from 9 doubles per correspondence (n_i and sym(B_i)) kept in shared memory.
#pragma once

#include "pnec_batch.cuh"
#include "pnec_lm.cuh"

namespace pnec {

// Eigenvector of the smallest eigenvalue of a symmetric 3x3 (xx, xy, xz, yy, yz, zz), unit
// norm, by cyclic Jacobi rotations.  Sign: first non-negligible component positive is NOT
// enforced — like Eigen's solvers the sign is arbitrary; callers compare modulo sign.
__device__ __forceinline__ void sym3_smallest_eigvec(const double m[6], double v[3], double &lambda) {
  double a[3][3] = {{m[0], m[1], m[2]}, {m[1], m[3], m[4]}, {m[2], m[4], m[5]}};
  double q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-34 * diag || off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0, r = (pq == 0) ? 1 : 2;
      const double apq = a[p][r];
      if (apq == 0.0) continue;
      const double theta = (a[r][r] - a[p][p]) / (2.0 * apq);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = rsqrt(t * t + 1.0), s = t * c;
#pragma unroll
      for (int k = 0; k < 3; ++k) {  // A <- A J
        const double akp = a[k][p], akr = a[k][r];
        a[k][p] = c * akp - s * akr;
        a[k][r] = s * akp + c * akr;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {  // A <- J^T A
        const double apk = a[p][k], ark = a[r][k];
        a[p][k] = c * apk - s * ark;
        a[r][k] = s * apk + c * ark;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {  // Q <- Q J
        const double qkp = q[k][p], qkr = q[k][r];
        q[k][p] = c * qkp - s * qkr;
        q[k][r] = s * qkp + c * qkr;
      }
    }
  }
  int j = 0;
  if (a[1][1] < a[j][j]) j = 1;
  if (a[2][2] < a[j][j]) j = 2;
  lambda = a[j][j];
  const double x = q[0][j], y = q[1][j], z = q[2][j];
  const double inv = rsqrt(x * x + y * y + z * z);
  v[0] = x * inv; v[1] = y * inv; v[2] = z * inv;
}

// Rotation matrix (row-major) of the normalised stored quaternion: Sophus::SE3d::rotationMatrix().
__device__ __forceinline__ void pose_rotation(const double *pose7, double R[9]) {
  const double qn = sqrt(pose7[0] * pose7[0] + pose7[1] * pose7[1] + pose7[2] * pose7[2] + pose7[3] * pose7[3]);
  const double x[6] = {0.0, 0.0, pose7[0] / qn, pose7[1] / qn, pose7[2] / qn, pose7[3] / qn};
  PoseConst pc;
  make_pose_const(x, pc);
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = pc.R[k];
}

template <int NW>
__device__ __forceinline__ void block_sum6(double v[6], double (*s_red)[6], int warp, int lane) {
#pragma unroll
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) s_red[warp][k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += s_red[w][k];
    v[k] = t;
  }
  __syncthreads();
}

// What a full scan of the sphere found for one frame pair, keyed by the rotation it was made with.
// The sphere costs depend on the rotation and the data only, so a later call with the bit-identical
// quaternion (the weighted-eigensolver iterations of PNEC::WeightedEigensolver converge to one
// within 1-3 rounds) can reuse them: the result is exactly what rescanning would give.
struct ScfScanCache {
  double q[4];
  double best_cost;
  int best_idx;  // 1-based sphere index
  int valid;
};

struct ScfArgs {
  BatchView bv;           // poses: rotation quaternion + start translation
  const double *sphere;   // [samples][3] fibonacci_sphere(samples)
  double *out_t;          // translation of pair b at out_t + out_stride * b
  int out_stride;         // 3 for a [B][3] array, 7 to write into the translation slot of a pose array
  double *out_cost;       // [B] obj_fun at the result, or nullptr
  double reg;
  int samples, steps;
  int cap_elems;          // correspondences that fit the dynamic shared memory
  ScfScanCache *cache;    // [B] or nullptr: reuse / record the sphere scan per rotation
  const int *q_same;      // [B] or nullptr: this call's rotation equals the previous round's bit for bit
  int *fixed;             // [B] or nullptr: in: pair already at a fixed point of the iteration (skip);
                          //     out: set when q_same and the translation did not move either
};

// per correspondence: n = f1 x R f2 and sym(B), B = [f1]x R S R^T [f1]x^T + reg I
__device__ __forceinline__ void scf_terms(const double R[9], double reg, const double f1[3],
                                          const double f2[3], const double c9[9], double out[9]) {
  double g[3], n[3];
  rot(R, f2, g);
  cross3(f1, g, n);
  // G = [f1]x R: column j of G is f1 x (column j of R)
  double G[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double col[3] = {R[j], R[3 + j], R[6 + j]};
    double gc[3];
    cross3(f1, col, gc);
    G[0][j] = gc[0]; G[1][j] = gc[1]; G[2][j] = gc[2];
  }
  double s6[6];
  pack_sym(c9, s6);
  // B = G S G^T: rows of G are vectors u_r; B_rs = u_r^T S u_s
  double Su[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r) sym_mul6(s6, G[r], Su[r]);
  out[0] = n[0]; out[1] = n[1]; out[2] = n[2];
  out[3] = dot3(G[0], Su[0]) + reg;  // xx
  out[4] = dot3(G[0], Su[1]);        // xy
  out[5] = dot3(G[0], Su[2]);        // xz
  out[6] = dot3(G[1], Su[1]) + reg;  // yy
  out[7] = dot3(G[1], Su[2]);        // yz
  out[8] = dot3(G[2], Su[2]) + reg;  // zz
}

// sum_i (t.n_i)^2 / (t^T B_i t) over the correspondences in shared memory, in index order
// (obj_fun, scf.cc:43-51)
__device__ __forceinline__ double scf_objective(const double *terms, int n, const double t[3]) {
  const double txx = t[0] * t[0], txy = 2.0 * t[0] * t[1], txz = 2.0 * t[0] * t[2];
  const double tyy = t[1] * t[1], tyz = 2.0 * t[1] * t[2], tzz = t[2] * t[2];
  double cost = 0.0;
  for (int i = 0; i < n; ++i) {
    const double *w = terms + 9 * i;  // same address in every lane: broadcast
    const double e = t[0] * w[0] + t[1] * w[1] + t[2] * w[2];
    const double den = w[3] * txx + w[4] * txy + w[5] * txz + w[6] * tyy + w[7] * tyz + w[8] * tzz;
    cost = fma(e * e, fast_rcp(den), cost);
  }
  return cost;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) scf_kernel(const __grid_constant__ ScfArgs args) {
  constexpr int NT = NW * 32;
  __shared__ double s_red[NW][6];
  __shared__ double s_best_cost[NW];
  __shared__ int s_best_idx[NW];
  __shared__ double s_t[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = blockIdx.x;
  long long s, e;
  problem_range(args.bv, b, s, e);
  const int n = static_cast<int>(e - s);
  const double *pose = args.bv.poses + 7 * b;
  double *terms = dyn_smem;  // [n][9]
  double *ot = args.out_t + static_cast<long long>(args.out_stride) * b;
  if (args.fixed && args.fixed[b]) return;  // fixed point of the iteration: the result is already in place
  if (n <= 0 || n > args.cap_elems) {
    // nothing to minimise (or a pair beyond the shared-memory capacity, rejected on the host)
    if (tid == 0) { ot[0] = pose[4]; ot[1] = pose[5]; ot[2] = pose[6]; if (args.out_cost) args.out_cost[b] = 0.0; }
    return;
  }
  double R[9];
  pose_rotation(pose, R);
  for (int i = tid; i < n; i += NT) {
    double a1[3], a2[3], c9[9], w[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) { a1[k] = args.bv.f1[3 * (s + i) + k]; a2[k] = args.bv.f2[3 * (s + i) + k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) c9[k] = args.bv.ct[9 * (s + i) + k];
    scf_terms(R, args.reg, a1, a2, c9, w);
#pragma unroll
    for (int k = 0; k < 9; ++k) terms[9 * i + k] = w[k];
  }
  __syncthreads();

  // ---- scan: candidate 0 is the given translation, 1..samples the Fibonacci sphere; the first
  // strict minimum wins (pnec.cc:332-340).  Candidate 0 is summed by the whole CTA; the sphere
  // candidates are summed one per thread in index order, or taken from the cache when this exact
  // rotation was scanned before.
  const double t0[3] = {pose[4], pose[5], pose[6]};
  double cost0;
  {
    const double txx = t0[0] * t0[0], txy = 2.0 * t0[0] * t0[1], txz = 2.0 * t0[0] * t0[2];
    const double tyy = t0[1] * t0[1], tyz = 2.0 * t0[1] * t0[2], tzz = t0[2] * t0[2];
    double part = 0.0;
    for (int i = tid; i < n; i += NT) {
      const double *w = terms + 9 * i;
      const double e = t0[0] * w[0] + t0[1] * w[1] + t0[2] * w[2];
      const double den = w[3] * txx + w[4] * txy + w[5] * txz + w[6] * tyy + w[7] * tyz + w[8] * tzz;
      part = fma(e * e, fast_rcp(den), part);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_best_cost[warp] = part;
    __syncthreads();
    cost0 = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) cost0 += s_best_cost[w];
    __syncthreads();
  }
  bool reuse = false;
  double best = CUDART_INF;
  int best_idx = 0x7fffffff;
  if (args.cache) {
    const ScfScanCache &c = args.cache[b];
    reuse = c.valid && __double_as_longlong(c.q[0]) == __double_as_longlong(pose[0]) &&
            __double_as_longlong(c.q[1]) == __double_as_longlong(pose[1]) &&
            __double_as_longlong(c.q[2]) == __double_as_longlong(pose[2]) &&
            __double_as_longlong(c.q[3]) == __double_as_longlong(pose[3]);
    if (reuse) { best = c.best_cost; best_idx = c.best_idx; }
  }
  if (!reuse) {  // block-uniform
    for (int c = 1 + tid; c <= args.samples; c += NT) {
      const double t[3] = {args.sphere[3 * (c - 1)], args.sphere[3 * (c - 1) + 1], args.sphere[3 * (c - 1) + 2]};
      const double cost = scf_objective(terms, n, t);
      if (cost < best || (cost == best && c < best_idx)) { best = cost; best_idx = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double oc = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
      if (oc < best || (oc == best && oi < best_idx)) { best = oc; best_idx = oi; }
    }
    if (lane == 0) { s_best_cost[warp] = best; s_best_idx[warp] = best_idx; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < NW; ++w)
        if (s_best_cost[w] < best || (s_best_cost[w] == best && s_best_idx[w] < best_idx)) {
          best = s_best_cost[w]; best_idx = s_best_idx[w];
        }
      if (args.cache) {
        ScfScanCache &c = args.cache[b];
        c.q[0] = pose[0]; c.q[1] = pose[1]; c.q[2] = pose[2]; c.q[3] = pose[3];
        c.best_cost = best; c.best_idx = best_idx; c.valid = 1;
      }
    }
  }
  if (tid == 0) {
    if (best_idx != 0x7fffffff && best < cost0) {
      s_t[0] = args.sphere[3 * (best_idx - 1)]; s_t[1] = args.sphere[3 * (best_idx - 1) + 1]; s_t[2] = args.sphere[3 * (best_idx - 1) + 2];
    } else {
      s_t[0] = t0[0]; s_t[1] = t0[1]; s_t[2] = t0[2];
    }
  }
  __syncthreads();

  // ---- SCF: t <- eigenvector of the smallest eigenvalue of E(t) = sum_i A_i / (t^T B_i t)
  // (alt_construct_E with frac[i] == 0, scf.cc:109-126: `frac.resize(n)` then push_back)
  for (int it = 0; it < args.steps; ++it) {
    const double t[3] = {s_t[0], s_t[1], s_t[2]};
    const double txx = t[0] * t[0], txy = 2.0 * t[0] * t[1], txz = 2.0 * t[0] * t[2];
    const double tyy = t[1] * t[1], tyz = 2.0 * t[1] * t[2], tzz = t[2] * t[2];
    double E[6] = {0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += NT) {
      const double *w = terms + 9 * i;
      const double den = w[3] * txx + w[4] * txy + w[5] * txz + w[6] * tyy + w[7] * tyz + w[8] * tzz;
      const double inv = fast_rcp(den);
      E[0] = fma(w[0] * w[0], inv, E[0]); E[1] = fma(w[0] * w[1], inv, E[1]); E[2] = fma(w[0] * w[2], inv, E[2]);
      E[3] = fma(w[1] * w[1], inv, E[3]); E[4] = fma(w[1] * w[2], inv, E[4]); E[5] = fma(w[2] * w[2], inv, E[5]);
    }
    block_sum6<NW>(E, s_red, warp, lane);
    if (tid == 0) {
      double v[3], lam;
      sym3_smallest_eigvec(E, v, lam);
      s_t[0] = v[0]; s_t[1] = v[1]; s_t[2] = v[2];
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (args.fixed && args.q_same && args.q_same[b] &&
        __double_as_longlong(s_t[0]) == __double_as_longlong(t0[0]) &&
        __double_as_longlong(s_t[1]) == __double_as_longlong(t0[1]) &&
        __double_as_longlong(s_t[2]) == __double_as_longlong(t0[2]))
      args.fixed[b] = 1;  // (rotation, translation) -> itself: every further round repeats it
    ot[0] = s_t[0]; ot[1] = s_t[1]; ot[2] = s_t[2];
    if (args.out_cost) {
      const double t[3] = {s_t[0], s_t[1], s_t[2]};
      args.out_cost[b] = scf_objective(terms, n, t);
    }
  }
}

// TranslationFromM(ComposeM(bvs_1, bvs_2, R)), common.cc:127-181.  ComposeM starts its loop at
// i = 1 (common.cc:131): the first correspondence of a pair is skipped, as in the reference.
__global__ void __launch_bounds__(128) nec_translation_kernel(BatchView bv, double *out_t, int out_stride, double *out_M) {
  __shared__ double s_red[4][6];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = blockIdx.x;
  long long s, e;
  problem_range(bv, b, s, e);
  double R[9];
  pose_rotation(bv.poses + 7 * b, R);
  double M[6] = {0, 0, 0, 0, 0, 0};
  for (long long i = s + 1 + tid; i < e; i += 128) {
    double f1[3], f2[3], g[3], n[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { f1[k] = bv.f1[3 * i + k]; f2[k] = bv.f2[3 * i + k]; }
    rot(R, f2, g);
    cross3(f1, g, n);
    M[0] = fma(n[0], n[0], M[0]); M[1] = fma(n[0], n[1], M[1]); M[2] = fma(n[0], n[2], M[2]);
    M[3] = fma(n[1], n[1], M[3]); M[4] = fma(n[1], n[2], M[4]); M[5] = fma(n[2], n[2], M[5]);
  }
  block_sum6<4>(M, s_red, warp, lane);
  if (tid == 0) {
    double v[3], lam;
    sym3_smallest_eigvec(M, v, lam);
    double *ot = out_t + static_cast<long long>(out_stride) * b;
    ot[0] = v[0]; ot[1] = v[1]; ot[2] = v[2];
    if (out_M) {
#pragma unroll
      for (int k = 0; k < 6; ++k) out_M[6 * b + k] = M[k];
    }
  }
}

}  // namespace pnec
