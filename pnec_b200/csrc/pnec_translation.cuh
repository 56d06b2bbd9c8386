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

// The same eigenvector without the rotation sweeps, for the SCF loop where one lane solves ten
// eigenproblems in a row while its CTA waits: smallest eigenvalue from the trigonometric closed
// form of the characteristic cubic, eigenvector as the largest cross product of two rows of
// (A - lambda I), then one Rayleigh-quotient refinement (the closed form is only accurate to
// eps * |A|; the Rayleigh quotient of the first vector is accurate to second order).  Accuracy of
// the direction: eps * |A| / (lambda_2 - lambda_1), the same conditioning as any method.  Falls back
// to the Jacobi routine when the matrix is (numerically) of rank <= 1 or not finite.
__device__ __forceinline__ void sym3_null_vector(const double m[6], double lam, double v[3], double &nrm2) {
  const double r0[3] = {m[0] - lam, m[1], m[2]};
  const double r1[3] = {m[1], m[3] - lam, m[4]};
  const double r2[3] = {m[2], m[4], m[5] - lam};
  double c01[3], c02[3], c12[3];
  cross3(r0, r1, c01);
  cross3(r0, r2, c02);
  cross3(r1, r2, c12);
  const double n01 = dot3(c01, c01), n02 = dot3(c02, c02), n12 = dot3(c12, c12);
  nrm2 = n01;
  v[0] = c01[0]; v[1] = c01[1]; v[2] = c01[2];
  if (n02 > nrm2) { nrm2 = n02; v[0] = c02[0]; v[1] = c02[1]; v[2] = c02[2]; }
  if (n12 > nrm2) { nrm2 = n12; v[0] = c12[0]; v[1] = c12[1]; v[2] = c12[2]; }
}

__device__ __noinline__ void sym3_smallest_eigvec_jacobi(const double m[6], double v[3], double &lambda) {
  sym3_smallest_eigvec(m, v, lambda);
}

__device__ __forceinline__ void sym3_smallest_eigvec_fast(const double m_in[6], double v[3], double &lambda) {
  // scale to unit size: the closed form cubes the entries
  double big = fmax(fmax(fabs(m_in[0]), fabs(m_in[3])), fabs(m_in[5]));
  big = fmax(big, fmax(fmax(fabs(m_in[1]), fabs(m_in[2])), fabs(m_in[4])));
  if (!(big > 0.0) || !(big < CUDART_INF)) { sym3_smallest_eigvec_jacobi(m_in, v, lambda); return; }
  const double inv_big = fast_rcp(big);
  double m[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) m[k] = m_in[k] * inv_big;
  const double q = (m[0] + m[3] + m[5]) * (1.0 / 3.0);
  const double p1 = m[1] * m[1] + m[2] * m[2] + m[4] * m[4];
  const double d0 = m[0] - q, d1 = m[3] - q, d2 = m[5] - q;
  const double p2 = d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * p1;
  double lam = q;
  if (p2 > 0.0) {
    const double ip = rsqrt(p2 * (1.0 / 6.0));  // 1 / p
    const double b0 = d0 * ip, b3 = d1 * ip, b5 = d2 * ip, b1 = m[1] * ip, b2 = m[2] * ip, b4 = m[4] * ip;
    const double det = b0 * (b3 * b5 - b4 * b4) - b1 * (b1 * b5 - b4 * b2) + b2 * (b1 * b4 - b3 * b2);
    const double r = fmin(1.0, fmax(-1.0, 0.5 * det));
    // r -> 1: the two smallest eigenvalues approach each other (the translation is then
    // ill-determined, e.g. near-pure rotation) and the smallest root of the cubic below becomes a
    // double root: rotation sweeps instead
    if (1.0 - r < 1e-4) { sym3_smallest_eigvec_jacobi(m_in, v, lambda); return; }
    // eigenvalues of (A - q I) / p are 2 y with 4 y^3 - 3 y = r (y = cos of the trigonometric
    // form); the smallest one lies in [-1, -1/2], where the cubic is increasing and concave, so
    // Newton from y = -1 climbs to it monotonically.
    double y = -1.0;
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
      const double f = fma(fma(4.0 * y, y, -3.0), y, -r);
      const double fp = fma(12.0 * y, y, -3.0);
      const double dy = f * fast_rcp(fp);
      y -= dy;
      if (fabs(dy) <= 4e-16) break;
    }
    lam = fma(2.0 * y, p2 * (1.0 / 6.0) * ip, q);  // q + 2 p y, p = p2/6 * (1/p)
  }
  double u[3], n2;
  sym3_null_vector(m, lam, u, n2);
  if (!(n2 > 1e-60)) { sym3_smallest_eigvec_jacobi(m_in, v, lambda); return; }
  // Rayleigh quotient of u, then the null vector again
  double Au[3];
  sym_mul6(m, u, Au);
  const double rq = dot3(u, Au) * fast_rcp(n2);
  double w[3], n2b;
  sym3_null_vector(m, rq, w, n2b);
  if (!(n2b > 1e-60)) { w[0] = u[0]; w[1] = u[1]; w[2] = u[2]; n2b = n2; }
  const double inv = rsqrt(n2b);
  v[0] = w[0] * inv; v[1] = w[1] * inv; v[2] = w[2] * inv;
  double Av[3];
  sym_mul6(m, v, Av);
  lambda = dot3(v, Av) * big;
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

// What a scan of the sphere established for one frame pair, keyed by the rotation it was made with.
// The sphere costs depend on the rotation and the data only, so a later call with the bit-identical
// quaternion (the weighted-eigensolver rounds of PNEC::WeightedEigensolver converge to one within
// 1-3 rounds) can reuse it: the outcome is exactly what rescanning would give.
//   idx > 0:  the sphere minimum is `value`, first reached at sphere point idx (1-based)
//   idx == 0: every sphere point costs at least `value` (all were pruned against it)
struct ScfScanCache {
  double q[4];
  double value;
  int idx;
  int valid;
};

constexpr int kScfPrefix = 32;      // correspondences every sphere candidate is summed over before pruning
constexpr int kScfMaxSurvivors = 512;

struct ScfArgs {
  BatchView bv;           // poses: rotation quaternion + start translation
  const double *sphere;   // [samples][3] fibonacci_sphere(samples)
  double *out_t;          // translation of pair b at out_t + out_stride * b
  int out_stride;         // 3 for a [B][3] array, 7 to write into the translation slot of a pose array
  double *out_cost;       // [B] obj_fun at the result, or nullptr
  double reg;
  int samples, steps;
  int cap_elems;          // correspondences that fit the dynamic shared memory
  double *spill;          // [total][9] terms of pairs above cap_elems (HBM / L2 instead of shared
                          // memory: slower, same arithmetic), or nullptr if every pair fits
  ScfScanCache *cache;    // [B] or nullptr: reuse / record the sphere scan per rotation
  const int *q_same;      // [B] or nullptr: this call's rotation equals the previous round's bit for bit
  const double *prev_poses;  // [B][7] or nullptr: the previous round's poses.  When given, the start
                          // translation is taken from THEM (bv.poses then only supplies the rotation),
                          // q_same is decided by comparing the two quaternions, and a pair at a fixed
                          // point copies its translation over instead of leaving the output untouched
  int *fixed;             // [B] or nullptr: in: pair already at a fixed point of the iteration (skip);
                          //     out: set when q_same and the translation did not move either
  // Two passes: the first (4 warps per pair) hands pairs whose scan cannot prune (flat cost
  // landscape, e.g. near-pure rotation: hundreds of survivors) to a list; the second runs that list
  // with 16 warps per pair on a few persistent CTAs, so that one such pair no longer sets the
  // duration of the whole launch.
  int *defer_list;        // pass 1: [B] pairs handed to pass 2, or nullptr (no deferral)
  int *defer_count;       // pass 1: number of entries of defer_list
  int defer_threshold;    // pass 1: survivors above which a pair is deferred
  const int *work_list;   // pass 2: pairs to process (CTAs take entries through work_cursor), else nullptr
  const int *work_count;
  int *work_cursor;
  long long *dbg;         // [B][4] or nullptr: path (0 fixed, 1 scan reused, 2 scanned), survivors,
                          //     cycles of the scan, cycles of the whole CTA (PNEC_B200_SCF_DEBUG)
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

// One Rayleigh quotient (t.n)^2 / (t^T B t) added to `acc`, with the rounding sequence spelled out
// (explicit fma / __dmul_rn: nothing for the compiler to contract differently in different kernels or
// call sites), so that every path that sums a candidate's terms in the same order gets the same bits.
struct ScfDir {
  double t0, t1, t2, txx, txy, txz, tyy, tyz, tzz;
};
__device__ __forceinline__ ScfDir scf_dir(const double t[3]) {
  ScfDir d;
  d.t0 = t[0]; d.t1 = t[1]; d.t2 = t[2];
  d.txx = __dmul_rn(t[0], t[0]); d.txy = __dmul_rn(2.0 * t[0], t[1]); d.txz = __dmul_rn(2.0 * t[0], t[2]);
  d.tyy = __dmul_rn(t[1], t[1]); d.tyz = __dmul_rn(2.0 * t[1], t[2]); d.tzz = __dmul_rn(t[2], t[2]);
  return d;
}
__device__ __forceinline__ double scf_add_term(const ScfDir &d, const double *w, double acc) {
  const double e = fma(d.t2, w[2], fma(d.t1, w[1], __dmul_rn(d.t0, w[0])));
  // t^T B t as three independent pairs (a shorter dependent chain than one long fma ladder)
  const double den = __dadd_rn(__dadd_rn(fma(w[4], d.txy, __dmul_rn(w[3], d.txx)), fma(w[6], d.tyy, __dmul_rn(w[5], d.txz))),
                               fma(w[8], d.tzz, __dmul_rn(w[7], d.tyz)));
  return fma(__dmul_rn(e, e), fast_rcp(den), acc);
}

// sum_i (t.n_i)^2 / (t^T B_i t) over the correspondences in shared memory, in index order
// (obj_fun, scf.cc:43-51)
__device__ __forceinline__ double scf_objective(const double *terms, int n, const double t[3]) {
  const ScfDir d = scf_dir(t);
  double cost = 0.0;
  for (int i = 0; i < n; ++i) cost = scf_add_term(d, terms + 9 * i, cost);  // same address in every lane: broadcast
  return cost;
}

// The same sum, abandoned as soon as it exceeds `bound` (checked every 4 terms).  The terms are
// non-negative, so the partial sums only grow: a sum that is abandoned would have ended above the
// bound as well, and one that is not abandoned is returned unchanged (same operations, same order).
__device__ __forceinline__ double scf_objective_bounded(const double *terms, int n, const double t[3], double bound) {
  const ScfDir d = scf_dir(t);
  double cost = 0.0;
  for (int i0 = 0; i0 < n; i0 += 4) {
    const int i1 = min(i0 + 4, n);
    for (int i = i0; i < i1; ++i) cost = scf_add_term(d, terms + 9 * i, cost);
    if (cost > bound) break;
  }
  return cost;
}

// The rest of a candidate's sum, correspondences [begin, n), by one lane in index order.
__device__ __forceinline__ double scf_objective_tail(const double *terms, int begin, int n, const double t[3]) {
  const ScfDir d = scf_dir(t);
  double cost = 0.0;
  for (int i = begin; i < n; ++i) cost = scf_add_term(d, terms + 9 * i, cost);
  return cost;
}

// One 64-correspondence chunk [base, base + 64) of a candidate's sum the way a warp computes it —
// lane l sums elements base + l and base + 32 + l, then a xor-butterfly over the lanes — but
// evaluated by ONE lane (the butterfly's additions replayed in the same order: after the step with
// offset o the lower o slots hold what lanes 0 .. o-1 hold).  Same bits as the warp version.
__device__ __forceinline__ double scf_chunk_one_lane(const double *terms, int base, int n, const ScfDir &d) {
  double p[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) {
    double c = 0.0;
    if (base + l < n) c = scf_add_term(d, terms + 9 * (base + l), c);
    if (base + 32 + l < n) c = scf_add_term(d, terms + 9 * (base + 32 + l), c);
    p[l] = c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int l = 0; l < o; ++l) p[l] += p[l + o];
  return p[0];
}

// ---- the contraction-sensitive arithmetic of a pair, compiled ONCE
//
// scf_pair is instantiated in several kernels (4-warp pass 1, 16-warp pass 2, the fused rounds kernel
// of small batches).  Plain expressions such as a*b + c*d may be contracted into different fma's in
// different instantiations, which changes last bits; the frame solve relies on a pair giving the SAME
// bits whichever kernel processes it (scan cache, fixed points, two passes).  Everything that is not
// already spelled out with explicit roundings (scf_add_term) therefore lives in these three
// out-of-line functions: one body, one SASS, called from every kernel.  Each is called once per
// thread and phase (not per correspondence), so the call costs nothing measurable.

// terms[9 i .. 9 i + 8] for the correspondences tid, tid + nt, ... of a pair
__device__ __noinline__ void scf_fill_terms(const double *f1, const double *f2, const double *ct, long long s, int n,
                                            int tid, int nt, const double *pose, double reg, double *terms) {
  double R[9];
  pose_rotation(pose, R);
  for (int i = tid; i < n; i += nt) {
    double a1[3], a2[3], c9[9], w[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) { a1[k] = f1[3 * (s + i) + k]; a2[k] = f2[3 * (s + i) + k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) c9[k] = ct[9 * (s + i) + k];
    scf_terms(R, reg, a1, a2, c9, w);
#pragma unroll
    for (int k = 0; k < 9; ++k) terms[9 * i + k] = w[k];
  }
}

// this thread's share of E(t) = sum_i n_i n_i^T / (t^T B_i t): correspondences tid, tid + nta, ...
__device__ __noinline__ void scf_step_partial(const double *terms, int n, int tid, int nta, double t0, double t1,
                                              double t2, double *E /* [6] */) {
  const double txx = t0 * t0, txy = 2.0 * t0 * t1, txz = 2.0 * t0 * t2;
  const double tyy = t1 * t1, tyz = 2.0 * t1 * t2, tzz = t2 * t2;
  double e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0, e5 = 0;
  for (int i = tid; i < n; i += nta) {
    const double *w = terms + 9 * i;
    const double den = w[3] * txx + w[4] * txy + w[5] * txz + w[6] * tyy + w[7] * tyz + w[8] * tzz;
    const double inv = fast_rcp(den);
    e0 = fma(w[0] * w[0], inv, e0); e1 = fma(w[0] * w[1], inv, e1); e2 = fma(w[0] * w[2], inv, e2);
    e3 = fma(w[1] * w[1], inv, e3); e4 = fma(w[1] * w[2], inv, e4); e5 = fma(w[2] * w[2], inv, e5);
  }
  E[0] = e0; E[1] = e1; E[2] = e2; E[3] = e3; E[4] = e4; E[5] = e5;
}

__device__ __noinline__ void scf_eigvec(const double *Es /* [6] */, double *v /* [3] */) {
  double m[6] = {Es[0], Es[1], Es[2], Es[3], Es[4], Es[5]}, vv[3], lam;
  sym3_smallest_eigvec_fast(m, vv, lam);
  v[0] = vv[0]; v[1] = vv[1]; v[2] = vv[2];
}

template <int NW, bool LANE_SERIAL = false>
__device__ __forceinline__ void scf_pair(const ScfArgs &args, const long long b) {
  constexpr int NT = NW * 32;
  // Sums whose value depends on how they are split over threads (cost0, E of the SCF steps) are
  // always split over the first NWA = 4 warps, so that the 4-warp and the 16-warp launch of the
  // same pair produce the same bits; the extra warps only serve the scan.
  constexpr int NWA = NW < 4 ? NW : 4;
  constexpr int NTA = NWA * 32;
  __shared__ double s_red[NW][6];
  __shared__ double s_best_cost[NW];
  __shared__ int s_best_idx[NW];
  __shared__ double s_t[3];
  __shared__ unsigned long long s_bound;  // bits of the running upper bound (costs are >= 0: ordered as integers)
  __shared__ int s_nsurv;
  __shared__ int s_surv_idx[kScfMaxSurvivors];
  __shared__ double s_surv_part[kScfMaxSurvivors];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long s, e;
  problem_range(args.bv, b, s, e);
  const int n = static_cast<int>(e - s);
  const double *pose = args.bv.poses + 7 * b;
  // [n][9]: shared memory, or this pair's slice of the spill array when it does not fit
  double *terms = (n <= args.cap_elems || !args.spill) ? dyn_smem : args.spill + 9 * s;
  double *ot = args.out_t + static_cast<long long>(args.out_stride) * b;
  const long long clk0 = clock64();
  const double *tsrc = args.prev_poses ? args.prev_poses + 7 * b + 4 : pose + 4;  // start translation
  if (args.fixed && args.fixed[b]) {  // fixed point of the iteration: the result is the previous round's
    if (tid == 0) {
      if (args.prev_poses) { ot[0] = tsrc[0]; ot[1] = tsrc[1]; ot[2] = tsrc[2]; }
      if (args.dbg) { args.dbg[4 * b] = 0; args.dbg[4 * b + 1] = 0; args.dbg[4 * b + 2] = 0; args.dbg[4 * b + 3] = 0; }
    }
    return;
  }
  if (n <= 0 || (n > args.cap_elems && !args.spill)) {
    // nothing to minimise (a pair beyond the shared-memory capacity without a spill array is rejected on the host)
    if (tid == 0) { ot[0] = tsrc[0]; ot[1] = tsrc[1]; ot[2] = tsrc[2]; if (args.out_cost) args.out_cost[b] = 0.0; }
    return;
  }
  scf_fill_terms(args.bv.f1, args.bv.f2, args.bv.ct, s, n, tid, NT, pose, args.reg, terms);
  __syncthreads();

  // ---- scan: candidate 0 is the given translation, 1..samples the Fibonacci sphere; the first
  // strict minimum wins (pnec.cc:332-340).
  //
  // Every term (t.n_i)^2 / (t^T B_i t) is non-negative (B_i is positive definite: reg > 0), so a
  // partial sum is a lower bound of a candidate's cost and a candidate whose partial sum already
  // EXCEEDS the cost of a fully evaluated one can never be the minimum: branch and bound, exact.
  //   0. cost0 = cost of the given translation (whole CTA), the first bound;
  //   1. every sphere candidate is summed over the first kScfPrefix correspondences; those not
  //      above cost0 survive (typically none or a handful: the sphere is 0.16 rad coarse);
  //   2. survivors are completed one per warp, tightening the bound as they finish.
  // The set of completed candidates depends on timing, the minimum and its index do not: pruning
  // is strict (ties are always evaluated) and every bound is the complete cost of a real candidate.
  const double t0[3] = {tsrc[0], tsrc[1], tsrc[2]};
  double cost0;
  {
    double part = 0.0;
    if (warp < NWA) {
      const ScfDir d0 = scf_dir(t0);
      for (int i = tid; i < n; i += NTA) part = scf_add_term(d0, terms + 9 * i, part);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) s_best_cost[warp] = part;
    }
    if (tid == 0) s_nsurv = 0;
    __syncthreads();
    cost0 = 0.0;
#pragma unroll
    for (int w = 0; w < NWA; ++w) cost0 += s_best_cost[w];
    __syncthreads();
  }
  // what is known about the sphere: minimum `sph_value` at point `sph_idx` (> 0), or only the lower
  // bound `sph_value` (sph_idx == 0)
  double sph_value = CUDART_INF;
  int sph_idx = 0;
  bool rescan = true;
  if (args.cache) {
    const ScfScanCache &c = args.cache[b];
    const bool same = c.valid && __double_as_longlong(c.q[0]) == __double_as_longlong(pose[0]) &&
                      __double_as_longlong(c.q[1]) == __double_as_longlong(pose[1]) &&
                      __double_as_longlong(c.q[2]) == __double_as_longlong(pose[2]) &&
                      __double_as_longlong(c.q[3]) == __double_as_longlong(pose[3]);
    // a lower bound only settles the comparison when it is not below the new cost0
    if (same && (c.idx > 0 || c.value >= cost0)) {
      rescan = false;
      sph_value = c.value;
      sph_idx = c.idx;
    }
  }
  if (rescan) {  // block-uniform
    const bool prune = cost0 >= 0.0 && cost0 < CUDART_INF && args.reg > 0.0 && n > kScfPrefix;
    const int npre = prune ? kScfPrefix : n;
    if (tid == 0) s_bound = static_cast<unsigned long long>(__double_as_longlong(prune ? cost0 : CUDART_INF));
    double best = CUDART_INF;  // complete costs seen by this thread / warp
    int best_idx = 0x7fffffff;
    for (int c = 1 + tid; c <= args.samples; c += NT) {
      const double t[3] = {args.sphere[3 * (c - 1)], args.sphere[3 * (c - 1) + 1], args.sphere[3 * (c - 1) + 2]};
      // with pruning on, a candidate is dropped at the first checkpoint above cost0 (most are after
      // four terms: the sphere is coarse); survivors carry the full 32-term prefix
      const double part = prune ? scf_objective_bounded(terms, npre, t, cost0) : scf_objective(terms, npre, t);
      if (npre == n) {
        if (part < best || (part == best && c < best_idx)) { best = part; best_idx = c; }
      } else if (!(part > cost0)) {
        const int slot = atomicAdd(&s_nsurv, 1);
        if (slot < kScfMaxSurvivors) { s_surv_idx[slot] = c; s_surv_part[slot] = part; }
      }
    }
    __syncthreads();
    const int nsurv = s_nsurv;
    if (args.defer_list && nsurv > args.defer_threshold) {  // block-uniform: nothing has been written yet
      if (tid == 0) args.defer_list[atomicAdd(args.defer_count, 1)] = static_cast<int>(b);
      return;
    }
    if (nsurv > kScfMaxSurvivors) {
      // more survivors than the list holds (a poor start translation): plain scan of the rest
      if (tid == 0) s_bound = static_cast<unsigned long long>(__double_as_longlong(CUDART_INF));
      best = CUDART_INF; best_idx = 0x7fffffff;
      for (int c = 1 + tid; c <= args.samples; c += NT) {
        const double t[3] = {args.sphere[3 * (c - 1)], args.sphere[3 * (c - 1) + 1], args.sphere[3 * (c - 1) + 2]};
        const double cost = scf_objective(terms, npre, t) + scf_objective_tail(terms, npre, n, t);
        if (cost < best || (cost == best && c < best_idx)) { best = cost; best_idx = c; }
      }
    } else {
      if (LANE_SERIAL && nsurv > 8 * NW) {  // compile-time: only the pass-2 kernel carries this path
        // many survivors (a flat cost landscape): one candidate per LANE, no shuffles; every chunk is
        // summed in the order a warp would use (scf_chunk_one_lane), so a candidate's cost does not
        // depend on which of the two paths completed it
        for (int sv = tid; sv < nsurv; sv += NT) {
          const int c = s_surv_idx[sv];
          double cost = s_surv_part[sv];
          const double t[3] = {args.sphere[3 * (c - 1)], args.sphere[3 * (c - 1) + 1], args.sphere[3 * (c - 1) + 2]};
          const ScfDir d = scf_dir(t);
          bool alive = !(cost > __longlong_as_double(static_cast<long long>(*reinterpret_cast<volatile unsigned long long *>(&s_bound))));
          for (int base = npre; alive && base < n; base += 64) {
            cost += scf_chunk_one_lane(terms, base, n, d);
            alive = !(cost > __longlong_as_double(static_cast<long long>(*reinterpret_cast<volatile unsigned long long *>(&s_bound))));
          }
          if (!alive) continue;
          if (cost < best || (cost == best && c < best_idx)) { best = cost; best_idx = c; }
          atomicMin(&s_bound, static_cast<unsigned long long>(__double_as_longlong(cost)));
        }
      } else {
      for (int sv = warp; sv < nsurv; sv += NW) {  // warp-uniform
        const int c = s_surv_idx[sv];
        const double part = s_surv_part[sv];
        unsigned long long bits = 0;
        if (lane == 0) bits = *reinterpret_cast<volatile unsigned long long *>(&s_bound);
        bits = __shfl_sync(0xffffffffu, bits, 0);  // one read per warp: the decision is warp-uniform
        if (part > __longlong_as_double(static_cast<long long>(bits))) continue;
        const double t[3] = {args.sphere[3 * (c - 1)], args.sphere[3 * (c - 1) + 1], args.sphere[3 * (c - 1) + 2]};
        // the rest of the sum by the whole warp, 64 correspondences at a time, re-checking the
        // bound after every chunk: cost = part + chunk sums in order (fixed, timing-independent)
        const ScfDir d = scf_dir(t);
        double cost = part;
        bool alive = true;
        for (int base = npre; base < n; base += 64) {
          double chunk = 0.0;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int i = base + 32 * k + lane;
            if (i < n) chunk = scf_add_term(d, terms + 9 * i, chunk);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) chunk += __shfl_xor_sync(0xffffffffu, chunk, o);
          cost += chunk;
          if (lane == 0) bits = *reinterpret_cast<volatile unsigned long long *>(&s_bound);
          bits = __shfl_sync(0xffffffffu, bits, 0);
          if (cost > __longlong_as_double(static_cast<long long>(bits))) { alive = false; break; }
        }
        if (!alive) continue;
        if (cost < best || (cost == best && c < best_idx)) { best = cost; best_idx = c; }
        if (lane == 0) atomicMin(&s_bound, static_cast<unsigned long long>(__double_as_longlong(cost)));
      }
      }
    }
    // every lane of a warp holds the same (best, best_idx) in the survivor path, its own in the others
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double oc = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
      if (oc < best || (oc == best && oi < best_idx)) { best = oc; best_idx = oi; }
    }
    if (lane == 0) { s_best_cost[warp] = best; s_best_idx[warp] = best_idx; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < NW; ++w)
      if (w != warp && (s_best_cost[w] < best || (s_best_cost[w] == best && s_best_idx[w] < best_idx))) {
        best = s_best_cost[w]; best_idx = s_best_idx[w];
      }
    // pruned candidates cost more than min(cost0, best): `best` is the sphere minimum if it is below
    // cost0, otherwise all that is known is that the sphere does not go below cost0
    if (best_idx != 0x7fffffff && (best < cost0 || !prune)) { sph_value = best; sph_idx = best_idx; }
    else { sph_value = cost0; sph_idx = 0; }
    if (tid == 0 && args.cache) {
      ScfScanCache &c = args.cache[b];
      c.q[0] = pose[0]; c.q[1] = pose[1]; c.q[2] = pose[2]; c.q[3] = pose[3];
      c.value = sph_value; c.idx = sph_idx; c.valid = 1;
    }
  }
  const long long clk1 = clock64();
  const int dbg_nsurv = rescan ? s_nsurv : 0;
  if (tid == 0) {
    if (sph_idx > 0 && sph_value < cost0) {
      s_t[0] = args.sphere[3 * (sph_idx - 1)]; s_t[1] = args.sphere[3 * (sph_idx - 1) + 1]; s_t[2] = args.sphere[3 * (sph_idx - 1) + 2];
    } else {
      s_t[0] = t0[0]; s_t[1] = t0[1]; s_t[2] = t0[2];
    }
  }
  __syncthreads();

  // ---- SCF: t <- eigenvector of the smallest eigenvalue of E(t) = sum_i A_i / (t^T B_i t)
  // (alt_construct_E with frac[i] == 0, scf.cc:109-126: `frac.resize(n)` then push_back).
  // E is even in t and the step is a deterministic map, so once a step returns +-(its input) every
  // later step returns the same vector, and once it returns +-(the input of the step before) the
  // sequence alternates between two vectors: the loop stops there with exactly the vector the
  // remaining steps would have produced.
  {
    double prev[3] = {0, 0, 0};  // input of the previous step (tid 0)
    bool have_prev = false;
    for (int it = 0; it < args.steps; ++it) {
      const double t[3] = {s_t[0], s_t[1], s_t[2]};
      if (warp < NWA) {
        double E[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        scf_step_partial(terms, n, tid, NTA, t[0], t[1], t[2], E);
        // transposing warp reduction: lane L ends with the warp sum of E[4 b4 + 2 b3 + b2]
        exchange_step<8, 16>(E, lane);
        exchange_step<4, 8>(E, lane);
        exchange_step<2, 4>(E, lane);
        double sum = E[0];
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        const int eidx = 4 * ((lane >> 4) & 1) + 2 * ((lane >> 3) & 1) + ((lane >> 2) & 1);
        if ((lane & 3) == 0 && eidx < 6) s_red[warp][eidx] = sum;
      }
      __syncthreads();
      if (tid == 0) {
        double Es[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          double a = 0.0;
#pragma unroll
          for (int w = 0; w < NWA; ++w) a += s_red[w][k];
          Es[k] = a;
        }
        double v[3];
        scf_eigvec(Es, v);
        auto same_up_to_sign = [](const double a[3], const double c[3]) {
          const bool p = __double_as_longlong(a[0]) == __double_as_longlong(c[0]) &&
                         __double_as_longlong(a[1]) == __double_as_longlong(c[1]) &&
                         __double_as_longlong(a[2]) == __double_as_longlong(c[2]);
          const bool m = __double_as_longlong(a[0]) == __double_as_longlong(-c[0]) &&
                         __double_as_longlong(a[1]) == __double_as_longlong(-c[1]) &&
                         __double_as_longlong(a[2]) == __double_as_longlong(-c[2]);
          return p || m;
        };
        int stop = 0;
        const int remaining = args.steps - 1 - it;
        if (same_up_to_sign(v, t)) {
          stop = 1;  // fixed point: every further step returns v
        } else if (have_prev && same_up_to_sign(v, prev)) {
          stop = 1;  // period two: v, t', v, t', ... with t' = step(v) = the current input's successor
          if (remaining & 1) { v[0] = t[0]; v[1] = t[1]; v[2] = t[2]; }
        }
        // the successor of +-prev is t (as computed one step ago), so an odd number of remaining
        // steps ends on t and an even number on v
        prev[0] = t[0]; prev[1] = t[1]; prev[2] = t[2];
        have_prev = true;
        s_t[0] = v[0]; s_t[1] = v[1]; s_t[2] = v[2];
        s_nsurv = stop;  // reuse as the stop flag (the scan is over)
      }
      __syncthreads();
      if (s_nsurv) break;
    }
  }
  if (tid == 0) {
    bool q_same = args.q_same && args.q_same[b];
    if (args.prev_poses) {
      const double *pq = args.prev_poses + 7 * b;
      q_same = __double_as_longlong(pq[0]) == __double_as_longlong(pose[0]) &&
               __double_as_longlong(pq[1]) == __double_as_longlong(pose[1]) &&
               __double_as_longlong(pq[2]) == __double_as_longlong(pose[2]) &&
               __double_as_longlong(pq[3]) == __double_as_longlong(pose[3]);
    }
    if (args.fixed && q_same &&
        __double_as_longlong(s_t[0]) == __double_as_longlong(t0[0]) &&
        __double_as_longlong(s_t[1]) == __double_as_longlong(t0[1]) &&
        __double_as_longlong(s_t[2]) == __double_as_longlong(t0[2]))
      args.fixed[b] = 1;  // (rotation, translation) -> itself: every further round repeats it
    if (args.dbg) {
      args.dbg[4 * b] = rescan ? 2 : 1;
      args.dbg[4 * b + 1] = dbg_nsurv;
      args.dbg[4 * b + 2] = clk1 - clk0;
      args.dbg[4 * b + 3] = clock64() - clk0;
    }
    ot[0] = s_t[0]; ot[1] = s_t[1]; ot[2] = s_t[2];
    if (args.out_cost) {
      const double t[3] = {s_t[0], s_t[1], s_t[2]};
      args.out_cost[b] = scf_objective(terms, n, t);
    }
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) scf_kernel(const __grid_constant__ ScfArgs args) {
  scf_pair<NW>(args, blockIdx.x);
}

// pass 2: persistent CTAs over the deferred pairs
template <int NW>
__global__ void __launch_bounds__(NW * 32) scf_list_kernel(const __grid_constant__ ScfArgs args) {
  __shared__ int s_work;
  const int count = *args.work_count;
  for (;;) {
    __syncthreads();  // the previous pair is finished with the shared memory (and s_work)
    if (threadIdx.x == 0) s_work = atomicAdd(args.work_cursor, 1);
    __syncthreads();
    const int k = s_work;
    if (k >= count) return;
    scf_pair<NW, true>(args, args.work_list[k]);
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

// The same for large batches: a warp per pair, four independent pairs per CTA.  With a CTA per pair
// 127 threads wait while one solves the 3x3 eigenproblem, and the loads of a pair are one round trip
// per stride (3.0 TB/s at 10 000 x 512).  Here a lane plays the four threads lane, lane + 32, lane + 64,
// lane + 96 of the CTA-per-pair kernel -- four accumulator sets, 24 loads in flight -- and the sums are
// combined exactly as block_sum6<4> combines them (butterfly per set, then sets 0..3 in order), so both
// kernels give the same bits; no CTA-wide barrier, a warp's serial tail costs 31 lanes, not 127.
__global__ void __launch_bounds__(128) nec_translation_warp_kernel(BatchView bv, double *out_t, int out_stride,
                                                                    double *out_M) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long b = static_cast<long long>(blockIdx.x) * 4 + warp;
  if (b >= bv.num_problems) return;
  long long s, e;
  problem_range(bv, b, s, e);
  double R[9];
  pose_rotation(bv.poses + 7 * b, R);
  double M[4][6];
#pragma unroll
  for (int w = 0; w < 4; ++w)
#pragma unroll
    for (int k = 0; k < 6; ++k) M[w][k] = 0.0;
  for (long long i0 = s + 1 + lane; i0 < e; i0 += 128) {
    double f1[4][3], f2[4][3];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const long long i = i0 + 32 * w;
      if (i < e) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { f1[w][k] = bv.f1[3 * i + k]; f2[w][k] = bv.f2[3 * i + k]; }
      }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      if (i0 + 32 * w < e) {
        double g[3], n[3];
        rot(R, f2[w], g);
        cross3(f1[w], g, n);
        M[w][0] = fma(n[0], n[0], M[w][0]); M[w][1] = fma(n[0], n[1], M[w][1]); M[w][2] = fma(n[0], n[2], M[w][2]);
        M[w][3] = fma(n[1], n[1], M[w][3]); M[w][4] = fma(n[1], n[2], M[w][4]); M[w][5] = fma(n[2], n[2], M[w][5]);
      }
    }
  }
  double T[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      double v = M[w][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      t += v;
    }
    T[k] = t;
  }
  if (lane == 0) {
    double v[3], lam;
    sym3_smallest_eigvec(T, v, lam);
    double *ot = out_t + static_cast<long long>(out_stride) * b;
    ot[0] = v[0]; ot[1] = v[1]; ot[2] = v[2];
    if (out_M) {
#pragma unroll
      for (int k = 0; k < 6; ++k) out_M[6 * b + k] = T[k];
    }
  }
}

}  // namespace pnec
