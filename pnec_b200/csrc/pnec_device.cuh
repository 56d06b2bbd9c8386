// pnec_device.cuh — device-side arithmetic of the PNEC frame-pair solver (sm_100a).
//
// Closed forms of what the reference evaluates numerically:
//   residuals   include/optimization/pnec_residual.h:50-150, nec_residual.h:47-69
//   Jacobian    ceres::NumericDiffCostFunction<.., CENTRAL, 1,1,1,4> composed with
//               ceres::EigenQuaternionManifold (src/optimization/pnec_ceres.cc:92-106)
//   LM          ceres::Solve with default options (src/optimization/pnec_ceres.cc:110)
// Tangent coordinates x = (theta, phi, d1, d2, d3): t = t(theta, phi) as in
// pnec_residual.h:89-91 and q <- [sin|d| d/|d|, cos|d|] (x) q, i.e. R <- Exp(2d) R.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/pnec_b200.h"

namespace pnec {

constexpr int kNumAcc = 21;  // 15 (JtJ upper) + 5 (Jtr) + 1 (sum r^2)
constexpr int kAccPad = 24;  // padded for the transposing warp reduction

// ------------------------------------------------------------------ small math

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double a[3], const double b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
// o = R v, R row-major
__device__ __forceinline__ void rot(const double R[9], const double v[3], double o[3]) {
  o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
// o = R^T v
__device__ __forceinline__ void rot_t(const double R[9], const double v[3], double o[3]) {
  o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
// symmetric part of a column-major 3x3 times v.  x^T S x only sees sym(S), and so
// does its derivative, so this is exact for whatever the caller stored in S.
__device__ __forceinline__ void sym_mul(const double c[9], const double v[3], double o[3]) {
  const double xy = 0.5 * (c[1] + c[3]);
  const double xz = 0.5 * (c[2] + c[6]);
  const double yz = 0.5 * (c[5] + c[7]);
  o[0] = c[0] * v[0] + xy * v[1] + xz * v[2];
  o[1] = xy * v[0] + c[4] * v[1] + yz * v[2];
  o[2] = xz * v[0] + yz * v[1] + c[8] * v[2];
}

// Packed symmetric part (xx, xy, xz, yy, yz, zz) of a column-major 3x3.
__device__ __forceinline__ void pack_sym(const double c[9], double s[6]) {
  s[0] = c[0];
  s[1] = 0.5 * (c[1] + c[3]);
  s[2] = 0.5 * (c[2] + c[6]);
  s[3] = c[4];
  s[4] = 0.5 * (c[5] + c[7]);
  s[5] = c[8];
}
__device__ __forceinline__ void sym_mul6(const double s[6], const double v[3], double o[3]) {
  o[0] = s[0] * v[0] + s[1] * v[1] + s[2] * v[2];
  o[1] = s[1] * v[0] + s[3] * v[1] + s[4] * v[2];
  o[2] = s[2] * v[0] + s[4] * v[1] + s[5] * v[2];
}

// Per-problem constants of one evaluation point.
struct PoseConst {
  double R[9];    // Eigen::Quaternion::toRotationMatrix(), q taken as stored
  double t[3];    // t(theta, phi)
  double tth[3];  // dt/dtheta
  double tph[3];  // dt/dphi
};

// x = (theta, phi, qx, qy, qz, qw)
__device__ __forceinline__ void make_pose_const(const double x[6], PoseConst &pc) {
  double st, ct, sp, cp;
  sincos(x[0], &st, &ct);
  sincos(x[1], &sp, &cp);
  pc.t[0] = st * cp;  pc.t[1] = st * sp;  pc.t[2] = ct;
  pc.tth[0] = ct * cp; pc.tth[1] = ct * sp; pc.tth[2] = -st;
  pc.tph[0] = -st * sp; pc.tph[1] = st * cp; pc.tph[2] = 0.0;
  const double qx = x[2], qy = x[3], qz = x[4], qw = x[5];
  const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  pc.R[0] = 1.0 - (tyy + tzz); pc.R[1] = txy - twz;         pc.R[2] = txz + twy;
  pc.R[3] = txy + twz;         pc.R[4] = 1.0 - (txx + tzz); pc.R[5] = tyz - twx;
  pc.R[6] = txz - twy;         pc.R[7] = tyz + twx;         pc.R[8] = 1.0 - (txx + tyy);
}

// pnec::common::AnglesFromVec, src/common/common.cc:103-116
__device__ __forceinline__ void angles_from_vec(const double v[3], double &theta, double &phi) {
  const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (n == 0.0) {
    theta = 0.0;
    phi = 0.0;
    return;
  }
  theta = acos(v[2] / n);
  phi = (fabs(theta) < 1e-10) ? 0.0 : atan2(v[1] / n, v[0] / n);
}

// ceres::EigenQuaternionManifold::Plus on the quaternion block, Euclidean on theta, phi.
__device__ __forceinline__ void state_plus(const double x[6], const double d[5], double out[6]) {
  out[0] = x[0] + d[0];
  out[1] = x[1] + d[1];
  const double nd = sqrt(d[2] * d[2] + d[3] * d[3] + d[4] * d[4]);
  if (nd == 0.0) {
    out[2] = x[2]; out[3] = x[3]; out[4] = x[4]; out[5] = x[5];
    return;
  }
  double sn, cs;
  sincos(nd, &sn, &cs);
  const double s = sn / nd;
  const double dw = cs, dx = s * d[2], dy = s * d[3], dz = s * d[4];
  const double xx = x[2], xy = x[3], xz = x[4], xw = x[5];
  out[5] = dw * xw - dx * xx - dy * xy - dz * xz;
  out[2] = dw * xx + dx * xw + dy * xz - dz * xy;
  out[3] = dw * xy - dx * xz + dy * xw + dz * xx;
  out[4] = dw * xz + dx * xy - dy * xx + dz * xw;
}

// -------------------------------------------------- residual + Jacobian row
//
// With g = R f2, a = t x f1, e = a . g (numerator of every variant):
//   de/dt = f1 x g                   de/dw = g x a      (w: full-angle left perturbation)
// TARGET     s2 = b^T S b + reg,  b = R^T a
//            (1/2) ds2/dt = f1 x (R S b)      (1/2) ds2/dw = (R S b) x a
// HOST       s2 = c^T S c + reg,  c = t x h,  h = R f1
// SYMMETRIC  s2 = b^T S2 b + c^T S1 c + reg,  c = t x h,  h = g
//            (1/2) d(c^T S c)/dt = h x (S c)  (1/2) d(c^T S c)/dw = (t.h) S c - (S c . h) t
//   r = e / s,   dr = (de - (e / s2) (1/2) ds2) / s
// Covariances arrive as their packed symmetric part (pack_sym): x^T S x only sees sym(S).
// The row returned is (dr/dtheta, dr/dphi, dr/dw) — the quaternion tangent
// columns are 2 dr/dw; the factor is applied once after the reduction.
template <int V>
__device__ __forceinline__ void residual_row(const PoseConst &pc, double reg, const double f1[3],
                                             const double f2[3], const double ct[6],
                                             const double ch[6], double &r, double row[5]) {
  double g[3], a[3];
  rot(pc.R, f2, g);
  cross3(pc.t, f1, a);
  const double e = dot3(a, g);
  double drdt[3], drdw[3];
  if (V == PNEC_VARIANT_NEC) {
    r = e;
    cross3(f1, g, drdt);
    cross3(g, a, drdw);
  } else if (V == PNEC_VARIANT_TARGET) {
    double b[3], Sb[3], RSb[3], p[3];
    rot_t(pc.R, a, b);
    sym_mul6(ct, b, Sb);
    const double s2 = dot3(b, Sb) + reg;
    const double is = rsqrt(s2);
    r = e * is;
    const double k = r * is;  // e / s2
    rot(pc.R, Sb, RSb);
    p[0] = is * (g[0] - k * RSb[0]);
    p[1] = is * (g[1] - k * RSb[1]);
    p[2] = is * (g[2] - k * RSb[2]);
    cross3(f1, p, drdt);
    cross3(p, a, drdw);
  } else {
    double s2 = reg;
    double hdt[3] = {0.0, 0.0, 0.0}, hdw[3] = {0.0, 0.0, 0.0};  // (1/2) ds2
    if (V == PNEC_VARIANT_SYMMETRIC) {
      double b[3], Sb[3], RSb[3];
      rot_t(pc.R, a, b);
      sym_mul6(ct, b, Sb);
      s2 += dot3(b, Sb);
      rot(pc.R, Sb, RSb);
      cross3(f1, RSb, hdt);
      cross3(RSb, a, hdw);
    }
    double h[3], c[3], Sc[3], hx[3];
    if (V == PNEC_VARIANT_HOST) {
      rot(pc.R, f1, h);
    } else {
      h[0] = g[0]; h[1] = g[1]; h[2] = g[2];
    }
    cross3(pc.t, h, c);
    sym_mul6(V == PNEC_VARIANT_HOST ? ct : ch, c, Sc);
    s2 += dot3(c, Sc);
    cross3(h, Sc, hx);
    const double th = dot3(pc.t, h), sh = dot3(Sc, h);
    hdt[0] += hx[0]; hdt[1] += hx[1]; hdt[2] += hx[2];
    hdw[0] += th * Sc[0] - sh * pc.t[0];
    hdw[1] += th * Sc[1] - sh * pc.t[1];
    hdw[2] += th * Sc[2] - sh * pc.t[2];
    const double is = rsqrt(s2);
    r = e * is;
    const double k = r * is;
    double det[3], dew[3];
    cross3(f1, g, det);
    cross3(g, a, dew);
    drdt[0] = is * (det[0] - k * hdt[0]);
    drdt[1] = is * (det[1] - k * hdt[1]);
    drdt[2] = is * (det[2] - k * hdt[2]);
    drdw[0] = is * (dew[0] - k * hdw[0]);
    drdw[1] = is * (dew[1] - k * hdw[1]);
    drdw[2] = is * (dew[2] - k * hdw[2]);
  }
  row[0] = dot3(drdt, pc.tth);
  row[1] = drdt[0] * pc.tph[0] + drdt[1] * pc.tph[1];  // tph[2] == 0
  row[2] = drdw[0];
  row[3] = drdw[1];
  row[4] = drdw[2];
}

// Residual only (cost passes of the LM loop): same expressions as residual_row.
template <int V>
__device__ __forceinline__ double residual_only(const PoseConst &pc, double reg, const double f1[3],
                                                const double f2[3], const double ct[6],
                                                const double ch[6]) {
  double g[3], a[3];
  rot(pc.R, f2, g);
  cross3(pc.t, f1, a);
  const double e = dot3(a, g);
  if (V == PNEC_VARIANT_NEC) return e;
  double s2 = reg;
  if (V == PNEC_VARIANT_TARGET || V == PNEC_VARIANT_SYMMETRIC) {
    double b[3], Sb[3];
    rot_t(pc.R, a, b);
    sym_mul6(ct, b, Sb);
    s2 = dot3(b, Sb) + reg;
  }
  if (V == PNEC_VARIANT_HOST || V == PNEC_VARIANT_SYMMETRIC) {
    double h[3], c[3], Sc[3];
    if (V == PNEC_VARIANT_HOST) {
      rot(pc.R, f1, h);
    } else {
      h[0] = g[0]; h[1] = g[1]; h[2] = g[2];
    }
    cross3(pc.t, h, c);
    sym_mul6(V == PNEC_VARIANT_HOST ? ct : ch, c, Sc);
    s2 += dot3(c, Sc);
  }
  return e * rsqrt(s2);
}

// acc += (J^T J upper, J^T r, r^2)
__device__ __forceinline__ void accumulate(double acc[kNumAcc], double r, const double row[5]) {
  int k = 0;
#pragma unroll
  for (int a = 0; a < 5; ++a) {
#pragma unroll
    for (int b = a; b < 5; ++b) {
      acc[k] = fma(row[a], row[b], acc[k]);
      ++k;
    }
  }
#pragma unroll
  for (int a = 0; a < 5; ++a) acc[15 + a] = fma(row[a], r, acc[15 + a]);
  acc[20] = fma(r, r, acc[20]);
}

// The quaternion tangent columns carry a factor 2 (delta is a half angle); apply
// it to the reduced sums: H[tp, d] *= 2, H[d, d] *= 4, g[d] *= 2, cost = sum r^2 / 2.
__device__ __forceinline__ double acc_scale(int idx) {
  // idx in the packed order (0,0..4),(1,1..4),(2,2..4),(3,3..4),(4,4), g0..g4, cost
  switch (idx) {
    case 0: case 1: case 5: return 1.0;               // (0,0) (0,1) (1,1)
    case 2: case 3: case 4: case 6: case 7: case 8: return 2.0;  // (0|1, d)
    case 9: case 10: case 11: case 12: case 13: case 14: return 4.0;  // (d, d)
    case 15: case 16: return 1.0;
    case 17: case 18: case 19: return 2.0;
    default: return 0.5;  // 20: cost
  }
}

// ----------------------------------------------------------- warp reduction
//
// Transposing reduction: 24 values per lane in, after 24 exchange steps (12+6+3+2+1)
// lane L owns the warp-wide sum of value index 12 b4 + 6 b3 + 3 b2 + 2 b1 + b0
// (lanes with b1 = b0 = 1 own nothing).  Fixed order => deterministic.
__device__ __forceinline__ int warp_reduce_owner_index(int lane) {
  const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1;
  const int low = lane & 3;
  return (low == 3) ? -1 : 12 * b4 + 6 * b3 + 3 * b2 + low;
}

template <int COUNT, int OFFSET>
__device__ __forceinline__ void exchange_step(double *v, int lane) {
  constexpr int HALF = COUNT / 2;
  const bool upper = (lane & OFFSET) != 0;
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    const double send = upper ? v[i] : v[i + HALF];
    const double keep = upper ? v[i + HALF] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFFSET);
  }
}

__device__ __forceinline__ double warp_transpose_reduce(const double acc[kNumAcc], int lane) {
  double v[kAccPad];
#pragma unroll
  for (int i = 0; i < kNumAcc; ++i) v[i] = acc[i];
#pragma unroll
  for (int i = kNumAcc; i < kAccPad; ++i) v[i] = 0.0;
  exchange_step<24, 16>(v, lane);
  exchange_step<12, 8>(v, lane);
  exchange_step<6, 4>(v, lane);
  v[3] = 0.0;
  exchange_step<4, 2>(v, lane);
  exchange_step<2, 1>(v, lane);
  return v[0];
}

// ------------------------------------------------ mbarrier / bulk-copy (TMA) PTX

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared through the TMA engine (UBLKCP in SASS).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// 1-D bulk async copy shared -> global (TMA engine), tracked by bulk async-groups.
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until the shared-memory sources of all committed bulk stores have been read
__device__ __forceinline__ void bulk_wait_group_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

}  // namespace pnec
