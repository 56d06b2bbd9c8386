// pnec_compat.hpp — the reference's C++ pose-optimization API over the B200 C-ABI.
//
// Same namespaces, class names, method names and argument order as
//   pnec::optimization::PNECCeres          include/optimization/pnec_ceres.h:48-89
//   pnec::optimization::NECCeres           include/optimization/nec_ceres.h:46-78
//   pnec::rel_pose_estimation::Options     include/rel_pose_estimation/pnec_config.h:46-65
//   pnec::rel_pose_estimation::PNEC        include/rel_pose_estimation/pnec.h:48-114
//   pnec::common::{NoiseFrame, AnglesFromVec, SkewFromVector, RotationalDifference,
//                  TranslationalDifference, CostFunction}   include/common/common.h:57-120
// so a call site of the reference compiles against this header after swapping the
// include.  Eigen / Sophus / opengv are not required: every container argument is a
// template that only needs .data()/.size() over 24-byte vectors (Eigen::Vector3d,
// opengv::bearingVector_t, pnec::Vec3) or 72-byte column-major matrices
// (Eigen::Matrix3d, pnec::Mat3), which is the memory the reference already holds.
// Poses are pnec::SE3 = {unit quaternion x,y,z,w ; translation}, the memory order of
// Sophus::SE3d; anything with .unit_quaternion()/.translation() or
// .rotationMatrix()/.translation() converts through the templated constructors.
//
// Header-only; link with libpnec_b200.so.
#ifndef PNEC_COMPAT_HPP_
#define PNEC_COMPAT_HPP_

#include <array>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../pnec_b200.h"

// Interop with the reference's own pose / vector types when their headers are on the include path
// (they are not needed otherwise).
#if defined(__has_include)
#if __has_include(<Eigen/Core>) && __has_include(<Eigen/Geometry>)
#include <Eigen/Core>
#include <Eigen/Geometry>
#define PNEC_COMPAT_HAS_EIGEN 1
#endif
#if defined(PNEC_COMPAT_HAS_EIGEN) && __has_include(<sophus/se3.hpp>)
#include <sophus/se3.hpp>
#define PNEC_COMPAT_HAS_SOPHUS 1
#endif
#endif

namespace pnec {

// ------------------------------------------------------------------ POD types

struct Vec3 {
  double v[3];
  Vec3() : v{0, 0, 0} {}
  Vec3(double x, double y, double z) : v{x, y, z} {}
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double *data() { return v; }
  const double *data() const { return v; }
  double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
static_assert(sizeof(Vec3) == 24, "Vec3 must be layout-identical to Eigen::Vector3d");

struct Mat3 {  // column-major, like Eigen::Matrix3d
  double m[9];
  Mat3() : m{0, 0, 0, 0, 0, 0, 0, 0, 0} {}
  double &operator()(int r, int c) { return m[c * 3 + r]; }
  double operator()(int r, int c) const { return m[c * 3 + r]; }
  double *data() { return m; }
  const double *data() const { return m; }
  static Mat3 Identity() {
    Mat3 I;
    I(0, 0) = I(1, 1) = I(2, 2) = 1.0;
    return I;
  }
  Vec3 operator*(const Vec3 &x) const {
    Vec3 o;
    for (int r = 0; r < 3; ++r) o[r] = (*this)(r, 0) * x[0] + (*this)(r, 1) * x[1] + (*this)(r, 2) * x[2];
    return o;
  }
};
static_assert(sizeof(Mat3) == 72, "Mat3 must be layout-identical to Eigen::Matrix3d");

struct Quat {  // x, y, z, w == Eigen::Quaterniond::coeffs()
  double c[4];
  Quat() : c{0, 0, 0, 1} {}
  // Eigen's constructor order: (w, x, y, z)
  Quat(double w, double x, double y, double z) : c{x, y, z, w} {}
  double x() const { return c[0]; }
  double y() const { return c[1]; }
  double z() const { return c[2]; }
  double w() const { return c[3]; }
  double norm() const { return std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]); }
  Quat normalized() const {
    const double n = norm();
    Quat q = *this;
    if (n > 0)
      for (double &e : q.c) e /= n;
    return q;
  }
  // Eigen::Quaternion::toRotationMatrix()
  Mat3 toRotationMatrix() const {
    const double tx = 2 * c[0], ty = 2 * c[1], tz = 2 * c[2];
    const double twx = tx * c[3], twy = ty * c[3], twz = tz * c[3];
    const double txx = tx * c[0], txy = ty * c[0], txz = tz * c[0];
    const double tyy = ty * c[1], tyz = tz * c[1], tzz = tz * c[2];
    Mat3 R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz;       R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;       R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;       R(2, 1) = tyz + twx;       R(2, 2) = 1 - (txx + tyy);
    return R;
  }
  // Eigen::Quaterniond(const Matrix3d&)
  static Quat FromRotationMatrix(const Mat3 &M) {
    Quat q;
    double t = M(0, 0) + M(1, 1) + M(2, 2);
    if (t > 0) {
      t = std::sqrt(t + 1.0);
      q.c[3] = 0.5 * t;
      t = 0.5 / t;
      q.c[0] = (M(2, 1) - M(1, 2)) * t;
      q.c[1] = (M(0, 2) - M(2, 0)) * t;
      q.c[2] = (M(1, 0) - M(0, 1)) * t;
    } else {
      int i = 0;
      if (M(1, 1) > M(0, 0)) i = 1;
      if (M(2, 2) > M(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
      q.c[i] = 0.5 * t;
      t = 0.5 / t;
      q.c[3] = (M(k, j) - M(j, k)) * t;
      q.c[j] = (M(j, i) + M(i, j)) * t;
      q.c[k] = (M(k, i) + M(i, k)) * t;
    }
    return q;
  }
};
static_assert(sizeof(Quat) == 32, "Quat must be layout-identical to Eigen::Quaterniond");

namespace detail {
template <class...>
using void_t = void;
// anything that looks like Sophus::SE3d: .unit_quaternion() with x()/y()/z()/w(), .translation() indexable
template <class S, class = void>
struct is_se3_like : std::false_type {};
template <class S>
struct is_se3_like<S, void_t<decltype(std::declval<const S &>().unit_quaternion().x()),
                             decltype(std::declval<const S &>().unit_quaternion().w()),
                             decltype(std::declval<const S &>().translation()[0])>> : std::true_type {};
}  // namespace detail

struct SE3 {  // memory order of Sophus::SE3d: unit quaternion (x,y,z,w), translation
  Quat q;
  Vec3 t;
  SE3() {}
  SE3(const Quat &q_, const Vec3 &t_) : q(q_.normalized()), t(t_) {}
  SE3(const Mat3 &R, const Vec3 &t_) : q(Quat::FromRotationMatrix(R).normalized()), t(t_) {}
  // from the reference's pose type (Sophus::SE3d) or anything with the same accessors: a call site
  // that passes `const Sophus::SE3d &initial_pose` compiles unchanged
  template <class S, class = typename std::enable_if<detail::is_se3_like<S>::value && !std::is_same<S, SE3>::value>::type>
  SE3(const S &o) {  // NOLINT(google-explicit-constructor): the conversion is the point
    const auto &oq = o.unit_quaternion();
    q = Quat(oq.w(), oq.x(), oq.y(), oq.z()).normalized();
    const auto &ot = o.translation();
    t = Vec3(ot[0], ot[1], ot[2]);
  }
  const Quat &unit_quaternion() const { return q; }
  Mat3 rotationMatrix() const { return q.toRotationMatrix(); }
  const Vec3 &translation() const { return t; }
  Vec3 &translation() { return t; }
  std::array<double, 16> matrix() const {  // 4x4, row-major
    const Mat3 R = rotationMatrix();
    return {R(0, 0), R(0, 1), R(0, 2), t[0], R(1, 0), R(1, 1), R(1, 2), t[1],
            R(2, 0), R(2, 1), R(2, 2), t[2], 0, 0, 0, 1};
  }
  const double *data() const { return q.c; }
#ifdef PNEC_COMPAT_HAS_SOPHUS
  // back to the reference's type: `Sophus::SE3d rel_pose = pnec.Solve(...)` compiles unchanged
  operator Sophus::SE3d() const {  // NOLINT(google-explicit-constructor)
    return Sophus::SE3d(Eigen::Quaterniond(q.c[3], q.c[0], q.c[1], q.c[2]), Eigen::Vector3d(t[0], t[1], t[2]));
  }
#endif
};
static_assert(sizeof(SE3) == 56, "SE3 must be 7 contiguous doubles");

namespace detail {

// One handle per calling thread: the reference's solver objects are per-call locals
// (`PNECCeres optimizer;`, pnec.cc:355), so concurrent callers never contend; here every thread gets
// its own device scratch and the handle-wide mutex of the C-ABI is never shared.
inline pnec_handle *Handle() {
  struct Holder {
    pnec_handle *h = nullptr;
    Holder() {
      const char *dev = std::getenv("PNEC_B200_DEVICE");
      if (pnec_create(dev ? std::atoi(dev) : 0, &h) != PNEC_OK)
        throw std::runtime_error(std::string("pnec_b200: ") + pnec_last_error());
    }
    ~Holder() { pnec_destroy(h); }
  };
  static thread_local Holder holder;  // no CPU fallback: construction throws without a B200
  return holder.h;
}

template <class Container>
const double *AsDoubles(const Container &c, std::size_t elem_bytes) {
  static_assert(sizeof(typename Container::value_type) % sizeof(double) == 0, "not a double array");
  if (sizeof(typename Container::value_type) != elem_bytes)
    throw std::invalid_argument("pnec_b200: element type has the wrong size");
  return reinterpret_cast<const double *>(c.data());
}

template <class T>
void Pose7(const T &quat_xyzw, const double *t, double out[7]) {
  const double *q = reinterpret_cast<const double *>(&quat_xyzw);
  for (int i = 0; i < 4; ++i) out[i] = q[i];
  for (int i = 0; i < 3; ++i) out[4 + i] = t[i];
}

struct SolveInfo {
  int32_t status = PNEC_STATUS_EMPTY;
  int32_t iterations = 0;
  double cost = 0.0, initial_cost = 0.0;
};

inline void SolveOne(int variant, const pnec_solver_opts &base, double regularization,
                     std::size_t n, const double *f1, const double *f2, const double *ct,
                     const double *ch, const double init7[7], double out7[7], SolveInfo *info) {
  pnec_solver_opts o = base;
  o.variant = variant;
  o.regularization = regularization;
  pnec_batch b{};
  b.num_problems = 1;
  b.n_per_problem = static_cast<int64_t>(n);
  b.offsets = nullptr;
  b.memspace = PNEC_MEM_HOST;
  b.bvs_host = f1;
  b.bvs_target = f2;
  b.covs_target = ct;
  b.covs_host = ch;
  b.poses = init7;
  pnec_solve_out out{};
  SolveInfo local;
  out.poses = out7;
  out.status = &local.status;
  out.iterations = &local.iterations;
  out.cost = &local.cost;
  out.initial_cost = &local.initial_cost;
  if (pnec_solve_batch(Handle(), &b, &o, &out, nullptr) != PNEC_OK)
    throw std::runtime_error(std::string("pnec_solve_batch: ") + pnec_last_error());
  if (info) *info = local;
}

}  // namespace detail

// --------------------------------------------------------------------- common

namespace common {

enum NoiseFrame { Host, Target, Both };  // include/common/common.h:59
enum CameraModel { Omnidirectional, Pinhole };

// src/common/common.cc:96-101
inline Mat3 SkewFromVector(const Vec3 &v) {
  Mat3 S;
  S(0, 1) = -v[2]; S(0, 2) = v[1];
  S(1, 0) = v[2];  S(1, 2) = -v[0];
  S(2, 0) = -v[1]; S(2, 1) = v[0];
  return S;
}

// src/common/common.cc:103-116
template <class V3>
inline void AnglesFromVec(const V3 &vector, double &theta, double &phi) {
  const double *v = reinterpret_cast<const double *>(&vector);
  const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (n == 0) {
    theta = 0.0;
    phi = 0.0;
  } else {
    theta = std::acos(v[2] / n);
    phi = (std::fabs(theta) < 1e-10) ? 0.0 : std::atan2(v[1] / n, v[0] / n);
  }
}

// src/common/common.cc:210-214, degrees
inline double RotationalDifference(const Quat &rotation_1, const Quat &rotation_2) {
  const Quat a = rotation_1.normalized(), b = rotation_2.normalized();
  const double ax = -a.c[0], ay = -a.c[1], az = -a.c[2], aw = a.c[3];
  const double w = aw * b.c[3] - ax * b.c[0] - ay * b.c[1] - az * b.c[2];
  const double x = aw * b.c[0] + ax * b.c[3] + ay * b.c[2] - az * b.c[1];
  const double y = aw * b.c[1] - ax * b.c[2] + ay * b.c[3] + az * b.c[0];
  const double z = aw * b.c[2] + ax * b.c[1] - ay * b.c[0] + az * b.c[3];
  return std::fabs(2.0 * std::atan2(std::sqrt(x * x + y * y + z * z), std::fabs(w))) * 180.0 / M_PI;
}
inline double RotationalDifference(const SE3 &a, const SE3 &b) { return RotationalDifference(a.q, b.q); }

// src/common/common.cc:216-235, degrees (including the doubled translation_1 test)
inline double TranslationalDifference(const Vec3 &translation_1, const Vec3 &translation_2,
                                      bool both_directions = true) {
  double error;
  const double n1 = translation_1.norm(), n2 = translation_2.norm();
  if (n1 < 1e-10 || n1 < 1e-10) {
    error = M_PI / 2;
  } else {
    const double c = (translation_1[0] * translation_2[0] + translation_1[1] * translation_2[1] +
                      translation_1[2] * translation_2[2]) / (n1 * n2);
    error = both_directions ? std::min(std::acos(c), std::acos(-c)) : std::acos(c);
  }
  return error * 180.0 / M_PI;
}

// src/common/common.cc:237-259 (mean PNEC energy, no regularisation) on the GPU
template <class BVs, class Covs>
inline double CostFunction(const BVs &bvs_1, const BVs &bvs_2, const Covs &covs, const SE3 &camera_pose) {
  pnec_batch b{};
  b.num_problems = 1;
  b.n_per_problem = static_cast<int64_t>(bvs_1.size());
  b.memspace = PNEC_MEM_HOST;
  b.bvs_host = detail::AsDoubles(bvs_1, 24);
  b.bvs_target = detail::AsDoubles(bvs_2, 24);
  b.covs_target = detail::AsDoubles(covs, 72);
  b.poses = camera_pose.data();
  double out = 0.0;
  if (pnec_cost_function_batch(detail::Handle(), &b, &out, nullptr) != PNEC_OK)
    throw std::runtime_error(std::string("pnec_cost_function_batch: ") + pnec_last_error());
  return out;
}

// src/common/common.cc:460-465
template <class V2>
inline Vec3 Unproject(const V2 &img_pt, const Mat3 &K_inv) {
  const double *p = reinterpret_cast<const double *>(&img_pt);
  Vec3 v = K_inv * Vec3(p[0], p[1], 1.0);
  const double n = v.norm();
  return Vec3(v[0] / n, v[1] / n, v[2] / n);
}

// src/common/common.cc:527-550: image-space covariances -> bearing-vector covariances (GPU)
template <class BVs, class Covs>
inline std::vector<Mat3> UnscentedTransform(const BVs &mus, const Covs &covs, const Mat3 &K_inv,
                                            double kappa = 1.0, CameraModel camera_model = Pinhole) {
  std::vector<Mat3> out(covs.size());
  if (mus.size() != covs.size()) {
    // the reference logs a warning and returns the covariances unchanged (common.cc:532-538)
    for (std::size_t i = 0; i < covs.size(); ++i)
      out[i] = *reinterpret_cast<const Mat3 *>(reinterpret_cast<const double *>(covs.data()) + 9 * i);
    return out;
  }
  if (mus.size() == 0) return out;
  const int model = (camera_model == Omnidirectional) ? PNEC_CAMERA_OMNIDIRECTIONAL : PNEC_CAMERA_PINHOLE;
  if (pnec_unscented_transform_batch(detail::Handle(), static_cast<int64_t>(mus.size()), PNEC_MEM_HOST,
                                     detail::AsDoubles(mus, 24), detail::AsDoubles(covs, 72), K_inv.data(),
                                     kappa, model, out[0].data(), nullptr) != PNEC_OK)
    throw std::runtime_error(std::string("pnec_unscented_transform_batch: ") + pnec_last_error());
  return out;
}
// src/common/common.cc:467-525: single point
inline Mat3 UnscentedTransform(const Vec3 &mu, const Mat3 &cov, const Mat3 &K_inv, double kappa = 1.0,
                               CameraModel camera_model = Pinhole) {
  return UnscentedTransform(std::vector<Vec3>{mu}, std::vector<Mat3>{cov}, K_inv, kappa, camera_model)[0];
}

// include/common/timing.h:48-67 (fields filled by the timed PNEC::Solve overload)
struct FrameTiming {
  explicit FrameTiming(int id) : id_(id) {}
  int id_;
  std::chrono::milliseconds frame_loading_{0}, feature_creation_{0}, nec_es_{0}, it_es_{0},
      avg_it_es_{0}, ceres_{0};
};

}  // namespace common

// --------------------------------------------------------------- optimization

namespace optimization {

// The slice of ceres::Solver::Options the refinement reads; defaults are Ceres'.
struct SolverOptions {
  int max_num_iterations = 50;
  int max_num_consecutive_invalid_steps = 5;
  bool jacobi_scaling = true;
  double function_tolerance = 1e-6;
  double gradient_tolerance = 1e-10;
  double parameter_tolerance = 1e-8;
  double initial_trust_region_radius = 1e4;
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double min_lm_diagonal = 1e-6;
  double max_lm_diagonal = 1e32;

  pnec_solver_opts ToAbi() const {
    pnec_solver_opts o;
    pnec_solver_opts_default(&o);
    o.max_num_iterations = max_num_iterations;
    o.max_num_consecutive_invalid_steps = max_num_consecutive_invalid_steps;
    o.jacobi_scaling = jacobi_scaling ? 1 : 0;
    o.function_tolerance = function_tolerance;
    o.gradient_tolerance = gradient_tolerance;
    o.parameter_tolerance = parameter_tolerance;
    o.initial_trust_region_radius = initial_trust_region_radius;
    o.max_trust_region_radius = max_trust_region_radius;
    o.min_trust_region_radius = min_trust_region_radius;
    o.min_relative_decrease = min_relative_decrease;
    o.min_lm_diagonal = min_lm_diagonal;
    o.max_lm_diagonal = max_lm_diagonal;
    return o;
  }
};

namespace detail_opt {
// State + accessors shared by PNECCeres and NECCeres (pnec_ceres.cc:43-68,170-206).
class CeresLike {
 public:
  CeresLike() : orientation_(1.0, 0.0, 0.0, 0.0), theta_(0.0), phi_(0.0) {}
  CeresLike(const SE3 &init, const SolverOptions &options) : options_(options) {
    orientation_ = init.unit_quaternion();
    common::AnglesFromVec(init.translation(), theta_, phi_);
  }
  // sic: the reference drops `options` in this constructor (pnec_ceres.cc:57-59)
  CeresLike(const Quat &orientation, double theta, double phi, const SolverOptions &)
      : orientation_(orientation), theta_(theta), phi_(phi) {}
  CeresLike(const Quat &orientation, const Vec3 &translation, const SolverOptions &options)
      : orientation_(orientation), options_(options) {
    common::AnglesFromVec(translation, theta_, phi_);
  }

  void InitValues(const Quat orientation, double theta, double phi) {
    orientation_ = orientation;
    theta_ = theta;
    phi_ = phi;
  }
  void InitValues(const SE3 &init) {
    orientation_ = init.unit_quaternion();
    common::AnglesFromVec(init.translation(), theta_, phi_);
  }
  template <class V3>
  void InitValues(const Quat &orientation, const V3 &translation) {
    orientation_ = orientation;
    common::AnglesFromVec(translation, theta_, phi_);
  }
  void SetOptions(const SolverOptions &options) { options_ = options; }

  Mat3 Orientation() const { return orientation_.normalized().toRotationMatrix(); }
  Vec3 Translation() const {
    return Vec3(std::sin(theta_) * std::cos(phi_), std::sin(theta_) * std::sin(phi_), std::cos(theta_));
  }
  SE3 Result() const { return SE3(orientation_.normalized(), Translation()); }

  // More than the reference exposes (its Summary is private and never read).
  int Status() const { return info_.status; }
  int Iterations() const { return info_.iterations; }
  double FinalCost() const { return info_.cost; }
  double InitialCost() const { return info_.initial_cost; }

 protected:
  void Run(int variant, double reg, std::size_t n, const double *f1, const double *f2,
           const double *ct, const double *ch) {
    // the C-ABI start pose is (q, t); t(theta, phi) reproduces theta_, phi_
    double init7[7], out7[7];
    const Vec3 t = Translation();
    detail::Pose7(orientation_, t.data(), init7);
    detail::SolveOne(variant, options_.ToAbi(), reg, n, f1, f2, ct, ch, init7, out7, &info_);
    // keep the optimiser's own (non-normalised) state semantics: Result() normalises
    orientation_ = Quat(out7[3], out7[0], out7[1], out7[2]);
    const Vec3 to(out7[4], out7[5], out7[6]);
    common::AnglesFromVec(to, theta_, phi_);
  }
  Quat orientation_;
  double theta_, phi_;
  SolverOptions options_;
  detail::SolveInfo info_;
};
}  // namespace detail_opt

class PNECCeres : public detail_opt::CeresLike {
 public:
  PNECCeres() {}
  explicit PNECCeres(const SE3 &init, const SolverOptions &options = SolverOptions()) : CeresLike(init, options) {}
  PNECCeres(const Quat &orientation, double theta, double phi, const SolverOptions &options = SolverOptions())
      : CeresLike(orientation, theta, phi, options) {}
  PNECCeres(const Quat &orientation, const Vec3 &translation, const SolverOptions &options = SolverOptions())
      : CeresLike(orientation, translation, options) {}

  // src/optimization/pnec_ceres.cc:70-111
  template <class BVs, class Covs>
  void Optimize(const BVs &bvs_1, const BVs &bvs_2, const Covs &covs, double regularization,
                common::NoiseFrame noise_frame = common::Target) {
    const int variant = (noise_frame == common::Host) ? PNEC_VARIANT_HOST : PNEC_VARIANT_TARGET;
    Run(variant, regularization, bvs_1.size(), detail::AsDoubles(bvs_1, 24),
        detail::AsDoubles(bvs_2, 24), detail::AsDoubles(covs, 72), nullptr);
  }
  // src/optimization/pnec_ceres.cc:113-168 (PNECSymmetrical)
  template <class BVs, class Covs>
  void Optimize(const BVs &bvs_1, const BVs &bvs_2, const Covs &covs_1, const Covs &covs_2,
                double regularization) {
    Run(PNEC_VARIANT_SYMMETRIC, regularization, bvs_1.size(), detail::AsDoubles(bvs_1, 24),
        detail::AsDoubles(bvs_2, 24), detail::AsDoubles(covs_2, 72), detail::AsDoubles(covs_1, 72));
  }
};

class NECCeres : public detail_opt::CeresLike {
 public:
  NECCeres() {}
  explicit NECCeres(const SE3 &init, const SolverOptions &options = SolverOptions()) : CeresLike(init, options) {}
  NECCeres(const Quat &orientation, double theta, double phi, const SolverOptions &options = SolverOptions())
      : CeresLike(orientation, theta, phi, options) {}
  NECCeres(const Quat &orientation, const Vec3 &translation, const SolverOptions &options = SolverOptions())
      : CeresLike(orientation, translation, options) {}

  // src/optimization/nec_ceres.cc:73-101
  template <class BVs>
  void Optimize(const BVs &bvs_1, const BVs &bvs_2) {
    Run(PNEC_VARIANT_NEC, 0.0, bvs_1.size(), detail::AsDoubles(bvs_1, 24),
        detail::AsDoubles(bvs_2, 24), nullptr, nullptr);
  }
};

}  // namespace optimization

// ------------------------------------------------ translation given rotation

namespace common {
// src/common/common.cc:127-181 composed: the NEC translation for a given rotation (GPU).
// (ComposeM skips the first correspondence, as in the reference.)
template <class BVs>
inline Vec3 TranslationFromM_ComposeM(const BVs &bvs_1, const BVs &bvs_2, const Mat3 &rotation) {
  const SE3 pose(rotation, Vec3(0, 0, 1));
  pnec_batch b{};
  b.num_problems = 1;
  b.n_per_problem = static_cast<int64_t>(bvs_1.size());
  b.memspace = PNEC_MEM_HOST;
  b.bvs_host = detail::AsDoubles(bvs_1, 24);
  b.bvs_target = detail::AsDoubles(bvs_2, 24);
  b.poses = pose.data();
  Vec3 t;
  if (pnec_nec_translation_batch(detail::Handle(), &b, t.data(), nullptr, nullptr) != PNEC_OK)
    throw std::runtime_error(std::string("pnec_nec_translation_batch: ") + pnec_last_error());
  return t;
}

}  // namespace common

namespace optimization {
// The translation half of PNEC::WeightedEigensolver for a given rotation (pnec.cc:317-343):
// builds A_i / B_i, scans {initial_translation} + fibonacci_sphere(500) and runs scf(..., 10)
// (src/optimization/scf.cc).  Sign of the result is arbitrary, as with Eigen's eigenvectors.
template <class BVs, class Covs>
inline Vec3 ScfTranslation(const BVs &bvs_1, const BVs &bvs_2, const Covs &projected_covariances,
                           const Mat3 &rotation, const Vec3 &initial_translation, double regularization,
                           int fibonacci_samples = 500, int scf_steps = 10) {
  const SE3 pose0(rotation, initial_translation);
  pnec_batch b{};
  b.num_problems = 1;
  b.n_per_problem = static_cast<int64_t>(bvs_1.size());
  b.memspace = PNEC_MEM_HOST;
  b.bvs_host = detail::AsDoubles(bvs_1, 24);
  b.bvs_target = detail::AsDoubles(bvs_2, 24);
  b.covs_target = detail::AsDoubles(projected_covariances, 72);
  b.poses = pose0.data();
  Vec3 t;
  if (pnec_scf_translation_batch(detail::Handle(), &b, regularization, fibonacci_samples, scf_steps,
                                 t.data(), nullptr, nullptr) != PNEC_OK)
    throw std::runtime_error(std::string("pnec_scf_translation_batch: ") + pnec_last_error());
  return t;
}
}  // namespace optimization

// ------------------------------------------------------------------- features

namespace features {

// include/frames/keypoints.h:50-69.  The reference's constructor unprojects one keypoint on the
// CPU with the Camera singleton's intrinsics; here a whole frame is unprojected in one GPU call.
struct KeyPoint {
  double point_[2] = {0, 0};
  Vec3 bearing_vector_;
  double img_covariance_[4] = {0, 0, 0, 0};  // column-major 2x2
  Mat3 bv_covariance_;
};

// KeyPoint::Unproject (src/frames/keypoints.cc:49-62) for every keypoint: fills
// bearing_vector_ and bv_covariance_ from point_ and img_covariance_.
inline void UnprojectKeyPoints(std::vector<KeyPoint> &keypoints, const Mat3 &K_inv) {
  const std::size_t n = keypoints.size();
  if (n == 0) return;
  std::vector<double> pts(2 * n), c2(4 * n), bvs(3 * n), covs(9 * n);
  for (std::size_t i = 0; i < n; ++i) {
    pts[2 * i] = keypoints[i].point_[0];
    pts[2 * i + 1] = keypoints[i].point_[1];
    for (int k = 0; k < 4; ++k) c2[4 * i + k] = keypoints[i].img_covariance_[k];
  }
  if (pnec_keypoints_unproject_batch(detail::Handle(), static_cast<int64_t>(n), PNEC_MEM_HOST, pts.data(),
                                     c2.data(), K_inv.data(), bvs.data(), covs.data(), nullptr) != PNEC_OK)
    throw std::runtime_error(std::string("pnec_keypoints_unproject_batch: ") + pnec_last_error());
  for (std::size_t i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) keypoints[i].bearing_vector_[k] = bvs[3 * i + k];
    for (int k = 0; k < 9; ++k) keypoints[i].bv_covariance_.m[k] = covs[9 * i + k];
  }
}

}  // namespace features

// -------------------------------------------------------- rel_pose_estimation

namespace rel_pose_estimation {

// include/rel_pose_estimation/pnec_config.h:46-65
struct Options {
  bool use_nec_ = false;
  common::NoiseFrame noise_frame_ = common::Target;
  double regularization_ = 1.0e-13;
  std::size_t weighted_iterations_ = 10;
  bool use_scf_ = true;
  bool use_ceres_ = true;
  optimization::SolverOptions ceres_options_ = optimization::SolverOptions();
  bool use_ransac_ = true;
  int max_ransac_iterations_ = 5000;
  int ransac_sample_size_ = 10;
  int min_matches_ = 30;
  int min_inliers_ = 10;
  int min_matches_further_ = 20;
  // extension (not in the reference): key of RANSAC's random stream.  opengv seeds from time(0) and
  // rand(), so the reference is not reproducible run to run; this implementation is, per seed.
  std::uint64_t ransac_seed_ = 1;
};

class PNEC {
 public:
  explicit PNEC(const Options &options) : options_(options) {}

  // src/rel_pose_estimation/pnec.cc:77-124: Eigensolver (+ RANSAC and InlierExtraction with
  // use_ransac_, the default) -> (WeightedEigensolver) -> CeresSolver / NECCeresSolver, every stage on
  // the GPU in one pnec_frame_solve_batch call.  `inliers` receives ransac.inliers_ (ascending indices
  // into bvs1 / bvs2) and is cleared without RANSAC (pnec.cc:277).  Like the reference, the refinement
  // runs with default Ceres options (pnec.cc:355 constructs `PNECCeres optimizer;`) and the TARGET
  // noise frame, whatever options_ holds.
  template <class BVs, class Covs>
  SE3 Solve(const BVs &bvs1, const BVs &bvs2, const Covs &projected_covs, const SE3 &initial_pose) {
    std::vector<int> inliers;
    return Solve(bvs1, bvs2, projected_covs, initial_pose, inliers);
  }
  template <class BVs, class Covs>
  SE3 Solve(const BVs &bvs1, const BVs &bvs2, const Covs &projected_covs, const SE3 &initial_pose,
            std::vector<int> &inliers) {
    return SolveImpl(bvs1, bvs2, detail::AsDoubles(projected_covs, 72), initial_pose, FrameOpts(), &inliers, nullptr);
  }
  // The timed overloads (pnec.cc:127-208): nec_es_ = Eigensolver + InlierExtraction, it_es_ =
  // WeightedEigensolver, avg_it_es_ = it_es_ / weighted_iterations_, ceres_ = the refinement --
  // measured with CUDA events around the same stages on the device, truncated to milliseconds like
  // the reference's duration_cast (a single pair takes well under one: expect zeros).
  template <class BVs, class Covs>
  SE3 Solve(const BVs &bvs1, const BVs &bvs2, const Covs &projected_covs, const SE3 &initial_pose,
            common::FrameTiming &timing) {
    std::vector<int> inliers;
    return Solve(bvs1, bvs2, projected_covs, initial_pose, inliers, timing);
  }
  template <class BVs, class Covs>
  SE3 Solve(const BVs &bvs1, const BVs &bvs2, const Covs &projected_covs, const SE3 &initial_pose,
            std::vector<int> &inliers, common::FrameTiming &timing) {
    float ms[3] = {0, 0, 0};
    SE3 r = SolveImpl(bvs1, bvs2, detail::AsDoubles(projected_covs, 72), initial_pose, FrameOpts(), &inliers, ms);
    last_stage_ms_[0] = ms[0]; last_stage_ms_[1] = ms[1]; last_stage_ms_[2] = ms[2];
    const auto to_ms = [](double v) { return std::chrono::milliseconds(static_cast<long long>(v)); };
    timing.nec_es_ = to_ms(ms[0]);
    if (!options_.use_nec_) {
      timing.it_es_ = to_ms(options_.weighted_iterations_ > 1 ? ms[1] : 0.0);
      if (options_.weighted_iterations_ > 1)
        timing.avg_it_es_ = to_ms(ms[1] / static_cast<double>(options_.weighted_iterations_));
    }
    timing.ceres_ = to_ms(options_.use_ceres_ ? ms[2] : 0.0);
    return r;
  }
  // stage times of the last timed Solve in (fractional) milliseconds: nec_es, it_es, ceres
  const double *LastStageMilliseconds() const { return last_stage_ms_; }

  // PNEC::Eigensolver (pnec.cc:231-281): with use_ransac_ opengv's RANSAC over the eigensolver,
  // optimizeModelCoefficients on the inliers and the translation from ComposeM over the inliers;
  // without, opengv::relative_pose::eigensolver started at initial_pose and the translation from
  // ComposeM over everything (`inliers` cleared).
  template <class BVs>
  SE3 Eigensolver(const BVs &bvs1, const BVs &bvs2, const SE3 &initial_pose, std::vector<int> &inliers) {
    pnec_frame_opts fo = FrameOpts();
    fo.use_nec = 1;
    fo.use_ceres = 0;
    return SolveImpl(bvs1, bvs2, static_cast<const double *>(nullptr), initial_pose, fo, &inliers, nullptr);
  }

  // PNEC::WeightedEigensolver (pnec.cc:283-348): `initial_pose` supplies the weights of every
  // iteration and the first iteration's start.
  template <class BVs, class Covs>
  SE3 WeightedEigensolver(const BVs &bvs1, const BVs &bvs2, const Covs &projected_covariances,
                          const SE3 &initial_pose) {
    SE3 rel = initial_pose;
    pnec_handle *h = detail::Handle();
    for (std::size_t it = 0; it + 1 < options_.weighted_iterations_; ++it) {
      pnec_batch b = MakeBatch(bvs1, bvs2, detail::AsDoubles(projected_covariances, 72), rel);
      SE3 next;
      if (pnec_eigensolver_batch(h, &b, initial_pose.data(), options_.regularization_,
                                 reinterpret_cast<double *>(&next), nullptr, nullptr, nullptr) != PNEC_OK)
        throw std::runtime_error(std::string("pnec_eigensolver_batch: ") + pnec_last_error());
      b.poses = next.data();  // rotation of this iteration + previous translation as the scan's first candidate
      Vec3 t;
      if (pnec_scf_translation_batch(h, &b, options_.regularization_, 500, 10, t.data(), nullptr, nullptr) != PNEC_OK)
        throw std::runtime_error(std::string("pnec_scf_translation_batch: ") + pnec_last_error());
      next.t = t;
      rel = next;
    }
    return rel;
  }

  // src/rel_pose_estimation/pnec.cc:350-370 — ignores options_.ceres_options_ and
  // options_.noise_frame_ exactly like the reference (default-constructed optimizer).
  template <class BVs, class Covs>
  SE3 CeresSolver(const BVs &bvs1, const BVs &bvs2, const Covs &projected_covariances, const SE3 &initial_pose) {
    optimization::PNECCeres optimizer;
    optimizer.InitValues(Quat::FromRotationMatrix(initial_pose.rotationMatrix()), initial_pose.translation());
    optimizer.Optimize(bvs1, bvs2, projected_covariances, options_.regularization_);
    last_status_ = optimizer.Status();
    last_iterations_ = optimizer.Iterations();
    return optimizer.Result();
  }
  // src/rel_pose_estimation/pnec.cc:372-392
  template <class BVs, class Covs>
  SE3 CeresSolverFull(const BVs &bvs1, const BVs &bvs2, const Covs &projected_covariances,
                      double regularization, const SE3 &initial_pose) {
    optimization::PNECCeres optimizer;
    optimizer.InitValues(Quat::FromRotationMatrix(initial_pose.rotationMatrix()), initial_pose.translation());
    optimizer.Optimize(bvs1, bvs2, projected_covariances, regularization);
    last_status_ = optimizer.Status();
    last_iterations_ = optimizer.Iterations();
    return optimizer.Result();
  }
  // src/rel_pose_estimation/pnec.cc:394-411
  template <class BVs>
  SE3 NECCeresSolver(const BVs &bvs1, const BVs &bvs2, const SE3 &initial_pose) {
    optimization::NECCeres optimizer;
    optimizer.InitValues(Quat::FromRotationMatrix(initial_pose.rotationMatrix()), initial_pose.translation());
    optimizer.Optimize(bvs1, bvs2);
    last_status_ = optimizer.Status();
    last_iterations_ = optimizer.Iterations();
    return optimizer.Result();
  }

  // Batched extension: B frame pairs per call (what the reference does in a loop,
  // src/run_simulation.cc:326-341).  Flat C-ABI layout, host memory.
  void CeresSolverBatch(std::size_t num_problems, const int64_t *offsets, std::size_t n_per_problem,
                        const double *bvs1, const double *bvs2, const double *projected_covariances,
                        const SE3 *initial_poses, SE3 *results, int32_t *status = nullptr,
                        int32_t *iterations = nullptr) {
    pnec_solver_opts o;
    pnec_solver_opts_default(&o);
    o.variant = PNEC_VARIANT_TARGET;
    o.regularization = options_.regularization_;
    pnec_batch b{};
    b.num_problems = static_cast<int64_t>(num_problems);
    b.n_per_problem = static_cast<int64_t>(n_per_problem);
    b.offsets = offsets;
    b.memspace = PNEC_MEM_HOST;
    b.bvs_host = bvs1;
    b.bvs_target = bvs2;
    b.covs_target = projected_covariances;
    b.poses = reinterpret_cast<const double *>(initial_poses);
    pnec_solve_out out{};
    out.poses = reinterpret_cast<double *>(results);
    out.status = status;
    out.iterations = iterations;
    if (pnec_solve_batch(detail::Handle(), &b, &o, &out, nullptr) != PNEC_OK)
      throw std::runtime_error(std::string("pnec_solve_batch: ") + pnec_last_error());
  }

  // Batched Solve(): B frame pairs per call, flat C-ABI layout, host memory.
  void SolveBatch(std::size_t num_problems, const int64_t *offsets, std::size_t n_per_problem, const double *bvs1,
                  const double *bvs2, const double *projected_covariances, const SE3 *initial_poses, SE3 *results,
                  SE3 *eigensolver_results = nullptr, int32_t *status = nullptr, int32_t *iterations = nullptr,
                  int32_t *num_inliers = nullptr, int32_t *inlier_index = nullptr) {
    const pnec_frame_opts fo = FrameOpts();
    pnec_batch b{};
    b.num_problems = static_cast<int64_t>(num_problems);
    b.n_per_problem = static_cast<int64_t>(n_per_problem);
    b.offsets = offsets;
    b.memspace = PNEC_MEM_HOST;
    b.bvs_host = bvs1;
    b.bvs_target = bvs2;
    b.covs_target = projected_covariances;
    b.poses = reinterpret_cast<const double *>(initial_poses);
    pnec_frame_out o{};
    o.poses = reinterpret_cast<double *>(results);
    o.es_poses = reinterpret_cast<double *>(eigensolver_results);
    o.status = status;
    o.iterations = iterations;
    o.num_inliers = num_inliers;    // [B]
    o.inlier_index = inlier_index;  // [total], pair b's list at its offset (pnec_b200.h)
    if (pnec_frame_solve_batch(detail::Handle(), &b, &fo, &o, nullptr) != PNEC_OK)
      throw std::runtime_error(std::string("pnec_frame_solve_batch: ") + pnec_last_error());
  }

  int LastStatus() const { return last_status_; }
  int LastIterations() const { return last_iterations_; }
  // ES_solution of the last Solve() (pnec.cc:86)
  const SE3 &LastEigensolverPose() const { return last_es_; }

 protected:
  pnec_frame_opts FrameOpts() const {
    pnec_frame_opts fo;
    pnec_frame_opts_default(&fo);
    fo.use_nec = options_.use_nec_ ? 1 : 0;
    fo.use_ceres = options_.use_ceres_ ? 1 : 0;
    fo.weighted_iterations = static_cast<int32_t>(options_.weighted_iterations_);
    fo.use_ransac = options_.use_ransac_ ? 1 : 0;
    fo.max_ransac_iterations = options_.max_ransac_iterations_;
    fo.ransac_sample_size = options_.ransac_sample_size_;
    fo.ransac_seed = options_.ransac_seed_;
    fo.ceres.regularization = options_.regularization_;
    return fo;
  }
  template <class BVs>
  static pnec_batch MakeBatch(const BVs &bvs1, const BVs &bvs2, const double *covs, const SE3 &pose) {
    pnec_batch b{};
    b.num_problems = 1;
    b.n_per_problem = static_cast<int64_t>(bvs1.size());
    b.memspace = PNEC_MEM_HOST;
    b.bvs_host = detail::AsDoubles(bvs1, 24);
    b.bvs_target = detail::AsDoubles(bvs2, 24);
    b.covs_target = covs;
    b.poses = pose.data();
    return b;
  }
  template <class BVs>
  SE3 SolveImpl(const BVs &bvs1, const BVs &bvs2, const double *covs, const SE3 &initial_pose,
                const pnec_frame_opts &fo, std::vector<int> *inliers, float *stage_ms) {
    pnec_batch b = MakeBatch(bvs1, bvs2, covs, initial_pose);
    SE3 out, es;
    int32_t status = PNEC_STATUS_EMPTY, iters = 0, num_inliers = 0;
    const std::size_t n = bvs1.size();
    std::vector<int32_t> index(fo.use_ransac && inliers ? n : 0);
    pnec_frame_out o{};
    o.poses = reinterpret_cast<double *>(&out);
    o.es_poses = reinterpret_cast<double *>(&es);
    o.status = &status;
    o.iterations = &iters;
    o.num_inliers = &num_inliers;
    o.inlier_index = index.empty() ? nullptr : index.data();
    o.stage_ms = stage_ms;
    if (pnec_frame_solve_batch(detail::Handle(), &b, &fo, &o, nullptr) != PNEC_OK)
      throw std::runtime_error(std::string("pnec_frame_solve_batch: ") + pnec_last_error());
    if (inliers) inliers->assign(index.begin(), index.begin() + (index.empty() ? 0 : num_inliers));
    last_status_ = status;
    last_iterations_ = iters;
    last_es_ = es;
    return out;
  }

  Options options_;
  int last_status_ = PNEC_STATUS_EMPTY;
  int last_iterations_ = 0;
  SE3 last_es_;
  double last_stage_ms_[3] = {0, 0, 0};
};

}  // namespace rel_pose_estimation
}  // namespace pnec

#endif  // PNEC_COMPAT_HPP_
