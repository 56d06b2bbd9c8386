// Minimal stand-in for <sophus/se3.hpp>: SE3d with Sophus' memory order (unit quaternion, translation)
// and the accessors the reference uses.  TEST FIXTURE ONLY (see ../Eigen/Core).
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
namespace Sophus {
class SE3d {
 public:
  SE3d() {}
  SE3d(const Eigen::Quaterniond &q, const Eigen::Vector3d &t) : q_(q), t_(t) {}
  const Eigen::Quaterniond &unit_quaternion() const { return q_; }
  const Eigen::Vector3d &translation() const { return t_; }

 private:
  Eigen::Quaterniond q_;
  Eigen::Vector3d t_;
};
}  // namespace Sophus
