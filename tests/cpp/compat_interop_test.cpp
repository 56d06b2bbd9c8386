// A call site written against the reference's own types -- opengv::bearingVectors_t
// (std::vector<Eigen::Vector3d>), std::vector<Eigen::Matrix3d>, Sophus::SE3d in and out
// (include/rel_pose_estimation/pnec.h:53-75) -- compiled against pnec_compat.hpp.  Eigen / Sophus are
// not installed in this image: tests/cpp/mock/ holds minimal headers with the same names, layouts
// and accessors, enough to prove the __has_include interop compiles and converts correctly.
// Input / output as compat_test.cpp.
#include <cstdio>
#include <vector>

#include "pnec/pnec_compat.hpp"

#if !defined(PNEC_COMPAT_HAS_EIGEN) || !defined(PNEC_COMPAT_HAS_SOPHUS)
#error "compile with -I tests/cpp/mock"
#endif

namespace opengv {
typedef std::vector<Eigen::Vector3d> bearingVectors_t;
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  long long n = 0;
  if (std::fread(&n, sizeof(n), 1, f) != 1) return 4;
  opengv::bearingVectors_t bvs1(n), bvs2(n);
  std::vector<Eigen::Matrix3d> covs(n), covs_h(n);
  double p[7];
  bool ok = std::fread(bvs1.data(), 24, n, f) == (size_t)n && std::fread(bvs2.data(), 24, n, f) == (size_t)n &&
            std::fread(covs.data(), 72, n, f) == (size_t)n && std::fread(covs_h.data(), 72, n, f) == (size_t)n &&
            std::fread(p, 8, 7, f) == 7;
  std::fclose(f);
  if (!ok) return 5;
  const Sophus::SE3d initial_pose(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]));
  // src/run_simulation.cc:74-86
  pnec::rel_pose_estimation::PNEC pnec((pnec::rel_pose_estimation::Options()));
  std::vector<int> inliers;
  Sophus::SE3d rel_pose = pnec.Solve(bvs1, bvs2, covs, initial_pose, inliers);
  const Eigen::Quaterniond &q = rel_pose.unit_quaternion();
  const Eigen::Vector3d &t = rel_pose.translation();
  std::printf("SolveSophus %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d %d\n", q.x(), q.y(), q.z(), q.w(), t[0], t[1],
              t[2], pnec.LastStatus(), (int)inliers.size());
  Sophus::SE3d refined = pnec.CeresSolver(bvs1, bvs2, covs, initial_pose);
  std::printf("CeresSolverSophus %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d %d\n", refined.unit_quaternion().x(),
              refined.unit_quaternion().y(), refined.unit_quaternion().z(), refined.unit_quaternion().w(),
              refined.translation()[0], refined.translation()[1], refined.translation()[2], pnec.LastStatus(),
              pnec.LastIterations());
  return 0;
}
