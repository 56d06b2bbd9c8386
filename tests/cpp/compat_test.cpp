// Exercises the reference's C++ call shapes through include/pnec/pnec_compat.hpp.
// Input (binary, from the Python test): int64 n, then f1[n*3], f2[n*3], cov_t[n*9],
// cov_h[n*9], init pose7.  Output (text): one line per API with 7 pose numbers + status + iters.
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "pnec/pnec_compat.hpp"

using pnec::Mat3;
using pnec::SE3;
using pnec::Vec3;

static void print_pose(const char *tag, const SE3 &p, int status, int iters) {
  std::printf("%s %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d %d\n", tag, p.q.c[0], p.q.c[1], p.q.c[2],
              p.q.c[3], p.t[0], p.t[1], p.t[2], status, iters);
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  long long n = 0;
  if (std::fread(&n, sizeof(n), 1, f) != 1) return 4;
  std::vector<Vec3> bvs1(n), bvs2(n);
  std::vector<Mat3> covs_t(n), covs_h(n);
  double init7[7];
  bool ok = std::fread(bvs1.data(), sizeof(Vec3), n, f) == (size_t)n &&
            std::fread(bvs2.data(), sizeof(Vec3), n, f) == (size_t)n &&
            std::fread(covs_t.data(), sizeof(Mat3), n, f) == (size_t)n &&
            std::fread(covs_h.data(), sizeof(Mat3), n, f) == (size_t)n &&
            std::fread(init7, sizeof(double), 7, f) == 7;
  std::fclose(f);
  if (!ok) return 5;
  const SE3 init(pnec::Quat(init7[3], init7[0], init7[1], init7[2]), Vec3(init7[4], init7[5], init7[6]));

  // call shapes of src/run_simulation.cc:157-180 (Ablation) and src/rel_pose_estimation/pnec.cc:350-411
  pnec::rel_pose_estimation::Options options;
  options.use_ransac_ = false;
  options.weighted_iterations_ = 0;
  pnec::rel_pose_estimation::PNEC pnec(options);
  SE3 a = pnec.CeresSolver(bvs1, bvs2, covs_t, init);
  print_pose("CeresSolver", a, pnec.LastStatus(), pnec.LastIterations());
  SE3 b = pnec.CeresSolverFull(bvs1, bvs2, covs_t, 1e-10, init);
  print_pose("CeresSolverFull", b, pnec.LastStatus(), pnec.LastIterations());
  SE3 c = pnec.NECCeresSolver(bvs1, bvs2, init);
  print_pose("NECCeresSolver", c, pnec.LastStatus(), pnec.LastIterations());
  SE3 d = pnec.Solve(bvs1, bvs2, covs_t, init);
  print_pose("Solve", d, pnec.LastStatus(), pnec.LastIterations());
  {
    // the whole PNEC::Solve pipeline (pnec.cc:77-124) with the reference's defaults minus RANSAC
    pnec::rel_pose_estimation::Options full;
    full.use_ransac_ = false;
    pnec::rel_pose_estimation::PNEC solver(full);
    std::vector<int> inliers(3, 7);
    SE3 r = solver.Solve(bvs1, bvs2, covs_t, init, inliers);
    print_pose("SolveFull", r, solver.LastStatus(), (int)inliers.size());
    print_pose("SolveFullES", solver.LastEigensolverPose(), 0, 0);
    SE3 es = solver.Eigensolver(bvs1, bvs2, init, inliers);
    print_pose("Eigensolver", es, 0, 0);
    SE3 w = solver.WeightedEigensolver(bvs1, bvs2, covs_t, es);
    print_pose("WeightedEigensolver", w, 0, 0);
    full.use_nec_ = true;
    pnec::rel_pose_estimation::PNEC nec(full);
    print_pose("SolveNEC", nec.Solve(bvs1, bvs2, covs_t, init), nec.LastStatus(), nec.LastIterations());
  }
  {
    // the reference's DEFAULT options (use_ransac_ = true, pnec_config.h:58), as run_simulation's
    // PNEC() uses them (src/run_simulation.cc:74-86), incl. the timed overload (pnec.cc:135-208)
    pnec::rel_pose_estimation::PNEC solver((pnec::rel_pose_estimation::Options()));
    std::vector<int> inliers(3, 7);
    SE3 r = solver.Solve(bvs1, bvs2, covs_t, init, inliers);
    print_pose("SolveDefault", r, solver.LastStatus(), (int)inliers.size());
    print_pose("SolveDefaultES", solver.LastEigensolverPose(), 0, 0);
    std::printf("SolveDefaultInliers");
    for (int i : inliers) std::printf(" %d", i);
    std::printf("\n");
    pnec::common::FrameTiming timing(0);
    std::vector<int> inliers2;
    SE3 r2 = solver.Solve(bvs1, bvs2, covs_t, init, inliers2, timing);
    const double *ms = solver.LastStageMilliseconds();
    std::printf("SolveTimed %d %d %.6f %.6f %.6f %lld\n", (int)(r2.q.c[0] == r.q.c[0] && r2.t[2] == r.t[2]),
                (int)(inliers2 == inliers), ms[0], ms[1], ms[2],
                (long long)(timing.nec_es_.count() + timing.it_es_.count() + timing.ceres_.count()));
    std::vector<int> es_inliers;
    SE3 es = solver.Eigensolver(bvs1, bvs2, init, es_inliers);
    print_pose("EigensolverRansac", es, 0, (int)es_inliers.size());
    // re-entrancy: concurrent callers (one handle per thread) get the sequential results
    SE3 t_out[4];
    std::vector<std::thread> threads;
    for (int k = 0; k < 4; ++k)
      threads.emplace_back([&, k] {
        pnec::rel_pose_estimation::PNEC local((pnec::rel_pose_estimation::Options()));
        for (int rep = 0; rep < 3; ++rep) t_out[k] = local.Solve(bvs1, bvs2, covs_t, init);
      });
    for (auto &t : threads) t.join();
    int same = 0;
    for (int k = 0; k < 4; ++k)
      same += (int)(t_out[k].q.c[0] == r.q.c[0] && t_out[k].q.c[3] == r.q.c[3] && t_out[k].t[1] == r.t[1]);
    std::printf("Threads %d\n", same);
  }

  // lower level: include/optimization/pnec_ceres.h:50-81
  pnec::optimization::PNECCeres opt;
  opt.InitValues(pnec::Quat::FromRotationMatrix(init.rotationMatrix()), init.translation());
  opt.Optimize(bvs1, bvs2, covs_t, 1e-13, pnec::common::Host);
  print_pose("PNECCeresHost", opt.Result(), opt.Status(), opt.Iterations());
  pnec::optimization::PNECCeres sym(init);
  sym.Optimize(bvs1, bvs2, covs_h, covs_t, 1e-13);
  print_pose("PNECCeresSymmetric", sym.Result(), sym.Status(), sym.Iterations());
  std::printf("CostFunction %.17g\n", pnec::common::CostFunction(bvs1, bvs2, covs_t, a));

  // include/common/common.h:103-113: covariance propagation for the first 3 correspondences
  {
    std::vector<Vec3> mus(bvs2.begin(), bvs2.begin() + 3);
    for (auto &m : mus) { m[0] *= 800; m[1] *= 800; m[2] *= 800; }
    std::vector<Mat3> img(3);
    for (auto &c : img) { c(0, 0) = 0.7; c(1, 1) = 0.4; c(0, 1) = c(1, 0) = 0.1; }
    const auto proj = pnec::common::UnscentedTransform(mus, img, Mat3::Identity(), 1.0, pnec::common::Pinhole);
    const Mat3 one = pnec::common::UnscentedTransform(mus[1], img[1], Mat3::Identity(), 1.0, pnec::common::Pinhole);
    std::printf("UnscentedTransform %.17g %.17g %.17g %d\n", proj[1](0, 0), proj[1](0, 1), proj[1](2, 2),
                (int)(one(0, 0) == proj[1](0, 0)));
  }

  return 0;
}
