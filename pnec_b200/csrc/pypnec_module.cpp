// pypnec — the reference's Python binding (python/pypnec.cpp:50-82, 244-257) over the
// B200 C-ABI.  Same module name, function names, positional argument order and return
// shapes:
//   pypnec.pyceres(host_bvs, target_bvs, host_covariances, target_covariances,
//                  init_pose(4x4), regularization) -> (4,4)      [PNECSymmetrical]
//   pypnec.pyceresnec(host_bvs, target_bvs, init_pose(4x4)) -> (4,4)
// Inputs may be lists of (3,)/(3,3) arrays (what the reference's Eigen casters take)
// or (N,3)/(N,3,3) ndarrays.  Batched extensions: pyceres_batch / pyceresnec_batch /
// pyceres_target_batch.  The KLT / demo helpers of the reference module (add, mat2,
// matrices, cppimg, KLTMatching, KLTImageMatching) are image-side and out of scope.
#include <algorithm>
#include <cstdint>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <cstring>
#include <stdexcept>
#include <vector>

#include "pnec/pnec_compat.hpp"

namespace py = pybind11;
using arr = py::array_t<double, py::array::c_style | py::array::forcecast>;

namespace {

// (N,3) contiguous doubles from a list of (3,) arrays or an (N,3) ndarray
arr as_vectors(const py::object &o, const char *what) {
  arr a = arr::ensure(o);
  if (!a) throw std::invalid_argument(std::string(what) + ": cannot convert to float64 array");
  if (a.ndim() == 1 && a.shape(0) == 0) return arr(std::vector<py::ssize_t>{0, 3});
  if (a.ndim() != 2 || a.shape(1) != 3)
    throw std::invalid_argument(std::string(what) + ": expected N vectors of length 3");
  return a;
}

// (N,3,3) row-indexed numpy matrices -> ABI layout: N x 9 column-major
std::vector<double> as_covariances(const py::object &o, py::ssize_t n, const char *what) {
  arr a = arr::ensure(o);
  if (!a) throw std::invalid_argument(std::string(what) + ": cannot convert to float64 array");
  if (n == 0) return {};
  if (a.ndim() != 3 || a.shape(0) != n || a.shape(1) != 3 || a.shape(2) != 3)
    throw std::invalid_argument(std::string(what) + ": expected N matrices of shape (3,3)");
  std::vector<double> out(static_cast<size_t>(n) * 9);
  auto r = a.unchecked<3>();
  for (py::ssize_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c)
      for (int rr = 0; rr < 3; ++rr) out[i * 9 + c * 3 + rr] = r(i, rr, c);
  return out;
}

// init_pose handling of python/pypnec.cpp:56-61: quaternion of the 3x3 block, normalised,
// to a rotation matrix, to Sophus::SE3d, and back to a quaternion for InitValues.
pnec::SE3 pose_from_matrix(const py::object &o) {
  arr a = arr::ensure(o);
  if (!a || a.ndim() != 2 || a.shape(0) != 4 || a.shape(1) != 4)
    throw std::invalid_argument("init_pose: expected a 4x4 matrix");
  auto r = a.unchecked<2>();
  pnec::Mat3 R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R(i, j) = r(i, j);
  const pnec::Mat3 Rn = pnec::Quat::FromRotationMatrix(R).normalized().toRotationMatrix();
  return pnec::SE3(Rn, pnec::Vec3(r(0, 3), r(1, 3), r(2, 3)));
}

py::array_t<double> matrix_from_pose(const pnec::SE3 &p) {
  py::array_t<double> out({4, 4});
  const auto m = p.matrix();
  std::memcpy(out.mutable_data(), m.data(), 16 * sizeof(double));
  return out;
}

struct Span3 {  // minimal .data()/.size() view over (N,3) doubles
  using value_type = pnec::Vec3;
  const pnec::Vec3 *p;
  size_t n;
  const pnec::Vec3 *data() const { return p; }
  size_t size() const { return n; }
};
struct Span9 {
  using value_type = pnec::Mat3;
  const pnec::Mat3 *p;
  size_t n;
  const pnec::Mat3 *data() const { return p; }
  size_t size() const { return n; }
};

py::array_t<double> pyceres(const py::object &host_bvs, const py::object &target_bvs,
                            const py::object &host_covariances, const py::object &target_covariances,
                            const py::object &init_pose, double regularization) {
  arr f1 = as_vectors(host_bvs, "host_bvs"), f2 = as_vectors(target_bvs, "target_bvs");
  const py::ssize_t n = f1.shape(0);
  if (f2.shape(0) != n) throw std::invalid_argument("host_bvs and target_bvs differ in length");
  std::vector<double> c1 = as_covariances(host_covariances, n, "host_covariances");
  std::vector<double> c2 = as_covariances(target_covariances, n, "target_covariances");
  const pnec::SE3 sp_init_pose = pose_from_matrix(init_pose);
  pnec::optimization::PNECCeres optimizer;
  optimizer.InitValues(pnec::Quat::FromRotationMatrix(sp_init_pose.rotationMatrix()),
                       sp_init_pose.translation());
  Span3 s1{reinterpret_cast<const pnec::Vec3 *>(f1.data()), static_cast<size_t>(n)};
  Span3 s2{reinterpret_cast<const pnec::Vec3 *>(f2.data()), static_cast<size_t>(n)};
  Span9 m1{reinterpret_cast<const pnec::Mat3 *>(c1.data()), static_cast<size_t>(n)};
  Span9 m2{reinterpret_cast<const pnec::Mat3 *>(c2.data()), static_cast<size_t>(n)};
  {
    py::gil_scoped_release release;
    optimizer.Optimize(s1, s2, m1, m2, regularization);
  }
  return matrix_from_pose(optimizer.Result());
}

py::array_t<double> pyceresnec(const py::object &host_bvs, const py::object &target_bvs,
                               const py::object &init_pose) {
  arr f1 = as_vectors(host_bvs, "host_bvs"), f2 = as_vectors(target_bvs, "target_bvs");
  const py::ssize_t n = f1.shape(0);
  if (f2.shape(0) != n) throw std::invalid_argument("host_bvs and target_bvs differ in length");
  const pnec::SE3 sp_init_pose = pose_from_matrix(init_pose);
  pnec::optimization::NECCeres optimizer;
  optimizer.InitValues(pnec::Quat::FromRotationMatrix(sp_init_pose.rotationMatrix()),
                       sp_init_pose.translation());
  Span3 s1{reinterpret_cast<const pnec::Vec3 *>(f1.data()), static_cast<size_t>(n)};
  Span3 s2{reinterpret_cast<const pnec::Vec3 *>(f2.data()), static_cast<size_t>(n)};
  {
    py::gil_scoped_release release;
    optimizer.Optimize(s1, s2);
  }
  return matrix_from_pose(optimizer.Result());
}

// Addition to the reference's module: the whole frame solve, PNEC::Solve
// (src/rel_pose_estimation/pnec.cc:77-124), with the Options fields it reads as keyword arguments
// (defaults as pnec_config.h:46-65, RANSAC included).  Returns (pose 4x4, inlier indices).
py::tuple pysolve_inliers(const py::object &host_bvs, const py::object &target_bvs,
                          const py::object &target_covariances, const py::object &init_pose, double regularization,
                          int weighted_iterations, bool use_nec, bool use_ceres, bool use_ransac,
                          int max_ransac_iterations, int ransac_sample_size, std::uint64_t ransac_seed) {
  arr f1 = as_vectors(host_bvs, "host_bvs"), f2 = as_vectors(target_bvs, "target_bvs");
  const py::ssize_t n = f1.shape(0);
  if (f2.shape(0) != n) throw std::invalid_argument("host_bvs and target_bvs differ in length");
  if (weighted_iterations < 0) throw std::invalid_argument("weighted_iterations < 0");
  std::vector<double> c2 = as_covariances(target_covariances, n, "target_covariances");
  const pnec::SE3 sp_init_pose = pose_from_matrix(init_pose);
  pnec::rel_pose_estimation::Options options;
  options.use_ransac_ = use_ransac;
  options.max_ransac_iterations_ = max_ransac_iterations;
  options.ransac_sample_size_ = ransac_sample_size;
  options.ransac_seed_ = ransac_seed;
  options.use_nec_ = use_nec;
  options.use_ceres_ = use_ceres;
  options.weighted_iterations_ = static_cast<std::size_t>(weighted_iterations);
  options.regularization_ = regularization;
  pnec::rel_pose_estimation::PNEC solver(options);
  Span3 s1{reinterpret_cast<const pnec::Vec3 *>(f1.data()), static_cast<size_t>(n)};
  Span3 s2{reinterpret_cast<const pnec::Vec3 *>(f2.data()), static_cast<size_t>(n)};
  Span9 m2{reinterpret_cast<const pnec::Mat3 *>(c2.data()), static_cast<size_t>(n)};
  pnec::SE3 result;
  std::vector<int> inliers;
  {
    py::gil_scoped_release release;
    result = solver.Solve(s1, s2, m2, sp_init_pose, inliers);
  }
  py::array_t<int> idx(static_cast<py::ssize_t>(inliers.size()));
  std::copy(inliers.begin(), inliers.end(), idx.mutable_data());
  return py::make_tuple(matrix_from_pose(result), idx);
}

py::array_t<double> pysolve(const py::object &host_bvs, const py::object &target_bvs,
                            const py::object &target_covariances, const py::object &init_pose,
                            double regularization, int weighted_iterations, bool use_nec, bool use_ceres,
                            bool use_ransac, int max_ransac_iterations, int ransac_sample_size,
                            std::uint64_t ransac_seed) {
  return pysolve_inliers(host_bvs, target_bvs, target_covariances, init_pose, regularization, weighted_iterations,
                         use_nec, use_ceres, use_ransac, max_ransac_iterations, ransac_sample_size, ransac_seed)[0]
      .cast<py::array_t<double>>();
}

// Batched: bvs (B,N,3), covariances (B,N,3,3) or None, init_poses (B,4,4) -> (B,4,4)
py::array_t<double> solve_batch(int variant, const py::object &host_bvs, const py::object &target_bvs,
                                const py::object &host_covariances, const py::object &target_covariances,
                                const py::object &init_poses, double regularization) {
  arr f1 = arr::ensure(host_bvs), f2 = arr::ensure(target_bvs), ip = arr::ensure(init_poses);
  if (!f1 || !f2 || !ip || f1.ndim() != 3 || f2.ndim() != 3 || f1.shape(2) != 3 || ip.ndim() != 3)
    throw std::invalid_argument("expected bvs of shape (B,N,3) and init_poses of shape (B,4,4)");
  const py::ssize_t B = f1.shape(0), N = f1.shape(1);
  if (f2.shape(0) != B || f2.shape(1) != N || ip.shape(0) != B || ip.shape(1) != 4 || ip.shape(2) != 4)
    throw std::invalid_argument("batch shapes do not match");
  std::vector<double> ct, ch;
  auto covs = [&](const py::object &o, const char *what) {
    arr a = arr::ensure(o);
    if (!a || a.ndim() != 4 || a.shape(0) != B || a.shape(1) != N || a.shape(2) != 3 || a.shape(3) != 3)
      throw std::invalid_argument(std::string(what) + ": expected shape (B,N,3,3)");
    std::vector<double> out(static_cast<size_t>(B * N) * 9);
    auto r = a.unchecked<4>();
    for (py::ssize_t b = 0; b < B; ++b)
      for (py::ssize_t i = 0; i < N; ++i)
        for (int c = 0; c < 3; ++c)
          for (int rr = 0; rr < 3; ++rr) out[((b * N + i) * 9) + c * 3 + rr] = r(b, i, rr, c);
    return out;
  };
  if (variant != PNEC_VARIANT_NEC) ct = covs(target_covariances, "target_covariances");
  if (variant == PNEC_VARIANT_SYMMETRIC) ch = covs(host_covariances, "host_covariances");
  std::vector<double> init(static_cast<size_t>(B) * 7), out(static_cast<size_t>(B) * 7);
  for (py::ssize_t b = 0; b < B; ++b) {
    py::object m = py::reinterpret_borrow<py::object>(ip[py::int_(b)]);
    const pnec::SE3 p0 = pose_from_matrix(m);
    const pnec::Quat q = pnec::Quat::FromRotationMatrix(p0.rotationMatrix());
    for (int k = 0; k < 4; ++k) init[b * 7 + k] = q.c[k];
    for (int k = 0; k < 3; ++k) init[b * 7 + 4 + k] = p0.t[k];
  }
  pnec_solver_opts o;
  pnec_solver_opts_default(&o);
  o.variant = variant;
  o.regularization = regularization;
  pnec_batch bt{};
  bt.num_problems = B;
  bt.n_per_problem = N;
  bt.memspace = PNEC_MEM_HOST;
  bt.bvs_host = f1.data();
  bt.bvs_target = f2.data();
  bt.covs_target = ct.empty() ? nullptr : ct.data();
  bt.covs_host = ch.empty() ? nullptr : ch.data();
  bt.poses = init.data();
  pnec_solve_out so{};
  so.poses = out.data();
  int rc;
  {
    py::gil_scoped_release release;
    rc = pnec_solve_batch(pnec::detail::Handle(), &bt, &o, &so, nullptr);
  }
  if (rc != PNEC_OK) throw std::runtime_error(std::string("pnec_solve_batch: ") + pnec_last_error());
  py::array_t<double> res({B, py::ssize_t(4), py::ssize_t(4)});
  for (py::ssize_t b = 0; b < B; ++b) {
    pnec::SE3 p(pnec::Quat(out[b * 7 + 3], out[b * 7 + 0], out[b * 7 + 1], out[b * 7 + 2]),
                pnec::Vec3(out[b * 7 + 4], out[b * 7 + 5], out[b * 7 + 6]));
    const auto m = p.matrix();
    std::memcpy(res.mutable_data() + b * 16, m.data(), 16 * sizeof(double));
  }
  return res;
}

}  // namespace

PYBIND11_MODULE(pypnec, m) {
  m.doc() = "PNEC frame-pair refinement on B200 (drop-in for tum-vision/pnec's pypnec)";
  m.def("pyceres", &pyceres, py::arg("host_bvs"), py::arg("target_bvs"), py::arg("host_covariances"),
        py::arg("target_covariances"), py::arg("init_pose"), py::arg("regularization"),
        "ceres");  // docstrings as in python/pypnec.cpp:252-253
  m.def("pyceresnec", &pyceresnec, py::arg("host_bvs"), py::arg("target_bvs"), py::arg("init_pose"),
        "ceres nec");
  m.def("pysolve", &pysolve, py::arg("host_bvs"), py::arg("target_bvs"), py::arg("target_covariances"),
        py::arg("init_pose"), py::arg("regularization") = 1.0e-13, py::arg("weighted_iterations") = 10,
        py::arg("use_nec") = false, py::arg("use_ceres") = true, py::arg("use_ransac") = true,
        py::arg("max_ransac_iterations") = 5000, py::arg("ransac_sample_size") = 10, py::arg("ransac_seed") = 1,
        "PNEC::Solve: (RANSAC over the) NEC eigensolver, weighted eigensolver + SCF, refinement");
  m.def("pysolve_inliers", &pysolve_inliers, py::arg("host_bvs"), py::arg("target_bvs"),
        py::arg("target_covariances"), py::arg("init_pose"), py::arg("regularization") = 1.0e-13,
        py::arg("weighted_iterations") = 10, py::arg("use_nec") = false, py::arg("use_ceres") = true,
        py::arg("use_ransac") = true, py::arg("max_ransac_iterations") = 5000, py::arg("ransac_sample_size") = 10,
        py::arg("ransac_seed") = 1, "PNEC::Solve -> (pose, inliers)");
  m.def("pyceres_batch",
        [](const py::object &a, const py::object &b, const py::object &c, const py::object &d,
           const py::object &e, double reg) { return solve_batch(PNEC_VARIANT_SYMMETRIC, a, b, c, d, e, reg); },
        py::arg("host_bvs"), py::arg("target_bvs"), py::arg("host_covariances"),
        py::arg("target_covariances"), py::arg("init_poses"), py::arg("regularization"));
  m.def("pyceres_target_batch",
        [](const py::object &a, const py::object &b, const py::object &d, const py::object &e, double reg) {
          return solve_batch(PNEC_VARIANT_TARGET, a, b, py::none(), d, e, reg);
        },
        py::arg("host_bvs"), py::arg("target_bvs"), py::arg("target_covariances"),
        py::arg("init_poses"), py::arg("regularization"));
  m.def("pyceresnec_batch",
        [](const py::object &a, const py::object &b, const py::object &e) {
          return solve_batch(PNEC_VARIANT_NEC, a, b, py::none(), py::none(), e, 0.0);
        },
        py::arg("host_bvs"), py::arg("target_bvs"), py::arg("init_poses"));
}
