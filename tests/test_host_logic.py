"""CPU: the C-ABI library loads and exports what include/pnec_b200.h declares, option
defaults are Ceres', argument validation works without a GPU, generator invariants."""
import ctypes
import os
import re

import numpy as np
import pytest

from pnec_b200 import api, distributed
from pnec_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "pnec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pnec_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = api.load_library()
    declared = _declared_functions()
    assert set(declared) == set(api.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pnec_b200.h but not exported"


def test_version_and_status_strings():
    lib = api.load_library()
    assert lib.pnec_version() >= 1
    for s, name in enumerate(api.STATUS_NAMES):
        text = lib.pnec_status_string(s).decode()
        assert text != "unknown", name
    assert lib.pnec_status_string(99).decode() == "unknown"


def test_default_options_are_ceres_defaults():
    """ceres::Solver::Options() as the reference uses it (pnec_ceres.cc:43-48) +
    Options::regularization_ (pnec_config.h:50)."""
    o = api.default_opts()
    assert o.variant == api.TARGET
    assert o.max_num_iterations == 50
    assert o.max_num_consecutive_invalid_steps == 5
    assert o.jacobi_scaling == 1
    assert o.regularization == 1e-13
    assert o.function_tolerance == 1e-6
    assert o.gradient_tolerance == 1e-10
    assert o.parameter_tolerance == 1e-8
    assert o.initial_trust_region_radius == 1e4
    assert o.max_trust_region_radius == 1e16
    assert o.min_trust_region_radius == 1e-32
    assert o.min_relative_decrease == 1e-3
    assert o.min_lm_diagonal == 1e-6
    assert o.max_lm_diagonal == 1e32


def test_options_struct_matches_oracle_layout():
    import oracle

    o = oracle.default_opts()
    p = api.default_opts()
    for name, _ in api.SolverOpts._fields_:
        assert getattr(o, name) == getattr(p, name), name
    assert ctypes.sizeof(api.SolverOpts) == 4 * 4 + 10 * 8


def test_frame_options_are_the_reference_defaults():
    """pnec::rel_pose_estimation::Options (pnec_config.h:46-65) as PNEC::Solve reads it, RANSAC
    included (on by default, 5000 iterations, sample size 10), plus the literals of the call sites."""
    import oracle

    f = api.default_frame_opts()
    assert (f.use_nec, f.use_ceres, f.weighted_iterations, f.use_ransac) == (0, 1, 10, 1)
    assert (f.max_ransac_iterations, f.ransac_sample_size) == (5000, 10)  # pnec_config.h:59-60
    assert f.ransac_threshold == 1e-6  # pnec.cc:250
    assert f.ransac_probability == 0.99
    assert (f.fibonacci_samples, f.scf_steps) == (500, 10)  # literals of pnec.cc:331 and :342
    assert f.ceres.regularization == 1e-13 and f.ceres.max_num_iterations == 50
    assert ctypes.sizeof(api.FrameOpts) == 6 * 4 + ctypes.sizeof(api.SolverOpts) + 2 * 4 + 5 * 8
    o = oracle.default_frame_opts()
    for name in ("use_nec", "use_ceres", "weighted_iterations", "fibonacci_samples", "scf_steps", "use_ransac"):
        assert getattr(o, name) == getattr(f, name), name
    assert (o.ransac.max_iterations, o.ransac.sample_size, o.ransac.threshold, o.ransac.probability,
            o.ransac.max_variation, o.ransac.seed) == (f.max_ransac_iterations, f.ransac_sample_size,
                                                       f.ransac_threshold, f.ransac_probability,
                                                       f.ransac_max_variation, f.ransac_seed)
    g = api.default_frame_opts(use_ransac=0, weighted_iterations=3, regularization=1e-10)
    assert g.weighted_iterations == 3 and g.ceres.regularization == 1e-10 and g.use_ransac == 0
    with pytest.raises(AttributeError):
        api.default_frame_opts(no_such_field=1)


def test_frame_entry_points_reject_null_arguments():
    lib = api.load_library()
    assert lib.pnec_frame_solve_batch(None, None, None, None, None) == -1
    assert lib.pnec_ransac_batch(None, None, None, 0, None, None, None, None, None) == -1
    assert lib.pnec_keypoints_to_batch(None, None, None, None) == -1
    assert lib.pnec_solve_from_keypoints_batch(None, None, None, None, None) == -1
    assert lib.pnec_frame_solve_from_keypoints_batch(None, None, None, None, None) == -1
    assert lib.pnec_eigensolver_batch(None, None, None, 0.0, None, None, None, None) == -1
    lib.pnec_frame_opts_default(None)  # must be a no-op


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly when there is no device."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.PnecError, match="no CUDA device"):
        api.Handle(0)


def test_null_arguments_are_rejected():
    lib = api.load_library()
    h = ctypes.c_void_p()
    assert lib.pnec_create(0, None) == -1
    assert b"NULL" in lib.pnec_last_error()
    assert lib.pnec_solve_batch(None, None, None, None, None) == -1
    assert lib.pnec_eval_batch(None, None, 1, 0.0, None, None) == -1
    assert lib.pnec_launch_count(None) == 0
    lib.pnec_destroy(None)  # must be a no-op


# ------------------------------------------------------------------ generator


@pytest.mark.parametrize("camera", [syn.OMNIDIRECTIONAL, syn.PINHOLE])
@pytest.mark.parametrize("noise_type", syn.NOISE_TYPES)
def test_generator_invariants(camera, noise_type):
    b = syn.make_batch(5, 33, seed=2, camera=camera, noise_type=noise_type)
    assert b.bvs_host.shape == (165, 3) and b.covs_target.shape == (165, 9)
    np.testing.assert_allclose(np.linalg.norm(b.bvs_host, axis=1), 1.0, atol=1e-14)
    np.testing.assert_allclose(np.linalg.norm(b.bvs_target, axis=1), 1.0, atol=1e-14)
    S = b.covs_target.reshape(-1, 3, 3)
    np.testing.assert_allclose(S, S.transpose(0, 2, 1), atol=1e-22)
    w = np.linalg.eigvalsh(S)
    assert (w[:, 0] > -1e-18).all() and (w[:, 2] > 0).all()
    np.testing.assert_allclose(np.linalg.norm(b.init_poses[:, :4], axis=1), 1.0, atol=1e-14)
    np.testing.assert_allclose(np.linalg.norm(b.init_poses[:, 4:], axis=1), 1.0, atol=1e-14)
    # start pose within 0.01 rad of the ground truth (sim_common.cc:212-219)
    from conftest import rotation_angle

    assert max(rotation_angle(a, g) for a, g in zip(b.init_poses, b.gt_poses)) <= 0.01 + 1e-12


def test_generator_ragged_layout():
    counts = np.array([3, 0, 17, 1, 64])
    b = syn.make_batch(5, 0, seed=4, counts=counts)
    assert b.offsets.tolist() == [0, 3, 3, 20, 21, 85]
    assert b.total == 85 and b.range(2) == (3, 20)
    same = syn.make_batch(5, 0, seed=4, counts=counts)
    assert np.array_equal(b.bvs_target, same.bvs_target)


def test_kitti_like_counts():
    c = syn.kitti_like_counts()
    assert c.shape == (4540,) and c.min() >= 500 and c.max() <= 3500
    assert 1900 < c.mean() < 2100


def test_quaternion_matrix_round_trip():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((200, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    R = syn.quaternion_to_matrix(q)
    q2 = syn.matrix_to_quaternion(R)
    sign = np.sign(np.sum(q * q2, axis=1, keepdims=True))
    np.testing.assert_allclose(q2 * sign, q, atol=1e-14)


# ------------------------------------------------------------------- sharding


def test_shard_bounds_uniform():
    assert distributed.shard_bounds(10, 1) == [(0, 10)]
    assert distributed.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert distributed.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert distributed.shard_bounds(100000, 8)[3] == (37500, 50000)


def test_shard_bounds_ragged_balances_correspondences():
    counts = syn.kitti_like_counts(num_pairs=500)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    for world in (2, 4, 8):
        bounds = distributed.shard_bounds(500, world, offsets)
        assert bounds[0][0] == 0 and bounds[-1][1] == 500
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        loads = [offsets[b1] - offsets[b0] for b0, b1 in bounds]
        assert max(loads) - min(loads) <= 2 * counts.max()


def test_shard_arrays_rebases_offsets():
    offsets = np.array([0, 2, 5, 9, 10])
    f1 = np.arange(30.0).reshape(10, 3)
    poses = np.arange(28.0).reshape(4, 7)
    lf1, lposes, loff = distributed.shard_arrays((1, 3), 0, offsets, f1, poses=poses)
    assert loff.tolist() == [0, 3, 7]
    assert np.array_equal(lf1, f1[2:9]) and np.array_equal(lposes, poses[1:3])
    lf1, lposes, loff = distributed.shard_arrays((2, 4), 5, None, np.arange(60.0).reshape(20, 3), poses=poses)
    assert loff is None and lf1.shape == (10, 3)


def test_compat_header_compiles_against_sophus_style_call_sites(tmp_path):
    """include/pnec/pnec_compat.hpp with the reference's own types at the call site (Sophus::SE3d,
    std::vector<Eigen::Vector3d>): Eigen / Sophus are absent from the image, so minimal mock headers
    (tests/cpp/mock) stand in for them; this only proves the interop overloads compile and link."""
    import subprocess

    lib = os.path.join(ROOT, "pnec_b200", "lib")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    for src, extra in (("compat_interop_test.cpp", ["-I", os.path.join(ROOT, "tests", "cpp", "mock")]),
                       ("compat_test.cpp", ["-pthread"])):
        subprocess.run([cxx, "-std=c++17", "-O0", "-Wall", "-I", os.path.join(ROOT, "include"), *extra,
                        os.path.join(ROOT, "tests", "cpp", src), "-o", str(tmp_path / src.replace(".cpp", "")),
                        "-L", lib, "-lpnec_b200", f"-Wl,-rpath,{lib}"], check=True)
