"""GPU (-m gpu): solver inputs built on the device from keypoints -- Frame2Frame::GetFeatures
(src/rel_pose_estimation/frame2frame.cc:359-392) over KeyPoint::Unproject (src/frames/keypoints.cc:49-62)
-- and the solve / frame solve fed from them (pnec_*_from_keypoints_batch).

Exactness: the result equals, bit for bit, unprojecting the same keypoints with
pnec_keypoints_unproject_batch and handing the 120-byte-per-correspondence arrays to the batch entry
points (one out-of-line device function computes a keypoint in both kernels).  Parity: against the
oracle's unprojection + solve within the north-star tolerance."""
import numpy as np
import pytest

import oracle
from conftest import max_pose_diff
from pnec_b200 import api
from pnec_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-6
DIR_TOL = 1e-6
K = np.array([[800.0, 0.0, 320.0], [0.0, 790.0, 240.0], [0.0, 0.0, 1.0]])
K_INV = np.ascontiguousarray(np.linalg.inv(K).T).reshape(9)  # column-major, the C-ABI's convention


@pytest.fixture(scope="module")
def handle():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a B200"
    return api.Handle(0)


def dev(a):
    import torch

    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def keypoint_workload(B, N, seed, counts=None):
    """Pinhole frame pairs as keypoints: pixel positions of both frames and 2x2 image covariances of the
    target frame's keypoints (projected from a synthetic pinhole batch; the covariances are fresh
    draws of the simulator's anisotropic family, src/simulation/standard_experiments.cc:86-123)."""
    batch = syn.make_batch(B, N, seed=seed, camera=syn.PINHOLE, counts=counts)
    rng = np.random.default_rng(seed + 1)
    pix = lambda f: (f[:, :2] / f[:, 2:3]) @ K[:2, :2].T + K[:2, 2]
    hp, tp = pix(batch.bvs_host), pix(batch.bvs_target)
    cov2 = syn.sample_covariances_2d(rng, (1, batch.total), 1.0, "anisotropic_inhomogenous")[0]
    c4 = np.swapaxes(cov2, -1, -2).reshape(-1, 4)  # column-major 2x2
    return batch, hp, tp, c4


def unproject_on_device(handle, pts, c4):
    return handle.keypoints_unproject(pts, c4, K_INV)


@pytest.mark.parametrize("mode", ["host", "device"])
@pytest.mark.parametrize("layout", ["identity", "indexed", "packed", "ragged"])
def test_solve_from_keypoints_is_exact_and_matches_oracle(handle, mode, layout):
    B, N = 40, 150
    counts = np.array([7 + 13 * (i % 11) for i in range(B)], dtype=np.int64) if layout == "ragged" else None
    batch, hp, tp, c4 = keypoint_workload(B, N, 51, counts)
    total = batch.total
    kw = dict(offsets=batch.offsets) if layout == "ragged" else dict(n_per_problem=N)
    # reference route: unproject on the device, then the 120 B / correspondence entry point
    f1, _ = unproject_on_device(handle, hp, c4)
    f2, ct = unproject_on_device(handle, tp, c4)
    ref = handle.solve_batch(f1, f2, ct, None, batch.init_poses, api.default_opts(api.TARGET), **kw)
    host_index = target_index = None
    hp_t, tp_t, c_t, packed = hp, tp, c4, False
    if layout == "indexed":  # shuffled tables with extra unmatched keypoints, as a frame's keypoint list has
        rng = np.random.default_rng(3)
        perm_h, perm_t = rng.permutation(total + 50), rng.permutation(total + 70)
        hp_t = np.zeros((total + 50, 2)); tp_t = np.zeros((total + 70, 2)); c_t = np.ones((total + 70, 4))
        host_index, target_index = perm_h[:total].astype(np.int32), perm_t[:total].astype(np.int32)
        hp_t[host_index], tp_t[target_index], c_t[target_index] = hp, tp, c4
    if layout == "packed":
        assert np.array_equal(c4[:, 1], c4[:, 2])
        c_t, packed = np.ascontiguousarray(c4[:, [0, 1, 3]]), True
    w = dev if mode == "device" else (lambda a: a)
    res = handle.solve_from_keypoints(w(hp_t), w(tp_t), w(c_t), w(batch.init_poses), K_INV, api.default_opts(api.TARGET),
                                      host_index=w(host_index), target_index=w(target_index), packed=packed, **kw)
    got = res.poses.cpu().numpy() if mode == "device" else res.poses
    np.testing.assert_array_equal(got, ref.poses)
    its = res.iterations.cpu().numpy() if mode == "device" else res.iterations
    np.testing.assert_array_equal(its, ref.iterations)
    # parity with the oracle's unprojection + refinement
    o1, _ = oracle.keypoints_unproject(hp, c4, K_INV)
    o2, oc = oracle.keypoints_unproject(tp, c4, K_INV)
    oref, info = oracle.solve_batch(o1, o2, oc, None, batch.init_poses, oracle.default_opts(oracle.TARGET),
                                    num_threads=oracle.max_threads(), **kw)
    well = np.ones(B, bool) if counts is None else counts >= 10
    r, t = max_pose_diff(got[well], oref[well])
    assert r <= ROT_TOL and t <= DIR_TOL, (r, t)


def test_chunked_host_keypoints_agree_bitwise(monkeypatch):
    B, N = 300, 96
    batch, hp, tp, c4 = keypoint_workload(B, N, 52)
    monkeypatch.setenv("PNEC_B200_H2D_CHUNKS", "1")
    one = api.Handle(0).solve_from_keypoints(hp, tp, c4, batch.init_poses, K_INV, n_per_problem=N)
    monkeypatch.setenv("PNEC_B200_H2D_CHUNKS", "4")
    cut = api.Handle(0).solve_from_keypoints(hp, tp, c4, batch.init_poses, K_INV, n_per_problem=N)
    for name in ("poses", "status", "iterations", "cost", "initial_cost"):
        assert np.array_equal(getattr(one, name), getattr(cut, name)), name


def test_symmetric_variant_from_keypoints(handle):
    B, N = 12, 80
    batch, hp, tp, c4 = keypoint_workload(B, N, 53)
    ch4 = np.roll(c4, 5, axis=0)
    f1, ch = unproject_on_device(handle, hp, ch4)
    f2, ct = unproject_on_device(handle, tp, c4)
    ref = handle.solve_batch(f1, f2, ct, ch, batch.init_poses, api.default_opts(api.SYMMETRIC), n_per_problem=N)
    res = handle.solve_from_keypoints(hp, tp, c4, batch.init_poses, K_INV, api.default_opts(api.SYMMETRIC),
                                      host_covs2=ch4, n_per_problem=N)
    np.testing.assert_array_equal(res.poses, ref.poses)
    with pytest.raises(api.PnecError):
        handle.solve_from_keypoints(hp, tp, c4, batch.init_poses, K_INV, api.default_opts(api.SYMMETRIC), n_per_problem=N)


@pytest.mark.parametrize("mode", ["host", "device"])
def test_frame_solve_from_keypoints(handle, mode):
    """PNEC::Solve with default options (RANSAC on) fed from keypoints == fed from the unprojected arrays."""
    B, N = 48, 200
    batch, hp, tp, c4 = keypoint_workload(B, N, 54)
    f1, _ = unproject_on_device(handle, hp, c4)
    f2, ct = unproject_on_device(handle, tp, c4)
    ref = handle.frame_solve_batch(f1, f2, ct, batch.init_poses, api.default_frame_opts(), n_per_problem=N)
    w = dev if mode == "device" else (lambda a: a)
    res = handle.frame_solve_from_keypoints(w(hp), w(tp), w(c4), w(batch.init_poses), K_INV, api.default_frame_opts(),
                                            n_per_problem=N)
    g = (lambda x: x.cpu().numpy()) if mode == "device" else (lambda x: x)
    np.testing.assert_array_equal(g(res.poses), ref.poses)
    np.testing.assert_array_equal(g(res.num_inliers), ref.num_inliers)
    np.testing.assert_array_equal(g(res.es_poses), ref.es_poses)
    assert (g(res.num_inliers) > 0.5 * N).all()


def test_keypoint_batch_validation(handle):
    batch, hp, tp, c4 = keypoint_workload(4, 20, 55)
    with pytest.raises(api.PnecError):  # table shorter than the batch without an index
        handle.solve_from_keypoints(hp[:-5], tp, c4, batch.init_poses, K_INV, n_per_problem=20)
    bad = np.arange(80, dtype=np.int32)
    bad[3] = 999
    with pytest.raises(api.PnecError):
        handle.solve_from_keypoints(hp, tp, c4, batch.init_poses, K_INV, host_index=bad, n_per_problem=20)
