"""GPU (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle.

Tolerances (BASELINE.json north_star): rotation within 1e-6 rad, translation direction
within 1e-6 rad (modulo sign) of the Ceres-semantics oracle on identical inputs; the
fused evaluation (cost, J^T r, J^T J) within 1e-9 relative of the oracle's closed form.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import direction_angle, max_pose_diff, rotation_angle
from pnec_b200 import api
from pnec_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROT_TOL = 1e-6  # rad
DIR_TOL = 1e-6  # rad
EVAL_RTOL = 1e-9

VARIANTS = {"nec": api.NEC, "target": api.TARGET, "host": api.HOST, "symmetric": api.SYMMETRIC}
CASES = ["c1_iso_omni_n100", "c2_aniso_omni_n512", "aniso_pinhole_n64", "aniso_omni_n10"]


@pytest.fixture(scope="module")
def handle():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a B200"
    return api.Handle(0)


def dev(a):
    import torch

    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def covs_for(variant, ct, ch):
    return (None if variant == api.NEC else ct), (ch if variant == api.SYMMETRIC else None)


def oracle_solve(variant, f1, f2, ct, ch, init, **kw):
    return oracle.solve_batch(f1, f2, ct, ch, init, oracle.default_opts(variant),
                              num_threads=oracle.max_threads(), **kw)


# ------------------------------------------------------------------ fixtures


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("vname", list(VARIANTS))
def test_solve_matches_committed_oracle_solutions(handle, golden_solutions, name, vname):
    g = golden_solutions
    variant = VARIANTS[vname]
    n = int(g[f"{name}/n"])
    ct, ch = covs_for(variant, g[f"{name}/cov_t"], g[f"{name}/cov_h"])
    res = handle.solve_batch(g[f"{name}/f1"], g[f"{name}/f2"], ct, ch, g[f"{name}/init"],
                             api.default_opts(variant), n_per_problem=n)
    r, t = max_pose_diff(res.poses, g[f"{name}/{vname}/poses"])
    assert r <= ROT_TOL and t <= DIR_TOL, (r, t)
    assert np.array_equal(res.iterations, g[f"{name}/{vname}/iterations"])
    assert np.array_equal(res.status, g[f"{name}/{vname}/status"])
    np.testing.assert_allclose(res.cost, g[f"{name}/{vname}/final_cost"], rtol=1e-9)
    np.testing.assert_allclose(res.initial_cost, g[f"{name}/{vname}/initial_cost"], rtol=1e-11)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("vname", list(VARIANTS))
def test_eval_matches_committed_oracle_evaluations(handle, golden_solutions, name, vname):
    g = golden_solutions
    variant = VARIANTS[vname]
    n = int(g[f"{name}/n"])
    ct, ch = covs_for(variant, g[f"{name}/cov_t"], g[f"{name}/cov_h"])
    ev = handle.eval_batch(g[f"{name}/f1"], g[f"{name}/f2"], ct, ch, g[f"{name}/init"], variant,
                           1e-13, n_per_problem=n)
    # committed evaluations use central differences (what Ceres does): ~1e-8 agreement
    np.testing.assert_allclose(ev.cost, g[f"{name}/{vname}/eval_cost"], rtol=1e-12)
    G, H = g[f"{name}/{vname}/eval_gradient"], g[f"{name}/{vname}/eval_jtj"]
    for i in range(G.shape[0]):
        np.testing.assert_allclose(ev.gradient[i], G[i], rtol=0, atol=5e-8 * np.abs(G[i]).max())
        np.testing.assert_allclose(ev.jtj[i], H[i], rtol=0, atol=5e-8 * np.abs(H[i]).max())


# ------------------------------------------------------- fresh seeded inputs


@pytest.mark.parametrize("vname", list(VARIANTS))
@pytest.mark.parametrize("B,N,camera", [(48, 100, syn.OMNIDIRECTIONAL), (12, 512, syn.OMNIDIRECTIONAL),
                                        (40, 37, syn.PINHOLE)])
def test_solve_and_eval_against_oracle(handle, vname, B, N, camera):
    variant = VARIANTS[vname]
    b = syn.with_host_covariances(syn.make_batch(B, N, seed=100 + N, camera=camera), camera=camera)
    ct, ch = covs_for(variant, b.covs_target, b.covs_host)
    res = handle.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(ct), dev(ch), dev(b.init_poses),
                             api.default_opts(variant), n_per_problem=N)
    ref, info = oracle_solve(variant, b.bvs_host, b.bvs_target, ct, ch, b.init_poses, n_per_problem=N)
    poses = res.poses.cpu().numpy()
    r, t = max_pose_diff(poses, ref)
    assert r <= ROT_TOL and t <= DIR_TOL, (r, t)
    assert np.array_equal(res.iterations.cpu().numpy(), info["iterations"])
    assert np.array_equal(res.status.cpu().numpy(), info["status"])
    np.testing.assert_allclose(res.cost.cpu().numpy(), info["final_cost"], rtol=1e-9)
    # fused evaluation vs the oracle's closed form
    ev = handle.eval_batch(dev(b.bvs_host), dev(b.bvs_target), dev(ct), dev(ch), dev(b.init_poses),
                           variant, 1e-13, n_per_problem=N)
    for i in range(min(B, 6)):
        s, e = b.range(i)
        o = oracle.evaluate(variant, b.bvs_host[s:e], b.bvs_target[s:e], None if ct is None else ct[s:e],
                            None if ch is None else ch[s:e], 1e-13, b.init_poses[i], oracle.JAC_ANALYTIC)
        assert ev.cost[i].item() == pytest.approx(o.cost, rel=EVAL_RTOL)
        np.testing.assert_allclose(ev.gradient[i].cpu().numpy(), o.gradient, rtol=0,
                                   atol=EVAL_RTOL * np.abs(o.gradient).max())
        np.testing.assert_allclose(ev.jtj[i].cpu().numpy(), o.jtj, rtol=0, atol=EVAL_RTOL * np.abs(o.jtj).max())


@pytest.mark.parametrize("path", ["stream", "resident"])
def test_ragged_batch_with_edge_sizes(handle, path, monkeypatch):
    """Empty problems, single correspondences, odd offsets, sizes around tile and
    resident-capacity boundaries, one problem larger than shared memory.  Run through both
    solve kernels: the streaming one (default once a pair exceeds 896 correspondences) and
    the shared-memory-resident one with its global-memory fallback for oversized pairs."""
    counts = np.array([0, 1, 2, 3, 5, 0, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 257, 511, 513,
                       1000, 1023, 1887, 1889, 1890, 1891, 2500, 0, 77, 4100, 6])
    monkeypatch.setenv("PNEC_B200_STREAM_MIN_N", "896" if path == "stream" else "100000000")
    handle = api.Handle(0)  # the switches are read when a handle is created
    B = len(counts)
    b = syn.make_batch(B, 0, seed=77, counts=counts)
    res = handle.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), None,
                             dev(b.init_poses), api.default_opts(api.TARGET), offsets=b.offsets)
    ref, info = oracle_solve(api.TARGET, b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses,
                             offsets=b.offsets)
    poses = res.poses.cpu().numpy()
    status = res.status.cpu().numpy()
    iters = res.iterations.cpu().numpy()
    cost = res.cost.cpu().numpy()
    assert (status[counts == 0] == 7).all() and (iters[counts == 0] == 0).all()
    # fewer than 5 correspondences leave the 5-DoF problem rank deficient: the LM path is then
    # decided by rounding noise, so only well-posed problems are compared step for step
    well = counts >= 10
    assert np.array_equal(status[well], info["status"][well])
    assert np.array_equal(iters[well], info["iterations"][well])
    r, t = max_pose_diff(poses[well], ref[well])
    assert r <= ROT_TOL and t <= DIR_TOL, (r, t)
    np.testing.assert_allclose(cost[well], info["final_cost"][well], rtol=1e-9)
    ill = (counts > 0) & ~well
    assert np.isfinite(poses[ill]).all() and (status[ill] <= 5).all()
    assert (cost[ill] <= res.initial_cost.cpu().numpy()[ill]).all()
    # empty problems return the start pose (normalised)
    for i in np.nonzero(counts == 0)[0]:
        assert rotation_angle(poses[i], b.init_poses[i]) < 1e-12
        assert direction_angle(poses[i][4:], b.init_poses[i][4:], False) < 1e-7
    ev = handle.eval_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), None,
                           dev(b.init_poses), api.TARGET, 1e-13, offsets=b.offsets)
    cost = ev.cost.cpu().numpy()
    np.testing.assert_allclose(cost[counts > 0], info["initial_cost"][counts > 0], rtol=1e-11)
    assert (cost[counts == 0] == 0).all() and (ev.jtj.cpu().numpy()[counts == 0] == 0).all()


def test_host_and_device_memspace_agree_bitwise(handle):
    b = syn.make_batch(20, 96, seed=5)
    o = api.default_opts(api.TARGET)
    host = handle.solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses, o, n_per_problem=96)
    devr = handle.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), None,
                              dev(b.init_poses), o, n_per_problem=96)
    assert np.array_equal(host.poses, devr.poses.cpu().numpy())
    assert np.array_equal(host.iterations, devr.iterations.cpu().numpy())
    again = handle.solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses, o, n_per_problem=96)
    assert np.array_equal(host.poses, again.poses), "solve is not deterministic"


@pytest.mark.parametrize("ragged", [False, True])
@pytest.mark.parametrize("chunks", [2, 4, 8])
def test_chunked_host_path_agrees_bitwise(handle, monkeypatch, ragged, chunks):
    """HOST batches above 64 MB are cut into chunks on separate streams (H2D / solve / D2H overlap);
    forced here on small batches, uniform and ragged (with empty and odd-sized pairs)."""
    counts = np.array([5, 40, 1, 0, 33, 64, 12, 100, 7, 0, 51, 2, 96], dtype=np.int64) if ragged else None
    B = len(counts) if ragged else 13
    b = syn.make_batch(B, 40, seed=17, counts=counts)
    o = api.default_opts(api.TARGET)
    kw = dict(offsets=b.offsets) if ragged else dict(n_per_problem=40)
    # (the switches are read when a handle is created)
    monkeypatch.setenv("PNEC_B200_H2D_CHUNKS", "1")
    one = api.Handle(0).solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses, o, **kw)
    monkeypatch.setenv("PNEC_B200_H2D_CHUNKS", str(chunks))
    cut = api.Handle(0).solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses, o, **kw)
    for name in ("poses", "status", "iterations", "cost", "initial_cost"):
        assert np.array_equal(getattr(one, name), getattr(cut, name)), name


def test_pageable_host_batch_through_the_staging_ring(handle, monkeypatch):
    """A 72 MB batch in pageable (numpy) memory: chunked streams + pinned staging ring with copy
    threads; the results must equal the device-resident call bit for bit, with and without the ring."""
    B, N = 3000, 200
    b = syn.make_batch(B, N, seed=23)
    o = api.default_opts(api.TARGET)
    devr = handle.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), None, dev(b.init_poses), o,
                              n_per_problem=N)
    host = handle.solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses, o, n_per_problem=N)
    monkeypatch.setenv("PNEC_B200_NO_STAGER", "1")
    plain = api.Handle(0).solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses, o, n_per_problem=N)
    for r in (host, plain):
        assert np.array_equal(r.poses, devr.poses.cpu().numpy())
        assert np.array_equal(r.iterations, devr.iterations.cpu().numpy())


def test_unaligned_device_pointers_take_the_plain_copy_path(handle):
    """Base pointers that are 8- but not 16-byte aligned cannot use cp.async.bulk."""
    import torch

    b = syn.make_batch(6, 70, seed=6)

    def shifted(a):
        flat = torch.empty(a.size + 1, dtype=torch.float64, device="cuda")
        view = flat[1:].view(a.shape)
        view.copy_(torch.from_numpy(a))
        assert view.data_ptr() % 16 == 8
        return view

    res = handle.solve_batch(shifted(b.bvs_host), shifted(b.bvs_target), shifted(b.covs_target), None,
                             dev(b.init_poses), api.default_opts(api.TARGET), n_per_problem=70)
    ref = handle.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), None,
                             dev(b.init_poses), api.default_opts(api.TARGET), n_per_problem=70)
    assert np.array_equal(res.poses.cpu().numpy(), ref.poses.cpu().numpy())
    ev = handle.eval_batch(shifted(b.bvs_host), shifted(b.bvs_target), shifted(b.covs_target), None,
                           dev(b.init_poses), api.TARGET, 1e-13, n_per_problem=70)
    ev2 = handle.eval_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), None,
                            dev(b.init_poses), api.TARGET, 1e-13, n_per_problem=70)
    np.testing.assert_allclose(ev.jtj.cpu().numpy(), ev2.jtj.cpu().numpy(), rtol=1e-13)


def test_solver_options_are_honoured(handle):
    b = syn.make_batch(16, 80, seed=8)
    for kw in (dict(max_num_iterations=1), dict(function_tolerance=1e-12, max_num_iterations=8),
               dict(initial_trust_region_radius=1e-2), dict(jacobi_scaling=0)):
        res = handle.solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses,
                                 api.default_opts(api.TARGET, **kw), n_per_problem=80)
        ref, info = oracle.solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses,
                                       oracle.default_opts(oracle.TARGET, **kw), n_per_problem=80)
        r, t = max_pose_diff(res.poses, ref)
        assert r <= ROT_TOL and t <= DIR_TOL, (kw, r, t)
        assert np.array_equal(res.iterations, info["iterations"]), kw
        assert np.array_equal(res.status, info["status"]), kw


def test_argument_errors(handle):
    b = syn.make_batch(2, 10, seed=1)
    with pytest.raises(api.PnecError, match="covs_target"):
        handle.solve_batch(b.bvs_host, b.bvs_target, None, None, b.init_poses, api.default_opts(api.TARGET),
                           n_per_problem=10)
    with pytest.raises(api.PnecError, match="covs_host"):
        handle.solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses,
                           api.default_opts(api.SYMMETRIC), n_per_problem=10)
    with pytest.raises(api.PnecError, match="variant"):
        handle.solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses,
                           api.default_opts(9), n_per_problem=10)
    with pytest.raises(api.PnecError, match="non-decreasing"):
        handle.solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses,
                           api.default_opts(api.TARGET), offsets=np.array([0, 12, 8]))


def test_cost_function_metric(handle):
    b = syn.make_batch(5, 60, seed=12)
    got = handle.cost_function_batch(b.bvs_host, b.bvs_target, b.covs_target, b.gt_poses, n_per_problem=60)
    for i in range(5):
        f1, f2, ct, _ = b.problem(i)
        assert got[i] == pytest.approx(oracle.cost_function(f1, f2, ct, b.gt_poses[i]), rel=1e-11)


@pytest.mark.parametrize("camera", [api.CAMERA_OMNIDIRECTIONAL, api.CAMERA_PINHOLE])
def test_unscented_transform_matches_oracle(handle, golden_ut, camera):
    """pnec_unscented_transform_batch vs the C restatement of common.cc:467-525 (fp64, 1e-12
    of the largest entry) and vs the reference's python outputs on the committed vectors."""
    rng = np.random.default_rng(17)
    n = 5000
    col = lambda M: np.swapaxes(M, -1, -2).reshape(-1, 9)
    c3 = np.zeros((n, 3, 3))
    c3[:, :2, :2] = syn.sample_covariances_2d(rng, (1, n), 1.5, "anisotropic_inhomogenous")[0]
    if camera == api.CAMERA_OMNIDIRECTIONAL:
        mus = syn._uniform_sphere(rng, (n,)) * 800.0
        mus[:, 2] = np.abs(mus[:, 2]) + 1.0
        ez = np.broadcast_to(np.array([0.0, 0.0, 1.0]), mus.shape)
        Rp = syn.rotation_between_points(ez, syn._normalize(mus))
        c3 = Rp @ c3 @ np.swapaxes(Rp, -1, -2)
        K = None
    else:
        mus = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n), np.ones(n)], -1)
        Kinv = np.linalg.inv(np.array([[718.856, 0, 607.19], [0, 718.856, 185.2157], [0, 0, 1.0]]))
        K = np.ascontiguousarray(Kinv.T).reshape(9)  # column-major
    ref = oracle.unscented_transform(mus, col(c3), K, 1.0, camera)
    got_h = handle.unscented_transform(mus, col(c3), K, 1.0, camera)
    got_d = handle.unscented_transform(dev(mus), dev(col(c3)), K, 1.0, camera).cpu().numpy()
    assert np.array_equal(got_h, got_d)
    np.testing.assert_allclose(got_h, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    assert np.abs(got_h - got_h.reshape(-1, 3, 3).transpose(0, 2, 1).reshape(-1, 9)).max() < 1e-20
    u = golden_ut
    if camera == api.CAMERA_OMNIDIRECTIONAL:
        g = handle.unscented_transform(u["mus"], col(u["covs_omni"]), None, 1.0, camera)
        np.testing.assert_allclose(g, col(u["ut_omni"]), rtol=0, atol=1e-12 * np.abs(u["ut_omni"]).max())
    else:
        g = handle.unscented_transform(u["mus_pinhole"], col(u["covs_local"]), None, 1.0, camera)
        np.testing.assert_allclose(g, col(u["ut_pinhole"]), rtol=0, atol=1e-12 * np.abs(u["ut_pinhole"]).max())
    assert handle.unscented_transform(np.zeros((0, 3)), np.zeros((0, 9))).shape == (0, 9)
    with pytest.raises(api.PnecError, match="camera model"):
        handle.unscented_transform(mus, col(c3), K, 1.0, 5)


def test_keypoints_unproject_matches_oracle(handle):
    """pnec_keypoints_unproject_batch == KeyPoint::Unproject (keypoints.cc:49-62) restated in
    the oracle; full TMA tiles, a ragged tail and unaligned device pointers."""
    import torch

    rng = np.random.default_rng(23)
    n = 128 * 37 + 51
    pts = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n)], -1)
    c2 = np.swapaxes(syn.sample_covariances_2d(rng, (1, n), 0.8, "anisotropic_inhomogenous")[0], -1, -2).reshape(n, 4)
    Kinv = np.linalg.inv(np.array([[718.856, 0, 607.19], [0, 718.856, 185.2157], [0, 0, 1.0]])).T.reshape(9)
    rb, rc = oracle.keypoints_unproject(pts, c2, Kinv)
    hb, hc = handle.keypoints_unproject(pts, c2, Kinv)
    db, dc = handle.keypoints_unproject(dev(pts), dev(c2), Kinv)
    assert np.array_equal(hb, db.cpu().numpy()) and np.array_equal(hc, dc.cpu().numpy())
    np.testing.assert_allclose(hb, rb, rtol=0, atol=1e-15)
    np.testing.assert_allclose(hc, rc, rtol=0, atol=1e-12 * np.abs(rc).max())
    flat = torch.empty(2 * n + 1, dtype=torch.float64, device="cuda")
    shifted = flat[1:].view(n, 2)
    shifted.copy_(torch.from_numpy(pts))
    ub, uc = handle.keypoints_unproject(shifted, dev(c2), Kinv)  # plain (non-TMA) path
    np.testing.assert_allclose(ub.cpu().numpy(), hb, rtol=0, atol=1e-15)
    np.testing.assert_allclose(uc.cpu().numpy(), hc, rtol=0, atol=1e-13 * np.abs(hc).max())
    # and the produced inputs drive the solver: bearing covariances are symmetric PSD, (numerically) rank 2
    w = np.linalg.eigvalsh(hc.reshape(n, 3, 3))
    assert (w[:, 2] > 0).all() and (np.abs(w[:, 0]) < 1e-5 * w[:, 2]).all()


def test_translation_given_rotation_matches_oracle(handle):
    """pnec_scf_translation_batch / pnec_nec_translation_batch vs the oracle's restatement of
    pnec.cc:317-343 + scf.cc and common.cc:127-181: direction within 1e-6 rad modulo sign (the
    eigenvector sign is arbitrary in the reference too)."""
    counts = np.array([40, 100, 257, 512, 33, 1000, 64, 5])
    b = syn.make_batch(len(counts), 0, seed=91, counts=counts)
    # start translations away from the optimum so that the Fibonacci scan matters
    poses = b.gt_poses.copy()
    poses[:, 4:] = syn._normalize(np.random.default_rng(2).standard_normal((len(counts), 3)))
    t, c = handle.scf_translation_batch(b.bvs_host, b.bvs_target, b.covs_target, poses, 1e-13, 500, 10,
                                        offsets=b.offsets)
    td, cd = handle.scf_translation_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), dev(poses),
                                          1e-13, 500, 10, offsets=b.offsets)
    assert np.array_equal(t, td.cpu().numpy()) and np.array_equal(c, cd.cpu().numpy())
    tn, M = handle.nec_translation_batch(b.bvs_host, b.bvs_target, poses, offsets=b.offsets)
    for i in range(len(counts)):
        f1, f2, ct, _ = b.problem(i)
        rt, rc = oracle.scf_translation(f1, f2, ct, poses[i], 1e-13, 500, 10)
        assert direction_angle(t[i], rt) <= DIR_TOL, i
        assert c[i] == pytest.approx(rc, rel=1e-9)
        rn, rM = oracle.nec_translation(f1, f2, poses[i])
        np.testing.assert_allclose(M[i], rM, rtol=0, atol=1e-12 * np.abs(rM).max())
        if counts[i] >= 10:
            assert direction_angle(tn[i], rn) <= DIR_TOL, i
    # fewer scan samples / steps are honoured
    t0, _ = handle.scf_translation_batch(b.bvs_host, b.bvs_target, b.covs_target, poses, 1e-13, 0, 0,
                                         offsets=b.offsets)
    np.testing.assert_array_equal(t0, poses[:, 4:])
    t3, _ = handle.scf_translation_batch(b.bvs_host, b.bvs_target, b.covs_target, poses, 1e-10, 50, 3,
                                         offsets=b.offsets)
    r3, _ = oracle.scf_translation(*b.problem(3)[:3], poses[3], 1e-10, 50, 3)
    assert direction_angle(t3[3], r3) <= DIR_TOL
    # pairs beyond the shared-memory capacity spill to device memory (tests/test_gpu_frame.py covers parity)
    big = syn.make_batch(1, 4000, seed=1)
    tb, _ = handle.scf_translation_batch(big.bvs_host, big.bvs_target, big.covs_target, big.gt_poses, n_per_problem=4000)
    assert np.isfinite(tb).all() and abs(np.linalg.norm(tb[0]) - 1.0) < 1e-12


# -------------------------------------- BASELINE full size: structural properties


@pytest.fixture(scope="module")
def c2_batch():
    """BASELINE config C2: 10 000 frame pairs x 512 correspondences, anisotropic."""
    return syn.make_batch(10000, 512, seed=2024, noise_type="anisotropic_inhomogenous")


def test_c2_full_size_properties(handle, c2_batch):
    b = c2_batch
    B, N = 10000, 512
    f1, f2, ct, init = dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), dev(b.init_poses)
    o = api.default_opts(api.TARGET)
    res = handle.solve_batch(f1, f2, ct, None, init, o, n_per_problem=N)
    poses, status = res.poses.cpu().numpy(), res.status.cpu().numpy()
    cost, cost0 = res.cost.cpu().numpy(), res.initial_cost.cpu().numpy()
    # 1. every problem terminates, the cost never increases, output is a unit pose
    assert (status <= 4).all() and (status <= 3).mean() > 0.999
    assert (cost <= cost0 * (1 + 1e-12)).all()
    np.testing.assert_allclose(np.linalg.norm(poses[:, :4], axis=1), 1.0, atol=1e-14)
    np.testing.assert_allclose(np.linalg.norm(poses[:, 4:], axis=1), 1.0, atol=1e-14)
    # 2. the returned cost is the cost of the returned pose (fused eval == solve bookkeeping)
    ev = handle.eval_batch(f1, f2, ct, None, res.poses, api.TARGET, 1e-13, n_per_problem=N)
    np.testing.assert_allclose(ev.cost.cpu().numpy(), cost, rtol=1e-9)
    # 3. idempotence: restarting from the solution (almost always) returns it unchanged,
    #    and never moves further than the early-stopping slack of function_tolerance
    again = handle.solve_batch(f1, f2, ct, None, res.poses, o, n_per_problem=N)
    p2 = again.poses.cpu().numpy()
    rot = np.array([rotation_angle(a, c) for a, c in zip(p2, poses)])
    tra = np.array([direction_angle(a[4:], c[4:]) for a, c in zip(p2, poses)])
    # (frame pairs with almost no translation have a flat valley in t: bound those by percentile)
    assert np.percentile(rot, 90) < 1e-9 and rot.max() < 1e-3
    assert np.percentile(tra, 90) < 1e-7 and np.percentile(tra, 99) < 1e-2
    assert (again.cost.cpu().numpy() <= cost * (1 + 1e-12)).all()
    # 4. gauge: the energy is even in t, the t-columns of the gradient are odd
    import torch

    flipped = res.poses.clone()
    flipped[:, 4:] *= -1
    evf = handle.eval_batch(f1, f2, ct, None, flipped, api.TARGET, 1e-13, n_per_problem=N)
    np.testing.assert_allclose(evf.cost.cpu().numpy(), ev.cost.cpu().numpy(), rtol=1e-12)
    np.testing.assert_allclose(evf.jtj.cpu().numpy()[:, 9:], ev.jtj.cpu().numpy()[:, 9:], rtol=1e-9, atol=1e-6)
    # 5. additivity: JtJ of a problem == sum of JtJ over its two halves (same pose)
    half = handle.eval_batch(f1, f2, ct, None, res.poses.repeat_interleave(2, dim=0), api.TARGET, 1e-13,
                             n_per_problem=N // 2)
    np.testing.assert_allclose(half.jtj.cpu().numpy().reshape(B, 2, 15).sum(1), ev.jtj.cpu().numpy(),
                               rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(half.cost.cpu().numpy().reshape(B, 2).sum(1), ev.cost.cpu().numpy(), rtol=1e-12)
    # 6. accuracy: the refinement lands near the ground truth (statistical sanity)
    gt = b.gt_poses
    dq = np.abs(np.sum(poses[:, :4] * gt[:, :4], axis=1))
    assert np.median(2 * np.arccos(np.clip(dq, -1, 1))) < 5e-4


def test_c2_full_batch_matches_oracle(handle, c2_batch):
    """All 10 000 frame pairs of BASELINE config C2 against the oracle (about a second of CPU
    time on the box's cores): poses within the bar, identical iteration counts and statuses."""
    b = c2_batch
    N = 512
    res = handle.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), None,
                             dev(b.init_poses), api.default_opts(api.TARGET), n_per_problem=N)
    ref, info = oracle_solve(api.TARGET, b.bvs_host, b.bvs_target, b.covs_target, None,
                             b.init_poses, n_per_problem=N)
    r, t = max_pose_diff(res.poses.cpu().numpy(), ref)
    assert r <= ROT_TOL and t <= DIR_TOL, (r, t)
    assert np.array_equal(res.iterations.cpu().numpy(), info["iterations"])
    assert np.array_equal(res.status.cpu().numpy(), info["status"])
    np.testing.assert_allclose(res.cost.cpu().numpy(), info["final_cost"], rtol=1e-8)


# ------------------------------------------- reference-shaped host interfaces


def test_pypnec_module_matches_oracle():
    """pypnec.pyceres / pyceresnec with the reference's signatures (python/pypnec.cpp:50-82)."""
    from pnec_b200 import pypnec

    N = 90
    b = syn.with_host_covariances(syn.make_batch(1, N, seed=31))
    f1 = [v for v in b.bvs_host]
    f2 = [v for v in b.bvs_target]
    cov_h = [c.reshape(3, 3).T.copy() for c in b.covs_host]
    cov_t = [c.reshape(3, 3).T.copy() for c in b.covs_target]
    init = np.eye(4)
    init[:3, :3] = syn.quaternion_to_matrix(b.init_poses[0][:4])
    init[:3, 3] = b.init_poses[0][4:]
    out = pypnec.pyceres(f1, f2, cov_h, cov_t, init, 1e-13)
    assert out.shape == (4, 4) and np.allclose(out[3], [0, 0, 0, 1])
    ref, _ = oracle.solve(b.bvs_host, b.bvs_target, b.covs_target, b.covs_host, b.init_poses[0],
                          oracle.default_opts(oracle.SYMMETRIC))
    got = np.concatenate([syn.matrix_to_quaternion(out[:3, :3]), out[:3, 3]])
    assert rotation_angle(got, ref) <= ROT_TOL and direction_angle(got[4:], ref[4:]) <= DIR_TOL
    assert np.linalg.norm(out[:3, 3]) == pytest.approx(1.0, abs=1e-14)
    out_nec = pypnec.pyceresnec(np.asarray(f1), np.asarray(f2), init)
    refn, _ = oracle.solve(b.bvs_host, b.bvs_target, None, None, b.init_poses[0], oracle.default_opts(oracle.NEC))
    gotn = np.concatenate([syn.matrix_to_quaternion(out_nec[:3, :3]), out_nec[:3, 3]])
    assert rotation_angle(gotn, refn) <= ROT_TOL and direction_angle(gotn[4:], refn[4:]) <= DIR_TOL
    # batched extension
    outs = pypnec.pyceres_target_batch(b.bvs_host.reshape(1, N, 3), b.bvs_target.reshape(1, N, 3),
                                       np.asarray(cov_t).reshape(1, N, 3, 3), init[None], 1e-13)
    reft, _ = oracle.solve(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses[0],
                           oracle.default_opts(oracle.TARGET))
    gott = np.concatenate([syn.matrix_to_quaternion(outs[0, :3, :3]), outs[0, :3, 3]])
    assert rotation_angle(gott, reft) <= ROT_TOL and direction_angle(gott[4:], reft[4:]) <= DIR_TOL
    with pytest.raises(ValueError):
        pypnec.pyceresnec(f1, f2[:-1], init)
    # addition: the whole frame solve (PNEC::Solve), without and with (the default) RANSAC
    full = pypnec.pysolve(f1, f2, cov_t, init, 1e-13, use_ransac=False)
    fref, _ = oracle.frame_solve_batch(b.bvs_host, b.bvs_target, b.covs_target, b.init_poses,
                                       oracle.default_frame_opts(use_ransac=0), n_per_problem=N)
    gotf = np.concatenate([syn.matrix_to_quaternion(full[:3, :3]), full[:3, 3]])
    assert rotation_angle(gotf, fref[0]) <= ROT_TOL and direction_angle(gotf[4:], fref[0][4:]) <= DIR_TOL
    full, inl = pypnec.pysolve_inliers(f1, f2, cov_t, init)
    fref, _, rmask, rni, _ = oracle.frame_solve_batch(b.bvs_host, b.bvs_target, b.covs_target, b.init_poses,
                                                      oracle.default_frame_opts(), n_per_problem=N, return_ransac=True)
    gotf = np.concatenate([syn.matrix_to_quaternion(full[:3, :3]), full[:3, 3]])
    assert np.array_equal(inl, np.nonzero(rmask)[0]) and len(inl) == rni[0]
    assert rotation_angle(gotf, fref[0]) <= ROT_TOL and direction_angle(gotf[4:], fref[0][4:]) <= DIR_TOL


def test_cpp_compat_api_matches_oracle(tmp_path):
    """The reference's C++ call shapes (PNEC::CeresSolver & co) through pnec_compat.hpp."""
    exe = tmp_path / "compat_test"
    lib = os.path.join(ROOT, "pnec_b200", "lib")
    subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O1", "-pthread",
                    "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "compat_test.cpp"),
                    "-o", str(exe), "-L", lib, "-lpnec_b200", f"-Wl,-rpath,{lib}"], check=True)
    N = 120
    b = syn.with_host_covariances(syn.make_batch(1, N, seed=41))
    blob = tmp_path / "in.bin"
    with open(blob, "wb") as f:
        f.write(np.int64(N).tobytes())
        for a in (b.bvs_host, b.bvs_target, b.covs_target, b.covs_host, b.init_poses[0]):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    out = subprocess.run([str(exe), str(blob)], check=True, capture_output=True, text=True).stdout
    lines = {l.split()[0]: l.split()[1:] for l in out.strip().splitlines()}
    init = b.init_poses[0]

    def check(tag, variant, reg, ct, ch):
        pose = np.array(lines[tag][:7], dtype=np.float64)
        ref, info = oracle.solve(b.bvs_host, b.bvs_target, ct, ch, init, oracle.default_opts(variant, reg))
        assert rotation_angle(pose, ref) <= ROT_TOL and direction_angle(pose[4:], ref[4:]) <= DIR_TOL, tag
        assert int(lines[tag][7]) == info.status and int(lines[tag][8]) == info.iterations, tag
        return pose

    p = check("CeresSolver", oracle.TARGET, 1e-13, b.covs_target, None)
    check("CeresSolverFull", oracle.TARGET, 1e-10, b.covs_target, None)
    check("NECCeresSolver", oracle.NEC, 0.0, None, None)
    check("Solve", oracle.TARGET, 1e-13, b.covs_target, None)
    check("PNECCeresHost", oracle.HOST, 1e-13, b.covs_target, None)
    check("PNECCeresSymmetric", oracle.SYMMETRIC, 1e-13, b.covs_target, b.covs_host)
    assert float(lines["CostFunction"][0]) == pytest.approx(
        oracle.cost_function(b.bvs_host, b.bvs_target, b.covs_target, p), rel=1e-10)
    # PNEC(Options()) -- RANSAC on, the reference's default (pnec_config.h:58; run_simulation.cc:74-86)
    dref, des, rmask, rni, _ = oracle.frame_solve_batch(b.bvs_host, b.bvs_target, b.covs_target, b.init_poses,
                                                        oracle.default_frame_opts(), n_per_problem=N, return_ransac=True)
    for tag, ref in (("SolveDefault", dref[0]), ("SolveDefaultES", des[0]), ("EigensolverRansac", des[0])):
        pose = np.array(lines[tag][:7], dtype=np.float64)
        assert rotation_angle(pose, ref) <= ROT_TOL and direction_angle(pose[4:], ref[4:]) <= DIR_TOL, tag
    assert int(lines["SolveDefault"][8]) == rni[0] == int(lines["EigensolverRansac"][8])
    assert [int(v) for v in lines["SolveDefaultInliers"]] == list(np.nonzero(rmask)[0])
    same_pose, same_inliers, ms0, ms1, ms2, total_ms = lines["SolveTimed"]
    assert same_pose == "1" and same_inliers == "1"
    assert float(ms0) > 0 and float(ms1) > 0 and float(ms2) > 0 and int(total_ms) <= float(ms0) + float(ms1) + float(ms2)
    assert lines["Threads"] == ["4"]
    # the whole PNEC::Solve pipeline and its stages (pnec.cc:77-124, 273-348)
    fref, fes = oracle.frame_solve_batch(b.bvs_host, b.bvs_target, b.covs_target, b.init_poses,
                                         oracle.default_frame_opts(use_ransac=0), n_per_problem=N)
    nref, _ = oracle.frame_solve_batch(b.bvs_host, b.bvs_target, b.covs_target, b.init_poses,
                                       oracle.default_frame_opts(use_ransac=0, use_nec=1), n_per_problem=N)
    wref = oracle.weighted_eigensolver(b.bvs_host, b.bvs_target, b.covs_target, fes[0])
    for tag, ref in (("SolveFull", fref[0]), ("SolveFullES", fes[0]), ("Eigensolver", fes[0]),
                     ("WeightedEigensolver", wref), ("SolveNEC", nref[0])):
        pose = np.array(lines[tag][:7], dtype=np.float64)
        assert rotation_angle(pose, ref) <= ROT_TOL and direction_angle(pose[4:], ref[4:]) <= DIR_TOL, tag
    assert lines["SolveFull"][8] == "0"  # inliers cleared, pnec.cc:277
    # the same call site written against Sophus::SE3d / Eigen containers (mock headers, tests/cpp/mock)
    exe2 = tmp_path / "compat_interop"
    subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O1",
                    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp", "mock"),
                    os.path.join(ROOT, "tests", "cpp", "compat_interop_test.cpp"),
                    "-o", str(exe2), "-L", lib, "-lpnec_b200", f"-Wl,-rpath,{lib}"], check=True)
    out2 = subprocess.run([str(exe2), str(blob)], check=True, capture_output=True, text=True).stdout
    lines2 = {l.split()[0]: l.split()[1:] for l in out2.strip().splitlines()}
    pose = np.array(lines2["SolveSophus"][:7], dtype=np.float64)
    assert rotation_angle(pose, dref[0]) <= ROT_TOL and direction_angle(pose[4:], dref[0][4:]) <= DIR_TOL
    assert int(lines2["SolveSophus"][8]) == rni[0]
    assert lines2["CeresSolverSophus"][:7] == lines["CeresSolver"][:7]
    mu = b.bvs_target[1] * 800.0
    img = np.array([[0.7, 0.1, 0.0], [0.1, 0.4, 0.0], [0.0, 0.0, 0.0]])
    ut = oracle.unscented_transform(mu[None], img.T.reshape(1, 9), None, 1.0, oracle.PINHOLE)[0].reshape(3, 3).T
    got = [float(v) for v in lines["UnscentedTransform"][:3]]
    np.testing.assert_allclose(got, [ut[0, 0], ut[0, 1], ut[2, 2]], rtol=1e-9, atol=1e-20)
    assert lines["UnscentedTransform"][3] == "1"
