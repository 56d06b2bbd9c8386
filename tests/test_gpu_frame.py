"""GPU (-m gpu): the stages of PNEC::Solve in front of the refinement — NEC eigensolver rotation,
weighted eigensolver + SCF, and the whole frame solve — through the C-ABI against the CPU oracle
(oracle/pnec_oracle_frame.c) on identical inputs.

Tolerances (BASELINE.json north_star): rotation within 1e-6 rad, translation direction within
1e-6 rad modulo sign.
"""
import numpy as np
import pytest

import oracle
from conftest import direction_angle, max_pose_diff, rotation_angle
from pnec_b200 import api
from pnec_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-6  # rad
DIR_TOL = 1e-6  # rad


@pytest.fixture(scope="module")
def handle():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a B200"
    return api.Handle(0)


def dev(a):
    import torch

    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def perturbed(batch, seed=0):
    """The same batch with every bearing-vector coordinate moved by at most one ulp."""
    rng = np.random.default_rng(seed)
    import copy

    p = copy.copy(batch)
    p.bvs_host = batch.bvs_host * (1.0 + rng.uniform(-1, 1, batch.bvs_host.shape) * 2.0 ** -52)
    p.bvs_target = batch.bvs_target * (1.0 + rng.uniform(-1, 1, batch.bvs_target.shape) * 2.0 ** -52)
    return p


def well_posed(ref_a, ref_b, rot_tol=1e-8, dir_tol=1e-8, min_fraction=0.85):
    """Frame pairs on which the REFERENCE ALGORITHM ITSELF (the oracle) reproduces its result when the
    inputs move by one ulp.  The closed-form eigenvalue derivative of the eigensolver is singular when
    the two smallest eigenvalues of M coincide (near-pure rotation), and the NEC translation is
    undefined there: such pairs amplify rounding noise above the parity tolerance in any
    implementation, the reference's included, and cannot serve as parity cases."""
    ok = np.array([rotation_angle(a, b) <= rot_tol and direction_angle(a[4:], b[4:]) <= dir_tol
                   for a, b in zip(ref_a, ref_b)])
    assert ok.mean() >= min_fraction, f"only {ok.mean():.2f} of the frame pairs are well posed"
    return ok


def oracle_es(batch, weights_from=None, reg=1e-13):
    out = np.zeros((batch.num_problems, 7))
    infos = []
    for b in range(batch.num_problems):
        f1, f2, ct, _ = batch.problem(b)
        w = None if weights_from is None else oracle.weights(f1, ct, weights_from[b], reg)
        q, info = oracle.eigensolver(f1, f2, batch.init_poses[b], w)
        out[b, :4] = q
        out[b, 4:] = batch.init_poses[b, 4:]
        infos.append(info.lm_info)
    return out, np.array(infos)


@pytest.mark.parametrize("n,noise,camera", [(100, 1.0, syn.OMNIDIRECTIONAL), (512, 1.0, syn.OMNIDIRECTIONAL),
                                            (64, 0.5, syn.PINHOLE), (10, 1.0, syn.OMNIDIRECTIONAL)])
def test_eigensolver_rotation_matches_oracle(handle, n, noise, camera):
    batch = syn.make_batch(96, n, seed=21 + n, noise_level=noise, camera=camera)
    ref, ref_info = oracle_es(batch)
    ok = well_posed(ref, oracle_es(perturbed(batch))[0])
    poses, info, ev = handle.eigensolver_batch(batch.bvs_host, batch.bvs_target, batch.init_poses, n_per_problem=n)
    r = max(rotation_angle(a, b) for a, b in zip(poses[ok], ref[ok]))
    assert r <= ROT_TOL, r
    # the others still land within the reference's own noise ball
    assert max(rotation_angle(a, b) for a, b in zip(poses, ref)) <= 1e-4
    # translation is passed through
    np.testing.assert_array_equal(poses[:, 4:], batch.init_poses[:, 4:])
    assert np.all((info >= 1) & (info <= 8))
    assert np.all(np.isfinite(ev))


def test_eigensolver_device_pointers_ragged(handle):
    counts = np.array([5, 17, 64, 333, 1, 0, 700, 31, 32, 33], dtype=np.int64)
    batch = syn.make_batch(len(counts), 0, seed=5, counts=counts)
    ref, _ = oracle_es(batch)
    refp, _ = oracle_es(perturbed(batch))
    ok = (counts >= 6) & np.array([rotation_angle(a, b) <= 1e-8 for a, b in zip(ref, refp)])
    assert ok.sum() >= 6
    poses, info, ev = handle.eigensolver_batch(dev(batch.bvs_host), dev(batch.bvs_target), dev(batch.init_poses),
                                               offsets=batch.offsets)
    poses = poses.cpu().numpy()
    for b in np.nonzero(ok)[0]:
        assert rotation_angle(poses[b], ref[b]) <= ROT_TOL, b


def test_weighted_eigensolver_rotation_matches_oracle(handle):
    n = 256
    batch = syn.make_batch(64, n, seed=33)
    # weights from the ground truth poses, start at the perturbed ones
    ref, _ = oracle_es(batch, weights_from=batch.gt_poses)
    ok = well_posed(ref, oracle_es(perturbed(batch), weights_from=batch.gt_poses)[0])
    poses, info, ev = handle.eigensolver_batch(batch.bvs_host, batch.bvs_target, batch.init_poses,
                                               covs_target=batch.covs_target, weight_poses=batch.gt_poses,
                                               n_per_problem=n)
    r = max(rotation_angle(a, b) for a, b in zip(poses[ok], ref[ok]))
    assert r <= ROT_TOL, r


FRAME_CONFIGS = {
    "default": dict(),  # eigensolver -> 9 weighted iterations -> Ceres (TARGET)
    "nec_ceres": dict(use_nec=1),
    "nec_only": dict(use_nec=1, use_ceres=0),
    "es_then_ceres": dict(weighted_iterations=1),
    "ceres_only": dict(weighted_iterations=0),
    "weighted_no_ceres": dict(use_ceres=0, weighted_iterations=3),
}


@pytest.mark.parametrize("cfg", list(FRAME_CONFIGS))
def test_frame_solve_matches_oracle(handle, cfg):
    n, B = 200, 48
    batch = syn.make_batch(B, n, seed=77)
    kw = FRAME_CONFIGS[cfg]
    def run_oracle(bt):
        return oracle.frame_solve_batch(bt.bvs_host, bt.bvs_target, bt.covs_target, bt.init_poses,
                                        oracle.default_frame_opts(use_ransac=0, **kw), n_per_problem=n,
                                        num_threads=oracle.max_threads())

    ref, ref_es = run_oracle(batch)
    ref_p, ref_es_p = run_oracle(perturbed(batch))
    ok = well_posed(ref, ref_p) & well_posed(ref_es, ref_es_p)
    res = handle.frame_solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses,
                                   api.default_frame_opts(use_ransac=0, **kw), n_per_problem=n)
    r, t = max_pose_diff(res.es_poses[ok], ref_es[ok])
    assert r <= ROT_TOL and t <= DIR_TOL, ("eigensolver stage", r, t)
    r, t = max_pose_diff(res.poses[ok], ref[ok])
    assert r <= ROT_TOL and t <= DIR_TOL, (cfg, r, t)


def test_frame_solve_device_matches_host_call(handle):
    n, B = 128, 40
    batch = syn.make_batch(B, n, seed=78)
    opts = api.default_frame_opts(use_ransac=0, weighted_iterations=4)
    host = handle.frame_solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses, opts,
                                    n_per_problem=n)
    devr = handle.frame_solve_batch(dev(batch.bvs_host), dev(batch.bvs_target), dev(batch.covs_target),
                                    dev(batch.init_poses), opts, n_per_problem=n)
    np.testing.assert_array_equal(host.poses, devr.poses.cpu().numpy())
    np.testing.assert_array_equal(host.es_poses, devr.es_poses.cpu().numpy())
    np.testing.assert_array_equal(host.status, devr.status.cpu().numpy())


def test_frame_solve_improves_on_eigensolver(handle):
    """Size-independent property at the C2 shape: the refined pose is at least as close to the
    ground truth (on average) as the NEC eigensolver's, and everything converged."""
    n, B = 512, 256
    batch = syn.make_batch(B, n, seed=79)
    res = handle.frame_solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses,
                                   api.default_frame_opts(use_ransac=0), n_per_problem=n)
    e_es = np.mean([rotation_angle(a, b) for a, b in zip(res.es_poses, batch.gt_poses)])
    e_fin = np.mean([rotation_angle(a, b) for a, b in zip(res.poses, batch.gt_poses)])
    assert e_fin <= e_es * 1.02, (e_es, e_fin)
    assert np.all(res.status <= 3)


def test_frame_solve_rejects_bad_ransac_settings(handle):
    batch = syn.make_batch(2, 16, seed=1)
    for kw in (dict(ransac_sample_size=0), dict(ransac_sample_size=33), dict(ransac_probability=1.0),
               dict(ransac_threshold=0.0), dict(max_ransac_iterations=-1)):
        with pytest.raises(api.PnecError):
            handle.frame_solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses,
                                     api.default_frame_opts(use_ransac=1, **kw), n_per_problem=16)


@pytest.mark.parametrize("name", ["omni_n200", "pinhole_n96"])
def test_frame_stages_match_committed_fixtures(handle, golden_frame, name):
    """CUDA path vs tests/golden/oracle_frame.npz; frame pairs on which the algorithm itself is not
    stable under a one-ulp input perturbation (`*_ulp` entries) are not parity cases."""
    g = golden_frame
    n = int(g[f"{name}/n"])
    f1, f2, cov, init, gt = (g[f"{name}/{k}"] for k in ("f1", "f2", "cov", "init", "gt"))

    def stable(a, b):
        return np.array([rotation_angle(x, y) <= 1e-8 and direction_angle(x[4:], y[4:]) <= 1e-8
                         for x, y in zip(a, b)])

    as_pose = lambda q: np.concatenate([q, np.tile([0.0, 0.0, 1.0], (q.shape[0], 1))], axis=1)
    poses, _, _ = handle.eigensolver_batch(f1, f2, init, n_per_problem=n)
    ok = stable(as_pose(g[f"{name}/es_quat"]), as_pose(g[f"{name}/es_quat_ulp"]))
    assert ok.mean() > 0.8
    assert max(rotation_angle(a, b) for a, b in zip(poses[ok], as_pose(g[f"{name}/es_quat"])[ok])) <= ROT_TOL
    poses, _, _ = handle.eigensolver_batch(f1, f2, init, covs_target=cov, weight_poses=gt, n_per_problem=n)
    ok = stable(as_pose(g[f"{name}/weighted_es_quat"]), as_pose(g[f"{name}/weighted_es_quat_ulp"]))
    assert max(rotation_angle(a, b) for a, b in zip(poses[ok], as_pose(g[f"{name}/weighted_es_quat"])[ok])) <= ROT_TOL
    for cname, kw in {"default": {}, "nec_ceres": {"use_nec": 1}, "es_then_ceres": {"weighted_iterations": 1},
                      "weighted3_no_ceres": {"use_ceres": 0, "weighted_iterations": 3}}.items():
        res = handle.frame_solve_batch(f1, f2, cov, init, api.default_frame_opts(use_ransac=0, **kw), n_per_problem=n)
        ok = stable(g[f"{name}/{cname}/poses"], g[f"{name}/{cname}/poses_ulp"]) & \
            stable(g[f"{name}/es_pose"], g[f"{name}/es_pose_ulp"])
        assert ok.mean() > 0.8, cname
        r, t = max_pose_diff(res.poses[ok], g[f"{name}/{cname}/poses"][ok])
        assert r <= ROT_TOL and t <= DIR_TOL, (cname, r, t)
        r, t = max_pose_diff(res.es_poses[ok], g[f"{name}/es_pose"][ok])
        assert r <= ROT_TOL and t <= DIR_TOL, (cname, "es", r, t)


def test_frame_solve_shortcuts_are_exact(handle, monkeypatch):
    """The scan cache / fixed-point skipping / two-pass launch must not change a single bit (small
    batch: the fused rounds kernel; the per-round kernels at bench sizes are covered by
    test_gpu_parity_large.py).  The switches are read when a handle is created."""
    n, B = 160, 64
    batch = syn.make_batch(B, n, seed=80)
    args = (batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses)
    fast = handle.frame_solve_batch(*args, api.default_frame_opts(use_ransac=0), n_per_problem=n)
    monkeypatch.setenv("PNEC_B200_NO_FRAME_SHORTCUTS", "1")
    monkeypatch.setenv("PNEC_B200_SCF_DEFER", "0")
    plain = api.Handle(0).frame_solve_batch(*args, api.default_frame_opts(use_ransac=0), n_per_problem=n)
    np.testing.assert_array_equal(fast.poses, plain.poses)
    np.testing.assert_array_equal(fast.iterations, plain.iterations)


def test_large_pairs_spill_to_device_memory(handle):
    """Pairs beyond the shared-memory capacity of the SCF stage (~3000 correspondences)."""
    counts = np.array([3600, 200, 4100], dtype=np.int64)
    batch = syn.make_batch(len(counts), 0, seed=91, counts=counts)
    for b in range(len(counts)):
        f1, f2, ct, _ = batch.problem(b)
        ref_t, ref_c = oracle.scf_translation(f1, f2, ct, batch.init_poses[b])
        t, c = handle.scf_translation_batch(f1, f2, ct, batch.init_poses[b:b + 1], n_per_problem=len(f1))
        assert direction_angle(t[0], ref_t) <= DIR_TOL
        assert c[0] == pytest.approx(ref_c, rel=1e-9)
    ref, ref_es = oracle.frame_solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses,
                                           oracle.default_frame_opts(use_ransac=0, weighted_iterations=3), offsets=batch.offsets,
                                           num_threads=oracle.max_threads())
    res = handle.frame_solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses,
                                   api.default_frame_opts(use_ransac=0, weighted_iterations=3), offsets=batch.offsets)
    r, t = max_pose_diff(res.poses, ref)
    assert r <= ROT_TOL and t <= DIR_TOL, (r, t)


def test_frame_solve_ragged_device_batch(handle):
    """KITTI-shaped ragged batch (C4 of BASELINE.json in miniature) with device pointers."""
    counts = np.array([310, 97, 512, 1033, 64, 200, 777, 45], dtype=np.int64)
    batch = syn.make_batch(len(counts), 0, seed=93, camera=syn.PINHOLE, counts=counts)
    fo = dict(weighted_iterations=4, max_num_iterations=20)
    ref, ref_es = oracle.frame_solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses,
                                           oracle.default_frame_opts(use_ransac=0, **fo), offsets=batch.offsets,
                                           num_threads=oracle.max_threads())
    p = perturbed(batch)
    ref_p, _ = oracle.frame_solve_batch(p.bvs_host, p.bvs_target, p.covs_target, p.init_poses,
                                        oracle.default_frame_opts(use_ransac=0, **fo), offsets=batch.offsets,
                                        num_threads=oracle.max_threads())
    ok = well_posed(ref, ref_p, min_fraction=0.7)
    res = handle.frame_solve_batch(dev(batch.bvs_host), dev(batch.bvs_target), dev(batch.covs_target),
                                   dev(batch.init_poses), api.default_frame_opts(use_ransac=0, **fo), offsets=batch.offsets)
    r, t = max_pose_diff(res.poses.cpu().numpy()[ok], ref[ok])
    assert r <= ROT_TOL and t <= DIR_TOL, (r, t)


def test_frame_solve_degenerate_inputs_terminate(handle):
    """Empty pairs, pairs below the minimal sample size and a pure rotation (the eigensolver's
    singular case) must terminate with a pose per pair; empty pairs keep their start pose."""
    counts = np.array([0, 1, 3, 6, 40, 0, 25], dtype=np.int64)
    batch = syn.make_batch(len(counts), 0, seed=12, counts=counts)
    # pair 4: pure rotation (no translation): f2 = R^T f1 exactly
    s, e = batch.range(4)
    R = syn.quaternion_to_matrix(batch.gt_poses[4, :4])
    batch.bvs_target[s:e] = batch.bvs_host[s:e] @ R
    res = handle.frame_solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses,
                                   api.default_frame_opts(use_ransac=0), offsets=batch.offsets)
    assert res.poses.shape == (len(counts), 7)
    assert (res.status[counts == 0] == 7).all()  # PNEC_STATUS_EMPTY
    for b in np.nonzero(counts == 0)[0]:
        q0 = batch.init_poses[b, :4] / np.linalg.norm(batch.init_poses[b, :4])
        np.testing.assert_allclose(res.poses[b, :4], q0, atol=1e-15)
    ok = counts >= 25
    assert np.isfinite(res.poses[ok]).all()
    # the pure rotation is recovered even though its translation is undefined
    assert rotation_angle(res.poses[4], batch.gt_poses[4]) < 1e-6
    ref, _ = oracle.frame_solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses,
                                      oracle.default_frame_opts(use_ransac=0), offsets=batch.offsets)
    assert rotation_angle(res.poses[6], ref[6]) <= ROT_TOL


def test_fused_rounds_kernel_matches_per_round_kernels(handle, monkeypatch):
    """Small batches run all weighted rounds of a pair in one launch (pnec_frame.cuh): same device
    functions and order of operations as the per-round kernels, compiled into a different kernel, so
    the results may differ in the last bit (the compiler's choice of which product of a sum of
    products to fuse) but no more."""
    counts = np.array([200, 64, 333, 0, 12, 512, 97, 150], dtype=np.int64)
    batch = syn.make_batch(len(counts), 0, seed=95, counts=counts)
    args = (batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses)
    for kw in (dict(), dict(weighted_iterations=2), dict(use_ceres=0)):
        monkeypatch.setenv("PNEC_B200_FUSED_ROUNDS_MAX_PAIRS", "512")
        fused = api.Handle(0).frame_solve_batch(*args, api.default_frame_opts(use_ransac=0, **kw), offsets=batch.offsets)
        monkeypatch.setenv("PNEC_B200_FUSED_ROUNDS_MAX_PAIRS", "0")
        rounds = api.Handle(0).frame_solve_batch(*args, api.default_frame_opts(use_ransac=0, **kw), offsets=batch.offsets)
        np.testing.assert_allclose(fused.poses, rounds.poses, rtol=0, atol=1e-13)  # (rotation LM: two instantiations)
        np.testing.assert_array_equal(fused.es_poses, rounds.es_poses)


def test_rotation_lm_wide_turns_are_bit_identical(handle):
    """es_lm_group evaluates a trial point and its three forward-difference columns in ONE turn as soon
    as at most two of a warp's eight pairs are unfinished (wide turns, pnec_eigensolver.cuh).  The
    arithmetic per pair is lmdif's either way: a pair solved alone in its call (wide from the first
    turn), next to one other pair, and inside a batch of 96 (narrow turns until the warp's tail) must
    give the same bits -- rotation, MINPACK info code and smallest eigenvalue."""
    n = 160
    batch = syn.make_batch(96, n, seed=77, noise_level=1.0)
    full, info, ev = handle.eigensolver_batch(batch.bvs_host, batch.bvs_target, batch.init_poses, n_per_problem=n)
    for group in ([0], [5], [37, 38], [90, 3], list(range(8, 11))):
        rows = np.concatenate([np.arange(b * n, (b + 1) * n) for b in group])
        p, i, e = handle.eigensolver_batch(batch.bvs_host[rows], batch.bvs_target[rows], batch.init_poses[group],
                                           n_per_problem=n)
        np.testing.assert_array_equal(p, full[group])
        np.testing.assert_array_equal(i, info[group])
        np.testing.assert_array_equal(e, ev[group])


def test_nec_translation_warp_kernel_is_bit_identical(handle):
    """Large batches run TranslationFromM(ComposeM(...)) with a warp per pair, small ones with a CTA per
    pair (pnec_translation.cuh); the sums are combined in the same order, so the same pair gives the
    same bits in a batch of 1 500 (warp kernel: >= 8 pairs per SM) and in a batch of 40."""
    counts = np.tile(np.array([130, 1, 64, 257, 2, 33, 96, 512, 129, 5], dtype=np.int64), 150)
    batch = syn.make_batch(len(counts), 0, seed=41, counts=counts)
    big, big_m = handle.nec_translation_batch(batch.bvs_host, batch.bvs_target, batch.init_poses, offsets=batch.offsets)
    k = 40
    n = int(batch.offsets[k])
    small, small_m = handle.nec_translation_batch(batch.bvs_host[:n], batch.bvs_target[:n], batch.init_poses[:k],
                                                  offsets=batch.offsets[: k + 1])
    np.testing.assert_array_equal(big[:k], small)
    np.testing.assert_array_equal(big_m[:k], small_m)
    for b in (0, 3, 7, 1203):
        sl = slice(int(batch.offsets[b]), int(batch.offsets[b + 1]))
        rt, rm = oracle.nec_translation(batch.bvs_host[sl], batch.bvs_target[sl], batch.init_poses[b])
        np.testing.assert_allclose(big_m[b], rm, rtol=1e-12, atol=1e-15)
