"""GPU (-m gpu): parity cases the round-1 suite did not reach.

* the code path the frame bench times: B >= 4096 runs 4 chunks on separate streams, per-round
  scf_kernel + scf_list_kernel, and the rotation chain on a priority stream (pnec_capi.cu,
  pnec_frame_solve_batch) -- the whole of BASELINE config C2 against the oracle, and the
  shortcut devices (scan cache, fixed points, two passes, chunking, rotation chain ahead)
  compared bit for bit against the plain sequence at B >= 4096 and at 513 <= B < 1024;
* the start pose of the reference's VO loop (src/odometry/frame_processing.cc:97-102: previous
  rotation, ZERO translation => AnglesFromVec's zero-vector branch, src/common/common.cc:104-106);
* a start at the poles of the S^2 chart (t = +-e_z: d t / d phi = 0, the LM runs with a zero
  column that only min_lm_diagonal regularises), all four residual variants, solve and eval.

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
VARIANTS = {"nec": api.NEC, "target": api.TARGET, "host": api.HOST, "symmetric": api.SYMMETRIC}


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
    import copy

    rng = np.random.default_rng(seed)
    p = copy.copy(batch)
    p.bvs_host = batch.bvs_host * (1.0 + rng.uniform(-1, 1, batch.bvs_host.shape) * 2.0 ** -52)
    p.bvs_target = batch.bvs_target * (1.0 + rng.uniform(-1, 1, batch.bvs_target.shape) * 2.0 ** -52)
    return p


def pose_diffs(a, b):
    r = np.array([rotation_angle(x, y) for x, y in zip(a, b)])
    t = np.array([direction_angle(x[4:], y[4:]) for x, y in zip(a, b)])
    return r, t


def check_against_oracle(got, ref, ref_ulp, min_fraction, what, k=100.0, k_dir=1000.0):
    """Pairs on which the oracle reproduces itself to 1e-8 under a one-ulp input perturbation must
    meet the parity bar, and at least `min_fraction` of the batch must be such pairs.  The others
    amplify rounding noise in ANY implementation, the reference's included (near-pure rotation: the
    two smallest eigenvalues of M meet, the eigenvalue derivative is singular and the translation
    direction is not determined by the data); they are bounded by k times the oracle's own
    self-difference -- one sample of its noise ball; k_dir for the translation direction, which a
    near-degenerate M leaves almost free -- unless that bound is vacuous (> 0.1 rad)."""
    sr, st = pose_diffs(ref, ref_ulp)
    ok = (sr <= 1e-8) & (st <= 1e-8)
    assert ok.mean() >= min_fraction, f"{what}: only {ok.mean():.4f} of the pairs are well posed"
    r, t = pose_diffs(got, ref)
    assert r[ok].max() <= ROT_TOL and t[ok].max() <= DIR_TOL, (what, r[ok].max(), t[ok].max())
    bad = ~ok
    if bad.any():
        rb, tb = np.maximum(k * sr[bad], ROT_TOL), np.maximum(k_dir * st[bad], DIR_TOL)
        assert ((r[bad] <= rb) | (rb > 0.1)).all(), (what, "ill-posed rot", r[bad], sr[bad])
        assert ((t[bad] <= tb) | (tb > 0.1)).all(), (what, "ill-posed dir", t[bad], st[bad])
    return ok


# --------------------------------------------------------------- the benched path


@pytest.fixture(scope="module")
def c2_batch():
    """BASELINE config C2: 10 000 frame pairs x 512 correspondences, anisotropic."""
    return syn.make_batch(10000, 512, seed=2024, noise_type="anisotropic_inhomogenous")


def test_c2_full_batch_frame_solve_matches_oracle(handle, c2_batch):
    """PNEC::Solve (no RANSAC) for all of C2 through the 4-chunk / priority-stream / two-pass path
    that bench.py's frame_pipeline times, against the oracle (a few seconds of CPU on the box)."""
    b, N = c2_batch, 512
    fo = oracle.default_frame_opts(use_ransac=0)
    run = lambda bt: oracle.frame_solve_batch(bt.bvs_host, bt.bvs_target, bt.covs_target, bt.init_poses, fo,
                                              n_per_problem=N, num_threads=oracle.max_threads())
    ref, ref_es = run(b)
    ref_p, ref_es_p = run(perturbed(b))
    res = handle.frame_solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), dev(b.init_poses),
                                   api.default_frame_opts(use_ransac=0), n_per_problem=N)
    check_against_oracle(res.es_poses.cpu().numpy(), ref_es, ref_es_p, 0.99, "eigensolver stage")
    check_against_oracle(res.poses.cpu().numpy(), ref, ref_p, 0.99, "frame solve")
    assert (res.status.cpu().numpy() <= 4).all()
    # the HOST call (chunked H2D) returns the same bits as the device call
    host = handle.frame_solve_batch(b.bvs_host, b.bvs_target, b.covs_target, b.init_poses,
                                    api.default_frame_opts(use_ransac=0), n_per_problem=N)
    np.testing.assert_array_equal(host.poses, res.poses.cpu().numpy())


PLAIN_ENV = {"PNEC_B200_NO_FRAME_SHORTCUTS": "1", "PNEC_B200_SCF_DEFER": "0", "PNEC_B200_FRAME_CHUNKS": "1",
             "PNEC_B200_NO_LM_AHEAD": "1", "PNEC_B200_FUSED_ROUNDS_MAX_PAIRS": "0"}


@pytest.mark.parametrize("B,N", [(4608, 96), (640, 160), (1500, 64)])
def test_frame_solve_devices_are_exact_at_bench_sizes(monkeypatch, B, N):
    """Scan cache, fixed-point skipping, two-pass SCF, chunking over streams and the rotation chain
    running ahead must not change a single bit: B >= 4096 (4 chunks), 1024 <= B < 4096 (2 chunks),
    513 <= B < 1024 (per-round kernels, one chunk) against the plain sequence of launches.  The
    switches are read when a handle is created, hence the fresh handles."""
    batch = syn.make_batch(B, N, seed=300 + N)
    args = (dev(batch.bvs_host), dev(batch.bvs_target), dev(batch.covs_target), dev(batch.init_poses))
    for k in PLAIN_ENV:
        monkeypatch.delenv(k, raising=False)
    fast = api.Handle(0).frame_solve_batch(*args, api.default_frame_opts(use_ransac=0), n_per_problem=N)
    for k, v in PLAIN_ENV.items():
        monkeypatch.setenv(k, v)
    plain = api.Handle(0).frame_solve_batch(*args, api.default_frame_opts(use_ransac=0), n_per_problem=N)
    np.testing.assert_array_equal(fast.es_poses.cpu().numpy(), plain.es_poses.cpu().numpy())
    np.testing.assert_array_equal(fast.poses.cpu().numpy(), plain.poses.cpu().numpy())
    np.testing.assert_array_equal(fast.iterations.cpu().numpy(), plain.iterations.cpu().numpy())
    np.testing.assert_array_equal(fast.status.cpu().numpy(), plain.status.cpu().numpy())


@pytest.mark.parametrize("n,camera", [(100, syn.OMNIDIRECTIONAL), (256, syn.PINHOLE)])
def test_frame_solve_tight_allowance(handle, n, camera):
    """At N >= 100 at most 1 % of a batch may be excluded as ill posed (measured: 0.3-0.7 %,
    profiles/frame_parity_stress_r01.jsonl), and the excluded pairs are bounded too."""
    B = 600
    batch = syn.make_batch(B, n, seed=500 + n, camera=camera)
    fo = oracle.default_frame_opts(use_ransac=0)
    run = lambda bt: oracle.frame_solve_batch(bt.bvs_host, bt.bvs_target, bt.covs_target, bt.init_poses, fo,
                                              n_per_problem=n, num_threads=oracle.max_threads())
    ref, ref_es = run(batch)
    ref_p, ref_es_p = run(perturbed(batch))
    res = handle.frame_solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses,
                                   api.default_frame_opts(use_ransac=0), n_per_problem=n)
    check_against_oracle(res.es_poses, ref_es, ref_es_p, 0.99, "eigensolver stage")
    check_against_oracle(res.poses, ref, ref_p, 0.99, "frame solve")


# ------------------------------------------------------- the VO loop's start pose


def vo_start_poses(batch):
    """frame_processing.cc:97-102: Sophus::SE3d(prev_rel_rotation, Vector3d(0, 0, 0)) -- the
    previous pair's relative rotation (identity for the first pair) and a ZERO translation."""
    init = np.zeros_like(batch.init_poses)
    init[0, 3] = 1.0
    init[1:, :4] = batch.gt_poses[:-1, :4]
    return init


@pytest.mark.parametrize("kw", [dict(), dict(weighted_iterations=0), dict(use_nec=1)],
                         ids=["default", "ceres_from_start", "nec"])
def test_c4_miniature_with_the_vo_start_pose(handle, kw):
    """BASELINE config C4 in miniature: ragged KITTI-shaped counts, pinhole camera, 20 LM iterations,
    started where the reference's VO loop starts.  `weighted_iterations=0` hands that start --
    translation (0, 0, 0), i.e. theta = phi = 0 by AnglesFromVec's zero branch -- straight to the
    refinement."""
    counts = syn.kitti_like_counts(20, seed=11)
    batch = syn.make_batch(len(counts), 0, seed=401, camera=syn.PINHOLE, counts=counts)
    init = vo_start_poses(batch)
    fo = dict(max_num_iterations=20, **kw)
    run = lambda bt: oracle.frame_solve_batch(bt.bvs_host, bt.bvs_target, bt.covs_target, init,
                                              oracle.default_frame_opts(use_ransac=0, **fo), offsets=batch.offsets,
                                              num_threads=oracle.max_threads())
    ref, ref_es = run(batch)
    ref_p, _ = run(perturbed(batch))
    res = handle.frame_solve_batch(dev(batch.bvs_host), dev(batch.bvs_target), dev(batch.covs_target), dev(init),
                                   api.default_frame_opts(use_ransac=0, **fo), offsets=batch.offsets)
    # a refinement started at the wrong rotation AND at the pole is a long, sensitive descent: fewer
    # pairs reproduce under one ulp there (the others are still bounded by the oracle's self-difference)
    check_against_oracle(res.poses.cpu().numpy(), ref, ref_p, 0.5 if "weighted_iterations" in kw else 0.85,
                         f"C4 miniature {kw}")
    assert np.isfinite(res.poses.cpu().numpy()).all()


# --------------------------------------------------------------- chart poles


@pytest.mark.parametrize("vname", list(VARIANTS))
@pytest.mark.parametrize("pole", [(0.0, 0.0, 1.0), (0.0, 0.0, 0.0), (0.0, 0.0, -1.0)], ids=["+ez", "zero", "-ez"])
def test_solve_and_eval_started_at_a_chart_pole(handle, vname, pole):
    """AnglesFromVec (common.cc:103-116) at t = +e_z and t = 0 (the VO loop's start): theta = 0,
    phi = 0 and d t / d phi = 0 exactly, so the phi column of the Jacobian vanishes and the LM step
    leans on min_lm_diagonal.  Solve (poses, iterations, status) and the fused evaluation against the
    oracle.  At t = -e_z (theta = pi, not special-cased by the reference) sin(theta) is 1.2e-16
    instead of 0: the phi column is rounding noise, and the first step moves phi by that noise over
    min_lm_diagonal / radius = 1e-10 -- the reference's own trajectory is decided by garbage there, so
    the solve is only checked for its outcome (finite, same final cost), the evaluation exactly."""
    variant = VARIANTS[vname]
    B, N = 24, 150
    b = syn.with_host_covariances(syn.make_batch(B, N, seed=611))
    ct = None if variant == api.NEC else b.covs_target
    ch = b.covs_host if variant == api.SYMMETRIC else None
    init = b.init_poses.copy()
    init[:, 4:] = pole
    south = pole[2] < 0
    res = handle.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(ct), dev(ch), dev(init),
                             api.default_opts(variant), n_per_problem=N)
    run = lambda bt: oracle.solve_batch(bt.bvs_host, bt.bvs_target, ct, ch, init, oracle.default_opts(variant),
                                        n_per_problem=N, num_threads=oracle.max_threads())
    ref, info = run(b)
    ref_p, info_p = run(perturbed(b))
    poses = res.poses.cpu().numpy()
    assert np.isfinite(poses).all()
    if south:
        # some of these descents end in poor local minima (cost 1e5 .. 1e7 instead of ~2e2), a different
        # one per implementation; those that reach the basin of the solution agree on its cost
        cost, rcost = res.cost.cpu().numpy(), info["final_cost"]
        good = (cost < 10 * np.median(rcost)) & (rcost < 10 * np.median(rcost))
        assert good.mean() >= 0.5
        np.testing.assert_allclose(cost[good], rcost[good], rtol=1e-4)
        assert (res.cost.cpu().numpy() <= res.initial_cost.cpu().numpy() * (1 + 1e-12)).all()
    else:
        # HOST started this far off wanders for up to 50 iterations and amplifies one ulp to 1e-3 rad on
        # some pairs: those are bounded by the oracle's own self-difference, the others meet the bar
        ok = check_against_oracle(poses, ref, ref_p, 0.5 if variant == api.HOST else 0.9, f"pole start {vname}")
        ok &= info["iterations"] == info_p["iterations"]
        assert np.array_equal(res.iterations.cpu().numpy()[ok], info["iterations"][ok])
        assert np.array_equal(res.status.cpu().numpy()[ok], info["status"][ok])
    ev = handle.eval_batch(dev(b.bvs_host), dev(b.bvs_target), dev(ct), dev(ch), dev(init), variant, 1e-13,
                           n_per_problem=N)
    for i in range(4):
        s, e = b.range(i)
        o = oracle.evaluate(variant, b.bvs_host[s:e], b.bvs_target[s:e], None if ct is None else ct[s:e],
                            None if ch is None else ch[s:e], 1e-13, init[i], oracle.JAC_ANALYTIC)
        assert ev.cost[i].item() == pytest.approx(o.cost, rel=1e-9)
        gmax = np.abs(o.gradient).max()
        np.testing.assert_allclose(ev.gradient[i].cpu().numpy(), o.gradient, rtol=0, atol=1e-9 * gmax)
        np.testing.assert_allclose(ev.jtj[i].cpu().numpy(), o.jtj, rtol=0, atol=1e-9 * np.abs(o.jtj).max())
        # the phi column vanishes at the pole (exactly at theta = 0, to rounding at theta = pi)
        assert abs(ev.gradient[i, 1].item()) <= (1e-15 * gmax if south else 0.0)
