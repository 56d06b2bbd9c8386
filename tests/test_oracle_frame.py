"""CPU: the oracle's restatement of the stages of PNEC::Solve in front of the refinement
(oracle/pnec_oracle_frame.c) — pinned where something independent exists in this image:

* the Levenberg-Marquardt (Eigen's port of MINPACK lmdif, as opengv's eigensolver configures it)
  against the Fortran MINPACK behind scipy.optimize.leastsq, on the oracle's own residuals;
* M(c) and lambda_min against numpy (np.cross, eigvalsh), the analytic gradient against central
  differences of eigvalsh;
* the weights against the reference's numpy energy denominators (scripts/pnec/common.py through the
  committed golden file) and the literal formula of common.cc:183-208.

That this IS what opengv executes stays unpinned: opengv is neither in the reference tree nor in
this image (oracle/pnec_oracle_frame.c header).
"""
import os

import numpy as np
import pytest
from scipy.optimize import leastsq

import oracle
from conftest import direction_angle, rotation_angle
from pnec_b200 import synthetic as syn


@pytest.fixture(scope="module")
def batch():
    return syn.make_batch(12, 150, seed=3, noise_level=1.0)


def reduced_rotation(c):
    x, y, z = c
    return np.array([[1 + x * x - y * y - z * z, 2 * (x * y - z), 2 * (x * z + y)],
                     [2 * (x * y + z), 1 - x * x + y * y - z * z, 2 * (y * z - x)],
                     [2 * (x * z - y), 2 * (y * z + x), 1 - x * x - y * y + z * z]])


def m_numpy(f1, f2, c, w=None):
    f2w = f2 if w is None else f2 * np.sqrt(w)[:, None]
    n = np.cross(f1, (reduced_rotation(c) @ f2w.T).T)
    return n.T @ n


def test_m_and_smallest_eigenvalue_match_numpy(batch):
    rng = np.random.default_rng(0)
    for k in range(batch.num_problems):
        f1, f2, ct, _ = batch.problem(k)
        c = rng.uniform(-0.3, 0.3, 3)
        w = rng.uniform(0.5, 2.0, f1.shape[0]) if k % 2 else None
        ev, jac, M = oracle.es_smallest_ev(f1, f2, c, w)
        Mn = m_numpy(f1, f2, c, w)
        np.testing.assert_allclose(M, Mn, rtol=0, atol=1e-13 * np.abs(Mn).max())
        assert abs(ev - np.linalg.eigvalsh(Mn)[0]) <= 1e-12 * np.abs(Mn).max()
        g = np.zeros(3)
        for j in range(3):
            h = 1e-6
            cp, cm = c.copy(), c.copy()
            cp[j] += h
            cm[j] -= h
            g[j] = (np.linalg.eigvalsh(m_numpy(f1, f2, cp, w))[0] - np.linalg.eigvalsh(m_numpy(f1, f2, cm, w))[0]) / (2 * h)
        np.testing.assert_allclose(jac, g, rtol=1e-6, atol=1e-7 * np.abs(g).max())


def test_lm_is_minpack_lmdif(batch):
    """Same iterates as Fortran MINPACK (scipy.optimize.leastsq, Dfun=None -> lmdif): identical x,
    identical termination code and function-evaluation count."""
    eps = np.finfo(float).eps
    for k in range(batch.num_problems):
        f1, f2, _, _ = batch.problem(k)
        x0 = batch.init_poses[k, :3] / batch.init_poses[k, 3]
        fun = lambda x: oracle.es_smallest_ev(f1, f2, x)[1]
        xs, _, infod, _, ier = leastsq(fun, x0, ftol=5e-5, xtol=10 * eps, gtol=0.0, maxfev=100, epsfcn=None,
                                       factor=100, full_output=True)
        xo, info, nfev = oracle.es_lm(f1, f2, x0, maxfev=100, fev_per_jacobian=3)
        np.testing.assert_allclose(xo, xs, rtol=0, atol=1e-13)
        assert info == ier and nfev == infod["nfev"], (k, info, ier, nfev, infod["nfev"])


def test_lm_minpack_other_settings(batch):
    eps = np.finfo(float).eps
    f1, f2, _, _ = batch.problem(0)
    x0 = np.array([0.3, -0.2, 0.25])  # far start: exercises the step-bound logic of lmpar
    fun = lambda x: oracle.es_smallest_ev(f1, f2, x)[1]
    for ftol, xtol, factor, maxfev in [(1e-8, 1e-8, 100.0, 400), (5e-5, 10 * eps, 0.1, 400), (1e-12, 1e-12, 1.0, 30)]:
        xs, _, infod, _, ier = leastsq(fun, x0, ftol=ftol, xtol=xtol, gtol=0.0, maxfev=maxfev, factor=factor,
                                       full_output=True)
        xo, info, nfev = oracle.es_lm(f1, f2, x0, ftol=ftol, xtol=xtol, factor=factor, maxfev=maxfev,
                                      fev_per_jacobian=3)
        np.testing.assert_allclose(xo, xs, rtol=0, atol=1e-12)
        assert info == ier and nfev == infod["nfev"]


def test_eigensolver_finds_the_rotation_on_noise_free_data():
    b = syn.make_batch(8, 60, seed=9, noise_level=1e-9)
    for k in range(b.num_problems):
        f1, f2, _, _ = b.problem(k)
        q, info = oracle.eigensolver(f1, f2, b.init_poses[k])
        assert rotation_angle(np.r_[q, 0, 0, 1], b.gt_poses[k]) < 1e-7
        assert abs(info.smallest_ev) < 1e-12
        pose, _ = oracle.nec_eigensolver_pose(f1, f2, b.init_poses[k])
        if np.linalg.norm(b.gt_poses[k, 4:]) > 0:
            assert direction_angle(pose[4:], b.gt_poses[k, 4:]) < 1e-5


def test_weights_follow_common_cc(batch):
    f1, f2, ct, _ = batch.problem(1)
    pose = batch.gt_poses[1]
    w = oracle.weights(f1, ct, pose, 1e-13)
    R = syn.quaternion_to_matrix(pose[:4])
    t = pose[4:]
    for i in range(0, f1.shape[0], 17):
        S = ct[i].reshape(3, 3).T  # column-major storage
        tt = (t @ syn.skew(f1[i]) @ R)
        assert w[i] == pytest.approx(1e-8 / (tt @ S @ tt + 1e-13), rel=1e-12)


def test_weighted_eigensolver_iterations_are_consistent(batch):
    """One weighted iteration by hand (weights -> eigensolver -> SCF) equals weighted_eigensolver(2)."""
    f1, f2, ct, _ = batch.problem(2)
    es, _ = oracle.nec_eigensolver_pose(f1, f2, batch.init_poses[2])
    w = oracle.weights(f1, ct, es, 1e-13)
    q, _ = oracle.eigensolver(f1, f2, es, w)
    t, _ = oracle.scf_translation(f1, f2, ct, np.r_[q, es[4:]], 1e-13, 500, 10)
    pose = oracle.weighted_eigensolver(f1, f2, ct, es, weighted_iterations=2)
    np.testing.assert_allclose(pose[:4], q, atol=1e-15)
    np.testing.assert_allclose(pose[4:], t, atol=1e-15)


def test_frame_solve_option_matrix(batch):
    """PNEC::Solve's branches (pnec.cc:93-123)."""
    n = batch.n_per_problem
    args = (batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses)
    es_only, es = oracle.frame_solve_batch(*args, oracle.default_frame_opts(use_ransac=0, use_nec=1, use_ceres=0), n_per_problem=n)
    np.testing.assert_array_equal(es_only, es)
    wi1, _ = oracle.frame_solve_batch(*args, oracle.default_frame_opts(use_ransac=0, weighted_iterations=1, use_ceres=0), n_per_problem=n)
    np.testing.assert_array_equal(wi1, es)
    wi0, _ = oracle.frame_solve_batch(*args, oracle.default_frame_opts(use_ransac=0, weighted_iterations=0, use_ceres=0), n_per_problem=n)
    np.testing.assert_allclose(wi0, batch.init_poses, atol=1e-15)
    # weighted_iterations = 0 + Ceres == the plain refinement of the start pose
    ref, _ = oracle.solve_batch(batch.bvs_host, batch.bvs_target, batch.covs_target, None, batch.init_poses,
                                oracle.default_opts(oracle.TARGET), n_per_problem=n)
    got, _ = oracle.frame_solve_batch(*args, oracle.default_frame_opts(use_ransac=0, weighted_iterations=0), n_per_problem=n)
    np.testing.assert_allclose(got, ref, atol=1e-14)
    # the full pipeline ends closer to the ground truth than the NEC eigensolver alone, on average
    full, _ = oracle.frame_solve_batch(*args, oracle.default_frame_opts(use_ransac=0), n_per_problem=n)
    e_es = np.mean([rotation_angle(a, g) for a, g in zip(es, batch.gt_poses)])
    e_full = np.mean([rotation_angle(a, g) for a, g in zip(full, batch.gt_poses)])
    assert e_full < e_es


def test_oracle_reproduces_committed_frame_fixtures(golden_frame):
    """tests/golden/oracle_frame.npz (tests/golden/make_golden.py) guards the restatement against drift."""
    g = golden_frame
    for name in ("omni_n200", "pinhole_n96"):
        n = int(g[f"{name}/n"])
        f1, f2, cov, init = g[f"{name}/f1"], g[f"{name}/f2"], g[f"{name}/cov"], g[f"{name}/init"]
        for k in (0, 7, 23):
            q, _ = oracle.eigensolver(f1[k * n:(k + 1) * n], f2[k * n:(k + 1) * n], init[k])
            assert rotation_angle(np.r_[q, 0, 0, 1], np.r_[g[f"{name}/es_quat"][k], 0, 0, 1]) < 1e-12
        poses, es = oracle.frame_solve_batch(f1, f2, cov, init, oracle.default_frame_opts(use_ransac=0), n_per_problem=n,
                                             num_threads=oracle.max_threads())
        stable = np.array([rotation_angle(a, b) < 1e-8 and direction_angle(a[4:], b[4:]) < 1e-8
                           for a, b in zip(g[f"{name}/default/poses"], g[f"{name}/default/poses_ulp"])])
        assert stable.mean() > 0.8
        for k in np.nonzero(stable)[0]:
            assert rotation_angle(poses[k], g[f"{name}/default/poses"][k]) < 1e-9
            assert direction_angle(poses[k][4:], g[f"{name}/default/poses"][k][4:]) < 1e-9


def _with_outliers(b, fraction, seed):
    rng = np.random.default_rng(seed)
    truth = np.ones(b.total, bool)
    for k in range(b.num_problems):
        s, e = b.range(k)
        bad = s + rng.choice(e - s, int(round(fraction * (e - s))), replace=False)
        v = rng.standard_normal((len(bad), 3))
        b.bvs_target[bad] = v / np.linalg.norm(v, axis=1, keepdims=True)
        truth[bad] = False
    return truth


def test_ransac_restatement_separates_outliers():
    """The restated opengv RANSAC over the eigensolver (pnec.cc:239-272) on pairs with 25 % gross
    outliers finds the inlier set, stops after the number of iterations opengv's formula prescribes,
    and recovers the rotation the plain eigensolver loses."""
    b = syn.make_batch(4, 300, seed=41, noise_level=0.25)
    truth_all = _with_outliers(b, 0.25, 0)
    for k in range(b.num_problems):
        f1, f2, _, _ = b.problem(k)
        s, e = b.range(k)
        truth = truth_all[s:e]
        pose, mask, iters = oracle.ransac_eigensolver(f1, f2, b.init_poses[k], pair_index=k)
        assert (mask & ~truth).sum() <= 1 and (mask & truth).sum() >= 0.95 * truth.sum()
        w = mask.mean()
        assert iters >= np.log(0.01) / np.log(1 - w ** 10) - 1e-9 and iters <= 5001
        plain, _ = oracle.nec_eigensolver_pose(f1, f2, b.init_poses[k])
        assert rotation_angle(pose, b.gt_poses[k]) < 1e-3 < rotation_angle(plain, b.gt_poses[k])
        # deterministic in (seed, pair index); a different seed draws different samples
        again, mask2, _ = oracle.ransac_eigensolver(f1, f2, b.init_poses[k], pair_index=k)
        np.testing.assert_array_equal(pose, again)
        np.testing.assert_array_equal(mask, mask2)
        other, _, it2 = oracle.ransac_eigensolver(f1, f2, b.init_poses[k], pair_index=k, seed=77)
        assert not np.array_equal(other, pose) and rotation_angle(other, pose) < 1e-3


def test_ransac_sequential_state_is_statistically_equivalent():
    """opengv carries two pieces of state from one iteration to the next -- the shuffled index array of
    drawIndexSample and, through the adapter, the previous model's rotation as the next start
    (`sequential = 1`).  The CUDA path implements the independent-hypotheses form (`sequential = 0`).
    On the same pairs the two reach the same inlier sets (up to borderline correspondences), rotations
    within the spread between two seeds of either, and the same distribution of iteration counts."""
    b = syn.make_batch(24, 256, seed=43, noise_level=0.5)
    truth = _with_outliers(b, 0.2, 1)
    its = {0: [], 1: []}
    agree, rot = [], []
    for k in range(b.num_problems):
        f1, f2, _, _ = b.problem(k)
        res = {q: oracle.ransac_eigensolver(f1, f2, b.init_poses[k], pair_index=k, sequential=q) for q in (0, 1)}
        for q in (0, 1):
            its[q].append(res[q][2])
            s, e = b.range(k)
            assert (res[q][1] & ~truth[s:e]).sum() <= 2
        agree.append((res[0][1] == res[1][1]).mean())
        rot.append(rotation_angle(res[0][0], res[1][0]))
    assert np.mean(agree) > 0.98 and np.median(rot) < 2e-4
    assert abs(np.median(its[0]) - np.median(its[1])) <= 0.35 * np.median(its[0])


def test_ransac_degenerate_inputs():
    """Fewer correspondences than the sample size: no model (the reference would index an empty vector);
    max_iterations caps the loop at max + 1 hypotheses."""
    b = syn.make_batch(1, 9, seed=3)
    f1, f2, _, _ = b.problem(0)
    pose, mask, iters = oracle.ransac_eigensolver(f1, f2, b.init_poses[0])
    assert iters == 0 and mask.sum() == 0
    np.testing.assert_allclose(pose[:4], b.init_poses[0, :4] / np.linalg.norm(b.init_poses[0, :4]), atol=1e-15)
    b = syn.make_batch(1, 200, seed=4)
    _with_outliers(b, 0.6, 2)
    f1, f2, _, _ = b.problem(0)
    _, _, iters = oracle.ransac_compute_model(f1, f2, b.init_poses[0], max_ransac_iterations=7)
    assert iters == 8


def test_ransac_fixture_is_what_the_oracle_computes():
    """tests/golden/oracle_ransac.npz (minted by make_golden.py) pins the oracle's RANSAC path against
    accidental change: inlier masks and iteration counts exactly, poses to 1e-9 (libm differences)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_ransac.npz"))
    N = int(g["n"])
    poses, es, mask, ni, it = oracle.frame_solve_batch(g["f1"], g["f2"], g["cov"], g["init"],
                                                       oracle.default_frame_opts(), n_per_problem=N,
                                                       return_ransac=True)
    stable = g["iterations"] == g["iterations_ulp"]
    assert stable.mean() > 0.8
    same = (it == g["iterations"]) & (ni == g["num_inliers"])
    assert same[stable].all()
    for k in np.nonzero(stable & same)[0]:
        if np.array_equal(g["mask"][k * N:(k + 1) * N], g["mask_ulp"][k * N:(k + 1) * N]):
            np.testing.assert_array_equal(mask[k * N:(k + 1) * N], g["mask"][k * N:(k + 1) * N])
            assert rotation_angle(poses[k], g["poses"][k]) < 1e-9
