"""CPU: the oracle against the reference's golden vectors and against itself."""
import numpy as np
import pytest

import oracle
from conftest import direction_angle, max_pose_diff, rotation_angle
from pnec_b200 import synthetic as syn


def test_functors_match_reference_numpy_energy(golden_energy):
    """oracle functors == scripts/pnec/common.py:13-86 (reference outputs, committed)."""
    g = golden_energy
    for k in range(g["f1"].shape[0]):
        p = g["poses"][k]
        e = oracle.energy(oracle.TARGET, g["f1"][k], g["f2"][k], g["cov_colmajor"][k], None,
                          float(g["regs"][k]), p[:4], p[4:])
        assert e == pytest.approx(float(g["pnec_energy_rotations"][k]), rel=1e-12)
        assert e == pytest.approx(float(g["pnec_energy_translations"][k]), rel=1e-12)
        en = oracle.energy(oracle.NEC, g["f1"][k], g["f2"][k], None, None, 0.0, p[:4], p[4:])
        assert en == pytest.approx(float(g["nec_energy_rotations"][k]), rel=1e-12)


def test_unscented_transform_matches_reference(golden_ut):
    """synthetic.unscented_transform == scripts/pnec/math.py:73-123 on the cases where the
    python and C++ references coincide (diagonal local covariance)."""
    u = golden_ut
    mine = syn.unscented_transform(u["mus"], u["covs_omni"], syn.OMNIDIRECTIONAL)
    np.testing.assert_allclose(mine, u["ut_omni"], rtol=0, atol=1e-12 * np.abs(u["ut_omni"]).max())
    mine = syn.unscented_transform(u["mus_pinhole"], u["covs_local"], syn.PINHOLE)
    np.testing.assert_allclose(mine, u["ut_pinhole"], rtol=0, atol=1e-12 * np.abs(u["ut_pinhole"]).max())
    ez = np.broadcast_to(np.array([0.0, 0.0, 1.0]), u["mus"].shape)
    rbp = syn.rotation_between_points(ez, syn._normalize(u["mus"]))
    np.testing.assert_allclose(rbp, u["rotation_between_points"], atol=1e-14)


def test_oracle_unscented_transform_matches_reference(golden_ut):
    """C restatement of common.cc:467-525 vs scripts/pnec/math.py:73-123 (diagonal local
    covariance, where the python and C++ references coincide) and vs the numpy generator."""
    u = golden_ut
    col = lambda M: np.swapaxes(M, -1, -2).reshape(-1, 9)
    o = oracle.unscented_transform(u["mus"], col(u["covs_omni"]), None, 1.0, oracle.OMNIDIRECTIONAL)
    np.testing.assert_allclose(o, col(u["ut_omni"]), rtol=0, atol=1e-12 * np.abs(u["ut_omni"]).max())
    o = oracle.unscented_transform(u["mus_pinhole"], col(u["covs_local"]), None, 1.0, oracle.PINHOLE)
    np.testing.assert_allclose(o, col(u["ut_pinhole"]), rtol=0, atol=1e-12 * np.abs(u["ut_pinhole"]).max())
    rng = np.random.default_rng(3)
    mus = syn._uniform_sphere(rng, (64,)) * 800.0
    mus[:, 2] = np.abs(mus[:, 2])
    c3 = np.zeros((64, 3, 3))
    c3[:, :2, :2] = syn.sample_covariances_2d(rng, (1, 64), 1.0, "anisotropic_inhomogenous")[0]
    ez = np.broadcast_to(np.array([0.0, 0.0, 1.0]), mus.shape)
    Rp = syn.rotation_between_points(ez, syn._normalize(mus))
    c3o = Rp @ c3 @ np.swapaxes(Rp, -1, -2)
    a = oracle.unscented_transform(mus, col(c3o), None, 1.0, oracle.OMNIDIRECTIONAL)
    np.testing.assert_allclose(a, col(syn.unscented_transform(mus, c3o, syn.OMNIDIRECTIONAL)), rtol=0, atol=1e-20)


def test_oracle_keypoints_unproject_is_unproject_plus_unscented():
    """KeyPoint::Unproject (keypoints.cc:49-62) == Unproject + pinhole unscented transform of
    (x, y, 1) with the 2x2 covariance embedded in a 3x3."""
    rng = np.random.default_rng(5)
    n = 40
    pts = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n)], -1)
    c2 = syn.sample_covariances_2d(rng, (1, n), 0.8, "anisotropic_inhomogenous")[0]
    K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.2157], [0, 0, 1.0]])
    Kinv = np.linalg.inv(K)
    bvs, covs = oracle.keypoints_unproject(pts, np.swapaxes(c2, -1, -2).reshape(n, 4), Kinv.T.reshape(9))
    mu = np.concatenate([pts, np.ones((n, 1))], -1)
    ray = mu @ Kinv.T
    np.testing.assert_allclose(bvs, ray / np.linalg.norm(ray, axis=1, keepdims=True), atol=1e-15)
    c3 = np.zeros((n, 3, 3))
    c3[:, :2, :2] = c2
    # numpy generator with K_inv applied by hand: sigma points in pixel space, rays through K_inv
    ref = oracle.unscented_transform(mu, np.swapaxes(c3, -1, -2).reshape(n, 9), Kinv.T.reshape(9), 1.0, oracle.PINHOLE)
    np.testing.assert_array_equal(covs, ref)
    w = np.linalg.eigvalsh(covs.reshape(n, 3, 3))
    assert (w[:, 2] > 0).all() and (np.abs(w[:, 0]) < 1e-5 * w[:, 2]).all()  # ~rank 2: tangent plane


def test_oracle_scf_pieces_match_reference(golden_scf):
    """obj_fun and fibonacci_sphere vs scripts/pnec/scf.py (committed outputs).  The C++
    fibonacci_sphere casts to float (scf.cc:59,63), the python does not: 1e-6 agreement."""
    g = golden_scf
    for x, o in zip(g["X"], g["obj_fun"]):
        assert oracle.scf_objective(g["Ai"], g["Bi"], x) == pytest.approx(float(o), rel=1e-13)
    pts = oracle.fibonacci_sphere(500)
    np.testing.assert_allclose(pts, g["fibonacci_500"], atol=2e-6)
    np.testing.assert_allclose(np.linalg.norm(pts, axis=1), 1.0, atol=1e-15)
    assert pts[0].tolist() == [0.0, 1.0, 0.0]


def test_oracle_translation_given_rotation():
    """SCF translation (pnec.cc:317-343 + scf.cc) and NEC translation (common.cc:127-181) at
    the ground-truth rotation recover the ground-truth direction up to the noise level."""
    b = syn.make_batch(4, 300, seed=8)
    for i in range(4):
        f1, f2, ct, _ = b.problem(i)
        t, cost = oracle.scf_translation(f1, f2, ct, b.gt_poses[i], 1e-13, 500, 10)
        assert np.linalg.norm(t) == pytest.approx(1.0, abs=1e-14)
        assert direction_angle(t, b.gt_poses[i][4:]) < 5e-3
        # the SCF fixed point does not depend on where the scan starts from
        p2 = b.gt_poses[i].copy()
        p2[4:] = [0.3, -0.2, 0.93]
        t2, cost2 = oracle.scf_translation(f1, f2, ct, p2, 1e-13, 500, 10)
        assert direction_angle(t, t2) < 1e-9 and cost2 == pytest.approx(cost, rel=1e-12)
        tn, M = oracle.nec_translation(f1, f2, b.gt_poses[i])
        assert direction_angle(tn, b.gt_poses[i][4:]) < 2e-2
        # ComposeM skips correspondence 0 (common.cc:131)
        R = syn.quaternion_to_matrix(b.gt_poses[i][:4])
        n = np.cross(f1[1:], f2[1:] @ R.T)
        Mref = n.T @ n
        np.testing.assert_allclose(M, [Mref[0, 0], Mref[0, 1], Mref[0, 2], Mref[1, 1], Mref[1, 2], Mref[2, 2]],
                                   rtol=1e-12, atol=1e-18)
        w, V = np.linalg.eigh(Mref)
        assert direction_angle(tn, V[:, 0]) < 1e-9


VARIANTS = {"nec": oracle.NEC, "target": oracle.TARGET, "host": oracle.HOST,
            "symmetric": oracle.SYMMETRIC}
CASES = ["c1_iso_omni_n100", "c2_aniso_omni_n512", "aniso_pinhole_n64", "aniso_omni_n10"]


def _case(g, name, vname):
    variant = VARIANTS[vname]
    ct = None if variant == oracle.NEC else g[f"{name}/cov_t"]
    ch = g[f"{name}/cov_h"] if variant == oracle.SYMMETRIC else None
    return variant, g[f"{name}/f1"], g[f"{name}/f2"], ct, ch, g[f"{name}/init"], int(g[f"{name}/n"])


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("vname", list(VARIANTS))
def test_oracle_regression_fixtures(golden_solutions, name, vname):
    """The oracle reproduces its own committed outputs (guards against drift)."""
    g = golden_solutions
    variant, f1, f2, ct, ch, init, n = _case(g, name, vname)
    poses, info = oracle.solve_batch(f1, f2, ct, ch, init, oracle.default_opts(variant), n_per_problem=n)
    r, t = max_pose_diff(poses, g[f"{name}/{vname}/poses"])
    assert r < 1e-12 and t < 1e-7
    assert np.array_equal(info["iterations"], g[f"{name}/{vname}/iterations"])
    assert np.array_equal(info["status"], g[f"{name}/{vname}/status"])
    np.testing.assert_allclose(info["final_cost"], g[f"{name}/{vname}/final_cost"], rtol=1e-10)


@pytest.mark.parametrize("vname", list(VARIANTS))
def test_numeric_and_analytic_jacobians_agree(golden_solutions, vname):
    """Central differences (what Ceres does for the reference) vs the closed form."""
    g = golden_solutions
    name = "c1_iso_omni_n100"
    variant, f1, f2, ct, ch, init, n = _case(g, name, vname)
    for i in range(init.shape[0]):
        s = slice(i * n, (i + 1) * n)
        a = oracle.evaluate(variant, f1[s], f2[s], None if ct is None else ct[s],
                            None if ch is None else ch[s], 1e-13, init[i], oracle.JAC_NUMERIC_CENTRAL)
        b = oracle.evaluate(variant, f1[s], f2[s], None if ct is None else ct[s],
                            None if ch is None else ch[s], 1e-13, init[i], oracle.JAC_ANALYTIC)
        assert a.cost == b.cost
        np.testing.assert_allclose(a.gradient, b.gradient, rtol=0, atol=2e-8 * np.abs(b.gradient).max())
        np.testing.assert_allclose(a.jtj, b.jtj, rtol=0, atol=2e-8 * np.abs(b.jtj).max())
    pn, infon = oracle.solve_batch(f1, f2, ct, ch, init, oracle.default_opts(variant), n_per_problem=n)
    pa, infoa = oracle.solve_batch(f1, f2, ct, ch, init,
                                   oracle.default_opts(variant, jacobian_mode=oracle.JAC_ANALYTIC),
                                   n_per_problem=n)
    r, t = max_pose_diff(pn, pa)
    assert r < 1e-9 and t < 1e-7
    assert np.array_equal(infon["iterations"], infoa["iterations"])


def test_converged_optimum_matches_scipy(golden_solutions):
    """Independent pin of the minimiser: with tolerances tightened, the oracle lands on the
    same optimum as scipy's MINPACK LM run on the same residual in a local chart."""
    scipy_opt = pytest.importorskip("scipy.optimize")
    g = golden_solutions
    name = "c2_aniso_omni_n512"
    variant, f1, f2, ct, ch, init, n = _case(g, name, "target")
    i = 0
    s = slice(i * n, (i + 1) * n)
    tight = oracle.default_opts(variant, function_tolerance=1e-16, parameter_tolerance=1e-14,
                                gradient_tolerance=1e-14, max_num_iterations=200)
    pose, info = oracle.solve(f1[s], f2[s], ct[s], None, init[i], tight)

    def residuals(x):
        th, ph, rx, ry, rz = x
        t = np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
        ang = np.linalg.norm([rx, ry, rz])
        dR = np.eye(3) if ang == 0 else syn.angle_axis_to_matrix(np.array(ang), np.array([rx, ry, rz]) / ang)
        R = dR @ syn.quaternion_to_matrix(init[i][:4])
        S = ct[s].reshape(-1, 3, 3).transpose(0, 2, 1)
        g2 = f2[s] @ R.T
        num = np.einsum("j,ij->i", t, np.cross(f1[s], g2))
        b = np.cross(t[None, :], f1[s]) @ R
        den = np.einsum("ij,ijk,ik->i", b, S, b) + 1e-13
        return num / np.sqrt(den)

    th0, ph0 = oracle.angles_from_vec(init[i][4:])
    sol = scipy_opt.least_squares(residuals, [th0, ph0, 0, 0, 0], method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15)
    th, ph, rx, ry, rz = sol.x
    ang = np.linalg.norm([rx, ry, rz])
    R = syn.angle_axis_to_matrix(np.array(ang), np.array([rx, ry, rz]) / ang) @ syn.quaternion_to_matrix(init[i][:4])
    p_scipy = np.concatenate([syn.matrix_to_quaternion(R), [np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)]])
    assert rotation_angle(pose, p_scipy) < 1e-8
    assert direction_angle(pose[4:], p_scipy[4:]) < 1e-7
    assert 0.5 * np.sum(sol.fun ** 2) == pytest.approx(info.final_cost, rel=1e-9)


def test_noise_free_data_has_zero_residual_at_ground_truth():
    b = syn.make_batch(3, 50, seed=9, noise_level=1e-30)
    for i in range(3):
        f1, f2, ct, _ = b.problem(i)
        e = oracle.energy(oracle.NEC, f1, f2, None, None, 0.0, b.gt_poses[i][:4], b.gt_poses[i][4:])
        assert e < 1e-25


def test_empty_problem_returns_start_pose():
    init = np.array([0.0, 0.0, 0.0, 2.0, 0.0, 3.0, 4.0])
    pose, info = oracle.solve(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 9)), None, init,
                              oracle.default_opts(oracle.TARGET))
    assert info.status == 7 and info.iterations == 0
    np.testing.assert_allclose(pose, [0, 0, 0, 1, 0, 0.6, 0.8], atol=1e-15)


def test_metrics_follow_reference_definitions():
    """RotationalDifference / TranslationalDifference, src/common/common.cc:210-235."""
    a = np.array([0, 0, 0, 1.0, 1, 0, 0])
    ang = 0.3
    b = np.array([0, 0, np.sin(ang / 2), np.cos(ang / 2), -1, 0, 0])
    assert oracle.rotational_difference_deg(a, b) == pytest.approx(np.rad2deg(ang), rel=1e-12)
    assert oracle.translational_difference_deg(a[4:], b[4:], True) == pytest.approx(0.0, abs=1e-6)
    assert oracle.translational_difference_deg(a[4:], b[4:], False) == pytest.approx(180.0)
    assert oracle.translational_difference_deg(np.zeros(3), b[4:], True) == pytest.approx(90.0)


def test_angles_from_vec_chart():
    """pnec::common::AnglesFromVec, src/common/common.cc:103-116, incl. the pole."""
    assert oracle.angles_from_vec([0.0, 0.0, 0.0]) == (0.0, 0.0)
    assert oracle.angles_from_vec([0.0, 0.0, 2.0]) == (0.0, 0.0)
    th, ph = oracle.angles_from_vec([0.0, 3.0, 0.0])
    assert th == pytest.approx(np.pi / 2) and ph == pytest.approx(np.pi / 2)


def test_threaded_batch_equals_serial(golden_solutions):
    g = golden_solutions
    variant, f1, f2, ct, ch, init, n = _case(g, "c1_iso_omni_n100", "target")
    p1, i1 = oracle.solve_batch(f1, f2, ct, ch, init, oracle.default_opts(variant), n_per_problem=n, num_threads=1)
    p2, i2 = oracle.solve_batch(f1, f2, ct, ch, init, oracle.default_opts(variant), n_per_problem=n, num_threads=4)
    assert np.array_equal(p1, p2) and np.array_equal(i1, i2)
