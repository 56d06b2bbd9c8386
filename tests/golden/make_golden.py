"""Generates tests/golden/*.npz.  Run HERE (the container that has /root/reference):

    python tests/golden/make_golden.py

Two kinds of fixtures:

1. reference_energy.npz / reference_ut.npz — outputs of the REFERENCE's own
   Python code, imported from /root/reference/scripts (numpy only):
     pnec.common.pnec_energy_rotations     scripts/pnec/common.py:13-37
     pnec.common.nec_energy_rotations      scripts/pnec/common.py:40-58
     pnec.common.pnec_energy_translations  scripts/pnec/common.py:61-86
     pnec.math.unscented_transform         scripts/pnec/math.py:73-123
     pnec.math.rotation_between_points     scripts/pnec/math.py:42-64
   They pin the oracle's residual functors (Target variant with reg, NEC) and
   the synthetic generator's unscented transform.  The reference cannot travel
   to the GPU box, hence the committed vectors.

2. oracle_solutions.npz — seeded synthetic frame pairs with the oracle's own
   outputs (start/final pose, iterations, status, cost; evaluation cost /
   gradient / JtJ).  The reference has no tests or golden vectors for the LM
   path, so these are regression fixtures minted by the
   oracle, not reference outputs: they guard the oracle against drift and give
   the GPU tests fixed known answers.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF_SCRIPTS = "/root/reference/scripts"


def reference_fixtures():
    sys.path.insert(0, REF_SCRIPTS)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from pnec.common import (nec_energy_rotations, pnec_energy_rotations,
                                 pnec_energy_translations)
        import pnec.math as rmath
    from pnec_b200 import synthetic as syn

    rng = np.random.default_rng(20221017)
    K, N = 6, 40
    out = {}
    batch = syn.make_batch(K, N, seed=3, noise_type="anisotropic_inhomogenous")
    f1 = batch.bvs_host.reshape(K, N, 3)
    f2 = batch.bvs_target.reshape(K, N, 3)
    # ABI layout is column-major; the python reference indexes sigma[row, col]
    sig = batch.covs_target.reshape(K, N, 3, 3).transpose(0, 1, 3, 2).copy()
    # make them slightly NON-symmetric copies too?  No: the reference assumes
    # symmetric covariances (common.cc:518-522 constructs them so).
    poses = batch.init_poses.copy()
    R = syn.quaternion_to_matrix(poses[:, :4])
    t = poses[:, 4:]
    regs = np.array([1e-13, 1e-10, 0.0, 1e-13, 1e-8, 1e-13])
    e_pnec = np.zeros(K)
    e_nec = np.zeros(K)
    e_pnec_t = np.zeros(K)
    for k in range(K):
        e_pnec[k] = pnec_energy_rotations(R[k][None, None], t[k], f1[k], f2[k], sig[k], regs[k])[0, 0]
        e_nec[k] = nec_energy_rotations(R[k][None, None], t[k], f1[k], f2[k])[0, 0]
        e_pnec_t[k] = pnec_energy_translations(t[k][None, None], R[k], f1[k], f2[k], sig[k], regs[k])[0, 0]
    out.update(f1=f1, f2=f2, cov_colmajor=batch.covs_target.reshape(K, N, 9), poses=poses,
               regs=regs, pnec_energy_rotations=e_pnec, nec_energy_rotations=e_nec,
               pnec_energy_translations=e_pnec_t)
    np.savez(os.path.join(HERE, "reference_energy.npz"), **out)

    # unscented transform + rotation_between_points.  The python reference takes
    # sigma-point offsets from ROWS of the Cholesky factor (math.py:102-107)
    # whereas the C++ takes COLUMNS (common.cc:496-505); they coincide for a
    # diagonal local covariance, which is what is pinned here, plus one
    # non-diagonal case against which only symmetric invariants are compared.
    M = 24
    mus = syn._uniform_sphere(rng, (M,)) * 800.0
    mus[:, 2] = np.abs(mus[:, 2])  # keep away from the -z singularity of the rotation
    d = rng.random((M, 2)) * 2.0 + 0.1
    covs_local = np.zeros((M, 3, 3))
    covs_local[:, 0, 0] = d[:, 0]
    covs_local[:, 1, 1] = d[:, 1]
    ut_omni = np.zeros((M, 3, 3))
    ut_pin = np.zeros((M, 3, 3))
    rbp = np.zeros((M, 3, 3))
    covs_omni = np.zeros((M, 3, 3))
    mus_pin = mus / mus[:, 2:3] * 800.0
    for i in range(M):
        rot = rmath.rotation_between_points(np.array([0.0, 0.0, 1.0]), mus[i] / np.linalg.norm(mus[i]))
        rbp[i] = rot
        covs_omni[i] = rot @ covs_local[i] @ rot.T
        ut_omni[i] = rmath.unscented_transform(mus[i], covs_omni[i], True, 1.0)
        ut_pin[i] = rmath.unscented_transform(mus_pin[i], covs_local[i], False, 1.0)
    np.savez(os.path.join(HERE, "reference_ut.npz"), mus=mus, mus_pinhole=mus_pin,
             covs_local=covs_local, covs_omni=covs_omni, rotation_between_points=rbp,
             ut_omni=ut_omni, ut_pinhole=ut_pin)


def reference_scf_fixtures():
    """scripts/pnec/scf.py: fibonacci_sphere (double arithmetic; the C++ of scf.cc:53-72 casts
    to float, so agreement is ~1e-7, not exact) and obj_fun (the sum of Rayleigh quotients)."""
    sys.path.insert(0, REF_SCRIPTS)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pnec.scf as rscf
    rng = np.random.default_rng(77)
    k = 12
    v = rng.standard_normal((k, 3))
    Ai = v[:, :, None] * v[:, None, :]
    L = rng.standard_normal((k, 3, 3))
    Bi = L @ np.swapaxes(L, -1, -2) + 0.1 * np.eye(3)
    X = rng.standard_normal((9, 3))
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    obj = rscf.obj_fun(X, Ai, Bi, n=3, k=k)
    np.savez(os.path.join(HERE, "reference_scf.npz"), fibonacci_500=rscf.fibonacci_sphere(500),
             Ai=Ai, Bi=Bi, X=X, obj_fun=obj)


def oracle_fixtures():
    import oracle
    from pnec_b200 import synthetic as syn

    cases = {
        # name: (B, N, camera, noise_type, noise_level, seed)
        "c1_iso_omni_n100": (4, 100, syn.OMNIDIRECTIONAL, "isotropic_homogenous", 1.0, 1),
        "c2_aniso_omni_n512": (4, 512, syn.OMNIDIRECTIONAL, "anisotropic_inhomogenous", 1.0, 2),
        "aniso_pinhole_n64": (4, 64, syn.PINHOLE, "anisotropic_inhomogenous", 0.5, 3),
        "aniso_omni_n10": (4, 10, syn.OMNIDIRECTIONAL, "anisotropic_inhomogenous", 2.0, 4),
    }
    out = {}
    for name, (B, N, cam, nt, nl, seed) in cases.items():
        b = syn.make_batch(B, N, seed=seed, camera=cam, noise_type=nt, noise_level=nl)
        b = syn.with_host_covariances(b, seed=seed + 100, camera=cam)
        out[f"{name}/f1"] = b.bvs_host
        out[f"{name}/f2"] = b.bvs_target
        out[f"{name}/cov_t"] = b.covs_target
        out[f"{name}/cov_h"] = b.covs_host
        out[f"{name}/init"] = b.init_poses
        out[f"{name}/gt"] = b.gt_poses
        out[f"{name}/n"] = np.int64(N)
        for vname, variant in (("nec", oracle.NEC), ("target", oracle.TARGET),
                               ("host", oracle.HOST), ("symmetric", oracle.SYMMETRIC)):
            ct = None if variant == oracle.NEC else b.covs_target
            ch = b.covs_host if variant == oracle.SYMMETRIC else None
            o = oracle.default_opts(variant)
            poses, info = oracle.solve_batch(b.bvs_host, b.bvs_target, ct, ch, b.init_poses, o,
                                             n_per_problem=N)
            out[f"{name}/{vname}/poses"] = poses
            out[f"{name}/{vname}/status"] = info["status"]
            out[f"{name}/{vname}/iterations"] = info["iterations"]
            out[f"{name}/{vname}/final_cost"] = info["final_cost"]
            out[f"{name}/{vname}/initial_cost"] = info["initial_cost"]
            ev_c, ev_g, ev_h = [], [], []
            for i in range(B):
                s, e = b.range(i)
                ev = oracle.evaluate(variant, b.bvs_host[s:e], b.bvs_target[s:e],
                                     None if ct is None else ct[s:e],
                                     None if ch is None else ch[s:e], 1e-13, b.init_poses[i],
                                     oracle.JAC_NUMERIC_CENTRAL)
                ev_c.append(ev.cost); ev_g.append(ev.gradient); ev_h.append(ev.jtj)
            out[f"{name}/{vname}/eval_cost"] = np.array(ev_c)
            out[f"{name}/{vname}/eval_gradient"] = np.array(ev_g)
            out[f"{name}/{vname}/eval_jtj"] = np.array(ev_h)
    np.savez_compressed(os.path.join(HERE, "oracle_solutions.npz"), **out)


def oracle_frame_fixtures():
    """oracle_frame.npz: the front stages of PNEC::Solve (oracle/pnec_oracle_frame.c) on seeded
    frame pairs: eigensolver rotations (plain and weighted), the NEC eigensolver pose, the
    weighted-eigensolver pose and the result of the whole pipeline for several option sets, plus
    the same outputs on inputs moved by one ulp (the tests only compare pairs on which the
    algorithm itself is stable under that perturbation).  Minted by the oracle (the LM inside is
    pinned against MINPACK, tests/test_oracle_frame.py); not reference outputs: opengv is not
    available."""
    import oracle
    from pnec_b200 import synthetic as syn

    out = {}
    configs = {"default": {}, "nec_ceres": {"use_nec": 1}, "es_then_ceres": {"weighted_iterations": 1},
               "weighted3_no_ceres": {"use_ceres": 0, "weighted_iterations": 3}}
    for name, (B, N, cam, seed) in {"omni_n200": (24, 200, syn.OMNIDIRECTIONAL, 301),
                                     "pinhole_n96": (24, 96, syn.PINHOLE, 302)}.items():
        b = syn.make_batch(B, N, seed=seed, camera=cam)
        rng = np.random.default_rng(seed)
        f1p = b.bvs_host * (1.0 + rng.uniform(-1, 1, b.bvs_host.shape) * 2.0 ** -52)
        f2p = b.bvs_target * (1.0 + rng.uniform(-1, 1, b.bvs_target.shape) * 2.0 ** -52)
        out[f"{name}/f1"], out[f"{name}/f2"], out[f"{name}/cov"] = b.bvs_host, b.bvs_target, b.covs_target
        out[f"{name}/init"], out[f"{name}/gt"], out[f"{name}/n"] = b.init_poses, b.gt_poses, np.int64(N)
        for tag, (f1, f2) in {"": (b.bvs_host, b.bvs_target), "_ulp": (f1p, f2p)}.items():
            es_q, w_q = np.zeros((B, 4)), np.zeros((B, 4))
            for k in range(B):
                s, e = b.range(k)
                es_q[k], _ = oracle.eigensolver(f1[s:e], f2[s:e], b.init_poses[k])
                w = oracle.weights(f1[s:e], b.covs_target[s:e], b.gt_poses[k], 1e-13)
                w_q[k], _ = oracle.eigensolver(f1[s:e], f2[s:e], b.init_poses[k], w)
            out[f"{name}/es_quat{tag}"] = es_q
            out[f"{name}/weighted_es_quat{tag}"] = w_q  # weights from the ground-truth poses
            for cname, kw in configs.items():
                poses, es = oracle.frame_solve_batch(f1, f2, b.covs_target, b.init_poses,
                                                     oracle.default_frame_opts(use_ransac=0, **kw), n_per_problem=N)
                out[f"{name}/{cname}/poses{tag}"] = poses
                out[f"{name}/es_pose{tag}"] = es
    np.savez_compressed(os.path.join(HERE, "oracle_frame.npz"), **out)


def oracle_ransac_fixtures():
    """oracle_ransac.npz: PNEC::Solve with the reference's default options (RANSAC over the eigensolver,
    pnec.cc:239-272) on seeded frame pairs, half of them with 20 % gross outliers, as minted by the
    oracle's restatement of opengv's Ransac<EigensolverSacProblem> (independent-hypotheses mode, seed 1):
    inlier masks, iteration counts, winning models, eigensolver and final poses, and the same on inputs
    moved by one ulp (pairs on which the algorithm is not stable under that are no parity cases)."""
    import oracle
    from pnec_b200 import synthetic as syn

    out = {}
    B, N = 32, 160
    b = syn.make_batch(B, N, seed=611, noise_level=0.5)
    rng = np.random.default_rng(612)
    for k in range(0, B, 2):
        bad = k * N + rng.choice(N, N // 5, replace=False)
        v = rng.standard_normal((len(bad), 3))
        b.bvs_target[bad] = v / np.linalg.norm(v, axis=1, keepdims=True)
    f1p = b.bvs_host * (1.0 + rng.uniform(-1, 1, b.bvs_host.shape) * 2.0 ** -52)
    f2p = b.bvs_target * (1.0 + rng.uniform(-1, 1, b.bvs_target.shape) * 2.0 ** -52)
    out["f1"], out["f2"], out["cov"], out["init"], out["n"] = b.bvs_host, b.bvs_target, b.covs_target, b.init_poses, np.int64(N)
    for tag, (f1, f2) in {"": (b.bvs_host, b.bvs_target), "_ulp": (f1p, f2p)}.items():
        poses, es, mask, ni, it = oracle.frame_solve_batch(f1, f2, b.covs_target, b.init_poses,
                                                           oracle.default_frame_opts(), n_per_problem=N,
                                                           return_ransac=True)
        models = np.array([oracle.ransac_compute_model(f1[k * N:(k + 1) * N], f2[k * N:(k + 1) * N], b.init_poses[k],
                                                       pair_index=k)[0] for k in range(B)])
        out[f"poses{tag}"], out[f"es_poses{tag}"], out[f"mask{tag}"] = poses, es, mask
        out[f"num_inliers{tag}"], out[f"iterations{tag}"], out[f"models{tag}"] = ni, it, models
    np.savez_compressed(os.path.join(HERE, "oracle_ransac.npz"), **out)


if __name__ == "__main__":
    reference_fixtures()
    reference_scf_fixtures()
    oracle_fixtures()
    oracle_frame_fixtures()
    oracle_ransac_fixtures()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
