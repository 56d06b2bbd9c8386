"""RANSAC pass 1 as three kernels per round vs the one-kernel form: outputs compared, stage timed."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
def handle(**env):
    for k, v in env.items(): os.environ[k] = str(v)
    h = api.Handle(0)
    for k in env: os.environ.pop(k)
    return h
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def case(name, B, N, outliers, seed, **okw):
    batch = syn.make_batch(B, N, seed=seed, noise_type="anisotropic_inhomogenous", noise_level=1.0)
    f2 = batch.bvs_target.copy()
    if outliers > 0:
        rng = np.random.default_rng(seed)
        bad = rng.random(B * N) < outliers
        v = rng.standard_normal((int(bad.sum()), 3)); f2[bad] = v / np.linalg.norm(v, axis=1, keepdims=True)
    f1, f2d, init = dev(batch.bvs_host), dev(f2), dev(batch.init_poses)
    one = handle(PNEC_B200_RANSAC_SPLIT=0, PNEC_B200_RANSAC_WARPS=1)
    split = handle(PNEC_B200_RANSAC_SPLIT=1, PNEC_B200_RANSAC_WARPS=1)
    o = api.default_frame_opts(**okw)
    r0 = one.ransac_batch(f1, f2d, init, o, n_per_problem=N)
    r1 = split.ransac_batch(f1, f2d, init, o, n_per_problem=N)
    torch.cuda.synchronize()
    m0, n0, i0, x0 = [t.cpu().numpy() for t in r0]
    m1, n1, i1, x1 = [t.cpu().numpy() for t in r1]
    same_it = (i0 == i1); same_n = (n0 == n1)
    idx_same = np.array([np.array_equal(x0[k*N:k*N+n0[k]], x1[k*N:k*N+n1[k]]) for k in range(B)])
    row = dict(case=name, B=B, pairs_same_iterations=float(same_it.mean()), pairs_same_inlier_count=float(same_n.mean()),
               pairs_same_inlier_set=float(idx_same.mean()), max_model_diff=float(np.abs(m0 - m1).max()),
               models_bit_identical=float((m0 == m1).all(axis=1).mean()),
               ms_one_kernel=timed(lambda: one.ransac_batch(f1, f2d, init, o, n_per_problem=N)),
               ms_split=timed(lambda: split.ransac_batch(f1, f2d, init, o, n_per_problem=N)))
    print(json.dumps(row), flush=True)
case("C2 clean", 10000, 512, 0.0, 1)
case("C2 25% outliers", 10000, 512, 0.25, 1)
case("2000x200 10% outliers", 2000, 200, 0.10, 5)
case("300x64, max 20 iterations", 300, 64, 0.3, 7, max_ransac_iterations=20)
case("50x9 (fewer than the sample)", 50, 9, 0.0, 9)
