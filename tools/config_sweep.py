"""Times K1 (eval) and the full solve on the other BASELINE.json configs (C1, C3 shard, C4, C5 sweep).
Parity-test cases, not bench lines; the table goes to profiles/."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn

dev = torch.device("cuda", 0)
h = api.Handle(0)
T = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)

def tile_batch(base, B):
    """B problems by repeating a smaller generated batch (kernels do not exploit repeats)."""
    rep = (B + base.num_problems - 1) // base.num_problems
    n = base.n_per_problem
    f = lambda a, per: np.tile(a, (rep, 1))[: B * per]
    return f(base.bvs_host, n), f(base.bvs_target, n), f(base.covs_target, n), f(base.init_poses, 1)

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

rows = []
def run(name, f1, f2, ct, init, opts, **kw):
    total = f1.shape[0]; B = init.shape[0]
    d = [T(x) for x in (f1, f2, ct, init)]
    ev_ms = timeit(lambda: h.eval_batch(d[0], d[1], d[2], None, d[3], api.TARGET, 1e-13, **kw))
    so_ms = timeit(lambda: h.solve_batch(d[0], d[1], d[2], None, d[3], opts, **kw))
    res = h.solve_batch(d[0], d[1], d[2], None, d[3], opts, **kw); torch.cuda.synchronize()
    it = res.iterations.float().mean().item()
    row = dict(config=name, problems=B, correspondences=total, MB=total * 120 / 1e6, k1_ms=ev_ms,
               k1_GBps=total * 120 / ev_ms / 1e6, solve_ms=so_ms, solves_per_s=B / so_ms * 1e3,
               solve_single_pass_GBps=total * 120 / so_ms / 1e6, mean_iterations=it)
    rows.append(row)
    print(json.dumps(row), flush=True)

opts = api.default_opts(api.TARGET)
# C1: one frame pair, 100 correspondences
b = syn.make_batch(1, 100, seed=1, noise_type="isotropic_homogenous")
run("C1 1x100", b.bvs_host, b.bvs_target, b.covs_target, b.init_poses, opts, n_per_problem=100)
# C3: one 1/8 shard of 100k x 256
base = syn.make_batch(500, 256, seed=3)
run("C3 shard 12500x256", *tile_batch(base, 12500), opts, n_per_problem=256)
# C4: KITTI-shaped ragged stream, 20 LM iterations
counts = syn.kitti_like_counts()
b = syn.make_batch(len(counts), 0, seed=4, camera=syn.PINHOLE, counts=counts)
run("C4 4540x~2000 ragged, max 20 it", b.bvs_host, b.bvs_target, b.covs_target, b.init_poses,
    api.default_opts(api.TARGET, max_num_iterations=20), offsets=b.offsets)
# C5: correspondence-count sweep at 4096 problems
os.environ["PNEC_B200_STREAM_MIN_N"] = "100000000"; h = api.Handle(0)  # switches are read at handle creation
run("C4 resident/global path", b.bvs_host, b.bvs_target, b.covs_target, b.init_poses,
    api.default_opts(api.TARGET, max_num_iterations=20), offsets=b.offsets)
os.environ.pop("PNEC_B200_STREAM_MIN_N"); h = api.Handle(0)  # switches are read at handle creation
for N in (64, 128, 256, 512, 768, 1024, 2048, 4096, 8192):
    base = syn.make_batch(128 if N <= 1024 else 32, N, seed=5)
    run(f"C5 4096x{N}", *tile_batch(base, 4096), opts, n_per_problem=N)
    if 512 <= N <= 2048:
        os.environ["PNEC_B200_STREAM_MIN_N"] = "100000000" if N > 940 else "0"; h = api.Handle(0)  # switches are read at handle creation
        run(f"C5 4096x{N} other path ({'resident' if N > 940 else 'stream'})", *tile_batch(base, 4096), opts, n_per_problem=N)
        os.environ.pop("PNEC_B200_STREAM_MIN_N"); h = api.Handle(0)  # switches are read at handle creation
# unscented transform: one covariance per C2 correspondence
n = 10000 * 512
rng = np.random.default_rng(0)
mus = syn._uniform_sphere(rng, (n,)) * 800.0
c3 = np.zeros((n, 9)); c3[:, 0] = 0.6; c3[:, 4] = 0.4; c3[:, 1] = c3[:, 3] = 0.1
dm, dc = T(mus), T(c3)
for cam, nm in ((api.CAMERA_PINHOLE, "pinhole"), (api.CAMERA_OMNIDIRECTIONAL, "omnidirectional")):
    ms = timeit(lambda: h.unscented_transform(dm, dc, None, 1.0, cam))
    row = dict(config=f"unscented transform {nm}, {n} points", points=n, ms=ms, points_per_s=n / ms * 1e3,
               GBps=n * 168 / ms / 1e6, algorithmic_bytes_per_point=168)
    rows.append(row); print(json.dumps(row), flush=True)
# keypoints -> solver inputs (KeyPoint::Unproject): 48 B in, 96 B out per keypoint
pts = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n)], -1)
c2 = np.tile(np.array([0.6, 0.1, 0.1, 0.4]), (n, 1))
Kinv = np.linalg.inv(np.array([[718.856, 0, 607.19], [0, 718.856, 185.2157], [0, 0, 1.0]])).T.reshape(9)
dp, dc2 = T(pts), T(c2)
ms = timeit(lambda: h.keypoints_unproject(dp, dc2, Kinv))
row = dict(config=f"keypoints unproject, {n} keypoints", points=n, ms=ms, points_per_s=n / ms * 1e3,
           GBps=n * 144 / ms / 1e6, algorithmic_bytes_per_point=144)
rows.append(row); print(json.dumps(row), flush=True)
# translation given rotation on the C2 shape
base = syn.make_batch(500, 512, seed=6)
f1_, f2_, ct_, init_ = (T(x) for x in tile_batch(base, 10000))
ms = timeit(lambda: h.scf_translation_batch(f1_, f2_, ct_, init_, 1e-13, 500, 10, n_per_problem=512), reps=5, warm=2)
row = dict(config="SCF translation (500-point scan + 10 SCF steps), 10000x512", points=10000, ms=ms,
           points_per_s=10000 / ms * 1e3, GBps=10000 * 512 * 120 / ms / 1e6, algorithmic_bytes_per_point=512 * 120,
           rayleigh_quotients_per_s=10000 * 512 * 501 / ms * 1e3)
rows.append(row); print(json.dumps(row), flush=True)
ms = timeit(lambda: h.nec_translation_batch(f1_, f2_, init_, n_per_problem=512))
row = dict(config="NEC translation (ComposeM + TranslationFromM), 10000x512", points=10000, ms=ms,
           points_per_s=10000 / ms * 1e3, GBps=10000 * 512 * 48 / ms / 1e6, algorithmic_bytes_per_point=512 * 48)
rows.append(row); print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(os.environ.get("CONFIGS_OUT", "gpurun_out/configs_r02.json"), "w"), indent=1)
