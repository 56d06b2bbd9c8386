"""RANSAC stage and default-options frame solve: device timings on the C2 batch and single-pair latency."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn

dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = []
B, N = 10000, 512
batch = syn.make_batch(B, N, seed=1)
f1, f2, ct, init = dev(batch.bvs_host), dev(batch.bvs_target), dev(batch.covs_target), dev(batch.init_poses)
for nw in ("1", "2", "4"):
    os.environ["PNEC_B200_RANSAC_WARPS"] = nw
    h = api.Handle(0)
    ms = timed(lambda: h.ransac_batch(f1, f2, init, n_per_problem=N))
    m, ni, it, idx = h.ransac_batch(f1, f2, init, n_per_problem=N)
    out.append({"what": "ransac stage C2", "warps": nw, "ms": ms, "mean_iters": float(it.float().mean()), "max_iters": int(it.max()),
                "mean_inlier_frac": float(ni.float().mean() / N)})
    print(out[-1], flush=True)
os.environ.pop("PNEC_B200_RANSAC_WARPS")
h = api.Handle(0)
for kw, name in ((dict(), "frame solve default (RANSAC)"), (dict(use_ransac=0), "frame solve no RANSAC")):
    o = api.default_frame_opts(**kw)
    ms = timed(lambda: h.frame_solve_batch(f1, f2, ct, init, o, n_per_problem=N))
    out.append({"what": name, "ms": ms}); print(out[-1], flush=True)
# outliers: 25 % on the C2 batch
rng = np.random.default_rng(0)
f2o = batch.bvs_target.copy()
bad = rng.random(B * N) < 0.25
v = rng.standard_normal((int(bad.sum()), 3)); f2o[bad] = v / np.linalg.norm(v, axis=1, keepdims=True)
f2od = dev(f2o)
for nw in ("1", "4"):
    os.environ["PNEC_B200_RANSAC_WARPS"] = nw
    h = api.Handle(0)
    ms = timed(lambda: h.ransac_batch(f1, f2od, init, n_per_problem=N), reps=3, warm=1)
    m, ni, it, idx = h.ransac_batch(f1, f2od, init, n_per_problem=N)
    out.append({"what": "ransac stage C2 + 25% outliers", "warps": nw, "ms": ms, "mean_iters": float(it.float().mean()), "max_iters": int(it.max())})
    print(out[-1], flush=True)
os.environ.pop("PNEC_B200_RANSAC_WARPS")
# single pair latency (HOST call), the VO calling pattern: N = 2000, clean and 40 % outliers
for frac in (0.0, 0.4):
    b1 = syn.make_batch(1, 2000, seed=3, camera=syn.PINHOLE)
    if frac:
        k = int(frac * 2000); sel = rng.choice(2000, k, replace=False)
        v = rng.standard_normal((k, 3)); b1.bvs_target[sel] = v / np.linalg.norm(v, axis=1, keepdims=True)
    for nw in ("1", "4"):
        os.environ["PNEC_B200_RANSAC_WARPS"] = nw
        h = api.Handle(0)
        o = api.default_frame_opts()
        fn = lambda: h.frame_solve_batch(b1.bvs_host, b1.bvs_target, b1.covs_target, b1.init_poses, o, n_per_problem=2000)
        for _ in range(3): fn()
        t0 = time.perf_counter()
        for _ in range(10): r = fn()
        ms = (time.perf_counter() - t0) / 10 * 1e3
        out.append({"what": f"single pair N=2000 outliers {frac}", "warps": nw, "ms_host_call": ms, "ransac_iters": int(r.ransac_iterations[0]), "inliers": int(r.num_inliers[0])})
        print(out[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/ransac_timing.jsonl", "w") as f:
    for o in out: f.write(json.dumps(o) + "\n")
