"""Where the RANSAC stage spends its time: pass 1 rounds vs pass 2 work items."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
B, N = 10000, 512
batch = syn.make_batch(B, N, seed=1)
rng = np.random.default_rng(0)
f2o = batch.bvs_target.copy()
bad = rng.random(B * N) < 0.25
v = rng.standard_normal((int(bad.sum()), 3)); f2o[bad] = v / np.linalg.norm(v, axis=1, keepdims=True)
f1, f2, f2od, init = dev(batch.bvs_host), dev(batch.bvs_target), dev(f2o), dev(batch.init_poses)
for name, tgt in (("clean", f2), ("25% outliers", f2od)):
    for nw, occ in (("1", "16"), ("4", "16")):
        for defer, max_it in (("0", 7), ("0", 55), ("-1", 247), ("-1", 5000)):
            os.environ["PNEC_B200_RANSAC_WARPS"] = nw; os.environ["PNEC_B200_RANSAC_DEFER"] = defer
            h = api.Handle(0)
            o = api.default_frame_opts(max_ransac_iterations=max_it)
            ms = timed(lambda: h.ransac_batch(f1, tgt, init, o, n_per_problem=N))
            m, ni, it, idx = h.ransac_batch(f1, tgt, init, o, n_per_problem=N)
            print(f"{name} nw={nw} occ={occ} defer={defer} max_it={max_it}: {ms:.2f} ms, total iterations {int(it.sum())}, at cap {int((it == max_it + 1).sum())}", flush=True)
