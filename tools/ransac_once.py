"""One RANSAC stage call on the C2 batch with 25 % outliers (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
B, N = int(os.environ.get("B", 10000)), 512
frac = float(os.environ.get("OUTLIERS", 0.25))
batch = syn.make_batch(B, N, seed=1)
rng = np.random.default_rng(0)
f2o = batch.bvs_target.copy()
bad = rng.random(B * N) < frac
v = rng.standard_normal((int(bad.sum()), 3)); f2o[bad] = v / np.linalg.norm(v, axis=1, keepdims=True)
f1, f2od, init = dev(batch.bvs_host), dev(f2o), dev(batch.init_poses)
h = api.Handle(0)
for _ in range(int(os.environ.get("REPS", 1))):
    m, ni, it, idx = h.ransac_batch(f1, f2od, init, api.default_frame_opts(), n_per_problem=N)
torch.cuda.synchronize()
print("iterations", int(it.sum()), "max", int(it.max()))
