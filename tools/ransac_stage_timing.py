import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
B, N = 10000, 512
batch = syn.make_batch(B, N, seed=1, noise_type="anisotropic_inhomogenous", noise_level=1.0)
rng = np.random.default_rng(0)
f2o = batch.bvs_target.copy()
bad = rng.random(B * N) < 0.25
v = rng.standard_normal((int(bad.sum()), 3)); f2o[bad] = v / np.linalg.norm(v, axis=1, keepdims=True)
f1, f2, f2od, ct, init = dev(batch.bvs_host), dev(batch.bvs_target), dev(f2o), dev(batch.covs_target), dev(batch.init_poses)
h = api.Handle(0)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
o = api.default_frame_opts()
print(json.dumps({"ransac stage clean ms": timed(lambda: h.ransac_batch(f1, f2, init, o, n_per_problem=N)),
                  "ransac stage 25% outliers ms": timed(lambda: h.ransac_batch(f1, f2od, init, o, n_per_problem=N), reps=2),
                  "frame default clean ms": timed(lambda: h.frame_solve_batch(f1, f2, ct, init, o, n_per_problem=N))}))
