import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
N = 512
h = api.Handle(0)
for B in (1, 8, 64, 256, 500):
    b = syn.make_batch(B, N, seed=5)
    f1, f2, init = dev(b.bvs_host), dev(b.bvs_target), dev(b.init_poses)
    print(json.dumps({"B": B, "ransac stage ms": round(timed(lambda: h.ransac_batch(f1, f2, init, n_per_problem=N)), 4)}), flush=True)
