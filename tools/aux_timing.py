import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
B, N = 10000, 512
b = syn.make_batch(B, N, seed=1)
f1, f2, ct, init = T(b.bvs_host), T(b.bvs_target), T(b.covs_target), T(b.init_poses)
h = api.Handle(0)
def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / reps
t = timeit(lambda: h.nec_translation_batch(f1, f2, init, n_per_problem=N))
print(json.dumps({"kernel": "nec_translation_kernel", "ms": round(t, 4), "GB/s": round(B * N * 48 / t / 1e6, 1)}))
t = timeit(lambda: h.cost_function_batch(f1, f2, ct, init, n_per_problem=N))
print(json.dumps({"kernel": "cost_kernel", "ms": round(t, 4), "GB/s": round(B * N * 120 / t / 1e6, 1)}))
