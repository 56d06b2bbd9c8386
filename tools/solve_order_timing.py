"""The refinement on BASELINE shapes with the pairs started in index order (PNEC_B200_SOLVE_ORDER=0) and
worst-conditioned first (default): device time per step, CUDA events, several seeds of the C2 shape."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
def handle(order):
    os.environ["PNEC_B200_SOLVE_ORDER"] = str(order)
    try:
        return api.Handle(0)
    finally:
        os.environ.pop("PNEC_B200_SOLVE_ORDER")
def timed(fn, reps=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
h0, h1 = handle(0), handle(1)
o = api.default_opts(api.TARGET)
for B, N, seed in ((10000, 512, 1), (10000, 512, 2), (10000, 512, 3), (10000, 512, 7), (12500, 256, 1), (12500, 256, 2), (2500, 512, 1), (5000, 200, 1), (4000, 1000, 1), (40000, 512, 1)):
    b = syn.make_batch(B, N, seed=seed)
    a = (dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), None, dev(b.init_poses), o)
    r = h1.solve_batch(*a, n_per_problem=N)
    it = r.iterations.cpu().numpy()
    t0 = timed(lambda: h0.solve_batch(*a, n_per_problem=N))
    t1 = timed(lambda: h1.solve_batch(*a, n_per_problem=N))
    print(json.dumps({"B": B, "N": N, "seed": seed, "index order ms": round(t0, 4), "conditioning order ms": round(t1, 4),
                      "pairs at 50 iterations": int((it == 50).sum()), "pairs >= 20": int((it >= 20).sum()),
                      "mean iterations": round(float(it.mean()), 3)}), flush=True)
