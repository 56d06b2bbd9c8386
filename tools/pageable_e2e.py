"""e2e of pnec_solve_batch with PAGEABLE host buffers (what a std::vector caller hands over) vs pinned."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
B, N = 10000, 512
b = syn.make_batch(B, N, seed=1)
h = api.Handle(0)
o = api.default_opts(api.TARGET)
def run(tag, f1, f2, ct, init, reps=10):
    for _ in range(2): h.solve_batch(f1, f2, ct, None, init, o, n_per_problem=N)
    t0 = time.perf_counter()
    for _ in range(reps): h.solve_batch(f1, f2, ct, None, init, o, n_per_problem=N)
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"buffers": tag, "ms_per_step": round(dt * 1e3, 2), "solves_per_s": round(B / dt), "h2d_GBps": round(B * N * 120 / dt / 1e9, 1)}), flush=True)
run("pageable", b.bvs_host, b.bvs_target, b.covs_target, b.init_poses)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
run("pinned", pin(b.bvs_host), pin(b.bvs_target), pin(b.covs_target), pin(b.init_poses))
