"""Workload for ncu: C2-shaped batch, a few eval (K1) and solve launches.
  ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 2 -c 1 -o gpurun_out/prof_eval python tools/profile_run.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn

B = int(os.environ.get("PROF_B", 10000)); N = int(os.environ.get("PROF_N", 512))
mode = os.environ.get("PROF_MODE", "both")
dev = torch.device("cuda", 0)
h = api.Handle(0)
# cheap generation: one 500-problem block tiled (the kernels do not care about repeats)
base = syn.make_batch(min(B, 500), N, seed=11)
rep = (B + base.num_problems - 1) // base.num_problems
T = lambda a: torch.from_numpy(np.ascontiguousarray(np.tile(a, (rep, 1))[: (B * N if a.shape[0] != base.num_problems else B)])).to(dev)
f1, f2, ct, init = T(base.bvs_host), T(base.bvs_target), T(base.covs_target), T(base.init_poses)
opts = api.default_opts(api.TARGET)
for _ in range(4):
    if mode in ("both", "eval"):
        h.eval_batch(f1, f2, ct, None, init, api.TARGET, 1e-13, n_per_problem=N)
    if mode in ("both", "solve"):
        h.solve_batch(f1, f2, ct, None, init, opts, n_per_problem=N)
torch.cuda.synchronize()
print("done", h.launch_count)
