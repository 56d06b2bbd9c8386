"""Debug: per-phase cycle counters of solve_kernel (needs the -DPNEC_PHASE_TIMING build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
B, N = int(os.environ.get("PROF_B", 10000)), int(os.environ.get("PROF_N", 512))
base = syn.make_batch(500, N, seed=11)
rep = max(1, B // 500)
dev = torch.device("cuda", 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(np.tile(a, (rep, 1))[: (B * N if a.shape[0] != 500 else B)])).to(dev)
f1, f2, ct, init = T(base.bvs_host), T(base.bvs_target), T(base.covs_target), T(base.init_poses)
h = api.Handle(0)
opts = api.default_opts(api.TARGET)
os.environ["PNEC_B200_DUMP_TIMING"] = "1"; h = api.Handle(0)  # switches are read at handle creation
for i in range(3):
    h.solve_batch(f1, f2, ct, None, init, opts, n_per_problem=N)
    torch.cuda.synchronize()
