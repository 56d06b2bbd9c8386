"""Debug: cycle probes of solve_slots_kernel (build with PNEC_B200_NVCC_EXTRA=-DPNEC_SLOT_TIMING)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
B, N = int(os.environ.get("PROF_B", 10000)), int(os.environ.get("PROF_N", 512))
b = syn.make_batch(B, N, seed=2024)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
f1, f2, ct, init = dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), dev(b.init_poses)
os.environ["PNEC_B200_DUMP_TIMING"] = "1"; os.environ["PNEC_B200_SOLVE_SLOTS"] = "2"
h = api.Handle(0)  # switches are read at handle creation
opts = api.default_opts(api.TARGET)
for i in range(3):
    h.solve_batch(f1, f2, ct, None, init, opts, n_per_problem=N)
    torch.cuda.synchronize()
