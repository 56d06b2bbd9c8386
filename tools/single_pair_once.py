"""One frame pair per call through the C-ABI with HOST buffers (for ncu launch lists): default options and no RANSAC."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pnec_b200 import api, synthetic as syn
N = int(os.environ.get("N", 512))
b = syn.make_batch(4, N, seed=3)
args = lambda k: (b.bvs_host[k*N:(k+1)*N], b.bvs_target[k*N:(k+1)*N], b.covs_target[k*N:(k+1)*N], b.init_poses[k:k+1])
h = api.Handle(0)
for use_ransac in (0, 1):
    fo = api.default_frame_opts(use_ransac=use_ransac)
    for k in range(3):
        h.frame_solve_batch(*args(k), fo, n_per_problem=N)
    t0 = time.perf_counter()
    for k in range(20):
        h.frame_solve_batch(*args(k % 4), fo, n_per_problem=N)
    print("use_ransac", use_ransac, "ms per call", (time.perf_counter() - t0) / 20 * 1e3, flush=True)
