"""Latency of one frame pair per call (the reference's calling pattern: PNEC::Solve per frame),
HOST buffers through the C-ABI, vs the oracle on one core."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from pnec_b200 import api, synthetic as syn
h = api.Handle(0)
for N in (100, 512, 2000):
    b = syn.make_batch(8, N, seed=3)
    args = lambda k: (b.bvs_host[k*N:(k+1)*N], b.bvs_target[k*N:(k+1)*N], b.covs_target[k*N:(k+1)*N], b.init_poses[k:k+1])
    fo, so = api.default_frame_opts(use_ransac=0), api.default_opts(api.TARGET)
    for _ in range(3):
        h.frame_solve_batch(*args(0), fo, n_per_problem=N); h.solve_batch(*args(0)[:3], None, args(0)[3], so, n_per_problem=N)
    reps = 50
    t0 = time.perf_counter()
    for r in range(reps): h.frame_solve_batch(*args(r % 8), fo, n_per_problem=N)
    t_frame = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for r in range(reps): h.solve_batch(*args(r % 8)[:3], None, args(r % 8)[3], so, n_per_problem=N)
    t_solve = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for r in range(8): oracle.frame_solve_batch(*args(r), oracle.default_frame_opts(use_ransac=0), n_per_problem=N, num_threads=1)
    c_frame = (time.perf_counter() - t0) / 8
    t0 = time.perf_counter()
    for r in range(8): oracle.solve_batch(*args(r)[:3], None, args(r)[3], oracle.default_opts(oracle.TARGET), n_per_problem=N, num_threads=1)
    c_solve = (time.perf_counter() - t0) / 8
    print(json.dumps(dict(N=N, gpu_frame_solve_ms=round(t_frame*1e3, 3), gpu_refinement_ms=round(t_solve*1e3, 3),
                          cpu_oracle_frame_solve_ms=round(c_frame*1e3, 3), cpu_oracle_refinement_ms=round(c_solve*1e3, 3))), flush=True)
