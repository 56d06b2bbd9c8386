"""Workload for ncu: the stages of the frame solve on a C2-shaped batch.
  ncu --set full --clock-control none --import-source on -k regex:scf_kernel -s 1 -c 1 -o gpurun_out/prof_scf python tools/profile_frame.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn

B = int(os.environ.get("PROF_B", 10000)); N = int(os.environ.get("PROF_N", 512))
mode = os.environ.get("PROF_MODE", "stages")
h = api.Handle(0)
base = syn.make_batch(min(B, 500), N, seed=11)
rep = (B + base.num_problems - 1) // base.num_problems
T = lambda a, per: torch.from_numpy(np.ascontiguousarray(np.tile(a, (rep, 1))[: B * per])).cuda()
f1, f2, ct, init = T(base.bvs_host, N), T(base.bvs_target, N), T(base.covs_target, N), T(base.init_poses, 1)
for _ in range(3):
    if mode == "stages":
        es, _, _ = h.eigensolver_batch(f1, f2, init, n_per_problem=N)
        h.eigensolver_batch(f1, f2, es, covs_target=ct, weight_poses=es, n_per_problem=N)
        h.scf_translation_batch(f1, f2, ct, es, n_per_problem=N)
    else:
        h.frame_solve_batch(f1, f2, ct, init, api.default_frame_opts(use_ransac=0), n_per_problem=N)
torch.cuda.synchronize()
print("done", h.launch_count)
