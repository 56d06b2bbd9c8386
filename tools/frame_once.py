import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
B, N = 10000, 512
h = api.Handle(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
base = syn.make_batch(1000, N, seed=2)
f1, f2, ct = (T(np.tile(a, (10, 1))) for a in (base.bvs_host, base.bvs_target, base.covs_target))
init = T(np.tile(base.init_poses, (10, 1)))
for _ in range(2):
    r = h.frame_solve_batch(f1, f2, ct, init, api.default_frame_opts(use_ransac=0), n_per_problem=N)
torch.cuda.synchronize()
