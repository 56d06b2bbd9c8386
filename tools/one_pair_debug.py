import sys
sys.path.insert(0,'/root/repo')
import numpy as np
from pnec_b200 import api, synthetic as syn
h=api.Handle(0)
N=512
b=syn.make_batch(1,N,seed=3)
for _ in range(2):
    h.frame_solve_batch(b.bvs_host,b.bvs_target,b.covs_target,b.init_poses,api.default_frame_opts(use_ransac=0),n_per_problem=N)
