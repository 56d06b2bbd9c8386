"""Small workloads through the round-2 kernels, for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
os.environ["PNEC_B200_SOLVE_SLOTS"] = "2"; os.environ["PNEC_B200_RANSAC_WARPS"] = "1"
h = api.Handle(0)
# slots kernel: odd N (heads / by-hand tail), ragged with empty pairs, symmetric variant
b = syn.make_batch(37, 333, seed=5)
h.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), None, dev(b.init_poses), api.default_opts(api.TARGET), n_per_problem=333)
counts = np.array([0, 5, 512, 1, 300, 0, 77, 560, 2, 129]); b = syn.make_batch(len(counts), 0, seed=6, counts=counts)
h.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), None, dev(b.init_poses), api.default_opts(api.TARGET), offsets=b.offsets)
b = syn.with_host_covariances(syn.make_batch(20, 200, seed=7))
for v in (api.NEC, api.HOST, api.SYMMETRIC):
    h.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(None if v == api.NEC else b.covs_target), dev(b.covs_host if v == api.SYMMETRIC else None),
                  dev(b.init_poses), api.default_opts(v), n_per_problem=200)
# RANSAC: split pass 1 + pass 2 + the frame solve behind it
b = syn.make_batch(24, 160, seed=8, noise_level=0.5)
rng = np.random.default_rng(1); bad = rng.random(24 * 160) < 0.3
v = rng.standard_normal((int(bad.sum()), 3)); b.bvs_target[bad] = v / np.linalg.norm(v, axis=1, keepdims=True)
h.frame_solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(b.covs_target), dev(b.init_poses), api.default_frame_opts(), n_per_problem=160)
torch.cuda.synchronize()
print("ok", h.launch_count)
