import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from pnec_b200 import api, synthetic as syn
N = int(os.environ.get("N", 512))
b = syn.make_batch(8, N, seed=3)
args = lambda k: (b.bvs_host[k*N:(k+1)*N], b.bvs_target[k*N:(k+1)*N], b.covs_target[k*N:(k+1)*N], b.init_poses[k:k+1])
h = api.Handle(0)
def t(fn, reps=40):
    for k in range(4): fn(k)
    t0 = time.perf_counter()
    for k in range(reps): fn(k % 8)
    return (time.perf_counter() - t0) / reps * 1e3
print("ransac_batch alone           %.3f ms" % t(lambda k: h.ransac_batch(args(k)[0], args(k)[1], args(k)[3], api.default_frame_opts(), n_per_problem=N)))
print("frame: ransac + ES only      %.3f ms" % t(lambda k: h.frame_solve_batch(*args(k), api.default_frame_opts(weighted_iterations=1, use_ceres=0), n_per_problem=N)))
print("frame: ES only (no ransac)   %.3f ms" % t(lambda k: h.frame_solve_batch(*args(k), api.default_frame_opts(use_ransac=0, weighted_iterations=1, use_ceres=0), n_per_problem=N)))
print("frame: no ransac, full       %.3f ms" % t(lambda k: h.frame_solve_batch(*args(k), api.default_frame_opts(use_ransac=0), n_per_problem=N)))
print("frame: default (ransac) full %.3f ms" % t(lambda k: h.frame_solve_batch(*args(k), api.default_frame_opts(), n_per_problem=N)))
print("refinement only              %.3f ms" % t(lambda k: h.solve_batch(*args(k)[:3], None, args(k)[3], api.default_opts(api.TARGET), n_per_problem=N)))
for w in ("1", "2"):
    os.environ["PNEC_B200_RANSAC_WARPS"] = w
    h2 = api.Handle(0)
    print("ransac_batch alone, warps=%s   %.3f ms" % (w, t(lambda k: h2.ransac_batch(args(k)[0], args(k)[1], args(k)[3], api.default_frame_opts(), n_per_problem=N))))
os.environ["PNEC_B200_RANSAC_DEFER"] = "0"
h3 = api.Handle(0)
print("ransac_batch alone, warps=2 no pass 2   %.3f ms" % t(lambda k: h3.ransac_batch(args(k)[0], args(k)[1], args(k)[3], api.default_frame_opts(), n_per_problem=N)))
# pinned host buffers (what packing small inputs into one pinned staging block would give at best)
import torch
os.environ.pop("PNEC_B200_RANSAC_WARPS", None); os.environ.pop("PNEC_B200_RANSAC_DEFER", None)
hp = api.Handle(0)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
P = [tuple(pin(x) for x in args(k)) for k in range(8)]
print("pinned inputs: frame default   %.3f ms" % t(lambda k: hp.frame_solve_batch(*P[k], api.default_frame_opts(), n_per_problem=N)))
print("pinned inputs: frame no ransac %.3f ms" % t(lambda k: hp.frame_solve_batch(*P[k], api.default_frame_opts(use_ransac=0), n_per_problem=N)))
