"""Is the frame solve (no RANSAC, default options) on the C2 batch bound by the few pairs whose rotation
LM runs into maxfev?  Times pnec_frame_solve_batch on the bench batch and on the same batch with
those pairs replaced by ordinary ones."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn

B = int(os.environ.get("FT_B", 10000)); N = int(os.environ.get("FT_N", 512))
h = api.Handle(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
b = syn.make_batch(B, N, seed=int(os.environ.get("FT_SEED", 1)))

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / reps

def sub(idx):
    idx = np.asarray(idx)
    rows = (idx[:, None] * N + np.arange(N)[None, :]).reshape(-1)
    return T(b.bvs_host[rows]), T(b.bvs_target[rows]), T(b.covs_target[rows]), T(b.init_poses[idx])

full = sub(np.arange(B))
_, info, _ = h.eigensolver_batch(full[0], full[1], full[3], n_per_problem=N)
info = info.cpu().numpy()
hard = np.nonzero(info == 5)[0]; easy = np.nonzero(info != 5)[0]
idx = np.arange(B); idx[hard] = easy[: len(hard)]
clean = sub(idx)
out = {"B": B, "N": N, "maxfev pairs": int(len(hard))}
for name, kw in (("default", {}), ("weighted_iterations=1", dict(weighted_iterations=1)), ("no refinement", dict(use_ceres=0))):
    o = api.default_frame_opts(use_ransac=0, **kw)
    out[name + " ms"] = round(timeit(lambda: h.frame_solve_batch(*full, o, n_per_problem=N)), 4)
    out[name + ", maxfev pairs replaced ms"] = round(timeit(lambda: h.frame_solve_batch(*clean, o, n_per_problem=N)), 4)
print(json.dumps(out))
