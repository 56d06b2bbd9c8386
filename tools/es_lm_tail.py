"""What bounds the rotation LM (es_lm_kernel) on the C2 shape: the few pairs that run into maxfev, or
the bulk?  Times pnec_eigensolver_batch on (a) the bench batch, (b) the same batch with the maxfev
pairs replaced by ordinary ones, (c) the maxfev pairs alone, (d) one ordinary pair alone.  CUDA events
on torch's current stream; the moments kernel (0.055 ms at 10 000 x 512) is part of every figure."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn

B = int(os.environ.get("FT_B", 10000)); N = int(os.environ.get("FT_N", 512))
dev = torch.device("cuda", 0)
h = api.Handle(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
b = syn.make_batch(B, N, seed=int(os.environ.get("FT_SEED", 1)))
f1, f2, init = T(b.bvs_host), T(b.bvs_target), T(b.init_poses)

def timeit(fn, reps=20, warm=3):
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
    return T(b.bvs_host[rows]), T(b.bvs_target[rows]), T(b.init_poses[idx])

_, info, _ = h.eigensolver_batch(f1, f2, init, n_per_problem=N)
info = info.cpu().numpy()
hard = np.nonzero(info == 5)[0]
easy = np.nonzero(info != 5)[0]
out = {"B": B, "N": N, "info_histogram": np.bincount(info, minlength=9).tolist()}
out["all pairs ms"] = timeit(lambda: h.eigensolver_batch(f1, f2, init, n_per_problem=N))
idx = np.arange(B); idx[hard] = easy[: len(hard)]
g = sub(idx)
out["maxfev pairs replaced ms"] = timeit(lambda: h.eigensolver_batch(*g, n_per_problem=N))
if len(hard):
    g2 = sub(hard)
    out["maxfev pairs alone ms (%d pairs)" % len(hard)] = timeit(lambda: h.eigensolver_batch(*g2, n_per_problem=N))
    g3 = sub(hard[:1])
    out["one maxfev pair ms"] = timeit(lambda: h.eigensolver_batch(*g3, n_per_problem=N))
g4 = sub(easy[:1])
out["one ordinary pair ms"] = timeit(lambda: h.eigensolver_batch(*g4, n_per_problem=N))
g5 = sub(easy[:1184])
out["1184 ordinary pairs ms"] = timeit(lambda: h.eigensolver_batch(*g5, n_per_problem=N))
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}))
