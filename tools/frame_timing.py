"""Stage timings of the whole frame solve (PNEC::Solve, pnec_frame_solve_batch) on the C2 shape:
eigensolver (moments + LM), NEC translation, weighted eigensolver, SCF, refinement.  CUDA events on
torch's current stream (the stream the C-ABI launches on).  Output: one JSON line per row."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn

B = int(os.environ.get("FT_B", 10000)); N = int(os.environ.get("FT_N", 512))
dev = torch.device("cuda", 0)
h = api.Handle(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
base = syn.make_batch(min(B, 1000), N, seed=2)
rep = (B + base.num_problems - 1) // base.num_problems
f1, f2, ct = (T(np.tile(a, (rep, 1))[: B * N]) for a in (base.bvs_host, base.bvs_target, base.covs_target))
init = T(np.tile(base.init_poses, (rep, 1))[:B])

def timeit(fn, reps=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

es, info, ev = h.eigensolver_batch(f1, f2, init, n_per_problem=N)
rows = {}
rows["eigensolver (moments + LM)"] = timeit(lambda: h.eigensolver_batch(f1, f2, init, n_per_problem=N))
rows["nec translation"] = timeit(lambda: h.nec_translation_batch(f1, f2, es, n_per_problem=N))
rows["weighted eigensolver (moments + LM)"] = timeit(
    lambda: h.eigensolver_batch(f1, f2, es, covs_target=ct, weight_poses=es, n_per_problem=N))
rows["scf translation (500 samples, 10 steps)"] = timeit(
    lambda: h.scf_translation_batch(f1, f2, ct, es, n_per_problem=N), reps=5)
rows["refinement (solve_batch)"] = timeit(lambda: h.solve_batch(f1, f2, ct, None, es, n_per_problem=N))
for name, kw in [("frame solve, default (10 weighted iterations)", {}),
                 ("frame solve, weighted_iterations=1", dict(weighted_iterations=1)),
                 ("frame solve, NEC", dict(use_nec=1))]:
    o = api.default_frame_opts(use_ransac=0, **kw)
    rows[name] = timeit(lambda: h.frame_solve_batch(f1, f2, ct, init, o, n_per_problem=N), reps=3, warm=1)
hist = np.bincount(info.cpu().numpy(), minlength=9).tolist()
for k, v in rows.items():
    print(json.dumps({"stage": k, "ms": round(v, 4), "pairs_per_s": round(B / v * 1e3, 1), "B": B, "N": N}), flush=True)
print(json.dumps({"eigensolver_lm_info_histogram": hist}))
