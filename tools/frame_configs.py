"""The frame solve on the other BASELINE.json shapes: C1 (1 x 100), a C3 shard (12 500 x 256),
C4 (KITTI-shaped ragged, 4540 x ~2000).  CUDA events; one JSON line per shape."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
h = api.Handle(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def run(name, b, **kw):
    d = [T(x) for x in (b.bvs_host, b.bvs_target, b.covs_target, b.init_poses)]
    rows = {}
    for tag, fo in (("default", {}), ("weighted_iterations=1", dict(weighted_iterations=1))):
        o = api.default_frame_opts(use_ransac=0, **fo)
        rows[tag] = timeit(lambda: h.frame_solve_batch(*d, o, **kw))
    rows["refinement only"] = timeit(lambda: h.solve_batch(d[0], d[1], d[2], None, d[3], api.default_opts(api.TARGET), **kw))
    B = b.num_problems
    print(json.dumps(dict(shape=name, pairs=B, correspondences=int(b.total),
                          **{f"{k} ms": round(v, 3) for k, v in rows.items()},
                          **{f"{k} pairs/s": round(B / v * 1e3) for k, v in rows.items()})), flush=True)
which = os.environ.get("FC", "c1,c3,c4").split(",")
if "c1" in which:
    run("C1 1x100 iso", syn.make_batch(1, 100, seed=1, noise_type="isotropic_homogenous"), n_per_problem=100)
if "c3" in which:
    base = syn.make_batch(12500, 256, seed=3)
    run("C3 shard 12500x256", base, n_per_problem=256)
if "c4" in which:
    counts = syn.kitti_like_counts()
    b = syn.make_batch(len(counts), 0, seed=4, camera=syn.PINHOLE, counts=counts)
    run("C4 4540 x ~2000 ragged pinhole", b, offsets=b.offsets)
