"""Solve-kernel warps-per-problem sweep for small/medium N (resident kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
dev = torch.device("cuda", 0); h = api.Handle(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
opts = api.default_opts(api.TARGET)
for N, B in ((64, 16384), (128, 16384), (256, 12500), (384, 10000), (512, 10000), (640, 8192), (768, 8192), (1024, 4096), (1536, 4096), (2048, 4096)):
    base = syn.make_batch(256, N, seed=5)
    rep = (B + 255) // 256
    f = lambda a, per: T(np.tile(a, (rep, 1))[: B * per])
    d = (f(base.bvs_host, N), f(base.bvs_target, N), f(base.covs_target, N), f(base.init_poses, 1))
    out = []
    for nw in (1, 2, 3, 4, 8):
        os.environ["PNEC_B200_SOLVE_WARPS"] = str(nw); h = api.Handle(0)  # switches are read at handle creation
        os.environ["PNEC_B200_STREAM_MIN_N"] = "100000000"; h = api.Handle(0)  # switches are read at handle creation
        try:
            ms = timeit(lambda: h.solve_batch(d[0], d[1], d[2], None, d[3], opts, n_per_problem=N))
            out.append(f"nw{nw}: {ms:.4f} ms ({B/ms/1e3:.1f}M/s)")
        except Exception as e:
            out.append(f"nw{nw}: fail")
    os.environ.pop("PNEC_B200_SOLVE_WARPS"); os.environ["PNEC_B200_STREAM_MIN_N"] = "0"; h = api.Handle(0)  # switches are read at handle creation
    ms = timeit(lambda: h.solve_batch(d[0], d[1], d[2], None, d[3], opts, n_per_problem=N))
    out.append(f"stream: {ms:.4f}")
    print(f"N={N:5d} B={B:6d}  " + "  ".join(out), flush=True)
