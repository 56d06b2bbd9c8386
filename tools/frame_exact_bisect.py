"""Which scheduling device of the frame solve changes bits?  fast (all on) vs each one switched off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn

SW = {"shortcuts": ("PNEC_B200_NO_FRAME_SHORTCUTS", "1"), "defer": ("PNEC_B200_SCF_DEFER", "0"),
      "chunks": ("PNEC_B200_FRAME_CHUNKS", "1"), "lm_ahead": ("PNEC_B200_NO_LM_AHEAD", "1")}
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for B, N, host in [(640, 160, False), (4608, 96, False), (600, 100, True)]:
    batch = syn.make_batch(B, N, seed=300 + N)
    args = (batch.bvs_host, batch.bvs_target, batch.covs_target, batch.init_poses)
    if not host:
        args = tuple(dev(a) for a in args)
    def run(env):
        for k, (name, _) in SW.items():
            os.environ.pop(name, None)
        for k in env:
            os.environ[SW[k][0]] = SW[k][1]
        try:
            r = api.Handle(0).frame_solve_batch(*args, api.default_frame_opts(use_ransac=0), n_per_problem=N)
        except Exception as e:
            print(B, N, host, env, "ERROR", e); return None
        p = r.poses if host else r.poses.cpu().numpy()
        return p
    base = run([])
    base2 = run([])
    if base is None: continue
    print(B, N, "run-to-run mismatches", int((base != base2).sum()))
    for k in SW:
        p = run([k])
        if p is not None:
            print(B, N, "off:", k, "mismatching pairs", int((p != base).any(axis=1).sum()), "max", float(np.abs(p - base).max()))
    p = run(list(SW))
    print(B, N, "all off: mismatching pairs", int((p != base).any(axis=1).sum()))
