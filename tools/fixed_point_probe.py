"""How quickly does the weighted-eigensolver iteration (pnec.cc:283-348) reach a bitwise fixed point?
Per iteration: pairs whose rotation / whole pose repeat the previous iteration bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn
B, N = 2000, int(os.environ.get("FT_N", 512))
h = api.Handle(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
b = syn.make_batch(B, N, seed=2)
f1, f2, ct, init = T(b.bvs_host), T(b.bvs_target), T(b.covs_target), T(b.init_poses)
es, _, _ = h.eigensolver_batch(f1, f2, init, n_per_problem=N)
t, _ = h.nec_translation_batch(f1, f2, es, n_per_problem=N)
es[:, 4:] = t
rel = es.clone()
for it in range(9):
    nxt, info, _ = h.eigensolver_batch(f1, f2, rel, covs_target=ct, weight_poses=es, n_per_problem=N)
    t, _ = h.scf_translation_batch(f1, f2, ct, nxt, n_per_problem=N)
    nxt[:, 4:] = t
    same_q = (nxt[:, :4] == rel[:, :4]).all(dim=1)
    same = (nxt == rel).all(dim=1)
    dq = (nxt[:, :4] - rel[:, :4]).abs().max(dim=1).values
    dt = torch.minimum((nxt[:, 4:] - rel[:, 4:]).abs().max(dim=1).values, (nxt[:, 4:] + rel[:, 4:]).abs().max(dim=1).values)
    print(f"it {it}: same rotation {int(same_q.sum())}/{B}, same pose {int(same.sum())}/{B}, "
          f"median |dq| {dq.median().item():.2e}, median |dt| {dt.median().item():.2e}, max |dt| {dt.max().item():.2e}, info hist {np.bincount(info.cpu().numpy(), minlength=9).tolist()}")
    rel = nxt
