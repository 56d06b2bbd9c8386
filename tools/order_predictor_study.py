"""Which cheap quantity at the START point predicts the frame pairs whose refinement runs long?  (CPU, oracle only.)

For a C2-shaped batch the oracle solves every pair (iteration counts), then J^T J of the first M correspondences
of a pair at its start pose is formed (oracle_eval) and several scores are ranked: for every pair with >= 20
iterations, the fraction of ordinary pairs (a random sample with < 10 iterations) that score LOWER.  A good
predictor puts all long pairs near 1.0.  The CUDA path orders its pairs by `mindiag/maxdiag`
(solve_score_kernel, pnec_b200/csrc/pnec_solve_slots.cuh).

    python tools/order_predictor_study.py [M=32] [seed=1]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from pnec_b200 import synthetic as syn

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B, N = 10000, 512
b = syn.make_batch(B, N, seed=seed)
_, info = oracle.solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses,
                             oracle.default_opts(oracle.TARGET), n_per_problem=N, num_threads=oracle.max_threads())
it = info["iterations"]
ct = oracle.covs_to_abi(b.covs_target) if b.covs_target.ndim == 3 else b.covs_target
longs = np.nonzero(it >= 20)[0]
ordinary = np.setdiff1d(np.random.default_rng(0).choice(B, 3000, replace=False), np.nonzero(it >= 10)[0])
names = ["initial cost / M", "|gradient|", "cond(JtJ) = lmax / lmin", "-(min diag / max diag)", "-lmin of Jacobi-scaled JtJ",
         "-smallest Cholesky pivot of Jacobi-scaled JtJ"]
feat = {}
for k in np.concatenate([longs, ordinary]):
    s = slice(k * N, k * N + M)
    r = oracle.evaluate(oracle.TARGET, b.bvs_host[s], b.bvs_target[s], ct[s], ct[s], 1e-13, b.init_poses[k],
                        oracle.JAC_NUMERIC_CENTRAL)
    H = np.zeros((5, 5)); H[np.triu_indices(5)] = r.jtj; H = H + H.T - np.diag(np.diag(H))
    w = np.linalg.eigvalsh(H)
    d = np.diag(H)
    Hs = H / np.sqrt(np.outer(d, d))
    try:
        piv = np.diag(np.linalg.cholesky(Hs)) ** 2
    except np.linalg.LinAlgError:
        piv = np.array([0.0])
    feat[k] = [r.cost / M, np.linalg.norm(r.gradient), w[-1] / max(w[0], 1e-300), -d.min() / d.max(),
               -np.linalg.eigvalsh(Hs)[0], -piv.min()]
F = lambda ks, j: np.array([feat[k][j] for k in ks])
out = {"B": B, "N": N, "seed": seed, "correspondences used": M, "pairs with >= 20 iterations": int(len(longs)),
       "their iteration counts": sorted(int(x) for x in it[longs]), "ordinary pairs sampled": int(len(ordinary)), "scores": {}}
for j, nm in enumerate(names):
    f = F(ordinary, j)
    pct = np.sort([(f < feat[k][j]).mean() for k in longs])
    out["scores"][nm] = {"lowest percentile of a long pair": round(float(pct[0]), 3),
                         "five lowest": [round(float(x), 3) for x in pct[:5]]}
print(json.dumps(out))
