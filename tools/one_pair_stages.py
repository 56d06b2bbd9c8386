import os, sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle
from conftest import rotation_angle, direction_angle
from pnec_b200 import api, synthetic as syn
b = syn.make_batch(1500, 200, seed=8, camera=syn.PINHOLE)
k = 320; N = 200
sl = slice(k*N, (k+1)*N)
f1, f2, ct, init = b.bvs_host[sl], b.bvs_target[sl], b.covs_target[sl], b.init_poses[k:k+1]
h = api.Handle(0)
def cmp(name, **kw):
    fo = oracle.default_frame_opts(use_ransac=0, **kw)
    ref, ref_es = oracle.frame_solve_batch(f1, f2, ct, init, fo, n_per_problem=N)
    res = h.frame_solve_batch(f1, f2, ct, init, api.default_frame_opts(use_ransac=0, **kw), n_per_problem=N)
    p, es = np.asarray(res.poses), np.asarray(res.es_poses)
    print(name, "final rot", rotation_angle(p[0], ref[0]), "dir", direction_angle(p[0][4:], ref[0][4:]), "| es rot", rotation_angle(es[0], ref_es[0]), "dir", direction_angle(es[0][4:], ref_es[0][4:]))
    return p[0], ref[0]
for wi in (1, 2, 3, 4, 6, 10):
    cmp(f"weighted_iterations={wi} no ceres", weighted_iterations=wi, use_ceres=0)
pg, po = cmp("default", )
# refinement alone from the oracle's pre-refinement pose
fo = oracle.default_frame_opts(use_ransac=0, use_ceres=0)
pre, _ = oracle.frame_solve_batch(f1, f2, ct, init, fo, n_per_problem=N)
r_o, info = oracle.solve_batch(f1, f2, ct, None, pre, oracle.default_opts(1), n_per_problem=N)
r_g = h.solve_batch(f1, f2, ct, None, pre, api.default_opts(api.TARGET), n_per_problem=N)
print("refinement from the oracle's start: rot", rotation_angle(np.asarray(r_g.poses)[0], r_o[0]), "dir", direction_angle(np.asarray(r_g.poses)[0][4:], r_o[0][4:]), "iters gpu", int(np.asarray(r_g.iterations)[0]), "oracle", int(info["iterations"][0]), "status", int(np.asarray(r_g.status)[0]), int(info["status"][0]), "cost", float(np.asarray(r_g.cost)[0]), float(info["final_cost"][0]))
print("gt", b.gt_poses[k]); print("gpu", pg); print("oracle", po)
