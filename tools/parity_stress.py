"""One-off parity stress: GPU solve vs oracle on many problems of several shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import oracle
from conftest import rotation_angle, direction_angle
from pnec_b200 import api, synthetic as syn
h = api.Handle(0)
dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
threads = len(os.sched_getaffinity(0))
def run(name, b, variant, **kw):
    ct = None if variant == api.NEC else b.covs_target
    ch = b.covs_host if variant == api.SYMMETRIC else None
    res = h.solve_batch(dev(b.bvs_host), dev(b.bvs_target), dev(ct), dev(ch), dev(b.init_poses), api.default_opts(variant), **kw)
    t0 = time.time()
    ref, info = oracle.solve_batch(b.bvs_host, b.bvs_target, ct, ch, b.init_poses, oracle.default_opts(variant), num_threads=threads, **kw)
    dt = time.time() - t0
    poses = res.poses.cpu().numpy(); its = res.iterations.cpu().numpy(); st = res.status.cpu().numpy()
    rot = np.array([rotation_angle(a, c) for a, c in zip(poses, ref)])
    tra = np.array([direction_angle(a[4:], c[4:]) for a, c in zip(poses, ref)])
    mism = np.nonzero(its != info["iterations"])[0]
    print(f"{name}: B={len(poses)} oracle {dt:.1f}s  max rot {rot.max():.2e} max t {tra.max():.2e}  iteration mismatches {len(mism)} status mismatches {int((st != info['status']).sum())}"
          f"  >1e-6: rot {int((rot > 1e-6).sum())} t {int((tra > 1e-6).sum())}", flush=True)
    for i in mism[:5]:
        print("   mismatch", i, its[i], info["iterations"][i], st[i], info["status"][i], rot[i], tra[i])
b = syn.make_batch(10000, 512, seed=2024)
run("C2 target", b, api.TARGET, n_per_problem=512)
b = syn.with_host_covariances(syn.make_batch(6000, 200, seed=7))
for v, nm in ((api.NEC, "nec"), (api.HOST, "host"), (api.SYMMETRIC, "symmetric")):
    run(f"6000x200 {nm}", b, v, n_per_problem=200)
for nt in syn.NOISE_TYPES:
    for cam in (syn.OMNIDIRECTIONAL, syn.PINHOLE):
        b = syn.make_batch(3000, 100, seed=11, camera=cam, noise_type=nt, noise_level=2.0)
        run(f"3000x100 {cam[:4]} {nt}", b, api.TARGET, n_per_problem=100)
counts = syn.kitti_like_counts(600)
b = syn.make_batch(600, 0, seed=5, camera=syn.PINHOLE, counts=counts)
run("KITTI-like ragged 600", b, api.TARGET, offsets=b.offsets)
