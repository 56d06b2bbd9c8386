"""solve_slots_kernel vs solve_kernel<V, 4, .>: bitwise comparison and timing on several shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnec_b200 import api, synthetic as syn

dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def handle(**env):
    for k, v in env.items():
        os.environ[k] = str(v)
    h = api.Handle(0)  # switches are read at handle creation
    for k in env:
        os.environ.pop(k)
    return h


def timed(fn, reps=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def case(name, b, variant, **kw):
    ct = None if variant == api.NEC else b.covs_target
    ch = b.covs_host if variant == api.SYMMETRIC else None
    args = (dev(b.bvs_host), dev(b.bvs_target), dev(ct), dev(ch), dev(b.init_poses), api.default_opts(variant))
    old = handle(PNEC_B200_SOLVE_SLOTS=0, PNEC_B200_SOLVE_WARPS=4)
    new = handle(PNEC_B200_SOLVE_SLOTS=2)
    dflt = handle(PNEC_B200_SOLVE_SLOTS=0)
    ro = old.solve_batch(*args, **kw)
    rn = new.solve_batch(*args, **kw)
    torch.cuda.synchronize()
    same = all(torch.equal(getattr(ro, f), getattr(rn, f)) for f in ("poses", "status", "iterations", "cost", "initial_cost"))
    if not same:
        rn2 = new.solve_batch(*args, **kw)
        torch.cuda.synchronize()
        bad = (ro.poses != rn.poses).any(dim=1) | (ro.iterations != rn.iterations) | (ro.cost != rn.cost) | (ro.initial_cost != rn.initial_cost)
        idx = torch.nonzero(bad).flatten()
        print("  mismatching pairs:", int(bad.sum()), "first:", idx[:8].tolist(),
              "max |dpose|", float((ro.poses - rn.poses).abs().max()), "iter diff", int((ro.iterations != rn.iterations).sum()),
              "init cost diff", int((ro.initial_cost != rn.initial_cost).sum()),
              "slots deterministic:", bool(torch.equal(rn.poses, rn2.poses)))
        for i in idx[:3].tolist():
            print("   pair", i, "old", ro.poses[i].tolist(), int(ro.iterations[i]), float(ro.initial_cost[i]))
            print("   pair", i, "new", rn.poses[i].tolist(), int(rn.iterations[i]), float(rn.initial_cost[i]))
    row = dict(case=name, B=int(b.num_problems), bit_identical=bool(same),
               ms_old4=timed(lambda: old.solve_batch(*args, out=ro, **kw)), ms_old_default=timed(lambda: dflt.solve_batch(*args, out=ro, **kw)),
               ms_slots=timed(lambda: new.solve_batch(*args, out=rn, **kw)),
               mean_iterations=float(rn.iterations.double().mean()))
    print(json.dumps(row), flush=True)
    return same


ok = True
b = syn.make_batch(10000, 512, seed=2024)
ok &= case("C2 target 10000x512", b, api.TARGET, n_per_problem=512)
if os.environ.get("SLOTS_QUICK"):
    sys.exit(0 if ok else 1)
b = syn.with_host_covariances(syn.make_batch(6000, 200, seed=7))
for v, nm in ((api.NEC, "nec"), (api.HOST, "host"), (api.SYMMETRIC, "symmetric"), (api.TARGET, "target")):
    ok &= case(f"6000x200 {nm}", b, v, n_per_problem=200)
b = syn.make_batch(12500, 256, seed=3)
ok &= case("C3 shard 12500x256", b, api.TARGET, n_per_problem=256)
b = syn.make_batch(3001, 333, seed=5)  # odd N: pairs start at odd correspondence indices, odd batch end
ok &= case("3001x333 odd", b, api.TARGET, n_per_problem=333)
counts = np.clip(syn.kitti_like_counts(3000) // 4, 0, 560)
counts[::97] = 0
b = syn.make_batch(3000, 0, seed=6, camera=syn.PINHOLE, counts=counts)
ok &= case("ragged 3000 (0..560)", b, api.TARGET, offsets=b.offsets)
b = syn.make_batch(7, 512, seed=8)
ok &= case("7x512 (fewer pairs than slots)", b, api.TARGET, n_per_problem=512)
print("ALL BIT-IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
