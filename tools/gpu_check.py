"""Ad-hoc GPU bring-up: parity of eval/solve vs the oracle and rough timings.
Usage (under gpurun): python tools/gpu_check.py [--big]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from pnec_b200 import api, synthetic as syn

dev = torch.device("cuda", 0)
h = api.Handle(0)
T = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rad = np.deg2rad

def check(variant, B, N, seed, camera=syn.OMNIDIRECTIONAL, counts=None):
    b = syn.make_batch(B, N, seed=seed, camera=camera, counts=counts)
    b = syn.with_host_covariances(b, seed=seed + 1, camera=camera)
    ct = None if variant == api.NEC else b.covs_target
    ch = b.covs_host if variant == api.SYMMETRIC else None
    kw = dict(offsets=b.offsets) if counts is not None else dict(n_per_problem=N)
    ev = h.eval_batch(T(b.bvs_host), T(b.bvs_target), T(ct), T(ch), T(b.init_poses), variant, 1e-13, **kw)
    res = h.solve_batch(T(b.bvs_host), T(b.bvs_target), T(ct), T(ch), T(b.init_poses), api.default_opts(variant), **kw)
    torch.cuda.synchronize()
    # host-memspace path too
    res_h = h.solve_batch(b.bvs_host, b.bvs_target, ct, ch, b.init_poses, api.default_opts(variant), **kw)
    poses = res.poses.cpu().numpy()
    assert np.array_equal(poses, res_h.poses), "host/device memspace mismatch"
    worst = dict(ev_cost=0, ev_g=0, ev_h=0, rot=0, tr=0, rot_a=0, tr_a=0, it_mismatch=0)
    on = oracle.default_opts(variant)
    oa = oracle.default_opts(variant, jacobian_mode=oracle.JAC_ANALYTIC)
    okw = dict(offsets=b.offsets) if counts is not None else dict(n_per_problem=N)
    ref, info = oracle.solve_batch(b.bvs_host, b.bvs_target, ct, ch, b.init_poses, on, num_threads=8, **okw)
    refa, infoa = oracle.solve_batch(b.bvs_host, b.bvs_target, ct, ch, b.init_poses, oa, num_threads=8, **okw)
    its = res.iterations.cpu().numpy(); sts = res.status.cpu().numpy()
    for i in range(B):
        s, e = b.range(i)
        if i < 8:
            o = oracle.evaluate(variant, b.bvs_host[s:e], b.bvs_target[s:e], None if ct is None else ct[s:e],
                                None if ch is None else ch[s:e], 1e-13, b.init_poses[i], oracle.JAC_ANALYTIC)
            worst['ev_cost'] = max(worst['ev_cost'], abs(ev.cost[i].item() - o.cost) / abs(o.cost))
            worst['ev_g'] = max(worst['ev_g'], np.abs(ev.gradient[i].cpu().numpy() - o.gradient).max() / np.abs(o.gradient).max())
            worst['ev_h'] = max(worst['ev_h'], np.abs(ev.jtj[i].cpu().numpy() - o.jtj).max() / np.abs(o.jtj).max())
        worst['rot'] = max(worst['rot'], rad(oracle.rotational_difference_deg(poses[i], ref[i])))
        worst['tr'] = max(worst['tr'], rad(oracle.translational_difference_deg(poses[i][4:], ref[i][4:])))
        worst['rot_a'] = max(worst['rot_a'], rad(oracle.rotational_difference_deg(poses[i], refa[i])))
        worst['tr_a'] = max(worst['tr_a'], rad(oracle.translational_difference_deg(poses[i][4:], refa[i][4:])))
    worst['it_mismatch'] = int((its != info['iterations']).sum())
    worst['st_mismatch'] = int((sts != info['status']).sum())
    print(f"variant={variant} B={B} N={N} ragged={counts is not None}:", {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in worst.items()},
          "iters", np.bincount(its), flush=True)

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b_.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b_))
    return float(np.median(ts)), float(np.min(ts))

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    for v in (api.TARGET, api.NEC, api.HOST, api.SYMMETRIC):
        check(v, 32, 100, 1)
        check(v, 16, 512, 2)
    check(api.TARGET, 24, 0, 3, counts=np.array([1, 2, 3, 5, 7, 10, 31, 33, 64, 65, 127, 129, 255, 300, 511, 513, 1000, 1023, 1500, 1887, 1889, 2500, 3000, 77]))
    check(api.TARGET, 8, 3000, 4)
    check(api.TARGET, 64, 64, 5, camera=syn.PINHOLE)
    if "--big" in sys.argv:
        B, N = 10000, 512
        t0 = time.time(); b = syn.make_batch(B, N, seed=11); print("gen s", time.time() - t0, flush=True)
        f1, f2, ct, init = T(b.bvs_host), T(b.bvs_target), T(b.covs_target), T(b.init_poses)
        opts = api.default_opts(api.TARGET)
        for cfg in ("11", "14", "17"):
            os.environ["PNEC_B200_EVAL_CFG"] = cfg; h = api.Handle(0)  # switches are read at handle creation
            med, mn = timeit(lambda: h.eval_batch(f1, f2, ct, None, init, api.TARGET, 1e-13, n_per_problem=N))
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for _ in range(50): h.eval_batch(f1, f2, ct, None, init, api.TARGET, 1e-13, n_per_problem=N)
            b_.record(); torch.cuda.synchronize(); loop = a_.elapsed_time(b_) / 50
            print(f'   back-to-back: {loop:.4f} ms -> {B*N*120/loop/1e6:.1f} GB/s')
            print(f"eval cfg{cfg}: med {med:.4f} ms min {mn:.4f} ms -> {B*N*120/mn/1e6:.1f} GB/s", flush=True)
        for nw in ("2", "4", "8"):
            os.environ["PNEC_B200_SOLVE_WARPS"] = nw; h = api.Handle(0)  # switches are read at handle creation
            med, mn = timeit(lambda: h.solve_batch(f1, f2, ct, None, init, opts, n_per_problem=N))
            print(f"solve nw{nw}: med {med:.4f} ms min {mn:.4f} ms -> {B/med*1e3:.3e} solves/s", flush=True)
        os.environ.pop("PNEC_B200_SOLVE_WARPS"); h = api.Handle(0)  # switches are read at handle creation
        res = h.solve_batch(f1, f2, ct, None, init, opts, n_per_problem=N); torch.cuda.synchronize()
        print("iters hist", np.bincount(res.iterations.cpu().numpy()), "status", np.bincount(res.status.cpu().numpy()), flush=True)
        t0 = time.time()
        ref, info = oracle.solve_batch(b.bvs_host[:512 * N], b.bvs_target[:512 * N], b.covs_target[:512 * N], None, b.init_poses[:512],
                                       oracle.default_opts(oracle.TARGET), n_per_problem=N, num_threads=oracle.max_threads())
        dt = time.time() - t0
        print(f"oracle 512 problems, {oracle.max_threads()} threads: {dt:.3f}s -> {512/dt:.1f} solves/s", flush=True)
        poses = res.poses.cpu().numpy()[:512]
        wr = max(rad(oracle.rotational_difference_deg(poses[i], ref[i])) for i in range(512))
        wt = max(rad(oracle.translational_difference_deg(poses[i][4:], ref[i][4:])) for i in range(512))
        print("C2 subsample parity: rot", wr, "tr", wt, "iter mismatch", int((res.iterations.cpu().numpy()[:512] != info['iterations']).sum()), flush=True)
