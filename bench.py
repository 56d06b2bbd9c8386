#!/usr/bin/env python
"""bench.py — frame-pair solves/sec of the batched PNEC refinement on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

One "step" = one batched on-device LM solve of the workload (BASELINE.json config C2:
10 000 synthetic frame pairs x 512 correspondences per GPU, anisotropic-inhomogeneous
per-point 3x3 covariances, Target PNEC residual, Ceres-default LM options), inputs
resident in HBM.  For N > 1 the batch is N independent shards (weak scaling) and every
step ends with the NCCL all-gather of the poses.  Prints ONE JSON line (rank 0).

  value      whole-job solves/s, device-timed (CUDA events), max over ranks
  e2e        same metric through the C-ABI with HOST (pinned) buffers: H2D of the inputs
             and D2H of poses/status inside the timed region.  The covariances cross the link in
             the C-ABI's packed symmetric layout (96 B per correspondence instead of 120; the
             kernels only ever use sym(S), results are bit-identical -- asserted).
             e2e_full_layout: the same call with the reference's own 3x3 layout (120 B).
             e2e_keypoints: a C2-shaped PINHOLE batch handed over as the keypoints the
             reference's frames hold (pixel + 2x2 image covariance, 64 B per correspondence;
             Frame2Frame::GetFeatures / KeyPoint::Unproject run on the device), through
             pnec_solve_from_keypoints_batch
  roofline   the fused residual+Jacobian+JtJ kernel (K1, pnec_eval_batch) timed live:
             algorithmic bytes B*N*120 / CUDA-event time vs the measured HBM peak
  cpu_baseline   the CPU oracle (port of the reference's Ceres path) on the host cores
  frame_pipeline the whole frame-to-frame solve (PNEC::Solve without RANSAC: NEC eigensolver ->
             9 weighted eigensolver + SCF rounds -> refinement, pnec_frame_solve_batch) on the
             same workload, next to the oracle's restatement of it on the host cores (N = 1 only;
             an extra object, the headline metric above is unchanged)
  frame_pipeline_ransac  the same with the reference's DEFAULT options (RANSAC over the eigensolver,
             5000 iterations max, sample size 10), and its CPU figure

--config c3 runs BASELINE config C3 instead: ONE batch of 100 000 frame pairs x 256 correspondences
sharded contiguously over the N ranks (strong scaling) through pnec_b200.distributed.solve_sharded.

--impl reference times the reference's CPU algorithm (the oracle port; the real Ceres
stack cannot be built here) with all host threads on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 10000
N_CORR = 512
BYTES_PER_CORR = 120  # 3 + 3 + 9 doubles
METRIC = "frame-pair solves/sec (batched)"
UNIT = "solves/s"
WORKLOAD = ("C2: 10000 synthetic frame pairs x 512 correspondences per GPU, anisotropic-inhomogeneous "
            "per-point 3x3 covariances, Target PNEC residual, Ceres-default LM (max 50 iterations, "
            "function_tolerance 1e-6)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3"])
    ap.add_argument("--problems", type=int, default=B_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--corr", type=int, default=N_CORR, help=argparse.SUPPRESS)
    return ap.parse_args()


def make_workload(B, N, seed):
    from pnec_b200 import synthetic

    return synthetic.make_batch(B, N, seed=seed, camera=synthetic.OMNIDIRECTIONAL,
                                noise_type="anisotropic_inhomogenous", noise_level=1.0)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes(which="k1_eval"):
    """dram read+write bytes per launch from the committed ncu capture (profiles/summary.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "summary.json")) as f:
            return json.load(f)[which]["dram_bytes_per_launch"]
    except Exception:
        return None


# ------------------------------------------------------------------ CPU arms


def host_threads():
    """All host cores this process may use (torchrun pins OMP_NUM_THREADS=1: ignore that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_rate(batch, n_problems, threads):
    """solves/s of the oracle (numeric-diff Jacobian + Ceres-style LM) on n_problems of batch."""
    import oracle

    N = batch.n_per_problem
    sl = slice(0, n_problems * N)
    o = oracle.default_opts(oracle.TARGET)
    t0 = time.perf_counter()
    oracle.solve_batch(batch.bvs_host[sl], batch.bvs_target[sl], batch.covs_target[sl], None,
                       batch.init_poses[:n_problems], o, n_per_problem=N, num_threads=threads)
    return n_problems / (time.perf_counter() - t0)


def cpu_frame_rate(batch, n_problems, threads):
    """pairs/s of the oracle's PNEC::Solve restatement (oracle/pnec_oracle_frame.c)."""
    import oracle

    N = batch.n_per_problem
    sl = slice(0, n_problems * N)
    t0 = time.perf_counter()
    oracle.frame_solve_batch(batch.bvs_host[sl], batch.bvs_target[sl], batch.covs_target[sl],
                             batch.init_poses[:n_problems], oracle.default_frame_opts(use_ransac=0), n_per_problem=N,
                             num_threads=threads)
    return n_problems / (time.perf_counter() - t0)


def cpu_frame_ransac_rate(batch, n_problems, threads):
    """pairs/s of the oracle's PNEC::Solve restatement with the reference's default options."""
    import oracle

    N = batch.n_per_problem
    sl = slice(0, n_problems * N)
    t0 = time.perf_counter()
    oracle.frame_solve_batch(batch.bvs_host[sl], batch.bvs_target[sl], batch.covs_target[sl],
                             batch.init_poses[:n_problems], oracle.default_frame_opts(), n_per_problem=N,
                             num_threads=threads)
    return n_problems / (time.perf_counter() - t0)


def bind_to_gpu_numa_node(index):
    """CPU affinity of this rank = the cores NVML reports as local to its GPU, so that pinned buffers
    are first-touched on the GPU's NUMA node.  Returns a description for the JSON line."""
    info = {"cpus_before": len(os.sched_getaffinity(0))}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        info["cpus_after"] = len(os.sched_getaffinity(0))
        try:
            info["numa_node"] = int(open(f"/sys/bus/pci/devices/{pynvml.nvmlDeviceGetPciInfo(h).busId.decode().lower()[4:]}/numa_node").read())
        except Exception:
            info["numa_node"] = None
    except Exception as e:  # affinity is an optimisation, never a requirement
        info["error"] = str(e)[:80]
    return info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle

    oracle.build()
    threads = host_threads()
    batch = make_workload(min(args.problems, 2048), args.corr, seed=1)
    # bounded sample: size each step for about a quarter of a second of wall clock on this box
    rate = 0.0
    for _ in range(max(args.warmup, 1)):
        rate = cpu_oracle_rate(batch, min(batch.num_problems, 8 * threads), threads)
    sample = int(min(batch.num_problems, max(4 * threads, rate * 0.25)))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_rate(batch, sample, threads)
    dt = time.perf_counter() - t0
    value = args.steps * sample / dt
    sample_txt = (f"{sample} of the {args.problems} C2 frame pairs per step (first {sample} problems, "
                  f"{args.corr} correspondences each), OpenMP over problems")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "arm": "CPU oracle: port of the reference's Ceres path "
                   "(per-correspondence functor, central numeric differentiation, Ceres-default LM); "
                   "the reference itself cannot be built in this image (no Ceres/Eigen/Sophus/opengv)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------- clocks


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
            0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------- main


def run_b200(args):
    import torch
    import torch.distributed as dist

    from pnec_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun --nproc-per-node N")
    torch.cuda.set_device(local_rank)
    topology = bind_to_gpu_numa_node(local_rank)  # before any pinned allocation (first touch)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, N = args.problems, args.corr
    batch = make_workload(B, N, seed=1 + rank)
    h = api.Handle(local_rank)
    opts = api.default_opts(api.TARGET)

    to_dev = lambda a: torch.from_numpy(a).to(dev)
    f1, f2, ct, init = (to_dev(batch.bvs_host), to_dev(batch.bvs_target), to_dev(batch.covs_target),
                        to_dev(batch.init_poses))
    new_out = lambda: api.SolveResult(torch.empty((B, 7), dtype=torch.float64, device=dev),
                                      torch.empty((B,), dtype=torch.int32, device=dev),
                                      torch.empty((B,), dtype=torch.int32, device=dev),
                                      torch.empty((B,), dtype=torch.float64, device=dev),
                                      torch.empty((B,), dtype=torch.float64, device=dev))
    # N > 1: every step ends with the all-gather of its poses, taken off the critical path: it runs on
    # a side stream while the next step solves (results are double-buffered), and the timed region
    # ends only when the last gather has landed.
    outs = [new_out(), new_out()]
    out = outs[0]
    gathered = [torch.empty((world * B, 7), dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    solved = [torch.cuda.Event(), torch.cuda.Event()]
    gathered_ev = [torch.cuda.Event(), torch.cuda.Event()]
    step_no = [0]

    def step():
        k = step_no[0] & 1
        step_no[0] += 1
        main = torch.cuda.current_stream(dev)
        if world > 1:
            main.wait_event(gathered_ev[k])  # the gather that last read this buffer pair is done
        h.solve_batch(f1, f2, ct, None, init, opts, n_per_problem=N, out=outs[k])
        if world > 1:
            solved[k].record(main)
            with torch.cuda.stream(side):
                side.wait_event(solved[k])
                dist.all_gather_into_tensor(gathered[k], outs[k].poses)
                gathered_ev[k].record(side)

    def sync_all():
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(side)
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing
    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    launches0 = h.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            step()
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(side)  # the last gathers are inside the timed region
        e1.record()
        sync_all()
    launches = h.launch_count - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step * 1e-3)
    iters = outs[0].iterations.float().mean().item()

    # ---- K1 roofline kernel, timed live (same inputs, CUDA events on the launch stream)
    ev_out = api.EvalResult(torch.empty((B,), dtype=torch.float64, device=dev),
                            torch.empty((B, 5), dtype=torch.float64, device=dev),
                            torch.empty((B, 15), dtype=torch.float64, device=dev))
    out = outs[0]
    k1 = lambda: h.eval_batch(f1, f2, ct, None, init, api.TARGET, 1e-13, n_per_problem=N, out=ev_out)
    for _ in range(5):
        k1()
    torch.cuda.synchronize()
    k1_reps = 50
    e0.record()
    for _ in range(k1_reps):
        k1()
    e1.record()
    torch.cuda.synchronize()
    k1_ms = e0.elapsed_time(e1) / k1_reps
    peak, peak_src = measured_peak_gbs()
    achieved = B * N * BYTES_PER_CORR / (k1_ms * 1e-3) / 1e9
    # `roofline` is the kernel BASELINE.json's second metric names (the fused residual + Jacobian +
    # JtJ kernel, K1).  It is launched by this measurement loop, not by the timed solve step, whose
    # kernel (solve_slots_kernel, after the start-point kernel and the two small kernels that order the
    # pairs by the conditioning of their normal equations: 4 launches per step)
    # is reported next to it as `roofline_step_kernel`.
    roofline = {"bound": "hbm", "kernel": "eval_warp_kernel (fused residual + Jacobian + JtJ/Jtr, K1)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_bytes(), "peak_source": peak_src, "ms_per_launch": k1_ms,
                "algorithmic_bytes_per_launch": B * N * BYTES_PER_CORR, "launches_timed": k1_reps,
                "in_timed_step": False}
    step_gbs = B * N * BYTES_PER_CORR / (ms_per_step * 1e-3) / 1e9
    roofline_step = {"bound": "hbm", "kernel": "solve_slots_kernel (whole LM solve, inputs read from HBM once; "
                                                "preceded by solve_prep_kernel, solve_score_kernel and solve_order_scatter_kernel, "
                                                "~25 microseconds together, inside the timed step)",
                     "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                     "traffic": ncu_traffic_bytes("solve"), "ms_per_launch": ms_per_step,
                     "algorithmic_bytes_per_launch": B * N * BYTES_PER_CORR,
                     "mean_lm_iterations": iters,
                     "note": "latency bound, not HBM bound: ~4.2 evaluations + ~3.3 serial LM updates per pair on "
                             "one pass of data, four pairs resident per SM; the pairs that run into max_num_iterations "
                             "(0.18 ms each) are started first (DESIGN.md section 3)"}

    # ---- end to end through the C-ABI with HOST buffers (pinned), copies inside the timed region
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    new_hout = lambda: api.SolveResult(torch.empty((B, 7), dtype=torch.float64).pin_memory().numpy(),
                                       torch.empty((B,), dtype=torch.int32).pin_memory().numpy(),
                                       torch.empty((B,), dtype=torch.int32).pin_memory().numpy(),
                                       torch.empty((B,), dtype=torch.float64).pin_memory().numpy(),
                                       torch.empty((B,), dtype=torch.float64).pin_memory().numpy())
    e2e_steps = max(5, min(args.steps, 20))
    ref_poses = outs[0].poses.cpu().numpy()

    def time_e2e(fn, hout, expect):
        for _ in range(3):
            fn()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()  # synchronous: returns with the poses in host memory
        dt = max_over_ranks(time.perf_counter() - t0)
        assert np.array_equal(hout.poses, expect), "host and device paths disagree"
        return dt

    hf1, hf2, hinit = pin(batch.bvs_host), pin(batch.bvs_target), pin(batch.init_poses)
    d2h_bytes = B * (56 + 4 + 4 + 8 + 8)
    # (a) the reference's own layout: 3x3 covariances, 120 B per correspondence
    hct = pin(batch.covs_target)
    hout = new_hout()
    dt = time_e2e(lambda: h.solve_batch(hf1, hf2, hct, None, hinit, opts, n_per_problem=N, out=hout), hout, ref_poses)
    e2e_full = {"value": world * B * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": B * N * BYTES_PER_CORR + B * 56,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
                "host_buffers": "pinned", "layout": "f1, f2, 3x3 covariance (120 B per correspondence)"}
    del hct
    # (b) packed symmetric covariances: the same numbers the kernels reduce a 3x3 to, 96 B per correspondence
    c9 = batch.covs_target
    hpk = pin(np.stack([c9[:, 0], 0.5 * (c9[:, 1] + c9[:, 3]), 0.5 * (c9[:, 2] + c9[:, 6]), c9[:, 4],
                        0.5 * (c9[:, 5] + c9[:, 7]), c9[:, 8]], axis=1))
    dt = time_e2e(lambda: h.solve_batch(hf1, hf2, hpk, None, hinit, opts, n_per_problem=N, out=hout), hout, ref_poses)
    e2e = {"value": world * B * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": B * N * 96 + B * 56,
           "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
           "host_buffers": "pinned",
           "layout": "f1, f2, packed symmetric covariance (PNEC_COV_PACKED, 96 B per correspondence); "
                     "poses bit-identical to the device-resident and to the 3x3-layout call (asserted)"}
    del hpk, hf1, hf2
    # (c) from keypoints: a C2-shaped pinhole batch as the fields a KeyPoint is built from
    e2e_kp = None
    if args.config == "c2":
        from pnec_b200 import synthetic

        kb = synthetic.make_batch(B, N, seed=101 + rank, camera=synthetic.PINHOLE)
        Kc = np.array([[800.0, 0.0, 320.0], [0.0, 800.0, 240.0], [0.0, 0.0, 1.0]])
        kinv = np.ascontiguousarray(np.linalg.inv(Kc).T).reshape(9)
        pix = lambda f: (f[:, :2] / f[:, 2:3]) * 800.0 + Kc[:2, 2]
        rng = np.random.default_rng(7 + rank)
        cov2 = synthetic.sample_covariances_2d(rng, (1, B * N), 1.0, "anisotropic_inhomogenous")[0]
        hp, tp = pin(pix(kb.bvs_host)), pin(pix(kb.bvs_target))
        c3 = pin(np.stack([cov2[:, 0, 0], cov2[:, 0, 1], cov2[:, 1, 1]], axis=1))
        kinit = pin(kb.init_poses)
        # the 120 B route on the same keypoints: unproject on the device, then the batch entry point
        kf1, _ = h.keypoints_unproject(to_dev(pix(kb.bvs_host)), to_dev(np.swapaxes(cov2, -1, -2).reshape(-1, 4)), kinv)
        kf2, kct = h.keypoints_unproject(to_dev(pix(kb.bvs_target)), to_dev(np.swapaxes(cov2, -1, -2).reshape(-1, 4)), kinv)
        kref = h.solve_batch(kf1, kf2, kct, None, to_dev(kb.init_poses), opts, n_per_problem=N).poses.cpu().numpy()
        del kf1, kf2, kct
        dt = time_e2e(lambda: h.solve_from_keypoints(hp, tp, c3, kinit, kinv, opts, packed=True, n_per_problem=N,
                                                     out=hout), hout, kref)
        e2e_kp = {"value": world * B * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": B * N * 56 + B * 56,
                  "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
                  "host_buffers": "pinned",
                  "workload": "C2-shaped pinhole batch (10000 x 512, anisotropic-inhomogeneous 2x2 image covariances) "
                              "handed over as keypoints: pixel of the host keypoint (16 B), pixel + packed 2x2 covariance "
                              "of the target keypoint (40 B); unprojection + unscented transform on the device "
                              "(pnec_solve_from_keypoints_batch); poses bit-identical to unprojecting first and "
                              "calling pnec_solve_batch with 120 B per correspondence (asserted)"}
        del hp, tp, c3

    # ---- the whole frame solve (SURVEY.md section 8f rows 1-2 + the refinement), device resident
    frame = None
    if world == 1:
        fopts = api.default_frame_opts(use_ransac=0)
        fstep = lambda: h.frame_solve_batch(f1, f2, ct, init, fopts, n_per_problem=N)
        for _ in range(3):
            fstep()
        torch.cuda.synchronize()
        f_reps = 10
        l0 = h.launch_count
        e0.record()
        for _ in range(f_reps):
            fstep()
        e1.record()
        torch.cuda.synchronize()
        f_ms = e0.elapsed_time(e1) / f_reps
        frame = {"value": B / (f_ms * 1e-3), "unit": "frame pairs/s", "ms_per_step": f_ms,
                 "gpu_launches_per_step": (h.launch_count - l0) // f_reps,
                 "config": "PNEC::Solve, use_ransac_=false, weighted_iterations_=10, use_ceres_=true "
                           "(pnec_frame_solve_batch) on the C2 batch, inputs resident in HBM"}

        ropts = api.default_frame_opts()
        rstep = lambda: h.frame_solve_batch(f1, f2, ct, init, ropts, n_per_problem=N)
        for _ in range(2):
            rres = rstep()
        torch.cuda.synchronize()
        l0 = h.launch_count
        e0.record()
        for _ in range(3):
            rstep()
        e1.record()
        torch.cuda.synchronize()
        r_ms = e0.elapsed_time(e1) / 3
        frame_ransac = {"value": B / (r_ms * 1e-3), "unit": "frame pairs/s", "ms_per_step": r_ms,
                        "gpu_launches_per_step": (h.launch_count - l0) // 3,
                        "mean_ransac_iterations": float(rres.ransac_iterations.float().mean().item()),
                        "pairs_at_max_iterations": int((rres.ransac_iterations > 5000).sum().item()),
                        "mean_inlier_fraction": float(rres.num_inliers.float().mean().item() / N),
                        "config": "PNEC::Solve with Options() defaults: use_ransac_=true (max 5000 iterations, sample "
                                  "size 10, threshold 1e-6), weighted_iterations_=10, use_ceres_=true, on the C2 "
                                  "batch, inputs resident in HBM"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "problems_per_gpu": B, "correspondences": N,
                   "l2": f"inputs are {B * N * BYTES_PER_CORR / 1e6:.0f} MB per GPU, larger than the 126 MB L2 "
                         "(no flush needed)",
                   "parallelism": f"{world} independent shard(s), NCCL all-gather of poses only" if world > 1
                   else "single GPU"},
        "clocks": clocks.summary(), "e2e": e2e, "e2e_full_layout": e2e_full, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_step_kernel": roofline_step, "topology": topology,
    }
    if e2e_kp is not None:
        line["e2e_keypoints"] = e2e_kp
    if world == 1:
        import oracle

        threads = host_threads()
        sample = min(B, 8192)
        rate_all = cpu_oracle_rate(batch, sample, threads)
        rate_one = cpu_oracle_rate(batch, min(B, 512), 1)
        line["cpu_baseline"] = {
            "value": rate_all, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {sample} of the {B} C2 frame pairs, OpenMP over problems; "
                      f"1 thread: {rate_one:.1f} solves/s on the first {min(B, 512)}",
            "value_1core": rate_one}
        fsample = min(B, 32 * threads)
        frame["cpu_baseline"] = {"value": cpu_frame_rate(batch, fsample, threads), "unit": "frame pairs/s",
                                 "cores": threads, "kind": "port",
                                 "sample": f"first {fsample} of the {B} C2 frame pairs, OpenMP over pairs"}
        line["frame_pipeline"] = frame
        rsample = min(B, 8 * threads)
        frame_ransac["cpu_baseline"] = {"value": cpu_frame_ransac_rate(batch, rsample, threads), "unit": "frame pairs/s",
                                        "cores": threads, "kind": "port",
                                        "sample": f"first {rsample} of the {B} C2 frame pairs, OpenMP over pairs"}
        line["frame_pipeline_ransac"] = frame_ransac
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_c3(args):
    """BASELINE config C3: ONE batch of 100 000 frame pairs x 256 correspondences, cut contiguously over the
    ranks (pnec_b200.distributed.shard_bounds), every rank solves its shard, poses / status / iterations are
    all-gathered (distributed.gather_results): strong scaling, the split the reference does with process
    fan-out over disjoint inputs (run_simulation.sh:58-65).  Every rank generates only its own shard."""
    import torch
    import torch.distributed as dist

    from pnec_b200 import api, distributed

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B_total, N = 100000, 256
    bounds = distributed.shard_bounds(B_total, world)
    b0, b1 = bounds[rank]
    batch = make_workload(b1 - b0, N, seed=1000 + rank)
    h = api.Handle(local_rank)
    opts = api.default_opts(api.TARGET)
    to_dev = lambda a: torch.from_numpy(a).to(dev)
    f1, f2, ct, init = (to_dev(batch.bvs_host), to_dev(batch.bvs_target), to_dev(batch.covs_target),
                        to_dev(batch.init_poses))

    def step():
        res = h.solve_batch(f1, f2, ct, None, init, opts, n_per_problem=N)
        if world > 1:
            return (distributed.gather_results(res.poses, bounds), distributed.gather_results(res.status, bounds),
                    distributed.gather_results(res.iterations, bounds))
        return res.poses, res.status, res.iterations

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            out = step()
        e1.record()
        sync_all()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert out[0].shape == (B_total, 7) and bool((out[1] <= 4).all())
    if rank == 0:
        ms_per_step = ms / args.steps
        print(json.dumps({
            "metric": METRIC, "value": B_total / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3: one batch of 100000 synthetic frame pairs x 256 correspondences (anisotropic-"
                                   "inhomogeneous 3x3 covariances, Target residual, Ceres-default LM), contiguous shard "
                                   "per GPU, all-gather of poses / status / iterations",
                       "pairs_per_gpu": b1 - b0, "correspondences": N,
                       "l2": f"{(b1 - b0) * N * BYTES_PER_CORR / 1e6:.0f} MB of inputs per GPU (> 126 MB L2)"},
            "clocks": clocks.summary(), "mean_lm_iterations": float(out[2].float().mean().item())}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "c3":
        return run_c3(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
