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
             and D2H of poses/status inside the timed region
  roofline   the fused residual+Jacobian+JtJ kernel (K1, pnec_eval_batch) timed live:
             algorithmic bytes B*N*120 / CUDA-event time vs the measured HBM peak
  cpu_baseline   the CPU oracle (port of the reference's Ceres path) on the host cores
  frame_pipeline the whole frame-to-frame solve (PNEC::Solve without RANSAC: NEC eigensolver ->
             9 weighted eigensolver + SCF rounds -> refinement, pnec_frame_solve_batch) on the
             same workload, next to the oracle's restatement of it on the host cores (N = 1 only;
             an extra object, the headline metric above is unchanged)

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
    out = api.SolveResult(torch.empty((B, 7), dtype=torch.float64, device=dev),
                          torch.empty((B,), dtype=torch.int32, device=dev),
                          torch.empty((B,), dtype=torch.int32, device=dev),
                          torch.empty((B,), dtype=torch.float64, device=dev),
                          torch.empty((B,), dtype=torch.float64, device=dev))
    gathered = torch.empty((world * B, 7), dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        h.solve_batch(f1, f2, ct, None, init, opts, n_per_problem=N, out=out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out.poses)

    def sync_all():
        if world > 1:
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
        e1.record()
        sync_all()
    launches = h.launch_count - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step * 1e-3)
    iters = out.iterations.float().mean().item()

    # ---- K1 roofline kernel, timed live (same inputs, CUDA events on the launch stream)
    ev_out = api.EvalResult(torch.empty((B,), dtype=torch.float64, device=dev),
                            torch.empty((B, 5), dtype=torch.float64, device=dev),
                            torch.empty((B, 15), dtype=torch.float64, device=dev))
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
    # one kernel per step is reported next to it as `roofline_step_kernel`.
    roofline = {"bound": "hbm", "kernel": "eval_warp_kernel (fused residual + Jacobian + JtJ/Jtr, K1)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_bytes(), "peak_source": peak_src, "ms_per_launch": k1_ms,
                "algorithmic_bytes_per_launch": B * N * BYTES_PER_CORR, "launches_timed": k1_reps,
                "in_timed_step": False}
    step_gbs = B * N * BYTES_PER_CORR / (ms_per_step * 1e-3) / 1e9
    roofline_step = {"bound": "hbm", "kernel": "solve_kernel (whole LM solve, inputs read from HBM once)",
                     "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                     "traffic": ncu_traffic_bytes("solve"), "ms_per_launch": ms_per_step,
                     "algorithmic_bytes_per_launch": B * N * BYTES_PER_CORR,
                     "mean_lm_iterations": iters,
                     "note": "latency/fp64-issue bound, not HBM bound: ~4.3 evaluations + ~3.3 serial LM "
                             "updates per pair on one pass of data (DESIGN.md section 3)"}

    # ---- end to end through the C-ABI with HOST buffers (pinned), copies inside the timed region
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    hf1, hf2, hct, hinit = pin(batch.bvs_host), pin(batch.bvs_target), pin(batch.covs_target), pin(batch.init_poses)
    hout = api.SolveResult(torch.empty((B, 7), dtype=torch.float64).pin_memory().numpy(),
                           torch.empty((B,), dtype=torch.int32).pin_memory().numpy(),
                           torch.empty((B,), dtype=torch.int32).pin_memory().numpy(),
                           torch.empty((B,), dtype=torch.float64).pin_memory().numpy(),
                           torch.empty((B,), dtype=torch.float64).pin_memory().numpy())
    e2e_step = lambda: h.solve_batch(hf1, hf2, hct, None, hinit, opts, n_per_problem=N, out=hout)
    for _ in range(3):
        e2e_step()
    e2e_steps = max(5, min(args.steps, 20))
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()  # synchronous: returns with the poses in host memory
    dt = max_over_ranks(time.perf_counter() - t0)
    assert np.array_equal(hout.poses, out.poses.cpu().numpy()), "host and device paths disagree"
    e2e = {"value": world * B * e2e_steps / dt, "unit": UNIT,
           "h2d_bytes_per_step": B * N * BYTES_PER_CORR + B * 56,
           "d2h_bytes_per_step": B * (56 + 4 + 4 + 8 + 8), "ms_per_step": 1e3 * dt / e2e_steps,
           "steps": e2e_steps, "host_buffers": "pinned"}

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
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "roofline_step_kernel": roofline_step,
    }
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
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
