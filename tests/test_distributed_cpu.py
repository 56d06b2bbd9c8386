"""CPU, world_size 2, gloo: the sharded solve + gather reproduces the single-process result.
The per-shard solver here is the oracle (test infrastructure); on GPUs it is the CUDA handle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ragged, out_dir, frame=False):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import oracle
    from pnec_b200 import distributed
    from pnec_b200 import synthetic as syn

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, N = 7, 40
        counts = np.array([5, 40, 1, 0, 33, 64, 12]) if ragged else None
        b = syn.make_batch(B, N, seed=21, counts=counts)

        def solve_fn(f1, f2, ct, ch, poses, offsets=None, n_per_problem=None):
            p, info = oracle.solve_batch(f1, f2, ct, ch, poses, oracle.default_opts(oracle.TARGET),
                                         offsets=offsets, n_per_problem=n_per_problem)
            return (torch.from_numpy(p), torch.from_numpy(info["status"].copy()),
                    torch.from_numpy(info["iterations"].copy()))

        if frame:
            # the whole frame solve (PNEC::Solve) shards the same way: pairs are independent
            fo = oracle.default_frame_opts(use_ransac=0, weighted_iterations=3)

            def frame_fn(f1, f2, ct, ch, poses, offsets=None, n_per_problem=None):
                p, es = oracle.frame_solve_batch(f1, f2, ct, poses, fo, offsets=offsets, n_per_problem=n_per_problem)
                z = torch.zeros(p.shape[0], dtype=torch.int32)
                return torch.from_numpy(p), torch.from_numpy(es), z

            poses, es, _ = distributed.solve_sharded(
                frame_fn, B, N, b.offsets, b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses)
            ref, ref_es = oracle.frame_solve_batch(b.bvs_host, b.bvs_target, b.covs_target, b.init_poses, fo,
                                                   offsets=b.offsets, n_per_problem=N)
            assert np.array_equal(poses.numpy(), ref) and np.array_equal(es.numpy(), ref_es)
            open(os.path.join(out_dir, f"ok{rank}"), "w").close()
            return
        poses, status, iters = distributed.solve_sharded(
            solve_fn, B, N, b.offsets, b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses)
        ref, info = oracle.solve_batch(b.bvs_host, b.bvs_target, b.covs_target, None, b.init_poses,
                                       oracle.default_opts(oracle.TARGET), offsets=b.offsets,
                                       n_per_problem=N)
        assert np.array_equal(poses.numpy(), ref)
        assert np.array_equal(status.numpy(), info["status"])
        assert np.array_equal(iters.numpy(), info["iterations"])
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_sharded_solve_matches_single_process(tmp_path, ragged):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), ragged, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_sharded_frame_solve_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), True, str(tmp_path), True), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def _gather_worker(rank, world, port, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from pnec_b200 import distributed

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for B in (8, 7, 2, 1):  # equal shards (one collective into the result) and unequal ones (padded)
            bounds = distributed.shard_bounds(B, world)
            b0, b1 = bounds[rank]
            full_p = torch.arange(B * 7, dtype=torch.float64).reshape(B, 7) * 0.5
            full_s = torch.arange(B, dtype=torch.int32) * 3
            got_p = distributed.gather_results(full_p[b0:b1].clone(), bounds)
            got_s = distributed.gather_results(full_s[b0:b1].clone(), bounds)
            assert torch.equal(got_p, full_p) and torch.equal(got_s, full_s) and got_s.dtype == torch.int32
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_gather_results_equal_and_unequal_shards(tmp_path):
    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
