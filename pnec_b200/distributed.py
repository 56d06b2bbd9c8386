"""Multi-GPU: frame pairs are independent problems, so the batch is split
contiguously over ranks (balanced by correspondence count for ragged batches),
every rank solves its shard with no data-path communication, and the only
collective is the final gather of 7 doubles + status/iterations per problem.

The reference's analogue is process-level fan-out over disjoint inputs
(run_simulation.sh:58-65); there is no collective to mirror.
One process per GPU; plumbing is torch.distributed (NCCL on GPUs, gloo in tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np


def shard_bounds(num_problems: int, world_size: int,
                 offsets: Optional[np.ndarray] = None) -> List[Tuple[int, int]]:
    """Contiguous [b0, b1) per rank.  Uniform batches: equal problem counts (first
    ranks take the remainder).  Ragged: cut where the running correspondence count
    crosses k/world of the total, so every rank streams about the same bytes."""
    B, W = int(num_problems), int(world_size)
    if W < 1:
        raise ValueError("world_size must be >= 1")
    if offsets is None:
        base, rem = divmod(B, W)
        cuts = [0]
        for r in range(W):
            cuts.append(cuts[-1] + base + (1 if r < rem else 0))
    else:
        offsets = np.asarray(offsets, dtype=np.int64)
        if offsets.shape != (B + 1,):
            raise ValueError("offsets must have num_problems + 1 entries")
        start, total = int(offsets[0]), int(offsets[-1] - offsets[0])
        cuts = [0]
        for r in range(1, W):
            target = start + (total * r) // W
            c = int(np.searchsorted(offsets, target, side="left"))
            cuts.append(min(max(c, cuts[-1]), B))
        cuts.append(B)
    return [(cuts[r], cuts[r + 1]) for r in range(W)]


def shard_arrays(bounds: Tuple[int, int], n_per_problem: int, offsets: Optional[np.ndarray],
                 *per_corr, poses=None):
    """Slices per-correspondence arrays (first dim = correspondences) and poses to one
    rank's shard.  Returns (sliced per_corr..., poses, local_offsets or None)."""
    b0, b1 = bounds
    if offsets is None:
        c0, c1 = b0 * n_per_problem, b1 * n_per_problem
        local_offsets = None
    else:
        c0, c1 = int(offsets[b0]), int(offsets[b1])
        local_offsets = np.asarray(offsets[b0 : b1 + 1], dtype=np.int64) - c0
    out = [None if a is None else a[c0:c1] for a in per_corr]
    p = None if poses is None else poses[b0:b1]
    return (*out, p, local_offsets)


def gather_results(local, bounds: List[Tuple[int, int]], group=None):
    """All-gathers a per-problem tensor (first dim = local problems) from every rank and
    returns the full-batch tensor on every rank.  Shards may differ in size, so each is
    padded to the largest shard (56 B per pose: the volume is negligible)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [b1 - b0 for b0, b1 in bounds]
    assert len(sizes) == world and local.shape[0] == sizes[dist.get_rank(group)]
    m = max(sizes) if sizes else 0
    if m > 0 and all(s == m for s in sizes):
        # equal shards (the usual case): one collective straight into the result, no padding, no concatenation
        out = local.new_empty((world * m,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = local.new_zeros((m,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def solve_sharded(solve_fn: Callable, num_problems: int, n_per_problem: int,
                  offsets: Optional[np.ndarray], f1, f2, ct, ch, init_poses, group=None):
    """Runs `solve_fn(f1, f2, ct, ch, poses, offsets=..., n_per_problem=...)` -> (poses,
    status, iterations) tensors on this rank's shard and gathers the full batch."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bounds = shard_bounds(num_problems, world, offsets)
    lf1, lf2, lct, lch, lposes, loff = shard_arrays(bounds[rank], n_per_problem, offsets,
                                                    f1, f2, ct, ch, poses=init_poses)
    poses, status, iters = solve_fn(lf1, lf2, lct, lch, lposes, offsets=loff,
                                    n_per_problem=n_per_problem)
    return (gather_results(poses, bounds, group), gather_results(status, bounds, group),
            gather_results(iters, bounds, group))
