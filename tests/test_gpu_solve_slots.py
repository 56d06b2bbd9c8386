"""GPU (-m gpu): solve_slots_kernel (pnec_b200/csrc/pnec_solve_slots.cuh), the refinement kernel of large
batches -- PNECCeres::Optimize / ceres::Solve (src/optimization/pnec_ceres.cc:70-168) with evaluation and
LM update on different warps.

* the default handle takes it for BASELINE config C2 (what bench.py times; test_gpu_parity.py compares
  that call with the oracle on all 10 000 pairs);
* against solve_kernel<V, 4, .> (one CTA per pair), which assigns correspondences to lanes and adds
  partial sums in the same order: bit for bit, all four residual variants, pairs that start at odd
  correspondence indices, ragged batches with empty pairs, fewer pairs than slots;
* two streams on one handle at the same time (every stream has its own work counter).
"""
import os

import numpy as np
import pytest

from pnec_b200 import api
from pnec_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

VARIANTS = {"nec": api.NEC, "target": api.TARGET, "host": api.HOST, "symmetric": api.SYMMETRIC}
FIELDS = ("poses", "status", "iterations", "cost", "initial_cost")


def dev(a):
    import torch

    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def handle_with(**env):
    """The switches are read when a handle is created."""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return api.Handle(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k)
            else:
                os.environ[k] = v


@pytest.fixture(scope="module")
def slots():
    return handle_with(PNEC_B200_SOLVE_SLOTS=2)  # whenever the pairs fit, whatever the batch size


@pytest.fixture(scope="module")
def per_pair():
    return handle_with(PNEC_B200_SOLVE_SLOTS=0, PNEC_B200_SOLVE_WARPS=4)


def args_of(b, variant):
    ct = None if variant == api.NEC else b.covs_target
    ch = b.covs_host if variant == api.SYMMETRIC else None
    return dev(b.bvs_host), dev(b.bvs_target), dev(ct), dev(ch), dev(b.init_poses), api.default_opts(variant)


def assert_same_bits(r0, r1):
    import torch

    for f in FIELDS:
        assert torch.equal(getattr(r0, f), getattr(r1, f)), f


def test_default_selection_takes_the_slots_kernel_on_c2(per_pair):
    """The default handle on a C2 batch: 4 launches (start points, the two kernels that order the pairs
    by the conditioning of their normal equations, slots kernel), the bits of the kernel with a CTA per
    pair -- and of the slots kernel taking the pairs in index order (2 launches).
    (tests/test_gpu_parity.py::test_c2_full_batch_matches_oracle compares this very path
    with the oracle: the default handle takes it for 10 000 x 512.)"""
    import torch

    h = api.Handle(0)
    b = syn.make_batch(10000, 512, seed=1, noise_type="anisotropic_inhomogenous", noise_level=1.0)
    a = args_of(b, api.TARGET)
    n0 = h.launch_count
    res = h.solve_batch(*a, n_per_problem=512)
    torch.cuda.synchronize()
    assert h.launch_count - n0 == 4
    assert_same_bits(per_pair.solve_batch(*a, n_per_problem=512), res)
    plain = handle_with(PNEC_B200_SOLVE_ORDER=0)
    n0 = plain.launch_count
    res_plain = plain.solve_batch(*a, n_per_problem=512)
    torch.cuda.synchronize()
    assert plain.launch_count - n0 == 2
    assert_same_bits(res_plain, res)
    its = res.iterations.cpu().numpy()
    assert its.max() == 50 and 3.0 < its.mean() < 3.6  # the batch of bench.py: a few pairs run into max_num_iterations


@pytest.mark.parametrize("vname", sorted(VARIANTS))
def test_bitwise_against_the_kernel_with_a_cta_per_pair(slots, per_pair, vname):
    b = syn.with_host_covariances(syn.make_batch(1500, 200, seed=7))
    a = args_of(b, VARIANTS[vname])
    assert_same_bits(per_pair.solve_batch(*a, n_per_problem=200), slots.solve_batch(*a, n_per_problem=200))


def test_bitwise_at_bench_shape(slots, per_pair):
    b = syn.make_batch(3000, 512, seed=2024)
    a = args_of(b, api.TARGET)
    assert_same_bits(per_pair.solve_batch(*a, n_per_problem=512), slots.solve_batch(*a, n_per_problem=512))


def test_pairs_starting_at_odd_correspondences_and_an_odd_batch_end(slots, per_pair):
    """N odd: every second pair starts 8 bytes off the 16-byte grid of the bulk copies (it is loaded from
    the even element in front of it and masked), and the very last element of the batch is copied by hand."""
    b = syn.make_batch(1201, 333, seed=5)
    a = args_of(b, api.TARGET)
    assert_same_bits(per_pair.solve_batch(*a, n_per_problem=333), slots.solve_batch(*a, n_per_problem=333))


def test_ragged_batch_with_empty_and_tiny_pairs(slots, per_pair):
    counts = np.clip(syn.kitti_like_counts(900) // 4, 0, 560)
    counts[::37] = 0
    counts[5::41] = 3  # rank deficient: compared like everything else, bit for bit
    counts[7::43] = 1
    b = syn.make_batch(len(counts), 0, seed=6, camera=syn.PINHOLE, counts=counts)
    a = args_of(b, api.TARGET)
    r0 = per_pair.solve_batch(*a, offsets=b.offsets)
    r1 = slots.solve_batch(*a, offsets=b.offsets)
    assert_same_bits(r0, r1)
    st = r1.status.cpu().numpy()
    assert (st[counts == 0] == 7).all()  # PNEC_STATUS_EMPTY


def test_fewer_pairs_than_slots_and_a_single_pair(slots, per_pair):
    for B in (1, 2, 7):
        b = syn.make_batch(B, 512, seed=8 + B)
        a = args_of(b, api.TARGET)
        assert_same_bits(per_pair.solve_batch(*a, n_per_problem=512), slots.solve_batch(*a, n_per_problem=512))


def test_pairs_too_large_for_a_slot_fall_back(slots, per_pair):
    """600 correspondences do not fit twice two slots per SM: the handle takes solve_kernel (1 launch)."""
    import torch

    b = syn.make_batch(300, 600, seed=9)
    a = args_of(b, api.TARGET)
    n0 = slots.launch_count
    r1 = slots.solve_batch(*a, n_per_problem=600)
    torch.cuda.synchronize()
    assert slots.launch_count - n0 == 1
    assert_same_bits(per_pair.solve_batch(*a, n_per_problem=600), r1)


def test_two_streams_on_one_handle(slots, per_pair):
    """Launches on different streams overlap; each stream draws pairs from its own counter."""
    import torch

    b1 = syn.make_batch(2500, 512, seed=31)
    b2 = syn.make_batch(2500, 448, seed=32)
    a1, a2 = args_of(b1, api.TARGET), args_of(b2, api.TARGET)
    want1 = per_pair.solve_batch(*a1, n_per_problem=512)
    want2 = per_pair.solve_batch(*a2, n_per_problem=448)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    got1, got2 = [], []
    for _ in range(3):
        with torch.cuda.stream(s1):
            got1.append(slots.solve_batch(*a1, n_per_problem=512))
        with torch.cuda.stream(s2):
            got2.append(slots.solve_batch(*a2, n_per_problem=448))
    torch.cuda.synchronize()
    for g in got1:
        assert_same_bits(want1, g)
    for g in got2:
        assert_same_bits(want2, g)
