"""CPU tests of the N>1 host logic (gloo, world_size 2): shard bounds, the rebalancing planner and the
reduction layout every rank applies identically."""
import os
import socket

import numpy as np
import pytest


def test_shard_bounds_and_plan_are_consistent():
    from pyvibdmc_b200.distributed import plan_rebalance, shard_bounds
    assert shard_bounds(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sum(c for _, c in shard_bounds(1_000_003, 8)) == 1_000_003
    assert plan_rebalance([1000, 1001, 999, 1000]) == []                      # within tolerance: nothing moves
    pops = [1300, 700, 1000, 1000]
    moves = plan_rebalance(pops)
    after = list(pops)
    for s, d, c in moves:
        assert c > 0 and s != d
        after[s] -= c
        after[d] += c
    assert after == [1000, 1000, 1000, 1000] and sum(after) == sum(pops)
    rng = np.random.default_rng(0)
    for _ in range(50):
        pops = rng.integers(0, 5000, size=8).tolist()
        after = list(pops)
        for s, d, c in plan_rebalance(pops, tolerance=0.0):
            after[s] -= c
            after[d] += c
        assert max(after) - min(after) <= 1 and sum(after) == sum(pops) and min(after) >= 0


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from pyvibdmc_b200 import _capi
    from pyvibdmc_b200.distributed import plan_rebalance
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 1. every rank derives the same plan from the all-gathered populations
    n = torch.tensor([1500 if rank == 0 else 500], dtype=torch.int64)
    gathered = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(gathered, n)
    plan = plan_rebalance([int(t) for t in gathered])
    # 2. the per-step reduction: ranks fill only their own min/max slots, a SUM all-reduce merges everything
    sums = torch.zeros(_capi.NSUMS, dtype=torch.float64)
    C = _capi
    sums[C.SUM_CV:C.SUM_CV + 3] = torch.tensor(C.sums_encode(10.0 * (rank + 1)))
    sums[C.SUM_C:C.SUM_C + 3] = torch.tensor(C.sums_encode(100.0 * (rank + 1)))
    sums[C.SUM_NIN] = 100.0 * (rank + 1)
    sums[C.SUM_EXT + 4 * rank + 0], sums[C.SUM_EXT + 4 * rank + 1] = -1.0 - rank, 2.0 + rank
    dist.all_reduce(sums)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[len(plan)] + [x for m in plan for x in m], sums.numpy()]))
    dist.destroy_process_group()


def test_gloo_world2_ranks_agree(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b)
    assert list(a[:4]) == [1, 0, 1, 500]                    # one move: rank 0 -> rank 1, 500 walkers
    from pyvibdmc_b200 import _capi as C
    sums = a[4:]
    cv, c = C.sums_decode(sums[C.SUM_CV:C.SUM_CV + 3]), C.sums_decode(sums[C.SUM_C:C.SUM_C + 3])
    assert cv == 30.0 and c == 300.0
    e = C.SUM_EXT
    assert sums[e] == -1.0 and sums[e + 1] == 2.0 and sums[e + 4] == -2.0 and sums[e + 5] == 3.0
    vref = cv / c - 0.1 * ((c - 300) / 300)
    assert vref == 0.1


def test_reduction_chunks_sum_exactly_in_any_order():
    """The three chunks of every floating sum add exactly whatever the association (ring, tree, rank order):
    eight ranks' contributions, all 40 320 / sampled orders and pairings give the same three doubles, and they decode to the
    correctly rounded exact total."""
    import itertools
    from fractions import Fraction
    from pyvibdmc_b200 import _capi as C
    rng = np.random.default_rng(5)
    for trial in range(20):
        xs = rng.normal(0, 1, 8) * 10.0 ** rng.integers(-3, 6)
        xs[rng.integers(0, 8)] *= -1
        enc = np.array([C.sums_encode(x) for x in xs])
        ref = None
        for perm in itertools.islice(itertools.permutations(range(8)), 0, 40320, 997):
            tot = np.zeros(3)
            for r in perm:                                  # left-to-right in this order
                tot = tot + enc[r]
            pair = (enc[list(perm[:4])].sum(0)) + (enc[list(perm[4:])].sum(0))   # a tree
            if ref is None:
                ref = tot.copy()
            assert np.array_equal(tot, ref) and np.array_equal(pair, ref)
        exact = sum(Fraction(float(x)) for x in xs)
        assert C.sums_decode(ref) == float(exact)
