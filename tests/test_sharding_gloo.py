"""CPU: the multi-GPU path is sharding with no collective on the data path
(SURVEY.md 8e).  World-size-2 gloo run of the host logic bench.py uses: each
rank takes a contiguous shard, generates exactly its slice of the seeded batch,
solves it (here with the oracle standing in for the GPU) and rank 0 gathers."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from embedded_qp_solver_b200 import problems  # noqa: E402


def test_shard_bounds_cover_batch_without_overlap():
    for count in (0, 1, 7, 32, 1000003):
        for world in (1, 2, 3, 8):
            edges = [problems.shard_bounds(count, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == count
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_shards_regenerate_the_same_problems():
    P, q, _, _ = problems.wellcond(10, 4, seed=5)
    lo, hi = problems.shard_bounds(10, 2, 1)
    P1, q1, _, _ = problems.wellcond(hi - lo, 4, seed=5, first=lo)
    assert np.array_equal(P[lo:hi], P1) and np.array_equal(q[lo:hi], q1)


def _worker(rank, world, port, count, n, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    lo, hi = problems.shard_bounds(count, world, rank)
    P, q, lb, ub = problems.wellcond(hi - lo, n, seed=77, first=lo)
    o = oracle.Oracle().admm(P.astype(np.float64), q.astype(np.float64), lb, ub)
    x = torch.from_numpy(o["x"])
    iters = torch.from_numpy(o["iters"].astype(np.int64))
    # host-side gather of results: the only exchange there is
    sizes = [problems.shard_bounds(count, world, r) for r in range(world)]
    big = max(b - a for a, b in sizes)  # ragged shards: pad to the largest, trim on rank 0
    xpad = torch.zeros((big, n), dtype=torch.float64)
    xpad[: hi - lo] = x
    ipad = torch.zeros(big, dtype=torch.int64)
    ipad[: hi - lo] = iters
    xs = [torch.empty((big, n), dtype=torch.float64) for _ in sizes] if rank == 0 else None
    its = [torch.empty(big, dtype=torch.int64) for _ in sizes] if rank == 0 else None
    dist.gather(xpad, xs, dst=0)
    dist.gather(ipad, its, dst=0)
    if rank == 0:
        xs = [t[: b - a] for t, (a, b) in zip(xs, sizes)]
        its = [t[: b - a] for t, (a, b) in zip(its, sizes)]
    # max-over-ranks timing reduction as in bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    if rank == 0:
        np.save(os.path.join(tmp, "x.npy"), torch.cat(xs).numpy())
        np.save(os.path.join(tmp, "it.npy"), torch.cat(its).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_shard_and_gather(tmp_path):
    import oracle
    oracle.build(ref=False)
    count, n, world = 37, 8, 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, count, n, str(tmp_path)), nprocs=world, join=True)
    P, q, lb, ub = problems.wellcond(count, n, seed=77)
    o = oracle.Oracle().admm(P.astype(np.float64), q.astype(np.float64), lb, ub)
    assert np.array_equal(np.load(tmp_path / "x.npy"), o["x"])
    assert np.array_equal(np.load(tmp_path / "it.npy"), o["iters"])
