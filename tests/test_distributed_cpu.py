"""CPU, world_size 2, gloo: host-side logic of the multi-GPU path (row
sharding, ragged gathers, joint-mode all-reduce, max-over-ranks timing)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, results):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from babe_b200 import distributed as bd
    r, l, w = bd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = bd.shard_rows(5, rank, world)
    full = torch.arange(5 * 3, dtype=torch.float32).reshape(5, 3)
    got = bd.gather_rows(full[lo:hi].clone())
    assert torch.equal(got, full)
    p = torch.full((2, 4), float(rank))
    ps = bd.gather_params(p)
    assert ps.shape == (world, 2, 4) and float(ps[1, 0, 0]) == 1.0
    stats = torch.ones(3, 7, dtype=torch.float64) * (rank + 1)
    bd.sum_over_ranks_(stats)
    assert float(stats[0, 0]) == 3.0
    assert bd.max_over_ranks(10.0 + rank, "cpu") == 11.0
    bd.barrier()
    dist.destroy_process_group()
    results.put(rank)


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]


def test_shard_rows_cover():
    from babe_b200.distributed import shard_rows
    for n in (1, 7, 8, 64):
        for w in (1, 2, 4, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
