"""Multi-GPU path on the CPU: page-range sharding with world_size 2 over gloo.  The data path has
no collective; the only exchange is the max-over-ranks of the timing scalar, as in bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ocr_system_b200.pipeline import batch_ranges, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 63, 64, 10000, 10001):
        for g in (1, 2, 4, 8):
            covered = []
            for r in range(g):
                lo, hi = shard_range(n, r, g)
                assert 0 <= lo <= hi <= n
                covered += list(range(lo, hi))
            assert covered == list(range(n))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_batch_ranges_cover_the_shard_and_restart_by_page_index():
    """The stream of config 5 is restartable by page index (checkpoint / resume row): the batches of a rank
    partition its shard, and a run resumed at any recorded page index processes exactly the pages that were left."""
    for n, g, b in [(10000, 8, 64), (10001, 4, 64), (63, 2, 64), (0, 2, 64), (130, 1, 64), (5, 8, 2)]:
        for r in range(g):
            lo, hi = shard_range(n, r, g)
            rs = list(batch_ranges(n, r, g, b))
            assert [p for a, z in rs for p in range(a, z)] == list(range(lo, hi))
            assert all(0 < z - a <= b for a, z in rs) and all(z - a == b for a, z in rs[:-1])
            for done in {lo, hi, lo + (hi - lo) // 2, min(hi, lo + b), min(hi, lo + b + 1)}:
                rest = list(batch_ranges(n, r, g, b, resume_from=done))
                assert [p for a, z in rest for p in range(a, z)] == list(range(done, hi))
    with pytest.raises(ValueError):
        list(batch_ranges(100, 1, 2, 64, resume_from=10))      # page 10 belongs to rank 0
    with pytest.raises(ValueError):
        list(batch_ranges(100, 0, 2, 0))


def _worker(rank, world, port, n_pages, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_pages, rank, world)
    # per-rank "work": page indices double as seeds (bench.py: seed0 = rank * batch)
    local = torch.tensor([sum(range(lo, hi)), hi - lo], dtype=torch.float64)
    t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)   # pretend elapsed seconds
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                      # the only collective: timing
    gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, local)
    if rank == 0:
        q.put((float(t.item()), [g.tolist() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_pages = 10001
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pages, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, parts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == pytest.approx(0.020)
    assert sum(p[1] for p in parts) == n_pages
    assert sum(p[0] for p in parts) == sum(range(n_pages))
