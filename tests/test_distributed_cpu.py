"""world_size-2 gloo tests of the N>1 host logic (shards, max-over-ranks timing, id gathering) on CPU."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from efficientconformer_b200.distributed import shard_bounds, shard_batch, max_over_ranks, sum_over_ranks, gather_ragged_ids


def test_shard_bounds_partition():
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mel = torch.arange(10 * 3, dtype=torch.float32).reshape(10, 3)
        lens = torch.arange(10)
        m, l = shard_batch([mel, lens], rank, world)
        frames = float(l.sum())
        t_max = max_over_ranks(1.0 + rank)            # the slowest rank defines the step time
        total = sum_over_ranks(frames)                # whole-job units
        ids = gather_ragged_ids([[rank, 7], [rank] * (rank + 1)])
        q.put((rank, m.shape[0], t_max, total, ids))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29731
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [5, 5]
    assert all(r[2] == 2.0 for r in res)
    assert all(r[3] == float(sum(range(10))) for r in res)
    assert res[0][4] == res[1][4] == [[0, 7], [0], [1, 7], [1, 1]]
