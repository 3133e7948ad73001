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


def _syncbn_worker(rank, world, port, q):
    """SyncBatchNorm exchange logic over gloo (the merge kernel is swapped for its torch-CPU restatement: no GPU here)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_shim
    from efficientconformer_b200 import distributed as D
    D._ops = cpu_ops_shim
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        full = 3.0 + 2.0 * torch.randn(70, 6, generator=g)              # 70 frames x 6 channels, |mean| > std
        out = {}
        for uniform, bounds in ((True, [(0, 35), (35, 70)]), (False, [(0, 50), (50, 70)])):
            lo, hi = bounds[rank]
            x = full[lo:hi]
            stats = torch.stack([x.mean(0), ((x - x.mean(0)) ** 2).sum(0)])
            red = D.make_sync_bn_reducer(None, "cpu", uniform=uniform)        # no peer memory on CPU: every rank agrees on the NCCL / gloo reducer
            assert type(red) is D.SyncBatchNormReducer
            n = red.forward_stats(stats, float(hi - lo))
            sums = torch.stack([x.sum(0), (x * x).sum(0)])
            red.backward_sums(sums)
            out[uniform] = (n, stats.numpy().copy(), sums.numpy().copy())      # by value: the worker may exit before the parent reads
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_sync_batchnorm_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_syncbn_worker, args=(r, 2, 29741, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    full = 3.0 + 2.0 * torch.randn(70, 6, generator=g)
    ref = torch.stack([full.mean(0), ((full - full.mean(0)) ** 2).sum(0)])
    ref_sums = torch.stack([full.sum(0), (full * full).sum(0)])
    for rank in (0, 1):
        for uniform in (True, False):
            n, stats, sums = res[rank][uniform]
            stats, sums = torch.from_numpy(stats), torch.from_numpy(sums)
            assert n == 70.0
            assert torch.allclose(stats, ref, rtol=1e-5, atol=1e-5), (rank, uniform)
            assert torch.allclose(sums, ref_sums, rtol=1e-5)


def _dp_params():
    from efficientconformer_b200.config import CTC_SMALL_ENCODER_PARAMS as P
    p = dict(P)
    p.update(Pdrop=0.0, num_blocks=3, strided_blocks=[1], expand_blocks=[1], dim_model=[24, 32], att_group_size=[3, 1], subsampling_filters=[8],
             num_heads=4, kernel_size=7)
    return p


def _dp_run(params, mel, y, yl, reducer):
    """Train-mode forward + CTC loss + backward of the training schedule (efficientconformer_b200/training.py) on the torch-CPU operator
    table; returns (loss, gradients, running statistics)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_shim
    from efficientconformer_b200 import training
    from efficientconformer_b200.model_ctc import ModelCTC
    from efficientconformer_b200.synthetic import seeded_state_dict
    from oracle import conformer_oracle as O
    training._ops = cpu_ops_shim
    V2 = 32
    model = ModelCTC(params, {"vocab_size": V2})
    model.load_state_dict(seeded_state_dict(params, V2, seed=0, prefix_encoder="encoder."), strict=False)
    model.train()
    path = training.TrainingPath(model.encoder, model.fc, stats_reducer=reducer)
    with torch.no_grad():
        x, logits, out_len, tape = path.forward(mel, None, "tf32", want_logits=True)
    lg = logits.detach().double().requires_grad_(True)
    B, T_out = lg.shape[0], lg.shape[1]
    loss, _ = O.ctc_loss(lg, torch.full((B,), T_out), y, yl)
    loss.backward()
    with torch.no_grad():
        grads = path.backward(tape, None, lg.grad)
    stats = {k: v.clone() for k, v in model.state_dict().items() if "running" in k}
    return float(loss.detach()), {k: v.double().clone() for k, v in grads.items()}, stats


def _dp_batch():
    from efficientconformer_b200.synthetic import synthetic_mel, synthetic_targets
    mel = synthetic_mel(4, 97, seed=21)
    y, yl = synthetic_targets(torch.full((4,), 25), 32, seed=4)
    return mel, y, yl


def _dp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import cpu_ops_shim
        from efficientconformer_b200 import distributed as D
        D._ops = cpu_ops_shim
        mel, y, yl = _dp_batch()
        lo, hi = rank * 2, rank * 2 + 2
        loss, grads, stats = _dp_run(_dp_params(), mel[lo:hi], y[lo:hi], yl[lo:hi], D.SyncBatchNormReducer(None, "cpu", uniform=True))
        flat = torch.cat([grads[k].reshape(-1) for k in sorted(grads)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)                  # the one gradient bucket; the mean (1 / world) is folded into Adam
        q.put((rank, loss, (flat / world).numpy().copy(), {k: v.numpy().copy() for k, v in stats.items()}))   # by value
    finally:
        dist.destroy_process_group()


def test_two_rank_data_parallel_training_step_equals_whole_batch():
    """The N > 1 training path end to end over gloo (world size 2, CPU operator table): utterance shards + SyncBatchNorm exchange
    between the conv stages + one all-reduced gradient bucket reproduce the single-process step on the whole batch -- losses,
    every parameter gradient and the BatchNorm running statistics."""
    torch.set_num_threads(2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, 29751, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = q.get(timeout=300)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mel, y, yl = _dp_batch()
    loss, grads, stats = _dp_run(_dp_params(), mel, y, yl, None)
    ref = torch.cat([grads[k].reshape(-1) for k in sorted(grads)])
    assert abs(0.5 * (res[0][1] + res[1][1]) - loss) < 1e-6 * abs(loss)
    g0, g1 = torch.from_numpy(res[0][2]), torch.from_numpy(res[1][2])
    assert torch.equal(g0, g1)                                            # every rank ends with the same averaged gradient
    err = float((g0 - ref).norm() / ref.norm())
    assert err < 1e-5, err                                                # fp32 statistics exchange bounds the agreement
    for k, v in stats.items():
        for r in (0, 1):
            assert torch.allclose(torch.from_numpy(res[r][3][k]), v, rtol=1e-5, atol=1e-6), (k, r)
