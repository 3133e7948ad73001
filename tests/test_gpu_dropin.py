"""The drop-in boundary on the B200, through the REFERENCE'S OWN code: the unmodified reference checkout staged
under baseline/_ref/ (tools/stage_reference.py; git-ignored, travels to the GPU box) is imported, `patch_reference()` rebinds
`ConformerEncoder`, `LossCTC` and `ModelCTC.gready_search_decoding`, and the reference's `create_model(config)`, `model.forward(batch)`,
`model.criterion`, `loss.backward()` under `GradScaler`, `optimizer.step()`, `scheduler.step()`, `gready_search_decoding`, strict
`load_state_dict` and `distribute_strategy` (SyncBatchNorm + DistributedDataParallel x 2) run unchanged on top of the CUDA path.
Checked against the reference's own modules run on the same GPU in fp32 (same weights, same batch) and against the CPU oracle."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import stage_reference as SR  # noqa: E402

from efficientconformer_b200.config import CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V  # noqa: E402
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_targets  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"
STAGED = os.path.isdir(os.path.join(SR.DST, "models"))
needs_ref = pytest.mark.skipif(not STAGED, reason="baseline/_ref not staged (python tools/stage_reference.py where /root/reference exists)")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _config(pdrop=0.0, spec_augment=False):
    cfg = json.load(open(os.path.join(SR.DST, "configs", "EfficientConformerCTCSmall.json")))
    cfg["encoder_params"]["Pdrop"] = pdrop
    cfg["encoder_params"]["spec_augment"] = spec_augment
    return cfg


def _reference_and_dropin(cfg, device=DEV):
    """(reference model, drop-in model) built by the reference's own create_model from the same config, same seeded weights."""
    import importlib
    SR.import_reference()
    import models.encoders, models.losses, models.model_ctc     # noqa: E401
    for m in (models.encoders, models.losses, models.model_ctc):
        importlib.reload(m)                                      # undo a previous patch: the unpatched classes first
    import functions
    importlib.reload(functions)
    ref = functions.create_model(cfg)
    import efficientconformer_b200 as ec
    patched = ec.patch_reference()
    assert "models.model_ctc.ConformerEncoder" in patched and "models.model_ctc.LossCTC" in patched
    drop = functions.create_model(cfg)
    assert type(drop.encoder).__module__ == "efficientconformer_b200.encoders" and type(ref.encoder).__module__ == "models.encoders"
    sd = seeded_state_dict(cfg["encoder_params"], cfg["tokenizer_params"]["vocab_size"], seed=0, prefix_encoder="encoder.")
    missing = ref.load_state_dict(sd, strict=False)
    assert all(k.startswith("encoder.preprocessing.") for k in missing.missing_keys) and not missing.unexpected_keys
    print(drop.load_state_dict(ref.state_dict(), strict=True))   # 632 names and shapes identical: strict load of a reference checkpoint
    return ref.to(device), drop.to(device)


def _batch(B=3, seconds=(2.0, 1.4, 0.9), seed=3, device=DEV):
    L = int(16000 * max(seconds))
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(seed))
    x_len = torch.tensor([int(16000 * s) for s in seconds[:B]])
    f_len = ((((x_len // 160 + 1) - 1) // 2 + 1 - 1) // 2 + 1 - 1) // 2 + 1
    y, y_len = synthetic_targets(f_len, V, seed=4)
    return [t.to(device) for t in (x, y, x_len, y_len)]


@needs_ref
def test_reference_main_path_through_the_dropin_eval_and_greedy():
    ref, drop = _reference_and_dropin(_config())
    ref.eval(); drop.eval()
    batch = _batch()
    with torch.no_grad():
        lr, llr, _ = ref.forward(batch)
        ld, lld, att = drop.forward(batch)
    assert torch.equal(llr.cpu(), lld.cpu()) and len(att) == 15
    e = rel_l2(ld, lr)
    print(f"drop-in eval logits vs the reference modules on the same GPU (fp32): rel-L2 {e:.3e}")
    assert e < 1e-3
    loss_r, loss_d = ref.criterion(batch, (lr, llr, None)), drop.criterion(batch, (ld, lld, None))
    assert abs(float(loss_r) - float(loss_d)) < 1e-3 * abs(float(loss_r))
    # greedy decoding: device kernel vs the reference's host loop (no tokenizer file offline: both return token-id lists)
    ref.tokenizer = drop.tokenizer = type("Ids", (), {"decode": staticmethod(lambda ids: ids)})()
    with torch.no_grad():
        ids_r, ids_d = ref.gready_search_decoding(batch[0], batch[2]), drop.gready_search_decoding(batch[0], batch[2])
    top2 = lr.topk(2, dim=-1).values
    if bool(((top2[..., 0] - top2[..., 1]) > 1e-3 * lr.abs().max()).all()):
        assert ids_r == ids_d
    else:                                                       # near-tied random-init frames may flip: compare lengths only
        assert [abs(len(a) - len(b)) <= 2 for a, b in zip(ids_r, ids_d)] == [True] * len(ids_r)


@needs_ref
def test_reference_training_loop_through_the_dropin():
    """The reference's inner loop (models/model.py:239-259) on the drop-in vs on the reference modules, same weights / batch, Pdrop = 0,
    SpecAugment off (the two random parts): loss, every gradient norm, the parameters after optimizer.step(), the schedule."""
    cfg = _config()
    ref, drop = _reference_and_dropin(cfg)
    ref.train(); drop.train()
    batch = _batch()
    stats = {}
    for name, model in (("reference", ref), ("dropin", drop)):
        scaler = torch.cuda.amp.GradScaler(enabled=True)         # as the reference trainer; the scale cancels in scaler.step
        with torch.cuda.amp.autocast(enabled=False):             # fp32 here: the comparison below is at the 1e-3 level
            pred = model.forward(batch)
            loss = model.criterion(batch, pred)
        scaler.scale(loss / 2).backward()                        # accumulated_steps = 2 of the shipped config: two micro-batches
        with torch.cuda.amp.autocast(enabled=False):
            loss2 = model.criterion(batch, model.forward(batch))
        scaler.scale(loss2 / 2).backward()
        inv = 1.0 / scaler.get_scale()
        grads = {k: (p.grad.detach().float() * inv).clone() for k, p in model.named_parameters()}
        scaler.step(model.optimizer); scaler.update(); model.optimizer.zero_grad(); model.scheduler.step()
        stats[name] = (float(loss), grads, {k: v.detach().clone() for k, v in model.state_dict().items()}, model.scheduler.model_step,
                       model.optimizer.param_groups[0]["lr"])
    (l_r, g_r, sd_r, st_r, lr_r), (l_d, g_d, sd_d, st_d, lr_d) = stats["reference"], stats["dropin"]
    assert abs(l_r - l_d) < 1e-3 * abs(l_r), (l_r, l_d)
    assert st_r == st_d == 1 and lr_r == lr_d
    norms = sorted(float(g.norm()) for g in g_r.values())
    floor = 1e-4 * norms[len(norms) // 2]
    worst = (0.0, "")
    for k, g in g_r.items():
        if float(g.norm()) < floor:
            continue
        worst = max(worst, (abs(float(g_d[k].norm()) - float(g.norm())) / float(g.norm()), k))
    print(f"drop-in training: loss {l_d:.5f} (reference {l_r:.5f}), worst gradient-norm error {worst[0]:.3e} ({worst[1]})")
    assert worst[0] < 4e-2, worst                                # attention-core gradients are bf16 grade in the default mode
    for k in sd_r:
        if k.endswith(("running_mean", "running_var")):
            assert rel_l2(sd_d[k], sd_r[k]) < 1e-3, k
        if k.endswith("num_batches_tracked"):
            assert int(sd_d[k]) == int(sd_r[k]) == 2, k


@needs_ref
def test_dropin_train_forward_applies_specaugment_and_autocast_runs():
    cfg = _config(pdrop=0.1, spec_augment=True)
    _, drop = _reference_and_dropin(cfg)
    drop.train()
    batch = _batch()
    # SpecAugment is inside the train-mode forward (reference models/encoders.py:103-104): two forwards differ, eval forwards do not
    seen = []
    enc = drop.encoder
    orig = enc.augment.forward
    enc.augment.forward = lambda x, x_len: seen.append(orig(x, x_len)) or seen[-1]
    # The reference trainer's GradScaler protocol (models/model.py:239-259): its fp16 autocast overflows the reference's own nn.Linear
    # head at the initial scale 65536 (inf in fc.weight.grad), scaler.step() skips that step and update() halves the scale -- the
    # drop-in must behave the same way: encoder gradients stay finite throughout and a step is taken once the scale has settled.
    scaler = torch.cuda.amp.GradScaler()
    enc_names = {id(p) for p in drop.encoder.parameters()}
    taken, scales = False, []
    for it in range(8):
        drop.optimizer.zero_grad()
        with torch.cuda.amp.autocast(enabled=True):              # mixed_precision: true in the shipped config -> bf16 operand mode
            pred = drop.forward(batch)
            loss = drop.criterion(batch, pred)
        scaler.scale(loss).backward()
        assert torch.isfinite(loss)
        finite = all(torch.isfinite(p.grad).all() for p in drop.parameters() if p.grad is not None)
        head_only = all(torch.isfinite(p.grad).all() for p in drop.parameters() if p.grad is not None and id(p) in enc_names)
        assert head_only, "the CUDA path produced a non-finite encoder gradient from finite logit gradients"
        scales.append(scaler.get_scale())
        scaler.step(drop.optimizer); scaler.update()
        if finite:
            taken = True
            break
    print(f"GradScaler scales tried {scales}; step taken: {taken}")
    assert taken, scales
    assert len(seen) == len(scales)                              # exactly one SpecAugment call per train-mode forward
    seen[:] = seen[-1:]
    mel = seen[0]
    B, F, T = mel.shape
    lens = (batch[2] // 160 + 1).tolist()
    for b in range(B):
        valid = mel[b, :, :lens[b]]
        zero_f = (valid == 0).all(dim=1).sum().item()            # <= mF * (F - 1) masked mel bins
        zero_t = (valid == 0).all(dim=0).sum().item()            # <= mT * int(pS * len) masked frames
        assert 0 <= zero_f <= 2 * 27 and 0 <= zero_t <= 5 * int(0.05 * lens[b]) + 1, (b, zero_f, zero_t)
    assert sum(((mel[b, :, :lens[b]] == 0).all(dim=0).sum().item() + (mel[b] == 0).all(dim=1).sum().item()) for b in range(B)) > 0


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cfg = _config()
        _, drop = _reference_and_dropin(cfg, dev)
        drop.train()
        x, y, x_len, y_len = _batch(B=4, seconds=(1.5, 1.5, 1.5, 1.5), device=dev)
        res = {}
        if rank == 0:                                            # the whole batch on one GPU, plain BatchNorm
            loss = drop.criterion([x, y, x_len, y_len], drop.forward([x, y, x_len, y_len]))
            loss.backward()
            res["full"] = (float(loss), {k: p.grad.detach().cpu().numpy().copy() for k, p in drop.named_parameters()},
                           {k: v.cpu().numpy().copy() for k, v in drop.state_dict().items() if "running" in k})
        _, ddp = _reference_and_dropin(cfg, dev)
        ddp.train()
        ddp.distribute_strategy(rank)                            # reference models/model_ctc.py:70-75: SyncBatchNorm + DDP x 2
        assert any(isinstance(m, torch.nn.SyncBatchNorm) for m in ddp.encoder.modules())
        sl = slice(rank * 2, rank * 2 + 2)
        b = [x[sl], y[sl], x_len[sl], y_len[sl]]
        loss = ddp.criterion(b, ddp.forward(b))
        loss.backward()                                          # DDP all-reduces (averages) the gradients
        res["dp"] = (float(loss), {k.replace(".module", ""): p.grad.detach().cpu().numpy().copy() for k, p in ddp.named_parameters()},
                     {k.replace(".module", ""): v.cpu().numpy().copy() for k, v in ddp.state_dict().items() if "running" in k})
        res["rank"] = rank
        q.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@needs_ref
def test_reference_distribute_strategy_two_gpus_matches_single_gpu_whole_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, 29791, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        o = q.get(timeout=600)
        res[o["rank"]] = o
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    lf, gf, sf = res[0]["full"]
    (l0, g0, s0), (l1, g1, s1) = res[0]["dp"], res[1]["dp"]
    assert abs(0.5 * (l0 + l1) - lf) < 1e-3 * abs(lf), (l0, l1, lf)              # SyncBatchNorm: shard losses average to the whole-batch loss
    t = torch.from_numpy
    for k in s0:
        assert rel_l2(t(s0[k]), t(s1[k])) < 1e-6 and rel_l2(t(s0[k]), t(sf[k])) < 1e-3, k
    norms = sorted(float(t(g).norm()) for g in gf.values())
    floor = 1e-4 * norms[len(norms) // 2]
    # parameters whose exact gradient is zero (softmax is invariant to a per-row constant: key / positional biases and the positional
    # columns that multiply the constant cos ~ 1 sinusoids; biases in front of a BatchNorm): their gradients are rounding noise whose
    # value depends on the batch split (same list as tests/test_gpu_training.py)
    noise = ("convolution_module.layers.4.bias", "subsampling_module.layers.0.0.bias", "key_layer.bias", "pos_layer.bias", "pos_layer.weight")
    worst = (0.0, "")
    for k, g in gf.items():
        assert rel_l2(t(g0[k]), t(g1[k])) < 1e-6 or float(t(g).norm()) < floor, k   # DDP: replicas hold the same averaged gradient
        if float(t(g).norm()) >= floor and not k.endswith(noise):
            worst = max(worst, (abs(float(t(g0[k]).norm()) - float(t(g).norm())) / float(t(g).norm()), k))
    print(f"2-rank distribute_strategy vs whole batch: losses {l0:.5f} {l1:.5f} | {lf:.5f}; worst gradient-norm error {worst[0]:.3e} ({worst[1]})")
    assert worst[0] < 4e-2, worst
