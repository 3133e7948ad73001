"""The ASSEMBLED training step on the B200 (BASELINE.json configs[1]) through the drop-in API:
`ModelCTC(...).train()`, `forward`, `LossCTC`, `loss.backward()` -- against the REAL reference's loss.backward() (golden gradients
of tests/golden/ctc_small_train_b2_t500.pt, produced by tests/golden/make_golden_train.py from /root/reference), plus the pieces
that close the step: counter-based dropout, flat-arena Adam with the Transformer schedule (against torch.optim.Adam, the
optimiser the reference constructs, models/model.py:88-93) and the graph-captured CTCTrainStep."""
import math
import os

import pytest
import torch

from efficientconformer_b200.config import CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel, synthetic_targets

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _model(precision, pdrop=0.0, params=None, vocab=V, seed=0):
    from efficientconformer_b200.model_ctc import ModelCTC
    p = dict(params or P); p["Pdrop"] = pdrop
    model = ModelCTC(p, {"vocab_size": vocab}, precision=precision)
    model.load_state_dict(seeded_state_dict(params or P, vocab, seed=seed, prefix_encoder="encoder."), strict=False)
    return model.to(DEV).train()


# gradient tolerances (relative L2 per tensor / relative error of the gradient norm).  The reference is fp32; TF32 operand rounding
# gives 7e-4 on the logits and accumulates through the 15-block backward; bf16 operands (the reference's own mixed-precision
# training uses fp16 autocast) are one decimal looser.
# Measured on the B200: tf32 logits 1.2e-3 / loss 8e-6 / worst gradient-norm error 9e-4 / worst tensor 2.1e-3;
# bf16 1.0e-2 / 1e-5 / 1.3e-2 / 1.6e-2.  The gates below are ~3x those.
# bf16x2 (the default split mode): forward operands carry 16 significant bits -> logits / loss must meet the north-star 1e-3;
# the attention core (forward TF32, backward bf16 operands) bounds the gradient accuracy.
TOL = {"tf32": dict(logits=2e-3, loss=1e-3, norm=3e-3, full=6e-3, stats=2e-4),
       "bf16x2": dict(logits=1e-3, loss=1e-3, norm=4e-2, full=5e-2, stats=2e-4),
       "bf16": dict(logits=1.5e-2, loss=1e-3, norm=4e-2, full=5e-2, stats=2e-3)}


@pytest.mark.parametrize("prec", ["bf16x2", "tf32", "bf16"])
def test_training_step_matches_reference_backward(prec, golden_dir):
    g = torch.load(os.path.join(golden_dir, "ctc_small_train_b2_t500.pt"))
    tol = TOL[prec]
    model = _model(prec)
    mel = synthetic_mel(2, 500, seed=g["mel_seed"]).to(DEV)
    logits, out_len, _ = model.forward_mel(mel, g["mel_len"].to(DEV))
    assert logits.requires_grad and logits.shape == g["logits"].shape
    assert out_len.tolist() == [63, 48]
    e_logits = rel_l2(logits.detach(), g["logits"])
    loss = model.criterion((None, g["targets"].to(DEV), None, g["target_len"].to(DEV)), (logits, out_len, None))
    e_loss = abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters()}
    assert set(grads) == set(g["grad_norms"])
    floor = 1e-4 * sorted(g["grad_norms"].values())[len(g["grad_norms"]) // 2]
    errs = []
    for k, ref_norm in g["grad_norms"].items():
        assert grads[k] is not None and grads[k].shape == dict(model.named_parameters())[k].shape, k
        gn = float(grads[k].double().norm())
        assert math.isfinite(gn), k
        if ref_norm < floor:
            # exactly zero in exact arithmetic (biases in front of BatchNorm, key / positional biases): rounding noise only
            assert gn < 50 * floor, (k, gn, floor)
            continue
        errs.append((abs(gn - ref_norm) / ref_norm, k))
    errs.sort(reverse=True)
    full = sorted(((rel_l2(grads[k], ref), k) for k, ref in g["grads"].items() if g["grad_norms"][k] >= floor), reverse=True)
    sd = model.state_dict()
    e_stats = max(rel_l2(sd[k], ref) for k, ref in g["running_stats"].items())
    print(f"\n[{prec}] logits {e_logits:.3e} loss {e_loss:.3e} worst grad-norm err {errs[0][0]:.3e} ({errs[0][1]}) "
          f"median {errs[len(errs) // 2][0]:.3e}; worst full-tensor rel-L2 {full[0][0]:.3e} ({full[0][1]}); running stats {e_stats:.3e}")
    assert e_logits < tol["logits"], e_logits
    assert e_loss < tol["loss"], e_loss
    assert errs[0][0] < tol["norm"], errs[:10]
    assert full[0][0] < tol["full"], full[:10]
    assert e_stats < tol["stats"], e_stats
    assert int(sd["encoder.blocks.3.convolution_module.layers.5.num_batches_tracked"]) == 1


def test_dropout_kernels_statistics_and_mask_consistency():
    from efficientconformer_b200 import ops
    from efficientconformer_b200.training import DropoutState
    drop = DropoutState(0.1, DEV, seed=7)
    drop.begin_step()
    n = 1 << 20
    x = torch.ones(n, device=DEV)
    y = ops.dropout_f32(x, drop, 3)
    keep = (y != 0)
    frac = float(keep.float().mean())
    assert abs(frac - 0.9) < 3e-3, frac
    assert torch.allclose(y[keep], torch.full_like(y[keep], 65536.0 / round(0.9 * 65536)))
    assert abs(float(y.mean()) - 1.0) < 4e-3                                  # unbiased
    # the same (step, site) draws the same mask in every variant: forward on activations, gradient re-masking, fused residual
    for prec in ("tf32", "bf16", "bf16x2"):
        ya = ops.unpack(ops.dropout_act(ops.cast(x, prec), drop, 3, prec), prec)
        assert torch.equal(ya != 0, keep)
        yg = ops.unpack(ops.dropout_cast_scaled(x, prec, 0.5, drop, 3), prec)
        assert torch.equal(yg != 0, keep)
        assert torch.allclose(yg, 0.5 * y, rtol=1e-2)
    r = torch.randn(n, device=DEV)
    yr = ops.dropout_residual(x, drop, 3, 0.5, r)
    assert torch.allclose(yr, r + 0.5 * y, rtol=1e-6, atol=1e-6)
    # other site / next step: independent masks
    y2 = ops.dropout_f32(x, drop, 4)
    agree = float(((y2 != 0) == keep).float().mean())
    assert abs(agree - (0.81 + 0.01)) < 5e-3, agree
    drop.begin_step()
    y3 = ops.dropout_f32(x, drop, 3)
    assert abs(float(((y3 != 0) == keep).float().mean()) - 0.82) < 5e-3
    # odd length (tail group) and no low-order structure along rows of width 120
    z = ops.dropout_f32(torch.ones(1001 * 120 + 3, device=DEV), drop, 9)
    assert z.shape[0] == 1001 * 120 + 3 and abs(float((z != 0).float().mean()) - 0.9) < 5e-3
    cols = (z[:1001 * 120].view(1001, 120) != 0).float().mean(0)
    assert float(cols.min()) > 0.85 and float(cols.max()) < 0.95


def test_fused_swish_dropout_and_arena_operand_kernels():
    """Swish fused with its dropout (forward and backward) equals the two-kernel sequence with the same mask; the one-launch
    arena cast / multi-tensor transposed cast equal the per-tensor operators."""
    from efficientconformer_b200 import ops, trainer
    from efficientconformer_b200.training import DropoutState, TrainingPath
    drop = DropoutState(0.1, DEV, seed=3)
    drop.begin_step()
    g = torch.Generator().manual_seed(8)
    z32, dy = 2 * torch.randn(4001, 480, generator=g).to(DEV), torch.randn(4001, 480, generator=g).to(DEV)
    for prec in ("tf32", "bf16", "bf16x2"):
        z = ops.cast(z32, prec)
        fused = ops.unpack(ops.swish_dropout_fwd(z, drop, 5, prec), prec)
        ref = ops.unpack(ops.dropout_act(ops.swish_fwd(z, prec), drop, 5, prec), prec)
        assert torch.equal(fused == 0, ref == 0)
        assert rel_l2(fused, ref) < (1e-3 if prec != "bf16" else 6e-3)
        fb = ops.unpack(ops.swish_dropout_bwd(z, dy, drop, 5, prec), prec)
        rb = ops.unpack(ops.swish_bwd(z, ops.dropout_f32(dy, drop, 5), prec), prec)
        assert torch.equal(fb == 0, rb == 0)
        assert rel_l2(fb, rb) < (1e-3 if prec != "bf16" else 6e-3)
    # arena operands
    model = _model("bf16", 0.0, _small_params(), vocab=32)
    path = TrainingPath(model.encoder, model.fc)
    flat = trainer.FlatParams(trainer._qkv_adjacent_order(path.param_list()), DEV)
    for prec in ("tf32", "bf16", "bf16x2"):
        w = trainer.ArenaWeights(flat, prec, DEV)
        assert w.supports(model.encoder)
        w.refresh()
        blk = model.encoder.blocks[1]
        for weight in (blk.feed_forward_module1.layers[1].weight, blk.convolution_module.layers[2].weight, blk.conv_res[1].weight,
                       model.encoder.linear.weight, model.fc.weight, blk.multi_head_self_attention_module.mhsa.output_layer.weight):
            w2 = weight.detach().reshape(weight.shape[0], -1)
            for mine, ref in ((w.act(weight), ops.cast_weight(w2, prec)), (w.act_t(weight), ops.transpose_cast(w2, prec))):
                assert torch.equal(mine, ref)
                if prec == "bf16x2":      # the swapped plane sits right behind the operand in both layouts
                    twin = lambda t: torch.as_strided(t, t.shape, t.stride(), t.storage_offset() + t.numel())
                    assert torch.equal(twin(mine), twin(ref))
        mh = blk.multi_head_self_attention_module.mhsa
        w32, b = ops.concat_qkv(mh)
        assert torch.equal(w.qkv_act(mh), ops.cast_weight(w32, prec)) and torch.equal(w.qkv_act_t(mh), ops.transpose_cast(w32, prec))
        assert torch.equal(w.qkv_bias(mh), b)


def test_flat_adam_matches_torch_adam_with_transformer_schedule():
    """ec_adam_step over a flat arena == torch.optim.Adam driven by the reference's Transformer schedule (models/schedules.py:99-123;
    compile() ends with scheduler.step(), models/model.py:150, so the n-th optimiser step runs with lr(s = n)) on the same
    gradients (fp64 CPU reference)."""
    from efficientconformer_b200 import ops
    g = torch.Generator().manual_seed(3)
    n = 100003
    p0 = torch.randn(n, generator=g)
    tp = dict(beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-6, K=2.0, schedule_dim=240.0, warmup_steps=5.0)
    ref_p = p0.double().clone().requires_grad_(True)
    lr_of = lambda s_: 2.0 * 240.0 ** -0.5 * min(s_ ** -0.5, s_ * 5.0 ** -1.5)
    opt = torch.optim.Adam([ref_p], lr=lr_of(1), betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)      # "Init LR": scheduler.step() in compile()
    pad = (n + 63) // 64 * 64
    p = torch.zeros(pad, device=DEV); p[:n] = p0.to(DEV)
    m, v = torch.zeros(pad, device=DEV), torch.zeros(pad, device=DEV)
    state = torch.zeros(4, dtype=torch.int32, device=DEV)
    state[0:1].view(torch.float32).fill_(lr_of(1)); state[2] = 1             # what CTCTrainStep._set_schedule_step(0) writes
    world = 2.0
    model_step = 0
    for it in range(8):
        grad = torch.randn(n, generator=g) * (1.0 + it)
        ref_p.grad = grad.double().clone()
        opt.step()
        model_step += 1; s = model_step + 1                                   # scheduler.step() AFTER optimizer.step()
        opt.param_groups[0]["lr"] = lr_of(s)
        gd = torch.zeros(pad, device=DEV); gd[:n] = (grad * world).to(DEV)    # a SUM all-reduce over 2 ranks; mean folded into Adam
        ops.adam_step(p[:n], gd[:n], m[:n], v[:n], state, tp["beta1"], tp["beta2"], tp["eps"], tp["weight_decay"], grad_scale=1.0 / world,
                      schedule=1, K=tp["K"], dim=tp["schedule_dim"], warmup=tp["warmup_steps"])
        lr_dev = float(state[0:1].view(torch.float32).item())
        assert abs(lr_dev - opt.param_groups[0]["lr"]) < 1e-6 * opt.param_groups[0]["lr"], (it, lr_dev)
        assert int(state[1]) == it + 1 and int(state[2]) == it + 2
        if it == 0:
            assert not torch.equal(p[:n].cpu(), p0)                           # the first optimiser step already moves the parameters
    assert rel_l2(p[:n], ref_p.detach()) < 1e-6
    assert float((p[:n].cpu().double() - ref_p.detach()).abs().max()) < 1e-5


def _small_params():
    p = dict(P)
    p.update(num_blocks=3, strided_blocks=[1], expand_blocks=[1], dim_model=[64, 96], att_group_size=[3, 1], subsampling_filters=[16],
             num_heads=4, kernel_size=15)
    return p


@pytest.mark.parametrize("prec", ["bf16x2", "tf32", "bf16"])
def test_train_step_graph_replay_equals_autograd_plus_torch_adam(prec):
    """CTCTrainStep (flat arenas, packed gradient bucket, device-side Adam + schedule, CUDA-graph replay) against the drop-in route
    the reference trainer takes: forward -> LossCTC -> loss.backward() -> torch.optim.Adam.step() -> scheduler.step().
    Adam turns every gradient element into a step of about +-lr whatever its magnitude, so elements whose gradient is rounding
    noise (biases in front of BatchNorm, key / positional biases, the positional-weight columns that multiply the constant cos ~ 1
    sinusoid columns) move in implementation-dependent directions; the comparison therefore uses the losses, the first / second
    moment estimates (linear / quadratic in the gradients) and the update direction of the well-conditioned tensors."""
    from efficientconformer_b200.trainer import CTCTrainStep
    sp = _small_params()
    tp = dict(optimizer="Adam", beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-6, lr_schedule="Transformer", schedule_dim=96,
              warmup_steps=300, K=2)
    lr_of = lambda s: 2 * 96 ** -0.5 * min(s ** -0.5, s * 300 ** -1.5)
    B, T = 3, 161
    mels = [synthetic_mel(B, T, seed=40 + i).to(DEV) for i in range(4)]
    mel_len = torch.tensor([161, 120, 77], device=DEV)
    out_len = ((((mel_len - 1) // 2 + 1) - 1) // 2 + 1)
    y, yl = synthetic_targets(out_len.cpu(), 32, seed=4)
    y, yl = y.to(DEV), yl.to(DEV)
    # route A: autograd node + torch optimiser
    a = _model(prec, 0.0, sp, vocab=32)
    init = {k: v.detach().clone() for k, v in a.state_dict().items()}
    opt = torch.optim.Adam(a.parameters(), lr=lr_of(1), betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    losses_a, model_step = [], 0
    for mel in mels:
        logits, ol, _ = a.forward_mel(mel, mel_len)
        loss = a.criterion((None, y, None, yl), (logits, ol, None))
        loss.backward()
        opt.step(); opt.zero_grad()
        model_step += 1
        opt.param_groups[0]["lr"] = lr_of(model_step + 1)
        losses_a.append(float(loss))
    # route B: native step, graph replay; route C: native step, eager
    results = {}
    for graph in (True, False):
        b = _model(prec, 0.0, sp, vocab=32)
        step = CTCTrainStep(b, tp, precision=prec, use_cuda_graph=graph)
        losses_b = [float(step.step(mel, mel_len, y, yl)) for mel in mels]
        assert step.steps_done() == len(mels)
        assert abs(step.lr() - lr_of(len(mels) + 1)) < 1e-6 * lr_of(len(mels) + 1)
        moments = {n: (step.flat.exp_avg[o:o + s].clone(), step.flat.exp_avg_sq[o:o + s].clone())
                   for n, o, s in zip(step.flat.names, step.flat.offsets, step.flat.sizes)}
        results[graph] = (losses_b, {k: v.detach().clone() for k, v in b.state_dict().items()}, moments)
    assert results[True][0] == results[False][0]                              # replay == eager, bit for bit
    for k, v in results[True][1].items():
        assert torch.equal(v, results[False][1][k]), k
    for la, lb in zip(losses_a, results[True][0]):
        assert abs(la - lb) < 1e-3 * abs(la), (losses_a, results[True][0])
    sd_a, sd_b, mom = a.state_dict(), results[True][1], results[True][2]
    noise = ("convolution_module.layers.4.bias", "subsampling_module.layers.0.0.bias", "key_layer.bias", "pos_layer.bias", "pos_layer.weight")
    worst_m, worst_v, worst_dir = (0.0, ""), (0.0, ""), 1.0
    table = []
    for n, p in a.named_parameters():
        if n.endswith(noise):
            continue
        st = opt.state[p]
        em, ev = rel_l2(mom[n][0], st["exp_avg"].reshape(-1)), rel_l2(mom[n][1], st["exp_avg_sq"].reshape(-1))
        worst_m = max(worst_m, (em, n))
        worst_v = max(worst_v, (ev, n))
        da, db = (sd_a[n] - init[n]).double().reshape(-1), (sd_b[n] - init[n]).double().reshape(-1)
        assert float(da.norm()) > 0 and float(db.norm()) > 0, n
        cos = float(da @ db / (da.norm() * db.norm()))
        worst_dir = min(worst_dir, cos)
        table.append((em, ev, cos, n))
    for row in sorted(table, reverse=True)[:6]:
        print("   moment rel-L2 %.2e / %.2e  update cosine %.4f  %s" % row)
    print(f"\n[{prec}] losses {results[True][0]} | autograd+torch.optim {losses_a} | moments rel-L2 {worst_m} / {worst_v}, "
          f"worst update cosine {worst_dir:.4f}")
    # bf16x2: the attention backward rounds its operands to bf16, which amplifies the round-off difference between the two Adam
    # implementations on the attention parameters (5e-3 measured on mhsa.v; everything else < 1e-3)
    # bf16: the two routes stay bit-identical until the first parameter that the two Adam implementations round differently flips a
    # bf16 operand (seen at step 3-4 of 4: losses identical for three steps, moments 6e-3 afterwards)
    gate = 1e-2 if prec in ("bf16x2", "bf16") else 2e-3
    assert worst_m[0] < gate and worst_v[0] < gate, (worst_m, worst_v)
    assert worst_dir > 0.98, worst_dir
    for k in sd_a:
        if k.endswith("num_batches_tracked"):
            assert int(sd_a[k]) == int(sd_b[k]) == len(mels), k
        if k.endswith(("running_mean", "running_var")):
            assert rel_l2(sd_b[k], sd_a[k]) < 6e-4, k      # the parameters move from the first step on, in two Adam implementations (measured 2-3e-4)


def test_training_with_dropout_is_reproducible_and_finite():
    from efficientconformer_b200.trainer import CTCTrainStep
    sp = _small_params()
    tp = dict(optimizer="Adam", beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-6, lr_schedule="Transformer", schedule_dim=96,
              warmup_steps=3, K=2)
    mel = synthetic_mel(4, 200, seed=9).to(DEV)
    y, yl = synthetic_targets(torch.full((4,), 50), 32, seed=4)
    runs = []
    for seed in (1, 1, 2):
        m = _model("bf16", 0.1, sp, vocab=32)
        step = CTCTrainStep(m, tp, precision="bf16", use_cuda_graph=True, dropout_seed=seed)
        runs.append([float(step.step(mel, None, y.to(DEV), yl.to(DEV))) for _ in range(4)])
    assert all(math.isfinite(v) for r in runs for v in r)
    assert runs[0] == runs[1]                                                 # same seed: same masks, forward and backward
    assert runs[0] != runs[2]
    assert len(set(runs[0])) == 4                                             # fresh masks every replay + parameters moving


@pytest.mark.parametrize("graph", [True, False])
def test_eval_after_native_training_sees_the_updated_weights(graph):
    """CTCTrainStep writes parameters and BatchNorm running statistics through raw pointers (no tensor._version bump);
    train -> eval -> train -> eval must equal a FRESH model loaded from the trained state_dict, in graph and eager mode."""
    from efficientconformer_b200.trainer import CTCTrainStep
    from efficientconformer_b200.model_ctc import ModelCTC
    sp = _small_params()
    tp = dict(optimizer="Adam", beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-6, lr_schedule="Transformer", schedule_dim=96,
              warmup_steps=3, K=2)                                           # short warm-up: the weights move visibly within a few steps
    mel = synthetic_mel(3, 161, seed=9).to(DEV)
    mel_len = torch.tensor([161, 120, 77], device=DEV)
    out_len = ((((mel_len - 1) // 2 + 1) - 1) // 2 + 1)
    y, yl = synthetic_targets(out_len.cpu(), 32, seed=4)
    m = _model("bf16x2", 0.0, sp, vocab=32)
    step = CTCTrainStep(m, tp, precision="bf16x2", use_cuda_graph=graph)
    seen = []
    for round_ in range(2):
        for _ in range(3):
            step.step(mel, mel_len, y.to(DEV), yl.to(DEV))
        m.eval()
        with torch.no_grad():
            logits = m.forward_mel(mel, mel_len)[0].clone()
        p = dict(sp); p["Pdrop"] = 0.0
        fresh = ModelCTC(p, {"vocab_size": 32}, precision="bf16x2")
        fresh.load_state_dict({k: v.detach().clone() for k, v in m.state_dict().items()}, strict=True)
        with torch.no_grad():
            ref = fresh.to(DEV).eval().forward_mel(mel, mel_len)[0]
        assert torch.equal(logits, ref), (graph, round_, rel_l2(logits, ref))
        seen.append(logits)
        m.train()
    assert rel_l2(seen[1], seen[0]) > 1e-3                                   # ... and the second evaluation is not the first one's weights


def test_optimizer_checkpoint_round_trip_and_accumulated_steps():
    """CTCTrainStep.state_dict() carries a torch.optim.Adam-layout optimizer state, the schedule step and the
    dropout counter (the reference saves optimizer.state_dict() and scheduler.model_step, models/model.py:346-376); a run resumed from
    it continues bit for bit.  accumulated_steps = 2 (the shipped config): parameters move on every second call only, and two
    accumulated half-batches give the gradient step of ... the same two micro-batches (loss / 2 each, reference models/model.py:245)."""
    from efficientconformer_b200.trainer import CTCTrainStep
    sp = _small_params()
    tp = dict(optimizer="Adam", beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-6, lr_schedule="Transformer", schedule_dim=96,
              warmup_steps=50, K=2)
    mels = [synthetic_mel(3, 161, seed=70 + i).to(DEV) for i in range(6)]
    y, yl = synthetic_targets(torch.full((3,), 41), 32, seed=4)
    y, yl = y.to(DEV), yl.to(DEV)
    a = _model("bf16x2", 0.1, sp, vocab=32)
    sa = CTCTrainStep(a, tp, precision="bf16x2", dropout_seed=3)
    for mel in mels[:3]:
        sa.step(mel, None, y, yl)
    ckpt_opt = sa.state_dict()
    ckpt_model = {k: v.detach().clone() for k, v in a.state_dict().items()}
    assert ckpt_opt["model_step"] == 3 and len(ckpt_opt["optimizer"]["state"]) == len(list(a.parameters()))
    tail_a = [float(sa.step(mel, None, y, yl)) for mel in mels[3:]]
    b = _model("bf16x2", 0.1, sp, vocab=32)
    b.load_state_dict(ckpt_model, strict=True)
    sb = CTCTrainStep(b, tp, precision="bf16x2", dropout_seed=3)
    sb.load_state_dict(ckpt_opt)
    assert sb.steps_done() == 3 and abs(sb.lr() - sa._lr_of(4)) < 1e-12 + 1e-6 * sa._lr_of(4)
    tail_b = [float(sb.step(mel, None, y, yl)) for mel in mels[3:]]
    assert tail_a == tail_b                                                   # resumed run == uninterrupted run, bit for bit
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k]), k
    # the optimizer state loads into torch.optim.Adam as it is (what the reference's Model.load does)
    opt = torch.optim.Adam(b.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    opt.load_state_dict(sb.state_dict()["optimizer"])
    # gradient accumulation
    tp2 = dict(tp); tp2["accumulated_steps"] = 2
    c = _model("bf16x2", 0.0, sp, vocab=32)
    sc = CTCTrainStep(c, tp2, precision="bf16x2")
    p0 = sc.flat.params.clone()
    sc.step(mels[0], None, y, yl)
    assert torch.equal(sc.flat.params, p0) and sc.steps_done() == 0           # first micro-batch: gradients only
    sc.step(mels[1], None, y, yl)
    assert not torch.equal(sc.flat.params, p0) and sc.steps_done() == 1


# ---- two GPUs: utterance shards + SyncBatchNorm + one all-reduced gradient bucket == the single-GPU step on the whole batch ----------
def _dp_worker(rank, world, port, prec, graph, q, overlap=False):
    import torch.distributed as dist
    os.environ["EFFCONF_BUCKET_OVERLAP"] = "1" if overlap else "0"     # gradient buckets leaving during the backward, or one after it
    from efficientconformer_b200.trainer import CTCTrainStep
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        sp = _small_params()
        tp = dict(optimizer="Adam", beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-6, lr_schedule="Transformer", schedule_dim=96,
                  warmup_steps=300, K=2)
        B, T, V2 = 4, 120, 32
        mels = [synthetic_mel(B, T, seed=60 + i) for i in range(3)]
        y, yl = synthetic_targets(torch.full((B,), 30), V2, seed=4)
        lo, hi = rank * (B // world), (rank + 1) * (B // world)

        def run(data_parallel, sl):
            from efficientconformer_b200.model_ctc import ModelCTC
            p = dict(sp); p["Pdrop"] = 0.0
            m = ModelCTC(p, {"vocab_size": V2}, precision=prec)
            m.load_state_dict(seeded_state_dict(sp, V2, seed=0, prefix_encoder="encoder."), strict=False)
            m = m.to(dev).train()
            st = CTCTrainStep(m, tp, precision=prec, use_cuda_graph=graph, sync_bn=True, data_parallel=data_parallel)
            assert bool(st._buckets) == (overlap and data_parallel)
            ls = [float(st.step(mel[sl].to(dev), None, y[sl].to(dev), yl[sl].to(dev))) for mel in mels]
            # numpy: pickled by value (tensors travel through a queue as shared-memory handles that die with the worker)
            return ls, st.flat.exp_avg.cpu().numpy().copy(), st.flat.params.cpu().numpy().copy(), \
                {k: v.cpu().numpy().copy() for k, v in m.state_dict().items() if "running" in k}
        dp = run(True, slice(lo, hi))
        out = {"rank": rank, "dp": dp}
        if rank == 0:
            out["full"] = run(False, slice(0, B))            # the whole batch on one GPU, no collectives
        q.put(out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("graph,overlap", [(False, False), (True, False), (False, True), (True, True)])
def test_two_gpu_data_parallel_step_matches_single_gpu_full_batch(graph, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, 29761 + int(graph) + 2 * int(overlap), "tf32", graph, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        o = q.get(timeout=300)
        res[o["rank"]] = o
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    t = torch.from_numpy
    unpack = lambda r: (r[0], t(r[1]), t(r[2]), {k: t(v) for k, v in r[3].items()})
    l0, m0, p0, s0 = unpack(res[0]["dp"])
    l1, m1, p1, s1 = unpack(res[1]["dp"])
    lf, mf, pf, sf = unpack(res[0]["full"])
    assert torch.equal(p0, p1) and torch.equal(m0, m1)                        # replicas stay identical
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k                                   # SyncBatchNorm: same running statistics on every rank
        assert rel_l2(s0[k], sf[k]) < 3e-4, k                                 # == statistics of the whole batch (3 Adam steps apart: 1.2e-4 measured)
    for a, b, f in zip(l0, l1, lf):
        assert abs(0.5 * (a + b) - f) < 1e-4 * abs(f), (l0, l1, lf)           # mean of the shard losses == loss of the whole batch
    print(f"\n[2 GPUs graph={graph} overlapped buckets={overlap}] shard losses {l0} {l1} | whole batch {lf} | exp_avg rel-L2 {rel_l2(m0, mf):.2e}")
    assert rel_l2(m0, mf) < 2e-3
