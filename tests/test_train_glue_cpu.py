"""CPU check of the training step's TAPE LOGIC (efficientconformer_b200/training.py): the same forward / backward schedule the GPU
runs, with the operator table swapped for exact torch-CPU restatements (tests/cpu_ops_shim.py), must reproduce the REAL reference's
loss.backward() -- the golden gradients of tests/golden/ctc_small_train_b2_t500.pt (every parameter's gradient norm, full tensors
for a subset, updated BatchNorm running statistics).  The CUDA operators themselves are held to autograd one by one in
tests/test_gpu_backward.py, and the assembled step on the GPU in tests/test_gpu_training.py."""
import os

import pytest
import torch

from efficientconformer_b200.config import CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel
from oracle import conformer_oracle as O


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture()
def cpu_ops(monkeypatch):
    import cpu_ops_shim
    from efficientconformer_b200 import training
    monkeypatch.setattr(training, "_ops", cpu_ops_shim)
    return training


def test_training_tape_reproduces_reference_gradients(cpu_ops, golden_dir):
    training = cpu_ops
    from efficientconformer_b200.model_ctc import ModelCTC
    g = torch.load(os.path.join(golden_dir, "ctc_small_train_b2_t500.pt"))
    params = dict(P); params["Pdrop"] = 0.0
    model = ModelCTC(params, {"vocab_size": V}, precision="tf32")
    sd = seeded_state_dict(P, V, seed=0, prefix_encoder="encoder.")
    model.load_state_dict(sd, strict=False)
    model.train()
    path = training.TrainingPath(model.encoder, model.fc)
    mel = synthetic_mel(2, 500, seed=g["mel_seed"])
    with torch.no_grad():
        x, logits, out_len, tape = path.forward(mel, g["mel_len"], "tf32", want_logits=True)
    assert rel_l2(logits, g["logits"]) < 2e-5
    lg = logits.detach().double().requires_grad_(True)
    loss, _ = O.ctc_loss(lg, out_len, g["targets"], g["target_len"])
    assert abs(float(loss) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    loss.backward()
    with torch.no_grad():
        grads = path.backward(tape, None, lg.grad)
    names = [n for n, _ in path.param_list()]
    assert set(names) == set(g["grad_norms"]) and set(grads) == set(names)
    shapes = dict(path.param_list())
    floor = 1e-4 * sorted(g["grad_norms"].values())[len(g["grad_norms"]) // 2]
    worst = 0.0
    for k, ref_norm in g["grad_norms"].items():
        assert tuple(grads[k].shape) == tuple(shapes[k].shape), k
        gn = float(grads[k].double().norm())
        if ref_norm < floor:
            assert gn < floor, k
            continue
        worst = max(worst, abs(gn - ref_norm) / ref_norm)
    assert worst < 2e-3, worst
    for k, ref in g["grads"].items():
        if g["grad_norms"][k] >= floor:
            assert rel_l2(grads[k], ref) < 2e-3, k
    new_sd = model.state_dict()
    for k, ref in g["running_stats"].items():
        assert rel_l2(new_sd[k], ref) < 1e-5, k
    assert int(new_sd["encoder.subsampling_module.layers.0.1.num_batches_tracked"]) == 1


def test_training_tape_reproduces_reference_gradients_medium_config(cpu_ops, golden_dir):
    """Same check against the REAL reference on the second shipped config (EfficientConformerCTCMedium: 16 blocks, widths 180 / 256 /
    360, strided blocks 4 and 10; golden from tests/golden/make_golden_train.py --medium; lengths 250 / 163 sit off the group-of-3 and
    stride-2 grids): loss, all 620 gradient norms, a subset of full gradients, running statistics."""
    training = cpu_ops
    from efficientconformer_b200.config import SHIPPED_ENCODER_PARAMS
    from efficientconformer_b200.model_ctc import ModelCTC
    g = torch.load(os.path.join(golden_dir, "ctc_medium_train_b2_t250.pt"))
    PM, VM = SHIPPED_ENCODER_PARAMS["EfficientConformerCTCMedium"]
    params = dict(PM); params["Pdrop"] = 0.0
    model = ModelCTC(params, {"vocab_size": VM})
    model.load_state_dict(seeded_state_dict(PM, VM, seed=0, prefix_encoder="encoder."), strict=False)
    model.train()
    path = training.TrainingPath(model.encoder, model.fc)
    mel = synthetic_mel(2, 250, seed=g["mel_seed"])
    with torch.no_grad():
        x, logits, out_len, tape = path.forward(mel, g["mel_len"], "tf32", want_logits=True)
    assert rel_l2(logits, g["logits"]) < 2e-5
    lg = logits.detach().double().requires_grad_(True)
    loss, _ = O.ctc_loss(lg, out_len, g["targets"], g["target_len"])
    assert abs(float(loss.detach()) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    loss.backward()
    with torch.no_grad():
        grads = path.backward(tape, None, lg.grad)
    assert set(grads) == set(g["grad_norms"])
    floor = 1e-4 * sorted(g["grad_norms"].values())[len(g["grad_norms"]) // 2]
    worst = 0.0
    for k, ref_norm in g["grad_norms"].items():
        gn = float(grads[k].double().norm())
        if ref_norm < floor:
            assert gn < floor, k
            continue
        worst = max(worst, abs(gn - ref_norm) / ref_norm)
    assert worst < 2e-3, worst
    for k, ref in g["grads"].items():
        if g["grad_norms"][k] >= floor:
            assert rel_l2(grads[k], ref) < 2e-3, k
    new_sd = model.state_dict()
    for k, ref in g["running_stats"].items():
        assert rel_l2(new_sd[k], ref) < 1e-5, k


def test_autograd_node_fills_parameter_grads(cpu_ops):
    """EncoderTrainFn: loss.backward() through the single autograd node assigns .grad on every parameter (and only where
    requires_grad), for the encoder-only output as well (Transducer-style callers)."""
    training = cpu_ops
    from efficientconformer_b200.encoders import ConformerEncoder
    params = dict(P); params.update(Pdrop=0.0, num_blocks=2, strided_blocks=[1], expand_blocks=[1], dim_model=[24, 32], att_group_size=[3, 1],
                                    subsampling_filters=[8], num_heads=4, kernel_size=7)
    enc = ConformerEncoder(params).train()
    path = training.TrainingPath(enc, None)
    enc.blocks[0].feed_forward_module1.layers[1].weight.requires_grad_(False)
    mel = synthetic_mel(2, 37, seed=3)
    plist = [p for _, p in path.param_list()]
    x, logits, out_len = training.EncoderTrainFn.apply(path, mel, torch.tensor([37, 20]), "tf32", False, *plist)
    assert logits is None and x.shape == (2, 10, 32) and out_len.tolist() == [10, 5]
    (x.double() ** 2).sum().backward()
    for n, p in path.param_list():
        if n == "encoder.blocks.0.feed_forward_module1.layers.1.weight":
            assert p.grad is None
        else:
            assert p.grad is not None and p.grad.shape == p.shape and bool(torch.isfinite(p.grad).all()), n


VARIANTS = {
    # scaled-down relatives of the shipped Efficient Conformer configs (reference configs/EfficientConformer*.json): the block layout rules
    # (strided / expand indices, grouped attention in stage 0, head counts, tap counts) are what the tape logic depends on, not the widths
    "medium_like_three_stages": dict(num_blocks=6, strided_blocks=[1, 3], expand_blocks=[1, 3], dim_model=[16, 24, 32], num_heads=4,
                                     att_group_size=[3, 1, 1], kernel_size=15, subsampling_filters=[8]),
    "large_like_eight_heads_stride_first_block": dict(num_blocks=4, strided_blocks=[0, 2], expand_blocks=[0, 2], dim_model=[32, 48, 64], num_heads=8,
                                                      att_group_size=[3, 1, 1], kernel_size=15, subsampling_filters=[6]),
    "single_stage_no_stride_k31": dict(num_blocks=2, strided_blocks=[], expand_blocks=[], dim_model=[24], num_heads=4, att_group_size=[1],
                                       kernel_size=31, subsampling_filters=[4]),
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_training_tape_equals_oracle_autograd_on_other_block_layouts(cpu_ops, name):
    """Every parameter gradient of the training schedule equals torch autograd over the oracle's train-mode forward (itself pinned to
    the reference's loss.backward() above) for other block layouts, ragged lengths included (x_len hitting the group-of-3 and
    stride-2 boundaries); vocabulary 1000 as in the Transducer encoders' sibling CTC heads is irrelevant to the path, 64 is used."""
    training = cpu_ops
    from efficientconformer_b200.model_ctc import ModelCTC
    from efficientconformer_b200.synthetic import synthetic_targets
    params = dict(P); params.update(VARIANTS[name]); params["Pdrop"] = 0.0
    Vv = 64
    sd = seeded_state_dict(params, Vv, seed=3, prefix_encoder="encoder.")
    model = ModelCTC(params, {"vocab_size": Vv})
    model.load_state_dict(sd, strict=False)
    model.train()
    path = training.TrainingPath(model.encoder, model.fc)
    mel = synthetic_mel(3, 131, seed=17)
    mel_len = torch.tensor([131, 100, 58])
    with torch.no_grad():
        x, logits, out_len, tape = path.forward(mel, mel_len, "tf32", want_logits=True)
    leaf = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.double() if v.is_floating_point() else v)
            for k, v in sd.items()}
    ref_logits, ref_len = O.model_ctc_forward_mel(leaf, params, mel.double(), mel_len, bn={"updates": {}})
    assert torch.equal(out_len, ref_len)
    assert rel_l2(logits, ref_logits.detach()) < 1e-9
    y, yl = synthetic_targets(ref_len, Vv, seed=4)
    ref_loss, _ = O.ctc_loss(ref_logits, ref_len, y, yl)
    ref_loss.backward()
    lg = logits.detach().double().requires_grad_(True)
    loss, _ = O.ctc_loss(lg, out_len, y, yl)
    loss.backward()
    with torch.no_grad():
        grads = path.backward(tape, None, lg.grad)
    names = [n for n, _ in path.param_list()]
    assert set(grads) == set(names)
    scale = max(float(leaf[k].grad.norm()) for k in names)
    for k in names:
        ref = leaf[k].grad
        assert tuple(grads[k].shape) == tuple(ref.shape), k
        # the shim exchanges BatchNorm statistics as fp32 (mean, M2) pairs like the CUDA stages: 1e-7-level noise on exact-zero gradients
        assert float((grads[k].double() - ref).norm()) < 1e-7 * scale + 1e-6 * float(ref.norm()), k
