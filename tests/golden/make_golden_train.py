"""Golden vectors of the TRAINING step of the hot path, produced by the REAL reference in the
authoring container: EfficientConformerCTCSmall in .train() mode (batch-statistics BatchNorm, running-stat updates) with
Pdrop = 0 and no SpecAugment (the two random parts cannot be pinned), BASELINE config-1 shape (B=2, 80x500 mel, ragged):
CTC loss, the gradient of every parameter (L2 norm for all 630 tensors, full tensors for a representative subset) and the
updated BatchNorm running statistics.  The oracle's autograd (tests/test_oracle_golden.py) and, once built, the CUDA backward
are held to these.

    python tests/golden/make_golden_train.py
"""
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
for n in ("jiwer", "ctcdecode", "warp_rnnt", "kenlm"):
    sys.modules[n] = types.ModuleType(n)
sys.modules["ctcdecode"].CTCBeamDecoder = object
REF = "/root/reference"
sys.path.insert(0, REF)

from functions import create_model  # noqa: E402

from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel, synthetic_targets  # noqa: E402

FULL = ["fc.bias", "fc.weight", "encoder.linear.bias", "encoder.subsampling_module.layers.0.0.weight",
        "encoder.subsampling_module.layers.0.1.weight", "encoder.subsampling_module.layers.0.1.bias",
        "encoder.blocks.0.feed_forward_module1.layers.0.weight", "encoder.blocks.0.feed_forward_module1.layers.1.bias",
        "encoder.blocks.0.multi_head_self_attention_module.mhsa.u", "encoder.blocks.0.multi_head_self_attention_module.mhsa.v",
        "encoder.blocks.0.multi_head_self_attention_module.mhsa.pos_layer.bias",
        "encoder.blocks.0.multi_head_self_attention_module.mhsa.query_layer.weight",
        "encoder.blocks.4.convolution_module.layers.4.weight", "encoder.blocks.4.convolution_module.layers.5.weight",
        "encoder.blocks.4.convolution_module.layers.5.bias", "encoder.blocks.4.conv_res.1.weight",
        "encoder.blocks.9.multi_head_self_attention_module.mhsa.key_layer.bias", "encoder.blocks.14.norm.weight",
        "encoder.blocks.14.feed_forward_module2.layers.4.bias"]


def main(config="EfficientConformerCTCSmall", out_name="ctc_small_train_b2_t500.pt", B=2, T=500, lens=(500, 377), full_names=None):
    full_names = FULL if full_names is None else full_names
    cfg = json.load(open(f"{REF}/configs/{config}.json"))
    cfg["encoder_params"]["Pdrop"] = 0.0
    cwd = os.getcwd(); os.chdir(REF)
    try:
        model = create_model(cfg)
    finally:
        os.chdir(cwd)
    params = dict(cfg["encoder_params"]); params["Pdrop"] = 0.1       # names / shapes do not depend on Pdrop
    V = cfg["tokenizer_params"]["vocab_size"]
    sd = seeded_state_dict(params, V, seed=0, prefix_encoder="encoder.")
    miss = model.load_state_dict(sd, strict=False)
    assert all("preprocessing." in k for k in miss.missing_keys) and not miss.unexpected_keys
    model.train()
    mel = synthetic_mel(B, T, seed=1)
    mel_len = torch.tensor(list(lens))
    enc = model.encoder
    h, l = enc.subsampling_module(mel, mel_len)
    mask = enc.padding_mask(h, l)
    h = enc.dropout(enc.linear(h.transpose(1, 2))) if hasattr(enc, "dropout") else enc.linear(h.transpose(1, 2))
    for blk in enc.blocks:
        h, _, _ = blk(h, mask)
        if blk.stride > 1:
            mask = mask[:, :, ::blk.stride, ::blk.stride]
            l = torch.div(l - 1, blk.stride, rounding_mode="floor") + 1
    logits = model.fc(h)
    y, y_len = synthetic_targets(l, V, seed=4)
    loss = model.criterion((None, y, None, y_len), (logits, l, None))
    loss.backward()
    norms, full = {}, {}
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        norms[k] = float(p.grad.double().norm())
        if k in full_names:
            full[k] = p.grad.clone()
    new_sd = model.state_dict()
    stats = {k: new_sd[k].clone() for k in new_sd if k.endswith("running_mean") or k.endswith("running_var")}
    torch.save({"mel_seed": 1, "mel_len": mel_len, "targets": y, "target_len": y_len, "loss": loss.detach(), "logits": logits.detach(),
                "grad_norms": norms, "grads": full, "running_stats": stats, "config": config, "shape": (B, T)}, f"{HERE}/{out_name}")
    print(config, "loss", float(loss), "params", len(norms), "size KB", os.path.getsize(f"{HERE}/{out_name}") // 1024)


if __name__ == "__main__":
    if "--medium" in sys.argv:
        # second shipped config (16 blocks, widths 180 / 256 / 360, strided blocks 4 and 10), lengths off the group-of-3 / stride-2 grid
        main("EfficientConformerCTCMedium", "ctc_medium_train_b2_t250.pt", B=2, T=250, lens=(250, 163),
             full_names=["fc.bias", "encoder.linear.bias", "encoder.blocks.10.conv_res.1.bias", "encoder.blocks.4.convolution_module.layers.5.weight",
                         "encoder.blocks.0.multi_head_self_attention_module.mhsa.u", "encoder.blocks.15.norm.weight"])
    else:
        main()
