"""Generate the golden fixtures in this directory from the REAL reference (authoring container only).

Imports burchim/EfficientConformer unmodified from /root/reference (stubbing the four uninstalled third-party
modules that the encoder/CTC path never calls: jiwer, ctcdecode, warp_rnnt, kenlm), loads the deterministic
synthetic weights of efficientconformer_b200.synthetic, runs the reference modules on seeded inputs and stores
only small outputs.  /root/reference does not exist on the GPU box, so tests read these files, never the reference.

    python tests/golden/make_golden.py
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

from functions import create_model  # noqa: E402  (reference functions.py:34)

from efficientconformer_b200.synthetic import (seeded_state_dict, synthetic_mel, synthetic_audio,  # noqa: E402
                                               ragged_lengths, synthetic_targets)
from efficientconformer_b200.config import resolve_blocks  # noqa: E402


def build(cfg_name, seed=0):
    cfg = json.load(open(f"{REF}/configs/{cfg_name}.json"))
    cwd = os.getcwd(); os.chdir(REF)
    try:
        model = create_model(cfg).eval()
    finally:
        os.chdir(cwd)
    sd = seeded_state_dict(cfg["encoder_params"], cfg["tokenizer_params"]["vocab_size"], seed=seed, prefix_encoder="encoder.")
    missing = model.load_state_dict(sd, strict=False)
    assert all(k.startswith("encoder.preprocessing.") for k in missing.missing_keys), missing
    assert not missing.unexpected_keys, missing
    return cfg, model, sd


def mel_level_forward(model, mel, mel_len, taps=None):
    """The reference encoder from the mel spectrogram on (reference models/encoders.py:106-142)."""
    enc = model.encoder
    h, l = enc.subsampling_module(mel, mel_len)
    mask = enc.padding_mask(h, l)
    h = enc.linear(h.transpose(1, 2))
    if taps is not None:
        taps["linear"] = h.clone()
    for i, blk in enumerate(enc.blocks):
        h, _, _ = blk(h, mask)
        if taps is not None and i in taps["_want"]:
            taps[f"blocks.{i}"] = h.clone()
        if blk.stride > 1:
            mask = mask[:, :, ::blk.stride, ::blk.stride]
            l = torch.div(l - 1, blk.stride, rounding_mode="floor") + 1
    return model.fc(h), l


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)

    # ---- 1. state_dict layouts ------------------------------------------------------------------------
    layouts = {}
    for name in ("EfficientConformerCTCSmall", "ConformerCTCSmall", "EfficientConformerCTCMedium"):
        cfg = json.load(open(f"{REF}/configs/{name}.json"))
        cwd = os.getcwd(); os.chdir(REF)
        model = create_model(cfg)
        os.chdir(cwd)
        layouts[name] = {"encoder_params": cfg["encoder_params"], "vocab_size": cfg["tokenizer_params"]["vocab_size"],
                         "keys": [[k, list(v.shape)] for k, v in model.state_dict().items()]}
    json.dump(layouts, open(f"{HERE}/state_dict_layouts.json", "w"))

    cfg, model, sd = build("EfficientConformerCTCSmall")
    params = cfg["encoder_params"]
    V = cfg["tokenizer_params"]["vocab_size"]

    # ---- 2. BASELINE config 1: CTCSmall, B=2, 80x500 mel, ragged lengths ---------------------------------
    B, T = 2, 500
    mel = synthetic_mel(B, T, seed=1)
    mel_len = torch.tensor([500, 377])
    taps = {"_want": (0, 4, 9)}
    logits, out_len = mel_level_forward(model, mel, mel_len, taps)
    y, y_len = synthetic_targets(out_len, V, seed=4)
    loss = model.criterion((None, y, None, y_len), (logits, out_len, None))
    per_utt = torch.nn.CTCLoss(blank=0, reduction="none")(logits.log_softmax(-1).transpose(0, 1), y, out_len, y_len)
    preds = logits.log_softmax(dim=-1).argmax(dim=-1)
    greedy = []
    for b in range(B):                       # the reference's own collapse loop (reference models/model_ctc.py:105-133)
        blank, lst = False, []
        for t in range(int(out_len[b])):
            if preds[b, t] == 0:
                blank = True; continue
            if len(lst) == 0:
                lst.append(preds[b, t].item())
            elif lst[-1] != preds[b, t] or blank:
                lst.append(preds[b, t].item())
            blank = False
        greedy.append(lst)
    del taps["_want"]
    torch.save({"mel_seed": 1, "mel_len": mel_len, "logits": logits, "out_len": out_len, "targets": y, "target_len": y_len,
                "loss": loss, "loss_per_utt": per_utt, "greedy": greedy, "taps": {k: v for k, v in taps.items()}},
               f"{HERE}/ctc_small_b2_t500.pt")

    # ---- 3. audio-level forward through the unchanged ModelCTC.forward ----------------------------------
    audio = synthetic_audio(2, 200, seed=7)
    audio_len = torch.tensor([audio.shape[1], audio.shape[1] - 4321])
    lg, ln, _ = model.forward((audio, None, audio_len, None))
    mel_ref, mel_len_ref = model.encoder.preprocessing(audio, audio_len)
    torch.save({"audio_seed": 7, "t_mel": 200, "audio_len": audio_len, "logits": lg, "out_len": ln,
                "mel": mel_ref.half(), "mel_len": mel_len_ref,
                "window": model.encoder.preprocessing.Spectrogram.window.clone(),
                "fb": model.encoder.preprocessing.MelScale.fb.clone()}, f"{HERE}/ctc_small_audio_b2_t200.pt")

    # ---- 4. per-module goldens on awkward lengths --------------------------------------------------------
    specs = resolve_blocks(params)
    mods = {}
    enc = model.encoder
    cases = {0: (1, 2, 4, 17, 250, 251), 4: (1, 2, 17, 250, 251), 5: (1, 16, 125, 126), 9: (1, 2, 125, 126), 10: (1, 63)}
    for bi, Ts in cases.items():
        blk = enc.blocks[bi]
        D = specs[bi].dim_model
        for Tq in Ts:
            g = torch.Generator().manual_seed(100 * bi + Tq)
            x = torch.randn(2, Tq, D, generator=g)
            x_len = torch.tensor([Tq, max(1, (2 * Tq) // 3)])
            mask = enc.padding_mask(x.transpose(1, 2), x_len)
            out, att, _ = blk(x, mask)
            entry = {"x_len": x_len, "block": out}
            if Tq in (17, 16, 63, 2):
                ffn = blk.feed_forward_module1(x)
                mh, att_m, _ = blk.multi_head_self_attention_module(x, mask)
                cv = blk.convolution_module(x)
                entry.update({"ffn1": ffn, "mhsa": mh, "conv": cv, "att_w": att_m})
            mods[f"b{bi}_T{Tq}"] = entry
    # subsampling + linear on odd mel lengths
    for Tm in (1, 2, 7, 64, 101):
        g = torch.Generator().manual_seed(9000 + Tm)
        m = torch.randn(2, 80, Tm, generator=g)
        ml = torch.tensor([Tm, max(1, Tm // 2)])
        h, l = enc.subsampling_module(m, ml)
        mods[f"sub_T{Tm}"] = {"out": enc.linear(h.transpose(1, 2)), "out_len": l}
    torch.save(mods, f"{HERE}/ctc_small_modules.pt")

    # ---- 5. CTC loss / greedy goldens on small random logits (incl. repeated labels, U=0 edge) -----------
    g = torch.Generator().manual_seed(55)
    lg = 3.0 * torch.randn(6, 40, 16, generator=g)
    ll = torch.tensor([40, 33, 21, 9, 40, 5])
    yy = torch.tensor([[3, 3, 3, 4, 5, 5, 1, 2, 9, 9, 9, 9], [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12], [7, 7, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                       [1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0], [15, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [2, 2, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0]])
    yl = torch.tensor([12, 12, 2, 4, 1, 3])     # last one infeasible (needs 5 frames incl. 2 blanks: 2*3-1=5 -> feasible edge)
    per = torch.nn.CTCLoss(blank=0, reduction="none", zero_infinity=False)(lg.log_softmax(-1).transpose(0, 1), yy, ll, yl)
    torch.save({"logits": lg, "logits_len": ll, "targets": yy, "target_len": yl, "loss_per_utt": per, "loss": per.mean()},
               f"{HERE}/ctc_loss_small.pt")
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print(f"  {f}: {os.path.getsize(os.path.join(HERE, f)) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
