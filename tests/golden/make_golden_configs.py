"""Golden outputs of the OTHER shipped encoder configs (BASELINE.json configs 3-5 and the rest of reference configs/*.json),
produced by the REAL reference in the authoring container (see make_golden.py for the import stubs).

For every config: deterministic synthetic weights (efficientconformer_b200.synthetic, seed 11), a seeded ragged mel batch,
and the reference's mel-level forward (reference models/encoders.py:106-142; + ModelCTC.fc for CTC models).  Only the small
outputs are stored; the GPU box regenerates weights and inputs from the seeds.

    python tests/golden/make_golden_configs.py
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

from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel  # noqa: E402
from efficientconformer_b200.config import SHIPPED_ENCODER_PARAMS  # noqa: E402

SEED = 11
# config name -> (batch, mel frames, mel lengths)
CASES = {
    "EfficientConformerCTCLarge": (2, 250, [250, 171]),            # BASELINE config 3 (dims 360/512/720, head dim 135)
    "EfficientConformerCTCMedium": (2, 131, [131, 64]),
    "EfficientConformerTransducerMedium": (2, 250, [250, 99]),     # BASELINE config 4: encoder only
    "EfficientConformerTransducerLarge": (1, 97, [97]),
    "ConformerCTCLarge": (2, 203, [203, 150]),                     # BASELINE config 5 family (two Conv2d layers, k = 31, G = 1)
    "ConformerCTCLarge@long": (1, 1000, [1000]),
    "ConformerCTCMedium": (2, 64, [64, 17]),
    "ConformerCTCSmall": (3, 97, [97, 50, 3]),
    "ConformerTransducerSmall": (2, 64, [64, 33]),                 # 6 heads of 24
    "ConformerTransducerLarge": (1, 41, [41]),
}


def mel_forward(enc, mel, mel_len):
    h, l = enc.subsampling_module(mel, mel_len)
    mask = enc.padding_mask(h, l)
    h = enc.linear(h.transpose(1, 2))
    for blk in enc.blocks:
        h, _, _ = blk(h, mask)
        if blk.stride > 1:
            mask = mask[:, :, ::blk.stride, ::blk.stride]
            l = torch.div(l - 1, blk.stride, rounding_mode="floor") + 1
    return h, l


def main():
    torch.set_grad_enabled(False)
    out = {}
    models = {}
    for case, (B, T, lens) in CASES.items():
        name = case.split("@")[0]
        cfg = json.load(open(f"{REF}/configs/{name}.json"))
        params, vocab = SHIPPED_ENCODER_PARAMS[name]
        assert cfg["encoder_params"] == params and cfg["tokenizer_params"]["vocab_size"] == vocab
        is_ctc = cfg["model_type"] == "CTC"
        if name not in models:
            cwd = os.getcwd(); os.chdir(REF)
            try:
                model = create_model(cfg).eval()
            finally:
                os.chdir(cwd)
            if is_ctc:
                sd = seeded_state_dict(params, vocab, seed=SEED, prefix_encoder="encoder.")
                miss = model.load_state_dict(sd, strict=False)
            else:
                sd = seeded_state_dict(params, None, seed=SEED)
                miss = model.encoder.load_state_dict(sd, strict=False)
            assert all("preprocessing." in k for k in miss.missing_keys) and not miss.unexpected_keys, miss
            models[name] = model
        model = models[name]
        mel = synthetic_mel(B, T, seed=SEED + T)
        mel_len = torch.tensor(lens)
        x, x_len = mel_forward(model.encoder, mel, mel_len)
        entry = {"batch": B, "t_mel": T, "mel_seed": SEED + T, "mel_len": mel_len, "out_len": x_len, "weights_seed": SEED}
        if is_ctc:
            entry["logits"] = model.fc(x)          # CTC models: logits pin the whole path
        else:
            entry["x"] = x                         # Transducer models: the encoder output is the boundary
        out[case] = entry
        print(case, tuple(x.shape), x_len.tolist(), float(x.abs().max()))
        if case == name and f"{name}@long" not in CASES:
            del models[name]
    torch.save(out, f"{HERE}/other_configs.pt")
    print(f"other_configs.pt: {os.path.getsize(f'{HERE}/other_configs.pt') / 1024:.0f} KB")


if __name__ == "__main__":
    main()
