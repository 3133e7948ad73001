"""Golden vectors for the device front end, generated in the authoring container from the REAL reference module
(models/modules.py AudioPreprocessing, imported from /root/reference; it calls torchaudio's Spectrogram / MelScale) and from torchaudio's
own mask_along_axis fed with preset uniforms (the arithmetic SpecAugment's FrequencyMasking / TimeMasking run).
    python tests/golden/make_golden_frontend.py"""
import os
import sys

import torch

REF = os.environ.get("EFFCONF_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    from models.modules import AudioPreprocessing            # the reference's own module
    import torchaudio
    import torchaudio.functional as AF
    g = torch.Generator().manual_seed(21)
    out = {"source": "reference models/modules.py AudioPreprocessing + torchaudio " + torchaudio.__version__, "logmel": {}, "spans": []}
    for name, (B, L, normalize) in {"plain": (2, 8000, False), "short": (3, 1234, False), "normalized": (2, 4800, True)}.items():
        # speech-like dynamic range: noise with a slow envelope plus two tones, last utterance zero padded as collate_fn_pad does
        t = torch.arange(L) / 16000.0
        audio = torch.randn(B, L, generator=g) * (0.02 + 0.3 * torch.sin(2 * torch.pi * 3.0 * t).abs())
        audio = audio + 0.5 * torch.sin(2 * torch.pi * 440.0 * t) + 0.1 * torch.sin(2 * torch.pi * 3000.0 * t)
        lens = torch.tensor([L] + [L - 317 * b for b in range(1, B)])
        for b in range(B):
            audio[b, lens[b]:] = 0
        pre = AudioPreprocessing(16000, 512, 25, 10, 80, normalize, -5.6501, 4.2280)
        with torch.no_grad():
            mel, mel_len = pre(audio, lens)
        out["logmel"][name] = {"audio": audio, "audio_len": lens, "normalize": normalize, "mean": -5.6501, "std": 4.2280, "mel": mel, "mel_len": mel_len}
    # mask_along_axis with preset uniforms: (u1, u2, param, size) -> [start, end)
    real_rand = torch.rand
    for (u1, u2, param, size) in [(0.0, 0.0, 27, 80), (0.999999, 0.999999, 27, 80), (0.5, 0.25, 27, 80), (0.37, 0.91, 50, 1000), (0.73, 0.05, 7, 143),
                                  (0.11, 0.66, 0, 9), (0.6180339, 0.4142135, 80, 1600), (0.25, 0.75, 13, 13)]:
        seq = iter([torch.tensor([u1]), torch.tensor([u2])])
        torch.rand = lambda *a, **k: next(seq)
        try:
            y = AF.mask_along_axis(torch.ones(1, size, 4), param, 0.0, 1)          # mask along the axis of length `size`
        finally:
            torch.rand = real_rand
        zero = (y[0, :, 0] == 0).nonzero().flatten()
        start, end = (int(zero[0]), int(zero[-1]) + 1) if zero.numel() else (None, None)
        out["spans"].append({"u1": u1, "u2": u2, "param": param, "size": size, "start": start, "end": end, "width": int(zero.numel())})
    torch.save(out, os.path.join(HERE, "frontend_small.pt"))
    print({k: tuple(v["mel"].shape) for k, v in out["logmel"].items()}, out["spans"])


if __name__ == "__main__":
    main()
