"""Target for ncu: the one-kernel log-mel front end + SpecAugment at the bench shape (32 x 159840 samples -> 32 x 80 x 1000), and, for
comparison in the same launch list, the torchaudio ops on the same CUDA tensor.  Numbers printed under ncu are never bench values."""
import os
import sys
import torch

sys.path.insert(0, ".")
from efficientconformer_b200.encoders import _PreprocessingHolder, SpecAugment  # noqa: E402

pre = _PreprocessingHolder({"sample_rate": 16000, "n_fft": 512, "win_length_ms": 25, "hop_length_ms": 10, "n_mels": 80, "normalize": False,
                            "mean": 0.0, "std": 1.0}).cuda()
aug = SpecAugment(True, 2, 27, 5, 0.05)
audio = torch.randn(32, 159840, device="cuda") * 0.1
lens = torch.full((32,), 159840, dtype=torch.int64, device="cuda")
for it in range(3):
    mel, ml = pre(audio, lens)
    y = aug(mel, ml)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
for _ in range(20):
    mel, ml = pre(audio, lens)
ev[1].record()
for _ in range(20):
    y = aug(mel, ml)
ev[2].record()
torch.cuda.synchronize()
print("device front end: logmel %.1f us, specaugment (advance + clone + kernel) %.1f us" % (ev[0].elapsed_time(ev[1]) * 50, ev[1].elapsed_time(ev[2]) * 50))
os.environ["EFFCONF_DEVICE_FRONTEND"] = "0"
for it in range(3):
    mel2, _ = pre(audio, lens)
    y2 = aug(mel2, ml)
torch.cuda.synchronize()
ev[0].record()
for _ in range(20):
    mel2, _ = pre(audio, lens)
ev[1].record()
for _ in range(20):
    y2 = aug(mel2, ml)
ev[2].record()
torch.cuda.synchronize()
print("torchaudio / torch ops on the GPU: logmel %.1f us, specaugment %.1f us" % (ev[0].elapsed_time(ev[1]) * 50, ev[1].elapsed_time(ev[2]) * 50))
print("max abs difference", float((mel - mel2).abs().max()))
