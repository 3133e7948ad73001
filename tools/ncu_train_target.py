"""Target for ncu: one eager (no CUDA graph) TRAINING step of the bench workload (train-mode forward + CTC loss/grad + backward +
gradient pack + Adam) between cudaProfilerStart/Stop.  Use with `ncu --profile-from-start off ...`.
Numbers printed under ncu are never bench values."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientconformer_b200 import ModelCTC, CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V  # noqa: E402
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel, synthetic_targets  # noqa: E402
from efficientconformer_b200.trainer import CTCTrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--pdrop", type=float, default=0.1)
a = ap.parse_args()
params = dict(P); params["Pdrop"] = a.pdrop
m = ModelCTC(params, {"vocab_size": V}, precision=a.precision)
m.load_state_dict(seeded_state_dict(P, V, seed=0, prefix_encoder="encoder."), strict=False)
m = m.cuda().train()
tp = dict(optimizer="Adam", beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-6, lr_schedule="Transformer", schedule_dim=240,
          warmup_steps=10000, K=2)
step = CTCTrainStep(m, tp, precision=a.precision, use_cuda_graph=False)
mel = synthetic_mel(a.batch, a.frames, seed=1).cuda()
t_out = ((((a.frames - 1) // 2 + 1) - 1) // 2 + 1 - 1) // 2 + 1
y, yl = synthetic_targets(torch.full((a.batch,), t_out), V, seed=4)
y, yl = y.cuda(), yl.cuda()
for _ in range(2):
    step.step(mel, None, y, yl)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = step.step(mel, None, y, yl)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss))
