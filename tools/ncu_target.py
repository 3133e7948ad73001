"""Target for ncu: one eager (no CUDA graph) forward + CTC loss of the bench workload between cudaProfilerStart/Stop.
Use with `ncu --profile-from-start off ...`.  Numbers printed under ncu are never bench values."""
import argparse
import sys
import torch

sys.path.insert(0, ".")
from efficientconformer_b200 import ModelCTC, CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V  # noqa: E402
from efficientconformer_b200.model_ctc import ctc_loss  # noqa: E402
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel, synthetic_targets  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=1000)
a = ap.parse_args()
torch.set_grad_enabled(False)
sd = seeded_state_dict(P, V, seed=0, prefix_encoder="encoder.")
m = ModelCTC(P, {"vocab_size": V}, precision=a.precision)
m.load_state_dict(sd, strict=False)
m = m.cuda().eval()
m.encoder.use_cuda_graph = False
mel = synthetic_mel(a.batch, a.frames, seed=1).cuda()
ln = torch.full((a.batch,), a.frames, dtype=torch.int64, device="cuda")
lg, ol, _ = m.forward_mel(mel, ln)
y, yl = synthetic_targets(ol.cpu(), V, seed=4)
y, yl = y.cuda(), yl.cuda()
ctc_loss(lg, ol, y, yl)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
lg, ol, _ = m.forward_mel(mel, ln)
loss = ctc_loss(lg, ol, y, yl)[0]
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss))
