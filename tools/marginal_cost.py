"""What each kernel category really costs inside the replayed CUDA graph (PDL overlap included): the forward is re-captured with
one category's launches skipped (ec_engine_set_skip_mask; outputs are garbage, timing is what is measured) and the step time
compared with the full step.  The per-launch CUDA-event / ncu durations are serialised numbers and overstate these."""
import argparse
import statistics
import sys
import torch

sys.path.insert(0, ".")
from efficientconformer_b200 import ModelCTC, _lib, CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V  # noqa: E402
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=1000)
a = ap.parse_args()
torch.set_grad_enabled(False)
L = _lib.lib()
sd = seeded_state_dict(P, V, seed=0, prefix_encoder="encoder.")
m = ModelCTC(P, {"vocab_size": V}, precision=a.precision)
m.load_state_dict(sd, strict=False)
m = m.cuda().eval()
mel = synthetic_mel(a.batch, a.frames, seed=1).cuda()
ln = torch.full((a.batch,), a.frames, dtype=torch.int64, device="cuda")
m.forward_mel(mel, ln)
enc = m.encoder
prec = _lib.PRECISIONS[a.precision]
eng = enc._engines[prec][0]
names = [L.ec_profile_category_name(i).decode() for i in range(L.ec_profile_categories())]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")


def step_ms(mask, reps=30):
    L.ec_engine_set_skip_mask(eng, mask)
    enc._plans.clear()                       # re-capture the graph with this mask
    for _ in range(3):
        m.forward_mel(mel, ln)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); enc.forward_mel(mel, ln, want_logits=True, clone=False); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


full = step_ms(0)
print(f"full forward (graph replay, L2 flushed, incl. input copy): {full * 1e3:.1f} us")
rows = []
for i, n in enumerate(names):
    if n in ("misc", "layernorm", "im2col_3x3s2", "gemm_sub_conv2_swish", "gemm_ffn_w1_swish", "gemm_ffn_w2_res"):
        continue
    t = step_ms(1 << i)
    rows.append((full - t, n))
for d, n in sorted(rows, reverse=True):
    print(f"  {n:22s} marginal {d * 1e3:7.1f} us  ({100 * d / full:4.1f} %)")
print(f"  sum of marginals {sum(d for d, _ in rows) * 1e3:.1f} us")
L.ec_engine_set_skip_mask(eng, 0)
