"""A/B of the inference engine's fusion options on the north-star forward (B = 32 x 80 x 1000): device time per forward + CTC loss
(CUDA-graph replay, L2 flushed between steps) with the fused-LayerNorm GEMM epilogues on / off, per operand mode."""
import sys
import torch

sys.path.insert(0, ".")
from efficientconformer_b200 import ModelCTC, CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V, _lib  # noqa: E402
from efficientconformer_b200.model_ctc import ctc_loss  # noqa: E402
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel, synthetic_targets  # noqa: E402

torch.set_grad_enabled(False)
L = _lib.lib()
sd = seeded_state_dict(P, V, seed=0, prefix_encoder="encoder.")
mel = synthetic_mel(32, 1000, seed=1).cuda()
ln = torch.full((32,), 1000, dtype=torch.int64, device="cuda")
y, yl = synthetic_targets(torch.full((32,), 125), V, seed=4)
y, yl = y.cuda(), yl.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for prec in (sys.argv[1:] or ["bf16x2", "tf32", "bf16"]):
    for fuse_ln in (1, 0):
        m = ModelCTC(P, {"vocab_size": V}, precision=prec)
        m.load_state_dict(sd, strict=False)
        m = m.cuda().eval()
        m.forward_mel(mel, ln)
        eng = m.encoder._engines[_lib.PRECISIONS[prec]][0]
        L.ec_engine_set_fuse_ln(eng, fuse_ln)
        m.encoder._plans.clear()

        def step():
            lg, ol, _ = m.forward_mel(mel, ln)
            return ctc_loss(lg, ol, y, yl)[0]
        for _ in range(5):
            step(); flush.zero_()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
        for a, b in ev:
            flush.zero_(); a.record(); step(); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        print(f"{prec:7s} fuse_ln={fuse_ln}  launches {L.ec_engine_last_launches(eng):4d}  median {ms[25]:.3f} ms  min {ms[0]:.3f} ms", flush=True)
        del m
