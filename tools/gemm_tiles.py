"""Tile-shape study of the plain tcgen05 GEMM on the shapes of one CTCSmall training step (B = 32 x 1000): device time per launch
(CUDA events, 30 launches, L2-warm like the step's producer -> consumer chains) for every admissible N tile width."""
import sys
import torch

sys.path.insert(0, ".")
from efficientconformer_b200 import ops, _lib  # noqa: E402

L = _lib.lib()
dev = "cuda"
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x2"
# (tag, M, N, K, outputs): f = fp32 out, a = activation-type out, r = residual
shapes = []
for M, D in ((16000, 120), (8000, 168), (4000, 240)):
    shapes += [(f"W1 fwd", M, 4 * D, D, "a"), ("W2 fwd", M, D, 4 * D, "fr"), ("qkv fwd", M, 3 * D, D, "a"), ("out/pw2 fwd", M, D, D, "fr"),
               ("pw1 fwd", M, 2 * D, D, "a"), ("dgrad W2 (ds)", M, 4 * D, D, "f"), ("dgrad W1", M, D, 4 * D, "f"), ("dgrad qkv", M, D, 3 * D, "f"),
               ("dgrad pw1", M, D, 2 * D, "f")]
for tag, M, N, K, outs in shapes:
    a = ops.cast(torch.randn(M, K, device=dev), prec)
    w = ops.cast_weight(torch.randn(N, K, device=dev) / K ** 0.5, prec)
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev) if "r" in outs else None
    row = []
    for bn in (0, 32, 64, 96, 128, 160, 192, 224, 256):
        if bn and bn >= N + 32:
            continue
        L.ec_debug_gemm_block_n(bn)
        try:
            fn = lambda: ops.gemm(a, w, bias, prec, residual=res, want_f32="f" in outs, want_act="a" in outs)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                fn()
            e1.record(); torch.cuda.synchronize()
            row.append(f"{bn}:{1e3 * e0.elapsed_time(e1) / 30:5.1f}")
        except RuntimeError as ex:
            row.append(f"{bn}:  x  ")
    L.ec_debug_gemm_block_n(0)
    print(f"{prec} {tag:14s} M={M:5d} N={N:4d} K={K:4d} [{outs:2s}]  " + "  ".join(row), flush=True)
