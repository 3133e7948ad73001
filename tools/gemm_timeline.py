"""Latency anatomy of the tcgen05 GEMM: SM-clock stamps of CTA (0,0) at the role hand-over points (debug hook
ec_debug_gemm_timeline).  Prints microseconds since kernel start assuming the reported SM clock."""
import ctypes as C
import sys
import torch

sys.path.insert(0, ".")
from efficientconformer_b200 import ops, _lib  # noqa: E402

L = _lib.lib()
names = ["start", "setup", "depwait", "tma0", "stage0", "mma_done", "epi_ready", "acc_full", "chunks", "ln", "stores", "end"]
mhz = 1965.0
dev = "cuda"


def run(tag, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    L.ec_debug_gemm_timeline(1, None)
    fn()
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 12)()
    L.ec_debug_gemm_timeline(0, out)
    t0 = out[0]
    print(f"{tag:44s} " + " ".join(f"{n}={(out[i] - t0) / mhz:6.2f}" for i, n in enumerate(names) if i))


for prec in (sys.argv[1:] or ["bf16", "tf32"]):
    M = 16000
    a120 = ops.cast(torch.randn(M, 120, device=dev), prec); a480 = ops.cast(torch.randn(M, 480, device=dev), prec)
    w480 = ops.cast_weight(torch.randn(480, 120, device=dev), prec); w120 = ops.cast_weight(torch.randn(120, 480, device=dev), prec)
    wsq = ops.cast_weight(torch.randn(120, 120, device=dev), prec)
    b480, b120 = torch.randn(480, device=dev), torch.randn(120, device=dev)
    res = torch.randn(M, 120, device=dev)
    g1, b1 = torch.ones(120, device=dev), torch.zeros(120, device=dev)
    r999 = ops.cast(torch.randn(999, 120, device=dev), prec)
    run(f"{prec} pos 999x120x120", lambda: ops.gemm(r999, wsq, b120, prec))
    run(f"{prec} out-proj 16000x120x120 f32", lambda: ops.gemm(a120, wsq, b120, prec))
    run(f"{prec} W1 16000x480x120 swish->act", lambda: ops.gemm(a120, w480, b480, prec, act=1, want_f32=False, want_act=True))
    run(f"{prec} W2 16000x120x480 +res f32", lambda: ops.gemm(a480, w120, b120, prec, alpha=0.5, residual=res))
    run(f"{prec} W2 16000x120x480 +res LN1", lambda: ops.gemm_ln(a480, w120, b120, prec, g1, b1, mode=1, alpha=0.5, residual=res))
    run(f"{prec} W2 16000x120x480 +res LN2", lambda: ops.gemm_ln(a480, w120, b120, prec, g1, b1, g1, b1, mode=2, alpha=0.5, residual=res))
    run(f"{prec} att-out 16000x120x120 +res LN1", lambda: ops.gemm_ln(a120, wsq, b120, prec, g1, b1, mode=1, alpha=1.0, residual=res))
    a240 = ops.cast(torch.randn(4000, 240, device=dev), prec); w240 = ops.cast_weight(torch.randn(240, 240, device=dev), prec)
    b240 = torch.randn(240, device=dev); res240 = torch.randn(4000, 240, device=dev); g240, z240 = torch.ones(240, device=dev), torch.zeros(240, device=dev)
    run(f"{prec} att-out 4000x240x240 +res LN1", lambda: ops.gemm_ln(a240, w240, b240, prec, g240, z240, mode=1, alpha=1.0, residual=res240))
    w720 = ops.cast_weight(torch.randn(720, 240, device=dev), prec); b720 = torch.randn(720, device=dev)
    run(f"{prec} qkv 4000x720x240 ->act", lambda: ops.gemm(a240, w720, b720, prec, want_f32=False, want_act=True))
