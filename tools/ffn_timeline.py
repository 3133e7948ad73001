"""Latency anatomy of the fused feed-forward cluster kernel: SM-clock stamps of CTA 0 (debug hook ec_debug_ffn_timeline)
and CUDA-event timing of the op at the three CTCSmall stage shapes, next to the unfused W1 / W2+LN GEMM pair."""
import ctypes as C
import sys
import torch

sys.path.insert(0, ".")
from efficientconformer_b200 import ops, _lib  # noqa: E402

L = _lib.lib()
names = ["start", "setup", "depwait", "x", "w1_0", "h0_acc", "h0_tile", "hN_tile", "y_full", "B0", "pushed", "B1", "reduced", "B2",
         "normed", "ln2", "stored", "end"]
mhz = 1965.0
dev = "cuda"


def timeline(tag, fn, detail=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    L.ec_debug_ffn_timeline(1, None)
    fn()
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 288)()
    L.ec_debug_ffn_timeline(0, out)
    t0 = out[0]
    print(f"{tag:30s} " + " ".join(f"{n}={(out[i] - t0) / mhz:5.2f}" for i, n in enumerate(names) if i))
    if detail:
        print("      phase B (warp 2): " + " ".join(f"{n}={(out[18 + i] - t0) / mhz:5.2f}" for i, n in enumerate(
            ["loaded", "res", "stats", "stored"]) if out[18 + i] > t0))
        cols = ["w1_req", "g1_rdy", "g1_iss", "g1_cmt", "acc", "loaded", "actd", "free", "written", "w2_rdy", "h_rdy", "g2_iss", "g2_cmt"]
        idx = [7, 0, 1, 2, 8, 9, 10, 11, 12, 3, 4, 5, 6]
        print("      chunk " + " ".join(f"{c:>7s}" for c in cols))
        for k in range(16):
            v = [out[32 + 16 * k + i] for i in range(16)]
            if v[0] <= t0:
                break
            f = lambda x: f"{(x - t0) / mhz:7.2f}" if x > t0 else "     - "
            print(f"      {k:5d} " + " ".join(f(v[i]) for i in idx))


def bench(fn, iters=200):
    for _ in range(10):
        fn()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    del flush
    return e0.elapsed_time(e1) / iters * 1e3


for (M, D) in ((16000, 120), (8000, 168), (4000, 240)):
    Hd = 4 * D
    x = ops.cast(torch.randn(M, D, device=dev), "bf16")
    w1 = ops.cast(torch.randn(Hd, D, device=dev) / D ** 0.5, "bf16")
    w2 = ops.cast(torch.randn(D, Hd, device=dev) / Hd ** 0.5, "bf16")
    b1, b2 = torch.randn(Hd, device=dev), torch.randn(D, device=dev)
    res = torch.randn(M, D, device=dev)
    g1, be1 = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    h = torch.empty(M, Hd, dtype=torch.bfloat16, device=dev)

    def unfused():
        _, ha = ops.gemm(x, w1, b1, "bf16", act=1, want_f32=False, want_act=True)
        ops.gemm_ln(ha, w2, b2, "bf16", g1, be1, mode=1, alpha=0.5, residual=res)

    print(f"--- M={M} D={D} hidden={Hd}: unfused pair {bench(unfused):6.1f} us (incl. allocator)")
    for cs in (1, 2, 4):
        try:
            fn1 = lambda: ops.ffn_fused(x, w1, b1, w2, b2, res, g1, be1, mode=1, cluster=cs)
            fn2 = lambda: ops.ffn_fused(x, w1, b1, w2, b2, res, g1, be1, g1, be1, mode=2, cluster=cs)
            fn1()
        except RuntimeError as e:
            print(f"cluster {cs}: {str(e).split(': ')[-1]}")
            continue
        print(f"cluster {cs}: mode1 {bench(fn1):6.1f} us   mode2 {bench(fn2):6.1f} us")
        timeline(f"  cs={cs} mode1", fn1, detail=True)
        timeline(f"  cs={cs} mode2", fn2)
