"""First-contact probe for a fresh B200 box: runs each kernel family once with diagnostics that are more useful than a
pytest assert when a descriptor / layout is wrong (prints error maps instead of failing fast)."""
import math
import sys
import time
import torch

sys.path.insert(0, ".")
from efficientconformer_b200 import ops, _lib  # noqa: E402

dev = "cuda"
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
_lib.check(_lib.lib().ec_device_check())


def probe_gemm(prec, M, N, K):
    g = torch.Generator().manual_seed(0)
    a = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
    aa, ww = ops.cast(a, prec), ops.cast(w, prec)
    out, _ = ops.gemm(aa, ww, None, prec)
    torch.cuda.synchronize()
    ref = aa.double() @ ww.double().t()
    err = (out.double() - ref).abs()
    rel = float(err.norm() / ref.norm())
    print(f"gemm {prec} M={M} N={N} K={K}: rel-L2 {rel:.3e} max {float(err.max()):.3e}")
    if rel > 1e-4:
        bad = (err > 1e-3)
        print("   bad fraction", float(bad.float().mean()), "bad rows", bad.any(1).nonzero().flatten()[:16].tolist(),
              "bad cols", bad.any(0).nonzero().flatten()[:16].tolist())
        print("   out[0,:8]", out[0, :8].tolist()); print("   ref[0,:8]", ref[0, :8].float().tolist())
        # which K prefix does out correspond to?
        for kk in (8, 16, 32, 64, K // 2):
            if kk < K:
                r2 = aa[:, :kk].double() @ ww[:, :kk].double().t()
                print(f"   rel-L2 vs first {kk} of K: {float((out.double() - r2).norm() / r2.norm()):.3e}")
    return rel


ok = True
for prec in ("tf32", "bf16"):
    for (M, N, K) in [(128, 128, 32), (128, 128, 64), (128, 256, 128), (256, 240, 120), (1000, 480, 120), (16000, 120, 480)]:
        try:
            ok &= probe_gemm(prec, M, N, K) < 1e-4
        except Exception as e:  # noqa
            print("EXC", prec, M, N, K, repr(e)); ok = False
print("GEMM PROBE", "OK" if ok else "FAILED")

# quick timing of a few shapes (warm L2): events around 20 launches
for prec in ("tf32", "bf16"):
    for (M, N, K) in [(16000, 480, 120), (16000, 120, 480), (16000, 360, 120), (8000, 672, 168), (4000, 960, 240), (16000, 120, 4800)]:
        a = ops.cast(torch.randn(M, K, device=dev), prec); w = ops.cast(torch.randn(N, K, device=dev), prec)
        b = torch.randn(N, device=dev)
        for _ in range(3):
            ops.gemm(a, w, b, prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.gemm(a, w, b, prec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"time gemm {prec} {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s (incl. host launch + output alloc)")
