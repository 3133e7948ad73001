"""Seeded random-shape sweeps of the CUDA operators (awkward M / N / K / T / x_len that the fixed cases do not hit):
each draw is checked against the same fp64 restatements as tests/test_gpu_ops.py."""
import math
import random

import pytest
import torch

from test_gpu_ops import rel_l2, rnd, tf32_round, _attention_reference

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from efficientconformer_b200 import ops as o
    return o


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_random_gemm_shapes(ops, prec):
    rng = random.Random(1234)
    for trial in range(24):
        M = rng.choice([1, 31, 128, 129, 777, 4000, 8191])
        N = rng.choice([8, 24, 120, 168, 240, 256, 264, 360, 504, 672, 960])
        K = rng.choice([8, 40, 120, 168, 240, 480, 672, 1000])
        if prec == "bf16" and (K * 2) % 16:
            continue
        g = torch.Generator().manual_seed(trial)
        a = ops.cast(torch.randn(M, K, generator=g).to(DEV), prec)
        w = ops.cast((torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV), prec)
        bias = torch.randn(N, generator=g).to(DEV)
        res = torch.randn(M, N, generator=g).to(DEV) if trial % 2 else None
        act = trial % 3 == 0
        ref = a.double() @ w.double().t() + bias.double()
        if act:
            ref = ref * torch.sigmoid(ref)
        ref = 0.75 * ref + (res.double() if res is not None else 0)
        out, out_a = ops.gemm(a, w, bias, prec, alpha=0.75, act=1 if act else 0, residual=res, want_act=True)
        tol = 5e-6 if prec == "tf32" or not act else 1.5e-3           # bf16 mode uses the hardware tanh in Swish
        assert rel_l2(out, ref) < tol, (trial, M, N, K, act)
        assert torch.equal(out_a.float().cpu(), rnd(prec, out.cpu())), (trial, M, N, K)


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_random_gemm_layernorm_shapes(ops, prec):
    rng = random.Random(99)
    for trial in range(16):
        fps = rng.choice([1, 7, 63, 125, 250])
        M = fps * rng.choice([1, 3, 16, 64])
        N = rng.choice([8, 24, 120, 168, 176, 240, 256])
        K = rng.choice([16, 120, 240, 672, 960])
        if prec == "bf16" and (K * 2) % 16:
            continue
        stride = rng.choice([1, 2])
        g = torch.Generator().manual_seed(1000 + trial)
        a = ops.cast(torch.randn(M, K, generator=g).to(DEV), prec)
        w = ops.cast((torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV), prec)
        bias = torch.randn(N, generator=g).to(DEV)
        res = (torch.randn(M, N, generator=g) * 3 + 1.5).to(DEV)
        g1, b1, g2, b2 = [(1 + 0.2 * torch.randn(N, generator=g)).to(DEV) if i % 2 == 0 else (0.2 * torch.randn(N, generator=g)).to(DEV) for i in range(4)]
        x = a.double() @ w.double().t() + bias.double() + res.double()
        ln = lambda t, gg, bb: torch.nn.functional.layer_norm(t, (N,), gg.double(), bb.double(), 1e-6)
        out, y, cp = ops.gemm_ln(a, w, bias, prec, g1, b1, mode=1, residual=res, copy_stride=stride, frames_per_seq=fps)
        assert rel_l2(out, x) < 5e-6, (trial, M, N, K)
        assert rel_l2(y.float(), ln(x, g1, b1)) < (5e-4 if prec == "tf32" else 4e-3), (trial, M, N, K)
        sel = x.reshape(M // fps, fps, N)[:, ::stride].reshape(-1, N)
        assert rel_l2(cp.float(), sel) < (5e-4 if prec == "tf32" else 4e-3)
        out2, y2, _ = ops.gemm_ln(a, w, bias, prec, g1, b1, g2, b2, mode=2, residual=res)
        assert rel_l2(out2, ln(x, g1, b1)) < 1e-5, (trial, M, N, K)
        assert rel_l2(y2.float(), ln(ln(x, g1, b1), g2, b2)) < (5e-4 if prec == "tf32" else 4e-3)


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_random_attention_shapes(ops, prec):
    rng = random.Random(7)
    for trial in range(22):
        # (D, H, G): the head layouts of every shipped config family (head dims 90/42/60/75/64, then 135/44/24/64/90 of the
        # Medium / Large / Conformer families)
        D, H, G = rng.choice([(120, 4, 3), (168, 4, 1), (240, 4, 1), (100, 4, 3), (256, 4, 1)]) if trial < 14 else \
            rng.choice([(360, 8, 3), (180, 4, 3), (176, 4, 1), (144, 6, 1), (512, 8, 1), (720, 8, 1)])
        T = rng.choice([1, 2, 3, 5, 63, 64, 65, 128, 191, 192, 193, 400])
        B = rng.choice([1, 2, 3])
        if prec == "bf16" and ((G * D) // H) % 2:
            continue                                      # odd head dims take the TF32 kernel inside the engine
        g = torch.Generator().manual_seed(500 + trial)
        qkv = rnd(prec, torch.randn(B, T, 3 * D, generator=g)).to(DEV)
        Tp = T + (-T) % G
        E = rnd(prec, torch.randn(2 * Tp - G, D, generator=g)).to(DEV)
        u, v = (0.3 * torch.randn(D, generator=g)).to(DEV), (0.3 * torch.randn(D, generator=g)).to(DEV)
        x_len = torch.tensor([rng.randint(1, T) for _ in range(B)], device=DEV)
        out = ops.relpos_attention(qkv, E, u, v, x_len, H, G, prec)
        ref = _attention_reference(qkv, E, u, v, x_len, H, G)
        assert rel_l2(out.float(), ref) < (2e-3 if prec == "tf32" else 1e-2), (trial, B, T, D, H, G)


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_random_dwconv_shapes(ops, prec):
    rng = random.Random(3)
    for trial in range(18):
        C = rng.choice([8, 120, 168, 176, 240, 256]) if trial < 12 else rng.choice([360, 512, 720, 260])   # > 256: channel tiles
        k = rng.choice([15, 31])
        stride = rng.choice([1, 2])
        T = rng.choice([1, 2, 13, 64, 65, 127, 500])
        B = rng.choice([1, 3])
        g = torch.Generator().manual_seed(300 + trial)
        x = ops.cast(torch.randn(B, T, C, generator=g).to(DEV), prec)
        w = (torch.randn(C, 1, k, generator=g) / math.sqrt(k)).to(DEV)
        b, gam, bet = (0.1 * torch.randn(C, generator=g)).to(DEV), (1 + 0.1 * torch.randn(C, generator=g)).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
        rm, rv = (0.1 * torch.randn(C, generator=g)).to(DEV), (0.5 + torch.rand(C, generator=g)).to(DEV)
        wf, bf = ops.fold_bn(w, b, gam, bet, rm, rv)
        y = ops.dwconv_bn_swish(x, wf, bf, stride, prec)
        pad = (k - 1) // 2
        conv = torch.nn.functional.conv1d(torch.nn.functional.pad(x.double().transpose(1, 2), (pad, pad)), w.double(), b.double(), stride=stride, groups=C)
        bn = (conv - rm.double()[None, :, None]) / torch.sqrt(rv.double()[None, :, None] + 1e-5) * gam.double()[None, :, None] + bet.double()[None, :, None]
        ref = (bn * torch.sigmoid(bn)).transpose(1, 2)
        assert y.shape == ref.shape
        assert rel_l2(y.float(), ref) < (5e-4 if prec == "tf32" else 5e-3), (trial, B, T, C, k, stride)
