"""Per-kernel parity on the B200: every CUDA operator (called through the C ABI) against (a) an exact fp64 torch
restatement fed with the SAME rounded operands (tight tolerance: validates tcgen05 descriptors / layouts / indexing) and
(b) the CPU oracle's module functions (tolerance = operand format noise)."""
import math

import pytest
import torch

from efficientconformer_b200.config import CTC_SMALL_ENCODER_PARAMS as P, resolve_blocks
from efficientconformer_b200.synthetic import seeded_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def tf32_round(x):
    xi = x.float().contiguous().view(torch.int32)
    return ((xi + 0x1000) & ~0x1FFF).view(torch.float32)


def split_round(x):
    """hi + lo with hi = bf16(x), lo = bf16(x - hi): the value a split-mode (bf16x2) operand element carries."""
    hi = x.float().to(torch.bfloat16).float()
    return hi + (x.float() - hi).to(torch.bfloat16).float()


def rnd(prec, x):
    if prec == "bf16x2":
        return split_round(x)
    return tf32_round(x) if prec == "tf32" else x.to(torch.bfloat16).float()


PRECS = ["tf32", "bf16", "bf16x2"]


@pytest.fixture(scope="module")
def ops():
    from efficientconformer_b200 import ops as o
    return o


@pytest.mark.parametrize("prec", PRECS)
def test_cast_rounding(ops, prec):
    x = torch.randn(1000, 37, device=DEV) * 3
    y = ops.unpack(ops.cast(x, prec), prec)
    if prec == "bf16x2":
        assert torch.equal(y.cpu(), split_round(x.cpu()))
        assert rel_l2(y, x) < 2e-5            # 16 significant bits
    elif prec == "tf32":
        # round-to-nearest (ties away) onto 10 mantissa bits
        assert torch.equal(y.cpu(), tf32_round(x.cpu()))
    else:
        assert torch.equal(y.cpu(), x.cpu().to(torch.bfloat16).float())


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 16, 64), (256, 240, 120), (16000, 480, 120), (8000, 168, 672), (4000, 256, 240),
                                   (999, 120, 120), (1, 120, 120), (130, 504, 168), (333, 960, 240), (700, 120, 4800), (77, 8, 40)])
def test_gemm_exact_operands(ops, prec, M, N, K):
    if prec == "bf16" and (K * 2) % 16:
        pytest.skip("bf16 row pitch must be a multiple of 16 bytes")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    aa, ww = ops.cast(a, prec), ops.cast_weight(w, prec)
    ref = ops.unpack(aa, prec).double() @ ops.unpack(ww, prec).double().t() + bias.double()
    out, _ = ops.gemm(aa, ww, bias, prec)
    tol = 2e-6 if K <= 1024 else 2e-5      # tensor-core fp32 accumulation over long K is slightly lossier than an IEEE fp32 chain
    if prec == "bf16x2":
        tol *= 2                           # four partial products per element pair are accumulated
    assert rel_l2(out, ref) < tol, "plain"
    out2, out2a = ops.gemm(aa, ww, bias, prec, alpha=0.5, act=0, residual=res, want_act=True)
    ref2 = 0.5 * ref + res.double()
    assert rel_l2(out2, ref2) < tol, "residual"
    assert torch.equal(ops.unpack(out2a, prec).cpu(), rnd(prec, out2.cpu())), "activation-type copy is the rounded fp32 output"
    out3, _ = ops.gemm(aa, ww, bias, prec, act=1)
    ref3 = ref * torch.sigmoid(ref)
    assert rel_l2(out3, ref3) < 3 * tol, "swish"


class _Drop:
    """The DropoutState the operators take: p, device {seed, step} counter."""
    def __init__(self, ops, p, seed=77):
        self.p = p
        self.counter = ops.dropout_counter(DEV, seed) if p > 0 else None


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
@pytest.mark.parametrize("M,N,K", [(16000, 480, 120), (8000, 168, 672), (4000, 240, 960), (999, 120, 120), (130, 504, 168), (1, 120, 480),
                                   (333, 256, 240), (257, 96, 64)])
def test_gemm_train_epilogues_match_the_element_kernels(ops, prec, p_drop, M, N, K):
    """ec_op_gemm_train (dropout / Swish side output / Swish-dropout backward folded into the GEMM epilogue) against the unfused chain
    GEMM -> element kernel it replaces: same masks (same counter hash), same values up to the fp32 rounding of one extra product."""
    if prec == "bf16" and ((K * 2) % 16 or (N * 2) % 16):
        pytest.skip("bf16 row pitch must be a multiple of 16 bytes")
    g = torch.Generator(device="cpu").manual_seed(M * 5 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    aa, ww = ops.cast(a, prec), ops.cast_weight(w, prec)
    drop = _Drop(ops, p_drop)
    # (1) projection + dropout + alpha + residual (feed-forward W2, attention output, pointwise conv 2, encoder.linear)
    y = ops.gemm(aa, ww, bias, prec)[0]
    ref = ops.dropout_residual(y, drop, 7, 0.5, res) if p_drop > 0 else 0.5 * y + res
    out = ops.gemm_train(aa, ww, bias, prec, drop, alpha=0.5, residual=res, site=7)[0]
    assert rel_l2(out, ref) < 1e-6
    if p_drop > 0:
        kept_ref, kept = (ref != res), (out != res)
        assert torch.equal(kept_ref, kept)                         # the very same mask
        assert abs(float(kept.float().mean()) - (1 - p_drop)) < 0.02 or M * N < 5000
        ref_nores = ops.dropout_f32(y, drop, 9)
        out_nores = ops.gemm_train(aa, ww, bias, prec, drop, site=9)[0]
        assert rel_l2(out_nores, ref_nores) < 1e-6 and torch.equal(out_nores == 0, ref_nores == 0)
    # (2) feed-forward W1: pre-activation z and s = dropout(Swish(z)) from one epilogue
    z_ref = ops.gemm(aa, ww, bias, prec, want_f32=False, want_act=True)[1]
    s_ref = ops.swish_dropout_fwd(z_ref, drop, 11, prec)
    _, z, s = ops.gemm_train(aa, ww, bias, prec, drop, want_f32=False, want_act=True, want_act2=True, site2=11)
    assert torch.equal(z.view(torch.int32) if z.dtype == torch.float32 else z.view(torch.int16),
                       z_ref.view(torch.int32) if z.dtype == torch.float32 else z_ref.view(torch.int16))
    sv, sv_ref = ops.unpack(s, prec), ops.unpack(s_ref, prec)
    assert torch.equal(sv == 0, sv_ref == 0) or p_drop == 0
    assert rel_l2(sv, sv_ref) < (1e-2 if prec == "bf16" else 2e-3 if prec == "tf32" else 3e-5)   # hardware tanh vs exp sigmoid, then rounding
    # (3) data gradient through dropout(Swish(z)): dz = mask * Swish'(z) * (dy W^T) straight into the activation type
    dy = torch.randn(M, K, generator=g).to(DEV)                   # reuse [M, K] x [N, K]^T: "dy" [M, K], weight [N, K] -> ds [M, N]
    ds = ops.gemm(ops.cast(dy, prec), ww, None, prec)[0]
    dz_ref = ops.swish_dropout_bwd(z_ref, ds, drop, 11, prec)
    dz = ops.gemm_train(ops.cast(dy, prec), ww, None, prec, drop, want_f32=False, want_act=True, aux=z_ref, site_aux=11)[1]
    dv, dv_ref = ops.unpack(dz, prec), ops.unpack(dz_ref, prec)
    assert torch.equal(dv == 0, dv_ref == 0) or p_drop == 0
    assert rel_l2(dv, dv_ref) < (8e-3 if prec == "bf16" else 1e-3 if prec == "tf32" else 3e-5)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
@pytest.mark.parametrize("M,N,K", [(16000, 120, 480), (8000, 168, 168), (4000, 240, 960), (999, 120, 120), (130, 256, 64), (1, 96, 96)])
def test_gemm_ln_train_matches_gemm_train_plus_layernorm(ops, prec, p_drop, M, N, K):
    """ec_op_gemm_ln_train (projection + dropout + residual + the next module's LayerNorm in one epilogue) == ec_op_gemm_train followed
    by the stand-alone LayerNorm kernel: same dropout mask, same fp32 output, LayerNorm output up to one rounding of the operand type."""
    if prec == "bf16" and ((K * 2) % 16 or (N * 2) % 16):
        pytest.skip("bf16 row pitch must be a multiple of 16 bytes")
    g = torch.Generator(device="cpu").manual_seed(M + N * 3 + K * 7)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias, res = torch.randn(N, generator=g).to(DEV), torch.randn(M, N, generator=g).to(DEV)
    lg, lb = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV), (0.1 * torch.randn(N, generator=g)).to(DEV)
    aa, ww = ops.cast(a, prec), ops.cast_weight(w, prec)
    drop = _Drop(ops, p_drop)
    ref = ops.gemm_train(aa, ww, bias, prec, drop, alpha=0.5, residual=res, site=5)[0]
    ref_ln = ops.layernorm(ref, lg, lb, prec, want_f32=False, want_act=True)[0]
    out, ln = ops.gemm_ln_train(aa, ww, bias, prec, lg, lb, drop, alpha=0.5, residual=res, site=5)
    assert rel_l2(out, ref) < 1e-6
    if p_drop > 0:
        assert torch.equal(out == res, ref == res)                 # the very same mask
    assert rel_l2(ops.unpack(ln, prec), ops.unpack(ref_ln, prec)) < (8e-3 if prec == "bf16" else 1e-3 if prec == "tf32" else 3e-5)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("M,N,K,fps,stride", [(16000, 120, 480, 500, 2), (8000, 168, 168, 250, 2), (4000, 240, 960, 125, 1), (999, 120, 4800, 333, 2),
                                              (37, 256, 64, 37, 1), (130, 8, 40, 13, 2)])
def test_gemm_fused_layernorm(ops, prec, M, N, K, fps, stride):
    if prec == "bf16" and (K * 2) % 16:
        pytest.skip("bf16 row pitch must be a multiple of 16 bytes")
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = (torch.randn(M, N, generator=g) * 2 + 0.3).to(DEV)
    g1, b1 = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV), (0.1 * torch.randn(N, generator=g)).to(DEV)
    g2, b2 = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV), (0.1 * torch.randn(N, generator=g)).to(DEV)
    aa, ww = ops.cast(a, prec), ops.cast_weight(w, prec)
    x_ref = 0.5 * (ops.unpack(aa, prec).double() @ ops.unpack(ww, prec).double().t() + bias.double()) + res.double()
    ln = lambda x, gg, bb: torch.nn.functional.layer_norm(x, (N,), gg.double(), bb.double(), 1e-6)
    # mode 1: out = x, ln_out = LN1(x), strided compaction copy of x
    out, y, cp = ops.gemm_ln(aa, ww, bias, prec, g1, b1, mode=1, alpha=0.5, residual=res, copy_stride=stride, frames_per_seq=fps)
    tol = 2e-6 if K <= 1024 else 2e-5
    assert rel_l2(out, x_ref) < tol
    y_ref = ln(x_ref, g1, b1)
    y = ops.unpack(y, prec)
    assert rel_l2(y, y_ref) < (4e-4 if prec != "bf16" else 4e-3)          # + rounding to the activation type
    assert rel_l2(y, rnd(prec, y_ref.float().cpu())) < 2e-4 + (0 if prec != "bf16" else 2e-3)
    sel = x_ref.reshape(M // fps, fps, N)[:, ::stride].reshape(-1, N)
    assert cp.shape == sel.shape
    assert rel_l2(ops.unpack(cp, prec), sel) < (4e-4 if prec != "bf16" else 4e-3)
    # mode 2: out = LN1(x) in place, ln_out = LN2(out)
    out2, y2, _ = ops.gemm_ln(aa, ww, bias, prec, g1, b1, g2, b2, mode=2, alpha=0.5, residual=res)
    assert rel_l2(out2, y_ref) < 5e-6 + tol
    assert rel_l2(ops.unpack(y2, prec), ln(y_ref, g2, b2)) < (4e-4 if prec != "bf16" else 4e-3)
    # mode 2 with identity second stage: activation-type copy of the normalised output
    out3, y3, _ = ops.gemm_ln(aa, ww, bias, prec, g1, b1, None, None, mode=2, alpha=0.5, residual=res)
    assert torch.equal(out3, out2)
    assert torch.equal(ops.unpack(y3, prec).cpu(), rnd(prec, out3.cpu()))


@pytest.mark.parametrize("M,D,hidden,clusters", [(16000, 120, 480, (0, 1, 2)), (8000, 168, 672, (0, 2)), (4000, 240, 960, (0, 2, 4)),
                                                 (999, 120, 480, (0, 2)), (37, 256, 1024, (0, 2, 4)), (130, 64, 256, (0, 1, 2, 4)),
                                                 (300, 16, 64, (0, 1))])
def test_ffn_fused(ops, M, D, hidden, clusters):
    """Fused feed-forward cluster kernel vs the module's formula (reference models/modules.py:367-398, blocks.py:123-150)."""
    g = torch.Generator(device="cpu").manual_seed(M + D + hidden)
    x = torch.randn(M, D, generator=g).to(DEV)
    w1 = (torch.randn(hidden, D, generator=g) / math.sqrt(D)).to(DEV)
    w2 = (torch.randn(D, hidden, generator=g) / math.sqrt(hidden)).to(DEV)
    b1, b2 = (0.3 * torch.randn(hidden, generator=g)).to(DEV), (0.3 * torch.randn(D, generator=g)).to(DEV)
    res = (torch.randn(M, D, generator=g) * 2 + 0.3).to(DEV)
    g1, be1 = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV), (0.1 * torch.randn(D, generator=g)).to(DEV)
    g2, be2 = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV), (0.1 * torch.randn(D, generator=g)).to(DEV)
    xa, w1a, w2a = ops.cast(x, "bf16"), ops.cast(w1, "bf16"), ops.cast(w2, "bf16")
    h = xa.double() @ w1a.double().t() + b1.double()
    h = (h * torch.sigmoid(h)).float().bfloat16().double()          # the hidden activation is a bf16 operand of the second product
    delta_ref = 0.5 * (h @ w2a.double().t() + b2.double())
    x_ref = res.double() + delta_ref
    ln = lambda t, gg, bb: torch.nn.functional.layer_norm(t, (D,), gg.double(), bb.double(), 1e-6)
    y_ref = ln(x_ref, g1, be1)
    first = None
    for cs in clusters:
        out, y = ops.ffn_fused(xa, w1a, b1, w2a, b2, res, g1, be1, mode=1, cluster=cs)
        assert rel_l2(out.double() - res.double(), delta_ref) < 6e-3, f"ffn delta, cluster {cs}"
        assert rel_l2(out, x_ref) < 2e-3
        assert rel_l2(y.float(), y_ref) < 5e-3
        out2, y2 = ops.ffn_fused(xa, w1a, b1, w2a, b2, res, g1, be1, g2, be2, mode=2, cluster=cs)
        assert rel_l2(out2, y_ref) < 2e-3
        assert rel_l2(y2.float(), ln(y_ref, g2, be2)) < 5e-3
        out3, y3 = ops.ffn_fused(xa, w1a, b1, w2a, b2, res, g1, be1, None, None, mode=2, cluster=cs)
        assert torch.equal(out3, out2)
        assert torch.equal(y3.float().cpu(), rnd("bf16", out3.cpu()))
        out4, y4 = ops.ffn_fused(xa, w1a, b1, w2a, b2, res, g1, be1, None, None, mode=2, cluster=cs, want_ln=False)
        assert y4 is None and torch.equal(out4, out2)
        # in-place operand / LayerNorm output buffer, as the engine uses it
        xin = xa.clone()
        M_, D_ = xin.shape
        from efficientconformer_b200 import _lib
        of = torch.empty(M_, D_, dtype=torch.float32, device=DEV)
        _lib.check(_lib.lib().ec_op_ffn(xin.data_ptr(), w1a.data_ptr(), b1.data_ptr(), w2a.data_ptr(), b2.data_ptr(), M_, D_, hidden,
                                        res.data_ptr(), of.data_ptr(), 1, g1.data_ptr(), be1.data_ptr(), None, None, 1e-6, xin.data_ptr(),
                                        cs, torch.cuda.current_stream().cuda_stream))
        assert torch.equal(of, out) and torch.equal(xin, y)
        if first is None:
            first = out
        else:     # the split over the cluster changes the summation order of the second product only
            assert rel_l2(out, first) < 1e-5


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("M,C,K", [(256, 120, 120), (1000, 168, 120), (517, 240, 168), (64, 8, 16), (300, 360, 360)])
def test_pointwise_glu(ops, prec, M, C, K):
    g = torch.Generator(device="cpu").manual_seed(C + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(2 * C, K, 1, generator=g) / math.sqrt(K)).to(DEV)
    b = torch.randn(2 * C, generator=g).to(DEV)
    aa = ops.cast(a, prec)
    h = ops.unpack(aa, prec).double() @ rnd(prec, w[:, :, 0].cpu()).to(DEV).double().t() + b.double()
    ref = h[:, :C] * torch.sigmoid(h[:, C:])
    out = ops.pointwise_glu(aa, w, b, prec)
    tol = 2e-6 if prec != "bf16" else 4e-3     # output is rounded to the activation type
    out = ops.unpack(out, prec)
    assert rel_l2(out, ref) < (1e-3 if prec != "bf16" else 4e-3)
    assert rel_l2(out, rnd(prec, ref.float().cpu())) < 1e-4 + tol


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("rows,dim", [(16000, 120), (1000, 168), (37, 240), (5, 1024), (3, 8)])
def test_layernorm(ops, prec, rows, dim):
    g = torch.Generator(device="cpu").manual_seed(rows + dim)
    x = (torch.randn(rows, dim, generator=g) * 2 + 0.5).to(DEV)
    gamma = (1 + 0.1 * torch.randn(dim, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(dim, generator=g)).to(DEV)
    ya, yf = ops.layernorm(x, gamma, beta, prec)
    ref = torch.nn.functional.layer_norm(x.double(), (dim,), gamma.double(), beta.double(), 1e-6)
    assert rel_l2(yf, ref) < 1e-6
    assert torch.equal(ops.unpack(ya, prec).cpu(), rnd(prec, yf.cpu()))


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,T,C,k,stride", [(4, 500, 120, 15, 1), (3, 251, 168, 15, 2), (2, 250, 240, 15, 2), (2, 1, 120, 15, 1),
                                            (2, 7, 240, 15, 2), (1, 130, 256, 31, 1), (2, 65, 176, 31, 2)])
def test_dwconv_bn_swish(ops, prec, B, T, C, k, stride):
    g = torch.Generator(device="cpu").manual_seed(T + C + k)
    x = torch.randn(B, T, C, generator=g).to(DEV)
    w = (torch.randn(C, 1, k, generator=g) / math.sqrt(k)).to(DEV)
    b, gam, bet = (0.1 * torch.randn(C, generator=g)).to(DEV), (1 + 0.1 * torch.randn(C, generator=g)).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
    rm, rv = (0.1 * torch.randn(C, generator=g)).to(DEV), (0.5 + torch.rand(C, generator=g)).to(DEV)
    xa = ops.cast(x, prec)
    wf, bf = ops.fold_bn(w, b, gam, bet, rm, rv)
    y = ops.dwconv_bn_swish(xa, wf, bf, stride, prec)
    xin = ops.unpack(xa, prec).double().transpose(1, 2)
    pad = (k - 1) // 2
    conv = torch.nn.functional.conv1d(torch.nn.functional.pad(xin, (pad, pad)), w.double(), b.double(), stride=stride, groups=C)
    bn = (conv - rm.double()[None, :, None]) / torch.sqrt(rv.double()[None, :, None] + 1e-5) * gam.double()[None, :, None] + bet.double()[None, :, None]
    ref = (bn * torch.sigmoid(bn)).transpose(1, 2)
    assert y.shape == ref.shape
    assert rel_l2(ops.unpack(y, prec), ref) < (5e-4 if prec != "bf16" else 4e-3)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,T", [(2, 101), (3, 64), (1, 1), (2, 2), (2, 33)])
def test_subsample_conv(ops, prec, B, T):
    C, F = 120, 80
    g = torch.Generator(device="cpu").manual_seed(T)
    mel = torch.randn(B, F, T, generator=g).to(DEV)
    w = (torch.randn(C, 1, 3, 3, generator=g) / 3).to(DEV)
    b, gam, bet = (0.1 * torch.randn(C, generator=g)).to(DEV), (1 + 0.1 * torch.randn(C, generator=g)).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
    rm, rv = (0.1 * torch.randn(C, generator=g)).to(DEV), (0.5 + torch.rand(C, generator=g)).to(DEV)
    wf, bf = ops.fold_bn(w, b, gam, bet, rm, rv)
    y = ops.subsample_conv(mel, wf, bf, prec)
    conv = torch.nn.functional.conv2d(mel.double().unsqueeze(1), w.double(), b.double(), stride=2, padding=1)
    bn = torch.nn.functional.batch_norm(conv, rm.double(), rv.double(), gam.double(), bet.double(), False, 0.0, 1e-5)
    ref = (bn * torch.sigmoid(bn)).reshape(B, C * (F // 2), -1).transpose(1, 2)
    assert y.shape == ref.shape
    assert rel_l2(ops.unpack(y, prec), ref) < (5e-4 if prec != "bf16" else 4e-3)


def _attention_reference(qkv, E, u, v, x_len, H, G):
    """fp64 closed form of SURVEY.md section 8 row a9 on the given (already projected) q|k|v and E."""
    B, T, D3 = qkv.shape
    D = D3 // 3
    d = G * D // H
    q, k, vv = qkv.double().split(D, dim=-1)
    Pd = (-T) % G
    if Pd:
        q, k, vv = (torch.nn.functional.pad(t, (0, 0, 0, Pd)) for t in (q, k, vv))
    Tg = (T + Pd) // G
    qu = (q + u.double()).reshape(B, Tg, H, d).transpose(1, 2)
    qv = (q + v.double()).reshape(B, Tg, H, d).transpose(1, 2)
    kk = k.reshape(B, Tg, H, d).transpose(1, 2)
    vh = vv.reshape(B, Tg, H, d).transpose(1, 2)
    Eh = E.double().reshape(2 * Tg - 1, H, d).transpose(0, 1)
    sk = qu @ kk.transpose(2, 3)
    se = qv @ Eh.transpose(1, 2).unsqueeze(0)
    idx = (Tg - 1) + torch.arange(Tg, device=qkv.device)[None, :] - torch.arange(Tg, device=qkv.device)[:, None]
    se = torch.gather(se, 3, idx[None, None].expand(B, H, Tg, Tg))
    s = (sk + se) / d ** 0.5
    if x_len is not None:
        masked = (torch.arange(Tg, device=qkv.device) * G)[None, :] >= x_len[:, None]
        s = s.masked_fill(masked[:, None, None, :], float("-inf"))
    w = s.softmax(-1)
    return (w @ vh).transpose(1, 2).reshape(B, Tg * G, D)[:, :T]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,T,D,H,G", [(2, 500, 120, 4, 3), (2, 251, 120, 4, 3), (3, 250, 168, 4, 1), (2, 125, 240, 4, 1), (2, 1, 120, 4, 3),
                                       (2, 2, 120, 4, 3), (1, 64, 168, 4, 1), (2, 65, 240, 4, 1), (1, 700, 168, 4, 1), (2, 33, 120, 8, 3)])
def test_relpos_attention(ops, prec, B, T, D, H, G):
    if prec == "bf16" and ((G * D) // H) % 2:
        pytest.skip("the bf16 attention kernel needs an even head dim (all CTC/Transducer Small/Medium configs)")
    g = torch.Generator(device="cpu").manual_seed(T * 3 + D)
    # contract: q|k|v and E arrive TF32-rounded (the producing GEMM epilogues round their fp32 output, round_out)
    qkv = rnd(prec, torch.randn(B, T, 3 * D, generator=g)).to(DEV)
    Tp = T + (-T) % G
    E = rnd(prec, torch.randn(2 * Tp - G, D, generator=g)).to(DEV)
    u, v = (0.3 * torch.randn(D, generator=g)).to(DEV), (0.3 * torch.randn(D, generator=g)).to(DEV)
    x_len = torch.tensor([T] + [max(1, (2 * T) // 3)] * (B - 1), device=DEV)
    for xl in (x_len, None):
        out = ops.relpos_attention(qkv, E, u, v, xl, H, G, prec)
        ref = _attention_reference(qkv, E, u, v, xl, H, G)
        assert out.shape == ref.shape
        err = rel_l2(ops.unpack(out, prec), ref)
        assert err < (2e-3 if prec != "bf16" else 1e-2), (err, xl is None)      # split mode: fp16 core = TF32-grade accuracy      # bf16 path also rounds qu/qv, P and the output to bf16


def test_ctc_loss_and_greedy_on_device(golden_dir):
    import os
    from efficientconformer_b200.model_ctc import ctc_loss, greedy_ids
    g = torch.load(os.path.join(golden_dir, "ctc_loss_small.pt"))
    mean, per = ctc_loss(g["logits"].to(DEV), g["logits_len"], g["targets"], g["target_len"])
    assert torch.allclose(per.cpu(), g["loss_per_utt"], rtol=2e-5, atol=2e-5)
    assert abs(float(mean) - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))
    from oracle import conformer_oracle as O
    assert greedy_ids(g["logits"].to(DEV), g["logits_len"]) == O.greedy_ids(g["logits"], g["logits_len"])
    # larger random case against the oracle restatement
    gen = torch.Generator().manual_seed(5)
    lg = 2 * torch.randn(8, 125, 256, generator=gen)
    ll = torch.tensor([125, 120, 99, 64, 33, 17, 125, 2])
    U = torch.clamp(ll // 3, min=1)
    y = torch.randint(1, 256, (8, int(U.max())), generator=gen)
    mean, per = ctc_loss(lg.to(DEV), ll, y, U)
    m2, p2 = O.ctc_loss(lg, ll, y, U)
    assert torch.allclose(per.cpu(), p2, rtol=1e-4, atol=1e-4)
    assert greedy_ids(lg.to(DEV), ll) == O.greedy_ids(lg, ll)
