"""Parity of the backward kernels (training step) on the B200, each through the C ABI against
torch.autograd over the CPU oracle's restatement of the same forward op (the oracle's autograd is pinned against the real
reference's loss.backward() in tests/test_oracle_golden.py::test_training_step_loss_gradients_and_running_stats)."""
import os
import random

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _oracle_ctc_grad(logits, ll, y, yl):
    from oracle import conformer_oracle as O
    lg = logits.double().clone().requires_grad_(True)
    loss, per = O.ctc_loss(lg, ll, y, yl)
    loss.backward()
    return loss.detach(), per.detach(), lg.grad


def test_ctc_gradient_known_answers(golden_dir):
    """The reference's CTC known answers (repeated labels, ragged lengths): loss and d loss / d logits."""
    from efficientconformer_b200.model_ctc import ctc_loss_and_grad
    g = torch.load(os.path.join(golden_dir, "ctc_loss_small.pt"))
    mean, per, grad = ctc_loss_and_grad(g["logits"].to(DEV), g["logits_len"], g["targets"], g["target_len"])
    ref_mean, ref_per, ref_grad = _oracle_ctc_grad(g["logits"], g["logits_len"], g["targets"], g["target_len"])
    assert torch.allclose(per.cpu(), g["loss_per_utt"], rtol=2e-5, atol=1e-4)
    assert rel_l2(grad, ref_grad) < 2e-5
    for b in range(g["logits"].shape[0]):                       # padded frames get exactly zero gradient
        assert float(grad[b, int(g["logits_len"][b]):].abs().max() if int(g["logits_len"][b]) < grad.shape[1] else 0.0) == 0.0
    # every frame's gradient sums to zero over the classes (softmax minus a distribution), up to fp32 rounding
    assert float(grad.sum(-1).abs().max()) < 1e-5


def test_ctc_gradient_random_shapes():
    """Seeded sweep: T up to 250, U up to 100 (1..7 extended states per lane), repeats, single-label and U = T/2 cases,
    peaky logits; plus the autograd node used by LossCTC."""
    from efficientconformer_b200.model_ctc import ctc_loss_and_grad, LossCTC
    rng = random.Random(11)
    for trial in range(10):
        B = rng.choice([1, 3, 8])
        T = rng.choice([1, 2, 17, 63, 125, 200, 250])
        V = rng.choice([5, 32, 256])
        g = torch.Generator().manual_seed(900 + trial)
        scale = rng.choice([1.0, 4.0, 12.0])
        logits = scale * torch.randn(B, T, V, generator=g)
        ll = torch.tensor([rng.randint(1, T) for _ in range(B)])
        ll[0] = T
        Umax = max(1, min(100, T // 2))
        yl = torch.tensor([rng.randint(1, max(1, min(Umax, int(ll[b]) // 2))) for b in range(B)])
        y = torch.randint(1, min(V, 4) if trial % 3 == 0 else V, (B, int(yl.max())), generator=g)   # small alphabets force repeats
        mean, per, grad = ctc_loss_and_grad(logits.to(DEV), ll, y, yl)
        ref_mean, ref_per, ref_grad = _oracle_ctc_grad(logits, ll, y, yl)
        finite = torch.isfinite(ref_per)
        assert torch.equal(torch.isfinite(per.cpu()), finite), (trial, per, ref_per)
        assert torch.allclose(per.cpu()[finite], ref_per.float()[finite], rtol=1e-4, atol=1e-3), (trial, per, ref_per)
        if bool(finite.all()):
            # fp32 log-domain recursion (like torch's own CUDA CTC): with peaky logits log p(l|x) reaches -1e4 and one fp32 ulp of
            # alpha + beta is ~1e-3 in the exponent of the occupancy -> the north-star tolerance (1e-3) for the stress cases
            tol = 1e-4 if scale == 1.0 else 1.5e-3
            assert rel_l2(grad, ref_grad) < tol, (trial, B, T, V, scale, rel_l2(grad, ref_grad))
    # long transcripts / long utterances: 9..16 states per lane (128..255 labels), and T x (2U+1) beyond the shared-memory
    # budget (emissions kept in the L2 scratch) -- every batch nn.CTCLoss accepts at the shipped 16 s / vocab-256 settings
    for trial, (B, T, U, V) in enumerate([(2, 201, 150, 256), (3, 260, 255, 64), (2, 640, 120, 256), (1, 520, 200, 32), (2, 300, 128, 256)]):
        g = torch.Generator().manual_seed(1900 + trial)
        logits = torch.randn(B, T, V, generator=g)
        ll = torch.tensor([T] + [rng.randint(2 * U if 2 * U <= T else T, T) for _ in range(B - 1)])
        yl = torch.tensor([U] + [rng.randint(1, max(1, min(U, int(ll[b]) // 2))) for b in range(1, B)])
        y = torch.randint(1, V, (B, U), generator=g)
        mean, per, grad = ctc_loss_and_grad(logits.to(DEV), ll, y, yl)
        ref_mean, ref_per, ref_grad = _oracle_ctc_grad(logits, ll, y, yl)
        finite = torch.isfinite(ref_per)
        assert torch.equal(torch.isfinite(per.cpu()), finite), (trial, per, ref_per)
        assert torch.allclose(per.cpu()[finite], ref_per.float()[finite], rtol=1e-4, atol=1e-3), (trial, per, ref_per)
        if bool(finite.all()):
            # fp32 log-domain recursion: |log2 p(l|x)| reaches 5e3 at T = 640, where one fp32 ulp of alpha + beta is 5e-4 in the exponent
            # of the occupancies (torch's own fp32 CUDA CTC has the same floor); the loss itself is held to 1e-4 above
            assert rel_l2(grad, ref_grad) < (2e-4 if T <= 250 else 3e-3), (trial, B, T, U, V, rel_l2(grad, ref_grad))
    # autograd node: loss.backward() through LossCTC fills logits.grad
    logits = torch.randn(2, 40, 16, generator=torch.Generator().manual_seed(5)).to(DEV).requires_grad_(True)
    ll, yl = torch.tensor([40, 31]), torch.tensor([7, 5])
    y = torch.randint(1, 16, (2, 7), generator=torch.Generator().manual_seed(6))
    loss = LossCTC()((None, y, None, yl), (logits, ll, None))
    (3.0 * loss).backward()
    _, _, ref_grad = _oracle_ctc_grad(logits.detach().cpu(), ll, y, yl)
    assert rel_l2(logits.grad, 3.0 * ref_grad) < 1e-4


@pytest.fixture(scope="module")
def ops():
    from efficientconformer_b200 import ops as o
    return o


def test_layernorm_backward(ops):
    rng = random.Random(21)
    for trial in range(8):
        rows = rng.choice([1, 7, 64, 1000, 4001, 16000])
        dim = rng.choice([100, 120, 168, 240, 256, 360, 512, 720])
        g = torch.Generator().manual_seed(700 + trial)
        x = (2.0 * torch.randn(rows, dim, generator=g) + 0.7)
        dy = torch.randn(rows, dim, generator=g)
        gamma = 1 + 0.2 * torch.randn(dim, generator=g)
        beta = 0.1 * torch.randn(dim, generator=g)
        xr = x.double().requires_grad_(True); gr = gamma.double().requires_grad_(True); br = beta.double().requires_grad_(True)
        torch.nn.functional.layer_norm(xr, (dim,), gr, br, 1e-6).backward(dy.double())
        dx, dg, db = ops.layernorm_bwd(x.to(DEV), dy.to(DEV), gamma.to(DEV))
        assert rel_l2(dx, xr.grad) < 2e-5, (trial, rows, dim)
        assert rel_l2(dg, gr.grad) < 2e-5 and rel_l2(db, br.grad) < 2e-5, (trial, rows, dim)
        # residual-branch accumulation and bit reproducibility
        acc = torch.randn(rows, dim, generator=g).to(DEV)
        ref_acc = acc.double().cpu() + xr.grad
        dx2, dg2, db2 = ops.layernorm_bwd(x.to(DEV), dy.to(DEV), gamma.to(DEV), dx_accum=acc)
        assert rel_l2(dx2, ref_acc) < 2e-5
        assert torch.equal(dg, dg2) and torch.equal(db, db2)
        # the emitted operand of the next GEMM: act(scale * dropout mask * accumulated dx) == the separate ec_op_dropout pass on dx
        if dim % 4 == 0:
            for prec in ("bf16x2", "bf16", "tf32"):
                if prec == "bf16" and dim % 8:
                    continue
                for p_drop in (0.0, 0.1):
                    drop = type("D", (), {"p": p_drop, "counter": ops.dropout_counter(DEV, 5) if p_drop > 0 else None})()
                    acc3 = acc.clone() if trial % 2 else None
                    dx3, dg3, db3, em = ops.layernorm_bwd(x.to(DEV), dy.to(DEV), gamma.to(DEV), dx_accum=acc3, emit=(prec, 0.5, drop, 13))
                    ref_em = ops.dropout_cast_scaled(dx3, prec, 0.5, drop, 13)
                    assert torch.equal(ops.unpack(em, prec), ops.unpack(ref_em, prec)), (trial, rows, dim, prec, p_drop)
                    assert torch.equal(dg3, dg) and torch.equal(db3, db)


@pytest.mark.parametrize("prec", ["bf16x2", "tf32", "bf16"])
def test_linear_data_gradient_bias_gradient_and_activations(ops, prec):
    """dX = dY . W on the forward tcgen05 kernel with the transposed weight copy; bias gradient = column sums;
    Swish / GLU backward."""
    tol = 1e-3 if prec != "bf16" else 8e-3
    rng = random.Random(31)
    for trial in range(6):
        M = rng.choice([5, 128, 1000, 4000, 16000])
        N, K = rng.choice([(480, 120), (120, 480), (360, 120), (240, 240), (256, 240), (2880, 720)])
        g = torch.Generator().manual_seed(800 + trial)
        dy = torch.randn(M, N, generator=g)
        w = torch.randn(N, K, generator=g) / N ** 0.5
        res = torch.randn(M, K, generator=g)
        dy_act = ops.cast(dy.to(DEV), prec)
        dx = ops.linear_dgrad(dy_act, w.to(DEV), prec, residual=res.to(DEV))
        dyv = ops.unpack(dy_act, prec).double().cpu()
        ref = dyv @ w.double() + res.double()
        assert rel_l2(dx, ref) < tol, (trial, M, N, K)
        bsum = ops.colsum(dy_act, prec)
        assert rel_l2(bsum, dyv.sum(0)) < 1e-5
        assert rel_l2(ops.colsum(dy.to(DEV), prec, is_f32=True), dy.double().sum(0)) < 1e-5
        # activations
        z = ops.cast((2 * torch.randn(M, N, generator=g)).to(DEV), prec)
        zr = ops.unpack(z, prec).double().cpu().requires_grad_(True)
        (zr * torch.sigmoid(zr)).backward(dy.double())
        assert rel_l2(ops.unpack(ops.swish_bwd(z, dy.to(DEV), prec), prec), zr.grad) < (5.2e-4 if prec != "bf16" else 4e-3)
        C = N // 2
        zg = ops.cast((2 * torch.randn(M, 2 * C, generator=g)).to(DEV), prec)
        zgr = ops.unpack(zg, prec).double().cpu().requires_grad_(True)
        (zgr[:, :C] * torch.sigmoid(zgr[:, C:])).backward(dy[:, :C].double())
        assert rel_l2(ops.unpack(ops.glu_bwd(zg, dy[:, :C].contiguous().to(DEV), prec), prec), zgr.grad) < (6e-4 if prec != "bf16" else 4e-3)


@pytest.mark.parametrize("prec", ["bf16x2", "bf16", "tf32"])
def test_linear_weight_gradient_tcgen05(ops, prec):
    """dW = dY^T . X with MN-major tcgen05 operands and split-M reduction (+ the bias gradient from the ones operand of the same kernel),
    against fp64 on the same rounded operands.  bf16x2: ONE pass over the packed operands (2 x 2 accumulator blocks summed in the epilogue)."""
    rng = random.Random(41)
    shapes = [(16000, 480, 120), (16000, 120, 480), (8000, 504, 168), (4000, 240, 960), (4000, 256, 240), (1000, 120, 4800),
              (77, 120, 120), (1, 360, 120), (130, 2880, 720), (16000, 360, 120)]
    for trial, (M, N, K) in enumerate(shapes):
        g = torch.Generator().manual_seed(1000 + trial)
        dy = ops.cast(torch.randn(M, N, generator=g).to(DEV), prec)
        x = ops.cast(torch.randn(M, K, generator=g).to(DEV), prec)
        dw = ops.linear_wgrad(dy, x, prec)
        dyv, xv = ops.unpack(dy, prec).double().cpu(), ops.unpack(x, prec).double().cpu()
        ref = dyv.t() @ xv
        assert rel_l2(dw, ref) < 2e-5, (prec, M, N, K, rel_l2(dw, ref))
        dwb, db = ops.linear_wgrad_bias(dy, x, prec)
        assert torch.equal(dwb, dw), (prec, M, N, K)
        assert rel_l2(db, dyv.sum(0)) < 2e-5, (prec, M, N, K, rel_l2(db, dyv.sum(0)))
        acc = torch.randn(N, K, generator=g).to(DEV)
        ref2 = acc.double().cpu() + ref
        dw2 = ops.linear_wgrad(dy, x, prec, dw_accum=acc)
        assert rel_l2(dw2, ref2) < 2e-5
        assert torch.equal(ops.linear_wgrad(dy, x, prec), dw)                # fixed-order reduction: bit reproducible


@pytest.mark.parametrize("prec", ["bf16x2", "tf32", "bf16"])
def test_dwconv_batchnorm_swish_training_forward_backward(ops, prec):
    """Train-mode depthwise conv -> BatchNorm1d (batch statistics incl. padded frames, running-stat update) -> Swish, forward and
    every gradient, against autograd over torch's own conv1d / batch_norm in fp64 on the same rounded input."""
    import torch.nn.functional as F
    rng = random.Random(51)
    for trial in range(8):
        C = rng.choice([8, 120, 168, 240, 360, 720])
        k = rng.choice([15, 31])
        stride = rng.choice([1, 2])
        T = rng.choice([1, 2, 13, 64, 127, 500]) if trial else 500
        B = rng.choice([1, 3, 32]) if T < 500 else 4
        g = torch.Generator().manual_seed(1200 + trial)
        x = ops.cast(torch.randn(B, T, C, generator=g).to(DEV), prec)
        w = (torch.randn(C, 1, k, generator=g) / k ** 0.5)
        b = 0.1 * torch.randn(C, generator=g)
        gam, bet = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
        rm, rv = 0.1 * torch.randn(C, generator=g), 0.5 + torch.rand(C, generator=g)
        To = (T - 1) // stride + 1
        dh = torch.randn(B, To, C, generator=g)
        # fp64 reference
        xr = ops.unpack(x, prec).double().cpu().requires_grad_(True)
        wr, br, gr, ber = [t.double().clone().requires_grad_(True) for t in (w, b, gam, bet)]
        rm_r, rv_r = rm.double().clone(), rv.double().clone()
        pad = (k - 1) // 2
        conv = F.conv1d(F.pad(xr.transpose(1, 2), (pad, pad)), wr, br, stride=stride, groups=C)
        if B * To > 1:
            bn = F.batch_norm(conv, rm_r, rv_r, gr, ber, training=True, momentum=0.1, eps=1e-5)
        else:
            continue                                   # torch refuses a single value per channel in training mode
        out = (bn * torch.sigmoid(bn)).transpose(1, 2)
        out.backward(dh.double())
        rm_d, rv_d = rm.clone().to(DEV), rv.clone().to(DEV)
        h, saved = ops.DwConvTrain.forward(x, w.to(DEV), b.to(DEV), gam.to(DEV), bet.to(DEV), rm_d, rv_d, stride, prec)
        tol_h = 5e-4 if prec != "bf16" else 5e-3
        assert rel_l2(ops.unpack(h, prec), out.detach()) < tol_h, (trial, B, T, C, k, stride)
        assert rel_l2(rm_d, rm_r) < 1e-5 and rel_l2(rv_d, rv_r) < 1e-4
        dx, dw, db, dgam, dbet = ops.DwConvTrain.backward(dh.to(DEV), saved)
        case = (trial, B, T, C, k, stride)
        assert rel_l2(dx, xr.grad) < 2e-4, case
        assert rel_l2(dw, wr.grad[:, 0, :]) < 2e-4, case
        assert rel_l2(dgam, gr.grad) < 2e-4 and rel_l2(dbet, ber.grad) < 2e-4, case
        assert float(db.abs().max()) < 1e-3 * max(1.0, float(dbet.abs().max())), case     # exactly zero in exact arithmetic


@pytest.mark.parametrize("prec", ["bf16x2", "tf32", "bf16"])
def test_relpos_attention_backward(ops, prec):
    """Gradients of the (grouped) relative-position attention core w.r.t. q|k|v, the projected positional rows E and the u / v
    biases, against autograd over the fp64 closed form on the same rounded operands: every shipped
    head layout, padded groups (T % G != 0), ragged key masks."""
    from test_gpu_ops import _attention_reference
    rng = random.Random(61)
    layouts = [(120, 4, 3), (168, 4, 1), (240, 4, 1), (360, 8, 3), (176, 4, 1), (144, 6, 1), (100, 4, 3), (512, 8, 1)]
    for trial in range(10):
        D, H, G = layouts[trial % len(layouts)]
        T = rng.choice([1, 2, 5, 17, 64, 125, 250, 251, 500])
        B = rng.choice([1, 2, 3])
        if prec == "bf16" and D % 8:
            continue
        g = torch.Generator().manual_seed(1500 + trial)
        # split mode: plain fp16 operands where the 16-bit kernels apply
        qkv = ops.cast_attn_operand(torch.randn(B, T, 3 * D, generator=g).to(DEV), prec, D, H, G)
        Tp = T + (-T) % G
        E = ops.cast_attn_operand(torch.randn(2 * Tp - G, D, generator=g).to(DEV), prec, D, H, G)
        u, v = 0.3 * torch.randn(D, generator=g), 0.3 * torch.randn(D, generator=g)
        x_len = torch.tensor([rng.randint(1, T) for _ in range(B)])
        x_len[0] = T
        d_out = torch.randn(B, T, D, generator=g)
        qr, Er = ops.attn_operand_values(qkv, prec).double().cpu().requires_grad_(True), ops.attn_operand_values(E, prec).double().cpu().requires_grad_(True)
        ur, vr = u.double().requires_grad_(True), v.double().requires_grad_(True)
        _attention_reference(qr, Er, ur, vr, x_len, H, G).backward(d_out.double())
        dqkv, dE, du, dv = ops.relpos_attention_bwd(qkv, E, u.to(DEV), v.to(DEV), x_len.to(DEV), H, G, d_out.to(DEV), prec)
        case = (prec, trial, B, T, D, H, G)
        # tf32 mode: fp32 CUDA-core kernels (parity path).  bf16 mode: batched tensor-core GEMMs with bf16 Qu / Qv / P / dS operands
        # (fp32 accumulation): the operand rounding (2^-9 relative per element) bounds the error, as in the forward kernel.
        tol = 2e-5 if prec == "tf32" else 1.5e-2      # bf16x2: the packed operands are rounded to bf16 for the attention backward
        errs = [rel_l2(dqkv, qr.grad), rel_l2(dE, Er.grad), rel_l2(du, ur.grad), rel_l2(dv, vr.grad)]
        print(case, ["%.2e" % e for e in errs])
        assert max(errs) < tol, (case, errs)


@pytest.mark.parametrize("prec", ["bf16x2", "tf32", "bf16"])
def test_relpos_attention_backward_activation_type_output(ops, prec):
    """ec_op_relpos_attention_bwd_act writes dq | dk | dv straight in the activation type: equal to casting the fp32 result, for the
    paired (even dims) and the scalar (odd head dim) unpack kernels."""
    for (B, T, D, H, G) in ((3, 50, 120, 4, 3), (2, 33, 168, 4, 1), (2, 40, 100, 4, 3)):
        if prec == "bf16" and D % 8:
            continue
        g = torch.Generator().manual_seed(B * 100 + T)
        qkv = ops.cast_attn_operand(torch.randn(B, T, 3 * D, generator=g).to(DEV), prec, D, H, G)
        Tp = T + (-T) % G
        E = ops.cast_attn_operand((0.3 * torch.randn(2 * Tp - G, D, generator=g)).to(DEV), prec, D, H, G)
        u, v = (0.2 * torch.randn(D, generator=g)).to(DEV), (0.2 * torch.randn(D, generator=g)).to(DEV)
        x_len = torch.tensor([T] + [max(1, T - 7 * b) for b in range(1, B)], device=DEV)
        dout = torch.randn(B, T, D, generator=g).to(DEV)
        dqkv, dE, du, dv = ops.relpos_attention_bwd(qkv, E, u, v, x_len, H, G, dout, prec)
        dq2, dE2, du2, dv2 = ops.relpos_attention_bwd_act(qkv, E, u, v, x_len, H, G, dout, prec)
        assert torch.equal(ops.unpack(dq2, prec), ops.unpack(ops.cast(dqkv, prec), prec)), (prec, B, T, D, H, G)
        assert torch.equal(dE, dE2) and torch.equal(du, du2) and torch.equal(dv, dv2)


@pytest.mark.parametrize("prec", ["bf16x2", "tf32", "bf16"])
def test_subsampling_conv2d_batchnorm2d_training_forward_backward(ops, prec):
    """Train-mode Conv2d(1->C,3x3,s2) -> BatchNorm2d (batch statistics) -> Swish in the layout of the following Linear, and the
    weight / bias / BatchNorm gradients, against fp64 autograd over torch's conv2d / batch_norm."""
    import torch.nn.functional as F
    rng = random.Random(71)
    for trial in range(5):
        C = rng.choice([8, 120, 180])
        T = rng.choice([2, 7, 64, 101, 500])
        B = rng.choice([1, 2, 4])
        Fm = 80
        g = torch.Generator().manual_seed(1700 + trial)
        mel = torch.randn(B, Fm, T, generator=g)
        w = torch.randn(C, 1, 3, 3, generator=g) / 3
        b = 0.1 * torch.randn(C, generator=g)
        gam, bet = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
        rm, rv = 0.1 * torch.randn(C, generator=g), 0.5 + torch.rand(C, generator=g)
        To = (T - 1) // 2 + 1
        da = torch.randn(B * To, C * (Fm // 2), generator=g)
        wr, br, gr, ber = [t.double().clone().requires_grad_(True) for t in (w, b, gam, bet)]
        rm_r, rv_r = rm.double().clone(), rv.double().clone()
        conv = F.conv2d(mel.double().unsqueeze(1), wr, br, stride=2, padding=1)             # (B, C, F/2, T/2)
        bn = F.batch_norm(conv, rm_r, rv_r, gr, ber, training=True, momentum=0.1, eps=1e-5)
        out = (bn * torch.sigmoid(bn)).reshape(B, C * (Fm // 2), To).transpose(1, 2).reshape(B * To, -1)
        out.backward(da.double())
        rm_d, rv_d = rm.clone().to(DEV), rv.clone().to(DEV)
        a, saved = ops.SubsampleTrain.forward(mel.to(DEV), w.to(DEV), b.to(DEV), gam.to(DEV), bet.to(DEV), rm_d, rv_d, prec)
        case = (prec, trial, B, T, C)
        assert rel_l2(ops.unpack(a, prec), out.detach()) < (5e-4 if prec != "bf16" else 5e-3), case
        assert rel_l2(rm_d, rm_r) < 1e-5 and rel_l2(rv_d, rv_r) < 1e-4, case
        dw, db, dgam, dbet = ops.SubsampleTrain.backward(da.to(DEV), saved)
        assert rel_l2(dw, wr.grad) < 2e-4, (case, rel_l2(dw, wr.grad))
        assert rel_l2(dgam, gr.grad) < 2e-4 and rel_l2(dbet, ber.grad) < 2e-4, case
        assert float(db.abs().max()) < 1e-3 * max(1.0, float(dbet.abs().max())), case
