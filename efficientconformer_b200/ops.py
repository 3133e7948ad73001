"""Tensor-level wrappers of the single-operator C entry points (ec_op_*), used by the unit parity tests.
Every function takes/returns CUDA tensors and launches on the current stream; nothing here has a fallback."""
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr, act_dtype, weight_planes, PRECISIONS, PREC_BF16X2


def _p(precision):
    return PRECISIONS[precision] if isinstance(precision, str) else precision


def cast(x, precision):
    """fp32 -> activation type (TF32-rounded fp32 or bf16)."""
    pr = _p(precision)
    x = x.float().contiguous()
    out = torch.empty(x.shape, dtype=act_dtype(pr), device=x.device)
    check(lib().ec_op_cast(pr, ptr(x), ptr(out), x.numel(), stream_ptr()))
    return out


def cast_weight(w, precision):
    """fp32 [N, K] -> GEMM weight operand.  Returns the [N, K] plane-0 view of a [planes, N, K] buffer: in the split mode the plane
    with the swapped halves sits right behind it in the same storage (which the view keeps alive) -- the layout ec_op_gemm expects."""
    pr = _p(precision)
    w = w.float().contiguous()
    out = torch.empty((weight_planes(pr),) + tuple(w.shape), dtype=act_dtype(pr), device=w.device)
    check(lib().ec_op_cast_weight(pr, ptr(w), w.numel(), ptr(out), stream_ptr()))
    return out[0]


def unpack(x_act, precision):
    """Values of an activation-type tensor as fp32 (tests / debugging): identity for TF32 words, hi + lo for split-mode pairs."""
    pr = _p(precision)
    if pr != PREC_BF16X2:
        return x_act.float()
    bits = x_act.contiguous().view(torch.int32)
    hi = (bits << 16).view(torch.float32)
    lo = (bits & -65536).view(torch.float32)
    return hi + lo


def layernorm(x, gamma, beta, precision, eps=1e-6, want_f32=True, want_act=True):
    pr = _p(precision)
    x = x.float().contiguous()
    rows, dim = x.numel() // x.shape[-1], x.shape[-1]
    ya = torch.empty(x.shape, dtype=act_dtype(pr), device=x.device) if want_act else None
    yf = torch.empty_like(x) if want_f32 else None
    check(lib().ec_op_layernorm(pr, ptr(x), rows, dim, ptr(gamma.float().contiguous()), ptr(beta.float().contiguous()), eps,
                                ptr(ya), ptr(yf), stream_ptr()))
    return ya, yf


def attn_operands_f16(precision, dim, heads, group):
    """True when the q|k|v / E operands of the attention core are plain fp16 tensors: split mode with a head layout the 16-bit
    attention kernels support (ec_attention_operand_kind == 2); otherwise they are in the activation type."""
    return lib().ec_attention_operand_kind(_p(precision), dim, heads, group) == 2


def cast_attn_operand(x, precision, dim, heads, group):
    """fp32 -> the storage the attention entry points expect for q|k|v / E in this mode (tests)."""
    if attn_operands_f16(precision, dim, heads, group):
        return x.float().to(torch.float16).contiguous()
    return cast(x, precision)


def attn_operand_values(x, precision):
    """fp32 values of an attention operand tensor (tests)."""
    return x.float() if x.dtype == torch.float16 else unpack(x, precision)


def gemm(a_act, w_act, bias, precision, alpha=1.0, act=0, residual=None, want_f32=True, want_act=False, act_f16=False):
    """a_act [M,K] in the activation type, w_act [N,K] a weight operand (cast_weight / transpose_cast).
    act_f16 (split mode): the activation-type output is plain fp16 (attention operands)."""
    pr = _p(precision)
    M, K = a_act.shape
    N = w_act.shape[0]
    act_f16 = bool(act_f16) and pr == PREC_BF16X2
    of = torch.empty(M, N, dtype=torch.float32, device=a_act.device) if want_f32 else None
    oa = torch.empty(M, N, dtype=torch.float16 if act_f16 else act_dtype(pr), device=a_act.device) if want_act else None
    check(lib().ec_op_gemm_ex(pr, ptr(a_act), ptr(w_act), M, N, K, ptr(bias), alpha, act, ptr(residual), ptr(of), ptr(oa),
                              1 if act_f16 else 0, stream_ptr()))
    return of, oa


def gemm_train(a_act, w_act, bias, precision, drop=None, alpha=1.0, residual=None, want_f32=True, want_act=False, site=0, want_act2=False,
               site2=0, aux=None, site_aux=0):
    """Training-step GEMM with the surrounding element work in its epilogue (ec_op_gemm_train):
        z = a @ w^T + bias;  out_act = act(z);  out_act2 = act(dropout_site2(Swish(out_act)));
        aux given: z <- z * dropout_site_aux-mask * Swish'(aux)  (data gradient through dropout(Swish));
        out_f32 = alpha * dropout_site(z) + residual.
    `drop` is the DropoutState (None or p == 0: the sites are ignored).  Returns (out_f32, out_act, out_act2)."""
    pr = _p(precision)
    M, K = a_act.shape
    N = w_act.shape[0]
    live = drop is not None and drop.p > 0.0
    ctr = drop.counter if live else None
    p = drop.p if live else 0.0
    s0, s2, sa = (site, site2, site_aux) if live else (0, 0, 0)
    of = torch.empty(M, N, dtype=torch.float32, device=a_act.device) if want_f32 else None
    oa = torch.empty(M, N, dtype=act_dtype(pr), device=a_act.device) if want_act else None
    o2 = torch.empty(M, N, dtype=act_dtype(pr), device=a_act.device) if want_act2 else None
    check(lib().ec_op_gemm_train(pr, ptr(a_act), ptr(w_act), M, N, K, ptr(bias), alpha, ptr(residual), ptr(of), ptr(oa), ptr(ctr), p, s0,
                                 ptr(o2), s2, ptr(aux), sa, stream_ptr()))
    return of, oa, o2


def gemm_ln_train(a_act, w_act, bias, precision, ln_g, ln_b, drop=None, alpha=1.0, residual=None, site=0, eps=1e-6):
    """Training-step projection + dropout + residual with the NEXT module's LayerNorm fused (N <= 256, ec_op_gemm_ln_train):
    -> (out_f32 [M, N], ln_out [M, N] activation type)."""
    pr = _p(precision)
    M, K = a_act.shape
    N = w_act.shape[0]
    live = drop is not None and drop.p > 0.0
    of = torch.empty(M, N, dtype=torch.float32, device=a_act.device)
    ya = torch.empty(M, N, dtype=act_dtype(pr), device=a_act.device)
    check(lib().ec_op_gemm_ln_train(pr, ptr(a_act), ptr(w_act), M, N, K, ptr(bias), alpha, ptr(residual), ptr(of), ptr(ln_g.float().contiguous()),
                                    ptr(ln_b.float().contiguous()), eps, ptr(ya), ptr(drop.counter) if live else None, drop.p if live else 0.0,
                                    site if live else 0, stream_ptr()))
    return of, ya


def gemm_ln(a_act, w_act, bias, precision, g1, b1, g2=None, b2=None, mode=1, alpha=1.0, residual=None, eps=1e-6,
            copy_stride=0, frames_per_seq=0):
    """GEMM with the fused LayerNorm epilogue.  Returns (out_f32, ln_out_act, copy_out_act or None)."""
    pr = _p(precision)
    M, K = a_act.shape
    N = w_act.shape[0]
    of = torch.empty(M, N, dtype=torch.float32, device=a_act.device)
    ya = torch.empty(M, N, dtype=act_dtype(pr), device=a_act.device)
    cp, fops = None, 0
    if copy_stride:
        fops = (frames_per_seq - 1) // copy_stride + 1
        cp = torch.zeros((M // frames_per_seq) * fops, N, dtype=act_dtype(pr), device=a_act.device)
    check(lib().ec_op_gemm_ln(pr, ptr(a_act), ptr(w_act), M, N, K, ptr(bias), alpha, ptr(residual), ptr(of), mode, ptr(g1), ptr(b1),
                              ptr(g2), ptr(b2), eps, ptr(ya), ptr(cp), copy_stride, frames_per_seq, fops, stream_ptr()))
    return of, ya, cp


def ffn_fused(x_act, w1, b1, w2, b2, residual, g1, be1, g2=None, be2=None, mode=1, eps=1e-6, cluster=0, want_ln=True):
    """Fused feed-forward module (bf16 operands).  Returns (out_f32, ln_out bf16 or None)."""
    M, D = x_act.shape
    hidden = w1.shape[0]
    for t in (x_act, w1, w2):
        if t.dtype != torch.bfloat16:
            raise ValueError("ffn_fused takes bf16 operands")
    of = torch.empty(M, D, dtype=torch.float32, device=x_act.device)
    ya = torch.empty(M, D, dtype=torch.bfloat16, device=x_act.device) if want_ln else None
    check(lib().ec_op_ffn(ptr(x_act), ptr(w1), ptr(b1), ptr(w2), ptr(b2), M, D, hidden, ptr(residual), ptr(of), mode, ptr(g1), ptr(be1),
                          ptr(g2), ptr(be2), eps, ptr(ya), cluster, stream_ptr()))
    return of, ya


def pointwise_glu(a_act, w_raw, b_raw, precision):
    pr = _p(precision)
    M, K = a_act.shape
    Cc = w_raw.shape[0] // 2
    rows = lib().ec_op_glu_scratch_rows(Cc)
    ws = torch.empty(weight_planes(pr) * rows, K, dtype=act_dtype(pr), device=a_act.device)
    bs = torch.empty(rows, dtype=torch.float32, device=a_act.device)
    out = torch.empty(M, Cc, dtype=act_dtype(pr), device=a_act.device)
    w2 = w_raw.float().reshape(2 * Cc, K).contiguous()
    check(lib().ec_op_pointwise_glu(pr, ptr(a_act), ptr(w2), ptr(b_raw.float().contiguous()), M, Cc, K, ptr(ws), ptr(bs), ptr(out),
                                    stream_ptr()))
    return out


def fold_bn(w, b, g, beta, rm, rv, eps=1e-5):
    Cc = w.shape[0]
    taps = w.numel() // Cc
    wo = torch.empty(Cc, taps, dtype=torch.float32, device=w.device)
    bo = torch.empty(Cc, dtype=torch.float32, device=w.device)
    check(lib().ec_op_fold_bn(ptr(w.float().contiguous()), ptr(b.float().contiguous()), ptr(g.float().contiguous()),
                              ptr(beta.float().contiguous()), ptr(rm.float().contiguous()), ptr(rv.float().contiguous()), eps, Cc, taps,
                              ptr(wo), ptr(bo), stream_ptr()))
    return wo, bo


def relpos_attention(qkv, E, u, v, x_len, heads, group, precision):
    """qkv [B,T,3D], E [2Tp-G, D] (cast here to the activation type the kernel expects: TF32-rounded fp32 or bf16),
    x_len int32 [B] or None -> [B,T,D] activation type."""
    pr = _p(precision)
    B, T, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty(B, T, D, dtype=act_dtype(pr), device=qkv.device)
    xl = x_len.to(torch.int32).contiguous() if x_len is not None else None
    qkv, E = cast_attn_operand(qkv, pr, D, heads, group), cast_attn_operand(E, pr, D, heads, group)
    check(lib().ec_op_relpos_attention(pr, ptr(qkv), ptr(E), ptr(u.float().contiguous()),
                                       ptr(v.float().contiguous()), ptr(xl), B, T, D, heads, group, ptr(out), stream_ptr()))
    return out


def dwconv_bn_swish(x_act, w_folded, b_folded, stride, precision):
    pr = _p(precision)
    B, T, Cc = x_act.shape
    k = w_folded.shape[1]
    To = (T - 1) // stride + 1
    y = torch.empty(B, To, Cc, dtype=act_dtype(pr), device=x_act.device)
    check(lib().ec_op_dwconv_bn_swish(pr, ptr(x_act.contiguous()), ptr(w_folded), ptr(b_folded), B, T, Cc, k, stride, ptr(y), stream_ptr()))
    return y


def subsample_conv(mel, w_folded, b_folded, precision):
    pr = _p(precision)
    B, F, T = mel.shape
    Cc = w_folded.shape[0]
    To = (T - 1) // 2 + 1
    y = torch.empty(B, To, Cc * (F // 2), dtype=act_dtype(pr), device=mel.device)
    check(lib().ec_op_subsample_conv(pr, ptr(mel.float().contiguous()), ptr(w_folded), ptr(b_folded), B, F, T, Cc, ptr(y), stream_ptr()))
    return y


# ---- backward operators (training step) --------------------------------------------------------------------------------
def layernorm_bwd(x, dy, gamma, eps=1e-6, dx_accum=None, emit=None):
    """x, dy [rows, dim] fp32 -> (dx, dgamma, dbeta).  dx_accum: fp32 tensor the x-gradient is ADDED to (residual branch).
    emit = (precision, scale, drop, site): also returns, as a 4th value, the activation-type tensor
    act(scale * dropout_site-mask * dx_total) -- the operand dropout_cast_scaled(dx_total, ...) would produce in a separate pass
    (drop None or p == 0: a plain scaled cast)."""
    x, dy = x.float().contiguous(), dy.float().contiguous()
    rows, dim = x.numel() // x.shape[-1], x.shape[-1]
    dx = dx_accum if dx_accum is not None else torch.empty_like(x)
    dg = torch.empty(dim, dtype=torch.float32, device=x.device)
    db = torch.empty(dim, dtype=torch.float32, device=x.device)
    work = torch.empty(lib().ec_op_layernorm_bwd_work_bytes(dim), dtype=torch.uint8, device=x.device)
    if emit is None:
        check(lib().ec_op_layernorm_bwd(ptr(x), ptr(dy), rows, dim, ptr(gamma.float().contiguous()), eps, ptr(dx), 1 if dx_accum is not None else 0,
                                        ptr(dg), ptr(db), ptr(work), stream_ptr()))
        return dx, dg, db
    precision, scale, drop, site = emit
    pr = _p(precision)
    live = drop is not None and drop.p > 0.0
    out = torch.empty(x.shape, dtype=act_dtype(pr), device=x.device)
    check(lib().ec_op_layernorm_bwd_emit(ptr(x), ptr(dy), rows, dim, ptr(gamma.float().contiguous()), eps, ptr(dx), 1 if dx_accum is not None else 0,
                                         ptr(dg), ptr(db), ptr(work), pr, ptr(out), float(scale), ptr(drop.counter) if live else None,
                                         drop.p if live else 0.0, site if live else 0, stream_ptr()))
    return dx, dg, db, out


def colsum(m, precision, is_f32=None):
    """Column sums (bias gradient) of a 2-D fp32 or activation-type matrix.  fp32-typed tensors are plain fp32 values unless the
    precision is the split mode, whose packed (hi, lo) words also travel as fp32-typed tensors: pass is_f32=True for true fp32 there."""
    pr = _p(precision)
    m = m.contiguous()
    rows, cols = m.shape
    if is_f32 is None:
        is_f32 = m.dtype == torch.float32 and pr != PREC_BF16X2
    out = torch.empty(cols, dtype=torch.float32, device=m.device)
    work = torch.empty(lib().ec_op_colsum_work_bytes(cols), dtype=torch.uint8, device=m.device)
    check(lib().ec_op_colsum(pr, ptr(m), 1 if is_f32 else 0, rows, cols, ptr(out), ptr(work), stream_ptr()))
    return out


def transpose_cast(w, precision):
    """fp32 [rows, cols] -> weight operand [cols, rows] (plane-0 view of [planes, cols, rows], see cast_weight)."""
    pr = _p(precision)
    w = w.float().contiguous()
    rows, cols = w.shape
    out = torch.empty(weight_planes(pr), cols, rows, dtype=act_dtype(pr), device=w.device)
    check(lib().ec_op_transpose_cast(pr, ptr(w), rows, cols, ptr(out), stream_ptr()))
    return out[0]


def linear_dgrad(dy_act, w_fp32, precision, residual=None):
    """dX = dY . W  (+ residual): the forward tcgen05 GEMM on the transposed weight copy.  dy_act [M, N] activation type,
    w_fp32 [N, K] fp32 (the nn.Linear weight) -> dX [M, K] fp32."""
    wt = transpose_cast(w_fp32, precision)                      # [K, N] = the "weight" of a GEMM that reduces over N
    return gemm(dy_act, wt, None, precision, residual=residual)[0]


def swish_bwd(z_act, dy, precision):
    pr = _p(precision)
    z_act, dy = z_act.contiguous(), dy.float().contiguous()
    out = torch.empty_like(z_act)
    check(lib().ec_op_swish_bwd(pr, ptr(z_act), ptr(dy), z_act.numel(), ptr(out), stream_ptr()))
    return out


def glu_bwd(zg_act, dy, precision):
    """zg_act [rows, 2C] = [a | g] (pre-GLU pointwise output), dy [rows, C] fp32 -> [da | dg] [rows, 2C] activation type."""
    pr = _p(precision)
    zg_act, dy = zg_act.contiguous(), dy.float().contiguous()
    rows, C2 = zg_act.shape
    out = torch.empty_like(zg_act)
    check(lib().ec_op_glu_bwd(pr, ptr(zg_act), ptr(dy), rows, C2 // 2, ptr(out), stream_ptr()))
    return out


def linear_wgrad(dy_act, x_act, precision, dw_accum=None):
    """dW [N, K] fp32 (+)= dY[M, N]^T . X[M, K] (both activation type): tcgen05 with MN-major operands."""
    pr = _p(precision)
    dy_act, x_act = dy_act.contiguous(), x_act.contiguous()
    M, N = dy_act.shape
    K = x_act.shape[1]
    dw = dw_accum if dw_accum is not None else torch.empty(N, K, dtype=torch.float32, device=dy_act.device)
    work = torch.empty(lib().ec_op_wgrad_work_bytes(pr, M, N, K), dtype=torch.uint8, device=dy_act.device)
    check(lib().ec_op_wgrad(pr, ptr(dy_act), ptr(x_act), M, N, K, ptr(dw), 1 if dw_accum is not None else 0, ptr(work), stream_ptr()))
    return dw


def linear_wgrad_bias(dy_act, x_act, precision):
    """(dW [N, K], db [N]) fp32 of a Linear from ONE kernel: dW = dY^T . X on tcgen05, db = column sums of dY through a ones operand."""
    pr = _p(precision)
    dy_act, x_act = dy_act.contiguous(), x_act.contiguous()
    M, N = dy_act.shape
    K = x_act.shape[1]
    dw = torch.empty(N, K, dtype=torch.float32, device=dy_act.device)
    db = torch.empty(N, dtype=torch.float32, device=dy_act.device)
    work = torch.empty(lib().ec_op_wgrad_work_bytes(pr, M, N, K), dtype=torch.uint8, device=dy_act.device)
    check(lib().ec_op_wgrad_bias(pr, ptr(dy_act), ptr(x_act), M, N, K, ptr(dw), 0, ptr(db), ptr(work), stream_ptr()))
    return dw, db


class DwConvTrain:
    """Training-mode depthwise conv -> BatchNorm1d (batch statistics, running-stat update) -> Swish and its backward
    (reference models/modules.py:515-517 under .train()).  `reduce_stats` (SyncBatchNorm, efficientconformer_b200/distributed.py) merges the
    [2, C] (mean, M2) forward statistics across ranks between the stages -- `forward_stats(stats, count)` returns the global frame
    count -- and all-reduces the backward sums -- `backward_sums(sums)`."""

    @staticmethod
    def forward(x_act, w, b, gamma, beta, running_mean, running_var, stride, precision, eps=1e-5, momentum=0.1, reduce_stats=None):
        pr = _p(precision)
        L = lib()
        B, T, Cc = x_act.shape
        K = w.shape[-1]
        To = (T - 1) // stride + 1
        dev = x_act.device
        w2, b2 = w.reshape(Cc, K).float().contiguous(), b.float().contiguous()
        y = torch.empty(B, To, Cc, dtype=torch.float32, device=dev)
        sums = torch.empty(2, Cc, dtype=torch.float32, device=dev)
        work = torch.empty(L.ec_op_conv_train_work_bytes(Cc, K), dtype=torch.uint8, device=dev)
        check(L.ec_op_dwconv_raw(pr, ptr(x_act.contiguous()), ptr(w2), ptr(b2), B, T, Cc, K, stride, ptr(y), ptr(sums), ptr(work), stream_ptr()))
        count = float(B * To)
        if reduce_stats is not None:
            count = reduce_stats.forward_stats(sums, count)
        mean = torch.empty(Cc, dtype=torch.float32, device=dev)
        rstd = torch.empty(Cc, dtype=torch.float32, device=dev)
        check(L.ec_op_bn_finalize(ptr(sums), Cc, count, eps, momentum, ptr(mean), ptr(rstd), ptr(running_mean), ptr(running_var), stream_ptr()))
        h = torch.empty(B, To, Cc, dtype=act_dtype(pr), device=dev)
        g, be = gamma.float().contiguous(), beta.float().contiguous()
        check(L.ec_op_bn_swish_fwd(pr, ptr(y), B * To, Cc, ptr(mean), ptr(rstd), ptr(g), ptr(be), ptr(h), stream_ptr()))
        return h, (x_act, w2, y, mean, rstd, g, be, count, stride, pr)

    @staticmethod
    def backward(dh, saved, reduce_stats=None):
        """dh [B, T_out, C] fp32 -> (dx fp32 [B,T,C], dw [C,K], db [C], dgamma [C], dbeta [C])."""
        x_act, w2, y, mean, rstd, g, be, count, stride, pr = saved
        L = lib()
        B, T, Cc = x_act.shape
        K = w2.shape[1]
        To = y.shape[1]
        dev = y.device
        dh = dh.float().contiguous()
        work = torch.empty(L.ec_op_conv_train_work_bytes(Cc, K), dtype=torch.uint8, device=dev)
        sums = torch.empty(2, Cc, dtype=torch.float32, device=dev)
        check(L.ec_op_bn_swish_bwd_stats(ptr(y), ptr(dh), B * To, Cc, ptr(mean), ptr(rstd), ptr(g), ptr(be), ptr(sums), ptr(work), stream_ptr()))
        dbeta, dgamma = sums[0].clone(), sums[1].clone()          # local sums are this rank's parameter gradients
        if reduce_stats is not None:
            reduce_stats.backward_sums(sums)
        dy = torch.empty_like(y)
        check(L.ec_op_bn_swish_bwd_apply(ptr(y), ptr(dh), B * To, Cc, ptr(mean), ptr(rstd), ptr(g), ptr(be), ptr(sums), count, ptr(dy), stream_ptr()))
        dx = torch.empty(B, T, Cc, dtype=torch.float32, device=dev)
        dw = torch.empty(Cc, K, dtype=torch.float32, device=dev)
        db = torch.empty(Cc, dtype=torch.float32, device=dev)
        check(L.ec_op_dwconv_bwd(pr, ptr(dy), ptr(x_act), ptr(w2), B, T, Cc, K, stride, ptr(dx), ptr(dw), ptr(db), ptr(work), stream_ptr()))
        return dx, dw, db, dgamma, dbeta


def relpos_attention_bwd(qkv_act, E_act, u, v, x_len, heads, group, d_out, precision):
    """Gradients of relpos_attention: qkv_act [B,T,3D], E_act [2Tp-G, D] (activation type), d_out [B,T,D] fp32
    -> (dqkv [B,T,3D], dE [2Tp-G, D], du [D], dv [D]) fp32."""
    pr = _p(precision)
    B, T, D3 = qkv_act.shape
    D = D3 // 3
    dev = qkv_act.device
    xl = x_len.to(torch.int32).contiguous() if x_len is not None else None
    dqkv = torch.empty(B, T, D3, dtype=torch.float32, device=dev)
    dE = torch.empty(E_act.shape, dtype=torch.float32, device=dev)
    du = torch.empty(D, dtype=torch.float32, device=dev)
    dv = torch.empty(D, dtype=torch.float32, device=dev)
    work = torch.empty(lib().ec_op_relpos_attention_bwd_work_bytes(B, T, D, heads, group), dtype=torch.uint8, device=dev)
    check(lib().ec_op_relpos_attention_bwd(pr, ptr(qkv_act.contiguous()), ptr(E_act.contiguous()), ptr(u.float().contiguous()),
                                           ptr(v.float().contiguous()), ptr(xl), B, T, D, heads, group, ptr(d_out.float().contiguous()),
                                           ptr(dqkv), ptr(dE), ptr(du), ptr(dv), ptr(work), stream_ptr()))
    return dqkv, dE, du, dv


def relpos_attention_bwd_act(qkv_act, E_act, u, v, x_len, heads, group, d_out, precision):
    """relpos_attention_bwd with dq | dk | dv delivered in the activation type (what the QKV weight / data gradient GEMMs read):
    -> (dqkv_act [B,T,3D], dE fp32, du, dv)."""
    pr = _p(precision)
    B, T, D3 = qkv_act.shape
    D = D3 // 3
    dev = qkv_act.device
    xl = x_len.to(torch.int32).contiguous() if x_len is not None else None
    tc = pr != PRECISIONS["tf32"]
    dqkv32 = None if tc else torch.empty(B, T, D3, dtype=torch.float32, device=dev)     # scratch of the TF32 CUDA-core path
    dqkv = torch.empty(B, T, D3, dtype=act_dtype(pr), device=dev)
    dE = torch.empty(E_act.shape, dtype=torch.float32, device=dev)
    du = torch.empty(D, dtype=torch.float32, device=dev)
    dv = torch.empty(D, dtype=torch.float32, device=dev)
    work = torch.empty(lib().ec_op_relpos_attention_bwd_work_bytes(B, T, D, heads, group), dtype=torch.uint8, device=dev)
    if tc and (D % 2 or ((group * D) // heads) % 2 or (T + group - 1) // group > 1024):
        dqkv32 = torch.empty(B, T, D3, dtype=torch.float32, device=dev)                  # odd head layouts / very long utterances: fp32 first
    check(lib().ec_op_relpos_attention_bwd_act(pr, ptr(qkv_act.contiguous()), ptr(E_act.contiguous()), ptr(u.float().contiguous()),
                                               ptr(v.float().contiguous()), ptr(xl), B, T, D, heads, group, ptr(d_out.float().contiguous()),
                                               ptr(dqkv32), ptr(dqkv), ptr(dE), ptr(du), ptr(dv), ptr(work), stream_ptr()))
    return dqkv, dE, du, dv


def cast_scaled(x, precision, scale):
    pr = _p(precision)
    x = x.float().contiguous()
    out = torch.empty(x.shape, dtype=act_dtype(pr), device=x.device)
    check(lib().ec_op_cast_scaled(pr, ptr(x), float(scale), x.numel(), ptr(out), stream_ptr()))
    return out


def swish_fwd(z_act, precision):
    out = torch.empty_like(z_act)
    check(lib().ec_op_swish_fwd(_p(precision), ptr(z_act.contiguous()), z_act.numel(), ptr(out), stream_ptr()))
    return out


def glu_fwd(zg_act, precision):
    rows, C2 = zg_act.shape
    out = torch.empty(rows, C2 // 2, dtype=zg_act.dtype, device=zg_act.device)
    check(lib().ec_op_glu_fwd(_p(precision), ptr(zg_act.contiguous()), rows, C2 // 2, ptr(out), stream_ptr()))
    return out


def strided_rows(x, stride, precision):
    """x [B, T, D] fp32 -> activation-type copy of frames 0, s, 2s, ... : [B, (T-1)//s+1, D]."""
    pr = _p(precision)
    B, T, D = x.shape
    out = torch.empty(B, (T - 1) // stride + 1, D, dtype=act_dtype(pr), device=x.device)
    check(lib().ec_op_strided_rows(pr, ptr(x.float().contiguous()), B, T, D, stride, ptr(out), stream_ptr()))
    return out


def strided_rows_bwd(d, dx, stride):
    """dx [B, T, D] fp32 += scatter of d [B, T_out, D] to frames 0, s, 2s, ..."""
    B, T, D = dx.shape
    check(lib().ec_op_strided_rows_bwd(ptr(d.float().contiguous()), B, T, D, stride, ptr(dx), stream_ptr()))
    return dx


class SubsampleTrain:
    """Training-mode Conv2d(1 -> C, 3x3, s2) -> BatchNorm2d (batch statistics over (b, f, t), running-stat update) -> Swish
    (reference models/modules.py:226-249 under .train()), output in the layout of the following Linear ([B*T/2, C*F/2]),
    and its backward (weight / bias / BatchNorm gradients; the mel input needs no gradient)."""

    @staticmethod
    def forward(mel, w, b, gamma, beta, running_mean, running_var, precision, eps=1e-5, momentum=0.1, reduce_stats=None):
        pr = _p(precision)
        L = lib()
        B, F, T = mel.shape
        Cc = w.shape[0]
        F2, To = F // 2, (T - 1) // 2 + 1
        cols, rows = Cc * F2, B * To
        dev = mel.device
        mel = mel.float().contiguous()
        y = torch.empty(rows, cols, dtype=torch.float32, device=dev)
        check(L.ec_op_subsample_conv_raw(ptr(mel), ptr(w.reshape(Cc, 9).float().contiguous()), ptr(b.float().contiguous()), B, F, T, Cc,
                                         ptr(y), stream_ptr()))
        col_stats = torch.empty(2, cols, dtype=torch.float32, device=dev)
        work = torch.empty(L.ec_op_col_stats_work_bytes(cols), dtype=torch.uint8, device=dev)
        check(L.ec_op_col_stats(ptr(y), rows, cols, ptr(col_stats), ptr(work), stream_ptr()))
        stats = torch.empty(2, Cc, dtype=torch.float32, device=dev)
        check(L.ec_op_group_stats_merge(ptr(col_stats), Cc, F2, rows, ptr(stats), stream_ptr()))
        count = float(rows * F2)
        if reduce_stats is not None:
            count = reduce_stats.forward_stats(stats, count)
        ch = torch.empty(4, Cc, dtype=torch.float32, device=dev)           # mean, rstd, gamma, beta per channel
        ch[2].copy_(gamma); ch[3].copy_(beta)
        check(L.ec_op_bn_finalize(ptr(stats), Cc, count, eps, momentum, ptr(ch[0]), ptr(ch[1]), ptr(running_mean), ptr(running_var), stream_ptr()))
        colv = torch.empty(4, cols, dtype=torch.float32, device=dev)        # the same, expanded to the c*F2 + f columns
        check(L.ec_op_group_expand(ptr(ch), 4, Cc, F2, ptr(colv), stream_ptr()))
        a = torch.empty(rows, cols, dtype=act_dtype(pr), device=dev)
        check(L.ec_op_bn_swish_fwd(pr, ptr(y), rows, cols, ptr(colv[0]), ptr(colv[1]), ptr(colv[2]), ptr(colv[3]), ptr(a), stream_ptr()))
        return a, (mel, y, colv, count, Cc, F2, (B, F, T))

    @staticmethod
    def backward(da, saved, reduce_stats=None):
        """da [B*T/2, C*F/2] fp32 -> (dw [C,1,3,3], db [C], dgamma [C], dbeta [C])."""
        mel, y, colv, count, Cc, F2, (B, F, T) = saved
        L = lib()
        rows, cols = y.shape
        dev = y.device
        da = da.float().contiguous()
        work = torch.empty(max(L.ec_op_conv_train_work_bytes(cols, 1), L.ec_op_subsample_wgrad_work_bytes(Cc, F, B, T)), dtype=torch.uint8, device=dev)
        sums_col = torch.empty(2, cols, dtype=torch.float32, device=dev)
        check(L.ec_op_bn_swish_bwd_stats(ptr(y), ptr(da), rows, cols, ptr(colv[0]), ptr(colv[1]), ptr(colv[2]), ptr(colv[3]), ptr(sums_col),
                                         ptr(work), stream_ptr()))
        sums = torch.empty(2, Cc, dtype=torch.float32, device=dev)
        check(L.ec_op_group_sum(ptr(sums_col), 2, Cc, F2, ptr(sums), stream_ptr()))
        dbeta, dgamma = sums[0].clone(), sums[1].clone()
        if reduce_stats is not None:
            reduce_stats.backward_sums(sums)
        check(L.ec_op_group_expand(ptr(sums), 2, Cc, F2, ptr(sums_col), stream_ptr()))
        dy = torch.empty_like(y)
        check(L.ec_op_bn_swish_bwd_apply(ptr(y), ptr(da), rows, cols, ptr(colv[0]), ptr(colv[1]), ptr(colv[2]), ptr(colv[3]), ptr(sums_col),
                                         count, ptr(dy), stream_ptr()))
        dw = torch.empty(Cc, 1, 3, 3, dtype=torch.float32, device=dev)
        db = torch.empty(Cc, dtype=torch.float32, device=dev)
        check(L.ec_op_subsample_wgrad(ptr(dy), ptr(mel), B, F, T, Cc, ptr(dw), ptr(db), ptr(work), stream_ptr()))
        return dw, db, dgamma, dbeta


# ---- pieces the assembled training step needs (efficientconformer_b200/training.py) -----------------------------------------
def relpos_attention_act(qkv_act, E_act, u, v, x_len, heads, group, precision):
    """relpos_attention on operands that are already in the activation type (what the QKV / pos GEMM epilogues emit)."""
    pr = _p(precision)
    B, T, D3 = qkv_act.shape
    D = D3 // 3
    out = torch.empty(B, T, D, dtype=act_dtype(pr), device=qkv_act.device)
    xl = x_len.to(torch.int32).contiguous() if x_len is not None else None
    check(lib().ec_op_relpos_attention(pr, ptr(qkv_act.contiguous()), ptr(E_act.contiguous()), ptr(u.float().contiguous()),
                                       ptr(v.float().contiguous()), ptr(xl), B, T, D, heads, group, ptr(out), stream_ptr()))
    return out


def concat_qkv(mhsa):
    """fp32 staging copy [3D, D] of Wq | Wk | Wv and [3D] of the biases (device-to-device memcpys, no kernel): one GEMM computes q|k|v
    and one transposed copy serves the data gradient."""
    D = mhsa.query_layer.weight.shape[0]
    dev = mhsa.query_layer.weight.device
    w = torch.empty(3 * D, D, dtype=torch.float32, device=dev)
    b = torch.empty(3 * D, dtype=torch.float32, device=dev)
    for j, layer in enumerate((mhsa.query_layer, mhsa.key_layer, mhsa.value_layer)):
        w[j * D:(j + 1) * D].copy_(layer.weight.detach())
        b[j * D:(j + 1) * D].copy_(layer.bias.detach())
    return w, b


def own_f32(x):
    """A private contiguous fp32 copy the backward may accumulate into."""
    return x.detach().float().clone(memory_format=torch.contiguous_format)


def zeros_f32(rows, cols, device):
    return torch.zeros(rows, cols, dtype=torch.float32, device=device)      # a memset node


def dropout_counter(device, seed):
    """Device {seed, step} pair of the counter-based dropout (ec_op_dropout*)."""
    t = torch.zeros(2, dtype=torch.int64, device=device)
    t[0] = int(seed) & 0x7FFFFFFFFFFFFFFF
    return t


def dropout_advance(counter):
    check(lib().ec_op_dropout_advance(ptr(counter), stream_ptr()))


def _dropout(x, drop, site, precision, src_f32, dst_f32, scale=1.0):
    pr = _p(precision)
    x = x.contiguous()
    dt = torch.float32 if dst_f32 else act_dtype(pr)
    out = torch.empty(x.shape, dtype=dt, device=x.device)
    check(lib().ec_op_dropout(pr, ptr(x), 1 if src_f32 else 0, float(scale), x.numel(), ptr(out), 1 if dst_f32 else 0, drop.p,
                              ptr(drop.counter), site, stream_ptr()))
    return out


def dropout_f32(x, drop, site):
    """fp32 -> fp32 (residual stream / gradients); identity when p == 0."""
    if drop.p == 0.0:
        return x
    return _dropout(x.float(), drop, site, PRECISIONS["tf32"], True, True)


def dropout_act(x_act, drop, site, precision):
    if drop.p == 0.0:
        return x_act
    return _dropout(x_act, drop, site, precision, False, False)


def dropout_cast_scaled(x, precision, scale, drop, site):
    """fp32 gradient -> activation type, scaled, with the forward mask of `site` re-applied."""
    if drop.p == 0.0:
        return cast_scaled(x, precision, scale) if scale != 1.0 else cast(x, precision)
    return _dropout(x.float(), drop, site, precision, True, False, scale)


def dropout_residual(y, drop, site, alpha, residual):
    y, residual = y.float().contiguous(), residual.float().contiguous()
    out = torch.empty_like(y)
    check(lib().ec_op_dropout_residual(ptr(y), ptr(residual), float(alpha), y.numel(), ptr(out), drop.p, ptr(drop.counter), site, stream_ptr()))
    return out


def stats_merge_ranks(gathered, counts, out):
    """gathered [W, 2, C] fp32, counts [W] fp32 (device) -> out [2, C] (in place)."""
    W, _, Cc = gathered.shape
    check(lib().ec_op_stats_merge_ranks(ptr(gathered.contiguous()), ptr(counts), W, Cc, ptr(out), stream_ptr()))
    return out


def adam_step(params, grads, exp_avg, exp_avg_sq, state, beta1, beta2, eps, weight_decay, grad_scale=1.0, schedule=0, K=0.0, dim=1.0,
              warmup=1.0):
    """One torch.optim.Adam step over flat fp32 arenas (ec_adam_step); `state` int32[4] device = {lr bits, t, s, 0}."""
    check(lib().ec_adam_step(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), params.numel(), ptr(state), beta1, beta2, eps,
                             weight_decay, grad_scale, schedule, K, dim, warmup, stream_ptr()))


def swish_dropout_fwd(z_act, drop, site, precision):
    """Swish followed by the feed-forward module's first dropout, one kernel (plain Swish when p == 0)."""
    if drop.p == 0.0:
        return swish_fwd(z_act, precision)
    z_act = z_act.contiguous()
    out = torch.empty_like(z_act)
    check(lib().ec_op_swish_dropout(_p(precision), ptr(z_act), None, z_act.numel(), ptr(out), drop.p, ptr(drop.counter), site, stream_ptr()))
    return out


def swish_dropout_bwd(z_act, dy, drop, site, precision):
    """Gradient through dropout(Swish(z)) with the forward mask of `site` re-applied, one kernel."""
    if drop.p == 0.0:
        return swish_bwd(z_act, dy, precision)
    z_act, dy = z_act.contiguous(), dy.float().contiguous()
    out = torch.empty_like(z_act)
    check(lib().ec_op_swish_dropout(_p(precision), ptr(z_act), ptr(dy), z_act.numel(), ptr(out), drop.p, ptr(drop.counter), site, stream_ptr()))
    return out


def cast_into(x_f32, out_act, precision):
    """fp32 -> activation type into a preallocated tensor (the whole flat parameter arena in one launch)."""
    check(lib().ec_op_cast(_p(precision), ptr(x_f32), ptr(out_act), x_f32.numel(), stream_ptr()))
    return out_act


def cast_multi(src_arena, desc_dev, n, dst_arena, precision, ctas_per_tensor=16):
    """desc_dev int64 [n, 4] = (src offset, rows, cols, dst offset): every forward weight operand in one launch."""
    check(lib().ec_op_cast_multi(_p(precision), ptr(src_arena), ptr(desc_dev), n, ctas_per_tensor, ptr(dst_arena), stream_ptr()))
    return dst_arena


def transpose_cast_multi(src_arena, desc_dev, n, dst_arena, precision, ctas_per_tensor=32):
    """desc_dev int64 [n, 4] = (src offset, rows, cols, dst offset): every W^T operand of the data-gradient GEMMs in one launch."""
    check(lib().ec_op_transpose_cast_multi(_p(precision), ptr(src_arena), ptr(desc_dev), n, ctas_per_tensor, ptr(dst_arena), stream_ptr()))
    return dst_arena


# ---- front end on the device (csrc/frontend.cu) -----------------------------------------------------------
def logmel(audio, window, fb, krange, hop, normalize=False, mean=0.0, std=1.0):
    """audio (B, L) fp32 -> (B, n_mels, L // hop + 1) fp32: reference models/modules.py:87-106 (STFT power -> mel -> log [-> normalise]).
    window (n_fft,) = analysis window centred / zero padded to n_fft; fb (n_fft // 2 + 1, n_mels); krange (n_mels, 2) int32 or None."""
    audio = audio.contiguous()
    B, L = audio.shape
    n_fft, n_mels = window.numel(), fb.shape[1]
    out = torch.empty(B, n_mels, L // hop + 1, dtype=torch.float32, device=audio.device)
    check(lib().ec_op_logmel(ptr(audio), B, L, n_fft, int(hop), ptr(window), ptr(fb), ptr(krange) if krange is not None else None, n_mels,
                             1 if normalize else 0, float(mean), float(std), ptr(out), stream_ptr()))
    return out


def specaugment_(mel, x_len, mF, F, mT, pS, counter):
    """In place on mel (B, n_mels, T) fp32: reference models/modules.py:136-151 with counter-based draws ({seed, step} device pair)."""
    B, n_mels, T = mel.shape
    check(lib().ec_op_specaugment(ptr(mel), ptr(x_len) if x_len is not None else None, B, n_mels, T, int(mF), int(F), int(mT), float(pS),
                                  ptr(counter), stream_ptr()))
    return mel
