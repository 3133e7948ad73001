"""TEST INFRASTRUCTURE ONLY -- a torch-CPU operator table with the signatures of efficientconformer_b200.ops, used by
tests/test_train_glue_cpu.py to check the TAPE LOGIC of efficientconformer_b200/training.py (which tensors are saved, which
gradient goes where, residual accumulation order) on a machine without a GPU.  It is never imported by the product package;
the product operators (efficientconformer_b200/ops.py) have no CPU implementation and raise without the CUDA library.
Every function is exact fp32/fp64 torch math: "activation type" is plain fp32 here (no TF32 / bf16 rounding)."""
import torch
import torch.nn.functional as F

DT = torch.float64        # the glue is checked in double so that only ordering mistakes, not rounding, can show up


def _d(x):
    return x.detach().to(DT)


def cast(x, precision):
    return _d(x).clone()


def cast_weight(x, precision):
    return _d(x).clone()


def cast_scaled(x, precision, scale):
    return _d(x) * scale


def layernorm(x, gamma, beta, precision, eps=1e-6, want_f32=True, want_act=True):
    y = F.layer_norm(_d(x), (x.shape[-1],), _d(gamma), _d(beta), eps)
    return (y.clone() if want_act else None), (y if want_f32 else None)


def attn_operands_f16(precision, dim, heads, group):
    return False


def gemm(a_act, w_act, bias, precision, alpha=1.0, act=0, residual=None, want_f32=True, want_act=False, act_f16=False):
    y = _d(a_act) @ _d(w_act).t()
    if bias is not None:
        y = y + _d(bias)
    if act == 1:
        y = y * torch.sigmoid(y)
    y = alpha * y
    if residual is not None:
        y = y + _d(residual)
    return (y if want_f32 else None), (y.clone() if want_act else None)


def gemm_train(a_act, w_act, bias, precision, drop=None, alpha=1.0, residual=None, want_f32=True, want_act=False, site=0, want_act2=False,
               site2=0, aux=None, site_aux=0):
    assert drop is None or drop.p == 0.0
    z = _d(a_act) @ _d(w_act).t()
    if bias is not None:
        z = z + _d(bias)
    if aux is not None:
        s = torch.sigmoid(_d(aux))
        z = z * (s + _d(aux) * s * (1 - s))
    y = alpha * z
    if residual is not None:
        y = y + _d(residual)
    return (y if want_f32 else None), (z.clone() if want_act else None), (z * torch.sigmoid(z) if want_act2 else None)


def gemm_ln_train(a_act, w_act, bias, precision, ln_g, ln_b, drop=None, alpha=1.0, residual=None, site=0, eps=1e-6):
    out = gemm_train(a_act, w_act, bias, precision, drop, alpha=alpha, residual=residual, site=site)[0]
    return out, F.layer_norm(out, (out.shape[-1],), _d(ln_g), _d(ln_b), eps)


def swish_fwd(z, precision):
    return z * torch.sigmoid(z)


def glu_fwd(zg, precision):
    C = zg.shape[1] // 2
    return zg[:, :C] * torch.sigmoid(zg[:, C:])


def swish_bwd(z, dy, precision):
    s = torch.sigmoid(z)
    return _d(dy) * (s + z * s * (1 - s))


def glu_bwd(zg, dy, precision):
    C = zg.shape[1] // 2
    a, g = zg[:, :C], zg[:, C:]
    s = torch.sigmoid(g)
    return torch.cat([_d(dy) * s, _d(dy) * a * s * (1 - s)], dim=1)


def layernorm_bwd(x, dy, gamma, eps=1e-6, dx_accum=None, emit=None):
    xr = _d(x).clone().requires_grad_(True)
    g = _d(gamma).clone().requires_grad_(True)
    b = torch.zeros_like(g).requires_grad_(True)
    with torch.enable_grad():
        F.layer_norm(xr, (x.shape[-1],), g, b, eps).backward(_d(dy))
    dx = xr.grad
    if dx_accum is not None:
        dx_accum += xr.grad                      # in place, like the kernel's accumulate flag
        dx = dx_accum
    if emit is not None:
        precision, scale, drop, site = emit
        assert drop is None or drop.p == 0.0
        return dx, g.grad, b.grad, dx.clone() * scale
    return dx, g.grad, b.grad


def colsum(m, precision):
    return _d(m).sum(0)


def linear_dgrad(dy_act, w_fp32, precision, residual=None):
    y = _d(dy_act) @ _d(w_fp32)
    return y + _d(residual) if residual is not None else y


def linear_wgrad(dy_act, x_act, precision, dw_accum=None):
    dw = _d(dy_act).t() @ _d(x_act)
    if dw_accum is not None:
        dw_accum += dw
        return dw_accum
    return dw


def linear_wgrad_bias(dy_act, x_act, precision):
    return _d(dy_act).t() @ _d(x_act), _d(dy_act).sum(0)


def _attention(qkv, E, u, v, x_len, H, G):
    """Closed form of SURVEY.md section 8 row a9 (same as tests/test_gpu_ops._attention_reference)."""
    B, T, D3 = qkv.shape
    D = D3 // 3
    d = G * D // H
    q, k, vv = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    P = (-T) % G
    if P:
        q, k, vv = (F.pad(t, (0, 0, 0, P)) for t in (q, k, vv))
    Tp = T + P
    Tg = Tp // G
    qu = (q + u).reshape(B, Tg, H, d).transpose(1, 2)
    qv = (q + v).reshape(B, Tg, H, d).transpose(1, 2)
    kk = k.reshape(B, Tg, H, d).transpose(1, 2)
    vh = vv.reshape(B, Tg, H, d).transpose(1, 2)
    Eh = E.reshape(2 * Tg - 1, H, d).transpose(0, 1)
    s_k = qu @ kk.transpose(2, 3)
    rel = qv @ Eh.transpose(1, 2).unsqueeze(0)
    idx = (Tg - 1) + torch.arange(Tg)[None, :] - torch.arange(Tg)[:, None]
    s_e = torch.gather(rel, 3, idx[None, None].expand(B, H, Tg, Tg))
    s = (s_k + s_e) / d ** 0.5
    if x_len is not None:
        masked = (torch.arange(Tg) * G)[None, :] >= x_len[:, None]
        s = s + masked[:, None, None, :].to(s.dtype) * -1e9
    w = s.softmax(-1)
    return (w @ vh).transpose(1, 2).reshape(B, Tp, D)[:, :T]


def relpos_attention_act(qkv_act, E_act, u, v, x_len, heads, group, precision):
    return _attention(_d(qkv_act), _d(E_act), _d(u), _d(v), x_len, heads, group).contiguous()


def relpos_attention_bwd_act(qkv_act, E_act, u, v, x_len, heads, group, d_out, precision):
    return relpos_attention_bwd(qkv_act, E_act, u, v, x_len, heads, group, d_out, precision)


def relpos_attention_bwd(qkv_act, E_act, u, v, x_len, heads, group, d_out, precision):
    leaves = [_d(t).clone().requires_grad_(True) for t in (qkv_act, E_act, u, v)]
    with torch.enable_grad():
        _attention(*leaves, x_len, heads, group).backward(_d(d_out))
    return tuple(t.grad for t in leaves)


class _StagedBatchNorm:
    """Train-mode BatchNorm + Swish in the same STAGES as the CUDA operators (raw statistics -> optional cross-rank exchange ->
    normalise; backward sums -> optional exchange -> apply), so that the SyncBatchNorm protocol of efficientconformer_b200/distributed.py
    can be exercised end to end over gloo.  `y` is [rows, C] (channel last); without a reducer this equals F.batch_norm(training=True)."""

    @staticmethod
    def forward(y, gamma, beta, running_mean, running_var, eps, momentum, reduce_stats):
        rows = y.shape[0]
        mean = y.mean(0)
        stats = torch.stack([mean, ((y - mean) ** 2).sum(0)]).to(torch.float32)       # the exchange runs on fp32 (mean, M2) pairs
        count = float(rows)
        if reduce_stats is not None:
            count = reduce_stats.forward_stats(stats, count)
            mean, m2 = stats[0].to(DT), stats[1].to(DT)
        else:
            mean, m2 = mean, ((y - mean) ** 2).sum(0)
        var = m2 / count
        rstd = 1.0 / torch.sqrt(var + eps)
        running_mean.copy_(((1 - momentum) * _d(running_mean) + momentum * mean).to(running_mean.dtype))
        running_var.copy_(((1 - momentum) * _d(running_var) + momentum * var * (count / max(count - 1.0, 1.0))).to(running_var.dtype))
        xhat = (y - mean) * rstd
        z = xhat * gamma + beta
        return z * torch.sigmoid(z), (xhat, z, rstd, gamma, count)

    @staticmethod
    def backward(dh, saved, reduce_stats):
        xhat, z, rstd, gamma, count = saved
        s = torch.sigmoid(z)
        dz = dh * (s + z * s * (1 - s))
        sums = torch.stack([dz.sum(0), (dz * xhat).sum(0)]).to(torch.float32)
        dbeta, dgamma = sums[0].to(DT).clone(), sums[1].to(DT).clone()                 # this rank's parameter gradients (local sums)
        if reduce_stats is not None:
            reduce_stats.backward_sums(sums)
        g = sums.to(DT)
        dy = gamma * rstd * (dz - g[0] / count - xhat * g[1] / count)
        return dy, dgamma, dbeta


class DwConvTrain:
    @staticmethod
    def forward(x_act, w, b, gamma, beta, running_mean, running_var, stride, precision, eps=1e-5, momentum=0.1, reduce_stats=None):
        xr, wr, br = [_d(t).clone().requires_grad_(True) for t in (x_act, w, b)]
        k, C = w.shape[-1], w.shape[0]
        with torch.enable_grad():
            conv = F.conv1d(F.pad(xr.transpose(1, 2), ((k - 1) // 2, (k - 1) // 2)), wr.reshape(C, 1, k), br, stride=stride, groups=C)
            y = conv.transpose(1, 2)                                                  # [B, To, C]
        B, To, _ = y.shape
        h, bn_saved = _StagedBatchNorm.forward(y.detach().reshape(B * To, C), _d(gamma), _d(beta), running_mean, running_var, eps, momentum,
                                               reduce_stats)
        return h.reshape(B, To, C).contiguous(), ((xr, wr, br), y, bn_saved)

    @staticmethod
    def backward(dh, saved, reduce_stats=None):
        (xr, wr, br), y, bn_saved = saved
        B, To, C = y.shape
        dy, dgamma, dbeta = _StagedBatchNorm.backward(_d(dh).reshape(B * To, C), bn_saved, reduce_stats)
        with torch.enable_grad():
            y.backward(dy.reshape(B, To, C))
        return xr.grad, wr.grad.reshape(C, -1), br.grad, dgamma, dbeta


class SubsampleTrain:
    @staticmethod
    def forward(mel, w, b, gamma, beta, running_mean, running_var, precision, eps=1e-5, momentum=0.1, reduce_stats=None):
        wr, br = [_d(t).clone().requires_grad_(True) for t in (w, b)]
        B, Fm, T = mel.shape
        C = w.shape[0]
        with torch.enable_grad():
            conv = F.conv2d(_d(mel).unsqueeze(1), wr, br, stride=2, padding=1)       # [B, C, F/2, To]
        F2, To = conv.shape[2], conv.shape[3]
        y = conv.detach().permute(0, 2, 3, 1).reshape(B * F2 * To, C)                 # statistics over (b, f, t) per channel
        h, bn_saved = _StagedBatchNorm.forward(y, _d(gamma), _d(beta), running_mean, running_var, eps, momentum, reduce_stats)
        out = h.reshape(B, F2, To, C).permute(0, 2, 3, 1).reshape(B * To, C * F2)     # [(b, t), c * F2 + f]: layout of the following Linear
        return out.contiguous(), ((wr, br), conv, bn_saved, (B, C, F2, To))

    @staticmethod
    def backward(da, saved, reduce_stats=None):
        (wr, br), conv, bn_saved, (B, C, F2, To) = saved
        dh = _d(da).reshape(B, To, C, F2).permute(0, 3, 1, 2).reshape(B * F2 * To, C)
        dy, dgamma, dbeta = _StagedBatchNorm.backward(dh, bn_saved, reduce_stats)
        with torch.enable_grad():
            conv.backward(dy.reshape(B, F2, To, C).permute(0, 3, 1, 2))
        return wr.grad, br.grad, dgamma, dbeta


def strided_rows(x, stride, precision):
    return _d(x)[:, ::stride].contiguous()


def strided_rows_bwd(d, dx, stride):
    dx[:, ::stride] += _d(d)
    return dx


def concat_qkv(mhsa):
    w = torch.cat([_d(l.weight) for l in (mhsa.query_layer, mhsa.key_layer, mhsa.value_layer)], 0)
    b = torch.cat([_d(l.bias) for l in (mhsa.query_layer, mhsa.key_layer, mhsa.value_layer)], 0)
    return w, b


def own_f32(x):
    return _d(x).clone()


def zeros_f32(rows, cols, device):
    return torch.zeros(rows, cols, dtype=DT)


# dropout with p == 0 only (masks are a device-side hash; their statistics are tested on the GPU)
def dropout_counter(device, seed):
    raise AssertionError("the CPU glue test runs with Pdrop = 0")


def dropout_f32(x, drop, site):
    assert drop.p == 0.0
    return x


def dropout_act(x, drop, site, precision):
    assert drop.p == 0.0
    return x


def dropout_cast_scaled(x, precision, scale, drop, site):
    assert drop.p == 0.0
    return _d(x) * scale


def stats_merge_ranks(gathered, counts, out):
    """Chan merge of per-rank (mean, M2) pairs (what ec_op_stats_merge_ranks computes)."""
    g, n = gathered.double(), counts.double()
    tot = n.sum()
    mean = (g[:, 0] * n[:, None]).sum(0) / tot
    m2 = g[:, 1].sum(0) + (n[:, None] * (g[:, 0] - mean) ** 2).sum(0)
    out[0].copy_(mean.to(out.dtype)); out[1].copy_(m2.to(out.dtype))
    return out


def swish_dropout_fwd(z, drop, site, precision):
    assert drop.p == 0.0
    return swish_fwd(z, precision)


def swish_dropout_bwd(z, dy, drop, site, precision):
    assert drop.p == 0.0
    return swish_bwd(z, dy, precision)


def transpose_cast(w, precision):
    return _d(w).t().contiguous()
