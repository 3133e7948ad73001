"""Transducer joint network and RNN-T loss of the reference (reference models/joint_networks.py:32-105, models/losses.py:22-46,
models/transducer.py:88-106) on the B200, forward and backward (BASELINE.json configs[3]).

`JointNetwork` keeps the reference's constructor, `forward(f, g)` and `state_dict` names (`linear_encoder / linear_decoder /
linear_joint .weight / .bias`); the modules below are parameter holders, the arithmetic runs in the CUDA library:
    fe = Linear_enc(f), gd = Linear_dec(g)                       tcgen05 GEMM (ec_op_gemm)
    H[(b,t,u)] = act(fe[b,t] + gd[b,u]) in the activation type   ec_op_joint_hidden (the reference `repeat`s both operands to
                                                                 (B, T, U+1, J) fp32 before adding them)
    logits = H W_joint^T + b                                     tcgen05 GEMM
    loss = mean_b -log p(y_b | x_b)                              ec_rnnt_loss(_grad): log-sum-exp + gather of the two values per lattice
                                                                 node the recursion reads, alpha / beta wavefronts, gradient with the
                                                                 log_softmax folded in (no (B,T,U+1,V) log-prob tensor)
Autograd: `JointNetwork.forward` returns logits carrying ONE autograd node whose backward runs the data / weight gradient GEMMs
(tcgen05) and `ec_op_joint_hidden_bwd`; `LossRNNT` returns the loss with a node that hands back the precomputed d loss / d logits --
the reference's own two-call structure (`Transducer.forward` -> `criterion`), so its trainer works unchanged.  The prediction network
(`models/decoders.py`, an LSTM) is outside the path and stays PyTorch.  There is no CPU fallback."""
import torch
import torch.nn as nn

from . import _lib
from . import ops as _ops

_ACTS = {None: 0, "tanh": 1, "relu": 2, "swish": 3}


def _pad_rows(w, bias, mult=8):
    """TMA needs 16-byte row pitches on the [rows, V] logits: pad the vocabulary with zero rows (sliced off again by the caller)."""
    V = w.shape[0]
    if V % mult == 0:
        return w, bias, V
    Vp = (V + mult - 1) // mult * mult
    return torch.cat([w, w.new_zeros(Vp - V, w.shape[1])]), torch.cat([bias, bias.new_zeros(Vp - V)]), V


class _JointFn(torch.autograd.Function):
    """logits = Linear_joint(act(Linear_enc(f)[:, :, None] + Linear_dec(g)[:, None])) with a hand-scheduled CUDA backward."""

    @staticmethod
    def forward(ctx, f, g, we, be, wd, bd, wj, bj, act_id, pr):
        B, T, _ = f.shape
        U1 = g.shape[1]
        J = we.shape[0]
        prec = _lib.PRECISIONS[pr]
        f_act, g_act = _ops.cast(f.reshape(B * T, -1), pr), _ops.cast(g.reshape(B * U1, -1), pr)
        fe = _ops.gemm(f_act, _ops.cast_weight(we, pr), be, pr)[0]
        gd = _ops.gemm(g_act, _ops.cast_weight(wd, pr), bd, pr)[0]
        H = torch.empty(B * T * U1, J, dtype=_lib.act_dtype(prec), device=f.device)
        _lib.check(_lib.lib().ec_op_joint_hidden(prec, _lib.ptr(fe), _lib.ptr(gd), B, T, U1, J, act_id, _lib.ptr(H), _lib.stream_ptr()))
        wjp, bjp, V = _pad_rows(wj, bj)
        logits = _ops.gemm(H, _ops.cast_weight(wjp, pr), bjp, pr)[0]
        ctx.save_for_backward(f_act, g_act, H, we, wd, wjp)
        ctx.meta = (B, T, U1, J, V, act_id, pr)
        return logits.view(B, T, U1, -1)[..., :V]

    @staticmethod
    def backward(ctx, d_logits):
        f_act, g_act, H, we, wd, wjp = ctx.saved_tensors
        B, T, U1, J, V, act_id, pr = ctx.meta
        if act_id == 3:
            raise NotImplementedError("joint backward: Swish would need the stored pre-activation (no shipped config uses it)")
        prec = _lib.PRECISIONS[pr]
        Vp = wjp.shape[0]
        dl = d_logits.reshape(B * T * U1, V).float()
        if Vp != V:
            dl = torch.cat([dl, dl.new_zeros(dl.shape[0], Vp - V)], dim=1)
        dl_act = _ops.cast(dl.contiguous(), pr)
        dwj, dbj = _ops.linear_wgrad_bias(dl_act, H, pr)                        # [Vp, J], [Vp]
        dH = _ops.gemm(dl_act, _ops.transpose_cast(wjp, pr), None, pr)[0]       # [rows, J] fp32
        dfe = torch.empty(B * T, J, dtype=torch.float32, device=dH.device)
        dgd = torch.empty(B * U1, J, dtype=torch.float32, device=dH.device)
        _lib.check(_lib.lib().ec_op_joint_hidden_bwd(prec, _lib.ptr(H), _lib.ptr(dH), B, T, U1, J, act_id, _lib.ptr(dfe), _lib.ptr(dgd), _lib.stream_ptr()))
        dfe_act, dgd_act = _ops.cast(dfe, pr), _ops.cast(dgd, pr)
        dwe, dbe = _ops.linear_wgrad_bias(dfe_act, f_act, pr)
        dwd, dbd = _ops.linear_wgrad_bias(dgd_act, g_act, pr)
        df = _ops.gemm(dfe_act, _ops.transpose_cast(we, pr), None, pr)[0].view(B, T, -1) if ctx.needs_input_grad[0] else None
        dg = _ops.gemm(dgd_act, _ops.transpose_cast(wd, pr), None, pr)[0].view(B, U1, -1) if ctx.needs_input_grad[1] else None
        return df, dg, dwe, dbe, dwd, dbd, dwj[:V], dbj[:V], None, None


class JointNetwork(nn.Module):
    """Drop-in for reference models/joint_networks.py JointNetwork (joint_mode "sum", the mode of every shipped config)."""

    def __init__(self, dim_encoder, dim_decoder, vocab_size, params, precision="bf16x2"):
        super().__init__()
        if params["act"] not in _ACTS:
            raise ValueError("joint activation must be tanh, relu, swish or None")
        if params["joint_mode"] != "sum":
            raise NotImplementedError('joint_mode "concat" is not used by any shipped config')
        if params["dim_model"] is None:
            raise NotImplementedError("joint network without projection layers is not used by any shipped config")
        J = params["dim_model"]
        self.linear_encoder = nn.Linear(dim_encoder, J)
        self.linear_decoder = nn.Linear(dim_decoder, J)
        self.linear_joint = nn.Linear(J, vocab_size)
        self.joint_mode = "sum"
        self.act_id = _ACTS[params["act"]]
        self.precision = precision

    def forward(self, f, g):
        """Training / eval-loss form: f (B, T, Denc), g (B, U+1, Ddec) -> logits (B, T, U+1, V).  Decoding form: f (B, Denc), g (B, Ddec)
        -> (B, V) (reference models/joint_networks.py:80-105)."""
        for t in (f, g):
            if not t.is_cuda:
                raise RuntimeError("effconf_b200 JointNetwork runs on CUDA sm_100 only (no CPU path)")
        squeeze = f.dim() == 2
        if squeeze:
            f, g = f.unsqueeze(1), g.unsqueeze(1)
        logits = _JointFn.apply(f.float(), g.float(), self.linear_encoder.weight, self.linear_encoder.bias, self.linear_decoder.weight,
                                self.linear_decoder.bias, self.linear_joint.weight, self.linear_joint.bias, self.act_id, self.precision)
        return logits[:, 0, 0] if squeeze else logits


def _rnnt(logits, labels, frame_len, label_len, blank, want_grad):
    if not logits.is_cuda:
        raise RuntimeError("effconf_b200 RNN-T loss runs on CUDA only")
    logits = logits.detach().float().contiguous()
    B, T, U1, V = logits.shape
    dev = logits.device
    labels = labels.to(dev, torch.int64).contiguous()
    if labels.dim() != 2 or labels.shape[1] < U1 - 1:
        raise ValueError("labels must hold at least U = logits.shape[2] - 1 columns")
    frame_len = frame_len.to(dev, torch.int64).contiguous()
    label_len = label_len.to(dev, torch.int64).contiguous()
    L = _lib.lib()
    scratch = torch.empty(L.ec_rnnt_scratch_bytes(B, T, U1), dtype=torch.uint8, device=dev)
    per = torch.empty(B, dtype=torch.float32, device=dev)
    mean = torch.empty((), dtype=torch.float32, device=dev)
    grad = torch.empty_like(logits) if want_grad else None
    with torch.cuda.device(dev):
        if want_grad:
            _lib.check(L.ec_rnnt_loss_grad(_lib.ptr(logits), B, T, U1, V, _lib.ptr(labels), labels.shape[1], _lib.ptr(frame_len), _lib.ptr(label_len),
                                           blank, _lib.ptr(scratch), _lib.ptr(per), _lib.ptr(mean), 1.0 / B, _lib.ptr(grad), _lib.stream_ptr()))
        else:
            _lib.check(L.ec_rnnt_loss(_lib.ptr(logits), B, T, U1, V, _lib.ptr(labels), labels.shape[1], _lib.ptr(frame_len), _lib.ptr(label_len), blank,
                                      _lib.ptr(scratch), _lib.ptr(per), _lib.ptr(mean), _lib.stream_ptr()))
    return mean, per, grad


def rnnt_loss(logits, labels, frame_len, label_len, blank=0):
    """logits (B, T, U+1, V) fp32 CUDA, labels (B, U) int, frame_len / label_len (B,) -> (mean loss (), per-utterance losses (B,))."""
    mean, per, _ = _rnnt(logits, labels, frame_len, label_len, blank, False)
    return mean, per


def rnnt_loss_and_grad(logits, labels, frame_len, label_len, blank=0):
    """-> (mean loss, per-utterance losses, d mean / d logits (B, T, U+1, V))."""
    return _rnnt(logits, labels, frame_len, label_len, blank, True)


class _RNNTLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, frame_len, label_len):
        mean, _, grad = _rnnt(logits, labels, frame_len, label_len, 0, True)
        ctx.save_for_backward(grad)
        return mean

    @staticmethod
    def backward(ctx, d_loss):
        (grad,) = ctx.saved_tensors
        return grad * d_loss, None, None, None


class LossRNNT(nn.Module):
    """Drop-in for reference models/losses.py LossRNNT.forward(batch, pred): mean negative log likelihood (blank 0, no frame averaging)."""

    def forward(self, batch, pred):
        x, y, x_len, y_len = batch
        outputs_pred, f_len, _ = pred
        if torch.is_grad_enabled() and outputs_pred.requires_grad:
            return _RNNTLossFn.apply(outputs_pred, y, f_len, y_len)
        return rnnt_loss(outputs_pred, y, f_len, y_len)[0]
