"""Host-side mirror of the reference CTC task model for the hot path (reference models/model_ctc.py:37-136,
models/losses.py:48-71): `ModelCTC.forward(batch) -> (logits, logits_len, attentions)`, `LossCTC()(batch, pred)`,
`gready_search_decoding(x, x_len)` (token ids; tokenizer.decode applied when a tokenizer is given).

The trainer base class (optimiser, schedules, checkpoints, WER: reference models/model.py) is out of scope and stays
the reference's own Python; `patch_reference()` in dropin.py swaps only the encoder class into it."""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .encoders import ConformerEncoder


class LossCTC(nn.Module):
    """reference models/losses.py:48-71: log_softmax -> CTCLoss(blank=0, reduction='none', zero_infinity=False) -> mean,
    on device in one call.  When the logits require grad the loss carries an autograd node whose backward is the CUDA
    alpha-beta kernel (ec_ctc_loss_grad): `loss.backward()` fills `logits.grad` like the reference's criterion does."""

    def forward(self, batch, pred):
        _, y, _, y_len = batch
        logits, f_len, _ = pred
        if torch.is_grad_enabled() and logits.requires_grad:
            return _CTCLossFn.apply(logits, f_len, y, y_len)
        return ctc_loss(logits, f_len, y, y_len)[0]


def _scratch(B, T, V, device):
    return torch.empty(_lib.lib().ec_ctc_scratch_bytes(B, T, V), dtype=torch.uint8, device=device)


def ctc_loss(logits, logits_len, targets, target_len):
    """Returns (mean loss (), per-utterance losses (B,)), fp32 CUDA tensors."""
    if not logits.is_cuda:
        raise RuntimeError("effconf_b200 CTC loss runs on CUDA only")
    logits = logits.float().contiguous()
    B, T, V = logits.shape
    dev = logits.device
    logits_len = logits_len.to(dev, torch.int64).contiguous()
    targets = targets.to(dev, torch.int64).contiguous()
    target_len = target_len.to(dev, torch.int64).contiguous()
    per = torch.empty(B, dtype=torch.float32, device=dev)
    mean = torch.empty((), dtype=torch.float32, device=dev)
    scratch = _scratch(B, T, V, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ec_ctc_loss(_lib.ptr(logits), B, T, V, _lib.ptr(logits_len), _lib.ptr(targets), targets.shape[1],
                                          _lib.ptr(target_len), _lib.ptr(scratch), _lib.ptr(per), _lib.ptr(mean), _lib.stream_ptr()))
    return mean, per


def ctc_loss_and_grad(logits, logits_len, targets, target_len):
    """Returns (mean loss (), per-utterance losses (B,), d mean / d logits (B, T, V)), fp32 CUDA tensors (ec_ctc_loss_grad)."""
    if not logits.is_cuda:
        raise RuntimeError("effconf_b200 CTC loss runs on CUDA only")
    logits = logits.detach().float().contiguous()
    B, T, V = logits.shape
    dev = logits.device
    logits_len = logits_len.to(dev, torch.int64).contiguous()
    targets = targets.to(dev, torch.int64).contiguous()
    target_len = target_len.to(dev, torch.int64).contiguous()
    per = torch.empty(B, dtype=torch.float32, device=dev)
    mean = torch.empty((), dtype=torch.float32, device=dev)
    grad = torch.empty(B, T, V, dtype=torch.float32, device=dev)
    scratch = _scratch(B, T, V, dev)
    L = _lib.lib()
    work = torch.empty(L.ec_ctc_grad_work_bytes(B, T, targets.shape[1]), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.ec_ctc_loss_grad(_lib.ptr(logits), B, T, V, _lib.ptr(logits_len), _lib.ptr(targets), targets.shape[1],
                                      _lib.ptr(target_len), _lib.ptr(scratch), _lib.ptr(work), _lib.ptr(per), _lib.ptr(mean),
                                      _lib.ptr(grad), _lib.stream_ptr()))
    return mean, per, grad


class _CTCLossFn(torch.autograd.Function):
    """autograd node of LossCTC: forward and gradient come from one kernel sequence (the gradient is kept for backward)."""

    @staticmethod
    def forward(ctx, logits, logits_len, targets, target_len):
        mean, _, grad = ctc_loss_and_grad(logits, logits_len, targets, target_len)
        ctx.save_for_backward(grad)
        ctx.in_dtype = logits.dtype
        return mean

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return (grad * grad_out).to(ctx.in_dtype), None, None, None


def greedy_ids(logits, logits_len):
    """Collapsed greedy token ids per utterance (reference models/model_ctc.py:99-133), one D2H copy in total."""
    logits = logits.float().contiguous()
    B, T, V = logits.shape
    dev = logits.device
    logits_len = logits_len.to(dev, torch.int64).contiguous()
    ids = torch.empty(B, T, dtype=torch.int32, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    scratch = _scratch(B, T, V, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ec_ctc_greedy(_lib.ptr(logits), B, T, V, _lib.ptr(logits_len), _lib.ptr(scratch), _lib.ptr(ids),
                                            _lib.ptr(counts), _lib.stream_ptr()))
    ids_h, counts_h = ids.cpu(), counts.cpu()
    return [ids_h[b, :int(counts_h[b])].tolist() for b in range(B)]


class ModelCTC(nn.Module):
    """Encoder + fc + CTC criterion with the reference's forward contract (reference models/model_ctc.py:37-68)."""

    def __init__(self, encoder_params, tokenizer_params, training_params=None, decoding_params=None, name="model",
                 precision: str = "auto", tokenizer=None):
        super().__init__()
        if encoder_params["arch"] != "Conformer":
            raise Exception("Unknown encoder architecture:", encoder_params["arch"])
        self.encoder = ConformerEncoder(encoder_params, precision=precision)
        d_last = encoder_params["dim_model"][-1] if isinstance(encoder_params["dim_model"], list) else encoder_params["dim_model"]
        self.fc = nn.Linear(d_last, tokenizer_params["vocab_size"])
        self.encoder.attach_head(self.fc)
        self.criterion = LossCTC()
        self.tokenizer = tokenizer
        self.name = name

    def forward(self, batch):
        x, _, x_len, _ = batch
        mel, mel_len = self.encoder.preprocessing(x.float(), x_len)
        if self.training:                                   # reference models/encoders.py:103-104 (SpecAugment inside the train-mode forward)
            mel = self.encoder.augment(mel, mel_len)
        return self.forward_mel(mel.contiguous(), mel_len)

    def forward_mel(self, mel, mel_len=None):
        _, out_len, logits = self.encoder.forward_mel(mel, mel_len, want_logits=True)
        return logits, out_len, [None] * len(self.encoder.blocks)

    def gready_search_decoding(self, x, x_len):
        logits, logits_len, _ = self.forward((x, None, x_len, None))
        if logits_len is None:
            logits_len = torch.full((logits.shape[0],), logits.shape[1], dtype=torch.int64, device=logits.device)
        ids = greedy_ids(logits, logits_len)
        return self.tokenizer.decode(ids) if self.tokenizer is not None else ids
