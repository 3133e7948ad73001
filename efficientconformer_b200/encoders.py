"""Drop-in replacement for the reference `ConformerEncoder` (reference models/encoders.py:44-142).

Same constructor (`ConformerEncoder(params)` with the `encoder_params` dict), same
`forward(x, x_len) -> (x, x_len, attentions)`, same `state_dict()` names / shapes (verified against
the reference in tests/test_oracle_golden.py and tests/test_host_logic.py), a `.blocks` list whose items carry `.stride`.
The modules below are PARAMETER HOLDERS only: nothing here runs PyTorch math on the hot path.  `forward` hands raw
device pointers to the sm_100a CUDA library (efficientconformer_b200/csrc, C ABI in include/effconf_b200.h); the audio
front end (STFT -> mel -> log, reference models/modules.py:87-106) is one kernel for audio on the GPU and torchaudio for CPU tensors.

`.eval()`: the fused inference engine (ec_engine_forward, CUDA-graph replay) -- what `Model.evaluate`, `gready_search_decoding`
and `eval_time_encoder` use.  `.train()`: the training operator schedule of efficientconformer_b200/training.py (batch-statistics
BatchNorm, dropout, one autograd node whose backward is the CUDA backward schedule); SpecAugment
(one kernel, batched; csrc/frontend.cu) is applied inside `forward` when training, as the reference does.  There is no fallback path: without the CUDA library or on a
non-sm_100 device, forward raises.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .config import resolve_blocks, BlockSpec


# --------------------------------------------------------------------------------------------------------------------
# parameter holders (never called) laid out so that state_dict() keys equal the reference's
# --------------------------------------------------------------------------------------------------------------------
def _seq(n, mods):
    """nn.Sequential of length n with parameterless placeholders except at the given indices."""
    return nn.Sequential(*[mods.get(i, nn.Identity()) for i in range(n)])


class _FeedForwardHolder(nn.Module):     # reference models/modules.py:362-395: layers.0 LN, .1 Linear, .4 Linear
    def __init__(self, dim, ratio):
        super().__init__()
        self.layers = _seq(6, {0: nn.LayerNorm(dim, eps=1e-6), 1: nn.Linear(dim, ratio * dim), 4: nn.Linear(ratio * dim, dim)})


class _RelPosAttentionHolder(nn.Module):  # reference models/attentions.py:451-478 (+ grouped :622-643)
    def __init__(self, dim, heads, group):
        super().__init__()
        self.query_layer = nn.Linear(dim, dim)
        self.key_layer = nn.Linear(dim, dim)
        self.value_layer = nn.Linear(dim, dim)
        self.output_layer = nn.Linear(dim, dim)
        self.pos_layer = nn.Linear(dim, dim)
        self.u = nn.Parameter(torch.empty(dim))
        self.v = nn.Parameter(torch.empty(dim))
        dh = dim // heads                      # the reference initialises u, v as (H, D/H) xavier-uniform matrices
        with torch.no_grad():
            nn.init.xavier_uniform_(self.u.view(heads, dh))
            nn.init.xavier_uniform_(self.v.view(heads, dh))


class _AttentionModuleHolder(nn.Module):  # reference models/modules.py:397-470
    def __init__(self, dim, heads, group):
        super().__init__()
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.mhsa = _RelPosAttentionHolder(dim, heads, group)


class _ConvModuleHolder(nn.Module):       # reference models/modules.py:490-522: layers.0 LN, .2 pw1, .4 depthwise, .5 BN, .7 pw2
    def __init__(self, dim, dim_expand, kernel_size, stride):
        super().__init__()
        self.layers = _seq(10, {
            0: nn.LayerNorm(dim, eps=1e-6),
            2: nn.Conv1d(dim, 2 * dim_expand, 1),
            4: nn.Conv1d(dim_expand, dim_expand, kernel_size, stride=stride, groups=dim_expand),
            5: nn.BatchNorm1d(dim_expand),
            7: nn.Conv1d(dim_expand, dim_expand, 1),
        })


class ConformerBlockHolder(nn.Module):    # reference models/blocks.py:32-117
    def __init__(self, spec: BlockSpec):
        super().__init__()
        D, De = spec.dim_model, spec.dim_expand
        self.feed_forward_module1 = _FeedForwardHolder(D, spec.ff_ratio)
        self.multi_head_self_attention_module = _AttentionModuleHolder(D, spec.num_heads, spec.group_size)
        self.convolution_module = _ConvModuleHolder(D, De, spec.kernel_size, spec.conv_stride)
        self.feed_forward_module2 = _FeedForwardHolder(De, spec.ff_ratio)
        self.norm = nn.LayerNorm(De, eps=1e-6)
        if D != De:
            self.conv_res = _seq(3, {1: nn.Conv1d(D, De, 1, stride=spec.conv_stride)})
        self.stride = spec.conv_stride    # att_stride is 1 for every supported config (reference blocks.py:117)
        self.spec = spec


def _device_front_end():
    """EFFCONF_DEVICE_FRONTEND=0 keeps torchaudio / torch ops for the front end on CUDA tensors as well (diagnostics)."""
    import os
    return os.environ.get("EFFCONF_DEVICE_FRONTEND", "1") != "0"


class _PreprocessingHolder(nn.Module):
    """reference models/modules.py:55-106 (AudioPreprocessing): Spectrogram + MelScale + log.  The torchaudio transforms are kept as the
    holders of the module's state (Hann window, mel filter bank) and as the path for CPU tensors; audio that is
    already on the GPU goes through ONE kernel (ec_op_logmel, csrc/frontend.cu: framing, window, shared-memory FFT, power, mel, log)
    with the same window and filter-bank buffers."""

    def __init__(self, params):
        super().__init__()
        import torchaudio
        self.win_length = int(params["sample_rate"] * params["win_length_ms"]) // 1000
        self.hop_length = int(params["sample_rate"] * params["hop_length_ms"]) // 1000
        self.Spectrogram = torchaudio.transforms.Spectrogram(params["n_fft"], self.win_length, self.hop_length)
        self.MelScale = torchaudio.transforms.MelScale(params["n_mels"], params["sample_rate"], f_min=0, f_max=8000,
                                                       n_stft=params["n_fft"] // 2 + 1)
        self.normalize, self.mean, self.std = params["normalize"], params["mean"], params["std"]

    def _device_tables(self, device):
        win, fb = self.Spectrogram.window, self.MelScale.fb
        key = (str(device), win.data_ptr(), fb.data_ptr())
        if getattr(self, "_tables_key", None) != key:
            n_fft = self.Spectrogram.n_fft
            left = (n_fft - win.numel()) // 2
            w = torch.nn.functional.pad(win.detach().float().to(device), (left, n_fft - win.numel() - left)).contiguous()
            f = fb.detach().float().to(device).contiguous()
            nz = f > 0                                        # filter m covers the bins [first, last] with a non-zero weight
            idx = torch.arange(f.shape[0], device=device)[:, None]
            lo = torch.where(nz, idx, f.shape[0]).amin(0)
            hi = torch.where(nz, idx + 1, 0).amax(0)
            kr = torch.stack([torch.minimum(lo, hi), hi], 1).to(torch.int32).contiguous()
            self._tables, self._tables_key = (w, f, kr), key
        return self._tables

    def forward(self, x, x_len):
        if x.is_cuda and _device_front_end():
            from . import ops
            w, f, kr = self._device_tables(x.device)
            with torch.cuda.device(x.device):
                x = ops.logmel(x.detach().float().reshape(-1, x.shape[-1]), w, f, kr, self.hop_length, self.normalize, self.mean,
                               self.std).reshape(*x.shape[:-1], f.shape[1], -1)
            if x_len is not None:
                x_len = torch.div(x_len, self.hop_length, rounding_mode="floor") + 1
            return x, x_len
        x = self.MelScale(self.Spectrogram(x))
        x = (x.float() + 1e-9).log().type(x.dtype)
        if x_len is not None:
            x_len = torch.div(x_len, self.hop_length, rounding_mode="floor") + 1
        if self.normalize:
            x = (x - self.mean) / self.std
        return x, x_len


class SpecAugment(nn.Module):
    """SpecAugment with the reference's semantics (reference models/modules.py:108-151), called inside
    the train-mode forward like reference models/encoders.py:103-104:
      * mF frequency masks shared by the whole batch, width floor(U(0, F)), start floor(U(0, n_mels - width))
        (torchaudio FrequencyMasking(F, iid_masks=False) -> mask_along_axis);
      * per utterance b, mT time masks inside its valid frames [0, x_len[b]), width floor(U(0, int(pS * x_len[b]))),
        start floor(U(0, x_len[b] - width)); masked cells are set to 0.
    The reference loops over the batch with one host read of x_len[b] per utterance; here all B * mT masks are drawn and applied with
    batched device ops (no host synchronisation).  Random streams differ from the reference's, the distribution is the same.
    fp32 CUDA tensors: one kernel (ec_op_specaugment) draws every mask from the counter-based hash the dropout sites use (seeded from
    torch.initial_seed(), {seed, step} on the device) and writes only the masked cells of a copy; other tensors: the torch ops below."""

    def __init__(self, spec_augment, mF, F, mT, pS):
        super().__init__()
        self.spec_augment, self.mF, self.F, self.mT, self.pS = bool(spec_augment), int(mF), int(F), int(mT), float(pS)

    def forward(self, x, x_len=None):
        if not self.spec_augment:
            return x
        B, n_mels, T = x.shape
        dev = x.device
        if x.is_cuda and x.dtype == torch.float32 and _device_front_end():
            from . import ops
            with torch.cuda.device(dev):
                ctr = getattr(self, "_counter", None)
                if ctr is None or ctr.device != dev:
                    ctr = self._counter = ops.dropout_counter(dev, torch.initial_seed())
                ops.dropout_advance(ctr)
                lens = None if x_len is None else x_len.to(device=dev, dtype=torch.int64).contiguous()
                return ops.specaugment_(x.detach().clone(memory_format=torch.contiguous_format), lens, self.mF, min(self.F, n_mels), self.mT,
                                        self.pS, ctr)
        keep = torch.ones(B, n_mels, T, dtype=torch.bool, device=dev)
        if self.mF > 0:
            value = torch.rand(self.mF, device=dev) * self.F
            start = (torch.rand(self.mF, device=dev) * (n_mels - value)).long()
            end = start + value.long()
            f = torch.arange(n_mels, device=dev)[None, :]
            fmask = ((f >= start[:, None]) & (f < end[:, None])).any(0)                     # (n_mels,)
            keep &= ~fmask[None, :, None]
        if self.mT > 0:
            lens = (x_len.to(dev) if x_len is not None else torch.full((B,), T, device=dev)).to(torch.float32)
            t_param = torch.floor(self.pS * lens)                                           # int(pS * x_len[b])
            value = torch.rand(B, self.mT, device=dev) * t_param[:, None]
            start = (torch.rand(B, self.mT, device=dev) * (lens[:, None] - value)).long()
            end = start + value.long()
            t = torch.arange(T, device=dev)[None, None, :]
            tmask = ((t >= start[:, :, None]) & (t < end[:, :, None])).any(1)              # (B, T)
            keep &= ~tmask[:, None, :]
        return x * keep.to(x.dtype)


class _SubsamplingHolder(nn.Module):      # reference models/modules.py:201-230 (one or two strided Conv2d layers)
    def __init__(self, params):
        super().__init__()
        filters, ks = params["subsampling_filters"], params["subsampling_kernel_size"]
        self.layers = nn.ModuleList([nn.Sequential(
            nn.Conv2d(1 if l == 0 else filters[l - 1], filters[l], ks, stride=2, padding=(ks - 1) // 2),
            nn.BatchNorm2d(filters[l]), nn.Identity()) for l in range(params["subsampling_layers"])])


def relative_sinusoid_rows(t_pad: int, dim: int, group: int, max_len: int) -> torch.Tensor:
    """fp32 rows [max_len - t_pad + G//2, max_len - G%2 + t_pad - G//2) of the reference's relative sinusoidal table
    (reference models/attentions.py:1209-1257, 1268-1315); row r encodes position max_len-1-r.  Computed with the same
    fp32 expression as the reference so the values are identical."""
    r0 = max_len - t_pad + group // 2
    r1 = max_len - group % 2 + t_pad - group // 2
    if r0 < 0:
        raise ValueError(f"sequence of {t_pad} frames exceeds max_pos_encoding {max_len}")
    pos = (max_len - 1 - torch.arange(r0, r1, dtype=torch.float32)).unsqueeze(1)
    angles = pos / 10000 ** (2 * torch.arange(0, dim // 2, dtype=torch.float32).unsqueeze(0) / dim)
    table = torch.zeros(r1 - r0, dim, dtype=torch.float32)
    table[:, 0::2] = angles.sin()
    table[:, 1::2] = angles.cos()
    return table


# --------------------------------------------------------------------------------------------------------------------
class _ShapePlan:
    """Device buffers for one (precision, batch, t_mel): workspace, relative tables, static I/O and the CUDA graph."""
    __slots__ = ("workspace", "tables", "table_ptrs", "mel", "x_len", "out_x", "logits", "out_len", "graph", "t_out", "has_len")


class ConformerEncoder(nn.Module):
    """B200-native Efficient Conformer encoder with the reference's API (reference models/encoders.py:44-142)."""

    def __init__(self, params, precision: str = "auto", use_cuda_graph: bool = True):
        super().__init__()
        if params.get("subsampling_module") != "Conv2d" or params.get("subsampling_layers") not in (1, 2) \
                or params.get("subsampling_kernel_size") != 3:
            raise NotImplementedError("supported front ends: one (Efficient Conformer) or two (Conformer) 3x3 Conv2d subsampling layers")
        if params.get("subsampling_norm") != "batch" or params.get("subsampling_act") != "swish":
            raise NotImplementedError("only subsampling_norm=batch / subsampling_act=swish (all shipped configs)")
        self.params = dict(params)
        self.specs = resolve_blocks(params)
        self.preprocessing = _PreprocessingHolder(params)
        self.augment = SpecAugment(params.get("spec_augment", False), params.get("mF", 0), params.get("F", 0), params.get("mT", 0),
                                   params.get("pS", 0.0))
        self.subsampling_module = _SubsamplingHolder(params)
        feat = params["subsampling_filters"][-1] * params["n_mels"] // 2 ** params["subsampling_layers"]
        self.linear = nn.Linear(feat, self.specs[0].dim_model)
        self.blocks = nn.ModuleList([ConformerBlockHolder(s) for s in self.specs])
        assert precision in ("auto", "tf32", "bf16", "bf16x2")
        self.precision = precision
        self.use_cuda_graph = use_cuda_graph
        self._head = None            # optional (weight, bias) Parameters of the CTC fc layer, attached by ModelCTC
        self._engines = {}           # precision id -> [engine handle, arena tensor, weights fingerprint]
        self._plans = {}
        self._weights_epoch = 0      # bumped by writers that bypass tensor._version (CTCTrainStep updates parameters through raw pointers)

    # ---- engine plumbing ------------------------------------------------------------------------------------------
    def attach_head(self, fc: nn.Linear):
        """Lets the engine also run the CTC `fc` GEMM (reference models/model_ctc.py:49,66)."""
        object.__setattr__(self, "_head", fc)
        self._drop_engines()

    def _drop_engines(self):
        for eng, _, _ in self._engines.values():
            _lib.lib().ec_engine_destroy(eng)
        self._engines.clear()
        self._plans.clear()

    def __del__(self):
        try:
            self._drop_engines()
        except Exception:
            pass

    def _config_struct(self):
        cfg = _lib.Config()
        cfg.n_mels = self.params["n_mels"]
        cfg.sub_filters = self.params["subsampling_filters"][0]
        cfg.sub_layers = self.params["subsampling_layers"]
        cfg.sub_filters2 = self.params["subsampling_filters"][1] if cfg.sub_layers == 2 else 0
        cfg.num_blocks = len(self.specs)
        cfg.vocab = self._head.out_features if self._head is not None else 0
        for i, s in enumerate(self.specs):
            b = cfg.blocks[i]
            b.dim_model, b.dim_expand, b.num_heads, b.kernel_size = s.dim_model, s.dim_expand, s.num_heads, s.kernel_size
            b.group_size, b.conv_stride, b.ff_ratio = s.group_size, s.conv_stride, s.ff_ratio
        return cfg

    def _raw_weights(self):
        raw = _lib.RawWeights()
        keep = []                    # keeps fp32-contiguous temporaries alive until prepare has been enqueued

        def p(t):
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.detach().float().contiguous(); keep.append(t)
            return t.data_ptr()
        sub = self.subsampling_module.layers[0]
        raw.sub_conv_w, raw.sub_conv_b = p(sub[0].weight), p(sub[0].bias)
        raw.sub_bn_w, raw.sub_bn_b, raw.sub_bn_rm, raw.sub_bn_rv = p(sub[1].weight), p(sub[1].bias), p(sub[1].running_mean), p(sub[1].running_var)
        if len(self.subsampling_module.layers) == 2:
            sub = self.subsampling_module.layers[1]
            raw.sub2_conv_w, raw.sub2_conv_b = p(sub[0].weight), p(sub[0].bias)
            raw.sub2_bn_w, raw.sub2_bn_b, raw.sub2_bn_rm, raw.sub2_bn_rv = p(sub[1].weight), p(sub[1].bias), p(sub[1].running_mean), p(sub[1].running_var)
        raw.lin_w, raw.lin_b = p(self.linear.weight), p(self.linear.bias)
        if self._head is not None:
            raw.fc_w, raw.fc_b = p(self._head.weight), p(self._head.bias)
        for i, blk in enumerate(self.blocks):
            r = raw.blocks[i]
            for tag, holder in (("ffn1", blk.feed_forward_module1), ("ffn2", blk.feed_forward_module2)):
                f = getattr(r, tag); L = holder.layers
                f.ln_w, f.ln_b, f.w1, f.b1, f.w2, f.b2 = p(L[0].weight), p(L[0].bias), p(L[1].weight), p(L[1].bias), p(L[4].weight), p(L[4].bias)
            m = blk.multi_head_self_attention_module
            r.att_ln_w, r.att_ln_b, r.u, r.v = p(m.norm.weight), p(m.norm.bias), p(m.mhsa.u), p(m.mhsa.v)
            r.wq, r.bq = p(m.mhsa.query_layer.weight), p(m.mhsa.query_layer.bias)
            r.wk, r.bk = p(m.mhsa.key_layer.weight), p(m.mhsa.key_layer.bias)
            r.wv, r.bv = p(m.mhsa.value_layer.weight), p(m.mhsa.value_layer.bias)
            r.wo, r.bo = p(m.mhsa.output_layer.weight), p(m.mhsa.output_layer.bias)
            r.wpos, r.bpos = p(m.mhsa.pos_layer.weight), p(m.mhsa.pos_layer.bias)
            L = blk.convolution_module.layers
            r.conv_ln_w, r.conv_ln_b = p(L[0].weight), p(L[0].bias)
            r.pw1_w, r.pw1_b, r.dw_w, r.dw_b = p(L[2].weight), p(L[2].bias), p(L[4].weight), p(L[4].bias)
            r.bn_w, r.bn_b, r.bn_rm, r.bn_rv = p(L[5].weight), p(L[5].bias), p(L[5].running_mean), p(L[5].running_var)
            r.pw2_w, r.pw2_b = p(L[7].weight), p(L[7].bias)
            r.norm_w, r.norm_b = p(blk.norm.weight), p(blk.norm.bias)
            if hasattr(blk, "conv_res"):
                r.res_w, r.res_b = p(blk.conv_res[1].weight), p(blk.conv_res[1].bias)
        return raw, keep

    def mark_weights_changed(self):
        """Parameters / running statistics were updated in place by device code that does not bump tensor._version (the native training
        step): the inference engines re-fold and re-cast their weights on the next eval-mode forward."""
        self._weights_epoch += 1

    def _fingerprint(self):
        ts = list(self.parameters()) + [b for b in self.buffers()]
        if self._head is not None:
            ts += [self._head.weight, self._head.bias]
        return (self._weights_epoch,) + tuple((t.data_ptr(), t._version) for t in ts)

    def _engine(self, prec: int, device):
        L = _lib.lib()
        if prec not in self._engines:
            _lib.check(L.ec_device_check())
            handle = C.c_void_p()
            cfg = self._config_struct()
            _lib.check(L.ec_engine_create(C.byref(cfg), prec, C.byref(handle)))
            arena = torch.empty(L.ec_engine_weight_bytes(handle), dtype=torch.uint8, device=device)
            self._engines[prec] = [handle, arena, None]
        ent = self._engines[prec]
        fp = self._fingerprint()
        if ent[2] != fp:             # first use or parameters changed: refold / recast the weights on device
            raw, keep = self._raw_weights()
            _lib.check(L.ec_engine_prepare(ent[0], C.byref(raw), _lib.ptr(ent[1]), _lib.stream_ptr()))
            torch.cuda.current_stream().synchronize()   # `keep` temporaries may be freed after this point
            ent[2] = fp
        return ent[0]

    def _plan(self, prec: int, eng, B: int, T: int, device, want_logits: bool):
        key = (prec, B, T, want_logits)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        L = _lib.lib()
        nb = len(self.specs)
        rows = (C.c_int32 * nb)(); frames = (C.c_int32 * nb)()
        _lib.check(L.ec_engine_relpos_rows(eng, T, rows, frames))
        plan = _ShapePlan()
        plan.workspace = torch.empty(L.ec_engine_workspace_bytes(eng, B, T), dtype=torch.uint8, device=device)
        plan.tables, cache = [], {}
        for i, s in enumerate(self.specs):
            G = s.group_size
            t_pad = frames[i] + (-frames[i]) % G
            k = (t_pad, s.dim_model, G, s.max_pos)
            if k not in cache:
                tab32 = relative_sinusoid_rows(*k).to(device)
                assert tab32.shape[0] == rows[i]
                tab = torch.empty(tab32.shape, dtype=_lib.act_dtype(prec), device=device)
                _lib.check(L.ec_op_cast(prec, _lib.ptr(tab32), _lib.ptr(tab), tab32.numel(), _lib.stream_ptr()))
                cache[k] = tab
            plan.tables.append(cache[k])
        plan.table_ptrs = (C.c_void_p * nb)(*[t.data_ptr() for t in plan.tables])
        plan.t_out = L.ec_engine_out_frames(eng, T)
        d_last = self.specs[-1].dim_expand
        plan.mel = torch.empty(B, self.params["n_mels"], T, dtype=torch.float32, device=device)
        plan.x_len = torch.empty(B, dtype=torch.int64, device=device)
        plan.out_x = torch.empty(B, plan.t_out, d_last, dtype=torch.float32, device=device)
        plan.logits = torch.empty(B, plan.t_out, self._head.out_features, dtype=torch.float32, device=device) if want_logits else None
        plan.out_len = torch.empty(B, dtype=torch.int64, device=device)
        plan.graph = {}
        self._plans[key] = plan
        return plan

    def _launch(self, eng, plan, B, T, has_len):
        _lib.check(_lib.lib().ec_engine_forward(eng, B, T, _lib.ptr(plan.mel), _lib.ptr(plan.x_len) if has_len else None,
                                                plan.table_ptrs, _lib.ptr(plan.workspace), _lib.ptr(plan.out_x),
                                                _lib.ptr(plan.logits), _lib.ptr(plan.out_len), _lib.stream_ptr()))

    def _select_precision(self):
        if self.precision != "auto":
            return _lib.PRECISIONS[self.precision]
        # default: the split mode (packed bf16 hi/lo operands, 16 significant bits -- the mode that meets the 1e-3 parity gate);
        # under torch.autocast the caller has asked for reduced precision: plain bf16 operands
        return _lib.PREC_BF16 if torch.is_autocast_enabled() else _lib.PREC_BF16X2

    # ---- public API -----------------------------------------------------------------------------------------------
    def forward_mel(self, mel, mel_len=None, want_logits: bool = False, clone: bool = True):
        """The hot path from the mel spectrogram on: mel (B, n_mels, T) fp32 CUDA, mel_len (B,) int64 or None.
        Returns (x (B,T_out,D_last) fp32, x_len_out, logits or None)."""
        if self.training:
            return self._forward_mel_train(mel, mel_len, want_logits)
        if not mel.is_cuda:
            raise RuntimeError("effconf_b200 runs on CUDA sm_100 only (the reference's --cpu path is the oracle's job)")
        if want_logits and self._head is None:
            raise RuntimeError("no fc head attached")
        B, F, T = mel.shape
        assert F == self.params["n_mels"]
        with torch.cuda.device(mel.device):
            prec = self._select_precision()
            eng = self._engine(prec, mel.device)
            plan = self._plan(prec, eng, B, T, mel.device, want_logits)
            plan.mel.copy_(mel)
            has_len = mel_len is not None
            if has_len:
                plan.x_len.copy_(mel_len)
            if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
                g = plan.graph.get(has_len)
                if g is None:
                    self._launch(eng, plan, B, T, has_len)          # warm-up (sets kernel attributes) outside capture
                    torch.cuda.current_stream().synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._launch(eng, plan, B, T, has_len)
                    plan.graph[has_len] = g
                g.replay()
            else:
                self._launch(eng, plan, B, T, has_len)
            x = plan.out_x.clone() if clone else plan.out_x
            lg = (plan.logits.clone() if clone else plan.logits) if want_logits else None
            out_len = plan.out_len.clone() if has_len else None
        return x, out_len, lg

    def _sync_batchnorm_group(self):
        """(found, process_group) of the first nn.SyncBatchNorm among the holders: the reference's distribute_strategy runs
        SyncBatchNorm.convert_sync_batchnorm over the encoder (reference models/model_ctc.py:73), which swaps the BatchNorm holders."""
        for m in self.modules():
            if isinstance(m, nn.SyncBatchNorm):
                return True, m.process_group
        return False, None

    def training_path(self, device=None):
        """The train-mode operator schedule (efficientconformer_b200/training.py) bound to this encoder's parameters."""
        from .training import TrainingPath
        import torch.distributed as dist
        reducer = self.__dict__.get("_stats_reducer")
        has_sync, group = self._sync_batchnorm_group()
        if reducer is None and has_sync and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            # never run local statistics silently under SyncBatchNorm holders: install the cross-rank exchange.  Ranks may hold
            # different frame counts under the reference's DistributedSampler + collate_fn_pad, hence uniform=False.
            from .distributed import SyncBatchNormReducer
            reducer = SyncBatchNormReducer(group, device if device is not None else next(self.parameters()).device, uniform=False)
            object.__setattr__(self, "_stats_reducer", reducer)
            object.__setattr__(self, "_training_path", None)
        tp = self.__dict__.get("_training_path")
        if tp is None or tp.head is not self._head:
            tp = TrainingPath(self, self._head, stats_reducer=reducer)
            object.__setattr__(self, "_training_path", tp)
        return tp

    def set_stats_reducer(self, reducer):
        """SyncBatchNorm: `reducer(stats, count)` merges BatchNorm statistics across ranks (efficientconformer_b200/distributed.py)."""
        object.__setattr__(self, "_stats_reducer", reducer)
        object.__setattr__(self, "_training_path", None)

    def _forward_mel_train(self, mel, mel_len, want_logits):
        """`.train()` semantics (batch-statistics BatchNorm with running-stat updates, dropout) -- reference models/encoders.py:106-142
        under model.train().  With grad enabled the whole path is ONE autograd node whose backward is the CUDA backward schedule."""
        from .training import EncoderTrainFn
        if not mel.is_cuda:
            raise RuntimeError("effconf_b200 runs on CUDA sm_100 only (the reference's --cpu path is the oracle's job)")
        if want_logits and self._head is None:
            raise RuntimeError("no fc head attached")
        path = self.training_path(mel.device)
        prec = self._select_precision()
        self.mark_weights_changed()          # BatchNorm running statistics are updated through raw pointers below
        mel = mel.float().contiguous()
        if mel_len is not None:
            mel_len = mel_len.to(mel.device)
        with torch.cuda.device(mel.device):
            _lib.check(_lib.lib().ec_device_check())
            plist = [p for _, p in path.param_list()]
            if torch.is_grad_enabled() and any(p.requires_grad for p in plist):
                x, logits, out_len = EncoderTrainFn.apply(path, mel, mel_len, prec, want_logits, *plist)
            else:
                with torch.no_grad():
                    x, logits, out_len, _ = path.forward(mel, mel_len, prec, want_logits)
        return x, out_len, logits

    def forward(self, x, x_len=None):
        """reference signature: x (B, L_audio) float, x_len (B,) long or None -> (x, x_len, attentions)."""
        mel, mel_len = self.preprocessing(x.float(), x_len)
        if self.training:                                   # reference models/encoders.py:103-104
            mel = self.augment(mel, mel_len)
        out, out_len, _ = self.forward_mel(mel.contiguous(), mel_len)
        # the reference returns one (B,H,T',T') attention map per block that no caller reads; they are never materialised here
        return out, out_len, [None] * len(self.blocks)
