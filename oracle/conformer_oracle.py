"""CPU oracle for the Efficient Conformer encoder hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch functional restatement (plain torch CPU tensor algebra, fp32 or fp64) of the reference's
algorithm for the one path this repository accelerates: Conv2dSubsampling -> Linear -> N x ConformerBlock
-> fc -> CTC loss / greedy ids.  Nothing here imports the reference; every function cites the reference
file:line it restates.  Parity pinning: this oracle is checked against golden vectors produced by the real
reference (imported from /root/reference in the authoring container by tests/golden/make_golden.py, which is
committed next to the fixtures) in tests/test_oracle_golden.py, so parity is PINNED.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module; the product package (efficientconformer_b200/) never does and fails loudly without its CUDA library.

Weights are passed as a dict using the reference's own state_dict names.
"""
import math
import torch
import torch.nn.functional as F

from types import SimpleNamespace


def resolve_blocks(params):
    """Per-block hyper-parameters, restated from the reference's own constructor (reference models/encoders.py:75-95): entry i of a
    list-valued parameter is selected by the number of expand / strided blocks strictly before (`>`) or up to (`>=`) block i.
    Independent of the product's efficientconformer_b200.config.resolve_blocks (tests/test_host_logic.py compares the two)."""
    def pick(v, idx):
        return v[idx] if isinstance(v, list) else v
    expand, strided = params.get("expand_blocks", []), params.get("strided_blocks", [])
    out = []
    for i in range(params["num_blocks"]):
        gt_e, ge_e = sum(i > e for e in expand), sum(i >= e for e in expand)
        gt_s = sum(i > st for st in strided)
        dim_model, dim_expand = pick(params["dim_model"], gt_e), pick(params["dim_model"], ge_e)
        heads, group = pick(params["num_heads"], gt_e), pick(params.get("att_group_size", 1), gt_s)
        out.append(SimpleNamespace(
            dim_model=dim_model, dim_expand=dim_expand, num_heads=heads, kernel_size=pick(params["kernel_size"], ge_e), group_size=group,
            max_pos=params["max_pos_encoding"] // params.get("stride", 2) ** gt_s,
            conv_stride=pick(params["conv_stride"], gt_s) if i in strided else 1, ff_ratio=params["ff_ratio"],
            dim_head=(group * dim_model) // heads,                   # reference models/attentions.py:640 (grouped) / :47
            has_conv_res_proj=dim_model != dim_expand))              # reference models/blocks.py:105-109
    return out


# ----------------------------------------------------------------------------------------------------------
# operand rounding emulation (predicts the error of tensor-core operand formats; fp32 accumulate)
# ----------------------------------------------------------------------------------------------------------
def _round_mantissa(x: torch.Tensor, keep_bits: int) -> torch.Tensor:
    """Round-to-nearest (ties away, like cvt.rna) an fp32 tensor to `keep_bits` explicit mantissa bits."""
    if x.dtype != torch.float32:
        return x
    drop = 23 - keep_bits
    xi = x.contiguous().view(torch.int32)
    xi = (xi + (1 << (drop - 1))) & ~((1 << drop) - 1)
    return xi.view(torch.float32)


class Numerics:
    """mode None: exact dtype math.  'tf32': GEMM operands rounded to 10 mantissa bits.  'bf16': 7 bits."""

    def __init__(self, mode=None):
        assert mode in (None, "tf32", "bf16")
        self.mode = mode

    def r(self, x):
        if self.mode == "tf32":
            return _round_mantissa(x, 10)
        if self.mode == "bf16":
            return x.to(torch.bfloat16).to(x.dtype)
        return x

    def mm(self, a, b):
        return torch.matmul(self.r(a), self.r(b))

    def linear(self, x, w, b):
        y = torch.matmul(self.r(x), self.r(w).t())
        return y + b if b is not None else y


EXACT = Numerics(None)


def swish(x):                       # reference models/activations.py:28-29
    return x * torch.sigmoid(x)


def layer_norm(x, w, b, eps=1e-6):  # nn.LayerNorm(dim, eps=1e-6): reference models/modules.py:386,433,511, blocks.py:96
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def batch_norm(x, sd, p, dims, bn=None, eps=1e-5, momentum=0.1):
    """nn.BatchNorm1d / BatchNorm2d over the channel dim of `x` (channel = the dim not in `dims` besides broadcasting).

    Eval mode (bn is None): running statistics (reference models/modules.py:228,517 under model.eval()).
    Train mode (bn = {"updates": {}}): statistics of the whole batch INCLUDING padded frames (SURVEY.md section 8 row a11:
    the reference never masks), biased variance for the normalisation, running statistics updated with momentum 0.1 and
    the unbiased variance (torch.nn.modules.batchnorm semantics); the new running statistics are collected in bn["updates"]."""
    g, b = sd[f"{p}.weight"].to(x.dtype), sd[f"{p}.bias"].to(x.dtype)
    rm, rv = sd[f"{p}.running_mean"].to(x.dtype), sd[f"{p}.running_var"].to(x.dtype)
    shape = [1] * x.dim()
    cdim = [d for d in range(x.dim()) if d not in dims][0]
    shape[cdim] = -1
    if bn is None:
        mean, var = rm, rv
    else:
        mean = x.mean(dim=dims)
        var = ((x - mean.view(shape)) ** 2).mean(dim=dims)
        n = x.numel() // x.shape[cdim]
        bn["updates"][f"{p}.running_mean"] = ((1 - momentum) * rm + momentum * mean).detach()
        bn["updates"][f"{p}.running_var"] = ((1 - momentum) * rv + momentum * var * n / max(n - 1, 1)).detach()
    return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps) * g.view(shape) + b.view(shape)


# ----------------------------------------------------------------------------------------------------------
# front end
# ----------------------------------------------------------------------------------------------------------
def conv2d_subsampling(sd, params, mel, x_len, nm=EXACT, prefix="", bn=None):
    """reference models/modules.py:232-249 (Conv2dSubsampling.forward), eval-mode BatchNorm2d.

    mel (B, n_mels, T) -> (B, C*n_mels/2^L, T') with feature index c*F' + f; x_len -> (x_len-1)//2+1 per layer."""
    x = mel.unsqueeze(1)
    for l in range(params["subsampling_layers"]):
        p = f"{prefix}subsampling_module.layers.{l}"
        ks = params["subsampling_kernel_size"]
        x = F.conv2d(x, sd[f"{p}.0.weight"].to(x.dtype), sd[f"{p}.0.bias"].to(x.dtype), stride=2, padding=(ks - 1) // 2)
        x = batch_norm(x, sd, f"{p}.1", (0, 2, 3), bn)
        x = swish(x)
        if x_len is not None:
            x_len = torch.div(x_len - 1, 2, rounding_mode="floor") + 1
    B, C, Fq, T = x.shape
    return x.reshape(B, C * Fq, T), x_len


# ----------------------------------------------------------------------------------------------------------
# block modules
# ----------------------------------------------------------------------------------------------------------
def feed_forward(sd, p, x, nm=EXACT):
    """reference models/modules.py:378-395: LN -> Linear(D,4D) -> Swish -> Linear(4D,D) (dropout = identity in eval)."""
    c = lambda k: sd[f"{p}.layers.{k}"].to(x.dtype)
    h = layer_norm(x, c("0.weight"), c("0.bias"))
    h = swish(nm.linear(h, c("1.weight"), c("1.bias")))
    return nm.linear(h, c("4.weight"), c("4.bias"))


def relative_sinusoid_rows(t_pad: int, dim: int, group: int, max_len: int, dtype=torch.float32):
    """Rows of the relative sinusoidal table a length-t_pad sequence reads.

    reference models/attentions.py:1209-1257 (G=1) and :1268-1315 (grouped): table row r holds position
    max_len-1-r; the slice for a (padded) sequence of t_pad frames is rows
    [max_len - t_pad + G//2, max_len - G%2 + t_pad - G//2).  The table itself is computed in fp32 exactly as
    the reference does (pos / 10000**(2i/D), sin on even / cos on odd columns)."""
    if group % 2 == 0:
        raise NotImplementedError("even attention group sizes are not used by any shipped config")
    r0 = max_len - t_pad + group // 2
    r1 = max_len - group % 2 + t_pad - group // 2
    assert r0 >= 0, "sequence longer than max_pos_encoding"
    pos = (max_len - 1 - torch.arange(r0, r1, dtype=torch.float32)).unsqueeze(1)
    angles = pos / 10000 ** (2 * torch.arange(0, dim // 2, dtype=torch.float32).unsqueeze(0) / dim)
    R = torch.zeros(r1 - r0, dim, dtype=torch.float32)
    R[:, 0::2] = angles.sin()
    R[:, 1::2] = angles.cos()
    return R.to(dtype)


def relpos_attention(sd, p, x, x_len, spec, nm=EXACT):
    """reference models/attentions.py:549-620 (RelPosMultiHeadSelfAttention.forward, G=1) and :645-718
    (GroupedRelPosMultiHeadSelfAttention.forward), closed form of SURVEY.md section 8 row a9.

    x: LN'ed input (B,T,D).  Returns the output-projected attention (B,T,D) and the weights (B,H,T',T')."""
    B, T, D = x.shape
    H, G = spec.num_heads, spec.group_size
    d = (G * D) // H
    c = lambda k: sd[f"{p}.{k}"].to(x.dtype)
    q = nm.linear(x, c("query_layer.weight"), c("query_layer.bias"))
    k = nm.linear(x, c("key_layer.weight"), c("key_layer.bias"))
    v = nm.linear(x, c("value_layer.weight"), c("value_layer.bias"))
    P = (-T) % G                                    # reference attentions.py:107-121 (pad): zero rows appended AFTER projection
    if P:
        q, k, v = (F.pad(t, (0, 0, 0, P)) for t in (q, k, v))
    Tp = T + P
    Tg = Tp // G
    qu = (q + c("u")).reshape(B, Tg, H, d).transpose(1, 2)      # attentions.py:674-675, 681-682
    qv = (q + c("v")).reshape(B, Tg, H, d).transpose(1, 2)
    kk = k.reshape(B, Tg, H, d).transpose(1, 2)
    vv = v.reshape(B, Tg, H, d).transpose(1, 2)
    R = relative_sinusoid_rows(Tp, D, G, spec.max_pos, x.dtype)
    E = nm.linear(R, c("pos_layer.weight"), c("pos_layer.bias"))   # attentions.py:678 (batch independent)
    E = E.reshape(2 * Tg - 1, H, d).transpose(0, 1)             # (H, 2T'-1, d)
    s_k = nm.mm(qu, kk.transpose(2, 3))                         # (B,H,T',T')
    s_e_rel = nm.mm(qv, E.transpose(1, 2).unsqueeze(0))         # (B,H,T',2T'-1)
    # rel_to_abs (attentions.py:526-547): S_E[i,j] = rel[i, T'-1 + j - i]
    idx = (Tg - 1) + torch.arange(Tg)[None, :] - torch.arange(Tg)[:, None]
    s_e = torch.gather(s_e_rel, 3, idx[None, None].expand(B, H, Tg, Tg))
    s = (s_k + s_e) / d ** 0.5                                  # attentions.py:692 (grouped head dim)
    if x_len is not None:                                       # attentions.py:695-701; mask key group j iff j*G >= x_len
        key_pos = torch.arange(Tg) * G
        masked = key_pos[None, :] >= x_len[:, None]             # (B,T')
        s = s + masked[:, None, None, :].to(s.dtype) * -1e9
    w = s.softmax(dim=-1)
    o = nm.mm(w, vv).transpose(1, 2).reshape(B, Tp, D)[:, :T]   # attentions.py:707-713
    return nm.linear(o, c("output_layer.weight"), c("output_layer.bias")), w


def conv_module(sd, p, x, spec, nm=EXACT, bn=None):
    """reference models/modules.py:507-525 (ConvolutionModule) with Conv1d 'same' pre-padding
    (models/layers.py:99-100,131-136), Glu (activations.py:37-39), eval-mode BatchNorm1d (eps 1e-5)."""
    c = lambda k: sd[f"{p}.layers.{k}"].to(x.dtype)
    De, ks, st = spec.dim_expand, spec.kernel_size, spec.conv_stride
    h = layer_norm(x, c("0.weight"), c("0.bias"))
    h = nm.linear(h, c("2.weight")[:, :, 0], c("2.bias"))       # pointwise conv == per-frame linear
    h = h[..., :De] * torch.sigmoid(h[..., De:])                # GLU over channels: first half * sigmoid(second half)
    B, T, _ = h.shape
    pad = (ks - 1) // 2
    hp = F.pad(h, (0, 0, pad, pad))                             # zero halo in time
    To = (T - 1) // st + 1
    w = c("4.weight")[:, 0, :]                                  # (De, ks)
    out = torch.zeros(B, To, De, dtype=x.dtype)
    for kk in range(ks):                                        # out[t,c] = b_c + sum_k w[c,k] * in[t*s + k - pad, c]
        out = out + hp[:, kk: kk + (To - 1) * st + 1: st, :] * w[:, kk]
    out = out + c("4.bias")
    out = batch_norm(out, sd, f"{p}.layers.5", (0, 1), bn)
    out = swish(out)
    return nm.linear(out, c("7.weight")[:, :, 0], c("7.bias"))


def conformer_block(sd, p, x, x_len, spec, nm=EXACT, taps=None, bn=None):
    """reference models/blocks.py:119-137."""
    x = x + 0.5 * feed_forward(sd, f"{p}.feed_forward_module1", x, nm)
    m = f"{p}.multi_head_self_attention_module"
    a_in = layer_norm(x, sd[f"{m}.norm.weight"].to(x.dtype), sd[f"{m}.norm.bias"].to(x.dtype))
    att, w = relpos_attention(sd, f"{m}.mhsa", a_in, x_len, spec, nm)
    x = x + att                                                 # att_res = Identity (att_stride == 1)
    if taps is not None:
        taps[f"{p}.after_mhsa"] = x
    cm = conv_module(sd, f"{p}.convolution_module", x, spec, nm, bn)
    if spec.has_conv_res_proj:                                  # blocks.py:105-109: strided pointwise conv on the block input
        res = nm.linear(x[:, ::spec.conv_stride], sd[f"{p}.conv_res.1.weight"].to(x.dtype)[:, :, 0], sd[f"{p}.conv_res.1.bias"].to(x.dtype))
    elif spec.conv_stride > 1:                                  # blocks.py:110-113: MaxPool1d(kernel 1, stride s)
        res = x[:, ::spec.conv_stride]
    else:
        res = x
    x = res + cm
    if taps is not None:
        taps[f"{p}.after_conv"] = x
    x = x + 0.5 * feed_forward(sd, f"{p}.feed_forward_module2", x, nm)
    x = layer_norm(x, sd[f"{p}.norm.weight"].to(x.dtype), sd[f"{p}.norm.bias"].to(x.dtype))
    return x, w


# ----------------------------------------------------------------------------------------------------------
# encoder / CTC model
# ----------------------------------------------------------------------------------------------------------
def encoder_forward_mel(sd, params, mel, x_len, nm=EXACT, taps=None, prefix="", bn=None):
    """reference models/encoders.py:106-142 from the mel spectrogram on (no SpecAugment / dropout: Pdrop = 0 semantics).

    mel (B, n_mels, T), x_len (B,) in mel frames or None.  Returns (x (B,T_out,D_last), x_len_out).
    bn = None: eval-mode BatchNorm; bn = {"updates": {}}: train-mode batch statistics (see batch_norm).  Every operation is a
    differentiable torch op, so torch.autograd over this function is the oracle of the backward pass as well."""
    sdp = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    x, x_len = conv2d_subsampling(sdp, params, mel, x_len, nm, bn=bn)
    x = nm.linear(x.transpose(1, 2), sdp["linear.weight"].to(mel.dtype), sdp["linear.bias"].to(mel.dtype))
    if taps is not None:
        taps["linear"] = x
    for i, spec in enumerate(resolve_blocks(params)):
        x, _ = conformer_block(sdp, f"blocks.{i}", x, x_len, spec, nm, taps, bn)
        if taps is not None:
            taps[f"blocks.{i}"] = x
        if spec.conv_stride > 1 and x_len is not None:
            x_len = torch.div(x_len - 1, spec.conv_stride, rounding_mode="floor") + 1
    return x, x_len


def audio_to_mel(sd, params, audio, audio_len, prefix=""):
    """reference models/modules.py:87-106 (AudioPreprocessing.forward).  STFT (torch.stft, hann window from the
    state_dict, centre/reflect padding, power 2) -> mel filterbank matmul -> log(x+1e-9).  The window / filterbank
    buffers are the torchaudio ones stored in the reference state_dict."""
    n_fft = params["n_fft"]
    win = int(params["sample_rate"] * params["win_length_ms"]) // 1000
    hop = int(params["sample_rate"] * params["hop_length_ms"]) // 1000
    spec = torch.stft(audio, n_fft, hop_length=hop, win_length=win, window=sd[f"{prefix}preprocessing.Spectrogram.window"].to(audio.dtype),
                      center=True, pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    power = spec.abs() ** 2
    mel = torch.matmul(power.transpose(-1, -2), sd[f"{prefix}preprocessing.MelScale.fb"].to(audio.dtype)).transpose(-1, -2)
    mel = (mel.float() + 1e-9).log().to(audio.dtype)
    if audio_len is not None:
        audio_len = torch.div(audio_len, hop, rounding_mode="floor") + 1
    if params["normalize"]:
        mel = (mel - params["mean"]) / params["std"]
    return mel, audio_len


def model_ctc_forward_mel(sd, params, mel, x_len, nm=EXACT, taps=None, bn=None):
    """reference models/model_ctc.py:57-68 from the mel spectrogram on; `sd` uses `encoder.` / `fc.` prefixes.
    Keys collected in bn["updates"] are relative to the encoder (no `encoder.` prefix)."""
    x, x_len = encoder_forward_mel(sd, params, mel, x_len, nm, taps, prefix="encoder.", bn=bn)
    logits = nm.linear(x, sd["fc.weight"].to(x.dtype), sd["fc.bias"].to(x.dtype))
    return logits, x_len


def _lse0(stacked):
    """logsumexp over dim 0 that is differentiable when every entry is -inf (torch.logsumexp back-propagates NaN there):
    value -inf, zero gradient."""
    m = stacked.max(dim=0).values
    dead = torch.isinf(m) & (m < 0)
    m_safe = torch.where(dead, torch.zeros_like(m), m).detach()
    tot = torch.exp(stacked - m_safe).sum(dim=0)
    return torch.where(dead, torch.full_like(m, float("-inf")), m_safe + torch.log(tot.clamp_min(1e-300)))


def ctc_loss(logits, logits_len, targets, target_len):
    """reference models/losses.py:56-71: log_softmax -> CTCLoss(blank=0, reduction='none', zero_infinity=False) -> mean.

    The CTC forward (alpha) recursion is restated here in the log domain (Graves et al. 2006, eq. 6-8; the algorithm
    torch's ctc_loss implements): extended label sequence l' = [blank, y1, blank, ..., yU, blank];
    alpha_t(s) = logsumexp(alpha_{t-1}(s), alpha_{t-1}(s-1), [alpha_{t-1}(s-2) if l'_s != blank and l'_s != l'_{s-2}]) + lp_t(l'_s);
    loss_b = -logsumexp(alpha_{T_b-1}(S-1), alpha_{T_b-1}(S-2))."""
    lp = torch.log_softmax(logits.double(), dim=-1)
    B, T, _ = logits.shape
    neg_inf = float("-inf")
    logits_len = logits_len.long().clamp(max=T)
    target_len = target_len.long()
    Umax = max(int(target_len.max()), 0)
    S = 2 * Umax + 1
    # the recursion only moves upward in s, so states above an utterance's own 2U+1 never feed its read-out:
    # all utterances share the widest extended sequence and are advanced together (vectorised over the batch)
    ext = torch.zeros(B, S, dtype=torch.long)
    if Umax:
        ext[:, 1::2] = targets[:, :Umax].long()
    lp_ext = lp.gather(2, ext[:, None, :].expand(B, T, S))
    allow_skip = torch.zeros(B, S, dtype=torch.bool)
    if S > 2:
        allow_skip[:, 2:] = (ext[:, 2:] != 0) & (ext[:, 2:] != ext[:, :-2])
    alpha = torch.full((B, S), neg_inf, dtype=torch.float64)
    alpha[:, 0] = lp_ext[:, 0, 0]
    if S > 1:
        alpha[:, 1] = lp_ext[:, 0, 1]
    pad1 = torch.full((B, 1), neg_inf, dtype=torch.float64)
    pad2 = torch.full((B, 2), neg_inf, dtype=torch.float64)
    for t in range(1, T):
        a1 = torch.cat([pad1, alpha[:, :-1]], dim=1)
        a2 = torch.cat([pad2, alpha[:, :-2]], dim=1) if S > 2 else torch.full_like(alpha, neg_inf)
        a2 = torch.where(allow_skip, a2, torch.full_like(a2, neg_inf))
        new = _lse0(torch.stack([alpha, a1, a2])) + lp_ext[:, t]
        alpha = torch.where((t < logits_len)[:, None], new, alpha)
    last = 2 * target_len                                        # index S_b - 1
    end1 = alpha.gather(1, last[:, None])[:, 0]
    end2 = torch.where(last > 0, alpha.gather(1, (last - 1).clamp(min=0)[:, None])[:, 0], torch.full_like(end1, neg_inf))
    losses = -_lse0(torch.stack([end1, end2]))
    losses = torch.where(logits_len > 0, losses, torch.full_like(losses, float("inf")))
    return losses.mean().to(logits.dtype), losses.to(logits.dtype)


def greedy_ids(logits, logits_len):
    """reference models/model_ctc.py:90-133: argmax over log_softmax (== argmax of logits), then per utterance
    skip blanks (id 0) and emit a token when it differs from the last emitted one or a blank intervened."""
    preds = logits.argmax(dim=-1)
    out = []
    for b in range(preds.shape[0]):
        seq, prev = [], 0
        for t in range(int(logits_len[b])):
            tok = int(preds[b, t])
            if tok != 0 and tok != prev:
                seq.append(tok)
            prev = tok
        out.append(seq)
    return out
