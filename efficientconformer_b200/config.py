"""Host-side resolution of the reference's `encoder_params` dict into per-block shapes.

Mirrors the index rules of the reference constructor (reference models/encoders.py:80-94):
`dim_model`/`num_heads` are indexed by the number of expand blocks strictly before the block,
`dim_expand`/`kernel_size` by the number of expand blocks at or before it, and group size /
max position / strides by the strided blocks.  Pure Python: no torch, no CUDA.
"""
from dataclasses import dataclass, asdict
from typing import List

# The encoder_params of configs/EfficientConformerCTCSmall.json (reference configs/EfficientConformerCTCSmall.json:5-43).
CTC_SMALL_ENCODER_PARAMS = {
    "arch": "Conformer", "num_blocks": 15, "dim_model": [120, 168, 240], "ff_ratio": 4, "num_heads": 4,
    "kernel_size": 15, "Pdrop": 0.1, "conv_stride": 2, "att_stride": 1, "strided_blocks": [4, 9],
    "expand_blocks": [4, 9], "att_group_size": [3, 1, 1], "relative_pos_enc": True, "max_pos_encoding": 10000,
    "subsampling_module": "Conv2d", "subsampling_layers": 1, "subsampling_filters": [120],
    "subsampling_kernel_size": 3, "subsampling_norm": "batch", "subsampling_act": "swish",
    "sample_rate": 16000, "win_length_ms": 25, "hop_length_ms": 10, "n_fft": 512, "n_mels": 80,
    "normalize": False, "mean": -5.6501, "std": 4.2280, "spec_augment": True, "mF": 2, "F": 27, "mT": 5, "pS": 0.05,
}
CTC_SMALL_VOCAB = 256


def _variant(**over):
    p = dict(CTC_SMALL_ENCODER_PARAMS)
    p.update(over)
    return p


def _conformer(num_blocks, dim, heads, **over):
    """The non-progressive Conformer family (reference configs/Conformer*.json): two Conv2d subsampling layers (1/4 frame rate),
    constant width, depthwise k = 31, no strided / expand blocks, ungrouped relative attention."""
    p = _variant(num_blocks=num_blocks, dim_model=dim, num_heads=heads, kernel_size=31, subsampling_layers=2,
                 subsampling_filters=[dim, dim], **over)
    for k in ("conv_stride", "att_stride", "strided_blocks", "expand_blocks", "att_group_size"):
        del p[k]
    return p


# encoder_params of every shipped ASR config (reference configs/*.json), keyed by the config file name; value = (params, vocab)
SHIPPED_ENCODER_PARAMS = {
    "EfficientConformerCTCSmall": (CTC_SMALL_ENCODER_PARAMS, 256),
    "EfficientConformerCTCMedium": (_variant(num_blocks=16, dim_model=[180, 256, 360], strided_blocks=[4, 10], expand_blocks=[4, 10],
                                             subsampling_filters=[180]), 256),
    "EfficientConformerCTCLarge": (_variant(num_blocks=16, dim_model=[360, 512, 720], num_heads=8, strided_blocks=[4, 10],
                                            expand_blocks=[4, 10], subsampling_filters=[360], mT=10), 256),
    "EfficientConformerTransducerSmall": (_variant(dim_model=[100, 140, 200], subsampling_filters=[100], mT=10), 1000),
    "EfficientConformerTransducerMedium": (_variant(dim_model=[180, 256, 360], subsampling_filters=[180], mT=10), 1000),
    "EfficientConformerTransducerLarge": (_variant(dim_model=[360, 512, 720], num_heads=8, subsampling_filters=[360], mT=10), 1000),
    "ConformerCTCSmall": (_conformer(16, 176, 4), 256),
    "ConformerCTCMedium": (_conformer(18, 256, 4), 256),
    "ConformerCTCLarge": (_conformer(18, 512, 8, mT=10), 256),
    "ConformerTransducerSmall": (_conformer(16, 144, 6, mT=10), 1000),
    "ConformerTransducerMedium": (_conformer(16, 256, 4, mT=10), 1000),
    "ConformerTransducerLarge": (_conformer(17, 512, 8, mT=10), 1000),
}


@dataclass
class BlockSpec:
    dim_model: int      # D  : width of FFN1 + attention
    dim_expand: int     # D' : width after the convolution module
    num_heads: int      # H
    kernel_size: int    # depthwise kernel k
    group_size: int     # attention group size G
    max_pos: int        # rows of the relative sinusoid table are 2*max_pos-1
    conv_stride: int    # 1 or 2
    ff_ratio: int

    @property
    def dim_head(self) -> int:          # reference models/attentions.py:640 (grouped) / :47
        return (self.group_size * self.dim_model) // self.num_heads

    @property
    def has_conv_res_proj(self) -> bool:  # reference models/blocks.py:105-109
        return self.dim_model != self.dim_expand


def _pick(value, idx):
    return value[idx] if isinstance(value, list) else value


def resolve_blocks(params: dict) -> List[BlockSpec]:
    expand = params.get("expand_blocks", [])
    strided = params.get("strided_blocks", [])
    specs = []
    for i in range(params["num_blocks"]):
        gt_e = sum(1 for e in expand if i > e)
        ge_e = sum(1 for e in expand if i >= e)
        gt_s = sum(1 for s in strided if i > s)
        if params.get("att_kernel_size", None) is not None or params.get("linear_att", False):
            raise NotImplementedError("local / linear attention variants are outside the hot-path scope (SURVEY.md §2 row 4)")
        if not params["relative_pos_enc"]:
            raise NotImplementedError("only relative_pos_enc=true configs are in scope (all shipped ASR configs)")
        if params.get("causal", False):
            raise NotImplementedError("causal attention is not used by any shipped config")
        att_stride = _pick(params["att_stride"], gt_s) if i in strided else 1
        if att_stride != 1:
            raise NotImplementedError("att_stride > 1 is not used by any shipped config")
        specs.append(BlockSpec(
            dim_model=_pick(params["dim_model"], gt_e),
            dim_expand=_pick(params["dim_model"], ge_e),
            num_heads=_pick(params["num_heads"], gt_e),
            kernel_size=_pick(params["kernel_size"], ge_e),
            group_size=_pick(params.get("att_group_size", 1), gt_s),
            max_pos=params["max_pos_encoding"] // params.get("stride", 2) ** gt_s,
            conv_stride=_pick(params["conv_stride"], gt_s) if i in strided else 1,
            ff_ratio=params["ff_ratio"],
        ))
    return specs


def state_dict_layout(params: dict, vocab_size: int = None):
    """Ordered (name, shape) list of the reference `ConformerEncoder.state_dict()`,
    followed by `fc.*` when vocab_size is given (reference models/model_ctc.py:49)."""
    out = []
    n_fft = params["n_fft"]
    win = int(params["sample_rate"] * params["win_length_ms"]) // 1000
    out.append(("preprocessing.Spectrogram.window", (win,)))
    out.append(("preprocessing.MelScale.fb", (n_fft // 2 + 1, params["n_mels"])))
    if params["subsampling_module"] != "Conv2d":
        raise NotImplementedError("only Conv2d subsampling is in scope")
    filters = params["subsampling_filters"]
    ks = params["subsampling_kernel_size"]
    for l in range(params["subsampling_layers"]):
        cin = 1 if l == 0 else filters[l - 1]
        p = f"subsampling_module.layers.{l}"
        out += [(f"{p}.0.weight", (filters[l], cin, ks, ks)), (f"{p}.0.bias", (filters[l],))]
        if params["subsampling_norm"] == "batch":
            out += [(f"{p}.1.weight", (filters[l],)), (f"{p}.1.bias", (filters[l],)),
                    (f"{p}.1.running_mean", (filters[l],)), (f"{p}.1.running_var", (filters[l],)),
                    (f"{p}.1.num_batches_tracked", ())]
        else:
            raise NotImplementedError("only subsampling_norm=batch is in scope")
    specs = resolve_blocks(params)
    d0 = specs[0].dim_model
    out += [("linear.weight", (d0, filters[-1] * params["n_mels"] // 2 ** params["subsampling_layers"])),
            ("linear.bias", (d0,))]
    for i, s in enumerate(specs):
        b = f"blocks.{i}"
        D, De, F = s.dim_model, s.dim_expand, s.ff_ratio
        def ffn(tag, d):
            p = f"{b}.{tag}.layers"
            return [(f"{p}.0.weight", (d,)), (f"{p}.0.bias", (d,)),
                    (f"{p}.1.weight", (F * d, d)), (f"{p}.1.bias", (F * d,)),
                    (f"{p}.4.weight", (d, F * d)), (f"{p}.4.bias", (d,))]
        out += ffn("feed_forward_module1", D)
        m = f"{b}.multi_head_self_attention_module"
        out += [(f"{m}.norm.weight", (D,)), (f"{m}.norm.bias", (D,)),
                (f"{m}.mhsa.u", (D,)), (f"{m}.mhsa.v", (D,))]
        for nm in ("query", "key", "value", "output", "pos"):
            out += [(f"{m}.mhsa.{nm}_layer.weight", (D, D)), (f"{m}.mhsa.{nm}_layer.bias", (D,))]
        c = f"{b}.convolution_module.layers"
        out += [(f"{c}.0.weight", (D,)), (f"{c}.0.bias", (D,)),
                (f"{c}.2.weight", (2 * De, D, 1)), (f"{c}.2.bias", (2 * De,)),
                (f"{c}.4.weight", (De, 1, s.kernel_size)), (f"{c}.4.bias", (De,)),
                (f"{c}.5.weight", (De,)), (f"{c}.5.bias", (De,)),
                (f"{c}.5.running_mean", (De,)), (f"{c}.5.running_var", (De,)),
                (f"{c}.5.num_batches_tracked", ()),
                (f"{c}.7.weight", (De, De, 1)), (f"{c}.7.bias", (De,))]
        out += ffn("feed_forward_module2", De)
        out += [(f"{b}.norm.weight", (De,)), (f"{b}.norm.bias", (De,))]
        if s.has_conv_res_proj:
            out += [(f"{b}.conv_res.1.weight", (De, D, 1)), (f"{b}.conv_res.1.bias", (De,))]
    if vocab_size is not None:
        out += [("fc.weight", (vocab_size, specs[-1].dim_expand)), ("fc.bias", (vocab_size,))]
    return out


def stage_lengths(params: dict, t_mel: int):
    """Frame count entering each block and leaving the encoder (reference models/modules.py:243, encoders.py:140)."""
    t = t_mel
    for _ in range(params["subsampling_layers"]):
        t = (t - 1) // 2 + 1
    lens = []
    for s in resolve_blocks(params):
        lens.append(t)
        if s.conv_stride > 1:
            t = (t - 1) // s.conv_stride + 1
    return lens, t


def describe(params: dict):
    return [asdict(s) for s in resolve_blocks(params)]
