"""Deterministic synthetic weights and inputs (no datasets / checkpoints exist offline).

Every tensor is drawn from its own torch CPU generator seeded by crc32(name) ^ seed, so the same values are
reproduced on any machine with the same torch build, independent of key order.  Non-trivial LayerNorm /
BatchNorm statistics are used on purpose (default init would hide scale/shift bugs)."""
import math
import zlib
import torch

from .config import state_dict_layout


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def seeded_state_dict(params: dict, vocab_size=None, seed: int = 0, prefix_encoder: str = ""):
    """State dict with the reference's names/shapes (config.state_dict_layout).  When vocab_size is given the
    encoder keys get `prefix_encoder` (use "encoder." for a ModelCTC-style dict) and fc.* is appended."""
    sd = {}
    for name, shape in state_dict_layout(params, vocab_size):
        g = _gen(name, seed)
        leaf = name.split(".")[-1]
        if name.startswith("preprocessing."):
            continue  # torchaudio buffers: filled by the module that owns them
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif leaf in ("u", "v"):
            t = 0.2 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and leaf == "weight":          # LayerNorm / BatchNorm gain
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        else:                                               # Linear / Conv weights: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        key = name if (vocab_size is None or name.startswith("fc.")) else prefix_encoder + name
        sd[key] = t
    return sd


def synthetic_mel(batch: int, t_mel: int, n_mels: int = 80, seed: int = 1):
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    return torch.randn(batch, n_mels, t_mel, generator=g)


def synthetic_audio(batch: int, t_mel: int, hop: int = 160, seed: int = 1):
    """N(0,1) audio of length (T-1)*hop gives exactly T mel frames (reference models/modules.py:80,100)."""
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    return torch.randn(batch, (t_mel - 1) * hop, generator=g)


def ragged_lengths(batch: int, t_max: int, seed: int = 3, min_frac: float = 0.4):
    """Descending lengths, first == t_max (collate_fn_pad pads to the batch max; reference utils/preprocessing.py:33-38)."""
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    lens = (torch.rand(batch, generator=g) * (1 - min_frac) + min_frac) * t_max
    lens = lens.long().clamp(1, t_max)
    lens[0] = t_max
    return torch.sort(lens, descending=True).values


def synthetic_targets(out_len: torch.Tensor, vocab: int, seed: int = 4, frac: float = 1.0 / 3.0):
    """Labels U{1..V-1} with y_len = floor(out_len*frac) >= 1 so that CTC is feasible even with repeats
    (needs 2U+1 <= T in the worst case of all-equal labels)."""
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    y_len = torch.clamp((out_len.float() * frac).long(), min=1)
    u_max = int(y_len.max())
    y = torch.randint(1, vocab, (out_len.numel(), u_max), generator=g)
    for b in range(out_len.numel()):
        y[b, int(y_len[b]):] = 0
    return y, y_len
