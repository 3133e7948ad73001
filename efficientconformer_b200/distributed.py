"""Data-parallel plumbing for the utterance-sharded forward path.

The reference's only strategy is one process per GPU over utterances (reference main.py:33-35, functions.py:168
DistributedSampler).  The eval forward has no exchange step, so ranks only need (a) a disjoint, balanced shard of the
utterances, (b) a barrier and (c) max / sum reductions of scalars for timing and metric aggregation.  Backend agnostic:
NCCL on the GPU box, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous balanced shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank: int, world: int):
    """Slice every tensor of a collated batch along dim 0 for this rank."""
    n = tensors[0].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return [t[lo:hi] for t in tensors]


def max_over_ranks(value: float, device="cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def sum_over_ranks(value: float, device="cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t)


def gather_ragged_ids(ids, device="cpu"):
    """All ranks' greedy id lists in rank order (the reference uses all_gather_object for transcripts, models/model.py:465-466)."""
    if not (dist.is_available() and dist.is_initialized()):
        return ids
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, ids)
    return [x for part in out for x in part]
