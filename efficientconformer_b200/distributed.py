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


# ---- training: SyncBatchNorm statistics exchange (reference models/model_ctc.py:70-75: convert_sync_batchnorm + DDP) -----------
from . import ops as _ops_module

_ops = _ops_module       # test seam (tests/test_distributed_cpu.py runs the exchange logic over gloo with a torch-CPU merge)


class SyncBatchNormReducer:
    """The two exchange steps of a synchronised BatchNorm layer, called between the stages of the train-mode conv kernels
    (ops.DwConvTrain / ops.SubsampleTrain):

      forward_stats(stats [2, C], count)  per-rank (mean, centred sum of squares M2) over `count` local frames are all-gathered and
                                          Chan-merged in place (ec_op_stats_merge_ranks); returns the global frame count
      backward_sums(sums [2, C])          (sum dz, sum dz * xhat) all-reduced (SUM) in place

    With `uniform=True` every rank is known to hold the same number of frames (fixed-shape batches: the benchmark, bucketed
    training), so no host synchronisation is needed; otherwise the per-rank counts travel with the statistics and the global
    count is read back once per layer (a device->host sync, as the reference's SyncBatchNorm also gathers counts)."""

    def __init__(self, group=None, device="cpu", uniform=True):
        self.group = group
        self.world = dist.get_world_size(group)
        self.device = device
        self.uniform = uniform

    def forward_stats(self, stats, count):
        C = stats.shape[1]
        payload = torch.empty(2 * C + 1, dtype=torch.float32, device=stats.device)
        payload[:2 * C].copy_(stats.reshape(-1))
        payload[2 * C:].fill_(float(count))
        flat = torch.empty(self.world * (2 * C + 1), dtype=torch.float32, device=stats.device)
        dist.all_gather_into_tensor(flat, payload, group=self.group)
        gathered = flat.view(self.world, 2 * C + 1)
        counts = gathered[:, 2 * C].contiguous()
        _ops.stats_merge_ranks(gathered[:, :2 * C].reshape(self.world, 2, C).contiguous(), counts, stats)
        if self.uniform:
            return float(count) * self.world
        return float(counts.sum().item())

    def backward_sums(self, sums):
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        return sums
