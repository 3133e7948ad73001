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


class P2PSyncBatchNormReducer(SyncBatchNormReducer):
    """The same two exchange steps as ONE kernel each over NVLink / NVSwitch peer memory (csrc/p2p_exchange.cu) instead of
    {pack, NCCL collective, unpack, merge}: the statistics are stored straight into every rank's mailbox (symmetric memory mapped by
    torch.distributed._symmetric_memory), flags are raised with release / acquire at system scope, and every rank merges the W payloads
    of its own mailbox in rank order (bit-identical results on all ranks).  Construction raises when symmetric memory cannot be set up
    on this machine (no P2P mapping, unsupported driver): callers fall back to the NCCL reducer (`make_sync_bn_reducer`)."""

    def __init__(self, group=None, device="cuda", uniform=True):
        super().__init__(group, device, uniform)
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        L = _lib.lib()
        self._lib, self._C = _lib, C
        n_bytes = L.ec_p2p_mailbox_bytes(self.world)
        self.mailbox = symm.empty(n_bytes // 4, dtype=torch.float32, device=device)
        self.mailbox.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm.rendezvous(self.mailbox, group if group is not None else dist.group.WORLD)
        if self.handle.world_size != self.world:
            raise RuntimeError("symmetric-memory rendezvous returned a different world size")
        self.rank = self.handle.rank
        self.peer_ptrs = (C.c_ulonglong * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self.max_floats = L.ec_p2p_max_payload_floats()
        self.handle.barrier()                          # every mailbox is zeroed before anybody stores into it
        torch.cuda.synchronize(device)
        self._count_out = torch.zeros(1, dtype=torch.float32, device=device)

    def _exchange(self, data, n, mode, count):
        if n > self.max_floats:
            raise RuntimeError(f"BatchNorm payload of {n} floats exceeds the peer mailbox slot ({self.max_floats})")
        _l = self._lib
        _l.check(_l.lib().ec_p2p_bn_exchange(self.peer_ptrs, self.rank, self.world, _l.ptr(data), n, mode, float(count),
                                             _l.ptr(self._count_out), _l.stream_ptr()))

    def forward_stats(self, stats, count):
        assert stats.is_contiguous() and stats.dtype == torch.float32
        self._exchange(stats, stats.numel(), 1, count)
        if self.uniform:
            return float(count) * self.world
        return float(self._count_out.item())

    def backward_sums(self, sums):
        assert sums.is_contiguous() and sums.dtype == torch.float32
        self._exchange(sums, sums.numel(), 0, 0.0)
        return sums

    def error(self):
        """1 when a peer failed to arrive within the kernel's time-out at any exchange so far (synchronises)."""
        out = self._C.c_int(0)
        self._lib.check(self._lib.lib().ec_p2p_error(self.peer_ptrs, self.rank, self.world, self._C.byref(out)))
        return out.value


def make_sync_bn_reducer(group, device, uniform=True):
    """Peer-memory exchange when this machine supports it (and EFFCONF_P2P_BN != 0), else the NCCL reducer.  All ranks must take
    the same branch: the decision is agreed with one all_reduce(MIN)."""
    import os
    want = os.environ.get("EFFCONF_P2P_BN", "1") != "0" and torch.device(device).type == "cuda"
    red, ok = None, 0
    if want:
        try:
            red = P2PSyncBatchNormReducer(group, device, uniform)
            ok = 1
        except Exception as ex:                          # noqa: BLE001 -- any failure to map peer memory means "use NCCL"
            import warnings
            warnings.warn(f"peer-memory SyncBatchNorm exchange unavailable ({type(ex).__name__}: {ex}); using NCCL collectives")
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 1:
        return red
    return SyncBatchNormReducer(group, device, uniform)
