"""The reference trainer's inner step for the CTC model, B200-native (reference models/model.py:239-259 and :88-93,120-126):

    pred = model.forward(batch); loss = criterion(batch, pred); loss.backward()
    optimizer.step(); optimizer.zero_grad(); scheduler.step()

as ONE stream-ordered launch sequence: train-mode forward (training.TrainingPath) -> CTC loss + gradient (ec_ctc_loss_grad) ->
hand-scheduled backward -> gradients gathered into one flat fp32 bucket (ec_op_pack_flat) -> data-parallel all-reduce of that
single bucket over NCCL (reference: DistributedDataParallel, models/model_ctc.py:70-75) -> Adam with the Transformer schedule over
the flat parameter arena (ec_adam_step).  Parameters of the holder modules are re-pointed to views of the flat arena, so
`state_dict()` / `load_state_dict()` keep working and the next forward reads the updated weights without copies.  With a fixed
batch shape the whole step (collectives excluded) is captured once into a CUDA graph and replayed.

Everything outside this step -- epochs, data loading, checkpoints, evaluation, WER -- stays the reference's own Python."""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib
from . import ops as _ops_module
from .model_ctc import ctc_loss_and_grad
from .training import TrainingPath

_ops = _ops_module
_ALIGN = 64     # floats: every tensor starts on a 256-byte boundary of the arena (TMA / vector-access friendly)


class FlatParams:
    """One contiguous fp32 arena holding every parameter of `named` (list of (name, Parameter)); each Parameter's storage is
    re-pointed to its slice.  `grads`, `exp_avg`, `exp_avg_sq` are arenas of the same layout."""

    def __init__(self, named, device):
        self.names, self.offsets, self.sizes, self.shapes = [], [], [], []
        off = 0
        self.tensors = [p for _, p in named]
        for n, p in named:
            self.names.append(n); self.offsets.append(off); self.sizes.append(p.numel()); self.shapes.append(tuple(p.shape))
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off
        self.params = torch.zeros(off, dtype=torch.float32, device=device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=device)
        with torch.no_grad():
            for (n, p), o, s in zip(named, self.offsets, self.sizes):
                view = self.params[o:o + s].view(p.shape)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
        self.index = {n: i for i, n in enumerate(self.names)}
        self.offsets_dev = torch.tensor(self.offsets, dtype=torch.int64, device=device)
        self.sizes_dev = torch.tensor(self.sizes, dtype=torch.int64, device=device)
        # gradient-pointer tables: a ring of (pinned host, device, event) slots, so that an eager step that runs ahead of the GPU never
        # rewrites a table whose host-to-device copy has not executed yet
        pin = torch.cuda.is_available()
        self._slots = []
        for _ in range(12):
            host = torch.empty(len(self.names), dtype=torch.int64)
            self._slots.append([host.pin_memory() if pin else host, torch.empty(len(self.names), dtype=torch.int64, device=device), None])
        self._slot = 0

    def grad_view(self, name):
        i = self.index[name]
        return self.grads[self.offsets[i]:self.offsets[i] + self.sizes[i]].view(self.shapes[i])

    def new_tables(self):
        """A private (pinned host, device) pointer-table pair for ONE captured graph: its memcpy node re-reads the pinned table on every
        replay, so the table must never be rewritten while the graph lives.  Allocate outside the capture."""
        host = torch.empty(len(self.names), dtype=torch.int64)
        return (host.pin_memory() if torch.cuda.is_available() else host, torch.empty(len(self.names), dtype=torch.int64, device=self.grads.device))

    def arena_range(self, lo, hi):
        """Float offsets [a, b) of the arena covered by the parameters lo .. hi-1 (arena order)."""
        return self.offsets[lo], (self.offsets[hi] if hi < len(self.names) else self.total)

    def pack(self, grads, accumulate=False, tables=None, lo=0, hi=None):
        """Gather the per-parameter gradient tensors (dict name -> contiguous fp32 tensor) into `self.grads` (accumulate: add to it);
        lo / hi restrict the call to the parameters lo .. hi-1 of the arena order (one gradient bucket)."""
        keep = []
        hi = len(self.names) if hi is None else hi
        if hi <= lo:
            return keep
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        if capturing and tables is None:
            raise RuntimeError("FlatParams.pack under CUDA-graph capture needs private pointer tables (new_tables())")
        if tables is not None:
            (host, dev), ev = tables, None
        else:
            host, dev, ev = self._slots[self._slot]
            if ev is not None:
                ev.synchronize()                  # the copy that last read this host table has executed
        for i in range(lo, hi):
            n = self.names[i]
            g = grads[n]
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous(); keep.append(g)
            assert g.numel() == self.sizes[i], n
            host[i] = g.data_ptr()
        dev[lo:hi].copy_(host[lo:hi], non_blocking=True)
        if torch.cuda.is_available() and tables is None:
            ev = torch.cuda.Event(); ev.record()
            self._slots[self._slot][2] = ev
            self._slot = (self._slot + 1) % len(self._slots)
        at = lambda t: C.c_void_p(t.data_ptr() + 8 * lo)
        _lib.check(_lib.lib().ec_op_pack_flat_acc(at(dev), at(self.offsets_dev), at(self.sizes_dev), hi - lo,
                                                  _lib.ptr(self.grads), 1 if accumulate else 0, _lib.stream_ptr()))
        return keep


def _qkv_adjacent_order(named):
    """Reorder (name, Parameter) so that the query / key / value weights of every attention module are consecutive in the arena:
    Wq | Wk | Wv is then ONE [3D, D] matrix (no per-step concatenation) when D*D is a multiple of the arena alignment."""
    by_name = dict(named)
    out, used = [], set()
    for n, p in named:
        if n in used:
            continue
        out.append((n, p)); used.add(n)
        if n.endswith("mhsa.query_layer.weight"):
            for other in ("key", "value"):
                k = n.replace("query_layer", f"{other}_layer")
                if k in by_name and k not in used:
                    out.append((k, by_name[k])); used.add(k)
    return out


class ArenaWeights:
    """GEMM operands of every weight, produced from the flat fp32 parameter arena by TWO launches per step: one cast of the whole
    arena (forward operands, [N, K]) and one multi-tensor transposed cast (data-gradient operands, [K, N]).  Views are looked up
    by Parameter identity."""

    def __init__(self, flat, precision, device):
        self.flat = flat
        self.pr = _lib.PRECISIONS[precision] if isinstance(precision, str) else precision
        dt = _lib.act_dtype(self.pr)
        # split mode: every weight operand is [2, N, K] (second plane = swapped halves), so operand offsets are twice the fp32 ones
        pl = self.planes = _lib.weight_planes(self.pr)
        self.arena = torch.zeros(pl * flat.total, dtype=dt, device=device)
        self.arena_t = torch.zeros(pl * flat.total, dtype=dt, device=device)
        self._fwd, self._bwd, self._qkv, desc, desc_f = {}, {}, {}, [], []
        params = dict(zip(flat.names, flat.tensors))
        merged = set()
        for n in flat.names:
            if not n.endswith("mhsa.query_layer.weight"):
                continue
            names = [n, n.replace("query_layer", "key_layer"), n.replace("query_layer", "value_layer")]
            i = [flat.index[x] for x in names]
            D = flat.shapes[i[0]][0]
            if all(flat.offsets[i[j]] == flat.offsets[i[0]] + j * D * D for j in range(3)):
                o = flat.offsets[i[0]]
                self._qkv[id(params[n])] = (self.arena[pl * o:pl * o + 3 * D * D].view(3 * D, D),
                                            self.arena_t[pl * o:pl * o + 3 * D * D].view(D, 3 * D),
                                            torch.zeros(3 * D, dtype=torch.float32, device=device))
                desc.append((o, 3 * D, D, pl * o))
                desc_f.append((o, 3 * D, D, pl * o))
                merged.update(names)
        for n, o, shape in zip(flat.names, flat.offsets, flat.shapes):
            is_matrix = len(shape) == 2 or (len(shape) == 3 and shape[-1] == 1)
            if not is_matrix:
                continue
            N, K = shape[0], shape[1]
            if n not in merged:
                self._fwd[id(params[n])] = self.arena[pl * o:pl * o + N * K].view(N, K)
                self._bwd[id(params[n])] = self.arena_t[pl * o:pl * o + N * K].view(K, N)
                desc.append((o, N, K, pl * o))
                desc_f.append((o, N, K, pl * o))
            elif pl == 1:
                self._fwd[id(params[n])] = self.arena[o:o + N * K].view(N, K)
        self.n_desc = len(desc)
        self.desc = torch.tensor(desc, dtype=torch.int64, device=device).contiguous()
        self.desc_f = torch.tensor(desc_f, dtype=torch.int64, device=device).contiguous()

    def refresh(self, side=None):
        """Run once per step, before the forward: the parameters changed in the previous optimiser step.  The transposed operands are
        only read by the backward: with `side` (a stream forked from the current one; the backward joins it before its first data
        gradient) they are produced beside the forward."""
        if self.planes == 1:
            _ops.cast_into(self.flat.params, self.arena, self.pr)
        else:
            _ops.cast_multi(self.flat.params, self.desc_f, self.n_desc, self.arena, self.pr)
        if side is None:
            _ops.transpose_cast_multi(self.flat.params, self.desc, self.n_desc, self.arena_t, self.pr)
            return
        side.wait_stream(torch.cuda.current_stream(self.arena.device))
        with torch.cuda.stream(side):
            _ops.transpose_cast_multi(self.flat.params, self.desc, self.n_desc, self.arena_t, self.pr)

    def act(self, weight):
        return self._fwd[id(weight)]

    def act_t(self, weight):
        return self._bwd[id(weight)]

    def _qkv_entry(self, mhsa):
        ent = self._qkv.get(id(mhsa.query_layer.weight))
        if ent is None:
            raise RuntimeError("Wq | Wk | Wv are not adjacent in the parameter arena (model dim not a multiple of 8?)")
        return ent

    def qkv_act(self, mhsa):
        return self._qkv_entry(mhsa)[0]

    def qkv_act_t(self, mhsa):
        return self._qkv_entry(mhsa)[1]

    def qkv_bias(self, mhsa):
        b = self._qkv_entry(mhsa)[2]
        D = b.numel() // 3
        for j, layer in enumerate((mhsa.query_layer, mhsa.key_layer, mhsa.value_layer)):
            b[j * D:(j + 1) * D].copy_(layer.bias.detach())       # device-to-device memcpy nodes
        return b

    def supports(self, encoder):
        return all(id(b.multi_head_self_attention_module.mhsa.query_layer.weight) in self._qkv for b in encoder.blocks)


class CTCTrainStep:
    """The reference trainer's optimisation step for ModelCTC (reference models/model.py:239-259).  `training_params` is the reference
    config's dict: optimizer Adam (beta1, beta2, eps, weight_decay), lr_schedule Transformer (schedule_dim, warmup_steps, K) or
    Constant (lr_value), accumulated_steps (micro-batches per optimiser step, default 1).  One CUDA graph per (batch shape, role of the
    micro-step) is captured on first use and replayed afterwards."""

    def __init__(self, model, training_params, precision="bf16x2", process_group=None, sync_bn=True, use_cuda_graph=True, dropout_seed=0,
                 data_parallel=True, max_graphs=8):
        if training_params.get("optimizer", "Adam") != "Adam":
            raise NotImplementedError("the shipped configs train with Adam (reference models/model.py:88-93)")
        sched = training_params.get("lr_schedule", "Transformer")
        if sched not in ("Transformer", "Constant"):
            raise NotImplementedError("lr schedules: Transformer (all shipped ASR configs) or Constant")
        self.model = model
        self.tp = training_params
        self.precision = precision
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (data_parallel and dist.is_available() and dist.is_initialized()) else 1
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("CTCTrainStep runs on CUDA sm_100 only")
        self.accum = int(training_params.get("accumulated_steps", 1))
        if self.accum < 1:
            raise ValueError("accumulated_steps must be >= 1")
        reducer = None
        if self.world > 1 and sync_bn:
            from .distributed import make_sync_bn_reducer
            reducer = make_sync_bn_reducer(process_group, self.device)      # peer-memory exchange kernel, or NCCL collectives
        self.reducer = reducer
        self.path = TrainingPath(model.encoder, model.fc, stats_reducer=reducer, dropout_seed=dropout_seed)
        self.flat = FlatParams(_qkv_adjacent_order(self.path.param_list()), self.device)
        model.encoder._drop_engines()                       # inference arenas were prepared from the old parameter storage
        weights = ArenaWeights(self.flat, precision, self.device)
        self.weights = weights if weights.supports(model.encoder) else None
        self.path.weights = self.weights
        # device state {lr bits, adam step t, schedule step s, 0}.  The reference's compile() ends with scheduler.step() ("Init LR",
        # models/model.py:150): the n-th optimiser step runs with lr(s = n) and scheduler.model_step == n afterwards.
        self.state = torch.zeros(4, dtype=torch.int32, device=self.device)
        self.schedule = 1 if sched == "Transformer" else 0
        self._set_schedule_step(0)
        self.use_cuda_graph = use_cuda_graph
        # EFFCONF_BUCKET_OVERLAP=1: three gradient buckets, the first two all-reduced on a communication stream while the backward of
        # the earlier blocks runs.  Off by default: measured neutral at 2 GPUs (13.18 vs 13.23 ms per step; the NCCL kernel takes SMs
        # from the single-wave backward GEMMs while it overlaps them) and not measured at 8.
        self._buckets = self._plan_buckets(model) if self.world > 1 and os.environ.get("EFFCONF_BUCKET_OVERLAP", "0") == "1" else {}
        self._comm, self._pending = None, None
        self.max_graphs = max_graphs
        self._graphs = {}                                   # (shape key, accumulate, final) -> [graph(s), static inputs, keep-alive]
        self._micro = 0                                     # micro-batches accumulated since the last optimiser step
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.launches_per_step = None

    def _lr_of(self, s):
        tp = self.tp
        if not self.schedule:
            return float(tp["lr_value"])
        return float(tp["K"]) * float(tp["schedule_dim"]) ** -0.5 * min(s ** -0.5, s * float(tp["warmup_steps"]) ** -1.5)

    def _set_schedule_step(self, model_step, adam_t=None):
        """State after `model_step` optimiser steps: lr = lr(s = model_step + 1), as the reference's scheduler holds it."""
        st = torch.zeros(4, dtype=torch.int32)
        st[0:1].view(torch.float32).fill_(self._lr_of(model_step + 1))
        st[1] = model_step if adam_t is None else adam_t
        st[2] = model_step + 1
        self.state.copy_(st)

    # ---- pieces -----------------------------------------------------------------------------------------------------------
    def _plan_buckets(self, model):
        """Gradient buckets for the overlapped data-parallel all-reduce: the arena is cut where the last third and the middle third of
        the blocks begin, so that bucket 0 = {blocks >= b1, head} is complete -- and its all-reduce can start on the communication
        stream -- while the backward of the earlier blocks is still running; the last bucket (first blocks + front end) goes out
        after the backward.  Returns {block index: (lo, hi) parameter range} (empty: one bucket after the backward)."""
        n = len(model.encoder.blocks)
        names = self.flat.names
        plan, hi = {}, len(names)
        for b in sorted({(2 * n) // 3, n // 3}, reverse=True):
            if b <= 0 or b >= n:
                continue
            first = [i for i, nm in enumerate(names) if nm.startswith(f"encoder.blocks.{b}.")]
            if not first:
                return {}
            lo = min(first)
            for i, nm in enumerate(names):                   # contiguity: everything from `lo` on belongs to blocks >= b or the head
                later = nm.startswith("fc.") or (nm.startswith("encoder.blocks.") and int(nm.split(".")[2]) >= b)
                if (i >= lo) != later:
                    return {}
            plan[b] = (lo, hi)
            hi = lo
        return plan

    def _forward_backward(self, mel, mel_len, targets, target_len, accumulate, tables=None, final=False):
        if self.weights is not None:
            self.weights.refresh(side=self.path.side_stream(mel.device))
        x, logits, out_len, tape = self.path.forward(mel, mel_len, self.precision, want_logits=True)
        if out_len is None:
            out_len = torch.full((mel.shape[0],), logits.shape[1], dtype=torch.int64, device=mel.device)
        mean, per, dlogits = ctc_loss_and_grad(logits, out_len, targets, target_len)
        hook, keep, works, rest_hi = None, [], [], len(self.flat.names)
        capturing = mel.is_cuda and torch.cuda.is_current_stream_capturing()
        if final and self.world > 1 and self._buckets and mel.is_cuda and (self.reducer is not None or not capturing):
            # buckets leave on the communication stream as soon as their blocks' gradients are enqueued (main + weight-gradient streams)
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=mel.device)
            comm, rest_hi = self._comm, min(lo for lo, _ in self._buckets.values())

            def hook(i, grads):
                rng = self._buckets.get(i)
                if rng is None:
                    return
                comm.wait_stream(torch.cuda.current_stream(mel.device))
                side = self.path.side_stream(mel.device)
                if side is not None:
                    comm.wait_stream(side)
                with torch.cuda.stream(comm):
                    keep.extend(self.flat.pack(grads, accumulate=accumulate, tables=tables, lo=rng[0], hi=rng[1]))
                    a, b = self.flat.arena_range(*rng)
                    works.append(dist.all_reduce(self.flat.grads[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        grads = self.path.backward(tape, None, dlogits, stage_hook=hook)
        keep.extend(self.flat.pack(grads, accumulate=accumulate, tables=tables, lo=0, hi=rest_hi))
        if hook is not None:
            keep.append(grads)                               # read by the communication stream: alive until the step has joined it
            self._pending = (works, self.flat.arena_range(0, rest_hi))
        self.loss.copy_(mean)
        return keep

    def _optimizer(self):
        tp = self.tp
        # loss / accumulated_steps (reference models/model.py:245) and the data-parallel mean, folded into one gradient scale
        _ops.adam_step(self.flat.params, self.flat.grads, self.flat.exp_avg, self.flat.exp_avg_sq, self.state, float(tp["beta1"]),
                       float(tp["beta2"]), float(tp["eps"]), float(tp["weight_decay"]), grad_scale=1.0 / (self.world * self.accum),
                       schedule=self.schedule, K=float(tp.get("K", 0.0)), dim=float(tp.get("schedule_dim", 1.0)),
                       warmup=float(tp.get("warmup_steps", 1.0)))

    def _all_reduce(self):
        if self.world <= 1:
            return
        if self._pending is not None:                        # overlapped buckets are in flight: the last one, then join
            (works, (a, b)), self._pending = self._pending, None
            dist.all_reduce(self.flat.grads[a:b], op=dist.ReduceOp.SUM, group=self.group)
            for w in works:
                w.wait()
            torch.cuda.current_stream(self.device).wait_stream(self._comm)
            return
        dist.all_reduce(self.flat.grads, op=dist.ReduceOp.SUM, group=self.group)         # ONE bucket; the mean is folded into Adam

    def _step_eager(self, mel, mel_len, targets, target_len, accumulate, final):
        with torch.no_grad():
            keep = self._forward_backward(mel, mel_len, targets, target_len, accumulate, final=final)
            if final:
                self._all_reduce()
                self._optimizer()
        del keep

    # ---- public -----------------------------------------------------------------------------------------------------------
    def step(self, mel, mel_len, targets, target_len):
        """mel (B, n_mels, T) fp32, mel_len (B,) int64 or None, targets (B, U) int64, target_len (B,) int64 -- CUDA tensors.
        Returns the (device) mean CTC loss of this (micro-)batch.  With accumulated_steps == A every A-th call is an optimiser step:
        parameters, Adam moments and the learning-rate schedule advance; BatchNorm running statistics advance on every call."""
        for t in (mel, targets, target_len):
            if not t.is_cuda:
                raise RuntimeError("CTCTrainStep takes CUDA tensors (copy the batch with non_blocking H2D first)")
        accumulate, final = self._micro > 0, self._micro + 1 == self.accum
        self._micro = 0 if final else self._micro + 1
        if final:
            self.model.encoder.mark_weights_changed()       # the inference engines re-prepare their weights on the next eval forward
        if not self.use_cuda_graph:
            self._step_eager(mel, mel_len, targets, target_len, accumulate, final)
            return self.loss
        key = (tuple(mel.shape), mel_len is not None, tuple(targets.shape), accumulate, final)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= self.max_graphs:        # bounded cache: drop the least recently used graph
                self._graphs.pop(next(iter(self._graphs)))
            ent = self._capture(mel, mel_len, targets, target_len, accumulate, final)
        else:
            self._graphs.pop(key)                           # re-insert: most recently used last
        self._graphs[key] = ent
        (g1, g2), (s_mel, s_len, s_y, s_yl), _ = ent
        s_mel.copy_(mel, non_blocking=True)
        if s_len is not None:
            s_len.copy_(mel_len, non_blocking=True)
        s_y.copy_(targets, non_blocking=True); s_yl.copy_(target_len, non_blocking=True)
        g1.replay()
        if g2 is not None:                                  # world > 1 with collectives kept outside the graph
            self._all_reduce()
            g2.replay()
        return self.loss

    def _capture(self, mel, mel_len, targets, target_len, accumulate, final):
        s_mel, s_y, s_yl = mel.clone(), targets.clone(), target_len.clone()
        s_len = mel_len.clone() if mel_len is not None else None
        # snapshot everything a step mutates: the warm-up step below must not count as a training step
        mutated = (self.flat.params, self.flat.grads, self.flat.exp_avg, self.flat.exp_avg_sq, self.state)
        snap = [t.clone() for t in mutated]
        bufs = [b for b in self.model.buffers()]
        snap_b = [b.clone() for b in bufs]
        drop = self.path._dropout_state(self.device)
        snap_c = drop.counter.clone() if drop.counter is not None else None
        self._step_eager(s_mel, s_len, s_y, s_yl, accumulate, final)   # warm-up: kernel attributes, allocator pools, NCCL communicators
        torch.cuda.synchronize()
        tables = self.flat.new_tables()
        split = final and self.world > 1 and self.reducer is None      # no SyncBN collectives inside: keep NCCL outside the graphs
        g1, g2 = torch.cuda.CUDAGraph(), None
        # other threads (the NCCL watchdog) may touch the CUDA API while this thread captures
        mode = {"capture_error_mode": "thread_local"} if self.world > 1 else {}
        with torch.no_grad():
            with torch.cuda.graph(g1, **mode):
                keep = self._forward_backward(s_mel, s_len, s_y, s_yl, accumulate, tables=tables, final=final)
                if final and not split:
                    self._all_reduce()
                    self._optimizer()
            if split:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, **mode):
                    self._optimizer()
        with torch.no_grad():
            for t, sn in zip(mutated, snap):
                t.copy_(sn)
            for b, sn in zip(bufs, snap_b):
                b.copy_(sn)
            if snap_c is not None:
                drop.counter.copy_(snap_c)
        torch.cuda.synchronize()
        return [(g1, g2), (s_mel, s_len, s_y, s_yl), (keep, tables)]

    def close(self):
        """Drop the captured graphs (they hold NCCL kernels when world > 1): call before destroy_process_group()."""
        torch.cuda.synchronize()
        self._graphs.clear()

    # ---- checkpointing (reference models/model.py:346-376 saves optimizer.state_dict() and scheduler.model_step) --------------------
    def _torch_param_order(self):
        return [n for n, _ in self.model.named_parameters()]

    def state_dict(self):
        """{"optimizer": a torch.optim.Adam state_dict over model.parameters() (same layout the reference saves), "model_step": int,
        "dropout_counter": tensor or None}."""
        names = self._torch_param_order()
        t = int(self.state[1].item())
        state = {}
        for idx, n in enumerate(names):
            i = self.flat.index[n]
            o, sz, shape = self.flat.offsets[i], self.flat.sizes[i], self.flat.shapes[i]
            state[idx] = {"step": torch.tensor(float(t)), "exp_avg": self.flat.exp_avg[o:o + sz].view(shape).clone(),
                          "exp_avg_sq": self.flat.exp_avg_sq[o:o + sz].view(shape).clone()}
        tp = self.tp
        group = {"lr": self.lr(), "betas": (float(tp["beta1"]), float(tp["beta2"])), "eps": float(tp["eps"]),
                 "weight_decay": float(tp["weight_decay"]), "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "params": list(range(len(names)))}
        drop = self.path._dropout_state(self.device)
        return {"optimizer": {"state": state if t > 0 else {}, "param_groups": [group]}, "model_step": self.steps_done(),
                "dropout_counter": drop.counter.clone() if drop.counter is not None else None}

    def load_state_dict(self, sd):
        """Accepts state_dict() of this class or a checkpoint of the reference ({"optimizer_state_dict", "model_step"} entries map to
        "optimizer" / "model_step")."""
        opt = sd.get("optimizer", sd.get("optimizer_state_dict"))
        names = self._torch_param_order()
        t = 0
        with torch.no_grad():
            self.flat.exp_avg.zero_(); self.flat.exp_avg_sq.zero_()
            for idx, st in (opt["state"] if opt is not None else {}).items():
                i = self.flat.index[names[int(idx)]]
                o, sz = self.flat.offsets[i], self.flat.sizes[i]
                self.flat.exp_avg[o:o + sz].copy_(st["exp_avg"].reshape(-1))
                self.flat.exp_avg_sq[o:o + sz].copy_(st["exp_avg_sq"].reshape(-1))
                t = max(t, int(float(st["step"])))
        model_step = int(sd.get("model_step", t))
        self._set_schedule_step(model_step, adam_t=t)
        ctr = sd.get("dropout_counter")
        drop = self.path._dropout_state(self.device)
        if ctr is not None and drop.counter is not None:
            drop.counter.copy_(ctr)
        self._micro = 0

    # ---- introspection ------------------------------------------------------------------------------------------------------
    def lr(self):
        """Learning rate the NEXT optimiser step will use."""
        return float(self.state[0:1].view(torch.float32).item())

    def steps_done(self):
        """Optimiser steps taken == the reference scheduler's model_step."""
        return int(self.state[2].item()) - 1 if self.schedule else int(self.state[1].item())
