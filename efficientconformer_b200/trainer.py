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
        self._ptr_host = torch.empty(len(self.names), dtype=torch.int64).pin_memory() if torch.cuda.is_available() else \
            torch.empty(len(self.names), dtype=torch.int64)
        self._ptr_dev = torch.empty(len(self.names), dtype=torch.int64, device=device)

    def grad_view(self, name):
        i = self.index[name]
        return self.grads[self.offsets[i]:self.offsets[i] + self.sizes[i]].view(self.shapes[i])

    def pack(self, grads):
        """Gather the per-parameter gradient tensors (dict name -> contiguous fp32 tensor) into `self.grads`."""
        keep = []
        for i, n in enumerate(self.names):
            g = grads[n]
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous(); keep.append(g)
            assert g.numel() == self.sizes[i], n
            self._ptr_host[i] = g.data_ptr()
        self._ptr_dev.copy_(self._ptr_host, non_blocking=True)
        _lib.check(_lib.lib().ec_op_pack_flat(_lib.ptr(self._ptr_dev), _lib.ptr(self.offsets_dev), _lib.ptr(self.sizes_dev), len(self.names),
                                              _lib.ptr(self.grads), _lib.stream_ptr()))
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
        self.arena = torch.zeros(flat.total, dtype=dt, device=device)
        self.arena_t = torch.zeros(flat.total, dtype=dt, device=device)
        self._fwd, self._bwd, self._qkv, desc = {}, {}, {}, []
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
                self._qkv[id(params[n])] = (self.arena[o:o + 3 * D * D].view(3 * D, D), self.arena_t[o:o + 3 * D * D].view(D, 3 * D),
                                            torch.zeros(3 * D, dtype=torch.float32, device=device))
                desc.append((o, 3 * D, D, o))
                merged.update(names)
        for n, o, shape in zip(flat.names, flat.offsets, flat.shapes):
            is_matrix = len(shape) == 2 or (len(shape) == 3 and shape[-1] == 1)
            if not is_matrix:
                continue
            N, K = shape[0], shape[1]
            self._fwd[id(params[n])] = self.arena[o:o + N * K].view(N, K)
            if n not in merged:
                self._bwd[id(params[n])] = self.arena_t[o:o + N * K].view(K, N)
                desc.append((o, N, K, o))
        self.n_desc = len(desc)
        self.desc = torch.tensor(desc, dtype=torch.int64, device=device).contiguous()

    def refresh(self):
        """Run once per step, before the forward: the parameters changed in the previous optimiser step."""
        _ops.cast_into(self.flat.params, self.arena, self.pr)
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
    """One optimisation step of ModelCTC on a fixed batch shape.  `training_params` is the reference config's dict
    (optimizer Adam: beta1, beta2, eps, weight_decay; lr_schedule Transformer: schedule_dim, warmup_steps, K)."""

    def __init__(self, model, training_params, precision="bf16", process_group=None, sync_bn=True, use_cuda_graph=True, dropout_seed=0,
                 data_parallel=True):
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
        reducer = None
        if self.world > 1 and sync_bn:
            from .distributed import SyncBatchNormReducer
            reducer = SyncBatchNormReducer(process_group, self.device)
        self.reducer = reducer
        self.path = TrainingPath(model.encoder, model.fc, stats_reducer=reducer, dropout_seed=dropout_seed)
        self.flat = FlatParams(_qkv_adjacent_order(self.path.param_list()), self.device)
        model.encoder._drop_engines()                       # inference arenas were prepared from the old parameter storage
        weights = ArenaWeights(self.flat, precision, self.device)
        self.weights = weights if weights.supports(model.encoder) else None
        self.path.weights = self.weights
        self.state = torch.zeros(4, dtype=torch.int32, device=self.device)
        if sched == "Constant":
            self.state[0:1].view(torch.float32).fill_(float(training_params["lr_value"]))
        self.schedule = 1 if sched == "Transformer" else 0
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._static = None
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.launches_per_step = None

    # ---- pieces -----------------------------------------------------------------------------------------------------------
    def _forward_backward(self, mel, mel_len, targets, target_len):
        if self.weights is not None:
            self.weights.refresh()
        x, logits, out_len, tape = self.path.forward(mel, mel_len, self.precision, want_logits=True)
        if out_len is None:
            out_len = torch.full((mel.shape[0],), logits.shape[1], dtype=torch.int64, device=mel.device)
        mean, per, dlogits = ctc_loss_and_grad(logits, out_len, targets, target_len)
        grads = self.path.backward(tape, None, dlogits)
        keep = self.flat.pack(grads)
        self.loss.copy_(mean)
        return keep

    def _optimizer(self):
        tp = self.tp
        _ops.adam_step(self.flat.params, self.flat.grads, self.flat.exp_avg, self.flat.exp_avg_sq, self.state, float(tp["beta1"]),
                       float(tp["beta2"]), float(tp["eps"]), float(tp["weight_decay"]), grad_scale=1.0 / self.world, schedule=self.schedule,
                       K=float(tp.get("K", 0.0)), dim=float(tp.get("schedule_dim", 1.0)), warmup=float(tp.get("warmup_steps", 1.0)))

    def _all_reduce(self):
        if self.world > 1:
            dist.all_reduce(self.flat.grads, op=dist.ReduceOp.SUM, group=self.group)     # ONE bucket; the mean is folded into Adam

    def _step_eager(self, mel, mel_len, targets, target_len):
        with torch.no_grad():
            keep = self._forward_backward(mel, mel_len, targets, target_len)
            self._all_reduce()
            self._optimizer()
        del keep

    # ---- public -----------------------------------------------------------------------------------------------------------
    def step(self, mel, mel_len, targets, target_len):
        """mel (B, n_mels, T) fp32, mel_len (B,) int64 or None, targets (B, U) int64, target_len (B,) int64 -- CUDA tensors.
        Returns the (device) mean CTC loss of this batch; parameters, Adam moments, BatchNorm running statistics and the
        learning-rate schedule have advanced by one step."""
        for t in (mel, targets, target_len):
            if not t.is_cuda:
                raise RuntimeError("CTCTrainStep takes CUDA tensors (copy the batch with non_blocking H2D first)")
        if not self.use_cuda_graph:
            self._step_eager(mel, mel_len, targets, target_len)
            return self.loss
        key = (tuple(mel.shape), mel_len is not None, tuple(targets.shape))
        if self._graph is None or self._static[0] != key:
            self._capture(key, mel, mel_len, targets, target_len)
        _, s_mel, s_len, s_y, s_yl = self._static
        s_mel.copy_(mel, non_blocking=True)
        if s_len is not None:
            s_len.copy_(mel_len, non_blocking=True)
        s_y.copy_(targets, non_blocking=True); s_yl.copy_(target_len, non_blocking=True)
        self._graph[0].replay()
        if self._graph[1] is not None:                      # world > 1 with collectives kept outside the graph
            self._all_reduce()
            self._graph[1].replay()
        return self.loss

    def _capture(self, key, mel, mel_len, targets, target_len):
        s_mel, s_y, s_yl = mel.clone(), targets.clone(), target_len.clone()
        s_len = mel_len.clone() if mel_len is not None else None
        self._static = (key, s_mel, s_len, s_y, s_yl)
        # snapshot everything a step mutates: the warm-up step below must not count as a training step
        snap = [t.clone() for t in (self.flat.params, self.flat.exp_avg, self.flat.exp_avg_sq, self.state)]
        bufs = [b for b in self.model.buffers()]
        snap_b = [b.clone() for b in bufs]
        drop = self.path._dropout_state(self.device)
        snap_c = drop.counter.clone() if drop.counter is not None else None
        self._step_eager(s_mel, s_len, s_y, s_yl)           # warm-up: kernel attributes, allocator pools, NCCL communicators
        torch.cuda.synchronize()
        split = self.world > 1 and self.reducer is None     # no SyncBN collectives inside: keep NCCL outside the graphs
        g1, g2 = torch.cuda.CUDAGraph(), None
        # other threads (the NCCL watchdog) may touch the CUDA API while this thread captures
        mode = {"capture_error_mode": "thread_local"} if self.world > 1 else {}
        with torch.no_grad():
            with torch.cuda.graph(g1, **mode):
                keep = self._forward_backward(s_mel, s_len, s_y, s_yl)
                if not split:
                    self._all_reduce()
                    self._optimizer()
            if split:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, **mode):
                    self._optimizer()
        self._keep = keep
        self._graph = (g1, g2)
        with torch.no_grad():
            for t, s in zip((self.flat.params, self.flat.exp_avg, self.flat.exp_avg_sq, self.state), snap):
                t.copy_(s)
            for b, s in zip(bufs, snap_b):
                b.copy_(s)
            if snap_c is not None:
                drop.counter.copy_(snap_c)
        torch.cuda.synchronize()

    # ---- introspection ------------------------------------------------------------------------------------------------------
    def lr(self):
        return float(self.state[0:1].view(torch.float32).item())

    def steps_done(self):
        return int(self.state[1].item())
