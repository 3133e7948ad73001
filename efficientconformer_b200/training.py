"""Training step of the hot path (BASELINE.json configs[1]): the train-mode forward of
`ConformerEncoder` (+ the CTC `fc` head) and its hand-scheduled backward, launched operator by operator through the C ABI.

What the reference executes for this (reference models/model.py:239-259 -> models/model_ctc.py:57-68 ->
models/encoders.py:106-142 -> models/blocks.py:119-137 under `.train()`, then `loss.backward()`):
  * BatchNorm with BATCH statistics over every (padded) frame, running statistics updated (momentum 0.1, unbiased variance);
  * dropout (p = encoder_params["Pdrop"]) after `encoder.linear`, twice per feed-forward module, after the attention
    output projection and at the end of the convolution module (reference models/encoders.py:119, modules.py:389,391,486,521);
  * autograd through all of it.
Here the forward keeps an explicit tape (pre-activation tensors, LayerNorm inputs, GEMM operands in the activation type) and
`backward` walks it in reverse with the backward operators of include/effconf_b200.h; residual-branch gradients are accumulated
in place by the kernels (`accumulate` flags), not by framework adds.  Nothing here computes with PyTorch: tensors are only
allocated (caching allocator) and handed to the library as pointers.  There is no fallback: without the CUDA library every
operator raises.

Autograd integration (the drop-in side): `EncoderTrainFn` is ONE autograd node for the whole path, so the reference trainer's
`loss.backward()`, `GradScaler`, DDP gradient hooks and `optimizer.step()` work unchanged on the parameters of the holder modules.
"""
import torch

from . import ops as _ops_module
from . import _lib
from .encoders import relative_sinusoid_rows

_ops = _ops_module       # test seam: tests/test_train_glue_cpu.py swaps in a torch-CPU operator table to check the tape logic

# Which element operators are folded into the GEMM epilogues (ec_op_gemm_train) instead of running as their own bandwidth kernels:
#   w1   feed-forward W1: pre-activation z and dropout(Swish(z)) from one epilogue
#   res  dropout + alpha + residual behind W2 / attention output / pointwise conv 2 / encoder.linear
#   dz   Swish-dropout backward in the data-gradient GEMM of W2
#   ln   every LayerNorm backward also emits the masked, scaled, activation-type operand of the next GEMM of the backward chain
#   lnf  forward: the LayerNorm of the NEXT module is computed in the epilogue of the producing projection (ec_op_gemm_ln_train, rows of
#        up to 256 features): feed-forward 1 -> attention norm, attention output -> conv-module norm, pointwise conv 2 -> feed-forward 2 norm.
#        OFF by default: measured 12.91 ms vs 12.76 ms per step (45 launches fewer, but the LayerNorm passes inside a one-CTA-per-SM GEMM
#        epilogue cost more than the stand-alone kernel at full occupancy); kept selectable and tested.
# Chosen by measurement on the B200 (profiles/r2); EFFCONF_TRAIN_FUSE=w1,res,dz overrides.
import os as _os
FUSE = set(filter(None, _os.environ.get("EFFCONF_TRAIN_FUSE", "w1,res,dz,ln").split(",")))


class _null_context:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class DropoutState:
    """Counter-based dropout: mask bit of element i of site s at step n = hash(seed, n, s, i) < keep.  The step counter lives on the
    device so that a captured CUDA graph draws fresh masks on every replay; the backward recomputes the same bits from (site, i)."""

    def __init__(self, p: float, device, seed: int = 0):
        self.p = float(p)
        self.seed = int(seed)
        self.site = 0
        self.counter = None
        if self.p > 0.0:
            self.counter = _ops.dropout_counter(device, seed)

    def begin_step(self):
        self.site = 0
        if self.counter is not None:
            _ops.dropout_advance(self.counter)

    def next_site(self):
        self.site += 1
        return self.site


def _bump_batches_tracked(bn, sink):
    """nn.BatchNorm*.num_batches_tracked += 1 in training mode (a state_dict entry of the reference): collected in `sink`, all 16
    counters advance in ONE multi-tensor launch at the end of the forward (16 separate one-element kernels sat on the critical path)."""
    if getattr(bn, "num_batches_tracked", None) is not None:
        sink.append(bn.num_batches_tracked)


def _len_after_stride(length, stride):
    """reference models/modules.py:243, encoders.py:140: x_len = (x_len - 1) // s + 1 (B integers of length bookkeeping)."""
    return torch.div(length - 1, stride, rounding_mode="floor") + 1


def _w2(t):
    """Conv1d(k=1) weights [N, K, 1] are used as [N, K] matrices."""
    return t.reshape(t.shape[0], -1)


class TrainingPath:
    """Owns nothing but references to the parameter holders of a ConformerEncoder (+ optional fc head)."""

    def __init__(self, encoder, head=None, stats_reducer=None, dropout_seed: int = 0):
        if len(encoder.subsampling_module.layers) != 1:
            raise NotImplementedError("training mode supports the one-layer Conv2d front end (Efficient Conformer family)")
        self.encoder = encoder
        self.head = head
        self.specs = encoder.specs
        self.stats_reducer = stats_reducer      # distributed.SyncBatchNormReducer (forward_stats / backward_sums) or None
        self.p_drop = float(encoder.params.get("Pdrop", 0.0))
        self.dropout_seed = dropout_seed
        self._drop = None
        self._tables = {}
        self.weights = None                      # optional arena-backed operand provider (trainer.ArenaWeights)
        # Weight gradients are off the critical path of the backward (only the optimiser reads them): they run on a forked side
        # stream (event fork / join, capturable), overlapping the data-gradient chain on the main stream.
        self.side_wgrad = True
        self._side = None
        self._side_keep = []
        self._side_pos = None

    # ---- GEMM operands of the weights ----------------------------------------------------------------------------------------
    # Default: one cast / transposed cast per weight and use.  CTCTrainStep installs a provider backed by its flat parameter arena
    # (one cast launch and one multi-tensor transpose launch per step for all weights).
    def _w(self, weight, pr):
        """[N, K] operand of the forward GEMM (activation type)."""
        if self.weights is not None:
            return self.weights.act(weight)
        return _ops.cast_weight(_w2(weight), pr)

    def _wt(self, weight, pr):
        """[K, N] operand of the data-gradient GEMM dX = dY . W (activation type)."""
        if self.weights is not None:
            return self.weights.act_t(weight)
        return _ops.transpose_cast(_w2(weight), pr)

    def _wgrad(self, dy_act, x_act, pr, cast_first=False):
        """(dW, db) of a Linear on the side stream.  The operands are kept alive until the join at the end of backward().
        cast_first: dy_act is still an fp32 gradient; its cast to the activation type runs on the side stream as well."""
        if not self.side_wgrad or not dy_act.is_cuda:
            return _ops.linear_wgrad_bias(_ops.cast(dy_act, pr) if cast_first else dy_act, x_act, pr)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dy_act.device)
        main = torch.cuda.current_stream(dy_act.device)
        self._side.wait_stream(main)             # dy was produced on the main stream
        with torch.cuda.stream(self._side):
            d = _ops.cast(dy_act, pr) if cast_first else dy_act
            out = _ops.linear_wgrad_bias(d, x_act, pr)
        self._side_keep.append((dy_act, d, x_act))
        return out

    def side_stream(self, device):
        """The forked stream the weight gradients run on (None when they run in line); whoever enqueues other work on it must make sure
        the main stream joins it: backward() does so before its first launch and after its last."""
        if not self.side_wgrad or device.type != "cuda":
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(device=device)
        return self._side

    def _join_side(self, device, force=False):
        if self._side is not None and (self._side_keep or force):
            torch.cuda.current_stream(device).wait_stream(self._side)
        self._side_keep = []

    def _dgrad(self, dy_act, weight, pr, residual=None):
        return _ops.gemm(dy_act, self._wt(weight, pr), None, pr, residual=residual)[0]

    def _qkv(self, mhsa, pr):
        """([3D, D] forward operand, [3D] fp32 bias, handle for the backward) of the concatenated Q | K | V projection."""
        if self.weights is not None:
            return self.weights.qkv_act(mhsa), self.weights.qkv_bias(mhsa), mhsa
        w32, b = _ops.concat_qkv(mhsa)
        return _ops.cast_weight(w32, pr), b, w32

    def _qkv_dgrad(self, dqkv_act, handle, pr):
        if self.weights is not None:
            return _ops.gemm(dqkv_act, self.weights.qkv_act_t(handle), None, pr)[0]
        return _ops.gemm(dqkv_act, _ops.transpose_cast(handle, pr), None, pr)[0]

    # ---- helpers ------------------------------------------------------------------------------------------------------
    def param_list(self):
        """(name, Parameter) in a fixed order: encoder parameters, then the head's."""
        named = [("encoder." + k, p) for k, p in self.encoder.named_parameters()]
        if self.head is not None:
            named += [("fc." + k, p) for k, p in self.head.named_parameters()]
        return named

    def _table(self, t_pad, spec, pr, device):
        key = (t_pad, spec.dim_model, spec.group_size, spec.max_pos, pr, str(device))
        tab = self._tables.get(key)
        if tab is None:
            tab = _ops.cast(relative_sinusoid_rows(t_pad, spec.dim_model, spec.group_size, spec.max_pos).to(device), pr)
            self._tables[key] = tab
        return tab

    def _dropout_state(self, device):
        if self._drop is None:
            self._drop = DropoutState(self.p_drop, device, self.dropout_seed)
        return self._drop

    # ---- forward ------------------------------------------------------------------------------------------------------
    def forward(self, mel, mel_len, precision, want_logits=True):
        """mel (B, n_mels, T) fp32, mel_len (B,) integer tensor on the same device or None.
        Returns (x (B, T_out, D_last) fp32, logits (B, T_out, V) fp32 or None, out_len or None, tape)."""
        o = _ops
        pr = _lib.PRECISIONS[precision] if isinstance(precision, str) else precision
        enc = self.encoder
        drop = self._dropout_state(mel.device)
        drop.begin_step()
        B, F, T = mel.shape
        tape = {"pr": pr, "B": B, "blocks": []}

        sub = enc.subsampling_module.layers[0]
        a, sub_saved = o.SubsampleTrain.forward(mel, sub[0].weight, sub[0].bias, sub[1].weight, sub[1].bias, sub[1].running_mean,
                                                sub[1].running_var, pr, reduce_stats=self.stats_reducer)
        self._tracked = []
        _bump_batches_tracked(sub[1], self._tracked)
        T0 = (T - 1) // 2 + 1
        cur_len = None
        if mel_len is not None:
            cur_len = _len_after_stride(mel_len, 2)
        w_lin = self._w(enc.linear.weight, pr)
        site0 = drop.next_site()
        if "res" in FUSE:
            x = o.gemm_train(a, w_lin, enc.linear.bias, pr, drop, site=site0)[0]   # (B*T0, D0) fp32, dropout in the epilogue
        else:
            x = o.dropout_f32(o.gemm(a, w_lin, enc.linear.bias, pr)[0], drop, site0)
        tape["front"] = (a, sub_saved, w_lin, site0)

        Tc = T0
        n_blocks = len(self.specs)
        x_act_last = None
        pos = self._positional_projections(T0, pr, mel.device)
        for i, (spec, blk) in enumerate(zip(self.specs, enc.blocks)):
            last = i == n_blocks - 1
            if i == 0:
                self._join_pos(mel.device)
            x, Tn, bt, x_act = self._block_forward(blk, spec, x, B, Tc, cur_len, pr, drop, want_act_out=last and want_logits and self.head is not None,
                                                   pos=pos[i])
            tape["blocks"].append(bt)
            if spec.conv_stride > 1 and cur_len is not None:
                cur_len = _len_after_stride(cur_len, spec.conv_stride)
            Tc = Tn
            x_act_last = x_act
        D_last = self.specs[-1].dim_expand
        logits = None
        if want_logits:
            if self.head is None:
                raise RuntimeError("no fc head attached")
            w_fc = self._w(self.head.weight, pr)
            logits = o.gemm(x_act_last, w_fc, self.head.bias, pr)[0].view(B, Tc, -1)
            tape["head"] = (x_act_last, w_fc)
        tape["T_out"] = Tc
        if self._tracked:
            torch._foreach_add_(self._tracked, 1)
            self._tracked = []
        return x.view(B, Tc, D_last), logits, cur_len, tape

    def _proj_drop_res(self, a_act, w_act, bias, pr, drop, site, alpha, residual, next_ln=None):
        """out = residual + alpha * dropout_site(a W^T + bias): one GEMM with the dropout in its epilogue, or GEMM + element kernel.
        next_ln = (gamma, beta) of the LayerNorm that consumes `out`: returns (out, LN(out) in the activation type), from the same GEMM
        epilogue when the row fits one tile (<= 256 features), else from the LayerNorm kernel."""
        o = _ops
        if next_ln is not None and "lnf" in FUSE and "res" in FUSE and w_act.shape[0] <= 256:
            return o.gemm_ln_train(a_act, w_act, bias, pr, next_ln[0], next_ln[1], drop, alpha=alpha, residual=residual, site=site)
        if "res" in FUSE or drop.p == 0.0:
            out = o.gemm_train(a_act, w_act, bias, pr, drop, alpha=alpha, residual=residual, site=site)[0]
        else:
            out = o.dropout_residual(o.gemm(a_act, w_act, bias, pr)[0], drop, site, alpha, residual)
        if next_ln is None:
            return out
        return out, o.layernorm(out, next_ln[0], next_ln[1], pr, want_f32=False, want_act=True)[0]

    def _ffn_forward(self, holder, x, pr, drop, alpha=0.5, h0=None, next_ln=None):
        """reference models/modules.py:378-395 + the half-step residual of models/blocks.py:122,132.  h0: the module's LayerNorm output
        when the producer of x already computed it; next_ln: see _proj_drop_res (then returns (out, saved, LN_next(out)))."""
        o = _ops
        L = holder.layers
        if h0 is None:
            h0 = o.layernorm(x, L[0].weight, L[0].bias, pr, want_f32=False, want_act=True)[0]
        w1 = self._w(L[1].weight, pr)
        s1 = drop.next_site()
        # W1: the pre-activation z (kept for the backward) and s = dropout(Swish(z)) leave the same epilogue
        if "w1" in FUSE:
            _, z, s = o.gemm_train(h0, w1, L[1].bias, pr, drop, want_f32=False, want_act=True, want_act2=True, site2=s1)
        else:
            z = o.gemm(h0, w1, L[1].bias, pr, want_f32=False, want_act=True)[1]
            s = o.swish_dropout_fwd(z, drop, s1, pr)
        w2 = self._w(L[4].weight, pr)
        s2 = drop.next_site()
        # W2: x + alpha * dropout(s W2^T + b2) in the epilogue
        saved = lambda: (x, h0, z, s, s1, s2, alpha)
        if next_ln is not None:
            out, nxt = self._proj_drop_res(s, w2, L[4].bias, pr, drop, s2, alpha, x, next_ln=next_ln)
            return out, saved(), nxt
        out = self._proj_drop_res(s, w2, L[4].bias, pr, drop, s2, alpha, x)
        return out, saved()

    def _ffn_backward(self, holder, saved, d_out, pr, grads, prefix, dy=None, emit_next=None):
        """d_out: gradient w.r.t. the module output (fp32, modified in place); returns the gradient w.r.t. its input x.
        dy: the activation-type operand alpha * dropout_s2-mask * d_out when the producer of d_out already emitted it (LayerNorm
        backward, ops.layernorm_bwd(emit=...)).  emit_next = (scale, site): also return that operand for the next consumer of dx."""
        o = _ops
        x, h0, z, s, s1, s2, alpha = saved
        L = holder.layers
        drop = self._drop
        if dy is None:
            dy = o.dropout_cast_scaled(d_out, pr, alpha, drop, s2)             # d(W2 output) in the activation type
        dw_, db_ = self._wgrad(dy, s, pr)
        grads[f"{prefix}.layers.4.weight"], grads[f"{prefix}.layers.4.bias"] = dw_, db_
        # dz = dropout-mask * Swish'(z) * (dy W2) in the data-gradient GEMM's epilogue, straight into the activation type
        if "dz" in FUSE:
            dz = o.gemm_train(dy, self._wt(L[4].weight, pr), None, pr, drop, want_f32=False, want_act=True, aux=z, site_aux=s1)[1]
        else:
            dz = o.swish_dropout_bwd(z, self._dgrad(dy, L[4].weight, pr), drop, s1, pr)
        dw_, db_ = self._wgrad(dz, h0, pr)
        grads[f"{prefix}.layers.1.weight"], grads[f"{prefix}.layers.1.bias"] = dw_, db_
        dh0 = self._dgrad(dz, L[1].weight, pr)
        if emit_next is not None and "ln" in FUSE:
            dx, dg, db, dnext = o.layernorm_bwd(x, dh0, L[0].weight, dx_accum=d_out, emit=(pr, emit_next[0], drop, emit_next[1]))
        else:
            (dx, dg, db), dnext = o.layernorm_bwd(x, dh0, L[0].weight, dx_accum=d_out), None
        grads[f"{prefix}.layers.0.weight"], grads[f"{prefix}.layers.0.bias"] = dg, db
        return dx, dnext

    def _positional_projections(self, T0, pr, device):
        """[(R_i, E_i)] for every block: E_i = pos_layer_i(R_i) depends on the weights only (reference models/attentions.py:678 repeats it per
        utterance), so all of them run on a forked side stream at the start of the forward, beside the front end and the first modules."""
        o = _ops
        main = None
        if device.type == "cuda":
            if self._side_pos is None:
                self._side_pos = torch.cuda.Stream(device=device)
            main = torch.cuda.current_stream(device)
            self._side_pos.wait_stream(main)           # the weight operands were refreshed on the main stream
        out, T = [], T0
        ctx = torch.cuda.stream(self._side_pos) if main is not None else _null_context()
        with ctx:
            for spec, blk in zip(self.specs, self.encoder.blocks):
                m = blk.multi_head_self_attention_module
                D, H, G = spec.dim_model, spec.num_heads, spec.group_size
                R = self._table(T + (-T) % G, spec, pr, device)
                E = o.gemm(R, self._w(m.mhsa.pos_layer.weight, pr), m.mhsa.pos_layer.bias, pr, want_f32=False, want_act=True,
                           act_f16=o.attn_operands_f16(pr, D, H, G))[1]
                out.append((R, E))
                T = (T - 1) // spec.conv_stride + 1
        return out

    def _join_pos(self, device):
        if device.type == "cuda" and self._side_pos is not None:
            torch.cuda.current_stream(device).wait_stream(self._side_pos)

    def _block_forward(self, blk, spec, x, B, T, cur_len, pr, drop, want_act_out, pos):
        o = _ops
        D, De, H, G, st = spec.dim_model, spec.dim_expand, spec.num_heads, spec.group_size, spec.conv_stride
        if st > 1 and not spec.has_conv_res_proj:
            raise NotImplementedError("strided block without channel expansion (MaxPool residual) is not used by any shipped config")
        m = blk.multi_head_self_attention_module
        x1, ffn1, a_in = self._ffn_forward(blk.feed_forward_module1, x, pr, drop, next_ln=(m.norm.weight, m.norm.bias))
        # ---- attention module (reference models/modules.py:472-488, attentions.py:549-718)
        wqkv, bqkv, qkv_handle = self._qkv(m.mhsa, pr)
        ab16 = o.attn_operands_f16(pr, D, H, G)                      # split mode: plain fp16 q|k|v and E for the attention core
        qkv = o.gemm(a_in, wqkv, bqkv, pr, want_f32=False, want_act=True, act_f16=ab16)[1]
        R, E = pos
        att = o.relpos_attention_act(qkv.view(B, T, 3 * D), E, m.mhsa.u, m.mhsa.v, cur_len, H, G, pr)
        wo = self._w(m.mhsa.output_layer.weight, pr)
        s_att = drop.next_site()
        # ---- convolution module (reference models/modules.py:507-525) + block residual (blocks.py:98-114,129)
        Lc = blk.convolution_module.layers
        x2, c_in = self._proj_drop_res(att.view(B * T, D), wo, m.mhsa.output_layer.bias, pr, drop, s_att, 1.0, x1, next_ln=(Lc[0].weight, Lc[0].bias))
        wpw1 = self._w(Lc[2].weight, pr)
        zg = o.gemm(c_in, wpw1, Lc[2].bias, pr, want_f32=False, want_act=True)[1]
        gl = o.glu_fwd(zg, pr)
        h, dw_saved = o.DwConvTrain.forward(gl.view(B, T, De), Lc[4].weight, Lc[4].bias, Lc[5].weight, Lc[5].bias, Lc[5].running_mean,
                                            Lc[5].running_var, st, pr, reduce_stats=self.stats_reducer)
        _bump_batches_tracked(Lc[5], self._tracked)
        To = (T - 1) // st + 1
        xs = None
        if spec.has_conv_res_proj:
            xs = o.strided_rows(x2.view(B, T, D), st, pr).view(B * To, D)
            wres = self._w(blk.conv_res[1].weight, pr)
            res = o.gemm(xs, wres, blk.conv_res[1].bias, pr)[0]
        else:
            res = x2
        wpw2 = self._w(Lc[7].weight, pr)
        s_conv = drop.next_site()
        L2 = blk.feed_forward_module2.layers
        x3, h0_2 = self._proj_drop_res(h.view(B * To, De), wpw2, Lc[7].bias, pr, drop, s_conv, 1.0, res, next_ln=(L2[0].weight, L2[0].bias))
        x4, ffn2 = self._ffn_forward(blk.feed_forward_module2, x3, pr, drop, h0=h0_2)
        x_act, x5 = o.layernorm(x4, blk.norm.weight, blk.norm.bias, pr, want_f32=True, want_act=want_act_out)
        tape = dict(ffn1=ffn1, x1=x1, a_in=a_in, qkv=qkv, qkv_handle=qkv_handle, E=E, R=R, att=att, cur_len=cur_len, s_att=s_att, x2=x2, c_in=c_in, zg=zg,
                    h=h, dw_saved=dw_saved, xs=xs, s_conv=s_conv, ffn2=ffn2, x4=x4, T=T, To=To)
        return x5, To, tape, x_act

    # ---- backward -----------------------------------------------------------------------------------------------------
    def backward(self, tape, d_x=None, d_logits=None, stage_hook=None):
        """d_x (B, T_out, D_last) and / or d_logits (B, T_out, V) fp32 -> {parameter name: gradient} (fp32, reference names with
        `encoder.` / `fc.` prefixes).  stage_hook(i, grads) is called after block i's backward has been enqueued: every gradient of
        the head and of the blocks >= i is then in `grads` (the weight gradients possibly still running on side_stream())."""
        o = _ops
        pr, B = tape["pr"], tape["B"]
        enc = self.encoder
        grads = {}
        dx = None
        if self._side is not None and (d_logits if d_logits is not None else d_x).is_cuda:
            self._join_side((d_logits if d_logits is not None else d_x).device, force=True)   # e.g. the transposed weight operands of this step
        if d_x is not None:
            dx = o.own_f32(d_x).view(-1, d_x.shape[-1])
        if d_logits is not None:
            x_act, w_fc = tape["head"]
            dl = o.cast(d_logits.reshape(-1, d_logits.shape[-1]), pr)
            dw_, db_ = self._wgrad(dl, x_act, pr)
            grads["fc.weight"], grads["fc.bias"] = dw_, db_
            dx = self._dgrad(dl, self.head.weight, pr, residual=dx)
        if dx is None:
            raise RuntimeError("backward needs a gradient for the encoder output or the logits")
        a, sub_saved, w_lin, site0 = tape["front"]
        d_act = None
        for i in reversed(range(len(self.specs))):
            # block 0 hands the front end its operand (dropout mask of encoder.linear's dropout re-applied)
            dx, d_act = self._block_backward(enc.blocks[i], self.specs[i], tape["blocks"][i], dx, B, pr, grads, f"encoder.blocks.{i}",
                                             emit_next=(1.0, site0) if i == 0 else None)
            if stage_hook is not None:
                stage_hook(i, grads)
        if d_act is None:
            d_act = o.dropout_cast_scaled(dx, pr, 1.0, self._drop, site0)
        dw_, db_ = self._wgrad(d_act, a, pr)
        grads["encoder.linear.weight"], grads["encoder.linear.bias"] = dw_, db_
        da = self._dgrad(d_act, enc.linear.weight, pr)
        dw, db, dgam, dbet = o.SubsampleTrain.backward(da, sub_saved, reduce_stats=self.stats_reducer)
        p = "encoder.subsampling_module.layers.0"
        grads[f"{p}.0.weight"], grads[f"{p}.0.bias"], grads[f"{p}.1.weight"], grads[f"{p}.1.bias"] = dw, db, dgam, dbet
        self._join_side(da.device)               # every weight gradient is complete before the caller reads `grads`
        return grads

    def _block_backward(self, blk, spec, t, d_out, B, pr, grads, p, emit_next=None):
        """Returns (gradient w.r.t. the block input, its emitted operand for `emit_next` or None).  Every LayerNorm backward of the
        chain also emits the activation-type, dropout-masked operand its consumer needs ("ln" in FUSE), which used to be one
        ec_op_dropout pass per site."""
        o = _ops
        D, De, H, G, st = spec.dim_model, spec.dim_expand, spec.num_heads, spec.group_size, spec.conv_stride
        T, To = t["T"], t["To"]
        drop = self._drop
        fuse_ln = "ln" in FUSE
        # block LayerNorm (reference models/blocks.py:135)
        ffn2 = t["ffn2"]
        if fuse_ln:
            dx4, dg, db, dy_ffn2 = o.layernorm_bwd(t["x4"], d_out, blk.norm.weight, emit=(pr, ffn2[6], drop, ffn2[5]))
        else:
            (dx4, dg, db), dy_ffn2 = o.layernorm_bwd(t["x4"], d_out, blk.norm.weight), None
        grads[f"{p}.norm.weight"], grads[f"{p}.norm.bias"] = dg, db
        dx3, dy = self._ffn_backward(blk.feed_forward_module2, ffn2, dx4, pr, grads, f"{p}.feed_forward_module2", dy=dy_ffn2,
                                     emit_next=(1.0, t["s_conv"]))
        # convolution module + residual
        Lc = blk.convolution_module.layers
        c = f"{p}.convolution_module.layers"
        if dy is None:
            dy = o.dropout_cast_scaled(dx3, pr, 1.0, drop, t["s_conv"])        # gradient of the pw2 output (after its dropout)
        h2 = t["h"].view(B * To, De)
        dw_, db_ = self._wgrad(dy, h2, pr)
        grads[f"{c}.7.weight"], grads[f"{c}.7.bias"] = dw_.view(De, De, 1), db_
        dh = self._dgrad(dy, Lc[7].weight, pr)
        dgl, dw_dw, db_dw, dgam, dbet = o.DwConvTrain.backward(dh.view(B, To, De), t["dw_saved"], reduce_stats=self.stats_reducer)
        grads[f"{c}.4.weight"], grads[f"{c}.4.bias"] = dw_dw.view(De, 1, -1), db_dw
        grads[f"{c}.5.weight"], grads[f"{c}.5.bias"] = dgam, dbet
        dzg = o.glu_bwd(t["zg"], dgl.view(B * T, De), pr)
        dw_, db_ = self._wgrad(dzg, t["c_in"], pr)
        grads[f"{c}.2.weight"], grads[f"{c}.2.bias"] = dw_.view(2 * De, D, 1), db_
        dc_in = self._dgrad(dzg, Lc[2].weight, pr)
        if spec.has_conv_res_proj:
            # the residual branch sees the un-dropped gradient dx3
            dres = o.cast(dx3, pr) if drop.p > 0.0 else dy
            dw_, db_ = self._wgrad(dres, t["xs"], pr)
            grads[f"{p}.conv_res.1.weight"], grads[f"{p}.conv_res.1.bias"] = dw_.view(De, D, 1), db_
            dxs = self._dgrad(dres, blk.conv_res[1].weight, pr)
            acc = o.zeros_f32(B * T, D, dx3.device)
            o.strided_rows_bwd(dxs.view(B, To, D), acc.view(B, T, D), st)
        else:
            acc = dx3                                                         # identity residual: gradient passes straight through
        if fuse_ln:
            dx2, dg, db, do = o.layernorm_bwd(t["x2"], dc_in, Lc[0].weight, dx_accum=acc, emit=(pr, 1.0, drop, t["s_att"]))
        else:
            dx2, dg, db = o.layernorm_bwd(t["x2"], dc_in, Lc[0].weight, dx_accum=acc)
            do = o.dropout_cast_scaled(dx2, pr, 1.0, drop, t["s_att"])
        grads[f"{c}.0.weight"], grads[f"{c}.0.bias"] = dg, db
        # attention module
        m = blk.multi_head_self_attention_module
        a = f"{p}.multi_head_self_attention_module"
        att2 = t["att"].view(B * T, D)
        dw_, db_ = self._wgrad(do, att2, pr)
        grads[f"{a}.mhsa.output_layer.weight"], grads[f"{a}.mhsa.output_layer.bias"] = dw_, db_
        datt = self._dgrad(do, m.mhsa.output_layer.weight, pr)
        dqkv_act, dE, du, dv = o.relpos_attention_bwd_act(t["qkv"].view(B, T, 3 * D), t["E"], m.mhsa.u, m.mhsa.v, t["cur_len"], H, G,
                                                          datt.view(B, T, D), pr)          # dq | dk | dv leave in the activation type
        dqkv_act = dqkv_act.view(B * T, 3 * D)
        grads[f"{a}.mhsa.u"], grads[f"{a}.mhsa.v"] = du, dv
        dw_, db_ = self._wgrad(dE, t["R"], pr, cast_first=True)         # cast + weight gradient of the positional projection: side stream
        grads[f"{a}.mhsa.pos_layer.weight"], grads[f"{a}.mhsa.pos_layer.bias"] = dw_, db_
        dwqkv, dbqkv = self._wgrad(dqkv_act, t["a_in"], pr)            # [3D, D]: rows q | k | v
        for j, nm in enumerate(("query", "key", "value")):
            grads[f"{a}.mhsa.{nm}_layer.weight"] = dwqkv[j * D:(j + 1) * D]
            grads[f"{a}.mhsa.{nm}_layer.bias"] = dbqkv[j * D:(j + 1) * D]
        da_in = self._qkv_dgrad(dqkv_act, t["qkv_handle"], pr)
        ffn1 = t["ffn1"]
        if fuse_ln:
            dx1, dg, db, dy_ffn1 = o.layernorm_bwd(t["x1"], da_in, m.norm.weight, dx_accum=dx2, emit=(pr, ffn1[6], drop, ffn1[5]))
        else:
            (dx1, dg, db), dy_ffn1 = o.layernorm_bwd(t["x1"], da_in, m.norm.weight, dx_accum=dx2), None
        grads[f"{a}.norm.weight"], grads[f"{a}.norm.bias"] = dg, db
        return self._ffn_backward(blk.feed_forward_module1, ffn1, dx1, pr, grads, f"{p}.feed_forward_module1", dy=dy_ffn1, emit_next=emit_next)


class EncoderTrainFn(torch.autograd.Function):
    """One autograd node for the whole train-mode hot path: forward runs the CUDA operator sequence and keeps the tape; backward
    runs the hand-scheduled CUDA backward and hands each parameter its gradient."""

    @staticmethod
    def forward(ctx, path, mel, mel_len, precision, want_logits, *params):
        x, logits, out_len, tape = path.forward(mel, mel_len, precision, want_logits)
        ctx.set_materialize_grads(False)             # an unused output (x when only the logits feed the loss) arrives as None
        ctx.path, ctx.tape = path, tape
        ctx.names = [n for n, _ in path.param_list()]
        if out_len is not None:
            ctx.mark_non_differentiable(out_len)
        if logits is None:
            return x, None, out_len
        return x, logits, out_len

    @staticmethod
    def backward(ctx, d_x, d_logits, _d_len=None):
        tape, ctx.tape = ctx.tape, None
        if tape is None:
            raise RuntimeError("the training tape of this forward has already been consumed (retain_graph is not supported)")
        grads = ctx.path.backward(tape, d_x, d_logits)
        out = []
        for i, n in enumerate(ctx.names):
            out.append(grads[n] if ctx.needs_input_grad[5 + i] else None)
        return (None, None, None, None, None, *out)
