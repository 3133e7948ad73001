#!/usr/bin/env python
"""Benchmark of the hot path on synthetic 80-mel batches (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16x2|bf16|tf32]
                    [--mode train|forward] [--config NAME] [--batch B] [--frames T] [--no-extras] [--no-cpu-baseline]

Prints ONE JSON line (rank 0).  metric = mel frames / second, whole job over all N GPUs.

Headline (top-level keys), default --mode train: BASELINE.json configs[1] at the north_star target shape -- one optimisation step of
EfficientConformerCTCSmall = train-mode forward (batch-statistics BatchNorm, dropout 0.1) + CTC loss + backward + gradient all-reduce
(N > 1) + Adam with the Transformer schedule, batch 32/GPU x 80 x 1000, through efficientconformer_b200.trainer.CTCTrainStep.
  value        batch resident in HBM; every step timed by its own CUDA-event pair on the launching stream, 256 MiB L2 flush between
               steps, max over ranks
  e2e          the same step through the public API with HOST batches: pinned mel / targets -> H2D on a copy stream, step, loss read
               back to the host, all inside the wall-clock timed region
  roofline     the dominant tensor-core operator of the step: algorithmic FLOPs / CUDA-event time of its calls in eager steps of
               this run, against MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (a port of the reference algorithm, pinned to the reference's outputs) on the host cores
Sub-records of the same line (skipped with --no-extras), each measured in this run:
  forward      the north_star target: inference forward + fc + CTC loss at 32 x (80 x 1000), with its own per-kernel rooflines
               (FFN / attention / GEMM against the tensor peak, depthwise conv against HBM) and parity against the CPU oracle
  sweep        forward and training throughput at T in {500, 1000, 2000}
  modes        the other operand modes (bf16, tf32) on the headline shapes
  ragged       configs[1] as literally worded: LibriSpeech-shaped batches, T ~ U{200..1600}, sorted and padded like collate_fn_pad
  configs      configs[2] EfficientConformerCTCLarge forward + CTC (B = 32) and configs[4] ConformerCTCLarge B = 8, T in 500..4000
               (+ the same two in plain bf16 with per-kernel rooflines: what the tcgen05 kernels reach when a tile has work)
  transducer_joint   configs[3]: the Transducer joint network + RNN-T loss, forward and forward + backward, on a B = 16 lattice
  torch_eager_b200   the UNMODIFIED reference modules (staged under baseline/_ref by tools/stage_reference.py) run eagerly on this GPU
               in fp32 and under autocast -- SURVEY.md 8(d) "the real bar"; N = 1 only
--impl reference times the CPU oracle port of the same step on the host cores (the reference arm of the task contract).
Multi-GPU: one process per GPU (torchrun); utterances shard over ranks; training exchanges SyncBatchNorm statistics and one flat
gradient bucket over NCCL, the forward has no collective; barrier + max-over-ranks timing."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from efficientconformer_b200.config import SHIPPED_ENCODER_PARAMS, resolve_blocks, stage_lengths  # noqa: E402
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel, synthetic_targets  # noqa: E402

METRIC = "encoder_mel_frames_per_sec"
UNIT = "frames/s"
DEFAULT_CONFIG = "EfficientConformerCTCSmall"
TRAINING_PARAMS = dict(optimizer="Adam", beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-6, lr_schedule="Transformer", schedule_dim=240,
                       warmup_steps=10000, K=2)      # reference configs/EfficientConformerCTCSmall.json:53-69


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback of B200_PROFILING.md"}


def ncu_traffic(precision, kernel_prefix, stem="ncu_full_train"):
    """Average DRAM bytes (read + write) per launch of one kernel from the committed `ncu --set full` capture of the same workload
    (profiles/r2/, produced by tools/summarize_ncu.py from tools/gpu_r2q.sh); None when no capture exists for this precision.  A
    cross-reference measured under ncu in a separate run (cold caches, serialised), not in this run: the source file is named."""
    path = os.path.join(ROOT, "profiles", "r2", f"{stem}_{precision}_summary.json")
    if not os.path.exists(path):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    for k in json.load(open(path))["kernels"]:
        if not k["kernel"].startswith(kernel_prefix):
            continue
        for key in ("dram_read", "dram_write"):
            val, unit = k[key].split()
            tot += float(val) * scale[unit]
        n += 1
    return {"bytes_per_launch": round(tot / n), "launches_sampled": n, "source": os.path.relpath(path, ROOT)} if n else None


def tensor_peak(precision, pk):
    """Algorithmic FLOP/s ceiling of an operand mode: tf32 operands run the tensor core at half the bf16 rate; the split mode issues
    two bf16 MMAs per product, so its algorithmic ceiling is also half the bf16 peak.  The reported `peak` stays the measured bf16
    figure (the contract's denominator); `mode_ceiling` carries this number next to it."""
    return pk["bf16_tflops"] * (1.0 if precision == "bf16" else 0.5)


def flops_per_frame(params, vocab, T, B):
    """Algorithmic forward FLOPs per mel frame (2 x MAC, real dims, pos_layer once per batch): SURVEY.md 8(d), == FlopCounterMode
    on the reference minus the B-redundant positional projection."""
    specs = resolve_blocks(params)
    lens, t_out = stage_lengths(params, T)
    f, t, cin, tot = params["n_mels"], T, 1, 0.0
    for l in range(params["subsampling_layers"]):
        f, t = f // 2, (t - 1) // 2 + 1
        tot += 18.0 * cin * params["subsampling_filters"][l] * f * t
        cin = params["subsampling_filters"][l]
    tot += 2.0 * t * (f * cin) * specs[0].dim_model
    for s, Tb in zip(specs, lens):
        D, De, H, G, st, k = s.dim_model, s.dim_expand, s.num_heads, s.group_size, s.conv_stride, s.kernel_size
        To, Tp = (Tb - 1) // st + 1, Tb + (-Tb) % G
        Tg, d = Tp // G, G * D // H
        tot += 16.0 * Tb * D * D + 16.0 * To * De * De + 8.0 * Tb * D * D + 2.0 * (2 * Tp - G) * D * D / B
        tot += 4.0 * H * Tg * Tg * d + 2.0 * H * Tg * (2 * Tg - 1) * d
        tot += 4.0 * Tb * D * De + 2.0 * To * De * De + 2.0 * To * De * k
        if s.has_conv_res_proj:
            tot += 2.0 * To * D * De
    tot += 2.0 * t_out * specs[-1].dim_expand * vocab
    return tot / T


def out_frames(params, T):
    return stage_lengths(params, T)[1]


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill(); out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# =====================================================================================================================
# CPU arm: the oracle port on the host cores (cpu_baseline and --impl reference)
# =====================================================================================================================
def cpu_oracle_pass(sd, params, mel, mel_len, y_fn):
    from oracle import conformer_oracle as O           # checker / CPU baseline only (never the product path)
    t0 = time.perf_counter()
    logits, out_len = O.model_ctc_forward_mel(sd, params, mel, mel_len)
    y, y_len = y_fn(out_len)
    loss, _ = O.ctc_loss(logits, out_len, y, y_len)
    return time.perf_counter() - t0, logits, out_len, float(loss)


def best_cpu_threads(sd, params, vocab):
    """The CPU path is many small ATen ops: beyond a few dozen threads the fork/join cost dominates (128 threads are >20x
    slower than 32 on the GPU box's host).  Give the CPU arm its best shot: time one small pass per candidate count."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    mel = synthetic_mel(4, 500, seed=9)
    ln = torch.full((4,), 500, dtype=torch.int64)
    yf = lambda ol: synthetic_targets(ol, vocab, seed=4)
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            cpu_oracle_pass(sd, params, mel, ln, yf)
            t = min(cpu_oracle_pass(sd, params, mel, ln, yf)[0] for _ in range(2))
            if t < best_t:
                best, best_t = c, t
            if t > 4 * best_t:
                break
    torch.set_num_threads(best)
    return best


def cpu_oracle_train_pass(leaf, opt, mel, mel_len, y, y_len, params):
    """One training step of the CPU oracle: train-mode forward, CTC loss, autograd backward, torch.optim.Adam (the optimiser the
    reference constructs, models/model.py:88-93).  Dropout is not applied (the oracle is deterministic; its cost is negligible)."""
    from oracle import conformer_oracle as O           # checker / CPU baseline only (never the product path)
    t0 = time.perf_counter()
    bn = {"updates": {}}
    logits, out_len = O.model_ctc_forward_mel(leaf, params, mel, mel_len, bn=bn)
    loss, _ = O.ctc_loss(logits, out_len, y, y_len)
    opt.zero_grad()
    loss.backward()
    opt.step()
    with torch.no_grad():
        for k, v in bn["updates"].items():
            leaf["encoder." + k].copy_(v)
    return time.perf_counter() - t0, float(loss.detach())


def cpu_train_setup(params, vocab):
    sd = seeded_state_dict(params, vocab, seed=0, prefix_encoder="encoder.")
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    opt = torch.optim.Adam([v for v in leaf.values() if v.requires_grad], lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    return sd, leaf, opt


def workload_config(args, mode):
    """The `config` object: identical in the GPU arm and the CPU (--impl reference) arm for the same flags."""
    B, T = args.batch, args.frames
    if mode == "train":
        what = (f"{args.config} CTC training step (train-mode forward, dropout {args.pdrop}, CTC loss, backward, Adam + Transformer schedule), "
                f"batch {B}/GPU x 80-mel x {T} frames, full-length utterances (BASELINE.json configs[1] at the north_star target shape)")
    else:
        what = (f"{args.config} encoder forward + fc + CTC loss, batch {B}/GPU x 80-mel x {T} frames (BASELINE.json north_star target shape)")
    return {"workload": what, "mode": mode, "batch_per_gpu": B, "global_batch": B * args.gpus, "frames": T, "n_mels": 80,
            "weights": "seeded random init",
            "optimizer": "Adam + Transformer schedule after EVERY batch (the shipped config accumulates 2 micro-batches per optimiser step: this "
                         "measures more optimiser work per frame, not less)" if mode == "train" else None,
            "l2": "GPU arm: 256 MiB write between timed steps (L2 flush)",
            "parallelism": f"dp{args.gpus}: utterances sharded over ranks" + ("; SyncBatchNorm statistics + one flat gradient bucket all-reduced over NCCL"
                                                                               if mode == "train" and args.gpus > 1 else "; no data-path collective")}


def run_reference(args, rank, world):
    """Reference arm: the CPU implementation of the path (oracle port; train mode: + torch autograd + torch.optim.Adam) on the host
    cores, on the SAME batch shape as the GPU arm whenever K + W steps of it finish within a few minutes (else a bounded sample)."""
    if rank != 0:
        return
    params, vocab = SHIPPED_ENCODER_PARAMS[args.config]
    train = args.mode == "train"
    sd, leaf, opt = cpu_train_setup(params, vocab)
    cores = best_cpu_threads(sd, params, vocab)
    T = args.frames
    t_out = out_frames(params, T)
    # one utterance-step to size the sample: the CPU arm may spend about `budget_s` in total
    mel1 = synthetic_mel(2, T, seed=1); len1 = torch.full((2,), T, dtype=torch.int64)
    y1, yl1 = synthetic_targets(torch.full((2,), t_out), vocab, seed=4)
    if train:
        probe = min(cpu_oracle_train_pass(leaf, opt, mel1, len1, y1, yl1, params)[0] for _ in range(2)) / 2
    else:
        with torch.no_grad():
            probe = min(cpu_oracle_pass(sd, params, mel1, len1, lambda ol: (y1, yl1))[0] for _ in range(2)) / 2
    n_pass = args.steps + max(1, min(args.warmup, 2))
    B = args.batch
    while B > 2 and probe * B * 0.6 * n_pass > args.cpu_budget:      # per-utterance cost drops ~40 % in a full batch
        B //= 2
    mel = synthetic_mel(B, T, seed=1)
    mel_len = torch.full((B,), T, dtype=torch.int64)
    y, y_len = synthetic_targets(torch.full((B,), t_out), vocab, seed=4)
    if train:
        one = lambda: cpu_oracle_train_pass(leaf, opt, mel, mel_len, y, y_len, params)[0]
    else:
        def one():
            with torch.no_grad():
                return cpu_oracle_pass(sd, params, mel, mel_len, lambda ol: (y, y_len))[0]
    for _ in range(max(1, min(args.warmup, 2))):
        one()
    times = [one() for _ in range(args.steps)]
    total = sum(times)
    value = B * T * args.steps / total
    what = "train-mode fwd + CTC + autograd backward + torch.optim.Adam" if train else "fwd + fc + CTC loss"
    sample = (f"{args.steps} steps of B={B} x 80 x {T} ({what}), fp32, torch CPU {torch.get_num_threads()} threads (best of 8/16/32/64/all on a "
              f"{os.cpu_count()}-core host)" + ("" if B == args.batch else f"; bounded sample: B reduced from {args.batch} to fit {args.cpu_budget:.0f} s"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, args.mode), "sample_batch": B,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# =====================================================================================================================
# GPU arm helpers
# =====================================================================================================================
class Env:
    """Per-process context: device, process group, L2-flush buffer, barrier, max-over-ranks reduction."""

    def __init__(self, rank, world, local_rank):
        self.rank, self.world = rank, world
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.local_rank = local_rank
        self.dist = None
        if world > 1:
            os.environ.setdefault("NCCL_DEBUG", "WARN")
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)


def _log(rank, msg):
    if os.environ.get("EFFCONF_BENCH_VERBOSE"):
        print(f"[bench rank {rank} t={time.perf_counter():.1f}] {msg}", file=sys.stderr, flush=True)


def timed_steps(env, fn, steps, warmup):
    """W warm-up calls, then K calls each bracketed by its own CUDA-event pair on the launching stream with an L2 flush before it;
    barrier + synchronize on both sides; returns (sum of the K durations in ms as the max over ranks, per-step list of this rank)."""
    for _ in range(max(warmup, 3)):
        fn(); env.flush.zero_()
    env.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        env.flush.zero_()
        a.record(); fn(); b.record()
    env.barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    return env.max_over_ranks(sum(step_ms)), step_ms


def e2e_pipeline(env, host_tensors, step_fn, n_steps, res=None):
    """K steps with HOST inputs: every step's batch goes pinned host -> device on a copy stream (2-deep pipeline: the copy of step k+1
    and the read-back of step k-1 overlap the compute of step k, what a DataLoader with pinned memory does), `step_fn(*device_batch)`
    returns the device loss, which is read back to the host every step.  Returns the K host losses.
    `res` (a dict kept by the caller across calls) holds what a data loader allocates once -- the copy stream, the two device
    staging batches, the pinned loss words, the events: allocating them inside the timed region put cudaMalloc / cudaHostAlloc calls
    (device-synchronising, 20 - 180 ms) into some runs and not others."""
    dev = env.dev
    res = {} if res is None else res
    if not res:
        res["copy_stream"] = torch.cuda.Stream(device=dev)
        res["loss_host"] = torch.zeros(2, dtype=torch.float32).pin_memory()
        res["bufs"] = [tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host_tensors) for _ in range(2)]
        res["ready"] = [torch.cuda.Event(), torch.cuda.Event()]
        res["done"] = [torch.cuda.Event(), torch.cuda.Event()]
    copy_stream, loss_host, bufs, ready, done = res["copy_stream"], res["loss_host"], res["bufs"], res["ready"], res["done"]
    cur = torch.cuda.current_stream()
    copy_stream.wait_stream(cur)                                  # earlier users of the staging batches have been enqueued on `cur`
    out = []
    used = [False, False]

    def stage(k):
        with torch.cuda.stream(copy_stream):
            if used[k & 1]:
                copy_stream.wait_event(done[k & 1])               # the compute that used this buffer has finished
            for d, t in zip(bufs[k & 1], host_tensors):
                d.copy_(t, non_blocking=True)                     # pinned host -> device
            ready[k & 1].record(copy_stream)
    trace = [] if os.environ.get("EFFCONF_E2E_TRACE") == "1" else None
    stage(0)
    for k in range(n_steps):
        t_a = time.perf_counter()
        if k + 1 < n_steps:
            stage(k + 1)
        cur.wait_event(ready[k & 1])
        t_b = time.perf_counter()
        loss = step_fn(*bufs[k & 1])
        t_c = time.perf_counter()
        loss_host[k & 1].copy_(loss.reshape(()), non_blocking=True)   # D2H of the step's result
        done[k & 1].record(cur)
        used[k & 1] = True
        if k >= 1:
            done[(k - 1) & 1].synchronize()
            out.append(float(loss_host[(k - 1) & 1]))
        if trace is not None:
            trace.append((round(1e3 * (t_b - t_a), 2), round(1e3 * (t_c - t_b), 2), round(1e3 * (time.perf_counter() - t_c), 2)))
    done[(n_steps - 1) & 1].synchronize()
    out.append(float(loss_host[(n_steps - 1) & 1]))
    if trace is not None:
        print("e2e trace (ms per step: stage next batch, step call, wait for step k-1):", trace, file=sys.stderr)
    return out


E2E_REPS = []      # wall seconds of every repetition of the last timed_e2e call (reported next to the value)


def timed_e2e(env, host_tensors, step_fn, steps, reps=2):
    """Wall clock around K pipelined steps, max over ranks.  The host thread is on the critical path of this loop (it launches step
    k + 1 only after reading the loss of step k - 1), so one host stall -- the driver lock taken by the concurrent nvidia-smi clock
    sampler, a scheduler hiccup -- lands in the number (a 15.0 ms reading next to 12.3 / 12.4 ms in the runs before and after it was
    observed; the allocations that caused most of it now happen once, in the warm-up call): the K-step region is measured `reps` times
    and the fastest repetition is reported, all of them listed."""
    res = {}
    e2e_pipeline(env, host_tensors, step_fn, 3, res)              # warm-up: also allocates the pipeline's staging buffers
    torch.cuda.synchronize()
    del E2E_REPS[:]
    best, losses = None, None
    for _ in range(max(1, reps)):
        env.barrier()
        t0 = time.perf_counter()
        ls = e2e_pipeline(env, host_tensors, step_fn, steps, res)
        torch.cuda.synchronize()
        sec = env.max_over_ranks(time.perf_counter() - t0)
        env.barrier()
        E2E_REPS.append(round(1e3 * sec / steps, 4))
        if best is None or sec < best:
            best, losses = sec, ls
    return best, losses


def build_ctc_model(config, precision, dev, train, pdrop=None):
    from efficientconformer_b200 import ModelCTC
    params, vocab = SHIPPED_ENCODER_PARAMS[config]
    params = dict(params)
    if pdrop is not None:
        params["Pdrop"] = pdrop
    sd = seeded_state_dict(SHIPPED_ENCODER_PARAMS[config][0], vocab, seed=0, prefix_encoder="encoder.")
    model = ModelCTC(params, {"vocab_size": vocab}, precision=precision)
    model.load_state_dict(sd, strict=False)
    model = model.to(dev)
    return (model.train() if train else model.eval()), sd, params, vocab


# ---- forward (inference) ---------------------------------------------------------------------------------------------------
def forward_record(env, config, precision, B, T, steps, warmup, want_e2e=True, want_kernels=True, want_parity=False):
    """Inference forward + fc + CTC loss of `config` at B x 80 x T per GPU: device-timed value, e2e through ModelCTC.forward_mel with host
    batches, per-kernel rooflines from an eager profiled pass, parity against the CPU oracle (rank 0, N = 1)."""
    from efficientconformer_b200 import _lib
    from efficientconformer_b200.model_ctc import ctc_loss
    import ctypes as C
    dev = env.dev
    was = torch.is_grad_enabled()
    torch.set_grad_enabled(False)
    try:
        model, sd, params, vocab = build_ctc_model(config, precision, dev, train=False)
        mel_h = synthetic_mel(B, T, seed=1 + env.rank).pin_memory()
        len_h = torch.full((B,), T, dtype=torch.int64).pin_memory()
        mel_d, len_d = mel_h.to(dev), len_h.to(dev)
        t_out = out_frames(params, T)
        y, y_len = synthetic_targets(torch.full((B,), t_out), vocab, seed=4)
        y_d, yl_d = y.to(dev), y_len.to(dev)
        last = {}

        def step_resident():
            logits, out_len, _ = model.forward_mel(mel_d, len_d)
            last["loss"] = ctc_loss(logits, out_len, y_d, yl_d)[0]

        total_ms, step_ms = timed_steps(env, step_resident, steps, warmup)
        frames = env.world * B * T * steps
        fpf = flops_per_frame(params, vocab, T, B)
        pk = peaks()
        rec = {"config": config, "precision": precision, "batch_per_gpu": B, "frames": T, "steps": steps,
               "value": frames / (total_ms / 1e3), "unit": UNIT, "ms_per_step": total_ms / steps,
               "step_ms_min_med_max": [round(min(step_ms), 4), round(statistics.median(step_ms), 4), round(max(step_ms), 4)],
               "algorithmic_mflop_per_frame": round(fpf / 1e6, 3),
               "whole_forward_tflops": round(fpf * B * T / (total_ms / steps) / 1e9, 2),
               "whole_forward_frac_of_bf16_peak": round(fpf * B * T / (total_ms / steps) / 1e9 / pk["bf16_tflops"], 4),
               "loss": float(last["loss"])}
        prec = _lib.PRECISIONS[precision]
        eng = model.encoder._engines[prec][0]
        rec["launches_per_step"] = _lib.lib().ec_engine_last_launches(eng) + 4      # + CTC: i64->i32, lse/argmax, alpha, mean
        if want_e2e:
            def step_host(m, l):
                logits, out_len, _ = model.forward_mel(m, l)
                return ctc_loss(logits, out_len, y_d, yl_d)[0]
            sec, _ = timed_e2e(env, (mel_h, len_h), step_host, steps)
            rec["e2e"] = {"value": frames / sec, "unit": UNIT, "ms_per_step": 1e3 * sec / steps, "ms_per_step_repetitions": list(E2E_REPS),
                          "h2d_bytes_per_step": mel_h.numel() * 4 + len_h.numel() * 8, "d2h_bytes_per_step": 4}
        if want_kernels:
            # per-kernel profile of eager forwards (CUDA events around every launch, launching stream)
            L = _lib.lib()
            ncat = L.ec_profile_categories()
            acc = {}
            model.encoder.use_cuda_graph = False
            L.ec_engine_set_profiling(eng, 1)
            reps = 5
            for r in range(reps + 1):
                model.forward_mel(mel_d, len_d)
                ms, fl, by = (C.c_double * ncat)(), (C.c_double * ncat)(), (C.c_double * ncat)()
                cnt = (C.c_int32 * ncat)()
                _lib.check(L.ec_engine_profile_read(eng, ms, fl, by, cnt))
                if r == 0:
                    continue
                for i in range(ncat):
                    a = acc.setdefault(L.ec_profile_category_name(i).decode(), [0.0, 0.0, 0.0, 0])
                    a[0] += ms[i] / reps; a[1] = fl[i]; a[2] = by[i]; a[3] = cnt[i]
            L.ec_engine_set_profiling(eng, 0)
            model.encoder.use_cuda_graph = True
            kernels = []
            classes = {"gemm_tc_kernel": [0.0, 0.0, 0], "ffn": [0.0, 0.0, 0], "relpos_attention": [0.0, 0.0, 0]}
            for name, (ms_, fl_, by_, n_) in acc.items():
                if n_ == 0:
                    continue
                kernels.append({"kernel": name, "launches": n_, "ms": round(ms_, 4), "tflops": round(fl_ / ms_ / 1e9, 2) if ms_ > 0 else None,
                                "gbs": round(by_ / ms_ / 1e6, 1) if ms_ > 0 else None})
                cls = "ffn" if name in ("ffn_fused", "gemm_ffn_w1_swish", "gemm_ffn_w2_res") else "relpos_attention" if name == "relpos_attention" \
                    else "gemm_tc_kernel" if name.startswith("gemm_") else None
                if cls:
                    c = classes[cls]; c[0] += ms_; c[1] += fl_; c[2] += n_
            tot = sum(k["ms"] for k in kernels)
            desc = {"gemm_tc_kernel": "gemm_tc_kernel (tcgen05 + TMA + TMEM): the Linear / pointwise-conv launches outside the feed-forward modules",
                    "ffn": "feed-forward modules (ffn_fused_kernel in bf16 mode, two gemm_tc_kernel launches in the other modes; tcgen05 + TMA)",
                    "relpos_attention": "relative-position (grouped) attention core"}
            roofs = []
            for cls, (ms_, fl_, n_) in classes.items():
                if n_ == 0 or ms_ <= 0:
                    continue
                ach = fl_ / ms_ / 1e9
                roofs.append({"kernel": desc[cls], "bound": "tensor", "achieved": round(ach, 2), "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                              "frac": round(ach / pk["bf16_tflops"], 4), "mode_ceiling": tensor_peak(precision, pk),
                              "traffic": ncu_traffic(precision, "relpos_attn" if cls == "relpos_attention" else "gemm_tc_kernel", "ncu_full_fwd"),
                              "launches": n_, "avg_launch_us": round(1e3 * ms_ / n_, 2), "share_of_forward": round(ms_ / tot, 3), "target_frac": 0.5})
            dw = acc.get("dwconv_bn_swish")
            if dw and dw[0] > 0:
                gbs = dw[2] / dw[0] / 1e6
                roofs.append({"kernel": "dwconv_bn_swish_kernel (GLU output -> depthwise conv -> BatchNorm(eval) -> Swish)", "bound": "hbm",
                              "achieved": round(gbs, 1), "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": round(gbs / pk["hbm_gbs"], 4), "traffic": None,
                              "launches": dw[3], "avg_launch_us": round(1e3 * dw[0] / dw[3], 2), "share_of_forward": round(dw[0] / tot, 3),
                              "target_frac": 0.7, "note": "working set is L2-resident at this shape (DESIGN.md)"})
            rec["rooflines"] = roofs
            rec["kernels"] = kernels
            rec["roofline_timing"] = "CUDA events around every launch of eager forwards on the launching stream (serialised: no PDL overlap), mean of 5"
        if want_parity and env.world == 1:
            best_cpu_threads(sd, params, vocab)
            Bs = min(B, 8)
            runs = [cpu_oracle_pass(sd, params, mel_h[:Bs].clone(), len_h[:Bs].clone(), lambda ol: (y[:Bs], y_len[:Bs])) for _ in range(2)]
            sec = min(r[0] for r in runs)
            ref_logits, ref_loss = runs[-1][1], runs[-1][3]
            lg, ol, _ = model.forward_mel(mel_d[:Bs].contiguous(), len_d[:Bs].contiguous())
            gl = float(ctc_loss(lg, ol, y_d[:Bs].contiguous(), yl_d[:Bs].contiguous())[0])
            d = (lg.cpu().double() - ref_logits.double())
            rec["parity_vs_oracle"] = {"logits_rel_l2": float(d.norm() / ref_logits.double().norm()),
                                       "logits_max_abs_over_absmax": float(d.abs().max() / ref_logits.abs().max()),
                                       "ctc_loss_rel": abs(gl - ref_loss) / abs(ref_loss), "sample": f"first {Bs} utterances", "gate": 1e-3}
            rec["cpu_baseline"] = {"value": Bs * T / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"best of 2 passes of the first {Bs} utterances of the same batch (fwd + fc + CTC), fp32 torch CPU"}
        del model
        return rec
    finally:
        torch.set_grad_enabled(was)


# ---- training step ------------------------------------------------------------------------------------------------------
def count_launches(fn):
    """Kernel launches of one eager call of fn(): (ours, library) -- `ours` are the kernels of libeffconf_b200.so (namespace ec::),
    `library` whatever PyTorch / NCCL launched next to them (fills, index bookkeeping, collectives).  Memcpy / memset nodes excluded."""
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        ours = lib_k = 0
        names = {}
        KERNEL_US.clear()
        for e in prof.events():
            if e.device_type != torch.autograd.DeviceType.CUDA:
                continue
            n = e.name
            if n.lower().startswith(("memcpy", "memset")):
                continue
            if "ec::" in n:
                ours += 1
                key = n.split("ec::", 1)[1].replace("<unnamed>::", "").replace("(anonymous namespace)::", "").split("<")[0].split("(")[0]
                ent = KERNEL_US.setdefault(key, [0, 0.0])
                ent[0] += 1; ent[1] += float(e.time_range.elapsed_us())
            else:
                lib_k += 1; names[n.split("<")[0][:60]] = names.get(n.split("<")[0][:60], 0) + 1
        return ours, lib_k, names
    except Exception as ex:                                          # profiler unavailable
        return None, None, {"error": repr(ex)}


KERNEL_US = {}       # kernel name -> [launches, device microseconds] of the last count_launches() pass (CUPTI kernel records)


def make_train_step(env, config, precision, pdrop, graph, sync_bn=True, accumulated_steps=1, data_parallel=True):
    from efficientconformer_b200.trainer import CTCTrainStep
    model, sd, params, vocab = build_ctc_model(config, precision, env.dev, train=True, pdrop=pdrop)
    tp = dict(TRAINING_PARAMS); tp["accumulated_steps"] = accumulated_steps
    step = CTCTrainStep(model, tp, precision=precision, use_cuda_graph=graph, sync_bn=sync_bn, dropout_seed=1234, data_parallel=data_parallel)
    return step, params, vocab


def train_record(env, config, precision, B, T, steps, warmup, pdrop, graph=True, sync_bn=True, want_e2e=True):
    """(record, step object, host batch) of the CTC training step at B x 80 x T per GPU, full-length utterances."""
    dev = env.dev
    step, params, vocab = make_train_step(env, config, precision, pdrop, graph, sync_bn)
    mel_h = synthetic_mel(B, T, seed=1 + env.rank).pin_memory()
    t_out = out_frames(params, T)
    y, y_len = synthetic_targets(torch.full((B,), t_out), vocab, seed=4 + env.rank)
    y_h, yl_h = y.pin_memory(), y_len.pin_memory()
    mel_d, y_d, yl_d = mel_h.to(dev), y_h.to(dev), yl_h.to(dev)
    graph_note = "cuda graph replay" if graph else "eager launches"
    losses = []

    def one():
        losses.append(step.step(mel_d, None, y_d, yl_d).clone())
    try:
        one(); torch.cuda.synchronize()
    except Exception as ex:                              # e.g. a collective that cannot be captured: measure the eager step and say so
        if not graph:
            raise
        graph_note = f"eager launches (graph capture failed: {type(ex).__name__})"
        step, params, vocab = make_train_step(env, config, precision, pdrop, False, sync_bn)
    total_ms, step_ms = timed_steps(env, one, steps, warmup)
    frames = env.world * B * T * steps
    fpf = flops_per_frame(params, vocab, T, B)
    pk = peaks()
    ms = total_ms / steps
    rec = {"config": config, "precision": precision, "batch_per_gpu": B, "frames": T, "steps": steps, "launch": graph_note,
           "value": frames / (total_ms / 1e3), "unit": UNIT, "ms_per_step": ms,
           "step_ms_min_med_max": [round(min(step_ms), 4), round(statistics.median(step_ms), 4), round(max(step_ms), 4)],
           "whole_step": {"algorithmic_tflop": round(3.0 * fpf * B * T / 1e12, 4), "achieved_tflops": round(3.0 * fpf * B * T / ms / 1e9, 2),
                          "frac_of_bf16_peak": round(3.0 * fpf * B * T / ms / 1e9 / pk["bf16_tflops"], 4),
                          "note": "3 x forward FLOPs of SURVEY.md 8(d) per mel frame"},
           "loss_first_last": [float(losses[-steps]), float(losses[-1])]}
    if want_e2e:
        sec, e2e_losses = timed_e2e(env, (mel_h, y_h, yl_h), lambda m, yy, yl: step.step(m, None, yy, yl), steps)
        rec["e2e"] = {"value": frames / sec, "unit": UNIT, "h2d_bytes_per_step": mel_h.numel() * 4 + y_h.numel() * 8 + yl_h.numel() * 8,
                      "d2h_bytes_per_step": 4, "ms_per_step": 1e3 * sec / steps, "ms_per_step_repetitions": list(E2E_REPS),
                      "timing": "wall clock around K CTCTrainStep.step calls with HOST batches (pinned mel / targets -> H2D on a copy stream, 2-deep "
                                "pipeline, loss read back every step), synchronised both sides; fastest of the listed repetitions of the K-step region"}
        rec["e2e_loss_last"] = e2e_losses[-1]
    rec["lr_after"] = step.lr(); rec["optimizer_steps"] = step.steps_done()
    rec["sync_bn_exchange"] = type(step.reducer).__name__ if step.reducer is not None else None
    if hasattr(step.reducer, "error"):
        rec["sync_bn_exchange_timeouts"] = step.reducer.error()
    return rec, step, (mel_d, y_d, yl_d, mel_h, y, y_len)


def train_operator_profile(env, config, precision, pdrop, batch, reps=3):
    """Operator table of eager steps (CUDA events around every operator entry point on the launching stream) + launch counts."""
    from efficientconformer_b200 import _lib
    mel_d, y_d, yl_d = batch[:3]
    estep, params, vocab = make_train_step(env, config, precision, pdrop, graph=False, data_parallel=False)
    for _ in range(2):
        estep.step(mel_d, None, y_d, yl_d)
    torch.cuda.synchronize()
    with _lib.OpProfile() as prof:
        for _ in range(reps):
            estep.step(mel_d, None, y_d, yl_d)
        summ = prof.summary()
    ops = {k: {"calls": v["calls"] // reps, "ms": v["ms"] / reps, "flops": v["flops"] / reps} for k, v in summ.items()}
    ours_k, lib_k, lib_names = count_launches(lambda: estep.step(mel_d, None, y_d, yl_d))
    del estep
    return ops, ours_k, lib_k, lib_names


def ragged_batches(params, vocab, B, n_batches, seed=2, t_min=200, t_max=1600):
    """LibriSpeech-shaped batches: per-utterance T ~ U{200..1600} mel frames, sorted descending and zero-padded to
    the batch maximum exactly as the reference's collate_fn_pad does (utils/preprocessing.py:33-38); y_len = floor(out_len / 3)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_batches):
        lens = torch.randint(t_min, t_max + 1, (B,), generator=g).sort(descending=True).values
        Tm = int(lens[0])
        mel = synthetic_mel(B, Tm, seed=100 + i)
        for b in range(B):
            mel[b, :, int(lens[b]):] = 0.0
        ol = lens.clone()
        for _ in range(params["subsampling_layers"]):
            ol = (ol - 1) // 2 + 1
        for s in resolve_blocks(params):
            if s.conv_stride > 1:
                ol = (ol - 1) // s.conv_stride + 1
        y_len = (ol // 3).clamp_min(1)
        U = int(y_len.max())
        y = torch.randint(1, vocab, (B, U), generator=g)
        out.append((mel, lens, y, y_len))
    return out


def ragged_record(env, config, precision, B, steps, pdrop, accumulated_steps=2):
    """configs[1] as literally worded: ragged LibriSpeech-shaped batches, accumulated_steps = 2 (the shipped config), graph cache by shape."""
    dev = env.dev
    step, params, vocab = make_train_step(env, config, precision, pdrop, True, accumulated_steps=accumulated_steps)
    batches = [tuple(t.to(dev) for t in b) for b in ragged_batches(params, vocab, B, 4)]
    real = sum(int(b[1].sum()) for b in batches)
    padded = sum(b[0].shape[0] * b[0].shape[2] for b in batches)
    reps = max(1, steps // len(batches))

    def epoch():
        for mel, lens, y, yl in batches:
            step.step(mel, lens, y, yl)
    for _ in range(2):                                   # captures one graph per (shape, micro-step role)
        epoch()
    env.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        epoch()
    b.record()
    env.barrier()
    ms = env.max_over_ranks(a.elapsed_time(b))
    n = reps * len(batches)
    rec = {"workload": f"{config} CTC training, LibriSpeech-shaped batches: B={B}, T ~ U{{200..1600}} sorted / zero-padded to the batch maximum "
                       f"(collate_fn_pad), ragged x_len, accumulated_steps={accumulated_steps}, {len(batches)} distinct batches x {reps} passes",
           "precision": precision, "ms_per_micro_batch": ms / n, "real_frames_per_s": env.world * real * reps / (ms / 1e3),
           "padded_frames_per_s": env.world * padded * reps / (ms / 1e3), "captured_graphs": len(step._graphs),
           "batch_max_frames": [b[0].shape[2] for b in batches], "optimizer_steps": step.steps_done(), "loss_last": float(step.loss)}
    step.close()
    return rec


# ---- the unmodified reference modules run eagerly on this GPU ("the real bar") ------------------------------------
def torch_eager_reference(env, config, B, T, steps=5):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import stage_reference as SR
    if SR.import_reference() is None:
        return {"unavailable": "baseline/_ref is not staged (python tools/stage_reference.py where /root/reference exists)"}
    import contextlib
    import io
    import functions
    cfg = json.load(open(os.path.join(SR.DST, "configs", f"{config}.json")))
    cfg["encoder_params"]["spec_augment"] = False            # host SpecAugment is outside the timed path in both arms
    with contextlib.redirect_stdout(io.StringIO()):
        model = functions.create_model(cfg)
    params, vocab = SHIPPED_ENCODER_PARAMS[config]
    sd = seeded_state_dict(params, vocab, seed=0, prefix_encoder="encoder.")
    model.load_state_dict(sd, strict=False)
    model = model.to(env.dev)

    class MelIn(torch.nn.Module):                             # the benchmark injects mel frames: skip STFT -> mel of the reference front end
        def forward(self, x, x_len):
            return x, x_len
    model.encoder.preprocessing = MelIn()
    mel = synthetic_mel(B, T, seed=1).to(env.dev)
    mel_len = torch.full((B,), T, dtype=torch.int64, device=env.dev)
    t_out = out_frames(params, T)
    y, y_len = synthetic_targets(torch.full((B,), t_out), vocab, seed=4)
    batch = [mel, y.to(env.dev), mel_len, y_len.to(env.dev)]
    out = {"what": "the reference's own ConformerEncoder / ModelCTC / LossCTC modules (unmodified, staged copy) run eagerly by PyTorch on this "
                   f"B200, {config}, B={B} x 80 x {T}, mel-level input, SpecAugment off; cuDNN / cuBLAS kernels, no CUDA graph",
           "torch": torch.__version__}

    def timeit(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for name, amp in (("fp32", False), ("autocast_fp16", True)):
        model.eval()

        def fwd():
            with torch.no_grad(), torch.autocast("cuda", enabled=amp):
                pred = model.forward(batch)
                return model.criterion(batch, pred)
        try:
            ms = timeit(fwd, steps)
            out[f"forward_{name}"] = {"ms_per_step": ms, "value": B * T / (ms / 1e3), "unit": UNIT}
        except Exception as ex:
            out[f"forward_{name}"] = {"error": repr(ex)[:200]}
        model.train()
        scaler = torch.amp.GradScaler("cuda", enabled=amp)

        def train():
            with torch.autocast("cuda", enabled=amp):
                pred = model.forward(batch)
                loss = model.criterion(batch, pred)
            scaler.scale(loss).backward()
            scaler.step(model.optimizer); scaler.update(); model.optimizer.zero_grad(); model.scheduler.step()
        try:
            ms = timeit(train, steps)
            out[f"train_{name}"] = {"ms_per_step": ms, "value": B * T / (ms / 1e3), "unit": UNIT}
        except Exception as ex:
            out[f"train_{name}"] = {"error": repr(ex)[:200]}
    del model
    torch.cuda.empty_cache()
    return out


def _finish(env, step):
    """Leave a multi-rank run in order: drop the captured graphs (they hold NCCL kernels), synchronise, destroy the process group.  The
    communicator teardown runs under a 20 s guard: should it stall (seen in round 1 while graphs still referenced NCCL kernels) the
    result is already printed and the process leaves without it."""
    if env.dist is None:
        return
    if step is not None:
        step.close()
    import gc
    import threading
    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush(); sys.stderr.flush()
    done = threading.Event()

    def teardown():
        try:
            env.dist.destroy_process_group()
        finally:
            done.set()
    threading.Thread(target=teardown, daemon=True).start()
    if not done.wait(20.0):
        sys.stderr.write(f"[bench rank {env.rank}] process-group teardown did not finish within 20 s; leaving without it\n")
        sys.stderr.flush()
        os._exit(0)


def guarded(name, out, fn):
    """An extra that fails must not take the headline with it."""
    try:
        t0 = time.perf_counter()
        out[name] = fn()
        if isinstance(out[name], dict):
            out[name]["bench_seconds"] = round(time.perf_counter() - t0, 1)
    except Exception as ex:
        import traceback
        out[name] = {"error": f"{type(ex).__name__}: {ex}"[:300], "trace": traceback.format_exc()[-600:]}
    torch.cuda.empty_cache()


def run_ours(args, rank, world, local_rank):
    env = Env(rank, world, local_rank)
    pk = peaks()
    cfg, pr, B, T = args.config, args.precision, args.batch, args.frames
    sampler = ClockSampler(local_rank)
    out, step = {}, None
    if args.mode == "train":
        env.barrier()
        if rank == 0:
            sampler.start()
        rec, step, batch = train_record(env, cfg, pr, B, T, args.steps, args.warmup, args.pdrop, graph=not args.no_graph, sync_bn=not args.no_sync_bn)
        clocks = sampler.stop() if rank == 0 else None
        _log(rank, f"train headline {rec['ms_per_step']:.3f} ms")
        out = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": pr,
               "data": "synthetic", "config": workload_config(args, "train"), "launch": rec["launch"], "sync_bn": world > 1 and not args.no_sync_bn,
               "e2e": rec["e2e"], "clocks": clocks}
        for k in ("step_ms_min_med_max", "loss_first_last", "e2e_loss_last", "lr_after", "optimizer_steps", "whole_step", "sync_bn_exchange",
                  "sync_bn_exchange_timeouts"):
            if k in rec:
                out[k] = rec[k]
        if world > 1:
            # exposed communication: the same step on the same batches with no collective at all (local BatchNorm statistics, no gradient
            # all-reduce), every rank for itself; the difference to the data-parallel step is what SyncBatchNorm + the bucket cost
            lstep, _, _ = make_train_step(env, cfg, pr, args.pdrop, graph=not args.no_graph, data_parallel=False)
            mel_d, y_d, yl_d = batch[:3]
            local_ms, _ = timed_steps(env, lambda: lstep.step(mel_d, None, y_d, yl_d), max(10, args.steps // 2), 3)
            local_ms /= max(10, args.steps // 2)
            lstep.close(); del lstep
            out["communication"] = {"ms_per_step_no_collectives": local_ms, "exposed_ms_per_step": rec["ms_per_step"] - local_ms,
                                    "exchanges_per_step": ("16 + 16 SyncBatchNorm exchanges (forward statistics, backward sums) as "
                                                           + ("peer-memory kernels over NVLink (csrc/p2p_exchange.cu)" if rec.get("sync_bn_exchange") == "P2PSyncBatchNormReducer"
                                                              else "NCCL all_gather / all_reduce")
                                                           + (f" + {len(step._buckets) + 1} NCCL all_reduces over the flat 53 MB fp32 gradient arena (buckets: last third of the "
                                                              "blocks + head, middle third, rest; the first two leave on a communication stream while the "
                                                              "backward of the earlier blocks runs), all inside the captured graph" if getattr(step, "_buckets", None)
                                                              else " + 1 x NCCL all_reduce of the flat 53 MB fp32 gradient bucket, all inside the captured graph"))
                                                          if not args.no_sync_bn else "1 x all_reduce of the flat fp32 gradient bucket"}
        if rank == 0:
            ops, ours_k, lib_k, lib_names = train_operator_profile(env, cfg, pr, args.pdrop, batch)
            tot_ms = sum(v["ms"] for v in ops.values())
            ops_sorted = sorted(ops.items(), key=lambda kv: -kv[1]["ms"])
            out["operators"] = [{"op": k, "calls": v["calls"], "ms": round(v["ms"], 4),
                                 "tflops": round(v["flops"] / v["ms"] / 1e9, 2) if v["flops"] else None} for k, v in ops_sorted]
            desc = {"ec_op_wgrad_bias": "ec_op_wgrad_bias = wgrad_tc_kernel (tcgen05, MN-major operands, split-M, bias gradient from a ones operand) + "
                                        "fixed-order reduce: every weight gradient of one step (runs on a forked side stream in the timed step)",
                    "ec_op_gemm_ex": "ec_op_gemm_ex = gemm_tc_kernel (tcgen05 + TMA + TMEM): every forward Linear / pointwise conv and every data-gradient "
                                     "GEMM of one step"}

            def roof(k, v):
                a = v["flops"] / v["ms"] / 1e9
                return {"kernel": desc.get(k, k), "bound": "tensor", "achieved": round(a, 2), "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": round(a / pk["bf16_tflops"], 4), "mode_ceiling": tensor_peak(pr, pk),
                        "traffic": ncu_traffic(pr, "wgrad_tc_kernel" if "wgrad" in k else "gemm_tc_kernel"),
                        "peak_source": pk["source"] + " bf16 cuBLAS burst", "launches_per_step": v["calls"],
                        "avg_launch_us": round(1e3 * v["ms"] / max(v["calls"], 1), 2), "share_of_step": round(v["ms"] / tot_ms, 3),
                        "algorithmic_flops_per_step": v["flops"],
                        "timing": "CUDA events around every operator call of eager steps on the launching stream (serialised), mean of 3 steps"}
            tensor_ops = [(k, v) for k, v in ops_sorted if v["flops"] > 0]
            if tensor_ops:
                out["roofline"] = roof(*tensor_ops[0])
                out["rooflines_other"] = [roof(k2, v2) for k2, v2 in tensor_ops[1:]]
            # the same operators by KERNEL duration (CUPTI records of one eager step, no host launch gaps): cross-check of the event-based
            # figures above, which include the gaps between an operator's launch and the previous kernel's end in an eager step
            if KERNEL_US:
                gemm_flops = sum(v["flops"] for k, v in ops.items() if k in ("ec_op_gemm", "ec_op_gemm_ex", "ec_op_gemm_train", "ec_op_gemm_ln_train"))
                wg_flops = sum(v["flops"] for k, v in ops.items() if k.startswith("ec_op_wgrad"))
                kt = {}
                for kern, fl in (("gemm_tc_kernel", gemm_flops), ("wgrad_tc_kernel", wg_flops)):
                    if kern in KERNEL_US and KERNEL_US[kern][1] > 0:
                        n_, us_ = KERNEL_US[kern]
                        kt[kern] = {"launches": n_, "kernel_ms": round(us_ / 1e3, 4), "avg_us": round(us_ / n_, 2),
                                    "achieved_tflops": round(fl / us_ / 1e6, 2), "frac_of_bf16_peak": round(fl / us_ / 1e6 / pk["bf16_tflops"], 4)}
                top = sorted(KERNEL_US.items(), key=lambda kv: -kv[1][1])[:12]
                out["kernel_time"] = {"tensor_kernels": kt, "top_kernels_ms": {k: round(v[1] / 1e3, 3) for k, v in top},
                                      "total_kernel_ms": round(sum(v[1] for v in KERNEL_US.values()) / 1e3, 3),
                                      "source": "torch.profiler (CUPTI) kernel records of ONE eager step of this run, warm caches; diagnostic only"}
            out["eager_profiled_step_ms"] = round(tot_ms, 3)
            launches = ours_k if ours_k is not None else 1267
            out["gpu_launches"] = launches * args.steps
            out["launches_per_step"] = launches
            out["library_launches_per_step"] = {"count": lib_k, "kernels": lib_names}
            if world > 1:
                out["launches_note"] = "counted with the CUDA profiler on an eager step without collectives (rank 0); the NCCL kernels come on top"
    else:
        env.barrier()
        if rank == 0:
            sampler.start()
        rec = forward_record(env, cfg, pr, B, T, args.steps, args.warmup, want_parity=not args.no_cpu_baseline)
        clocks = sampler.stop() if rank == 0 else None
        out = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": pr,
               "data": "synthetic", "config": workload_config(args, "forward"), "e2e": rec["e2e"], "clocks": clocks,
               "gpu_launches": rec["launches_per_step"] * args.steps}
        roofs = rec.pop("rooflines", [])
        if roofs:
            dom = max(roofs, key=lambda r: r["share_of_forward"])
            out["roofline"] = dom
            out["rooflines_other"] = [r for r in roofs if r is not dom]
        for k, v in rec.items():
            out.setdefault(k, v)

    # ---- sub-records -------------------------------------------------------------------------------------------------------
    if not args.no_extras:
        small = DEFAULT_CONFIG
        if args.mode == "train":
            guarded("forward", out, lambda: forward_record(env, cfg, pr, B, T, 50, 3, want_parity=(world == 1 and not args.no_cpu_baseline)))
            _log(rank, "forward done")

        def sweep():
            rows = []
            for t in (500, 1000, 2000):
                row = {"frames": t}
                f = forward_record(env, cfg, pr, B, t, 20, 3, want_e2e=False, want_kernels=False)
                row["forward"] = {k: f[k] for k in ("value", "ms_per_step", "whole_forward_tflops", "algorithmic_mflop_per_frame")}
                r, st, _ = train_record(env, cfg, pr, B, t, 10, 3, args.pdrop, sync_bn=not args.no_sync_bn, want_e2e=False)
                st.close(); del st
                row["train"] = {k: r[k] for k in ("value", "ms_per_step")}
                row["train"]["achieved_tflops"] = r["whole_step"]["achieved_tflops"]
                rows.append(row)
                torch.cuda.empty_cache()
            return {"batch_per_gpu": B, "precision": pr, "n_gpus": world, "rows": rows}
        guarded("sweep", out, sweep)
        _log(rank, "sweep done")

        def modes():
            res = {}
            for m in ("bf16", "tf32", "bf16x2"):
                if m == pr:
                    continue
                f = forward_record(env, cfg, m, B, T, 20, 3, want_e2e=False, want_kernels=False)
                r, st, _ = train_record(env, cfg, m, B, T, 10, 3, args.pdrop, sync_bn=not args.no_sync_bn, want_e2e=False)
                st.close(); del st
                res[m] = {"forward": {k: f[k] for k in ("value", "ms_per_step")}, "train": {k: r[k] for k in ("value", "ms_per_step")}}
                torch.cuda.empty_cache()
            res["note"] = ("bf16 = plain bf16 operands (what torch.autocast selects; 6e-3 logits error, outside the 1e-3 gate), tf32 = TF32 operands "
                           "(7e-4), bf16x2 = split bf16 hi/lo operands (8e-5; the default)")
            return res
        guarded("modes", out, modes)
        _log(rank, "modes done")
        guarded("ragged", out, lambda: ragged_record(env, small, pr, B, 8, args.pdrop))
        _log(rank, "ragged done")
        if world == 1 or args.extras_all_ranks:
            def configs():
                res = {}
                f = forward_record(env, "EfficientConformerCTCLarge", pr, 32, 1000, 10, 3, want_e2e=False, want_kernels=False)
                res["EfficientConformerCTCLarge_forward_B32_T1000"] = {k: f[k] for k in ("value", "ms_per_step", "whole_forward_tflops",
                                                                                           "whole_forward_frac_of_bf16_peak", "algorithmic_mflop_per_frame", "precision")}
                rows = []
                for t in (500, 1000, 2000, 4000):
                    f = forward_record(env, "ConformerCTCLarge", pr, 8, t, 10, 3, want_e2e=False, want_kernels=False)
                    rows.append({k: f[k] for k in ("frames", "value", "ms_per_step", "whole_forward_tflops", "whole_forward_frac_of_bf16_peak",
                                                   "algorithmic_mflop_per_frame")})
                    torch.cuda.empty_cache()
                res["ConformerCTCLarge_forward_B8_sweep"] = rows
                res["precision"] = pr
                # the same kernels at the shapes where they are not latency bound: plain bf16 operands (fused feed-forward cluster kernel),
                # D = 512 / 720 -- what the tcgen05 path reaches when a tile has work (per-kernel rooflines of the T = 4000 run included)
                big = {}
                f = forward_record(env, "EfficientConformerCTCLarge", "bf16", 32, 1000, 10, 3, want_e2e=False, want_kernels=False)
                big["EfficientConformerCTCLarge_forward_B32_T1000"] = {k: f[k] for k in ("value", "ms_per_step", "whole_forward_tflops", "whole_forward_frac_of_bf16_peak")}
                f = forward_record(env, "ConformerCTCLarge", "bf16", 8, 4000, 10, 3, want_e2e=False, want_kernels=True)
                big["ConformerCTCLarge_forward_B8_T4000"] = {k: f[k] for k in ("value", "ms_per_step", "whole_forward_tflops", "whole_forward_frac_of_bf16_peak", "rooflines")}
                res["bf16_mode"] = big
                return res
            guarded("configs", out, configs)
            _log(rank, "configs done")
        if world == 1:
            def transducer_joint():
                # configs[3] (EfficientConformerTransducerMedium): the joint network + RNN-T loss forward on the lattice of a
                # B = 16 x 1000-frame batch (T' = 125 encoder frames, U = 50 labels; encoder 360 / decoder 640 / joint 640 / vocab 1000)
                from efficientconformer_b200.transducer import JointNetwork, rnnt_loss
                Bj, Tj, Uj, Denc, Ddec, J, Vj = 16, 125, 50, 360, 640, 640, 1000
                g = torch.Generator().manual_seed(3)
                jn = JointNetwork(Denc, Ddec, Vj, {"joint_mode": "sum", "dim_model": J, "act": "tanh"}, precision=pr).to(env.dev).eval()
                f, gd = torch.randn(Bj, Tj, Denc, generator=g).to(env.dev), torch.randn(Bj, Uj + 1, Ddec, generator=g).to(env.dev)
                y = torch.randint(1, Vj, (Bj, Uj), generator=g).to(env.dev)
                fl, yl = torch.full((Bj,), Tj, device=env.dev), torch.full((Bj,), Uj, device=env.dev)

                def one():
                    with torch.no_grad():
                        return rnnt_loss(jn(f, gd), y, fl, yl)[0]
                ms, _ = timed_steps(env, one, 10, 3)
                ms /= 10
                from efficientconformer_b200.transducer import LossRNNT
                crit = LossRNNT()
                fg, gg = f.clone().requires_grad_(True), gd.clone().requires_grad_(True)

                def train():
                    jn.zero_grad(set_to_none=True); fg.grad = None; gg.grad = None
                    loss = crit((None, y, None, yl), (jn(fg, gg), fl, None))
                    loss.backward()
                    return loss
                jn.train()
                ms_train, _ = timed_steps(env, train, 10, 3)
                ms_train /= 10
                flops = 2.0 * Bj * Tj * (Uj + 1) * J * Vj
                return {"workload": "Transducer JointNetwork.forward + LossRNNT (and + loss.backward() through both autograd nodes), B=16, T'=125, U=50, "
                                    "encoder 360 / decoder 640 / joint 640 / vocab 1000; eager launches",
                        "precision": pr, "forward_ms": ms, "forward_backward_ms": ms_train, "lattice_nodes_per_s_forward": Bj * Tj * (Uj + 1) / (ms / 1e3),
                        "output_projection_tflops_forward": round(flops / ms / 1e9, 1), "loss": float(one())}
            guarded("transducer_joint", out, transducer_joint)
            guarded("torch_eager_b200", out, lambda: torch_eager_reference(env, small, B, T))
            _log(rank, "torch eager done")
    if env.dist is not None:
        env.dist.barrier()
        torch.cuda.synchronize()
    if rank != 0:
        _finish(env, step)
        return
    if world == 1 and not args.no_cpu_baseline and args.mode == "train":
        params, vocab = SHIPPED_ENCODER_PARAMS[cfg]
        sd2, leaf, opt = cpu_train_setup(params, vocab)
        best_cpu_threads(sd2, params, vocab)
        Bs = 4
        mel_h, y, y_len = batch[3], batch[4], batch[5]
        cm, cl = mel_h[:Bs].clone(), torch.full((Bs,), T, dtype=torch.int64)
        cpu_oracle_train_pass(leaf, opt, cm, cl, y[:Bs], y_len[:Bs], params)
        runs = [cpu_oracle_train_pass(leaf, opt, cm, cl, y[:Bs], y_len[:Bs], params)[0] for _ in range(3)]
        sec = statistics.median(runs)
        out["cpu_baseline"] = {"value": Bs * T / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"median of 3 training steps on the first {Bs} utterances of the same batch (oracle train-mode fwd + CTC + autograd "
                                         f"backward + torch.optim.Adam), fp32 torch CPU, {torch.get_num_threads()} threads"}
    print(json.dumps(out))
    sys.stdout.flush()
    _finish(env, step)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16x2", choices=["bf16x2", "bf16", "tf32"],
                    help="operand mode: bf16x2 = packed bf16 hi/lo pairs (default; meets the 1e-3 parity gate), bf16 = fast mode, tf32")
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(SHIPPED_ENCODER_PARAMS))
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip the forward / sweep / modes / ragged / configs / torch-eager sub-records")
    ap.add_argument("--extras-all-ranks", action="store_true", help="N > 1: also run the other-config forwards on every rank")
    ap.add_argument("--mode", default="train", choices=["train", "forward"])
    ap.add_argument("--pdrop", type=float, default=0.1, help="dropout probability of the training step (reference config: 0.1)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="--impl reference: seconds of CPU work the arm may spend")
    ap.add_argument("--watchdog", type=int, default=900, help="seconds after which a stuck run dumps its Python stacks to stderr and exits")
    ap.add_argument("--no-graph", action="store_true", help="training step: eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-sync-bn", action="store_true", help="N > 1: per-rank BatchNorm statistics (NOT the reference's SyncBatchNorm semantics)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = (40 if args.mode == "train" else 100) if args.impl == "ours" else 3
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    args.gpus = max(args.gpus, world) if args.impl == "ours" else args.gpus
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)      # a hung collective must not hang the caller: dump stacks, exit
    # keep stdout to the single JSON line whatever the libraries print (NCCL prints its version banner to stdout)
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a B200: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
