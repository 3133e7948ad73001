#!/usr/bin/env python
"""Benchmark of the hot path on synthetic 80-mel batches of EfficientConformerCTCSmall.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|tf32] [--mode train|forward]

--mode train (default; BASELINE.json configs[1]): one optimisation step = train-mode forward (batch-statistics BatchNorm, dropout
0.1) + CTC loss + backward + gradient all-reduce (N > 1) + Adam with the Transformer schedule, through
efficientconformer_b200.trainer.CTCTrainStep (CUDA-graph replay).  --mode forward: inference forward + fc + CTC loss (the round-1
forward numbers).  The remaining text describes the keys of the JSON line, which are the same in both modes.

Prints ONE JSON line (rank 0).  metric = mel frames / second (BASELINE.json), whole job over all N GPUs.
  value     inputs already resident in HBM; every step timed by its own CUDA-event pair on the launching stream, L2 flushed
            (256 MiB write) between steps, max over ranks
  e2e       the same step through the public API with HOST inputs: pinned mel -> H2D, ModelCTC.forward_mel, ctc_loss,
            loss.item() (D2H) inside the timed region
  roofline  dominant kernel (the tcgen05 GEMM, all its launches in one forward): algorithmic FLOPs / CUDA-event time
            from an eager profiled pass inside this run, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (port of the reference algorithm, pinned to the reference's outputs) on the host cores
--impl reference times that CPU oracle as the reference arm (the reference itself is Python and cannot travel to the GPU box).
Multi-GPU: one process per GPU (torchrun); utterances shard over ranks with no data-path collective (forward has no
exchange step); barrier + max-over-ranks timing over NCCL."""
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

from efficientconformer_b200.config import CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V  # noqa: E402
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel, synthetic_targets  # noqa: E402

METRIC = "encoder_mel_frames_per_sec"
UNIT = "frames/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(precision, kernel_prefix="gemm_tc_kernel", stem="r1_ncu_full_block0"):
    """Average DRAM bytes (read + write) per launch of one kernel from the committed `ncu --set full` capture of block 0
    (profiles/, produced by tools/summarize_ncu.py); None when no capture exists for this precision."""
    path = os.path.join(ROOT, "profiles", f"{stem}_{precision}_summary.json")
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


def cpu_oracle_pass(sd, mel, mel_len, y_fn):
    from oracle import conformer_oracle as O           # checker / CPU baseline only (never the product path)
    t0 = time.perf_counter()
    logits, out_len = O.model_ctc_forward_mel(sd, P, mel, mel_len)
    y, y_len = y_fn(out_len)
    loss, _ = O.ctc_loss(logits, out_len, y, y_len)
    return time.perf_counter() - t0, logits, out_len, float(loss)


def best_cpu_threads(sd):
    """The CPU path is many small ATen ops: beyond a few dozen threads the fork/join cost dominates (128 threads are >20x
    slower than 32 on the GPU box's host).  Give the CPU arm its best shot: time one small pass per candidate count."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    mel = synthetic_mel(4, 500, seed=9)
    ln = torch.full((4,), 500, dtype=torch.int64)
    yf = lambda ol: synthetic_targets(ol, V, seed=4)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        cpu_oracle_pass(sd, mel, ln, yf)
        t = min(cpu_oracle_pass(sd, mel, ln, yf)[0] for _ in range(2))
        if t < best_t:
            best, best_t = c, t
        if t > 4 * best_t:
            break
    torch.set_num_threads(best)
    return best


def run_reference(args, rank, world):
    """Reference arm: the CPU implementation of the path (oracle port) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    torch.set_grad_enabled(False)
    sd = seeded_state_dict(P, V, seed=0, prefix_encoder="encoder.")
    cores = best_cpu_threads(sd)
    B = args.batch if (args.steps + args.warmup) <= 40 else max(2, args.batch // 4)
    mel = synthetic_mel(B, args.frames, seed=1)
    mel_len = torch.full((B,), args.frames, dtype=torch.int64)
    yf = lambda ol: synthetic_targets(ol, V, seed=4)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_oracle_pass(sd, mel, mel_len, yf)
    times = [cpu_oracle_pass(sd, mel, mel_len, yf)[0] for _ in range(args.steps)]
    total = sum(times)
    value = B * args.frames * args.steps / total
    sample = (f"{args.steps} passes of B={B} x 80 x {args.frames} (fwd + fc + CTC loss), fp32, torch CPU {torch.get_num_threads()} threads "
              f"(best of 8/16/32/64/all on a {os.cpu_count()}-core host)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, B, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, batch_per_gpu, where):
    return {"workload": f"EfficientConformerCTCSmall encoder fwd + fc + CTC loss, batch {batch_per_gpu}/GPU x 80-mel x {args.frames} frames "
                        f"(BASELINE.json north_star target shape; configs[1] without backward, see DESIGN.md)",
            "global_batch": batch_per_gpu * (args.gpus if where != "cpu" else 1), "frames": args.frames, "n_mels": 80,
            "weights": "seeded random init", "l2": "256 MiB write between timed steps (L2 flush)", "parallelism": f"dp{args.gpus} (utterance shards, no collective)"}


def run_ours(args, rank, world, local_rank):
    from efficientconformer_b200 import ModelCTC, _lib
    from efficientconformer_b200.model_ctc import ctc_loss
    import ctypes as C
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"            # keep stdout to the single JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    B, T = args.batch, args.frames
    sd = seeded_state_dict(P, V, seed=0, prefix_encoder="encoder.")
    model = ModelCTC(P, {"vocab_size": V}, precision=args.precision)
    model.load_state_dict(sd, strict=False)
    model = model.to(dev).eval()
    mel_h = synthetic_mel(B, T, seed=1 + rank).pin_memory()
    len_h = torch.full((B,), T, dtype=torch.int64).pin_memory()
    mel_d, len_d = mel_h.to(dev), len_h.to(dev)
    t_out = (((T - 1) // 2 + 1 - 1) // 2 + 1 - 1) // 2 + 1
    y, y_len = synthetic_targets(torch.full((B,), t_out), V, seed=4)
    y_d, yl_d = y.to(dev), y_len.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_resident():
        logits, out_len, _ = model.forward_mel(mel_d, len_d)
        return ctc_loss(logits, out_len, y_d, yl_d)[0]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident(); flush.zero_()
    prec = _lib.PRECISIONS[args.precision]
    eng = model.encoder._engines[prec][0]
    launches_per_step = _lib.lib().ec_engine_last_launches(eng) + 4          # + CTC: i64->i32, lse/argmax, alpha, mean
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.zero_()
        a.record(); loss = step_resident(); b.record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    # ---- end to end through the public API with host buffers ----
    # Every step copies its own inputs from pinned host memory (copy stream), runs ModelCTC.forward_mel + ctc_loss and reads
    # the loss back to the host.  The copy of step k+1 and the read-back of step k-1 overlap the compute of step k (a 2-deep
    # input pipeline, what a DataLoader with pinned memory does); all K copies, K computes and K read-backs are inside the
    # timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()

    def e2e_loop(n_steps):
        cur = torch.cuda.current_stream()
        bufs, ready, done = [None, None], [torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()]
        losses = []

        def stage(k):
            with torch.cuda.stream(copy_stream):
                if bufs[k & 1] is not None:
                    copy_stream.wait_event(done[k & 1])                   # the compute that used this buffer has finished
                bufs[k & 1] = (mel_h.to(dev, non_blocking=True), len_h.to(dev, non_blocking=True))
                ready[k & 1].record(copy_stream)
        stage(0)
        for k in range(n_steps):
            if k + 1 < n_steps:
                stage(k + 1)
            cur.wait_event(ready[k & 1])
            m, l = bufs[k & 1]
            logits, out_len, _ = model.forward_mel(m, l)
            loss = ctc_loss(logits, out_len, y_d, yl_d)[0]
            loss_host[k & 1].copy_(loss, non_blocking=True)               # D2H of the step's result
            done[k & 1].record(cur)
            if k >= 1:
                done[(k - 1) & 1].synchronize()
                losses.append(float(loss_host[(k - 1) & 1]))
        done[(n_steps - 1) & 1].synchronize()
        losses.append(float(loss_host[(n_steps - 1) & 1]))
        return losses

    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    e2e_losses = e2e_loop(args.steps)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    barrier()
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    e2e_s = float(e2e_s)

    # ---- per-kernel profile of one eager forward (CUDA events around every launch, same stream) ----
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
            name = L.ec_profile_category_name(i).decode()
            a = acc.setdefault(name, [0.0, 0.0, 0.0, 0])
            a[0] += ms[i] / reps; a[1] = fl[i]; a[2] = by[i]; a[3] = cnt[i]
    L.ec_engine_set_profiling(eng, 0)
    model.encoder.use_cuda_graph = True
    pk = peaks()
    tensor_peak = pk["bf16_tflops"] * (0.5 if args.precision == "tf32" else 1.0)     # kind::tf32 runs at half the bf16 rate
    kernels = []
    # kernel classes = device functions: every gemm_* category is one gemm_tc_kernel, the fused FFN and attention are their own
    classes = {"gemm_tc_kernel": [0.0, 0.0, 0], "ffn_fused_kernel": [0.0, 0.0, 0], "relpos_attn_kernel": [0.0, 0.0, 0]}
    for name, (ms_, fl_, by_, n_) in acc.items():
        if n_ == 0:
            continue
        ent = {"kernel": name, "launches": n_, "ms": round(ms_, 4), "tflops": round(fl_ / ms_ / 1e9, 2) if ms_ > 0 else None,
               "gbs": round(by_ / ms_ / 1e6, 1) if ms_ > 0 else None}
        kernels.append(ent)
        cls = "gemm_tc_kernel" if name.startswith("gemm_") else "ffn_fused_kernel" if name == "ffn_fused" else \
            "relpos_attn_kernel" if name == "relpos_attention" else None
        if cls:
            c = classes[cls]; c[0] += ms_; c[1] += fl_; c[2] += n_
    fwd_profiled_ms = sum(k["ms"] for k in kernels)
    desc = {"gemm_tc_kernel": "gemm_tc_kernel (tcgen05 + TMA, every Linear / pointwise conv launch of one forward)",
            "ffn_fused_kernel": "ffn_fused_kernel (tcgen05 + TMA cluster kernel: W1 -> Swish -> W2 -> residual -> LayerNorm, 30 launches per forward)",
            "relpos_attn_kernel": "relpos_attn kernels (mma.sync bf16/tf32, TMA-staged)"}

    def tensor_roofline(cls):
        ms_, fl_, n_ = classes[cls]
        ach = fl_ / ms_ / 1e9 if ms_ > 0 else 0.0
        return {"kernel": desc[cls], "bound": "tensor", "achieved": round(ach, 2), "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": round(ach / tensor_peak, 4), "traffic": ncu_traffic(args.precision, cls.replace("relpos_attn_kernel", "relpos_attn")),
                "peak_source": pk["source"] + (" bf16 cuBLAS burst" if args.precision == "bf16" else " bf16 cuBLAS burst / 2 (tf32 operands)"),
                "launches_per_forward": n_, "avg_launch_us": round(1e3 * ms_ / max(n_, 1), 2),
                "share_of_forward": round(ms_ / fwd_profiled_ms, 3) if fwd_profiled_ms else None,
                "timing": "CUDA events around every launch of an eager forward on the launching stream (serialised: no PDL overlap)"}
    dominant = max(classes, key=lambda c: classes[c][0])
    roofline = tensor_roofline(dominant)
    extra_rooflines = [tensor_roofline(c) for c in classes if c != dominant and classes[c][2] > 0]
    dw = acc.get("dwconv_bn_swish")
    if dw and dw[0] > 0:
        gbs = dw[2] / dw[0] / 1e6
        extra_rooflines.append({"kernel": "dwconv_bn_swish", "bound": "hbm", "achieved": round(gbs, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                                "frac": round(gbs / pk["hbm_gbs"], 4), "note": "working set is L2-resident at this shape (see DESIGN.md)"})

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    frames_total = world * B * T * args.steps
    value = frames_total / (total_ms / 1e3)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic", "config": workload_config(args, B, "gpu"),
        "e2e": {"value": frames_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": mel_h.numel() * 4 + len_h.numel() * 8, "d2h_bytes_per_step": 4,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "timing": "wall clock around K API calls, synchronised both sides; 2-deep pinned-memory input pipeline on a copy stream"},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "clocks": clocks, "roofline": roofline, "rooflines_other": extra_rooflines, "kernels": kernels,
        "step_ms_min_med_max": [round(min(step_ms), 4), round(statistics.median(step_ms), 4), round(max(step_ms), 4)],
        "loss": float(loss),
    }
    if world == 1 and not args.no_cpu_baseline:
        best_cpu_threads(sd)
        Bs = 8
        yf = lambda ol: (y[:Bs], y_len[:Bs])
        cpu_oracle_pass(sd, mel_h[:Bs].clone(), len_h[:Bs].clone(), yf)
        runs = [cpu_oracle_pass(sd, mel_h[:Bs].clone(), len_h[:Bs].clone(), yf) for _ in range(3)]
        sec = statistics.median(r[0] for r in runs)
        ref_logits, ref_len, ref_loss = runs[-1][1], runs[-1][2], runs[-1][3]
        out["cpu_baseline"] = {"value": Bs * T / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"median of 3 passes of the first {Bs} utterances of the same batch (fwd + fc + CTC), fp32 torch CPU, {torch.get_num_threads()} threads"}
        lg, ol, _ = model.forward_mel(mel_d[:Bs].contiguous(), len_d[:Bs].contiguous())
        gl = float(ctc_loss(lg, ol, y_d[:Bs].contiguous(), yl_d[:Bs].contiguous())[0])
        d = (lg.cpu().double() - ref_logits.double())
        out["parity_vs_oracle"] = {"logits_rel_l2": float(d.norm() / ref_logits.double().norm()),
                                   "logits_max_abs_over_absmax": float(d.abs().max() / ref_logits.abs().max()),
                                   "ctc_loss_rel": abs(gl - ref_loss) / abs(ref_loss), "sample": f"first {Bs} utterances", "gate": 1e-3,
                                   "note": "tf32 operand mode meets the 1e-3 gate; bf16 mode is the fast mode (reference's own bf16 autocast deviates 1.1e-2)"}
    print(json.dumps(out))


# =====================================================================================================================
# training step (BASELINE.json configs[1]; reference models/model.py:239-259)
# =====================================================================================================================
TRAINING_PARAMS = dict(optimizer="Adam", beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-6, lr_schedule="Transformer", schedule_dim=240,
                       warmup_steps=10000, K=2)      # reference configs/EfficientConformerCTCSmall.json:53-69
TRAIN_FLOP_PER_FRAME = {500: 6.148e6, 1000: 6.616e6, 1600: 7.198e6, 2000: 7.587e6}   # forward; training = 3x


def train_workload_config(args, batch_per_gpu, world, where, extra=None):
    cfg = {"workload": f"EfficientConformerCTCSmall CTC training step (train-mode forward, dropout {args.pdrop}, CTC loss, backward, Adam + Transformer "
                       f"schedule), batch {batch_per_gpu}/GPU x 80-mel x {args.frames} frames, full-length utterances (BASELINE.json configs[1] at the "
                       f"north_star target shape)",
           "global_batch": batch_per_gpu * world, "frames": args.frames, "n_mels": 80, "weights": "seeded random init",
           "optimizer": "Adam + Transformer schedule applied after EVERY batch (the reference config accumulates 2 micro-batches per optimiser "
                        "step: this measures more optimiser work per frame, not less)",
           "l2": "256 MiB write between timed steps (L2 flush)" if where == "gpu" else "n/a (CPU)",
           "parallelism": f"dp{world}: utterances sharded over ranks; SyncBatchNorm statistics + one flat gradient bucket all-reduced over NCCL"
                          if world > 1 else "dp1"}
    if extra:
        cfg.update(extra)
    return cfg


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


def cpu_train_setup():
    sd = seeded_state_dict(P, V, seed=0, prefix_encoder="encoder.")
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    opt = torch.optim.Adam([v for v in leaf.values() if v.requires_grad], lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    return sd, leaf, opt


def run_reference_train(args, rank, world):
    """Reference arm for the training step: the CPU implementation (oracle port + torch autograd + torch.optim.Adam) on the host cores."""
    if rank != 0:
        return
    sd, leaf, opt = cpu_train_setup()
    with torch.no_grad():
        cores = best_cpu_threads(sd)
    B = max(2, min(args.batch, 8)) if (args.steps + args.warmup) <= 12 else 4
    mel = synthetic_mel(B, args.frames, seed=1)
    mel_len = torch.full((B,), args.frames, dtype=torch.int64)
    t_out = (((args.frames - 1) // 2 + 1 - 1) // 2 + 1 - 1) // 2 + 1
    y, y_len = synthetic_targets(torch.full((B,), t_out), V, seed=4)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_oracle_train_pass(leaf, opt, mel, mel_len, y, y_len, P)
    times = [cpu_oracle_train_pass(leaf, opt, mel, mel_len, y, y_len, P)[0] for _ in range(args.steps)]
    total = sum(times)
    value = B * args.frames * args.steps / total
    sample = (f"{args.steps} training steps of B={B} x 80 x {args.frames} (train-mode fwd + CTC + autograd backward + torch.optim.Adam), fp32, "
              f"torch CPU {torch.get_num_threads()} threads (best of 8/16/32/64/all on a {os.cpu_count()}-core host)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": train_workload_config(args, B, 1, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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
        for e in prof.events():
            if e.device_type != torch.autograd.DeviceType.CUDA:
                continue
            n = e.name
            if n.lower().startswith(("memcpy", "memset")):
                continue
            if "ec::" in n:
                ours += 1
            else:
                lib_k += 1; names[n.split("<")[0][:60]] = names.get(n.split("<")[0][:60], 0) + 1
        return ours, lib_k, names
    except Exception as ex:                                          # profiler unavailable: report the committed ncu count
        return None, None, {"error": repr(ex)}


def _finish(dist, step):
    """Leave a multi-rank run: drop the captured graphs (they hold NCCL kernels), then exit without waiting on communicator teardown."""
    if dist is None:
        return
    step.close()
    import gc
    gc.collect()
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(0)


def _log(rank, msg):
    if os.environ.get("EFFCONF_BENCH_VERBOSE"):
        print(f"[bench rank {rank} t={time.perf_counter():.1f}] {msg}", file=sys.stderr, flush=True)


def run_train(args, rank, world, local_rank):
    from efficientconformer_b200 import ModelCTC, _lib
    from efficientconformer_b200.trainer import CTCTrainStep
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, T = args.batch, args.frames
    params = dict(P); params["Pdrop"] = args.pdrop
    sd = seeded_state_dict(P, V, seed=0, prefix_encoder="encoder.")

    def make_step(graph):
        model = ModelCTC(params, {"vocab_size": V}, precision=args.precision)
        model.load_state_dict(sd, strict=False)
        model = model.to(dev).train()
        return CTCTrainStep(model, TRAINING_PARAMS, precision=args.precision, use_cuda_graph=graph, sync_bn=not args.no_sync_bn,
                            dropout_seed=1234)
    use_graph = not args.no_graph
    _log(rank, "process group ready")
    step = make_step(use_graph)
    _log(rank, "step object built")
    mel_h = synthetic_mel(B, T, seed=1 + rank).pin_memory()
    t_out = (((T - 1) // 2 + 1 - 1) // 2 + 1 - 1) // 2 + 1
    y, y_len = synthetic_targets(torch.full((B,), t_out), V, seed=4 + rank)
    y_h, yl_h = y.pin_memory(), y_len.pin_memory()
    mel_d, y_d, yl_d = mel_h.to(dev), y_h.to(dev), yl_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    graph_note = "cuda graph replay" if use_graph else "eager launches"
    try:
        for _ in range(max(args.warmup, 3)):
            step.step(mel_d, None, y_d, yl_d); flush.zero_()
        torch.cuda.synchronize()
    except Exception as ex:                              # e.g. a collective that cannot be captured: measure the eager step and say so
        if not use_graph:
            raise
        graph_note = f"eager launches (graph capture failed: {type(ex).__name__})"
        use_graph = False
        step = make_step(False)
        for _ in range(max(args.warmup, 3)):
            step.step(mel_d, None, y_d, yl_d); flush.zero_()
    _log(rank, f"warm-up done ({graph_note})")
    sampler = ClockSampler(local_rank)
    barrier()
    _log(rank, "barrier passed")
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    losses = []
    for a, b in ev:
        flush.zero_()
        a.record(); loss = step.step(mel_d, None, y_d, yl_d); b.record()
        losses.append(loss.clone())
    barrier()
    _log(rank, "timed steps done")
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    losses = [float(l) for l in losses]
    _log(rank, f"ms/step {total_ms / args.steps:.3f}")

    # ---- end to end: every step copies its batch (mel, targets, lengths) from pinned host memory, runs the step, reads the loss back ----
    copy_stream = torch.cuda.Stream(device=dev)
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()

    def e2e_loop(n_steps):
        cur = torch.cuda.current_stream()
        bufs, ready, done = [None, None], [torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()]
        out = []

        def stage(k):
            with torch.cuda.stream(copy_stream):
                if bufs[k & 1] is not None:
                    copy_stream.wait_event(done[k & 1])
                bufs[k & 1] = (mel_h.to(dev, non_blocking=True), y_h.to(dev, non_blocking=True), yl_h.to(dev, non_blocking=True))
                ready[k & 1].record(copy_stream)
        stage(0)
        for k in range(n_steps):
            if k + 1 < n_steps:
                stage(k + 1)
            cur.wait_event(ready[k & 1])
            m, yy, yl = bufs[k & 1]
            ls = step.step(m, None, yy, yl)
            loss_host[k & 1].copy_(ls, non_blocking=True)
            done[k & 1].record(cur)
            if k >= 1:
                done[(k - 1) & 1].synchronize()
                out.append(float(loss_host[(k - 1) & 1]))
        done[(n_steps - 1) & 1].synchronize()
        out.append(float(loss_host[(n_steps - 1) & 1]))
        return out

    e2e_loop(3)
    barrier()
    _log(rank, "e2e warm-up done")
    t0 = time.perf_counter()
    e2e_losses = e2e_loop(args.steps)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    barrier()
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    e2e_s = float(e2e_s)
    _log(rank, "e2e done")

    # ---- operator profile of eager steps: CUDA events around every operator entry point, launching stream (rank 0, no collectives) ----
    prof_ops, ours_k, lib_k, lib_names = {}, None, None, {}
    if world == 1:
        estep = make_step(False)
        for _ in range(2):
            estep.step(mel_d, None, y_d, yl_d)
        torch.cuda.synchronize()
        reps = 3
        with _lib.OpProfile() as prof:
            for _ in range(reps):
                estep.step(mel_d, None, y_d, yl_d)
            summ = prof.summary()
        for k, v in summ.items():
            prof_ops[k] = {"calls": v["calls"] // reps, "ms": v["ms"] / reps, "flops": v["flops"] / reps}
        ours_k, lib_k, lib_names = count_launches(lambda: estep.step(mel_d, None, y_d, yl_d))
        del estep
    if dist is not None:
        # Tear-down: NCCL communicator destruction blocks while captured graphs still reference its kernels, so the graphs go first;
        # should the destruction stall anyway, the result is already out and the process leaves without it.
        dist.barrier()
        torch.cuda.synchronize()
        _log(rank, "collectives done")
    if rank != 0:
        _finish(dist, step)
        return
    pk = peaks()
    tensor_peak = pk["bf16_tflops"] * (0.5 if args.precision == "tf32" else 1.0)
    frames_total = world * B * T * args.steps
    out = {
        "metric": METRIC, "value": frames_total / (total_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": train_workload_config(args, B, world, "gpu", {"launch": graph_note, "sync_bn": world > 1 and not args.no_sync_bn}),
        "e2e": {"value": frames_total / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": mel_h.numel() * 4 + y_h.numel() * 8 + yl_h.numel() * 8, "d2h_bytes_per_step": 4,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "timing": "wall clock around K CTCTrainStep.step calls with HOST batches (pinned mel / targets -> H2D on a copy stream, 2-deep "
                          "pipeline, loss read back every step), synchronised both sides"},
        "clocks": clocks,
        "step_ms_min_med_max": [round(min(step_ms), 4), round(statistics.median(step_ms), 4), round(max(step_ms), 4)],
        "loss_first_last": [losses[0], losses[-1]], "e2e_loss_last": e2e_losses[-1], "lr_after": step.lr(), "optimizer_steps": step.steps_done(),
    }
    if prof_ops:
        tot_ms = sum(v["ms"] for v in prof_ops.values())
        ops_sorted = sorted(prof_ops.items(), key=lambda kv: -kv[1]["ms"])
        out["operators"] = [{"op": k, "calls": v["calls"], "ms": round(v["ms"], 4),
                             "tflops": round(v["flops"] / v["ms"] / 1e9, 2) if v["flops"] else None} for k, v in ops_sorted]
        tensor_ops = [(k, v) for k, v in ops_sorted if v["flops"] > 0]
        k, v = tensor_ops[0]
        desc = {"ec_op_wgrad": "ec_op_wgrad = wgrad_tc_kernel (tcgen05, MN-major operands, split-M) + fixed-order reduce: every weight gradient of one step",
                "ec_op_gemm": "ec_op_gemm = gemm_tc_kernel (tcgen05 + TMA): every forward Linear / pointwise conv and every data-gradient GEMM of one step"}
        ach = v["flops"] / v["ms"] / 1e9

        def roof(k, v):
            a = v["flops"] / v["ms"] / 1e9
            kern = "wgrad_tc_kernel" if k == "ec_op_wgrad" else "gemm_tc_kernel"
            return {"kernel": desc.get(k, k), "bound": "tensor", "achieved": round(a, 2), "peak": tensor_peak, "unit": "TFLOP/s",
                    "frac": round(a / tensor_peak, 4), "traffic": ncu_traffic(args.precision, kern, "r1_ncu_full_train"),
                    "peak_source": pk["source"] + (" bf16 cuBLAS burst" if args.precision == "bf16" else " bf16 cuBLAS burst / 2 (tf32 operands)"),
                    "launches_per_step": v["calls"], "avg_launch_us": round(1e3 * v["ms"] / max(v["calls"], 1), 2),
                    "share_of_step": round(v["ms"] / tot_ms, 3),
                    "algorithmic_flops_per_step": v["flops"],
                    "timing": "CUDA events around every operator call of an eager step on the launching stream (serialised), mean of 3 steps"}
        out["roofline"] = roof(k, v)
        out["rooflines_other"] = [roof(k2, v2) for k2, v2 in tensor_ops[1:]]
        step_flops = 3.0 * TRAIN_FLOP_PER_FRAME.get(T, 6.616e6) * B * T
        out["whole_step"] = {"algorithmic_tflop": round(step_flops / 1e12, 4), "achieved_tflops": round(step_flops / (total_ms / args.steps) / 1e9, 2),
                             "frac_of_tensor_peak": round(step_flops / (total_ms / args.steps) / 1e9 / tensor_peak, 4),
                             "note": "3 x forward FLOPs of SURVEY.md 8(d) per mel frame; the step is launch / latency bound, not math bound"}
        out["eager_profiled_step_ms"] = round(tot_ms, 3)
    launches = ours_k if ours_k is not None else 2048
    out["gpu_launches"] = launches * args.steps
    out["launches_per_step"] = launches
    out["library_launches_per_step"] = {"count": lib_k, "kernels": lib_names}
    if world == 1 and not args.no_cpu_baseline:
        sd2, leaf, opt = cpu_train_setup()
        with torch.no_grad():
            best_cpu_threads(sd2)
        Bs = 4
        cm, cl = mel_h[:Bs].clone(), torch.full((Bs,), T, dtype=torch.int64)
        cpu_oracle_train_pass(leaf, opt, cm, cl, y[:Bs], y_len[:Bs], P)
        runs = [cpu_oracle_train_pass(leaf, opt, cm, cl, y[:Bs], y_len[:Bs], P)[0] for _ in range(3)]
        sec = statistics.median(runs)
        out["cpu_baseline"] = {"value": Bs * T / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"median of 3 training steps on the first {Bs} utterances of the same batch (oracle train-mode fwd + CTC + autograd "
                                         f"backward + torch.optim.Adam), fp32 torch CPU, {torch.get_num_threads()} threads"}
    print(json.dumps(out))
    sys.stdout.flush()
    _finish(dist, step)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16x2", choices=["bf16x2", "bf16", "tf32"],
                    help="operand mode: bf16x2 = packed bf16 hi/lo pairs (default; meets the 1e-3 parity gate), bf16 = fast mode, tf32")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="train", choices=["train", "forward"])
    ap.add_argument("--pdrop", type=float, default=0.1, help="dropout probability of the training step (reference config: 0.1)")
    ap.add_argument("--watchdog", type=int, default=600, help="seconds after which a stuck run dumps its Python stacks to stderr and exits")
    ap.add_argument("--no-graph", action="store_true", help="training step: eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-sync-bn", action="store_true", help="N > 1: per-rank BatchNorm statistics (NOT the reference's SyncBatchNorm semantics)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = (40 if args.mode == "train" else 100) if args.impl == "ours" else 3
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        (run_reference_train if args.mode == "train" else run_reference)(args, rank, world)
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
    (run_train if args.mode == "train" else run_ours)(args, rank, world, local)


if __name__ == "__main__":
    main()
