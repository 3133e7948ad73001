"""Timing probe of the native training step (CTCTrainStep) on one GPU: eager vs CUDA-graph replay, CUDA events, peak memory.
    python tools/train_probe.py [B] [T] [precision] [pdrop]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientconformer_b200.config import CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V  # noqa: E402
from efficientconformer_b200.model_ctc import ModelCTC  # noqa: E402
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel, synthetic_targets  # noqa: E402
from efficientconformer_b200.trainer import CTCTrainStep  # noqa: E402

TP = dict(optimizer="Adam", beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-6, lr_schedule="Transformer", schedule_dim=240,
          warmup_steps=10000, K=2)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    pdrop = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
    dev = "cuda"
    params = dict(P); params["Pdrop"] = pdrop
    out = {"B": B, "T": T, "precision": prec, "pdrop": pdrop}
    mel = synthetic_mel(B, T, seed=1).to(dev)
    t_out = ((((T - 1) // 2 + 1) - 1) // 2 + 1 - 1) // 2 + 1
    y, yl = synthetic_targets(torch.full((B,), t_out), V, seed=4)
    y, yl = y.to(dev), yl.to(dev)
    for graph in (False, True):
        model = ModelCTC(params, {"vocab_size": V}, precision=prec)
        model.load_state_dict(seeded_state_dict(P, V, seed=0, prefix_encoder="encoder."), strict=False)
        model = model.to(dev).train()
        step = CTCTrainStep(model, TP, precision=prec, use_cuda_graph=graph)
        torch.cuda.reset_peak_memory_stats()
        t0 = time.perf_counter()
        for _ in range(3):
            loss = step.step(mel, None, y, yl)
        torch.cuda.synchronize()
        out[f"warm_s_graph{int(graph)}"] = time.perf_counter() - t0
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        t0 = time.perf_counter()
        losses = []
        for a, b in ev:
            a.record(); loss = step.step(mel, None, y, yl); b.record()
            losses.append(loss.clone())
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / len(ev) * 1e3
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        out[f"graph{int(graph)}"] = {"ms_med": ms[len(ms) // 2], "ms_min": ms[0], "wall_ms": wall, "loss": [float(l) for l in losses][:4],
                                      "peak_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
