"""Debug: which gradients of the drop-in are non-finite under autocast + GradScaler (tests/test_gpu_dropin.py)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_dropin as T
for pdrop, aug, scale in ((0.1, True, 65536.0), (0.0, False, 65536.0), (0.0, False, 1.0), (0.1, False, 1.0), (0.0, True, 1.0)):
    cfg = T._config(pdrop=pdrop, spec_augment=aug)
    _, drop = T._reference_and_dropin(cfg)
    drop.train()
    batch = T._batch()
    with torch.cuda.amp.autocast(enabled=True):
        pred = drop.forward(batch)
        loss = drop.criterion(batch, pred)
    (loss * scale).backward()
    bad = [(k, int((~torch.isfinite(p.grad)).sum()), p.grad.numel()) for k, p in drop.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    mx = max(float(p.grad.abs().max()) for k, p in drop.named_parameters() if p.grad is not None and torch.isfinite(p.grad).all())
    print(f"pdrop {pdrop} aug {aug} scale {scale}: loss {float(loss):.4f} logits finite {bool(torch.isfinite(pred[0]).all())} non-finite grads {len(bad)} max finite |g| {mx:.3e}")
    for b in bad[:12]:
        print("   ", b)
