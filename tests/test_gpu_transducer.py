"""Transducer joint network + RNN-T loss on the B200: forward against the golden logits of the REAL
reference JointNetwork and the torchaudio RNN-T loss values (tests/golden/make_golden_rnnt.py) and against the CPU oracle on seeded inputs
at the shipped EfficientConformerTransducerMedium dimensions; backward (loss gradient, joint gradients) against torch autograd over the
oracle; edge cases of the lattice (single frame, empty transcript, ragged lengths, repeated labels)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _joint(c, precision):
    from efficientconformer_b200.transducer import JointNetwork
    B, T, U, Denc, Ddec, J, V = c["dims"]
    jn = JointNetwork(Denc, Ddec, V, {"joint_mode": "sum", "dim_model": J, "act": "tanh"}, precision=precision)
    print(jn.load_state_dict({k[len("joint_network."):]: v for k, v in c["state_dict"].items()}, strict=True))   # the reference's names / shapes
    return jn.to(DEV).eval()


@pytest.mark.parametrize("precision,tol", [("bf16x2", 1e-4), ("tf32", 2e-3), ("bf16", 2e-2)])
def test_joint_and_rnnt_loss_against_reference_golden(golden_dir, precision, tol):
    from efficientconformer_b200.transducer import rnnt_loss, LossRNNT
    cases = torch.load(os.path.join(golden_dir, "rnnt_joint_small.pt"))
    for name, c in cases.items():
        jn = _joint(c, precision)
        with torch.no_grad():
            logits = jn(c["f"].to(DEV), c["g"].to(DEV))
        ref = c["logits"]
        assert logits.shape[:1] + logits.shape[2:] == ref.shape[:1] + ref.shape[2:]
        e = rel_l2(logits[:, :ref.shape[1]], ref)
        print(f"[{precision}] {name}: joint logits vs the reference module rel-L2 {e:.3e}")
        assert e < tol, (name, e)
        # the loss kernels on the REFERENCE's logits: isolates them from the operand rounding of the joint
        if name == "small":
            mean, per = rnnt_loss(ref.to(DEV), c["y"], c["f_len"], c["y_len"])
            assert torch.allclose(per.cpu(), c["loss_per_utt"], rtol=2e-5, atol=1e-4), (per, c["loss_per_utt"])
            assert abs(float(mean) - float(c["loss_mean"])) < 2e-5 * abs(float(c["loss_mean"]))
        # end to end: joint + loss through the reference-facing LossRNNT.forward(batch, pred)
        with torch.no_grad():
            loss = LossRNNT()((None, c["y"].to(DEV), None, c["y_len"].to(DEV)), (logits, c["f_len"].to(DEV), None))
        assert abs(float(loss) - float(c["loss_mean"])) < max(tol, 1e-4) * abs(float(c["loss_mean"])), (name, float(loss), float(c["loss_mean"]))
        # decoding form: (B, Denc), (B, Ddec) -> (B, V)
        with torch.no_grad():
            step = jn(c["f"][:, 3].to(DEV), c["g"][:, 2].to(DEV))
        assert rel_l2(step, logits[:, 3, 2]) < 1e-6


def test_transducer_medium_dimensions_against_oracle():
    """Seeded inputs at the shipped EfficientConformerTransducerMedium dimensions (encoder 360, decoder 640, joint 640, vocab 1000),
    ragged frame / label lengths: joint logits and loss against the CPU oracle (fp64 recursion)."""
    from oracle import rnnt_oracle as R
    from efficientconformer_b200.transducer import JointNetwork, rnnt_loss
    B, T, U, Denc, Ddec, J, V = 4, 40, 17, 360, 640, 640, 1000
    g = torch.Generator().manual_seed(3)
    jn = JointNetwork(Denc, Ddec, V, {"joint_mode": "sum", "dim_model": J, "act": "tanh"})
    sd = {k: (torch.randn(v.shape, generator=g) / (v.shape[-1] ** 0.5 if v.dim() == 2 else 10.0)) for k, v in jn.state_dict().items()}
    jn.load_state_dict(sd)
    jn = jn.to(DEV).eval()
    f, gd = torch.randn(B, T, Denc, generator=g), torch.randn(B, U + 1, Ddec, generator=g)
    y = torch.randint(1, V, (B, U), generator=g)
    f_len, y_len = torch.tensor([40, 33, 21, 1]), torch.tensor([17, 9, 0, 3])
    with torch.no_grad():
        logits = jn(f.to(DEV), gd.to(DEV))
    ref = R.joint_forward({"joint_network." + k: v for k, v in sd.items()}, f, gd)
    assert rel_l2(logits, ref) < 1e-4
    mean, per = rnnt_loss(logits, y, f_len, y_len)
    ref_mean, ref_per = R.rnnt_loss(ref.double(), y, f_len, y_len)
    print("per-utterance losses", per.tolist(), ref_per.tolist())
    assert torch.allclose(per.cpu().double(), ref_per, rtol=1e-4, atol=1e-3)
    assert abs(float(mean) - float(ref_mean)) < 1e-4 * abs(float(ref_mean))
    # property at scale: the loss of an utterance does not depend on what is batched next to it (padding is never read)
    mean1, per1 = rnnt_loss(logits[1:2, :33, :10].contiguous(), y[1:2, :9], f_len[1:2], y_len[1:2])
    assert abs(float(per1[0]) - float(per[1])) < 1e-5 * abs(float(per[1]))


def test_rnnt_gradient_against_oracle_autograd():
    """d(mean nll) / d logits (log_softmax folded in) against torch autograd over the fp64 oracle recursion: ragged lattices, a single-frame
    utterance, an empty transcript, repeated labels."""
    from oracle import rnnt_oracle as R
    from efficientconformer_b200.transducer import rnnt_loss_and_grad
    g = torch.Generator().manual_seed(21)
    for (B, T, U, V, f_len, y_len) in ((4, 13, 6, 11, [13, 9, 1, 5], [6, 0, 3, 6]), (2, 40, 17, 1000, [40, 31], [17, 12]), (1, 1, 1, 8, [1], [1])):
        logits = 2.0 * torch.randn(B, T, U + 1, V, generator=g)
        y = torch.randint(1, min(V, 4), (B, U), generator=g)                  # small alphabet: repeated labels
        fl, yl = torch.tensor(f_len), torch.tensor(y_len)
        ref_in = logits.double().requires_grad_(True)
        ref_mean, ref_per = R.rnnt_loss(ref_in, y, fl, yl)
        ref_mean.backward()
        mean, per, grad = rnnt_loss_and_grad(logits.to(DEV), y, fl, yl)
        assert torch.allclose(per.cpu().double(), ref_per.detach(), rtol=1e-5, atol=1e-4)
        e = rel_l2(grad, ref_in.grad)
        print(f"RNN-T gradient B={B} T={T} U={U} V={V}: rel-L2 {e:.3e}")
        assert e < 1e-4, (B, T, U, V, e)        # fp32 log-domain exponents at |log p| ~ 3e2: 3e-5 measured at V = 1000
        assert float(grad[:, :, :, :].abs().sum()) > 0 and torch.isfinite(grad).all()
        for b in range(B):                                                    # nothing outside the valid lattice
            assert float(grad[b, f_len[b]:].abs().max()) == 0.0 if f_len[b] < T else True
            assert float(grad[b, :, y_len[b] + 1:].abs().max()) == 0.0 if y_len[b] < U else True


@pytest.mark.parametrize("precision,tol", [("bf16x2", 2e-4), ("tf32", 5e-3)])
def test_joint_and_loss_training_step_against_oracle_autograd(precision, tol):
    """The reference's two-call structure, Transducer.forward -> criterion (models/transducer.py:88-106, models/losses.py:22-46), through the
    drop-in JointNetwork + LossRNNT with loss.backward(): loss and the gradients of f, g and the six joint parameters against torch
    autograd over the CPU oracle (fp64)."""
    from oracle import rnnt_oracle as R
    from efficientconformer_b200.transducer import JointNetwork, LossRNNT
    for (B, T, U, Denc, Ddec, J, V) in ((3, 21, 7, 48, 40, 64, 37), (2, 30, 11, 360, 640, 640, 1000)):
        g = torch.Generator().manual_seed(B * 7 + T)
        jn = JointNetwork(Denc, Ddec, V, {"joint_mode": "sum", "dim_model": J, "act": "tanh"}, precision=precision)
        sd = {k: (torch.randn(v.shape, generator=g) / (v.shape[-1] ** 0.5 if v.dim() == 2 else 10.0)) for k, v in jn.state_dict().items()}
        jn.load_state_dict(sd)
        jn = jn.to(DEV).train()
        f, gd = torch.randn(B, T, Denc, generator=g), torch.randn(B, U + 1, Ddec, generator=g)
        y = torch.randint(1, V, (B, U), generator=g)
        f_len = torch.tensor([T] + [max(1, T - 4 * b) for b in range(1, B)])
        y_len = torch.tensor([U] + [max(0, U - 3 * b) for b in range(1, B)])
        # oracle
        leaf = {"joint_network." + k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
        fr, gr = f.double().requires_grad_(True), gd.double().requires_grad_(True)
        ref_mean, _ = R.rnnt_loss(R.joint_forward(leaf, fr, gr), y, f_len, y_len)
        ref_mean.backward()
        # drop-in
        fd, gdd = f.to(DEV).requires_grad_(True), gd.to(DEV).requires_grad_(True)
        logits = jn(fd, gdd)
        loss = LossRNNT()((None, y.to(DEV), None, y_len.to(DEV)), (logits, f_len.to(DEV), None))
        (2.0 * loss).backward()                                               # an upstream factor (GradScaler) must pass through
        assert abs(float(loss) - float(ref_mean)) < max(tol, 1e-4) * abs(float(ref_mean))
        worst = (rel_l2(fd.grad, 2 * fr.grad), "f")
        worst = max(worst, (rel_l2(gdd.grad, 2 * gr.grad), "g"))
        for k, p in jn.named_parameters():
            worst = max(worst, (rel_l2(p.grad, 2 * leaf["joint_network." + k].grad), k))
        print(f"[{precision}] joint training step {(B, T, U, Denc, Ddec, J, V)}: loss {float(loss):.5f} (oracle {float(ref_mean):.5f}), worst gradient rel-L2 {worst[0]:.3e} ({worst[1]})")
        assert worst[0] < tol, worst


def test_no_cpu_fallback():
    from efficientconformer_b200.transducer import JointNetwork, rnnt_loss
    jn = JointNetwork(8, 8, 16, {"joint_mode": "sum", "dim_model": 8, "act": "tanh"})
    with pytest.raises(RuntimeError):
        jn(torch.zeros(1, 2, 8), torch.zeros(1, 3, 8))                      # CPU tensors: no CPU path
    with pytest.raises(RuntimeError):
        rnnt_loss(torch.zeros(1, 2, 3, 8), torch.ones(1, 2, dtype=torch.long), torch.tensor([2]), torch.tensor([2]))
