"""Pin the CPU oracle (oracle/conformer_oracle.py) against golden vectors produced by the real reference
(tests/golden/make_golden.py).  Tolerances: fp32 oracle vs fp32 reference, different op order only."""
import json
import os

import pytest
import torch

from efficientconformer_b200.config import (CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V, resolve_blocks,
                                            state_dict_layout, stage_lengths)
from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_mel, synthetic_audio
from oracle import conformer_oracle as O


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def sd():
    return seeded_state_dict(P, V, seed=0, prefix_encoder="encoder.")


def test_state_dict_layout_matches_reference(golden_dir):
    layouts = json.load(open(os.path.join(golden_dir, "state_dict_layouts.json")))
    for name, entry in layouts.items():
        ref = {k: tuple(s) for k, s in entry["keys"]}
        mine = {("encoder." + k if not k.startswith("fc.") else k): tuple(s)
                for k, s in state_dict_layout(entry["encoder_params"], entry["vocab_size"])}
        assert mine == ref, (name, set(mine) ^ set(ref))
        # order of the reference state_dict is reproduced too
        assert [k for k, _ in entry["keys"]] == list(mine.keys()), name


def test_ctc_small_params_are_the_shipped_config(golden_dir):
    layouts = json.load(open(os.path.join(golden_dir, "state_dict_layouts.json")))
    assert layouts["EfficientConformerCTCSmall"]["encoder_params"] == P
    specs = resolve_blocks(P)
    assert [s.dim_head for s in specs] == [90] * 5 + [42] * 5 + [60] * 5
    assert [s.conv_stride for s in specs] == [1, 1, 1, 1, 2, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1]
    assert stage_lengths(P, 1000)[1] == 125


def test_config1_logits_loss_greedy(sd, golden_dir):
    g = torch.load(os.path.join(golden_dir, "ctc_small_b2_t500.pt"))
    mel = synthetic_mel(2, 500, seed=g["mel_seed"])
    taps = {}
    logits, out_len = O.model_ctc_forward_mel(sd, P, mel, g["mel_len"], taps=taps)
    assert torch.equal(out_len, g["out_len"])
    assert rel_l2(logits, g["logits"]) < 2e-5
    for k, v in g["taps"].items():
        assert rel_l2(taps[k], v) < 2e-5, k
    loss, per = O.ctc_loss(logits, out_len, g["targets"], g["target_len"])
    assert abs(float(loss) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    assert torch.allclose(per, g["loss_per_utt"], rtol=1e-5)
    assert O.greedy_ids(g["logits"], g["out_len"]) == g["greedy"]


def test_config1_fp64_oracle_close_to_fp32_reference(sd, golden_dir):
    g = torch.load(os.path.join(golden_dir, "ctc_small_b2_t500.pt"))
    mel = synthetic_mel(2, 500, seed=g["mel_seed"]).double()
    logits, _ = O.model_ctc_forward_mel(sd, P, mel, g["mel_len"])
    assert rel_l2(logits.float(), g["logits"]) < 2e-5


def test_audio_front_end(sd, golden_dir):
    g = torch.load(os.path.join(golden_dir, "ctc_small_audio_b2_t200.pt"))
    audio = synthetic_audio(2, g["t_mel"], seed=g["audio_seed"])
    sd2 = dict(sd)
    sd2["encoder.preprocessing.Spectrogram.window"] = g["window"]
    sd2["encoder.preprocessing.MelScale.fb"] = g["fb"]
    mel, mel_len = O.audio_to_mel(sd2, P, audio, g["audio_len"], prefix="encoder.")
    assert torch.equal(mel_len, g["mel_len"])
    assert mel.shape[-1] == g["t_mel"]
    assert (mel - g["mel"].float()).abs().max() < 2e-2          # golden mel stored as fp16
    logits, out_len = O.model_ctc_forward_mel(sd2, P, mel, mel_len)
    assert torch.equal(out_len, g["out_len"])
    assert rel_l2(logits, g["logits"]) < 1e-4


def test_modules_awkward_lengths(sd, golden_dir):
    g = torch.load(os.path.join(golden_dir, "ctc_small_modules.pt"))
    enc = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    specs = resolve_blocks(P)
    for key, e in g.items():
        if key.startswith("sub_"):
            Tm = int(key[5:])
            gen = torch.Generator().manual_seed(9000 + Tm)
            m = torch.randn(2, 80, Tm, generator=gen)
            ml = torch.tensor([Tm, max(1, Tm // 2)])
            h, l = O.conv2d_subsampling(enc, P, m, ml)
            out = O.EXACT.linear(h.transpose(1, 2), enc["linear.weight"], enc["linear.bias"])
            assert torch.equal(l, e["out_len"])
            assert rel_l2(out, e["out"]) < 1e-5, key
            continue
        bi, Tq = int(key[1:key.index("_")]), int(key[key.index("T") + 1:])
        spec = specs[bi]
        gen = torch.Generator().manual_seed(100 * bi + Tq)
        x = torch.randn(2, Tq, spec.dim_model, generator=gen)
        out, w = O.conformer_block(enc, f"blocks.{bi}", x, e["x_len"], spec)
        assert out.shape == e["block"].shape, key
        assert rel_l2(out, e["block"]) < 2e-5, key
        if "ffn1" in e:
            p = f"blocks.{bi}"
            assert rel_l2(O.feed_forward(enc, f"{p}.feed_forward_module1", x), e["ffn1"]) < 1e-5, key
            m = f"{p}.multi_head_self_attention_module"
            a_in = O.layer_norm(x, enc[f"{m}.norm.weight"], enc[f"{m}.norm.bias"])
            att, w = O.relpos_attention(enc, f"{m}.mhsa", a_in, e["x_len"], spec)
            assert rel_l2(att, e["mhsa"]) < 1e-5, key
            assert (w - e["att_w"]).abs().max() < 1e-5, key
            assert rel_l2(O.conv_module(enc, f"{p}.convolution_module", x, spec), e["conv"]) < 1e-5, key


def test_ctc_loss_and_greedy_known_answers(golden_dir):
    g = torch.load(os.path.join(golden_dir, "ctc_loss_small.pt"))
    loss, per = O.ctc_loss(g["logits"], g["logits_len"], g["targets"], g["target_len"])
    assert torch.allclose(per, g["loss_per_utt"], rtol=1e-5, atol=1e-5)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    # collapse rule on hand-made sequences: merge repeats, then drop blanks
    lg = torch.full((1, 8, 4), -5.0)
    for t, tok in enumerate([1, 1, 0, 1, 2, 2, 0, 3]):
        lg[0, t, tok] = 5.0
    assert O.greedy_ids(lg, torch.tensor([8])) == [[1, 1, 2, 3]]
    assert O.greedy_ids(lg, torch.tensor([2])) == [[1]]
    assert O.greedy_ids(lg, torch.tensor([0])) == [[]]


def test_operand_rounding_emulation_predicts_tf32_within_gate(sd, golden_dir):
    """The parity gate is 1e-3 (BASELINE.json north_star).  TF32 operands (what the tcgen05 kind::tf32 path
    feeds the tensor cores) must leave headroom; bf16 operands do not, which is why parity runs in tf32 mode."""
    g = torch.load(os.path.join(golden_dir, "ctc_small_b2_t500.pt"))
    mel = synthetic_mel(2, 500, seed=g["mel_seed"])
    lt, _ = O.model_ctc_forward_mel(sd, P, mel, g["mel_len"], nm=O.Numerics("tf32"))
    lb, _ = O.model_ctc_forward_mel(sd, P, mel, g["mel_len"], nm=O.Numerics("bf16"))
    et, eb = rel_l2(lt, g["logits"]), rel_l2(lb, g["logits"])
    print("emulated rel-L2: tf32 %.2e  bf16 %.2e" % (et, eb))
    assert et < 1e-3
    assert eb > et


# ---- the other shipped configs (BASELINE.json configs 3-5 and the rest of reference configs/*.json) ----------------------
OTHER_CASES = ["EfficientConformerCTCLarge", "EfficientConformerCTCMedium", "EfficientConformerTransducerMedium",
               "EfficientConformerTransducerLarge", "ConformerCTCLarge", "ConformerCTCLarge@long", "ConformerCTCMedium",
               "ConformerCTCSmall", "ConformerTransducerSmall", "ConformerTransducerLarge"]


def test_shipped_config_table_matches_reference(golden_dir):
    from efficientconformer_b200.config import SHIPPED_ENCODER_PARAMS
    layouts = json.load(open(os.path.join(golden_dir, "state_dict_layouts.json")))
    for name, entry in layouts.items():
        assert SHIPPED_ENCODER_PARAMS[name] == (entry["encoder_params"], entry["vocab_size"]), name
    assert len(SHIPPED_ENCODER_PARAMS) == 12


@pytest.mark.parametrize("case", OTHER_CASES)
def test_other_configs_against_reference_golden(golden_dir, case):
    """Two-layer Conv2d front end, k = 31, widths up to 720, head dims 24 ... 135, T' up to 250: the oracle reproduces the
    real reference's outputs for every shipped encoder family (fp32 vs fp32)."""
    from efficientconformer_b200.config import SHIPPED_ENCODER_PARAMS
    g = torch.load(os.path.join(golden_dir, "other_configs.pt"))[case]
    params, vocab = SHIPPED_ENCODER_PARAMS[case.split("@")[0]]
    mel = synthetic_mel(g["batch"], g["t_mel"], seed=g["mel_seed"])
    if "logits" in g:
        sd2 = seeded_state_dict(params, vocab, seed=g["weights_seed"], prefix_encoder="encoder.")
        out, out_len = O.model_ctc_forward_mel(sd2, params, mel, g["mel_len"])
        ref = g["logits"]
    else:
        sd2 = seeded_state_dict(params, None, seed=g["weights_seed"])
        out, out_len = O.encoder_forward_mel(sd2, params, mel, g["mel_len"])
        ref = g["x"]
    assert torch.equal(out_len, g["out_len"])
    assert rel_l2(out, ref) < 2e-5


# ---- training step: the oracle's train-mode forward + autograd against the real reference ----
def test_training_step_loss_gradients_and_running_stats(sd, golden_dir):
    """Train-mode BatchNorm (batch statistics incl. padded frames, running-stat update), Pdrop = 0: CTC loss, the gradient of
    every parameter and the updated running statistics equal the reference's loss.backward() (fp32 vs fp32)."""
    g = torch.load(os.path.join(golden_dir, "ctc_small_train_b2_t500.pt"))
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    bn = {"updates": {}}
    mel = synthetic_mel(2, 500, seed=g["mel_seed"])
    logits, out_len = O.model_ctc_forward_mel(leaf, P, mel, g["mel_len"], bn=bn)
    assert rel_l2(logits.detach(), g["logits"]) < 2e-5
    loss, _ = O.ctc_loss(logits, out_len, g["targets"], g["target_len"])
    assert abs(float(loss) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    loss.backward()
    worst = 0.0
    # gradients that are zero in exact arithmetic (biases in front of BatchNorm, key / positional biases under the softmax's
    # shift invariance) are fp32 noise of magnitude ~1e-6 in the reference: compared against an absolute floor
    floor = 1e-4 * sorted(g["grad_norms"].values())[len(g["grad_norms"]) // 2]
    for k, ref_norm in g["grad_norms"].items():
        gn = float(leaf[k].grad.double().norm())
        assert gn == gn, k
        if ref_norm < floor:
            assert gn < floor, k
            continue
        worst = max(worst, abs(gn - ref_norm) / ref_norm)
    assert worst < 2e-3, worst                      # fp32 reference backward (op order differs); norms of all 582 gradients
    for k, ref in g["grads"].items():
        if g["grad_norms"][k] >= floor:
            assert rel_l2(leaf[k].grad, ref) < 2e-3, k
    for k, ref in g["running_stats"].items():
        assert rel_l2(bn["updates"][k[len("encoder."):]], ref) < 1e-5, k


def test_training_step_medium_config_loss_gradients_and_running_stats(golden_dir):
    """The oracle's train-mode forward + autograd against the real reference's loss.backward() on the second shipped config
    (EfficientConformerCTCMedium, B=2, 80x250, lengths 250 / 163): loss, all 620 gradient norms, running statistics."""
    from efficientconformer_b200.config import SHIPPED_ENCODER_PARAMS
    PM, VM = SHIPPED_ENCODER_PARAMS["EfficientConformerCTCMedium"]
    g = torch.load(os.path.join(golden_dir, "ctc_medium_train_b2_t250.pt"))
    sdm = seeded_state_dict(PM, VM, seed=0, prefix_encoder="encoder.")
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sdm.items()}
    bn = {"updates": {}}
    logits, out_len = O.model_ctc_forward_mel(leaf, PM, synthetic_mel(2, 250, seed=g["mel_seed"]), g["mel_len"], bn=bn)
    assert rel_l2(logits.detach(), g["logits"]) < 2e-5
    loss, _ = O.ctc_loss(logits, out_len, g["targets"], g["target_len"])
    assert abs(float(loss.detach()) - float(g["loss"])) / abs(float(g["loss"])) < 1e-5
    loss.backward()
    floor = 1e-4 * sorted(g["grad_norms"].values())[len(g["grad_norms"]) // 2]
    worst = 0.0
    for k, ref_norm in g["grad_norms"].items():
        gn = float(leaf[k].grad.double().norm())
        if ref_norm < floor:
            assert gn < floor, k
            continue
        worst = max(worst, abs(gn - ref_norm) / ref_norm)
    assert worst < 2e-3, worst
    for k, ref in g["running_stats"].items():
        assert rel_l2(bn["updates"][k[len("encoder."):]], ref) < 1e-5, k


def test_rnnt_oracle_against_reference_joint_and_torchaudio_loss(golden_dir):
    """the joint restatement against the real reference JointNetwork's logits, the RNN-T recursion against
    torchaudio.functional.rnnt_loss values (tests/golden/make_golden_rnnt.py)."""
    from oracle import rnnt_oracle as R
    cases = torch.load(os.path.join(golden_dir, "rnnt_joint_small.pt"))
    for name, c in cases.items():
        logits = R.joint_forward(c["state_dict"], c["f"], c["g"])
        ref = c["logits"]
        assert rel_l2(logits[:, :ref.shape[1]], ref) < 2e-6, name
        mean, per = R.rnnt_loss(logits.double(), c["y"], c["f_len"], c["y_len"])
        assert torch.allclose(per.float(), c["loss_per_utt"], rtol=2e-5, atol=1e-4), (name, per, c["loss_per_utt"])
        assert abs(float(mean) - float(c["loss_mean"])) < 2e-5 * abs(float(c["loss_mean"]))


def test_frontend_oracle_against_reference_golden(golden_dir):
    """oracle/frontend_oracle.py pinned: log-mel against the REAL reference AudioPreprocessing, the SpecAugment span arithmetic against
    torchaudio's mask_along_axis with preset uniforms (tests/golden/make_golden_frontend.py)."""
    import numpy as np
    from oracle import frontend_oracle as FO
    g = torch.load(os.path.join(golden_dir, "frontend_small.pt"))
    for name, c in g["logmel"].items():
        mel, mel_len = FO.logmel(c["audio"].numpy(), normalize=c["normalize"], mean=c["mean"], std=c["std"], audio_len=c["audio_len"].numpy())
        ref = c["mel"].double().numpy()
        assert mel.shape == ref.shape and (mel_len == c["mel_len"].numpy()).all()
        assert np.abs(mel - ref).max() < 2e-4 and np.linalg.norm(mel - ref) / np.linalg.norm(ref) < 5e-6, name
    for s in g["spans"]:
        st, en = FO.span_from_uniforms(s["u1"], s["u2"], s["param"], s["size"])
        assert max(en - st, 0) == s["width"] and (s["width"] == 0 or (st, en) == (s["start"], s["end"])), s
    # the counter-based application: masks inside the valid frames, frequency masks shared by the batch, deterministic in (seed, step)
    mel = np.ones((3, 80, 200), dtype=np.float32)
    a = FO.specaugment_apply(mel, [200, 120, 7], 5, 1, 2, 27, 5, 0.05)
    b = FO.specaugment_apply(mel, [200, 120, 7], 5, 1, 2, 27, 5, 0.05)
    c2 = FO.specaugment_apply(mel, [200, 120, 7], 5, 2, 2, 27, 5, 0.05)
    assert (a == b).all() and not (a == c2).all()
    assert ((a == 0).all(2) == (a == 0).all(2)[0:1]).all()
    assert not (a[1] == 0).all(0)[120:].any() and not (a[2] == 0).all(0)[7:].any()


def test_specaugment_torch_path_follows_the_pinned_mask_arithmetic():
    """The CPU-tensor route of encoders.SpecAugment (batched torch ops) fed with preset uniforms draws exactly the spans the pinned
    arithmetic (oracle span_from_uniforms == torchaudio mask_along_axis, see the golden above) gives: frequency masks shared by the
    batch, time masks per utterance inside its valid frames with parameter int(pS * x_len[b])."""
    import numpy as np
    from oracle import frontend_oracle as FO
    from efficientconformer_b200.encoders import SpecAugment
    B, F, T, mF, Fp, mT, pS = 3, 80, 400, 2, 27, 5, 0.05
    lens = torch.tensor([400, 250, 37])
    rng = np.random.default_rng(7)
    draws = [torch.tensor(rng.random(mF), dtype=torch.float32), torch.tensor(rng.random(mF), dtype=torch.float32),
             torch.tensor(rng.random((B, mT)), dtype=torch.float32), torch.tensor(rng.random((B, mT)), dtype=torch.float32)]
    seq = iter(draws)
    real_rand = torch.rand
    torch.rand = lambda *a, **k: next(seq)
    try:
        y = SpecAugment(True, mF, Fp, mT, pS)(torch.ones(B, F, T), lens)
    finally:
        torch.rand = real_rand
    ref = np.ones((B, F, T), dtype=np.float32)
    for i in range(mF):
        s0, e0 = FO.span_from_uniforms(float(draws[0][i]), float(draws[1][i]), Fp, F)
        ref[:, s0:e0, :] = 0
    for b in range(B):
        ln = int(lens[b]); Tp = int(np.float32(pS) * np.float32(ln))
        for j in range(mT):
            s0, e0 = FO.span_from_uniforms(float(draws[2][b, j]), float(draws[3][b, j]), Tp, ln)
            ref[b, :, s0:min(e0, ln)] = 0
    assert torch.equal(y, torch.from_numpy(ref))
    assert not SpecAugment(False, mF, Fp, mT, pS)(torch.ones(B, F, T), lens).eq(0).any()
