"""Front end on the B200 (csrc/frontend.cu): the one-kernel log-mel features against the golden output of
the REAL reference AudioPreprocessing (tests/golden/make_golden_frontend.py) and against the float64 oracle at the bench shape, the
drop-in `forward(audio)` with and without it, and the SpecAugment kernel against the bit-exact restatement of its counter-based draws
(oracle/frontend_oracle.py; the mask arithmetic itself is pinned to torchaudio's mask_along_axis in tests/test_oracle_golden.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _holder(normalize=False):
    from efficientconformer_b200.encoders import _PreprocessingHolder
    return _PreprocessingHolder({"sample_rate": 16000, "n_fft": 512, "win_length_ms": 25, "hop_length_ms": 10, "n_mels": 80,
                                 "normalize": normalize, "mean": -5.6501, "std": 4.2280}).to(DEV)


def test_logmel_against_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "frontend_small.pt"))
    for name, c in g["logmel"].items():
        pre = _holder(c["normalize"])
        mel, mel_len = pre(c["audio"].to(DEV), c["audio_len"].to(DEV))
        ref = c["mel"]
        assert mel.shape == ref.shape and mel.dtype == torch.float32 and torch.equal(mel_len.cpu(), c["mel_len"])
        d = (mel.cpu().double() - ref.double()).abs().max().item()
        e = float((mel.cpu().double() - ref.double()).norm() / ref.double().norm())
        print(f"{name}: log-mel vs the reference module max abs {d:.3e} rel-L2 {e:.3e}")
        assert d < 1e-3 and e < 1e-5, (name, d, e)          # fp32 shared-memory FFT vs torch's CPU FFT; log domain


@pytest.mark.parametrize("B,L", [(32, 159840), (3, 16000), (1, 257), (2, 4001)])
def test_logmel_against_oracle(B, L):
    """Bench shape (32 x ~10 s -> 1000 frames), ragged frame counts off the 8-frame CTA grid, the shortest audio torch.stft accepts."""
    from oracle import frontend_oracle as FO
    g = torch.Generator().manual_seed(L)
    audio = torch.randn(B, L, generator=g) * 0.1
    audio[-1, L // 2:] = 0                                   # a zero-padded tail (collate_fn_pad): exact log(1e-9) frames
    nb = min(B, 4)                                           # the oracle's float64 frames of 4 utterances are enough at the big shape
    pre = _holder()
    mel, _ = pre(audio.to(DEV), None)
    assert mel.shape == (B, 80, L // 160 + 1)
    ref, _ = FO.logmel(audio[-nb:].numpy())
    got = mel[-nb:].cpu().double().numpy()
    d = np.abs(got - ref).max()
    e = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(f"B={B} L={L}: log-mel vs oracle max abs {d:.3e} rel-L2 {e:.3e}")
    assert d < 1e-3 and e < 1e-5
    # the torchaudio ops on the same CUDA tensor (what the reference runs on a GPU): same result
    os.environ["EFFCONF_DEVICE_FRONTEND"] = "0"
    try:
        lib_mel, _ = pre(audio.to(DEV), None)
    finally:
        del os.environ["EFFCONF_DEVICE_FRONTEND"]
    assert (lib_mel - mel).abs().max().item() < 1e-3


def test_logmel_rejects_bad_shapes():
    pre = _holder()
    with pytest.raises(RuntimeError, match="reflect"):
        pre(torch.zeros(1, 200, device=DEV), None)           # torch.stft refuses this too (reflect padding needs > n_fft / 2 samples)


def test_drop_in_forward_with_the_device_front_end(golden_dir):
    """reference contract `forward((audio, _, audio_len, _))` on the golden of the real reference: the device front end is the default."""
    from efficientconformer_b200 import ModelCTC, CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V
    from efficientconformer_b200.synthetic import seeded_state_dict, synthetic_audio
    g = torch.load(os.path.join(golden_dir, "ctc_small_audio_b2_t200.pt"))
    sd = seeded_state_dict(P, V, seed=0, prefix_encoder="encoder.")
    m = ModelCTC(P, {"vocab_size": V}, precision="bf16x2")
    m.load_state_dict(sd, strict=False)
    m = m.to(DEV).eval()
    audio = synthetic_audio(2, g["t_mel"], seed=g["audio_seed"]).to(DEV)
    logits, out_len, _ = m.forward((audio, None, g["audio_len"].to(DEV), None))
    e = float((logits.cpu().double() - g["logits"].double()).norm() / g["logits"].double().norm())
    print(f"audio -> logits with the one-kernel front end vs the reference: rel-L2 {e:.3e}")
    assert torch.equal(out_len.cpu(), g["out_len"]) and e < 1e-3


@pytest.mark.parametrize("with_len", [True, False])
def test_specaugment_kernel_is_bit_exact_against_the_oracle(with_len):
    from efficientconformer_b200 import ops
    from oracle import frontend_oracle as FO
    B, F, T = 6, 80, 333
    g = torch.Generator().manual_seed(3)
    mel = torch.randn(B, F, T, generator=g)
    lens = torch.tensor([333, 200, 77, 20, 1, 0]) if with_len else None
    seed = 1234567
    ctr = ops.dropout_counter(torch.device(DEV), seed)
    for step in range(1, 6):
        ops.dropout_advance(ctr)
        x = mel.to(DEV).clone()
        ops.specaugment_(x, None if lens is None else lens.to(DEV), 2, 27, 5, 0.05, ctr)
        ref = FO.specaugment_apply(mel.numpy(), None if lens is None else lens.numpy(), seed, step, 2, 27, 5, 0.05)
        assert torch.equal(x.cpu(), torch.from_numpy(ref)), step
        assert (x == 0).any()                                # something was masked


def test_specaugment_module_statistics_and_streams():
    """The module path: fresh masks every call, frequency masks shared by the batch, time masks inside the valid frames, widths
    distributed like torchaudio's (floor of U * param)."""
    from efficientconformer_b200.encoders import SpecAugment
    aug = SpecAugment(True, 2, 27, 5, 0.05)
    B, F, T = 8, 80, 1000
    lens = torch.tensor([1000, 900, 800, 700, 600, 500, 400, 300], device=DEV)
    x = torch.ones(B, F, T, device=DEV)
    f_widths, outs = [], []
    for _ in range(200):
        y = aug(x, lens)
        assert y.data_ptr() != x.data_ptr() and bool((x == 1).all())      # the input is not modified
        fmask = (y == 0).all(2)                                           # (B, F): rows masked over all frames
        assert bool((fmask == fmask[0:1]).all())                          # shared by the batch
        f_widths.append(int(fmask[0].sum()))
        tmask = (y == 0).all(1)                                           # (B, T)
        for b in range(B):
            assert not bool(tmask[b, int(lens[b]):].any())                # never past the valid frames
            assert int(tmask[b].sum()) <= 5 * int(0.05 * int(lens[b]))
        outs.append(y)
    assert any(not torch.equal(outs[0], o) for o in outs[1:])
    mean_w = sum(f_widths) / len(f_widths)                                # two masks of E[floor(27 U)] = 13 each, minus overlaps
    print("mean masked mel bins per step", mean_w)
    assert 18.0 < mean_w < 27.0


def test_device_batch_prefetcher_preserves_order_and_values_under_overlap():
    """reference models/model.py:229 `batch = [elt.to(device) for elt in batch]` replaced by the pinned, stream-overlapped pipeline:
    same batches, same order, while a compute stream is busy and the allocator recycles the staging memory."""
    from efficientconformer_b200 import DeviceBatchPrefetcher
    g = torch.Generator().manual_seed(5)
    host = [(torch.randn(4, 16000 + 160 * i, generator=g), torch.randint(1, 30, (4, 12), generator=g), torch.full((4,), 16000 + 160 * i),
             torch.full((4,), 12), "meta%d" % i) for i in range(9)]
    busy = torch.randn(2048, 2048, device=DEV)
    seen = 0
    for k, batch in enumerate(DeviceBatchPrefetcher(host, DEV, depth=2)):
        x, y, xl, yl, meta = batch
        for _ in range(3):
            busy = busy @ busy * 1e-3                       # keep the consumer stream behind the copy stream
        assert x.is_cuda and y.is_cuda and meta == "meta%d" % k
        assert torch.equal(x.cpu(), host[k][0]) and torch.equal(y.cpu(), host[k][1]) and torch.equal(xl.cpu(), host[k][2])
        seen += 1
    assert seen == len(host)
    with pytest.raises(RuntimeError, match="CUDA path only"):
        DeviceBatchPrefetcher(host, "cpu")
