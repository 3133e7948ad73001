"""CPU oracle for the device front end  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in numpy float64:
  * reference models/modules.py:87-106 AudioPreprocessing.forward.  The transforms it calls live in a third-party dependency that is not
    vendored under /root/reference (torchaudio; requirements.txt lists it unpinned, 2.11.0 is installed in the authoring container), so
    their published algorithms are restated:
      - torchaudio.transforms.Spectrogram(n_fft, win_length, hop_length) = torch.stft(center=True, pad_mode='reflect', periodic Hann
        window of win_length samples zero padded symmetrically to n_fft, onesided) -> |.|^2;
      - torchaudio.transforms.MelScale(n_mels, sample_rate, f_min=0, f_max=8000, n_stft) = torchaudio.functional.melscale_fbanks with
        the HTK mel scale, no area normalisation: triangular filters between n_mels + 2 points equally spaced in mel;
    then log(x + 1e-9), lengths x_len // hop + 1 and the optional (x - mean) / std.
  * reference models/modules.py:136-151 SpecAugment.forward with torchaudio.functional.mask_along_axis arithmetic
    (value = U * mask_param, min_value = U' * (size - value), start = long(min_value), end = start + long(value), cells [start, end) <- 0),
    drawing U, U' from the product's counter-based hash (csrc/ec_common.cuh splitmix64 / site_key, csrc/frontend.cu augment_span) instead
    of torch's generator: integer / single-precision arithmetic restated bit for bit.
Pinning: log-mel against the REAL reference module and the mask arithmetic against torchaudio's own mask_along_axis fed with preset
uniforms (tests/golden/make_golden_frontend.py imports both; fixture tests/golden/frontend_small.pt).  The random STREAM of SpecAugment is
the product's own (the reference uses torch's global generator): stream parity is not defined, distribution parity is by construction."""
import numpy as np

MASK64 = (1 << 64) - 1
AUGMENT_SITE = 0x5AE5A06


def hann_window_padded(win_length, n_fft):
    """torch.hann_window(win_length, periodic=True) centred in n_fft samples (torch.stft pads the window on both sides)."""
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
    left = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[left:left + win_length] = w
    return out


def melscale_fbanks_htk(n_freqs, f_min, f_max, n_mels, sample_rate):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk'): (n_freqs, n_mels)."""
    all_freqs = np.linspace(0.0, sample_rate // 2, n_freqs)
    hz_to_mel = lambda f: 2595.0 * np.log10(1.0 + f / 700.0)
    mel_to_hz = lambda m: 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    m_pts = np.linspace(hz_to_mel(f_min), hz_to_mel(f_max), n_mels + 2)
    f_pts = mel_to_hz(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]                          # (n_mels + 1)
    slopes = f_pts[None, :] - all_freqs[:, None]             # (n_freqs, n_mels + 2)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))


def logmel(audio, sample_rate=16000, n_fft=512, win_length_ms=25, hop_length_ms=10, n_mels=80, normalize=False, mean=0.0, std=1.0,
           audio_len=None):
    """reference models/modules.py:87-106.  audio (B, L) -> (mel (B, n_mels, L // hop + 1) float64, mel_len or None)."""
    audio = np.asarray(audio, dtype=np.float64)
    win_length = int(sample_rate * win_length_ms) // 1000
    hop = int(sample_rate * hop_length_ms) // 1000
    B, L = audio.shape
    T = L // hop + 1
    w = hann_window_padded(win_length, n_fft)
    padded = np.pad(audio, ((0, 0), (n_fft // 2, n_fft // 2)), mode="reflect")
    idx = np.arange(T)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = padded[:, idx] * w                              # (B, T, n_fft)
    power = np.abs(np.fft.rfft(frames, axis=-1)) ** 2        # (B, T, n_fft // 2 + 1)
    fb = melscale_fbanks_htk(n_fft // 2 + 1, 0.0, 8000.0, n_mels, sample_rate)
    mel = np.log(power @ fb + 1e-9).transpose(0, 2, 1)
    if normalize:
        mel = (mel - mean) / std
    mel_len = None if audio_len is None else np.asarray(audio_len) // hop + 1
    return mel, mel_len


# ---- SpecAugment ---------------------------------------------------------------------------------------------------------------
def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK64
    return x ^ (x >> 31)


def site_key(seed, step, site):
    return splitmix64(seed ^ ((step * 0xD1B54A32D192ED03) & MASK64)) ^ ((site * 0x9FB21C651E98DF25) & MASK64)


def span_from_uniforms(u1, u2, param, size):
    """torchaudio.functional.mask_along_axis: value = U * param; min_value = U' * (size - value); [long(min_value), + long(value))."""
    f = np.float32
    value = f(u1) * f(param)
    minv = f(u2) * (f(size) - value)
    start = int(minv)
    return start, start + int(value)


def draw_span(key, draw, param, size):
    h = splitmix64(key ^ ((draw * 0xA24BAED4963EE407) & MASK64))
    u1 = np.float32(h >> 40) * np.float32(1.0 / 16777216.0)
    u2 = np.float32((h >> 16) & 0xFFFFFF) * np.float32(1.0 / 16777216.0)
    return span_from_uniforms(u1, u2, param, size)


def specaugment_spans(seed, step, lens, n_mels, mF, F, mT, pS):
    """-> (frequency spans [(s, e)] * mF shared by the batch, time spans per utterance [[(s, e)] * mT] * B)."""
    key = site_key(seed, step, AUGMENT_SITE)
    fspans = [draw_span(key, i, F, n_mels) for i in range(mF)]
    tspans = []
    for b, ln in enumerate(lens):
        ln = int(ln)
        Tp = int(np.float32(pS) * np.float32(ln))            # int(pS * x_len[b]) in single precision, as the tensor product is
        tspans.append([draw_span(key, mF + b * mT + j, Tp, ln) for j in range(mT)])
    return fspans, tspans


def specaugment_apply(mel, lens, seed, step, mF, F, mT, pS):
    """reference models/modules.py:136-151 on mel (B, n_mels, T) with the counter-based draws; returns a masked copy."""
    out = np.array(mel, copy=True)
    B, n_mels, T = out.shape
    lens = [T] * B if lens is None else [min(max(int(v), 0), T) for v in lens]
    fspans, tspans = specaugment_spans(seed, step, lens, n_mels, mF, F, mT, pS)
    for s, e in fspans:
        out[:, max(s, 0):min(e, n_mels), :] = 0
    for b in range(B):
        for s, e in tspans[b]:
            out[b, :, max(s, 0):min(e, lens[b])] = 0
    return out
