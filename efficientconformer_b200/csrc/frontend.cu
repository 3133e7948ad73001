// Front end on the device:
//   * log-mel features: reference models/modules.py:87-106 AudioPreprocessing.forward =
//         torchaudio Spectrogram(n_fft, win_length, hop, power 2, center / reflect pad, one-sided)  ->  MelScale(fb [n_freq, n_mels])
//         -> log(x + 1e-9) -> optional (x - mean) / std
//     as ONE kernel: framing with reflect padding, window, radix-2 Stockham FFT in shared memory (natural order in and out, no
//     bit-reversed scatter: its 16-way bank conflicts were 60 % of the first version's shared-memory wavefronts; two real frames per
//     complex transform),
//     power spectrum, mel projection, log, normalisation; [B, L] audio in, [B, n_mels, T] fp32 out, nothing else touches HBM
//     (torchaudio on a CUDA tensor runs pad + as_strided + cuFFT + abs/pow + matmul + add + log: seven launches and ~20x the bytes);
//   * SpecAugment: reference models/modules.py:136-151 (mF frequency masks shared by the batch, mT time masks per utterance inside its
//     valid frames; torchaudio mask_along_axis arithmetic: value = U * param, start = floor(U' * (size - value)), end = start +
//     floor(value)).  The reference loops over the batch in Python with one host read of x_len[b] per utterance; here one launch draws
//     all masks from the counter-based hash of ec_common.cuh (the {seed, step} pair lives on the device, so a captured graph draws
//     fresh masks on every replay) and writes ONLY the masked cells.
#include "ec_common.cuh"

namespace ec {

namespace {
constexpr int kFeFrames = 8;                 // frames per CTA (4 complex transforms): 32-byte output runs per mel bin
constexpr unsigned kAugmentSite = 0x5AE5A06u;

__device__ __forceinline__ int reflect_index(int i, int L) {
  if (i < 0) i = -i;
  if (i >= L) i = 2 * (L - 1) - i;
  return i;
}
}  // namespace

// grid (ceil(T / 8), B), block N / 2 threads; dynamic shared memory: z[2][2][N] (ping-pong re / im) twc[N/2] tws[N/2] pw[2][N/2 + 1] mel[n_mels][8]
__global__ void logmel_kernel(const float* __restrict__ audio, int L, int T, int hop, int N, int log2n, const float* __restrict__ window,
                              const float* __restrict__ fb, const int* __restrict__ krange, int n_mels, float scale, float shift,
                              float* __restrict__ out) {
  extern __shared__ float fe_smem[];
  const int H = N / 2, NF = H + 1;
  float* zbuf = fe_smem;                     // buffer p: re at zbuf + 2 p N, im at zbuf + (2 p + 1) N
  float* twc = zbuf + 4 * N;
  float* tws = twc + H;
  float* pw = tws + H;                       // [2][NF]
  float* melbuf = pw + 2 * NF;               // [n_mels][kFeFrames]
  const int tid = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * kFeFrames;
  const float* x = audio + static_cast<size_t>(b) * L;
  {
    float s, c;
    sincospif(-2.0f * static_cast<float>(tid) / static_cast<float>(N), &s, &c);
    twc[tid] = c; tws[tid] = s;
  }
  for (int pair = 0; pair < kFeFrames / 2; ++pair) {
    const int ta = t0 + 2 * pair, tb = ta + 1;
    __syncthreads();                         // previous pair's spectrum / power fully consumed
    // windowed frames in natural order: frame a -> real part, frame b -> imaginary part
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = tid + h * H;
      const float w = __ldg(window + i);
      zbuf[i] = ta < T ? w * __ldg(x + reflect_index(ta * hop + i - H, L)) : 0.f;
      zbuf[N + i] = tb < T ? w * __ldg(x + reflect_index(tb * hop + i - H, L)) : 0.f;
    }
    __syncthreads();
    // Stockham autosort, decimation in time: stage with sub-transform length Ns reads (j, j + N/2) and writes (j0, j0 + Ns)
    int cur = 0;
    for (int s = 0; s < log2n; ++s) {
      const int Ns = 1 << s;
      const int k = tid & (Ns - 1), j0 = ((tid - k) << 1) + k;
      const int tw = k << (log2n - 1 - s);
      const float c = twc[tw], sn = tws[tw];
      const float* ir = zbuf + 2 * cur * N;
      const float* ii = ir + N;
      float* orr = zbuf + 2 * (cur ^ 1) * N;
      float* oi = orr + N;
      const float ar = ir[tid], ai = ii[tid], br = ir[tid + H], bi = ii[tid + H];
      const float pr = br * c - bi * sn, pi = br * sn + bi * c;
      orr[j0] = ar + pr; oi[j0] = ai + pi;
      orr[j0 + Ns] = ar - pr; oi[j0 + Ns] = ai - pi;
      cur ^= 1;
      __syncthreads();
    }
    const float* zr = zbuf + 2 * cur * N;
    const float* zi = zr + N;
    // X_a[k] = (Z[k] + conj(Z[N-k])) / 2,  X_b[k] = (Z[k] - conj(Z[N-k])) / (2i);  power = |X|^2
    for (int k = tid; k < NF; k += H) {
      const int n = (N - k) & (N - 1);
      const float ar = 0.5f * (zr[k] + zr[n]), ai = 0.5f * (zi[k] - zi[n]);
      const float br = 0.5f * (zi[k] + zi[n]), bi = 0.5f * (zr[n] - zr[k]);
      pw[k] = ar * ar + ai * ai;
      pw[NF + k] = br * br + bi * bi;
    }
    __syncthreads();
    for (int q = tid; q < 2 * n_mels; q += H) {
      const int which = q >= n_mels ? 1 : 0, m = q - which * n_mels;
      const float* p = pw + which * NF;
      // filter m is a triangle over the bins [krange[2m], krange[2m+1]) (zero elsewhere): same terms, same order as the dense product
      const int k_lo = krange ? max(krange[2 * m], 0) : 0, k_hi = krange ? min(krange[2 * m + 1], NF) : NF;
      float acc = 0.f;
      for (int k = k_lo; k < k_hi; ++k) acc = fmaf(__ldg(fb + static_cast<size_t>(k) * n_mels + m), p[k], acc);
      melbuf[m * kFeFrames + 2 * pair + which] = logf(acc + 1e-9f) * scale + shift;
    }
  }
  __syncthreads();
  for (int q = tid; q < n_mels * kFeFrames; q += H) {
    const int m = q / kFeFrames, j = q - m * kFeFrames;
    if (t0 + j < T) out[(static_cast<size_t>(b) * n_mels + m) * T + t0 + j] = melbuf[q];
  }
}

static int launch_logmel(const float* audio, int B, int L, int n_fft, int hop, const float* window, const float* fb, const int* krange, int n_mels,
                         int normalize, float mean, float stdv, float* out, cudaStream_t stream) {
  EC_REQUIRE(audio && window && fb && out && B > 0 && hop > 0 && n_mels > 0, "logmel: bad argument");
  EC_REQUIRE(n_fft >= 64 && n_fft <= 2048 && (n_fft & (n_fft - 1)) == 0, "logmel: n_fft must be a power of two in [64, 2048]");
  EC_REQUIRE(L > n_fft / 2, "logmel: reflect padding needs more than n_fft / 2 samples (as torch.stft does)");
  EC_REQUIRE(!normalize || stdv != 0.f, "logmel: std must not be zero");
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  const int T = L / hop + 1;
  const size_t smem = (static_cast<size_t>(5) * n_fft + 2 * (n_fft / 2 + 1) + static_cast<size_t>(n_mels) * kFeFrames) * sizeof(float);
  EC_REQUIRE(smem <= 48 * 1024, "logmel: n_mels too large for the shared-memory tile");
  const float scale = normalize ? 1.f / stdv : 1.f, shift = normalize ? -mean / stdv : 0.f;
  dim3 grid(cdiv(T, kFeFrames), B);
  logmel_kernel<<<grid, n_fft / 2, smem, stream>>>(audio, L, T, hop, n_fft, log2n, window, fb, krange, n_mels, scale, shift, out);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---- SpecAugment ----------------------------------------------------------------------------------------------------
// Draw i of the step: two 24-bit uniforms from one hash.  Draws 0 .. mF-1 are the frequency masks (shared by the batch), draw
// mF + b*mT + j is time mask j of utterance b.  All float arithmetic is single, unfused operations so that the CPU restatement
// used by the tests reproduces the mask boundaries bit for bit.
__device__ __forceinline__ void augment_span(unsigned long long key, unsigned draw, int param, int size, int* start, int* end) {
  const unsigned long long h = splitmix64(key ^ (static_cast<unsigned long long>(draw) * 0xA24BAED4963EE407ull));
  const float u1 = static_cast<float>(static_cast<unsigned>(h >> 40)) * (1.f / 16777216.f);
  const float u2 = static_cast<float>(static_cast<unsigned>(h >> 16) & 0xFFFFFFu) * (1.f / 16777216.f);
  const float value = __fmul_rn(u1, static_cast<float>(param));
  const float minv = __fmul_rn(u2, __fsub_rn(static_cast<float>(size), value));
  *start = static_cast<int>(minv);
  *end = *start + static_cast<int>(value);
}

// grid (B), block 256.  mel [B, F, T] in place.
__global__ void specaugment_kernel(float* __restrict__ mel, const long long* __restrict__ x_len, int F, int T, int mF, int Fp, int mT,
                                   float pS, const unsigned long long* __restrict__ ctr) {
  extern __shared__ int aug_spans[];         // [mF + mT][2]
  const int b = blockIdx.x;
  const unsigned long long key = site_key(ctr, kAugmentSite);
  int len = x_len ? static_cast<int>(x_len[b]) : T;
  len = len < 0 ? 0 : (len > T ? T : len);
  if (threadIdx.x < mF) {
    augment_span(key, threadIdx.x, Fp, F, &aug_spans[2 * threadIdx.x], &aug_spans[2 * threadIdx.x + 1]);
  } else if (threadIdx.x < mF + mT) {
    const int j = threadIdx.x - mF;
    const int Tp = static_cast<int>(__fmul_rn(pS, static_cast<float>(len)));         // int(pS * x_len[b])
    augment_span(key, mF + b * mT + j, Tp, len, &aug_spans[2 * threadIdx.x], &aug_spans[2 * threadIdx.x + 1]);
  }
  __syncthreads();
  float* m = mel + static_cast<size_t>(b) * F * T;
  for (int i = 0; i < mF; ++i) {
    const int s = max(aug_spans[2 * i], 0), e = min(aug_spans[2 * i + 1], F);
    const int n = (e - s) * T;
    for (int q = threadIdx.x; q < n; q += blockDim.x) m[static_cast<size_t>(s) * T + q] = 0.f;
  }
  for (int i = mF; i < mF + mT; ++i) {
    const int s = max(aug_spans[2 * i], 0), e = min(aug_spans[2 * i + 1], len);
    const int w = e - s;
    if (w <= 0) continue;
    for (int q = threadIdx.x; q < w * F; q += blockDim.x) {
      const int f = q / w, t = s + q - f * w;
      m[static_cast<size_t>(f) * T + t] = 0.f;
    }
  }
}

static int launch_specaugment(float* mel, const long long* x_len, int B, int F, int T, int mF, int Fparam, int mT, float pS,
                       const unsigned long long* counter, cudaStream_t stream) {
  EC_REQUIRE(mel && counter && B > 0 && F > 0 && T > 0 && mF >= 0 && mT >= 0 && Fparam >= 0, "specaugment: bad argument");
  EC_REQUIRE(mF + mT <= 256, "specaugment: at most 256 masks per utterance");
  EC_REQUIRE(Fparam <= F, "specaugment: frequency mask parameter exceeds the number of mel bins");
  if (mF + mT == 0) return EC_OK;
  specaugment_kernel<<<B, 256, static_cast<size_t>(2) * (mF + mT) * sizeof(int), stream>>>(mel, x_len, F, T, mF, Fparam, mT, pS, counter);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec

#define EC_ST(s) reinterpret_cast<cudaStream_t>(s)
extern "C" {
int ec_op_logmel(const float* audio, int batch, int samples, int n_fft, int hop, const float* window, const float* fb, const int* krange,
                 int n_mels, int normalize, float mean, float stdv, float* out, void* stream) {
  return ec::launch_logmel(audio, batch, samples, n_fft, hop, window, fb, krange, n_mels, normalize, mean, stdv, out, EC_ST(stream));
}
int ec_op_specaugment(float* mel, const long long* x_len, int batch, int n_mels, int t, int mF, int F, int mT, float pS,
                      const unsigned long long* counter, void* stream) {
  return ec::launch_specaugment(mel, x_len, batch, n_mels, t, mF, F, mT, pS, counter, EC_ST(stream));
}
}
