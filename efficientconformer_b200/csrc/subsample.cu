// Conv2d(1 -> C, 3x3, stride 2, pad 1) + eval BatchNorm2d (folded) + Swish, written directly in the layout the
// following Linear consumes:  y[b, t, c*F/2 + f]   (reference models/modules.py:232-249 Conv2dSubsampling.forward,
// reshape (B,C,F/2,T/2)->(B,C*F/2,T/2) :245-247, then transpose + Linear at models/encoders.py:113-116).
//
// Producer kernel: materialises the GEMM A operand once (activation type); the K = C*F/2 GEMM follows.
// One CTA = 8 output frames of one utterance; thread = (pair of adjacent output frequencies, frame).  The thread's 5x3 mel
// patch lives in registers for the whole channel loop; the BatchNorm-folded taps are read from shared memory as warp-wide
// broadcasts (3 x LDS.128 per channel, shared by both outputs), so an output costs 9 FFMA + Swish + half a packed store.
#include "ec_common.cuh"
#include <algorithm>

namespace ec {

constexpr int kSubTT = 8;

// kRaw (training forward): plain convolution output in fp32, no BatchNorm fold, no Swish, no operand rounding -- the batch
// statistics are taken over this tensor before the normalisation kernel runs.
template <typename T, bool kRaw = false>
__global__ void __launch_bounds__(512) subsample_conv_kernel(const float* __restrict__ mel, const float* __restrict__ w, const float* __restrict__ bias,
                                                             int F, int T_in, int T_out, int C, T* __restrict__ y) {
  using Tr = ActTraits<T>;
  extern __shared__ __align__(16) float ss[];
  constexpr int TW = 2 * kSubTT + 1;            // staged mel frames
  float* ws = ss;                               // [C][12]: 9 taps, folded bias, 2 pad  (3 x float4)
  float* patch = ws + C * 12;                   // [(F+3)][TW+1], zero halo (one spare row for the odd-F2 pair)
  grid_dependency_wait();
  grid_launch_dependents();
  const int b = blockIdx.y, t0 = blockIdx.x * kSubTT;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  const float* melb = mel + static_cast<size_t>(b) * F * T_in;
  for (int i = tid; i < (F + 3) * TW; i += nthr) {
    const int fr = i / TW, tc = i % TW;
    const int f = fr - 1, t = 2 * t0 - 1 + tc;
    patch[fr * (TW + 1) + tc] = (f >= 0 && f < F && t >= 0 && t < T_in) ? __ldg(melb + static_cast<size_t>(f) * T_in + t) : 0.f;
  }
  for (int i = tid; i < C * 12; i += nthr) {
    const int c = i / 12, k = i % 12;
    ws[i] = k < 9 ? w[c * 9 + k] : (k == 9 ? bias[c] : 0.f);
  }
  __syncthreads();
  const int F2 = F / 2;
  // thread = (pair of adjacent output frequencies, frame): 5 x 3 mel patch in registers, taps broadcast from smem,
  // two outputs per channel written as one packed store
  const int fp = threadIdx.x, tl = threadIdx.y;
  const int f = 2 * fp;
  if (f >= F2 || t0 + tl >= T_out) return;
  float pv[15];
#pragma unroll
  for (int df = 0; df < 5; ++df)
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) pv[df * 3 + dt] = patch[(2 * f + df) * (TW + 1) + 2 * tl + dt];
  T* yo = y + (static_cast<size_t>(b) * T_out + t0 + tl) * (static_cast<size_t>(C) * F2) + f;
  const bool pair_ok = (f + 1 < F2) && (F2 % 2 == 0);
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float4 w0 = *reinterpret_cast<const float4*>(ws + c * 12);
    const float4 w1 = *reinterpret_cast<const float4*>(ws + c * 12 + 4);
    const float4 w2 = *reinterpret_cast<const float4*>(ws + c * 12 + 8);
    float a0 = w2.y, a1 = w2.y;                             // folded bias
    a0 = fmaf(w0.x, pv[0], a0); a0 = fmaf(w0.y, pv[1], a0); a0 = fmaf(w0.z, pv[2], a0);
    a0 = fmaf(w0.w, pv[3], a0); a0 = fmaf(w1.x, pv[4], a0); a0 = fmaf(w1.y, pv[5], a0);
    a0 = fmaf(w1.z, pv[6], a0); a0 = fmaf(w1.w, pv[7], a0); a0 = fmaf(w2.x, pv[8], a0);
    a1 = fmaf(w0.x, pv[6], a1); a1 = fmaf(w0.y, pv[7], a1); a1 = fmaf(w0.z, pv[8], a1);
    a1 = fmaf(w0.w, pv[9], a1); a1 = fmaf(w1.x, pv[10], a1); a1 = fmaf(w1.y, pv[11], a1);
    a1 = fmaf(w1.z, pv[12], a1); a1 = fmaf(w1.w, pv[13], a1); a1 = fmaf(w2.x, pv[14], a1);
    const float o0 = kRaw ? a0 : swish_fn<T>(a0), o1 = kRaw ? a1 : swish_fn<T>(a1);
    T* dst = yo + static_cast<size_t>(c) * F2;
    if constexpr (kRaw) {
      dst[0] = o0;
      if (f + 1 < F2) dst[1] = o1;
    } else if (pair_ok) {
      if constexpr (sizeof(T) == 2) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(o0, o1);
      else *reinterpret_cast<float2*>(dst) = make_float2(Tr::word(o0), Tr::word(o1));
    } else {
      dst[0] = Tr::to(o0);
      if (f + 1 < F2) dst[1] = Tr::to(o1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Two-layer front end of the Conformer family (subsampling_filters [C, C2]; reference models/modules.py:226-249 with
// num_layers == 2).  Layer 0 writes CHANNELS-LAST  y0[b, t, f, c]  so that the 3x3 patches of layer 1 are contiguous C-vectors;
// layer 1 (Conv2d(C -> C2, 3x3, stride 2) + BatchNorm2d + Swish: 9*C MACs per output, a real dense contraction) runs on the
// tcgen05 GEMM: im2col rows (b, t2, f2) x K = (kh, kw, c), folded weight permuted to the same K order at prepare time, Swish in
// the GEMM epilogue.  Its output [B*T2*F2, C2] viewed as [B*T2, F2*C2] is the operand of the Linear, whose weight columns are
// permuted from the reference's feature order c*F2 + f to f*C2 + c.
// ---------------------------------------------------------------------------------------------------------------
// thread = 8 consecutive channels (taps in registers), walking (frame, frequency) positions; the 3x3 mel patch is a
// shared-memory broadcast; one 16-byte (bf16) or two 16-byte (fp32) stores per position, coalesced over channels.
template <typename T>
__global__ void __launch_bounds__(256) subsample_conv_cl_kernel(const float* __restrict__ mel, const float* __restrict__ w,
                                                                const float* __restrict__ bias, int F, int T_in, int T_out, int C,
                                                                T* __restrict__ y) {
  using Tr = ActTraits<T>;
  extern __shared__ __align__(16) float ss[];
  constexpr int TW = 2 * kSubTT + 1;
  float* patch = ss;                             // [(F+2)][TW+1], zero halo
  const int b = blockIdx.y, t0 = blockIdx.x * kSubTT;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  const int cg = threadIdx.x;                    // channels [8cg, 8cg + 8)
  float wr[8][9], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[j][k] = __ldg(w + (8 * cg + j) * 9 + k);
    br[j] = __ldg(bias + 8 * cg + j);
  }
  grid_dependency_wait();
  grid_launch_dependents();
  const float* melb = mel + static_cast<size_t>(b) * F * T_in;
  for (int i = tid; i < (F + 2) * TW; i += nthr) {
    const int fr = i / TW, tc = i % TW;
    const int f = fr - 1, t = 2 * t0 - 1 + tc;
    patch[fr * (TW + 1) + tc] = (f >= 0 && f < F && t >= 0 && t < T_in) ? __ldg(melb + static_cast<size_t>(f) * T_in + t) : 0.f;
  }
  __syncthreads();
  const int F2 = F / 2;
  for (int idx = threadIdx.y; idx < kSubTT * F2; idx += blockDim.y) {
    const int tl = idx / F2, f = idx - tl * F2;
    if (t0 + tl >= T_out) break;
    float pv[9];
#pragma unroll
    for (int df = 0; df < 3; ++df)
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) pv[df * 3 + dt] = patch[(2 * f + df) * (TW + 1) + 2 * tl + dt];
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = br[j];
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fmaf(wr[j][k], pv[k], a);
      o[j] = swish_fn<T>(a);
    }
    T* dst = y + ((static_cast<size_t>(b) * T_out + t0 + tl) * F2 + f) * C + 8 * cg;
    if constexpr (sizeof(T) == 2) {
      uint4 pk;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(dst) = pk;
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(Tr::word(o[0]), Tr::word(o[1]), Tr::word(o[2]), Tr::word(o[3]));
      *reinterpret_cast<float4*>(dst + 4) = make_float4(Tr::word(o[4]), Tr::word(o[5]), Tr::word(o[6]), Tr::word(o[7]));
    }
  }
}

int launch_subsample_conv_cl(int precision, const SubsampleArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.F % 2 == 0, "n_mels must be even");
  EC_REQUIRE(a.C % 8 == 0 && a.C <= 2048, "the channels-last subsampling kernel needs C % 8 == 0");
  const int T_out = (a.T - 1) / 2 + 1;
  const int cg = a.C / 8;
  dim3 grid(cdiv(T_out, kSubTT), a.B);
  dim3 block(cg, std::max(1, 256 / cg));
  const size_t smem = sizeof(float) * ((a.F + 2) * (2 * kSubTT + 2));
  EC_REQUIRE(smem <= 48 * 1024, "subsampling patch does not fit in shared memory");
  EC_DISPATCH_PREC(precision, return launch_pdl(subsample_conv_cl_kernel<ActT>, grid, block, smem, stream, a.mel, a.w, a.b, a.F, a.T, T_out, a.C,
                                                reinterpret_cast<ActT*>(a.y)));
}

// im2col of a channels-last map for a 3x3 / stride 2 / pad 1 convolution: row (b, t2, f2), column (kh*3 + kw)*C + c holds
// y0[b, 2*t2 - 1 + kw, 2*f2 - 1 + kh, c] (kh walks frequency = Conv2d height, kw walks time = width), zero outside the map.
template <typename T>
__global__ void __launch_bounds__(256) im2col_3x3s2_kernel(const T* __restrict__ y0, int B, int T1, int F1, int C, int T2, int F2,
                                                           T* __restrict__ A) {
  constexpr int VEC = 16 / sizeof(T);
  grid_dependency_wait();
  grid_launch_dependents();
  const int cv = C / VEC;
  const size_t total = static_cast<size_t>(B) * T2 * F2 * 9 * cv;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = static_cast<int>(i % cv);
    size_t r = i / cv;
    const int k = static_cast<int>(r % 9);
    const size_t m = r / 9;
    const int f2 = static_cast<int>(m % F2);
    r = m / F2;
    const int t2 = static_cast<int>(r % T2), b = static_cast<int>(r / T2);
    const int kh = k / 3, kw = k - 3 * kh;
    const int f = 2 * f2 - 1 + kh, t = 2 * t2 - 1 + kw;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (f >= 0 && f < F1 && t >= 0 && t < T1)
      val = *reinterpret_cast<const uint4*>(y0 + ((static_cast<size_t>(b) * T1 + t) * F1 + f) * C + v * VEC);
    *reinterpret_cast<uint4*>(A + (m * 9 + k) * C + v * VEC) = val;
  }
}

int launch_im2col_3x3s2(int precision, const void* y0, int B, int T1, int F1, int C, void* A, cudaStream_t stream) {
  const int T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  const int vec = static_cast<int>(16 / act_esize(precision));
  const size_t vecs = static_cast<size_t>(B) * T2 * F2 * 9 * (C / vec);
  const int blocks = static_cast<int>(std::min<size_t>((vecs + 255) / 256, 148 * 32));
  EC_REQUIRE(C % vec == 0, "im2col needs whole 16-byte channel vectors");
  EC_DISPATCH_PREC(precision, return launch_pdl(im2col_3x3s2_kernel<ActT>, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const ActT*>(y0), B, T1,
                                                F1, C, T2, F2, reinterpret_cast<ActT*>(A)));
}

// Weight preparation of the second layer: eval BatchNorm2d folded into the taps and bias, K order (kh, kw, c_in).
template <typename T>
__global__ void conv2_weight_prep_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* g, const float* beta,
                                         const float* rm, const float* rv, float eps, int C2, int C, T* __restrict__ w_out,
                                         float* __restrict__ b_out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(C2) * 9 * C) return;
  const int co = static_cast<int>(i / (9 * C)), r = static_cast<int>(i % (9 * C)), k = r / C, ci = r % C;
  const float s = g[co] / sqrtf(rv[co] + eps);
  const T v = ActTraits<T>::to(w[(static_cast<size_t>(co) * C + ci) * 9 + k] * s);
  w_out[i] = v;
  if constexpr (IsSplit<T>::value) w_out[static_cast<size_t>(C2) * 9 * C + i] = SplitBf16{split_swap(v.bits)};   // swapped plane
  if (r == 0) b_out[co] = (b[co] - rm[co]) * s + beta[co];
}
int launch_conv2_weight_prep(int precision, const float* w, const float* b, const float* g, const float* beta, const float* rm,
                             const float* rv, float eps, int C2, int C, void* w_out, float* b_out, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(C2) * 9 * C;
  const int blocks = static_cast<int>((n + 255) / 256);
  EC_DISPATCH_PREC(precision, (conv2_weight_prep_kernel<ActT><<<blocks, 256, 0, stream>>>(w, b, g, beta, rm, rv, eps, C2, C, reinterpret_cast<ActT*>(w_out), b_out)));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// Linear weight [D, C*Fq] with reference feature order c*Fq + f  ->  [D, Fq*Cp] with order f*Cp + c (channels-last operand; Cp >= C,
// zero columns for the pad channels c >= C).
template <typename T>
__global__ void linear_weight_permute_kernel(const float* __restrict__ w, int D, int C, int Cp, int Fq, T* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t K = static_cast<size_t>(C) * Fq, Kp = static_cast<size_t>(Cp) * Fq;
  if (i >= D * Kp) return;
  const size_t d = i / Kp, r = i % Kp;
  const int f = static_cast<int>(r / Cp), c = static_cast<int>(r % Cp);
  const T v = ActTraits<T>::to(c < C ? w[d * K + static_cast<size_t>(c) * Fq + f] : 0.f);
  out[i] = v;
  if constexpr (IsSplit<T>::value) out[D * Kp + i] = SplitBf16{split_swap(v.bits)};   // swapped plane
}
int launch_linear_weight_permute(int precision, const float* w, int D, int C, int Fq, void* out, cudaStream_t stream, int Cp) {
  if (Cp < C) Cp = C;
  const size_t n = static_cast<size_t>(D) * Cp * Fq;
  const int blocks = static_cast<int>((n + 255) / 256);
  EC_DISPATCH_PREC(precision, (linear_weight_permute_kernel<ActT><<<blocks, 256, 0, stream>>>(w, D, C, Cp, Fq, reinterpret_cast<ActT*>(out))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// training forward: raw fp32 convolution (raw taps / bias), y [B, T_out, C*F/2]
int launch_subsample_conv_raw(const SubsampleArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.F % 2 == 0, "n_mels must be even");
  const int T_out = (a.T - 1) / 2 + 1;
  dim3 grid(cdiv(T_out, kSubTT), a.B);
  dim3 block((a.F / 2 + 1) / 2, kSubTT);
  EC_REQUIRE(block.x * block.y <= 512, "n_mels too large for the subsampling kernel");
  const size_t smem = sizeof(float) * ((a.F + 3) * (2 * kSubTT + 2) + a.C * 12);
  EC_REQUIRE(smem <= 48 * 1024, "subsampling patch does not fit in shared memory");
  return launch_pdl(subsample_conv_kernel<float, true>, grid, block, smem, stream, a.mel, a.w, a.b, a.F, a.T, T_out, a.C,
                    reinterpret_cast<float*>(a.y));
}

int launch_subsample_conv(int precision, const SubsampleArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.F % 2 == 0, "n_mels must be even");
  const int T_out = (a.T - 1) / 2 + 1;
  dim3 grid(cdiv(T_out, kSubTT), a.B);
  dim3 block((a.F / 2 + 1) / 2, kSubTT);
  EC_REQUIRE(block.x * block.y <= 512, "n_mels too large for the subsampling kernel");
  const size_t smem = sizeof(float) * ((a.F + 3) * (2 * kSubTT + 2) + a.C * 12);
  EC_REQUIRE(smem <= 48 * 1024, "subsampling patch does not fit in shared memory");
  EC_DISPATCH_PREC(precision, return launch_pdl(subsample_conv_kernel<ActT, false>, grid, block, smem, stream, a.mel, a.w, a.b, a.F, a.T, T_out,
                                                a.C, reinterpret_cast<ActT*>(a.y)));
}

}  // namespace ec
