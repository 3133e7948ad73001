// Conv2d(1 -> C, 3x3, stride 2, pad 1) + eval BatchNorm2d (folded) + Swish, written directly in the layout the
// following Linear consumes:  y[b, t, c*F/2 + f]   (reference models/modules.py:232-249 Conv2dSubsampling.forward,
// reshape (B,C,F/2,T/2)->(B,C*F/2,T/2) :245-247, then transpose + Linear at models/encoders.py:113-116).
//
// v1: producer kernel materialises the GEMM A operand once (activation type); the K = C*F/2 GEMM follows.
// One CTA = 16 output frames of one utterance: the (F+2) x 33 mel patch and all taps sit in shared memory,
// threads sweep the (frame, feature) outputs with the feature index fastest so stores are fully coalesced.
#include "ec_common.cuh"

namespace ec {

constexpr int kSubTT = 16;

template <typename T>
__global__ void __launch_bounds__(256) subsample_conv_kernel(const float* __restrict__ mel, const float* __restrict__ w, const float* __restrict__ bias,
                                                             int F, int T_in, int T_out, int C, T* __restrict__ y) {
  using Tr = ActTraits<T>;
  extern __shared__ float ss[];
  constexpr int TW = 2 * kSubTT + 1;           // staged mel frames
  float* patch = ss;                            // [(F+2)][TW+1]
  float* ws = patch + (F + 2) * (TW + 1);       // [C][9]
  float* bs = ws + C * 9;                       // [C]
  const int b = blockIdx.y, t0 = blockIdx.x * kSubTT;
  const int tid = threadIdx.x;
  const float* melb = mel + static_cast<size_t>(b) * F * T_in;
  for (int i = tid; i < (F + 2) * TW; i += 256) {
    const int fr = i / TW, tc = i % TW;
    const int f = fr - 1, t = 2 * t0 - 1 + tc;
    patch[fr * (TW + 1) + tc] = (f >= 0 && f < F && t >= 0 && t < T_in) ? __ldg(melb + static_cast<size_t>(f) * T_in + t) : 0.f;
  }
  for (int i = tid; i < C * 9; i += 256) ws[i] = w[i];
  for (int i = tid; i < C; i += 256) bs[i] = bias[i];
  __syncthreads();
  const int F2 = F / 2, feat = C * F2;
  const int n_t = min(kSubTT, T_out - t0);
  for (int i = tid; i < n_t * feat; i += 256) {
    const int tl = i / feat, col = i % feat;
    const int c = col / F2, f = col % F2;
    const float* pw = ws + c * 9;
    const float* pp = patch + (2 * f) * (TW + 1) + 2 * tl;
    float acc = bs[c];
#pragma unroll
    for (int df = 0; df < 3; ++df)
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) acc = fmaf(pw[df * 3 + dt], pp[df * (TW + 1) + dt], acc);
    y[(static_cast<size_t>(b) * T_out + t0 + tl) * feat + col] = Tr::to(swishf_(acc));
  }
}

int launch_subsample_conv(int precision, const SubsampleArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.F % 2 == 0, "n_mels must be even");
  const int T_out = (a.T - 1) / 2 + 1;
  dim3 grid(cdiv(T_out, kSubTT), a.B);
  const size_t smem = sizeof(float) * ((a.F + 2) * (2 * kSubTT + 2) + a.C * 10);
  EC_REQUIRE(smem <= 48 * 1024, "subsampling patch does not fit in shared memory");
  if (precision == EC_PREC_TF32)
    subsample_conv_kernel<float><<<grid, 256, smem, stream>>>(a.mel, a.w, a.b, a.F, a.T, T_out, a.C, reinterpret_cast<float*>(a.y));
  else if (precision == EC_PREC_BF16)
    subsample_conv_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(a.mel, a.w, a.b, a.F, a.T, T_out, a.C, reinterpret_cast<__nv_bfloat16*>(a.y));
  else EC_FAIL("unknown precision");
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec
