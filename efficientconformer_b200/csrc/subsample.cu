// Conv2d(1 -> C, 3x3, stride 2, pad 1) + eval BatchNorm2d (folded) + Swish, written directly in the layout the
// following Linear consumes:  y[b, t, c*F/2 + f]   (reference models/modules.py:232-249 Conv2dSubsampling.forward,
// reshape (B,C,F/2,T/2)->(B,C*F/2,T/2) :245-247, then transpose + Linear at models/encoders.py:113-116).
//
// Producer kernel: materialises the GEMM A operand once (activation type); the K = C*F/2 GEMM follows.
// One CTA = 8 output frames of one utterance; thread = (output frequency f, frame).  The thread's 3x3 mel patch lives in
// registers for the whole channel loop; the BatchNorm-folded taps are read from shared memory as warp-wide broadcasts
// (3 x LDS.128 per channel), so each output costs 9 FFMA + Swish + one store, consecutive f -> consecutive addresses.
#include "ec_common.cuh"

namespace ec {

constexpr int kSubTT = 8;

template <typename T>
__global__ void __launch_bounds__(512) subsample_conv_kernel(const float* __restrict__ mel, const float* __restrict__ w, const float* __restrict__ bias,
                                                             int F, int T_in, int T_out, int C, T* __restrict__ y) {
  using Tr = ActTraits<T>;
  extern __shared__ __align__(16) float ss[];
  constexpr int TW = 2 * kSubTT + 1;            // staged mel frames
  float* ws = ss;                               // [C][12]: 9 taps, folded bias, 2 pad  (3 x float4)
  float* patch = ws + C * 12;                   // [(F+2)][TW+1], zero halo
  grid_dependency_wait();
  grid_launch_dependents();
  const int b = blockIdx.y, t0 = blockIdx.x * kSubTT;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  const float* melb = mel + static_cast<size_t>(b) * F * T_in;
  for (int i = tid; i < (F + 2) * TW; i += nthr) {
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
  const int f = threadIdx.x, tl = threadIdx.y;
  if (f >= F2 || t0 + tl >= T_out) return;
  float pv[9];
#pragma unroll
  for (int df = 0; df < 3; ++df)
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) pv[df * 3 + dt] = patch[(2 * f + df) * (TW + 1) + 2 * tl + dt];
  T* yo = y + (static_cast<size_t>(b) * T_out + t0 + tl) * (static_cast<size_t>(C) * F2) + f;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float4 w0 = *reinterpret_cast<const float4*>(ws + c * 12);
    const float4 w1 = *reinterpret_cast<const float4*>(ws + c * 12 + 4);
    const float4 w2 = *reinterpret_cast<const float4*>(ws + c * 12 + 8);
    float acc = w2.y;                                       // folded bias
    acc = fmaf(w0.x, pv[0], acc); acc = fmaf(w0.y, pv[1], acc); acc = fmaf(w0.z, pv[2], acc);
    acc = fmaf(w0.w, pv[3], acc); acc = fmaf(w1.x, pv[4], acc); acc = fmaf(w1.y, pv[5], acc);
    acc = fmaf(w1.z, pv[6], acc); acc = fmaf(w1.w, pv[7], acc); acc = fmaf(w2.x, pv[8], acc);
    yo[static_cast<size_t>(c) * F2] = Tr::to(acc * __fdividef(1.0f, 1.0f + __expf(-acc)));
  }
}

int launch_subsample_conv(int precision, const SubsampleArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.F % 2 == 0, "n_mels must be even");
  const int T_out = (a.T - 1) / 2 + 1;
  dim3 grid(cdiv(T_out, kSubTT), a.B);
  dim3 block(a.F / 2, kSubTT);
  EC_REQUIRE(block.x * block.y <= 512, "n_mels too large for the subsampling kernel");
  const size_t smem = sizeof(float) * ((a.F + 2) * (2 * kSubTT + 2) + a.C * 12);
  EC_REQUIRE(smem <= 48 * 1024, "subsampling patch does not fit in shared memory");
  if (precision == EC_PREC_TF32)
    return launch_pdl(subsample_conv_kernel<float>, grid, block, smem, stream, a.mel, a.w, a.b, a.F, a.T, T_out, a.C,
                      reinterpret_cast<float*>(a.y));
  if (precision == EC_PREC_BF16)
    return launch_pdl(subsample_conv_kernel<__nv_bfloat16>, grid, block, smem, stream, a.mel, a.w, a.b, a.F, a.T, T_out, a.C,
                      reinterpret_cast<__nv_bfloat16*>(a.y));
  EC_FAIL("unknown precision");
}

}  // namespace ec
