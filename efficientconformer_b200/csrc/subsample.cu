// Conv2d(1 -> C, 3x3, stride 2, pad 1) + eval BatchNorm2d (folded) + Swish, written directly in the layout the
// following Linear consumes:  y[b, t, c*F/2 + f]   (reference models/modules.py:232-249 Conv2dSubsampling.forward,
// reshape (B,C,F/2,T/2)->(B,C*F/2,T/2) :245-247, then transpose + Linear at models/encoders.py:113-116).
//
// Producer kernel: materialises the GEMM A operand once (activation type); the K = C*F/2 GEMM follows.
// One CTA = 8 output frames of one utterance; thread = (pair of adjacent output frequencies, frame).  The thread's 5x3 mel
// patch lives in registers for the whole channel loop; the BatchNorm-folded taps are read from shared memory as warp-wide
// broadcasts (3 x LDS.128 per channel, shared by both outputs), so an output costs 9 FFMA + Swish + half a packed store.
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
    const float o0 = swish_fn<T>(a0), o1 = swish_fn<T>(a1);
    T* dst = yo + static_cast<size_t>(c) * F2;
    if (pair_ok) {
      if constexpr (sizeof(T) == 2) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(o0, o1);
      else *reinterpret_cast<float2*>(dst) = make_float2(Tr::to(o0), Tr::to(o1));
    } else {
      dst[0] = Tr::to(o0);
      if (f + 1 < F2) dst[1] = Tr::to(o1);
    }
  }
}

int launch_subsample_conv(int precision, const SubsampleArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.F % 2 == 0, "n_mels must be even");
  const int T_out = (a.T - 1) / 2 + 1;
  dim3 grid(cdiv(T_out, kSubTT), a.B);
  dim3 block((a.F / 2 + 1) / 2, kSubTT);
  EC_REQUIRE(block.x * block.y <= 512, "n_mels too large for the subsampling kernel");
  const size_t smem = sizeof(float) * ((a.F + 3) * (2 * kSubTT + 2) + a.C * 12);
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
