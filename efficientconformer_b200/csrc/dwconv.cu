// Fused depthwise Conv1d (k taps, stride s, 'same' zero halo) + eval BatchNorm1d (folded into taps/bias) + Swish.
//
// Restates reference models/modules.py:515-517 with Conv1d "same" pre-padding (models/layers.py:99-100,131-136):
//   out[b, t, c] = swish( b'_c + sum_k w'[c,k] * in[b, t*s + k - (K-1)/2, c] ),   T_out = (T-1)//s + 1,
// on channels-last activations (the GLU output of the pw1 GEMM epilogue), with NO length masking (padded frames
// are computed densely like the reference).  Bandwidth-bound: each input element is read from global once per CTA
// (16-byte vector loads into shared memory, halo included), each thread then slides a register window over time for
// one channel, so shared memory is read ~(R*s+K-1)/R times per output; the output is written once.  Models wider than
// 256 channels (Medium / Large: up to 720) are tiled along the channel dim (grid.z), 256 channels per CTA.
#include "ec_common.cuh"

namespace ec {

constexpr int kDwR = 16;        // outputs per thread (register run along time)
constexpr int kDwRuns = 4;      // runs per CTA  -> 64 output frames per CTA
constexpr int kDwMaxK = 31;
constexpr int kDwMaxTile = 256;  // channels per CTA

template <typename T, int STRIDE, int K>
__global__ void __launch_bounds__(1024) dwconv_bn_swish_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                               int T_in, int T_out, int C, int CT, T* __restrict__ y) {
  using Tr = ActTraits<T>;
  extern __shared__ __align__(16) uint8_t dw_smem[];
  T* tile = reinterpret_cast<T*>(dw_smem);
  grid_dependency_wait();
  grid_launch_dependents();
  const int b = blockIdx.y;
  const int to0 = blockIdx.x * (kDwR * kDwRuns);
  const int pad = (K - 1) / 2;
  const int rows = (kDwR * kDwRuns - 1) * STRIDE + K;    // input frames needed by this CTA
  const int ti0 = to0 * STRIDE - pad;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;

  // ---- global -> shared, 16-byte vectors; a CTA stages the (rows x cw) slab of its channel tile [c0, c0 + cw) ----
  constexpr int VEC = 16 / sizeof(T);
  const int c0 = blockIdx.z * CT, cw = min(CT, C - c0);
  const T* xb = x + static_cast<size_t>(b) * T_in * C + c0;
  if (C % VEC == 0 && CT % VEC == 0) {
    const int nv = cw / VEC, total = rows * nv;          // cw is a multiple of VEC because C and CT are
    for (int i = tid; i < total; i += nthr) {
      const int r = i / nv, v = i - r * nv;
      const int tt = ti0 + r;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (tt >= 0 && tt < T_in) val = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(tt) * C + v * VEC);
      *reinterpret_cast<uint4*>(tile + r * cw + v * VEC) = val;
    }
  } else {
    const int total = rows * cw;
    for (int i = tid; i < total; i += nthr) {
      const int r = i / cw;
      const int tt = ti0 + r;
      tile[i] = (tt >= 0 && tt < T_in) ? xb[static_cast<size_t>(tt) * C + (i - r * cw)] : Tr::to(0.f);
    }
  }
  __syncthreads();

  const int c = threadIdx.x;
  if (c >= cw) return;
  float wk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wk[k] = __ldg(w + (c0 + c) * K + k);
  const float bc = __ldg(bias + c0 + c);
  const int run0 = threadIdx.y * kDwR;                  // first local output of this thread
  float acc[kDwR];
#pragma unroll
  for (int r = 0; r < kDwR; ++r) acc[r] = bc;
  // input-stationary sweep: every staged input frame is read once and scattered into the outputs it feeds
  constexpr int in_rows = (kDwR - 1) * STRIDE + K;
  const T* col = tile + static_cast<size_t>(run0 * STRIDE) * cw + c;
#pragma unroll
  for (int i = 0; i < in_rows; ++i) {
    const float v = Tr::from(col[static_cast<size_t>(i) * cw]);
#pragma unroll
    for (int r = 0; r < kDwR; ++r) {
      const int k = i - r * STRIDE;            // compile-time after full unrolling
      if (k >= 0 && k < K) acc[r] = fmaf(v, wk[k], acc[r]);
    }
  }
  T* yb = y + static_cast<size_t>(b) * T_out * C + c0;
#pragma unroll
  for (int r = 0; r < kDwR; ++r) {
    const int to = to0 + run0 + r;
    if (to < T_out) yb[static_cast<size_t>(to) * C + c] = Tr::to(swish_fn<T>(acc[r]));
  }
}

template <typename T>
static int launch_dw_t(const DwConvArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.k % 2 == 1 && a.k <= kDwMaxK, "depthwise kernel size must be odd and <= 31");
  EC_REQUIRE(a.stride == 1 || a.stride == 2, "depthwise stride must be 1 or 2");
  const int T_out = (a.T - 1) / a.stride + 1;
  const int CT = a.C <= kDwMaxTile ? a.C : kDwMaxTile;   // channel tile of a CTA (wide models: several tiles along grid.z)
  const int cx = round_up(CT, 32);
  dim3 block(cx, kDwRuns);
  dim3 grid(cdiv(T_out, kDwR * kDwRuns), a.B, cdiv(a.C, CT));
  const int rows = (kDwR * kDwRuns - 1) * a.stride + a.k;
  const size_t smem = align_up(static_cast<size_t>(rows) * CT * sizeof(T), 16);
  EC_REQUIRE(smem <= 200 * 1024, "depthwise slab does not fit in shared memory");
  const T* x = reinterpret_cast<const T*>(a.x);
  T* y = reinterpret_cast<T*>(a.y);
#define EC_DW_CASE(S, KK)                                                                                              \
  if (a.stride == S && a.k == KK) {                                                                                    \
    static cudaError_t e = cudaFuncSetAttribute(dwconv_bn_swish_kernel<T, S, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    EC_CUDA(e);                                                                                                        \
    return launch_pdl(dwconv_bn_swish_kernel<T, S, KK>, grid, block, smem, stream, x, a.w, a.b, a.T, T_out, a.C, CT, y); \
  }
  EC_DW_CASE(1, 15) EC_DW_CASE(2, 15) EC_DW_CASE(1, 31) EC_DW_CASE(2, 31)
#undef EC_DW_CASE
  EC_FAIL("depthwise kernel size " + std::to_string(a.k) + " is not instantiated (15 and 31 are)");
}

int launch_dwconv_bn_swish(int precision, const DwConvArgs& a, cudaStream_t stream) {
  EC_DISPATCH_PREC(precision, return launch_dw_t<ActT>(a, stream));
}

}  // namespace ec
