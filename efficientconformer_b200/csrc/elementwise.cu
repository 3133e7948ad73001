// Bandwidth-bound row kernels: LayerNorm (+ optional strided compaction copy), operand casts, weight preparation.
#include "ec_common.cuh"

namespace ec {

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (reference nn.LayerNorm(dim, eps=1e-6): models/modules.py:386,433,511; blocks.py:96).
// One warp per row, the row lives in registers (dim <= 32*kMaxPerLane), two-pass mean / centred variance in fp32,
// warp-shuffle reductions, coalesced loads/stores.  Output: activation type (GEMM operand) or fp32 (block output).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kLnMaxPerLane = 32;   // dim <= 1024

template <typename T, int NPL>      // NPL = elements per lane kept in registers (dim <= 32 * NPL)
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int rows, int dim,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                        T* __restrict__ y_act, float* __restrict__ y_f32, T* __restrict__ copy_out,
                                                        int copy_stride, int frames_per_seq, int frames_out_per_seq) {
  using Tr = ActTraits<T>;
  grid_dependency_wait();
  grid_launch_dependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + static_cast<size_t>(warp) * dim;
  float v[NPL];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < dim) ? xr[c] : 0.f;
    sum += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mu = sum / dim;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    const float d = (c < dim) ? v[i] - mu : 0.f;
    sq += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / dim + eps);
  T* copy_row = nullptr;
  if (copy_out != nullptr) {
    const int seq = warp / frames_per_seq, t = warp % frames_per_seq;
    if (t % copy_stride == 0) copy_row = copy_out + (static_cast<size_t>(seq) * frames_out_per_seq + t / copy_stride) * dim;
  }
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    if (c < dim) {
      const float o = (v[i] - mu) * rstd * gamma[c] + beta[c];
      if (y_f32 != nullptr) y_f32[static_cast<size_t>(warp) * dim + c] = o;
      if (y_act != nullptr) y_act[static_cast<size_t>(warp) * dim + c] = Tr::to(o);
      if (copy_row != nullptr) copy_row[c] = Tr::to(v[i]);
    }
  }
}

template <typename T>
static int launch_layernorm_t(const LayerNormArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.dim <= 32 * kLnMaxPerLane, "LayerNorm dim too large");
  EC_REQUIRE(a.rows > 0, "empty LayerNorm");
  const int threads = 256, rows_per_block = threads / 32;
  dim3 grid(cdiv(a.rows, rows_per_block));
  T* copy = reinterpret_cast<T*>(a.copy_out);
  const int cs = a.copy_stride > 0 ? a.copy_stride : 1;
  const int fps = a.frames_per_seq > 0 ? a.frames_per_seq : a.rows;
  T* ya = reinterpret_cast<T*>(a.y_act);
  if (a.dim <= 128) return launch_pdl(layernorm_kernel<T, 4>, grid, dim3(threads), 0, stream, a.x, a.rows, a.dim, a.gamma, a.beta, a.eps, ya, a.y_f32, copy, cs, fps, a.frames_out_per_seq);
  if (a.dim <= 256) return launch_pdl(layernorm_kernel<T, 8>, grid, dim3(threads), 0, stream, a.x, a.rows, a.dim, a.gamma, a.beta, a.eps, ya, a.y_f32, copy, cs, fps, a.frames_out_per_seq);
  return launch_pdl(layernorm_kernel<T, kLnMaxPerLane>, grid, dim3(threads), 0, stream, a.x, a.rows, a.dim, a.gamma, a.beta, a.eps, ya, a.y_f32, copy, cs, fps, a.frames_out_per_seq);
}

int launch_layernorm(int precision, const LayerNormArgs& a, cudaStream_t stream) {
  EC_DISPATCH_PREC(precision, return launch_layernorm_t<ActT>(a, stream));
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 -> activation type (TF32 rounding or bf16)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void cast_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
  grid_dependency_wait();
  grid_launch_dependents();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = ActTraits<T>::to(src[i]);
}

int launch_cast_rows(int precision, const float* src, void* dst, size_t n, cudaStream_t stream) {
  if (n == 0) return EC_OK;
  const int threads = 256;
  const int blocks = static_cast<int>(std::min<size_t>((n + threads - 1) / threads, 148 * 8));
  EC_DISPATCH_PREC(precision, ((void)launch_dep(cast_kernel<ActT>, dim3(blocks), dim3(threads), 0, stream, src, reinterpret_cast<ActT*>(dst), n)));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// Weight operand of a GEMM: the activation-type cast and, in split mode, the second plane with the halves swapped at
// dst + twin_elems (see effconf_b200.h, EC_PREC_BF16X2).
__global__ void cast_weight_split_kernel(const float* __restrict__ src, uint32_t* __restrict__ dst, size_t n, size_t twin_elems) {
  grid_dependency_wait();
  grid_launch_dependents();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t v = split_pack(src[i]);
    dst[i] = v;
    dst[twin_elems + i] = split_swap(v);
  }
}
int launch_cast_weight(int precision, const float* src, void* dst, size_t n, size_t twin_elems, cudaStream_t stream) {
  if (precision != EC_PREC_BF16X2) return launch_cast_rows(precision, src, dst, n, stream);
  if (n == 0) return EC_OK;
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 8));
  (void)launch_dep(cast_weight_split_kernel, dim3(blocks), dim3(256), 0, stream, src, reinterpret_cast<uint32_t*>(dst), n, twin_elems);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Eval-mode BatchNorm folded into the preceding convolution (reference nn.BatchNorm1d/2d, eps 1e-5:
// models/modules.py:228,517):  y = (conv(x) - rm) / sqrt(rv + eps) * g + beta  ==  conv'(x) with
// w' = w * s, b' = (b - rm) * s + beta, s = g / sqrt(rv + eps).
// ---------------------------------------------------------------------------------------------------------------
__global__ void fold_bn_kernel(const float* w, const float* b, const float* g, const float* beta, const float* rm, const float* rv,
                               float eps, int C, int taps, float* w_out, float* b_out) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * taps) return;
  const int c = i / taps;
  const float s = g[c] / sqrtf(rv[c] + eps);
  w_out[i] = w[i] * s;
  if (i % taps == 0) b_out[c] = (b[c] - rm[c]) * s + beta[c];
}

int launch_fold_bn(const float* w, const float* b, const float* g, const float* beta, const float* rm, const float* rv,
                   float eps, int C, int taps, float* w_out, float* b_out, cudaStream_t stream) {
  (void)launch_dep(fold_bn_kernel, dim3(cdiv(C * taps, 256)), dim3(256), 0, stream, w, b, g, beta, rm, rv, eps, C, taps, w_out, b_out);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// GLU weight interleave: raw pointwise weight [2C, K] (value rows 0..C-1, gate rows C..2C-1; reference
// models/activations.py:37-39 chunks dim 1 in two) -> tiles of [nb value rows | nb gate rows], zero rows beyond C.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void glu_interleave_kernel(const float* __restrict__ w, const float* __restrict__ b, int C, int K, int nb, int tiles,
                                      T* __restrict__ w_out, float* __restrict__ b_out) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int row = blockIdx.x;               // output row in [0, tiles*2*nb)
  const int tile = row / (2 * nb), r = row % (2 * nb);
  const bool gate = r >= nb;
  const int ch = tile * nb + (gate ? r - nb : r);
  const bool valid = ch < C;
  const int src_row = gate ? C + ch : ch;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const T v = ActTraits<T>::to(valid ? w[static_cast<size_t>(src_row) * K + k] : 0.f);
    w_out[static_cast<size_t>(row) * K + k] = v;
    if constexpr (IsSplit<T>::value)     // swapped plane of the [2, rows, K] weight operand
      w_out[(static_cast<size_t>(gridDim.x) + row) * K + k] = SplitBf16{split_swap(v.bits)};
  }
  if (threadIdx.x == 0) b_out[row] = valid ? b[src_row] : 0.f;
}

int launch_glu_interleave(int precision, const float* w, const float* b, int channels, int K, int nb, int tiles,
                          void* w_out, float* b_out, cudaStream_t stream) {
  const int rows = tiles * 2 * nb;
  EC_DISPATCH_PREC(precision, ((void)launch_dep(glu_interleave_kernel<ActT>, dim3(rows), dim3(128), 0, stream, w, b, channels, K, nb, tiles, reinterpret_cast<ActT*>(w_out), b_out)));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Valid lengths seen by every block: out[0] = (x_len-1)//2+1 after the Conv2d subsampling (reference
// models/modules.py:243), out[i+1] = (out[i]-1)//stride_i+1 after block i (models/encoders.py:140; stride 1 keeps it).
// x_len == nullptr -> full length.  out is [(n_blocks+1), B].
// ---------------------------------------------------------------------------------------------------------------
__global__ void stage_lengths_kernel(const long long* x_len, int B, int t_mel, BlockStrides st, int* out) {
  grid_dependency_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long l = x_len != nullptr ? x_len[b] : t_mel;
  if (l > t_mel) l = t_mel;
  for (int i = 0; i < st.sub_layers; ++i) l = l > 0 ? (l - 1) / 2 + 1 : 0;   // one halving per Conv2d subsampling layer
  out[b] = static_cast<int>(l);
  for (int i = 0; i < st.n; ++i) {
    if (l > 0) l = (l - 1) / st.s[i] + 1;
    out[(i + 1) * B + b] = static_cast<int>(l);
  }
}

int launch_stage_lengths(const long long* x_len, int B, int t_mel, const BlockStrides& st, int* out, cudaStream_t stream) {
  (void)launch_dep(stage_lengths_kernel, dim3(cdiv(B, 128)), dim3(128), 0, stream, x_len, B, t_mel, st, out);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

__global__ void i64_to_i32_kernel(const long long* src, int n, int* dst, int clamp_max) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { long long v = src[i]; if (v > clamp_max) v = clamp_max; if (v < 0) v = 0; dst[i] = static_cast<int>(v); }
}
__global__ void i32_to_i64_kernel(const int* src, int n, long long* dst) {
  grid_dependency_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
int launch_i64_to_i32(const long long* src, int n, int* dst, int clamp_max, cudaStream_t stream) {
  (void)launch_dep(i64_to_i32_kernel, dim3(cdiv(n, 128)), dim3(128), 0, stream, src, n, dst, clamp_max);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_i32_to_i64(const int* src, int n, long long* dst, cudaStream_t stream) {
  (void)launch_dep(i32_to_i64_kernel, dim3(cdiv(n, 128)), dim3(128), 0, stream, src, n, dst);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec
