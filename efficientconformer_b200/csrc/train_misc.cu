// Small kernels the training-mode forward / backward needs next to the GEMMs (all bandwidth bound, fixed-order reductions):
//   swish_fwd / glu_fwd        the training forward keeps the PRE-activation tensors for the backward, so the activations run as
//                              their own element kernels instead of inside the GEMM epilogue (reference models/activations.py)
//   strided_rows               activation-type copy of every s-th frame (operand of conv_res, reference models/blocks.py:105-109)
//   col_stats / group merges   BatchNorm2d statistics of the Conv2d subsampling output: per-column (mean, M2), merged over the
//                              F/2 columns of a channel (feature index c*F/2 + f, reference models/modules.py:245-247)
//   subsample_conv_wgrad       weight / bias gradient of Conv2d(1 -> C, 3x3, stride 2) (reference models/modules.py:226)
#include "ec_common.cuh"
#include <algorithm>

namespace ec {

namespace {
constexpr int kCtas = 148;
__device__ __forceinline__ void cta_rows(size_t rows, size_t& r0, size_t& r1) {
  const size_t per = (rows + gridDim.x - 1) / gridDim.x;
  r0 = min(rows, per * blockIdx.x); r1 = min(rows, r0 + per);
}
}  // namespace

template <typename T>
__global__ void __launch_bounds__(256) swish_fwd_kernel(const T* __restrict__ z, size_t n, T* __restrict__ h) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = ActTraits<T>::from(z[i]);
    h[i] = ActTraits<T>::to(v / (1.f + __expf(-v)));
  }
}
template <typename T>
__global__ void __launch_bounds__(256) glu_fwd_kernel(const T* __restrict__ zg, size_t rows, int C, T* __restrict__ out) {
  const size_t n = rows * C, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const size_t r = i / C; const int c = static_cast<int>(i - r * C);
    const float a = ActTraits<T>::from(zg[r * 2 * C + c]), g = ActTraits<T>::from(zg[r * 2 * C + C + c]);
    out[i] = ActTraits<T>::to(a / (1.f + __expf(-g)));
  }
}
template <typename T>
__global__ void __launch_bounds__(256) strided_rows_kernel(const float* __restrict__ x, int B, int T_in, int T_out, int D, int stride,
                                                           T* __restrict__ out) {
  const size_t n = static_cast<size_t>(B) * T_out * D, st = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += st) {
    const int c = static_cast<int>(i % D);
    const size_t r = i / D;
    const int t = static_cast<int>(r % T_out), b = static_cast<int>(r / T_out);
    out[i] = ActTraits<T>::to(x[(static_cast<size_t>(b) * T_in + static_cast<size_t>(t) * stride) * D + c]);
  }
}
// scatter back: dx[b, t*s, :] += d[b, t, :]  (the gradient of the strided copy, added to the residual-stream gradient)
__global__ void __launch_bounds__(256) strided_rows_bwd_kernel(const float* __restrict__ d, int B, int T_in, int T_out, int D, int stride,
                                                               float* __restrict__ dx) {
  const size_t n = static_cast<size_t>(B) * T_out * D, st = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += st) {
    const int c = static_cast<int>(i % D);
    const size_t r = i / D;
    const int t = static_cast<int>(r % T_out), b = static_cast<int>(r / T_out);
    dx[(static_cast<size_t>(b) * T_in + static_cast<size_t>(t) * stride) * D + c] += d[i];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cast_scaled_kernel(const float* __restrict__ src, float scale, size_t n, T* __restrict__ dst) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = ActTraits<T>::to(scale * src[i]);
}
static int grid_for(size_t n);
int launch_cast_scaled(int precision, const float* src, float scale, size_t n, void* dst, cudaStream_t st) {
  EC_DISPATCH_PREC(precision, (cast_scaled_kernel<ActT><<<grid_for(n), 256, 0, st>>>(src, scale, n, reinterpret_cast<ActT*>(dst))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
static int grid_for(size_t n) { return static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16)); }

int launch_swish_fwd(int precision, const void* z, size_t n, void* h, cudaStream_t st) {
  EC_DISPATCH_PREC(precision, (swish_fwd_kernel<ActT><<<grid_for(n), 256, 0, st>>>(reinterpret_cast<const ActT*>(z), n, reinterpret_cast<ActT*>(h))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_glu_fwd(int precision, const void* zg, size_t rows, int C, void* out, cudaStream_t st) {
  EC_DISPATCH_PREC(precision, (glu_fwd_kernel<ActT><<<grid_for(rows * C), 256, 0, st>>>(reinterpret_cast<const ActT*>(zg), rows, C, reinterpret_cast<ActT*>(out))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_strided_rows(int precision, const float* x, int B, int T_in, int D, int stride, void* out, cudaStream_t st) {
  const int T_out = (T_in - 1) / stride + 1;
  const size_t n = static_cast<size_t>(B) * T_out * D;
  EC_DISPATCH_PREC(precision, (strided_rows_kernel<ActT><<<grid_for(n), 256, 0, st>>>(x, B, T_in, T_out, D, stride, reinterpret_cast<ActT*>(out))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_strided_rows_bwd(const float* d, int B, int T_in, int D, int stride, float* dx, cudaStream_t st) {
  const int T_out = (T_in - 1) / stride + 1;
  strided_rows_bwd_kernel<<<grid_for(static_cast<size_t>(B) * T_out * D), 256, 0, st>>>(d, B, T_in, T_out, D, stride, dx);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---- per-column statistics of an fp32 matrix: stats[0][c] = mean, stats[1][c] = centred sum of squares (two-pass per CTA + Chan) ----
__global__ void __launch_bounds__(128) col_stats_kernel(const float* __restrict__ y, size_t rows, int cols, float* __restrict__ partial) {
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= cols) return;
  size_t r0, r1; cta_rows(rows, r0, r1);
  float s1 = 0.f, s2 = 0.f;
  for (size_t r = r0; r < r1; ++r) s1 += y[r * cols + c];
  const float lm = r1 > r0 ? s1 / static_cast<float>(r1 - r0) : 0.f;
  for (size_t r = r0; r < r1; ++r) { const float d = y[r * cols + c] - lm; s2 = fmaf(d, d, s2); }
  partial[(static_cast<size_t>(blockIdx.x) * 2) * cols + c] = lm;
  partial[(static_cast<size_t>(blockIdx.x) * 2 + 1) * cols + c] = s2;
}
__global__ void col_stats_merge_kernel(const float* __restrict__ partial, int n_partial, size_t rows, int cols, float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const size_t per = (rows + n_partial - 1) / n_partial;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int p = 0; p < n_partial; ++p) {
    const size_t r0 = min(rows, per * p), r1 = min(rows, r0 + per);
    const double nb = static_cast<double>(r1 - r0);
    if (nb == 0.0) continue;
    const double mb = partial[(static_cast<size_t>(p) * 2) * cols + c], qb = partial[(static_cast<size_t>(p) * 2 + 1) * cols + c];
    const double tot = n + nb, dl = mb - mean;
    mean += dl * nb / tot; m2 += qb + dl * dl * n * nb / tot; n = tot;
  }
  stats[c] = static_cast<float>(mean); stats[cols + c] = static_cast<float>(m2);
}
// channel c owns the `group` consecutive columns c*group .. : merge their (mean, M2) (each over `rows` samples) -> [2][C]
__global__ void group_stats_merge_kernel(const float* __restrict__ col_stats, int C, int group, double rows, float* __restrict__ ch_stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int cols = C * group;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int f = 0; f < group; ++f) {
    const double mb = col_stats[c * group + f], qb = col_stats[cols + c * group + f];
    const double tot = n + rows, dl = mb - mean;
    mean += dl * rows / tot; m2 += qb + dl * dl * n * rows / tot; n = tot;
  }
  ch_stats[c] = static_cast<float>(mean); ch_stats[C + c] = static_cast<float>(m2);
}
// out[j][c*group + f] = in[j][c]  (per-channel vectors expanded to per-column vectors);  n_vec stacked vectors
__global__ void group_expand_kernel(const float* __restrict__ in, int n_vec, int C, int group, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec * C * group) return;
  const int j = i / (C * group), col = i - j * C * group;
  out[i] = in[j * C + col / group];
}
// out[j][c] = sum_f in[j][c*group + f]  (column sums folded into channel sums, fixed order)
__global__ void group_sum_kernel(const float* __restrict__ in, int n_vec, int C, int group, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec * C) return;
  const int j = i / C, c = i - j * C;
  float s = 0.f;
  for (int f = 0; f < group; ++f) s += in[static_cast<size_t>(j) * C * group + c * group + f];
  out[i] = s;
}

size_t col_stats_work_bytes(int cols) { return align_up(static_cast<size_t>(kCtas) * 2 * cols * sizeof(float), 256); }
int launch_col_stats(const float* y, size_t rows, int cols, float* stats, float* work, cudaStream_t st) {
  const int ctas = static_cast<int>(std::min<size_t>(kCtas, std::max<size_t>(rows, 1)));
  col_stats_kernel<<<dim3(ctas, cdiv(cols, 128)), 128, 0, st>>>(y, rows, cols, work);
  EC_CUDA(cudaGetLastError());
  col_stats_merge_kernel<<<cdiv(cols, 128), 128, 0, st>>>(work, ctas, rows, cols, stats);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_group_stats_merge(const float* col_stats, int C, int group, size_t rows, float* ch_stats, cudaStream_t st) {
  group_stats_merge_kernel<<<cdiv(C, 128), 128, 0, st>>>(col_stats, C, group, static_cast<double>(rows), ch_stats);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_group_expand(const float* in, int n_vec, int C, int group, float* out, cudaStream_t st) {
  group_expand_kernel<<<cdiv(n_vec * C * group, 256), 256, 0, st>>>(in, n_vec, C, group, out);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_group_sum(const float* in, int n_vec, int C, int group, float* out, cudaStream_t st) {
  group_sum_kernel<<<cdiv(n_vec * C, 256), 256, 0, st>>>(in, n_vec, C, group, out);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---- Conv2d(1 -> C, 3x3, stride 2, pad 1) weight / bias gradient ------------------------------------------------------------------
// dy [B, T_out, C*F2] (feature c*F2 + f), mel [B, F, T]:  dw[c][kh*3+kw] = sum_{b,t,f} dy * mel[b, 2f-1+kh, 2t-1+kw],  db[c] = sum dy
// Weight / bias gradient of Conv2d(1 -> C, 3x3, stride 2, pad 1):  dw[c, kh, kw] = sum_{b,t,f} dy[(b,t), c*F2+f] * mel[b, 2f-1+kh, 2t-1+kw].
// Pre-pass: zero-bordered transposed copy melT[b, ti+1, fi+1] (coalesced along frequency, no bounds checks in the main loop).
// Main kernel: thread = output column (c, f), CTA = row range; 9 + 1 accumulators; partial[cta][col][10]; then one block per channel
// adds the partials of its F2 columns in a fixed order.
__global__ void __launch_bounds__(256) subsample_melT_kernel(const float* __restrict__ mel, int B, int F, int T_in, float* __restrict__ melT) {
  const int Fp = F + 2, Tp = T_in + 2;
  const size_t n = static_cast<size_t>(B) * Tp * Fp, st = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += st) {
    const int fp = static_cast<int>(i % Fp);
    const size_t r = i / Fp;
    const int tp = static_cast<int>(r % Tp), b = static_cast<int>(r / Tp);
    const int fi = fp - 1, ti = tp - 1;
    melT[i] = (fi >= 0 && fi < F && ti >= 0 && ti < T_in) ? mel[(static_cast<size_t>(b) * F + fi) * T_in + ti] : 0.f;
  }
}
__global__ void __launch_bounds__(128) subsample_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ melT, int B, int F, int T_in,
                                                              int T_out, int C, float* __restrict__ partial) {
  const int F2 = F / 2, cols = C * F2, Fp = F + 2, Tp = T_in + 2;
  const int col = blockIdx.y * 128 + threadIdx.x;
  if (col >= cols) return;
  const int f = col % F2;
  float acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k] = 0.f;
  size_t r0, r1; cta_rows(static_cast<size_t>(B) * T_out, r0, r1);
  for (size_t r = r0; r < r1; ++r) {
    const int b = static_cast<int>(r / T_out), t = static_cast<int>(r - static_cast<size_t>(b) * T_out);
    const float d = dy[r * cols + col];
    // rows 2t-1 .. 2t+1 of the input are rows 2t .. 2t+2 of the bordered copy (2t+2 <= T_in+1 always); same for the frequencies
    const float* base = melT + (static_cast<size_t>(b) * Tp + 2 * t) * Fp + 2 * f;
    acc[9] += d;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) acc[kh * 3 + kw] = fmaf(d, base[kw * Fp + kh], acc[kh * 3 + kw]);
  }
  float* out = partial + (static_cast<size_t>(blockIdx.x) * cols + col) * 10;
#pragma unroll
  for (int k = 0; k < 10; ++k) out[k] = acc[k];
}
// block = channel c: 256 threads stride over the (partial, f) pairs, then a fixed-order shared-memory tree per tap
__global__ void __launch_bounds__(256) subsample_wgrad_reduce_kernel(const float* __restrict__ partial, int n_partial, int C, int F2,
                                                                     float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sm[256][11];
  const int c = blockIdx.x, cols = C * F2, tid = threadIdx.x;
  float acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k] = 0.f;
  for (int idx = tid; idx < n_partial * F2; idx += 256) {
    const int p = idx / F2, f = idx - p * F2;
    const float* src = partial + (static_cast<size_t>(p) * cols + c * F2 + f) * 10;
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] += src[k];
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) sm[tid][k] = acc[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s)
#pragma unroll
      for (int k = 0; k < 10; ++k) sm[tid][k] += sm[tid + s][k];
    __syncthreads();
  }
  if (tid < 9) dw[c * 9 + tid] = sm[0][tid];
  if (tid == 9) db[c] = sm[0][9];
}
static constexpr int kWgCtas = 592;
static size_t melT_bytes(int B, int F, int T) { return align_up(static_cast<size_t>(B) * (T + 2) * (F + 2) * sizeof(float), 256); }
static size_t wg_partial_bytes(int C, int F) { return align_up(static_cast<size_t>(kWgCtas) * C * (F / 2) * 10 * sizeof(float), 256); }
size_t subsample_wgrad_work_bytes(int C, int F, int B, int T) { return wg_partial_bytes(C, F) + melT_bytes(B, F, T); }
int launch_subsample_wgrad(const float* dy, const float* mel, int B, int F, int T, int C, float* dw, float* db, float* work, cudaStream_t st) {
  EC_REQUIRE(F % 2 == 0, "Conv2d subsampling weight gradient: even number of mel bins");
  const int T_out = (T - 1) / 2 + 1, cols = C * (F / 2);
  const int ctas = static_cast<int>(std::min<size_t>(kWgCtas, static_cast<size_t>(B) * T_out));
  float* melT = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(work) + wg_partial_bytes(C, F));
  subsample_melT_kernel<<<grid_for(static_cast<size_t>(B) * (T + 2) * (F + 2)), 256, 0, st>>>(mel, B, F, T, melT);
  EC_CUDA(cudaGetLastError());
  subsample_wgrad_kernel<<<dim3(ctas, cdiv(cols, 128)), 128, 0, st>>>(dy, melT, B, F, T, T_out, C, work);
  EC_CUDA(cudaGetLastError());
  subsample_wgrad_reduce_kernel<<<C, 256, 0, st>>>(work, ctas, C, F / 2, dw, db);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec
