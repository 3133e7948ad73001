// Training-mode kernels of the convolution module's depthwise stage:
//   depthwise Conv1d (k taps, stride s, 'same' zero halo) -> BatchNorm1d with BATCH statistics over all B*T_out frames, padded
//   frames included (the reference never masks) -> Swish             (reference models/modules.py:515-517, models/layers.py:96-136)
// and their backward.  Channels-last activations [B, T, C]; all kernels are bandwidth bound; every reduction over frames uses
// per-CTA partials added in a fixed order (bit-reproducible).  Between the statistics and the normalisation the host may
// merge the per-rank (count, mean, M2) statistics across ranks (SyncBatchNorm of the reference's distribute_strategy, models/model_ctc.py:73).
//
//   forward :  dwconv_raw        y = b + sum_k w[c,k] x[t*s + k - pad]            (fp32) + per-channel mean / centred sum of squares
//              bn_swish_fwd      h = swish((y - mean) * rstd * gamma + beta)      (activation type)
//              bn_running_update running statistics, momentum 0.1, unbiased variance
//   backward:  bn_swish_bwd_stats  dz = dh * swish'(z);  sum dz, sum dz * xhat    (= dbeta, dgamma)
//              bn_swish_bwd_apply  dy = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat))
//              dwconv_bwd_data     dx[ti] = sum_k w[c,k] dy[(ti + pad - k) / s]   (terms with a whole, in-range quotient)
//              dwconv_bwd_weight   dw[c,k] = sum_{b,t} dy[t] x[t*s + k - pad],  db[c] = sum dy
#include "ec_common.cuh"
#include <algorithm>

namespace ec {

namespace {
constexpr int kCtas = 888;      // 8 per SM: these kernels are latency bound on one dependent load chain per thread
constexpr int kMaxTaps = 31;

// rows [r0, r1) of this CTA for a row-strided split of `rows` over gridDim.x CTAs (contiguous ranges: coalesced, deterministic)
__device__ __forceinline__ void cta_rows(size_t rows, size_t& r0, size_t& r1) {
  const size_t per = (rows + gridDim.x - 1) / gridDim.x;
  r0 = min(rows, per * blockIdx.x); r1 = min(rows, r0 + per);
}
}  // namespace

// ---- forward: raw depthwise conv + statistics ------------------------------------------------------------------------------
// thread = channel (blockIdx.y tiles channels by 128), CTA = a contiguous range of output frames (flattened b*T_out + t)
template <typename T, int KT>
__global__ void __launch_bounds__(128) dwconv_raw_kernel_(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         int B, int T_in, int T_out, int C, int K, int stride, float* __restrict__ y,
                                                         float* __restrict__ partial /* [gridDim.x][2][C] */) {
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  float wk[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) wk[k] = k < K ? w[c * K + k] : 0.f;
  const float bc = bias[c];
  const int pad = (K - 1) / 2;
  size_t r0, r1; cta_rows(static_cast<size_t>(B) * T_out, r0, r1);
  float s1 = 0.f, s2 = 0.f;
  for (size_t r = r0; r < r1; ++r) {
    const int b = static_cast<int>(r / T_out), t = static_cast<int>(r - static_cast<size_t>(b) * T_out);
    const T* xb = x + static_cast<size_t>(b) * T_in * C + c;
    float acc = bc;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const int ti = t * stride + k - pad;
      if (k < K && ti >= 0 && ti < T_in) acc = fmaf(wk[k], ActTraits<T>::from(xb[static_cast<size_t>(ti) * C]), acc);
    }
    y[r * C + c] = acc;
    s1 += acc;
  }
  // second pass over this thread's own outputs: centred sum of squares about the local mean (sum of squares minus mean^2 loses
  // every digit when |mean| >> std, e.g. a tiny tap vector next to a large bias)
  const float lm = r1 > r0 ? s1 / static_cast<float>(r1 - r0) : 0.f;
  for (size_t r = r0; r < r1; ++r) { const float d = y[r * C + c] - lm; s2 = fmaf(d, d, s2); }
  partial[(static_cast<size_t>(blockIdx.x) * 2) * C + c] = lm;
  partial[(static_cast<size_t>(blockIdx.x) * 2 + 1) * C + c] = s2;
}

// (mean, M2) of the CTA row ranges merged with Chan's update: stats[0][c] = mean, stats[1][c] = M2.
// Block = 32 channels x 32 lanes: lane ty merges the contiguous chunk ty of the partials in CTA order, then the 32 chunk results
// are merged in lane order in double (fixed order: bit-reproducible; the dependent chain is n_partial / 32 + 32 steps).
__global__ void __launch_bounds__(1024) bn_stats_merge_kernel(const float* __restrict__ partial, int n_partial, size_t rows, int C,
                                                              float* __restrict__ stats) {
  __shared__ float sn[32][33], smean[32][33], sm2[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const size_t per = (rows + n_partial - 1) / n_partial;
  const int chunk = (n_partial + 31) / 32, p0 = ty * chunk, p1 = min(n_partial, p0 + chunk);
  float n = 0.f, mean = 0.f, m2 = 0.f;
  if (c < C) {
    for (int p = p0; p < p1; ++p) {
      const size_t r0 = min(rows, per * p), r1 = min(rows, r0 + per);
      const float nb = static_cast<float>(r1 - r0);
      if (nb == 0.f) continue;
      const float mb = partial[(static_cast<size_t>(p) * 2) * C + c], qb = partial[(static_cast<size_t>(p) * 2 + 1) * C + c];
      const float tot = n + nb, dl = mb - mean, f = nb / tot;
      mean = fmaf(dl, f, mean);
      m2 += qb + dl * dl * n * f;
      n = tot;
    }
  }
  sn[ty][tx] = n; smean[ty][tx] = mean; sm2[ty][tx] = m2;
  __syncthreads();
  if (ty == 0 && c < C) {
    double dn = 0.0, dmean = 0.0, dm2 = 0.0;
    for (int q = 0; q < 32; ++q) {
      const double nb = sn[q][tx];
      if (nb == 0.0) continue;
      const double tot = dn + nb, dl = static_cast<double>(smean[q][tx]) - dmean;
      dmean += dl * nb / tot;
      dm2 += static_cast<double>(sm2[q][tx]) + dl * dl * dn * nb / tot;
      dn = tot;
    }
    stats[c] = static_cast<float>(dmean);
    stats[C + c] = static_cast<float>(dm2);
  }
}

// out[i] = sum_p partial[p][i], i < n: block = 32 outputs x 32 lanes, lane ty adds chunk ty in CTA order, chunk sums added in lane order
__device__ __forceinline__ float chunked_sum_32x32(const float* __restrict__ partial, int n_partial, size_t stride, int i, bool ok,
                                                   float (*sm)[33]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int per = (n_partial + 31) / 32, p0 = ty * per, p1 = min(n_partial, p0 + per);
  float s = 0.f;
  if (ok)
    for (int p = p0; p < p1; ++p) s += partial[static_cast<size_t>(p) * stride + i];
  sm[ty][tx] = s;
  __syncthreads();
  float t = 0.f;
  if (ty == 0) {
#pragma unroll
    for (int q = 0; q < 32; ++q) t += sm[q][tx];
  }
  return t;
}
__global__ void __launch_bounds__(1024) conv_partial_reduce_kernel(const float* __restrict__ partial, int n_partial, int n_out, int dim,
                                                                   float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int i = blockIdx.x * 32 + (threadIdx.x & 31), n = n_out * dim;
  const float t = chunked_sum_32x32(partial, n_partial, static_cast<size_t>(n), i, i < n, sm);
  if ((threadIdx.x >> 5) == 0 && i < n) out[i] = t;
}

// stats [2][C] (mean, centred sum of squares M2) over `count` frames -> mean / rstd (biased variance) and the running-statistics update
__global__ void bn_finalize_kernel(const float* __restrict__ sums, int C, float count, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = sums[c];
  const float var = fmaxf(sums[C + c] / count, 0.f);
  mean[c] = m;
  rstd[c] = rsqrtf(var + eps);
  if (running_mean != nullptr) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (count > 1.f ? count / (count - 1.f) : 1.f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_swish_fwd_kernel(const float* __restrict__ y, size_t rows, int C, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, T* __restrict__ h) {
  const size_t n = rows * C, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = static_cast<int>(i % C);
    const float z = (y[i] - mean[c]) * rstd[c] * gamma[c] + beta[c];
    h[i] = ActTraits<T>::to(z / (1.f + __expf(-z)));
  }
}

// ---- backward ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float swish_grad(float z) { const float s = 1.f / (1.f + __expf(-z)); return s + z * s * (1.f - s); }

// thread = channel; partial[cta][0][c] = sum dz, partial[cta][1][c] = sum dz * xhat
__global__ void __launch_bounds__(128) bn_swish_bwd_stats_kernel(const float* __restrict__ y, const float* __restrict__ dh, size_t rows, int C,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 float* __restrict__ partial) {
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  const float mu = mean[c], rs = rstd[c], g = gamma[c], be = beta[c];
  size_t r0, r1; cta_rows(rows, r0, r1);
  float s1 = 0.f, s2 = 0.f;
  for (size_t r = r0; r < r1; ++r) {
    const float xh = (y[r * C + c] - mu) * rs;
    const float dz = dh[r * C + c] * swish_grad(xh * g + be);
    s1 += dz; s2 = fmaf(dz, xh, s2);
  }
  partial[(static_cast<size_t>(blockIdx.x) * 2) * C + c] = s1;
  partial[(static_cast<size_t>(blockIdx.x) * 2 + 1) * C + c] = s2;
}

// dy = gamma * rstd * (dz - sum_dz / R - xhat * sum_dzx / R)   (R = frames the statistics were taken over, all ranks)
__global__ void __launch_bounds__(256) bn_swish_bwd_apply_kernel(const float* __restrict__ y, const float* __restrict__ dh, size_t rows, int C,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ sums /* [2][C] */, float count,
                                                                 float* __restrict__ dy) {
  const size_t n = rows * C, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const float inv = 1.f / count;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = static_cast<int>(i % C);
    const float xh = (y[i] - mean[c]) * rstd[c];
    const float dz = dh[i] * swish_grad(xh * gamma[c] + beta[c]);
    dy[i] = gamma[c] * rstd[c] * (dz - sums[c] * inv - xh * sums[C + c] * inv);
  }
}

// dx[b, ti, c] = sum_k w[c,k] * dy[b, to, c],  to = (ti + pad - k) / s when divisible and 0 <= to < T_out
template <int KT>
__global__ void __launch_bounds__(128) dwconv_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, int B, int T_in,
                                                              int T_out, int C, int K, int stride, float* __restrict__ dx) {
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  float wk[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) wk[k] = k < K ? w[c * K + k] : 0.f;
  const int pad = (K - 1) / 2;
  size_t r0, r1; cta_rows(static_cast<size_t>(B) * T_in, r0, r1);
  for (size_t r = r0; r < r1; ++r) {
    const int b = static_cast<int>(r / T_in), ti = static_cast<int>(r - static_cast<size_t>(b) * T_in);
    const float* db = dy + static_cast<size_t>(b) * T_out * C + c;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const int num = ti + pad - k;
      // stride is 1 or 2 (checked by the launcher): no integer division in the tap loop
      const bool hit = k < K && num >= 0 && (stride == 1 || (num & 1) == 0);
      const int to = stride == 1 ? num : (num >> 1);
      if (hit && to < T_out) acc = fmaf(wk[k], db[static_cast<size_t>(to) * C], acc);
    }
    dx[r * C + c] = acc;
  }
}

// partial[cta][c][k] = sum over this CTA's output frames of dy * x[t*s + k - pad];  partial[cta][c][K] = sum dy
template <typename T, int KT>
__global__ void __launch_bounds__(128) dwconv_bwd_weight_kernel(const float* __restrict__ dy, const T* __restrict__ x, int B, int T_in, int T_out,
                                                                int C, int K, int stride, float* __restrict__ partial) {
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  const int pad = (K - 1) / 2;
  float acc[KT + 1];
#pragma unroll
  for (int k = 0; k <= KT; ++k) acc[k] = 0.f;
  size_t r0, r1; cta_rows(static_cast<size_t>(B) * T_out, r0, r1);
  for (size_t r = r0; r < r1; ++r) {
    const int b = static_cast<int>(r / T_out), t = static_cast<int>(r - static_cast<size_t>(b) * T_out);
    const T* xb = x + static_cast<size_t>(b) * T_in * C + c;
    const float d = dy[r * C + c];
    acc[KT] += d;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const int ti = t * stride + k - pad;
      if (k < K && ti >= 0 && ti < T_in) acc[k] = fmaf(d, ActTraits<T>::from(xb[static_cast<size_t>(ti) * C]), acc[k]);
    }
  }
  float* out = partial + (static_cast<size_t>(blockIdx.x) * C + c) * (K + 1);
#pragma unroll
  for (int k = 0; k < KT; ++k) if (k < K) out[k] = acc[k];
  out[K] = acc[KT];
}
// dw[c][k] / db[c] from the partials (chunked fixed-order sum, see chunked_sum_32x32)
__global__ void __launch_bounds__(1024) dwconv_wgrad_reduce_kernel(const float* __restrict__ partial, int n_partial, int C, int K,
                                                                   float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sm[32][33];
  const int i = blockIdx.x * 32 + (threadIdx.x & 31), n = C * (K + 1);
  const float t = chunked_sum_32x32(partial, n_partial, static_cast<size_t>(n), i, i < n, sm);
  if ((threadIdx.x >> 5) == 0 && i < n) {
    const int c = i / (K + 1), k = i - c * (K + 1);
    if (k < K) dw[c * K + k] = t; else db[c] = t;
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------------
size_t conv_train_work_bytes(int C, int K) { return align_up(static_cast<size_t>(kCtas) * C * std::max(K + 1, 2) * sizeof(float), 256); }

static int ctas_for(size_t rows) { return static_cast<int>(std::min<size_t>(kCtas, std::max<size_t>(rows, 1))); }

int launch_dwconv_raw(int precision, const void* x, const float* w, const float* bias, int B, int T, int C, int K, int stride, float* y,
                      float* sums /* [2][C] */, float* work, cudaStream_t st) {
  EC_REQUIRE(K % 2 == 1 && K <= kMaxTaps && (stride == 1 || stride == 2), "depthwise conv: odd k <= 31, stride 1 or 2");
  const int T_out = (T - 1) / stride + 1;
  const int ctas = ctas_for(static_cast<size_t>(B) * T_out);
  dim3 grid(ctas, cdiv(C, 128));
  const bool small = K <= 15;                                  // tap loops are fully unrolled: 15-tap (Efficient Conformer) or 31-tap instances
  if (small) EC_DISPATCH_PREC(precision, (dwconv_raw_kernel_<ActT, 15><<<grid, 128, 0, st>>>(reinterpret_cast<const ActT*>(x), w, bias, B, T, T_out, C, K, stride, y, work)));
  else EC_DISPATCH_PREC(precision, (dwconv_raw_kernel_<ActT, 31><<<grid, 128, 0, st>>>(reinterpret_cast<const ActT*>(x), w, bias, B, T, T_out, C, K, stride, y, work)));
  EC_CUDA(cudaGetLastError());
  bn_stats_merge_kernel<<<cdiv(C, 32), 1024, 0, st>>>(work, ctas, static_cast<size_t>(B) * T_out, C, sums);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_bn_finalize(const float* sums, int C, float count, float eps, float momentum, float* mean, float* rstd, float* running_mean,
                       float* running_var, cudaStream_t st) {
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(sums, C, count, eps, momentum, mean, rstd, running_mean, running_var);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_bn_swish_fwd(int precision, const float* y, size_t rows, int C, const float* mean, const float* rstd, const float* gamma,
                        const float* beta, void* h, cudaStream_t st) {
  const int blocks = static_cast<int>(std::min<size_t>((rows * C + 255) / 256, 148 * 16));
  EC_DISPATCH_PREC(precision, (bn_swish_fwd_kernel<ActT><<<blocks, 256, 0, st>>>(y, rows, C, mean, rstd, gamma, beta, reinterpret_cast<ActT*>(h))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_bn_swish_bwd_stats(const float* y, const float* dh, size_t rows, int C, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, float* sums /* [2][C]: dbeta, dgamma */, float* work, cudaStream_t st) {
  const int ctas = ctas_for(rows);
  bn_swish_bwd_stats_kernel<<<dim3(ctas, cdiv(C, 128)), 128, 0, st>>>(y, dh, rows, C, mean, rstd, gamma, beta, work);
  EC_CUDA(cudaGetLastError());
  conv_partial_reduce_kernel<<<cdiv(2 * C, 32), 1024, 0, st>>>(work, ctas, 2, C, sums);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_bn_swish_bwd_apply(const float* y, const float* dh, size_t rows, int C, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const float* sums, float count, float* dy, cudaStream_t st) {
  const int blocks = static_cast<int>(std::min<size_t>((rows * C + 255) / 256, 148 * 16));
  bn_swish_bwd_apply_kernel<<<blocks, 256, 0, st>>>(y, dh, rows, C, mean, rstd, gamma, beta, sums, count, dy);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_dwconv_bwd(int precision, const float* dy, const void* x, const float* w, int B, int T, int C, int K, int stride, float* dx,
                      float* dw, float* db, float* work, cudaStream_t st) {
  EC_REQUIRE(K % 2 == 1 && K <= kMaxTaps && (stride == 1 || stride == 2), "depthwise conv: odd k <= 31, stride 1 or 2");
  const int T_out = (T - 1) / stride + 1;
  if (dx != nullptr) {
    const dim3 gd(static_cast<unsigned>(std::min<size_t>(static_cast<size_t>(B) * T, 148 * 32)), cdiv(C, 128));
    if (K <= 15) dwconv_bwd_data_kernel<15><<<gd, 128, 0, st>>>(dy, w, B, T, T_out, C, K, stride, dx);
    else dwconv_bwd_data_kernel<31><<<gd, 128, 0, st>>>(dy, w, B, T, T_out, C, K, stride, dx);
    EC_CUDA(cudaGetLastError());
  }
  const int ctas = ctas_for(static_cast<size_t>(B) * T_out);
  dim3 grid(ctas, cdiv(C, 128));
  if (K <= 15) EC_DISPATCH_PREC(precision, (dwconv_bwd_weight_kernel<ActT, 15><<<grid, 128, 0, st>>>(dy, reinterpret_cast<const ActT*>(x), B, T, T_out, C, K, stride, work)));
  else EC_DISPATCH_PREC(precision, (dwconv_bwd_weight_kernel<ActT, 31><<<grid, 128, 0, st>>>(dy, reinterpret_cast<const ActT*>(x), B, T, T_out, C, K, stride, work)));
  EC_CUDA(cudaGetLastError());
  dwconv_wgrad_reduce_kernel<<<cdiv(C * (K + 1), 32), 1024, 0, st>>>(work, ctas, C, K, dw, db);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec
