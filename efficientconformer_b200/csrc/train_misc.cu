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
  grid_dependency_wait();
  grid_launch_dependents();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = ActTraits<T>::from(z[i]);
    h[i] = ActTraits<T>::to(v / (1.f + __expf(-v)));
  }
}
template <typename T>
__global__ void __launch_bounds__(256) glu_fwd_kernel(const T* __restrict__ zg, size_t rows, int C, T* __restrict__ out) {
  grid_dependency_wait();
  grid_launch_dependents();
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
  grid_dependency_wait();
  grid_launch_dependents();
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
  grid_dependency_wait();
  grid_launch_dependents();
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
  grid_dependency_wait();
  grid_launch_dependents();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = ActTraits<T>::to(scale * src[i]);
}
static int grid_for(size_t n);
int launch_cast_scaled(int precision, const float* src, float scale, size_t n, void* dst, cudaStream_t st) {
  EC_DISPATCH_PREC(precision, ((void)launch_dep(cast_scaled_kernel<ActT>, dim3(grid_for(n)), dim3(256), 0, st, src, scale, n, reinterpret_cast<ActT*>(dst))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
static int grid_for(size_t n) { return static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16)); }

int launch_swish_fwd(int precision, const void* z, size_t n, void* h, cudaStream_t st) {
  EC_DISPATCH_PREC(precision, ((void)launch_dep(swish_fwd_kernel<ActT>, dim3(grid_for(n)), dim3(256), 0, st, reinterpret_cast<const ActT*>(z), n, reinterpret_cast<ActT*>(h))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_glu_fwd(int precision, const void* zg, size_t rows, int C, void* out, cudaStream_t st) {
  EC_DISPATCH_PREC(precision, ((void)launch_dep(glu_fwd_kernel<ActT>, dim3(grid_for(rows * C)), dim3(256), 0, st, reinterpret_cast<const ActT*>(zg), rows, C, reinterpret_cast<ActT*>(out))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_strided_rows(int precision, const float* x, int B, int T_in, int D, int stride, void* out, cudaStream_t st) {
  const int T_out = (T_in - 1) / stride + 1;
  const size_t n = static_cast<size_t>(B) * T_out * D;
  EC_DISPATCH_PREC(precision, ((void)launch_dep(strided_rows_kernel<ActT>, dim3(grid_for(n)), dim3(256), 0, st, x, B, T_in, T_out, D, stride, reinterpret_cast<ActT*>(out))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_strided_rows_bwd(const float* d, int B, int T_in, int D, int stride, float* dx, cudaStream_t st) {
  const int T_out = (T_in - 1) / stride + 1;
  (void)launch_dep(strided_rows_bwd_kernel, dim3(grid_for(static_cast<size_t>(B) * T_out * D)), dim3(256), 0, st, d, B, T_in, T_out, D, stride, dx);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---- per-column statistics of an fp32 matrix: stats[0][c] = mean, stats[1][c] = centred sum of squares (two-pass per CTA + Chan) ----
// thread = 4 consecutive columns (128-bit loads), block = 32 column quads x 8 row lanes over a contiguous row range; two passes over the
// CTA's rows (mean, then centred squares: sum of squares minus mean^2 loses every digit when |mean| >> std), 4 rows in flight per
// thread; the 8 row lanes are Chan-merged in shared memory in lane order.  partial[cta.x] = (count, mean, M2) per column.
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ y, size_t rows, int cols, float* __restrict__ partial) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float sm[3][8][132];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y * 128 + tx * 4;
  const bool ok = c < cols;
  const size_t per = (rows + gridDim.x - 1) / gridDim.x;
  const size_t r0 = min(rows, per * blockIdx.x), r1 = min(rows, r0 + per);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  int cnt = 0;
  if (ok) {
    for (size_t r = r0 + ty; r < r1; r += 32) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const size_t rr = r + 8 * u; v[u] = rr < r1 ? *reinterpret_cast<const float4*>(y + rr * cols + c) : make_float4(0, 0, 0, 0); cnt += rr < r1; }
#pragma unroll
      for (int u = 0; u < 4; ++u) { s1[0] += v[u].x; s1[1] += v[u].y; s1[2] += v[u].z; s1[3] += v[u].w; }
    }
    float lm[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) lm[l] = cnt > 0 ? s1[l] / static_cast<float>(cnt) : 0.f;
    for (size_t r = r0 + ty; r < r1; r += 32) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const size_t rr = r + 8 * u; v[u] = rr < r1 ? *reinterpret_cast<const float4*>(y + rr * cols + c) : make_float4(lm[0], lm[1], lm[2], lm[3]); }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float d0 = v[u].x - lm[0], d1 = v[u].y - lm[1], d2 = v[u].z - lm[2], d3 = v[u].w - lm[3];
        s2[0] = fmaf(d0, d0, s2[0]); s2[1] = fmaf(d1, d1, s2[1]); s2[2] = fmaf(d2, d2, s2[2]); s2[3] = fmaf(d3, d3, s2[3]);
      }
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) { sm[0][ty][tx * 4 + l] = static_cast<float>(cnt); sm[1][ty][tx * 4 + l] = lm[l]; sm[2][ty][tx * 4 + l] = s2[l]; }
  } else {
#pragma unroll
    for (int l = 0; l < 4; ++l) { sm[0][ty][tx * 4 + l] = 0.f; sm[1][ty][tx * 4 + l] = 0.f; sm[2][ty][tx * 4 + l] = 0.f; }
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int col = threadIdx.x, cc = blockIdx.y * 128 + col;
    float n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float nb = sm[0][q][col];
      if (nb == 0.f) continue;
      const float tot = n + nb, dl = sm[1][q][col] - mean, f = nb / tot;
      mean = fmaf(dl, f, mean);
      m2 += sm[2][q][col] + dl * dl * n * f;
      n = tot;
    }
    if (cc < cols) {
      float* o = partial + static_cast<size_t>(blockIdx.x) * 3 * cols + cc;
      o[0] = n; o[cols] = mean; o[2 * static_cast<size_t>(cols)] = m2;
    }
  }
}
// (count, mean, M2) partials -> stats[0][c] = mean, stats[1][c] = M2: block = 32 columns x 32 lanes, lane ty merges the contiguous chunk
// ty of the partials in order, the 32 chunk results are merged as a fixed binary tree (chan_tree_merge_32x32)
__global__ void __launch_bounds__(1024) col_stats_merge_kernel(const float* __restrict__ partial, int n_partial, int cols, float* __restrict__ stats) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float sn[32][33], smean[32][33], sm2[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int chunk = (n_partial + 31) / 32, p0 = ty * chunk, p1 = min(n_partial, p0 + chunk);
  float n = 0.f, mean = 0.f, m2 = 0.f;
  if (c < cols) {
    for (int p = p0; p < p1; ++p) {
      const float* q = partial + static_cast<size_t>(p) * 3 * cols + c;
      const float nb = q[0];
      if (nb == 0.f) continue;
      const float tot = n + nb, dl = q[cols] - mean, f = nb / tot;
      mean = fmaf(dl, f, mean);
      m2 += q[2 * static_cast<size_t>(cols)] + dl * dl * n * f;
      n = tot;
    }
  }
  sn[ty][tx] = n; smean[ty][tx] = mean; sm2[ty][tx] = m2;
  chan_tree_merge_32x32(sn, smean, sm2);
  if (ty == 0 && c < cols) {
    stats[c] = smean[0][tx];
    stats[cols + c] = sm2[0][tx];
  }
}
// channel c owns the `group` consecutive columns c*group .. : merge their (mean, M2) (each over `rows` samples) -> [2][C]
__global__ void group_stats_merge_kernel(const float* __restrict__ col_stats, int C, int group, double rows, float* __restrict__ ch_stats) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int cols = C * group;
  // equal counts: mean of the column means, M2 = sum of the column M2 + rows * sum (column mean - mean)^2 (no dependent division chain)
  double sum = 0.0, m2 = 0.0;
  for (int f = 0; f < group; ++f) { sum += col_stats[c * group + f]; m2 += col_stats[cols + c * group + f]; }
  const double mean = sum / group;
  double dev = 0.0;
  for (int f = 0; f < group; ++f) { const double d = col_stats[c * group + f] - mean; dev += d * d; }
  ch_stats[c] = static_cast<float>(mean); ch_stats[C + c] = static_cast<float>(m2 + rows * dev);
}
// out[j][c*group + f] = in[j][c]  (per-channel vectors expanded to per-column vectors);  n_vec stacked vectors
__global__ void group_expand_kernel(const float* __restrict__ in, int n_vec, int C, int group, float* __restrict__ out) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec * C * group) return;
  const int j = i / (C * group), col = i - j * C * group;
  out[i] = in[j * C + col / group];
}
// out[j][c] = sum_f in[j][c*group + f]  (column sums folded into channel sums, fixed order)
__global__ void group_sum_kernel(const float* __restrict__ in, int n_vec, int C, int group, float* __restrict__ out) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec * C) return;
  const int j = i / C, c = i - j * C;
  float s = 0.f;
  for (int f = 0; f < group; ++f) s += in[static_cast<size_t>(j) * C * group + c * group + f];
  out[i] = s;
}

static constexpr int kColStatCtas = 592;
static int col_stat_ctas(size_t rows, int cols) {
  const int gy = cdiv(cols, 128);
  return static_cast<int>(std::max<size_t>(1, std::min<size_t>(std::max(1, kColStatCtas / gy), (rows + 31) / 32)));
}
size_t col_stats_work_bytes(int cols) { return align_up(static_cast<size_t>(kColStatCtas) * 3 * cols * sizeof(float), 256); }
int launch_col_stats(const float* y, size_t rows, int cols, float* stats, float* work, cudaStream_t st) {
  EC_REQUIRE(cols % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "column statistics: columns must be a multiple of 4, 16-byte aligned rows");
  const int ctas = col_stat_ctas(rows, cols);
  (void)launch_dep(col_stats_kernel, dim3(dim3(ctas, cdiv(cols, 128))), dim3(256), 0, st, y, rows, cols, work);
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(col_stats_merge_kernel, dim3(cdiv(cols, 32)), dim3(1024), 0, st, work, ctas, cols, stats);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_group_stats_merge(const float* col_stats, int C, int group, size_t rows, float* ch_stats, cudaStream_t st) {
  (void)launch_dep(group_stats_merge_kernel, dim3(cdiv(C, 128)), dim3(128), 0, st, col_stats, C, group, static_cast<double>(rows), ch_stats);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_group_expand(const float* in, int n_vec, int C, int group, float* out, cudaStream_t st) {
  (void)launch_dep(group_expand_kernel, dim3(cdiv(n_vec * C * group, 256)), dim3(256), 0, st, in, n_vec, C, group, out);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_group_sum(const float* in, int n_vec, int C, int group, float* out, cudaStream_t st) {
  (void)launch_dep(group_sum_kernel, dim3(cdiv(n_vec * C, 256)), dim3(256), 0, st, in, n_vec, C, group, out);
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
  grid_dependency_wait();
  grid_launch_dependents();
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
// Block = CPB channels x F2 frequency threads (CPB = 128 / F2), CTA = a contiguous row range: thread (c, f) accumulates its 9 tap
// gradients and the bias gradient over the rows, the F2 threads of a channel are then summed in shared memory in frequency order:
// partial[cta.x][c][10] (the round-1 kernel wrote one partial per COLUMN: 114 MB of partials at C = 120).
__global__ void __launch_bounds__(128) subsample_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ melT, int B, int F, int T_in,
                                                              int T_out, int C, int cpb, float* __restrict__ partial) {
  grid_dependency_wait();
  grid_launch_dependents();
  extern __shared__ float wg_sm[];              // [cpb * F2][10]
  const int F2 = F / 2, cols = C * F2, Fp = F + 2, Tp = T_in + 2;
  const int nthr = cpb * F2;
  const int c_local = threadIdx.x / F2, f = threadIdx.x - c_local * F2;
  const int c = blockIdx.y * cpb + c_local;
  const bool ok = threadIdx.x < nthr && c < C;
  const int col = c * F2 + f;
  float acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k] = 0.f;
  const size_t rows = static_cast<size_t>(B) * T_out;
  const size_t per = (rows + gridDim.x - 1) / gridDim.x;
  const size_t r0 = min(rows, per * blockIdx.x), r1 = min(rows, r0 + per);
  if (ok && r0 < r1) {
    // (utterance, frame) of the row advance with the loop: no 64-bit division per row (it was most of this kernel's instructions)
    int b = static_cast<int>(r0 / T_out), t = static_cast<int>(r0 - static_cast<size_t>(b) * T_out);
    const float* dp = dy + r0 * cols + col;
    // rows 2t-1 .. 2t+1 of the input are rows 2t .. 2t+2 of the bordered copy (2t+2 <= T_in+1 always); same for the frequencies
    const float* base = melT + (static_cast<size_t>(b) * Tp + 2 * t) * Fp + 2 * f;
    for (size_t r = r0; r < r1; ++r) {
      const float d = *dp;
      float m[9];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) m[kw * 3 + kh] = base[kw * Fp + kh];
      acc[9] += d;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) acc[kh * 3 + kw] = fmaf(d, m[kw * 3 + kh], acc[kh * 3 + kw]);
      dp += cols;
      if (++t == T_out) { t = 0; ++b; base = melT + static_cast<size_t>(b) * Tp * Fp + 2 * f; }
      else base += 2 * Fp;
    }
  }
  if (threadIdx.x < nthr) {
#pragma unroll
    for (int k = 0; k < 10; ++k) wg_sm[threadIdx.x * 10 + k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < cpb * 10) {
    const int cl = threadIdx.x / 10, k = threadIdx.x - cl * 10, cc = blockIdx.y * cpb + cl;
    if (cc < C) {
      float s_ = 0.f;
      for (int ff = 0; ff < F2; ++ff) s_ += wg_sm[(cl * F2 + ff) * 10 + k];
      partial[(static_cast<size_t>(blockIdx.x) * C + cc) * 10 + k] = s_;
    }
  }
}
// out[c][k] = sum over the CTA partials in order: block = 32 outputs x 32 lanes (chunked fixed-order sum)
__global__ void __launch_bounds__(1024) subsample_wgrad_reduce_kernel(const float* __restrict__ partial, int n_partial, int C,
                                                                      float* __restrict__ dw, float* __restrict__ db) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx, n = C * 10;
  const int per = (n_partial + 31) / 32, p0 = ty * per, p1 = min(n_partial, p0 + per);
  float s_ = 0.f;
  if (i < n)
    for (int p = p0; p < p1; ++p) s_ += partial[static_cast<size_t>(p) * n + i];
  sm[ty][tx] = s_;
  __syncthreads();
  if (ty == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) t += sm[q][tx];
    const int c = i / 10, k = i - c * 10;
    if (k < 9) dw[c * 9 + k] = t; else db[c] = t;
  }
}
static constexpr int kWgCtas = 592;
static size_t melT_bytes(int B, int F, int T) { return align_up(static_cast<size_t>(B) * (T + 2) * (F + 2) * sizeof(float), 256); }
static size_t wg_partial_bytes(int C, int F) { (void)F; return align_up(static_cast<size_t>(kWgCtas) * C * 10 * sizeof(float), 256); }
size_t subsample_wgrad_work_bytes(int C, int F, int B, int T) { return wg_partial_bytes(C, F) + melT_bytes(B, F, T); }
int launch_subsample_wgrad(const float* dy, const float* mel, int B, int F, int T, int C, float* dw, float* db, float* work, cudaStream_t st) {
  EC_REQUIRE(F % 2 == 0, "Conv2d subsampling weight gradient: even number of mel bins");
  const int T_out = (T - 1) / 2 + 1, cols = C * (F / 2);
  const int ctas = static_cast<int>(std::min<size_t>(kWgCtas, static_cast<size_t>(B) * T_out));
  float* melT = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(work) + wg_partial_bytes(C, F));
  (void)launch_dep(subsample_melT_kernel, dim3(grid_for(static_cast<size_t>(B) * (T + 2) * (F + 2))), dim3(256), 0, st, mel, B, F, T, melT);
  EC_CUDA(cudaGetLastError());
  const int F2 = F / 2;
  EC_REQUIRE(F2 <= 128, "Conv2d subsampling weight gradient: at most 256 mel bins");
  const int cpb = std::max(1, 128 / F2), gy = cdiv(C, cpb);
  const int ctas_x = std::max(1, std::min(ctas, kWgCtas));      // row ranges of ~27 rows: 2.9 M threads in flight, 2.8 MB of partials
  (void)cols;
  (void)launch_dep(subsample_wgrad_kernel, dim3(dim3(ctas_x, gy)), dim3(128), static_cast<size_t>(cpb) * F2 * 10 * sizeof(float), st, dy, melT, B, F, T, T_out, C, cpb, work);
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(subsample_wgrad_reduce_kernel, dim3(cdiv(C * 10, 32)), dim3(1024), 0, st, work, ctas_x, C, dw, db);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec
