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
constexpr int kRun = 16;        // output frames per thread run (register sliding window over the taps)
constexpr int kMaxRunCtas = 2048;   // data-gradient runs
constexpr int kStatRunCtas = 444;   // forward runs (3 CTAs / SM): every CTA leaves one (count, mean, M2) partial per channel
constexpr int kColCtas = 592;   // 4 per SM: CTAs of the column-reduction kernels
constexpr int kMaxTaps = 31;
}  // namespace

// ---- forward: raw depthwise conv + statistics ------------------------------------------------------------------------------
// thread = channel (coalesced 128-byte rows per warp), one RUN of kRun consecutive output frames of one sequence per iteration: the
// (kRun-1)*S + K input frames of the run are loaded once into a register window and every output reads its taps from it (1 load per
// output instead of K).  Per-thread statistics of its runs are exact two-pass (values are in registers), merged with Chan's update;
// partial[cta] = (count, mean, M2) per channel.  kFlip: taps reversed, input fp32, no bias / statistics = the stride-1 data gradient.
template <typename T, int KT, int S, bool kFlip>
__global__ void __launch_bounds__(128) dwconv_run_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         int B, int T_in, int T_out, int C, int K, float* __restrict__ y,
                                                         float* __restrict__ partial /* [gridDim.x][3][C] or nullptr */) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  float wk[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) wk[k] = k < K ? w[c * K + (kFlip ? K - 1 - k : k)] : 0.f;
  const float bc = (kFlip || bias == nullptr) ? 0.f : bias[c];
  const int pad = (K - 1) / 2;
  const int runs_per_seq = (T_out + kRun - 1) / kRun;
  const int n_runs = B * runs_per_seq;
  constexpr int W = (kRun - 1) * S + KT;
  float n_acc = 0.f, mean_acc = 0.f, m2_acc = 0.f;
  for (int run = blockIdx.x; run < n_runs; run += gridDim.x) {
    const int b = run / runs_per_seq, t0 = (run - b * runs_per_seq) * kRun;
    const T* xb = x + static_cast<size_t>(b) * T_in * C + c;
    const int ti0 = t0 * S - pad;
    float xin[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int ti = ti0 + j;
      xin[j] = (ti >= 0 && ti < T_in) ? ActTraits<T>::from(xb[static_cast<size_t>(ti) * C]) : 0.f;
    }
    const int nv = min(kRun, T_out - t0);
    float acc[kRun];
    float s1 = 0.f;
#pragma unroll
    for (int r = 0; r < kRun; ++r) {
      float a = bc;
#pragma unroll
      for (int k = 0; k < KT; ++k) a = fmaf(wk[k], xin[r * S + k], a);
      acc[r] = a;
      if (r < nv) { y[(static_cast<size_t>(b) * T_out + t0 + r) * C + c] = a; s1 += a; }
    }
    if (!kFlip) {
      const float lm = s1 / static_cast<float>(nv);
      float s2 = 0.f;
#pragma unroll
      for (int r = 0; r < kRun; ++r) if (r < nv) { const float d = acc[r] - lm; s2 = fmaf(d, d, s2); }
      const float nb = static_cast<float>(nv), tot = n_acc + nb, dl = lm - mean_acc, f = nb / tot;
      mean_acc = fmaf(dl, f, mean_acc);
      m2_acc += s2 + dl * dl * n_acc * f;
      n_acc = tot;
    }
  }
  if (!kFlip && partial != nullptr) {
    float* o = partial + static_cast<size_t>(blockIdx.x) * 3 * C + c;
    o[0] = n_acc; o[C] = mean_acc; o[2 * C] = m2_acc;
  }
}

// (count, mean, M2) partials merged with Chan's update: stats[0][c] = mean, stats[1][c] = M2.
// Block = 32 channels x 32 lanes: lane ty merges the contiguous chunk ty of the partials in CTA order, then the 32 chunk results
// are merged as a fixed binary tree (chan_tree_merge_32x32: bit-reproducible; the dependent chain is n_partial / 32 + 5 steps).
__global__ void __launch_bounds__(1024) bn_stats_merge_kernel(const float* __restrict__ partial, int n_partial, int C, float* __restrict__ stats) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float sn[32][33], smean[32][33], sm2[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int chunk = (n_partial + 31) / 32, p0 = ty * chunk, p1 = min(n_partial, p0 + chunk);
  float n = 0.f, mean = 0.f, m2 = 0.f;
  if (c < C) {
    for (int p = p0; p < p1; ++p) {
      const float* q = partial + static_cast<size_t>(p) * 3 * C + c;
      const float nb = q[0];
      if (nb == 0.f) continue;
      const float mb = q[C], qb = q[2 * C];
      const float tot = n + nb, dl = mb - mean, f = nb / tot;
      mean = fmaf(dl, f, mean);
      m2 += qb + dl * dl * n * f;
      n = tot;
    }
  }
  sn[ty][tx] = n; smean[ty][tx] = mean; sm2[ty][tx] = m2;
  chan_tree_merge_32x32(sn, smean, sm2);
  if (ty == 0 && c < C) {
    stats[c] = smean[0][tx];
    stats[C + c] = sm2[0][tx];
  }
}

// out[i] = sum_p partial[p][i], i < n: block = 32 outputs x 32 lanes, lane ty adds chunk ty in CTA order, chunk sums added in lane order
__device__ __forceinline__ float chunked_sum_32x32(const float* __restrict__ partial, int n_partial, size_t stride, int i, bool ok,
                                                   float (*sm)[33]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int per = (n_partial + 31) / 32, p0 = ty * per, p1 = min(n_partial, p0 + per);
  float s = 0.f;
  if (ok) {
#pragma unroll 8
    for (int p = p0; p < p1; ++p) s += partial[static_cast<size_t>(p) * stride + i];
  }
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
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float sm[32][33];
  const int i = blockIdx.x * 32 + (threadIdx.x & 31), n = n_out * dim;
  const float t = chunked_sum_32x32(partial, n_partial, static_cast<size_t>(n), i, i < n, sm);
  if ((threadIdx.x >> 5) == 0 && i < n) out[i] = t;
}

// stats [2][C] (mean, centred sum of squares M2) over `count` frames -> mean / rstd (biased variance) and the running-statistics update
__global__ void bn_finalize_kernel(const float* __restrict__ sums, int C, float count, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ running_mean, float* __restrict__ running_var) {
  grid_dependency_wait();
  grid_launch_dependents();
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

// Element kernels over [rows, C] with per-channel parameters: a thread owns 4 consecutive channels (128-bit accesses; C % 4 == 0)
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(round_tf32(v[0]), round_tf32(v[1]), round_tf32(v[2]), round_tf32(v[3]));
}
template <> __device__ __forceinline__ void store4<SplitBf16>(SplitBf16* p, const float (&v)[4]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(split_pack(v[0]), split_pack(v[1]), split_pack(v[2]), split_pack(v[3]));
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

template <typename T>
__global__ void __launch_bounds__(256) bn_swish_fwd_kernel(const float* __restrict__ y, size_t rows, int C, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, T* __restrict__ h) {
  grid_dependency_wait();
  grid_launch_dependents();
  const size_t n4 = rows * C / 4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n4; g += stride) {
    const int c = static_cast<int>((g * 4) % C);
    const float4 yv = reinterpret_cast<const float4*>(y)[g];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    float z[4] = {(yv.x - mu.x) * rs.x * ga.x + be.x, (yv.y - mu.y) * rs.y * ga.y + be.y, (yv.z - mu.z) * rs.z * ga.z + be.z,
                  (yv.w - mu.w) * rs.w * ga.w + be.w};
#pragma unroll
    for (int l = 0; l < 4; ++l) z[l] = z[l] / (1.f + __expf(-z[l]));
    store4<T>(h + g * 4, z);
  }
}

// ---- backward ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float swish_grad(float z) { const float s = 1.f / (1.f + __expf(-z)); return s + z * s * (1.f - s); }

// Column sums of dz = dh * swish'(z) and dz * xhat.  Block = 32 channel quads x 8 row lanes; the CTA owns a contiguous row range and
// its row lane ty the rows r0 + ty, r0 + ty + 8, ... (4 independent rows in flight per thread); lanes merged in shared memory in a fixed
// order.  partial[cta.x][0][c] = sum dz, partial[cta.x][1][c] = sum dz * xhat.
__global__ void __launch_bounds__(256) bn_swish_bwd_stats_kernel(const float* __restrict__ y, const float* __restrict__ dh, size_t rows, int C,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 float* __restrict__ partial) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float sm[2][8][132];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y * 128 + tx * 4;
  const bool ok = c < C;
  float4 mu = make_float4(0, 0, 0, 0), rs = mu, ga = mu, be = mu;
  if (ok) {
    mu = *reinterpret_cast<const float4*>(mean + c); rs = *reinterpret_cast<const float4*>(rstd + c);
    ga = *reinterpret_cast<const float4*>(gamma + c); be = *reinterpret_cast<const float4*>(beta + c);
  }
  const size_t per = (rows + gridDim.x - 1) / gridDim.x;
  const size_t r0 = min(rows, per * blockIdx.x), r1 = min(rows, r0 + per);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (ok) {
    for (size_t r = r0 + ty; r < r1; r += 32) {
      float4 yv[4], dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t rr = r + 8 * u;
        if (rr < r1) { yv[u] = *reinterpret_cast<const float4*>(y + rr * C + c); dv[u] = *reinterpret_cast<const float4*>(dh + rr * C + c); }
        else { yv[u] = mu; dv[u] = make_float4(0, 0, 0, 0); }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float xh[4] = {(yv[u].x - mu.x) * rs.x, (yv[u].y - mu.y) * rs.y, (yv[u].z - mu.z) * rs.z, (yv[u].w - mu.w) * rs.w};
        const float dd[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
        const float gg[4] = {ga.x, ga.y, ga.z, ga.w}, bb[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const float dz = dd[l] * swish_grad(xh[l] * gg[l] + bb[l]);
          s1[l] += dz; s2[l] = fmaf(dz, xh[l], s2[l]);
        }
      }
    }
  }
#pragma unroll
  for (int l = 0; l < 4; ++l) { sm[0][ty][tx * 4 + l] = s1[l]; sm[1][ty][tx * 4 + l] = s2[l]; }
  __syncthreads();
  {
    const int which = threadIdx.x >> 7, col = threadIdx.x & 127;       // 256 threads = 2 outputs x 128 channels
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sm[which][q][col];
    const int cc = blockIdx.y * 128 + col;
    if (cc < C) partial[(static_cast<size_t>(blockIdx.x) * 2 + which) * C + cc] = t;
  }
}

// dy = gamma * rstd * (dz - sum_dz / R - xhat * sum_dzx / R)   (R = frames the statistics were taken over, all ranks)
__global__ void __launch_bounds__(256) bn_swish_bwd_apply_kernel(const float* __restrict__ y, const float* __restrict__ dh, size_t rows, int C,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ sums /* [2][C] */, float count,
                                                                 float* __restrict__ dy) {
  grid_dependency_wait();
  grid_launch_dependents();
  const size_t n4 = rows * C / 4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const float inv = 1.f / count;
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n4; g += stride) {
    const int c = static_cast<int>((g * 4) % C);
    const float4 yv = reinterpret_cast<const float4*>(y)[g], dv = reinterpret_cast<const float4*>(dh)[g];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    const float4 sa = *reinterpret_cast<const float4*>(sums + c), sb = *reinterpret_cast<const float4*>(sums + C + c);
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w}, mm[4] = {mu.x, mu.y, mu.z, mu.w}, rr[4] = {rs.x, rs.y, rs.z, rs.w};
    const float gg[4] = {ga.x, ga.y, ga.z, ga.w}, bb[4] = {be.x, be.y, be.z, be.w}, a1[4] = {sa.x, sa.y, sa.z, sa.w}, a2[4] = {sb.x, sb.y, sb.z, sb.w};
    float o[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float xh = (yy[l] - mm[l]) * rr[l];
      const float dz = dd[l] * swish_grad(xh * gg[l] + bb[l]);
      o[l] = gg[l] * rr[l] * (dz - a1[l] * inv - xh * a2[l] * inv);
    }
    reinterpret_cast<float4*>(dy)[g] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// dx[b, ti, c] = sum_k w[c,k] * dy[b, to, c],  to = (ti + pad - k) / s when divisible and 0 <= to < T_out  (strided blocks; the stride-1
// data gradient is the forward run kernel with reversed taps)
template <int KT>
__global__ void __launch_bounds__(128) dwconv_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, int B, int T_in,
                                                              int T_out, int C, int K, int stride, float* __restrict__ dx) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= C) return;
  float wk[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) wk[k] = k < K ? w[c * K + k] : 0.f;
  const int pad = (K - 1) / 2;
  const size_t rows = static_cast<size_t>(B) * T_in;
  const size_t per = (rows + gridDim.x - 1) / gridDim.x;
  const size_t r0 = min(rows, per * blockIdx.x), r1 = min(rows, r0 + per);
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

// Weight / bias gradient of the depthwise conv.  Block = 128 channels x 4 run lanes; every thread walks runs of kRun output frames with
// the same register window as the forward, accumulating its K tap gradients and the bias gradient in registers; the 4 run lanes are
// merged in shared memory in a fixed order.  partial[cta][c][k] (k = K: sum dy).
template <typename T, int KT, int S>
__global__ void __launch_bounds__(512) dwconv_bwd_weight_kernel(const float* __restrict__ dy, const T* __restrict__ x, int B, int T_in, int T_out,
                                                                int C, int K, float* __restrict__ partial) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float sm[3][KT + 1][128];
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
  const int c = blockIdx.y * 128 + tx;
  const int pad = (K - 1) / 2;
  const int runs_per_seq = (T_out + kRun - 1) / kRun;
  const int n_runs = B * runs_per_seq;
  constexpr int W = (kRun - 1) * S + KT;
  float acc[KT + 1];
#pragma unroll
  for (int k = 0; k <= KT; ++k) acc[k] = 0.f;
  if (c < C) {
    for (int run = blockIdx.x * 4 + ty; run < n_runs; run += gridDim.x * 4) {
      const int b = run / runs_per_seq, t0 = (run - b * runs_per_seq) * kRun;
      const T* xb = x + static_cast<size_t>(b) * T_in * C + c;
      const float* db = dy + (static_cast<size_t>(b) * T_out + t0) * C + c;
      const int ti0 = t0 * S - pad;
      const int nv = min(kRun, T_out - t0);
      float xin[W], d[kRun];
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const int ti = ti0 + j;
        xin[j] = (ti >= 0 && ti < T_in) ? ActTraits<T>::from(xb[static_cast<size_t>(ti) * C]) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < kRun; ++r) d[r] = r < nv ? db[static_cast<size_t>(r) * C] : 0.f;
#pragma unroll
      for (int r = 0; r < kRun; ++r) {
        acc[KT] += d[r];
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[k] = fmaf(d[r], xin[r * S + k], acc[k]);
      }
    }
  }
  if (ty > 0) {
#pragma unroll
    for (int k = 0; k <= KT; ++k) sm[ty - 1][k][tx] = acc[k];
  }
  __syncthreads();
  if (ty == 0 && c < C) {
    float* out = partial + (static_cast<size_t>(blockIdx.x) * C + c) * (K + 1);
#pragma unroll
    for (int k = 0; k <= KT; ++k) {
      const float t = ((acc[k] + sm[0][k][tx]) + sm[1][k][tx]) + sm[2][k][tx];
      if (k < K) out[k] = t; else if (k == KT) out[K] = t;
    }
  }
}
// dw[c][k] / db[c] from the partials (chunked fixed-order sum, see chunked_sum_32x32)
__global__ void __launch_bounds__(1024) dwconv_wgrad_reduce_kernel(const float* __restrict__ partial, int n_partial, int C, int K,
                                                                   float* __restrict__ dw, float* __restrict__ db) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float sm[32][33];
  const int i = blockIdx.x * 32 + (threadIdx.x & 31), n = C * (K + 1);
  const float t = chunked_sum_32x32(partial, n_partial, static_cast<size_t>(n), i, i < n, sm);
  if ((threadIdx.x >> 5) == 0 && i < n) {
    const int c = i / (K + 1), k = i - c * (K + 1);
    if (k < K) dw[c * K + k] = t; else db[c] = t;
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------------
size_t conv_train_work_bytes(int C, int K) {
  const size_t a = static_cast<size_t>(kStatRunCtas) * 3 * C, b = static_cast<size_t>(kColCtas) * C * std::max(K + 1, 2);
  return align_up(std::max(a, b) * sizeof(float), 256);
}

static int run_ctas(int B, int T_out) { return std::max(1, std::min(kMaxRunCtas, B * cdiv(T_out, kRun))); }
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int launch_dwconv_raw(int precision, const void* x, const float* w, const float* bias, int B, int T, int C, int K, int stride, float* y,
                      float* sums /* [2][C] */, float* work, cudaStream_t st) {
  EC_REQUIRE(K % 2 == 1 && K <= kMaxTaps && (stride == 1 || stride == 2), "depthwise conv: odd k <= 31, stride 1 or 2");
  const int T_out = (T - 1) / stride + 1;
  const int ctas = std::min(run_ctas(B, T_out), kStatRunCtas);
  dim3 grid(ctas, cdiv(C, 128));
  // tap loops are fully unrolled: 15-tap (Efficient Conformer) or 31-tap instances, stride 1 or 2
#define EC_RAW(KT, S) EC_DISPATCH_PREC(precision, ((void)launch_dep(dwconv_run_kernel<ActT, KT, S, false>, dim3(grid), dim3(128), 0, st, reinterpret_cast<const ActT*>(x), w, bias, B, T, T_out, C, K, y, work)))
  if (K <= 15) { if (stride == 1) EC_RAW(15, 1); else EC_RAW(15, 2); }
  else { if (stride == 1) EC_RAW(31, 1); else EC_RAW(31, 2); }
#undef EC_RAW
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(bn_stats_merge_kernel, dim3(cdiv(C, 32)), dim3(1024), 0, st, work, ctas, C, sums);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_bn_finalize(const float* sums, int C, float count, float eps, float momentum, float* mean, float* rstd, float* running_mean,
                       float* running_var, cudaStream_t st) {
  (void)launch_dep(bn_finalize_kernel, dim3(cdiv(C, 128)), dim3(128), 0, st, sums, C, count, eps, momentum, mean, rstd, running_mean, running_var);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_bn_swish_fwd(int precision, const float* y, size_t rows, int C, const float* mean, const float* rstd, const float* gamma,
                        const float* beta, void* h, cudaStream_t st) {
  EC_REQUIRE(C % 4 == 0 && aligned16(y) && aligned16(h) && aligned16(mean) && aligned16(rstd) && aligned16(gamma) && aligned16(beta),
             "BatchNorm + Swish: channels must be a multiple of 4 and every pointer 16-byte aligned");
  const int blocks = static_cast<int>(std::min<size_t>((rows * C / 4 + 255) / 256, 148 * 8));
  EC_DISPATCH_PREC(precision, ((void)launch_dep(bn_swish_fwd_kernel<ActT>, dim3(blocks), dim3(256), 0, st, y, rows, C, mean, rstd, gamma, beta, reinterpret_cast<ActT*>(h))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_bn_swish_bwd_stats(const float* y, const float* dh, size_t rows, int C, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, float* sums /* [2][C]: dbeta, dgamma */, float* work, cudaStream_t st) {
  EC_REQUIRE(C % 4 == 0 && aligned16(y) && aligned16(dh) && aligned16(mean) && aligned16(rstd) && aligned16(gamma) && aligned16(beta),
             "BatchNorm + Swish backward: channels must be a multiple of 4 and every pointer 16-byte aligned");
  const int gy = cdiv(C, 128);
  const int ctas = static_cast<int>(std::max<size_t>(1, std::min<size_t>(std::max(1, kColCtas / gy), (rows + 31) / 32)));
  (void)launch_dep(bn_swish_bwd_stats_kernel, dim3(dim3(ctas, gy)), dim3(256), 0, st, y, dh, rows, C, mean, rstd, gamma, beta, work);
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(conv_partial_reduce_kernel, dim3(cdiv(2 * C, 32)), dim3(1024), 0, st, work, ctas, 2, C, sums);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_bn_swish_bwd_apply(const float* y, const float* dh, size_t rows, int C, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const float* sums, float count, float* dy, cudaStream_t st) {
  EC_REQUIRE(C % 4 == 0 && aligned16(y) && aligned16(dh) && aligned16(dy) && aligned16(mean) && aligned16(rstd) && aligned16(gamma) &&
             aligned16(beta) && aligned16(sums), "BatchNorm + Swish backward: channels must be a multiple of 4 and every pointer 16-byte aligned");
  const int blocks = static_cast<int>(std::min<size_t>((rows * C / 4 + 255) / 256, 148 * 8));
  (void)launch_dep(bn_swish_bwd_apply_kernel, dim3(blocks), dim3(256), 0, st, y, dh, rows, C, mean, rstd, gamma, beta, sums, count, dy);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_dwconv_bwd(int precision, const float* dy, const void* x, const float* w, int B, int T, int C, int K, int stride, float* dx,
                      float* dw, float* db, float* work, cudaStream_t st) {
  EC_REQUIRE(K % 2 == 1 && K <= kMaxTaps && (stride == 1 || stride == 2), "depthwise conv: odd k <= 31, stride 1 or 2");
  const int T_out = (T - 1) / stride + 1;
  // the weight / bias gradient only feeds the optimiser: it runs beside the data gradient on a library side stream (fork / join)
  SideStreams& ss = side_streams();
  const bool par = dx != nullptr && side_streams_enabled() && ss.init();
  cudaStream_t sw = par ? ss.s[3] : st;
  if (par) {
    EC_CUDA(cudaEventRecord(ss.fork_ev, st));
    EC_CUDA(cudaStreamWaitEvent(sw, ss.fork_ev, 0));
  }
  if (dx != nullptr) {
    if (stride == 1) {                 // correlation with the reversed taps: the forward run kernel on the fp32 gradient
      dim3 gd(run_ctas(B, T), cdiv(C, 128));
      if (K <= 15) (void)launch_dep(dwconv_run_kernel<float, 15, 1, true>, dim3(gd), dim3(128), 0, st, dy, w, nullptr, B, T, T, C, K, dx, nullptr);
      else (void)launch_dep(dwconv_run_kernel<float, 31, 1, true>, dim3(gd), dim3(128), 0, st, dy, w, nullptr, B, T, T, C, K, dx, nullptr);
    } else {
      const dim3 gd(static_cast<unsigned>(std::min<size_t>(static_cast<size_t>(B) * T, 148 * 32)), cdiv(C, 128));
      if (K <= 15) (void)launch_dep(dwconv_bwd_data_kernel<15>, dim3(gd), dim3(128), 0, st, dy, w, B, T, T_out, C, K, stride, dx);
      else (void)launch_dep(dwconv_bwd_data_kernel<31>, dim3(gd), dim3(128), 0, st, dy, w, B, T, T_out, C, K, stride, dx);
    }
    EC_CUDA(cudaGetLastError());
  }
  const int gy = cdiv(C, 128);
  const int ctas = std::max(1, std::min(std::max(1, kColCtas / 2 / gy), cdiv(B * cdiv(T_out, kRun), 4)));
  dim3 grid(ctas, gy);
#define EC_WG(KT, S) EC_DISPATCH_PREC(precision, ((void)launch_dep(dwconv_bwd_weight_kernel<ActT, KT, S>, dim3(grid), dim3(512), 0, sw, dy, reinterpret_cast<const ActT*>(x), B, T, T_out, C, K, work)))
  if (K <= 15) { if (stride == 1) EC_WG(15, 1); else EC_WG(15, 2); }
  else { if (stride == 1) EC_WG(31, 1); else EC_WG(31, 2); }
#undef EC_WG
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(dwconv_wgrad_reduce_kernel, dim3(cdiv(C * (K + 1), 32)), dim3(1024), 0, sw, work, ctas, C, K, dw, db);
  EC_CUDA(cudaGetLastError());
  if (par) {
    EC_CUDA(cudaEventRecord(ss.join_ev[3], sw));
    EC_CUDA(cudaStreamWaitEvent(st, ss.join_ev[3], 0));
  }
  return EC_OK;
}

}  // namespace ec
