// Kernels that close the training step around the forward / backward operators (all HBM-bandwidth bound element kernels):
//   dropout family      counter-based masks (reference nn.Dropout sites: models/encoders.py:119, modules.py:389,391,486,521).  The mask
//                       bit of element i at site s of step n is a pure function of (seed, n, s, i), so the backward recomputes it
//                       instead of storing masks, and a replayed CUDA graph draws fresh masks because (seed, n) live in device memory.
//   adam_step           torch.optim.Adam semantics (L2-style weight decay added to the gradient, bias correction) over ONE flat fp32
//                       parameter / gradient / moment arena (reference models/model.py:88-93 builds optim.Adam over all parameters);
//                       the learning rate and step counters live in device memory (graph replay).
//   transformer_lr      reference models/schedules.py:99-123: lr = K * d^-0.5 * min(s^-0.5, s * warmup^-1.5), advanced on device.
//   stats_merge_ranks   SyncBatchNorm forward: Chan merge of the per-rank (mean, M2, count) triples gathered over NCCL.
#include "ec_common.cuh"
#include <algorithm>

namespace ec {

inline int grid_for(size_t n) { return static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16)); }
__global__ void dropout_advance_kernel(unsigned long long* ctr) {
  grid_dependency_wait();
  grid_launch_dependents(); ctr[1] += 1ull; }

// dst[i] = TOut(scale * keep(i) / (1 - p) * src[i]); each thread owns groups of 4 consecutive elements
template <typename TOut, bool kRound> __device__ __forceinline__ TOut drop_store(float x);
template <> __device__ __forceinline__ float drop_store<float, false>(float x) { return x; }               // true fp32 (residual stream, gradients)
template <> __device__ __forceinline__ float drop_store<float, true>(float x) { return round_tf32(x); }    // TF32-mode GEMM operand
template <> __device__ __forceinline__ __nv_bfloat16 drop_store<__nv_bfloat16, true>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ SplitBf16 drop_store<SplitBf16, true>(float x) { return SplitBf16{split_pack(x)}; }
template <typename TIn, typename TOut, bool kRound>
__global__ void __launch_bounds__(256) dropout_kernel(const TIn* __restrict__ src, float scale, size_t n, TOut* __restrict__ dst,
                                                      const unsigned long long* __restrict__ ctr, unsigned site, unsigned keep16) {
  grid_dependency_wait();
  grid_launch_dependents();
  const unsigned long long key = site_key(ctr, site);
  const float inv_keep = scale * 65536.f / static_cast<float>(keep16);
  const size_t groups = (n + 3) / 4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const unsigned long long draw = splitmix64(key + g);
    const size_t i0 = g * 4;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const size_t i = i0 + l;
      if (i < n) dst[i] = drop_store<TOut, kRound>(keep_factor(draw, l, keep16, inv_keep) * ActTraits<TIn>::from(src[i]));
    }
  }
}
// out[i] = residual[i] + alpha * keep(i) / (1 - p) * y[i]
__global__ void __launch_bounds__(256) dropout_residual_kernel(const float* __restrict__ y, const float* __restrict__ residual, float alpha,
                                                               size_t n, float* __restrict__ out, const unsigned long long* __restrict__ ctr,
                                                               unsigned site, unsigned keep16) {
  grid_dependency_wait();
  grid_launch_dependents();
  const unsigned long long key = site_key(ctr, site);
  const float inv_keep = alpha * 65536.f / static_cast<float>(keep16);
  const size_t groups = (n + 3) / 4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const unsigned long long draw = splitmix64(key + g);
    const size_t i0 = g * 4;
    if (vec) {
      const float4 a = reinterpret_cast<const float4*>(y)[g], r = reinterpret_cast<const float4*>(residual)[g];
      float4 o;
      o.x = fmaf(keep_factor(draw, 0, keep16, inv_keep), a.x, r.x);
      o.y = fmaf(keep_factor(draw, 1, keep16, inv_keep), a.y, r.y);
      o.z = fmaf(keep_factor(draw, 2, keep16, inv_keep), a.z, r.z);
      o.w = fmaf(keep_factor(draw, 3, keep16, inv_keep), a.w, r.w);
      reinterpret_cast<float4*>(out)[g] = o;
    } else {
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const size_t i = i0 + l;
        if (i < n) out[i] = fmaf(keep_factor(draw, l, keep16, inv_keep), y[i], residual[i]);
      }
    }
  }
}


int launch_dropout_advance(unsigned long long* ctr, cudaStream_t st) {
  (void)launch_dep(dropout_advance_kernel, dim3(1), dim3(1), 0, st, ctr);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_dropout(int precision, const void* src, int src_f32, float scale, size_t n, void* dst, int dst_f32, float p,
                   const unsigned long long* ctr, unsigned site, cudaStream_t st) {
  EC_REQUIRE(p >= 0.f && p < 1.f, "dropout probability must be in [0, 1)");
  const unsigned k = keep16_of(p);
  const bool sf = src_f32 || precision == EC_PREC_TF32;
  const int grid = grid_for((n + 3) / 4);
  if (n == 0) return EC_OK;
  using bf = __nv_bfloat16;
  const float* s32 = static_cast<const float*>(src); const bf* s16 = static_cast<const bf*>(src); const SplitBf16* sx = static_cast<const SplitBf16*>(src);
  float* d32 = static_cast<float*>(dst); bf* d16 = static_cast<bf*>(dst); SplitBf16* dx = static_cast<SplitBf16*>(dst);
  if (dst_f32) {
    if (sf) (void)launch_dep(dropout_kernel<float, float, false>, dim3(grid), dim3(256), 0, st, s32, scale, n, d32, ctr, site, k);
    else if (precision == EC_PREC_BF16X2) (void)launch_dep(dropout_kernel<SplitBf16, float, false>, dim3(grid), dim3(256), 0, st, sx, scale, n, d32, ctr, site, k);
    else (void)launch_dep(dropout_kernel<bf, float, false>, dim3(grid), dim3(256), 0, st, s16, scale, n, d32, ctr, site, k);
  } else if (precision == EC_PREC_TF32) {
    (void)launch_dep(dropout_kernel<float, float, true>, dim3(grid), dim3(256), 0, st, s32, scale, n, d32, ctr, site, k);
  } else if (precision == EC_PREC_BF16X2) {
    if (sf) (void)launch_dep(dropout_kernel<float, SplitBf16, true>, dim3(grid), dim3(256), 0, st, s32, scale, n, dx, ctr, site, k);
    else (void)launch_dep(dropout_kernel<SplitBf16, SplitBf16, true>, dim3(grid), dim3(256), 0, st, sx, scale, n, dx, ctr, site, k);
  } else {
    if (sf) (void)launch_dep(dropout_kernel<float, bf, true>, dim3(grid), dim3(256), 0, st, s32, scale, n, d16, ctr, site, k);
    else (void)launch_dep(dropout_kernel<bf, bf, true>, dim3(grid), dim3(256), 0, st, s16, scale, n, d16, ctr, site, k);
  }
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_dropout_residual(const float* y, const float* residual, float alpha, size_t n, float* out, float p, const unsigned long long* ctr,
                            unsigned site, cudaStream_t st) {
  EC_REQUIRE(p >= 0.f && p < 1.f, "dropout probability must be in [0, 1)");
  if (n == 0) return EC_OK;
  (void)launch_dep(dropout_residual_kernel, dim3(grid_for((n + 3) / 4)), dim3(256), 0, st, y, residual, alpha, n, out, ctr, site, keep16_of(p));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---- Adam over a flat arena ------------------------------------------------------------------------------------------
// state (device): [0] lr (float bits), [1] adam step t (int), [2] schedule step s (int), [3] reserved
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, size_t n, const int* __restrict__ state, float beta1, float beta2,
                                                   float eps, float weight_decay, float grad_scale) {
  grid_dependency_wait();
  grid_launch_dependents();
  const float lr = __int_as_float(state[0]);
  const float t = static_cast<float>(state[1] + 1);
  const float bc1 = 1.f - powf(beta1, t), bc2 = 1.f - powf(beta2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const size_t n4 = n / 4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float* pe = &pp.x; float* me = &mm.x; float* ve = &vv.x; const float* ge = &gg.x;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float gr = fmaf(weight_decay, pe[l], grad_scale * ge[l]);
      me[l] = fmaf(1.f - beta1, gr - me[l], me[l]);                      // exp_avg.lerp_(grad, 1 - beta1)
      ve[l] = fmaf(1.f - beta2, gr * gr, beta2 * ve[l]);
      pe[l] -= step_size * me[l] / (sqrtf(ve[l]) * inv_sqrt_bc2 + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (size_t i = n4 * 4 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gr = fmaf(weight_decay, p[i], grad_scale * g[i]);
    const float mi = fmaf(1.f - beta1, gr - m[i], m[i]);
    const float vi = fmaf(1.f - beta2, gr * gr, beta2 * v[i]);
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}
// after the update: t += 1; schedule mode 1 (Transformer): s += 1, lr = K * d^-0.5 * min(s^-0.5, s * warmup^-1.5); mode 0: lr unchanged
__global__ void adam_advance_kernel(int* state, int mode, float K, float dim, float warmup) {
  grid_dependency_wait();
  grid_launch_dependents();
  state[1] += 1;
  if (mode == 1) {
    state[2] += 1;
    const double s = static_cast<double>(state[2]);
    const double lr = static_cast<double>(K) * pow(static_cast<double>(dim), -0.5) * fmin(pow(s, -0.5), s * pow(static_cast<double>(warmup), -1.5));
    state[0] = __float_as_int(static_cast<float>(lr));
  }
}
int launch_adam(float* p, const float* g, float* m, float* v, size_t n, int* state, float beta1, float beta2, float eps, float weight_decay,
                float grad_scale, int schedule, float K, float dim, float warmup, cudaStream_t st) {
  EC_REQUIRE(p && g && m && v && state, "null argument");
  EC_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
             "Adam arenas must be 16-byte aligned");
  if (n) (void)launch_dep(adam_kernel, dim3(grid_for((n + 3) / 4)), dim3(256), 0, st, p, g, m, v, n, state, beta1, beta2, eps, weight_decay, grad_scale);
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(adam_advance_kernel, dim3(1), dim3(1), 0, st, state, schedule, K, dim, warmup);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---- SyncBatchNorm forward: merge the (mean, M2) pairs of `world` ranks (Chan et al.), rank r holding counts[r] frames -----------
__global__ void stats_merge_ranks_kernel(const float* __restrict__ gathered, const float* __restrict__ counts, int world, int C,
                                         float* __restrict__ out) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int r = 0; r < world; ++r) {
    const double nr = counts[r];
    if (nr <= 0.0) continue;
    const double mr = gathered[(static_cast<size_t>(r) * 2) * C + c], qr = gathered[(static_cast<size_t>(r) * 2 + 1) * C + c];
    const double tot = n + nr, delta = mr - mean;
    m2 += qr + delta * delta * n * nr / tot;
    mean += delta * nr / tot;
    n = tot;
  }
  out[c] = static_cast<float>(mean);
  out[C + c] = static_cast<float>(m2);
}
int launch_stats_merge_ranks(const float* gathered, const float* counts, int world, int C, float* out, cudaStream_t st) {
  EC_REQUIRE(gathered && counts && out && world >= 1 && C >= 1, "bad argument");
  (void)launch_dep(stats_merge_ranks_kernel, dim3((C + 127) / 128), dim3(128), 0, st, gathered, counts, world, C, out);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec

using namespace ec;
#define EC_ST(s) reinterpret_cast<cudaStream_t>(s)
extern "C" {
int ec_op_dropout_advance(unsigned long long* counter, void* stream) {
  EC_REQUIRE(counter, "null argument");
  return launch_dropout_advance(counter, EC_ST(stream));
}
int ec_op_dropout(int precision, const void* src, int src_f32, float scale, size_t n, void* dst, int dst_f32, float p,
                  const unsigned long long* counter, unsigned site, void* stream) {
  EC_REQUIRE(src && dst && counter, "null argument");
  return launch_dropout(precision, src, src_f32, scale, n, dst, dst_f32, p, counter, site, EC_ST(stream));
}
int ec_op_dropout_residual(const float* y, const float* residual, float alpha, size_t n, float* out, float p,
                           const unsigned long long* counter, unsigned site, void* stream) {
  EC_REQUIRE(y && residual && out && counter, "null argument");
  return launch_dropout_residual(y, residual, alpha, n, out, p, counter, site, EC_ST(stream));
}
int ec_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, int* state, float beta1, float beta2,
                 float eps, float weight_decay, float grad_scale, int schedule, float sched_k, float sched_dim, float sched_warmup,
                 void* stream) {
  return launch_adam(params, grads, exp_avg, exp_avg_sq, n, state, beta1, beta2, eps, weight_decay, grad_scale, schedule, sched_k, sched_dim,
                     sched_warmup, EC_ST(stream));
}
int ec_op_stats_merge_ranks(const float* gathered, const float* counts, int world, int channels, float* out, void* stream) {
  return launch_stats_merge_ranks(gathered, counts, world, channels, out, EC_ST(stream));
}
}

// ---- pack: gather n separately allocated fp32 tensors into one flat arena (gradient bucket for the all-reduce / Adam) ----------
namespace ec {
template <bool kAcc>
__global__ void __launch_bounds__(256) pack_flat_kernel(const float* const* __restrict__ srcs, const long long* __restrict__ offsets,
                                                        const long long* __restrict__ sizes, float* __restrict__ arena) {
  grid_dependency_wait();
  grid_launch_dependents();
  const float* __restrict__ src = srcs[blockIdx.y];
  float* __restrict__ dst = arena + offsets[blockIdx.y];
  const long long n = sizes[blockIdx.y];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = kAcc ? dst[i] + src[i] : src[i];
}
}  // namespace ec
extern "C" int ec_op_pack_flat_acc(const float* const* srcs, const long long* offsets, const long long* sizes, int n, float* arena, int accumulate,
                                   void* stream) {
  EC_REQUIRE(srcs && offsets && sizes && arena && n >= 0 && n <= 65535, "bad argument");
  if (n == 0) return EC_OK;
  if (accumulate) (void)launch_dep(ec::pack_flat_kernel<true>, dim3(dim3(8, n)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), srcs, offsets, sizes, arena);
  else (void)launch_dep(ec::pack_flat_kernel<false>, dim3(dim3(8, n)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), srcs, offsets, sizes, arena);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
extern "C" int ec_op_pack_flat(const float* const* srcs, const long long* offsets, const long long* sizes, int n, float* arena, void* stream) {
  return ec_op_pack_flat_acc(srcs, offsets, sizes, n, arena, 0, stream);
}

// ---- Swish fused with its dropout (feed-forward module, reference models/modules.py:388-389) --------------------------------------
//   forward : h  = act_type(keep / (1 - p) * z * sigmoid(z))
//   backward: dz = act_type(keep / (1 - p) * dy * d/dz (z sigmoid z))      (same mask: same (step, site, element) hash)
namespace ec {
template <typename T, bool kBwd>
__global__ void __launch_bounds__(256) swish_dropout_kernel(const T* __restrict__ z, const float* __restrict__ dy, size_t n, T* __restrict__ out,
                                                            const unsigned long long* __restrict__ ctr, unsigned site, unsigned keep16) {
  grid_dependency_wait();
  grid_launch_dependents();
  const unsigned long long key = site_key(ctr, site);
  const float inv_keep = 65536.f / static_cast<float>(keep16);
  const size_t groups = (n + 3) / 4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const unsigned long long draw = splitmix64(key + g);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const size_t i = g * 4 + l;
      if (i >= n) break;
      const float zz = ActTraits<T>::from(z[i]);
      const float s = 1.f / (1.f + __expf(-zz));
      const float v = kBwd ? dy[i] * (s + zz * s * (1.f - s)) : zz * s;
      out[i] = ActTraits<T>::to(keep_factor(draw, l, keep16, inv_keep) * v);
    }
  }
}
// ---- multi-tensor transposed cast: for each descriptor (src offset, rows, cols, dst offset) of the flat fp32 parameter arena,
//      dst[c][r] = act_type(src[r][c]) in the transposed operand arena (the W^T operands of the data-gradient GEMMs, one launch) ----
template <typename T>
__global__ void __launch_bounds__(256) transpose_cast_multi_kernel(const float* __restrict__ src_arena, const long long* __restrict__ desc,
                                                                   T* __restrict__ dst_arena) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float tile[32][33];
  const long long* d = desc + static_cast<long long>(blockIdx.y) * 4;
  const int rows = static_cast<int>(d[1]), cols = static_cast<int>(d[2]);
  const int tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* __restrict__ src = src_arena + d[0];
  T* __restrict__ dst = dst_arena + d[3];
  for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int c0 = (t % tiles_c) * 32, r0 = (t / tiles_c) * 32;
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      tile[i][tx] = (r < rows && c < cols) ? src[static_cast<size_t>(r) * cols + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;
      if (c < cols && r < rows) {
        const T v = ActTraits<T>::to(tile[tx][i]);
        dst[static_cast<size_t>(c) * rows + r] = v;
        if constexpr (IsSplit<T>::value)      // swapped plane right behind the [cols, rows] operand
          dst[static_cast<size_t>(cols) * rows + static_cast<size_t>(c) * rows + r] = SplitBf16{split_swap(v.bits)};
      }
    }
    __syncthreads();
  }
}
// ---- multi-tensor weight cast: the [N, K] forward operands of all GEMM weights in one launch (same descriptors; split mode writes
//      the swapped plane right behind each tensor, so the destination offsets are twice the source offsets there) ----
template <typename T>
__global__ void __launch_bounds__(256) cast_multi_kernel(const float* __restrict__ src_arena, const long long* __restrict__ desc,
                                                         T* __restrict__ dst_arena) {
  grid_dependency_wait();
  grid_launch_dependents();
  const long long* d = desc + static_cast<long long>(blockIdx.y) * 4;
  const size_t n = static_cast<size_t>(d[1]) * static_cast<size_t>(d[2]);
  const float* __restrict__ src = src_arena + d[0];
  T* __restrict__ dst = dst_arena + d[3];
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const T v = ActTraits<T>::to(src[i]);
    dst[i] = v;
    if constexpr (IsSplit<T>::value) dst[n + i] = SplitBf16{split_swap(v.bits)};
  }
}
}  // namespace ec
extern "C" {
int ec_op_swish_dropout(int precision, const void* z, const float* dy, size_t n, void* out, float p, const unsigned long long* counter,
                        unsigned site, void* stream) {
  EC_REQUIRE(z && out && counter && p >= 0.f && p < 1.f, "bad argument");
  if (n == 0) return EC_OK;
  const int grid = ec::grid_for((n + 3) / 4);
  const unsigned k = ec::keep16_of(p);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dy) EC_DISPATCH_PREC(precision, ((void)launch_dep(ec::swish_dropout_kernel<ActT, true>, dim3(grid), dim3(256), 0, st, static_cast<const ActT*>(z), dy, n, static_cast<ActT*>(out), counter, site, k)));
  else EC_DISPATCH_PREC(precision, ((void)launch_dep(ec::swish_dropout_kernel<ActT, false>, dim3(grid), dim3(256), 0, st, static_cast<const ActT*>(z), dy, n, static_cast<ActT*>(out), counter, site, k)));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int ec_op_transpose_cast_multi(int precision, const float* src_arena, const long long* desc, int n, int ctas_per_tensor, void* dst_arena,
                               void* stream) {
  EC_REQUIRE(src_arena && desc && dst_arena && n >= 0 && n <= 65535 && ctas_per_tensor >= 1, "bad argument");
  if (n == 0) return EC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(ctas_per_tensor, n);
  EC_DISPATCH_PREC(precision, ((void)launch_dep(ec::transpose_cast_multi_kernel<ActT>, dim3(grid), dim3(256), 0, st, src_arena, desc, static_cast<ActT*>(dst_arena))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int ec_op_cast_multi(int precision, const float* src_arena, const long long* desc, int n, int ctas_per_tensor, void* dst_arena, void* stream) {
  EC_REQUIRE(src_arena && desc && dst_arena && n >= 0 && n <= 65535 && ctas_per_tensor >= 1, "bad argument");
  if (n == 0) return EC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(ctas_per_tensor, n);
  EC_DISPATCH_PREC(precision, ((void)launch_dep(ec::cast_multi_kernel<ActT>, dim3(grid), dim3(256), 0, st, src_arena, desc, static_cast<ActT*>(dst_arena))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
}
