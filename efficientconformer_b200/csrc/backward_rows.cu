// Row / element kernels of the training backward pass, bandwidth bound:
//   LayerNorm backward          (nn.LayerNorm(eps=1e-6): reference models/modules.py:386,433,511; blocks.py:96)
//   column sums (bias gradients of every Linear / pointwise conv: reference models/layers.py:67,136)
//   transposed operand-type copies (W^T for the data-gradient GEMMs, which reuse the forward tcgen05 GEMM: dX = dY . W)
//   Swish / GLU backward         (reference models/activations.py:28-29, 37-39)
// Every reduction has a fixed order (per-CTA partials + a second pass), so gradients are bit-reproducible.
#include "ec_common.cuh"
#include <cstdlib>
#include <algorithm>
#include <type_traits>

namespace ec {

constexpr int kBwdCtas = 592;     // 4 per SM: row-strided persistent CTAs for the column reductions

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm backward.  y = (x - mu) * rstd * gamma + beta  (statistics recomputed from x: cheaper than storing them).
//   g = dy * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat));  dgamma = sum_rows dy * xhat;  dbeta = sum_rows dy
// One warp per row, the row in registers (dim <= 32 * NPL); a warp walks rows blockIdx.x*8 + warp, + 8*gridDim.x, ... TWO at a time
// (both rows' loads are issued before the first reduction: the walk is latency bound); per-lane column partials stay in registers
// over the whole walk, are merged across the 8 warps in shared memory and written as per-CTA partial rows; partial_reduce_kernel
// adds the partial rows in CTA order.
// Optional second output (TE != void): the x-gradient AFTER the residual accumulation, scaled, with the dropout mask of `site`
// re-applied, in the activation type -- the operand of the next data-gradient / weight-gradient GEMM of the backward chain (what
// ec_op_dropout computed from dx in a separate pass).
// ---------------------------------------------------------------------------------------------------------------
struct LnEmit { void* out; float scale; const unsigned long long* ctr; unsigned site, keep16; };

template <typename TE> __device__ __forceinline__ void emit_store(void* out, size_t i, float v) { reinterpret_cast<TE*>(out)[i] = ActTraits<TE>::to(v); }
template <> __device__ __forceinline__ void emit_store<void>(void*, size_t, float) {}

template <int NPL, typename TE>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int rows, int dim,
                                                            const float* __restrict__ gamma, float eps, float* __restrict__ dx,
                                                            int accumulate, float* __restrict__ partial /* [gridDim.x][2][dim] */, const LnEmit em) {
  grid_dependency_wait();
  grid_launch_dependents();
  constexpr bool kEmit = !std::is_same<TE, void>::value;
  __shared__ float red[8][32 * NPL];        // cross-warp merge buffer, used for dgamma then dbeta
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gm[NPL], pg[NPL], pb[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    gm[i] = c < dim ? __ldg(gamma + c) : 0.f;
    pg[i] = 0.f; pb[i] = 0.f;
  }
  unsigned long long key = 0; float inv_keep = 0.f;
  if (kEmit) {
    inv_keep = em.scale;
    if (em.ctr != nullptr) { key = site_key(em.ctr, em.site); inv_keep = em.scale * 65536.f / static_cast<float>(em.keep16); }
  }
  const float inv_dim = 1.f / dim;
  const int step = 8 * gridDim.x;
  for (int row0 = blockIdx.x * 8 + warp; row0 < rows; row0 += 2 * step) {
    float v[2][NPL], d[2][NPL], a[2][NPL];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = row0 + h * step;
      const bool live = row < rows;
      const float* xr = x + static_cast<size_t>(row) * dim;
      const float* dr = dy + static_cast<size_t>(row) * dim;
      const float* ar = dx + static_cast<size_t>(row) * dim;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int c = lane + 32 * i;
        const bool ok = live && c < dim;
        v[h][i] = ok ? xr[c] : 0.f;
        d[h][i] = ok ? dr[c] : 0.f;
        a[h][i] = (ok && accumulate) ? ar[c] : 0.f;
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = row0 + h * step;
      if (row >= rows) break;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) sum += v[h][i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mu = sum * inv_dim;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int c = lane + 32 * i;
        const float t = c < dim ? v[h][i] - mu : 0.f;
        sq += t * t;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq * inv_dim + eps);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int c = lane + 32 * i;
        const float xh = c < dim ? (v[h][i] - mu) * rstd : 0.f;
        const float g = d[h][i] * gm[i];
        s1 += g; s2 += g * xh;
        pg[i] += d[h][i] * xh; pb[i] += d[h][i];
        v[h][i] = xh; d[h][i] = g;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
      const float m1 = s1 * inv_dim, m2 = s2 * inv_dim;
      float* out = dx + static_cast<size_t>(row) * dim;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int c = lane + 32 * i;
        if (c < dim) {
          const float r = a[h][i] + rstd * (d[h][i] - m1 - v[h][i] * m2);
          out[c] = r;
          if (kEmit) {
            const size_t idx = static_cast<size_t>(row) * dim + c;
            float f = inv_keep;
            if (em.ctr != nullptr) f = keep_factor(splitmix64(key + (idx >> 2)), static_cast<int>(idx & 3), em.keep16, inv_keep);
            emit_store<TE>(em.out, idx, f * r);
          }
        }
      }
    }
  }
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    if (which) __syncthreads();
#pragma unroll
    for (int i = 0; i < NPL; ++i) red[warp][lane + 32 * i] = which ? pb[i] : pg[i];
    __syncthreads();
    for (int col = threadIdx.x; col < 32 * NPL; col += 256) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][col];
      if (col < dim) partial[(static_cast<size_t>(blockIdx.x) * 2 + which) * dim + col] = s;
    }
  }
}

// out[j][c] = sum over the n_partial partial rows (j = 0 .. n_out-1 stacked outputs of width dim).  Block = 32 columns x 32 row
// lanes: lane ty adds the contiguous chunk ty of the partials in order, the 32 chunk sums are added in lane order: fixed order,
// bit-reproducible, and a 32x shorter dependent chain than one thread per column.
__global__ void __launch_bounds__(1024) partial_reduce_kernel(const float* __restrict__ partial, int n_partial, int n_out, int dim,
                                                              float* __restrict__ out0, float* __restrict__ out1) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx, n = n_out * dim;
  const int per = (n_partial + 31) / 32, p0 = ty * per, p1 = min(n_partial, p0 + per);
  float s = 0.f;
  if (i < n) {
    const int j = i / dim, c = i - j * dim;
#pragma unroll 8
    for (int p = p0; p < p1; ++p) s += partial[(static_cast<size_t>(p) * n_out + j) * dim + c];
  }
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) t += sm[q][tx];
    const int j = i / dim, c = i - j * dim;
    (j == 0 ? out0 : out1)[c] = t;
  }
}

size_t layernorm_bwd_work_bytes(int dim) { return align_up(static_cast<size_t>(kBwdCtas) * 2 * dim * sizeof(float), 256); }

template <typename TE>
static void launch_ln_bwd_t(int ctas, const float* x, const float* dy, int rows, int dim, const float* gamma, float eps, float* dx, int accumulate,
                            float* work, const LnEmit& em, cudaStream_t stream) {
  if (dim <= 128) (void)launch_dep(layernorm_bwd_kernel<4, TE>, dim3(ctas), dim3(256), 0, stream, x, dy, rows, dim, gamma, eps, dx, accumulate, work, em);
  else if (dim <= 256) (void)launch_dep(layernorm_bwd_kernel<8, TE>, dim3(ctas), dim3(256), 0, stream, x, dy, rows, dim, gamma, eps, dx, accumulate, work, em);
  else if (dim <= 512) (void)launch_dep(layernorm_bwd_kernel<16, TE>, dim3(ctas), dim3(256), 0, stream, x, dy, rows, dim, gamma, eps, dx, accumulate, work, em);
  else (void)launch_dep(layernorm_bwd_kernel<32, TE>, dim3(ctas), dim3(256), 0, stream, x, dy, rows, dim, gamma, eps, dx, accumulate, work, em);
}

int launch_layernorm_bwd(const float* x, const float* dy, int rows, int dim, const float* gamma, float eps, float* dx, int accumulate,
                         float* dgamma, float* dbeta, float* work, cudaStream_t stream, int emit_precision, void* emit_out, float emit_scale,
                         const unsigned long long* drop_ctr, float drop_p, unsigned drop_site) {
  EC_REQUIRE(rows > 0 && dim > 0 && dim <= 1024, "LayerNorm backward: bad shape (dim <= 1024)");
  EC_REQUIRE(x && dy && gamma && dx && dgamma && dbeta && work, "null argument");
  EC_REQUIRE(drop_ctr == nullptr || (drop_p >= 0.f && drop_p < 1.f && dim % 4 == 0), "bad dropout arguments (dim must be a multiple of 4)");
  // EFFCONF_LNBWD_CTAS (<= kBwdCtas): grid-size study of the row-strided LayerNorm backward (fewer CTAs = fewer partials to reduce)
  static const int cta_cap = [] { const char* e = getenv("EFFCONF_LNBWD_CTAS"); const int v = e ? atoi(e) : 0; return v > 0 ? std::min(v, kBwdCtas) : kBwdCtas; }();
  const int ctas = std::min(cta_cap, cdiv(rows, 8));
  LnEmit em{emit_out, emit_scale, drop_ctr, drop_site, drop_ctr != nullptr ? keep16_of(drop_p) : 65536u};
  if (emit_out == nullptr) launch_ln_bwd_t<void>(ctas, x, dy, rows, dim, gamma, eps, dx, accumulate, work, em, stream);
  else if (emit_precision == EC_PREC_TF32) launch_ln_bwd_t<float>(ctas, x, dy, rows, dim, gamma, eps, dx, accumulate, work, em, stream);
  else if (emit_precision == EC_PREC_BF16) launch_ln_bwd_t<__nv_bfloat16>(ctas, x, dy, rows, dim, gamma, eps, dx, accumulate, work, em, stream);
  else if (emit_precision == EC_PREC_BF16X2) launch_ln_bwd_t<SplitBf16>(ctas, x, dy, rows, dim, gamma, eps, dx, accumulate, work, em, stream);
  else EC_FAIL("unknown precision");
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(partial_reduce_kernel, dim3(cdiv(2 * dim, 32)), dim3(1024), 0, stream, work, ctas, 2, dim, dgamma, dbeta);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Column sums of a [rows, cols] matrix (fp32 or activation type): the bias gradient of a Linear / pointwise conv.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ m, int rows, int cols, float* __restrict__ partial) {
  grid_dependency_wait();
  grid_launch_dependents();
  // thread = column (blockIdx.y tiles the columns by 256), CTA blockIdx.x walks rows blockIdx.x, + gridDim.x, ...: coalesced rows
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) s += ActTraits<T>::from(m[static_cast<size_t>(r) * cols + c]);
  partial[static_cast<size_t>(blockIdx.x) * cols + c] = s;
}


size_t colsum_work_bytes(int cols) { return align_up(static_cast<size_t>(kBwdCtas) * cols * sizeof(float), 256); }

int launch_colsum(int precision, const void* m, int is_f32, int rows, int cols, float* out, float* work, cudaStream_t stream) {
  EC_REQUIRE(rows > 0 && cols > 0 && m && out && work, "column sum: bad arguments");
  const int ctas = std::min(kBwdCtas, rows);
  dim3 grid(ctas, cdiv(cols, 256));
  if (is_f32) precision = EC_PREC_TF32;
  EC_DISPATCH_PREC(precision, ((void)launch_dep(colsum_kernel<ActT>, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const ActT*>(m), rows, cols, work)));
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(partial_reduce_kernel, dim3(cdiv(cols, 32)), dim3(1024), 0, stream, work, ctas, 1, cols, out, out);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// dst[c][r] = to_act(src[r][c]): operand-type transposed copy of an fp32 matrix (W^T for the data-gradient GEMMs).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) transpose_cast_kernel(const float* __restrict__ src, int rows, int cols, T* __restrict__ dst) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
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
      if constexpr (IsSplit<T>::value)      // swapped plane of the [2, cols, rows] weight operand
        dst[static_cast<size_t>(cols) * rows + static_cast<size_t>(c) * rows + r] = SplitBf16{split_swap(v.bits)};
    }
  }
}

int launch_transpose_cast(int precision, const float* src, int rows, int cols, void* dst, cudaStream_t stream) {
  EC_REQUIRE(rows > 0 && cols > 0 && src && dst, "transpose: bad arguments");
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32));
  EC_DISPATCH_PREC(precision, ((void)launch_dep(transpose_cast_kernel<ActT>, dim3(grid), dim3(256), 0, stream, src, rows, cols, reinterpret_cast<ActT*>(dst))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Swish backward:  y = z * sigmoid(z)  ->  dz = dy * (s + z * s * (1 - s)),  s = sigmoid(z).       (z = pre-activation, saved)
// GLU backward:    y = a * sigmoid(g)  ->  da = dy * s,  dg = dy * a * s * (1 - s);  zg = [a | g] rows of width 2C, out [da | dg].
// Inputs fp32 or activation type for z / zg; dy fp32; outputs in the activation type (operands of the next gradient GEMM).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) swish_bwd_kernel(const T* __restrict__ z, const float* __restrict__ dy, size_t n, T* __restrict__ dz) {
  grid_dependency_wait();
  grid_launch_dependents();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float zz = ActTraits<T>::from(z[i]);
    const float s = 1.f / (1.f + __expf(-zz));
    dz[i] = ActTraits<T>::to(dy[i] * (s + zz * s * (1.f - s)));
  }
}
template <typename T>
__global__ void __launch_bounds__(256) glu_bwd_kernel(const T* __restrict__ zg, const float* __restrict__ dy, size_t rows, int C,
                                                      T* __restrict__ dzg) {
  grid_dependency_wait();
  grid_launch_dependents();
  const size_t n = rows * C;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const size_t r = i / C; const int c = static_cast<int>(i - r * C);
    const float a = ActTraits<T>::from(zg[r * 2 * C + c]), g = ActTraits<T>::from(zg[r * 2 * C + C + c]);
    const float s = 1.f / (1.f + __expf(-g));
    const float d = dy[i];
    dzg[r * 2 * C + c] = ActTraits<T>::to(d * s);
    dzg[r * 2 * C + C + c] = ActTraits<T>::to(d * a * s * (1.f - s));
  }
}

int launch_swish_bwd(int precision, const void* z, const float* dy, size_t n, void* dz, cudaStream_t stream) {
  EC_REQUIRE(z && dy && dz && n > 0, "swish backward: bad arguments");
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
  EC_DISPATCH_PREC(precision, ((void)launch_dep(swish_bwd_kernel<ActT>, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const ActT*>(z), dy, n, reinterpret_cast<ActT*>(dz))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
int launch_glu_bwd(int precision, const void* zg, const float* dy, size_t rows, int C, void* dzg, cudaStream_t stream) {
  EC_REQUIRE(zg && dy && dzg && rows > 0 && C > 0, "GLU backward: bad arguments");
  const int blocks = static_cast<int>(std::min<size_t>((rows * C + 255) / 256, 148 * 16));
  EC_DISPATCH_PREC(precision, ((void)launch_dep(glu_bwd_kernel<ActT>, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const ActT*>(zg), dy, rows, C, reinterpret_cast<ActT*>(dzg))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec
