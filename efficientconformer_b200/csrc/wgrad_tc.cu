// Weight (and bias) gradient of a Linear / pointwise conv on the tcgen05 tensor cores:
//     dW[N, K] = dY[M, N]^T . X[M, K],     db[N] = sum_m dY[m, n]
// (what autograd computes for reference models/layers.py:67 F.linear and :136 pointwise F.conv1d).
//
// The contraction runs over the M = B*T rows, which is the slow dimension of both row-major operands, so both tiles are
// MN-MAJOR operands of the MMA: a TMA box of 64 rows (m) x 128 bytes (64 bf16 columns) with the 128-byte swizzle lands in shared
// memory exactly as the canonical MN-major SWIZZLE_128B layout (8-row groups 1024 B apart = stride byte offset; the next 128-byte
// column block 8192 B further = leading byte offset), and the a_major / b_major bits of the instruction descriptor tell the tensor
// core to read it transposed -- no transposed copies of the activations are ever made.
//   D[n (128 TMEM lanes), k (<= 256 columns)] += sum over a 64-row stage of dY[m, n] * X[m, k]
// Bias gradient: a constant tile of ones is a second B operand (UMMA N = 16) accumulated into 16 spare TMEM columns by the CTAs of
// the first k tile -- the column sums of dY come out of the tensor core for one extra small MMA per k-step, no separate pass over dY.
// Split mode (EC_PREC_BF16X2): the packed (hi, lo) operands are read as bf16 matrices of twice the width, so the accumulator holds the
// four partial products of every weight in a 2 x 2 block (rows 2n, 2n+1 x columns 2k, 2k+1); the epilogue adds the block (column pairs
// in registers, row pairs by one shuffle): ONE pass gives the 16-bit-operand product.
// Epilogue: the accumulator chunk of a warp is staged through the drained operand ring so that every global store writes whole rows
// (512 contiguous bytes per instruction) instead of 4 bytes to 32 different rows.
// Split-M: grid.z CTAs reduce disjoint row ranges into fp32 partial tiles; wgrad_reduce_kernel adds them in a fixed order
// (bit-reproducible, optional accumulation into an existing gradient).  The TF32 parity mode splits its fp32 operands into bf16
// hi + lo planes and runs three bf16 passes (launch_wgrad).  Warp 0: TMA producer, warp 1: TMEM + MMA issue, warps 2-5: epilogue.
#include "ec_common.cuh"
#include "ec_tma.cuh"
#include <algorithm>
#include <cstdlib>

namespace ec {

namespace {
constexpr int kWgThreads = 192;
constexpr int kWgRows = 64;               // m rows per pipeline stage
constexpr int kWgBoxBytes = kWgRows * 128;
constexpr int kWgMaxStages = 6;
constexpr int kWgEB = 64;                 // bf16 elements per 128-byte box row

struct WgDev {
  int M, N, K;           // logical dims of the gradient
  int bn;                // UMMA N = bf16 view columns of X per CTA (multiple of 64, <= 256)
  int b_boxes;           // 128-byte column blocks of the X tile (the dY tile has two)
  int rows_per_split, stages, tmem_cols;
  int bias_col;          // TMEM column of the bias accumulator, or -1
  float* partial;        // [splits][N][K] then [splits][N] (bias)
  int splits;
};

// MN-major SWIZZLE_128B operand: [16,30) leading byte offset >> 4 = distance between 128-byte column blocks,
// [32,46) stride byte offset >> 4 = distance between 8-row groups along the contraction dim.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
}  // namespace

template <bool kSplit>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgDev p) {
  constexpr int KSTEP_ROWS = 16;                   // m rows consumed by one bf16 MMA
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  const int stage_bytes = (2 + p.b_boxes) * kWgBoxBytes;
  const uint32_t ones_addr = base + p.stages * stage_bytes;             // 8 KB tile of bf16 ones (bias gradient), 1024-byte aligned
  const uint32_t bars = ones_addr + kWgBoxBytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kWgMaxStages + s); };
  const uint32_t acc_full = bars + 8u * (2 * kWgMaxStages);
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(bp + p.stages * stage_bytes + kWgBoxBytes + 8 * (2 * kWgMaxStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128, k0 = blockIdx.y * p.bn;             // bf16 view coordinates
  const int m_begin = blockIdx.z * p.rows_per_split, m_end = min(p.M, m_begin + p.rows_per_split);
  const int n_stage = (m_end - m_begin + kWgRows - 1) / kWgRows;
  const bool do_bias = p.bias_col >= 0 && blockIdx.y == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY); tma_prefetch_desc(&tmX);
    for (int s = 0; s < kWgMaxStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), p.tmem_cols);
  if (warp >= 2 && do_bias) {                       // every entry is 1.0, so the swizzle permutation is irrelevant
    uint32_t* ones = reinterpret_cast<uint32_t*>(bp + p.stages * stage_bytes);
    for (int i = threadIdx.x - 64; i < kWgBoxBytes / 4; i += kWgThreads - 64) ones[i] = 0x3F803F80u;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  grid_dependency_wait();
  grid_launch_dependents();

  if (warp == 0) {
    for (int i = 0; i < n_stage; ++i) {
      const int s = i % p.stages;
      mbar_wait(empty(s), ((i / p.stages) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(full(s), stage_bytes);
        const uint32_t dst = base + s * stage_bytes;
        const int m = m_begin + i * kWgRows;           // rows beyond M are zero-filled by TMA
        for (int bx = 0; bx < 2; ++bx) tma_load_2d(dst + bx * kWgBoxBytes, &tmDY, full(s), n0 + bx * kWgEB, m);
        for (int bx = 0; bx < p.b_boxes; ++bx) tma_load_2d(dst + (2 + bx) * kWgBoxBytes, &tmX, full(s), k0 + bx * kWgEB, m);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(1u, 128, p.bn) | (1u << 15) | (1u << 16);      // a_major = b_major = MN
    const uint32_t idesc_bias = make_idesc(1u, 128, 16) | (1u << 15) | (1u << 16);
    for (int i = 0; i < n_stage; ++i) {
      const int s = i % p.stages;
      mbar_wait(full(s), (i / p.stages) & 1);
      tc_fence_after();
      const uint32_t a_addr = base + s * stage_bytes, b_addr = a_addr + 2 * kWgBoxBytes;
      // rows_per_split is a multiple of the stage, so only the global tail M is ever partial, and that tail is zero-filled by TMA
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kWgRows / KSTEP_ROWS; ++ks) {
          const uint64_t da = make_smem_desc_mn_sw128(a_addr + ks * KSTEP_ROWS * 128, kWgBoxBytes);
          const uint64_t db = make_smem_desc_mn_sw128(b_addr + ks * KSTEP_ROWS * 128, kWgBoxBytes);
          tc_mma<false>(tmem_base, da, db, idesc, (i | ks) != 0 ? 1u : 0u);
          if (do_bias)
            tc_mma<false>(tmem_base + p.bias_col, da, make_smem_desc_mn_sw128(ones_addr + ks * KSTEP_ROWS * 128, kWgBoxBytes), idesc_bias,
                          (i | ks) != 0 ? 1u : 0u);
        }
        tc_commit(empty(s));
        if (i == n_stage - 1) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue: TMEM lane = view row n0 + q*32 + lane.  Chunks of 32 view columns are staged in shared memory (the drained
    //      operand ring: every MMA has completed once acc_full fired and the producer issues no further loads), then whole logical
    //      rows go out with 16-byte stores (a warp instruction writes 512 contiguous bytes of one weight row) ----
    const int q = warp & 3;                             // warps 2..5 -> TMEM lane quarters 2, 3, 0, 1
    constexpr int kRowsL = kSplit ? 16 : 32;            // logical weight rows of this warp
    const int cols_l = kSplit ? p.bn / 2 : p.bn;        // logical weight columns of this CTA
    const int ld = cols_l + 4;
    float* tile = reinterpret_cast<float*>(bp) + (warp - 2) * kRowsL * ld;
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    if (n_stage > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    for (int c0 = 0; c0 < p.bn; c0 += 32) {
      uint32_t v[32];
      if (n_stage > 0) {
        tmem_ld_32x32(tq + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if constexpr (kSplit) {
        float sacc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = __uint_as_float(v[2 * j]) + __uint_as_float(v[2 * j + 1]);      // (n', 2k) + (n', 2k+1)
          sacc[j] = t + __shfl_xor_sync(0xffffffffu, t, 1);                                 // + row n' ^ 1
        }
        if ((lane & 1) == 0) {
          float* dst = tile + (lane >> 1) * ld + (c0 >> 1);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            *reinterpret_cast<float4*>(dst + 4 * j4) = make_float4(sacc[4 * j4], sacc[4 * j4 + 1], sacc[4 * j4 + 2], sacc[4 * j4 + 3]);
        }
      } else {
        float* dst = tile + lane * ld + c0;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<float4*>(dst + 4 * j4) = make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                                                                 __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3]));
      }
    }
    __syncwarp();
    const int nl0 = (n0 + q * 32) / (kSplit ? 2 : 1), kl0 = k0 / (kSplit ? 2 : 1);
    float* out = p.partial + static_cast<size_t>(blockIdx.z) * p.N * p.K;
    const bool vec = (p.K & 3) == 0;
    for (int r = 0; r < kRowsL; ++r) {
      const int n = nl0 + r;
      if (n >= p.N) break;
      float* orow = out + static_cast<size_t>(n) * p.K + kl0;
      for (int c = 4 * lane; c < cols_l; c += 128) {
        const float4 t4 = *reinterpret_cast<const float4*>(tile + r * ld + c);
        if (vec && kl0 + c + 3 < p.K) {
          *reinterpret_cast<float4*>(orow + c) = t4;
        } else {
          if (kl0 + c < p.K) orow[c] = t4.x;
          if (kl0 + c + 1 < p.K) orow[c + 1] = t4.y;
          if (kl0 + c + 2 < p.K) orow[c + 2] = t4.z;
          if (kl0 + c + 3 < p.K) orow[c + 3] = t4.w;
        }
      }
    }
    if (do_bias) {                                      // all 16 ones-columns hold the same column sum of dY
      uint32_t b16[16];
      float t = 0.f;
      if (n_stage > 0) {
        tmem_ld_32x16(tq + p.bias_col, b16);
        tmem_ld_wait();
        t = __uint_as_float(b16[0]);
      }
      float* pb = p.partial + static_cast<size_t>(p.splits) * p.N * p.K + static_cast<size_t>(blockIdx.z) * p.N;
      if constexpr (kSplit) {
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        const int n = nl0 + (lane >> 1);
        if ((lane & 1) == 0 && n < p.N) pb[n] = t;
      } else {
        const int n = nl0 + lane;
        if (n < p.N) pb[n] = t;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// dW[i] (+)= sum over splits of partial[s][i], in split order; the bias sums follow the weight partials
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, size_t n_w, int n_b, float* __restrict__ dw, int accumulate,
                                    float* __restrict__ db, int accumulate_b) {
  grid_dependency_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n_w) {
    float s = accumulate ? dw[i] : 0.f;
    for (int z = 0; z < splits; ++z) s += partial[static_cast<size_t>(z) * n_w + i];
    dw[i] = s;
  } else if (db != nullptr && i < n_w + n_b) {
    const size_t j = i - n_w;
    const float* pb = partial + static_cast<size_t>(splits) * n_w;
    float s = accumulate_b ? db[j] : 0.f;
    for (int z = 0; z < splits; ++z) s += pb[static_cast<size_t>(z) * n_b + j];
    db[j] = s;
  }
}

static int wgrad_bn(int Kv) { return std::min(256, round_up(Kv, kWgEB)); }
static int wgrad_splits(int M, int Nv, int Kv) {
  const int tiles = cdiv(Nv, 128) * cdiv(Kv, wgrad_bn(Kv));
  const int want = std::max(1, 148 / tiles);
  return std::max(1, std::min(want, cdiv(M, 4 * kWgRows)));
}
static size_t wgrad_partial_bytes(int precision, int M, int N, int K) {
  const int f = precision == EC_PREC_BF16X2 ? 2 : 1;
  return align_up(static_cast<size_t>(wgrad_splits(M, f * N, f * K)) * (static_cast<size_t>(N) * K + N) * sizeof(float), 256);
}
size_t wgrad_work_bytes(int precision, int M, int N, int K) {
  size_t b = wgrad_partial_bytes(precision, M, N, K);
  if (precision == EC_PREC_TF32)   // bf16 hi / lo copies of both operands + the column-sum partials of the bias gradient (see launch_wgrad)
    b += 2 * (align_up(static_cast<size_t>(M) * N * 2, 256) + align_up(static_cast<size_t>(M) * K * 2, 256)) + colsum_work_bytes(N);
  return b;
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 significant bits in two bf16 tensors
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ src, size_t n, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// One pass: dy / x are bf16 matrices [M, Nv] / [M, Kv] (Nv = 2N, Kv = 2K packed views when split).
template <bool kSplit>
static int launch_wgrad_pass(const void* dy, const void* x, int M, int N, int K, float* dw, int accumulate, float* db, int accumulate_b,
                             float* work, cudaStream_t stream) {
  const int f = kSplit ? 2 : 1, Nv = f * N, Kv = f * K;
  WgDev p{};
  p.M = M; p.N = N; p.K = K;
  p.bn = wgrad_bn(Kv);
  p.b_boxes = p.bn / kWgEB;
  const int splits = wgrad_splits(M, Nv, Kv);
  p.splits = splits;
  p.rows_per_split = round_up(cdiv(M, splits), kWgRows);
  const int stage_bytes = (2 + p.b_boxes) * kWgBoxBytes;
  p.stages = std::max(2, std::min(kWgMaxStages, (192 * 1024) / stage_bytes));
  p.bias_col = db != nullptr ? round_up(p.bn, 32) : -1;
  int cols = 32;
  while (cols < p.bn + (db != nullptr ? 32 : 0)) cols <<= 1;
  p.tmem_cols = cols;
  p.partial = work;
  // epilogue staging (4 warps x logical rows x (logical columns + 4) floats) aliases the operand ring
  const size_t staging = static_cast<size_t>(4) * (kSplit ? 16 : 32) * ((kSplit ? p.bn / 2 : p.bn) + 4) * sizeof(float);
  EC_REQUIRE(staging <= static_cast<size_t>(p.stages) * stage_bytes, "wgrad: epilogue staging does not fit in the operand ring");
  CUtensorMap tmDY, tmX;
  EC_TRY(make_map(&tmDY, false, dy, M, Nv, Nv, kWgEB, kWgRows, CU_TENSOR_MAP_SWIZZLE_128B));
  EC_TRY(make_map(&tmX, false, x, M, Kv, Kv, kWgEB, kWgRows, CU_TENSOR_MAP_SWIZZLE_128B));
  const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + kWgBoxBytes + 8 * (2 * kWgMaxStages + 1) + 16 + 1024;
  static cudaError_t attr = cudaFuncSetAttribute(wgrad_tc_kernel<kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  EC_CUDA(attr);
  dim3 grid(cdiv(Nv, 128), cdiv(Kv, p.bn), splits);
  EC_TRY(launch_pdl(wgrad_tc_kernel<kSplit>, grid, dim3(kWgThreads), smem, stream, tmDY, tmX, p));
  const size_t n_w = static_cast<size_t>(N) * K, n = n_w + (db != nullptr ? N : 0);
  EC_TRY(launch_pdl(wgrad_reduce_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, stream, static_cast<const float*>(work), splits,
                    n_w, N, dw, accumulate, db, accumulate_b));
  return EC_OK;
}

int launch_wgrad(int precision, const void* dy, const void* x, int M, int N, int K, float* dw, int accumulate, float* db, float* work,
                 cudaStream_t stream) {
  EC_REQUIRE(M > 0 && N > 0 && K > 0 && dy && x && dw && work, "wgrad: bad arguments");
  if (precision == EC_PREC_BF16) return launch_wgrad_pass<false>(dy, x, M, N, K, dw, accumulate, db, 0, work, stream);
  if (precision == EC_PREC_BF16X2) return launch_wgrad_pass<true>(dy, x, M, N, K, dw, accumulate, db, 0, work, stream);
  if (precision == EC_PREC_TF32) {
    // Parity mode.  The tensor core reads MN-major operands only for 16-bit types (kind::tf32 with a transposed operand produces
    // nothing -- measured), so the fp32 operands are split into bf16 hi + lo planes and the product is assembled from three bf16
    // passes, dY_hi^T X_hi + dY_hi^T X_lo + dY_lo^T X_hi: 16 significant operand bits, more than the 11 of TF32.
    uint8_t* wp = reinterpret_cast<uint8_t*>(work) + wgrad_partial_bytes(precision, M, N, K);
    __nv_bfloat16* dy_hi = reinterpret_cast<__nv_bfloat16*>(wp); wp += align_up(static_cast<size_t>(M) * N * 2, 256);
    __nv_bfloat16* dy_lo = reinterpret_cast<__nv_bfloat16*>(wp); wp += align_up(static_cast<size_t>(M) * N * 2, 256);
    __nv_bfloat16* x_hi = reinterpret_cast<__nv_bfloat16*>(wp); wp += align_up(static_cast<size_t>(M) * K * 2, 256);
    __nv_bfloat16* x_lo = reinterpret_cast<__nv_bfloat16*>(wp); wp += align_up(static_cast<size_t>(M) * K * 2, 256);
    const size_t n1 = static_cast<size_t>(M) * N, n2 = static_cast<size_t>(M) * K;
    split_bf16_kernel<<<static_cast<int>(std::min<size_t>((n1 + 255) / 256, 148 * 16)), 256, 0, stream>>>(reinterpret_cast<const float*>(dy), n1, dy_hi, dy_lo);
    split_bf16_kernel<<<static_cast<int>(std::min<size_t>((n2 + 255) / 256, 148 * 16)), 256, 0, stream>>>(reinterpret_cast<const float*>(x), n2, x_hi, x_lo);
    EC_CUDA(cudaGetLastError());
    if (db != nullptr) EC_TRY(launch_colsum(precision, dy, 1, M, N, db, reinterpret_cast<float*>(wp), stream));
    EC_TRY(launch_wgrad_pass<false>(dy_lo, x_hi, M, N, K, dw, accumulate, nullptr, 0, work, stream));   // small terms first
    EC_TRY(launch_wgrad_pass<false>(dy_hi, x_lo, M, N, K, dw, 1, nullptr, 0, work, stream));
    return launch_wgrad_pass<false>(dy_hi, x_hi, M, N, K, dw, 1, nullptr, 0, work, stream);
  }
  EC_FAIL("unknown precision");
}

}  // namespace ec
