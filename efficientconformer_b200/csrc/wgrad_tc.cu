// Weight gradient of a Linear / pointwise conv on the tcgen05 tensor cores:   dW[N, K] = dY[M, N]^T . X[M, K]
// (what autograd computes for the weights of reference models/layers.py:67 F.linear and :136 pointwise F.conv1d).
//
// The contraction runs over the M = B*T rows, which is the slow dimension of both row-major operands, so both tiles are
// MN-MAJOR operands of the MMA: a TMA box of 64 rows (m) x 128 bytes (64 bf16 / 32 tf32 columns) with the 128-byte swizzle
// lands in shared memory exactly as the canonical MN-major SWIZZLE_128B layout (8-row groups 1024 B apart = stride byte offset;
// the next 128-byte column block 8192 B further = leading byte offset), and the a_major / b_major bits of the instruction
// descriptor tell the tensor core to read it transposed -- no transposed copies of the activations are ever made.
//   D[n (128 TMEM lanes), k (<= 256 columns)] += sum over a 64-row stage of dY[m, n] * X[m, k]
// Split-M: grid.z CTAs reduce disjoint row ranges into fp32 partial tiles; wgrad_reduce_kernel adds them in a fixed order
// (bit-reproducible, optional accumulation into an existing gradient).  bf16 operands; the TF32 parity mode splits its fp32
// operands into bf16 hi + lo parts and runs three passes (launch_wgrad).  Warp 0: TMA producer, warp 1: TMEM + MMA issue,
// warps 2-5: epilogue (thread = weight row n).
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

struct WgDev {
  int M, N, K;
  int bn;                // UMMA N = weight columns per CTA (multiple of the box width, <= 256)
  int a_boxes, b_boxes;  // 128-byte column blocks of the dY / X tiles
  int rows_per_split, stages, tmem_cols;
  int epi;               // 0: thread-per-row stores (measured default); 1: experimental coalesced epilogue through shared memory
  float* partial;        // [splits][N][K]
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

template <typename T>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgDev p) {
  using Tr = ActTraits<T>;
  constexpr int EB = 128 / sizeof(T);              // elements per 128-byte box row
  constexpr int KSTEP_ROWS = Tr::kUmmaK;           // m rows consumed by one MMA (16 bf16 / 8 tf32)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  const int stage_bytes = (p.a_boxes + p.b_boxes) * kWgBoxBytes;
  const uint32_t bars = base + p.stages * stage_bytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kWgMaxStages + s); };
  const uint32_t acc_full = bars + 8u * (2 * kWgMaxStages);
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(bp + p.stages * stage_bytes + 8 * (2 * kWgMaxStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128, k0 = blockIdx.y * p.bn;
  const int m_begin = blockIdx.z * p.rows_per_split, m_end = min(p.M, m_begin + p.rows_per_split);
  const int n_stage = (m_end - m_begin + kWgRows - 1) / kWgRows;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY); tma_prefetch_desc(&tmX);
    for (int s = 0; s < kWgMaxStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), p.tmem_cols);
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
        const int m = m_begin + i * kWgRows;           // rows beyond M (or this split's end: harmless, see below) are zero-filled
        for (int bx = 0; bx < p.a_boxes; ++bx) tma_load_2d(dst + bx * kWgBoxBytes, &tmDY, full(s), n0 + bx * EB, m);
        for (int bx = 0; bx < p.b_boxes; ++bx) tma_load_2d(dst + (p.a_boxes + bx) * kWgBoxBytes, &tmX, full(s), k0 + bx * EB, m);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(Tr::kTf32 ? 2u : 1u, 128, p.bn) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
    for (int i = 0; i < n_stage; ++i) {
      const int s = i % p.stages;
      mbar_wait(full(s), (i / p.stages) & 1);
      tc_fence_after();
      const uint32_t a_addr = base + s * stage_bytes, b_addr = a_addr + p.a_boxes * kWgBoxBytes;
      // a stage whose rows run past this split's range would double count rows of the next split: only whole k-steps inside
      // [m_begin, m_end) are issued (rows_per_split is a multiple of the stage, so only the global tail M is ever partial, and
      // that tail is zero-filled by TMA)
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kWgRows / KSTEP_ROWS; ++ks) {
          const uint64_t da = make_smem_desc_mn_sw128(a_addr + ks * KSTEP_ROWS * 128, kWgBoxBytes);
          const uint64_t db = make_smem_desc_mn_sw128(b_addr + ks * KSTEP_ROWS * 128, kWgBoxBytes);
          tc_mma<Tr::kTf32>(tmem_base, da, db, idesc, (i | ks) != 0 ? 1u : 0u);
        }
        tc_commit(empty(s));
        if (i == n_stage - 1) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue: thread = weight row n0 + q*32 + lane; 32-column chunks straight to the fp32 partial tile ----
    const int q = warp & 3;                             // warps 2..5 -> TMEM lane quarters 2, 3, 0, 1
    const int n = n0 + q * 32 + lane;
    float* out = p.partial + (static_cast<size_t>(blockIdx.z) * p.N + n) * p.K;
    if (n_stage > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    for (int c0 = 0; c0 < p.bn; c0 += 32) {
      uint32_t v[32];
      if (n_stage > 0) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (p.epi == 0) {
        if (n < p.N) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int k = k0 + c0 + j;
            if (k < p.K) out[k] = __uint_as_float(v[j]);
          }
        }
      } else {
        // EXPERIMENTAL (EFFCONF_WGRAD_EPI=1, not yet measured): the 32 x 32 chunk of this warp is transposed through shared memory
        // (the drained operand ring: every MMA has completed once acc_full fired and the producer issues no further loads), so that
        // each store instruction writes 128 contiguous bytes of ONE weight row instead of 4 bytes of 32 different rows.
        float* tile = reinterpret_cast<float*>(bp) + (warp - 2) * (32 * 33);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        const int k = k0 + c0 + lane;
        for (int r = 0; r < 32; ++r) {
          const int nr = n0 + q * 32 + r;
          if (nr < p.N && k < p.K) p.partial[(static_cast<size_t>(blockIdx.z) * p.N + nr) * p.K + k] = tile[r * 33 + lane];
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// dW[i] (+)= sum over splits of partial[s][i], in split order
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, size_t n, float* __restrict__ dw, int accumulate) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = accumulate ? dw[i] : 0.f;
  for (int z = 0; z < splits; ++z) s += partial[static_cast<size_t>(z) * n + i];
  dw[i] = s;
}

static int wgrad_splits(int M, int N, int K, int bn) {
  const int tiles = cdiv(N, 128) * cdiv(K, bn);
  const int want = std::max(1, 148 / tiles);
  return std::max(1, std::min(want, cdiv(M, 4 * kWgRows)));
}
static int wgrad_bn(int precision, int K) {
  const int eb = precision == EC_PREC_TF32 ? 32 : 64;
  return std::min(256, round_up(K, eb));
}
static size_t wgrad_partial_bytes(int M, int N, int K) {
  const int bn = wgrad_bn(EC_PREC_BF16, K);
  return align_up(static_cast<size_t>(wgrad_splits(M, N, K, bn)) * N * K * sizeof(float), 256);
}
size_t wgrad_work_bytes(int precision, int M, int N, int K) {
  size_t b = wgrad_partial_bytes(M, N, K);
  if (precision == EC_PREC_TF32 || precision == EC_PREC_BF16X2)   // bf16 hi / lo copies of both operands (see launch_wgrad)
    b += 2 * (align_up(static_cast<size_t>(M) * N * 2, 256) + align_up(static_cast<size_t>(M) * K * 2, 256));
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

// split mode: the packed (hi, lo) pairs are taken apart into the same two bf16 planes
__global__ void __launch_bounds__(256) unpack_bf16_kernel(const uint32_t* __restrict__ src, size_t n, __nv_bfloat16* __restrict__ hi,
                                                          __nv_bfloat16* __restrict__ lo) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t v = src[i];
    hi[i] = __ushort_as_bfloat16(static_cast<unsigned short>(v & 0xffffu));
    lo[i] = __ushort_as_bfloat16(static_cast<unsigned short>(v >> 16));
  }
}

template <typename T>
static int launch_wgrad_t(int precision, const void* dy, const void* x, int M, int N, int K, float* dw, int accumulate, float* work,
                          cudaStream_t stream) {
  constexpr int EB = 128 / sizeof(T);
  WgDev p{};
  p.M = M; p.N = N; p.K = K;
  p.bn = wgrad_bn(precision, K);
  p.a_boxes = 128 / EB; p.b_boxes = p.bn / EB;
  const int splits = wgrad_splits(M, N, K, p.bn);
  p.rows_per_split = round_up(cdiv(M, splits), kWgRows);
  const int stage_bytes = (p.a_boxes + p.b_boxes) * kWgBoxBytes;
  p.stages = std::max(2, std::min(kWgMaxStages, (200 * 1024) / stage_bytes));
  int cols = 32;
  while (cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  p.partial = work;
  static const int epi_mode = [] { const char* e = getenv("EFFCONF_WGRAD_EPI"); return (e && e[0] == '1') ? 1 : 0; }();
  p.epi = epi_mode;
  CUtensorMap tmDY, tmX;
  const bool f32 = precision == EC_PREC_TF32;
  EC_TRY(make_map(&tmDY, f32, dy, M, N, N, EB, kWgRows, CU_TENSOR_MAP_SWIZZLE_128B));
  EC_TRY(make_map(&tmX, f32, x, M, K, K, EB, kWgRows, CU_TENSOR_MAP_SWIZZLE_128B));
  const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + 8 * (2 * kWgMaxStages + 1) + 16 + 1024;
  static cudaError_t attr = cudaFuncSetAttribute(wgrad_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  EC_CUDA(attr);
  dim3 grid(cdiv(N, 128), cdiv(K, p.bn), splits);
  EC_TRY(launch_pdl(wgrad_tc_kernel<T>, grid, dim3(kWgThreads), smem, stream, tmDY, tmX, p));
  const size_t n = static_cast<size_t>(N) * K;
  wgrad_reduce_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, stream>>>(work, splits, n, dw, accumulate);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

int launch_wgrad(int precision, const void* dy, const void* x, int M, int N, int K, float* dw, int accumulate, float* work,
                 cudaStream_t stream) {
  EC_REQUIRE(M > 0 && N > 0 && K > 0 && dy && x && dw && work, "wgrad: bad arguments");
  if (precision == EC_PREC_BF16) return launch_wgrad_t<__nv_bfloat16>(precision, dy, x, M, N, K, dw, accumulate, work, stream);
  if (precision == EC_PREC_TF32 || precision == EC_PREC_BF16X2) {
    // Parity / split modes.  The tensor core reads MN-major operands only for 16-bit types (kind::tf32 with a transposed operand produces
    // nothing -- measured), so the fp32 operands are split into bf16 hi + lo parts and the product is assembled from three bf16
    // passes, dY_hi^T X_hi + dY_hi^T X_lo + dY_lo^T X_hi: 16 significant operand bits, more than the 11 of TF32.
    uint8_t* wp = reinterpret_cast<uint8_t*>(work) + wgrad_partial_bytes(M, N, K);
    __nv_bfloat16* dy_hi = reinterpret_cast<__nv_bfloat16*>(wp); wp += align_up(static_cast<size_t>(M) * N * 2, 256);
    __nv_bfloat16* dy_lo = reinterpret_cast<__nv_bfloat16*>(wp); wp += align_up(static_cast<size_t>(M) * N * 2, 256);
    __nv_bfloat16* x_hi = reinterpret_cast<__nv_bfloat16*>(wp); wp += align_up(static_cast<size_t>(M) * K * 2, 256);
    __nv_bfloat16* x_lo = reinterpret_cast<__nv_bfloat16*>(wp);
    const size_t n1 = static_cast<size_t>(M) * N, n2 = static_cast<size_t>(M) * K;
    const int g1 = static_cast<int>(std::min<size_t>((n1 + 255) / 256, 148 * 16)), g2 = static_cast<int>(std::min<size_t>((n2 + 255) / 256, 148 * 16));
    if (precision == EC_PREC_TF32) {
      split_bf16_kernel<<<g1, 256, 0, stream>>>(reinterpret_cast<const float*>(dy), n1, dy_hi, dy_lo);
      split_bf16_kernel<<<g2, 256, 0, stream>>>(reinterpret_cast<const float*>(x), n2, x_hi, x_lo);
    } else {
      unpack_bf16_kernel<<<g1, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(dy), n1, dy_hi, dy_lo);
      unpack_bf16_kernel<<<g2, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(x), n2, x_hi, x_lo);
    }
    EC_CUDA(cudaGetLastError());
    EC_TRY(launch_wgrad_t<__nv_bfloat16>(EC_PREC_BF16, dy_lo, x_hi, M, N, K, dw, accumulate, work, stream));   // small terms first
    EC_TRY(launch_wgrad_t<__nv_bfloat16>(EC_PREC_BF16, dy_hi, x_lo, M, N, K, dw, 1, work, stream));
    return launch_wgrad_t<__nv_bfloat16>(EC_PREC_BF16, dy_hi, x_hi, M, N, K, dw, 1, work, stream);
  }
  EC_FAIL("unknown precision");
}

}  // namespace ec
