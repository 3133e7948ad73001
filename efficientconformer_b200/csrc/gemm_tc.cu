// tcgen05 GEMM with fused epilogues:
//     out = alpha * act(A @ W^T + bias) [+ residual]          (optionally GLU over column halves)
//     [+ LayerNorm(out) / LayerNorm(LayerNorm(out)) written in the activation type for the next GEMM]
//
// Replaces every dense contraction on the hot path (reference models/layers.py:67 F.linear, :136 pointwise F.conv1d):
// FFN projections, QKV / positional / output projections, pointwise convs of the convolution module, conv_res,
// the subsampling Linear and the CTC fc head -- and, through the fused epilogue, the five nn.LayerNorm(eps=1e-6) of a
// block (reference models/modules.py:386,433,511; models/blocks.py:96,135).
//
// One CTA = one 128 x BLOCK_N output tile over the whole K.  Warp roles (192 threads):
//   warp 0   TMA producer  : cp.async.bulk.tensor loads of the A (128 x 128B) and W (BLOCK_N x 128B) k-slices, 128B swizzle
//   warp 1   MMA issuer    : allocates TMEM, one thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 32B), fp32 accumulator in TMEM
//   warps 2-5 epilogue     : tcgen05.ld (thread = output row) -> padded smem transpose -> (lane = output column)
//                            bias / Swish / GLU / alpha / residual in registers with 32 independent rows in flight ->
//                            coalesced fp32 / activation-type stores; optional per-row LayerNorm statistics by a
//                            register reduce-scatter across the warp, then normalise sweeps over the CTA's own output.
// K and N tails are zero-filled by TMA out-of-bounds handling, so D, 4D, head dims etc. need no host-side padding
// (only 16-byte row pitches).  Operand type float => kind::tf32, __nv_bfloat16 => kind::f16.
#include "ec_common.cuh"
#include <mutex>

namespace ec {

struct GemmDev {
  int M, N, K;
  int block_n;       // UMMA N (multiple of 16, <= 256)
  int num_k_blocks, stages;
  int pipe_bytes;    // bytes of the operand ring (the fused-LayerNorm row tile aliases it after the mainloop)
  int tmem_cols;
  const float* bias;
  float alpha;
  int act;
  int glu_nb, glu_channels;
  const float* residual; int ld_res;
  float* out_f32; int ld_out;
  void* out_act; int ld_act;
  int round_out;     // round the fp32 output to TF32 (it feeds a TF32 mma.sync consumer)
  // fused LayerNorm (kLN instantiation only; requires a single N tile)
  int ln_mode;       // 1: y = LN1(out);  2: out <- LN1(out), y = LN2(out) (LN2 identity when ln2_g == nullptr)
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  float ln_eps;
  void* ln_out; int ld_ln;                 // y, activation type (may be null)
  void* copy_out; int copy_stride, frames_per_seq, frames_out_per_seq;   // strided compaction of `out` (LN1 mode), activation type
};

constexpr int kBlockM = 128;
constexpr int kATileBytes = kBlockM * 128;
constexpr int kStagingBytes = 4 * 32 * 33 * 4;
constexpr int kMaxStages = 8;

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

template <typename T, bool kLN>
__global__ void __launch_bounds__(192, kLN ? 1 : 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  using Tr = ActTraits<T>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stage_bytes = kATileBytes + p.block_n * 128;
  float* staging = reinterpret_cast<float*>(base_ptr + p.pipe_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + p.pipe_bytes + kStagingBytes);
  const uint32_t bars_addr = base + p.pipe_bytes + kStagingBytes;
  // bars[0..kMaxStages) full, [kMaxStages..2kMaxStages) empty, [2kMaxStages] tmem_full, then the TMEM address holder
  auto full_bar = [&](int s) { return bars_addr + 8u * s; };
  auto empty_bar = [&](int s) { return bars_addr + 8u * (kMaxStages + s); };
  const uint32_t tmem_full_bar = bars_addr + 8u * (2 * kMaxStages);
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(bars + 2 * kMaxStages + 1);

  const int m0 = blockIdx.x * kBlockM;
  const int tile_n = blockIdx.y;
  const int w_row0 = tile_n * p.block_n;   // first W row (and, for plain GEMMs, first output column) of this tile

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on we read its output.
  grid_dependency_wait();
  grid_launch_dependents();

  if (warp_idx == 0) {
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>(stage_bytes);
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), tx);
        const uint32_t a_dst = base + s * stage_bytes;
        tma_load_2d(a_dst, &tmA, full_bar(s), kb * Tr::kBlockK, m0);
        tma_load_2d(a_dst + kATileBytes, &tmB, full_bar(s), kb * Tr::kBlockK, w_row0);
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(Tr::kTf32 ? 2u : 1u, kBlockM, p.block_n);
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_src = base + s * stage_bytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 4 x 32-byte K slices inside the 128-byte swizzle row
          const uint64_t da = make_smem_desc_sw128(a_src + k * 32);
          const uint64_t db = make_smem_desc_sw128(a_src + kATileBytes + k * 32);
          tc_mma<Tr::kTf32>(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(empty_bar(s));       // frees the smem slot when these MMAs retire
      }
      tc_commit(tmem_full_bar);        // accumulator complete
    }
  } else {
    // ---------------- epilogue: warps 2..5 own TMEM lane quarters (warp_idx % 4) ----------------
    const int q = warp_idx & 3;
    float* stg = staging + q * (32 * 33);
    const bool glu = p.glu_nb > 0;
    const int cols = glu ? p.glu_nb : p.block_n;            // logical output columns of this tile
    const int out_col0 = glu ? tile_n * p.glu_nb : w_row0;
    const int n_limit = glu ? p.glu_channels : p.N;
    const int row0 = m0 + q * 32;
    const int rows_valid = min(32, p.M - row0);             // may be <= 0 for the tail tile
    T* out_act = reinterpret_cast<T*>(p.out_act);
    // kLN: the warp's 32 x N fp32 output rows are kept in shared memory (aliasing the drained pipeline stages; odd row
    // pitch => conflict-free in both the lane = column and the thread = row orientation) until the LayerNorm(s) are done.
    const int tile_pitch = p.N | 1;
    float* tile = reinterpret_cast<float*>(base_ptr) + static_cast<size_t>(q) * 32 * tile_pitch;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < cols; c0 += 32) {
      const int n = out_col0 + c0 + lane;
      const bool n_ok = (c0 + lane) < cols && n < n_limit;
      // per-column constants first (their latency hides behind the TMEM load + transpose)
      float bias_a = 0.f, bias_g = 0.f;
      if (p.bias != nullptr && n_ok) {
        if (glu) { bias_a = __ldg(p.bias + w_row0 + c0 + lane); bias_g = __ldg(p.bias + w_row0 + p.glu_nb + c0 + lane); }
        else bias_a = __ldg(p.bias + n);
      }
      float rr[32];
      const bool has_res = p.residual != nullptr;
      if (has_res) {
        const float* rp = p.residual + static_cast<size_t>(row0) * p.ld_res + n;
#pragma unroll
        for (int r = 0; r < 32; ++r) rr[r] = (n_ok && r < rows_valid) ? __ldg(rp + static_cast<size_t>(r) * p.ld_res) : 0.f;
      }
      uint32_t v[32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0);
      tmem_ld_32x32(taddr, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      float t[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) t[r] = stg[r * 33 + lane] + bias_a;
      __syncwarp();
      if (glu) {
        tmem_ld_32x32(taddr + static_cast<uint32_t>(p.glu_nb), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r) t[r] *= fast_sigmoid(stg[r * 33 + lane] + bias_g);
        __syncwarp();
      }
      if (p.act == GEMM_ACT_SWISH) {
#pragma unroll
        for (int r = 0; r < 32; ++r) t[r] *= fast_sigmoid(t[r]);
      }
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        t[r] *= p.alpha;
        if (has_res) t[r] += rr[r];
        if (p.round_out) t[r] = round_tf32(t[r]);
      }
      if constexpr (kLN) {
        if (n_ok) {
#pragma unroll
          for (int r = 0; r < 32; ++r) tile[r * tile_pitch + n] = t[r];
        }
      } else if (n_ok) {
        if (p.out_f32 != nullptr) {
          float* op = p.out_f32 + static_cast<size_t>(row0) * p.ld_out + n;
#pragma unroll
          for (int r = 0; r < 32; ++r) if (r < rows_valid) op[static_cast<size_t>(r) * p.ld_out] = t[r];
        }
        if (out_act != nullptr) {
          T* op = out_act + static_cast<size_t>(row0) * p.ld_act + n;
#pragma unroll
          for (int r = 0; r < 32; ++r) if (r < rows_valid) op[static_cast<size_t>(r) * p.ld_act] = Tr::to(t[r]);
        }
      }
    }
    if constexpr (kLN) {
      // ---- fused LayerNorm(s) over the rows this warp owns (every column of a row is in this CTA) ----
      const float inv_n = 1.0f / static_cast<float>(p.N);
      float mean, rstd;                                    // of row `lane` (thread = row orientation)
      auto row_stats = [&]() {                             // two-pass (mean, then centred squares) like nn.LayerNorm
        __syncwarp();
        const float* tr = tile + lane * tile_pitch;
        float s = 0.f;
        for (int n = 0; n < p.N; ++n) s += tr[n];
        mean = s * inv_n;
        float sq = 0.f;
        for (int n = 0; n < p.N; ++n) { const float dlt = tr[n] - mean; sq = fmaf(dlt, dlt, sq); }
        rstd = rsqrtf(sq * inv_n + p.ln_eps);
        __syncwarp();
      };
      row_stats();
      const float* g_fin = p.ln1_g; const float* b_fin = p.ln1_b;
      if (p.ln_mode == 2) {                                // block norm in place, then the next module's LayerNorm
        for (int c0 = 0; c0 < p.N; c0 += 32) {
          const int n = c0 + lane;
          if (n < p.N) {
            const float g = __ldg(p.ln1_g + n), b = __ldg(p.ln1_b + n);
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const float m_r = __shfl_sync(0xffffffffu, mean, r), rs_r = __shfl_sync(0xffffffffu, rstd, r);
              tile[r * tile_pitch + n] = (tile[r * tile_pitch + n] - m_r) * rs_r * g + b;
            }
          } else {
#pragma unroll
            for (int r = 0; r < 32; ++r) { __shfl_sync(0xffffffffu, mean, r); __shfl_sync(0xffffffffu, rstd, r); }
          }
        }
        g_fin = p.ln2_g; b_fin = p.ln2_b;
        if (g_fin != nullptr) row_stats(); else __syncwarp();
      }
      T* ln_out = reinterpret_cast<T*>(p.ln_out);
      T* copy_out = reinterpret_cast<T*>(p.copy_out);
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        const int n = c0 + lane;
        const bool n_ok = n < p.N;
        const float g = (g_fin != nullptr && n_ok) ? __ldg(g_fin + n) : 1.f, b = (g_fin != nullptr && n_ok) ? __ldg(b_fin + n) : 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const float m_r = __shfl_sync(0xffffffffu, mean, r), rs_r = __shfl_sync(0xffffffffu, rstd, r);
          if (n_ok && r < rows_valid) {
            const int m = row0 + r;
            const float x = tile[r * tile_pitch + n];
            p.out_f32[static_cast<size_t>(m) * p.ld_out + n] = x;
            if (ln_out != nullptr) ln_out[static_cast<size_t>(m) * p.ld_ln + n] = Tr::to(g_fin != nullptr ? (x - m_r) * rs_r * g + b : x);
            if (copy_out != nullptr) {
              const int seq = m / p.frames_per_seq, tt = m - seq * p.frames_per_seq;
              if (tt % p.copy_stride == 0)
                copy_out[(static_cast<size_t>(seq) * p.frames_out_per_seq + tt / p.copy_stride) * p.N + n] = Tr::to(x);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2D K-major operand [rows, K] -> tensor map with a (128-byte x box_rows) box and 128B swizzle.
static int make_operand_map(CUtensorMap* map, int precision, const void* ptr, int rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  EC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  const int esize = precision == EC_PREC_TF32 ? 4 : 2;
  const size_t pitch = static_cast<size_t>(K) * esize;
  EC_REQUIRE(pitch % 16 == 0, "GEMM operand row pitch (K * element size) must be a multiple of 16 bytes");
  EC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand must be 16-byte aligned");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esize), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, precision == EC_PREC_TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
  return EC_OK;
}

static int pick_block_n(int N) {
  const int tiles = cdiv(N, 256);
  return round_up(cdiv(N, tiles), 16);
}

template <typename T, bool kLN>
static int launch_gemm_t(int precision, const GemmArgs& a, cudaStream_t stream) {
  using Tr = ActTraits<T>;
  EC_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "empty GEMM");
  GemmDev p{};
  p.M = a.M; p.N = a.N; p.K = a.K;
  int tiles_n;
  if (a.glu_nb > 0) {
    EC_REQUIRE(a.glu_nb % 8 == 0 && 2 * a.glu_nb <= 256 && (2 * a.glu_nb) % 16 == 0, "invalid GLU tile width");
    p.block_n = 2 * a.glu_nb;
    EC_REQUIRE(a.N % p.block_n == 0, "GLU weight rows must be a whole number of tiles");
    tiles_n = a.N / p.block_n;
  } else {
    p.block_n = pick_block_n(a.N);
    tiles_n = cdiv(a.N, p.block_n);
  }
  p.num_k_blocks = cdiv(a.K, Tr::kBlockK);
  const int stage_bytes = kATileBytes + p.block_n * 128;
  const int fixed = kStagingBytes + (2 * kMaxStages + 2) * 8 + 1024;
  // up to 148 CTAs: one CTA per SM anyway -> deep ring (hides the TMA->MMA->refill round trip); otherwise 2 CTAs per SM
  const int ctas = cdiv(a.M, kBlockM) * tiles_n;
  const int budget = (ctas <= 148 || kLN) ? 208 * 1024 : 113 * 1024;
  int stages = (budget - fixed) / stage_bytes;
  if (stages < 2) stages = 2;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > p.num_k_blocks) stages = p.num_k_blocks;
  p.stages = stages;
  int cols = 32;
  while (cols < p.block_n) cols <<= 1;
  // the GLU epilogue reads 32-column chunks of the gate half; keep the last (partial) chunk inside the allocation
  if (a.glu_nb > 0 && a.glu_nb % 32 != 0 && 2 * a.glu_nb + 32 > cols) cols <<= 1;
  EC_REQUIRE(cols <= 512, "tile does not fit in tensor memory");
  p.tmem_cols = cols;
  p.bias = a.bias; p.alpha = a.alpha; p.act = a.act;
  p.glu_nb = a.glu_nb; p.glu_channels = a.glu_channels;
  p.residual = a.residual; p.ld_res = a.ld_res;
  p.out_f32 = a.out_f32; p.ld_out = a.ld_out;
  p.out_act = a.out_act; p.ld_act = a.ld_act;
  p.round_out = a.round_out;
  EC_REQUIRE(a.out_f32 != nullptr || a.out_act != nullptr, "GEMM needs at least one output");
  EC_REQUIRE(a.glu_nb == 0 || a.bias != nullptr, "GLU GEMM needs a bias");
  if (kLN) {
    EC_REQUIRE(a.ln_mode == 1 || a.ln_mode == 2, "bad LayerNorm mode");
    EC_REQUIRE(tiles_n == 1 && a.glu_nb == 0 && a.act == GEMM_ACT_NONE, "fused LayerNorm needs the whole row in one plain tile (N <= 256)");
    EC_REQUIRE(a.out_f32 != nullptr && a.ld_out == a.N, "fused LayerNorm normalises the fp32 output in place");
    EC_REQUIRE(a.ln1_g != nullptr && a.ln1_b != nullptr, "missing LayerNorm parameters");
    p.ln_mode = a.ln_mode; p.ln1_g = a.ln1_g; p.ln1_b = a.ln1_b; p.ln2_g = a.ln2_g; p.ln2_b = a.ln2_b; p.ln_eps = a.ln_eps;
    p.ln_out = a.ln_out; p.ld_ln = a.N;
    p.copy_out = a.copy_out; p.copy_stride = a.copy_stride > 0 ? a.copy_stride : 1;
    p.frames_per_seq = a.frames_per_seq > 0 ? a.frames_per_seq : a.M; p.frames_out_per_seq = a.frames_out_per_seq;
    EC_REQUIRE(a.copy_out == nullptr || a.ln_mode == 1, "the strided copy is only available with a single LayerNorm");
  }

  CUtensorMap tmA, tmB;
  EC_TRY(make_operand_map(&tmA, precision, a.A, a.M, a.K, kBlockM));
  EC_TRY(make_operand_map(&tmB, precision, a.W, a.N, a.K, p.block_n));

  size_t pipe_bytes = static_cast<size_t>(stages) * stage_bytes;
  if (kLN) pipe_bytes = std::max(pipe_bytes, static_cast<size_t>(kBlockM) * (a.N | 1) * sizeof(float));   // row tile aliases the stages
  p.pipe_bytes = static_cast<int>(pipe_bytes);
  const size_t smem = pipe_bytes + fixed;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel<T, kLN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  EC_CUDA(attr_err);
  dim3 grid(cdiv(a.M, kBlockM), tiles_n);
  EC_TRY(launch_pdl(gemm_tc_kernel<T, kLN>, grid, dim3(192), smem, stream, tmA, tmB, p));
  return EC_OK;
}

int launch_gemm(int precision, const GemmArgs& a, cudaStream_t stream) {
  if (precision == EC_PREC_TF32) return a.ln_mode ? launch_gemm_t<float, true>(precision, a, stream) : launch_gemm_t<float, false>(precision, a, stream);
  if (precision == EC_PREC_BF16)
    return a.ln_mode ? launch_gemm_t<__nv_bfloat16, true>(precision, a, stream) : launch_gemm_t<__nv_bfloat16, false>(precision, a, stream);
  EC_FAIL("unknown precision");
}

}  // namespace ec
