// Front end as ONE kernel (inference, one-layer Conv2d subsampling):
//     x[b, t, :] = Linear( reshape( Swish( BatchNorm2d( Conv2d(1 -> C, 3x3, stride 2, pad 1)(mel) ) ) ) )
// (reference models/modules.py:232-249 Conv2dSubsampling.forward + models/encoders.py:113-116 transpose + Linear).
//
// The reference -- and the two-kernel path of subsample.cu + gemm_tc.cu -- materialise the (B*T/2) x (C*F/2) operand of that Linear
// (154 MB bf16 / 307 MB packed at B = 32 x 1000: ~20x the algorithmic bytes of the front end, 252 MB written + 312 MB read in the ncu
// capture of round 2).  Here the operand never exists in global memory: the 16 producer warps of a CTA compute the convolution
// outputs of one 128-frame tile k-block by k-block straight into the 128B-swizzled shared-memory A tiles the tensor core reads, the
// TMA warp streams the matching Linear-weight k-blocks, one thread issues tcgen05.mma into a TMEM accumulator, and the producers
// finish as the epilogue (bias, fp32 rows).  Global traffic = mel in + weights + x out.
//
//   * K order: the Linear contracts over (c, f) pairs in any order as long as both operands agree, so the weight is permuted once at
//     prepare time to k' = f*Cp + c (launch_linear_weight_permute; Cp = C rounded up to the channels of one 16-byte operand chunk,
//     zero columns in the pad): a producer thread owns 4 frames x one chunk (4 / 8 consecutive k' of ONE mel bin) per k-block, with
//     the four 3x3 mel patches in registers and the taps as warp-wide shared-memory broadcasts, each feeding 4 outputs -- no bin
//     boundary inside a chunk, so the unrolled loops are branch free and the 16 / 32 outputs overlap in the pipeline.  The 16
//     producer warps form two groups that fill alternate k-blocks (A stage g <- k-blocks kb = g mod 2).
//   * The mel frames of the tile (2*128 + 1 of them) are staged once, split into even / odd frame planes so that the stride-2 tap
//     columns of 32 consecutive frames are conflict-free stride-1 reads.
//   * Synchronisation: full_a[g] (8 producer-warp arrivals after fence.proxy.async: generic-proxy stores -> tensor-core reads),
//     empty_a[g] / empty_w[s] (tcgen05.commit), full_w[s] (TMA bytes; the weight ring is as deep as shared memory allows), tmem_full.
#include "ec_common.cuh"
#include "ec_tma.cuh"
#include <algorithm>
#include <mutex>

namespace ec {

namespace {
constexpr int kSfRows = 128;                 // output frames per CTA
constexpr int kSfProducerWarps = 16;
constexpr int kSfThreads = 64 + 32 * kSfProducerWarps;
constexpr int kSfPitchE = 130, kSfPitchO = 128;      // even / odd mel-frame planes (129 / 128 entries per mel bin)
constexpr int kSfAStages = 2;                // A tiles: one per producer group (8 warps each); group g fills the k-blocks kb = g mod 2
constexpr int kSfRowsPerThread = 4;          // frames per producer thread: every tap load (a warp-wide 16-byte broadcast, 4 shared-memory
                                             // cycles) then feeds 4 outputs -- with 1 the kernel was bound by shared-memory return bandwidth
constexpr int kSfMaxWStages = 6;             // weight tiles: as deep as shared memory allows (the TMA round trip, ~1.5 us, is longer
                                             // than the producers need for a k-block: a 2-deep ring made the kernel latency bound)

struct SfDev {
  const float* mel; const float* w; const float* b; const float* lin_b; float* out;
  int F, T_in, T_out, C, Cp, D0, K, block_n, num_kb, w_stages, tmem_cols;
  int off_w, off_patch, off_taps, off_bars;  // byte offsets from the 1024-aligned base (A ring at 0, then the weight ring)
};

// One 16-byte chunk of an A-tile row = kChunk consecutive k of one frame (128B swizzle: chunk index ^ (row & 7)).
template <typename T> struct SfPack;
template <> struct SfPack<__nv_bfloat16> {
  static constexpr int kChunk = 8;
  __device__ static void store(uint8_t* a_tile, int r, int chunk, const float* v) {
    uint4 pk;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
    pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(a_tile + r * 128 + ((chunk ^ (r & 7)) << 4)) = pk;
  }
};
template <> struct SfPack<SplitBf16> {
  static constexpr int kChunk = 4;
  __device__ static void store(uint8_t* a_tile, int r, int chunk, const float* v) {
    *reinterpret_cast<uint4*>(a_tile + r * 128 + ((chunk ^ (r & 7)) << 4)) = make_uint4(split_pack(v[0]), split_pack(v[1]), split_pack(v[2]), split_pack(v[3]));
  }
};
template <> struct SfPack<float> {
  static constexpr int kChunk = 4;
  __device__ static void store(uint8_t* a_tile, int r, int chunk, const float* v) {
    *reinterpret_cast<float4*>(a_tile + r * 128 + ((chunk ^ (r & 7)) << 4)) = make_float4(round_tf32(v[0]), round_tf32(v[1]), round_tf32(v[2]), round_tf32(v[3]));
  }
};
}  // namespace

template <typename T>
__global__ void __launch_bounds__(kSfThreads, 1)
subsample_linear_fused_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW2, const SfDev p) {
  using Tr = ActTraits<T>;
  constexpr bool kSplit = IsSplit<T>::value;
  constexpr int KB = Tr::kBlockK;                      // k elements per 128-byte k-block
  constexpr int NCH = SfPack<T>::kChunk;               // channels per producer thread and k-block (one 16-byte chunk per frame)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b_tile_bytes = p.block_n * 128;
  const int w_stage_bytes = (kSplit ? 2 : 1) * b_tile_bytes;
  float* patchE = reinterpret_cast<float*>(base_ptr + p.off_patch);
  float* patchO = patchE + (p.F + 2) * kSfPitchE;
  float* taps = reinterpret_cast<float*>(base_ptr + p.off_taps);          // [C][12]: 9 taps, folded bias, 2 pad
  const uint32_t bars_addr = base + p.off_bars;
  auto full_a = [&](int s) { return bars_addr + 8u * s; };
  auto empty_a = [&](int s) { return bars_addr + 8u * (kSfAStages + s); };
  auto full_w = [&](int s) { return bars_addr + 8u * (2 * kSfAStages + s); };
  auto empty_w = [&](int s) { return bars_addr + 8u * (2 * kSfAStages + kSfMaxWStages + s); };
  const uint32_t tmem_full_bar = bars_addr + 8u * (2 * kSfAStages + 2 * kSfMaxWStages);
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(base_ptr + p.off_bars + 8 * (2 * kSfAStages + 2 * kSfMaxWStages + 1));

  const int b = blockIdx.y, t0 = blockIdx.x * kSfRows;
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < kSfAStages; ++s) { mbar_init(full_a(s), kSfProducerWarps / kSfAStages); mbar_init(empty_a(s), 1); }
    for (int s = 0; s < p.w_stages; ++s) { mbar_init(full_w(s), 1); mbar_init(empty_w(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), p.tmem_cols);
  // taps are weights (input independent): staged before the dependency wait
  for (int i = threadIdx.x; i < p.Cp * 12; i += kSfThreads) {
    const int c = i / 12, k = i - c * 12;
    taps[i] = c < p.C ? (k < 9 ? p.w[c * 9 + k] : (k == 9 ? p.b[c] : 0.f)) : 0.f;     // pad channels: Swish(0) = 0 against zero weights
  }
  grid_dependency_wait();
  grid_launch_dependents();
  // mel frames 2*t0 - 1 .. 2*t0 + 255 of every bin (plus a zero bin above and below), even / odd frame planes
  {
    const float* melb = p.mel + static_cast<size_t>(b) * p.F * p.T_in;
    const int tbase = 2 * t0 - 1;
    for (int i = threadIdx.x; i < (p.F + 2) * 257; i += kSfThreads) {
      const int fr = i / 257, col = i - fr * 257;
      const int f = fr - 1, t = tbase + col;
      const float v = (f >= 0 && f < p.F && t >= 0 && t < p.T_in) ? __ldg(melb + static_cast<size_t>(f) * p.T_in + t) : 0.f;
      if (col & 1) patchO[fr * kSfPitchO + (col >> 1)] = v; else patchE[fr * kSfPitchE + (col >> 1)] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp_idx == 0) {
    // ---------------- TMA: Linear-weight k-blocks ----------------
    int s = 0; uint32_t ph = 1;
    for (int kb = 0; kb < p.num_kb; ++kb) {
      mbar_wait(empty_w(s), ph);
      if (elect_one()) {
        mbar_arrive_expect_tx(full_w(s), static_cast<uint32_t>(w_stage_bytes));
        const uint32_t dst = base + p.off_w + s * w_stage_bytes;
        tma_load_2d(dst, &tmW, full_w(s), kb * KB, 0);
        if constexpr (kSplit) tma_load_2d(dst + b_tile_bytes, &tmW2, full_w(s), kb * KB, 0);
      }
      __syncwarp();
      if (++s == p.w_stages) { s = 0; ph ^= 1; }
    }
  } else if (warp_idx == 1) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = make_idesc(Tr::kTf32 ? 2u : 1u, kSfRows, p.block_n);
    int sa = 0, sw = 0; uint32_t pa = 0, pw = 0;
    for (int kb = 0; kb < p.num_kb; ++kb) {
      mbar_wait(full_a(sa), pa);
      mbar_wait(full_w(sw), pw);
      tc_fence_after();
      const uint64_t da = make_smem_desc_sw128(base + sa * kATileBytes);
      const uint64_t db = make_smem_desc_sw128(base + p.off_w + sw * w_stage_bytes);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc_mma<Tr::kTf32>(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          if constexpr (kSplit)
            tc_mma<false>(tmem_base, da + 2 * k, make_smem_desc_sw128(base + p.off_w + sw * w_stage_bytes + b_tile_bytes) + 2 * k, idesc, 1u);
        }
        tc_commit(empty_a(sa));
        tc_commit(empty_w(sw));
        if (kb == p.num_kb - 1) tc_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++sa == kSfAStages) { sa = 0; pa ^= 1; }
      if (++sw == p.w_stages) { sw = 0; pw ^= 1; }
    }
  } else {
    // ---------------- producers: convolution outputs -> swizzled A tiles ----------------
    const int ptid = threadIdx.x - 64;
    const int g = ptid >> 8, rp = ptid & 31, chunk = (ptid >> 5) & 7;   // producer group; frames rp + 32 j; 16-byte chunk of the k-block
    const int Cp = p.Cp, F2 = p.F / 2;
    uint8_t* a_tile = base_ptr + g * kATileBytes;
    uint32_t ph = 1;
    for (int kb = g; kb < p.num_kb; kb += kSfAStages, ph ^= 1) {
      const int kk = kb * KB + chunk * NCH;
      const int f = kk / Cp, c0 = kk - f * Cp;                       // Cp % NCH == 0: the whole chunk lies in mel bin f
      float v[kSfRowsPerThread][NCH];
      if (f < F2) {
        float pv[kSfRowsPerThread][9];
#pragma unroll
        for (int j = 0; j < kSfRowsPerThread; ++j) {
          const int r = rp + 32 * j;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const int fr = 2 * f + kh;                               // mel bin 2f - 1 + kh, +1 for the zero bin below
            pv[j][kh * 3 + 0] = patchE[fr * kSfPitchE + r];
            pv[j][kh * 3 + 1] = patchO[fr * kSfPitchO + r];
            pv[j][kh * 3 + 2] = patchE[fr * kSfPitchE + r + 1];
          }
        }
        const float4* tp = reinterpret_cast<const float4*>(taps + c0 * 12);
#pragma unroll
        for (int e = 0; e < NCH; ++e) {
          const float4 w0 = tp[3 * e], w1 = tp[3 * e + 1], w2 = tp[3 * e + 2];
#pragma unroll
          for (int j = 0; j < kSfRowsPerThread; ++j) {
            float a = w2.y;                                          // folded bias
            a = fmaf(w0.x, pv[j][0], a); a = fmaf(w0.y, pv[j][1], a); a = fmaf(w0.z, pv[j][2], a);
            a = fmaf(w0.w, pv[j][3], a); a = fmaf(w1.x, pv[j][4], a); a = fmaf(w1.y, pv[j][5], a);
            a = fmaf(w1.z, pv[j][6], a); a = fmaf(w1.w, pv[j][7], a); a = fmaf(w2.x, pv[j][8], a);
            v[j][e] = a;
          }
        }
#pragma unroll
        for (int j = 0; j < kSfRowsPerThread; ++j)
#pragma unroll
          for (int e = 0; e < NCH; ++e) v[j][e] = swish_fn<T>(v[j][e]);
      } else {
#pragma unroll
        for (int j = 0; j < kSfRowsPerThread; ++j)
#pragma unroll
          for (int e = 0; e < NCH; ++e) v[j][e] = 0.f;               // k' beyond K: zero (the weight tile is zero-filled there as well)
      }
      mbar_wait(empty_a(g), ph);                                     // the MMAs that read this slot have retired
#pragma unroll
      for (int j = 0; j < kSfRowsPerThread; ++j) SfPack<T>::store(a_tile, rp + 32 * j, chunk, v[j]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_a(g));
    }
    // ---------------- epilogue: accumulator rows + bias -> x (fp32) ----------------
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp_idx & 3, sub = (warp_idx - 2) >> 2;           // TMEM lane quarter of this warp; chunk interleave among its 4 warps
    const int row = q * 32 + lane, t = t0 + row;
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int n_chunks = (p.D0 + 31) / 32;
    for (int ch = sub; ch < n_chunks; ch += 4) {
      uint32_t acc[32];
      tmem_ld_32x32(tbase + 32u * ch, acc);
      tmem_ld_wait();
      if (t < p.T_out) {
        float* o = p.out + (static_cast<size_t>(b) * p.T_out + t) * p.D0 + 32 * ch;
        const int nc = min(32, p.D0 - 32 * ch);
        if (nc == 32 && (p.D0 & 3) == 0) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.lin_b + 32 * ch) + j4);
            *reinterpret_cast<float4*>(o + 4 * j4) = make_float4(__uint_as_float(acc[4 * j4]) + bb.x, __uint_as_float(acc[4 * j4 + 1]) + bb.y,
                                                                  __uint_as_float(acc[4 * j4 + 2]) + bb.z, __uint_as_float(acc[4 * j4 + 3]) + bb.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < nc) o[j] = __uint_as_float(acc[j]) + __ldg(p.lin_b + 32 * ch + j);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------------
int subsample_fused_cpad(int precision, int C) { return round_up(C, precision == EC_PREC_BF16 ? SfPack<__nv_bfloat16>::kChunk : SfPack<float>::kChunk); }

struct SfPlan { int block_n, w_stages, tmem_cols, off_w, off_patch, off_taps, off_bars; size_t smem; bool ok; };
static SfPlan sf_plan(int precision, int F, int C, int D0) {
  SfPlan pl{};
  pl.ok = false;
  if (D0 > 256 || F % 2 != 0 || F / 2 > 128) return pl;
  pl.block_n = round_up(D0, 16);
  const int planes = precision == EC_PREC_BF16X2 ? 2 : 1;
  const int w_stage = planes * pl.block_n * 128;
  const size_t patch = static_cast<size_t>(F + 2) * (kSfPitchE + kSfPitchO) * 4, taps = static_cast<size_t>(subsample_fused_cpad(precision, C)) * 12 * 4;
  const size_t bars = 8 * (2 * kSfAStages + 2 * kSfMaxWStages + 1) + 16;
  const size_t fixed = align_up(patch, 16) + align_up(taps, 16) + bars + 1024 + kSfAStages * kATileBytes;
  if (fixed + 2 * static_cast<size_t>(w_stage) > 227 * 1024) return pl;
  pl.w_stages = std::min(static_cast<int>((227 * 1024 - fixed) / w_stage), kSfMaxWStages);
  pl.off_w = kSfAStages * kATileBytes;
  pl.off_patch = pl.off_w + pl.w_stages * w_stage;
  pl.off_taps = pl.off_patch + static_cast<int>(align_up(patch, 16));
  pl.off_bars = pl.off_taps + static_cast<int>(align_up(taps, 16));
  pl.smem = pl.off_bars + bars + 1024;
  int cols = 32;
  while (cols < pl.block_n) cols <<= 1;
  pl.tmem_cols = cols;
  pl.ok = pl.smem <= 227 * 1024;
  return pl;
}
bool subsample_fused_fits(int precision, int F, int C, int D0) { return sf_plan(precision, F, C, D0).ok; }

template <typename T>
static int launch_sf_t(int precision, const SubFusedArgs& a, cudaStream_t st) {
  const SfPlan pl = sf_plan(precision, a.F, a.C, a.D0);
  EC_REQUIRE(pl.ok, "fused front end: shape does not fit (D0 <= 256, shared memory)");
  const int T_out = (a.T - 1) / 2 + 1, Cp = subsample_fused_cpad(precision, a.C), K = Cp * (a.F / 2);
  SfDev p{};
  p.mel = a.mel; p.w = a.w; p.b = a.b; p.lin_b = a.lin_b; p.out = a.out;
  p.F = a.F; p.T_in = a.T; p.T_out = T_out; p.C = a.C; p.Cp = Cp; p.D0 = a.D0; p.K = K;
  p.block_n = pl.block_n; p.num_kb = cdiv(K, ActTraits<T>::kBlockK); p.w_stages = pl.w_stages; p.tmem_cols = pl.tmem_cols;
  p.off_w = pl.off_w; p.off_patch = pl.off_patch; p.off_taps = pl.off_taps; p.off_bars = pl.off_bars;
  CUtensorMap tmW, tmW2;
  EC_TRY(make_operand_map(&tmW, precision, a.lin_w_perm, a.D0, K, pl.block_n));
  tmW2 = tmW;
  if (IsSplit<T>::value) {
    const uint8_t* twin = reinterpret_cast<const uint8_t*>(a.lin_w_perm) + static_cast<size_t>(a.D0) * K * 4;
    EC_TRY(make_operand_map(&tmW2, precision, twin, a.D0, K, pl.block_n));
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(subsample_linear_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  EC_CUDA(attr_err);
  dim3 grid(cdiv(T_out, kSfRows), a.B);
  return launch_pdl(subsample_linear_fused_kernel<T>, grid, dim3(kSfThreads), pl.smem, st, tmW, tmW2, p);
}

int launch_subsample_linear_fused(int precision, const SubFusedArgs& a, cudaStream_t st) {
  EC_REQUIRE(a.mel && a.w && a.b && a.lin_w_perm && a.lin_b && a.out && a.B > 0 && a.T > 0, "fused front end: bad argument");
  if (precision == EC_PREC_BF16) return launch_sf_t<__nv_bfloat16>(precision, a, st);
  if (precision == EC_PREC_BF16X2) return launch_sf_t<SplitBf16>(precision, a, st);
  if (precision == EC_PREC_TF32) return launch_sf_t<float>(precision, a, st);
  EC_FAIL("unknown precision");
}

}  // namespace ec
