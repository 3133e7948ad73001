// tcgen05 GEMM with fused epilogues:
//     out = alpha * act(A @ W^T + bias) [+ residual]          (optionally GLU over column halves)
//     [+ LayerNorm(out) / LayerNorm(LayerNorm(out)) written in the activation type for the next GEMM]
//
// Replaces every dense contraction on the hot path (reference models/layers.py:67 F.linear, :136 pointwise F.conv1d):
// FFN projections, QKV / positional / output projections, pointwise convs of the convolution module, conv_res,
// the subsampling Linear and the CTC fc head -- and, through the fused epilogue, the five nn.LayerNorm(eps=1e-6) of a
// block (reference models/modules.py:386,433,511; models/blocks.py:96,135).
//
// One CTA = one 128 x BLOCK_N output tile over the whole K.  Warp roles (2 + 16 warps; 2 + 8 in the fused-LayerNorm variant):
//   warp 0    TMA producer : cp.async.bulk.tensor loads of the A (128 x 128B) and W (BLOCK_N x 128B) k-slices, 128B swizzle
//   warp 1    MMA issuer   : allocates TMEM, one thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 32B), fp32 accumulator in TMEM
//   warps 2-.. epilogue    : four (two) warps per TMEM lane quarter share its 32 rows and interleave the 32-column chunks: the
//                            per-row work is a dependent chain, so the epilogue is latency bound per warp, not throughput bound.
//                            thread = output row.  Per 32-column chunk: tcgen05.ld -> bias (smem broadcast) / Swish / GLU /
//                            alpha / residual (TMA-prefetched swizzled slab) in registers -> swizzled smem slab (conflict-free
//                            16-byte stores) -> TMA bulk tensor store (coalesced, clips the M / N tails).  LayerNorm row
//                            statistics are per-thread (chunk-wise Chan/Welford merge, no shuffles); the fp32 row tile stays
//                            in shared memory (aliasing the drained operand ring) for the normalise passes.
// K and N tails are zero-filled by TMA out-of-bounds handling, so D, 4D, head dims etc. need no host-side padding
// (only 16-byte row pitches).  Operand type float => kind::tf32, __nv_bfloat16 => kind::f16.
#include "ec_common.cuh"
#include "ec_tma.cuh"
#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace ec {

struct GemmDev {
  int M, N, K;
  int block_n;       // UMMA N (multiple of 16, <= 256)
  int num_k_blocks, stages;
  int pipe_bytes;    // operand ring bytes (epilogue staging aliases it after the mainloop)
  int warp_stage_bytes;
  int nbuf;          // output slab buffers per warp (1, 2, 4 or 8): the warp's i-th chunk uses buffer i % nbuf
  int res_depth;     // residual TMA ring depth per warp (1 or 2)
  int epi_batch;     // 1: one proxy fence + all TMA stores after the chunk loop (needs n_chunks <= nbuf)
  int tmem_cols;
  const float* bias;
  float alpha;
  int act;
  int glu_nb, glu_channels;
  int has_res, has_out_f32, has_out_act;
  int round_out;     // round the fp32 output to TF32 (it feeds a TF32 mma.sync consumer)
  int act_f16;       // split mode: the activation-type output is plain fp16 (64-byte slab rows)
  // fused LayerNorm (kLN instantiation only; requires a single N tile)
  int ln_mode;       // 1: y = LN1(out);  2: out <- LN1(out), y = LN2(out) (plain copy when ln2_g == nullptr)
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  float ln_eps;
  int has_ln_out;
  int dbg;
  void* copy_out; int copy_stride, frames_per_seq, frames_out_per_seq;   // strided compaction of `out` (LN mode 1), activation type
  // training-step epilogue (plain variant): counter-based dropout / Swish side output / Swish-dropout backward (see GemmArgs)
  const unsigned long long* drop_ctr; unsigned drop_keep16; float drop_inv_keep;
  unsigned drop_site, drop_site2, drop_site_aux;
  int has_out_act2, aux_mode;
  int res_tx;        // bytes of one residual / aux slab (4096; 2048 for a bf16 aux operand)
};

// Optional in-kernel timeline (SM clock stamps of CTA (0,0)), enabled through ec_debug_gemm_timeline for latency studies.
__device__ unsigned long long g_gemm_timeline[16];
static int g_timeline_enabled = 0;
static int g_block_n_override = 0;          // debug: force the N tile of the plain GEMM (multiple of 32 when N spans several tiles)
__device__ __forceinline__ void stamp(int enabled, int slot) {
  if (enabled && blockIdx.x == 0 && blockIdx.y == 0) g_gemm_timeline[slot] = clock64();
}

constexpr int kVecFloats = 288;              // bias / LayerNorm vectors in smem (256 + one chunk of slack)
constexpr int kVecBytes = 5 * kVecFloats * 4 + 2 * 4 * 4 * 32 * 3 * 4;   // + LayerNorm statistics exchange [stage][sub][quarter][lane][3]
constexpr int kMaxStages = 8;
constexpr int kNumBars = 2 * kMaxStages + 1 + 32;

// 16 epilogue warps: the four warps (q, sub) of TMEM lane quarter q share its 32 rows and take the 32-column chunks sub,
// sub + 4, ...; the LayerNorm statistics of a row are merged through shared memory in a fixed order.
template <typename T, bool kLN>
__global__ void __launch_bounds__(576, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmOutF,
               const __grid_constant__ CUtensorMap tmOutA, const __grid_constant__ CUtensorMap tmLn, const __grid_constant__ CUtensorMap tmOutA2,
               const GemmDev p) {
  using Tr = ActTraits<T>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool kSplit = IsSplit<T>::value;       // split mode: a second W tile (halves swapped) and a second MMA per k-slice
  const int b_tile_bytes = p.block_n * 128;
  const int stage_bytes = kATileBytes + (kSplit ? 2 : 1) * b_tile_bytes;
  constexpr int kEpiWarps = 16, kHalves = kEpiWarps / 4;                 // warps per lane quarter = chunk interleave factor
  const int res_bytes = p.has_res ? kEpiWarps * p.res_depth * kSlabBytes : 0;     // res_depth 0: residual lands in the x slabs
  uint8_t* res_ring = base_ptr + p.pipe_bytes;
  float* vecs = reinterpret_cast<float*>(base_ptr + p.pipe_bytes + res_bytes);          // bias | ln1_g | ln1_b | ln2_g | ln2_b
  uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + p.pipe_bytes + res_bytes + kVecBytes);
  const uint32_t bars_addr = base + p.pipe_bytes + res_bytes + kVecBytes;
  // bars: [0,8) full, [8,16) empty, [16] tmem_full, [17,25) residual slots (warp q, slot s -> 17 + 2q + s), then the TMEM address
  auto full_bar = [&](int s) { return bars_addr + 8u * s; };
  auto empty_bar = [&](int s) { return bars_addr + 8u * (kMaxStages + s); };
  const uint32_t tmem_full_bar = bars_addr + 8u * (2 * kMaxStages);
  auto res_bar = [&](int ew, int s) { return bars_addr + 8u * (2 * kMaxStages + 1 + 2 * ew + s); };
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(bars + kNumBars);

  if (threadIdx.x == 0) stamp(p.dbg, 0);
  const int m0 = blockIdx.x * kBlockM;
  const int tile_n = blockIdx.y;
  const int w_row0 = tile_n * p.block_n;   // first W row (and, for plain GEMMs, first output column) of this tile

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(res_bar(i >> 1, i & 1), 1);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (threadIdx.x == 0) stamp(p.dbg, 1);
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on we read its output.
  grid_dependency_wait();
  grid_launch_dependents();
  if (threadIdx.x == 0) stamp(p.dbg, 2);

  // Producer and MMA loops run warp-uniform; only the TMA / tcgen05 instructions sit under elect_one() (a lane-0 branch makes
  // the compiler wrap every uniform-datapath instruction in an ELECT/BRA loop, which doubles the issue time per MMA).
  if (warp_idx == 0) {
    const uint32_t tx = static_cast<uint32_t>(stage_bytes);
    int s = 0; uint32_t ph = 1;
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      mbar_wait(empty_bar(s), ph);
      if (elect_one()) {
        mbar_arrive_expect_tx(full_bar(s), tx);
        const uint32_t a_dst = base + s * stage_bytes;
        tma_load_2d(a_dst, &tmA, full_bar(s), kb * Tr::kBlockK, m0);
        tma_load_2d(a_dst + kATileBytes, &tmB, full_bar(s), kb * Tr::kBlockK, w_row0);
        if constexpr (kSplit) tma_load_2d(a_dst + kATileBytes + b_tile_bytes, &tmB2, full_bar(s), kb * Tr::kBlockK, w_row0);
        if (kb == 0) stamp(p.dbg, 3);
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp_idx == 1) {
    const uint32_t idesc = make_idesc(Tr::kTf32 ? 2u : 1u, kBlockM, p.block_n);
    int s = 0; uint32_t ph = 0;
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint64_t da = make_smem_desc_sw128(base + s * stage_bytes);
      const uint64_t db = make_smem_desc_sw128(base + s * stage_bytes + kATileBytes);
      if (elect_one()) {
        if (kb == 0) stamp(p.dbg, 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 4 x 32-byte K slices inside the 128-byte swizzle row: start address advances by 32 B
          tc_mma<Tr::kTf32>(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          if constexpr (kSplit)         // [ah al] . [wl wh]: the cross terms, against the swapped plane of the weight operand
            tc_mma<false>(tmem_base, da + 2 * k, make_smem_desc_sw128(base + s * stage_bytes + kATileBytes + b_tile_bytes) + 2 * k, idesc, 1u);
        }
        tc_commit(empty_bar(s));         // frees the smem slot when these MMAs retire
        if (kb == p.num_k_blocks - 1) { tc_commit(tmem_full_bar); stamp(p.dbg, 5); }   // accumulator complete
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  } else {
    // ---------------- epilogue: warps 2..5 own TMEM lane quarters (warp_idx % 4); thread = output row ----------------
    const int q = warp_idx & 3;
    const int ew = warp_idx - 2;                             // epilogue warp index
    const int half = ew >> 2;                                // which chunk parity this warp owns (kLN only)
    const int et = ew * 32 + lane;                           // index among the epilogue threads
    const bool glu = !kLN && p.glu_nb > 0;                   // (the fused-LayerNorm variant is plain: no GLU, no activation)
    const int cols = glu ? p.glu_nb : p.block_n;             // logical output columns of this tile (multiple of 32 unless last tile)
    const int out_col0 = glu ? tile_n * p.glu_nb : w_row0;
    const int n_limit = glu ? p.glu_channels : p.N;
    const int row0 = m0 + q * 32;
    float* sbias = vecs;
    float *sg1 = vecs + kVecFloats, *sb1 = vecs + 2 * kVecFloats, *sg2 = vecs + 3 * kVecFloats, *sb2 = vecs + 4 * kVecFloats;
    // ---- per-column vectors -> smem (zero beyond the valid range) ----
    for (int i = et; i < kVecFloats; i += 32 * kEpiWarps) {
      float bv = 0.f;
      if (p.bias != nullptr) {
        if (glu) { if (i < 2 * p.glu_nb) bv = __ldg(p.bias + w_row0 + i); }
        else if (i < p.block_n && out_col0 + i < n_limit) bv = __ldg(p.bias + out_col0 + i);
      }
      sbias[i] = bv;
      if constexpr (kLN) {
        const bool ok = i < p.N;
        sg1[i] = ok ? __ldg(p.ln1_g + i) : 0.f; sb1[i] = ok ? __ldg(p.ln1_b + i) : 0.f;
        sg2[i] = (ok && p.ln2_g != nullptr) ? __ldg(p.ln2_g + i) : 0.f; sb2[i] = (ok && p.ln2_g != nullptr) ? __ldg(p.ln2_b + i) : 0.f;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");      // epilogue warps only
    // ---- residual slabs: 2-deep TMA ring per warp ----
    const int n_chunks = (min(cols, n_limit - out_col0) + 31) / 32;
    // this warp's i-th chunk is column chunk c = half + i * kHalves
    const int my_chunks = (n_chunks - half + kHalves - 1) / kHalves;
    uint8_t* wstage = base_ptr + (kLN ? q : ew) * p.warp_stage_bytes;     // aliases the operand ring: only touched after tmem_full
    // residual: a per-warp TMA ring prefetched during the mainloop, or (res_depth 0, fused LayerNorm with a wide row) straight
    // into the resident x slab of the chunk once the operand ring is drained
    const bool res_direct = kLN && p.res_depth == 0;
    const int rdepth = res_direct ? 2 : p.res_depth;       // a warp owns at most two chunks in the direct mode
    uint8_t* my_res = res_ring + ew * p.res_depth * kSlabBytes;
    const int rmask = rdepth - 1;
    auto res_slab = [&](int i) { return res_direct ? wstage + (half + i * kHalves) * kSlabBytes : my_res + (i & rmask) * kSlabBytes; };
    auto issue_res = [&](int i) {
      mbar_arrive_expect_tx(res_bar(ew, i & rmask), static_cast<uint32_t>(p.res_tx));
      tma_load_2d(smem_u32(res_slab(i)), &tmRes, res_bar(ew, i & rmask), out_col0 + (half + i * kHalves) * 32, row0);
    };
    if (p.has_res && !res_direct && lane == 0) {
      if (my_chunks > 0) issue_res(0);
      if (my_chunks > 1 && rdepth > 1) issue_res(1);
    }
    // plain: [F slabs (nbuf, if fp32 output) | A slabs (nbuf)]; kLN: [x slabs (n_chunks) | A slabs (n_chunks)]
    constexpr int kASlab = sizeof(T) == 4 ? kSlabBytes : kSlabBytes / 2;
    const int nbuf = p.nbuf;
    uint8_t* slabA = wstage + (kLN ? n_chunks : (p.has_out_f32 ? nbuf : 0)) * kSlabBytes + (kLN ? half * nbuf * kASlab : 0);
    uint8_t* slabA2 = slabA + nbuf * kASlab;                 // second activation-type output (plain variant, training step)
    // counter-based dropout: keys of the three possible sites; element index = row * N + column (groups of 4 share one 64-bit draw)
    unsigned long long key_out = 0, key_act2 = 0, key_aux = 0;
    if (p.drop_ctr != nullptr && p.drop_site) key_out = site_key(p.drop_ctr, p.drop_site);      // (also in the fused-LayerNorm variant)
    if (!kLN && p.drop_ctr != nullptr) {
      if (p.drop_site2) key_act2 = site_key(p.drop_ctr, p.drop_site2);
      if (p.drop_site_aux) key_aux = site_key(p.drop_ctr, p.drop_site_aux);
    }
    const unsigned long long grow = static_cast<unsigned long long>(row0 + lane) * static_cast<unsigned long long>(n_limit);
    float* stat_x = vecs + 5 * kVecFloats;                   // [stage 2][sub 4][quarter 4][lane 32][3]
    const bool batch = p.epi_batch != 0;
    float mean = 0.f, m2 = 0.f, cnt = 0.f;                   // running LayerNorm statistics of this thread's row
    if (et == 0) stamp(p.dbg, 6);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (et == 0) stamp(p.dbg, 7);
    if (p.has_res && res_direct && lane == 0) {
      if (my_chunks > 0) issue_res(0);
      if (my_chunks > 1) issue_res(1);
    }
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t v[32];
    if (my_chunks > 0) tmem_ld_32x32(tbase + 32u * half, v);  // software pipeline: the next chunk's accumulator load is in flight
    for (int i = 0; i < my_chunks; ++i) {
      const int c = half + i * kHalves;
      const int c0 = c * 32;
      const uint32_t taddr = tbase + static_cast<uint32_t>(c0);
      tmem_ld_wait();
      float t[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 b4 = *reinterpret_cast<const float4*>(sbias + c0 + 4 * j4);
        t[4 * j4] = __uint_as_float(v[4 * j4]) + b4.x; t[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + b4.y;
        t[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + b4.z; t[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + b4.w;
      }
      if (glu) {
        tmem_ld_32x32(taddr + static_cast<uint32_t>(p.glu_nb), v);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sbias + p.glu_nb + c0 + 4 * j4);
          t[4 * j4] *= sigmoid_fn<T>(__uint_as_float(v[4 * j4]) + b4.x); t[4 * j4 + 1] *= sigmoid_fn<T>(__uint_as_float(v[4 * j4 + 1]) + b4.y);
          t[4 * j4 + 2] *= sigmoid_fn<T>(__uint_as_float(v[4 * j4 + 2]) + b4.z); t[4 * j4 + 3] *= sigmoid_fn<T>(__uint_as_float(v[4 * j4 + 3]) + b4.w);
        }
      }
      if (i + 1 < my_chunks) tmem_ld_32x32(taddr + 32u * kHalves, v);   // v is consumed: prefetch this warp's next chunk
      if (!kLN && p.act == GEMM_ACT_SWISH) {
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = swish_fn<T>(t[j]);
      }
      if constexpr (!kLN) {
        if (i >= nbuf) {                                   // the stores issued nbuf chunks ago (one bulk group per chunk) have drained this buffer
          if (lane == 0) { if (nbuf > 1) bulk_wait_read(nbuf - 1); else bulk_wait_read0(); }
          __syncwarp();
        }
        const unsigned long long g0 = (grow + static_cast<unsigned long long>(out_col0 + c0)) >> 2;   // first 4-element group of this chunk
        if (p.has_out_act2) {
          // side output h = dropout(Swish(z)) with z as the backward will read it (rounded to the activation type)
          float h[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) h[j] = swish_fn<T>(Tr::from(Tr::to(t[j])));
          if (p.drop_site2) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const unsigned long long draw = splitmix64(key_act2 + g0 + j4);
#pragma unroll
              for (int l = 0; l < 4; ++l) h[4 * j4 + l] *= keep_factor(draw, l, p.drop_keep16, p.drop_inv_keep);
            }
          }
          slab_store_act<T>(slabA2 + (i & (nbuf - 1)) * kASlab, lane, h);
        }
        if (p.aux_mode == 1) {
          // data gradient through dropout(Swish(z)): acc * keep / (1 - p) * d/dz (z sigmoid z); z tile arrives through the residual ring
          mbar_wait(res_bar(ew, i & rmask), (i / rdepth) & 1);
          float zz[32];
          slab_load_act<T>(res_slab(i), lane, zz);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sg = fast_sigmoid(zz[j]);
            t[j] *= sg + zz[j] * sg * (1.f - sg);
          }
          if (p.drop_site_aux) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const unsigned long long draw = splitmix64(key_aux + g0 + j4);
#pragma unroll
              for (int l = 0; l < 4; ++l) t[4 * j4 + l] *= keep_factor(draw, l, p.drop_keep16, p.drop_inv_keep);
            }
          }
          __syncwarp();
          if (lane == 0 && i + rdepth < my_chunks) issue_res(i + rdepth);
        }
      }
      if (p.drop_site) {                                   // nn.Dropout behind the projection (both variants)
        const unsigned long long gd0 = (grow + static_cast<unsigned long long>(out_col0 + c0)) >> 2;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const unsigned long long draw = splitmix64(key_out + gd0 + j4);
#pragma unroll
          for (int l = 0; l < 4; ++l) t[4 * j4 + l] *= keep_factor(draw, l, p.drop_keep16, p.drop_inv_keep);
        }
      }
      if (p.has_res && (kLN || p.aux_mode == 0)) {
        mbar_wait(res_bar(ew, i & rmask), (i / rdepth) & 1);
        float rr[32];
        slab_load_f32(res_slab(i), lane, rr);
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = fmaf(p.alpha, t[j], rr[j]);
        __syncwarp();
        if (lane == 0 && !res_direct && i + rdepth < my_chunks) issue_res(i + rdepth);
      } else if (p.alpha != 1.0f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] *= p.alpha;
      }
      if (p.round_out) {
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = round_tf32(t[j]);
      }
      if constexpr (kLN) {
        // chunk statistics (two-pass inside the chunk), merged into the running row statistics (Chan et al.)
        const int nc = min(32, p.N - c0);
        if (nc < 32) {                   // accumulator columns beyond the UMMA N are stale tensor memory: zero the tail
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = (j < nc) ? t[j] : 0.f;
        }
        float cm, cq;
        chunk_stats(t, nc, cm, cq);
        stats_merge(cnt, mean, m2, static_cast<float>(nc), cm, cq);
        slab_store_f32(wstage + c * kSlabBytes, lane, t);      // row tile stays resident for the normalise passes
        if (p.ln_mode == 1 && !batch) {                        // x itself is an output: store it now
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(&tmOutF, smem_u32(wstage + c * kSlabBytes), c0, row0); bulk_commit(); }
        }
      } else {
        const int buf = i & (nbuf - 1);
        uint8_t* sf = wstage + buf * kSlabBytes;
        uint8_t* sa = slabA + buf * kASlab;
        if (p.has_out_f32) slab_store_f32(sf, lane, t);
        if (p.has_out_act) {
          if (kSplit && p.act_f16) slab_store_act<__half>(sa, lane, t);
          else slab_store_act<T>(sa, lane, t);
        }
        if (!batch) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.has_out_f32) tma_store_2d(&tmOutF, smem_u32(sf), out_col0 + c0, row0);
            if (p.has_out_act) tma_store_2d(&tmOutA, smem_u32(sa), out_col0 + c0, row0);
            if (p.has_out_act2) tma_store_2d(&tmOutA2, smem_u32(slabA2 + buf * kASlab), out_col0 + c0, row0);
            bulk_commit();
          }
        }
      }
    }
    if (batch && (!kLN || p.ln_mode == 1)) {                 // one fence, then every slab of this warp in one burst
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        for (int c = half; c < n_chunks; c += kHalves) {
          if (kLN || p.has_out_f32) tma_store_2d(&tmOutF, smem_u32(wstage + c * kSlabBytes), out_col0 + c * 32, row0);
          if (!kLN && p.has_out_act) tma_store_2d(&tmOutA, smem_u32(slabA + c * kASlab), out_col0 + c * 32, row0);
          if (!kLN && p.has_out_act2) tma_store_2d(&tmOutA2, smem_u32(slabA2 + c * kASlab), out_col0 + c * 32, row0);
        }
        bulk_commit();
      }
    }
    if (et == 0) stamp(p.dbg, 8);
    if constexpr (kLN) {
      const float inv_n = 1.0f / static_cast<float>(p.N);
      // merge the (count, mean, M2) of the four warps that share these 32 rows (fixed order -> all get identical results)
      auto merge_pair = [&](int stage, float& mean_, float& m2_, float cnt_) {
        float* mine = stat_x + (((stage * 4 + half) * 4 + q) * 32 + lane) * 3;
        mine[0] = cnt_; mine[1] = mean_; mine[2] = m2_;
        asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
        float ct = 0.f, mu = 0.f, qq = 0.f;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          const float* o = stat_x + (((stage * 4 + hh) * 4 + q) * 32 + lane) * 3;
          const float oc = o[0], om = o[1], o2 = o[2];
          if (oc > 0.f) {
            const float tot = ct + oc, w_ = __fdividef(oc, tot), dl = om - mu;
            mu = fmaf(dl, w_, mu);
            qq += o2 + dl * dl * ct * w_;
            ct = tot;
          }
        }
        mean_ = mu; m2_ = qq;
      };
      merge_pair(0, mean, m2, cnt);
      float rstd = rsqrtf(m2 * inv_n + p.ln_eps);
      const float* gfin = sg1; const float* bfin = sb1;
      bool affine = true;
      if (p.ln_mode == 2) {            // block norm in place (fp32 output), then the next module's LayerNorm on top of it
        float mean2 = 0.f, m22 = 0.f, cnt2 = 0.f;
        for (int c = half; c < n_chunks; c += kHalves) {
          const int c0 = c * 32, nc = min(32, p.N - c0);
          float t[32];
          slab_load_f32(wstage + c * kSlabBytes, lane, t);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 g4 = *reinterpret_cast<const float4*>(sg1 + c0 + 4 * j4), b4 = *reinterpret_cast<const float4*>(sb1 + c0 + 4 * j4);
            t[4 * j4] = (t[4 * j4] - mean) * rstd * g4.x + b4.x; t[4 * j4 + 1] = (t[4 * j4 + 1] - mean) * rstd * g4.y + b4.y;
            t[4 * j4 + 2] = (t[4 * j4 + 2] - mean) * rstd * g4.z + b4.z; t[4 * j4 + 3] = (t[4 * j4 + 3] - mean) * rstd * g4.w + b4.w;
          }
          float cm, cq;
          chunk_stats(t, nc, cm, cq);
          stats_merge(cnt2, mean2, m22, static_cast<float>(nc), cm, cq);
          slab_store_f32(wstage + c * kSlabBytes, lane, t);
          if (!batch) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&tmOutF, smem_u32(wstage + c * kSlabBytes), c0, row0); bulk_commit(); }
          }
        }
        if (batch) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            for (int c = half; c < n_chunks; c += kHalves) tma_store_2d(&tmOutF, smem_u32(wstage + c * kSlabBytes), c * 32, row0);
            bulk_commit();
          }
        }
        if (p.ln2_g != nullptr) merge_pair(1, mean2, m22, cnt2);
        mean = mean2; rstd = rsqrtf(m22 * inv_n + p.ln_eps);
        gfin = sg2; bfin = sb2;
        affine = p.ln2_g != nullptr;
      }
      const bool do_copy = p.copy_out != nullptr;
      T* copy_row = nullptr;
      if (do_copy) {
        const int m = row0 + lane;
        if (m < p.M) {
          const int seq = m / p.frames_per_seq, tt = m - seq * p.frames_per_seq;
          if (tt % p.copy_stride == 0)
            copy_row = reinterpret_cast<T*>(p.copy_out) + (static_cast<size_t>(seq) * p.frames_out_per_seq + tt / p.copy_stride) * p.N;
        }
      }
      if (p.has_ln_out || do_copy) {
        for (int c = half, i = 0; c < n_chunks; c += kHalves, ++i) {
          const int c0 = c * 32, nc = min(32, p.N - c0);
          float t[32];
          slab_load_f32(wstage + c * kSlabBytes, lane, t);
          if (copy_row != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nc) copy_row[c0 + j] = Tr::to(t[j]);
          }
          if (p.has_ln_out) {
            if (affine) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 g4 = *reinterpret_cast<const float4*>(gfin + c0 + 4 * j4), b4 = *reinterpret_cast<const float4*>(bfin + c0 + 4 * j4);
                t[4 * j4] = (t[4 * j4] - mean) * rstd * g4.x + b4.x; t[4 * j4 + 1] = (t[4 * j4 + 1] - mean) * rstd * g4.y + b4.y;
                t[4 * j4 + 2] = (t[4 * j4 + 2] - mean) * rstd * g4.z + b4.z; t[4 * j4 + 3] = (t[4 * j4 + 3] - mean) * rstd * g4.w + b4.w;
              }
            }
            if (i >= nbuf) { if (lane == 0) { if (nbuf > 1) bulk_wait_read(nbuf - 1); else bulk_wait_read0(); } __syncwarp(); }
            uint8_t* sa = slabA + (i & (nbuf - 1)) * kASlab;   // normally one slab per owned chunk: no waits
            slab_store_act<T>(sa, lane, t);
            if (!batch) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) { tma_store_2d(&tmLn, smem_u32(sa), c0, row0); bulk_commit(); }
            }
          }
        }
        if (batch && p.has_ln_out) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            for (int c = half, i = 0; c < n_chunks; c += kHalves, ++i) tma_store_2d(&tmLn, smem_u32(slabA + i * kASlab), c * 32, row0);
            bulk_commit();
          }
        }
      }
    }
    if (et == 0) stamp(p.dbg, 9);
    if (lane == 0) bulk_wait_read0();    // the slabs have been read (shared memory may be released); the writes themselves
                                         // complete before the grid does
    if (et == 0) stamp(p.dbg, 10);
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  if (threadIdx.x == 32) stamp(p.dbg, 11);
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
// Output tiles are stored in 32-column slabs, so tile boundaries inside a row must be multiples of 32.
static int pick_block_n(int N, int cap = 256) {
  if (N <= cap) return round_up(N, 16);
  const int tiles = cdiv(N, cap);
  return std::min(cap, round_up(cdiv(N, tiles), 32));
}

template <typename T, bool kLN>
static int launch_gemm_t(int precision, const GemmArgs& a, cudaStream_t stream) {
  using Tr = ActTraits<T>;
  EC_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "empty GEMM");
  GemmDev p{};
  p.M = a.M; p.N = a.N; p.K = a.K;
  int tiles_n;
  const int out_cols = a.glu_nb > 0 ? a.glu_channels : a.N;
  if (a.glu_nb > 0) {
    EC_REQUIRE(a.glu_nb % 8 == 0 && 2 * a.glu_nb <= 256 && (2 * a.glu_nb) % 16 == 0, "invalid GLU tile width");
    p.block_n = 2 * a.glu_nb;
    EC_REQUIRE(a.N % p.block_n == 0, "GLU weight rows must be a whole number of tiles");
    tiles_n = a.N / p.block_n;
    EC_REQUIRE(tiles_n == 1 || a.glu_nb % 32 == 0, "multi-tile GLU needs a tile width that is a multiple of 32");
  } else {
    // split mode stages two W tiles per k-block: narrower plain tiles keep a 3-4 deep ring (the fused-LayerNorm variant needs the
    // whole row in one tile and runs 2 stages for wide rows)
    // Split mode stages two W tiles per k-block.  The N tile is chosen by a small cost model fitted to the tile study
    // (tools/gemm_tiles.py, profiles/r2/gemm_tiles_bf16x2.txt): time ~ waves x (fixed + per-column epilogue), one CTA per SM, with
    // wide tiles only for short contractions (K <= 256: 2-deep ring) -- e.g. W1 16000 x 480 x 120: 4 x 128 (27.8 us) -> 2 x 256
    // (21.1 us); 8000 x 672 x 168: 3 x 224 leaves a near-empty second wave, 4 x 192 fills it.
    p.block_n = pick_block_n(a.N, 256);
    if (IsSplit<T>::value && !kLN) {
      const int cap = a.K > 256 ? 128 : 256;
      const size_t fixed_est = ((a.residual != nullptr || a.aux_mode != 0) ? 16 * kSlabBytes : 0) + kVecBytes + kNumBars * 8 + 16 + 1024;
      const int m_tiles = cdiv(a.M, kBlockM);
      double best = 1e30; int best_bn = 0;
      auto consider = [&](int bn) {
        if (bn > cap || bn < 16) return;
        if (2 * (static_cast<size_t>(kATileBytes) + 2 * static_cast<size_t>(bn) * 128) + fixed_est > 227 * 1024) return;
        const int waves = cdiv(m_tiles * cdiv(a.N, bn), 148);
        const double cost = waves * (4.0 + 0.004 * a.K + 0.02 * bn);
        if (cost < best - 1e-9 || (cost < best + 1e-9 && bn > best_bn)) { best = cost; best_bn = bn; }
      };
      consider(round_up(a.N, 16));                       // the whole row in one tile
      for (int bn = 64; bn <= 256; bn += 32) if (bn < a.N) consider(bn);
      if (best_bn == 0) best_bn = pick_block_n(a.N, 128);
      p.block_n = best_bn;
    }
    if (!kLN && g_block_n_override > 0) p.block_n = std::min(round_up(a.N, 16), g_block_n_override);   // tile-shape studies (tools/gemm_tiles.py)
    tiles_n = cdiv(a.N, p.block_n);
  }
  p.num_k_blocks = cdiv(a.K, Tr::kBlockK);
  p.has_res = (a.residual != nullptr || a.aux_mode != 0); p.has_out_f32 = a.out_f32 != nullptr; p.has_out_act = a.out_act != nullptr;
  p.has_out_act2 = a.out_act2 != nullptr; p.aux_mode = a.aux_mode;
  p.res_tx = (a.aux_mode != 0 && sizeof(T) == 2) ? kSlabBytes / 2 : kSlabBytes;
  const bool train_epi = a.drop_ctr != nullptr || a.out_act2 != nullptr || a.aux_mode != 0;
  if (train_epi) {
    EC_REQUIRE(a.glu_nb == 0 && (!kLN || (a.out_act2 == nullptr && a.aux_mode == 0)),
               "the Swish side output / Swish backward belong to the plain GEMM (the fused-LayerNorm variant takes the dropout site only)");
    EC_REQUIRE(a.N % 4 == 0, "dropout masks are drawn in groups of 4 consecutive elements: N must be a multiple of 4");
    EC_REQUIRE(a.aux_mode == 0 || (a.aux_mode == 1 && a.aux_act != nullptr && a.residual == nullptr), "aux_mode 1 needs the saved pre-activation and no residual");
    EC_REQUIRE((a.drop_site | a.drop_site2 | a.drop_site_aux) == 0 || (a.drop_ctr != nullptr && a.drop_p >= 0.f && a.drop_p < 1.f), "dropout site without a counter / bad p");
    EC_REQUIRE(a.out_act2 == nullptr || (a.out_act != nullptr && a.act == GEMM_ACT_NONE && !a.act_f16), "the Swish side output comes with the pre-activation output");
    p.drop_ctr = a.drop_ctr; p.drop_keep16 = keep16_of(a.drop_p); p.drop_inv_keep = 65536.f / static_cast<float>(p.drop_keep16);
    p.drop_site = a.drop_site; p.drop_site2 = a.drop_site2; p.drop_site_aux = a.drop_site_aux;
  }
  const int stage_bytes = kATileBytes + (IsSplit<T>::value ? 2 : 1) * p.block_n * 128;
  const int n_chunks = cdiv(std::min(a.glu_nb > 0 ? a.glu_nb : p.block_n, out_cols), 32);
  const int a_slab = sizeof(T) == 4 ? kSlabBytes : kSlabBytes / 2;
  static const int epi_batch_env = [] { const char* e = getenv("EFFCONF_EPI_BATCH"); return (e != nullptr && e[0] == '1') ? 1 : 0; }();
  const int per_chunk = (p.has_out_f32 ? kSlabBytes : 0) + (p.has_out_act ? a_slab : 0) + (p.has_out_act2 ? a_slab : 0);
  p.res_depth = 1;
  if (kLN) {
    // per row quarter: persistent x slabs (one per chunk) + ln_out slabs for each of its four warps (one per owned chunk when
    // that fits, else one); the residual ring (depth 1, prefetched during the mainloop) is dropped for wide rows, whose residual
    // then lands in the x slabs once the operand ring is drained
    const int own = cdiv(n_chunks, 4);
    EC_REQUIRE(own <= 2, "fused LayerNorm row too wide");
    int na = own;
    auto total = [&](int na_, int depth) {
      return 4 * (n_chunks * kSlabBytes + 4 * na_ * a_slab) + (p.has_res ? 16 * depth * kSlabBytes : 0) + kVecBytes + kNumBars * 8 + 16 + 1024;
    };
    p.res_depth = 1;
    if (total(na, 1) > 227 * 1024) p.res_depth = 0;
    if (total(na, p.res_depth) > 227 * 1024) na = 1;
    EC_REQUIRE(total(na, p.res_depth) <= 227 * 1024, "fused LayerNorm tile does not fit in shared memory");
    p.nbuf = na;
    p.warp_stage_bytes = n_chunks * kSlabBytes + 4 * na * a_slab;
    p.epi_batch = (epi_batch_env && na >= own) ? 1 : 0;
  } else {
    p.nbuf = 2;                                             // each of the 16 epilogue warps owns at most ceil(n_chunks / 4) chunks
    p.warp_stage_bytes = p.nbuf * per_chunk;
    p.epi_batch = 0;
    p.res_depth = 1;
  }
  int fixed = (p.has_res ? 16 * p.res_depth * kSlabBytes : 0) + kVecBytes + kNumBars * 8 + 16 + 1024;
  // one CTA per SM (thread count): a deep ring hides the TMA->MMA->refill round trip
  const int budget = 224 * 1024;
  if (kLN && p.has_res && p.res_depth == 1 && (budget - fixed) / stage_bytes < 2) {   // wide split-mode rows: residual lands in the x slabs
    p.res_depth = 0;
    fixed = kVecBytes + kNumBars * 8 + 16 + 1024;
  }
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
  p.round_out = a.round_out;
  p.act_f16 = (IsSplit<T>::value && a.act_f16) ? 1 : 0;
  EC_REQUIRE(!a.act_f16 || (IsSplit<T>::value && !kLN), "act_f16 is an option of the plain split-mode GEMM");
  p.dbg = g_timeline_enabled;
  EC_REQUIRE(a.out_f32 != nullptr || a.out_act != nullptr, "GEMM needs at least one output");
  EC_REQUIRE(a.glu_nb == 0 || a.bias != nullptr, "GLU GEMM needs a bias");
  EC_REQUIRE(a.residual == nullptr || a.ld_res == out_cols, "residual must be dense [M, N]");
  if (kLN) {
    EC_REQUIRE(a.ln_mode == 1 || a.ln_mode == 2, "bad LayerNorm mode");
    EC_REQUIRE(tiles_n == 1 && a.glu_nb == 0 && a.act == GEMM_ACT_NONE && a.N <= 256, "fused LayerNorm needs the whole row in one plain tile (N <= 256)");
    EC_REQUIRE(a.out_f32 != nullptr && a.ld_out == a.N && a.out_act == nullptr, "fused LayerNorm writes the fp32 output and ln_out only");
    EC_REQUIRE(a.ln1_g != nullptr && a.ln1_b != nullptr, "missing LayerNorm parameters");
    p.ln_mode = a.ln_mode; p.ln1_g = a.ln1_g; p.ln1_b = a.ln1_b; p.ln2_g = a.ln2_g; p.ln2_b = a.ln2_b; p.ln_eps = a.ln_eps;
    p.has_ln_out = a.ln_out != nullptr;
    p.copy_out = a.copy_out; p.copy_stride = a.copy_stride > 0 ? a.copy_stride : 1;
    p.frames_per_seq = a.frames_per_seq > 0 ? a.frames_per_seq : a.M; p.frames_out_per_seq = a.frames_out_per_seq;
    EC_REQUIRE(a.copy_out == nullptr || a.ln_mode == 1, "the strided copy is only available with a single LayerNorm");
  }

  CUtensorMap tmA, tmB, tmB2, tmRes, tmOutF, tmOutA, tmLn, tmOutA2;
  EC_TRY(make_operand_map(&tmA, precision, a.A, a.M, a.K, kBlockM));
  EC_TRY(make_operand_map(&tmB, precision, a.W, a.N, a.K, p.block_n));
  tmB2 = tmB;
  if (IsSplit<T>::value) {            // plane 1 of the [2, N, K] weight operand: (lo, hi)
    const uint8_t* twin = reinterpret_cast<const uint8_t*>(a.W) + (a.w_twin_bytes != 0 ? a.w_twin_bytes : static_cast<size_t>(a.N) * a.K * 4);
    EC_TRY(make_operand_map(&tmB2, precision, twin, a.N, a.K, p.block_n));
  }
  const bool act_f32 = sizeof(T) == 4;
  tmRes = tmA; tmOutF = tmA; tmOutA = tmA; tmLn = tmA; tmOutA2 = tmA;       // placeholders for unused maps
  if (a.residual != nullptr) EC_TRY(make_slab_map(&tmRes, true, a.residual, a.M, out_cols, a.ld_res));
  if (a.aux_mode != 0) EC_TRY(make_slab_map(&tmRes, act_f32, a.aux_act, a.M, out_cols, out_cols));
  if (a.out_act2 != nullptr) EC_TRY(make_slab_map(&tmOutA2, act_f32, a.out_act2, a.M, out_cols, a.ld_act2));
  if (a.out_f32 != nullptr) EC_TRY(make_slab_map(&tmOutF, true, a.out_f32, a.M, out_cols, a.ld_out));
  if (a.out_act != nullptr) EC_TRY(make_slab_map(&tmOutA, act_f32 && !p.act_f16, a.out_act, a.M, out_cols, a.ld_act));
  if (kLN && a.ln_out != nullptr) EC_TRY(make_slab_map(&tmLn, act_f32, a.ln_out, a.M, a.N, a.N));

  if (!kLN) {
    if (static_cast<size_t>(16) * p.nbuf * per_chunk + fixed > 227 * 1024) { p.nbuf = 1; p.warp_stage_bytes = per_chunk; }
    p.epi_batch = (epi_batch_env && cdiv(n_chunks, 4) <= p.nbuf) ? 1 : 0;   // staging (16 warps) aliases the operand ring
  }
  size_t pipe_bytes = std::max(static_cast<size_t>(stages) * stage_bytes, static_cast<size_t>(kLN ? 4 : 16) * p.warp_stage_bytes);
  p.pipe_bytes = static_cast<int>(pipe_bytes);
  const size_t smem = pipe_bytes + fixed;
  EC_REQUIRE(smem <= 227 * 1024, "GEMM tile does not fit in shared memory");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel<T, kLN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  EC_CUDA(attr_err);
  dim3 grid(cdiv(a.M, kBlockM), tiles_n);
  EC_TRY(launch_pdl(gemm_tc_kernel<T, kLN>, grid, dim3(576), smem, stream, tmA, tmB, tmB2, tmRes, tmOutF, tmOutA, tmLn, tmOutA2, p));
  return EC_OK;
}

// Debug hook: enable/disable the timeline stamps and read the 12 stamps (SM cycles) of the last GEMM's CTA (0,0):
// 0 start, 1 setup done, 2 dependency wait done, 3 first TMA issued, 4 first stage landed, 5 last MMA committed,
// 6 epilogue ready, 7 accumulator complete, 8 chunk loop done, 9 LayerNorm passes done, 10 bulk stores complete, 11 end.
int gemm_timeline(int enable, unsigned long long* out12) {
  g_timeline_enabled = enable;
  if (out12 != nullptr) EC_CUDA(cudaMemcpyFromSymbol(out12, g_gemm_timeline, 12 * sizeof(unsigned long long)));
  return EC_OK;
}

int gemm_block_n_override(int block_n) { g_block_n_override = block_n; return EC_OK; }

int launch_gemm(int precision, const GemmArgs& a, cudaStream_t stream) {
  if (precision == EC_PREC_TF32) return a.ln_mode ? launch_gemm_t<float, true>(precision, a, stream) : launch_gemm_t<float, false>(precision, a, stream);
  if (precision == EC_PREC_BF16)
    return a.ln_mode ? launch_gemm_t<__nv_bfloat16, true>(precision, a, stream) : launch_gemm_t<__nv_bfloat16, false>(precision, a, stream);
  if (precision == EC_PREC_BF16X2)
    return a.ln_mode ? launch_gemm_t<SplitBf16, true>(precision, a, stream) : launch_gemm_t<SplitBf16, false>(precision, a, stream);
  EC_FAIL("unknown precision");
}

}  // namespace ec
