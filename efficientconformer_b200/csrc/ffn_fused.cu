// Fused half-step feed-forward module (fast / bf16 mode):
//     y = x + 0.5 * (Swish(LN(x) W1^T + b1) W2^T + b2)           followed by the LayerNorm(s) the next module needs
// (reference models/modules.py:367-398 FeedForwardModule, models/blocks.py:123-150 residual half steps and block norm).
//
// One thread-block CLUSTER owns a 128-row tile; CTA r of the cluster owns the hidden units [r*Hs, (r+1)*Hs).  Per CTA the
// hidden slice is walked in chunks of 64 units (one 128-byte bf16 K panel of the second product):
//     GEMM1_k : Hk[128 x 64]  = X[128 x D] . W1_k^T          tcgen05.mma kind::f16, accumulator in TMEM (double buffered)
//     epilogue: Hk -> +b1 -> Swish -> bf16 -> shared memory in the 128B-swizzled K-major layout an A operand needs
//     GEMM2_k : Y[128 x D]   += Hk . W2_k^T                   accumulator in TMEM, never leaves the SM
// so the (M x 4D) hidden activation never touches L2 / HBM.  Warp roles (576 threads): warp 0 TMA producer (X once, W1 / W2
// chunk rings), warp 1 MMA issuer (software-pipelined: GEMM1_{k+1} is issued before GEMM2_k), warps 2-17 epilogue
// (thread = row; the four warps of a TMEM lane quarter take the four 16-column quarters of a chunk).
//
// After the last chunk the partial Y tiles of the cluster are reduced through distributed shared memory: every CTA pushes
// 32-column chunks of its partial to the CTA that owns those columns (st.shared::cluster), the owner adds them in a fixed
// order, applies bias / 0.5 / residual, and the per-row LayerNorm statistics of the column slices are exchanged the same way
// (count-weighted Chan merge in a fixed order -> every CTA derives bit-identical row statistics).  Outputs leave through
// swizzled shared-memory slabs and TMA bulk tensor stores, as in gemm_tc.cu.
#include "ec_common.cuh"
#include "ec_tma.cuh"
#include <algorithm>
#include <string>

namespace ec {

constexpr int kHC = 64;                       // hidden units per chunk
constexpr int kFfnEpiThreads = 512;             // 16 epilogue warps
constexpr int kFfnThreads = 64 + kFfnEpiThreads;
constexpr int kFfnTmemH = 256;                // TMEM column of the first H buffer (Y occupies [0, 256))
constexpr int kFfnB1Floats = 1024;             // largest hidden slice of one CTA
constexpr int kFfnVecFloats = kFfnB1Floats + 5 * 288;  // b1 slice | b2 | ln1_g | ln1_b | ln2_g | ln2_b
constexpr int kFfnMaxRing = 4;
constexpr int kFfnBars = 1 + 2 * kFfnMaxRing + 2 + 4 + 1 + 16;
constexpr int kFfnFixedBytes = kFfnVecFloats * 4 + kFfnBars * 8 + 16;

struct FfnDev {
  int M, D, Hs, n_hc, CS;
  int kp1, ksteps1;                  // 128-byte K panels / UMMA K steps over the model dim
  int n2;                            // UMMA N of the second product (D rounded up to 16)
  int ns, nsh;                       // ring depths: weight stages {W1_i, W2_{i-1}}, H operand tiles
  int w1_slot, w2_slot, w_stage;
  int off_w, off_h;                  // mainloop layout (bytes from the 1024-aligned base); X tile at 0
  int n_chunks, chunks_per;          // 32-column output chunks, chunks owned per CTA
  int off_xt, off_ln, off_stats;     // final-phase layout (aliases the drained operand region); receive buffers at 0
  int off_fixed;                     // vectors + barriers behind max(mainloop, final) bytes
  const float *b1, *b2;
  int ln_mode; const float *ln1_g, *ln1_b, *ln2_g, *ln2_b; float ln_eps; int has_ln_out;
  int dbg;
};

// Optional in-kernel timeline (SM clock stamps of CTA 0), enabled through ec_debug_ffn_timeline.
__device__ unsigned long long g_ffn_timeline[288];   // [0,18) phase stamps; 32 + 16k + i: per-chunk stamps (k < 16)
static int g_ffn_timeline_enabled = 0;
__device__ __forceinline__ void fstamp(int enabled, int slot) {
  if (enabled && blockIdx.x == 0) g_ffn_timeline[slot] = clock64();
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kFfnEpiThreads) : "memory"); }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmRes,
                 const __grid_constant__ CUtensorMap tmOutF, const __grid_constant__ CUtensorMap tmLn, const FfnDev p) {
  using T = __nv_bfloat16;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = (blockIdx.x / p.CS) * kBlockM;
  const int h0 = static_cast<int>(rank) * p.Hs;                 // first hidden unit of this CTA
  const int hc_last = p.Hs - (p.n_hc - 1) * kHC;

  float* vecs = reinterpret_cast<float*>(base_ptr + p.off_fixed);
  const uint32_t bars_addr = base + p.off_fixed + kFfnVecFloats * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + p.off_fixed + kFfnVecFloats * 4);
  // barrier map
  const uint32_t x_full = bars_addr;
  // weight stage i = {W1 chunk i, W2 chunk i-1}: what MMA iteration i consumes (GEMM1_i, then GEMM2_{i-1}); stages 0 .. n_hc
  auto w_full = [&](int s) { return bars_addr + 8u * (1 + s); };
  auto w_empty = [&](int s) { return bars_addr + 8u * (1 + kFfnMaxRing + s); };
  constexpr int kB0 = 1 + 2 * kFfnMaxRing;
  auto ht_full = [&](int s) { return bars_addr + 8u * (kB0 + s); };          // H accumulator (TMEM) ready; "drained" is implied by
                                                                             // hs_full of the same chunk (same warps, program order)
  auto hs_full = [&](int s) { return bars_addr + 8u * (kB0 + 2 + s); };          // H operand tile (smem) written / consumed
  auto hs_empty = [&](int s) { return bars_addr + 8u * (kB0 + 4 + s); };
  const uint32_t y_full = bars_addr + 8u * (kB0 + 6);
  auto res_bar = [&](int ew_) { return bars_addr + 8u * (kB0 + 7 + ew_); };
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(bars + kFfnBars);

  if (threadIdx.x == 0) fstamp(p.dbg, 0);
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    mbar_init(x_full, 1);
    for (int s = 0; s < kFfnMaxRing; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(ht_full(s), 1); mbar_init(hs_full(s), 16); mbar_init(hs_empty(s), 1); }
    mbar_init(y_full, 1);
    for (int i = 0; i < 16; ++i) mbar_init(res_bar(i), 1);
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (threadIdx.x == 0) fstamp(p.dbg, 1);

  // Producer and MMA loops run warp-uniform (all 32 lanes wait on the barriers); only the TMA / tcgen05 instructions sit under
  // elect_one().  A lane-0 branch instead makes the compiler wrap every uniform-datapath instruction in an ELECT/BRA loop and
  // serialises the whole role behind one thread's scalar latency -- measured 2x slower than the tensor pipe for these tile sizes.
  if (warp_idx == 0) {
    // ---------------- TMA producer ----------------
    int sw = 0, phw = 1;                         // ring position advances by increments (no integer division on this path)
    auto load_stage = [&](int i) {   // W1 chunk i (i < n_hc) and W2 chunk i-1 (i >= 1) on one barrier
      mbar_wait(w_empty(sw), phw);
      if (elect_one()) {
        const uint32_t bytes = (i < p.n_hc ? p.w1_slot : 0) + (i >= 1 ? p.w2_slot : 0);
        mbar_arrive_expect_tx(w_full(sw), bytes);
        if (i < 16) fstamp(p.dbg, 32 + 16 * i + 7);
        const uint32_t dst = base + p.off_w + sw * p.w_stage;
        if (i < p.n_hc)
          for (int pn = 0; pn < p.kp1; ++pn) tma_load_2d(dst + pn * (kHC * 128), &tmW1, w_full(sw), pn * 64, h0 + i * kHC);
        if (i >= 1) tma_load_2d(dst + p.w1_slot, &tmW2, w_full(sw), h0 + (i - 1) * kHC, 0);
      }
      __syncwarp();
      if (++sw == p.ns) { sw = 0; phw ^= 1; }
    };
    load_stage(0);              // weights do not depend on the previous kernel: in flight before the dependency wait
    if (p.ns > 1) load_stage(1);
    grid_dependency_wait();
    grid_launch_dependents();
    if (elect_one()) {
      fstamp(p.dbg, 2);
      mbar_arrive_expect_tx(x_full, static_cast<uint32_t>(p.kp1 * kATileBytes));
      for (int pn = 0; pn < p.kp1; ++pn) tma_load_2d(base + pn * kATileBytes, &tmX, x_full, pn * 64, m0);
    }
    __syncwarp();
    for (int i = (p.ns > 1 ? 2 : 1); i <= p.n_hc; ++i) load_stage(i);
  } else if (warp_idx == 1) {
    // ---------------- MMA issuer ----------------
    grid_launch_dependents();
    const uint32_t idesc2 = make_idesc(1u, kBlockM, p.n2);
    // descriptors advance by adding (bytes >> 4) to the start-address field
    const uint64_t dx0 = make_smem_desc_sw128(base);
    int sw = 0, phw = 0, sh = 0, phh = 0;
    mbar_wait(x_full, 0);
    if (lane == 0) fstamp(p.dbg, 3);
    for (int k = 0; k <= p.n_hc; ++k) {
      mbar_wait(w_full(sw), phw);                        // W1_k and W2_{k-1}
      const uint32_t wst = base + p.off_w + sw * p.w_stage;
      if (k < p.n_hc) {
        // the H accumulator buffer (k & 1) was drained by the epilogue of chunk k-2, whose hs_full this warp has already seen
        const int hb = k & 1, hc = (k == p.n_hc - 1) ? hc_last : kHC;
        tc_fence_after();
        const uint32_t idesc1 = make_idesc(1u, kBlockM, hc);
        const uint64_t dw0 = make_smem_desc_sw128(wst);
        const uint32_t dacc = tmem_base + kFfnTmemH + hb * kHC;
        if (elect_one()) {
          if (k == 0) fstamp(p.dbg, 4);
          if (k < 16) fstamp(p.dbg, 32 + 16 * k + 0);
          int ks = 0;
          for (int pn = 0; pn < p.kp1; ++pn) {
            const uint64_t da = dx0 + static_cast<uint64_t>(pn * (kATileBytes >> 4)), db = dw0 + static_cast<uint64_t>(pn * ((kHC * 128) >> 4));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk, ++ks)
              if (ks < p.ksteps1) tc_mma<false>(dacc, da + 2 * kk, db + 2 * kk, idesc1, ks != 0 ? 1u : 0u);
          }
          if (k < 16) fstamp(p.dbg, 32 + 16 * k + 1);
          tc_commit(ht_full(hb));
          if (k == 0) tc_commit(w_empty(sw));            // stage 0 holds W1_0 only
          if (k < 16) fstamp(p.dbg, 32 + 16 * k + 2);
        }
        __syncwarp();
      }
      if (k >= 1) {
        const int j = k - 1, hc = (j == p.n_hc - 1) ? hc_last : kHC;
        if (lane == 0 && j < 16) fstamp(p.dbg, 32 + 16 * j + 3);
        mbar_wait(hs_full(sh), phh);
        tc_fence_after();
        const uint64_t da = make_smem_desc_sw128(base + p.off_h + sh * kATileBytes);
        const uint64_t db = make_smem_desc_sw128(wst + p.w1_slot);
        if (elect_one()) {
          if (j < 16) fstamp(p.dbg, 32 + 16 * j + 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (ks * 16 < hc) tc_mma<false>(tmem_base, da + 2 * ks, db + 2 * ks, idesc2, (j | ks) != 0 ? 1u : 0u);
          if (j < 16) fstamp(p.dbg, 32 + 16 * j + 5);
          tc_commit(w_empty(sw));                        // frees W1_k and W2_{k-1} when everything issued so far has retired
          tc_commit(hs_empty(sh));
          if (j == p.n_hc - 1) tc_commit(y_full);
          if (j < 16) fstamp(p.dbg, 32 + 16 * j + 6);
        }
        __syncwarp();
        if (++sh == p.nsh) { sh = 0; phh ^= 1; }
      }
      if (++sw == p.ns) { sw = 0; phw ^= 1; }
    }
  }

  // ---------------- epilogue warps: thread = row; (q, sub) = (TMEM lane quarter, 16-column quarter of a chunk) ----------------
  // 16 warps: the per-row work is a dependent chain, so the epilogue is latency bound per warp; four warps per lane quarter
  // quarter that chain (and the final phase gives every warp at most one 32-column output chunk).
  const int q = warp_idx & 3, ew = warp_idx - 2, sub = ew >> 2;
  float* sb1 = vecs;                                   // b1 slice of this CTA, zero padded
  float *sb2 = vecs + kFfnB1Floats, *sg1 = sb2 + 288, *sbb1 = sg1 + 288, *sg2 = sbb1 + 288, *sbb2 = sg2 + 288;
  const int row0 = m0 + q * 32;
  const bool dbt = threadIdx.x == 64;                  // the thread whose stamps are reported (warp 2: q = 2, sub = 0)
  const int CS = p.CS, cper = p.chunks_per;
  const int my_first = static_cast<int>(rank) * cper;
  const int owned = max(0, min(cper, p.n_chunks - my_first));
  const bool mine = warp_idx >= 2 && sub < owned;      // this warp's output chunk: local index sub (at most one per warp)
  const int c0 = (my_first + sub) * 32, nc = min(32, p.D - c0);
  if (warp_idx >= 2) {
    const int et = ew * 32 + lane;
    for (int i = et; i < kFfnB1Floats; i += kFfnEpiThreads) sb1[i] = i < p.Hs ? __ldg(p.b1 + h0 + i) : 0.f;
    for (int i = et; i < 288; i += kFfnEpiThreads) {
      const bool ok = i < p.D;
      sb2[i] = ok ? __ldg(p.b2 + i) : 0.f;
      sg1[i] = ok ? __ldg(p.ln1_g + i) : 0.f; sbb1[i] = ok ? __ldg(p.ln1_b + i) : 0.f;
      sg2[i] = (ok && p.ln2_g != nullptr) ? __ldg(p.ln2_g + i) : 0.f; sbb2[i] = (ok && p.ln2_g != nullptr) ? __ldg(p.ln2_b + i) : 0.f;
    }
    epi_sync();
    grid_dependency_wait();
    const uint32_t tH = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kFfnTmemH + sub * 16;
    int hs = 0, hs_par = 1;                              // H operand tile ring: slot and the parity of its "consumed" barrier
    for (int k = 0; k < p.n_hc; ++k) {
      const int hb = k & 1, hc = (k == p.n_hc - 1) ? hc_last : kHC;
      const bool active = sub * 16 < hc;
      uint32_t v[16];
      mbar_wait(ht_full(hb), (k >> 1) & 1);
      tc_fence_after();
      if (k == 0 && dbt) fstamp(p.dbg, 5);
      if (dbt && k < 16) fstamp(p.dbg, 32 + 16 * k + 8);
      if (active) { tmem_ld_32x16(tH + hb * kHC, v); tmem_ld_wait(); }
      if (dbt && k < 16) fstamp(p.dbg, 32 + 16 * k + 9);
      tc_fence_before();                                 // ordered before this warp's hs_full arrive below, which also tells the
                                                         // MMA warp that accumulator buffer hb is drained
      uint4 pk[2];
      if (active) {
        const float* bk = sb1 + k * kHC + sub * 16;
#pragma unroll
        for (int j8 = 0; j8 < 2; ++j8) {
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) t[j] = swish_fn<T>(__uint_as_float(v[8 * j8 + j]) + bk[8 * j8 + j]);
          __nv_bfloat162 h0_ = __floats2bfloat162_rn(t[0], t[1]), h1_ = __floats2bfloat162_rn(t[2], t[3]);
          __nv_bfloat162 h2_ = __floats2bfloat162_rn(t[4], t[5]), h3_ = __floats2bfloat162_rn(t[6], t[7]);
          pk[j8].x = *reinterpret_cast<uint32_t*>(&h0_); pk[j8].y = *reinterpret_cast<uint32_t*>(&h1_);
          pk[j8].z = *reinterpret_cast<uint32_t*>(&h2_); pk[j8].w = *reinterpret_cast<uint32_t*>(&h3_);
        }
      }
      if (dbt && k < 16) fstamp(p.dbg, 32 + 16 * k + 10);
      mbar_wait(hs_empty(hs), hs_par);                   // GEMM2 of the chunk that used this tile has retired (free on first use)
      if (dbt && k < 16) fstamp(p.dbg, 32 + 16 * k + 11);
      if (active) {
        const int r = q * 32 + lane;
        uint8_t* tile = base_ptr + p.off_h + hs * kATileBytes + r * 128;
#pragma unroll
        for (int j8 = 0; j8 < 2; ++j8) *reinterpret_cast<uint4*>(tile + (((sub * 2 + j8) ^ (r & 7)) << 4)) = pk[j8];
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(hs_full(hs));
      if (dbt) fstamp(p.dbg, k == 0 ? 6 : 7);
      if (dbt && k < 16) fstamp(p.dbg, 32 + 16 * k + 12);
      if (++hs == p.nsh) { hs = 0; hs_par ^= 1; }
    }
    mbar_wait(y_full, 0);                                // all MMAs of this CTA have retired: its operand region is free
    tc_fence_after();
    if (dbt) fstamp(p.dbg, 8);
  }

  // Exchange points: a cluster barrier when the row tile is split over several CTAs, else a barrier of the epilogue warps.
  auto sync_point = [&]() {
    if (CS > 1) cluster_sync_all();
    else if (warp_idx >= 2) epi_sync();
  };
  // ---- B0 (clusters only): every CTA has retired its MMAs -> peers may write into the (aliased) receive buffers ----
  if (CS > 1) cluster_sync_all();
  if (dbt) fstamp(p.dbg, 9);

  uint8_t* xt_slab = base_ptr + p.off_xt + (sub * 4 + q) * kSlabBytes;
  uint8_t* ln_slab = base_ptr + p.off_ln + (sub * 4 + q) * (kSlabBytes / 2);
  const int n_src = 4 * CS;                            // statistics sources: (cta, sub)
  float2* stats = reinterpret_cast<float2*>(base_ptr + p.off_stats);       // [stage 2][source 4*CS][row 128]
  const uint32_t stats_addr = base + p.off_stats;
  // columns owned by (cta c, sub s): fixed function of the shape -> the merge below needs no transmitted counts
  auto cnt_of = [&](int c, int s_) {
    const int own_c = max(0, min(cper, p.n_chunks - c * cper));
    return s_ < own_c ? min(32, p.D - (c * cper + s_) * 32) : 0;
  };
  auto push_stats = [&](int stage, float mean_, float m2_) {
    const uint32_t a = stats_addr + ((stage * n_src + static_cast<int>(rank) * 4 + sub) * 128 + q * 32 + lane) * 8;
    if (CS == 1) *reinterpret_cast<float2*>(base_ptr + (a - base)) = make_float2(mean_, m2_);
    else for (int dst = 0; dst < CS; ++dst) st_cluster_v2(mapa_shared(a, dst), mean_, m2_);
  };
  auto merge_stats = [&](int stage, float& mean_, float& rstd_) {
    float cnt = 0.f, mu = 0.f, m2_ = 0.f;
    for (int src = 0; src < n_src; ++src) {
      const float cs = static_cast<float>(cnt_of(src >> 2, src & 3));
      if (cs == 0.f) continue;
      const float2 s_ = stats[(stage * n_src + src) * 128 + q * 32 + lane];
      const float tot = cnt + cs, w_ = __fdividef(cs, tot), dl = s_.x - mu;
      mu = fmaf(dl, w_, mu);
      m2_ += s_.y + dl * dl * cnt * w_;
      cnt = tot;
    }
    mean_ = mu;
    rstd_ = rsqrtf(m2_ / static_cast<float>(p.D) + p.ln_eps);
  };

  const uint32_t tY = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  if (warp_idx >= 2) {
    // residual slab of the owned chunk lands directly in the x tile
    if (mine && lane == 0) {
      mbar_arrive_expect_tx(res_bar(ew), kSlabBytes);
      tma_load_2d(smem_u32(xt_slab), &tmRes, res_bar(ew), c0, row0);
    }
    if (CS > 1) {
      // ---- phase A: push the 32-column chunks of this CTA's partial Y to their owners ----
      for (int j = sub; j < p.n_chunks; j += 4) {
        uint32_t v[32];
        tmem_ld_32x32(tY + j * 32, v);
        tmem_ld_wait();
        const int o = j / cper, l = j - o * cper;
        const uint32_t dst = mapa_shared(base + ((static_cast<int>(rank) * cper + l) * 4 + q) * kSlabBytes, o);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          st_cluster_v4(dst + slab_f32_off(lane, j4), __uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                        __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3]));
      }
      tc_fence_before();
    }
    if (dbt) fstamp(p.dbg, 10);
  }
  if (CS > 1) cluster_sync_all();      // ---- B1 (clusters only): all partials have landed ----
  if (dbt) fstamp(p.dbg, 11);

  // ---- phase B: reduce, bias, half-step residual, first statistics.  The chunk stays in registers until it is stored. ----
  float t[32];
  if (mine) {
    if (CS == 1) {
      uint32_t v[32];
      tmem_ld_32x32(tY + sub * 32, v);
      tmem_ld_wait();
      tc_fence_before();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(v[j]);
    } else {
      float pp[32];
      slab_load_f32(base_ptr + (sub * 4 + q) * kSlabBytes, lane, t);
      for (int src = 1; src < CS; ++src) {
        slab_load_f32(base_ptr + ((src * cper + sub) * 4 + q) * kSlabBytes, lane, pp);
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] += pp[j];
      }
    }
    if (dbt) fstamp(p.dbg, 18);
    mbar_wait(res_bar(ew), 0);
    if (dbt) fstamp(p.dbg, 19);
    {
      float rr[32];
      slab_load_f32(xt_slab, lane, rr);
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = fmaf(0.5f, t[j] + sb2[c0 + j], rr[j]);
      if (nc < 32) {                   // accumulator columns beyond the row are stale tensor memory: zero them
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = (j < nc) ? t[j] : 0.f;
      }
    }
    float cm, cq;
    chunk_stats(t, nc, cm, cq);
    if (dbt) fstamp(p.dbg, 20);
    if (p.ln_mode == 1) {                // x itself is an output
      slab_store_f32(xt_slab, lane, t);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_2d(&tmOutF, smem_u32(xt_slab), c0, row0); bulk_commit(); }
    }
    push_stats(0, cm, cq);
    if (dbt) fstamp(p.dbg, 21);
  }
  if (dbt) fstamp(p.dbg, 12);
  sync_point();              // ---- B2: first statistics exchanged ----
  if (dbt) fstamp(p.dbg, 13);

  const bool second = p.ln_mode == 2 && p.ln2_g != nullptr;
  float mean = 0.f, rstd = 0.f;
  if (mine) {
    merge_stats(0, mean, rstd);
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = (t[j] - mean) * rstd * sg1[c0 + j] + sbb1[c0 + j];     // gamma = beta = 0 beyond the row
    if (p.ln_mode == 2) {                // block norm replaces x; it is the fp32 output
      slab_store_f32(xt_slab, lane, t);
      if (second) {
        float cm, cq;
        chunk_stats(t, nc, cm, cq);
        push_stats(1, cm, cq);
      }
    }
    const bool emit_ln = p.has_ln_out && !second;
    if (emit_ln) slab_store_act<T>(ln_slab, lane, t);
    if (p.ln_mode == 2 || emit_ln) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (p.ln_mode == 2) tma_store_2d(&tmOutF, smem_u32(xt_slab), c0, row0);
        if (emit_ln) tma_store_2d(&tmLn, smem_u32(ln_slab), c0, row0);
        bulk_commit();
      }
    }
  }
  if (dbt) fstamp(p.dbg, 14);
  if (second) {
    sync_point();            // ---- B3: statistics of the block-normalised rows exchanged ----
    if (mine) {
      merge_stats(1, mean, rstd);
      if (p.has_ln_out) {
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = (t[j] - mean) * rstd * sg2[c0 + j] + sbb2[c0 + j];
        slab_store_act<T>(ln_slab, lane, t);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { tma_store_2d(&tmLn, smem_u32(ln_slab), c0, row0); bulk_commit(); }
      }
    }
  }
  if (dbt) fstamp(p.dbg, 15);
  if (mine && lane == 0) bulk_wait_read0();            // the slabs have been read; the writes complete before the grid does
  if (dbt) fstamp(p.dbg, 16);
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) tmem_dealloc(tmem_base, 512);
  if (threadIdx.x == 32) fstamp(p.dbg, 17);
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
// A cluster split is usable when every CTA gets a whole number of 16-unit groups (<= 1024 units) and at most 4 output chunks.
static bool cluster_ok(int cs, int D, int hidden) {
  return hidden % cs == 0 && (hidden / cs) % 16 == 0 && hidden / cs <= kFfnB1Floats && cdiv(cdiv(D, 32), cs) <= 4;
}
static int pick_cluster(int m_tiles, int D, int hidden) {
  // widest usable split of the hidden dim (1, 2 or 4 CTAs per row tile) that still fits one wave of 148 SMs
  int best = 0;
  for (int cs = 1; cs <= 4; cs *= 2)
    if (cluster_ok(cs, D, hidden) && (best == 0 || m_tiles * cs <= 148)) best = cs;
  return best;
}

// Shared-memory plan for one launch; returns an empty string on success, else why the shape is not supported.
static std::string plan_ffn(int M, int D, int hidden, int cluster, FfnDev& p) {
  if (M <= 0 || D < 16 || D > 256 || D % 8 != 0) return "fused FFN needs 16 <= D <= 256, D % 8 == 0";
  if (hidden < 16 || hidden % 16 != 0) return "fused FFN needs hidden % 16 == 0";
  p.M = M; p.D = D;
  const int m_tiles = cdiv(M, kBlockM);
  p.CS = cluster > 0 ? cluster : pick_cluster(m_tiles, D, hidden);
  if (!(p.CS == 1 || p.CS == 2 || p.CS == 4) || !cluster_ok(p.CS, D, hidden))
    return "no usable cluster split (1, 2 or 4 CTAs; hidden/CTAs a multiple of 16 and <= 1024; <= 128 output columns per CTA)";
  p.Hs = hidden / p.CS;
  p.n_hc = cdiv(p.Hs, kHC);
  p.kp1 = cdiv(D, 64); p.ksteps1 = cdiv(D, 16);
  p.n2 = round_up(D, 16);
  p.w1_slot = p.kp1 * kHC * 128; p.w2_slot = p.n2 * 128;
  p.n_chunks = cdiv(D, 32); p.chunks_per = cdiv(p.n_chunks, p.CS);
  // final-phase layout: receive buffers (later ln_out slabs) | x tile | statistics
  const int recv_bytes = p.CS * p.chunks_per * 4 * kSlabBytes;
  p.off_xt = recv_bytes;
  p.off_ln = 0;                      // ln_out slabs reuse the receive buffers (dead once the partials are reduced)
  p.off_stats = p.off_xt + p.chunks_per * 4 * kSlabBytes;
  const int fin_bytes = p.off_stats + 2 * (4 * p.CS) * 128 * 8;
  // mainloop layout: X | W1 ring | W2 ring | H tiles; grow the rings while everything fits
  const int x_bytes = p.kp1 * kATileBytes;
  const int limit = 227 * 1024 - 1024 - kFfnFixedBytes;
  p.w_stage = p.w1_slot + p.w2_slot;
  auto main_bytes = [&](int n, int nh) { return x_bytes + n * p.w_stage + nh * kATileBytes; };
  int n = 2, nh = 1;
  if (main_bytes(n, nh) > limit || fin_bytes > limit) return "fused FFN tile does not fit in shared memory";
  // weight prefetch depth first (an L2 round trip is about two chunk times), then a second H tile
  while (n < kFfnMaxRing && n < p.n_hc + 1 && main_bytes(n + 1, nh) <= limit) ++n;
  if (main_bytes(n, 2) <= limit) nh = 2;
  p.ns = n; p.nsh = nh;
  p.off_w = x_bytes; p.off_h = p.off_w + n * p.w_stage;
  p.off_fixed = std::max(main_bytes(n, nh), fin_bytes);
  return "";
}

// Debug hook: stamps of CTA 0 of the last launch: 0 start, 1 setup done, 2 dependency wait done, 3 X landed, 4 first W1 landed,
// 5 first H accumulator ready, 6 first H tile written, 7 last H tile written, 8 Y complete, 9 B0, 10 partials pushed, 11 B1,
// 12 reduced + first statistics, 13 B2, 14 normalised, 15 (B3 +) second LayerNorm, 16 stores complete, 17 end;
// 32 + 16k + {0 GEMM1_k operands ready, 1 issued, 2 committed, 3 W2_k landed, 4 H_k tile ready, 5 GEMM2_k issued, 6 committed,
//            7 W1_k requested, 8 H_k accumulator seen, 9 loaded, 10 activated, 11 tile free, 12 tile written}.
int ffn_timeline(int enable, unsigned long long* out192) {
  g_ffn_timeline_enabled = enable;
  if (out192 != nullptr) EC_CUDA(cudaMemcpyFromSymbol(out192, g_ffn_timeline, 288 * sizeof(unsigned long long)));
  return EC_OK;
}

bool ffn_fused_fits(int M, int D, int hidden) {
  FfnDev p{};
  return plan_ffn(M, D, hidden, 0, p).empty();
}

int launch_ffn_fused(const FfnArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.ln_mode == 1 || a.ln_mode == 2, "bad LayerNorm mode");
  EC_REQUIRE(a.ln1_g != nullptr && a.ln1_b != nullptr && a.b1 != nullptr && a.b2 != nullptr, "missing FFN parameters");
  EC_REQUIRE(a.residual != nullptr && a.out_f32 != nullptr, "fused FFN needs the residual stream in and out");
  EC_REQUIRE(a.ln_mode == 2 || a.ln_out != nullptr, "LayerNorm mode 1 needs ln_out");
  FfnDev p{};
  const std::string why = plan_ffn(a.M, a.D, a.hidden, a.cluster, p);
  EC_REQUIRE(why.empty(), why);
  const int m_tiles = cdiv(a.M, kBlockM);
  p.b1 = a.b1; p.b2 = a.b2;
  p.ln_mode = a.ln_mode; p.ln1_g = a.ln1_g; p.ln1_b = a.ln1_b; p.ln2_g = a.ln2_g; p.ln2_b = a.ln2_b; p.ln_eps = a.ln_eps;
  p.has_ln_out = a.ln_out != nullptr;
  p.dbg = g_ffn_timeline_enabled;

  CUtensorMap tmX, tmW1, tmW2, tmRes, tmOutF, tmLn;
  EC_TRY(make_operand_map(&tmX, EC_PREC_BF16, a.x_act, a.M, a.D, kBlockM));
  EC_TRY(make_operand_map(&tmW1, EC_PREC_BF16, a.w1, a.hidden, a.D, kHC));
  EC_TRY(make_operand_map(&tmW2, EC_PREC_BF16, a.w2, a.D, a.hidden, p.n2));
  EC_TRY(make_slab_map(&tmRes, true, a.residual, a.M, a.D, a.D));
  EC_TRY(make_slab_map(&tmOutF, true, a.out_f32, a.M, a.D, a.D));
  tmLn = tmX;
  if (a.ln_out != nullptr) EC_TRY(make_slab_map(&tmLn, false, a.ln_out, a.M, a.D, a.D));

  const size_t smem = static_cast<size_t>(p.off_fixed) + kFfnFixedBytes + 1024;
  EC_REQUIRE(smem <= 227 * 1024, "fused FFN tile does not fit in shared memory");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  EC_CUDA(attr_err);
  EC_TRY(launch_pdl_cluster(ffn_fused_kernel, dim3(m_tiles * p.CS), dim3(kFfnThreads), smem, stream, p.CS, tmX, tmW1, tmW2, tmRes, tmOutF, tmLn, p));
  return EC_OK;
}

}  // namespace ec
