// TMA-staged variant of the bf16 relative-position (grouped) attention core.
//
// Same algorithm, indexing and math as attention_bf16.cu (reference models/attentions.py:549-620, 645-718; closed form in
// SURVEY.md section 8 row a9).  What changes is how K, V and the E band reach shared memory: the per-element cp.async gather
// (96 four-byte copies + address arithmetic per thread and key tile -- about 70 % of that kernel's instructions) is replaced by
// a handful of cp.async.bulk.tensor loads issued by one thread:
//   * a head's feature range [h*d, (h+1)*d) of the grouped (B, T/G, G*D) view is at most two runs that are contiguous inside one
//     frame of the q|k|v matrix; each run becomes a "panel": panel 0 up to 64 features (128-byte rows, SWIZZLE_128B), panel 1 up
//     to 32 features (64-byte rows, SWIZZLE_64B).  The frame stride G of the grouped view is the TMA element stride of the row
//     dimension, so one box per panel fetches the 64 keys of a tile.  A box must start on a 16-byte boundary of the row, so it
//     starts at the run's channel rounded down to a multiple of 8; the few leading columns belong to the previous head.
//   * the head dimension is a contraction index for both score products, so its order is free: the kernel works in "panel order"
//     (panel 0 columns, then panel 1 columns) for Q, K, E and V alike, and maps back to (frame, channel) only when it gathers Q
//     and scatters the output.  Columns a panel holds beyond the head's features are neighbouring heads' data; Q is zero there.
//   * swizzled rows keep every mma fragment load and ldmatrix conflict-free without padding.
// Zero padding of the grouped view (frames >= T of the last group) is restored by zeroing that key row after the copy lands.
#include "ec_common.cuh"
#include "ec_tma.cuh"
#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace ec {

// per head and panel: frame offset inside the group, first channel of the box (a multiple of 8: TMA needs a 16-byte aligned
// start in the contiguous dimension), how many leading box columns belong to the previous head, number of features
struct AttnPanels { int fo[2], ch[2], shift[2], n[2]; };
struct AttnDevT {
  const __nv_bfloat16* qkv; const float* u; const float* v; const int* x_len;
  int B, T, D, H, G, d, Tg;
  void* out; int ld_out;
  float scale_log2;
  AttnPanels pan[8];
};

// kHalf: fp16 operands (split mode: 11 significant bits at the same MMA rate), else bf16
template <bool kHalf>
__device__ __forceinline__ void mma_bf16_t(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (kHalf)
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool kHalf>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if constexpr (kHalf) { __half2 h = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
  else { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
}
template <bool kHalf>
__device__ __forceinline__ float2 unpack16x2(uint32_t w) {
  if constexpr (kHalf) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  else return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

constexpr int kTM = 64, kTN = 64, kTGW = 80, kTGStride = 81;
constexpr int tma_attn_ctas(int kt) { return kt <= 3 ? 4 : 3; }

// byte offset of 16-byte chunk `chunk` of row `row` in a panel: 128-byte rows XOR (row & 7), 64-byte rows XOR ((row >> 1) & 3)
__device__ __forceinline__ uint32_t p0_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ uint32_t p1_off(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

template <int KT0, int KT1, typename OutT>     // k-tiles of 16 features in panel 0 (<= 4) and panel 1 (<= 2)
__global__ void __launch_bounds__(128, tma_attn_ctas(KT0 + KT1))
relpos_attn_tma_kernel(const __grid_constant__ CUtensorMap tmKV0, const __grid_constant__ CUtensorMap tmKV1,
                       const __grid_constant__ CUtensorMap tmE0, const __grid_constant__ CUtensorMap tmE1, const AttnDevT p) {
  constexpr bool kHalf = IsSplit<OutT>::value;       // split mode: fp16 operands in, packed (hi, lo) pairs out
  constexpr int KT = KT0 + KT1, DP = KT * 16, PR = DP / 2, STRQ = DP + 8;
  constexpr int kK0 = kTN * 128, kK1 = KT1 ? kTN * 64 : 0;               // panel bytes of a 64-key tile
  constexpr int kE0 = 128 * 128, kE1 = KT1 ? 128 * 64 : 0;
  extern __shared__ uint8_t sm_raw[];
  const uint32_t raw = smem_u32(sm_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = sm_raw + (base - raw);
  // layout: K0 | K1 | E0 | E1 | V0 | V1 | score strip | table | barriers     (every panel base is a multiple of 1024)
  constexpr int oK0 = 0, oK1 = oK0 + kK0, oE0 = oK1 + kK1, oE1 = oE0 + kE0, oV0 = oE1 + kE1, oV1 = oV0 + kK0, oG = oV1 + kK1;
  constexpr int oTab = oG + 4 * 16 * kTGStride * 4, oBar = oTab + PR * 8;
  float* Gs = reinterpret_cast<float*>(bp + oG);
  int2* tab = reinterpret_cast<int2*>(bp + oTab);                        // [PR] (frame offset, channel) of panel-order feature pair
  const uint32_t ke_bar = base + oBar, v_bar = base + oBar + 8;
  // The query tile only passes through shared memory once (gather + bias add), then lives in A fragments for the whole key
  // loop; its staging area aliases V and the score strip, which are first written after the fragments have been read.
  __nv_bfloat16* Qu = reinterpret_cast<__nv_bfloat16*>(bp + oV0);
  __nv_bfloat16* Qv = Qu + kTM * STRQ;
  static_assert(2 * kTM * STRQ * 2 <= kK0 + kK1 + 4 * 16 * kTGStride * 4, "Q staging must fit in V + strip");

  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kTM;
  const int D = p.D, G = p.G, Tg = p.Tg, T = p.T;
  const size_t row3 = static_cast<size_t>(3) * D;
  const AttnPanels& hp = p.pan[h];
  for (int pr = tid; pr < PR; pr += 128) {           // input independent: built before the dependency wait
    const int c = 2 * pr, pn = c < KT0 * 16 ? 0 : 1, pc = c - pn * KT0 * 16;
    tab[pr] = make_int2((pc >= hp.shift[pn] && pc < hp.shift[pn] + hp.n[pn]) ? hp.fo[pn] : -1, hp.ch[pn] + pc);
  }
  if (tid == 0) {
    tma_prefetch_desc(&tmKV0); tma_prefetch_desc(&tmE0);
    if (KT1) { tma_prefetch_desc(&tmKV1); tma_prefetch_desc(&tmE1); }
    mbar_init(ke_bar, 1); mbar_init(v_bar, 1);
    fence_barrier_init();
  }
  grid_dependency_wait();
  grid_launch_dependents();
  __syncthreads();
  const int xl = p.x_len != nullptr ? p.x_len[b] : T;
  const __nv_bfloat16* qkv_b = p.qkv + static_cast<size_t>(b) * T * row3;

  // Copy schedule (no second buffer): K and the E band of tile j+1 are fetched while tile j does softmax and P.V; V of tile j
  // is fetched while tile j does its two score products.
  auto issue_ke = [&](int j0) {
    if (tid == 0) {
      mbar_arrive_expect_tx(ke_bar, kK0 + kK1 + kE0 + kE1);
      const int ebase = Tg - 1 + j0 - i0 - (kTM - 1);
      tma_load_2d(base + oK0, &tmKV0, ke_bar, D + hp.ch[0], b * T + G * j0 + hp.fo[0]);
      tma_load_2d(base + oE0, &tmE0, ke_bar, hp.fo[0] * D + hp.ch[0], ebase);     // grouped feature index = fo * D + channel
      if (KT1) {
        tma_load_2d(base + oK1, &tmKV1, ke_bar, D + hp.ch[1], b * T + G * j0 + hp.fo[1]);
        tma_load_2d(base + oE1, &tmE1, ke_bar, hp.fo[1] * D + hp.ch[1], ebase);
      }
    }
  };
  auto issue_v = [&](int j0) {
    if (tid == 0) {
      mbar_arrive_expect_tx(v_bar, kK0 + kK1);
      tma_load_2d(base + oV0, &tmKV0, v_bar, 2 * D + hp.ch[0], b * T + G * j0 + hp.fo[0]);
      if (KT1) tma_load_2d(base + oV1, &tmKV1, v_bar, 2 * D + hp.ch[1], b * T + G * j0 + hp.fo[1]);
    }
  };
  // zero padding of the grouped view: key row Tg-1 of a panel whose frame (Tg-1)*G + fo lies beyond the utterance
  auto zero_pad_row = [&](int j0, int o0, int o1) {
    const int r = Tg - 1 - j0;
    if (r >= 0 && r < kTN) {
      if ((Tg - 1) * G + hp.fo[0] >= T && tid < 8) *reinterpret_cast<uint4*>(bp + o0 + r * 128 + tid * 16) = make_uint4(0, 0, 0, 0);
      if (KT1 && (Tg - 1) * G + hp.fo[1] >= T && tid >= 8 && tid < 12) *reinterpret_cast<uint4*>(bp + o1 + r * 64 + (tid - 8) * 16) = make_uint4(0, 0, 0, 0);
    }
  };
  const bool padded = Tg * G > T;
  issue_ke(0);

  // ---- stage Qu / Qv = bf16(q + u), bf16(q + v) in panel order; a thread owns one feature pair and walks the rows ----
  {
    constexpr int RPP = 128 / PR;
    const bool stager = tid < PR * RPP;
    const int pr = tid % PR, rsub = tid / PR;
    const int2 te = tab[pr];
    const bool col_ok = stager && te.x >= 0;
    if (stager) {
      constexpr int NR = (kTM + RPP - 1) / RPP;
      float2 uu = make_float2(0.f, 0.f), vv = uu;
      if (col_ok) { uu = __ldg(reinterpret_cast<const float2*>(p.u + te.y)); vv = __ldg(reinterpret_cast<const float2*>(p.v + te.y)); }
      uint32_t qraw[NR];
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const int i = i0 + rsub + k * RPP, frame = i * G + te.x;
        qraw[k] = 0u;
        if (col_ok && rsub + k * RPP < kTM && i < Tg && frame < T) qraw[k] = __ldg(reinterpret_cast<const uint32_t*>(qkv_b + frame * row3 + te.y));
      }
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const int r = rsub + k * RPP, i = i0 + r;
        if (r < kTM) {
          uint32_t qu = 0, qv = 0;
          if (col_ok && i < Tg) {
            const float2 q = unpack16x2<kHalf>(qraw[k]);
            qu = pack2<kHalf>(q.x + uu.x, q.y + uu.y);
            qv = pack2<kHalf>(q.x + vv.x, q.y + vv.y);
          }
          *reinterpret_cast<uint32_t*>(Qu + r * STRQ + 2 * pr) = qu;
          *reinterpret_cast<uint32_t*>(Qv + r * STRQ + 2 * pr) = qv;
        }
      }
    }
  }
  __syncthreads();
  uint32_t qu[KT][4], qv[KT][4];
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    const __nv_bfloat16* qa = Qu + (w * 16 + g) * STRQ + kt * 16 + 2 * t;
    const __nv_bfloat16* qb = Qv + (w * 16 + g) * STRQ + kt * 16 + 2 * t;
    qu[kt][0] = *reinterpret_cast<const uint32_t*>(qa); qu[kt][1] = *reinterpret_cast<const uint32_t*>(qa + 8 * STRQ);
    qu[kt][2] = *reinterpret_cast<const uint32_t*>(qa + 8); qu[kt][3] = *reinterpret_cast<const uint32_t*>(qa + 8 * STRQ + 8);
    qv[kt][0] = *reinterpret_cast<const uint32_t*>(qb); qv[kt][1] = *reinterpret_cast<const uint32_t*>(qb + 8 * STRQ);
    qv[kt][2] = *reinterpret_cast<const uint32_t*>(qb + 8); qv[kt][3] = *reinterpret_cast<const uint32_t*>(qb + 8 * STRQ + 8);
  }

  float o[2 * KT][4];
#pragma unroll
  for (int n = 0; n < 2 * KT; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float* Gw = Gs + w * 16 * kTGStride;
  const int eo = (3 - w) * 16;
  // B-fragment word of k-tile kt for operand row `row` (row & 7 == g): panel, 16-byte chunk 2*kt (+1 for the upper 8 features)
  auto frag = [&](int o0, int o1, int row, int kt, int hi) -> uint32_t {
    const uint8_t* a = kt < KT0 ? bp + o0 + p0_off(row, 2 * kt + hi) : bp + o1 + p1_off(row, 2 * (kt - KT0) + hi);
    return *reinterpret_cast<const uint32_t*>(a + 4 * t);
  };

  uint32_t par = 0;
  for (int j0 = 0; j0 < Tg; j0 += kTN, par ^= 1) {
    mbar_wait(ke_bar, par);
    if (padded && j0 + kTN >= Tg) zero_pad_row(j0, oK0, oK1);
    __syncthreads();                                 // pad row visible; P.V of the last tile (and the Q fragment reads) done
    issue_v(j0);

    // ---- G = Qv_w . Eband_w^T (16 x 80) -> per-warp fp32 strip ----
    {
      float acc[kTGW / 8][4];
#pragma unroll
      for (int n = 0; n < kTGW / 8; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const uint32_t a0 = qv[kt][0], a1 = qv[kt][1], a2 = qv[kt][2], a3 = qv[kt][3];
#pragma unroll
        for (int n = 0; n < kTGW / 8; ++n) {
          const int row = eo + n * 8 + g;
          mma_bf16_t<kHalf>(acc[n], a0, a1, a2, a3, frag(oE0, oE1, row, kt, 0), frag(oE0, oE1, row, kt, 1));
        }
      }
#pragma unroll
      for (int n = 0; n < kTGW / 8; ++n) {
        float* g0 = Gw + g * kTGStride + n * 8 + 2 * t;
        g0[0] = acc[n][0]; g0[1] = acc[n][1];
        g0[8 * kTGStride] = acc[n][2]; g0[8 * kTGStride + 1] = acc[n][3];
      }
    }
    // ---- S = Qu_w . K^T (16 x 64) ----
    float s[kTN / 8][4];
#pragma unroll
    for (int n = 0; n < kTN / 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const uint32_t a0 = qu[kt][0], a1 = qu[kt][1], a2 = qu[kt][2], a3 = qu[kt][3];
#pragma unroll
      for (int n = 0; n < kTN / 8; ++n) {
        const int row = n * 8 + g;
        mma_bf16_t<kHalf>(s[n], a0, a1, a2, a3, frag(oK0, oK1, row, kt, 0), frag(oK0, oK1, row, kt, 1));
      }
    }
    __syncthreads();                                 // every warp is done with K and the E band (and has written its strip)
    if (j0 + kTN < Tg) issue_ke(j0 + kTN);
    // ---- relative shift, scale, key mask ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < kTN / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = g + ((e & 2) ? 8 : 0);
        const int jl = n * 8 + 2 * t + (e & 1);
        const int j = j0 + jl;
        const float val = (s[n][e] + Gw[r * kTGStride + jl - r + 15]) * p.scale_log2;
        const bool valid = j < Tg && j * G < xl;
        s[n][e] = valid ? val : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[n][e]);
      }
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      float m = mx[hrow];
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      const float m_new = fmaxf(m_run[hrow], m);
      m_use[hrow] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[hrow] = exp2f(m_run[hrow] - m_use[hrow]);
      m_run[hrow] = m_new;
      l_run[hrow] *= corr[hrow];
    }
#pragma unroll
    for (int n = 0; n < kTN / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pe = exp2f(s[n][e] - m_use[e >> 1]);
        s[n][e] = pe;
        l_run[e >> 1] += pe;
      }
    }
#pragma unroll
    for (int n = 0; n < 2 * KT; ++n) { o[n][0] *= corr[0]; o[n][1] *= corr[0]; o[n][2] *= corr[1]; o[n][3] *= corr[1]; }
    mbar_wait(v_bar, par);
    if (padded && j0 + kTN >= Tg) { zero_pad_row(j0, oV0, oV1); __syncthreads(); }
    // ---- O += P . V : P accumulators of two adjacent key n-tiles form one 16-key A fragment; V^T via ldmatrix.trans ----
#pragma unroll
    for (int kt2 = 0; kt2 < kTN / 16; ++kt2) {
      const uint32_t a0 = pack2<kHalf>(s[2 * kt2][0], s[2 * kt2][1]), a1 = pack2<kHalf>(s[2 * kt2][2], s[2 * kt2][3]);
      const uint32_t a2 = pack2<kHalf>(s[2 * kt2 + 1][0], s[2 * kt2 + 1][1]), a3 = pack2<kHalf>(s[2 * kt2 + 1][2], s[2 * kt2 + 1][3]);
      const int mi = lane >> 3, rr = lane & 7;
      const int row = kt2 * 16 + (mi & 1) * 8 + rr;  // key row this lane addresses; (mi >> 1) selects the upper 8 features
#pragma unroll
      for (int np = 0; np < KT; ++np) {            // pairs of 8-wide feature tiles
        uint32_t b0, b1, b2, b3;
        const uint32_t addr = np < KT0 ? base + oV0 + p0_off(row, 2 * np + (mi >> 1)) : base + oV1 + p1_off(row, 2 * (np - KT0) + (mi >> 1));
        ldsm_x4_trans(addr, b0, b1, b2, b3);
        mma_bf16_t<kHalf>(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16_t<kHalf>(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
  }

  // ---- normalise and scatter to (B, T, D) ----
  using Tr = ActTraits<OutT>;
  OutT* out = reinterpret_cast<OutT*>(p.out);
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {
    float l = l_run[hrow];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const int i = i0 + w * 16 + g + hrow * 8;
    if (i >= Tg) continue;
#pragma unroll
    for (int n = 0; n < 2 * KT; ++n) {
      const int2 e = tab[(n * 8 + 2 * t) >> 1];      // features c, c+1 of the panel order (same frame: widths and D are even)
      if (e.x >= 0) {
        const int frame = i * G + e.x;
        if (frame < T) {
          OutT* dst = out + (static_cast<size_t>(b) * T + frame) * p.ld_out + e.y;
          dst[0] = Tr::to(o[n][hrow * 2] * inv);
          dst[1] = Tr::to(o[n][hrow * 2 + 1] * inv);
        }
      }
    }
  }
}

// 2D map with a traversal stride on the row dimension: a (box_cols x box_rows) box whose rows are `row_stride` tensor rows apart.
static int make_strided_map(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_cols, int box_rows, int row_stride,
                            CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_fn();
  EC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  const size_t pitch = static_cast<size_t>(ld) * 2;
  EC_REQUIRE(pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "attention operands must be 16-byte aligned with 16-byte row pitch");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows * row_stride)};
  cuuint32_t estr[2] = {1, static_cast<cuuint32_t>(row_stride)};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (attention) failed with code " + std::to_string(static_cast<int>(r)));
  return EC_OK;
}

template <int KT0, int KT1, typename OutT>
static int launch_tma_inst(const AttnArgs& a, AttnDevT& p, cudaStream_t stream) {
  constexpr int KT = KT0 + KT1;
  CUtensorMap kv0, kv1, e0, e1;
  const int e_rows = 2 * p.Tg - 1;
  EC_TRY(make_strided_map(&kv0, a.qkv, a.B * a.T, 3 * a.D, 3 * a.D, 64, kTN, a.G, CU_TENSOR_MAP_SWIZZLE_128B));
  EC_TRY(make_strided_map(&e0, a.E, e_rows, a.G * a.D, a.G * a.D, 64, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B));
  kv1 = kv0; e1 = e0;
  if (KT1) {
    EC_TRY(make_strided_map(&kv1, a.qkv, a.B * a.T, 3 * a.D, 3 * a.D, 32, kTN, a.G, CU_TENSOR_MAP_SWIZZLE_64B));
    EC_TRY(make_strided_map(&e1, a.E, e_rows, a.G * a.D, a.G * a.D, 32, 128, 1, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  const size_t smem = 1024 + 2 * (kTN * 128 + (KT1 ? kTN * 64 : 0)) + 128 * 128 + (KT1 ? 128 * 64 : 0) + 4 * 16 * kTGStride * 4 + KT * 8 * 8 + 16;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(relpos_attn_tma_kernel<KT0, KT1, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  EC_CUDA(attr_err);
  dim3 grid(cdiv(p.Tg, kTM), p.H, p.B);
  return launch_pdl(relpos_attn_tma_kernel<KT0, KT1, OutT>, grid, dim3(128), smem, stream, kv0, kv1, e0, e1, p);
}

// Returns EC_OK after launching, or 1 (without setting an error) when the shape does not fit the panel scheme; the caller then
// uses the cp.async kernel of attention_bf16.cu.
int try_launch_relpos_attention_tma(const AttnArgs& a, cudaStream_t stream, bool* launched) {
  *launched = false;
  static const bool disabled = [] { const char* e = getenv("EFFCONF_ATTN_TMA"); return e != nullptr && e[0] == '0'; }();
  if (disabled) return EC_OK;
  if (a.G < 1 || a.G % 2 == 0 || (a.G * a.D) % a.H != 0 || a.H > 8) return EC_OK;
  AttnDevT p{};
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(a.qkv); p.u = a.u; p.v = a.v; p.x_len = a.x_len;
  p.B = a.B; p.T = a.T; p.D = a.D; p.H = a.H; p.G = a.G;
  p.d = (a.G * a.D) / a.H;
  if (p.d % 2 != 0 || a.D % 8 != 0 || (reinterpret_cast<uintptr_t>(a.qkv) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.E) & 15) != 0) return EC_OK;
  if (kTN * a.G > 256) return EC_OK;                 // TMA box limit on the strided row dimension
  const int P = (a.G - a.T % a.G) % a.G;
  p.Tg = (a.T + P) / a.G;
  p.out = a.out; p.ld_out = a.ld_out;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(p.d));
  // split every head's feature range into runs inside one frame, then into panels: the widest piece (<= 64) and one more (<= 32)
  int kt0 = 0, kt1 = 0;
  for (int h = 0; h < a.H; ++h) {
    struct Piece { int fo, ch, shift, n; } pc[4];
    int np = 0, pos = 0;                              // pos: feature index inside the head
    while (pos < p.d) {
      const int f = h * p.d + pos, fo = f / a.D;
      int ch = f % a.D, run = std::min(p.d - pos, a.D - ch);   // contiguous inside this frame
      pos += run;
      while (run > 0) {                               // a run that does not fit one 64-column box continues in the next panel
        const int shift = ch & 7, n = std::min(run, 64 - shift);
        if (np == 4) return EC_OK;
        pc[np++] = Piece{fo, ch - shift, shift, n};
        ch += n; run -= n;
      }
    }
    if (np > 2) return EC_OK;
    if (np == 2 && pc[0].shift + pc[0].n <= 32 && pc[1].shift + pc[1].n > 32) std::swap(pc[0], pc[1]);
    if (np == 2 && pc[1].shift + pc[1].n > 32) return EC_OK;
    for (int i = 0; i < 2; ++i) {
      const Piece q = i < np ? pc[i] : Piece{0, 0, 0, 0};
      if (q.shift % 2 != 0 || q.n % 2 != 0) return EC_OK;
      p.pan[h].fo[i] = q.fo; p.pan[h].ch[i] = q.ch; p.pan[h].shift[i] = q.shift; p.pan[h].n[i] = q.n;
    }
    kt0 = std::max(kt0, cdiv(pc[0].shift + pc[0].n, 16));
    kt1 = std::max(kt1, np == 2 ? cdiv(pc[1].shift + pc[1].n, 16) : 0);
  }
  int rc;
  if (kt1 == 0 && kt0 <= 2) rc = a.in_f16 ? launch_tma_inst<2, 0, SplitBf16>(a, p, stream) : launch_tma_inst<2, 0, __nv_bfloat16>(a, p, stream);
  else if (kt1 == 0 && kt0 == 3) rc = a.in_f16 ? launch_tma_inst<3, 0, SplitBf16>(a, p, stream) : launch_tma_inst<3, 0, __nv_bfloat16>(a, p, stream);
  else if (kt1 == 0 && kt0 == 4) rc = a.in_f16 ? launch_tma_inst<4, 0, SplitBf16>(a, p, stream) : launch_tma_inst<4, 0, __nv_bfloat16>(a, p, stream);
  else if (kt0 <= 4 && kt1 <= 2) rc = a.in_f16 ? launch_tma_inst<4, 2, SplitBf16>(a, p, stream) : launch_tma_inst<4, 2, __nv_bfloat16>(a, p, stream);
  else return EC_OK;
  *launched = rc == EC_OK;
  return rc;
}

}  // namespace ec
