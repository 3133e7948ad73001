// bf16-operand variant of the relative-position (grouped) attention core (fast mode).
//
// Same algorithm and indexing as attention.cu (reference models/attentions.py:549-620, 645-718; closed form in SURVEY.md
// section 8 row a9) with bf16 q|k|v and E produced directly by the QKV / pos GEMM epilogues:
//   * K / V / E-band tiles live in shared memory as bf16, the query tile in registers (A fragments) -> 3-4 CTAs per SM, which
//     is what lets the 384 / 512 CTA grids of the first two stages run as a single wave,
//   * mma.sync m16n8k16 bf16 with fp32 accumulation; fp32 online softmax; P is re-packed from the accumulator layout
//     straight into A fragments; V^T fragments come from ldmatrix.trans,
//   * K / V / E staging with cp.async (4- or 8-byte copies, zero fill) driven by a per-head (frame offset, channel)
//     lookup table built once per CTA.
#include "ec_common.cuh"
#include <mutex>

namespace ec {

struct AttnDevB {
  const __nv_bfloat16* qkv; const __nv_bfloat16* E; const float* u; const float* v; const int* x_len;
  int B, T, D, H, G, d, Tg;
  void* out; int ld_out;
  float scale_log2;
};

__device__ __forceinline__ void cp_async_b(uint32_t dst, const void* src, uint32_t src_bytes, int bytes) {
  if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_b() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
// kHalf: fp16 operands (split mode: 11 significant bits at the same MMA rate), else bf16
template <bool kHalf>
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
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
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool kHalf>
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  if constexpr (kHalf) { __half2 h = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
  else { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
}
template <bool kHalf>
__device__ __forceinline__ float2 unpack16x2(uint32_t w) {
  if constexpr (kHalf) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  else return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

constexpr int kBM = 64, kBN = 64, kBGW = 80, kBGStride = 81;

constexpr int attn_ctas_per_sm(int KT) { return KT <= 3 ? 4 : 3; }

template <int KT, typename OutT>     // KT = k-tiles of 16 covering the head dim
__global__ void __launch_bounds__(128, attn_ctas_per_sm(KT)) relpos_attn_bf16_kernel(const AttnDevB p) {
  constexpr int DP = KT * 16, STR = DP + 8;          // bf16 elements; row pitch 2*STR bytes keeps 32-bit fragment loads conflict-free
  constexpr int PR = DP / 2;                         // bf16 pairs per row
  constexpr bool kHalf = IsSplit<OutT>::value;       // split mode: fp16 operands in, packed (hi, lo) pairs out
  extern __shared__ __align__(16) uint8_t smb[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smb);
  __nv_bfloat16* Es = Ks + kBN * STR;
  __nv_bfloat16* Vs = Es + 128 * STR;
  float* Gs = reinterpret_cast<float*>(Vs + kBN * STR);
  // The query tile only passes through shared memory once (gather + bias add), then lives in A fragments for the whole key
  // loop; its staging area aliases V and the score strip, which are first written after the fragments have been read.
  __nv_bfloat16* Qu = Vs;
  __nv_bfloat16* Qv = Qu + kBM * STR;
  static_assert(sizeof(__nv_bfloat16) * kBM * STR <= sizeof(float) * 4 * 16 * kBGStride, "Q staging must fit in V + strip");
  int2* tab = reinterpret_cast<int2*>(Gs + 4 * 16 * kBGStride);     // [PR] (frame offset, channel) of feature pair c

  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kBM;
  const int D = p.D, G = p.G, d = p.d, Tg = p.Tg, T = p.T;
  const size_t row3 = static_cast<size_t>(3) * D;
  const int f0 = h * d;
  for (int pr = tid; pr < PR; pr += 128) {           // input independent: built before the dependency wait
    const int c = 2 * pr;
    int foff = f0 / D, ch = f0 - foff * D + c;
    while (ch >= D) { ch -= D; ++foff; }
    tab[pr] = make_int2(c < d ? foff : -1, ch);
  }
  grid_dependency_wait();
  grid_launch_dependents();
  __syncthreads();
  const int xl = p.x_len != nullptr ? p.x_len[b] : T;
  const __nv_bfloat16* qkv_b = p.qkv + static_cast<size_t>(b) * T * row3;
  // Staging map: every thread owns ONE feature pair (column) and walks the rows, so the (frame offset, channel) lookup, the
  // u / v biases and all column predicates are loop invariants; per copy only the row address and bounds change.
  constexpr int RPP = 128 / PR;                      // rows covered per pass by the PR*RPP active threads
  const bool stager = tid < PR * RPP;
  const int pr = tid % PR, rsub = tid / PR;
  const int2 te = tab[pr];
  const bool col_ok = stager && te.x >= 0;

  float o[2 * KT][4];
#pragma unroll
  for (int n = 0; n < 2 * KT; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float* Gw = Gs + w * 16 * kBGStride;
  const int eo = (3 - w) * 16;
  const size_t e_row = static_cast<size_t>(G) * D;

  // Copy schedule (no second buffer): K and the E band of tile j+1 are fetched while tile j does softmax and P.V; V of tile j
  // is fetched while tile j does its two score products.  Every cp.async batch therefore has a compute phase to hide behind.
  auto issue_ke = [&](int j0) {
    if (stager) {
      const int ebase = Tg - 1 + j0 - i0 - (kBM - 1);
      const __nv_bfloat16* kcol = qkv_b + D + te.y;
      for (int r = rsub; r < kBN; r += RPP) {
        const int j = j0 + r, frame = j * G + te.x;
        const bool ok = col_ok && j < Tg && frame < T;
        cp_async_b(smem_u32(Ks + r * STR + 2 * pr), ok ? kcol + frame * row3 : qkv_b, ok ? 4u : 0u, 4);
      }
      const __nv_bfloat16* ecol = p.E + f0 + 2 * pr;
      for (int r = rsub; r < 128; r += RPP) {
        const int ee = ebase + r;
        const bool ok = col_ok && ee >= 0 && ee <= 2 * Tg - 2;
        cp_async_b(smem_u32(Es + r * STR + 2 * pr), ok ? ecol + ee * e_row : p.E, ok ? 4u : 0u, 4);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto issue_v = [&](int j0) {
    if (stager) {
      const __nv_bfloat16* vcol = qkv_b + 2 * D + te.y;
      for (int r = rsub; r < kBN; r += RPP) {
        const int j = j0 + r, frame = j * G + te.x;
        const bool ok = col_ok && j < Tg && frame < T;
        cp_async_b(smem_u32(Vs + r * STR + 2 * pr), ok ? vcol + frame * row3 : qkv_b, ok ? 4u : 0u, 4);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_ke(0);

  // ---- stage Qu / Qv = bf16(q + u), bf16(q + v): all of a thread's rows are loaded before the first use ----
  if (stager) {
    constexpr int NR = (kBM + RPP - 1) / RPP;
    float2 uu = make_float2(0.f, 0.f), vv = uu;
    if (col_ok) { uu = __ldg(reinterpret_cast<const float2*>(p.u + te.y)); vv = __ldg(reinterpret_cast<const float2*>(p.v + te.y)); }
    uint32_t qraw[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int i = i0 + rsub + k * RPP, frame = i * G + te.x;
      qraw[k] = 0u;
      if (col_ok && rsub + k * RPP < kBM && i < Tg && frame < T) qraw[k] = __ldg(reinterpret_cast<const uint32_t*>(qkv_b + frame * row3 + te.y));
    }
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int r = rsub + k * RPP, i = i0 + r;
      if (r < kBM) {
        uint32_t qu = 0, qv = 0;
        if (col_ok && i < Tg) {
          const float2 q = unpack16x2<kHalf>(qraw[k]);
          qu = pack_bf16<kHalf>(q.x + uu.x, q.y + uu.y);
          qv = pack_bf16<kHalf>(q.x + vv.x, q.y + vv.y);
        }
        *reinterpret_cast<uint32_t*>(Qu + r * STR + 2 * pr) = qu;
        *reinterpret_cast<uint32_t*>(Qv + r * STR + 2 * pr) = qv;
      }
    }
  }

  __syncthreads();
  uint32_t qu[KT][4], qv[KT][4];
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    const __nv_bfloat16* qa = Qu + (w * 16 + g) * STR + kt * 16 + 2 * t;
    const __nv_bfloat16* qb = Qv + (w * 16 + g) * STR + kt * 16 + 2 * t;
    qu[kt][0] = *reinterpret_cast<const uint32_t*>(qa); qu[kt][1] = *reinterpret_cast<const uint32_t*>(qa + 8 * STR);
    qu[kt][2] = *reinterpret_cast<const uint32_t*>(qa + 8); qu[kt][3] = *reinterpret_cast<const uint32_t*>(qa + 8 * STR + 8);
    qv[kt][0] = *reinterpret_cast<const uint32_t*>(qb); qv[kt][1] = *reinterpret_cast<const uint32_t*>(qb + 8 * STR);
    qv[kt][2] = *reinterpret_cast<const uint32_t*>(qb + 8); qv[kt][3] = *reinterpret_cast<const uint32_t*>(qb + 8 * STR + 8);
  }

  for (int j0 = 0; j0 < Tg; j0 += kBN) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                 // K / E of this tile (and Q on the first pass) visible; P.V of the last tile done
    issue_v(j0);

    // ---- G = Qv_w . Eband_w^T (16 x 80) -> per-warp fp32 strip ----
    {
      float acc[kBGW / 8][4];
#pragma unroll
      for (int n = 0; n < kBGW / 8; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const uint32_t a0 = qv[kt][0], a1 = qv[kt][1], a2 = qv[kt][2], a3 = qv[kt][3];
#pragma unroll
        for (int n = 0; n < kBGW / 8; ++n) {
          const __nv_bfloat16* eb = Es + (eo + n * 8 + g) * STR + kt * 16 + 2 * t;
          mma_bf16<kHalf>(acc[n], a0, a1, a2, a3, *reinterpret_cast<const uint32_t*>(eb), *reinterpret_cast<const uint32_t*>(eb + 8));
        }
      }
#pragma unroll
      for (int n = 0; n < kBGW / 8; ++n) {
        float* g0 = Gw + g * kBGStride + n * 8 + 2 * t;
        g0[0] = acc[n][0]; g0[1] = acc[n][1];
        g0[8 * kBGStride] = acc[n][2]; g0[8 * kBGStride + 1] = acc[n][3];
      }
    }
    // ---- S = Qu_w . K^T (16 x 64) ----
    float s[kBN / 8][4];
#pragma unroll
    for (int n = 0; n < kBN / 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const uint32_t a0 = qu[kt][0], a1 = qu[kt][1], a2 = qu[kt][2], a3 = qu[kt][3];
#pragma unroll
      for (int n = 0; n < kBN / 8; ++n) {
        const __nv_bfloat16* kb = Ks + (n * 8 + g) * STR + kt * 16 + 2 * t;
        mma_bf16<kHalf>(s[n], a0, a1, a2, a3, *reinterpret_cast<const uint32_t*>(kb), *reinterpret_cast<const uint32_t*>(kb + 8));
      }
    }
    __syncthreads();                                 // every warp is done with K and the E band
    if (j0 + kBN < Tg) issue_ke(j0 + kBN);
    else asm volatile("cp.async.commit_group;" ::: "memory");      // keeps the group count uniform for the wait below
    // ---- relative shift, scale, key mask ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < kBN / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = g + ((e & 2) ? 8 : 0);
        const int jl = n * 8 + 2 * t + (e & 1);
        const int j = j0 + jl;
        const float val = (s[n][e] + Gw[r * kBGStride + jl - r + 15]) * p.scale_log2;
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
    for (int n = 0; n < kBN / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pe = exp2f(s[n][e] - m_use[e >> 1]);
        s[n][e] = pe;
        l_run[e >> 1] += pe;
      }
    }
#pragma unroll
    for (int n = 0; n < 2 * KT; ++n) { o[n][0] *= corr[0]; o[n][1] *= corr[0]; o[n][2] *= corr[1]; o[n][3] *= corr[1]; }
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();                                 // V of this tile visible
    // ---- O += P . V : P accumulators of two adjacent key n-tiles form one 16-key A fragment; V^T via ldmatrix.trans ----
#pragma unroll
    for (int kt2 = 0; kt2 < kBN / 16; ++kt2) {
      const uint32_t a0 = pack_bf16<kHalf>(s[2 * kt2][0], s[2 * kt2][1]), a1 = pack_bf16<kHalf>(s[2 * kt2][2], s[2 * kt2][3]);
      const uint32_t a2 = pack_bf16<kHalf>(s[2 * kt2 + 1][0], s[2 * kt2 + 1][1]), a3 = pack_bf16<kHalf>(s[2 * kt2 + 1][2], s[2 * kt2 + 1][3]);
      const int mi = lane >> 3, rr = lane & 7;
      const __nv_bfloat16* vrow = Vs + (kt2 * 16 + (mi & 1) * 8 + rr) * STR + (mi >> 1) * 8;
#pragma unroll
      for (int np = 0; np < KT; ++np) {            // pairs of 8-wide dim tiles
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(smem_u32(vrow + np * 16), b0, b1, b2, b3);
        mma_bf16<kHalf>(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16<kHalf>(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
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
      const int c = n * 8 + 2 * t;                 // features c, c+1 (same frame: d and D are even)
      if (c < d) {
        const int2 e = tab[c >> 1];
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

template <int KT, typename OutT>
static int launch_inst(const AttnDevB& p, cudaStream_t stream) {
  constexpr int STR = KT * 16 + 8;
  const size_t smem = sizeof(__nv_bfloat16) * (static_cast<size_t>(kBN) * STR * 2 + 128 * STR) +
                      sizeof(float) * 4 * 16 * kBGStride + sizeof(int2) * (KT * 8) + 16;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(relpos_attn_bf16_kernel<KT, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  EC_CUDA(attr_err);
  EC_REQUIRE(smem <= 227 * 1024, "attention tile does not fit in shared memory");
  dim3 grid(cdiv(p.Tg, kBM), p.H, p.B);
  return launch_pdl(relpos_attn_bf16_kernel<KT, OutT>, grid, dim3(128), smem, stream, p);
}

// bf16 q|k|v [B*T, 3D] and E [2Tp-G, D]; requires even d and D (every shipped config except the Large family).
int launch_relpos_attention_bf16(const AttnArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.G >= 1 && a.G % 2 == 1, "attention group size must be odd");
  EC_REQUIRE((a.G * a.D) % a.H == 0, "G*D must be divisible by H");
  AttnDevB p{};
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(a.qkv); p.E = reinterpret_cast<const __nv_bfloat16*>(a.E);
  p.u = a.u; p.v = a.v; p.x_len = a.x_len;
  p.B = a.B; p.T = a.T; p.D = a.D; p.H = a.H; p.G = a.G;
  p.d = (a.G * a.D) / a.H;
  EC_REQUIRE(p.d % 2 == 0 && a.D % 2 == 0, "the bf16 attention kernel needs even head dim and model dim");
  const int P = (a.G - a.T % a.G) % a.G;
  p.Tg = (a.T + P) / a.G;
  p.out = a.out; p.ld_out = a.ld_out;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(p.d));
  const int kt = cdiv(p.d, 16);
  switch (kt) {
    case 2: return a.in_f16 ? launch_inst<2, SplitBf16>(p, stream) : launch_inst<2, __nv_bfloat16>(p, stream);
    case 3: return a.in_f16 ? launch_inst<3, SplitBf16>(p, stream) : launch_inst<3, __nv_bfloat16>(p, stream);
    case 4: return a.in_f16 ? launch_inst<4, SplitBf16>(p, stream) : launch_inst<4, __nv_bfloat16>(p, stream);
    case 5: return a.in_f16 ? launch_inst<5, SplitBf16>(p, stream) : launch_inst<5, __nv_bfloat16>(p, stream);
    case 6: return a.in_f16 ? launch_inst<6, SplitBf16>(p, stream) : launch_inst<6, __nv_bfloat16>(p, stream);
    default: EC_FAIL("unsupported attention head dim " + std::to_string(p.d) + " for the bf16 kernel");
  }
}

}  // namespace ec
