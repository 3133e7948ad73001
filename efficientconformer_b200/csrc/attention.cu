// Relative-position (grouped) multi-head self-attention core, flash style, relative shift done in-kernel.
//
// Restates reference models/attentions.py:549-620 (RelPosMultiHeadSelfAttention.forward) and :645-718
// (GroupedRelPosMultiHeadSelfAttention.forward) between the Q/K/V projections and the output projection:
//   P = (-T) mod G zero frames appended after the projection (:107-121), qu = q + u, qv = q + v (:674-675),
//   the contiguous (B, Tp, D) buffer is reinterpreted as (B, T' = Tp/G, H, d = G*D/H) (:681-685),
//   S[i,j] = (qu_i . k_j + qv_i . E[h, T'-1+j-i]) / sqrt(d)                (:690-692 with rel_to_abs :526-547)
//   key group j masked iff j*G >= x_len[b] (:695-701), softmax over j, O = W v (:704-707), first T frames kept (:713).
// Never materialises the (B,H,T',2T'-1) relative scores nor the (B,H,T',T') weights.
//
// v1 implementation: one CTA = 64 query groups of one (batch, head); 4 warps x 16 rows; key tiles of 64;
// mma.sync m16n8k8 TF32 with fp32 accumulation and fp32 online softmax.  For each (query tile, key tile) the needed
// diagonal band of E (127 rows) is staged in shared memory, G = Qv . Eband^T is computed per warp (16 x 80), written to
// a per-warp smem strip and read back with the per-row shift  S_E[r, j] = G[r, j - r + 15].
#include "ec_common.cuh"
#include <mutex>

namespace ec {

struct AttnDev {
  const float* qkv; const float* E; const float* u; const float* v; const int* x_len;
  int B, T, D, H, G, d, Tg;
  void* out; int ld_out;
  float scale_log2;
};

__device__ __forceinline__ uint32_t tf32_bits(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// Ampere-style async copy global -> shared with zero fill (src_bytes in {0, size}); LDGSTS on sm_100a.
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int kAttBM = 64;

// BN = keys per tile: 64, or 32 for the widest heads (d = 135 of the Medium / Large grouped blocks) whose 64-key tiles do not fit
// in shared memory.  The staged E band has kAttBM + BN rows, the per-warp score strip is 16 x (BN + 16).
// kPacked (split mode): q|k|v and E arrive as packed (hi, lo) bf16 pairs; they are unpacked (16 significant bits) and rounded to
// TF32 while staging, so the contractions of the attention core run at TF32 precision like the parity mode.
__device__ __forceinline__ float unpack_tf32(float w) { return round_tf32(split_unpack(__float_as_uint(w))); }

template <int DPT, typename OutT, int BN, bool kPacked>
__global__ void __launch_bounds__(128) relpos_attn_kernel(const AttnDev p) {
  constexpr int DP = DPT * 8, STR = DP + 4;
  constexpr int kAttBN = BN, kGW = BN + 16, kGStride = kGW + 1, kBand = kAttBM + BN;
  extern __shared__ float sm[];
  float* Qu = sm;
  float* Qv = Qu + kAttBM * STR;
  float* Ks = Qv + kAttBM * STR;
  float* Vs = Ks + kAttBN * STR;
  float* Es = Vs + kAttBN * STR;
  float* Gs = Es + kBand * STR;

  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kAttBM;
  const int D = p.D, G = p.G, d = p.d, Tg = p.Tg, T = p.T;
  const int xl = p.x_len != nullptr ? p.x_len[b] : T;
  const size_t row3 = static_cast<size_t>(3) * D;
  const float* qkv_b = p.qkv + static_cast<size_t>(b) * T * row3;

  grid_dependency_wait();
  grid_launch_dependents();
  // (frame offset, channel) of this head's first feature inside a grouped row: flat f = h*d + c -> frame i*G + f/D, channel f%D
  const int f0 = h * d;
  const int foff0 = f0 / D, ch0 = f0 - foff0 * D;
  const bool vec2 = ((d | D) & 1) == 0;          // pairs (c, c+1) never straddle a frame and stay 8-byte aligned
  auto locate = [&](int c, int& foff, int& ch) {  // c < d <= G*D, so at most G-1 wraps
    foff = foff0; ch = ch0 + c;
    while (ch >= D) { ch -= D; ++foff; }
  };

  // ---- stage Qu / Qv (q + u, q + v rounded to tf32): 8-byte loads, 8 pairs in flight per thread ----
  // vec2 staging map: every thread owns one feature pair (column) and walks the rows -> lookups and biases are loop invariant
  constexpr int PRc = DP / 2, RPP = 128 / PRc;
  const bool stager = tid < PRc * RPP;
  const int prc = tid % PRc, rsub = tid / PRc, cc = 2 * prc;
  int cfoff = 0, cch = 0;
  if (cc < d) locate(cc, cfoff, cch);
  const bool col_ok = stager && cc < d;
  if (vec2) {
    if (stager) {
      float2 uu = make_float2(0.f, 0.f), vv = uu;
      if (col_ok) { uu = __ldg(reinterpret_cast<const float2*>(p.u + cch)); vv = __ldg(reinterpret_cast<const float2*>(p.v + cch)); }
#pragma unroll 4
      for (int r = rsub; r < kAttBM; r += RPP) {
        const int i = i0 + r;
        float2 qu = make_float2(0.f, 0.f), qv = qu;
        if (col_ok && i < Tg) {
          const int frame = i * G + cfoff;
          float2 q = make_float2(0.f, 0.f);                     // appended pad frames are exact zeros
          if (frame < T) q = __ldg(reinterpret_cast<const float2*>(qkv_b + frame * row3 + cch));
          if constexpr (kPacked) { q.x = split_unpack(__float_as_uint(q.x)); q.y = split_unpack(__float_as_uint(q.y)); }
          qu = make_float2(q.x + uu.x, q.y + uu.y); qv = make_float2(q.x + vv.x, q.y + vv.y);
        }
        *reinterpret_cast<float2*>(Qu + r * STR + cc) = make_float2(round_tf32(qu.x), round_tf32(qu.y));
        *reinterpret_cast<float2*>(Qv + r * STR + cc) = make_float2(round_tf32(qv.x), round_tf32(qv.y));
      }
    }
  } else {
    for (int idx = tid; idx < kAttBM * DP; idx += 128) {
      const int r = idx / DP, c = idx % DP;
      const int i = i0 + r;
      float qu = 0.f, qv = 0.f;
      if (c < d && i < Tg) {
        int foff, ch; locate(c, foff, ch);
        const int frame = i * G + foff;
        float q = frame < T ? __ldg(qkv_b + frame * row3 + ch) : 0.f;
        if constexpr (kPacked) q = split_unpack(__float_as_uint(q));
        qu = q + __ldg(p.u + ch);
        qv = q + __ldg(p.v + ch);
      }
      Qu[r * STR + c] = round_tf32(qu);
      Qv[r * STR + c] = round_tf32(qv);
    }
  }

  float o[DPT][4];
#pragma unroll
  for (int n = 0; n < DPT; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float* Gw = Gs + w * 16 * kGStride;
  const int eo = (3 - w) * 16;                       // this warp's first row inside the staged E band
  const size_t e_row = static_cast<size_t>(G) * D;

  for (int j0 = 0; j0 < Tg; j0 += kAttBN) {
    __syncthreads();
    // ---- stage the K, V tile and the E band with cp.async (global -> shared, zero-fill outside the valid region).
    // K / V / E arrive already rounded to TF32 by the producing GEMM epilogues (round_out), so no register pass is needed.
    const int ebase = Tg - 1 + j0 - i0 - (kAttBM - 1);
    if (vec2) {
      if (stager) {
        const float* kcol = qkv_b + D + cch;
        for (int r = rsub; r < kAttBN; r += RPP) {
          const int j = j0 + r, frame = j * G + cfoff;
          const bool ok = col_ok && j < Tg && frame < T;
          const float* src = ok ? kcol + frame * row3 : qkv_b;
          cp_async_8(smem_u32(Ks + r * STR + cc), src, ok ? 8u : 0u);
          cp_async_8(smem_u32(Vs + r * STR + cc), src + D, ok ? 8u : 0u);
        }
        const float* ecol = p.E + f0 + cc;
        for (int r = rsub; r < kBand; r += RPP) {
          const int e = ebase + r;
          const bool ok = col_ok && e >= 0 && e <= 2 * Tg - 2;
          cp_async_8(smem_u32(Es + r * STR + cc), ok ? ecol + e * e_row : p.E, ok ? 8u : 0u);
        }
      }
    } else {
      for (int idx = tid; idx < kAttBN * DP; idx += 128) {
        const int r = idx / DP, c = idx % DP;
        const int j = j0 + r;
        const float* src = qkv_b;
        uint32_t bytes = 0;
        if (c < d && j < Tg) {
          int foff, ch; locate(c, foff, ch);
          const int frame = j * G + foff;
          if (frame < T) { src = qkv_b + frame * row3 + D + ch; bytes = 4; }
        }
        cp_async_4(smem_u32(Ks + r * STR + c), src, bytes);
        cp_async_4(smem_u32(Vs + r * STR + c), src + D, bytes);
      }
      for (int idx = tid; idx < kBand * DP; idx += 128) {
        const int r = idx / DP, c = idx % DP;
        const int e = ebase + r;
        const bool ok = c < d && e >= 0 && e <= 2 * Tg - 2;
        cp_async_4(smem_u32(Es + r * STR + c), ok ? p.E + e * e_row + f0 + c : p.E, ok ? 4u : 0u);
      }
    }
    cp_async_wait_all();
    if constexpr (kPacked) {      // every thread converts, in place, exactly the words its own cp.async wrote (visible after the wait)
      if (vec2) {
        if (stager) {
          for (int r = rsub; r < kAttBN; r += RPP) {
            float2* k2 = reinterpret_cast<float2*>(Ks + r * STR + cc); float2* v2 = reinterpret_cast<float2*>(Vs + r * STR + cc);
            *k2 = make_float2(unpack_tf32(k2->x), unpack_tf32(k2->y)); *v2 = make_float2(unpack_tf32(v2->x), unpack_tf32(v2->y));
          }
          for (int r = rsub; r < kBand; r += RPP) {
            float2* e2 = reinterpret_cast<float2*>(Es + r * STR + cc);
            *e2 = make_float2(unpack_tf32(e2->x), unpack_tf32(e2->y));
          }
        }
      } else {
        for (int idx = tid; idx < kAttBN * DP; idx += 128) {
          const int r = idx / DP, c = idx % DP;
          Ks[r * STR + c] = unpack_tf32(Ks[r * STR + c]); Vs[r * STR + c] = unpack_tf32(Vs[r * STR + c]);
        }
        for (int idx = tid; idx < kBand * DP; idx += 128) {
          const int r = idx / DP, c = idx % DP;
          Es[r * STR + c] = unpack_tf32(Es[r * STR + c]);
        }
      }
    }
    __syncthreads();

    // ---- G = Qv_w . Eband_w^T  (16 x 80) -> per-warp smem strip ----
    {
      float acc[kGW / 8][4];
#pragma unroll
      for (int n = 0; n < kGW / 8; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
#pragma unroll
      for (int kt = 0; kt < DPT; ++kt) {
        const float* qa = Qv + (w * 16 + g) * STR + kt * 8 + t;
        const uint32_t a0 = __float_as_uint(qa[0]), a1 = __float_as_uint(qa[8 * STR]);
        const uint32_t a2 = __float_as_uint(qa[4]), a3 = __float_as_uint(qa[8 * STR + 4]);
#pragma unroll
        for (int n = 0; n < kGW / 8; ++n) {
          const float* eb = Es + (eo + n * 8 + g) * STR + kt * 8 + t;
          mma_tf32(acc[n], a0, a1, a2, a3, __float_as_uint(eb[0]), __float_as_uint(eb[4]));
        }
      }
#pragma unroll
      for (int n = 0; n < kGW / 8; ++n) {
        float* g0 = Gw + g * kGStride + n * 8 + 2 * t;
        g0[0] = acc[n][0]; g0[1] = acc[n][1];
        g0[8 * kGStride] = acc[n][2]; g0[8 * kGStride + 1] = acc[n][3];
      }
    }
    __syncwarp();

    // ---- S = Qu_w . K^T (16 x 64) ----
    float s[kAttBN / 8][4];
#pragma unroll
    for (int n = 0; n < kAttBN / 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < DPT; ++kt) {
      const float* qa = Qu + (w * 16 + g) * STR + kt * 8 + t;
      const uint32_t a0 = __float_as_uint(qa[0]), a1 = __float_as_uint(qa[8 * STR]);
      const uint32_t a2 = __float_as_uint(qa[4]), a3 = __float_as_uint(qa[8 * STR + 4]);
#pragma unroll
      for (int n = 0; n < kAttBN / 8; ++n) {
        const float* kb = Ks + (n * 8 + g) * STR + kt * 8 + t;
        mma_tf32(s[n], a0, a1, a2, a3, __float_as_uint(kb[0]), __float_as_uint(kb[4]));
      }
    }
    // ---- relative shift, scale, key mask ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < kAttBN / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = g + ((e & 2) ? 8 : 0);
        const int jl = n * 8 + 2 * t + (e & 1);
        const int j = j0 + jl;
        const float val = (s[n][e] + Gw[r * kGStride + jl - r + 15]) * p.scale_log2;
        const bool valid = j < Tg && j * G < xl;
        s[n][e] = valid ? val : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[n][e]);
      }
    }
    // ---- online softmax (base 2) ----
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
    for (int n = 0; n < kAttBN / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pe = exp2f(s[n][e] - m_use[e >> 1]);
        s[n][e] = pe;
        l_run[e >> 1] += pe;
      }
    }
#pragma unroll
    for (int n = 0; n < DPT; ++n) { o[n][0] *= corr[0]; o[n][1] *= corr[0]; o[n][2] *= corr[1]; o[n][3] *= corr[1]; }
    // ---- O += P . V  (k index of the k-tile is permuted: slot t <-> key 2t, slot t+4 <-> key 2t+1) ----
#pragma unroll
    for (int kt = 0; kt < kAttBN / 8; ++kt) {
      const uint32_t a0 = tf32_bits(s[kt][0]), a1 = tf32_bits(s[kt][2]), a2 = tf32_bits(s[kt][1]), a3 = tf32_bits(s[kt][3]);
      const float* vb = Vs + (kt * 8 + 2 * t) * STR + g;
#pragma unroll
      for (int n = 0; n < DPT; ++n) mma_tf32(o[n], a0, a1, a2, a3, __float_as_uint(vb[n * 8]), __float_as_uint(vb[STR + n * 8]));
    }
  }

  // ---- normalise and scatter to (B, T, D): grouped row i / head h / col c -> frame i*G + (h*d+c)/D, channel (h*d+c)%D ----
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
    for (int n = 0; n < DPT; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = n * 8 + 2 * t + e;
        if (c < d) {
          int foff, ch; locate(c, foff, ch);
          const int frame = i * G + foff;
          if (frame < T) out[(static_cast<size_t>(b) * T + frame) * p.ld_out + ch] = Tr::to(o[n][hrow * 2 + e] * inv);
        }
      }
    }
  }
}

template <int DPT, typename OutT, bool kPacked>
static int launch_attn_inst(const AttnDev& p, cudaStream_t stream) {
  constexpr int STR = DPT * 8 + 4;
  constexpr int BN = DPT > 16 ? 32 : 64;
  const size_t smem = sizeof(float) * (static_cast<size_t>(kAttBM) * STR * 2 + BN * STR * 2 + (kAttBM + BN) * STR + 4 * 16 * (BN + 17));
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(relpos_attn_kernel<DPT, OutT, BN, kPacked>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  EC_CUDA(attr_err);
  EC_REQUIRE(smem <= 227 * 1024, "attention tile does not fit in shared memory");
  dim3 grid(cdiv(p.Tg, kAttBM), p.H, p.B);
  return launch_pdl(relpos_attn_kernel<DPT, OutT, BN, kPacked>, grid, dim3(128), smem, stream, p);
}

template <typename T, bool kPacked = false>
static int launch_attn_t(const AttnDev& p, cudaStream_t stream) {
  const int dpt = cdiv(p.d, 8);
  switch (dpt) {
    case 3: return launch_attn_inst<3, T, kPacked>(p, stream);
    case 5: return launch_attn_inst<5, T, kPacked>(p, stream);
    case 6: return launch_attn_inst<6, T, kPacked>(p, stream);
    case 7: return launch_attn_inst<7, T, kPacked>(p, stream);
    case 8: return launch_attn_inst<8, T, kPacked>(p, stream);
    case 10: return launch_attn_inst<10, T, kPacked>(p, stream);
    case 12: return launch_attn_inst<12, T, kPacked>(p, stream);
    case 17: return launch_attn_inst<17, T, kPacked>(p, stream);
    default: EC_FAIL("unsupported attention head dim " + std::to_string(p.d));
  }
}

int launch_relpos_attention(int precision, const AttnArgs& a, cudaStream_t stream) {
  if ((precision == EC_PREC_BF16 && !a.in_f32) || (precision == EC_PREC_BF16X2 && a.in_f16)) {
    bool launched = false;                           // TMA-staged kernel when the head layout fits its panel scheme
    EC_TRY(try_launch_relpos_attention_tma(a, stream, &launched));
    return launched ? EC_OK : launch_relpos_attention_bf16(a, stream);
  }
  EC_REQUIRE(a.G >= 1 && a.G % 2 == 1, "attention group size must be odd");
  EC_REQUIRE((a.G * a.D) % a.H == 0, "G*D must be divisible by H");
  AttnDev p{};
  p.qkv = reinterpret_cast<const float*>(a.qkv); p.E = reinterpret_cast<const float*>(a.E); p.u = a.u; p.v = a.v; p.x_len = a.x_len;
  p.B = a.B; p.T = a.T; p.D = a.D; p.H = a.H; p.G = a.G;
  p.d = (a.G * a.D) / a.H;
  const int P = (a.G - a.T % a.G) % a.G;
  p.Tg = (a.T + P) / a.G;
  p.out = a.out; p.ld_out = a.ld_out;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(p.d));
  if (precision == EC_PREC_TF32) return launch_attn_t<float>(p, stream);
  if (precision == EC_PREC_BF16) return launch_attn_t<__nv_bfloat16>(p, stream);     // fp32 inputs, TF32 math, bf16 output (odd head dims)
  if (precision == EC_PREC_BF16X2) return launch_attn_t<SplitBf16, true>(p, stream);  // packed inputs and output, TF32 math
  EC_FAIL("unknown precision");
}

}  // namespace ec
