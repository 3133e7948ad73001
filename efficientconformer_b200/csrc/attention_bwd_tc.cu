// Tensor-core backward of the relative-position (grouped) attention core, bf16 operand mode (training step).
// Same mathematics as attention_bwd.cu (notation there); the CUDA-core version stays the TF32 parity path.  Here every
// contraction runs on the tensor cores as a BATCHED GEMM over (batch, head) with bf16 operands and fp32 accumulation
// (mma.sync m16n8k16 + ldmatrix, cp.async double buffering; these are T' x T' x d problems with d = 42..96 -- far below the size
// where a tcgen05/TMEM pipeline pays for its setup, and there are B*H = 128 of them per launch to fill the 148 SMs):
//     pack        q|k|v (grouped head reinterpretation, +u / +v, zero pad frames), dO, E  ->  dense [B*H, T', dp] bf16 operands
//     S1  = Qu K^T            Rel = Qv E_h^T            dP = dO V^T                     (NT GEMMs, fp32 out)
//     rows        P = softmax(scale (S1 + skew(Rel)) + key mask);  dS = scale P (dP - sum_j P dP);  dRel = unskew(dS)   (warp per row)
//     dV  = P^T dO            dK = dS^T Qu              dQu = dS K       dQv = dRel E_h       dE_b = dRel^T Qv       (TN / NN GEMMs)
//     unpack      dq = dQu + dQv, dk, dv scattered to [B*T, 3D] (real frames); du / dv / dE reduced over the batch in a fixed order
// No atomics anywhere: bit-reproducible.  P, dS and dRel are materialised in bf16 in the caller's workspace (a flash-style
// version that keeps them on chip needs the in-kernel relative shift of attention_tma.cu transposed; planned successor).
#include "ec_common.cuh"
#include <algorithm>

namespace ec {

namespace {
using bf16 = __nv_bfloat16;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;           // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int BM = 64, BN = 64, BK = 32;
constexpr int kPitchK = BK + 8;       // [rows][BK] tiles (K contiguous): 80-byte pitch, conflict-free ldmatrix
constexpr int kPitchM = BM + 8;       // [BK][64] tiles (M / N contiguous): 144-byte pitch

struct BGemm {
  const bf16* A; const bf16* B; void* C;
  int M, N, K, lda, ldb, ldc;
  long long sAb, sAh, sBb, sBh, sCb, sCh;   // element strides of the (batch b, head h) planes; z = b * H + h
  int H, c_bf16;
};

// C[z] (M x N) = op(A[z]) . op(B[z]):  TA = false: A is [M, K] (K contiguous);  TA = true: A is stored [K, M] (M contiguous).
//                                      BKN = false: B is [N, K] (K contiguous); BKN = true: B is stored [K, N] (N contiguous).
template <bool TA, bool BKN>
__global__ void __launch_bounds__(128) bgemm_kernel(const BGemm g) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ __align__(16) bf16 As[2][TA ? BK * kPitchM : BM * kPitchK];
  __shared__ __align__(16) bf16 Bs[2][BKN ? BK * kPitchM : BN * kPitchK];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int z = blockIdx.z, zb = z / g.H, zh = z - zb * g.H;
  const bf16* __restrict__ A = g.A + zb * g.sAb + zh * g.sAh;
  const bf16* __restrict__ B = g.B + zb * g.sBb + zh * g.sBh;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

  auto load = [&](int kt, int st) {
    const int k0 = kt * BK;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int id = tid + it * 128;
      if (!TA) {
        const int r = id >> 2, ck = (id & 3) * 8;
        const bool ok = (m0 + r < g.M) && (k0 + ck < g.K);
        cp_async16(smem_addr(&As[st][r * kPitchK + ck]), ok ? A + static_cast<long long>(m0 + r) * g.lda + k0 + ck : A, ok);
      } else {
        const int r = id >> 3, cm = (id & 7) * 8;
        const bool ok = (k0 + r < g.K) && (m0 + cm < g.M);
        cp_async16(smem_addr(&As[st][r * kPitchM + cm]), ok ? A + static_cast<long long>(k0 + r) * g.lda + m0 + cm : A, ok);
      }
      if (!BKN) {
        const int r = id >> 2, ck = (id & 3) * 8;
        const bool ok = (n0 + r < g.N) && (k0 + ck < g.K);
        cp_async16(smem_addr(&Bs[st][r * kPitchK + ck]), ok ? B + static_cast<long long>(n0 + r) * g.ldb + k0 + ck : B, ok);
      } else {
        const int r = id >> 3, cn = (id & 7) * 8;
        const bool ok = (k0 + r < g.K) && (n0 + cn < g.N);
        cp_async16(smem_addr(&Bs[st][r * kPitchM + cn]), ok ? B + static_cast<long long>(k0 + r) * g.ldb + n0 + cn : B, ok);
      }
    }
    cp_async_commit();
  };

  const int nk = (g.K + BK - 1) / BK;
  load(0, 0);
  for (int kt = 0; kt < nk; ++kt) {
    const int st = kt & 1;
    if (kt + 1 < nk) { load(kt + 1, st ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk += 16) {
      uint32_t af[2][4], bfr[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int mb = wm + mi * 16;
        if (!TA) ldsm_x4(smem_addr(&As[st][(mb + (lane & 15)) * kPitchK + kk + (lane >> 4) * 8]), af[mi]);
        else ldsm_x4_t(smem_addr(&As[st][(kk + (lane & 7) + ((lane >> 4) << 3)) * kPitchM + mb + ((lane >> 3) & 1) * 8]), af[mi]);
      }
#pragma unroll
      for (int nj = 0; nj < 2; ++nj) {           // each x4 covers two n8 blocks: regs {b0, b1} of block 2nj, {b0, b1} of block 2nj+1
        const int nb = wn + nj * 16;
        if (!BKN) ldsm_x4(smem_addr(&Bs[st][(nb + (lane & 7) + ((lane >> 4) << 3)) * kPitchK + kk + ((lane >> 3) & 1) * 8]), bfr[nj]);
        else ldsm_x4_t(smem_addr(&Bs[st][(kk + (lane & 15)) * kPitchM + nb + (lane >> 4) * 8]), bfr[nj]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int nj = 0; nj < 2; ++nj) {
          mma16816(acc[mi][2 * nj], af[mi], bfr[nj][0], bfr[nj][1]);
          mma16816(acc[mi][2 * nj + 1], af[mi], bfr[nj][2], bfr[nj][3]);
        }
    }
    __syncthreads();
  }
  // ---- epilogue: thread owns rows (lane / 4) and (lane / 4 + 8) of each m16 block, column pairs 2 * (lane % 4) ----
  const long long cz = zb * g.sCb + zh * g.sCh;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + wm + mi * 16 + (lane >> 2) + half * 8;
      if (m >= g.M) continue;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const int n = n0 + wn + nb * 8 + (lane & 3) * 2;
        const float v0 = acc[mi][nb][half * 2], v1 = acc[mi][nb][half * 2 + 1];
        const long long o = cz + static_cast<long long>(m) * g.ldc + n;
        if (g.c_bf16) {
          bf16* C = reinterpret_cast<bf16*>(g.C);
          if (n < g.N) C[o] = __float2bfloat16_rn(v0);
          if (n + 1 < g.N) C[o + 1] = __float2bfloat16_rn(v1);
        } else {
          float* C = reinterpret_cast<float*>(g.C);
          if (n < g.N) C[o] = v0;
          if (n + 1 < g.N) C[o + 1] = v1;
        }
      }
    }
}

struct TcDev {
  const void* qkv; const void* E;            // activation type: bf16, or packed (hi, lo) pairs in split mode (rounded to bf16 while packing)
  const float* u; const float* v; const int* x_len; const float* dO;
  int B, T, D, H, G, d, dp, Tg, Tp, R, Rp;
  float scale;
  bf16 *Qu, *Qv, *Kd, *Vd, *dOd, *Eh;        // [BH, Tg, dp] x5, [H, R, dp]
  float *S1, *Rel, *dP;                      // [BH, Tg, Tp], [BH, Tg, Rp], [BH, Tg, Tp]
  bf16 *P, *dS, *dRel;                       // [BH, Tg, Tp] x2, [BH, Tg, Rp]
  float *dV, *dK, *dQu, *dQv, *dEp;          // [BH, Tg, dp] x4, [BH, R, dp]
  float *uv_part;                            // [BH][2][dp]
  float *dqkv, *dE, *du, *dv;
  void* dqkv_act; int act_prec;              // optional: dq | dk | dv straight in the activation type (operand of the QKV weight / data gradient GEMMs)
};

// ---- pack: dense per-head operands.  One thread per (bh, i, PAIR of features c, c+1): d, D and f = h*d + c are even, so a pair
//      never straddles a head or a frame; 32-bit stores, half the index arithmetic per element ---------------------------------
template <typename TIn> __device__ __forceinline__ float2 load_pair(const TIn* p);
template <> __device__ __forceinline__ float2 load_pair<bf16>(const bf16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }
template <> __device__ __forceinline__ float2 load_pair<__half>(const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }
template <> __device__ __forceinline__ float2 load_pair<SplitBf16>(const SplitBf16* p) {
  const uint2 w = *reinterpret_cast<const uint2*>(p);
  return make_float2(split_unpack(w.x), split_unpack(w.y));
}
__device__ __forceinline__ void store_pair(bf16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }

template <typename TIn>
__global__ void __launch_bounds__(256) tc_pack_kernel(const TcDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int hp = p.dp / 2;
  const long long n = static_cast<long long>(p.B) * p.H * p.Tg * hp;
  const long long row3 = 3LL * p.D;
  for (long long idx2 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx2 < n; idx2 += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = 2 * static_cast<int>(idx2 % hp);
    const long long r = idx2 / hp;
    const int i = static_cast<int>(r % p.Tg);
    const int bh = static_cast<int>(r / p.Tg);
    const int b = bh / p.H, h = bh - b * p.H;
    float2 qu = make_float2(0.f, 0.f), qv = qu, k = qu, v = qu, go = qu;
    if (c < p.d) {
      const int f = h * p.d + c, fo = f / p.D, ch = f - fo * p.D;
      const int frame = i * p.G + fo;
      float2 q = make_float2(0.f, 0.f);
      if (frame < p.T) {
        const TIn* row = reinterpret_cast<const TIn*>(p.qkv) + (static_cast<long long>(b) * p.T + frame) * row3;
        q = load_pair<TIn>(row + ch); k = load_pair<TIn>(row + p.D + ch); v = load_pair<TIn>(row + 2 * p.D + ch);
        go = *reinterpret_cast<const float2*>(p.dO + (static_cast<long long>(b) * p.T + frame) * p.D + ch);
      }
      const float2 uu = *reinterpret_cast<const float2*>(p.u + ch), vv = *reinterpret_cast<const float2*>(p.v + ch);
      qu = make_float2(q.x + uu.x, q.y + uu.y); qv = make_float2(q.x + vv.x, q.y + vv.y);
    }
    const long long idx = r * p.dp + c;
    store_pair(p.Qu + idx, qu.x, qu.y); store_pair(p.Qv + idx, qv.x, qv.y);
    store_pair(p.Kd + idx, k.x, k.y); store_pair(p.Vd + idx, v.x, v.y); store_pair(p.dOd + idx, go.x, go.y);
  }
}
// scalar variants for odd head / model dims (one thread per element)
template <typename TIn>
__global__ void __launch_bounds__(256) tc_pack_scalar_kernel(const TcDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  const long long n = static_cast<long long>(p.B) * p.H * p.Tg * p.dp;
  const long long row3 = 3LL * p.D;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % p.dp);
    const long long r = idx / p.dp;
    const int i = static_cast<int>(r % p.Tg);
    const int bh = static_cast<int>(r / p.Tg);
    const int b = bh / p.H, h = bh - b * p.H;
    float qu = 0.f, qv = 0.f, k = 0.f, v = 0.f, go = 0.f;
    if (c < p.d) {
      const int f = h * p.d + c, fo = f / p.D, ch = f - fo * p.D;
      const int frame = i * p.G + fo;
      float q = 0.f;
      if (frame < p.T) {
        const TIn* row = reinterpret_cast<const TIn*>(p.qkv) + (static_cast<long long>(b) * p.T + frame) * row3;
        q = ActTraits<TIn>::from(row[ch]); k = ActTraits<TIn>::from(row[p.D + ch]); v = ActTraits<TIn>::from(row[2 * p.D + ch]);
        go = p.dO[(static_cast<long long>(b) * p.T + frame) * p.D + ch];
      }
      qu = q + p.u[ch]; qv = q + p.v[ch];
    }
    p.Qu[idx] = __float2bfloat16_rn(qu); p.Qv[idx] = __float2bfloat16_rn(qv);
    p.Kd[idx] = __float2bfloat16_rn(k); p.Vd[idx] = __float2bfloat16_rn(v); p.dOd[idx] = __float2bfloat16_rn(go);
  }
}
__global__ void __launch_bounds__(256) tc_unpack_scalar_kernel(const TcDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  const long long n = static_cast<long long>(p.B) * p.T * p.D;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(idx % p.D);
    const long long bt = idx / p.D;
    const int frame = static_cast<int>(bt % p.T), b = static_cast<int>(bt / p.T);
    const int i = frame / p.G, fo = frame - i * p.G;
    const int f = fo * p.D + ch, h = f / p.d, c = f - h * p.d;
    const long long src = ((static_cast<long long>(b) * p.H + h) * p.Tg + i) * p.dp + c;
    float* out = p.dqkv + bt * 3 * p.D;
    out[ch] = p.dQu[src] + p.dQv[src];
    out[p.D + ch] = p.dK[src];
    out[2 * p.D + ch] = p.dV[src];
  }
}
template <typename TIn>
__global__ void __launch_bounds__(256) tc_pack_e_kernel(const TcDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  const long long n = static_cast<long long>(p.H) * p.R * p.dp;
  const long long e_row = static_cast<long long>(p.H) * p.d;            // = G * D
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % p.dp);
    const long long r = idx / p.dp;
    const int e = static_cast<int>(r % p.R), h = static_cast<int>(r / p.R);
    p.Eh[idx] = __float2bfloat16_rn(c < p.d ? ActTraits<TIn>::from(reinterpret_cast<const TIn*>(p.E)[e * e_row + h * p.d + c]) : 0.f);
  }
}

// ---- rows: softmax, dS, dRel.  One warp per (bh, i); NPL = score elements per lane -----------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(256) tc_rows_kernel(const TcDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (row >= static_cast<long long>(p.B) * p.H * p.Tg) return;
  const int i = static_cast<int>(row % p.Tg);
  const int bh = static_cast<int>(row / p.Tg), b = bh / p.H;
  const int Tg = p.Tg, xl = p.x_len != nullptr ? p.x_len[b] : p.T;
  const float* __restrict__ s1 = p.S1 + row * p.Tp;
  const float* __restrict__ rel = p.Rel + row * p.Rp + (Tg - 1 - i);       // rel[j] = Rel[i, T'-1+j-i]
  const float* __restrict__ dp_ = p.dP + row * p.Tp;
  float s[NPL], g[NPL];
  float m = -INFINITY;
#pragma unroll
  for (int q = 0; q < NPL; ++q) {
    const int j = lane + 32 * q;
    float sv = -INFINITY, gv = 0.f;
    if (j < Tg) {
      gv = dp_[j];
      if (j * p.G < xl) sv = p.scale * (s1[j] + rel[j]);
    }
    s[q] = sv; g[q] = gv; m = fmaxf(m, sv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < NPL; ++q) { s[q] = (m == -INFINITY || s[q] == -INFINITY) ? 0.f : __expf(s[q] - m); sum += s[q]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  float delta = 0.f;
#pragma unroll
  for (int q = 0; q < NPL; ++q) { s[q] *= inv; delta = fmaf(s[q], g[q], delta); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
  bf16* __restrict__ Pr = p.P + row * p.Tp;
  bf16* __restrict__ dSr = p.dS + row * p.Tp;
  bf16* __restrict__ dRr = p.dRel + row * p.Rp;
  const bf16 zero = __float2bfloat16_rn(0.f);
  // zero the parts of the dRel row outside the band [T'-1-i, 2T'-2-i]
  for (int r = lane; r < Tg - 1 - i; r += 32) dRr[r] = zero;
  for (int r = 2 * Tg - 1 - i + lane; r < p.Rp; r += 32) dRr[r] = zero;
#pragma unroll
  for (int q = 0; q < NPL; ++q) {
    const int j = lane + 32 * q;
    if (j < Tg) {
      const bf16 ds = __float2bfloat16_rn(p.scale * s[q] * (g[q] - delta));
      Pr[j] = __float2bfloat16_rn(s[q]);
      dSr[j] = ds;
      dRr[Tg - 1 - i + j] = ds;
    } else if (j < p.Tp) {
      Pr[j] = zero; dSr[j] = zero;
    }
  }
}

// ---- unpack: dq = dQu + dQv, dk, dv -> dqkv [B*T, 3D] (real frames); one thread per channel pair (64-bit accesses) ----------------
template <typename TO> __device__ __forceinline__ void unpack_store(void* base, long long i, float a, float b);
template <> __device__ __forceinline__ void unpack_store<float>(void* base, long long i, float a, float b) {
  *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + i) = make_float2(a, b);
}
struct Tf32Out {};
template <> __device__ __forceinline__ void unpack_store<Tf32Out>(void* base, long long i, float a, float b) {
  *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + i) = make_float2(round_tf32(a), round_tf32(b));
}
template <> __device__ __forceinline__ void unpack_store<bf16>(void* base, long long i, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<bf16*>(base) + i) = __floats2bfloat162_rn(a, b);
}
template <> __device__ __forceinline__ void unpack_store<SplitBf16>(void* base, long long i, float a, float b) {
  *reinterpret_cast<uint2*>(reinterpret_cast<uint32_t*>(base) + i) = make_uint2(split_pack(a), split_pack(b));
}
template <typename TO>
__global__ void __launch_bounds__(256) tc_unpack_kernel(const TcDev p, void* __restrict__ out_base) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int hD = p.D / 2;
  const long long n = static_cast<long long>(p.B) * p.T * hD;
  for (long long idx2 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx2 < n; idx2 += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = 2 * static_cast<int>(idx2 % hD);
    const long long bt = idx2 / hD;
    const int frame = static_cast<int>(bt % p.T), b = static_cast<int>(bt / p.T);
    const int i = frame / p.G, fo = frame - i * p.G;
    const int f = fo * p.D + ch, h = f / p.d, c = f - h * p.d;
    const long long src = ((static_cast<long long>(b) * p.H + h) * p.Tg + i) * p.dp + c;
    const long long o = bt * 3 * p.D + ch;
    const float2 a = *reinterpret_cast<const float2*>(p.dQu + src), a2 = *reinterpret_cast<const float2*>(p.dQv + src);
    const float2 k = *reinterpret_cast<const float2*>(p.dK + src), v = *reinterpret_cast<const float2*>(p.dV + src);
    unpack_store<TO>(out_base, o, a.x + a2.x, a.y + a2.y);
    unpack_store<TO>(out_base, o + p.D, k.x, k.y);
    unpack_store<TO>(out_base, o + 2 * p.D, v.x, v.y);
  }
}
// column sums of dQu / dQv over the grouped rows of one (b, h): uv_part[bh][0 | 1][c].  Block = 32 columns x 32 row lanes, lane ty
// adds the contiguous row chunk ty, chunk sums added in lane order (fixed order).
__global__ void __launch_bounds__(1024) tc_uv_part_kernel(const TcDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float su_s[32][33], sv_s[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int bh = blockIdx.x, c = blockIdx.y * 32 + tx;
  const int per = (p.Tg + 31) / 32, i0 = ty * per, i1 = min(p.Tg, i0 + per);
  float su = 0.f, sv = 0.f;
  if (c < p.dp) {
    const float* qu = p.dQu + static_cast<long long>(bh) * p.Tg * p.dp + c;
    const float* qv = p.dQv + static_cast<long long>(bh) * p.Tg * p.dp + c;
    for (int i = i0; i < i1; ++i) { su += qu[static_cast<long long>(i) * p.dp]; sv += qv[static_cast<long long>(i) * p.dp]; }
  }
  su_s[ty][tx] = su; sv_s[ty][tx] = sv;
  __syncthreads();
  if (ty == 0 && c < p.dp) {
    su = 0.f; sv = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) { su += su_s[q][tx]; sv += sv_s[q][tx]; }
    p.uv_part[(static_cast<long long>(bh) * 2) * p.dp + c] = su;
    p.uv_part[(static_cast<long long>(bh) * 2 + 1) * p.dp + c] = sv;
  }
}
// du[ch] = sum_b sum_fo part[b, h(fo, ch), c(fo, ch)]: 32 channels x 32 batch lanes per block (fixed order)
__global__ void __launch_bounds__(1024) tc_uv_reduce_kernel(const TcDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  __shared__ float su_s[32][33], sv_s[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + tx;
  const int per = (p.B + 31) / 32, b0 = ty * per, b1 = min(p.B, b0 + per);
  float su = 0.f, sv = 0.f;
  if (ch < p.D)
    for (int b = b0; b < b1; ++b)
      for (int fo = 0; fo < p.G; ++fo) {
        const int f = fo * p.D + ch, h = f / p.d, c = f - h * p.d;
        const long long o = ((static_cast<long long>(b) * p.H + h) * 2) * p.dp + c;
        su += p.uv_part[o]; sv += p.uv_part[o + p.dp];
      }
  su_s[ty][tx] = su; sv_s[ty][tx] = sv;
  __syncthreads();
  if (ty == 0 && ch < p.D) {
    su = 0.f; sv = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) { su += su_s[q][tx]; sv += sv_s[q][tx]; }
    p.du[ch] = su; p.dv[ch] = sv;
  }
}
// dE[e, h*d + c] = sum_b dEp[b, h, e, c]
__global__ void __launch_bounds__(128) tc_de_reduce_kernel(const TcDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int e = blockIdx.x, h = blockIdx.y;
  for (int c = threadIdx.x; c < p.d; c += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < p.B; ++b) s += p.dEp[((static_cast<long long>(b) * p.H + h) * p.R + e) * p.dp + c];
    p.dE[static_cast<long long>(e) * p.H * p.d + h * p.d + c] = s;
  }
}

struct TcLayout { size_t off[20]; size_t total; };
TcLayout tc_layout(int B, int T, int D, int H, int G) {
  const int P = (G - T % G) % G, Tg = (T + P) / G, d = (G * D) / H;
  const int dp = round_up(d, 16), Tp = round_up(Tg, 8), R = 2 * Tg - 1, Rp = round_up(R, 8);
  const size_t BH = static_cast<size_t>(B) * H;
  const size_t dense = align_up(BH * Tg * dp * 2, 256), dense32 = align_up(BH * Tg * dp * 4, 256);
  const size_t sq32 = align_up(BH * Tg * Tp * 4, 256), sq16 = align_up(BH * Tg * Tp * 2, 256);
  const size_t rl32 = align_up(BH * Tg * Rp * 4, 256), rl16 = align_up(BH * Tg * Rp * 2, 256);
  TcLayout L{}; size_t o = 0; int k = 0;
  for (int q = 0; q < 5; ++q) { L.off[k++] = o; o += dense; }                       // 0-4  Qu Qv Kd Vd dOd
  L.off[k++] = o; o += align_up(static_cast<size_t>(H) * R * dp * 2, 256);          // 5    Eh
  L.off[k++] = o; o += sq32;  L.off[k++] = o; o += rl32;  L.off[k++] = o; o += sq32;  // 6-8  S1 Rel dP
  L.off[k++] = o; o += sq16;  L.off[k++] = o; o += sq16;  L.off[k++] = o; o += rl16;  // 9-11 P dS dRel
  for (int q = 0; q < 4; ++q) { L.off[k++] = o; o += dense32; }                     // 12-15 dV dK dQu dQv
  L.off[k++] = o; o += align_up(BH * R * dp * 4, 256);                              // 16   dEp
  L.off[k++] = o; o += align_up(BH * 2 * dp * 4, 256);                              // 17   uv_part
  L.total = o;
  return L;
}

template <bool TA, bool BKN>
int run_gemm(const BGemm& g, int BH, cudaStream_t st) {
  dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), BH);
  (void)launch_dep(bgemm_kernel<TA, BKN>, dim3(grid), dim3(128), 0, st, g);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
inline int egrid(long long n) { return static_cast<int>(std::min<long long>((n + 255) / 256, 148LL * 16)); }
}  // namespace

size_t attention_bwd_tc_work_bytes(int B, int T, int D, int H, int G) { return tc_layout(B, T, D, H, G).total; }

int launch_relpos_attention_bwd_tc(int precision, const AttnArgs& a, const float* dO, float* dqkv, float* dE, float* du, float* dv, void* work,
                                   cudaStream_t st, void* dqkv_act) {
  TcDev p{};
  p.qkv = a.qkv; p.E = a.E; p.u = a.u; p.v = a.v; p.x_len = a.x_len; p.dO = dO;
  p.B = a.B; p.T = a.T; p.D = a.D; p.H = a.H; p.G = a.G;
  p.d = (a.G * a.D) / a.H; p.dp = round_up(p.d, 16);
  const int Pad = (a.G - a.T % a.G) % a.G;
  p.Tg = (a.T + Pad) / a.G; p.Tp = round_up(p.Tg, 8); p.R = 2 * p.Tg - 1; p.Rp = round_up(p.R, 8);
  p.scale = 1.f / sqrtf(static_cast<float>(p.d));
  EC_REQUIRE(p.Tg <= 1024, "attention backward (tensor-core path): at most 1024 grouped frames per utterance");
  const TcLayout L = tc_layout(a.B, a.T, a.D, a.H, a.G);
  uint8_t* w = reinterpret_cast<uint8_t*>(work);
  auto at = [&](int k) { return w + L.off[k]; };
  p.Qu = reinterpret_cast<bf16*>(at(0)); p.Qv = reinterpret_cast<bf16*>(at(1)); p.Kd = reinterpret_cast<bf16*>(at(2));
  p.Vd = reinterpret_cast<bf16*>(at(3)); p.dOd = reinterpret_cast<bf16*>(at(4)); p.Eh = reinterpret_cast<bf16*>(at(5));
  p.S1 = reinterpret_cast<float*>(at(6)); p.Rel = reinterpret_cast<float*>(at(7)); p.dP = reinterpret_cast<float*>(at(8));
  p.P = reinterpret_cast<bf16*>(at(9)); p.dS = reinterpret_cast<bf16*>(at(10)); p.dRel = reinterpret_cast<bf16*>(at(11));
  p.dV = reinterpret_cast<float*>(at(12)); p.dK = reinterpret_cast<float*>(at(13)); p.dQu = reinterpret_cast<float*>(at(14));
  p.dQv = reinterpret_cast<float*>(at(15)); p.dEp = reinterpret_cast<float*>(at(16)); p.uv_part = reinterpret_cast<float*>(at(17));
  p.dqkv = dqkv; p.dE = dE; p.du = du; p.dv = dv; p.dqkv_act = dqkv_act; p.act_prec = precision;
  const int BH = a.B * a.H, Tg = p.Tg, dp = p.dp;
  const long long sD = static_cast<long long>(Tg) * dp, sS = static_cast<long long>(Tg) * p.Tp, sR = static_cast<long long>(Tg) * p.Rp;
  const long long sE = static_cast<long long>(p.R) * dp;
  const int H = a.H;

  const bool pairs = a.D % 2 == 0 && p.d % 2 == 0;      // feature pairs stay inside one head and one frame
  const int pgrid = egrid(static_cast<long long>(BH) * Tg * dp / (pairs ? 2 : 1));
#define EC_PACK(TIN) do { if (pairs) (void)launch_dep(tc_pack_kernel<TIN>, dim3(pgrid), dim3(256), 0, st, p); else (void)launch_dep(tc_pack_scalar_kernel<TIN>, dim3(pgrid), dim3(256), 0, st, p); \
                          (void)launch_dep(tc_pack_e_kernel<TIN>, dim3(egrid(static_cast<long long>(H) * p.R * dp)), dim3(256), 0, st, p); } while (0)
  if (precision == EC_PREC_BF16X2 && !a.in_f16) EC_PACK(SplitBf16);
  else if (precision == EC_PREC_BF16X2) EC_PACK(__half);        // fp16 q|k|v / E from the forward; the backward GEMMs run on bf16 copies
  else EC_PACK(bf16);
#undef EC_PACK
  EC_CUDA(cudaGetLastError());
  // Independent launches run as parallel branches (library-owned side streams, event fork / join: capturable); the critical path is
  // pack -> max(S1, Rel, dP) -> rows -> max(dV, dK, dQu, dQv) -> unpack, the parameter-gradient tail (dE, du, dv) runs beside it.
  SideStreams& ss = side_streams();
  const bool par = side_streams_enabled() && ss.init();
  cudaStream_t s1 = par ? ss.s[0] : st, s2 = par ? ss.s[1] : st, s3 = par ? ss.s[2] : st;
  // S1 = Qu K^T, Rel = Qv Eh^T, dP = dO V^T
  if (par) EC_REQUIRE(ss.fork(st, 2), "side-stream fork failed");
  EC_TRY((run_gemm<false, false>(BGemm{p.Qv, p.Eh, p.Rel, Tg, p.R, dp, dp, dp, p.Rp, sD * H, sD, 0, sE, sR * H, sR, H, 0}, BH, st)));
  EC_TRY((run_gemm<false, false>(BGemm{p.Qu, p.Kd, p.S1, Tg, Tg, dp, dp, dp, p.Tp, sD * H, sD, sD * H, sD, sS * H, sS, H, 0}, BH, s1)));
  EC_TRY((run_gemm<false, false>(BGemm{p.dOd, p.Vd, p.dP, Tg, Tg, dp, dp, dp, p.Tp, sD * H, sD, sD * H, sD, sS * H, sS, H, 0}, BH, s2)));
  if (par) EC_REQUIRE(ss.join(st, 2), "side-stream join failed");
  const long long rows = static_cast<long long>(BH) * Tg;
  const int rgrid = static_cast<int>((rows + 7) / 8);
  if (Tg <= 128) (void)launch_dep(tc_rows_kernel<4>, dim3(rgrid), dim3(256), 0, st, p);
  else if (Tg <= 256) (void)launch_dep(tc_rows_kernel<8>, dim3(rgrid), dim3(256), 0, st, p);
  else if (Tg <= 512) (void)launch_dep(tc_rows_kernel<16>, dim3(rgrid), dim3(256), 0, st, p);
  else (void)launch_dep(tc_rows_kernel<32>, dim3(rgrid), dim3(256), 0, st, p);
  EC_CUDA(cudaGetLastError());
  if (par) EC_REQUIRE(ss.fork(st, 3), "side-stream fork failed");
  // dQv = dRel Eh (longest contraction: R) on the caller's stream; dV = P^T dO, dK = dS^T Qu (contraction over the query rows: A stored
  // [K = i, M = j]) and dQu = dS K beside it
  EC_TRY((run_gemm<false, true>(BGemm{p.dRel, p.Eh, p.dQv, Tg, dp, p.R, p.Rp, dp, dp, sR * H, sR, 0, sE, sD * H, sD, H, 0}, BH, st)));
  EC_TRY((run_gemm<true, true>(BGemm{p.P, p.dOd, p.dV, Tg, dp, Tg, p.Tp, dp, dp, sS * H, sS, sD * H, sD, sD * H, sD, H, 0}, BH, s1)));
  EC_TRY((run_gemm<true, true>(BGemm{p.dS, p.Qu, p.dK, Tg, dp, Tg, p.Tp, dp, dp, sS * H, sS, sD * H, sD, sD * H, sD, H, 0}, BH, s2)));
  EC_TRY((run_gemm<false, true>(BGemm{p.dS, p.Kd, p.dQu, Tg, dp, Tg, p.Tp, dp, dp, sS * H, sS, sD * H, sD, sD * H, sD, H, 0}, BH, s2)));
  // parameter-gradient tail on the third branch: dE_b = dRel^T Qv (per (b, h) partials) reduced over b
  EC_TRY((run_gemm<true, true>(BGemm{p.dRel, p.Qv, p.dEp, p.R, dp, Tg, p.Rp, dp, dp, sR * H, sR, sD * H, sD, sE * H, sE, H, 0}, BH, s3)));
  (void)launch_dep(tc_de_reduce_kernel, dim3(dim3(p.R, H)), dim3(128), 0, s3, p);
  EC_CUDA(cudaGetLastError());
  if (par) EC_REQUIRE(ss.join(st, 2), "side-stream join failed");      // dV, dK, dQu, dQv complete
  // du / dv from dQu / dQv: parameter gradients, beside the unpack
  if (par) {
    EC_CUDA(cudaEventRecord(ss.fork_ev, st));                          // (after the join: dQu / dQv are complete at this point of `st`)
    EC_CUDA(cudaStreamWaitEvent(s1, ss.fork_ev, 0));
  }
  const int ugrid = egrid(static_cast<long long>(a.B) * a.T * a.D / 2);
  if (pairs && p.dqkv_act != nullptr) {             // the consumer only reads the activation-type copy: skip the fp32 tensor
    if (p.act_prec == EC_PREC_BF16) (void)launch_dep(tc_unpack_kernel<bf16>, dim3(ugrid), dim3(256), 0, st, p, p.dqkv_act);
    else if (p.act_prec == EC_PREC_BF16X2) (void)launch_dep(tc_unpack_kernel<SplitBf16>, dim3(ugrid), dim3(256), 0, st, p, p.dqkv_act);
    else (void)launch_dep(tc_unpack_kernel<Tf32Out>, dim3(ugrid), dim3(256), 0, st, p, p.dqkv_act);
  } else {
    EC_REQUIRE(p.dqkv != nullptr, "attention backward: no fp32 dqkv buffer");
    if (pairs) (void)launch_dep(tc_unpack_kernel<float>, dim3(ugrid), dim3(256), 0, st, p, p.dqkv);
    else (void)launch_dep(tc_unpack_scalar_kernel, dim3(egrid(static_cast<long long>(a.B) * a.T * a.D)), dim3(256), 0, st, p);
    EC_CUDA(cudaGetLastError());
    if (p.dqkv_act != nullptr) EC_TRY(launch_cast_rows(p.act_prec, p.dqkv, p.dqkv_act, static_cast<size_t>(a.B) * a.T * 3 * a.D, st));
  }
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(tc_uv_part_kernel, dim3(dim3(BH, cdiv(dp, 32))), dim3(1024), 0, s1, p);
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(tc_uv_reduce_kernel, dim3(cdiv(a.D, 32)), dim3(1024), 0, s1, p);
  EC_CUDA(cudaGetLastError());
  if (par) {
    EC_CUDA(cudaEventRecord(ss.join_ev[0], s1)); EC_CUDA(cudaStreamWaitEvent(st, ss.join_ev[0], 0));
    EC_CUDA(cudaEventRecord(ss.join_ev[2], s3)); EC_CUDA(cudaStreamWaitEvent(st, ss.join_ev[2], 0));
  }
  return EC_OK;
}

}  // namespace ec
