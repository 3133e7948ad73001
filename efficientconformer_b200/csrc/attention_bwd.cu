// Backward of the relative-position (grouped) multi-head self-attention core (training step):
// gradients of reference models/attentions.py:549-620 / 645-718 between the Q|K|V projections and the output projection.
// Notation of attention.cu / SURVEY.md row a9, per (batch b, head h), grouped rows i, j in [0, T'), head features c in [0, d):
//     S_ij = scale * (Qu_i . K_j + Qv_i . E[T'-1+j-i]),  P = softmax_j(S + key mask),  O_i = sum_j P_ij V_j
//     dV_j = sum_i P_ij dO_i          dP_ij = dO_i . V_j          dS_ij = scale * P_ij (dP_ij - sum_j' P_ij' dP_ij')
//     dQu_i = sum_j dS_ij K_j         dK_j = sum_i dS_ij Qu_i
//     dQv_i = sum_j dS_ij E[T'-1+j-i] dE[e] = sum_b sum_{j-i = e-(T'-1)} dS_ij Qv_i
//     dq = dQu + dQv (real frames only: the appended pad frames are constants),  du = sum_{b,i} dQu_i,  dv = sum_{b,i} dQv_i
// First implementation (correct, bandwidth-aware, CUDA cores): three kernels with P and dS materialised once as (B,H,T',T') fp32
// in a caller-provided workspace (14-32 MB per CTCSmall stage), shared-memory tiles for K / V / E, no atomics (every output
// element has one owner; du / dv use per-(b,h) partial rows reduced in a fixed order).  The tensor-core version (dS and P kept
// on chip, flash style) is the planned successor.
#include "ec_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace ec {

namespace {
constexpr int kRows = 16;      // query rows per CTA (kernel 1) / key rows per CTA (kernel 2)
constexpr int kTile = 32;      // keys (kernel 1) / queries (kernel 2) staged per step
constexpr int kThreads = 256;

struct BwdDev {
  const void* qkv; const void* E; const float* u; const float* v; const int* x_len; const float* dO;
  int B, T, D, H, G, d, Tg, dp;          // dp = padded row pitch of the shared-memory tiles (odd: conflict free)
  float scale;
  float *P, *dS;                         // [B,H,Tg,Tg]
  float *Qu, *Qv, *dOg, *dQu, *dQv;      // [B,H,Tg,d] grouped dense copies / results
  float* dqkv;                           // [B*T, 3D]
  float* dE;                             // [(2Tg-1), G*D]
  float *du_part, *dv_part;              // [B*H][D]
};

template <typename T>
__device__ __forceinline__ float ld_act(const void* p, size_t i) { return ActTraits<T>::from(reinterpret_cast<const T*>(p)[i]); }

// head feature c of head h -> (frame offset inside the group, channel)
__device__ __forceinline__ void locate(int h, int d, int D, int c, int& fo, int& ch) {
  const int f = h * d + c;
  fo = f / D; ch = f - fo * D;
}
}  // namespace

// ---- kernel 1: per (b, h, 16 query rows): P, dS, dQu, dQv (+ the grouped dense copies Qu, Qv, dO) ---------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) attn_bwd_rows_kernel(const BwdDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  extern __shared__ float sm[];
  const int Tg = p.Tg, d = p.d, dp = p.dp, D = p.D, G = p.G, Tt = p.T;
  float* Sb = sm;                                   // [kRows][Tg]   scores -> P
  float* Db = Sb + kRows * Tg;                      // [kRows][Tg]   dP -> dS
  float* Qus = Db + kRows * Tg;                     // [kRows][dp]
  float* Qvs = Qus + kRows * dp;
  float* dOs = Qvs + kRows * dp;
  float* Ks = dOs + kRows * dp;                     // [kTile][dp]   K tile, later V tile
  float* Es = Ks + kTile * dp;                      // [kRows + kTile - 1][dp]  E band of the tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = blockIdx.x * kRows, h = blockIdx.y, b = blockIdx.z;
  const int xl = p.x_len != nullptr ? p.x_len[b] : Tt;
  const size_t row3 = static_cast<size_t>(3) * D;
  const size_t qkv_b = static_cast<size_t>(b) * Tt * row3;
  const size_t bh = static_cast<size_t>(b) * p.H + h;
  const size_t e_row = static_cast<size_t>(G) * D;

  // ---- stage Qu, Qv, dO rows (grouped) and write the dense copies ----
  for (int idx = tid; idx < kRows * d; idx += kThreads) {
    const int r = idx / d, c = idx - r * d, i = i0 + r;
    float qu = 0.f, qv = 0.f, go = 0.f;
    if (i < Tg) {
      int fo, ch; locate(h, d, D, c, fo, ch);
      const int frame = i * G + fo;
      const float q = frame < Tt ? ld_act<T>(p.qkv, qkv_b + frame * row3 + ch) : 0.f;
      qu = q + p.u[ch]; qv = q + p.v[ch];
      go = frame < Tt ? p.dO[(static_cast<size_t>(b) * Tt + frame) * D + ch] : 0.f;
      const size_t o = (bh * Tg + i) * d + c;
      p.Qu[o] = qu; p.Qv[o] = qv; p.dOg[o] = go;
    }
    Qus[r * dp + c] = qu; Qvs[r * dp + c] = qv; dOs[r * dp + c] = go;
  }
  __syncthreads();

  const int ti = tid >> 4, tj = tid & 15;           // thread owns row ti and key columns tj, tj + 16 of every tile
  // ---- phase A: scores ----
  for (int j0 = 0; j0 < Tg; j0 += kTile) {
    for (int idx = tid; idx < kTile * d; idx += kThreads) {
      const int r = idx / d, c = idx - r * d, j = j0 + r;
      float kv = 0.f;
      if (j < Tg) {
        int fo, ch; locate(h, d, D, c, fo, ch);
        const int frame = j * G + fo;
        if (frame < Tt) kv = ld_act<T>(p.qkv, qkv_b + frame * row3 + D + ch);
      }
      Ks[r * dp + c] = kv;
    }
    const int eb = Tg - 1 + j0 - i0 - (kRows - 1);                      // band row 0 <-> e = eb
    for (int idx = tid; idx < (kRows + kTile - 1) * d; idx += kThreads) {
      const int r = idx / d, c = idx - r * d, e = eb + r;
      Es[r * dp + c] = (e >= 0 && e <= 2 * Tg - 2) ? ld_act<T>(p.E, e * e_row + h * d + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int jl = tj + 16 * half, j = j0 + jl;
      if (j < Tg) {
        const float* qu = Qus + ti * dp; const float* qv = Qvs + ti * dp;
        const float* kr = Ks + jl * dp; const float* er = Es + (jl - ti + kRows - 1) * dp;
        float acc = 0.f;
        for (int c = 0; c < d; ++c) acc = fmaf(qu[c], kr[c], fmaf(qv[c], er[c], acc));
        const bool valid = j * G < xl;
        Sb[ti * Tg + j] = valid ? acc * p.scale : -INFINITY;
      }
    }
    __syncthreads();
  }
  // ---- phase B: softmax rows (2 rows per warp) -> P (shared + global) ----
  for (int r = warp; r < kRows; r += kThreads / 32) {
    float* row = Sb + r * Tg;
    float m = -INFINITY;
    for (int j = lane; j < Tg; j += 32) m = fmaxf(m, row[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int j = lane; j < Tg; j += 32) { const float e = (m == -INFINITY) ? 0.f : __expf(row[j] - m); row[j] = e; s += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = s > 0.f ? 1.f / s : 0.f;
    const int i = i0 + r;
    for (int j = lane; j < Tg; j += 32) {
      const float pv = row[j] * inv;
      row[j] = pv;
      if (i < Tg) p.P[(bh * Tg + i) * Tg + j] = pv;
    }
  }
  __syncthreads();
  // ---- phase C: dP = dO . V^T ----
  for (int j0 = 0; j0 < Tg; j0 += kTile) {
    for (int idx = tid; idx < kTile * d; idx += kThreads) {
      const int r = idx / d, c = idx - r * d, j = j0 + r;
      float vv = 0.f;
      if (j < Tg) {
        int fo, ch; locate(h, d, D, c, fo, ch);
        const int frame = j * G + fo;
        if (frame < Tt) vv = ld_act<T>(p.qkv, qkv_b + frame * row3 + 2 * D + ch);
      }
      Ks[r * dp + c] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int jl = tj + 16 * half, j = j0 + jl;
      if (j < Tg) {
        const float* go = dOs + ti * dp; const float* vr = Ks + jl * dp;
        float acc = 0.f;
        for (int c = 0; c < d; ++c) acc = fmaf(go[c], vr[c], acc);
        Db[ti * Tg + j] = acc;
      }
    }
    __syncthreads();
  }
  // ---- phase D: dS = scale * P * (dP - sum_j P dP) (shared + global) ----
  for (int r = warp; r < kRows; r += kThreads / 32) {
    const float* pr = Sb + r * Tg; float* dr = Db + r * Tg;
    float dl = 0.f;
    for (int j = lane; j < Tg; j += 32) dl = fmaf(pr[j], dr[j], dl);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dl += __shfl_xor_sync(0xffffffffu, dl, o);
    const int i = i0 + r;
    for (int j = lane; j < Tg; j += 32) {
      const float ds = p.scale * pr[j] * (dr[j] - dl);
      dr[j] = ds;
      if (i < Tg) p.dS[(bh * Tg + i) * Tg + j] = ds;
    }
  }
  __syncthreads();
  // ---- phase E: dQu = dS . K, dQv_i = sum_j dS_ij E[T'-1+j-i]; thread = (row ti, features tj, tj + 16, ...) ----
  constexpr int kMaxC = 9;                          // d <= 144
  float aqu[kMaxC], aqv[kMaxC];
#pragma unroll
  for (int q = 0; q < kMaxC; ++q) { aqu[q] = 0.f; aqv[q] = 0.f; }
  for (int j0 = 0; j0 < Tg; j0 += kTile) {
    for (int idx = tid; idx < kTile * d; idx += kThreads) {
      const int r = idx / d, c = idx - r * d, j = j0 + r;
      float kv = 0.f;
      if (j < Tg) {
        int fo, ch; locate(h, d, D, c, fo, ch);
        const int frame = j * G + fo;
        if (frame < Tt) kv = ld_act<T>(p.qkv, qkv_b + frame * row3 + D + ch);
      }
      Ks[r * dp + c] = kv;
    }
    const int eb = Tg - 1 + j0 - i0 - (kRows - 1);
    for (int idx = tid; idx < (kRows + kTile - 1) * d; idx += kThreads) {
      const int r = idx / d, c = idx - r * d, e = eb + r;
      Es[r * dp + c] = (e >= 0 && e <= 2 * Tg - 2) ? ld_act<T>(p.E, e * e_row + h * d + c) : 0.f;
    }
    __syncthreads();
    const int jn = min(kTile, Tg - j0);
    for (int jl = 0; jl < jn; ++jl) {
      const float ds = Db[ti * Tg + j0 + jl];
      const float* kr = Ks + jl * dp; const float* er = Es + (jl - ti + kRows - 1) * dp;
#pragma unroll
      for (int q = 0; q < kMaxC; ++q) {
        const int c = tj + 16 * q;
        if (c < d) { aqu[q] = fmaf(ds, kr[c], aqu[q]); aqv[q] = fmaf(ds, er[c], aqv[q]); }
      }
    }
    __syncthreads();
  }
  const int i = i0 + ti;
  if (i < Tg) {
#pragma unroll
    for (int q = 0; q < kMaxC; ++q) {
      const int c = tj + 16 * q;
      if (c < d) {
        const size_t o = (bh * Tg + i) * d + c;
        p.dQu[o] = aqu[q]; p.dQv[o] = aqv[q];
        int fo, ch; locate(h, d, D, c, fo, ch);
        const int frame = i * G + fo;
        if (frame < Tt) p.dqkv[(static_cast<size_t>(b) * Tt + frame) * row3 + ch] = aqu[q] + aqv[q];
      }
    }
  }
}

// ---- kernel 2: per (b, h, 16 key rows): dV_j = sum_i P_ij dO_i, dK_j = sum_i dS_ij Qu_i ------------------------------------------
__global__ void __launch_bounds__(kThreads) attn_bwd_cols_kernel(const BwdDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  extern __shared__ float sm[];
  const int Tg = p.Tg, d = p.d, dp = p.dp, D = p.D, G = p.G, Tt = p.T;
  float* Gs = sm;                                   // [kTile][dp]  dO rows of the query tile
  float* Qs = Gs + kTile * dp;                      // [kTile][dp]  Qu rows
  float* Ps = Qs + kTile * dp;                      // [kTile][kRows + 1]  P block
  float* Ss = Ps + kTile * (kRows + 1);             // [kTile][kRows + 1]  dS block
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * kRows, h = blockIdx.y, b = blockIdx.z;
  const size_t bh = static_cast<size_t>(b) * p.H + h;
  const int tjr = tid >> 4, tc = tid & 15;          // thread owns key row tjr and features tc, tc + 16, ...
  constexpr int kMaxC = 9;
  float av[kMaxC], ak[kMaxC];
#pragma unroll
  for (int q = 0; q < kMaxC; ++q) { av[q] = 0.f; ak[q] = 0.f; }
  for (int i0 = 0; i0 < Tg; i0 += kTile) {
    for (int idx = tid; idx < kTile * d; idx += kThreads) {
      const int r = idx / d, c = idx - r * d, i = i0 + r;
      const size_t o = (bh * Tg + i) * d + c;
      Gs[r * dp + c] = i < Tg ? p.dOg[o] : 0.f;
      Qs[r * dp + c] = i < Tg ? p.Qu[o] : 0.f;
    }
    for (int idx = tid; idx < kTile * kRows; idx += kThreads) {
      const int r = idx / kRows, jl = idx - r * kRows, i = i0 + r, j = j0 + jl;
      const bool ok = i < Tg && j < Tg;
      Ps[r * (kRows + 1) + jl] = ok ? p.P[(bh * Tg + i) * Tg + j] : 0.f;
      Ss[r * (kRows + 1) + jl] = ok ? p.dS[(bh * Tg + i) * Tg + j] : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < kTile; ++r) {
      const float pv = Ps[r * (kRows + 1) + tjr], ds = Ss[r * (kRows + 1) + tjr];
      const float* go = Gs + r * dp; const float* qu = Qs + r * dp;
#pragma unroll
      for (int q = 0; q < kMaxC; ++q) {
        const int c = tc + 16 * q;
        if (c < d) { av[q] = fmaf(pv, go[c], av[q]); ak[q] = fmaf(ds, qu[c], ak[q]); }
      }
    }
    __syncthreads();
  }
  const int j = j0 + tjr;
  if (j < Tg) {
    const size_t row3 = static_cast<size_t>(3) * D;
#pragma unroll
    for (int q = 0; q < kMaxC; ++q) {
      const int c = tc + 16 * q;
      if (c < d) {
        int fo, ch; locate(h, d, D, c, fo, ch);
        const int frame = j * G + fo;
        if (frame < Tt) {
          float* dst = p.dqkv + (static_cast<size_t>(b) * Tt + frame) * row3 + ch;
          dst[D] = ak[q]; dst[2 * D] = av[q];
        }
      }
    }
  }
}

// ---- kernel 3: dE[e, h*d + c] = sum_b sum_i dS[b,h,i,i+e-(T'-1)] Qv[b,h,i,c];  du / dv partial rows per (b, h) -----------------------
__global__ void __launch_bounds__(128) attn_bwd_e_kernel(const BwdDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int Tg = p.Tg, d = p.d;
  const int e = blockIdx.x, h = blockIdx.y;
  const int off = e - (Tg - 1);                     // j = i + off
  const int ilo = max(0, -off), ihi = min(Tg, Tg - off);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < p.B; ++b) {
      const size_t bh = static_cast<size_t>(b) * p.H + h;
      const float* ds = p.dS + bh * Tg * Tg;
      const float* qv = p.Qv + bh * Tg * d + c;
      for (int i = ilo; i < ihi; ++i) acc = fmaf(ds[static_cast<size_t>(i) * Tg + i + off], qv[static_cast<size_t>(i) * d], acc);
    }
    p.dE[static_cast<size_t>(e) * (p.G * p.D) + h * d + c] = acc;
  }
}
// du_part[bh][ch] / dv_part[bh][ch] = sum_i dQu / dQv over the rows of one (b, h) (features of one head map to distinct channels
// only when d <= D; for grouped heads d = G*D/H may exceed D: several features share a channel and are added in feature order)
__global__ void __launch_bounds__(128) attn_bwd_uv_kernel(const BwdDev p) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int Tg = p.Tg, d = p.d, D = p.D;
  const int h = blockIdx.x, b = blockIdx.y;
  const size_t bh = static_cast<size_t>(b) * p.H + h;
  for (int ch = threadIdx.x; ch < D; ch += blockDim.x) {
    float su = 0.f, sv = 0.f;
    for (int c = 0; c < d; ++c) {
      if ((h * d + c) % D != ch) continue;
      const float* qu = p.dQu + bh * Tg * d + c; const float* qv = p.dQv + bh * Tg * d + c;
      for (int i = 0; i < Tg; ++i) { su += qu[static_cast<size_t>(i) * d]; sv += qv[static_cast<size_t>(i) * d]; }
    }
    p.du_part[bh * D + ch] = su; p.dv_part[bh * D + ch] = sv;
  }
}
__global__ void attn_bwd_uv_reduce_kernel(const float* __restrict__ du_part, const float* __restrict__ dv_part, int n, int D,
                                          float* __restrict__ du, float* __restrict__ dv) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= D) return;
  float su = 0.f, sv = 0.f;
  for (int q = 0; q < n; ++q) { su += du_part[static_cast<size_t>(q) * D + ch]; sv += dv_part[static_cast<size_t>(q) * D + ch]; }
  du[ch] = su; dv[ch] = sv;
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
static void bwd_layout(int B, int T, int D, int H, int G, size_t* off, size_t* total) {
  const int P = (G - T % G) % G, Tg = (T + P) / G, d = (G * D) / H;
  const size_t pp = align_up(static_cast<size_t>(B) * H * Tg * Tg * 4, 256), gd = align_up(static_cast<size_t>(B) * H * Tg * d * 4, 256);
  const size_t uv = align_up(static_cast<size_t>(B) * H * D * 4, 256);
  size_t o = 0;
  off[0] = o; o += pp;  off[1] = o; o += pp;                                   // P, dS
  for (int i = 2; i < 7; ++i) { off[i] = o; o += gd; }                          // Qu, Qv, dOg, dQu, dQv
  off[7] = o; o += uv;  off[8] = o; o += uv;                                    // du_part, dv_part
  *total = o;
}
size_t attention_bwd_work_bytes(int B, int T, int D, int H, int G) {
  size_t off[9], total; bwd_layout(B, T, D, H, G, off, &total);
  return std::max(total, attention_bwd_tc_work_bytes(B, T, D, H, G));
}
static bool tc_path_enabled() {
  static const bool on = [] { const char* e = getenv("EFFCONF_ATTN_BWD_TC"); return !(e && e[0] == '0'); }();
  return on;
}

int launch_relpos_attention_bwd(int precision, const AttnArgs& a, const float* dO, float* dqkv, float* dE, float* du, float* dv, void* work,
                                cudaStream_t stream, void* dqkv_act) {
  EC_REQUIRE(a.G >= 1 && a.G % 2 == 1 && (a.G * a.D) % a.H == 0, "attention backward: bad head layout");
  EC_REQUIRE(dO && (dqkv || dqkv_act) && dE && du && dv && work, "attention backward: null argument");
  // bf16 operand mode: batched tensor-core GEMMs (attention_bwd_tc.cu); the CUDA-core kernels below are the TF32 parity path
  // (split mode: the packed operands are rounded to bf16 while packing -- the gradients of the attention core are bf16 grade)
  if ((precision == EC_PREC_BF16 || precision == EC_PREC_BF16X2) && tc_path_enabled() && (a.T + a.G - 1) / a.G <= 1024)
    return launch_relpos_attention_bwd_tc(precision, a, dO, dqkv, dE, du, dv, work, stream, dqkv_act);
  EC_REQUIRE(dqkv != nullptr, "attention backward (CUDA-core parity path): the fp32 dqkv buffer is required");
  EC_REQUIRE(precision != EC_PREC_BF16X2, "attention backward: the split mode needs the tensor-core path (<= 1024 grouped frames)");
  BwdDev p{};
  p.qkv = a.qkv; p.E = a.E; p.u = a.u; p.v = a.v; p.x_len = a.x_len; p.dO = dO;
  p.B = a.B; p.T = a.T; p.D = a.D; p.H = a.H; p.G = a.G;
  p.d = (a.G * a.D) / a.H;
  const int P = (a.G - a.T % a.G) % a.G;
  p.Tg = (a.T + P) / a.G;
  p.dp = p.d | 1;
  p.scale = 1.f / sqrtf(static_cast<float>(p.d));
  EC_REQUIRE(p.d <= 144, "attention backward: head dim <= 144");
  size_t off[9], total; bwd_layout(a.B, a.T, a.D, a.H, a.G, off, &total);
  uint8_t* w = reinterpret_cast<uint8_t*>(work);
  p.P = reinterpret_cast<float*>(w + off[0]); p.dS = reinterpret_cast<float*>(w + off[1]);
  p.Qu = reinterpret_cast<float*>(w + off[2]); p.Qv = reinterpret_cast<float*>(w + off[3]); p.dOg = reinterpret_cast<float*>(w + off[4]);
  p.dQu = reinterpret_cast<float*>(w + off[5]); p.dQv = reinterpret_cast<float*>(w + off[6]);
  p.du_part = reinterpret_cast<float*>(w + off[7]); p.dv_part = reinterpret_cast<float*>(w + off[8]);
  p.dqkv = dqkv; p.dE = dE;
  const size_t sm1 = sizeof(float) * (2 * static_cast<size_t>(kRows) * p.Tg + 3 * kRows * p.dp + kTile * p.dp + (kRows + kTile - 1) * p.dp);
  EC_REQUIRE(sm1 <= 227 * 1024, "attention backward: sequence too long for the shared-memory score rows");
  dim3 g1(cdiv(p.Tg, kRows), p.H, p.B);
  if (precision == EC_PREC_TF32) {
    static cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    EC_CUDA(e1);
    (void)launch_dep(attn_bwd_rows_kernel<float>, dim3(g1), dim3(kThreads), sm1, stream, p);
  } else {
    static cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_rows_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    EC_CUDA(e2);
    (void)launch_dep(attn_bwd_rows_kernel<__nv_bfloat16>, dim3(g1), dim3(kThreads), sm1, stream, p);
  }
  EC_CUDA(cudaGetLastError());
  const size_t sm2 = sizeof(float) * (2 * static_cast<size_t>(kTile) * p.dp + 2 * kTile * (kRows + 1));
  (void)launch_dep(attn_bwd_cols_kernel, dim3(g1), dim3(kThreads), sm2, stream, p);
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(attn_bwd_e_kernel, dim3(dim3(2 * p.Tg - 1, p.H)), dim3(128), 0, stream, p);
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(attn_bwd_uv_kernel, dim3(dim3(p.H, p.B)), dim3(128), 0, stream, p);
  EC_CUDA(cudaGetLastError());
  (void)launch_dep(attn_bwd_uv_reduce_kernel, dim3(cdiv(p.D, 128)), dim3(128), 0, stream, p.du_part, p.dv_part, p.B * p.H, p.D, du, dv);
  EC_CUDA(cudaGetLastError());
  if (dqkv_act != nullptr) EC_TRY(launch_cast_rows(precision, dqkv, dqkv_act, static_cast<size_t>(a.B) * a.T * 3 * a.D, stream));
  return EC_OK;
}

}  // namespace ec
