// CTC head on device: per-frame log-sum-exp + argmax, CTC negative log-likelihood (alpha recursion), greedy collapse.
//
//   launch_logsoftmax_argmax : log_softmax denominators and argmax ids      (reference models/losses.py:66, model_ctc.py:99)
//   launch_ctc_loss          : nn.CTCLoss(blank=0, reduction='none', zero_infinity=False) then .mean()   (models/losses.py:54,65-69)
//   launch_greedy_collapse   : merge repeats then drop blanks == the reference's per-frame loop          (models/model_ctc.py:105-130)
// The reference loops over (b, t) on the host with one device sync per comparison; here it is one thread per utterance.
#include "ec_common.cuh"

namespace ec {

__global__ void __launch_bounds__(256) logsoftmax_argmax_kernel(const float* __restrict__ logits, int rows, int V,
                                                                float* __restrict__ lse, int* __restrict__ amax) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = logits + static_cast<size_t>(row) * V;
  float m = -INFINITY; int mi = 0x7fffffff;
  for (int c = lane; c < V; c += 32) {
    const float v = x[c];
    if (v > m) { m = v; mi = c; }           // strictly greater keeps the lowest index inside a lane (c ascending)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  float s = 0.f;
  for (int c = lane; c < V; c += 32) s += expf(x[c] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) { lse[row] = m + logf(s); amax[row] = mi; }
}

int launch_logsoftmax_argmax(const float* logits, int rows, int V, float* lse, int* argmax, cudaStream_t stream) {
  if (rows == 0) return EC_OK;
  (void)launch_dep(logsoftmax_argmax_kernel, dim3(cdiv(rows, 8)), dim3(256), 0, stream, logits, rows, V, lse, argmax);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// One CTA per utterance; alpha over the 2U+1 extended labels double-buffered in shared memory (Graves et al. 2006).
__global__ void __launch_bounds__(256) ctc_alpha_kernel(const float* __restrict__ logits, const float* __restrict__ lse, int T, int V,
                                                        const int* __restrict__ logits_len, const long long* __restrict__ targets,
                                                        int target_stride, const long long* __restrict__ target_len,
                                                        float* __restrict__ loss_per_utt) {
  grid_dependency_wait();
  grid_launch_dependents();
  extern __shared__ float alpha_sm[];
  const int b = blockIdx.x;
  const int U = static_cast<int>(target_len[b]);
  const int S = 2 * U + 1;
  int Tb = logits_len[b];
  if (Tb > T) Tb = T;
  float* a0 = alpha_sm;
  float* a1 = alpha_sm + S;
  int* ext = reinterpret_cast<int*>(alpha_sm + 2 * S);
  const long long* y = targets + static_cast<size_t>(b) * target_stride;
  for (int s = threadIdx.x; s < S; s += blockDim.x) ext[s] = (s & 1) ? static_cast<int>(y[s >> 1]) : 0;
  __syncthreads();
  const float* lg = logits + static_cast<size_t>(b) * T * V;
  const float* ls = lse + static_cast<size_t>(b) * T;
  if (Tb <= 0) {
    if (threadIdx.x == 0) loss_per_utt[b] = INFINITY;
    return;
  }
  for (int s = threadIdx.x; s < S; s += blockDim.x) a0[s] = s < 2 ? lg[ext[s]] - ls[0] : -INFINITY;
  __syncthreads();
  float* prev = a0; float* cur = a1;
  for (int t = 1; t < Tb; ++t) {
    const float* lgt = lg + static_cast<size_t>(t) * V;
    const float lst = ls[t];
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      const int e = ext[s];
      const float x0 = prev[s];
      const float x1 = s > 0 ? prev[s - 1] : -INFINITY;
      const float x2 = (s > 1 && e != 0 && e != ext[s - 2]) ? prev[s - 2] : -INFINITY;
      const float acc = lse3(x0, x1, x2);
      cur[s] = acc == -INFINITY ? -INFINITY : acc + (lgt[e] - lst);
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
  if (threadIdx.x == 0) {
    const float l = S > 1 ? lse3(prev[S - 1], prev[S - 2], -INFINITY) : prev[0];
    loss_per_utt[b] = -l;
  }
}

// Warp-per-utterance variant for 2U+1 <= 32*kCtcNS: every lane keeps kCtcNS consecutive extended-label states in
// registers, neighbours s-1 / s-2 come from the lane's own registers or two shuffles, the emission log-probs of the next
// frame are prefetched while the current frame is combined; no block-level barrier in the T-step recursion.
// log2-domain logsumexp of three values, branch free (an all -inf triple gives ex2 -> 0 and lg2(0) = -inf by itself), so the
// kCtcNS independent states of a lane interleave in the instruction stream instead of serialising behind a divergent early return
__device__ __forceinline__ float lse3_log2(float a, float b, float c) {
  // min/max network: m = largest, (o1, o2) = the other two; the largest contributes ex2(0) = 1, so only two ex2 are needed
  const float hi = fmaxf(a, b), o2 = fminf(a, b);
  const float m = fmaxf(hi, c), o1 = fminf(hi, c);
  const float ms = (m == -INFINITY) ? 0.f : m;
  float e1, e2, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(o1 - ms));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(o2 - ms));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((m == -INFINITY ? 0.f : 1.f) + e1 + e2));
  return ms + r;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kCtcGatherThreads = 256;   // the whole CTA gathers emissions; warp 0 then runs the recursion
template <int kCtcNS>
__global__ void __launch_bounds__(kCtcGatherThreads) ctc_alpha_warp_kernel(const float* __restrict__ logits, const float* __restrict__ lse, int T, int V,
                                                            const int* __restrict__ logits_len, const long long* __restrict__ targets,
                                                            int target_stride, const long long* __restrict__ target_len,
                                                            float* __restrict__ loss_per_utt) {
  extern __shared__ float lp_sm[];                 // [Tb][32*kCtcNS] emission log-probs of the extended labels, gathered up front
  constexpr int SP = 32 * kCtcNS;
  grid_dependency_wait();
  const int b = blockIdx.x, lane = threadIdx.x;
  const int U = static_cast<int>(target_len[b]);
  const int S = 2 * U + 1;
  int Tb = logits_len[b];
  if (Tb > T) Tb = T;
  if (Tb <= 0) { if (threadIdx.x == 0) loss_per_utt[b] = INFINITY; return; }
  const long long* y = targets + static_cast<size_t>(b) * target_stride;
  const float* lg = logits + static_cast<size_t>(b) * T * V;
  const float* ls = lse + static_cast<size_t>(b) * T;
  // gather phase: T*S independent loads spread over the whole CTA (the recursion below is one warp and only touches shared
  // memory); state s lives at column (s % kCtcNS) * 32 + s / kCtcNS so that a lane's kCtcNS states are conflict-free
  {
    __shared__ int lab[SP];
    for (int s = threadIdx.x; s < SP; s += blockDim.x) lab[s] = (s < S && (s & 1)) ? static_cast<int>(y[s >> 1]) : 0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < Tb * SP; idx += blockDim.x) {
      const int t = idx / SP, s = idx - t * SP;
      lp_sm[t * SP + (s % kCtcNS) * 32 + s / kCtcNS] = kLog2e * (__ldg(lg + static_cast<size_t>(t) * V + lab[s]) - __ldg(ls + t));
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  bool skip[kCtcNS];
#pragma unroll
  for (int k = 0; k < kCtcNS; ++k) {
    const int s = lane * kCtcNS + k;
    const int e = (s < S && (s & 1)) ? static_cast<int>(y[s >> 1]) : 0;
    const int e2 = (s >= 2 && s < S && (s & 1)) ? static_cast<int>(y[(s - 2) >> 1]) : 0;
    skip[k] = s >= 2 && s < S && e != 0 && e != e2;
  }
  float a[kCtcNS];
#pragma unroll
  for (int k = 0; k < kCtcNS; ++k) {
    const int s = lane * kCtcNS + k;
    a[k] = (s < 2 && s < S) ? lp_sm[k * 32 + lane] : -INFINITY;
  }
  // the recursion runs in the log2 domain (emissions were scaled by log2(e) in the gather): ex2 / lg2 are single MUFU operations
  float cur[kCtcNS];
#pragma unroll
  for (int k = 0; k < kCtcNS; ++k) cur[k] = Tb > 1 ? lp_sm[SP + k * 32 + lane] : 0.f;
  for (int t = 1; t < Tb; ++t) {
    float nxt[kCtcNS];                                  // next frame's emissions: independent of the recursion, issued first
    const int tn = t + 1 < Tb ? t + 1 : t;
#pragma unroll
    for (int k = 0; k < kCtcNS; ++k) nxt[k] = lp_sm[tn * SP + k * 32 + lane];
    float pm1 = __shfl_up_sync(0xffffffffu, a[kCtcNS - 1], 1), pm2 = __shfl_up_sync(0xffffffffu, a[kCtcNS - 2], 1);
    if (lane == 0) { pm1 = -INFINITY; pm2 = -INFINITY; }
    float na[kCtcNS];
#pragma unroll
    for (int k = 0; k < kCtcNS; ++k) {
      const float x1 = k >= 1 ? a[k - 1] : pm1;
      const float x2 = skip[k] ? (k >= 2 ? a[k - 2] : (k == 1 ? pm1 : pm2)) : -INFINITY;
      const float acc = lse3_log2(a[k], x1, x2);
      na[k] = (lane * kCtcNS + k >= S) ? -INFINITY : acc + cur[k];       // -inf + finite stays -inf
    }
#pragma unroll
    for (int k = 0; k < kCtcNS; ++k) { a[k] = na[k]; cur[k] = nxt[k]; }
  }
  float e1 = -INFINITY, e2 = -INFINITY;
#pragma unroll
  for (int k = 0; k < kCtcNS; ++k) {
    const int s = lane * kCtcNS + k;
    if (s == S - 1) e1 = a[k];
    if (s == S - 2) e2 = a[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { e1 = fmaxf(e1, __shfl_xor_sync(0xffffffffu, e1, o)); e2 = fmaxf(e2, __shfl_xor_sync(0xffffffffu, e2, o)); }
  if (lane == 0) loss_per_utt[b] = -lse3_log2(e1, e2, -INFINITY) * 0.6931471805599453f;
}

__global__ void mean_kernel(const float* x, int n, float* out) {
  grid_dependency_wait();
  grid_launch_dependents();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *out = s / n;
}

int launch_mean(const float* x, int n, float* out, cudaStream_t stream) {
  (void)launch_dep(mean_kernel, dim3(1), dim3(32), 0, stream, x, n, out);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

int launch_ctc_loss(const float* logits, const float* lse, int B, int T, int V, const int* logits_len, const long long* targets,
                    int target_stride, const long long* target_len, float* loss_per_utt, float* loss_mean, cudaStream_t stream) {
  EC_REQUIRE(B > 0 && target_stride >= 0, "bad CTC shapes");
  const size_t smem = sizeof(float) * 3 * (2 * static_cast<size_t>(target_stride) + 1);
  EC_REQUIRE(smem <= 48 * 1024, "CTC target too long for the shared-memory alpha buffers");
  const int s_max = 2 * target_stride + 1;
  // warp kernel: emissions of one utterance gathered into shared memory ([T][32*NS] floats)
#define EC_CTC_WARP(NS)                                                                                                     \
  {                                                                                                                        \
    const size_t sm = static_cast<size_t>(T) * 32 * NS * sizeof(float);                                                     \
    if (sm <= 200 * 1024) {                                                                                                \
      static cudaError_t attr = cudaFuncSetAttribute(ctc_alpha_warp_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      EC_CUDA(attr);                                                                                                       \
      (void)launch_dep(ctc_alpha_warp_kernel<NS>, dim3(B), dim3(kCtcGatherThreads), sm, stream, logits, lse, T, V, logits_len, targets, target_stride, target_len, loss_per_utt); \
      launched = true;                                                                                                     \
    }                                                                                                                      \
  }
  bool launched = false;
  if (s_max <= 64) EC_CTC_WARP(2)
  else if (s_max <= 96) EC_CTC_WARP(3)
  else if (s_max <= 128) EC_CTC_WARP(4)
  else if (s_max <= 192) EC_CTC_WARP(6)
  else if (s_max <= 256) EC_CTC_WARP(8)
#undef EC_CTC_WARP
  if (!launched) (void)launch_dep(ctc_alpha_kernel, dim3(B), dim3(256), smem, stream, logits, lse, T, V, logits_len, targets, target_stride, target_len, loss_per_utt);
  EC_CUDA(cudaGetLastError());
  if (loss_mean != nullptr) {
    (void)launch_dep(mean_kernel, dim3(1), dim3(32), 0, stream, loss_per_utt, B, loss_mean);
    EC_CUDA(cudaGetLastError());
  }
  return EC_OK;
}

__global__ void greedy_collapse_kernel(const int* __restrict__ amax, int B, int T, const int* __restrict__ logits_len,
                                       int* __restrict__ ids, int* __restrict__ counts) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int n = 0, prev = 0;
  int Tb = logits_len[b];
  if (Tb > T) Tb = T;
  const int* a = amax + static_cast<size_t>(b) * T;
  int* o = ids + static_cast<size_t>(b) * T;
  for (int t = 0; t < Tb; ++t) {
    const int tok = a[t];
    if (tok != 0 && tok != prev) o[n++] = tok;
    prev = tok;
  }
  counts[b] = n;
  for (int t = n; t < T; ++t) o[t] = 0;
}

int launch_greedy_collapse(const int* argmax, int B, int T, const int* logits_len, int* ids, int* counts, cudaStream_t stream) {
  (void)launch_dep(greedy_collapse_kernel, dim3(cdiv(B, 64)), dim3(64), 0, stream, argmax, B, T, logits_len, ids, counts);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec
