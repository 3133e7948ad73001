// Gradient of the CTC loss with respect to the logits (first kernel of the training backward pass).
//
// Restates what autograd computes for the reference's LossCTC.forward (reference models/losses.py:56-71):
//     loss = mean_b nll_b,   nll_b = nn.CTCLoss(blank=0, reduction='none', zero_infinity=False)(log_softmax(logits), ...)
// For utterance b, frame t < T_b and class c (Graves et al. 2006, eq. 14-16; the alpha-beta product includes the emission twice):
//     gamma_t(s) = exp(alpha_t(s) + beta_t(s) - lp_t(l'_s) + nll_b)                       posterior occupancy of extended state s
//     d loss / d logits[b,t,c] = ( softmax(logits[b,t])[c] - sum_{s : l'_s = c} gamma_t(s) ) / B
// and zero for the padded frames t >= T_b.  One CTA per utterance: the whole CTA gathers the emission log-probs of the extended
// labels into shared memory (log2 domain), warp 0 runs the alpha recursion forward (alpha rows go to an L2-resident scratch:
// stores are off the dependency chain) and the beta recursion backward (the alpha row of the next step is prefetched), leaving
// gamma in the scratch; then all warps write the gradient rows: softmax / B, minus the occupancies, scattered deterministically
// (blank: fixed-order warp reduction; labels: the first occurrence of a label walks the chain of its repeats).
#include "ec_common.cuh"
#include <algorithm>

namespace ec {

namespace {
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kThreads = 256;

__device__ __forceinline__ float lse3_log2(float a, float b, float c) {     // see ctc.cu
  const float hi = fmaxf(a, b), o2 = fminf(a, b);
  const float m = fmaxf(hi, c), o1 = fminf(hi, c);
  const float ms = (m == -INFINITY) ? 0.f : m;
  float e1, e2, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(o1 - ms));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(o2 - ms));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((m == -INFINITY ? 0.f : 1.f) + e1 + e2));
  return ms + r;
}
__device__ __forceinline__ float ex2f(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
}  // namespace

template <int NS>
__global__ void __launch_bounds__(kThreads) ctc_grad_kernel(const float* __restrict__ logits, const float* __restrict__ lse, int T, int V,
                                                            const int* __restrict__ logits_len, const long long* __restrict__ targets,
                                                            int target_stride, const long long* __restrict__ target_len,
                                                            float* __restrict__ work,        // [B][T][32*NS] alpha, then gamma
                                                            float* __restrict__ work_lp,     // [B][T][32*NS] emissions when they do not fit on chip
                                                            float grad_scale, float* __restrict__ loss_per_utt, float* __restrict__ grad) {
  grid_dependency_wait();
  grid_launch_dependents();
  extern __shared__ float lp_dyn[];                // [Tb][SP] emission log2-probs of the extended labels (or unused: work_lp)
  constexpr int SP = 32 * NS;
  // long utterances x long transcripts (T * (2U+1) floats beyond the shared-memory budget) keep the emissions in an L2-resident scratch
  float* __restrict__ lp_sm = work_lp != nullptr ? work_lp + static_cast<size_t>(blockIdx.x) * T * SP : lp_dyn;
  __shared__ int lab[SP];                          // extended labels l'_s (0 = blank)
  __shared__ int nxt_same[SP / 2 + 1];             // u -> next u' > u with the same label, or -1
  __shared__ unsigned char is_first[SP / 2 + 1];
  __shared__ float nll_sm;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int U = static_cast<int>(target_len[b]);
  const int S = 2 * U + 1;
  int Tb = logits_len[b];
  if (Tb > T) Tb = T;
  const long long* y = targets + static_cast<size_t>(b) * target_stride;
  const float* lg = logits + static_cast<size_t>(b) * T * V;
  const float* ls = lse + static_cast<size_t>(b) * T;
  float* wk = work + static_cast<size_t>(b) * T * SP;
  float* gr = grad + static_cast<size_t>(b) * T * V;
  if (Tb <= 0) {                                   // empty utterance: infinite loss, zero gradient
    if (tid == 0) loss_per_utt[b] = INFINITY;
    for (int i = tid; i < T * V; i += kThreads) gr[i] = 0.f;
    return;
  }
  for (int s = tid; s < SP; s += kThreads) lab[s] = (s < S && (s & 1)) ? static_cast<int>(y[s >> 1]) : 0;
  for (int u = tid; u < U; u += kThreads) {
    const long long me = y[u];
    int nx = -1; bool first = true;
    for (int v = u + 1; v < U; ++v) if (y[v] == me) { nx = v; break; }
    for (int v = 0; v < u; ++v) if (y[v] == me) { first = false; break; }
    nxt_same[u] = nx; is_first[u] = first ? 1 : 0;
  }
  __syncthreads();
  for (int idx = tid; idx < Tb * SP; idx += kThreads) {
    const int t = idx / SP, s = idx - t * SP;
    lp_sm[t * SP + (s % NS) * 32 + s / NS] = kLog2e * (__ldg(lg + static_cast<size_t>(t) * V + lab[s]) - __ldg(ls + t));
  }
  __syncthreads();

  if (warp == 0) {
    // state s = lane * NS + k lives in register k of lane `lane`, column k * 32 + lane of a row
    bool skip[NS], skip_from[NS];                  // alpha: s-2 -> s allowed;  beta: s -> s+2 allowed
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const int s = lane * NS + k;
      const int e = s < S ? lab[s] : 0;
      skip[k] = s >= 2 && s < S && e != 0 && e != lab[s - 2];
      skip_from[k] = s + 2 < S && lab[s + 2] != 0 && lab[s + 2] != e;
    }
    float a[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const int s = lane * NS + k;
      a[k] = (s < 2 && s < S) ? lp_sm[k * 32 + lane] : -INFINITY;
      wk[k * 32 + lane] = a[k];
    }
    float cur[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) cur[k] = Tb > 1 ? lp_sm[SP + k * 32 + lane] : 0.f;
    for (int t = 1; t < Tb; ++t) {
      float nx[NS];
      const int tn = t + 1 < Tb ? t + 1 : t;
#pragma unroll
      for (int k = 0; k < NS; ++k) nx[k] = lp_sm[tn * SP + k * 32 + lane];
      float pm1 = __shfl_up_sync(0xffffffffu, a[NS - 1], 1);
      float pm2 = NS >= 2 ? __shfl_up_sync(0xffffffffu, a[NS >= 2 ? NS - 2 : 0], 1) : __shfl_up_sync(0xffffffffu, a[0], 2);
      if (lane == 0) { pm1 = -INFINITY; pm2 = -INFINITY; }
      if (NS == 1 && lane == 1) pm2 = -INFINITY;
      float na[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        const float x1 = k >= 1 ? a[k - 1] : pm1;
        const float x2 = skip[k] ? (k >= 2 ? a[k - 2] : (k == 1 ? pm1 : pm2)) : -INFINITY;
        const float acc = lse3_log2(a[k], x1, x2);
        na[k] = (lane * NS + k >= S) ? -INFINITY : acc + cur[k];
      }
#pragma unroll
      for (int k = 0; k < NS; ++k) { a[k] = na[k]; cur[k] = nx[k]; wk[static_cast<size_t>(t) * SP + k * 32 + lane] = na[k]; }
    }
    float e1 = -INFINITY, e2 = -INFINITY;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const int s = lane * NS + k;
      if (s == S - 1) e1 = a[k];
      if (s == S - 2) e2 = a[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { e1 = fmaxf(e1, __shfl_xor_sync(0xffffffffu, e1, o)); e2 = fmaxf(e2, __shfl_xor_sync(0xffffffffu, e2, o)); }
    const float nll2 = -lse3_log2(e1, e2, -INFINITY);          // -log2 p(l | x)
    if (lane == 0) { loss_per_utt[b] = nll2 * 0.6931471805599453f; nll_sm = nll2; }

    // ---- beta backward; gamma_t(s) = 2^(alpha + beta - lp + nll2) replaces alpha_t(s) in the scratch ----
    float be[NS], al[NS], lpc[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const int s = lane * NS + k;
      lpc[k] = lp_sm[(Tb - 1) * SP + k * 32 + lane];
      be[k] = (s < S && s >= S - 2) ? lpc[k] : -INFINITY;
      al[k] = a[k];                                             // alpha of the last frame is still in registers
    }
    for (int t = Tb - 1; t >= 0; --t) {
      float al_n[NS], lp_n[NS];
      const int tp = t > 0 ? t - 1 : 0;
#pragma unroll
      for (int k = 0; k < NS; ++k) {                            // prefetch row t-1 (alpha from the scratch this warp wrote itself)
        al_n[k] = wk[static_cast<size_t>(tp) * SP + k * 32 + lane];
        lp_n[k] = lp_sm[tp * SP + k * 32 + lane];
      }
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        const float gm = ex2f(al[k] + be[k] - lpc[k] + nll2);
        wk[static_cast<size_t>(t) * SP + k * 32 + lane] = (lane * NS + k < S && be[k] != -INFINITY && al[k] != -INFINITY) ? gm : 0.f;
      }
      if (t == 0) break;
      float nb1 = __shfl_down_sync(0xffffffffu, be[0], 1);
      float nb2 = NS >= 2 ? __shfl_down_sync(0xffffffffu, be[NS >= 2 ? 1 : 0], 1) : __shfl_down_sync(0xffffffffu, be[0], 2);
      if (lane == 31) { nb1 = -INFINITY; nb2 = -INFINITY; }
      if (NS == 1 && lane == 30) nb2 = -INFINITY;
      float nbv[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        const float x1 = k + 1 < NS ? be[k + 1] : nb1;
        const float x2 = skip_from[k] ? (k + 2 < NS ? be[k + 2] : (k + 2 == NS ? nb1 : nb2)) : -INFINITY;
        const float acc = lse3_log2(be[k], x1, x2);
        nbv[k] = (lane * NS + k >= S) ? -INFINITY : acc + lp_n[k];
      }
#pragma unroll
      for (int k = 0; k < NS; ++k) { be[k] = nbv[k]; al[k] = al_n[k]; lpc[k] = lp_n[k]; }
    }
    __threadfence_block();
  }
  __syncthreads();

  // ---- gradient rows: softmax / B everywhere, minus the occupancies of the classes that appear in l' ----
  for (int i = tid; i < T * V; i += kThreads) {
    const int t = i / V;
    gr[i] = t < Tb ? grad_scale * __expf(__ldg(lg + i) - __ldg(ls + t)) : 0.f;
  }
  __syncthreads();
  for (int t = warp; t < Tb; t += kThreads / 32) {
    const float* gm = wk + static_cast<size_t>(t) * SP;
    float* row = gr + static_cast<size_t>(t) * V;
    float blank = 0.f;                                          // even states, fixed summation order
    for (int s = 2 * lane; s < S; s += 64) blank += gm[(s % NS) * 32 + s / NS];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) blank += __shfl_xor_sync(0xffffffffu, blank, o);
    if (lane == 0) row[0] -= grad_scale * blank;
    for (int u = lane; u < U; u += 32) {
      if (!is_first[u]) continue;
      float tot = 0.f;
      for (int v = u; v >= 0; v = nxt_same[v]) { const int s = 2 * v + 1; tot += gm[(s % NS) * 32 + s / NS]; }
      row[lab[2 * u + 1]] -= grad_scale * tot;
    }
  }
}

constexpr size_t kCtcSmemBudget = 200 * 1024;
size_t ctc_grad_work_bytes(int B, int T, int target_stride) {
  const int ns = std::max(cdiv(2 * target_stride + 1, 32), 1);
  const size_t plane = align_up(static_cast<size_t>(B) * T * 32 * ns * sizeof(float), 256);
  const bool global_lp = static_cast<size_t>(T) * 32 * ns * sizeof(float) > kCtcSmemBudget;
  return plane * (global_lp ? 2 : 1);
}

int launch_ctc_grad(const float* logits, const float* lse, int B, int T, int V, const int* logits_len, const long long* targets,
                    int target_stride, const long long* target_len, float* work, float grad_scale, float* loss_per_utt, float* grad,
                    cudaStream_t stream) {
  EC_REQUIRE(B > 0 && T > 0 && V > 0 && target_stride >= 0, "bad CTC shapes");
  const int ns = std::max(cdiv(2 * target_stride + 1, 32), 1);
  // 16 extended-label states per lane = transcripts of up to 255 labels (nn.CTCLoss needs U <= T, and T_out <= 201 for the 16 s
  // utterances of the shipped configs: train_audio_max_length 256000 samples)
  EC_REQUIRE(ns <= 16, "CTC gradient supports targets of up to 255 labels");
#define EC_CTC_GRAD(NS)                                                                                                              \
  case NS: {                                                                                                                         \
    size_t sm = static_cast<size_t>(T) * 32 * NS * sizeof(float);                                                                    \
    float* work_lp = nullptr;                                                                                                        \
    if (sm > kCtcSmemBudget) { work_lp = work + align_up(static_cast<size_t>(B) * T * 32 * NS * sizeof(float), 256) / sizeof(float); sm = 0; } \
    static cudaError_t attr = cudaFuncSetAttribute(ctc_grad_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kCtcSmemBudget)); \
    EC_CUDA(attr);                                                                                                                   \
    (void)launch_dep(ctc_grad_kernel<NS>, dim3(B), dim3(kThreads), sm, stream, logits, lse, T, V, logits_len, targets, target_stride, target_len, work, work_lp, \
                                                     grad_scale, loss_per_utt, grad);                                                \
    break;                                                                                                                           \
  }
  switch (ns) {
    EC_CTC_GRAD(1) EC_CTC_GRAD(2) EC_CTC_GRAD(3) EC_CTC_GRAD(4) EC_CTC_GRAD(5) EC_CTC_GRAD(6) EC_CTC_GRAD(7) EC_CTC_GRAD(8)
    EC_CTC_GRAD(9) EC_CTC_GRAD(10) EC_CTC_GRAD(11) EC_CTC_GRAD(12) EC_CTC_GRAD(13) EC_CTC_GRAD(14) EC_CTC_GRAD(15) EC_CTC_GRAD(16)
  }
#undef EC_CTC_GRAD
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec
