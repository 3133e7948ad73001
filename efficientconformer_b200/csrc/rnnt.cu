// Transducer joint network + RNN-T loss on the device (BASELINE.json configs[3]).
//
// Replaces, for the shipped `joint_mode: sum` / `act: tanh` joint (reference models/joint_networks.py:80-105) and the loss call
// warp_rnnt.rnnt_loss(log_softmax(logits), labels, frames_lengths, labels_lengths, average_frames=False, reduction='mean', blank=0)
// (reference models/losses.py:22-46):
//   joint_hidden   H[(b,t,u), :] = act_type(act(fe[b,t,:] + gd[b,u,:]))     fe = Linear_enc(f), gd = Linear_dec(g) come from the tcgen05
//                                                                          GEMM; the reference materialises the broadcast sum twice
//                                                                          in fp32 (`repeat` x 2), here it is written once as the
//                                                                          GEMM operand of the output projection
//   (logits = H W_joint^T + b on the tcgen05 GEMM, ec_op_gemm)
//   rnnt_lattice   per lattice node (b,t,u): log-sum-exp over the vocabulary, lp_blank = logit[blank] - lse, lp_label = logit[y_u] - lse:
//                  the only two numbers per node the loss reads (the reference's log_softmax writes all V of them)
//   rnnt_alpha     Graves 2012 eq. 16-18 forward variable over the (T, U+1) lattice, one CTA per utterance, anti-diagonal wavefront
//                  (thread = u), log domain in fp32;  nll_b = -(alpha(T_b-1, U_b) + lp_blank(T_b-1, U_b));  loss = mean_b nll_b
#include "ec_common.cuh"
#include <algorithm>

namespace ec {

enum JointAct { JOINT_ACT_NONE = 0, JOINT_ACT_TANH = 1, JOINT_ACT_RELU = 2, JOINT_ACT_SWISH = 3 };

template <typename T> __device__ __forceinline__ void store_act4(T* p, float a, float b, float c, float d);
template <> __device__ __forceinline__ void store_act4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(round_tf32(a), round_tf32(b), round_tf32(c), round_tf32(d));
}
template <> __device__ __forceinline__ void store_act4<SplitBf16>(SplitBf16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(split_pack(a), split_pack(b), split_pack(c), split_pack(d));
}
template <> __device__ __forceinline__ void store_act4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  __nv_bfloat162 x = __floats2bfloat162_rn(a, b), y = __floats2bfloat162_rn(c, d);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&x), *reinterpret_cast<uint32_t*>(&y));
}
__device__ __forceinline__ float joint_act(float x, int act) {
  if (act == JOINT_ACT_TANH) return tanhf(x);
  if (act == JOINT_ACT_RELU) return fmaxf(x, 0.f);
  if (act == JOINT_ACT_SWISH) return x / (1.f + __expf(-x));
  return x;
}

// thread = 4 consecutive hidden units of one lattice node; grid-stride over B*T*U1*J/4
template <typename T>
__global__ void __launch_bounds__(256) joint_hidden_kernel(const float* __restrict__ fe, const float* __restrict__ gd, int B, int Tn, int U1, int J,
                                                           int act, T* __restrict__ H) {
  const int J4 = J / 4;
  const size_t n = static_cast<size_t>(B) * Tn * U1 * J4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int j4 = static_cast<int>(i % J4);
    const size_t node = i / J4;
    const int u = static_cast<int>(node % U1);
    const size_t bt = node / U1;
    const int b = static_cast<int>(bt / Tn);
    const float4 a = *reinterpret_cast<const float4*>(fe + bt * J + 4 * j4);
    const float4 g = *reinterpret_cast<const float4*>(gd + (static_cast<size_t>(b) * U1 + u) * J + 4 * j4);
    store_act4<T>(H + node * J + 4 * j4, joint_act(a.x + g.x, act), joint_act(a.y + g.y, act), joint_act(a.z + g.z, act), joint_act(a.w + g.w, act));
  }
}

// warp = lattice node: log-sum-exp over V, then the blank and label log-probabilities
__global__ void __launch_bounds__(256) rnnt_lattice_kernel(const float* __restrict__ logits, size_t nodes, int Tn, int U1, int V,
                                                           const long long* __restrict__ labels, int label_stride, int blank,
                                                           float* __restrict__ lp_blank, float* __restrict__ lp_label, float* __restrict__ lse_out) {
  const int lane = threadIdx.x & 31;
  const size_t warps = static_cast<size_t>(gridDim.x) * (blockDim.x >> 5);
  for (size_t node = static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); node < nodes; node += warps) {
    const float* row = logits + node * V;
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, row[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(row[v] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const float lse = m + logf(s);
      const int u = static_cast<int>(node % U1);
      const size_t b = node / (static_cast<size_t>(Tn) * U1);
      lp_blank[node] = row[blank] - lse;
      if (lse_out != nullptr) lse_out[node] = lse;
      float ll = 0.f;
      if (u < U1 - 1) {
        const long long y = labels[b * label_stride + u];
        ll = (y >= 0 && y < V) ? row[y] - lse : -INFINITY;
      }
      lp_label[node] = ll;
    }
  }
}

__device__ __forceinline__ float log_add_exp(float a, float b) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  if (hi == -INFINITY) return -INFINITY;
  return hi + log1pf(expf(lo - hi));
}

// CTA = utterance; thread = label position u; anti-diagonal d = t + u.  alpha of the previous diagonal lives in shared memory.
__global__ void __launch_bounds__(1024) rnnt_alpha_kernel(const float* __restrict__ lp_blank, const float* __restrict__ lp_label, int Tn, int U1,
                                                          const int* __restrict__ frame_len, const int* __restrict__ label_len,
                                                          float* __restrict__ nll, float* __restrict__ alpha_out /* [B, T, U1] or nullptr */) {
  extern __shared__ float sh[];               // [2][U1]
  const int b = blockIdx.x;
  int Tb = frame_len[b], Ub = label_len[b];
  if (Tb > Tn) Tb = Tn;
  if (Ub > U1 - 1) Ub = U1 - 1;
  if (Tb <= 0 || Ub < 0) { if (threadIdx.x == 0) nll[b] = INFINITY; return; }
  const float* pb = lp_blank + static_cast<size_t>(b) * Tn * U1;
  const float* pl = lp_label + static_cast<size_t>(b) * Tn * U1;
  float* prev = sh;
  float* cur = sh + U1;
  for (int u = threadIdx.x; u < U1; u += blockDim.x) { prev[u] = -INFINITY; cur[u] = -INFINITY; }
  __syncthreads();
  if (threadIdx.x == 0) { prev[0] = 0.f; if (alpha_out != nullptr) alpha_out[static_cast<size_t>(b) * Tn * U1] = 0.f; }   // alpha(0, 0)
  __syncthreads();
  float last = (Tb == 1 && Ub == 0) ? 0.f : -INFINITY;      // alpha(T_b - 1, U_b)
  for (int d = 1; d <= Tb - 1 + Ub; ++d) {
    for (int u = threadIdx.x; u <= Ub; u += blockDim.x) {
      const int t = d - u;
      float a = -INFINITY;
      if (t >= 0 && t < Tb) {
        const float from_t = t > 0 ? prev[u] + pb[static_cast<size_t>(t - 1) * U1 + u] : -INFINITY;            // (t-1, u) -> blank
        const float from_u = u > 0 ? prev[u - 1] + pl[static_cast<size_t>(t) * U1 + u - 1] : -INFINITY;        // (t, u-1) -> label
        a = log_add_exp(from_t, from_u);
        if (t == Tb - 1 && u == Ub) last = a;
        if (alpha_out != nullptr) alpha_out[(static_cast<size_t>(b) * Tn + t) * U1 + u] = a;
      }
      cur[u] = a;
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
  // the thread that owned (T_b - 1, U_b) writes the result
  const int owner = Ub % blockDim.x;
  if (threadIdx.x == owner) nll[b] = -(last + pb[static_cast<size_t>(Tb - 1) * U1 + Ub]);
}

// backward variable: beta(T_b-1, U_b) = lp_blank(T_b-1, U_b);  beta(t,u) = logaddexp(beta(t+1,u) + lp_blank(t,u), beta(t,u+1) + lp_label(t,u))
// (Graves 2012 eq. 18), reverse anti-diagonal wavefront, stored for the gradient kernel
__global__ void __launch_bounds__(1024) rnnt_beta_kernel(const float* __restrict__ lp_blank, const float* __restrict__ lp_label, int Tn, int U1,
                                                         const int* __restrict__ frame_len, const int* __restrict__ label_len,
                                                         float* __restrict__ beta_out) {
  extern __shared__ float sh[];               // [2][U1]
  const int b = blockIdx.x;
  int Tb = frame_len[b], Ub = label_len[b];
  if (Tb > Tn) Tb = Tn;
  if (Ub > U1 - 1) Ub = U1 - 1;
  if (Tb <= 0 || Ub < 0) return;
  const float* pb = lp_blank + static_cast<size_t>(b) * Tn * U1;
  const float* pl = lp_label + static_cast<size_t>(b) * Tn * U1;
  float* bo = beta_out + static_cast<size_t>(b) * Tn * U1;
  float* prev = sh;                            // diagonal d + 1
  float* cur = sh + U1;
  for (int u = threadIdx.x; u < U1; u += blockDim.x) { prev[u] = -INFINITY; cur[u] = -INFINITY; }
  __syncthreads();
  const int dmax = Tb - 1 + Ub;
  if (threadIdx.x == 0) { const float v = pb[static_cast<size_t>(Tb - 1) * U1 + Ub]; prev[Ub] = v; bo[static_cast<size_t>(Tb - 1) * U1 + Ub] = v; }
  __syncthreads();
  for (int d = dmax - 1; d >= 0; --d) {
    for (int u = threadIdx.x; u <= Ub; u += blockDim.x) {
      const int t = d - u;
      float v = -INFINITY;
      if (t >= 0 && t < Tb) {
        const float to_t = t + 1 < Tb ? prev[u] + pb[static_cast<size_t>(t) * U1 + u] : -INFINITY;               // blank: (t,u) -> (t+1,u)
        const float to_u = u + 1 <= Ub ? prev[u + 1] + pl[static_cast<size_t>(t) * U1 + u] : -INFINITY;          // label: (t,u) -> (t,u+1)
        v = log_add_exp(to_t, to_u);
        bo[static_cast<size_t>(t) * U1 + u] = v;
      }
      cur[u] = v;
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
}

// d(mean nll) / d logits, log-softmax folded in (warp = lattice node):
//   grad[v] = scale * ( exp(c + logit_v - lse + beta(t,u)) - [v = blank] exp(c + lp_blank + beta(t+1,u)) - [v = y_u] exp(c + lp_label + beta(t,u+1)) )
// with c = alpha(t,u) + nll_b (= alpha / P in the log domain), beta(T_b, U_b) := 0 behind the final blank; zero outside the valid lattice.
__global__ void __launch_bounds__(256) rnnt_grad_kernel(const float* __restrict__ logits, size_t nodes, int Tn, int U1, int V,
                                                        const long long* __restrict__ labels, int label_stride, int blank,
                                                        const int* __restrict__ frame_len, const int* __restrict__ label_len,
                                                        const float* __restrict__ lp_blank, const float* __restrict__ lp_label,
                                                        const float* __restrict__ lse, const float* __restrict__ alpha, const float* __restrict__ beta,
                                                        const float* __restrict__ nll, float scale, float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const size_t warps = static_cast<size_t>(gridDim.x) * (blockDim.x >> 5);
  for (size_t node = static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); node < nodes; node += warps) {
    const int u = static_cast<int>(node % U1);
    const size_t bt = node / U1;
    const int t = static_cast<int>(bt % Tn), b = static_cast<int>(bt / Tn);
    int Tb = frame_len[b], Ub = label_len[b];
    if (Tb > Tn) Tb = Tn;
    if (Ub > U1 - 1) Ub = U1 - 1;
    float* g = grad + node * V;
    const float nl = nll[b];
    if (t >= Tb || u > Ub || !(nl < INFINITY)) {
      for (int v = lane; v < V; v += 32) g[v] = 0.f;
      continue;
    }
    const float c = alpha[node] + nl;
    const float bcur = beta[node];
    const float b_t = (t + 1 < Tb) ? beta[node + U1] : ((u == Ub) ? 0.f : -INFINITY);      // behind the blank transition
    const float b_u = (u + 1 <= Ub) ? beta[node + 1] : -INFINITY;                          // behind the label transition
    const float sub_blank = expf(c + lp_blank[node] + b_t);
    const float sub_label = (u < Ub) ? expf(c + lp_label[node] + b_u) : 0.f;
    const long long y = (u < Ub) ? labels[static_cast<size_t>(b) * label_stride + u] : -1;
    const float* row = logits + node * V;
    const float base = c + bcur - lse[node];
    for (int v = lane; v < V; v += 32) {
      float x = expf(base + row[v]);
      if (v == blank) x -= sub_blank;
      if (v == y) x -= sub_label;
      g[v] = scale * x;
    }
  }
}

// gradient through H = act(fe + gd) and the two broadcast sums: dpre = dH * act'(H);  dfe[b,t,:] = sum_u dpre,  dgd[b,u,:] = sum_t dpre.
// act' from the stored activation output: tanh 1 - H^2, relu [H > 0], none 1.  Thread = 4 hidden units of one (b,t) row / one (b,u) row.
template <typename T> __device__ __forceinline__ float4 load_act4(const T* p);
template <> __device__ __forceinline__ float4 load_act4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 load_act4<SplitBf16>(const SplitBf16* p) {
  const uint4 w = *reinterpret_cast<const uint4*>(p);
  return make_float4(split_unpack(w.x), split_unpack(w.y), split_unpack(w.z), split_unpack(w.w));
}
template <> __device__ __forceinline__ float4 load_act4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 w = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xffff0000u));
}
__device__ __forceinline__ float act_grad_from_output(float h, int act) {
  if (act == JOINT_ACT_TANH) return 1.f - h * h;
  if (act == JOINT_ACT_RELU) return h > 0.f ? 1.f : 0.f;
  return 1.f;
}
template <typename T, bool kOverU>
__global__ void __launch_bounds__(256) joint_hidden_bwd_kernel(const T* __restrict__ H, const float* __restrict__ dH, int B, int Tn, int U1, int J,
                                                               int act, float* __restrict__ out) {
  const int J4 = J / 4;
  const size_t rows = static_cast<size_t>(B) * (kOverU ? Tn : U1);
  const size_t n = rows * J4, stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int j4 = static_cast<int>(i % J4);
    const size_t r = i / J4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int steps = kOverU ? U1 : Tn;
    for (int k = 0; k < steps; ++k) {
      size_t node;
      if (kOverU) node = r * U1 + k;                                  // r = b*T + t, k = u
      else { const size_t b = r / U1, u = r - b * U1; node = (b * Tn + k) * U1 + u; }   // r = b*U1 + u, k = t
      const float4 h = load_act4<T>(H + node * J + 4 * j4);
      const float4 d = *reinterpret_cast<const float4*>(dH + node * J + 4 * j4);
      acc.x = fmaf(d.x, act_grad_from_output(h.x, act), acc.x); acc.y = fmaf(d.y, act_grad_from_output(h.y, act), acc.y);
      acc.z = fmaf(d.z, act_grad_from_output(h.z, act), acc.z); acc.w = fmaf(d.w, act_grad_from_output(h.w, act), acc.w);
    }
    *reinterpret_cast<float4*>(out + r * J + 4 * j4) = acc;
  }
}

__global__ void mean_kernel_rnnt(const float* __restrict__ x, int n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { float s = 0.f; for (int i = 0; i < n; ++i) s += x[i]; out[0] = s / n; }
}

}  // namespace ec

using namespace ec;
#define EC_ST(s) reinterpret_cast<cudaStream_t>(s)
extern "C" {
int ec_op_joint_hidden(int precision, const float* fe, const float* gd, int batch, int t, int u1, int dim_joint, int act, void* hidden, void* stream) {
  EC_REQUIRE(fe && gd && hidden && batch > 0 && t > 0 && u1 > 0 && dim_joint > 0, "bad argument");
  EC_REQUIRE(dim_joint % 4 == 0 && act >= 0 && act <= 3, "joint: hidden width must be a multiple of 4; act in {none, tanh, relu, swish}");
  const size_t n = static_cast<size_t>(batch) * t * u1 * (dim_joint / 4);
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
  EC_DISPATCH_PREC(precision, (joint_hidden_kernel<ActT><<<grid, 256, 0, EC_ST(stream)>>>(fe, gd, batch, t, u1, dim_joint, act, static_cast<ActT*>(hidden))));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
/* logits [B, T, U1, V] fp32 -> per-utterance negative log likelihoods and their mean.  scratch: ec_rnnt_scratch_bytes. */
size_t ec_rnnt_scratch_bytes(int batch, int t, int u1) {
  return align_up(static_cast<size_t>(5) * batch * t * u1 * sizeof(float), 256) + align_up(static_cast<size_t>(2) * batch * sizeof(int), 256);
}
static int rnnt_run(const float* logits, int batch, int t, int u1, int vocab, const long long* labels, int label_stride, const long long* frame_len,
                    const long long* label_len, int blank, void* scratch, float* loss_per_utt, float* loss_mean, float grad_scale, float* grad,
                    cudaStream_t st) {
  EC_REQUIRE(logits && labels && frame_len && label_len && scratch && loss_per_utt, "null argument");
  EC_REQUIRE(batch > 0 && t > 0 && u1 > 0 && vocab > 0 && blank >= 0 && blank < vocab && label_stride >= u1 - 1, "bad RNN-T shapes");
  const size_t nodes = static_cast<size_t>(batch) * t * u1;
  float* lp_blank = reinterpret_cast<float*>(scratch);
  float* lp_label = lp_blank + nodes;
  float* lse = lp_label + nodes;
  float* alpha = lse + nodes;
  float* beta = alpha + nodes;
  int* fl = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(scratch) + align_up(5 * nodes * sizeof(float), 256));
  int* ll = fl + batch;
  EC_TRY(launch_i64_to_i32(frame_len, batch, fl, t, st));
  EC_TRY(launch_i64_to_i32(label_len, batch, ll, u1 - 1, st));
  const int grid = static_cast<int>(std::min<size_t>((nodes + 7) / 8, 148 * 32));
  const bool want_grad = grad != nullptr;
  rnnt_lattice_kernel<<<grid, 256, 0, st>>>(logits, nodes, t, u1, vocab, labels, label_stride, blank, lp_blank, lp_label, want_grad ? lse : nullptr);
  EC_CUDA(cudaGetLastError());
  const int threads = std::min(1024, round_up(u1, 32));
  const size_t sm = 2 * static_cast<size_t>(u1) * sizeof(float);
  rnnt_alpha_kernel<<<batch, threads, sm, st>>>(lp_blank, lp_label, t, u1, fl, ll, loss_per_utt, want_grad ? alpha : nullptr);
  EC_CUDA(cudaGetLastError());
  if (loss_mean != nullptr) {
    mean_kernel_rnnt<<<1, 32, 0, st>>>(loss_per_utt, batch, loss_mean);
    EC_CUDA(cudaGetLastError());
  }
  if (want_grad) {
    rnnt_beta_kernel<<<batch, threads, sm, st>>>(lp_blank, lp_label, t, u1, fl, ll, beta);
    EC_CUDA(cudaGetLastError());
    rnnt_grad_kernel<<<grid, 256, 0, st>>>(logits, nodes, t, u1, vocab, labels, label_stride, blank, fl, ll, lp_blank, lp_label, lse, alpha, beta,
                                           loss_per_utt, grad_scale, grad);
    EC_CUDA(cudaGetLastError());
  }
  return EC_OK;
}
int ec_rnnt_loss(const float* logits, int batch, int t, int u1, int vocab, const long long* labels, int label_stride, const long long* frame_len,
                 const long long* label_len, int blank, void* scratch, float* loss_per_utt, float* loss_mean, void* stream) {
  return rnnt_run(logits, batch, t, u1, vocab, labels, label_stride, frame_len, label_len, blank, scratch, loss_per_utt, loss_mean, 0.f, nullptr,
                  EC_ST(stream));
}
/* same, plus grad [B, T, U1, V] = grad_scale * d(sum_b nll_b) / d logits (log-softmax included); pass grad_scale = 1 / B for the mean */
int ec_rnnt_loss_grad(const float* logits, int batch, int t, int u1, int vocab, const long long* labels, int label_stride,
                      const long long* frame_len, const long long* label_len, int blank, void* scratch, float* loss_per_utt, float* loss_mean,
                      float grad_scale, float* grad, void* stream) {
  EC_REQUIRE(grad != nullptr, "null gradient buffer");
  return rnnt_run(logits, batch, t, u1, vocab, labels, label_stride, frame_len, label_len, blank, scratch, loss_per_utt, loss_mean, grad_scale, grad,
                  EC_ST(stream));
}
/* gradient through hidden = act(fe + gd) and the broadcast sums: dfe [B*T, J] = sum_u dH * act'(hidden), dgd [B*U1, J] = sum_t ... */
int ec_op_joint_hidden_bwd(int precision, const void* hidden, const float* d_hidden, int batch, int t, int u1, int dim_joint, int act, float* dfe,
                           float* dgd, void* stream) {
  EC_REQUIRE(hidden && d_hidden && dfe && dgd && batch > 0 && t > 0 && u1 > 0, "bad argument");
  EC_REQUIRE(dim_joint % 4 == 0 && (act == JOINT_ACT_NONE || act == JOINT_ACT_TANH || act == JOINT_ACT_RELU),
             "joint backward: hidden width multiple of 4; act in {none, tanh, relu} (the derivative is taken from the stored output)");
  cudaStream_t st = EC_ST(stream);
  const size_t n1 = static_cast<size_t>(batch) * t * (dim_joint / 4), n2 = static_cast<size_t>(batch) * u1 * (dim_joint / 4);
  const int g1 = static_cast<int>(std::min<size_t>((n1 + 255) / 256, 148 * 16)), g2 = static_cast<int>(std::min<size_t>((n2 + 255) / 256, 148 * 16));
  EC_DISPATCH_PREC(precision, (joint_hidden_bwd_kernel<ActT, true><<<g1, 256, 0, st>>>(static_cast<const ActT*>(hidden), d_hidden, batch, t, u1, dim_joint, act, dfe)));
  EC_CUDA(cudaGetLastError());
  EC_DISPATCH_PREC(precision, (joint_hidden_bwd_kernel<ActT, false><<<g2, 256, 0, st>>>(static_cast<const ActT*>(hidden), d_hidden, batch, t, u1, dim_joint, act, dgd)));
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
}
