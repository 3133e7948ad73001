// Shared declarations of the effconf_b200 CUDA library (internal).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ec_ptx.cuh"
#include "../../include/effconf_b200.h"

namespace ec {

// ---- error convention: C-ABI functions return 0 on success; message retrievable with ec_last_error() ----------
void set_error(const std::string& msg);
#define EC_FAIL(msg)                                                                    \
  do {                                                                                  \
    ::ec::set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
    return EC_ERR;                                                                      \
  } while (0)
#define EC_REQUIRE(cond, msg) \
  do {                        \
    if (!(cond)) EC_FAIL(msg); \
  } while (0)
#define EC_CUDA(call)                                                  \
  do {                                                                 \
    cudaError_t _e = (call);                                           \
    if (_e != cudaSuccess) EC_FAIL(std::string("CUDA: ") + cudaGetErrorString(_e)); \
  } while (0)
#define EC_TRY(call)              \
  do {                            \
    int _s = (call);              \
    if (_s != EC_OK) return _s;   \
  } while (0)

// Kernel launch with the programmatic-stream-serialization attribute (PDL): the kernel may start while its predecessor
// drains; every kernel of the forward chain executes griddepcontrol.wait before touching predecessor output.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  EC_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
  return EC_OK;
}

// The same attribute for the kernels of the TRAINING step (element / reduction / batched-GEMM kernels of the backward and the train-mode
// forward).  Every one of them executes griddepcontrol.wait as its first statement (no read or write of global memory before it) and
// then griddepcontrol.launch_dependents, so its successor's launch and prologue overlap this kernel's run and only the successor's own
// wait orders the data.  EFFCONF_PDL_TRAIN=0 launches them as plain stream-ordered kernels (A/B measurement).
bool pdl_train_enabled();
template <typename... KArgs, typename... Args>
int launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (pdl_enabled() && pdl_train_enabled()) ? 1 : 0;
  EC_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
  return EC_OK;
}

// Same, as thread-block clusters of `cluster_x` CTAs along x.
template <typename... KArgs, typename... Args>
int launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  EC_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
  return EC_OK;
}

// Fork / join of independent launches onto library-owned side streams (non-blocking, created once per process): `fork(st)` makes the
// side streams wait for everything enqueued on `st` so far, `join(st, n)` makes `st` wait for the first n side streams.  Both are
// event record / wait pairs, so they are capturable: inside a CUDA graph the launches become parallel branches.  After `join` all
// side work is ordered before whatever the caller enqueues on `st` next (allocator-safe: the caller's stream semantics are preserved).
struct SideStreams {
  static constexpr int kN = 4;
  cudaStream_t s[kN]{};
  cudaEvent_t fork_ev{}, join_ev[kN]{};
  bool ok = false;
  bool init();
  bool fork(cudaStream_t st, int n);
  bool join(cudaStream_t st, int n);
};
SideStreams& side_streams();
bool side_streams_enabled();

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return cdiv(a, b) * b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- activation storage type: float (TF32 mode, values pre-rounded to 10 mantissa bits) or bf16 ----------------
template <typename T> struct ActTraits;
template <> struct ActTraits<float> {
  static constexpr bool kTf32 = true;
  static constexpr int kBlockK = 32;  // elements per 128-byte swizzle row
  static constexpr int kUmmaK = 8;
  __device__ static float to(float x) { return round_tf32(x); }
  __device__ static float from(float x) { return x; }
  __device__ static float word(float x) { return round_tf32(x); }     // the stored 32-bit word, carried in a float register
};
template <> struct ActTraits<__nv_bfloat16> {
  static constexpr bool kTf32 = false;
  static constexpr int kBlockK = 64;
  static constexpr int kUmmaK = 16;
  __device__ static __nv_bfloat16 to(float x) { return __float2bfloat16_rn(x); }
  __device__ static float from(__nv_bfloat16 x) { return __bfloat162float(x); }
};

// Split mode (EC_PREC_BF16X2): one element = the pair hi = bf16(x), lo = bf16(x - hi) in one 32-bit word (low half hi), 16
// significant bits.  Same 4-byte footprint, TMA boxes, slab layouts and workspace sizes as the TF32 mode; the tensor core reads a
// row of K packed elements as 2K bf16 values (kind::f16) against the two planes of the weight operand (see effconf_b200.h).
struct SplitBf16 { uint32_t bits; };
__device__ __forceinline__ uint32_t split_pack(float x) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  return static_cast<uint32_t>(__bfloat16_as_ushort(h)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l)) << 16);
}
__device__ __forceinline__ float split_unpack(uint32_t b) { return __uint_as_float(b << 16) + __uint_as_float(b & 0xffff0000u); }
__device__ __forceinline__ uint32_t split_swap(uint32_t b) { return __byte_perm(b, 0, 0x1032); }
template <> struct ActTraits<SplitBf16> {
  static constexpr bool kTf32 = false;
  static constexpr int kBlockK = 32;  // packed elements per 128-byte swizzle row (= 64 bf16 values)
  static constexpr int kUmmaK = 16;
  __device__ static SplitBf16 to(float x) { return SplitBf16{split_pack(x)}; }
  __device__ static float word(float x) { return __uint_as_float(split_pack(x)); }
  __device__ static float from(SplitBf16 x) { return split_unpack(x.bits); }
};
// fp16: the q|k|v / E operands of the attention core in split mode (11 significant bits = TF32 grade at the bf16 MMA rate; the values
// are O(1) activations, far inside the fp16 range)
template <> struct ActTraits<__half> {
  static constexpr bool kTf32 = false;
  __device__ static __half to(float x) { return __float2half_rn(x); }
  __device__ static float from(__half x) { return __half2float(x); }
};
template <typename T> struct IsSplit { static constexpr bool value = false; };
template <> struct IsSplit<SplitBf16> { static constexpr bool value = true; };
// bytes of one activation element / factor of the weight operand ([2, N, K] in split mode)
inline size_t act_esize(int precision) { return precision == EC_PREC_BF16 ? 2 : 4; }
inline size_t weight_planes(int precision) { return precision == EC_PREC_BF16X2 ? 2 : 1; }

// `body` sees the activation storage type of `precision` as ActT
#define EC_DISPATCH_PREC(precision, ...)                                                  \
  do {                                                                                    \
    if ((precision) == EC_PREC_TF32) { using ActT = float; __VA_ARGS__; }                    \
    else if ((precision) == EC_PREC_BF16) { using ActT = __nv_bfloat16; __VA_ARGS__; }       \
    else if ((precision) == EC_PREC_BF16X2) { using ActT = ::ec::SplitBf16; __VA_ARGS__; }   \
    else EC_FAIL("unknown precision");                                                    \
  } while (0)

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// x * sigmoid(x).  Parity (TF32) mode: exp + reciprocal.  bf16 mode: 0.5x + 0.5x*tanh(0.5x) with the hardware tanh
// (1 MUFU instead of 2; its 2^-11 relative error is below the bf16 rounding that follows).
template <typename T>
__device__ __forceinline__ float swish_fn(float x) {
  if constexpr (sizeof(T) == 4) {
    return x * fast_sigmoid(x);
  } else {
    const float h = 0.5f * x;
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
    return fmaf(h, th, h);
  }
}
template <typename T>
__device__ __forceinline__ float sigmoid_fn(float x) {
  if constexpr (sizeof(T) == 4) {
    return fast_sigmoid(x);
  } else {
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * x));
    return fmaf(0.5f, th, 0.5f);
  }
}
// ---- counter-based dropout (train_step.cu; also applied inside the GEMM epilogues of the training step) -------------------------
// The keep bit of element i at site s of step n is a pure function of (seed, n, s, i): ctr = {seed, step} lives in device memory, so a
// replayed CUDA graph draws fresh masks and the backward recomputes the forward mask instead of storing it.  One 64-bit draw covers the
// 4 consecutive elements 4g .. 4g+3 (16 bits each): keep iff bits < keep16.
// (count, mean, M2) <- (count, mean, M2) (+) (nb, mb, qb): Chan's parallel update, single precision
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb, float qb) {
  if (nb == 0.f) return;
  const float tot = n + nb, dl = mb - mean, f = nb / tot;
  mean = fmaf(dl, f, mean);
  m2 += qb + dl * dl * n * f;
  n = tot;
}
// 32 per-lane-row (count, mean, M2) triples per column (block = 32 columns x 32 rows, arrays [32][33]) merged as a fixed binary tree
// in shared memory: 5 dependent merges instead of a 32-step chain (which, in double precision with two divisions per step, was 70 % of
// the merge kernels' time); the result is left in row 0.  Fixed order: bit-reproducible.
__device__ __forceinline__ void chan_tree_merge_32x32(float (*sn)[33], float (*smean)[33], float (*sm2)[33]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    __syncthreads();
    if ((ty & (2 * off - 1)) == 0) {
      float n = sn[ty][tx], mean = smean[ty][tx], m2 = sm2[ty][tx];
      chan_merge(n, mean, m2, sn[ty + off][tx], smean[ty + off][tx], sm2[ty + off][tx]);
      sn[ty][tx] = n; smean[ty][tx] = mean; sm2[ty][tx] = m2;
    }
  }
  __syncthreads();
}
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned long long site_key(const unsigned long long* ctr, unsigned site) {
  return splitmix64(ctr[0] ^ (ctr[1] * 0xD1B54A32D192ED03ull)) ^ (static_cast<unsigned long long>(site) * 0x9FB21C651E98DF25ull);
}
__device__ __forceinline__ float keep_factor(unsigned long long draw, int lane, unsigned keep16, float inv_keep) {
  return ((draw >> (16 * lane)) & 0xFFFFull) < keep16 ? inv_keep : 0.f;
}
inline unsigned keep16_of(float p) {
  double keep = (1.0 - static_cast<double>(p)) * 65536.0 + 0.5;
  keep = keep < 1.0 ? 1.0 : (keep > 65536.0 ? 65536.0 : keep);
  return static_cast<unsigned>(keep);
}

// ---- kernel launch entry points (defined in the .cu files; all stream-ordered, never synchronise) --------------
enum GemmAct { GEMM_ACT_NONE = 0, GEMM_ACT_SWISH = 1 };

struct GemmArgs {
  const void* A;        // [M, K] row-major, activation type
  const void* W;        // [N, K] row-major (nn.Linear layout), activation type; split mode: plane 0 of the [2, N, K] operand
  size_t w_twin_bytes;  // split mode: byte distance from W to its swapped plane (0: N*K*4, i.e. a contiguous [2, N, K])
  int M, N, K;
  const float* bias;    // [N] (GLU: prepared interleaved order) or nullptr
  float alpha;          // out = alpha * act(acc + bias) + residual
  int act;              // GemmAct
  int glu_nb;           // 0: plain.  >0: W rows are tiles of [nb value rows | nb gate rows]; output channel count = glu_channels
  int glu_channels;
  const float* residual; int ld_res;   // fp32 [M, *] or nullptr
  float* out_f32; int ld_out;          // optional fp32 output
  void* out_act; int ld_act;           // optional activation-type output (rounded)
  int act_f16;                         // split mode only: out_act is written as PLAIN fp16 (q|k|v and E feed the 16-bit attention kernels)
  int round_out;                       // round the fp32 output to TF32 (feeds the TF32 mma.sync attention)
  // fused LayerNorm epilogue (N <= 256, plain epilogue): mode 1: ln_out = LN1(out); mode 2: out <- LN1(out), ln_out = LN2(out)
  // (LN2 = identity copy when ln2_g == nullptr).  copy_out: activation-type copy of every copy_stride-th frame of `out` (mode 1).
  int ln_mode; const float *ln1_g, *ln1_b, *ln2_g, *ln2_b; float ln_eps;
  void* ln_out;
  void* copy_out; int copy_stride, frames_per_seq, frames_out_per_seq;
  // training-step epilogue options (plain variant only; element index of the masks = row * N + column, N % 4 == 0)
  const unsigned long long* drop_ctr;  // device {seed, step} of the counter-based dropout (nullptr: no dropout anywhere in this launch)
  float drop_p;
  unsigned drop_site;                  // > 0: out = alpha * keep / (1 - p) * act(acc + bias) [+ residual]  (nn.Dropout after the projection)
  void* out_act2; int ld_act2;         // second activation-type output h = dropout_{site2}(Swish(z)), z = out_act as stored (pre-activation)
  unsigned drop_site2;
  const void* aux_act; int aux_mode;   // aux_mode 1: acc <- acc * keep_{site_aux} / (1 - p) * Swish'(aux): data gradient through dropout(Swish(z)),
  unsigned drop_site_aux;              //             aux = the saved pre-activation z [M, N] (activation type); excludes `residual`
};
int launch_gemm(int precision, const GemmArgs& a, cudaStream_t stream);
int gemm_timeline(int enable, unsigned long long* out12);
int gemm_block_n_override(int block_n);

// Fused feed-forward module (bf16 operands): out = residual + 0.5 * (Swish(x_act W1^T + b1) W2^T + b2), then
// mode 1: ln_out = LN1(out);  mode 2: out <- LN1(out), ln_out = LN2(out) (plain rounded copy when ln2_g == nullptr).
struct FfnArgs {
  const void* x_act;     // [M, D] bf16 (LayerNorm output of the module)
  const void* w1;        // [hidden, D] bf16
  const void* w2;        // [D, hidden] bf16
  const float *b1, *b2;
  int M, D, hidden;
  const float* residual; // [M, D] fp32
  float* out_f32;        // [M, D] fp32
  int ln_mode; const float *ln1_g, *ln1_b, *ln2_g, *ln2_b; float ln_eps;
  void* ln_out;          // [M, D] bf16 (may alias x_act)
  int cluster;           // CTAs per row tile (1, 2, 4); 0 = choose
};
int launch_ffn_fused(const FfnArgs& a, cudaStream_t stream);
int ffn_timeline(int enable, unsigned long long* out192);
bool ffn_fused_fits(int M, int D, int hidden);   // shape supported by the fused kernel

struct LayerNormArgs {
  const float* x; int rows, dim;
  const float* gamma; const float* beta; float eps;
  void* y_act; float* y_f32;           // outputs (either may be null): activation type (GEMM operand) / fp32 (residual stream)
  // optional compacted activation-type copy of every `copy_stride`-th frame of each sequence (conv_res operand)
  void* copy_out; int copy_stride; int frames_per_seq; int frames_out_per_seq;
};
int launch_layernorm(int precision, const LayerNormArgs& a, cudaStream_t stream);

struct AttnArgs {
  const void* qkv;       // [B*T, 3D] q | k | v (biases already added): fp32 TF32-rounded (EC_PREC_TF32) or bf16 (EC_PREC_BF16)
  const void* E;         // [2*Tp - G, D] projected relative sinusoid rows, same type as qkv
  const float* u; const float* v;   // [D]
  const int* x_len;      // [B] valid frames at this stage, or nullptr
  int B, T, D, H, G;
  void* out; int ld_out; // [B*T, D] activation type
  int in_f32;            // EC_PREC_BF16 only: q|k|v / E are fp32 (odd head dims fall back to the TF32 kernel with bf16 output)
  int in_f16;            // EC_PREC_BF16X2 only: q|k|v / E are plain fp16 (attn_operands_f16(): the 16-bit kernels apply), output packed
};
// Split mode: the attention core runs the 16-bit mma.sync kernels on FP16 q|k|v / E / P (11 significant bits: TF32-grade accuracy at
// the bf16 rate) whenever those kernels support the head layout; other layouts keep packed operands and the TF32 kernel.
inline bool attn_operands_f16(int D, int H, int G) { return D % 8 == 0 && ((G * D) / H) % 2 == 0; }
int launch_relpos_attention(int precision, const AttnArgs& a, cudaStream_t stream);
int launch_relpos_attention_bf16(const AttnArgs& a, cudaStream_t stream);
int try_launch_relpos_attention_tma(const AttnArgs& a, cudaStream_t stream, bool* launched);

// backward of the attention core (attention_bwd.cu): dqkv [B*T, 3D], dE [2Tp-G, D], du / dv [D], all fp32
size_t attention_bwd_work_bytes(int B, int T, int D, int H, int G);
size_t attention_bwd_tc_work_bytes(int B, int T, int D, int H, int G);
struct AttnArgs;
// dqkv_act (optional): dq | dk | dv in the activation type; the tensor-core path then skips the fp32 dqkv (which may be null)
int launch_relpos_attention_bwd_tc(int precision, const AttnArgs& a, const float* dO, float* dqkv, float* dE, float* du, float* dv, void* work,
                                   cudaStream_t stream, void* dqkv_act = nullptr);
int launch_relpos_attention_bwd(int precision, const AttnArgs& a, const float* dO, float* dqkv, float* dE, float* du, float* dv, void* work,
                                cudaStream_t stream, void* dqkv_act = nullptr);

struct DwConvArgs {
  const void* x;         // [B, T, C] activation type (GLU output)
  const float* w;        // [C, k] BatchNorm-folded depthwise taps
  const float* b;        // [C]    BatchNorm-folded bias
  int B, T, C, k, stride;
  void* y;               // [B, T_out, C] activation type = swish(conv)
};
int launch_dwconv_bn_swish(int precision, const DwConvArgs& a, cudaStream_t stream);

struct SubsampleArgs {
  const float* mel;      // [B, F, T] fp32
  const float* w;        // [C, 9] BatchNorm-folded 3x3 taps
  const float* b;        // [C]
  int B, F, T, C;
  void* y;               // [B, T_out, C*F/2] activation type, feature = c*(F/2) + f
};
int launch_subsample_conv(int precision, const SubsampleArgs& a, cudaStream_t stream);
// one-kernel front end (subsample_fused.cu): conv + BN + Swish producer warps -> swizzled A tiles -> tcgen05 Linear; lin_w_perm = the
// Linear weight with its columns permuted to k' = f*Cp + c (launch_linear_weight_permute with Cp = subsample_fused_cpad), activation
// type ([2, D0, K'] in split mode)
struct SubFusedArgs {
  const float* mel; const float* w; const float* b;      // [B, F, T]; BatchNorm-folded taps [C, 9] and bias [C]
  const void* lin_w_perm; const float* lin_b;
  int B, F, T, C, D0;
  float* out;                                            // [B * T_out, D0] fp32
};
bool subsample_fused_fits(int precision, int F, int C, int D0);
int subsample_fused_cpad(int precision, int C);     // channel pitch Cp of the permuted weight (C rounded up to a producer run)
int launch_subsample_linear_fused(int precision, const SubFusedArgs& a, cudaStream_t stream);
// two-layer front end: channels-last layer 0 (y = [B, T_out, F/2, C]), im2col for the 3x3/s2 layer 1, weight preparation
int launch_subsample_conv_cl(int precision, const SubsampleArgs& a, cudaStream_t stream);
int launch_im2col_3x3s2(int precision, const void* y0, int B, int T1, int F1, int C, void* A, cudaStream_t stream);
int launch_conv2_weight_prep(int precision, const float* w, const float* b, const float* g, const float* beta, const float* rm,
                             const float* rv, float eps, int C2, int C, void* w_out, float* b_out, cudaStream_t stream);
int launch_linear_weight_permute(int precision, const float* w, int D, int C, int Fq, void* out, cudaStream_t stream, int Cp = 0);

int launch_cast_rows(int precision, const float* src, void* dst, size_t n, cudaStream_t stream);   // fp32 -> activation type
// weight operand: cast + (split mode) the swapped plane at dst + twin_elems elements
int launch_cast_weight(int precision, const float* src, void* dst, size_t n, size_t twin_elems, cudaStream_t stream);
int launch_fold_bn(const float* w, const float* b, const float* g, const float* beta, const float* rm, const float* rv,
                   float eps, int C, int taps, float* w_out, float* b_out, cudaStream_t stream);
int launch_glu_interleave(int precision, const float* w, const float* b, int channels, int K, int nb, int tiles,
                          void* w_out, float* b_out, cudaStream_t stream);
struct BlockStrides { int n; int sub_layers; int s[EC_MAX_BLOCKS]; };
int launch_stage_lengths(const long long* x_len, int B, int t_mel, const BlockStrides& st, int* out, cudaStream_t stream);
int launch_i64_to_i32(const long long* src, int n, int* dst, int clamp_max, cudaStream_t stream);
int launch_i32_to_i64(const int* src, int n, long long* dst, cudaStream_t stream);

int launch_logsoftmax_argmax(const float* logits, int rows, int V, float* lse, int* argmax, cudaStream_t stream);
int launch_ctc_loss(const float* logits, const float* lse, int B, int T, int V, const int* logits_len, const long long* targets,
                    int target_stride, const long long* target_len, float* loss_per_utt, float* loss_mean, cudaStream_t stream);
int launch_mean(const float* x, int n, float* out, cudaStream_t stream);
// backward of the CTC loss (ctc_grad.cu): per-utterance losses and d(mean loss)/d(logits) * (grad_scale * B)
size_t ctc_grad_work_bytes(int B, int T, int target_stride);
int launch_ctc_grad(const float* logits, const float* lse, int B, int T, int V, const int* logits_len, const long long* targets,
                    int target_stride, const long long* target_len, float* work, float grad_scale, float* loss_per_utt, float* grad,
                    cudaStream_t stream);
// row / element kernels of the backward pass (backward_rows.cu)
size_t layernorm_bwd_work_bytes(int dim);
// emit_out (optional): activation-type copy of emit_scale * dropout_site-mask * (the accumulated x-gradient), see backward_rows.cu
int launch_layernorm_bwd(const float* x, const float* dy, int rows, int dim, const float* gamma, float eps, float* dx, int accumulate,
                         float* dgamma, float* dbeta, float* work, cudaStream_t stream, int emit_precision = 0, void* emit_out = nullptr,
                         float emit_scale = 1.f, const unsigned long long* drop_ctr = nullptr, float drop_p = 0.f, unsigned drop_site = 0);
size_t colsum_work_bytes(int cols);
int launch_colsum(int precision, const void* m, int is_f32, int rows, int cols, float* out, float* work, cudaStream_t stream);
int launch_transpose_cast(int precision, const float* src, int rows, int cols, void* dst, cudaStream_t stream);
int launch_swish_bwd(int precision, const void* z, const float* dy, size_t n, void* dz, cudaStream_t stream);
int launch_glu_bwd(int precision, const void* zg, const float* dy, size_t rows, int C, void* dzg, cudaStream_t stream);
// weight gradient dW[N,K] = dY[M,N]^T . X[M,K] on tcgen05 with MN-major operands (wgrad_tc.cu)
size_t wgrad_work_bytes(int precision, int M, int N, int K);
// db (may be null): bias gradient [N] = column sums of dY, from the same kernel
int launch_wgrad(int precision, const void* dy, const void* x, int M, int N, int K, float* dw, int accumulate, float* db, float* work,
                 cudaStream_t stream);
// training-mode depthwise conv + BatchNorm (batch statistics) + Swish and their backward (conv_train.cu)
size_t conv_train_work_bytes(int C, int K);
int launch_dwconv_raw(int precision, const void* x, const float* w, const float* bias, int B, int T, int C, int K, int stride, float* y,
                      float* sums, float* work, cudaStream_t st);
int launch_bn_finalize(const float* sums, int C, float count, float eps, float momentum, float* mean, float* rstd, float* running_mean,
                       float* running_var, cudaStream_t st);
int launch_bn_swish_fwd(int precision, const float* y, size_t rows, int C, const float* mean, const float* rstd, const float* gamma,
                        const float* beta, void* h, cudaStream_t st);
int launch_bn_swish_bwd_stats(const float* y, const float* dh, size_t rows, int C, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, float* sums, float* work, cudaStream_t st);
int launch_bn_swish_bwd_apply(const float* y, const float* dh, size_t rows, int C, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const float* sums, float count, float* dy, cudaStream_t st);
int launch_dwconv_bwd(int precision, const float* dy, const void* x, const float* w, int B, int T, int C, int K, int stride, float* dx,
                      float* dw, float* db, float* work, cudaStream_t st);
// training forward / backward helpers (train_misc.cu, subsample.cu)
int launch_subsample_conv_raw(const SubsampleArgs& a, cudaStream_t stream);
int launch_cast_scaled(int precision, const float* src, float scale, size_t n, void* dst, cudaStream_t st);
int launch_swish_fwd(int precision, const void* z, size_t n, void* h, cudaStream_t st);
int launch_glu_fwd(int precision, const void* zg, size_t rows, int C, void* out, cudaStream_t st);
int launch_strided_rows(int precision, const float* x, int B, int T_in, int D, int stride, void* out, cudaStream_t st);
int launch_strided_rows_bwd(const float* d, int B, int T_in, int D, int stride, float* dx, cudaStream_t st);
size_t col_stats_work_bytes(int cols);
int launch_col_stats(const float* y, size_t rows, int cols, float* stats, float* work, cudaStream_t st);
int launch_group_stats_merge(const float* col_stats, int C, int group, size_t rows, float* ch_stats, cudaStream_t st);
int launch_group_expand(const float* in, int n_vec, int C, int group, float* out, cudaStream_t st);
int launch_group_sum(const float* in, int n_vec, int C, int group, float* out, cudaStream_t st);
size_t subsample_wgrad_work_bytes(int C, int F, int B, int T);
int launch_subsample_wgrad(const float* dy, const float* mel, int B, int F, int T, int C, float* dw, float* db, float* work, cudaStream_t st);
int launch_greedy_collapse(const int* argmax, int B, int T, const int* logits_len, int* ids, int* counts, cudaStream_t stream);

}  // namespace ec
