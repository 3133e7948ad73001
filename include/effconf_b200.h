/* effconf_b200 -- C ABI of the B200-native Efficient Conformer encoder hot path.
 *
 * Drop-in boundary.  Every entry point takes plain pointers, sizes and a cudaStream_t
 * (passed as void*); no torch types.  All pointers are DEVICE pointers unless stated.  Calls are stream-ordered,
 * never synchronise, allocate nothing (caller provides weights arena + workspace), hold no global tensors and are
 * CUDA-graph capturable.  Return value: EC_OK (0) or EC_ERR (1); ec_last_error() gives the thread-local message.
 * There is NO CPU fallback: a device that is not sm_100 makes ec_engine_create fail.
 *
 * Reference interfaces replaced (burchim/EfficientConformer, paths relative to the reference root):
 *   ec_engine_forward      <- ConformerEncoder.forward after AudioPreprocessing  models/encoders.py:106-142
 *                             (+ ModelCTC.fc, models/model_ctc.py:66, when logits != NULL)
 *   ec_ctc_loss            <- LossCTC.forward                                     models/losses.py:56-71
 *   ec_ctc_greedy          <- ModelCTC.gready_search_decoding (ids, pre-tokenizer) models/model_ctc.py:99-133
 *   ec_op_*                <- the individual modules, for unit parity tests:
 *       ec_op_layernorm        nn.LayerNorm(eps=1e-6)                 models/modules.py:386,433,511; blocks.py:96
 *       ec_op_gemm, ec_op_pointwise_glu  Linear / pointwise Conv1d (+Swish/GLU/residual)  models/layers.py:67,136; modules.py:385-392,513-519
 *       ec_op_relpos_attention (Grouped)RelPosMultiHeadSelfAttention core       models/attentions.py:549-620, 645-718
 *       ec_op_dwconv_bn_swish  depthwise Conv1d + BatchNorm1d(eval) + Swish     models/modules.py:515-517
 *       ec_op_subsample_conv   Conv2d(1->C,3x3,s2)+BatchNorm2d(eval)+Swish      models/modules.py:226-249
 */
#ifndef EFFCONF_B200_H_
#define EFFCONF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EC_OK 0
#define EC_ERR 1
#define EC_MAX_BLOCKS 32

/* numerics mode of the tensor-core operands (accumulation, residual stream, LayerNorm, softmax are always fp32) */
#define EC_PREC_TF32 0 /* parity mode: operands rounded to TF32 (fp32 storage), tcgen05 kind::tf32 */
#define EC_PREC_BF16 1 /* fast mode: bf16 operands and bf16 inter-kernel activations, tcgen05 kind::f16 */
/* split mode (the default; meets the 1e-3 parity gate with ~50x margin): every operand element is the PAIR (hi, lo) of bf16
 * values hi = bf16(x), lo = bf16(x - hi) packed in one 32-bit word (low half hi), i.e. 16 significant bits.  A row of K packed
 * elements is read by the tensor core as 2K bf16 values; weights carry a second plane with the halves swapped, so that
 *   [ah al] . [wh wl] + [ah al] . [wl wh] = (ah + al)(wh + wl)
 * is assembled from two kind::f16 MMAs into the same fp32 TMEM accumulator.  Weight operands are therefore [2, N, K] (plane 0
 * = (hi, lo), plane 1 = (lo, hi)); activations [M, K]. */
#define EC_PREC_BF16X2 2

typedef struct ec_block_cfg {
  int32_t dim_model;   /* D  */
  int32_t dim_expand;  /* D' */
  int32_t num_heads;   /* H  */
  int32_t kernel_size; /* depthwise k (odd) */
  int32_t group_size;  /* attention group size G (odd) */
  int32_t conv_stride; /* 1 or 2 */
  int32_t ff_ratio;
  int32_t reserved;
} ec_block_cfg;

typedef struct ec_config {
  int32_t n_mels;       /* 80 */
  int32_t sub_filters;  /* C of the first Conv2d subsampling layer (1 -> C) */
  int32_t num_blocks;
  int32_t vocab;        /* fc rows; 0 = encoder only */
  int32_t sub_layers;   /* Conv2d subsampling layers: 1 (Efficient Conformer family) or 2 (Conformer family); 0 reads as 1 */
  int32_t sub_filters2; /* C of the second layer (C -> C2) when sub_layers == 2 */
  ec_block_cfg blocks[EC_MAX_BLOCKS];
} ec_config;

/* fp32 parameter pointers in the reference's own layouts (state_dict tensors, contiguous) */
typedef struct ec_ffn_raw { const float *ln_w, *ln_b, *w1, *b1, *w2, *b2; } ec_ffn_raw;
typedef struct ec_block_raw {
  ec_ffn_raw ffn1, ffn2;
  const float *att_ln_w, *att_ln_b, *u, *v, *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo, *wpos, *bpos;
  const float *conv_ln_w, *conv_ln_b, *pw1_w, *pw1_b, *dw_w, *dw_b, *bn_w, *bn_b, *bn_rm, *bn_rv, *pw2_w, *pw2_b;
  const float *norm_w, *norm_b;
  const float *res_w, *res_b; /* conv_res.1 (NULL when dim_model == dim_expand) */
} ec_block_raw;
typedef struct ec_raw_weights {
  const float *sub_conv_w, *sub_conv_b, *sub_bn_w, *sub_bn_b, *sub_bn_rm, *sub_bn_rv;
  /* second subsampling layer, Conv2d(C, C2, 3, stride 2) weight [C2, C, 3, 3] (NULL when sub_layers == 1) */
  const float *sub2_conv_w, *sub2_conv_b, *sub2_bn_w, *sub2_bn_b, *sub2_bn_rm, *sub2_bn_rv;
  const float *lin_w, *lin_b;
  const float *fc_w, *fc_b; /* NULL when vocab == 0 */
  ec_block_raw blocks[EC_MAX_BLOCKS];
} ec_raw_weights;

typedef struct ec_engine ec_engine;

const char* ec_last_error(void);
int ec_version(void);
/* queries the current device; fails unless compute capability is 10.x */
int ec_device_check(void);

int ec_engine_create(const ec_config* cfg, int precision, ec_engine** out);
void ec_engine_destroy(ec_engine* e);
/* bytes of the prepared-weights arena (device), and of the per-call workspace for a (batch, t_mel) shape */
size_t ec_engine_weight_bytes(const ec_engine* e);
size_t ec_engine_workspace_bytes(const ec_engine* e, int batch, int t_mel);
/* convert/fold/concatenate the raw fp32 parameters into the arena (eval-mode BatchNorm folded into the conv taps).
 * Must be re-run whenever parameters change.  The arena must stay alive while the engine is used. */
int ec_engine_prepare(ec_engine* e, const ec_raw_weights* raw, void* arena, void* stream);
/* number of rows / pointer slots of the relative sinusoid tables: block i needs a [rows_i, dim_model_i] table of the
 * activation type (fp32 TF32-rounded, or bf16), rows_i = 2*Tp_i - G_i, holding reference table rows
 * [max_len - Tp + G/2, max_len - G%2 + Tp - G/2)  (models/attentions.py:1309).  Built by the host (same fp32 formula). */
int ec_engine_relpos_rows(const ec_engine* e, int t_mel, int32_t* rows_per_block /* [num_blocks] */, int32_t* frames_per_block);

/* mel [B, n_mels, T] fp32; x_len [B] int64 mel-frame lengths or NULL (all T);
 * relpos[i]: table of block i (activation type);
 * out_x [B, T_out, D_last] fp32 (may be NULL if logits given); logits [B, T_out, vocab] fp32 or NULL;
 * out_len [B] int64 or NULL.  */
int ec_engine_forward(ec_engine* e, int batch, int t_mel, const float* mel, const long long* x_len,
                      const void* const* relpos, void* workspace, float* out_x, float* logits, long long* out_len,
                      void* stream);
int ec_engine_out_frames(const ec_engine* e, int t_mel);

/* Launch accounting and optional per-launch CUDA-event timing (the measurement hook that replaces the reference's
 * torch.autograd.profiler use in Model.eval_time_encoder, models/model.py:627-674).  With profiling enabled a forward
 * must be launched eagerly (not under graph capture); ec_engine_profile_read then returns, per kernel category,
 * the summed device time [ms], algorithmic FLOPs, algorithmic bytes and launch count (arrays of ec_profile_categories()). */
int ec_profile_categories(void);
const char* ec_profile_category_name(int cat);
int ec_engine_set_profiling(ec_engine* e, int enabled);
int ec_engine_last_launches(const ec_engine* e);
int ec_engine_profile_read(ec_engine* e, double* ms, double* flops, double* bytes, int32_t* launches);
/* Execution options.  fuse_ln (per engine, default 1): the five LayerNorms of a block run in the epilogue of the producing
 * GEMM instead of as separate kernels (needs model dims <= 256).  pdl (process wide, default 1, env EFFCONF_PDL=0 disables):
 * launch every kernel with programmatic stream serialization so its prologue overlaps the predecessor's tail. */
int ec_engine_set_fuse_ln(ec_engine* e, int enabled);
/* fuse_ffn (per engine, default 1; EC_PREC_BF16 with fuse_ln only): each feed-forward module runs as ONE cluster kernel
 * (W1 -> Swish -> W2 -> half-step residual -> LayerNorm) whose hidden activation never leaves the SM. */
int ec_engine_set_fuse_ffn(ec_engine* e, int enabled);
/* one-layer Conv2d subsampling + Linear as one kernel (the 4800-wide operand stays in shared memory); default on where it fits */
int ec_engine_set_fuse_front(ec_engine* e, int enabled);
int ec_set_pdl(int enabled);
/* Debug / measurement only: bit c set = the kernels of profile category c are NOT launched (outputs are garbage).  Used by
 * tools/marginal_cost.py to measure what each kernel category really costs inside the replayed CUDA graph (PDL overlap included). */
int ec_engine_set_skip_mask(ec_engine* e, unsigned mask);
/* Debug: enable in-kernel SM-clock stamps in the GEMM and read the 12 stamps of the last GEMM's CTA (0,0) (synchronises). */
int ec_debug_gemm_timeline(int enable, unsigned long long* out12);
/* debug: force the N tile width of the plain GEMM (0 = automatic); tile-shape studies only */
int ec_debug_gemm_block_n(int block_n);
/* Same for the fused feed-forward kernel: 192 stamp slots of CTA 0 (see ffn_fused.cu). */
int ec_debug_ffn_timeline(int enable, unsigned long long* out192);

/* CTC head.  logits [B, T, V] fp32, logits_len [B] int64, targets [B, target_stride] int64 (blank = 0), target_len [B] int64.
 * scratch: at least B*T*(sizeof(float)+sizeof(int)) + B*sizeof(int) bytes.  loss_per_utt [B], loss_mean [1] fp32. */
size_t ec_ctc_scratch_bytes(int batch, int t, int vocab);
int ec_ctc_loss(const float* logits, int batch, int t, int vocab, const long long* logits_len, const long long* targets,
                int target_stride, const long long* target_len, void* scratch, float* loss_per_utt, float* loss_mean,
                void* stream);
/* ids [B, T] int32 (collapsed token ids, zero padded), counts [B] int32 */
/* ---- backward operators (training step; each is tested against autograd over the CPU oracle, tests/test_gpu_backward.py) ----
 * ec_op_layernorm_bwd : x, dy [rows, dim] fp32, gamma [dim] -> dx [rows, dim] (accumulate != 0: dx += ...: the LayerNorm sits on
 *                       a residual branch), dgamma / dbeta [dim].  Row statistics are recomputed from x.  nn.LayerNorm(eps) backward.
 * ec_op_colsum        : out[c] = sum_r m[r, c]  (bias gradient of a Linear / pointwise conv); m fp32 (is_f32) or activation type.
 * ec_op_transpose_cast: dst [cols, rows] activation type = transpose of src [rows, cols] fp32.  The data gradient of a Linear,
 *                       dX = dY . W, is ec_op_gemm(A = dY, W = transpose_cast(W)) on the same tcgen05 kernel as the forward.
 * ec_op_swish_bwd     : dz = dy * d/dz (z sigmoid z);   ec_op_glu_bwd: zg = [a | g] (rows x 2C), dy (rows x C) -> [da | dg].
 * Reductions use per-CTA partials added in a fixed order: bit-reproducible.
 * ec_op_wgrad         : dW [N, K] fp32 (+)= dY[M, N]^T . X[M, K], both activation type, on tcgen05 with MN-major operands (no
 *                       transposed copies), split over M with a fixed-order reduction of the partial tiles. */
/* Training-forward element kernels (the pre-activation tensors are kept for the backward, so Swish / GLU run on their own), the
 * strided frame copy feeding conv_res and its scatter-add backward, and the BatchNorm2d pieces of the Conv2d subsampling layer:
 * raw fp32 convolution, per-column (mean, M2) statistics, merge / expand / sum between the F/2 columns of a channel and the
 * channel (feature index c*F/2 + f), weight / bias gradient of the 3x3 stride-2 convolution. */
int ec_op_cast_scaled(int precision, const float* src, float scale, size_t n, void* dst, void* stream); /* dst = act_type(scale * src) */
int ec_op_swish_fwd(int precision, const void* z, size_t n, void* h, void* stream);
int ec_op_glu_fwd(int precision, const void* zg, size_t rows, int channels, void* out, void* stream);
int ec_op_strided_rows(int precision, const float* x, int batch, int t, int dim, int stride, void* out, void* stream);
int ec_op_strided_rows_bwd(const float* d, int batch, int t, int dim, int stride, float* dx, void* stream);
int ec_op_subsample_conv_raw(const float* mel, const float* w, const float* b, int batch, int n_mels, int t, int channels, float* y, void* stream);
size_t ec_op_col_stats_work_bytes(int cols);
int ec_op_col_stats(const float* y, size_t rows, int cols, float* stats, void* work, void* stream);
int ec_op_group_stats_merge(const float* col_stats, int channels, int group, size_t rows, float* ch_stats, void* stream);
int ec_op_group_expand(const float* in, int n_vec, int channels, int group, float* out, void* stream);
int ec_op_group_sum(const float* in, int n_vec, int channels, int group, float* out, void* stream);
size_t ec_op_subsample_wgrad_work_bytes(int channels, int n_mels, int batch, int t);
int ec_op_subsample_wgrad(const float* dy, const float* mel, int batch, int n_mels, int t, int channels, float* dw, float* db, void* work,
                          void* stream);
/* Backward of ec_op_relpos_attention (same operand conventions: qkv [B*T, 3D] and E [2Tp-G, D] in the activation type):
 * d_out [B*T, D] fp32 = gradient of the attention output -> dqkv [B*T, 3D], dE [2Tp-G, D] (summed over the batch), du, dv [D], fp32.
 * First implementation on the CUDA cores with P and dS materialised in `work`; no atomics, bit-reproducible. */
size_t ec_op_relpos_attention_bwd_work_bytes(int batch, int t, int dim, int heads, int group);
int ec_op_relpos_attention_bwd(int precision, const void* qkv, const void* E, const float* u, const float* v, const int32_t* x_len, int batch,
                               int t, int dim, int heads, int group, const float* d_out, float* dqkv, float* dE, float* du, float* dv,
                               void* work, void* stream);
/* same, with dq | dk | dv delivered in the activation type (dqkv_act, [B*T, 3D]): the operand of the QKV weight- and data-gradient GEMMs.
 * The tensor-core path then never writes the fp32 tensor (dqkv_f32 may be NULL there; the TF32 CUDA-core path needs it as scratch). */
int ec_op_relpos_attention_bwd_act(int precision, const void* qkv, const void* E, const float* u, const float* v, const int32_t* x_len, int batch,
                                   int t, int dim, int heads, int group, const float* d_out, float* dqkv_f32, void* dqkv_act, float* dE,
                                   float* du, float* dv, void* work, void* stream);
/* Training-mode depthwise stage of the convolution module (reference models/modules.py:515-517 under model.train()) and its
 * backward, in stages so that the host can all-reduce the statistics across ranks between them (SyncBatchNorm):
 *   ec_op_dwconv_raw        y [B, T_out, C] fp32 = depthwise conv (raw taps w [C, k], bias); stats [2][C] = mean, centred sum of squares
 *                           M2 over the B*T_out frames (two-pass per CTA + Chan merge: robust when |mean| >> std)
 *   ec_op_bn_finalize       mean / rstd (biased variance M2 / count) from the (merged) stats over `count` frames; running stats updated in place when given
 *   ec_op_bn_swish_fwd      h = swish(BatchNorm(y))  (activation type)
 *   ec_op_bn_swish_bwd_stats  sums [2][C] = dbeta, dgamma  (= sum dz, sum dz * xhat with dz = dh * swish'(z))
 *   ec_op_bn_swish_bwd_apply  dy = gamma * rstd * (dz - sums[0] / count - xhat * sums[1] / count)
 *   ec_op_dwconv_bwd        dx [B, T, C] (may be NULL), dw [C, k], db [C] from dy [B, T_out, C] and the saved conv input x */
size_t ec_op_conv_train_work_bytes(int channels, int k);
int ec_op_dwconv_raw(int precision, const void* x, const float* w, const float* bias, int batch, int t, int channels, int k, int stride,
                     float* y, float* sums, void* work, void* stream);
int ec_op_bn_finalize(const float* sums, int channels, float count, float eps, float momentum, float* mean, float* rstd,
                      float* running_mean, float* running_var, void* stream);
int ec_op_bn_swish_fwd(int precision, const float* y, size_t rows, int channels, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, void* h, void* stream);
int ec_op_bn_swish_bwd_stats(const float* y, const float* dh, size_t rows, int channels, const float* mean, const float* rstd,
                             const float* gamma, const float* beta, float* sums, void* work, void* stream);
int ec_op_bn_swish_bwd_apply(const float* y, const float* dh, size_t rows, int channels, const float* mean, const float* rstd,
                             const float* gamma, const float* beta, const float* sums, float count, float* dy, void* stream);
int ec_op_dwconv_bwd(int precision, const float* dy, const void* x, const float* w, int batch, int t, int channels, int k, int stride,
                     float* dx, float* dw, float* db, void* work, void* stream);
size_t ec_op_wgrad_work_bytes(int precision, int M, int N, int K);
int ec_op_wgrad(int precision, const void* dy, const void* x, int M, int N, int K, float* dw, int accumulate, void* work, void* stream);
/* same, plus the bias gradient db [N] = column sums of dY from the same kernel (a ones tile as a second MMA operand) */
int ec_op_wgrad_bias(int precision, const void* dy, const void* x, int M, int N, int K, float* dw, int accumulate, float* db, void* work,
                     void* stream);
size_t ec_op_layernorm_bwd_work_bytes(int dim);
int ec_op_layernorm_bwd(const float* x, const float* dy, int rows, int dim, const float* gamma, float eps, float* dx, int accumulate,
                        float* dgamma, float* dbeta, void* work, void* stream);
/* same, plus (emit_out != NULL) the activation-type operand of the next GEMM of the backward chain:
 * emit_out[r, c] = act_type(emit_scale * keep_{drop_site}/(1-p) * dx[r, c]) with dx the value AFTER the residual accumulation -- what
 * ec_op_dropout (fp32 -> activation type, scaled) computed from dx in a separate pass; drop_counter NULL = no mask (plain scaled cast) */
int ec_op_layernorm_bwd_emit(const float* x, const float* dy, int rows, int dim, const float* gamma, float eps, float* dx, int accumulate,
                             float* dgamma, float* dbeta, void* work, int emit_precision, void* emit_out, float emit_scale,
                             const unsigned long long* drop_counter, float drop_p, unsigned drop_site, void* stream);
size_t ec_op_colsum_work_bytes(int cols);
int ec_op_colsum(int precision, const void* m, int is_f32, int rows, int cols, float* out, void* work, void* stream);
int ec_op_transpose_cast(int precision, const float* src, int rows, int cols, void* dst, void* stream);
int ec_op_swish_bwd(int precision, const void* z, const float* dy, size_t n, void* dz, void* stream);
int ec_op_glu_bwd(int precision, const void* zg, const float* dy, size_t rows, int channels, void* dzg, void* stream);
/* Backward of ec_ctc_loss (first kernel of the training backward pass): also writes grad_logits [B, T, V] fp32 =
 * d(mean_b nll_b) / d logits = (softmax - posterior occupancy of the class) / B for t < logits_len[b], 0 for padded frames
 * (what autograd gives for LossCTC.forward, reference models/losses.py:56-71).  work: ec_ctc_grad_work_bytes() device bytes. */
size_t ec_ctc_grad_work_bytes(int batch, int t, int target_stride);
int ec_ctc_loss_grad(const float* logits, int batch, int t, int vocab, const long long* logits_len, const long long* targets,
                     int target_stride, const long long* target_len, void* scratch, void* work, float* loss_per_utt,
                     float* loss_mean, float* grad_logits, void* stream);
int ec_ctc_greedy(const float* logits, int batch, int t, int vocab, const long long* logits_len, void* scratch,
                  int32_t* ids, int32_t* counts, void* stream);

/* ---- closing the training step (reference models/model.py:239-259: forward, loss.backward(), optimizer.step(), scheduler.step()) ----
 * Dropout (reference nn.Dropout sites models/encoders.py:119, modules.py:389,391,486,521): counter-based masks.  `counter` is a
 * device array {seed, step}; the keep bit of element i at `site` is a pure function of (seed, step, site, i), so the backward
 * re-applies the same mask without storing it and a replayed CUDA graph draws fresh masks after ec_op_dropout_advance.
 *   ec_op_dropout          dst = scale * keep / (1 - p) * src;  src fp32 (src_f32) or activation type, dst likewise (dst_f32)
 *   ec_op_dropout_residual out = residual + alpha * keep / (1 - p) * y   (fp32)
 * Adam (torch.optim.Adam semantics as the reference uses them, models/model.py:88-93: L2 weight decay added to the gradient, bias
 * correction) over ONE flat fp32 arena; `state` is a device int[4] = {lr as float bits, adam step t, schedule step s, 0}.  After the
 * update t += 1 and, with schedule == 1, the Transformer schedule of models/schedules.py:99-123 advances on device:
 * s += 1, lr = K * dim^-0.5 * min(s^-0.5, s * warmup^-1.5).  grad_scale multiplies the gradient first (1 / world size after a SUM all-reduce).
 * ec_op_stats_merge_ranks: SyncBatchNorm forward statistics -- gathered [world][2][C] (mean, centred sum of squares) and counts [world]
 * (frames per rank, device) merged into out [2][C]. */
int ec_op_dropout_advance(unsigned long long* counter, void* stream);
int ec_op_dropout(int precision, const void* src, int src_f32, float scale, size_t n, void* dst, int dst_f32, float p,
                  const unsigned long long* counter, unsigned site, void* stream);
int ec_op_dropout_residual(const float* y, const float* residual, float alpha, size_t n, float* out, float p,
                           const unsigned long long* counter, unsigned site, void* stream);
int ec_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, int* state, float beta1, float beta2,
                 float eps, float weight_decay, float grad_scale, int schedule, float sched_k, float sched_dim, float sched_warmup,
                 void* stream);
int ec_op_stats_merge_ranks(const float* gathered, const float* counts, int world, int channels, float* out, void* stream);
/* gather n separately allocated fp32 tensors (device pointer table srcs[n], element offsets[n] into the arena, sizes[n]) into one
 * flat arena: the gradient bucket that the data-parallel all-reduce (reference DistributedDataParallel, models/model_ctc.py:73-75)
 * and ec_adam_step operate on. */
/* Swish fused with the dropout that follows it in the feed-forward module: dy == NULL: out = keep/(1-p) * z*sigmoid(z);
 * dy != NULL (backward): out = keep/(1-p) * dy * d/dz(z*sigmoid(z)); z / out in the activation type, same mask as ec_op_dropout at `site`. */
int ec_op_swish_dropout(int precision, const void* z, const float* dy, size_t n, void* out, float p, const unsigned long long* counter,
                        unsigned site, void* stream);
/* W^T operands of all data-gradient GEMMs in one launch: desc[n][4] = {src element offset, rows, cols, dst element offset} over the flat
 * fp32 parameter arena; dst[c][r] = act_type(src[r][c]). */
int ec_op_transpose_cast_multi(int precision, const float* src_arena, const long long* desc, int n, int ctas_per_tensor, void* dst_arena,
                               void* stream);
/* forward operands [N, K] of all GEMM weights in one launch (same descriptor table; dst offsets in activation-type elements) */
int ec_op_cast_multi(int precision, const float* src_arena, const long long* desc, int n, int ctas_per_tensor, void* dst_arena, void* stream);
int ec_op_pack_flat(const float* const* srcs, const long long* offsets, const long long* sizes, int n, float* arena, void* stream);
/* accumulate != 0: arena += gathered gradients (gradient accumulation over micro-batches, reference models/model.py:245-259) */
int ec_op_pack_flat_acc(const float* const* srcs, const long long* offsets, const long long* sizes, int n, float* arena, int accumulate,
                        void* stream);

/* ---- single-operator entry points (unit parity tests; same kernels the engine launches) ---------------------------- */
int ec_op_cast(int precision, const float* src, void* dst, size_t n, void* stream);
/* GEMM weight operand ("W" of ec_op_gemm and friends): ec_weight_planes(precision) planes of n = N*K activation-type elements; plane 0
 * is ec_op_cast(src), plane 1 (EC_PREC_BF16X2 only) the same pairs with the halves swapped.  ec_op_transpose_cast, ec_op_cast_multi,
 * ec_op_transpose_cast_multi and ec_op_pointwise_glu's scratch follow the same [planes, N, K] convention. */
int ec_weight_planes(int precision);
int ec_op_cast_weight(int precision, const float* src, size_t n, void* dst, void* stream);
int ec_op_layernorm(int precision, const float* x, int rows, int dim, const float* gamma, const float* beta, float eps,
                    void* y_act /* activation type or NULL */, float* y_f32 /* or NULL */, void* stream);
/* out = alpha * act(A @ W^T + bias) + residual.  A [M,K], W [N,K] in the activation type (use ec_op_cast).  act: 0 none, 1 swish. */
int ec_op_gemm(int precision, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha, int act,
               const float* residual, float* out_f32, void* out_act, void* stream);
/* flags bit 0 (EC_PREC_BF16X2 only): out_act is written as PLAIN fp16 -- the q|k|v and E operands of the attention core in split mode */
int ec_op_gemm_ex(int precision, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha, int act,
                  const float* residual, float* out_f32, void* out_act, int flags, void* stream);
/* Training-step GEMM with the element work of the surrounding nn.Dropout / Swish modules folded into its epilogue (reference
 * models/modules.py:386-391 feed-forward module, :486 attention dropout, :521 conv-module dropout, models/encoders.py:119):
 *   z        = A W^T + bias
 *   out_act  = act_type(z)                                      (optional: the pre-activation the backward needs)
 *   out_act2 = act_type(keep_{site2}/(1-p) * Swish(out_act))    (optional: Swish + the module's first dropout = operand of the next GEMM)
 *   aux_act  : z <- z * keep_{site_aux}/(1-p) * Swish'(aux)     (optional: data gradient through dropout(Swish(.)), aux = saved pre-activation
 *                                                                [M, N] in the activation type; excludes `residual`)
 *   out_f32  = alpha * keep_{site}/(1-p) * z + residual         (optional; drop_site 0 = no dropout here)
 * drop_counter = the device {seed, step} pair of ec_op_dropout_advance (NULL when every site is 0); masks are the same function of
 * (seed, step, site, element = row * N + column) as ec_op_dropout / ec_op_swish_dropout draw, so either may compute a forward or a backward. */
int ec_op_gemm_train(int precision, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha,
                     const float* residual, float* out_f32, void* out_act, const unsigned long long* drop_counter, float drop_p,
                     unsigned drop_site, void* out_act2, unsigned drop_site2, const void* aux_act, unsigned drop_site_aux, void* stream);
/* Training-step projection with the LayerNorm of the NEXT module fused (N <= 256): out_f32 = alpha * keep_{site}/(1-p) * (A W^T + bias) +
 * residual (kept for the backward), ln_out = act_type(LN(out_f32; g1, b1, eps)) = the next GEMM's operand.  Replaces a projection, its
 * nn.Dropout, the residual add and the following nn.LayerNorm (reference models/modules.py:386-391,433,486,511,521; blocks.py:122-132). */
int ec_op_gemm_ln_train(int precision, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha, const float* residual,
                        float* out_f32, const float* g1, const float* b1, float eps, void* ln_out, const unsigned long long* drop_counter,
                        float drop_p, unsigned drop_site, void* stream);
/* Storage of the q|k|v and E operands that ec_op_relpos_attention / _bwd read in this mode and head layout: 0 = the activation type,
 * 1 = bf16 (EC_PREC_BF16), 2 = fp16 (EC_PREC_BF16X2 when dim % 8 == 0 and the head dim G*dim/heads is even: the 16-bit mma.sync kernels
 * then run on fp16 operands -- 11 significant bits, TF32-grade accuracy at the bf16 rate -- and write the packed output). */
int ec_attention_operand_kind(int precision, int dim, int heads, int group);
/* GEMM + fused LayerNorm epilogue (N <= 256): out_f32 = alpha*(A W^T + bias) + residual;
 * ln_mode 1: ln_out = LN(out; g1,b1) (activation type), optional copy_out = activation-type copy of every copy_stride-th frame of out;
 * ln_mode 2: out_f32 <- LN(out; g1,b1) in place, ln_out = LN(out_f32; g2,b2), or a plain activation-type copy when g2 == NULL. */
int ec_op_gemm_ln(int precision, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha,
                  const float* residual, float* out_f32, int ln_mode, const float* g1, const float* b1, const float* g2,
                  const float* b2, float eps, void* ln_out, void* copy_out, int copy_stride, int frames_per_seq,
                  int frames_out_per_seq, void* stream);
/* Fused feed-forward module (EC_PREC_BF16 operands; reference models/modules.py:367-398 + the half-step residual of
 * models/blocks.py:123-150): out_f32 = residual + 0.5*(Swish(x_act W1^T + b1) W2^T + b2); LayerNorm modes as ec_op_gemm_ln.
 * x_act [M,D], w1 [hidden,D], w2 [D,hidden] bf16; D <= 256.  cluster: CTAs per 128-row tile splitting the hidden dim (0 = auto). */
int ec_op_ffn(const void* x_act, const void* w1, const float* b1, const void* w2, const float* b2, int M, int D, int hidden,
              const float* residual, float* out_f32, int ln_mode, const float* g1, const float* be1, const float* g2,
              const float* be2, float eps, void* ln_out, int cluster, void* stream);
/* pointwise Conv1d(K -> 2*channels) + GLU: out[m, c] = (A w_c + b_c) * sigmoid(A w_{C+c} + b_{C+c}).  w_raw [2C, K], b_raw [2C] fp32
 * (reference layout); w_scratch / b_scratch hold ec_op_glu_scratch_rows(channels) rows of the interleaved copy. */
int ec_op_pointwise_glu(int precision, const void* A, const float* w_raw, const float* b_raw, int M, int channels, int K,
                        void* w_scratch, float* b_scratch, void* out_act, void* stream);
int ec_op_glu_scratch_rows(int channels);
/* eval BatchNorm folded into conv taps: w_out [C, taps], b_out [C] */
int ec_op_fold_bn(const float* w, const float* b, const float* g, const float* beta, const float* rm, const float* rv, float eps,
                  int C, int taps, float* w_out, float* b_out, void* stream);
/* qkv [B*T, 3D] and E [2Tp-G, D]: fp32 already rounded to TF32 for EC_PREC_TF32, bf16 for EC_PREC_BF16 (what the QKV / pos GEMM epilogues
 * emit); EC_PREC_BF16X2: plain fp16 when ec_attention_operand_kind() == 2 (ec_op_gemm_ex flag), packed pairs otherwise; out packed */
int ec_op_relpos_attention(int precision, const void* qkv, const void* E, const float* u, const float* v,
                           const int32_t* x_len, int batch, int t, int dim, int heads, int group, void* out, void* stream);
int ec_op_dwconv_bn_swish(int precision, const void* x, const float* w_folded, const float* b_folded, int batch, int t,
                          int channels, int k, int stride, void* y, void* stream);
int ec_op_subsample_conv(int precision, const float* mel, const float* w_folded, const float* b_folded, int batch,
                         int n_mels, int t, int channels, void* y, void* stream);

/* ---- SyncBatchNorm exchange over NVLink / NVSwitch peer memory (csrc/p2p_exchange.cu; reference models/model_ctc.py:70-75:
 * nn.SyncBatchNorm under DistributedDataParallel).  One kernel per exchange: store the payload into every rank's mailbox, raise flags
 * (release / acquire at system scope), wait, merge in rank order.  peer_ptrs [world] (HOST array) = this process's mapped addresses of
 * every rank's mailbox (ec_p2p_mailbox_bytes each, zero-initialised, symmetric memory); mode 0: data [n] <- sum over ranks;
 * mode 1: data [2, n/2] = (mean, M2) with `count` local frames <- Chan merge over ranks, total count to out_count (device, optional). */
size_t ec_p2p_mailbox_bytes(int world);
int ec_p2p_max_payload_floats(void);
int ec_p2p_bn_exchange(const unsigned long long* peer_ptrs, int rank, int world, float* data, int n, int mode, float count,
                       float* out_count, void* stream);
int ec_p2p_error(const unsigned long long* peer_ptrs, int rank, int world, int* out);

/* ---- Transducer joint network + RNN-T loss, forward (csrc/rnnt.cu) ----------------------------------------------
 * ec_op_joint_hidden : hidden[(b,t,u), :] = act_type(act(fe[b,t,:] + gd[b,u,:])), fe [B*T, J] = Linear_enc(f), gd [B*U1, J] = Linear_dec(g)
 *                      (both fp32, from ec_op_gemm); act 0 none, 1 tanh, 2 relu, 3 swish.  Replaces the two `repeat`s, the sum and the
 *                      activation of reference models/joint_networks.py:84-99 (joint_mode "sum"); the output projection
 *                      (joint_networks.py:102) is ec_op_gemm on `hidden`.
 * ec_rnnt_loss       : logits [B, T, U1, V] fp32 (U1 = U + 1), labels [B, label_stride] int64, frame_len / label_len [B] int64 ->
 *                      loss_per_utt [B] = -log p(labels | frames) (Graves 2012), loss_mean [1] = their mean: what the reference gets from
 *                      warp_rnnt.rnnt_loss(log_softmax(logits), ..., average_frames=False, reduction='mean', blank=0, gather=True)
 *                      (models/losses.py:22-46).  scratch: ec_rnnt_scratch_bytes(B, T, U1). */
int ec_op_joint_hidden(int precision, const float* fe, const float* gd, int batch, int t, int u1, int dim_joint, int act, void* hidden, void* stream);
size_t ec_rnnt_scratch_bytes(int batch, int t, int u1);
int ec_rnnt_loss(const float* logits, int batch, int t, int u1, int vocab, const long long* labels, int label_stride, const long long* frame_len,
                 const long long* label_len, int blank, void* scratch, float* loss_per_utt, float* loss_mean, void* stream);
/* ec_rnnt_loss_grad  : the same plus grad [B, T, U1, V] = grad_scale * d(sum_b nll_b) / d logits with the log_softmax folded in (alpha and
 *                      beta wavefronts, Graves 2012 eq. 16-20): what autograd computes through warp_rnnt.rnnt_loss and F.log_softmax
 *                      (models/losses.py:36-44); grad_scale = 1 / B for reduction = 'mean'.
 * ec_op_joint_hidden_bwd : gradient through hidden = act(fe + gd) and the two broadcasts (`repeat`s of models/joint_networks.py:88-89):
 *                      dfe [B*T, J] = sum_u d_hidden * act'(hidden), dgd [B*U1, J] = sum_t ...; act' from the stored output (none / tanh / relu). */
int ec_rnnt_loss_grad(const float* logits, int batch, int t, int u1, int vocab, const long long* labels, int label_stride,
                      const long long* frame_len, const long long* label_len, int blank, void* scratch, float* loss_per_utt, float* loss_mean,
                      float grad_scale, float* grad, void* stream);
int ec_op_joint_hidden_bwd(int precision, const void* hidden, const float* d_hidden, int batch, int t, int u1, int dim_joint, int act, float* dfe,
                           float* dgd, void* stream);

/* ---- Front end on the device (csrc/frontend.cu) -------------------------------------------------------------------
 * ec_op_logmel      : audio [B, samples] fp32 -> out [B, n_mels, samples / hop + 1] fp32 =
 *                       log(fb^T |STFT|^2 + 1e-9), then (x - mean) / std when `normalize`
 *                     = reference models/modules.py:87-106 AudioPreprocessing.forward (torchaudio Spectrogram(n_fft, win, hop, power 2,
 *                     centre / reflect padding, one-sided) -> MelScale -> log) in one kernel.  window [n_fft]: the analysis window already
 *                     centred and zero padded to n_fft (torch.stft's own padding of a win_length window); fb [n_fft/2 + 1, n_mels]: the
 *                     MelScale filter bank (a state buffer of the reference module); krange [n_mels][2] (optional, device): first / one past
 *                     last non-zero bin of each filter.  n_fft a power of two in [64, 2048]; samples > n_fft / 2 (as torch.stft).
 * ec_op_specaugment : mel [B, n_mels, T] fp32 in place <- reference models/modules.py:136-151 SpecAugment.forward: mF frequency masks shared
 *                     by the batch (torchaudio FrequencyMasking(F, iid_masks=False)), mT time masks per utterance inside its x_len[b] valid
 *                     frames with parameter int(pS * x_len[b]); mask_along_axis arithmetic: value = U * param, start = floor(U' * (size -
 *                     value)), end = start + floor(value); masked cells <- 0.  U, U' are counter-based draws of the device {seed, step} pair
 *                     of ec_op_dropout_advance (so a replayed CUDA graph draws new masks); x_len NULL = every utterance is T frames long. */
int ec_op_logmel(const float* audio, int batch, int samples, int n_fft, int hop, const float* window, const float* fb, const int* krange,
                 int n_mels, int normalize, float mean, float stdv, float* out, void* stream);
int ec_op_specaugment(float* mel, const long long* x_len, int batch, int n_mels, int t, int mF, int F, int mT, float pS,
                      const unsigned long long* counter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EFFCONF_B200_H_ */
