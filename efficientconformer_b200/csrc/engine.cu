// Engine: weight preparation, workspace planning and the kernel sequence of one encoder forward, plus the C ABI.
//
// Restates the control flow of reference models/encoders.py:106-142 (ConformerEncoder.forward after the audio front
// end) and models/blocks.py:119-137 (ConformerBlock.forward) as a fixed, stream-ordered launch sequence that is
// CUDA-graph capturable: no allocation, no host synchronisation, no data-dependent host control flow.
#include "ec_common.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ec {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

static int g_pdl = -1;   // -1: read EFFCONF_PDL (default on)
bool pdl_enabled() {
  if (g_pdl < 0) { const char* e = getenv("EFFCONF_PDL"); g_pdl = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return g_pdl != 0;
}

static int g_pdl_train = -1;
bool pdl_train_enabled() {
  if (g_pdl_train < 0) { const char* e = getenv("EFFCONF_PDL_TRAIN"); g_pdl_train = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return g_pdl_train != 0;
}

bool SideStreams::init() {
  if (ok) return true;
  for (int i = 0; i < kN; ++i) {
    if (cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&join_ev[i], cudaEventDisableTiming) != cudaSuccess) return false;
  }
  if (cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming) != cudaSuccess) return false;
  ok = true;
  return true;
}
bool SideStreams::fork(cudaStream_t st, int n) {
  if (!init() || n > kN) return false;
  if (cudaEventRecord(fork_ev, st) != cudaSuccess) return false;
  for (int i = 0; i < n; ++i)
    if (cudaStreamWaitEvent(s[i], fork_ev, 0) != cudaSuccess) return false;
  return true;
}
bool SideStreams::join(cudaStream_t st, int n) {
  for (int i = 0; i < n; ++i) {
    if (cudaEventRecord(join_ev[i], s[i]) != cudaSuccess) return false;
    if (cudaStreamWaitEvent(st, join_ev[i], 0) != cudaSuccess) return false;
  }
  return true;
}
SideStreams& side_streams() {
  static SideStreams ss;     // one device per process (one process per GPU)
  return ss;
}
static int g_side = -1;      // EFFCONF_SIDE_STREAMS=0 serialises everything on the caller's stream
bool side_streams_enabled() {
  if (g_side < 0) { const char* e = getenv("EFFCONF_SIDE_STREAMS"); g_side = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return g_side != 0;
}

struct Arena {   // bump allocator over a caller-provided buffer (or a dry run when base == nullptr)
  uint8_t* base; size_t off;
  explicit Arena(void* b) : base(reinterpret_cast<uint8_t*>(b)), off(0) {}
  void* take(size_t bytes) {
    off = align_up(off, 256);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

struct FfnW { float *ln_w, *ln_b; void* w1; float* b1; void* w2; float* b2; };
struct BlockW {
  FfnW ffn1, ffn2;
  float *att_ln_w, *att_ln_b, *u, *v; void* wqkv; float* bqkv; void* wo; float* bo; void* wpos; float* bpos;
  float *conv_ln_w, *conv_ln_b; void* pw1; float* pw1_b; float *dw_w, *dw_b; void* pw2; float* pw2_b;
  void* res_w; float* res_b;
  float *norm_w, *norm_b;
  int glu_nb, glu_tiles;
};
struct Weights {
  float *sub_w, *sub_b; void* sub2_w; float* sub2_b; void* lin_w; float* lin_b; void* fc_w; float* fc_b;
  void* lin_w_perm;      // one-layer front end: the Linear weight with columns in k' = f*Cp + c order (fused front-end kernel)
  BlockW blk[EC_MAX_BLOCKS];
};

}  // namespace ec

namespace ec {
// kernel categories of one forward (profiling / launch accounting)
enum ProfCat { PC_SUBSAMPLE = 0, PC_LIN, PC_LAYERNORM, PC_FFN_W1, PC_FFN_W2, PC_QKV, PC_POS, PC_ATTN, PC_OUT, PC_PW1_GLU, PC_DWCONV,
               PC_RES, PC_PW2, PC_FC, PC_MISC, PC_FFN_FUSED, PC_IM2COL, PC_SUB_CONV2, PC_COUNT };
static const char* kProfNames[PC_COUNT] = {"subsample_conv", "gemm_sub_linear", "layernorm", "gemm_ffn_w1_swish", "gemm_ffn_w2_res",
                                           "gemm_qkv", "gemm_pos", "relpos_attention", "gemm_att_out_res", "gemm_pw1_glu",
                                           "dwconv_bn_swish", "gemm_conv_res", "gemm_pw2_res", "gemm_fc", "misc", "ffn_fused", "im2col_3x3s2",
                                           "gemm_sub_conv2_swish"};
struct ProfEntry { int cat; double flops, bytes; cudaEvent_t e0, e1; };
}  // namespace ec

struct ec_engine {
  ec_config cfg;
  int precision;
  size_t esize;
  ec::Weights w;
  size_t weight_bytes;
  bool prepared;
  // profiling: CUDA events around every launch of the last forward (eager launches only, not under graph capture)
  bool prof_enabled = false;
  std::vector<ec::ProfEntry> prof;
  std::vector<cudaEvent_t> event_pool;
  size_t events_used = 0;
  int last_launches = 0;
  unsigned skip_mask = 0;  // debug (ec_engine_set_skip_mask): kernel categories NOT launched -- marginal-cost studies only, results are garbage
  bool fuse_ffn = true;    // bf16 mode: whole feed-forward module in one cluster kernel (needs fuse_ln)
  bool fuse_ln = true;     // LayerNorms in the epilogue of the producing GEMM (needs dim <= 256)
  bool fuse_front = true;  // Conv2d subsampling + Linear as one kernel (subsample_fused.cu; one-layer front end, first dim <= 256)
  // the positional projections E_i = pos_layer_i(R) depend on weights only: they run on a forked stream, off the critical path
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace ec {

// GLU tiles are [nb value rows | nb gate rows]; the epilogue stores 32-column slabs, so with more than one tile the
// tile width must be a multiple of 32 (a single tile is clipped by the tensor bound instead).
static void pick_glu(int channels, int* nb, int* tiles) {
  *tiles = cdiv(channels, 128);
  *nb = *tiles == 1 ? round_up(channels, 8) : round_up(cdiv(channels, *tiles), 32);
}

// width of the Linear that follows the Conv2d subsampling: C_last * n_mels / 2^layers (reference models/encoders.py:71)
static int sub_features(const ec_config& c) {
  return c.sub_layers == 2 ? c.sub_filters2 * (c.n_mels / 4) : c.sub_filters * (c.n_mels / 2);
}

// Lays the prepared-weights arena out (dry run when arena == nullptr); fills e->w with pointers.
static size_t layout_weights(ec_engine* e, void* arena) {
  Arena a(arena);
  const ec_config& c = e->cfg;
  const size_t es = e->esize;
  Weights& w = e->w;
  auto f32 = [&](size_t n) { return reinterpret_cast<float*>(a.take(n * 4)); };
  auto act = [&](size_t n) { return a.take(n * es * weight_planes(e->precision)); };   // GEMM weight operand (split mode: two planes)
  const int C = c.sub_filters, D0 = c.blocks[0].dim_model;
  w.sub_w = f32(C * 9); w.sub_b = f32(C);
  if (c.sub_layers == 2) { w.sub2_w = act(static_cast<size_t>(c.sub_filters2) * 9 * C); w.sub2_b = f32(c.sub_filters2); }
  else { w.sub2_w = nullptr; w.sub2_b = nullptr; }
  w.lin_w = act(static_cast<size_t>(D0) * sub_features(c)); w.lin_b = f32(D0);
  w.lin_w_perm = c.sub_layers == 1 ? act(static_cast<size_t>(D0) * subsample_fused_cpad(e->precision, C) * (c.n_mels / 2)) : nullptr;
  for (int i = 0; i < c.num_blocks; ++i) {
    const ec_block_cfg& bc = c.blocks[i];
    const int D = bc.dim_model, De = bc.dim_expand, Fr = bc.ff_ratio;
    BlockW& b = w.blk[i];
    auto ffn = [&](FfnW& f, int d) {
      f.ln_w = f32(d); f.ln_b = f32(d);
      f.w1 = act(static_cast<size_t>(Fr) * d * d); f.b1 = f32(Fr * d);
      f.w2 = act(static_cast<size_t>(Fr) * d * d); f.b2 = f32(d);
    };
    ffn(b.ffn1, D);
    b.att_ln_w = f32(D); b.att_ln_b = f32(D); b.u = f32(D); b.v = f32(D);
    b.wqkv = act(static_cast<size_t>(3) * D * D); b.bqkv = f32(3 * D);
    b.wo = act(static_cast<size_t>(D) * D); b.bo = f32(D);
    b.wpos = act(static_cast<size_t>(D) * D); b.bpos = f32(D);
    b.conv_ln_w = f32(D); b.conv_ln_b = f32(D);
    pick_glu(De, &b.glu_nb, &b.glu_tiles);
    b.pw1 = act(static_cast<size_t>(b.glu_tiles) * 2 * b.glu_nb * D); b.pw1_b = f32(b.glu_tiles * 2 * b.glu_nb);
    b.dw_w = f32(De * bc.kernel_size); b.dw_b = f32(De);
    b.pw2 = act(static_cast<size_t>(De) * De); b.pw2_b = f32(De);
    if (D != De) { b.res_w = act(static_cast<size_t>(De) * D); b.res_b = f32(De); } else { b.res_w = nullptr; b.res_b = nullptr; }
    ffn(b.ffn2, De);
    b.norm_w = f32(De); b.norm_b = f32(De);
  }
  if (c.vocab > 0) {
    const int Dl = c.blocks[c.num_blocks - 1].dim_expand;
    w.fc_w = act(static_cast<size_t>(c.vocab) * Dl); w.fc_b = f32(c.vocab);
  } else { w.fc_w = nullptr; w.fc_b = nullptr; }
  return align_up(a.off, 256);
}

struct Shapes {   // per-block frame counts for one (batch, t_mel)
  int t_sub1;                     // frames after the first Conv2d subsampling layer
  int t0;                         // frames after subsampling
  int t_in[EC_MAX_BLOCKS], t_out[EC_MAX_BLOCKS];
  int t_final;
};
static void compute_shapes(const ec_config& c, int t_mel, Shapes* s) {
  int t = (t_mel - 1) / 2 + 1;
  s->t_sub1 = t;
  if (c.sub_layers == 2) t = (t - 1) / 2 + 1;
  s->t0 = t;
  for (int i = 0; i < c.num_blocks; ++i) {
    s->t_in[i] = t;
    if (c.blocks[i].conv_stride > 1) t = (t - 1) / c.blocks[i].conv_stride + 1;
    s->t_out[i] = t;
  }
  s->t_final = t;
}

struct Workspace {
  int* lens; void* sub_a; void *sub_y0, *sub_col; float *xa, *xb; void *xn, *xs, *h; float* qkv; float* ebuf; void *o, *gl, *hc; float* r;
  size_t e_stride;   // bytes between the per-block E buffers
  size_t bytes;
};
static void layout_workspace(const ec_engine* e, int B, int t_mel, void* base, Workspace* ws) {
  const ec_config& c = e->cfg;
  const size_t es = e->esize;
  Shapes sh; compute_shapes(c, t_mel, &sh);
  size_t mx_x = static_cast<size_t>(B) * sh.t0 * c.blocks[0].dim_model, mx_h = 0, mx_qkv = 0, mx_e = 0, mx_g = 0, mx_hc = 0, mx_xs = 0;
  for (int i = 0; i < c.num_blocks; ++i) {
    const ec_block_cfg& b = c.blocks[i];
    const size_t Mi = static_cast<size_t>(B) * sh.t_in[i], Mo = static_cast<size_t>(B) * sh.t_out[i];
    mx_x = std::max(mx_x, std::max(Mi * b.dim_model, Mo * b.dim_expand));
    mx_h = std::max(mx_h, std::max(Mi * b.dim_model, Mo * b.dim_expand) * b.ff_ratio);
    mx_qkv = std::max(mx_qkv, Mi * 3 * b.dim_model);
    const int P = (b.group_size - sh.t_in[i] % b.group_size) % b.group_size;
    mx_e = std::max(mx_e, static_cast<size_t>(2 * (sh.t_in[i] + P) - b.group_size) * b.dim_model);
    mx_g = std::max(mx_g, Mi * b.dim_expand);
    mx_hc = std::max(mx_hc, Mo * b.dim_expand);
    if (b.dim_model != b.dim_expand) mx_xs = std::max(mx_xs, Mo * b.dim_model);
  }
  Arena a(base);
  ws->lens = reinterpret_cast<int*>(a.take(sizeof(int) * (c.num_blocks + 1) * B));
  ws->sub_a = a.take(static_cast<size_t>(B) * sh.t0 * sub_features(c) * es);       // operand of the subsampling Linear
  if (c.sub_layers == 2) {   // channels-last layer-0 map and the im2col operand of the layer-1 GEMM
    ws->sub_y0 = a.take(static_cast<size_t>(B) * sh.t_sub1 * (c.n_mels / 2) * c.sub_filters * es);
    ws->sub_col = a.take(static_cast<size_t>(B) * sh.t0 * (c.n_mels / 4) * 9 * c.sub_filters * es);
  } else { ws->sub_y0 = nullptr; ws->sub_col = nullptr; }
  ws->xa = reinterpret_cast<float*>(a.take(mx_x * 4));
  ws->xb = reinterpret_cast<float*>(a.take(mx_x * 4));
  ws->xn = a.take(mx_x * es);
  ws->xs = a.take(std::max<size_t>(mx_xs, 1) * es);
  ws->h = a.take(mx_h * es);
  ws->qkv = reinterpret_cast<float*>(a.take(mx_qkv * 4));
  ws->e_stride = align_up(mx_e * 4, 256);
  ws->ebuf = reinterpret_cast<float*>(a.take(ws->e_stride * c.num_blocks));
  ws->o = a.take(mx_x * es);
  ws->gl = a.take(mx_g * es);
  ws->hc = a.take(mx_hc * es);
  ws->r = reinterpret_cast<float*>(a.take(std::max<size_t>(mx_hc, 1) * 4));
  ws->bytes = align_up(a.off, 256);
}

// ---- launch accounting / optional per-launch CUDA-event timing -------------------------------------------------------
struct ProfScope {
  ec_engine* e; cudaStream_t st; bool on;
  ProfScope(ec_engine* e_, cudaStream_t st_, int cat, double flops, double bytes) : e(e_), st(st_), on(e_->prof_enabled) {
    e->last_launches++;
    if (!on) return;
    auto ev = [&]() {
      if (e->events_used == e->event_pool.size()) { cudaEvent_t x; cudaEventCreate(&x); e->event_pool.push_back(x); }
      return e->event_pool[e->events_used++];
    };
    ProfEntry pe{cat, flops, bytes, ev(), ev()};
    cudaEventRecord(pe.e0, st);
    e->prof.push_back(pe);
  }
  ~ProfScope() { if (on) cudaEventRecord(e->prof.back().e1, st); }
};

struct LnFuse {            // optional fused LayerNorm epilogue of a GEMM
  int mode = 0; const float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr; void* y = nullptr;
  void* copy_out = nullptr; int copy_stride = 1, fps = 0, fops = 0;
};
static int gemm(ec_engine* e, cudaStream_t st, int cat, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha,
                int act, const float* residual, float* out_f32, void* out_act, int glu_nb = 0, int glu_channels = 0,
                const LnFuse* ln = nullptr, int round_out = 0, int act_f16 = 0) {
  GemmArgs g{};
  g.round_out = round_out; g.act_f16 = act_f16;
  if (ln != nullptr && ln->mode != 0) {
    g.ln_mode = ln->mode; g.ln1_g = ln->g1; g.ln1_b = ln->b1; g.ln2_g = ln->g2; g.ln2_b = ln->b2; g.ln_eps = 1e-6f; g.ln_out = ln->y;
    g.copy_out = ln->copy_out; g.copy_stride = ln->copy_stride; g.frames_per_seq = ln->fps; g.frames_out_per_seq = ln->fops;
  }
  g.A = A; g.W = W; g.M = M; g.N = N; g.K = K; g.bias = bias; g.alpha = alpha; g.act = act;
  g.glu_nb = glu_nb; g.glu_channels = glu_channels;
  const int ncols = glu_nb > 0 ? glu_channels : N;
  g.residual = residual; g.ld_res = ncols; g.out_f32 = out_f32; g.ld_out = ncols; g.out_act = out_act; g.ld_act = ncols;
  // algorithmic work: real (unpadded) dims; every tensor crosses memory once
  const double n_real = glu_nb > 0 ? 2.0 * glu_channels : N;
  const double flops = 2.0 * M * n_real * K;
  const double bytes = (static_cast<double>(M) * K + n_real * K) * e->esize + static_cast<double>(M) * ncols *
                       ((out_f32 ? 4 : 0) + (out_act ? e->esize : 0) + (residual ? 4 : 0));
  if (e->skip_mask >> cat & 1u) return EC_OK;
  ProfScope ps(e, st, cat, flops, bytes);
  return launch_gemm(e->precision, g, st);
}
// whole feed-forward module in one launch (bf16 mode)
static int ffn_fused(ec_engine* e, cudaStream_t st, const void* x_act, const void* w1, const float* b1, const void* w2, const float* b2,
                     int M, int D, int hidden, const float* residual, float* out_f32, const LnFuse& ln) {
  FfnArgs f{};
  f.x_act = x_act; f.w1 = w1; f.b1 = b1; f.w2 = w2; f.b2 = b2; f.M = M; f.D = D; f.hidden = hidden;
  f.residual = residual; f.out_f32 = out_f32;
  f.ln_mode = ln.mode; f.ln1_g = ln.g1; f.ln1_b = ln.b1; f.ln2_g = ln.g2; f.ln2_b = ln.b2; f.ln_eps = 1e-6f; f.ln_out = ln.y;
  const double flops = 4.0 * M * D * hidden;
  const double bytes = 2.0 * (static_cast<double>(M) * D + 2.0 * D * hidden) + static_cast<double>(M) * D * (4 + 4 + (ln.y ? 2 : 0));
  if (e->skip_mask >> PC_FFN_FUSED & 1u) return EC_OK;
  ProfScope ps(e, st, PC_FFN_FUSED, flops, bytes);
  return launch_ffn_fused(f, st);
}
static int lnorm(ec_engine* e, cudaStream_t st, const float* x, int rows, int dim, const float* g, const float* b, void* y_act,
                 float* y_f32, void* copy_out = nullptr, int copy_stride = 1, int fps = 0, int fops = 0) {
  LayerNormArgs a{};
  a.x = x; a.rows = rows; a.dim = dim; a.gamma = g; a.beta = b; a.eps = 1e-6f; a.y_act = y_act; a.y_f32 = y_f32;
  a.copy_out = copy_out; a.copy_stride = copy_stride; a.frames_per_seq = fps; a.frames_out_per_seq = fops;
  const double el = static_cast<double>(rows) * dim;
  ProfScope ps(e, st, PC_LAYERNORM, 8.0 * el, el * (4 + (y_act ? e->esize : 0) + (y_f32 ? 4 : 0)) + (copy_out ? el / copy_stride * e->esize : 0));
  return launch_layernorm(e->precision, a, st);
}

}  // namespace ec

using namespace ec;

extern "C" {

const char* ec_last_error(void) { return g_last_error.c_str(); }
int ec_version(void) { return 100; }

int ec_device_check(void) {
  int dev = 0;
  EC_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  EC_CUDA(cudaGetDeviceProperties(&prop, dev));
  EC_REQUIRE(prop.major == 10, std::string("effconf_b200 needs an sm_100 (B200) device; found sm_") + std::to_string(prop.major) +
                                   std::to_string(prop.minor) + " -- there is no fallback path");
  return EC_OK;
}

int ec_engine_create(const ec_config* cfg, int precision, ec_engine** out) {
  EC_REQUIRE(cfg != nullptr && out != nullptr, "null argument");
  EC_REQUIRE(precision == EC_PREC_TF32 || precision == EC_PREC_BF16 || precision == EC_PREC_BF16X2, "unknown precision");
  EC_REQUIRE(cfg->num_blocks >= 1 && cfg->num_blocks <= EC_MAX_BLOCKS, "num_blocks out of range");
  EC_REQUIRE(cfg->n_mels > 0 && cfg->n_mels % 2 == 0 && cfg->sub_filters > 0, "bad front-end config");
  EC_REQUIRE(cfg->sub_layers >= 0 && cfg->sub_layers <= 2, "1 or 2 Conv2d subsampling layers are supported");
  if (cfg->sub_layers == 2)
    EC_REQUIRE(cfg->n_mels % 4 == 0 && cfg->sub_filters % 8 == 0 && cfg->sub_filters2 > 0, "two-layer subsampling needs n_mels % 4 == 0 and filters % 8 == 0");
  for (int i = 0; i < cfg->num_blocks; ++i) {
    const ec_block_cfg& b = cfg->blocks[i];
    EC_REQUIRE(b.dim_model > 0 && b.dim_expand > 0 && b.num_heads > 0 && b.ff_ratio > 0, "bad block dims");
    EC_REQUIRE(b.group_size % 2 == 1, "attention group size must be odd");
    EC_REQUIRE((b.group_size * b.dim_model) % b.num_heads == 0, "G*D must be divisible by H");
    EC_REQUIRE(b.conv_stride == 1 || b.conv_stride == 2, "conv_stride must be 1 or 2");
    EC_REQUIRE(b.conv_stride == 1 || b.dim_model != b.dim_expand, "strided block without expansion (MaxPool residual) is not implemented");
    if (i > 0) EC_REQUIRE(cfg->blocks[i - 1].dim_expand == b.dim_model, "block dims do not chain");
  }
  ec_engine* e = new ec_engine();
  e->cfg = *cfg; if (e->cfg.sub_layers == 0) e->cfg.sub_layers = 1;
  e->precision = precision; e->esize = act_esize(precision); e->prepared = false;
  // Fused-LayerNorm GEMM epilogues pay off with 2-byte (bf16) and TF32 operands; in the split mode the LayerNorm passes inside a
  // one-CTA-per-SM epilogue cost more than the stand-alone row kernel at full occupancy (B = 32 x 1000 forward: 2.97 ms fused,
  // 2.77 ms unfused; tools/fwd_options.py, profiles/r2/forward_fusion_options.txt).  ec_engine_set_fuse_ln overrides.
  e->fuse_ln = precision != EC_PREC_BF16X2;
  e->weight_bytes = layout_weights(e, nullptr);
  *out = e;
  return EC_OK;
}
void ec_engine_destroy(ec_engine* e) {
  if (e == nullptr) return;
  for (cudaEvent_t ev : e->event_pool) cudaEventDestroy(ev);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->side) cudaStreamDestroy(e->side);
  delete e;
}
size_t ec_engine_weight_bytes(const ec_engine* e) { return e->weight_bytes; }
size_t ec_engine_workspace_bytes(const ec_engine* e, int batch, int t_mel) {
  Workspace ws; layout_workspace(e, batch, t_mel, nullptr, &ws);
  return ws.bytes;
}
int ec_engine_out_frames(const ec_engine* e, int t_mel) {
  Shapes sh; compute_shapes(e->cfg, t_mel, &sh);
  return sh.t_final;
}
int ec_engine_relpos_rows(const ec_engine* e, int t_mel, int32_t* rows_per_block, int32_t* frames_per_block) {
  Shapes sh; compute_shapes(e->cfg, t_mel, &sh);
  for (int i = 0; i < e->cfg.num_blocks; ++i) {
    const int G = e->cfg.blocks[i].group_size, T = sh.t_in[i];
    const int P = (G - T % G) % G;
    if (rows_per_block) rows_per_block[i] = 2 * (T + P) - G;
    if (frames_per_block) frames_per_block[i] = T;
  }
  return EC_OK;
}

int ec_engine_prepare(ec_engine* e, const ec_raw_weights* raw, void* arena, void* stream_) {
  EC_REQUIRE(e && raw && arena, "null argument");
  EC_TRY(ec_device_check());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  layout_weights(e, arena);
  const ec_config& c = e->cfg;
  const int prec = e->precision;
  Weights& w = e->w;
  auto cp = [&](float* dst, const float* src, size_t n) -> int {
    EC_REQUIRE(src != nullptr, "missing raw weight pointer");
    EC_CUDA(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, st));
    return EC_OK;
  };
  // weight operand of a GEMM; twin = element distance to the swapped plane of the split mode (default: right behind the n elements)
  auto cast = [&](void* dst, const float* src, size_t n, size_t twin = 0) -> int {
    EC_REQUIRE(src != nullptr, "missing raw weight pointer");
    return launch_cast_weight(prec, src, dst, n, twin != 0 ? twin : n, st);
  };
  const int C = c.sub_filters, D0 = c.blocks[0].dim_model;
  EC_REQUIRE(raw->sub_conv_w && raw->sub_conv_b && raw->sub_bn_w && raw->sub_bn_b && raw->sub_bn_rm && raw->sub_bn_rv, "missing subsampling weights");
  EC_TRY(launch_fold_bn(raw->sub_conv_w, raw->sub_conv_b, raw->sub_bn_w, raw->sub_bn_b, raw->sub_bn_rm, raw->sub_bn_rv, 1e-5f, C, 9, w.sub_w, w.sub_b, st));
  if (c.sub_layers == 2) {
    EC_REQUIRE(raw->sub2_conv_w && raw->sub2_conv_b && raw->sub2_bn_w && raw->sub2_bn_b && raw->sub2_bn_rm && raw->sub2_bn_rv,
               "missing weights of the second subsampling layer");
    EC_TRY(launch_conv2_weight_prep(prec, raw->sub2_conv_w, raw->sub2_conv_b, raw->sub2_bn_w, raw->sub2_bn_b, raw->sub2_bn_rm, raw->sub2_bn_rv,
                                    1e-5f, c.sub_filters2, C, w.sub2_w, w.sub2_b, st));
    EC_REQUIRE(raw->lin_w != nullptr, "missing raw weight pointer");
    EC_TRY(launch_linear_weight_permute(prec, raw->lin_w, D0, c.sub_filters2, c.n_mels / 4, w.lin_w, st));
  } else {
    EC_TRY(cast(w.lin_w, raw->lin_w, static_cast<size_t>(D0) * sub_features(c)));
    EC_TRY(launch_linear_weight_permute(prec, raw->lin_w, D0, C, c.n_mels / 2, w.lin_w_perm, st, subsample_fused_cpad(prec, C)));
  }
  EC_TRY(cp(w.lin_b, raw->lin_b, D0));
  for (int i = 0; i < c.num_blocks; ++i) {
    const ec_block_cfg& bc = c.blocks[i];
    const ec_block_raw& r = raw->blocks[i];
    BlockW& b = w.blk[i];
    const int D = bc.dim_model, De = bc.dim_expand, Fr = bc.ff_ratio;
    auto ffn = [&](FfnW& f, const ec_ffn_raw& fr, int d) -> int {
      EC_TRY(cp(f.ln_w, fr.ln_w, d)); EC_TRY(cp(f.ln_b, fr.ln_b, d));
      EC_TRY(cast(f.w1, fr.w1, static_cast<size_t>(Fr) * d * d)); EC_TRY(cp(f.b1, fr.b1, Fr * d));
      EC_TRY(cast(f.w2, fr.w2, static_cast<size_t>(Fr) * d * d)); EC_TRY(cp(f.b2, fr.b2, d));
      return EC_OK;
    };
    EC_TRY(ffn(b.ffn1, r.ffn1, D));
    EC_TRY(cp(b.att_ln_w, r.att_ln_w, D)); EC_TRY(cp(b.att_ln_b, r.att_ln_b, D));
    EC_TRY(cp(b.u, r.u, D)); EC_TRY(cp(b.v, r.v, D));
    const size_t dd = static_cast<size_t>(D) * D;
    uint8_t* wqkv = reinterpret_cast<uint8_t*>(b.wqkv);
    EC_TRY(cast(wqkv, r.wq, dd, 3 * dd)); EC_TRY(cast(wqkv + dd * e->esize, r.wk, dd, 3 * dd)); EC_TRY(cast(wqkv + 2 * dd * e->esize, r.wv, dd, 3 * dd));
    EC_TRY(cp(b.bqkv, r.bq, D)); EC_TRY(cp(b.bqkv + D, r.bk, D)); EC_TRY(cp(b.bqkv + 2 * D, r.bv, D));
    EC_TRY(cast(b.wo, r.wo, dd)); EC_TRY(cp(b.bo, r.bo, D));
    EC_TRY(cast(b.wpos, r.wpos, dd)); EC_TRY(cp(b.bpos, r.bpos, D));
    EC_TRY(cp(b.conv_ln_w, r.conv_ln_w, D)); EC_TRY(cp(b.conv_ln_b, r.conv_ln_b, D));
    EC_REQUIRE(r.pw1_w && r.pw1_b, "missing pw1 weights");
    EC_TRY(launch_glu_interleave(prec, r.pw1_w, r.pw1_b, De, D, b.glu_nb, b.glu_tiles, b.pw1, b.pw1_b, st));
    EC_REQUIRE(r.dw_w && r.dw_b && r.bn_w && r.bn_b && r.bn_rm && r.bn_rv, "missing depthwise / BatchNorm weights");
    EC_TRY(launch_fold_bn(r.dw_w, r.dw_b, r.bn_w, r.bn_b, r.bn_rm, r.bn_rv, 1e-5f, De, bc.kernel_size, b.dw_w, b.dw_b, st));
    EC_TRY(cast(b.pw2, r.pw2_w, static_cast<size_t>(De) * De)); EC_TRY(cp(b.pw2_b, r.pw2_b, De));
    if (D != De) { EC_TRY(cast(b.res_w, r.res_w, static_cast<size_t>(De) * D)); EC_TRY(cp(b.res_b, r.res_b, De)); }
    EC_TRY(ffn(b.ffn2, r.ffn2, De));
    EC_TRY(cp(b.norm_w, r.norm_w, De)); EC_TRY(cp(b.norm_b, r.norm_b, De));
  }
  if (c.vocab > 0) {
    const int Dl = c.blocks[c.num_blocks - 1].dim_expand;
    EC_TRY(cast(w.fc_w, raw->fc_w, static_cast<size_t>(c.vocab) * Dl)); EC_TRY(cp(w.fc_b, raw->fc_b, c.vocab));
  }
  e->prepared = true;
  return EC_OK;
}

int ec_engine_forward(ec_engine* e, int B, int t_mel, const float* mel, const long long* x_len, const void* const* relpos,
                      void* workspace, float* out_x, float* logits, long long* out_len, void* stream_) {
  EC_REQUIRE(e && mel && relpos && workspace, "null argument");
  EC_REQUIRE(e->prepared, "ec_engine_prepare has not been called");
  EC_REQUIRE(B > 0 && t_mel > 0, "empty batch");
  EC_REQUIRE(out_x != nullptr || logits != nullptr, "no output requested");
  EC_REQUIRE(logits == nullptr || e->cfg.vocab > 0, "engine was built without an fc head");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const ec_config& c = e->cfg;
  const Weights& w = e->w;
  const int prec = e->precision;
  Shapes sh; compute_shapes(c, t_mel, &sh);
  Workspace ws; layout_workspace(e, B, t_mel, workspace, &ws);

  e->last_launches = 0;
  e->prof.clear(); e->events_used = 0;
  if (e->prof_enabled) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    EC_REQUIRE(cs == cudaStreamCaptureStatusNone, "profiling mode cannot be used under CUDA graph capture");
  }
  const double es = static_cast<double>(e->esize);

  BlockStrides bs{}; bs.n = c.num_blocks; bs.sub_layers = c.sub_layers;
  for (int i = 0; i < c.num_blocks; ++i) bs.s[i] = c.blocks[i].conv_stride;
  { ProfScope ps(e, st, PC_MISC, 0, 0); EC_TRY(launch_stage_lengths(x_len, B, t_mel, bs, ws.lens, st)); }

  // ---- fork: all positional projections (weights x constant tables) on the side stream ----
  if (e->side == nullptr) {
    EC_CUDA(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
    EC_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    EC_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  }
  EC_CUDA(cudaEventRecord(e->ev_fork, st));
  EC_CUDA(cudaStreamWaitEvent(e->side, e->ev_fork, 0));
  for (int i = 0; i < c.num_blocks; ++i) {
    const ec_block_cfg& bc = c.blocks[i];
    const int D = bc.dim_model, T = sh.t_in[i], G = bc.group_size, P = (G - T % G) % G, e_rows = 2 * (T + P) - G;
    EC_REQUIRE(relpos[i] != nullptr, "missing relative position table");
    float* eb = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws.ebuf) + i * ws.e_stride);
    // activation-type (bf16 / packed) E straight from the epilogue, else TF32-rounded fp32
    const bool a16 = prec == EC_PREC_BF16X2 || (prec == EC_PREC_BF16 && ((G * D) / bc.num_heads) % 2 == 0 && D % 2 == 0);
    const int af16 = (prec == EC_PREC_BF16X2 && attn_operands_f16(D, bc.num_heads, G)) ? 1 : 0;      // split mode: plain bf16 E
    EC_TRY(gemm(e, e->side, PC_POS, relpos[i], w.blk[i].wpos, e_rows, D, D, w.blk[i].bpos, 1.f, GEMM_ACT_NONE, nullptr, a16 ? nullptr : eb,
                a16 ? eb : nullptr, 0, 0, nullptr, 1, af16));
  }
  EC_CUDA(cudaEventRecord(e->ev_join, e->side));
  bool joined = false;

  // ---- front end: Conv2d+BN+Swish producer(s), then Linear (K = C_last * F / 2^layers) ----
  const int feat = sub_features(c);
  if (c.sub_layers == 2) {
    // layer 0 channels-last -> im2col -> layer 1 on the tensor cores (Swish in the epilogue); see subsample.cu
    const int C = c.sub_filters, C2 = c.sub_filters2, F1 = c.n_mels / 2, F2 = c.n_mels / 4, T1 = sh.t_sub1, T2 = sh.t0;
    {
      SubsampleArgs sa{mel, w.sub_w, w.sub_b, B, c.n_mels, t_mel, C, ws.sub_y0};
      ProfScope ps(e, st, PC_SUBSAMPLE, 18.0 * B * T1 * F1 * C, 4.0 * B * c.n_mels * t_mel + es * B * T1 * F1 * C);
      EC_TRY(launch_subsample_conv_cl(prec, sa, st));
    }
    {
      ProfScope ps(e, st, PC_IM2COL, 0, es * (static_cast<double>(B) * T1 * F1 * C + static_cast<double>(B) * T2 * F2 * 9 * C));
      EC_TRY(launch_im2col_3x3s2(prec, ws.sub_y0, B, T1, F1, C, ws.sub_col, st));
    }
    EC_TRY(gemm(e, st, PC_SUB_CONV2, ws.sub_col, w.sub2_w, B * T2 * F2, C2, 9 * C, w.sub2_b, 1.f, GEMM_ACT_SWISH, nullptr, nullptr, ws.sub_a));
  }
  const int D0 = c.blocks[0].dim_model;
  const bool front_fused = c.sub_layers == 1 && e->fuse_front && subsample_fused_fits(prec, c.n_mels, c.sub_filters, D0);
  if (c.sub_layers == 1 && !front_fused) {
    SubsampleArgs sa{mel, w.sub_w, w.sub_b, B, c.n_mels, t_mel, c.sub_filters, ws.sub_a};
    ProfScope ps(e, st, PC_SUBSAMPLE, 18.0 * B * sh.t0 * feat, 4.0 * B * c.n_mels * t_mel + es * B * sh.t0 * feat);
    if (!(e->skip_mask >> PC_SUBSAMPLE & 1u)) EC_TRY(launch_subsample_conv(prec, sa, st));
  }
  float* x = ws.xa; float* x_alt = ws.xb;

  // GEMM followed by the LayerNorm(s) the next module needs: in the GEMM's own epilogue when the row fits one tile (N <= 256) and
  // fuse_ln is on, else as separate row kernels with identical semantics (LnFuse modes 1 / 2 of GemmArgs).
  auto gemm_ln = [&](int cat, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha, const float* residual,
                     float* out_f32, const LnFuse& ln) -> int {
    if (e->fuse_ln && N <= 256) return gemm(e, st, cat, A, W, M, N, K, bias, alpha, GEMM_ACT_NONE, residual, out_f32, nullptr, 0, 0, &ln);
    EC_TRY(gemm(e, st, cat, A, W, M, N, K, bias, alpha, GEMM_ACT_NONE, residual, out_f32, nullptr));
    if (ln.mode == 1) return lnorm(e, st, out_f32, M, N, ln.g1, ln.b1, ln.y, nullptr, ln.copy_out, ln.copy_stride, ln.fps, ln.fops);
    // mode 2: out <- LN1(out) in place (a warp holds its whole row in registers before it writes), then y = LN2(out) or its rounded copy
    if (ln.g2 == nullptr) return lnorm(e, st, out_f32, M, N, ln.g1, ln.b1, ln.y, out_f32);
    EC_TRY(lnorm(e, st, out_f32, M, N, ln.g1, ln.b1, nullptr, out_f32));
    return lnorm(e, st, out_f32, M, N, ln.g2, ln.b2, ln.y, nullptr);
  };
  // whole feed-forward module: one cluster kernel when the shape fits (bf16 mode), else W1 GEMM (+Swish) and W2 GEMM (+residual, LN)
  auto ffn = [&](const FfnW& f, int M, int D, int hidden, const float* residual, float* out_f32, const LnFuse& ln) -> int {
    if (e->fuse_ln && e->fuse_ffn && prec == EC_PREC_BF16 && ffn_fused_fits(M, D, hidden))
      return ffn_fused(e, st, ws.xn, f.w1, f.b1, f.w2, f.b2, M, D, hidden, residual, out_f32, ln);
    EC_TRY(gemm(e, st, PC_FFN_W1, ws.xn, f.w1, M, hidden, D, f.b1, 1.f, GEMM_ACT_SWISH, nullptr, nullptr, ws.h));
    return gemm_ln(PC_FFN_W2, ws.h, f.w2, M, D, hidden, f.b2, 0.5f, residual, out_f32, ln);
  };

  {
    LnFuse ln;   // x0 = Linear(sub), xn = LN_ffn1(x0) for block 0
    ln.mode = 1; ln.g1 = w.blk[0].ffn1.ln_w; ln.b1 = w.blk[0].ffn1.ln_b; ln.y = ws.xn;
    if (front_fused) {
      // conv + BN + Swish producers feed the Linear's A tiles in shared memory: the (B*T/2) x (C*F/2) operand is never written
      {
        SubFusedArgs fa{mel, w.sub_w, w.sub_b, w.lin_w_perm, w.lin_b, B, c.n_mels, t_mel, c.sub_filters, D0, x};
        ProfScope ps(e, st, PC_SUBSAMPLE, (18.0 + 2.0 * D0) * B * sh.t0 * feat,
                     4.0 * B * c.n_mels * t_mel + es * static_cast<double>(D0) * feat + 4.0 * B * sh.t0 * D0);
        if (!(e->skip_mask >> PC_SUBSAMPLE & 1u)) EC_TRY(launch_subsample_linear_fused(prec, fa, st));
      }
      EC_TRY(lnorm(e, st, x, B * sh.t0, D0, ln.g1, ln.b1, ln.y, nullptr, ln.copy_out, ln.copy_stride, ln.fps, ln.fops));
    } else {
      EC_TRY(gemm_ln(PC_LIN, ws.sub_a, w.lin_w, B * sh.t0, D0, feat, w.lin_b, 1.f, nullptr, x, ln));
    }
  }

  for (int i = 0; i < c.num_blocks; ++i) {
    const ec_block_cfg& bc = c.blocks[i];
    const BlockW& b = w.blk[i];
    const int D = bc.dim_model, De = bc.dim_expand, Fr = bc.ff_ratio;
    const int T = sh.t_in[i], To = sh.t_out[i];
    const int M = B * T, Mo = B * To;
    const int* lens = ws.lens + static_cast<size_t>(i) * B;
    const bool proj = D != De;
    const bool last = i == c.num_blocks - 1;
    // FFN1: x1 = x + 0.5 * W2 swish(W1 LN(x)); also emits xn = LN_att(x1)
    {
      LnFuse ln;
      ln.mode = 1; ln.g1 = b.att_ln_w; ln.b1 = b.att_ln_b; ln.y = ws.xn;
      EC_TRY(ffn(b.ffn1, M, D, Fr * D, x, x_alt, ln));
    }
    std::swap(x, x_alt);
    // MHSA: x2 = x1 + Wo attn(LN(x1)); also emits xn = LN_conv(x2) and the strided copy xs (conv_res operand)
    // q|k|v and E feed the attention kernel: TF32-rounded fp32 in parity mode, bf16 in fast mode
    const bool a16 = prec == EC_PREC_BF16X2 || (prec == EC_PREC_BF16 && ((bc.group_size * D) / bc.num_heads) % 2 == 0 && D % 2 == 0);
    const int af16 = (prec == EC_PREC_BF16X2 && attn_operands_f16(D, bc.num_heads, bc.group_size)) ? 1 : 0;   // split mode: plain bf16 q|k|v
    EC_TRY(gemm(e, st, PC_QKV, ws.xn, b.wqkv, M, 3 * D, D, b.bqkv, 1.f, GEMM_ACT_NONE, nullptr, a16 ? nullptr : ws.qkv, a16 ? ws.qkv : nullptr, 0, 0, nullptr, 1, af16));
    const int G = bc.group_size, P = (G - T % G) % G, e_rows = 2 * (T + P) - G;
    if (!joined) { EC_CUDA(cudaStreamWaitEvent(st, e->ev_join, 0)); joined = true; }     // join: E_i of every block is ready
    {
      const float* eb = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(ws.ebuf) + i * ws.e_stride);
      AttnArgs aa{ws.qkv, eb, b.u, b.v, lens, B, T, D, bc.num_heads, G, ws.o, D, (prec == EC_PREC_BF16 && !a16) ? 1 : 0, af16};
      const double Tg = static_cast<double>(T + P) / G, dh = static_cast<double>(G) * D / bc.num_heads;
      ProfScope ps(e, st, PC_ATTN, B * bc.num_heads * (4.0 * Tg * Tg * dh + 2.0 * Tg * (2 * Tg - 1) * dh),
                   4.0 * M * 3 * D + 4.0 * e_rows * D + es * M * D);
      if (!(e->skip_mask >> PC_ATTN & 1u)) EC_TRY(launch_relpos_attention(prec, aa, st));
    }
    {
      LnFuse ln;
      ln.mode = 1; ln.g1 = b.conv_ln_w; ln.b1 = b.conv_ln_b; ln.y = ws.xn;
      if (proj) { ln.copy_out = ws.xs; ln.copy_stride = bc.conv_stride; ln.fps = T; ln.fops = To; }
      EC_TRY(gemm_ln(PC_OUT, ws.o, b.wo, M, D, D, b.bo, 1.f, x, x_alt, ln));
    }
    std::swap(x, x_alt);
    // Conv module: x3 = conv_res(x2) + pw2 swish(bn(dw(glu(pw1 LN(x2))))); also emits xn = LN_ffn2(x3)
    EC_TRY(gemm(e, st, PC_PW1_GLU, ws.xn, b.pw1, M, b.glu_tiles * 2 * b.glu_nb, D, b.pw1_b, 1.f, GEMM_ACT_NONE, nullptr, nullptr, ws.gl, b.glu_nb, De));
    {
      DwConvArgs da{ws.gl, b.dw_w, b.dw_b, B, T, De, bc.kernel_size, bc.conv_stride, ws.hc};
      ProfScope ps(e, st, PC_DWCONV, 2.0 * Mo * De * bc.kernel_size, es * (static_cast<double>(M) * De + static_cast<double>(Mo) * De));
      if (!(e->skip_mask >> PC_DWCONV & 1u)) EC_TRY(launch_dwconv_bn_swish(prec, da, st));
    }
    const float* res = x;
    if (proj) {
      EC_TRY(gemm(e, st, PC_RES, ws.xs, b.res_w, Mo, De, D, b.res_b, 1.f, GEMM_ACT_NONE, nullptr, ws.r, nullptr));
      res = ws.r;
    }
    {
      LnFuse ln;
      ln.mode = 1; ln.g1 = b.ffn2.ln_w; ln.b1 = b.ffn2.ln_b; ln.y = ws.xn;
      EC_TRY(gemm_ln(PC_PW2, ws.hc, b.pw2, Mo, De, De, b.pw2_b, 1.f, res, x_alt, ln));
    }
    std::swap(x, x_alt);
    // FFN2 + block LayerNorm: normalises in place (block norm) and emits the next GEMM's operand
    // (LN_ffn1 of the next block, or the rounded block output for the fc head)
    {
      float* y = (last && out_x != nullptr) ? out_x : x_alt;
      LnFuse ln;
      ln.mode = 2; ln.g1 = b.norm_w; ln.b1 = b.norm_b;
      if (!last) { ln.g2 = w.blk[i + 1].ffn1.ln_w; ln.b2 = w.blk[i + 1].ffn1.ln_b; ln.y = ws.xn; }
      else { ln.g2 = nullptr; ln.b2 = nullptr; ln.y = logits != nullptr ? ws.xn : nullptr; }
      EC_TRY(ffn(b.ffn2, Mo, De, Fr * De, x, y, ln));
      if (!(last && out_x != nullptr)) x_alt = x, x = y;
    }
  }
  if (logits != nullptr) {
    const int Dl = c.blocks[c.num_blocks - 1].dim_expand;
    EC_TRY(gemm(e, st, PC_FC, ws.xn, w.fc_w, B * sh.t_final, c.vocab, Dl, w.fc_b, 1.f, GEMM_ACT_NONE, nullptr, logits, nullptr));
  }
  if (out_len != nullptr) { ProfScope ps(e, st, PC_MISC, 0, 0); EC_TRY(launch_i32_to_i64(ws.lens + static_cast<size_t>(c.num_blocks) * B, B, out_len, st)); }
  return EC_OK;
}

// ---------------------------------------------------------------- profiling / accounting
int ec_profile_categories(void) { return PC_COUNT; }
const char* ec_profile_category_name(int cat) { return (cat >= 0 && cat < PC_COUNT) ? kProfNames[cat] : ""; }
int ec_engine_set_profiling(ec_engine* e, int enabled) { e->prof_enabled = enabled != 0; return EC_OK; }
/* option 0: fuse LayerNorm into the GEMM epilogues (default 1).  Global option via ec_set_pdl: programmatic dependent launch. */
int ec_engine_set_fuse_ln(ec_engine* e, int enabled) { e->fuse_ln = enabled != 0; return EC_OK; }
int ec_engine_set_fuse_ffn(ec_engine* e, int enabled) { e->fuse_ffn = enabled != 0; return EC_OK; }
int ec_engine_set_fuse_front(ec_engine* e, int enabled) { e->fuse_front = enabled != 0; return EC_OK; }
int ec_engine_set_skip_mask(ec_engine* e, unsigned mask) { e->skip_mask = mask; return EC_OK; }
int ec_debug_gemm_timeline(int enable, unsigned long long* out12) { return gemm_timeline(enable, out12); }
int ec_debug_gemm_block_n(int block_n) { return gemm_block_n_override(block_n); }
int ec_debug_ffn_timeline(int enable, unsigned long long* out192) { return ffn_timeline(enable, out192); }
int ec_set_pdl(int enabled) { g_pdl = enabled != 0 ? 1 : 0; return EC_OK; }
int ec_engine_last_launches(const ec_engine* e) { return e->last_launches; }
/* sums the CUDA-event durations of the last (eager, profiled) forward per category; synchronises the recorded events */
int ec_engine_profile_read(ec_engine* e, double* ms, double* flops, double* bytes, int32_t* launches) {
  for (int i = 0; i < PC_COUNT; ++i) { ms[i] = 0; flops[i] = 0; bytes[i] = 0; launches[i] = 0; }
  for (const ProfEntry& pe : e->prof) {
    EC_CUDA(cudaEventSynchronize(pe.e1));
    float t = 0.f;
    EC_CUDA(cudaEventElapsedTime(&t, pe.e0, pe.e1));
    ms[pe.cat] += t; flops[pe.cat] += pe.flops; bytes[pe.cat] += pe.bytes; launches[pe.cat] += 1;
  }
  return EC_OK;
}

// ---------------------------------------------------------------- CTC head
size_t ec_ctc_scratch_bytes(int batch, int t, int vocab) {
  (void)vocab;
  return align_up(static_cast<size_t>(batch) * t * 4, 256) * 2 + align_up(static_cast<size_t>(batch) * 4, 256);
}
static void ctc_scratch(void* scratch, int B, int T, float** lse, int** amax, int** len32) {
  uint8_t* p = reinterpret_cast<uint8_t*>(scratch);
  *lse = reinterpret_cast<float*>(p);
  *amax = reinterpret_cast<int*>(p + align_up(static_cast<size_t>(B) * T * 4, 256));
  *len32 = reinterpret_cast<int*>(p + 2 * align_up(static_cast<size_t>(B) * T * 4, 256));
}
int ec_ctc_loss(const float* logits, int B, int T, int V, const long long* logits_len, const long long* targets, int target_stride,
                const long long* target_len, void* scratch, float* loss_per_utt, float* loss_mean, void* stream_) {
  EC_REQUIRE(logits && logits_len && targets && target_len && scratch && loss_per_utt, "null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  float* lse; int* amax; int* len32;
  ctc_scratch(scratch, B, T, &lse, &amax, &len32);
  EC_TRY(launch_i64_to_i32(logits_len, B, len32, T, st));
  EC_TRY(launch_logsoftmax_argmax(logits, B * T, V, lse, amax, st));
  return launch_ctc_loss(logits, lse, B, T, V, len32, targets, target_stride, target_len, loss_per_utt, loss_mean, st);
}
size_t ec_ctc_grad_work_bytes(int batch, int t, int target_stride) { return ctc_grad_work_bytes(batch, t, target_stride); }
int ec_ctc_loss_grad(const float* logits, int B, int T, int V, const long long* logits_len, const long long* targets, int target_stride,
                     const long long* target_len, void* scratch, void* work, float* loss_per_utt, float* loss_mean, float* grad_logits,
                     void* stream_) {
  EC_REQUIRE(logits && logits_len && targets && target_len && scratch && work && loss_per_utt && grad_logits, "null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  float* lse; int* amax; int* len32;
  ctc_scratch(scratch, B, T, &lse, &amax, &len32);
  EC_TRY(launch_i64_to_i32(logits_len, B, len32, T, st));
  EC_TRY(launch_logsoftmax_argmax(logits, B * T, V, lse, amax, st));
  // gradient of the MEAN over the batch (reference models/losses.py:71): every utterance's gradient is scaled by 1 / B
  EC_TRY(launch_ctc_grad(logits, lse, B, T, V, len32, targets, target_stride, target_len, reinterpret_cast<float*>(work), 1.f / B,
                         loss_per_utt, grad_logits, st));
  if (loss_mean != nullptr) EC_TRY(launch_mean(loss_per_utt, B, loss_mean, st));
  return EC_OK;
}
int ec_ctc_greedy(const float* logits, int B, int T, int V, const long long* logits_len, void* scratch, int32_t* ids, int32_t* counts,
                  void* stream_) {
  EC_REQUIRE(logits && logits_len && scratch && ids && counts, "null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  float* lse; int* amax; int* len32;
  ctc_scratch(scratch, B, T, &lse, &amax, &len32);
  EC_TRY(launch_i64_to_i32(logits_len, B, len32, T, st));
  EC_TRY(launch_logsoftmax_argmax(logits, B * T, V, lse, amax, st));
  return launch_greedy_collapse(amax, B, T, len32, ids, counts, st);
}

// ---------------------------------------------------------------- backward operators (training step)
size_t ec_op_layernorm_bwd_work_bytes(int dim) { return layernorm_bwd_work_bytes(dim); }
int ec_op_layernorm_bwd(const float* x, const float* dy, int rows, int dim, const float* gamma, float eps, float* dx, int accumulate,
                        float* dgamma, float* dbeta, void* work, void* stream) {
  return launch_layernorm_bwd(x, dy, rows, dim, gamma, eps, dx, accumulate, dgamma, dbeta, reinterpret_cast<float*>(work),
                              reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_layernorm_bwd_emit(const float* x, const float* dy, int rows, int dim, const float* gamma, float eps, float* dx, int accumulate,
                             float* dgamma, float* dbeta, void* work, int emit_precision, void* emit_out, float emit_scale,
                             const unsigned long long* drop_counter, float drop_p, unsigned drop_site, void* stream) {
  return launch_layernorm_bwd(x, dy, rows, dim, gamma, eps, dx, accumulate, dgamma, dbeta, reinterpret_cast<float*>(work),
                              reinterpret_cast<cudaStream_t>(stream), emit_precision, emit_out, emit_scale, drop_counter, drop_p, drop_site);
}
#define EC_ST(s) reinterpret_cast<cudaStream_t>(s)
int ec_op_cast_scaled(int precision, const float* src, float scale, size_t n, void* dst, void* stream) {
  return launch_cast_scaled(precision, src, scale, n, dst, EC_ST(stream));
}
int ec_op_swish_fwd(int precision, const void* z, size_t n, void* h, void* stream) { return launch_swish_fwd(precision, z, n, h, EC_ST(stream)); }
int ec_op_glu_fwd(int precision, const void* zg, size_t rows, int channels, void* out, void* stream) {
  return launch_glu_fwd(precision, zg, rows, channels, out, EC_ST(stream));
}
int ec_op_strided_rows(int precision, const float* x, int batch, int t, int dim, int stride, void* out, void* stream) {
  return launch_strided_rows(precision, x, batch, t, dim, stride, out, EC_ST(stream));
}
int ec_op_strided_rows_bwd(const float* d, int batch, int t, int dim, int stride, float* dx, void* stream) {
  return launch_strided_rows_bwd(d, batch, t, dim, stride, dx, EC_ST(stream));
}
int ec_op_subsample_conv_raw(const float* mel, const float* w, const float* b, int batch, int n_mels, int t, int channels, float* y, void* stream) {
  SubsampleArgs a{mel, w, b, batch, n_mels, t, channels, y};
  return launch_subsample_conv_raw(a, EC_ST(stream));
}
size_t ec_op_col_stats_work_bytes(int cols) { return col_stats_work_bytes(cols); }
int ec_op_col_stats(const float* y, size_t rows, int cols, float* stats, void* work, void* stream) {
  return launch_col_stats(y, rows, cols, stats, reinterpret_cast<float*>(work), EC_ST(stream));
}
int ec_op_group_stats_merge(const float* col_stats, int channels, int group, size_t rows, float* ch_stats, void* stream) {
  return launch_group_stats_merge(col_stats, channels, group, rows, ch_stats, EC_ST(stream));
}
int ec_op_group_expand(const float* in, int n_vec, int channels, int group, float* out, void* stream) {
  return launch_group_expand(in, n_vec, channels, group, out, EC_ST(stream));
}
int ec_op_group_sum(const float* in, int n_vec, int channels, int group, float* out, void* stream) {
  return launch_group_sum(in, n_vec, channels, group, out, EC_ST(stream));
}
size_t ec_op_subsample_wgrad_work_bytes(int channels, int n_mels, int batch, int t) { return subsample_wgrad_work_bytes(channels, n_mels, batch, t); }
int ec_op_subsample_wgrad(const float* dy, const float* mel, int batch, int n_mels, int t, int channels, float* dw, float* db, void* work,
                          void* stream) {
  return launch_subsample_wgrad(dy, mel, batch, n_mels, t, channels, dw, db, reinterpret_cast<float*>(work), EC_ST(stream));
}
#undef EC_ST
size_t ec_op_relpos_attention_bwd_work_bytes(int batch, int t, int dim, int heads, int group) {
  return attention_bwd_work_bytes(batch, t, dim, heads, group);
}
int ec_op_relpos_attention_bwd(int precision, const void* qkv, const void* E, const float* u, const float* v, const int32_t* x_len, int batch,
                               int t, int dim, int heads, int group, const float* d_out, float* dqkv, float* dE, float* du, float* dv,
                               void* work, void* stream) {
  AttnArgs a{qkv, E, u, v, x_len, batch, t, dim, heads, group, nullptr, dim, 0,
             (precision == EC_PREC_BF16X2 && attn_operands_f16(dim, heads, group)) ? 1 : 0};
  return launch_relpos_attention_bwd(precision, a, d_out, dqkv, dE, du, dv, work, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_relpos_attention_bwd_act(int precision, const void* qkv, const void* E, const float* u, const float* v, const int32_t* x_len, int batch,
                                   int t, int dim, int heads, int group, const float* d_out, float* dqkv_f32, void* dqkv_act, float* dE,
                                   float* du, float* dv, void* work, void* stream) {
  AttnArgs a{qkv, E, u, v, x_len, batch, t, dim, heads, group, nullptr, dim, 0,
             (precision == EC_PREC_BF16X2 && attn_operands_f16(dim, heads, group)) ? 1 : 0};
  return launch_relpos_attention_bwd(precision, a, d_out, dqkv_f32, dE, du, dv, work, reinterpret_cast<cudaStream_t>(stream), dqkv_act);
}
size_t ec_op_conv_train_work_bytes(int channels, int k) { return conv_train_work_bytes(channels, k); }
int ec_op_dwconv_raw(int precision, const void* x, const float* w, const float* bias, int batch, int t, int channels, int k, int stride,
                     float* y, float* sums, void* work, void* stream) {
  return launch_dwconv_raw(precision, x, w, bias, batch, t, channels, k, stride, y, sums, reinterpret_cast<float*>(work),
                           reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_bn_finalize(const float* sums, int channels, float count, float eps, float momentum, float* mean, float* rstd,
                      float* running_mean, float* running_var, void* stream) {
  return launch_bn_finalize(sums, channels, count, eps, momentum, mean, rstd, running_mean, running_var, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_bn_swish_fwd(int precision, const float* y, size_t rows, int channels, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, void* h, void* stream) {
  return launch_bn_swish_fwd(precision, y, rows, channels, mean, rstd, gamma, beta, h, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_bn_swish_bwd_stats(const float* y, const float* dh, size_t rows, int channels, const float* mean, const float* rstd,
                             const float* gamma, const float* beta, float* sums, void* work, void* stream) {
  return launch_bn_swish_bwd_stats(y, dh, rows, channels, mean, rstd, gamma, beta, sums, reinterpret_cast<float*>(work),
                                   reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_bn_swish_bwd_apply(const float* y, const float* dh, size_t rows, int channels, const float* mean, const float* rstd,
                             const float* gamma, const float* beta, const float* sums, float count, float* dy, void* stream) {
  return launch_bn_swish_bwd_apply(y, dh, rows, channels, mean, rstd, gamma, beta, sums, count, dy, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_dwconv_bwd(int precision, const float* dy, const void* x, const float* w, int batch, int t, int channels, int k, int stride,
                     float* dx, float* dw, float* db, void* work, void* stream) {
  return launch_dwconv_bwd(precision, dy, x, w, batch, t, channels, k, stride, dx, dw, db, reinterpret_cast<float*>(work),
                           reinterpret_cast<cudaStream_t>(stream));
}
size_t ec_op_wgrad_work_bytes(int precision, int M, int N, int K) { return wgrad_work_bytes(precision, M, N, K); }
int ec_op_wgrad(int precision, const void* dy, const void* x, int M, int N, int K, float* dw, int accumulate, void* work, void* stream) {
  return launch_wgrad(precision, dy, x, M, N, K, dw, accumulate, nullptr, reinterpret_cast<float*>(work), reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_wgrad_bias(int precision, const void* dy, const void* x, int M, int N, int K, float* dw, int accumulate, float* db, void* work,
                     void* stream) {
  return launch_wgrad(precision, dy, x, M, N, K, dw, accumulate, db, reinterpret_cast<float*>(work), reinterpret_cast<cudaStream_t>(stream));
}
size_t ec_op_colsum_work_bytes(int cols) { return colsum_work_bytes(cols); }
int ec_op_colsum(int precision, const void* m, int is_f32, int rows, int cols, float* out, void* work, void* stream) {
  return launch_colsum(precision, m, is_f32, rows, cols, out, reinterpret_cast<float*>(work), reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_transpose_cast(int precision, const float* src, int rows, int cols, void* dst, void* stream) {
  return launch_transpose_cast(precision, src, rows, cols, dst, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_swish_bwd(int precision, const void* z, const float* dy, size_t n, void* dz, void* stream) {
  return launch_swish_bwd(precision, z, dy, n, dz, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_glu_bwd(int precision, const void* zg, const float* dy, size_t rows, int channels, void* dzg, void* stream) {
  return launch_glu_bwd(precision, zg, dy, rows, channels, dzg, reinterpret_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------- single-operator entry points
int ec_op_cast(int precision, const float* src, void* dst, size_t n, void* stream) {
  return launch_cast_rows(precision, src, dst, n, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_cast_weight(int precision, const float* src, size_t n, void* dst, void* stream) {
  return launch_cast_weight(precision, src, dst, n, n, reinterpret_cast<cudaStream_t>(stream));
}
int ec_weight_planes(int precision) { return static_cast<int>(weight_planes(precision)); }
int ec_op_layernorm(int precision, const float* x, int rows, int dim, const float* gamma, const float* beta, float eps, void* y_act,
                    float* y_f32, void* stream) {
  LayerNormArgs a{};
  a.x = x; a.rows = rows; a.dim = dim; a.gamma = gamma; a.beta = beta; a.eps = eps; a.y_act = y_act; a.y_f32 = y_f32;
  return launch_layernorm(precision, a, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_gemm(int precision, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha, int act,
               const float* residual, float* out_f32, void* out_act, void* stream) {
  GemmArgs g{};
  g.A = A; g.W = W; g.M = M; g.N = N; g.K = K; g.bias = bias; g.alpha = alpha; g.act = act;
  g.residual = residual; g.ld_res = N; g.out_f32 = out_f32; g.ld_out = N; g.out_act = out_act; g.ld_act = N;
  return launch_gemm(precision, g, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_gemm_ex(int precision, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha, int act,
                  const float* residual, float* out_f32, void* out_act, int flags, void* stream) {
  GemmArgs g{};
  g.A = A; g.W = W; g.M = M; g.N = N; g.K = K; g.bias = bias; g.alpha = alpha; g.act = act;
  g.residual = residual; g.ld_res = N; g.out_f32 = out_f32; g.ld_out = N; g.out_act = out_act; g.ld_act = N;
  g.act_f16 = (flags & 1) ? 1 : 0;
  return launch_gemm(precision, g, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_gemm_train(int precision, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha,
                     const float* residual, float* out_f32, void* out_act, const unsigned long long* drop_counter, float drop_p,
                     unsigned drop_site, void* out_act2, unsigned drop_site2, const void* aux_act, unsigned drop_site_aux, void* stream) {
  GemmArgs g{};
  g.A = A; g.W = W; g.M = M; g.N = N; g.K = K; g.bias = bias; g.alpha = alpha; g.act = GEMM_ACT_NONE;
  g.residual = residual; g.ld_res = N; g.out_f32 = out_f32; g.ld_out = N; g.out_act = out_act; g.ld_act = N;
  g.drop_ctr = drop_counter; g.drop_p = drop_p; g.drop_site = drop_site;
  g.out_act2 = out_act2; g.ld_act2 = N; g.drop_site2 = drop_site2;
  g.aux_act = aux_act; g.aux_mode = aux_act != nullptr ? 1 : 0; g.drop_site_aux = drop_site_aux;
  return launch_gemm(precision, g, reinterpret_cast<cudaStream_t>(stream));
}
int ec_attention_operand_kind(int precision, int dim, int heads, int group) {
  return precision == EC_PREC_BF16 ? 1 : (precision == EC_PREC_BF16X2 && attn_operands_f16(dim, heads, group)) ? 2 : 0;
}
int ec_op_gemm_ln(int precision, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha, const float* residual,
                  float* out_f32, int ln_mode, const float* g1, const float* b1, const float* g2, const float* b2, float eps, void* ln_out,
                  void* copy_out, int copy_stride, int frames_per_seq, int frames_out_per_seq, void* stream) {
  GemmArgs g{};
  g.A = A; g.W = W; g.M = M; g.N = N; g.K = K; g.bias = bias; g.alpha = alpha; g.act = GEMM_ACT_NONE;
  g.residual = residual; g.ld_res = N; g.out_f32 = out_f32; g.ld_out = N;
  g.ln_mode = ln_mode; g.ln1_g = g1; g.ln1_b = b1; g.ln2_g = g2; g.ln2_b = b2; g.ln_eps = eps; g.ln_out = ln_out;
  g.copy_out = copy_out; g.copy_stride = copy_stride; g.frames_per_seq = frames_per_seq; g.frames_out_per_seq = frames_out_per_seq;
  return launch_gemm(precision, g, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_gemm_ln_train(int precision, const void* A, const void* W, int M, int N, int K, const float* bias, float alpha, const float* residual,
                        float* out_f32, const float* g1, const float* b1, float eps, void* ln_out, const unsigned long long* drop_counter,
                        float drop_p, unsigned drop_site, void* stream) {
  GemmArgs g{};
  g.A = A; g.W = W; g.M = M; g.N = N; g.K = K; g.bias = bias; g.alpha = alpha; g.act = GEMM_ACT_NONE;
  g.residual = residual; g.ld_res = N; g.out_f32 = out_f32; g.ld_out = N;
  g.ln_mode = 1; g.ln1_g = g1; g.ln1_b = b1; g.ln_eps = eps; g.ln_out = ln_out;
  g.drop_ctr = drop_counter; g.drop_p = drop_p; g.drop_site = drop_site;
  return launch_gemm(precision, g, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_ffn(const void* x_act, const void* w1, const float* b1, const void* w2, const float* b2, int M, int D, int hidden,
              const float* residual, float* out_f32, int ln_mode, const float* g1, const float* be1, const float* g2, const float* be2,
              float eps, void* ln_out, int cluster, void* stream) {
  FfnArgs f{};
  f.x_act = x_act; f.w1 = w1; f.b1 = b1; f.w2 = w2; f.b2 = b2; f.M = M; f.D = D; f.hidden = hidden;
  f.residual = residual; f.out_f32 = out_f32; f.ln_mode = ln_mode; f.ln1_g = g1; f.ln1_b = be1; f.ln2_g = g2; f.ln2_b = be2;
  f.ln_eps = eps; f.ln_out = ln_out; f.cluster = cluster;
  return launch_ffn_fused(f, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_pointwise_glu(int precision, const void* A, const float* w_raw, const float* b_raw, int M, int channels, int K, void* w_scratch,
                        float* b_scratch, void* out_act, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  int nb, tiles; pick_glu(channels, &nb, &tiles);
  EC_TRY(launch_glu_interleave(precision, w_raw, b_raw, channels, K, nb, tiles, w_scratch, b_scratch, st));
  GemmArgs g{};
  g.A = A; g.W = w_scratch; g.M = M; g.N = tiles * 2 * nb; g.K = K; g.bias = b_scratch; g.alpha = 1.f; g.act = GEMM_ACT_NONE;
  g.glu_nb = nb; g.glu_channels = channels; g.out_act = out_act; g.ld_act = channels;
  return launch_gemm(precision, g, st);
}
int ec_op_glu_scratch_rows(int channels) { int nb, tiles; pick_glu(channels, &nb, &tiles); return tiles * 2 * nb; }
int ec_op_fold_bn(const float* w, const float* b, const float* g, const float* beta, const float* rm, const float* rv, float eps, int C,
                  int taps, float* w_out, float* b_out, void* stream) {
  return launch_fold_bn(w, b, g, beta, rm, rv, eps, C, taps, w_out, b_out, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_relpos_attention(int precision, const void* qkv, const void* E, const float* u, const float* v, const int32_t* x_len,
                           int batch, int t, int dim, int heads, int group, void* out, void* stream) {
  AttnArgs a{qkv, E, u, v, x_len, batch, t, dim, heads, group, out, dim, 0,
             (precision == EC_PREC_BF16X2 && attn_operands_f16(dim, heads, group)) ? 1 : 0};
  return launch_relpos_attention(precision, a, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_dwconv_bn_swish(int precision, const void* x, const float* w_folded, const float* b_folded, int batch, int t, int channels,
                          int k, int stride, void* y, void* stream) {
  DwConvArgs a{x, w_folded, b_folded, batch, t, channels, k, stride, y};
  return launch_dwconv_bn_swish(precision, a, reinterpret_cast<cudaStream_t>(stream));
}
int ec_op_subsample_conv(int precision, const float* mel, const float* w_folded, const float* b_folded, int batch, int n_mels, int t,
                         int channels, void* y, void* stream) {
  SubsampleArgs a{mel, w_folded, b_folded, batch, n_mels, t, channels, y};
  return launch_subsample_conv(precision, a, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
