// C ABI of libafr_sm100.so (declared in include/afr_sm100.h): context, workspaces, and the
// sequencing of the kernels that make up one training step / one batched render.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/afr_sm100.h"
#include "afr_gemm.cuh"
#include "afr_internal.h"

using namespace afr;

struct afr_ctx {
  afr_config cfg{};
  int num_sms = 0;                   // SMs of the device (workspace sizing)
  int sms = 0;                       // SMs the persistent kernels may occupy (afr_set_sm_limit)
  int K = 0;  // max_length * hidden : fc_output in_features
  int P = 0;  // sheet_h * sheet_w   : fc_output out_features
  SmallLayout lay{};
  Tensors params{}, grads{}, m{}, v{};
  bool has_params = false, has_grads = false, has_state = false, shadow_valid = false;
  // private device buffers
  __nv_bfloat16* feats = nullptr;    // [max_batch, K]
  // bf16 copy of fc_output.weight [P, K], double buffered (training contexts): the AdamW sweep of
  // step t writes buffer 1 - shadow_cur on its own stream while the dgrad GEMM of step t still
  // reads the buffer the forward used; the buffers swap once a sweep has covered all P rows.
  __nv_bfloat16* wshadow_buf[2] = {nullptr, nullptr};
  int shadow_cur = 0;                // what the next forward reads
  int shadow_fwd = 0;                // what the last training forward read (dgrad reads it too)
  int shadow_rows_swept = 0;         // rows already rewritten in buffer 1 - shadow_cur
  bool shadow_external = false;      // the two copies belong to the caller (afr_bind_shadow)
  __nv_bfloat16* dz = nullptr;       // [max_batch, P]   unscaled (y - t) * mask
  float* dfeat = nullptr;            // [max_batch, K]
  float* logits = nullptr;           // [max_batch, P]   generic path, allocated on first use
  float* loss_partials = nullptr;
  int loss_partials_cap = 0;
  float* bias_scratch = nullptr;     // [32, P]
  float* partials = nullptr;         // [num_sms, lay.total]
  float* fstate = nullptr;           // [max_batch, fsl.stride] front-end records (forward -> backward)
  FrontStateLayout fsl{};
  bool state_valid = false;          // fstate holds the records of the batch in (tokens, B, S)
  // state carried from forward to backward
  const long long* tokens = nullptr;
  long long token_stride = 0;
  int B = 0, S = 0;
  Dropout drop{};
  float grad_scale = 0.f;
  bool fwd_done = false;
  bool frontend_done = false;        // afr_train_frontend ran, afr_train_loss still to come
  bool coresident = false;           // afr_set_coresident
  int smem_reserve = 0;              // afr_set_smem_reserve: bytes per SM the GEMMs leave to a background kernel
  bool cta2 = true;                  // forward / dgrad / wgrad GEMMs run as CTA pairs (AFR_CTA2=0: single CTAs)
  long long launches = 0;
  std::string err;
  // ---- optional font conditioning (config 3): font_embedding [n_fonts, E] + its grad / Adam moments
  float *font_p = nullptr, *font_g = nullptr, *font_m = nullptr, *font_v = nullptr;
  int n_fonts = 0;
  const int* font_ids = nullptr;        // for the next forward call
  const int* font_ids_live = nullptr;   // of the training forward whose backward is pending
  // ---- wide front-end (embed_dim / heads / fc1 width other than 32 / 4 / 64): afr_wide.cu + GEMMs
  bool wide = false;
  int E = kE, H = kHeads, F = kF;
  struct Wide {
    float *e32 = nullptr, *qkv32 = nullptr, *a32 = nullptr, *f32 = nullptr;          // forward
    __nv_bfloat16 *e16 = nullptr, *ctx16 = nullptr, *h16 = nullptr;
    __nv_bfloat16 *win16 = nullptr, *wo16 = nullptr, *w116 = nullptr;                // bf16 weights
    float *xhat = nullptr, *rstd = nullptr, *dr32 = nullptr, *dctx32 = nullptr;      // training
    float2* stat = nullptr;
    uint32_t* abits = nullptr;
    uint8_t* ebits = nullptr;      // embedding-dropout keep bits, one byte per 8 channels
    __nv_bfloat16 *df16 = nullptr, *dr16 = nullptr, *dqkv16 = nullptr;
    float *splitk = nullptr, *ln_part = nullptr, *pos_part = nullptr, *emb_part = nullptr;
    int max_splits = 0, max_ln_parts = 0;
  } w;
};

namespace {

std::string g_create_err;

int fail(afr_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_err = msg;
  return code;
}
int fail_cuda(afr_ctx* c, cudaError_t e, const char* where) {
  return fail(c, AFR_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
#define AFR_CUDA(ctx, expr, where)                          \
  do {                                                      \
    cudaError_t e__ = (expr);                               \
    if (e__ != cudaSuccess) return fail_cuda(ctx, e__, where); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

Tensors to_tensors(const afr_tensors* t) {
  Tensors r{};
  r.pos = t->positional_encoding; r.emb = t->embedding_weight;
  r.win = t->in_proj_weight; r.bin = t->in_proj_bias;
  r.wo = t->out_proj_weight; r.bo = t->out_proj_bias;
  r.lnw = t->layer_norm_weight; r.lnb = t->layer_norm_bias;
  r.w1 = t->fc1_weight; r.b1 = t->fc1_bias;
  r.wout = t->fc_output_weight; r.bout = t->fc_output_bias;
  return r;
}
bool all_set(const afr_tensors* t) {
  const float* const* p = reinterpret_cast<const float* const*>(t);
  for (int i = 0; i < 12; ++i) if (p[i] == nullptr) return false;
  return true;
}
Dropout to_dropout(const afr_dropout* d) {
  Dropout r{};
  if (d == nullptr) return r;
  r.mode = d->mode; r.seed = d->seed; r.step = d->step; r.sample_offset = d->sample_offset;
  r.mask_embed = d->mask_embed; r.mask_attn = d->mask_attn; r.mask_fc1 = d->mask_fc1;
  r.p_embed = d->p_embed; r.p_attn = d->p_attn; r.p_fc1 = d->p_fc1;
  return r;
}

// Tile width: minimise waves * (128 + BN) -- per-tile cost is operand-feed bound (the A tile of
// 128 rows plus the B tile of BN rows per k-block), waves = ceil(tiles / SMs).
int env_int(const char* name) {
  const char* s = std::getenv(name);
  return s ? std::atoi(s) : 0;
}
int choose_bn(int M, int N, int num_sms, const char* env_name, bool cta2 = false) {
  const int forced = env_int(env_name);
  if (forced >= 32 && forced <= 256 && forced % 32 == 0) return forced;
  int best = 256;
  long long best_cost = -1;
  const int units = cta2 ? num_sms / 2 : num_sms;   // CTAs, or CTA pairs working on 256-row tiles
  for (int bn = 256; bn >= 128; bn -= 32) {
    const long long tiles = gemm_num_tiles(M, N, bn, cta2);
    const long long waves = (tiles + units - 1) / units;
    const long long cost = waves * ((cta2 ? 64 : 128) + bn);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  if (N < best) best = ((N + 31) / 32) * 32;
  return best;
}

int ensure_loss_partials(afr_ctx* c, int n) {
  if (n <= c->loss_partials_cap) return AFR_OK;
  if (c->loss_partials) cudaFree(c->loss_partials);
  c->loss_partials = nullptr;
  AFR_CUDA(c, cudaMalloc(&c->loss_partials, sizeof(float) * n), "cudaMalloc(loss_partials)");
  c->loss_partials_cap = n;
  return AFR_OK;
}

int check_batch(afr_ctx* c, int B, int S, const void* tokens) {
  if (!c->has_params) return fail(c, AFR_ERR_STATE, "parameters not bound (afr_bind_params)");
  if (tokens == nullptr) return fail(c, AFR_ERR_INVALID, "tokens is NULL");
  if (B < 1 || B > c->cfg.max_batch)
    return fail(c, AFR_ERR_INVALID, "B outside [1, max_batch]");
  if (S < 1 || S > c->cfg.max_length)
    return fail(c, AFR_ERR_INVALID, "S outside [1, max_length]");
  return AFR_OK;
}

int ensure_shadow(afr_ctx* c, cudaStream_t st) {
  if (c->shadow_rows_swept != 0)
    return fail(c, AFR_ERR_STATE, "a forward pass in the middle of an AdamW sweep over fc_output.weight "
                                  "(afr_adamw_rows must cover every row before the next forward)");
  if (c->shadow_valid) return AFR_OK;
  AFR_CUDA(c, launch_f32_to_bf16(c->params.wout, c->wshadow_buf[c->shadow_cur],
                                 static_cast<long long>(c->P) * c->K, st),
           "f32_to_bf16(fc_output.weight)");
  c->launches += 1;
  c->shadow_valid = true;
  return AFR_OK;
}

int run_frontend_wide(afr_ctx* c, const long long* tokens, long long stride, int B, int S,
                      const Dropout& drop, bool save_state, cudaStream_t st, float* feats_f32);
int run_frontend_backward_wide(afr_ctx* c, const long long* tokens, long long stride, int B, int S,
                               const Dropout& drop, const float* dfeat, cudaStream_t st);

int run_frontend(afr_ctx* c, const long long* tokens, long long stride, int B, int S,
                 const Dropout& drop, bool save_state, cudaStream_t st, float* feats_f32 = nullptr) {
  if (c->wide) {
    if (c->font_ids != nullptr)
      return fail(c, AFR_ERR_INVALID, "font conditioning is built for the reference widths only");
    return run_frontend_wide(c, tokens, stride, B, S, drop, save_state, st, feats_f32);
  }
  float* state = save_state ? c->fstate : nullptr;
  FontCond font{c->font_p, c->font_ids, c->n_fonts};
  if (c->font_ids != nullptr && c->font_p == nullptr)
    return fail(c, AFR_ERR_STATE, "font ids set but no font_embedding bound (afr_bind_font_embedding)");
  AFR_CUDA(c, launch_frontend_forward(c->params, tokens, stride, B, S, c->cfg.max_length,
                                      c->cfg.vocab, drop, c->feats, state, c->sms, st, feats_f32,
                                      save_state && c->smem_reserve > 0, &font),
           "frontend_forward");
  if (save_state) c->font_ids_live = c->font_ids;
  c->launches += 1;
  c->state_valid = state != nullptr;
  return AFR_OK;
}


// ---------------------------------------------------------------- wide front-end sequencing
WideDims wide_dims(const afr_ctx* c, int B, int S) {
  WideDims d{};
  d.B = B; d.S = S; d.L = c->cfg.max_length; d.E = c->E; d.H = c->H; d.dh = c->E / c->H; d.F = c->F;
  d.vocab = c->cfg.vocab;
  return d;
}
WideDrop wide_drop(const Dropout& drop) {
  auto thr = [](double p) { return static_cast<uint32_t>(p * 65536.0 + 0.5); };
  auto inv = [](double p) { return 1.0f / static_cast<float>(1.0 - p); };
  WideDrop w{};
  w.mode = drop.mode;
  w.k0 = static_cast<uint32_t>(drop.seed); w.k1 = static_cast<uint32_t>(drop.seed >> 32);
  w.step = static_cast<uint32_t>(drop.step); w.sample_offset = drop.sample_offset;
  w.thr_e = thr(drop.p_embed); w.thr_a = thr(drop.p_attn); w.thr_f = thr(drop.p_fc1);
  const bool on = drop.mode != 0;
  w.inv_e = on ? inv(drop.p_embed) : 1.f; w.inv_a = on ? inv(drop.p_attn) : 1.f; w.inv_f = on ? inv(drop.p_fc1) : 1.f;
  return w;
}

// D[M,N] (fp32, ld = ldo) = A * B^T (+ bias[n]) on the tcgen05 GEMM; k_splits > 1: partial sums
// into c->w.splitk, summed into `out` afterwards (weight gradients: K = batch x positions)
int wide_gemm(afr_ctx* c, const __nv_bfloat16* A, long long lda, bool a_mn, const __nv_bfloat16* B,
              long long ldb, bool b_mn, int M, int N, int K, float* out, long long ldo, const float* bias,
              int k_splits, cudaStream_t st, const char* what) {
  GemmEpilogue ep{};
  ep.kind = kEpiF32; ep.alpha = 1.f; ep.bias = bias; ep.use_tma_store = 1;
  const bool split = k_splits > 1;
  ep.cta2 = c->cta2 && !split;
  ep.k_splits = split ? k_splits : 0;
  ep.out = split ? c->w.splitk : out;
  ep.ldo = split ? N : ldo;
  const int bn = choose_bn(M, N, c->sms, "AFR_BN_WIDE", ep.cta2 != 0);
  const char* msg = nullptr;
  int splits_used = 1;
  cudaError_t e = launch_gemm_bf16(A, lda, a_mn, B, ldb, b_mn, M, N, K, bn, ep, c->sms, st, nullptr, &msg,
                                   &splits_used);
  if (e != cudaSuccess) return msg ? fail(c, AFR_ERR_INVALID, std::string(what) + ": " + msg) : fail_cuda(c, e, what);
  c->launches += 1;
  if (split) {
    // partial [splits][M rounded up to 128][N] -> out [M][N]
    AFR_CUDA(c, launch_wide_splitk_reduce(c->w.splitk, splits_used, M, N, out, st), "splitk_reduce");
    c->launches += 1;
  }
  return AFR_OK;
}

int run_frontend_wide(afr_ctx* c, const long long* tokens, long long stride, int B, int S,
                      const Dropout& drop, bool save_state, cudaStream_t st, float* feats_f32) {
  if (drop.mode == 2)
    return fail(c, AFR_ERR_INVALID, "injected dropout masks (mode 2) are not built for the wide front-end");
  auto& w = c->w;
  const WideDims d = wide_dims(c, B, S);
  const WideDrop dr = wide_drop(drop);
  const int E = c->E, F = c->F, R = B * S;
  AFR_CUDA(c, ensure_err_flag_public(), "err flag");
  AFR_CUDA(c, launch_wide_split_weight(c->params.win, 3 * E, E, w.win16, st), "split bf16(in_proj)");
  AFR_CUDA(c, launch_wide_split_weight(c->params.wo, E, E, w.wo16, st), "split bf16(out_proj)");
  AFR_CUDA(c, launch_wide_split_weight(c->params.w1, F, E, w.w116, st), "split bf16(fc1)");
  AFR_CUDA(c, launch_wide_embed(d, dr, tokens, stride, c->params.emb, c->params.pos, w.e32, w.e16,
                                save_state ? w.ebits : nullptr, frontend_error_flag(), c->sms, st), "wide_embed");
  c->launches += 4;
  int rc;
  // forward GEMMs: split-bf16 operands, K = 3E (x_hi w_hi + x_lo w_hi + x_hi w_lo)
  if ((rc = wide_gemm(c, w.e16, 3 * E, false, w.win16, 3 * E, false, R, 3 * E, 3 * E, w.qkv32, 3 * E,
                      c->params.bin, 1, st, "gemm(in_proj)"))) return rc;
  AFR_CUDA(c, launch_wide_attention_fwd(d, dr, w.qkv32, w.ctx16, save_state ? w.stat : nullptr,
                                        save_state ? w.abits : nullptr, st), "wide_attention_fwd");
  if ((rc = wide_gemm(c, w.ctx16, 3 * E, false, w.wo16, 3 * E, false, R, E, 3 * E, w.a32, E, c->params.bo, 1, st,
                      "gemm(out_proj)"))) return rc;
  AFR_CUDA(c, launch_wide_ln_fwd(R, E, w.e32, w.a32, c->params.lnw, c->params.lnb, save_state ? w.xhat : nullptr,
                                 save_state ? w.rstd : nullptr, w.h16, c->sms, st), "wide_ln_fwd");
  if ((rc = wide_gemm(c, w.h16, 3 * E, false, w.w116, 3 * E, false, R, F, 3 * E, w.f32, F, c->params.b1, 1, st,
                      "gemm(fc1)"))) return rc;
  AFR_CUDA(c, launch_wide_act_fwd(d, dr, w.f32, c->feats, feats_f32, c->sms, st), "wide_act_fwd");
  c->launches += 3;
  c->state_valid = save_state;
  return AFR_OK;
}

int run_frontend_backward_wide(afr_ctx* c, const long long* tokens, long long stride, int B, int S,
                               const Dropout& drop, const float* dfeat, cudaStream_t st) {
  auto& w = c->w;
  const WideDims d = wide_dims(c, B, S);
  const WideDrop dr = wide_drop(drop);
  const int E = c->E, F = c->F, R = B * S;
  int splits = R / 64 / 8;                      // >= 8 k-blocks of 64 rows per piece
  if (splits > w.max_splits) splits = w.max_splits;
  if (splits < 1) splits = 1;
  int rc;
  // fc1 + ReLU + dropout1
  AFR_CUDA(c, launch_wide_act_bwd(d, dr, w.f32, dfeat, w.df16, c->sms, st), "wide_act_bwd");
  AFR_CUDA(c, launch_wide_colsum_bf16(w.df16, R, F, w.ln_part, w.max_ln_parts, c->grads.b1, c->sms, st),
           "bias_grad(fc1)");
  c->launches += 3;
  // (the hi parts of the split operands: the first E columns of every 3E-wide row)
  if ((rc = wide_gemm(c, w.df16, F, false, w.w116, 3 * E, true, R, E, F, w.a32, E, nullptr, 1, st,
                      "gemm(d fc1 input)"))) return rc;
  if ((rc = wide_gemm(c, w.df16, F, true, w.h16, 3 * E, true, F, E, R, c->grads.w1, E, nullptr, splits, st,
                      "gemm(d fc1.weight)"))) return rc;
  // LayerNorm + residual
  AFR_CUDA(c, launch_wide_ln_bwd(R, E, w.a32, w.xhat, w.rstd, c->params.lnw, w.dr32, w.dr16, w.ln_part,
                                 w.max_ln_parts, c->grads.lnw, c->grads.lnb, c->sms, st), "wide_ln_bwd");
  AFR_CUDA(c, launch_wide_colsum_bf16(w.dr16, R, E, w.ln_part, w.max_ln_parts, c->grads.bo, c->sms, st),
           "bias_grad(out_proj)");
  c->launches += 4;
  // attention out-projection
  if ((rc = wide_gemm(c, w.dr16, E, false, w.wo16, 3 * E, true, R, E, E, w.dctx32, E, nullptr, 1, st,
                      "gemm(d ctx)"))) return rc;
  if ((rc = wide_gemm(c, w.dr16, E, true, w.ctx16, 3 * E, true, E, E, R, c->grads.wo, E, nullptr, splits, st,
                      "gemm(d out_proj.weight)"))) return rc;
  AFR_CUDA(c, launch_wide_attention_bwd(d, dr, w.qkv32, w.dctx32, w.ctx16, w.stat, w.abits, w.dqkv16, st),
           "wide_attention_bwd");
  AFR_CUDA(c, launch_wide_colsum_bf16(w.dqkv16, R, 3 * E, w.ln_part, w.max_ln_parts, c->grads.bin, c->sms, st),
           "bias_grad(in_proj)");
  c->launches += 3;
  // in-projection
  if ((rc = wide_gemm(c, w.dqkv16, 3 * E, false, w.win16, 3 * E, true, R, E, 3 * E, w.a32, E, nullptr, 1, st,
                      "gemm(d e)"))) return rc;
  if ((rc = wide_gemm(c, w.dqkv16, 3 * E, true, w.e16, 3 * E, true, 3 * E, E, R, c->grads.win, E, nullptr, splits, st,
                      "gemm(d in_proj.weight)"))) return rc;
  // embedding + positions
  AFR_CUDA(c, launch_wide_embed_bwd(d, dr, tokens, stride, w.dr32, w.a32, w.ebits, w.pos_part, w.emb_part,
                                    2 * c->num_sms, c->grads.pos, c->grads.emb, c->sms, st), "wide_embed_bwd");
  c->launches += 3;
  return AFR_OK;
}

AdamHyper make_hyper(double lr, double b1, double b2, double eps, double wd, long long step) {
  AdamHyper h{};
  const double bc1 = 1.0 - std::pow(b1, static_cast<double>(step));
  const double bc2 = 1.0 - std::pow(b2, static_cast<double>(step));
  h.decay = static_cast<float>(1.0 - lr * wd);
  h.beta1_w = static_cast<float>(1.0 - b1);
  h.beta2 = static_cast<float>(b2);
  h.one_m_beta2 = static_cast<float>(1.0 - b2);
  h.bc2_sqrt = static_cast<float>(std::sqrt(bc2));
  h.eps = static_cast<float>(eps);
  h.neg_step = static_cast<float>(-(lr / bc1));
  return h;
}

}  // namespace

extern "C" {

int afr_abi_version(void) { return AFR_ABI_VERSION; }

const char* afr_last_error(const afr_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

int afr_create(const afr_config* cfg, afr_ctx** out) {
  if (cfg == nullptr || out == nullptr) return fail(nullptr, AFR_ERR_INVALID, "null argument");
  *out = nullptr;
  const bool wide = cfg->embed_dim != kE || cfg->num_heads != kHeads || cfg->hidden != kF;
  if (wide) {
    const char* why = nullptr;
    if (!wide_shape_supported(cfg->embed_dim, cfg->num_heads, cfg->hidden, cfg->max_length, &why))
      return fail(nullptr, AFR_ERR_INVALID, std::string("unsupported front-end shape: ") + why);
  }
  if (cfg->max_length < 1 || cfg->max_length > kMaxL || cfg->vocab < 1 || cfg->max_batch < 1)
    return fail(nullptr, AFR_ERR_INVALID, "max_length must be in [1,128], vocab >= 1, max_batch >= 1");
  const long long P = static_cast<long long>(cfg->sheet_h) * cfg->sheet_w;
  if (cfg->sheet_h < 1 || cfg->sheet_w < 1 || (P % 32) != 0)
    return fail(nullptr, AFR_ERR_INVALID, "sheet_h*sheet_w must be a positive multiple of 32");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev)
    return fail(nullptr, AFR_ERR_CUDA, "no such CUDA device");
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess)
    return fail(nullptr, AFR_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(nullptr, AFR_ERR_UNSUPPORTED,
                "libafr_sm100 needs an sm_100 (Blackwell B200) device; there is no fallback path");
  DeviceGuard guard(cfg->device);
  afr_ctx* c = new afr_ctx();
  c->cfg = *cfg;
  c->num_sms = prop.multiProcessorCount;
  c->sms = c->num_sms;
  c->K = cfg->max_length * cfg->hidden;
  c->P = static_cast<int>(P);
  c->wide = wide;
  c->E = cfg->embed_dim; c->H = cfg->num_heads; c->F = cfg->hidden;
  { const char* e2 = std::getenv("AFR_CTA2"); c->cta2 = !(e2 && std::atoi(e2) == 0); }
  c->lay.init(cfg->max_length, cfg->vocab);
  c->fsl.init(cfg->max_length);
  if (cfg->training && !wide &&
      frontend_backward_smem_bytes(cfg->max_length, cfg->vocab) > prop.sharedMemPerBlockOptin) {
    delete c;
    return fail(nullptr, AFR_ERR_INVALID,
                "max_length too large for the training path (one sample's backward must fit in "
                "the 227 KB of shared memory: max_length <= 120)");
  }
  const size_t Bm = static_cast<size_t>(cfg->max_batch);
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  alloc(reinterpret_cast<void**>(&c->feats), Bm * c->K * 2);
  alloc(reinterpret_cast<void**>(&c->wshadow_buf[0]), static_cast<size_t>(c->P) * c->K * 2);
  if (cfg->training)
    alloc(reinterpret_cast<void**>(&c->wshadow_buf[1]), static_cast<size_t>(c->P) * c->K * 2);
  if (cfg->training) {
    alloc(reinterpret_cast<void**>(&c->dz), Bm * c->P * 2);
    alloc(reinterpret_cast<void**>(&c->dfeat), Bm * c->K * 4);
    alloc(reinterpret_cast<void**>(&c->bias_scratch), static_cast<size_t>(32) * c->P * 4);
    if (static_cast<long long>(cfg->vocab) * kE * c->num_sms * 4 > (1ll << 31)) {
      afr_destroy(c);
      return fail(nullptr, AFR_ERR_INVALID, "training path supports vocab up to ~100k rows");
    }
    if (!wide) {
      alloc(reinterpret_cast<void**>(&c->partials),
            static_cast<size_t>(c->num_sms) * c->lay.total * 4);
      alloc(reinterpret_cast<void**>(&c->fstate), Bm * c->fsl.stride * 4);
    }
  }
  if (wide) {
    const size_t R = Bm * cfg->max_length, E = c->E, F = c->F, H = c->H;
    auto& w = c->w;
    alloc(reinterpret_cast<void**>(&w.e32), R * E * 4);
    alloc(reinterpret_cast<void**>(&w.e16), R * 3 * E * 2);      // split bf16 rows [hi | lo | hi], see afr_wide.cu
    alloc(reinterpret_cast<void**>(&w.qkv32), R * 3 * E * 4);
    alloc(reinterpret_cast<void**>(&w.ctx16), R * 3 * E * 2);
    alloc(reinterpret_cast<void**>(&w.a32), R * E * 4);
    alloc(reinterpret_cast<void**>(&w.h16), R * 3 * E * 2);      // [hi | lo | hi], see afr_wide.cu
    alloc(reinterpret_cast<void**>(&w.f32), R * F * 4);
    alloc(reinterpret_cast<void**>(&w.win16), 3 * E * 3 * E * 2);   // split bf16 rows [hi | hi | lo]
    alloc(reinterpret_cast<void**>(&w.wo16), E * 3 * E * 2);
    alloc(reinterpret_cast<void**>(&w.w116), F * 3 * E * 2);     // [hi | hi | lo]
    if (cfg->training) {
      alloc(reinterpret_cast<void**>(&w.xhat), R * E * 4);
      alloc(reinterpret_cast<void**>(&w.rstd), R * 4);
      alloc(reinterpret_cast<void**>(&w.stat), Bm * H * cfg->max_length * sizeof(float2));
      alloc(reinterpret_cast<void**>(&w.abits), Bm * H * cfg->max_length * 4 * sizeof(uint32_t));
      alloc(reinterpret_cast<void**>(&w.ebits), R * E / 8);
      alloc(reinterpret_cast<void**>(&w.df16), R * F * 2);
      alloc(reinterpret_cast<void**>(&w.dr32), R * E * 4);
      alloc(reinterpret_cast<void**>(&w.dr16), R * E * 2);
      alloc(reinterpret_cast<void**>(&w.dctx32), R * E * 4);
      alloc(reinterpret_cast<void**>(&w.dqkv16), R * 3 * E * 2);
      w.max_splits = c->num_sms;
      const size_t max_mn = ((3 * E + 127) / 128 * 128) * E > ((F + 127) / 128 * 128) * E
                                ? ((3 * E + 127) / 128 * 128) * E : ((F + 127) / 128 * 128) * E;
      alloc(reinterpret_cast<void**>(&w.splitk), static_cast<size_t>(w.max_splits) * max_mn * 4);
      w.max_ln_parts = c->num_sms * 4;
      alloc(reinterpret_cast<void**>(&w.ln_part),
            static_cast<size_t>(w.max_ln_parts) * (3 * E > F ? 3 * E : F) * 4);   // LayerNorm / bias-gradient partial rows
      alloc(reinterpret_cast<void**>(&w.pos_part), static_cast<size_t>(2 * c->num_sms) * cfg->max_length * E * 4);
      if (static_cast<size_t>(cfg->vocab) * E * 4 <= 128 * 1024)
        alloc(reinterpret_cast<void**>(&w.emb_part), static_cast<size_t>(2 * c->num_sms) * cfg->vocab * E * 4);
    }
  }
  if (e != cudaSuccess) {
    std::string msg = std::string("cudaMalloc(workspace): ") + cudaGetErrorString(e);
    afr_destroy(c);
    return fail(nullptr, AFR_ERR_CUDA, msg);
  }
  *out = c;
  return AFR_OK;
}

int afr_destroy(afr_ctx* c) {
  if (c == nullptr) return AFR_OK;
  DeviceGuard guard(c->cfg.device);
  cudaFree(c->feats);
  if (!c->shadow_external) { cudaFree(c->wshadow_buf[0]); cudaFree(c->wshadow_buf[1]); }
  cudaFree(c->dz); cudaFree(c->dfeat);
  cudaFree(c->logits); cudaFree(c->loss_partials); cudaFree(c->bias_scratch);
  cudaFree(c->partials); cudaFree(c->fstate);
  {
    auto& w = c->w;
    void* ptrs[] = {w.e32, w.qkv32, w.a32, w.f32, w.e16, w.ctx16, w.h16, w.win16, w.wo16, w.w116, w.xhat, w.rstd,
                    w.dr32, w.dctx32, w.stat, w.abits, w.ebits, w.df16, w.dr16, w.dqkv16, w.splitk, w.ln_part, w.pos_part,
                    w.emb_part};
    for (void* q : ptrs) cudaFree(q);
  }
  delete c;
  return AFR_OK;
}

int afr_bind_params(afr_ctx* c, const afr_tensors* t) {
  if (!c || !t || !all_set(t)) return fail(c, AFR_ERR_INVALID, "afr_bind_params: null tensor");
  c->params = to_tensors(t);
  c->has_params = true;
  c->shadow_valid = false;
  return AFR_OK;
}
int afr_bind_grads(afr_ctx* c, const afr_tensors* t) {
  if (!c || !t || !all_set(t)) return fail(c, AFR_ERR_INVALID, "afr_bind_grads: null tensor");
  c->grads = to_tensors(t);
  c->has_grads = true;
  return AFR_OK;
}
int afr_bind_adam_state(afr_ctx* c, const afr_tensors* m, const afr_tensors* v) {
  if (!c || !m || !v || !all_set(m) || !all_set(v))
    return fail(c, AFR_ERR_INVALID, "afr_bind_adam_state: null tensor");
  c->m = to_tensors(m);
  c->v = to_tensors(v);
  c->has_state = true;
  return AFR_OK;
}

int afr_sync_shadow(afr_ctx* c, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->has_params) return fail(c, AFR_ERR_STATE, "parameters not bound");
  DeviceGuard guard(c->cfg.device);
  c->shadow_valid = false;
  return ensure_shadow(c, static_cast<cudaStream_t>(stream));
}

int afr_bind_shadow(afr_ctx* c, void* copy0, void* copy1) {
  if (!c || !copy0 || !copy1 || copy0 == copy1)
    return fail(c, AFR_ERR_INVALID, "afr_bind_shadow: need two distinct buffers");
  if ((reinterpret_cast<uintptr_t>(copy0) | reinterpret_cast<uintptr_t>(copy1)) & 15)
    return fail(c, AFR_ERR_INVALID, "afr_bind_shadow: buffers must be 16-byte aligned");
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  if (c->shadow_rows_swept != 0)
    return fail(c, AFR_ERR_STATE, "afr_bind_shadow in the middle of an AdamW sweep");
  DeviceGuard guard(c->cfg.device);
  __nv_bfloat16* fresh[2] = {static_cast<__nv_bfloat16*>(copy0), static_cast<__nv_bfloat16*>(copy1)};
  // A training step may be in flight (forward done, dgrad still to come): both copies move over
  // with their contents and roles, synchronously.
  AFR_CUDA(c, cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  const size_t bytes = static_cast<size_t>(c->P) * c->K * 2;
  for (int i = 0; i < 2; ++i)
    if (c->wshadow_buf[i] != nullptr && c->wshadow_buf[i] != fresh[i])
      AFR_CUDA(c, cudaMemcpy(fresh[i], c->wshadow_buf[i], bytes, cudaMemcpyDeviceToDevice),
               "cudaMemcpy(shadow)");
  if (!c->shadow_external) { cudaFree(c->wshadow_buf[0]); cudaFree(c->wshadow_buf[1]); }
  c->wshadow_buf[0] = fresh[0];
  c->wshadow_buf[1] = fresh[1];
  c->shadow_external = true;
  return AFR_OK;
}

int afr_bind_font_embedding(afr_ctx* c, int n_fonts, float* table, float* grad, float* exp_avg,
                            float* exp_avg_sq) {
  if (!c) return AFR_ERR_INVALID;
  if (n_fonts == 0 && table == nullptr) {     // unbind
    c->font_p = c->font_g = c->font_m = c->font_v = nullptr;
    c->n_fonts = 0; c->font_ids = c->font_ids_live = nullptr;
    return AFR_OK;
  }
  if (c->wide) return fail(c, AFR_ERR_INVALID, "font conditioning is built for the reference widths only");
  if (n_fonts < 1 || n_fonts > kMaxFonts || table == nullptr)
    return fail(c, AFR_ERR_INVALID, "afr_bind_font_embedding: 1..16 fonts and a non-null table");
  if ((exp_avg == nullptr) != (exp_avg_sq == nullptr) || (exp_avg != nullptr && grad == nullptr))
    return fail(c, AFR_ERR_INVALID, "afr_bind_font_embedding: Adam moments come in pairs and need a gradient");
  c->font_p = table; c->font_g = grad; c->font_m = exp_avg; c->font_v = exp_avg_sq; c->n_fonts = n_fonts;
  return AFR_OK;
}

int afr_set_font_ids(afr_ctx* c, const int32_t* font_ids) {
  if (!c) return AFR_ERR_INVALID;
  if (font_ids != nullptr && c->font_p == nullptr)
    return fail(c, AFR_ERR_STATE, "afr_set_font_ids before afr_bind_font_embedding");
  c->font_ids = font_ids;
  return AFR_OK;
}

int afr_set_sm_limit(afr_ctx* c, int sms) {
  if (!c) return AFR_ERR_INVALID;
  if (sms < 1 || sms > c->num_sms) sms = c->num_sms;
  c->sms = sms;
  return AFR_OK;
}

int afr_set_coresident(afr_ctx* c, int on) {
  if (!c) return AFR_ERR_INVALID;
  c->coresident = on != 0;
  return AFR_OK;
}

int afr_set_smem_reserve(afr_ctx* c, int bytes) {
  if (!c) return AFR_ERR_INVALID;
  if (bytes < 0 || bytes > 96 * 1024) return fail(c, AFR_ERR_INVALID, "afr_set_smem_reserve: 0..96 KB");
  c->smem_reserve = bytes;
  return AFR_OK;
}

int afr_shadow_index(const afr_ctx* c) { return c ? c->shadow_cur : AFR_ERR_INVALID; }

int afr_shadow_commit(afr_ctx* c) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  c->shadow_cur ^= 1;
  c->shadow_rows_swept = 0;
  c->shadow_valid = true;
  return AFR_OK;
}

int afr_forward_eval(afr_ctx* c, const int64_t* tokens, int64_t token_stride, int B, int S,
                     void* out, int out_kind, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  int rc = check_batch(c, B, S, tokens);
  if (rc) return rc;
  if (out == nullptr) return fail(c, AFR_ERR_INVALID, "out is NULL");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((rc = ensure_shadow(c, st))) return rc;
  Dropout off{};
  if ((rc = run_frontend(c, reinterpret_cast<const long long*>(tokens), token_stride, B, S, off,
                         false, st)))
    return rc;
  GemmEpilogue ep{};
  ep.out = out; ep.ldo = c->P; ep.bias = c->params.bout; ep.alpha = 1.f;
  if (out_kind == AFR_OUT_SHEET_F32) { ep.kind = kEpiF32; ep.clamp01 = 1; ep.use_tma_store = 1; }
  else if (out_kind == AFR_OUT_LOGITS_F32) { ep.kind = kEpiF32; ep.clamp01 = 0; ep.use_tma_store = 1; }
  else if (out_kind == AFR_OUT_SHEET_U8) { ep.kind = kEpiU8; }
  else return fail(c, AFR_ERR_INVALID, "unknown out_kind");
  if (env_int("AFR_NO_TMA_STORE")) ep.use_tma_store = 0;
  const char* msg = nullptr;
  ep.cta2 = c->cta2;
  const int bn = choose_bn(B, c->P, c->sms, "AFR_BN_FWD", c->cta2);
  cudaError_t e = launch_gemm_bf16(c->feats, c->K, false, c->wshadow_buf[c->shadow_cur], c->K, false, B, c->P, c->K, bn,
                                   ep, c->sms, st, nullptr, &msg);
  if (e != cudaSuccess) return msg ? fail(c, AFR_ERR_INVALID, msg) : fail_cuda(c, e, "gemm(forward)");
  c->launches += 1;
  return AFR_OK;
}

int afr_train_frontend(afr_ctx* c, const int64_t* tokens, int64_t token_stride, int B, int S,
                       const afr_dropout* dropout, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  int rc = check_batch(c, B, S, tokens);
  if (rc) return rc;
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  c->drop = to_dropout(dropout);
  c->tokens = reinterpret_cast<const long long*>(tokens);
  c->token_stride = token_stride;
  c->B = B; c->S = S;
  c->fwd_done = false;
  c->frontend_done = false;
  if ((rc = run_frontend(c, c->tokens, token_stride, B, S, c->drop, true, st))) return rc;
  c->frontend_done = true;
  return AFR_OK;
}

int afr_train_loss(afr_ctx* c, const void* targets, int target_kind, double loss_count,
                   float* loss_out, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->frontend_done) return fail(c, AFR_ERR_STATE, "afr_train_loss before afr_train_frontend");
  if (targets == nullptr || loss_out == nullptr || !(loss_count > 0))
    return fail(c, AFR_ERR_INVALID, "targets/loss_out NULL or loss_count <= 0");
  if (target_kind != AFR_TARGET_U8 && target_kind != AFR_TARGET_F32)
    return fail(c, AFR_ERR_INVALID, "unknown target_kind");
  if (reinterpret_cast<uintptr_t>(targets) & 15)
    return fail(c, AFR_ERR_INVALID, "targets must be 16-byte aligned");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_shadow(c, st);
  if (rc) return rc;
  c->shadow_fwd = c->shadow_cur;
  c->grad_scale = static_cast<float>(2.0 / loss_count);  // d/dy of mean((y-t)^2)
  const int B = c->B;
  const int bn = choose_bn(B, c->P, c->sms, "AFR_BN_FWD", c->cta2);
  const int tiles = gemm_num_tiles(B, c->P, bn, c->cta2) * (c->cta2 ? 2 : 1);   // 128-row blocks
  if ((rc = ensure_loss_partials(c, tiles * 4))) return rc;
  GemmEpilogue ep{};
  ep.cta2 = c->cta2;
  ep.kind = kEpiLoss; ep.out = c->dz; ep.ldo = c->P; ep.bias = c->params.bout; ep.alpha = 1.f;
  ep.target = targets; ep.target_is_f32 = target_kind == AFR_TARGET_F32;
  ep.loss_partials = c->loss_partials;
  const char* msg = nullptr;
  cudaError_t e = launch_gemm_bf16(c->feats, c->K, false, c->wshadow_buf[c->shadow_cur], c->K, false,
                                   B, c->P, c->K, bn, ep, c->sms, st, nullptr, &msg);
  if (e != cudaSuccess) return msg ? fail(c, AFR_ERR_INVALID, msg) : fail_cuda(c, e, "gemm(forward+loss)");
  AFR_CUDA(c, launch_loss_finalize(c->loss_partials, tiles * 4, loss_count, loss_out, st),
           "loss_finalize");
  c->launches += 2;
  c->fwd_done = true;
  c->frontend_done = false;
  return AFR_OK;
}

int afr_train_forward_loss(afr_ctx* c, const int64_t* tokens, int64_t token_stride, int B, int S,
                           const void* targets, int target_kind, const afr_dropout* dropout,
                           double loss_count, float* loss_out, void* stream) {
  int rc = afr_train_frontend(c, tokens, token_stride, B, S, dropout, stream);
  if (rc) return rc;
  return afr_train_loss(c, targets, target_kind, loss_count, loss_out, stream);
}

int afr_train_wgrad(afr_ctx* c, int row_begin, int row_end, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->fwd_done) return fail(c, AFR_ERR_STATE, "afr_train_wgrad before a training forward");
  if (!c->has_grads) return fail(c, AFR_ERR_STATE, "gradients not bound (afr_bind_grads)");
  if (row_begin < 0 || row_end > c->P || row_begin >= row_end || (row_begin % 32) != 0 ||
      ((row_end - row_begin) % 32) != 0)
    return fail(c, AFR_ERR_INVALID, "row range must be 32-aligned inside [0, H*W]");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = row_end - row_begin;
  // dW[rows, K] = scale * dZ[:, rows]^T feats : A = dZ (MN-major, K = batch), B = feats (MN-major)
  GemmEpilogue ep{};
  ep.kind = kEpiF32;
  ep.out = c->grads.wout + static_cast<long long>(row_begin) * c->K;
  ep.ldo = c->K; ep.alpha = c->grad_scale; ep.use_tma_store = env_int("AFR_NO_TMA_STORE") ? 0 : 1;
  const char* msg = nullptr;
  ep.cta2 = c->cta2;
  const int bn = choose_bn(rows, c->K, c->sms, "AFR_BN_WGRAD", c->cta2);
  cudaError_t e = launch_gemm_bf16(c->dz + row_begin, c->P, true, c->feats, c->K, true, rows, c->K,
                                   c->B, bn, ep, c->sms, st, nullptr, &msg);
  if (e != cudaSuccess) return msg ? fail(c, AFR_ERR_INVALID, msg) : fail_cuda(c, e, "gemm(wgrad)");
  // db[rows] = scale * sum_b dZ[b, rows]
  AFR_CUDA(c, launch_bias_grad(c->dz + row_begin, c->B, rows, c->grad_scale, c->bias_scratch,
                               c->grads.bout + row_begin, st, c->P),
           "bias_grad");
  c->launches += 3;
  return AFR_OK;
}

int afr_train_dgrad_gemm(afr_ctx* c, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->fwd_done) return fail(c, AFR_ERR_STATE, "afr_train_dgrad_gemm before a training forward");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // dfeat[B, K] = scale * dZ[B, P] W[P, K] : A = dZ (K-major over pixels), B = W (MN-major)
  GemmEpilogue ep{};
  ep.kind = kEpiF32; ep.out = c->dfeat; ep.ldo = c->K; ep.alpha = c->grad_scale;
  ep.use_tma_store = env_int("AFR_NO_TMA_STORE") ? 0 : 1;
  ep.compact = c->coresident ? 1 : 0;
  const char* msg = nullptr;
  ep.cta2 = c->cta2; ep.smem_reserve = c->smem_reserve;
  int bn = choose_bn(c->B, c->K, c->sms, "AFR_BN_DGRAD", c->cta2 && !c->coresident);
  if (c->coresident && bn > 128) bn = 128;
  cudaError_t e = launch_gemm_bf16(c->dz, c->P, false, c->wshadow_buf[c->shadow_fwd], c->K, true, c->B, c->K, c->P, bn,
                                   ep, c->sms, st, nullptr, &msg);
  if (e != cudaSuccess) return msg ? fail(c, AFR_ERR_INVALID, msg) : fail_cuda(c, e, "gemm(dgrad)");
  c->launches += 1;
  return AFR_OK;
}

int afr_train_frontend_backward(afr_ctx* c, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->fwd_done) return fail(c, AFR_ERR_STATE, "afr_train_frontend_backward before a training forward");
  if (!c->has_grads) return fail(c, AFR_ERR_STATE, "gradients not bound (afr_bind_grads)");
  if (!c->state_valid)
    return fail(c, AFR_ERR_STATE, "front-end records of this batch were overwritten by another forward");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->wide)
    return run_frontend_backward_wide(c, c->tokens, c->token_stride, c->B, c->S, c->drop, c->dfeat, st);
  int grid = 0;
  FontCond font{c->font_p, c->font_ids_live, c->n_fonts};
  AFR_CUDA(c, launch_frontend_backward(c->params, c->tokens, c->token_stride, c->B, c->S,
                                       c->cfg.max_length, c->cfg.vocab, c->drop, c->dfeat,
                                       c->fstate, c->partials, c->sms, &grid, c->sms, st,
                                       c->smem_reserve > 0, &font),
           "frontend_backward");
  AFR_CUDA(c, launch_small_grad_reduce(c->partials, grid, c->lay, c->grads, st,
                                       c->font_ids_live ? c->font_g : nullptr, c->n_fonts),
           "small_grad_reduce");
  c->launches += 2;
  return AFR_OK;
}

int afr_train_dgrad(afr_ctx* c, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->has_grads) return fail(c, AFR_ERR_STATE, "gradients not bound (afr_bind_grads)");
  int rc = afr_train_dgrad_gemm(c, stream);
  if (rc) return rc;
  return afr_train_frontend_backward(c, stream);
}

int afr_train_step(afr_ctx* c, const int64_t* tokens, int64_t token_stride, int B, int S,
                   const void* targets, int target_kind, const afr_dropout* dropout,
                   double loss_count, float* loss_out, void* stream) {
  int rc = afr_train_forward_loss(c, tokens, token_stride, B, S, targets, target_kind, dropout,
                                  loss_count, loss_out, stream);
  if (rc) return rc;
  if ((rc = afr_train_wgrad(c, 0, c->P, stream))) return rc;
  return afr_train_dgrad(c, stream);
}

int afr_forward_train(afr_ctx* c, const int64_t* tokens, int64_t token_stride, int B, int S,
                      const afr_dropout* dropout, float* sheet_out, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  int rc = check_batch(c, B, S, tokens);
  if (rc) return rc;
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  if (sheet_out == nullptr) return fail(c, AFR_ERR_INVALID, "sheet_out is NULL");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->logits == nullptr)
    AFR_CUDA(c, cudaMalloc(&c->logits, static_cast<size_t>(c->cfg.max_batch) * c->P * 4),
             "cudaMalloc(logits)");
  if ((rc = ensure_shadow(c, st))) return rc;
  c->shadow_fwd = c->shadow_cur;
  c->drop = to_dropout(dropout);
  c->tokens = reinterpret_cast<const long long*>(tokens);
  c->token_stride = token_stride;
  c->B = B; c->S = S;
  c->grad_scale = 1.f;
  if ((rc = run_frontend(c, c->tokens, token_stride, B, S, c->drop, true, st))) return rc;
  GemmEpilogue ep{};
  ep.kind = kEpiF32; ep.out = c->logits; ep.ldo = c->P; ep.bias = c->params.bout; ep.alpha = 1.f;
  ep.use_tma_store = env_int("AFR_NO_TMA_STORE") ? 0 : 1;
  const char* msg = nullptr;
  ep.cta2 = c->cta2;
  const int bn = choose_bn(B, c->P, c->sms, "AFR_BN_FWD", c->cta2);
  cudaError_t e = launch_gemm_bf16(c->feats, c->K, false, c->wshadow_buf[c->shadow_cur], c->K, false, B, c->P, c->K, bn,
                                   ep, c->sms, st, nullptr, &msg);
  if (e != cudaSuccess) return msg ? fail(c, AFR_ERR_INVALID, msg) : fail_cuda(c, e, "gemm(forward)");
  AFR_CUDA(c, launch_clamp01(c->logits, sheet_out, static_cast<long long>(B) * c->P, st), "clamp01");
  c->launches += 2;
  c->fwd_done = false;  // becomes true once afr_backward has produced dZ
  return AFR_OK;
}

int afr_backward(afr_ctx* c, const float* dsheet, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (c->logits == nullptr || c->tokens == nullptr)
    return fail(c, AFR_ERR_STATE, "afr_backward before afr_forward_train");
  if (dsheet == nullptr) return fail(c, AFR_ERR_INVALID, "dsheet is NULL");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AFR_CUDA(c, launch_clamp_backward(dsheet, c->logits, c->dz, static_cast<long long>(c->B) * c->P, st),
           "clamp_backward");
  c->launches += 1;
  c->fwd_done = true;
  int rc = afr_train_wgrad(c, 0, c->P, stream);
  if (rc) return rc;
  return afr_train_dgrad(c, stream);
}

int afr_adamw_rows(afr_ctx* c, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int64_t step, int row_begin, int row_end, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->has_params || !c->has_grads || !c->has_state)
    return fail(c, AFR_ERR_STATE, "params / grads / adam state not bound");
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  if (step < 1 || row_begin < 0 || row_end > c->P || row_begin >= row_end)
    return fail(c, AFR_ERR_INVALID, "bad step or row range");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AdamHyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step);
  const long long off = static_cast<long long>(row_begin) * c->K;
  const long long n = static_cast<long long>(row_end - row_begin) * c->K;
  AFR_CUDA(c, launch_adamw(c->params.wout + off, c->grads.wout + off, c->m.wout + off,
                           c->v.wout + off, n, h, c->wshadow_buf[1 - c->shadow_cur] + off,
                           c->num_sms, st),
           "adamw(fc_output.weight)");
  c->launches += 1;
  c->shadow_rows_swept += row_end - row_begin;
  if (c->shadow_rows_swept >= c->P) {   // sweep complete: the next forward reads the new weights
    c->shadow_cur ^= 1;
    c->shadow_rows_swept = 0;
    c->shadow_valid = true;
  }
  return AFR_OK;
}

int afr_adamw_rows_bg(afr_ctx* c, double lr, double beta1, double beta2, double eps,
                      double weight_decay, int64_t step, int row_begin, int row_end,
                      const float* grad_rows, int ctas, int stages, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->has_params || !c->has_state || (grad_rows == nullptr && !c->has_grads))
    return fail(c, AFR_ERR_STATE, "params / grads / adam state not bound");
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  if (step < 1 || row_begin < 0 || row_end > c->P || row_begin >= row_end)
    return fail(c, AFR_ERR_INVALID, "bad step or row range");
  if ((static_cast<long long>(row_end - row_begin) * c->K) % 4 != 0 ||
      (static_cast<long long>(row_begin) * c->K) % 4 != 0)
    return fail(c, AFR_ERR_INVALID, "afr_adamw_rows_bg: row range must cover whole 16-byte groups");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AdamHyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step);
  const long long off = static_cast<long long>(row_begin) * c->K;
  const long long n = static_cast<long long>(row_end - row_begin) * c->K;
  if (ctas < 1) ctas = c->num_sms;
  if (stages < 1) stages = 4;
  const float* g = grad_rows != nullptr ? grad_rows : c->grads.wout + off;
  AFR_CUDA(c, launch_adamw_ring(c->params.wout + off, g, c->m.wout + off, c->v.wout + off, n, h,
                                c->wshadow_buf[1 - c->shadow_cur] + off, ctas, stages, st),
           "adamw_ring(fc_output.weight rows)");
  c->launches += 1;
  c->shadow_rows_swept += row_end - row_begin;
  if (c->shadow_rows_swept >= c->P) {   // sweep complete: the next forward reads the new weights
    c->shadow_cur ^= 1;
    c->shadow_rows_swept = 0;
    c->shadow_valid = true;
  }
  return AFR_OK;
}

static int wgrad_to_impl(afr_ctx* c, int row_begin, int row_end, void* grad_rows, bool bf16_out, int with_bias,
                         void* stream);

int afr_train_wgrad_to(afr_ctx* c, int row_begin, int row_end, float* grad_rows, int with_bias,
                       void* stream) {
  return wgrad_to_impl(c, row_begin, row_end, grad_rows, false, with_bias, stream);
}

int afr_train_wgrad_to_bf16(afr_ctx* c, int row_begin, int row_end, void* grad_rows_bf16, int with_bias,
                            void* stream) {
  return wgrad_to_impl(c, row_begin, row_end, grad_rows_bf16, true, with_bias, stream);
}

static int wgrad_to_impl(afr_ctx* c, int row_begin, int row_end, void* grad_rows, bool bf16_out, int with_bias,
                         void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->fwd_done) return fail(c, AFR_ERR_STATE, "afr_train_wgrad_to before a training forward");
  if (grad_rows == nullptr) return fail(c, AFR_ERR_INVALID, "afr_train_wgrad_to: grad_rows is NULL");
  if (with_bias && !c->has_grads) return fail(c, AFR_ERR_STATE, "gradients not bound (afr_bind_grads)");
  if (row_begin < 0 || row_end > c->P || row_begin >= row_end || (row_begin % 32) != 0 ||
      ((row_end - row_begin) % 32) != 0)
    return fail(c, AFR_ERR_INVALID, "row range must be 32-aligned inside [0, H*W]");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = row_end - row_begin;
  GemmEpilogue ep{};
  ep.kind = kEpiF32;
  ep.out = grad_rows;
  ep.out_bf16 = bf16_out ? 1 : 0;
  ep.ldo = c->K; ep.alpha = c->grad_scale; ep.use_tma_store = env_int("AFR_NO_TMA_STORE") ? 0 : 1;
  const char* msg = nullptr;
  ep.cta2 = c->cta2; ep.smem_reserve = c->smem_reserve;
  const int bn = choose_bn(rows, c->K, c->sms, "AFR_BN_WGRAD", c->cta2);
  cudaError_t e = launch_gemm_bf16(c->dz + row_begin, c->P, true, c->feats, c->K, true, rows, c->K,
                                   c->B, bn, ep, c->sms, st, nullptr, &msg);
  if (e != cudaSuccess) return msg ? fail(c, AFR_ERR_INVALID, msg) : fail_cuda(c, e, "gemm(wgrad chunk)");
  c->launches += 1;
  if (with_bias) {
    AFR_CUDA(c, launch_bias_grad(c->dz + row_begin, c->B, rows, c->grad_scale, c->bias_scratch,
                                 c->grads.bout + row_begin, st, c->P),
             "bias_grad");
    c->launches += 2;
  }
  return AFR_OK;
}

int afr_train_wgrad_adamw(afr_ctx* c, double lr, double beta1, double beta2, double eps,
                          double weight_decay, int64_t step, int row_begin, int row_end,
                          void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->fwd_done) return fail(c, AFR_ERR_STATE, "afr_train_wgrad_adamw before a training forward");
  if (!c->has_params || !c->has_grads || !c->has_state)
    return fail(c, AFR_ERR_STATE, "params / grads / adam state not bound");
  if (step < 1 || row_begin < 0 || row_end > c->P || row_begin >= row_end || (row_begin % 32) != 0 ||
      ((row_end - row_begin) % 32) != 0)
    return fail(c, AFR_ERR_INVALID, "bad step, or row range not 32-aligned inside [0, H*W]");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = row_end - row_begin;
  const long long off = static_cast<long long>(row_begin) * c->K;
  GemmEpilogue ep{};
  ep.kind = kEpiAdamW;
  ep.out = c->wshadow_buf[1 - c->shadow_cur] + off;   // the copy the next forward will read
  ep.ldo = c->K; ep.alpha = c->grad_scale;
  ep.adam_p = c->params.wout + off; ep.adam_m = c->m.wout + off; ep.adam_v = c->v.wout + off;
  ep.hyper = make_hyper(lr, beta1, beta2, eps, weight_decay, step);
  const char* msg = nullptr;
  int bn = env_int("AFR_WA_BN");            // tuning knobs (tools/fused_sweep.py)
  if (bn < 32 || bn > 256 || (bn % 32) != 0) bn = 256;
  ep.compact = c->coresident ? 1 : 0;
  // the AdamW GEMM is HBM-bound and keeps single CTAs (8 epilogue warps + 2 x 48 KB ring measured
  // faster than pairs with a 4 x 32 KB ring: 0.73 vs 0.79 ms); AFR_WA_CTA2=1 to compare
  ep.cta2 = c->cta2 && env_int("AFR_WA_CTA2") != 0; ep.smem_reserve = c->smem_reserve;
  if (c->coresident && bn > 128) bn = 128;
  ep.adam_sets = env_int("AFR_WA_SETS");
  ep.adam_sub = env_int("AFR_WA_SUB");
  ep.adam_stages = env_int("AFR_WA_STAGES");
  if (c->K < bn) bn = c->K;
  cudaError_t e = launch_gemm_bf16(c->dz + row_begin, c->P, true, c->feats, c->K, true, rows, c->K,
                                   c->B, bn, ep, c->sms, st, nullptr, &msg);
  if (e != cudaSuccess) return msg ? fail(c, AFR_ERR_INVALID, msg) : fail_cuda(c, e, "gemm(wgrad+adamw)");
  c->launches += 1;
  c->shadow_rows_swept += rows;
  if (c->shadow_rows_swept >= c->P) {   // every row rewritten: the next forward reads the new weights
    c->shadow_cur ^= 1;
    c->shadow_rows_swept = 0;
    c->shadow_valid = true;
  }
  return AFR_OK;
}

int afr_train_bgrad(afr_ctx* c, int row_begin, int row_end, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->fwd_done) return fail(c, AFR_ERR_STATE, "afr_train_bgrad before a training forward");
  if (!c->has_grads) return fail(c, AFR_ERR_STATE, "gradients not bound (afr_bind_grads)");
  if (row_begin < 0 || row_end > c->P || row_begin >= row_end || (row_begin % 2) != 0)
    return fail(c, AFR_ERR_INVALID, "bad row range");
  DeviceGuard guard(c->cfg.device);
  AFR_CUDA(c, launch_bias_grad(c->dz + row_begin, c->B, row_end - row_begin, c->grad_scale,
                               c->bias_scratch, c->grads.bout + row_begin,
                               static_cast<cudaStream_t>(stream), c->P),
           "bias_grad");
  c->launches += 2;
  return AFR_OK;
}

int afr_adamw_rows_gather(afr_ctx* c, double lr, double beta1, double beta2, double eps,
                          double weight_decay, int64_t step, int row_begin, int row_end,
                          const void* const* peer_grads, void* const* peer_shadows, int world,
                          int ctas, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->has_params || !c->has_state)
    return fail(c, AFR_ERR_STATE, "params / adam state not bound");
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  if (step < 1 || row_begin < 0 || row_end > c->P || row_begin >= row_end)
    return fail(c, AFR_ERR_INVALID, "bad step or row range");
  if (peer_grads == nullptr || peer_shadows == nullptr || world < 1 || world > 8 || ctas < 1)
    return fail(c, AFR_ERR_INVALID, "afr_adamw_rows_gather: need 1..8 peers and ctas >= 1");
  const long long off = static_cast<long long>(row_begin) * c->K;
  const long long n = static_cast<long long>(row_end - row_begin) * c->K;
  const float* g[8];
  __nv_bfloat16* sh[8];
  for (int q = 0; q < world; ++q) {
    if (peer_grads[q] == nullptr || peer_shadows[q] == nullptr)
      return fail(c, AFR_ERR_INVALID, "afr_adamw_rows_gather: null peer pointer");
    g[q] = static_cast<const float*>(peer_grads[q]) + off;
    sh[q] = static_cast<__nv_bfloat16*>(peer_shadows[q]) + off;
  }
  DeviceGuard guard(c->cfg.device);
  const AdamHyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step);
  AFR_CUDA(c, launch_adamw_gather(c->params.wout + off, c->m.wout + off, c->v.wout + off, n, h, g, sh,
                                  world, ctas, static_cast<cudaStream_t>(stream)),
           "adamw_gather(fc_output.weight rows)");
  c->launches += 1;
  c->shadow_rows_swept += row_end - row_begin;   // completed by the peers' stores; afr_shadow_commit
  return AFR_OK;
}

int afr_adamw_rows_gather_nvls(afr_ctx* c, double lr, double beta1, double beta2, double eps,
                               double weight_decay, int64_t step, int row_begin, int row_end,
                               const void* grad_multicast, void* shadow_multicast, int ctas,
                               void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->has_params || !c->has_state)
    return fail(c, AFR_ERR_STATE, "params / adam state not bound");
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  if (step < 1 || row_begin < 0 || row_end > c->P || row_begin >= row_end)
    return fail(c, AFR_ERR_INVALID, "bad step or row range");
  if (grad_multicast == nullptr || shadow_multicast == nullptr || ctas < 1)
    return fail(c, AFR_ERR_INVALID, "afr_adamw_rows_gather_nvls: null multicast pointer or ctas < 1");
  const long long off = static_cast<long long>(row_begin) * c->K;
  const long long n = static_cast<long long>(row_end - row_begin) * c->K;
  DeviceGuard guard(c->cfg.device);
  const AdamHyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step);
  AFR_CUDA(c, launch_adamw_gather_nvls(c->params.wout + off, c->m.wout + off, c->v.wout + off, n, h,
                                       static_cast<const float*>(grad_multicast) + off,
                                       static_cast<__nv_bfloat16*>(shadow_multicast) + off, ctas,
                                       static_cast<cudaStream_t>(stream)),
           "adamw_gather_nvls(fc_output.weight rows)");
  c->launches += 1;
  c->shadow_rows_swept += row_end - row_begin;   // completed by the multicast stores; afr_shadow_commit
  return AFR_OK;
}


int afr_adamw_rows_gather_bf16(afr_ctx* c, double lr, double beta1, double beta2, double eps,
                               double weight_decay, int64_t step, int row_begin, int row_end,
                               const void* const* peer_grads, void* const* peer_shadows, int world,
                               int ctas, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->has_params || !c->has_state) return fail(c, AFR_ERR_STATE, "params / adam state not bound");
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  if (step < 1 || row_begin < 0 || row_end > c->P || row_begin >= row_end)
    return fail(c, AFR_ERR_INVALID, "bad step or row range");
  if (peer_grads == nullptr || peer_shadows == nullptr || world < 1 || world > 8 || ctas < 1)
    return fail(c, AFR_ERR_INVALID, "afr_adamw_rows_gather_bf16: need 1..8 peers and ctas >= 1");
  const long long off = static_cast<long long>(row_begin) * c->K;
  const long long n = static_cast<long long>(row_end - row_begin) * c->K;
  if ((n % 8) != 0 || (off % 8) != 0) return fail(c, AFR_ERR_INVALID, "row range must cover whole 16-byte bf16 groups");
  const __nv_bfloat16* g[8];
  __nv_bfloat16* sh[8];
  for (int q = 0; q < world; ++q) {
    if (peer_grads[q] == nullptr || peer_shadows[q] == nullptr)
      return fail(c, AFR_ERR_INVALID, "afr_adamw_rows_gather_bf16: null peer pointer");
    g[q] = static_cast<const __nv_bfloat16*>(peer_grads[q]) + off;
    sh[q] = static_cast<__nv_bfloat16*>(peer_shadows[q]) + off;
  }
  DeviceGuard guard(c->cfg.device);
  const AdamHyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step);
  AFR_CUDA(c, launch_adamw_gather_bf16(c->params.wout + off, c->m.wout + off, c->v.wout + off, n, h, g, sh,
                                       world, ctas, static_cast<cudaStream_t>(stream)),
           "adamw_gather_bf16(fc_output.weight rows)");
  c->launches += 1;
  c->shadow_rows_swept += row_end - row_begin;
  return AFR_OK;
}

int afr_adamw_rows_gather_nvls_bf16(afr_ctx* c, double lr, double beta1, double beta2, double eps,
                                    double weight_decay, int64_t step, int row_begin, int row_end,
                                    const void* grad_multicast, void* shadow_multicast, int ctas,
                                    void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->has_params || !c->has_state) return fail(c, AFR_ERR_STATE, "params / adam state not bound");
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  if (step < 1 || row_begin < 0 || row_end > c->P || row_begin >= row_end)
    return fail(c, AFR_ERR_INVALID, "bad step or row range");
  if (grad_multicast == nullptr || shadow_multicast == nullptr || ctas < 1)
    return fail(c, AFR_ERR_INVALID, "afr_adamw_rows_gather_nvls_bf16: null multicast pointer or ctas < 1");
  const long long off = static_cast<long long>(row_begin) * c->K;
  const long long n = static_cast<long long>(row_end - row_begin) * c->K;
  if ((n % 8) != 0 || (off % 8) != 0) return fail(c, AFR_ERR_INVALID, "row range must cover whole 16-byte bf16 groups");
  DeviceGuard guard(c->cfg.device);
  const AdamHyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step);
  AFR_CUDA(c, launch_adamw_gather_nvls_bf16(c->params.wout + off, c->m.wout + off, c->v.wout + off, n, h,
                                            static_cast<const __nv_bfloat16*>(grad_multicast) + off,
                                            static_cast<__nv_bfloat16*>(shadow_multicast) + off, ctas,
                                            static_cast<cudaStream_t>(stream)),
           "adamw_gather_nvls_bf16(fc_output.weight rows)");
  c->launches += 1;
  c->shadow_rows_swept += row_end - row_begin;
  return AFR_OK;
}

int afr_adamw_small(afr_ctx* c, double lr, double beta1, double beta2, double eps,
                    double weight_decay, int64_t step, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  if (!c->has_params || !c->has_grads || !c->has_state)
    return fail(c, AFR_ERR_STATE, "params / grads / adam state not bound");
  if (step < 1) return fail(c, AFR_ERR_INVALID, "step must be >= 1");
  DeviceGuard guard(c->cfg.device);
  const AdamHyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step);
  const int L = c->cfg.max_length, V = c->cfg.vocab, E = c->E, F = c->F;
  const int sizes[10] = {L * E, V * E, 3 * E * E, 3 * E, E * E, E, E, E, F * E, F};
  float* const* pp = reinterpret_cast<float* const*>(&c->params);
  float* const* gg = reinterpret_cast<float* const*>(&c->grads);
  float* const* mm = reinterpret_cast<float* const*>(&c->m);
  float* const* vv = reinterpret_cast<float* const*>(&c->v);
  SmallAdamJob jobs[12];
  for (int i = 0; i < 10; ++i) jobs[i] = SmallAdamJob{pp[i], gg[i], mm[i], vv[i], sizes[i]};
  jobs[10] = SmallAdamJob{c->params.bout, c->grads.bout, c->m.bout, c->v.bout, c->P};
  int njobs = 11;
  if (c->font_p != nullptr && c->font_m != nullptr)
    jobs[njobs++] = SmallAdamJob{c->font_p, c->font_g, c->font_m, c->font_v, c->n_fonts * E};
  AFR_CUDA(c, launch_adamw_small(jobs, njobs, h, static_cast<cudaStream_t>(stream)), "adamw(small)");
  c->launches += 1;
  return AFR_OK;
}

int afr_adamw_step(afr_ctx* c, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int64_t step, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  int rc = afr_adamw_rows(c, lr, beta1, beta2, eps, weight_decay, step, 0, c->P, stream);
  if (rc) return rc;
  return afr_adamw_small(c, lr, beta1, beta2, eps, weight_decay, step, stream);
}

int afr_check_tokens(afr_ctx* c, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  DeviceGuard guard(c->cfg.device);
  int* flag = frontend_error_flag();
  if (flag == nullptr) return AFR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int host = 0;
  AFR_CUDA(c, cudaMemcpyAsync(&host, flag, sizeof(int), cudaMemcpyDeviceToHost, st), "read token flag");
  AFR_CUDA(c, cudaStreamSynchronize(st), "cudaStreamSynchronize");
  if (host != 0) {
    AFR_CUDA(c, cudaMemsetAsync(flag, 0, sizeof(int), st), "reset token flag");
    return fail(c, AFR_ERR_TOKEN_RANGE, "token id outside [0, vocab) (index out of range in embedding)");
  }
  return AFR_OK;
}

int afr_workspace_ptr(afr_ctx* c, int which, void** ptr, size_t* bytes) {
  if (!c || !ptr || !bytes) return AFR_ERR_INVALID;
  const size_t Bm = static_cast<size_t>(c->cfg.max_batch);
  switch (which) {
    case 0: *ptr = c->feats; *bytes = Bm * c->K * 2; break;
    case 1: *ptr = c->dz; *bytes = c->dz ? Bm * c->P * 2 : 0; break;
    case 2: *ptr = c->dfeat; *bytes = c->dfeat ? Bm * c->K * 4 : 0; break;
    case 3: *ptr = c->wshadow_buf[c->shadow_cur]; *bytes = static_cast<size_t>(c->P) * c->K * 2; break;
    case 4: *ptr = c->logits; *bytes = c->logits ? Bm * c->P * 4 : 0; break;
    default: return fail(c, AFR_ERR_INVALID, "unknown workspace id");
  }
  return AFR_OK;
}

int afr_workspace_copy(afr_ctx* c, int which, void* dst, size_t bytes, void* stream) {
  void* src = nullptr;
  size_t avail = 0;
  int rc = afr_workspace_ptr(c, which, &src, &avail);
  if (rc) return rc;
  if (dst == nullptr || src == nullptr || bytes > avail)
    return fail(c, AFR_ERR_INVALID, "afr_workspace_copy: null pointer or size beyond the workspace");
  DeviceGuard guard(c->cfg.device);
  AFR_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice,
                              static_cast<cudaStream_t>(stream)),
           "cudaMemcpyAsync(workspace)");
  return AFR_OK;
}

int afr_gemm_tiles(afr_ctx* c, int B, int* out) {
  if (!c || !out || B < 1) return AFR_ERR_INVALID;
  out[0] = choose_bn(B, c->P, c->sms, "AFR_BN_FWD", c->cta2);
  out[1] = choose_bn(B, c->K, c->sms, "AFR_BN_DGRAD", c->cta2);
  out[2] = choose_bn(c->P, c->K, c->sms, "AFR_BN_WGRAD", c->cta2);
  return AFR_OK;
}

int64_t afr_launch_count(const afr_ctx* c) { return c ? c->launches : 0; }

int afr_debug_phase_cycles(unsigned long long* out, int reset) {
  if (out == nullptr) return fail(nullptr, AFR_ERR_INVALID, "afr_debug_phase_cycles: out is NULL");
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = read_phase_cycles(out, reset);
  if (e == cudaErrorNotSupported)
    return fail(nullptr, AFR_ERR_UNSUPPORTED, "library was not built with -DAFR_PHASE_TIMING");
  if (e != cudaSuccess) return fail_cuda(nullptr, e, "afr_debug_phase_cycles");
  return AFR_OK;
}

int afr_debug_frontend_forward(afr_ctx* c, const int64_t* tokens, int64_t token_stride, int B, int S,
                               const afr_dropout* dropout, float* feats_f32, void* stream) {
  if (!c) return AFR_ERR_INVALID;
  int rc = check_batch(c, B, S, tokens);
  if (rc) return rc;
  if (feats_f32 == nullptr) return fail(c, AFR_ERR_INVALID, "feats_f32 is NULL");
  DeviceGuard guard(c->cfg.device);
  c->tokens = reinterpret_cast<const long long*>(tokens);
  c->token_stride = token_stride;
  c->B = B; c->S = S;
  return run_frontend(c, c->tokens, token_stride, B, S, to_dropout(dropout), c->cfg.training != 0,
                      static_cast<cudaStream_t>(stream), feats_f32);
}

int afr_debug_frontend_backward(afr_ctx* c, const int64_t* tokens, int64_t token_stride, int B,
                                int S, const afr_dropout* dropout, const float* dfeat,
                                void* stream) {
  if (!c) return AFR_ERR_INVALID;
  int rc = check_batch(c, B, S, tokens);
  if (rc) return rc;
  if (!c->cfg.training) return fail(c, AFR_ERR_STATE, "context created with training = 0");
  if (!c->has_grads) return fail(c, AFR_ERR_STATE, "gradients not bound (afr_bind_grads)");
  if (dfeat == nullptr) return fail(c, AFR_ERR_INVALID, "dfeat is NULL");
  if (!c->state_valid || c->tokens != reinterpret_cast<const long long*>(tokens) || c->B != B || c->S != S)
    return fail(c, AFR_ERR_STATE,
                "afr_debug_frontend_backward needs afr_debug_frontend_forward on the same batch first");
  DeviceGuard guard(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->wide)
    return run_frontend_backward_wide(c, reinterpret_cast<const long long*>(tokens), token_stride, B, S,
                                      to_dropout(dropout), dfeat, st);
  int grid = 0;
  FontCond font{c->font_p, c->font_ids_live, c->n_fonts};
  AFR_CUDA(c, launch_frontend_backward(c->params, reinterpret_cast<const long long*>(tokens),
                                       token_stride, B, S, c->cfg.max_length, c->cfg.vocab,
                                       to_dropout(dropout), dfeat, c->fstate, c->partials,
                                       c->sms, &grid, c->sms, st, false, &font),
           "frontend_backward(debug)");
  AFR_CUDA(c, launch_small_grad_reduce(c->partials, grid, c->lay, c->grads, st,
                                       c->font_ids_live ? c->font_g : nullptr, c->n_fonts),
           "small_grad_reduce(debug)");
  c->launches += 2;
  return AFR_OK;
}

int afr_debug_div_sqrt(const float* a, const float* b, float* q, float* s, float* q_ieee,
                       float* s_ieee, int64_t n, void* stream) {
  if (!a || !b || !q || !s || !q_ieee || !s_ieee || n < 1)
    return fail(nullptr, AFR_ERR_INVALID, "afr_debug_div_sqrt: null pointer or n < 1");
  cudaError_t e = launch_div_sqrt_check(a, b, q, s, q_ieee, s_ieee, n, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(nullptr, e, "afr_debug_div_sqrt");
  return AFR_OK;
}

int afr_gemm_bf16(int device, const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb,
                  int b_mn, float* D, int64_t ldd, int M, int N, int K, int tile_n, float alpha,
                  int use_tma_store, void* stream) {
  if (!A || !B || !D) return fail(nullptr, AFR_ERR_INVALID, "afr_gemm_bf16: null pointer");
  int major = 0, sms = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
    return fail(nullptr, AFR_ERR_CUDA, "cudaDeviceGetAttribute failed");
  if (major != 10) return fail(nullptr, AFR_ERR_UNSUPPORTED, "needs an sm_100 device");
  DeviceGuard guard(device);
  GemmEpilogue ep{};
  ep.kind = kEpiF32; ep.out = D; ep.ldo = ldd; ep.alpha = alpha;
  ep.use_tma_store = use_tma_store & 1;
  ep.cta2 = (use_tma_store & 2) ? 1 : 0;
  const char* msg = nullptr;
  cudaError_t e = launch_gemm_bf16(static_cast<const __nv_bfloat16*>(A), lda, a_mn != 0,
                                   static_cast<const __nv_bfloat16*>(B), ldb, b_mn != 0, M, N, K,
                                   tile_n, ep, sms,
                                   static_cast<cudaStream_t>(stream), nullptr, &msg);
  if (e != cudaSuccess)
    return msg ? fail(nullptr, AFR_ERR_INVALID, msg) : fail_cuda(nullptr, e, "afr_gemm_bf16");
  return AFR_OK;
}

}  // extern "C"
