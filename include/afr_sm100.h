/* libafr_sm100.so -- C ABI of the B200-native hot path of chenglou/ai-font-renderer.
 *
 * The reference has no FFI of its own: its hot path is the body of the per-batch loop in
 * model.py:291-311 (zero_grad -> model(x) -> mse_loss -> backward -> AdamW.step) and the
 * no-grad forward used by validation (model.py:317-330) and render_strings (helpers.py:62-64),
 * all expressed as PyTorch eager ops. Each entry point below names the reference statements it
 * replaces. The binding a maintainer would add on the reference side (a ctypes stub inside
 * model.py / helpers.py) is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C: opaque context, raw device pointers, sizes, a cudaStream_t passed as void*.
 *  - every function returns 0 on success or a negative afr_status; afr_last_error() gives text.
 *    Nothing throws, aborts or synchronises the host (except afr_check_tokens, by contract).
 *  - the caller (PyTorch on the host side) owns parameters, gradients, optimizer state, inputs and
 *    outputs; the library owns only its context: TMA descriptors, the bf16 shadow of
 *    fc_output.weight, and the activations that live between forward and backward.
 *  - one context per (device, model shape); not re-entrant per context.
 *  - sm_100a only. There is no CPU path and no fallback: on any other device afr_create fails.
 */
#ifndef AFR_SM100_H_
#define AFR_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFR_ABI_VERSION 1

typedef struct afr_ctx afr_ctx;

typedef enum afr_status {
  AFR_OK = 0,
  AFR_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  AFR_ERR_CUDA = -2,        /* CUDA runtime or driver error (text in afr_last_error) */
  AFR_ERR_UNSUPPORTED = -3, /* not an sm_100 device */
  AFR_ERR_STATE = -4,       /* call order violated (e.g. params not bound) */
  AFR_ERR_TOKEN_RANGE = -5  /* a token id outside [0, vocab): reference raises IndexError, model.py:167 */
} afr_status;

/* Model shape. Defaults of the reference are in the comments (model.py:64-66,79-81,136,148). */
typedef struct afr_config {
  int device;      /* CUDA ordinal */
  int vocab;       /* nn.Embedding rows: 128 */
  int max_length;  /* MAX_CHARS_PER_SHEET: 100 (<= 128) */
  int embed_dim;   /* EMBEDDING_DIM: 32; other multiples of 32 up to 256 take the GEMM-based front-end */
  int num_heads;   /* NUM_ATTENTION_HEADS: 4; embed_dim / num_heads must be 8, 16 or 32 */
  int hidden;      /* fc1 width: 64; other multiples of 32 with the GEMM-based front-end */
  int sheet_h;     /* SHEET_HEIGHT: 80 */
  int sheet_w;     /* SHEET_WIDTH: 240; sheet_h*sheet_w must be a multiple of 32 */
  int max_batch;   /* largest B any call will pass; sizes the private workspaces */
  int training;    /* nonzero: also allocate the backward workspaces */
} afr_config;

/* The 12 fp32 tensors of the reference state_dict, in its order (helpers.py:76-79 saves them,
 * helpers.py:101 loads them). Row-major [out, in] like torch. Device pointers. */
typedef struct afr_tensors {
  float* positional_encoding; /* [max_length, 32]        model.py:140 */
  float* embedding_weight;    /* [vocab, 32]             model.py:136 */
  float* in_proj_weight;      /* [96, 32]                model.py:144 */
  float* in_proj_bias;        /* [96] */
  float* out_proj_weight;     /* [32, 32] */
  float* out_proj_bias;       /* [32] */
  float* layer_norm_weight;   /* [32]                    model.py:145 */
  float* layer_norm_bias;     /* [32] */
  float* fc1_weight;          /* [64, 32]                model.py:148 */
  float* fc1_bias;            /* [64] */
  float* fc_output_weight;    /* [H*W, 64*max_length]    model.py:152 */
  float* fc_output_bias;      /* [H*W] */
} afr_tensors;

/* Dropout at the three sites of forward (model.py:137/168, 144 -> torch functional.py attention
 * dropout, 149/184). mode 0 = off (model.eval()); mode 1 = built-in counter-based generator
 * (Philox4x32-10 keyed by seed, counter = element block / site / global sample index / step, so
 * masks do not depend on how the batch is sharded over GPUs and backward regenerates them);
 * mode 2 = caller-supplied keep-masks (1 = keep), used to reproduce a recorded torch run. */
typedef struct afr_dropout {
  int mode;
  uint64_t seed;
  uint64_t step;
  int64_t sample_offset;
  const uint8_t* mask_embed; /* [B, S, 32] */
  const uint8_t* mask_attn;  /* [B, 4, S, S] */
  const uint8_t* mask_fc1;   /* [B, S, 64] */
  double p_embed;            /* 0.2  (DROPOUT_RATE, model.py:80) */
  double p_attn;             /* 0.2 */
  double p_fc1;              /* 0.25 (model.py:149) */
} afr_dropout;

typedef enum afr_out_kind {
  AFR_OUT_SHEET_F32 = 0, /* clamp(z,0,1) as fp32 [B,H,W]: what forward() returns, model.py:199-204 */
  AFR_OUT_SHEET_U8 = 1,  /* (uint8)(clamp(z,0,1)*255), truncation: helpers.py:33 */
  AFR_OUT_LOGITS_F32 = 2 /* z before the clamp (parity tests) */
} afr_out_kind;

typedef enum afr_target_kind {
  AFR_TARGET_U8 = 0, /* 8-bit grey, compared as u8/255.0f exactly like helpers.py:121 */
  AFR_TARGET_F32 = 1 /* fp32 in [0,1], the TensorDataset layout of helpers.py:177-181 */
} afr_target_kind;

int afr_abi_version(void);

/* Lifetime. */
int afr_create(const afr_config* cfg, afr_ctx** out);
int afr_destroy(afr_ctx* ctx);
const char* afr_last_error(const afr_ctx* ctx); /* ctx may be NULL: error of the last afr_create */

/* Binding of caller-owned tensors. Replaces nothing in the reference: it is where
 * model.parameters() / p.grad / optimizer.state (model.py:273) become raw pointers. */
int afr_bind_params(afr_ctx* ctx, const afr_tensors* params);
int afr_bind_grads(afr_ctx* ctx, const afr_tensors* grads);
int afr_bind_adam_state(afr_ctx* ctx, const afr_tensors* exp_avg, const afr_tensors* exp_avg_sq);
/* Rebuild the private bf16 copy of fc_output.weight from the fp32 master; call after the caller
 * wrote the master itself (load_state_dict at helpers.py:101, a torch optimizer, init). */
int afr_sync_shadow(afr_ctx* ctx, void* stream);
/* Data-parallel runs with a row-sharded optimizer (training.py): the two bf16 copies of
 * fc_output.weight ([H*W, 64*max_length] each, 16-byte aligned) live in caller-owned memory, so
 * the caller's collective can all-gather the rows other ranks updated straight into the inactive
 * copy. afr_shadow_index returns which copy (0/1) the next forward reads; the AdamW sweep
 * (afr_adamw_rows) always writes the other one. A sweep that covered every row activates the
 * written copy by itself; a sharded sweep (own rows only) is completed by the caller's
 * all-gather and then activated with afr_shadow_commit. */
int afr_bind_shadow(afr_ctx* ctx, void* copy0, void* copy1);
/* Multi-font conditioning (BASELINE.json configs[2]; an EXTENSION: the reference has one font,
 * generate_font.ts:65-67 ships Montserrat unused): a thirteenth parameter font_embedding
 * [n_fonts <= 16, embed_dim] whose row font_ids[b] is added to every token embedding of sample b
 * before the embedding dropout (model.py:167-168 with `+ font_embedding(font)` in between).
 * afr_bind_font_embedding binds the table and, for training, its gradient (overwritten by
 * afr_train_frontend_backward) and, with the gradient, its two Adam moments (afr_adamw_small then
 * steps the table); (0, NULL, ...) unbinds. afr_set_font_ids names the device array int32 [B] that the
 * NEXT forward call (eval or training; the training backward reuses it) reads; NULL = no
 * conditioning. Reference widths only. */
int afr_bind_font_embedding(afr_ctx* ctx, int n_fonts, float* table, float* grad, float* exp_avg,
                            float* exp_avg_sq);
int afr_set_font_ids(afr_ctx* ctx, const int32_t* font_ids);
/* The GEMMs and the front-end kernels are persistent (one CTA per SM, static tile order): a CTA
 * that cannot become resident because a collective's CTAs hold its SM delays the whole kernel.
 * A data-parallel caller therefore leaves the collective its SMs: the persistent kernels launch
 * at most `sms` CTAs (values outside [1, #SMs] restore the default, all SMs). */
int afr_set_sm_limit(afr_ctx* ctx, int sms);
/* Room for a background kernel (afr_adamw_rows_bg: stages x 8 KB + 1 KB of shared memory, 4 warps
 * of 40 registers) beside the kernels that follow the wgrad GEMM in a training step: with
 * bytes > 0 (0..96 KB) the dgrad GEMM (and afr_train_wgrad_to) launch with an operand ring that
 * leaves `bytes` of the SM's shared memory free, and the front-end backward / training front-end
 * forward launch as register-capped variants (112 / 64 registers) so that a warp of the
 * background kernel fits on every scheduler partition. 0 restores the full footprints. */
int afr_set_smem_reserve(afr_ctx* ctx, int bytes);
int afr_shadow_index(const afr_ctx* ctx);
int afr_shadow_commit(afr_ctx* ctx);

/* Eval forward: model.py:158-204 under model.eval(), for validation (model.py:317-330) and
 * render_strings (helpers.py:62-64), batched. tokens: int64 [B, token_stride], first S columns
 * used, S <= max_length (positions >= S give zero features, model.py:190-193). */
int afr_forward_eval(afr_ctx* ctx, const int64_t* tokens, int64_t token_stride, int B, int S,
                     void* out, int out_kind, void* stream);

/* Fused training forward: model.py:299 + 304-306. Runs the forward, the clamp, the MSE partial
 * sums and emits d(loss)/d(logits) internally. *loss_out (device fp32) = sum over the local batch
 * of (y - t)^2 / loss_count; with loss_count = global_B * H * W the per-rank values add up to the
 * reference's mean loss. targets: [B, H*W] of target_kind. */
int afr_train_forward_loss(afr_ctx* ctx, const int64_t* tokens, int64_t token_stride, int B, int S,
                           const void* targets, int target_kind, const afr_dropout* dropout,
                           double loss_count, float* loss_out, void* stream);
/* The same in two calls, for callers that overlap something with the front-end: afr_train_frontend
 * is everything before fc_output (model.py:167-193; reads only the ten small parameters), so a
 * data-parallel caller lets it run while the all-gather of the fc_output weights updated in the
 * previous step is still in flight, then joins and calls afr_train_loss (model.py:196-202,
 * 304-306: the GEMM with the clamp / MSE / d(logits) epilogue). */
int afr_train_frontend(afr_ctx* ctx, const int64_t* tokens, int64_t token_stride, int B, int S,
                       const afr_dropout* dropout, void* stream);
int afr_train_loss(afr_ctx* ctx, const void* targets, int target_kind, double loss_count,
                   float* loss_out, void* stream);
/* Backward of fc_output w.r.t. its weight and bias (part of loss.backward(), model.py:309) for
 * pixel rows [row_begin, row_end) -- row ranges let a data-parallel caller all-reduce finished
 * buckets while later ones are still being computed. Overwrites the bound gradient rows. */
int afr_train_wgrad(afr_ctx* ctx, int row_begin, int row_end, void* stream);
/* Rest of loss.backward(): d(features) through fc_output, then fc1 / LayerNorm / attention /
 * embedding backward into the ten small bound gradients (overwritten). */
int afr_train_dgrad(afr_ctx* ctx, void* stream);
/* afr_train_dgrad in its two parts, for callers that put them on different streams:
 * afr_train_dgrad_gemm writes d(features) [B, 64*max_length] (library workspace) from
 * d(logits) and the bf16 weights the forward read; afr_train_frontend_backward consumes it. */
int afr_train_dgrad_gemm(afr_ctx* ctx, void* stream);
int afr_train_frontend_backward(afr_ctx* ctx, void* stream);
/* Co-resident launches (single GPU): with `on` != 0, afr_train_wgrad_adamw and
 * afr_train_dgrad_gemm use a footprint of half an SM each (128-wide tiles, two operand stages,
 * 256 tensor-memory columns, <= 128 registers), so that a caller who enqueues them on two streams
 * gets one CTA of each per SM: the dgrad GEMM (tensor-bound) runs under the HBM-bound AdamW GEMM.
 * Results are bit-identical to the default footprint. */
int afr_set_coresident(afr_ctx* ctx, int on);
/* Convenience: afr_train_forward_loss + afr_train_wgrad(0, H*W) + afr_train_dgrad. */
int afr_train_step(afr_ctx* ctx, const int64_t* tokens, int64_t token_stride, int B, int S,
                   const void* targets, int target_kind, const afr_dropout* dropout,
                   double loss_count, float* loss_out, void* stream);

/* Generic (unfused-loss) training path, so that forward() stays differentiable for any loss the
 * caller writes in PyTorch: forward keeps the logits, backward takes d(loss)/d(sheet) [B,H*W] fp32. */
int afr_forward_train(afr_ctx* ctx, const int64_t* tokens, int64_t token_stride, int B, int S,
                      const afr_dropout* dropout, float* sheet_out, void* stream);
int afr_backward(afr_ctx* ctx, const float* dsheet, void* stream);

/* optimizer.step() of optim.AdamW (model.py:273,310), decoupled weight decay, torch's arithmetic.
 * step is 1-based. The fc_output.weight sweep also refreshes the bf16 shadow.
 * afr_adamw_rows updates fc_output.weight rows [row_begin,row_end); afr_adamw_small the other
 * eleven tensors (fc_output.bias included); afr_adamw_step both, over everything. */
int afr_adamw_step(afr_ctx* ctx, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int64_t step, void* stream);
int afr_adamw_rows(afr_ctx* ctx, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int64_t step, int row_begin, int row_end, void* stream);
/* Row-sharded data parallel step over NVLink peer memory (no NCCL call): AdamW on the owned rows
 * [row_begin, row_end) whose gradient is the sum over `world` ranks, read directly from every
 * rank's dW buffer (peer_grads[q] = base of rank q's fc_output.weight.grad, peer-mapped), and whose
 * updated bf16 weights are stored into every rank's INACTIVE shadow copy (peer_shadows[q] = base
 * of that copy on rank q, peer-mapped; see afr_bind_shadow / afr_shadow_commit). The caller
 * provides the cross-rank barriers: all wgrad GEMMs done before, all stores delivered after.
 * Runs on `ctas` CTAs so the remaining SMs keep the compute kernels. */
int afr_adamw_rows_gather(afr_ctx* ctx, double lr, double beta1, double beta2, double eps,
                          double weight_decay, int64_t step, int row_begin, int row_end,
                          const void* const* peer_grads, void* const* peer_shadows, int world,
                          int ctas, void* stream);
/* The same through NVSwitch multicast (NVLS): grad_multicast / shadow_multicast are the MULTICAST
 * addresses of the symmetric gradient buffer and of the inactive bf16 copy (base of the tensor).
 * One multimem.ld_reduce returns the gradient summed over all ranks inside the switch, one
 * multimem.st delivers the updated bf16 weights to every rank: the kernel's link traffic no longer
 * grows with the number of ranks. The in-switch sum order is the hardware's, not the rank order
 * (fp32 results agree with afr_adamw_rows_gather to ~1e-7 relative). Same barriers as above. */
int afr_adamw_rows_gather_nvls(afr_ctx* ctx, double lr, double beta1, double beta2, double eps,
                               double weight_decay, int64_t step, int row_begin, int row_end,
                               const void* grad_multicast, void* shadow_multicast, int ctas,
                               void* stream);
/* afr_adamw_rows as a BACKGROUND kernel: a persistent launch of `ctas` small CTAs (128 threads,
 * <= 40 registers, `stages` x 8 KB of shared memory; 0 = one CTA per SM / 4 stages; built depths
 * 2, 3, 4, 6, 8, 12) in which every thread streams its own 16-byte groups of p / g / m / v through
 * a private slot of a cp.async ring, so that -- enqueued on its own stream -- it shares every SM
 * with the compute kernels of the step (front-end kernels: registers and shared memory are theirs,
 * the HBM bandwidth is idle; see afr_set_smem_reserve) instead of taking the GPU for itself.
 * training.py runs it with 8 stages after the dgrad GEMM. Bit-identical to
 * afr_adamw_rows. grad_rows: the gradient of rows [row_begin,row_end) ([rows, 64*max_length] fp32,
 * e.g. the L2-resident chunk afr_train_wgrad_to has just written), or NULL for the bound
 * fc_output.weight.grad. Replaces optimizer.step() (model.py:310) for those rows. */
int afr_adamw_rows_bg(afr_ctx* ctx, double lr, double beta1, double beta2, double eps,
                      double weight_decay, int64_t step, int row_begin, int row_end,
                      const float* grad_rows, int ctas, int stages, void* stream);
/* afr_train_wgrad into caller memory: d(loss)/d(fc_output.weight[row_begin:row_end]) is written to
 * grad_rows ([rows, 64*max_length] fp32) instead of the bound gradient tensor, so a caller can
 * cycle a few chunk buffers that stay in the 126 MB L2 between this GEMM and the optimizer kernel
 * that consumes them (the 491 MB gradient of model.py:309 is then never in HBM as a whole).
 * with_bias != 0 also writes fc_output.bias.grad[row_begin:row_end] (bound gradients). */
int afr_train_wgrad_to(afr_ctx* ctx, int row_begin, int row_end, float* grad_rows, int with_bias,
                       void* stream);
/* afr_train_wgrad_to with a bf16 destination: the data-parallel gradient that crosses NVLink, at
 * half the bytes of the fp32 form (the sum over ranks is still taken in fp32, below). */
int afr_train_wgrad_to_bf16(afr_ctx* ctx, int row_begin, int row_end, void* grad_rows_bf16, int with_bias,
                            void* stream);
/* afr_adamw_rows_gather / afr_adamw_rows_gather_nvls for bf16 gradient buffers (peer_grads[q] /
 * grad_multicast = base of a bf16 [H*W, 64*max_length] buffer written by afr_train_wgrad_to_bf16):
 * every rank ships half the bytes per step; the per-element sum over ranks is accumulated in fp32
 * (registers, or inside the NVSwitch: multimem.ld_reduce .acc::f32 .bf16x2). Replaces the
 * all-reduce of model.py:309-310's gradient in a data-parallel run. */
int afr_adamw_rows_gather_bf16(afr_ctx* ctx, double lr, double beta1, double beta2, double eps,
                               double weight_decay, int64_t step, int row_begin, int row_end,
                               const void* const* peer_grads, void* const* peer_shadows, int world,
                               int ctas, void* stream);
int afr_adamw_rows_gather_nvls_bf16(afr_ctx* ctx, double lr, double beta1, double beta2, double eps,
                                    double weight_decay, int64_t step, int row_begin, int row_end,
                                    const void* grad_multicast, void* shadow_multicast, int ctas,
                                    void* stream);
int afr_adamw_small(afr_ctx* ctx, double lr, double beta1, double beta2, double eps,
                    double weight_decay, int64_t step, void* stream);
/* loss.backward() w.r.t. fc_output.weight / .bias (model.py:309) AND optimizer.step() of
 * fc_output.weight (model.py:310) for pixel rows [row_begin, row_end) in ONE kernel: the wgrad
 * GEMM's accumulator is the gradient, its epilogue streams the matching tiles of the parameter and
 * the Adam moments through shared memory, applies afr_adamw_rows' arithmetic (bit-identical) and
 * writes p, exp_avg, exp_avg_sq and the bf16 copy. The 491 MB gradient is never written or re-read
 * (26 instead of 38 bytes of HBM traffic per parameter for the two reference statements), so the
 * bound fc_output.weight.grad is left UNTOUCHED by this call. Single-GPU only (a data-parallel
 * gradient must be summed over ranks first). Replaces afr_train_wgrad + afr_adamw_rows for the
 * weight; the bias gradient of the same rows comes from afr_train_bgrad. Row ranges must be
 * multiples of 32. */
int afr_train_wgrad_adamw(afr_ctx* ctx, double lr, double beta1, double beta2, double eps,
                          double weight_decay, int64_t step, int row_begin, int row_end,
                          void* stream);
/* fc_output.bias.grad[row_begin:row_end] alone (the part of afr_train_wgrad that
 * afr_train_wgrad_adamw leaves out): column sums of d(loss)/d(logits), deterministic order. */
int afr_train_bgrad(afr_ctx* ctx, int row_begin, int row_end, void* stream);

/* Synchronises the stream and reports AFR_ERR_TOKEN_RANGE if any token id seen since the last
 * call was outside [0, vocab) (the reference fails with IndexError at model.py:167). */
int afr_check_tokens(afr_ctx* ctx, void* stream);

/* Introspection for tests and benches. which: 0 = features bf16 [B, 64*max_length],
 * 1 = d(logits) residual bf16 [B, H*W] (unscaled (y-t)*mask), 2 = d(features) fp32,
 * 3 = bf16 shadow of fc_output.weight, 4 = logits fp32 of the generic path. */
int afr_workspace_ptr(afr_ctx* ctx, int which, void** ptr, size_t* bytes);
/* Device-to-device copy of the first `bytes` bytes of that workspace into caller memory. */
int afr_workspace_copy(afr_ctx* ctx, int which, void* dst, size_t bytes, void* stream);
/* Tile width (UMMA N) the three GEMMs will use for batch B: out[0..2] = forward, dgrad, wgrad. */
int afr_gemm_tiles(afr_ctx* ctx, int B, int* out_bn3);
/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t afr_launch_count(const afr_ctx* ctx);

/* Profiling builds only (nvcc -DAFR_PHASE_TIMING): copies 2 x 16 clock64() cycle counters (front-end
 * forward, front-end backward; summed over CTAs) into `out` and optionally resets them.
 * AFR_ERR_UNSUPPORTED in a normal build. Synchronises the device. */
int afr_debug_phase_cycles(unsigned long long* out, int reset);

/* Test hooks: the fp32 front-end (embedding / attention / LayerNorm / fc1, model.py:167-193) in
 * isolation, so its forward and backward can be checked at fp32 tolerance without the bf16 GEMM
 * in between. Forward writes the features as fp32 [B, 64*max_length]; backward takes
 * d(loss)/d(features) fp32 [B, 64*max_length] and overwrites the ten small bound gradients. */
int afr_debug_frontend_forward(afr_ctx* ctx, const int64_t* tokens, int64_t token_stride, int B,
                               int S, const afr_dropout* dropout, float* feats_f32, void* stream);
int afr_debug_frontend_backward(afr_ctx* ctx, const int64_t* tokens, int64_t token_stride, int B,
                                int S, const afr_dropout* dropout, const float* dfeat, void* stream);

/* Test hook for the AdamW arithmetic: the kernels compute the optimizer's two divisions and its
 * square root with branch-free round-to-nearest sequences (afr_internal.h). For n device floats
 * a[i], b[i]: q[i] = that division a/b, s[i] = that square root of |a|, and q_ieee / s_ieee the
 * results of the IEEE instructions (div.rn.f32 / sqrt.rn.f32) on the same inputs. */
int afr_debug_div_sqrt(const float* a, const float* b, float* q, float* s, float* q_ieee,
                       float* s_ieee, int64_t n, void* stream);

/* Diagnostic: D[M,N] (fp32, ld = ldd) = alpha * A * B^T with bf16 operands on the tcgen05 path.
 * a_mn_major / b_mn_major: operand stored [K, M] resp. [K, N] row-major instead of [M, K] / [N, K].
 * use_tma_store is a flag word: bit 0 = fp32 output through TMA stores, bit 1 = run as CTA pairs
 * (tcgen05 cta_group::2, 256-row tiles: what the forward / dgrad / wgrad / render GEMMs use unless
 * AFR_CTA2=0).
 * Exists so the tensor-core kernel can be tested against a plain matmul in isolation. */
int afr_gemm_bf16(int device, const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                  int b_mn_major, float* D, int64_t ldd, int M, int N, int K, int tile_n, float alpha,
                  int use_tma_store, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AFR_SM100_H_ */
