/*
 * gemmgan.h — C ABI of libgemmgan_sm100a.so
 *
 * Drop-in boundary for the WGAN-GP training step of GeMM-GAN. The reference
 * (pure PyTorch, /root/reference/src) has no native interface of its own: every
 * entry point below replaces a group of torch library calls made by the
 * reference's Python hot path, cited as `file:line` of the reference.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless a
 *     name ends in `_host`. The caller owns every buffer; the library never
 *     allocates or frees device memory except inside gg_engine_create/destroy
 *     for objects it returns a handle to (tensor-map caches, no tensors).
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as
 *     void*); no implicit synchronisation; re-entrant across streams.
 *   - return value 0 = success, negative = error; gg_last_error() returns a
 *     thread-local message. There is no CPU fallback: on a device that is not
 *     sm_100 every compute entry point returns GG_ERR_ARCH.
 *   - row-major, batch-first tensors; bool masks are uint8 with 1 = padding
 *     (src/multi_patch_multi_token_gan_dataloader.py:46-47).
 */
#ifndef GEMMGAN_H
#define GEMMGAN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_OK 0
#define GG_ERR_ARG (-1)
#define GG_ERR_ARCH (-2)
#define GG_ERR_CUDA (-3)
#define GG_ERR_WORKSPACE (-4)

#define GG_ABI_VERSION 4

const char* gg_last_error(void);
int gg_abi_version(void);
/* 0 when device `dev` can run the kernels (compute capability 10.x), else GG_ERR_ARCH. */
int gg_check_device(int dev);

/* ------------------------------------------------------------------ GEMM --
 * D[M,N] = epilogue( alpha * sum_seg A_seg * B_seg^T ), bf16 operands, fp32
 * accumulation in tensor memory (tcgen05.mma, TMA-fed).
 * Replaces nn.Linear / F.linear and the autograd-generated dgrad / wgrad GEMMs
 * (src/conditional_gan_cross_attention_with_film.py:56-72, 129-162, 226-231).
 *
 * Operand storage (row-major, leading dimension in elements, multiple of 8):
 *   a_mn_major = 0 : A_seg stored [M, K]   (forward X, dgrad dY)
 *   a_mn_major = 1 : A_seg stored [K, M]   (wgrad: dY^T without a transpose pass)
 *   b_mn_major = 0 : B_seg stored [N, K]   (forward W)
 *   b_mn_major = 1 : B_seg stored [K, N]   (dgrad W, wgrad X)
 * Up to two K segments are accumulated into one tile (torch.cat((x, c), 1) of
 * :157 / :226 becomes two segments; GP's Gram term rides on the W1 wgrad).
 */
#define GG_ACT_NONE 0
#define GG_ACT_LEAKY 1 /* v > 0 ? v : slope * v ; slope 0.0 = ReLU (:56-72) */
#define GG_ACT_FILM 2  /* columns [0,N/2): tanh ; [N/2,N): clamp(-5,5) (:129-134) */

typedef struct gg_epilogue {
  float alpha;           /* accumulator scale (1.0f if unused) */
  const float* bias;     /* [N] fp32, or NULL */
  const void* pre;       /* [M, pre_ld] added BEFORE the activation, or NULL */
  int64_t pre_ld;
  int32_t pre_f32;       /* 1 = fp32, 0 = bf16 */
  int32_t act;           /* GG_ACT_* */
  float slope;
  float drop_p;          /* inverted dropout after the activation; 0 = off */
  const uint64_t* rng;   /* device: {seed, step}; required when drop_p > 0 */
  uint32_t site;         /* dropout call-site id (decorrelates sites) */
  const void* mask;      /* [M, mask_ld]: v *= (mask > 0 ? mask_pos : mask_neg), or NULL */
  int64_t mask_ld;
  int32_t mask_f32;
  float mask_pos, mask_neg;
  const void* res;       /* [M, res_ld] added AFTER everything else, or NULL */
  int64_t res_ld;
  int32_t res_f32;
  void* out_bf16;        /* [*, ld_bf16] bf16 output or NULL. With a 16-byte aligned base and pitch the tile is written
                            by TMA stores, which clip at 16-byte granularity: when N is not a multiple of 8 (4 for
                            fp32) the row padding up to that multiple receives zeros */
  int64_t ld_bf16;
  float* out_f32;        /* [*, ld_f32] fp32 output or NULL */
  int64_t ld_f32;
  int32_t accum_f32;     /* 1: out_f32 += v instead of = v */
  /* output row remap (0 = identity): row = (m / row_div) * row_mul + row_add + m % row_div */
  int32_t row_div, row_mul, row_add;
} gg_epilogue;

typedef struct gg_gemm_seg {
  const void* a;
  const void* b;
  int64_t lda, ldb;
  int32_t K;
} gg_gemm_seg;

#define GG_IMPL_TCGEN05 0
#define GG_IMPL_SIMT_F32 1 /* CUDA-core fp32 check path on the same bf16 operands (tests only) */

typedef struct gg_gemm_desc {
  int32_t M, N;
  int32_t nseg;
  gg_gemm_seg seg[2];
  int32_t a_mn_major, b_mn_major;
  gg_epilogue epi;
  void* workspace;          /* split-K partials, or NULL (then no split-K) */
  int64_t workspace_bytes;
  int32_t impl;             /* GG_IMPL_* */
  int32_t force_splits;     /* 0 = heuristic */
  int32_t block_n;          /* 0 = heuristic; 64, 128 or 256 */
  int32_t light;            /* 0 = heuristic; 1 = force, -1 = forbid the two-CTAs-per-SM configuration (128-wide
                               tiles, no split-K) used for short-K products with many tiles */
  int32_t pair;             /* 0 = heuristic; 1 = force, -1 = forbid the CTA-pair configuration (tcgen05 cta_group::2:
                               256 x 256 tiles over the two SMs of a TPC, each CTA stages half of the B tile) */
  int32_t tf32_operands;    /* 1: A and B are FP32 in memory (pitches in fp32 elements, multiples of 4, both K-major) and
                               are multiplied as TF32 (tcgen05.mma kind::tf32, fp32 accumulate): fp32 tensors are read
                               once, with no bf16 copy (the gradient-penalty microbenchmark's real / fake profiles) */
} gg_gemm_desc;

int gg_gemm_bf16(const gg_gemm_desc* desc, void* stream);
/* Measurement hooks (bench.py): kernels launched by the library so far; CUDA-event timing of every
 * tcgen05 GEMM launch between begin/end (summed ms, 2*M*N*K FLOPs, launch count). */
long long gg_launch_count(int reset);
void gg_launch_count_add(long long n); /* a replayed CUDA graph adds the launches it contains */
int gg_gemm_profile_begin(void);
int gg_gemm_profile_end(double* ms, double* flops, long long* launches);
double gg_gemm_profile_bytes(void); /* algorithmic bytes (operands once + outputs once) of the last profiled region */
int gg_gemm_profile_dump(const char* csv_path); /* per-launch shapes and durations of the last region */
/* Diagnostics: CTA `cta` of every following GEMM launch stamps clock64() per pipeline role into device_buf
 * (6 x 512 int64: TMA issued / stage full / tile committed / accumulator full / tile stored / origin); NULL = off. */
int gg_gemm_set_trace(void* device_buf, int cta);
/* In-kernel timing of every following GEMM launch (also inside replayed CUDA graphs): slot i of device_buf
 * (2 x capacity uint64, preset to {UINT64_MAX, 0} by the caller before each run) receives {min start, max end} in
 * %globaltimer ns. gg_gemm_timer_slots returns the number of slots handed out and their algorithmic FLOPs / bytes. */
int gg_gemm_set_timer(void* device_buf, int capacity);
int gg_gemm_timer_slots(double* flops, double* bytes, int n);

/* -------------------------------------------------------- training engine --
 * One engine = one (generator, critic) pair of one model variant at a fixed per-rank batch
 * size. It executes the bodies of WGAN_GP.train_disc / train_gen / generate_samples
 * (src/conditional_gan_cross_attention_with_film.py:376-461, :601-608; film variant
 * conditional_gan_film.py:347-430; vanilla vanilla_gan_unconditional.py:329-418) as a fixed
 * sequence of kernels with a hand-written backward and double-backward (no autograd replay).
 *
 * Parameters, gradients and optimizer state stay in caller-owned flat fp32 buffers (the
 * PyTorch nn.Parameters are views into them); `off[slot]` locates each tensor, in the
 * reference's native [out, in] layout. All activations live in one caller-owned workspace.
 */
#define GG_VARIANT_VANILLA 0 /* vanilla_gan_unconditional.py: trunk only                      */
#define GG_VARIANT_FILM 1    /* conditional_gan_film.py: FiLM + encoder (no bias), CLS vector */
#define GG_VARIANT_PAPER 2   /* conditional_gan_cross_attention_with_film.py (paper model)    */
#define GG_VARIANT_CONCAT 4  /* conditional_gan_concat.py: c = Linear(text embedding) ('text'), or Linear(masked mean
                                patch embedding) ('image', :137-138); parameter slots GG_P_TEXT_W / GG_P_TEXT_B = encoder */
#define GG_VARIANT_IMG 5     /* conditional_gan_img_transformer.py: patch encoder Linear -> ReLU -> LayerNorm (:111-115), no
                                text, bias-free encoder layers, CLS vector as conditioning (:131-133) */
#define GG_VARIANT_LABEL 6   /* benchmark_generative_model.py (label-conditioned baseline, :101-236): c = [emb0[y0] | emb1[y1]],
                                two nn.Embedding tables of width E/2 per net; cfg.Dt / cfg.Dp = their vocabulary sizes;
                                labels staged with gg_engine_set_labels; no dropout, no attention                      */
#define GG_VARIANT_ATTN 7    /* conditional_gan_attention.py (:92-170): c = MultiheadAttention(query = text_encoder(text
                                embedding) [B, 1, E], keys / values = patches_encoder(patches), key_padding_mask),
                                no encoder layers, no dropout; the generator passes c through BatchNorm1d (attn_bn, :108,
                                :126). Slots: GG_P_TEXT_*, GG_P_PATCH_*, GG_P_P2T_* = `attention`, GG_P_BN_* = attn_bn;
                                running statistics through gg_engine_set_batchnorm                                     */
#define GG_VARIANT_CROSS 3   /* conditional_gan_cross_attention.py: the paper model's towers without FiLM and without
                                tower biases (:97-206); only row 0 of its multi-query cross-attentions reaches the
                                conditioning vector, so it runs as the same single-query tail */

#define GG_OPT_RMSPROP 0
#define GG_OPT_ADAM 1
#define GG_OPT_ADAMW 2

enum gg_param_slot {
  GG_P_FILM_W = 0, GG_P_FILM_B, GG_P_TEXT_W, GG_P_TEXT_B, GG_P_PATCH_W, GG_P_PATCH_B, GG_P_CLS,
  /* encoder layer l: GG_P_LAYER0 + 12*l + {IN_W, IN_B, OUT_W, OUT_B, FF1_W, FF1_B, FF2_W, FF2_B,
   *                                        N1_W, N1_B, N2_W, N2_B} */
  GG_P_LAYER0 = 7,
  GG_P_P2T_IN_W = GG_P_LAYER0 + 24, GG_P_P2T_IN_B, GG_P_P2T_OUT_W, GG_P_P2T_OUT_B,
  GG_P_T2P_IN_W, GG_P_T2P_IN_B, GG_P_T2P_OUT_W, GG_P_T2P_OUT_B,
  GG_P_TR0_W, GG_P_TR0_B, GG_P_TR1_W, GG_P_TR1_B, GG_P_FIN_W, GG_P_FIN_B,
  GG_NSLOTS,
  /* GG_VARIANT_IMG has no FiLM: its patch-encoder LayerNorm weight / bias [E] live in the FiLM slots */
  GG_P_PENC_LN_W = GG_P_FILM_W, GG_P_PENC_LN_B = GG_P_FILM_B,
  /* GG_VARIANT_LABEL: the two embedding tables [vocab_i, E/2] (fp32, gathered directly: no bf16 shadow) */
  GG_P_EMB0 = GG_P_TEXT_W, GG_P_EMB1 = GG_P_PATCH_W,
  /* GG_VARIANT_ATTN (generator only): BatchNorm1d weight / bias [E] in the FiLM slots */
  GG_P_BN_W = GG_P_FILM_W, GG_P_BN_B = GG_P_FILM_B
};
enum gg_layer_slot {
  GG_L_IN_W = 0, GG_L_IN_B, GG_L_OUT_W, GG_L_OUT_B, GG_L_FF1_W, GG_L_FF1_B, GG_L_FF2_W, GG_L_FF2_B,
  GG_L_N1_W, GG_L_N1_B, GG_L_N2_W, GG_L_N2_B, GG_L_COUNT
};

typedef struct gg_model_cfg {
  int32_t variant;
  int32_t B;              /* per-rank batch */
  int32_t G, L, E, H;     /* genes, latent, embedding, trunk hidden width */
  int32_t Dt, Dp;         /* text / patch feature widths (768 / 1024) */
  int32_t P, T;           /* patch tokens, text tokens (film: T = 1, text is [B, Dt]) */
  int32_t n_layers, n_heads, ffn;
  int32_t tower_bias;     /* 1: encoder layers carry biases (paper), 0: bias=False (film) */
  float slope;            /* LeakyReLU negative slope of the trunk (scripts use 0.0) */
  float dropout_p;        /* encoder dropout in train mode (reference: 0.1) */
  float gp_weight;        /* 10 */
  float clip_d, clip_g;   /* clip_grad_norm_ max norms; <= 0 disables (:414, :457) */
  float ln_eps;           /* 1e-5 */
  int32_t optimizer;      /* GG_OPT_* */
  int32_t gemm_impl;      /* GG_IMPL_TCGEN05, or GG_IMPL_SIMT_F32 for the check path */
  uint64_t seed;          /* dropout stream seed */
} gg_model_cfg;

typedef struct gg_net_buffers {
  float* params;          /* flat fp32, 16-byte aligned */
  float* grads;
  float* exp_avg;         /* Adam/AdamW first moment; NULL for RMSprop */
  float* exp_avg_sq;      /* second moment / RMSprop square_avg */
  float* step_count;      /* device float[1]: optimizer steps taken (Adam bias correction) */
  int64_t n_used;         /* [0, n_used) are trained; never-used tensors (the prototype encoder
                             layer, :114) sit behind and are left untouched like in the reference */
  int64_t off[GG_NSLOTS]; /* element offsets, -1 = tensor absent in this variant */
} gg_net_buffers;

#define GG_NET_GEN 0
#define GG_NET_DISC 1

/* device-resident results, float[GG_STATS_COUNT] (gg_engine_stats) */
#define GG_STAT_LOSS_REAL 0  /* -mean D(real)                         (:407, :41-46) */
#define GG_STAT_LOSS_FAKE 1  /*  mean D(fake)                                        */
#define GG_STAT_GP 2         /*  mean (||grad||-1)^2                  (:372-374)     */
#define GG_STAT_D_LOSS 3     /*  loss_real + loss_fake + gp_weight*gp (:409)         */
#define GG_STAT_G_LOSS 4     /* -mean D(G(z))                         (:452)         */
#define GG_STAT_D_GRAD_NORM 5
#define GG_STAT_D_CLIP_COEF 6
#define GG_STAT_G_GRAD_NORM 7
#define GG_STAT_G_CLIP_COEF 8
#define GG_STATS_COUNT 16

typedef struct gg_engine gg_engine;

int gg_engine_workspace_bytes(const gg_model_cfg* cfg, int64_t* bytes);
int gg_engine_create(const gg_model_cfg* cfg, const gg_net_buffers* gen, const gg_net_buffers* disc,
                     void* workspace, int64_t workspace_bytes, void* stream, gg_engine** out);
void gg_engine_destroy(gg_engine* e);
/* Concurrency inside one entry point: enabled != 0 (default) forks work nothing on the dependent chain waits
 * for (weight / bias gradients, the generator forward next to the critic tower) onto engine-owned side streams
 * and joins them before returning; 0 keeps every kernel on the caller's stream (profiling, debugging). The
 * setting must not change between capture and replay of a CUDA graph. */
int gg_engine_set_lanes(gg_engine* e, int enabled);
/* fp32 master weights -> bf16 shadows (call after any external change of `params`). */
int gg_engine_refresh_shadows(gg_engine* e, int net, void* stream);
/* Stages one batch (the dataloader tuple, already on the device, fp32 / uint8 masks):
 * casts to bf16 once per train() call (:465-469). Unused pointers may be NULL per variant. */
int gg_engine_set_batch(gg_engine* e, const float* genes, const float* patches, const uint8_t* patch_pad,
                        const float* text, const uint8_t* text_pad, void* stream);
/* GG_VARIANT_LABEL: stages the two categorical covariates of the batch (device int64 [B] each, the dataloader's
 * disease_type / primary_site columns; benchmark_generative_model.py:138-150). Values must lie in [0, vocab_i):
 * nn.Embedding raises on anything else, the caller checks (the kernel clamps instead of faulting). */
int gg_engine_set_labels(gg_engine* e, const int64_t* labels0, const int64_t* labels1, void* stream);
/* GG_VARIANT_ATTN: the generator's BatchNorm1d buffers (device fp32 [E] each, torch's running_mean / running_var, updated
 * in place by every training-mode generator forward exactly like nn.BatchNorm1d: running = (1 - momentum) * running +
 * momentum * batch statistic, the variance unbiased), its momentum and eps (torch defaults 0.1 / 1e-5). Must be called
 * before the first forward of an ATTN engine. */
int gg_engine_set_batchnorm(gg_engine* e, float* running_mean, float* running_var, float momentum, float eps);
/* train_disc minus optimizer: fills critic grads + stats. z [B,L], alpha [B,1] fp32. training=1
 * uses dropout_p (three independently-dropped critic tower passes, as the reference).
 * training | GG_TRAIN_GEN_EVAL: the generator forward inside the critic step runs in eval mode (no dropout in its tower,
 * BatchNorm on its running statistics, which are left alone). The reference's train_disc uses the generator in whatever
 * mode it was left in (:390; only train_gen calls gen.train(), :427), so the critic steps that follow a generate_samples
 * call (gen.eval(), :603) see it that way. */
#define GG_TRAIN_GEN_EVAL 2
int gg_engine_disc_grads(gg_engine* e, const float* z, const float* alpha, int training, void* stream);
/* train_gen minus optimizer: fills generator grads + stats. */
int gg_engine_gen_grads(gg_engine* e, const float* z, int training, void* stream);
/* The same steps in two halves, for data-parallel runs that overlap the gradient all-reduce with the rest of
 * the backward: phase 1 = forward + trunk backward (on return every trunk gradient — slots GG_P_TR0_W ..
 * GG_P_FIN_B, one contiguous range at the end of `grads` — is final), phase 2 = fusion-tower backward (the
 * remaining slots). Phase 2 must follow phase 1 of the same step on the same stream.
 * Finer cut: phase GG_PHASE_STAGE0 + s runs stage s of the tower backward alone (s = 0: cross-attention tail — slots
 * GG_P_P2T_* / GG_P_T2P_* and the text encoder final; s = 1 .. n_layers: encoder layers from the last to the first
 * — that layer's 12 slots final; s = n_layers + 1: CLS / patch encoder / FiLM — slots 0..6 final). OR-ing
 * GG_PHASE_NO_JOIN leaves the side lanes open when the call returns (a later call without it joins them); the
 * trainer then orders its communication stream behind them with gg_engine_lanes_signal and lets lane 0 run on. */
#define GG_PHASE_STAGE0 16
#define GG_PHASE_NO_JOIN 64
int gg_engine_disc_grads_phase(gg_engine* e, const float* z, const float* alpha, int training, int phase,
                               void* stream);
int gg_engine_gen_grads_phase(gg_engine* e, const float* z, int training, int phase, void* stream);
int gg_engine_lanes_signal(gg_engine* e, void* stream);
/* clip_grad_norm_ (if configured) + optimizer.step() on the flat buffers + shadow refresh.
 * In data-parallel runs the caller all-reduces `grads` between *_grads and this call. */
int gg_engine_optim_step(gg_engine* e, int net, float lr, void* stream);
/* generator forward only: out_f32 [B, G] (generate_samples :601-608; training=0 = eval mode). */
int gg_engine_generate(gg_engine* e, const float* z, float* out_f32, int training, void* stream);
/* critic forward only on `genes_f32` [B, G] with the staged conditioning: score_f32 [B]. */
int gg_engine_critic(gg_engine* e, const float* genes_f32, float* score_f32, int training, void* stream);
/* WGAN_GP.gradient_penalty (:351-374) on caller-provided real / fake [B, G] fp32 (real may be NULL =
 * the staged batch): gp_out[0] (device) = mean_b (||dD/dx_hat||_2 - 1)^2. */
int gg_engine_gradient_penalty(gg_engine* e, const float* real_f32, const float* fake_f32, const float* alpha,
                               int training, float* gp_out, void* stream);
/* The gradient penalty alone, value and gradients, on the unconditional critic (BASELINE.json config 5: the GP
 * microbenchmark). Replaces WGAN_GP.gradient_penalty (src/vanilla_gan_unconditional.py:304-327) + the GP part of
 * disc_loss.backward() (:381): gp_out[0] (device, may be NULL) = GP; gp_weight * dGP/dW1, dW2, dw3 are written
 * into the critic's gradient buffer (biases get no GP gradient: the masks are piecewise constant). */
int gg_engine_gp_step(gg_engine* e, const float* real_f32, const float* fake_f32, const float* alpha, float* gp_out,
                      void* stream);
/* Module-level autograd (SURVEY.md section 8 b: generator.forward / discriminator.forward of
 * src/conditional_gan_cross_attention_with_film.py:128-164 / :198-233 as free-standing differentiable nn.Modules, outside
 * WGAN_GP.train): gg_engine_generate_keep / gg_engine_critic_keep are gg_engine_generate / gg_engine_critic keeping what
 * the backward reads (one replica); the *_backward calls take the upstream gradient (d loss / d output, device fp32
 * [B, G] / [B]) and fill the net's gradient buffer (overwriting it, like the grads entry points), plus — when the
 * pointer is not NULL — the gradient w.r.t. the first input (z [B, L] / gene profiles [B, G], fp32). First order only:
 * the gradient penalty's double backward is gg_engine_disc_grads / gg_engine_gp_step. The staged batch must not change
 * between a *_keep forward and its backward; a later forward of the same net replaces what was kept. */
int gg_engine_generate_keep(gg_engine* e, const float* z, float* out_f32, int training, void* stream);
int gg_engine_generate_backward(gg_engine* e, const float* dout_f32, float* dz_f32, void* stream);
int gg_engine_critic_keep(gg_engine* e, const float* genes_f32, float* score_f32, int training, void* stream);
int gg_engine_critic_backward(gg_engine* e, const float* dscore_f32, float* dgenes_f32, void* stream);
/* Critic layer 1 on the two fp32 [B, K] gene matrices of the gradient penalty (real / fake of WGAN_GP.gradient_penalty,
 * src/vanilla_gan_unconditional.py:304-327), read in place: out[0:B] = x0 . w^T, out[B:2B] = x1 . w^T (fp32 [2B, 256]);
 * w = bf16 [256, K] with pitch ldw (elements). The fp32 -> bf16 conversion happens on chip and the weight k-blocks are
 * TMA-multicast across a cluster of row tiles (csrc/xw_f32.cu). x rows 16-byte aligned, K % 4 == 0. workspace (may be
 * NULL): split-K partial sums, up to 8 * 2B * 256 floats are used. */
int gg_xw_f32(const float* x0, const float* x1, int32_t B, int32_t K, const void* w_bf16, int64_t ldw, float* out,
              void* workspace, int64_t workspace_bytes, void* stream);
/* FiLM + patch encoder + CLS token of the paper / film models in one kernel (src/conditional_gan_cross_attention_with_film.py
 * :129-142: `patches = gamma * patches + beta`, `patches_encoder(patches)`, `torch.cat((cls, patches), 1)`): x0 [R, B, P + 1,
 * 256] bf16 with row (r, b, 0) = cls and row (r, b, 1 + j) = (gamma_b * patches[b, j] + beta_b) w^T + bias for each of the R
 * dropout replicas. patches bf16 [B * P, Dp] (Dp % 64 == 0), gamma_beta fp32 [B, 2 * Dp] (gamma | beta, already tanh'ed /
 * clamped), w bf16 [256, Dp] with pitch ldw, bias (may be NULL) / cls fp32 [256]. mod (may be NULL) receives the
 * modulated patches [B * P, Dp] bf16 that the weight gradient of the encoder reads in the backward. */
int gg_film_patch_encode(const void* patches_bf16, const float* gamma_beta, const void* w_bf16, int64_t ldw, const float* bias,
                         const float* cls, void* x0_bf16, void* mod_bf16, int32_t B, int32_t P, int32_t R, int32_t Dp,
                         void* stream);
/* `x = norm(x + dropout(linear(a)))` of a post-norm encoder layer (torch/nn/modules/transformer.py: norm1 / _sa_block's
 * out_proj, norm2 / linear2; reference :114-119) as one kernel for d_model = 256: z = res + dropout(a w^T + bias) (bf16, kept
 * for the backward), out = LayerNorm(z) * gamma + beta, mean / rstd per row (fp32). a bf16 [rows, K] (pitch lda, K % 64 == 0),
 * w bf16 [256, K] (pitch ldw), res / z / out bf16 [rows, 256]; bias / beta may be NULL. Dropout: stream {seed, step} at rng,
 * `site`, element index row * 256 + column (the indexing of the stand-alone LayerNorm kernels). */
int gg_gemm_layernorm(const void* a_bf16, int64_t lda, const void* w_bf16, int64_t ldw, int32_t K, const float* bias,
                      const void* res_bf16, const float* gamma, const float* beta, void* z_bf16, void* out_bf16, float* mean,
                      float* rstd, int64_t rows, float eps, float drop_p, const uint64_t* rng, uint32_t site, void* stream);
/* out[b, :] (fp32) = mean over the rows p with pad[b, p] == 0 of x[b, p, :] — the masked mean of
 * conditional_gan_concat.py:137-138 ('image' conditioning), taken BEFORE the affine encoder. pad may be NULL. */
int gg_masked_mean_rows(const float* x, const uint8_t* pad, float* out, int B, int P, int D, void* stream);
/* Device-side batch assembly: dst[r, 0:cols] = index[r] >= 0 ? src[index[r], 0:cols] : 0 (fp32 rows,
 * pitches in elements, index: device int64 [rows]). With the dataset resident in HBM (patch embeddings of every case
 * as one ragged matrix, gene profiles, token embeddings) one call per tensor builds the batch tuple of
 * MultiPatchMultiTokenGANDataset.__getitem__ + the DataLoader collation
 * (src/multi_patch_multi_token_gan_dataloader.py:25-55: the picked / zero-padded patch rows, -1 = padding row). */
int gg_gather_rows(const float* src, int64_t ld_src, const int64_t* index, float* dst, int64_t ld_dst, int64_t rows,
                   int32_t cols, void* stream);
float* gg_engine_stats(gg_engine* e);
/* named internal device buffers for tests ("fake_bf16", "score", "gp_norms", "cond_disc", ...);
 * returns NULL for unknown names. rows/cols/ld (elements) are optional outputs. */
void* gg_engine_buffer(gg_engine* e, const char* name, int64_t* rows, int64_t* cols, int64_t* ld,
                       int32_t* is_f32);

/* ----------------------------------------------------- standalone kernels --
 * Exposed for unit tests and micro-benchmarks; the engine calls the same code. */
int gg_optim_step(int kind, float* p, float* g, float* m, float* v, int64_t n, float lr, float max_norm,
                  float* step_count, float* norm_out2, float* scratch, void* stream);

/* Multi-head attention core softmax(q k^T / sqrt(hd) + key_padding_mask) v with dropout on the probabilities,
 * forward and hand-written backward (F.multi_head_attention_forward / SDPA inside nn.TransformerEncoderLayer and
 * nn.MultiheadAttention, :114-123, :144-152). bf16 tensors; rows of sequence b are b*L + i; heads are 64-column
 * slices (hd columns each). q / kv / mask may be shared by replicas through the *_mod fields (row block =
 * b % mod). Backward needs `stat` (2*nb*H*Lq floats) for the generic short-sequence path. */
typedef struct gg_attn_args {
  const void* q; int64_t ldq; int32_t q_mod;
  const void* k; const void* v; int64_t ldkv; int32_t kv_mod;
  const uint8_t* mask; int32_t mask_mod;      /* [mask_mod, Lk], 1 = padded key; may be NULL */
  int32_t nb, H, hd, Lq, Lk;
  float drop_p; const uint64_t* rng; uint32_t site;
  void* o; int64_t ldo;                       /* forward output [nb*Lq, H*hd] */
  const void* dout; int64_t lddo;             /* backward inputs / outputs */
  void* dq; int64_t lddq;
  void* dk; void* dv; int64_t lddkv;
  float* stat;
  /* optional (may be NULL): the keep bits of this site's dropout stream drawn once by gg_dropout_bits (bit idx = element
   * idx = ((b * H + h) * Lq + i) * Lk + j, 1 = kept), read by the 17..320-token self-attention kernels in forward and
   * backward instead of drawing Philox groups per score tile; results are identical with and without it */
  const uint32_t* dbits;
} gg_attn_args;
int gg_attention_fwd(const gg_attn_args* a, void* stream);
int gg_attention_bwd(const gg_attn_args* a, void* stream);
/* Keep bits of n_elems consecutive elements of the dropout stream {seed, step} at `rng` (device uint64[2]), site `site`,
 * probability p: out[w] bit k = element 32 * w + k kept. `out` holds gg_dropout_bits_words(n_elems) uint32 words (the
 * count includes the slack the attention kernels' unaligned 16-bit windows may touch). */
int64_t gg_dropout_bits_words(int64_t n_elems);
int gg_dropout_bits(const uint64_t* rng, uint32_t site, float p, int64_t n_elems, uint32_t* out, void* stream);

/* One post-norm nn.TransformerEncoderLayer forward (d_model 256, 4 heads x 64, ffn 512, relu, dropout p) as ONE kernel
 * for short sequences (S <= 16 tokens: the paper model's 8 patches + CLS): in-proj, masked softmax attention, out-proj,
 * residual + dropout + LayerNorm, ffn1 + relu + dropout, ffn2, residual + dropout + LayerNorm
 * (src/conditional_gan_cross_attention_with_film.py:114-119, :144; torch/nn/modules/transformer.py post-norm branch,
 * F.multi_head_attention_forward). x [nb * S, 256] bf16 (row = sequence * S + token); weights bf16 [out, in] with
 * pitches that are multiples of 8; biases / LayerNorm vectors fp32 (b_* and be* may be NULL: bias=False variants).
 * mask [mask_mod, S] uint8 (1 = padded key), sequence b uses row b % mask_mod; may be NULL. Dropout: sites
 * site .. site + 3 (attention probabilities, after out-proj, after relu, after ffn2) of the {seed, step} stream at
 * `rng`, indexed exactly as gg_attention_fwd / the GEMM epilogue / the LayerNorm kernel index them, so the unfused
 * backward regenerates the same masks. Rows < save_rows (negative: all) also write what the backward reads: qkv
 * [rows, 768], ao (attention output), z1 / z2 (pre-LayerNorm sums), x1, h [rows, 512], mean / rstd (fp32 [rows]);
 * with save_rows = 0 those pointers may be NULL and only `out` is written. */
typedef struct gg_enc_layer_params {
  int32_t nb, S, E, F, n_heads;
  int64_t save_rows;
  const void* x;
  const void* w_in; int64_t ld_in;
  const void* w_out; int64_t ld_out;
  const void* w_ff1; int64_t ld_ff1;
  const void* w_ff2; int64_t ld_ff2;
  const float *b_in, *b_out, *b_ff1, *b_ff2, *g1, *be1, *g2, *be2;
  const uint8_t* mask; int32_t mask_mod;
  float drop_p, eps;
  const uint64_t* rng; uint32_t site;
  void *qkv, *ao, *z1, *x1, *h, *z2, *out;
  float *mean1, *rstd1, *mean2, *rstd2;
  /* optional (all three or none): keep bits of the dropout sites site + 1 ([rows, E]), site + 2 ([rows, F]), site + 3
   * ([rows, E]) drawn beforehand by gg_dropout_bits (bit idx = row * width + column); the kernel then reads one 64-bit
   * word per thread and phase instead of running eight Philox groups. Same decisions, same results. */
  const uint32_t *dbits1, *dbits2, *dbits3;
} gg_enc_layer_params;
int gg_encoder_layer_fwd(const gg_enc_layer_params* p, void* stream);
/* The feed-forward half of the same layer's BACKWARD, dependent chain only, as one kernel: LayerNorm-2 backward of
 * dout -> dropout mask (site) -> x W2 -> relu / ffn-dropout mask from the stored activation h -> x W1 -> + residual:
 *   gz = LN2'(dout; z2, mean2, rstd2, gamma2); gh = (mask(gz) W2) * [h > 0] / (1 - p); gb = gz + gh W1.
 * Replaces autograd's LayerNormBackward / AddmmBackward / ReluBackward nodes of linear2 / linear1 inside
 * disc_loss.backward() / gen_loss.backward() (:412, :455; torch/nn/modules/transformer.py _ff_block). w2t / w1t are
 * TRANSPOSED bf16 copies of linear2.weight ([512, 256]) and linear1.weight ([256, 512]) (pitches multiples of 8).
 * Outputs: gh [rows, 512] (the weight gradient of linear1 reads it), gb [rows, 256] (gradient w.r.t. x1). The
 * LayerNorm parameter gradients and the tensors linear2's weight gradient reads stay with the unfused kernels. */
typedef struct gg_enc_ffn_bwd_params {
  int64_t rows;
  const void* dout;
  const void* z2;
  const float *mean2, *rstd2, *gamma2;
  const void* h;
  const void* w2t; int64_t ld_w2t;
  const void* w1t; int64_t ld_w1t;
  float drop_p; const uint64_t* rng; uint32_t site;
  void* gh;
  void* gb;
  const uint32_t* dbits; /* optional: keep bits of `site` ([rows, 256]) as handed to gg_encoder_layer_fwd (dbits3) */
} gg_enc_ffn_bwd_params;
int gg_encoder_ffn_bwd(const gg_enc_ffn_bwd_params* p, void* stream);
/* Diagnostics: CTA 0 of every following gg_encoder_layer_fwd launch stamps clock64() per pipeline role for its first
 * tile into device_buf (3 x 64 int64: TMA producer / MMA issuer / first epilogue warp); NULL = off. */
int gg_enc_layer_set_trace(void* device_buf);
/* Live profile (bench.py): summed CUDA-event duration, algorithmic FLOPs / bytes and count of the fused encoder-layer
 * launches between gg_gemm_profile_begin and gg_gemm_profile_end (call after the latter). */
int gg_enc_layer_profile(double* ms, double* flops, double* bytes, long long* launches);
/* Same for the grouped weight-gradient launches (wgrad_group.cu): 2*M*N*K flops, bf16 operands + fp32 outputs once. */
int gg_wgrad_group_profile(double* ms, double* flops, double* bytes, long long* launches);

/* Grouped weight gradients: out_i[M_i, N_i] (fp32, pitch ld) = dY_i^T X_i for up to 32 problems in ONE launch
 * (autograd's grad_output.t().mm(input) of every Linear of one backward pass, :412 / :455). dY_i is stored
 * [K_i rows, M_i], X_i [K_i rows, N_i], both bf16 with 16-byte aligned bases and pitches that are multiples of
 * 8. `workspace` (gg_wgrad_group_workspace_bytes(sum_i M_i*N_i) bytes) holds split-K partials behind 64 KiB of
 * arrival counters; the counters must be zero before the first launch and every launch leaves them zero.
 * Deterministic (partials are summed in split order). */
typedef struct gg_wgrad_item {
  const void* dy; int64_t ld_dy;
  const void* x; int64_t ld_x;
  int32_t M, N, K;
  float* out; int64_t ld;
  float* bias; /* optional [M] fp32: the bias gradient sum_k dY[k, m] (autograd's grad_output.sum(0)), formed by one
                  extra N = 16 tensor-core MMA per K step against an all-ones operand; NULL = not wanted */
} gg_wgrad_item;
int64_t gg_wgrad_group_workspace_bytes(int64_t sum_output_elems);
int gg_wgrad_group(const gg_wgrad_item* items, int n, void* workspace, int64_t workspace_bytes, void* stream);
/* Grouped column sums: out_i[N_i] (fp32) = sum over the rows of in_i [rows_i, N_i] (bf16, pitch ld) for up to 40
 * problems in ONE launch (every bias gradient of one backward pass). Same workspace / counter contract, sized by
 * gg_colsum_group_workspace_bytes(sum_i N_i). Deterministic. */
typedef struct gg_colsum_item {
  const void* in; int64_t ld; int64_t rows; int32_t N; float* out;
} gg_colsum_item;
int64_t gg_colsum_group_workspace_bytes(int64_t sum_columns);
int gg_colsum_group(const gg_colsum_item* items, int n, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------- evaluation metrics --
 * The sample-quality metrics fit() computes every 50 epochs on the generated profiles. fp32 row-major tensors,
 * fp32 CUDA-core arithmetic (nearest-neighbour comparisons do not survive bf16 operands); see
 * gemmgan_b200/csrc/evalmetrics.cu and the host mirror gemmgan_b200/evalmetrics.py.
 *
 * out[i, j] (pitch ldo) = distance between x[i, 0:d] and y[j, 0:d]:
 *   GG_DIST_L1   sum |a-b|      sklearn pairwise_distances(metric='l1') of compute_pairwise_distance
 *                               (src/distribution_distances.py:51-66)
 *   GG_DIST_SQL2 sum (a-b)^2    batch_pairwise_distances (src/unsupervised_metrics.py:114-138; the reference's
 *                               |u|^2 - 2uv + |v|^2 clamped at 0 -- formed here from the differences, which is the
 *                               same number without the cancellation)
 *   GG_DIST_L2   sqrt of it     (a[:, None] - b).pow(2).sum(2).sqrt() of dcr / nndr (src/privacy_evaluator.py:23-24,
 *                               :48, :55) without the [128, N, G] intermediate */
#define GG_DIST_L1 0
#define GG_DIST_SQL2 1
#define GG_DIST_L2 2
int gg_pairwise_distance(const float* x, int64_t ldx, const float* y, int64_t ldy, int32_t n, int32_t m, int32_t d,
                         int32_t metric, float* out, int64_t ldo, void* stream);
/* kth[i] = sorted(dist[i, 0:m])[k] (0-based rank, duplicates counted): get_kth_value(dist, k + 1)
 * (src/distribution_distances.py:69-83), np.partition(dist, seq)[:, k] (src/unsupervised_metrics.py:186-187), and
 * with k = 0 / 1 the first / second neighbour of dcr / nndr (src/privacy_evaluator.py:23-24, :49-52).
 * argmin (optional) [n] = first index of the row minimum (np.argmin, src/unsupervised_metrics.py:236-237). */
int gg_row_kth_smallest(const float* dist, int64_t ld, int32_t n, int32_t m, int32_t k, float* kth, int32_t* argmin,
                        void* stream);
/* Per row i of dist [n, m] (outputs may be NULL):
 *   row_any[i]    = any_j dist[i,j] < col_radius[j] (inclusive: <=)  -- recall of compute_prdc
 *                   (src/distribution_distances.py:124-127), np.any(distance <= D) of ManifoldEstimator.evaluate
 *                   (src/unsupervised_metrics.py:229-232)
 *   row_min[i], row_argmin[i] = min_j dist[i,j] and its first index -- coverage (:134-137), nearest_indices (:236)
 *   row_ratio[i]  = max_j col_radius[j] / (dist[i,j] + eps)          -- realism score (:234-235) */
int gg_row_membership(const float* dist, int64_t ld, int32_t n, int32_t m, const float* col_radius, int32_t inclusive,
                      float eps, uint8_t* row_any, float* row_min, int32_t* row_argmin, float* row_ratio, void* stream);
/* col_hits[j] += #{ i < n : dist[i,j] < row_radius[i] } (inclusive: <=): precision (.any(axis=0)) and density
 * (.sum(axis=0)) of compute_prdc (src/distribution_distances.py:119-132). Accumulates, so that the caller can feed
 * the rows in chunks; zero col_hits first. */
int gg_col_hits(const float* dist, int64_t ld, int32_t n, int32_t m, const float* row_radius, int32_t inclusive,
                int32_t* col_hits, void* stream);
/* out[s, c] = (x[s, c] - mean_c) / std_c over the n rows (population std); constant columns give 0
 * (standardize() inside pearson_correlation, src/corr_score.py:55-61). */
int gg_standardize_columns(const float* x, int64_t ldx, int32_t n, int32_t g, float* out, int64_t ldo, void* stream);
/* out[i, j] (pitch ldo) = sum_s a[s, i] * b[s, j] / n for standardised a [n, ga], b [n, gb]:
 * np.dot(x_.T, y_) / x.shape[0] of pearson_correlation (src/corr_score.py:63-68). */
int gg_gene_correlation(const float* a, int64_t lda, const float* b, int64_t ldb, int32_t n, int32_t ga, int32_t gb,
                        float* out, int64_t ldo, void* stream);
/* Fused gamma coefficient (gamma_coef / gamma_coeff_score, src/corr_score.py:71-120): over all gene pairs i < j with
 * cx = corr of genes i, j in xs [nx, g] and cy = the same in ys [ny, g] (both standardised), sums[0..5] (fp64,
 * device) = count, sum cx, sum cy, sum cx^2, sum cy^2, sum cx*cy. The Pearson correlation of the two
 * upper_diag_list()s of 1 - corr follows on the host (1 - c flips both lists, so the sign survives). Neither
 * [g, g] matrix is written. workspace: gg_gamma_moments_workspace_bytes(g) bytes. Deterministic. */
int64_t gg_gamma_moments_workspace_bytes(int32_t g);
int gg_gamma_moments(const float* xs, int64_t ldx, int32_t nx, const float* ys, int64_t ldy, int32_t ny, int32_t g,
                     void* workspace, int64_t workspace_bytes, double* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEMMGAN_H */
