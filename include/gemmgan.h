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

#define GG_ABI_VERSION 1

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
  void* out_bf16;        /* [*, ld_bf16] bf16 output or NULL */
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
} gg_gemm_desc;

int gg_gemm_bf16(const gg_gemm_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEMMGAN_H */
