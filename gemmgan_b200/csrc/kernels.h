// Internal launch API of the non-GEMM kernels (all enqueue on `st`, return GG_OK / error code).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gemmgan.h"

namespace gg {

typedef __nv_bfloat16 bf16;

int gemm_dispatch(const gg_gemm_desc* d, cudaStream_t st);

// ---- elementwise.cu -------------------------------------------------------------------------
// dst[r, 0:cols] (bf16, pitch ld_dst) = src[r, 0:cols] (fp32, pitch ld_src)
int k_cast_f32_bf16(const float* src, int64_t ld_src, bf16* dst, int64_t ld_dst, int64_t rows, int cols,
                    cudaStream_t st);
// key-padding mask with a never-padded CLS column in front: out[b, 0] = 0, out[b, 1+j] = in[b, j]
int k_mask_with_cls(const uint8_t* in, uint8_t* out, int B, int P, cudaStream_t st);
// mod[b, j, k] = gamma[b, k] * patches[b, j, k] + beta[b, k]; gb = [gamma | beta] fp32 [B, 2*Dp]
int k_film_apply(const bf16* patches, const float* gb, bf16* mod, int B, int P, int Dp, cudaStream_t st);
// dgb[b, k] = (sum_j dmod*patch) * (1 - gamma^2); dgb[b, Dp+k] = (sum_j dmod) * (|beta| < 5)
int k_film_bwd(const bf16* dmod, const bf16* patches, const float* gb, bf16* dgb, int B, int P, int Dp,
               cudaStream_t st);
// x[r, b, 0, :] = cls; x[r, b, 1+j, :] = x[0, b, 1+j, :] for r >= 1 (replica 0 rows 1.. are already there)
// src (optional) [B, S-1, E]: token rows of EVERY replica are taken from it instead
int k_assemble_tokens(bf16* x, const float* cls, int R, int B, int S, int E, cudaStream_t st,
                      const bf16* src = nullptr);
// out[i] = h[i] > 0 ? g[i] : 0   (ReLU backward from the stored output)
int k_relu_bwd(const bf16* g, const bf16* h, bf16* out, int64_t n, cudaStream_t st);
// dpe[b, j, :] = sum_r dx[r, b, 1+j, :]; dcls[:] = sum_{r,b} dx[r, b, 0, :]  (fp32 out for dcls)
int k_unassemble_tokens(const bf16* dx, bf16* dpe, float* dcls, int R, int B, int S, int E, cudaStream_t st);
// out[i] = sum_r in[r * n + i]
int k_sum_replicas(const bf16* in, bf16* out, int R, int64_t n, cudaStream_t st);
// out[r * n + i] = in[r * n + i] + add[i] for r < R (add: fp32, shared by the replicas)
int k_add_bcast_replicas(const bf16* in, const float* add, bf16* out, int R, int64_t n, cudaStream_t st);
// dst[b * stride_rows, :] += src[b, :]   (bf16, E columns)
int k_scatter_add_rows(bf16* dst, const bf16* src, int B, int stride_rows, int E, cudaStream_t st);
// dst[b, 0, :] = src[b, :], other token rows zero: dst [B, S, E]
int k_scatter_cls(bf16* dst, const bf16* src, int B, int S, int E, cudaStream_t st);
// out[n] (fp32) = scale * sum_rows w[row] * in[row, n]; in bf16 or fp32; w may be NULL (=1). Deterministic.
int k_colsum(const void* in, int in_f32, int64_t ld, int64_t rows, int N, const float* roww, float scale,
             float* out, int accumulate, float* scratch, cudaStream_t st);

// All bias gradients of one backward pass in one launch (deterministic; see elementwise.cu).
constexpr int COLSUM_GROUP_MAX = 40;
constexpr int COLSUM_GROUP_MAX_CHUNKS = 16;
typedef gg_colsum_item ColsumItem;
constexpr int64_t GROUP_COUNTER_BYTES = 64 * 1024;  // zeroed arrival counters at the head of a group workspace
int64_t colsum_group_workspace_bytes(int64_t max_total_columns);
int k_colsum_group(const ColsumItem* items, int n, void* workspace, int64_t workspace_bytes, cudaStream_t st);

// out[b, d] = mean over rows p with pad[b, p] == 0 of x[b, p, d]   (fp32 in / out; pad may be NULL)
// label-conditioned baseline: c[b] = [emb0[y0[b]] | emb1[y1[b]]] (bf16), and its gradient (deterministic row sums)
int k_embed_gather(const float* emb0, const float* emb1, const int64_t* y0, const int64_t* y1, int V0, int V1, bf16* c,
                   int B, int Eh, cudaStream_t st);
int k_embed_grad(const bf16* dc, const int64_t* y0, const int64_t* y1, int V0, int V1, float* g0, float* g1, int B,
                 int Eh, cudaStream_t st);
int k_masked_mean_rows(const float* x, const uint8_t* pad, float* out, int B, int P, int D, cudaStream_t st);
// dst[r, :] = index[r] >= 0 ? src[index[r], :] : 0   (fp32 rows; device-side batch assembly)
int k_gather_rows(const float* src, int64_t ld_src, const int64_t* index, float* dst, int64_t ld_dst, int64_t rows,
                  int cols, cudaStream_t st);

// BatchNorm1d over the rows of x [B, E] (conditional_gan_attention.py:108, :126); training: batch statistics + running update
int k_bn_fwd(const bf16* x, int64_t ldx, const float* gamma, const float* beta, float* run_mean, float* run_var,
             float momentum, float eps, int training, bf16* y, int64_t ldy, float* mean_out, float* rstd_out, int B, int E,
             cudaStream_t st);
int k_bn_bwd(const bf16* dy, int64_t lddy, const bf16* x, int64_t ldx, const float* mean, const float* rstd,
             const float* gamma, bf16* dx, int64_t lddx, float* dgamma, float* dbeta, int B, int E, cudaStream_t st);

// ---- xw_f32.cu: critic layer 1 on the two fp32 [B, K] gene matrices of the gradient penalty, read in place (fp32 -> bf16
// on chip, weight k-blocks TMA-multicast across a cluster): out [2B, 256] fp32
int k_xw_f32(const float* x0, const float* x1, int B, int K, const bf16* w, int64_t ldw, float* out, void* workspace,
             int64_t workspace_bytes, cudaStream_t st);

// ---- film_patch.cu: FiLM modulation in the A-operand path of the patch-encoder GEMM + bias + CLS rows + replica copies
int k_film_patch(const bf16* patches, const float* gb, const bf16* w, int64_t ldw, const float* bias, const float* cls, bf16* x0,
                 bf16* mod, int B, int P, int R, int Dp, cudaStream_t st);

// ---- gemm_ln.cu: z = res + dropout(a w^T + bias), out = LayerNorm(z) (256 columns) in one tcgen05 kernel
int k_gemm_ln(const bf16* a, int64_t lda, const bf16* w, int64_t ldw, int K, const float* bias, const bf16* res,
              const float* gamma, const float* beta, bf16* z, bf16* out, float* mean, float* rstd, int64_t rows, float eps,
              float drop_p, const uint64_t* rng, uint32_t site, cudaStream_t st);

// live profiler channel of the grouped weight-gradient kernel (gemm.cu; no-ops unless gg_gemm_profile_begin is active)
int prof_wgrad_begin(double flops, double bytes, cudaStream_t stream);
int prof_wgrad_end(cudaStream_t stream);

// ---- enc_layer.cu: one encoder layer forward as one tcgen05 kernel (S <= 16 tokens, E = 256, ffn = 512, 4 heads)
typedef gg_enc_layer_params EncLayerParams;
int k_enc_layer_fwd(const EncLayerParams& p, cudaStream_t st);
typedef gg_enc_ffn_bwd_params EncFfnBwdParams;
int k_enc_ffn_bwd(const EncFfnBwdParams& p, cudaStream_t st);

// ---- wgrad_group.cu: all single-segment weight gradients dW = dY^T X of one backward pass in one launch
constexpr int WGRAD_GROUP_MAX = 32;
constexpr int WGRAD_GROUP_MAX_SPLITS = 4;
typedef gg_wgrad_item WgradItem;
int64_t wgrad_group_workspace_bytes(int64_t max_output_elems);
int k_wgrad_group(const WgradItem* items, int n, void* workspace, int64_t workspace_bytes, cudaStream_t st);

// ---- trunk / gradient-penalty glue (elementwise.cu) ------------------------------------------
// First critic layer after the big GEMM. a1x [nx*B, H] fp32 holds x*W1x^T for the fake (and real) rows;
// a1c [R*B, H] fp32 (or NULL) holds c*W1c^T; bias b1 [H]. Writes h1 [npass*B, H] bf16 =
// leaky(a1x-mix + a1c + b1) for passes {fake, real, interp}; interp mixes alpha*real+(1-alpha)*fake.
int k_trunk1_combine(const float* a1x, const float* a1c, const float* b1, const float* alpha, bf16* h1,
                     int B, int H, int npass, int R, float slope, cudaStream_t st);
// score[m] = h2f[m, :] . w3 + b3
int k_rowdot_bias(const float* h2f, const float* w3, const float* b3, float* score, int rows, int H,
                  cudaStream_t st);
// u2[b, o] = (h2i[b, o] > 0 ? 1 : slope) * w3[o]      (bf16 out)
int k_gp_u2(const bf16* h2i, const float* w3, bf16* u2, int B, int H, float slope, cudaStream_t st);
// per row: n = sqrt(sum_h y*u1); r = gpw*(2/Bglobal)*(1-1/n); ru1 = r*u1 ; dv1 = m1 .* (r*y) ; pen[b] = (n-1)^2
int k_gp_rows(const float* y, const float* u1f, const bf16* h1i, float* norms, float* pen, bf16* ru1,
              bf16* dv1, int B, int H, float slope, float gp_weight, float inv_batch, cudaStream_t st);
// da2[m, o] = sign(m)/B * w3[o] * (h2[m, o] > 0 ? 1 : slope); rows [0,B): sign_fake, rows [B,2B): sign_real
int k_score_bwd(const bf16* h2, const float* w3, bf16* da2, float* roww, int rows, int B, int H, float slope,
                float sign_first, float sign_second, float inv_batch, cudaStream_t st);
int k_score_bwd_rows(const bf16* h2, const float* w3, const float* drow, bf16* da2, int rows, int H, float slope,
                     cudaStream_t st);
// stats: [0]=loss_real=-mean(score_real) [1]=loss_fake=mean(score_fake) [2]=gp=mean(pen) [3]=d_loss total
int k_disc_losses(const float* score, const float* pen, float* stats, int B, float gp_weight, float inv_batch,
                  cudaStream_t st);
int k_gen_loss(const float* score, float* stats, int B, float inv_batch, cudaStream_t st);
int k_fill_f32(float* p, float v, int64_t n, cudaStream_t st);
int k_bump_rng(uint64_t* rng, cudaStream_t st);

// ---- layernorm.cu ---------------------------------------------------------------------------
// z = x + dropout(y); out = LayerNorm(z) * w + b. One warp per row; E % 32 == 0, E <= 1024.
int k_add_ln_fwd(const bf16* x, const bf16* y, const float* w, const float* b, bf16* z, bf16* out,
                 float* mean, float* rstd, int64_t rows, int E, float eps, float drop_p, const uint64_t* rng,
                 uint32_t site, cudaStream_t st);
// dz = LN backward of dout (grad to the residual input); dy = dropout-mask(dz) (grad to the sublayer output,
// may be NULL when drop_p == 0); dw/db fp32 [E] (deterministic two-stage; scratch >= 2*E*nblocks floats)
int k_add_ln_bwd(const bf16* dout, const bf16* z, const float* mean, const float* rstd, const float* w,
                 bf16* dz, bf16* dy, float* dw, float* db, int64_t rows, int E, float drop_p,
                 const uint64_t* rng, uint32_t site, float* scratch, cudaStream_t st);
int k_ln_bwd_finish(const float* scratch, int64_t rows, int E, float* dw, float* db, cudaStream_t st);
int64_t ln_bwd_scratch_floats(int64_t rows, int E);

// ---- attention.cu ---------------------------------------------------------------------------
// public mirror: gg_attn_args (include/gemmgan.h); the internal view types the pointers
struct AttnArgs {
  const bf16* q; int64_t ldq; int q_mod;     // query row = (b % q_mod) * Lq + i
  const bf16* k; const bf16* v; int64_t ldkv; int kv_mod;   // key row = (b % kv_mod) * Lk + j
  const uint8_t* mask; int mask_mod;         // [mask_mod, Lk], 1 = padded key; may be NULL
  int nb, H, hd, Lq, Lk;
  float drop_p; const uint64_t* rng; uint32_t site;
  bf16* o; int64_t ldo;                      // forward output [nb*Lq, H*hd]
  // backward
  const bf16* dout; int64_t lddo;            // [nb*Lq, H*hd]
  bf16* dq; int64_t lddq;                    // [nb*Lq, H*hd]   (per replica even if q is shared)
  bf16* dk; bf16* dv; int64_t lddkv;         // [nb*Lk, ...]    (per replica even if kv is shared)
  float* stat;                               // backward scratch: 2 * nb * H * Lq floats (lse, delta)
  const uint32_t* dbits;                     // optional precomputed keep bits of the site (k_dropout_bits), or NULL
};
static_assert(sizeof(AttnArgs) == sizeof(gg_attn_args), "AttnArgs must mirror gg_attn_args");
int k_attention_fwd(const AttnArgs& a, cudaStream_t st);
int k_attention_bwd(const AttnArgs& a, cudaStream_t st);
// keep bits of a dropout site drawn once (AttnArgs::dbits); words = dropout_bits_words(n_elems)
int64_t dropout_bits_words(int64_t n_elems);
int k_dropout_bits(const uint64_t* rng, uint32_t site, float p, int64_t n_elems, uint32_t* out, cudaStream_t st);
int k_dropout_bits3(const uint64_t* rng, float p, const uint32_t (&site)[3], const int64_t (&n_elems)[3], uint32_t* const (&out)[3],
                    cudaStream_t st);

// ---- optim.cu -------------------------------------------------------------------------------
// total = sqrt(sum g^2) over n elements -> norm_out[0]; clip coefficient -> norm_out[1]
int k_grad_norm_clip(const float* g, int64_t n, float max_norm, float* norm_out, float* scratch,
                     cudaStream_t st);
// One fused pass over the flat buffers: g *= coef (if coef_ptr), optimizer update, step counter in state.
int k_optim_step(int kind, float* p, float* g, float* m, float* v, int64_t n, float lr, const float* coef_ptr,
                 float* step_count, cudaStream_t st, bool bump_step = true);
struct ShadowSeg {            // fp32 parameter sub-matrix -> bf16 shadow (padded pitch)
  int64_t p_off;              // offset of the parameter tensor in the flat buffer
  int32_t rows, cols;         // parameter shape [rows, cols]
  int32_t col0, ncols;        // column range copied
  int64_t s_off;              // offset (elements) in the shadow buffer
  int64_t s_ld;
  int32_t transpose;          // 1: the shadow holds the TRANSPOSE ([ncols, rows], pitch s_ld >= rows): K-major operand of
  int32_t pad_;               //    the dgrad products of the fused backward kernels
};
// bump_step (optional): step counter incremented by this launch (the engine's optimizer step folds it in here)
int k_refresh_shadows(const float* p, bf16* shadow, const ShadowSeg* segs_dev, int nseg, int max_rows,
                      cudaStream_t st, float* bump_step = nullptr);

}  // namespace gg
