// Host build of gemmgan_b200/csrc/elementwise.cu (see emu.h): the HBM-bound glue kernels of the step (casts, FiLM,
// token assembly, trunk-1 combine, gradient-penalty rows, loss reductions, column sums, embedding gather / gradient,
// masked patch mean). The internal launch API of kernels.h is exported as it is (C++ linkage would mangle it, so thin
// extern "C" wrappers named emu_<function> forward to gg::k_<function>).
#include "emu.h"

#include "../../gemmgan_b200/csrc/elementwise.cu"

using gg::bf16;
#define W(name, params, call) extern "C" int emu_##name params { return gg::k_##name call; }

W(cast_f32_bf16, (const float* src, int64_t ld_src, bf16* dst, int64_t ld_dst, int64_t rows, int cols),
  (src, ld_src, dst, ld_dst, rows, cols, nullptr))
W(mask_with_cls, (const uint8_t* in, uint8_t* out, int B, int P), (in, out, B, P, nullptr))
W(film_apply, (const bf16* patches, const float* gb, bf16* mod, int B, int P, int Dp), (patches, gb, mod, B, P, Dp, nullptr))
W(film_bwd, (const bf16* dmod, const bf16* patches, const float* gb, bf16* dgb, int B, int P, int Dp),
  (dmod, patches, gb, dgb, B, P, Dp, nullptr))
W(assemble_tokens, (bf16* x, const float* cls, int R, int B, int S, int E, const bf16* src), (x, cls, R, B, S, E, nullptr, src))
W(unassemble_tokens, (const bf16* dx, bf16* dpe, int R, int B, int S, int E), (dx, dpe, nullptr, R, B, S, E, nullptr))
W(relu_bwd, (const bf16* g, const bf16* h, bf16* out, int64_t n), (g, h, out, n, nullptr))
W(sum_replicas, (const bf16* in, bf16* out, int R, int64_t n), (in, out, R, n, nullptr))
W(colsum, (const void* in, int in_f32, int64_t ld, int64_t rows, int N, const float* roww, float scale, float* out,
           int accumulate, float* scratch),
  (in, in_f32, ld, rows, N, roww, scale, out, accumulate, scratch, nullptr))
W(colsum_group, (const gg::ColsumItem* items, int n, void* workspace, int64_t workspace_bytes),
  (items, n, workspace, workspace_bytes, nullptr))
extern "C" int64_t emu_colsum_group_workspace_bytes(int64_t cols) { return gg::colsum_group_workspace_bytes(cols); }
W(trunk1_combine, (const float* a1x, const float* a1c, const float* b1, const float* alpha, bf16* h1, int B, int H,
                   int npass, int R, float slope),
  (a1x, a1c, b1, alpha, h1, B, H, npass, R, slope, nullptr))
W(rowdot_bias, (const float* h2f, const float* w3, const float* b3, float* score, int rows, int H),
  (h2f, w3, b3, score, rows, H, nullptr))
W(gp_u2, (const bf16* h2i, const float* w3, bf16* u2, int B, int H, float slope), (h2i, w3, u2, B, H, slope, nullptr))
W(gp_rows, (const float* y, const float* u1f, const bf16* h1i, float* norms, float* pen, bf16* ru1, bf16* dv1, int B,
            int H, float slope, float gp_weight, float inv_batch),
  (y, u1f, h1i, norms, pen, ru1, dv1, B, H, slope, gp_weight, inv_batch, nullptr))
W(score_bwd, (const bf16* h2, const float* w3, bf16* da2, float* roww, int rows, int B, int H, float slope, float s0,
              float s1, float inv_batch),
  (h2, w3, da2, roww, rows, B, H, slope, s0, s1, inv_batch, nullptr))
W(disc_losses, (const float* score, const float* pen, float* stats, int B, float gp_weight, float inv_batch),
  (score, pen, stats, B, gp_weight, inv_batch, nullptr))
W(gen_loss, (const float* score, float* stats, int B, float inv_batch), (score, stats, B, inv_batch, nullptr))
W(embed_gather, (const float* e0, const float* e1, const int64_t* y0, const int64_t* y1, int V0, int V1, bf16* c, int B,
                 int Eh),
  (e0, e1, y0, y1, V0, V1, c, B, Eh, nullptr))
W(embed_grad, (const bf16* dc, const int64_t* y0, const int64_t* y1, int V0, int V1, float* g0, float* g1, int B, int Eh),
  (dc, y0, y1, V0, V1, g0, g1, B, Eh, nullptr))
W(masked_mean_rows, (const float* x, const uint8_t* pad, float* out, int B, int P, int D), (x, pad, out, B, P, D, nullptr))
W(gather_rows, (const float* src, int64_t ld_src, const int64_t* index, float* dst, int64_t ld_dst, int64_t rows, int cols),
  (src, ld_src, index, dst, ld_dst, rows, cols, nullptr))
W(bn_fwd, (const bf16* x, int64_t ldx, const float* gamma, const float* beta, float* run_mean, float* run_var, float momentum,
           float eps, int training, bf16* y, int64_t ldy, float* mean_out, float* rstd_out, int B, int E),
  (x, ldx, gamma, beta, run_mean, run_var, momentum, eps, training, y, ldy, mean_out, rstd_out, B, E, nullptr))
W(bn_bwd, (const bf16* dy, int64_t lddy, const bf16* x, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
           bf16* dx, int64_t lddx, float* dgamma, float* dbeta, int B, int E),
  (dy, lddy, x, ldx, mean, rstd, gamma, dx, lddx, dgamma, dbeta, B, E, nullptr))
