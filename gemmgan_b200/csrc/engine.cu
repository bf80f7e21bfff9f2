// gg_engine: the WGAN-GP training step of GeMM-GAN as a fixed kernel sequence.
//
// Restructuring relative to the reference's autograd execution (see DESIGN.md for the derivations):
//  * critic layer 1 is linear in the gene vector, so the interpolated pass re-uses the fake/real GEMM
//    results (a1x_interp = alpha*a1x_real + (1-alpha)*a1x_fake): no third [B,G] GEMM and no interpolated
//    [B,G] tensor (reference :358-360);
//  * gradient penalty via the Gram matrix M = W1x W1x^T: ||grad_b||^2 = u1_b M u1_b^T, and its W1 gradient
//    (U1^T diag(r) U1) W1x rides as a second K-segment on the W1 weight-gradient GEMM (reference :351-374
//    + the double backward inside :412);
//  * the three critic tower passes of train_disc (fake / real / interpolated, :403,:404,:360) differ only by
//    their dropout draws: they run as one batch of 3B rows (or one pass of B rows when dropout is off), and
//    only the fake and real replicas are back-propagated (the GP's tower gradient is identically zero);
//  * torch.cat((x, c)) (:157,:226) is a second K-segment, never materialised.
#include "host_util.h"
#include "kernels.h"

#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

namespace gg {

#define GG_TRY(x)            \
  do {                       \
    int _rc = (x);           \
    if (_rc) return _rc;     \
  } while (0)

struct Op {
  const bf16* p = nullptr;
  int64_t ld = 0;
};

struct Epi {
  gg_epilogue e;
  Epi() {
    memset(&e, 0, sizeof(e));
    e.alpha = 1.f;
    e.mask_pos = 1.f;
  }
  Epi& bias(const float* b) { e.bias = b; return *this; }
  Epi& act(int a, float s = 0.f) { e.act = a; e.slope = s; return *this; }
  Epi& drop(float p, const uint64_t* rng, uint32_t site) {
    if (p > 0.f) { e.drop_p = p; e.rng = rng; e.site = site; }
    return *this;
  }
  Epi& mask(const void* m, int64_t ld, float pos, float neg) {
    e.mask = m; e.mask_ld = ld; e.mask_f32 = 0; e.mask_pos = pos; e.mask_neg = neg; return *this;
  }
  Epi& res(const bf16* r, int64_t ld) { e.res = r; e.res_ld = ld; e.res_f32 = 0; return *this; }
  Epi& obf(bf16* o, int64_t ld) { e.out_bf16 = o; e.ld_bf16 = ld; return *this; }
  Epi& of32(float* o, int64_t ld) { e.out_f32 = o; e.ld_f32 = ld; return *this; }
  Epi& rowmap(int div, int mul, int add) { e.row_div = div; e.row_mul = mul; e.row_add = add; return *this; }
};

struct Arena {
  uint8_t* base;
  int64_t off = 0;
  explicit Arena(uint8_t* b) : base(b) {}
  template <class T>
  T* take(int64_t n) {
    off = round_up64(off, 256);
    T* p = reinterpret_cast<T*>(base + off);  // base may be null during the sizing pass
    off += n * static_cast<int64_t>(sizeof(T));
    return p;
  }
};

struct WShadow {
  bf16* p = nullptr;
  int64_t ld = 0;
};
struct NetShadow {
  WShadow w[GG_NSLOTS];
  WShadow wt[GG_NSLOTS];  // transposed copies (ffn weights of the encoder layers: dgrad operands of enc_ffn_bwd)
  WShadow tr0_c;  // conditioning column block of the first trunk layer
  bf16* base = nullptr;
  int64_t elems = 0;
  std::vector<ShadowSeg> segs;
  ShadowSeg* segs_dev = nullptr;
};

struct TowerLayer {
  uint32_t* dbits = nullptr;  // keep bits of the attention-probability dropout of this layer pass (17 .. 320 tokens)
  uint32_t* lbits[3] = {nullptr, nullptr, nullptr};  // fused layer (<= 16 tokens): keep bits of sites + 1, + 2, + 3
  bf16 *qkv, *ao, *z1, *x1, *h, *z2;
  float *mean1, *rstd1, *mean2, *rstd2;
};
struct Tower {
  int Rmax = 1;
  float* gb;
  bf16 *mod, *te, *X[3];
  TowerLayer L[2];
  bf16 *qp, *kvp, *ap, *pv, *qt, *kvt, *at, *c, *tmpE;
  float* ot = nullptr;  // T == 1: out-projection of the single text token's value vector [B, E] (shared by the replicas)
  bf16 *pe_h = nullptr, *pe_ln = nullptr, *zero = nullptr;  // IMG: relu(patch projection), its LayerNorm, zeros
  float *pe_mean = nullptr, *pe_rstd = nullptr;
};
struct LayerGrads {  // per layer: side-lane weight-gradient GEMMs read these while lane 0 moves on
  bf16 *gz2, *gy2, *gz1, *gy1, *gh, *gqkv;
  float *lnp2, *lnp1;  // LayerNorm weight/bias-gradient block partials (finished on a side lane)
};
struct GradScratch {
  bf16 *dc, *dat, *dqt, *dkvt, *dkvt_sum, *dp, *dap, *dqp, *dqp_sum, *dkvp;
  bf16* dcs = nullptr;  // T == 1: dc summed over the replicas [B, E]
  bf16 *ga, *gb, *gao, *dte, *dte0, *dpe, *dmod, *dgb;
  bf16 *dpe_z = nullptr, *dpe_h = nullptr;  // IMG: gradient before the patch LayerNorm / before the ReLU
  float* lnp_pe = nullptr;
  LayerGrads L[2];
};
struct TrunkBufs {
  float *a1x, *a1c, *h2f, *score, *Mg, *u1f, *y, *norms, *pen, *du2f, *roww;
  bf16 *h1, *h2, *Mgb, *u2, *u1b, *ru1, *dv1, *Qb, *da2, *da1;
  bf16 *hg1, *hg2, *dfake, *dag2, *dag1;
};

}  // namespace gg

using namespace gg;

struct gg_engine {
  gg_model_cfg cfg;
  gg_net_buffers nets[2];
  NetShadow sh[2];
  int S_ = 1, Gp = 0, F = 0, hd = 0;
  bool img = false;     // conditional_gan_img_transformer.py: patch encoder = Linear + ReLU + LayerNorm, CLS conditioning
  bool concat = false;  // conditional_gan_concat.py: the conditioning is one Linear of the staged text / mean-patch vector
  bool label = false;   // benchmark_generative_model.py: the conditioning is two gathered embedding rows (a concat-style
                        // engine: `concat` is set too; the staged vector / encoder GEMM become labels / a gather)
  int64_t* labels = nullptr;  // [2, B]
  // conditional_gan_attention.py: c = single-query MHA(text vector -> patches) [+ BatchNorm1d in the generator]
  bool attn = false;
  float *bn_run_mean = nullptr, *bn_run_var = nullptr;  // caller-owned running statistics (gg_engine_set_batchnorm)
  float bn_momentum = 0.1f, bn_eps = 1e-5f;
  float *bn_mean = nullptr, *bn_rstd = nullptr;  // statistics the last generator forward normalised with
  int bn_training = 1;
  float kept_p[2] = {-1.f, -1.f};  // dropout probability of the last *_keep forward per net (< 0: none yet)
  int kept_training[2] = {0, 0};
  bool cond = false, paper = false, film = false;  // paper: cross-attention tail; film: FiLM modulation of the patches
  // staged inputs
  bf16 *xfr, *patches, *text, *zbf, *xin;
  uint8_t *mask_s, *tpad;
  bool has_tpad = false, has_ppad = false;
  Tower tw[2];
  GradScratch gs;
  TrunkBufs tb;
  float *stats, *opt_step[2], *normbuf, *attn_stat;
  uint64_t* rng;
  // Lanes: lane 0 is the caller's stream (the dependent chain: forwards, dgrads, attention / LayerNorm
  // backward); lanes 1.. are engine-owned streams for work nothing on the chain waits for (weight- and
  // bias-gradient GEMMs / reductions, the generator forward next to the critic tower forward). Each lane
  // has its own split-K workspace and reduction scratch. Forks and joins are event edges, so the whole
  // step stays capturable into one CUDA graph; every entry point joins all lanes before it returns.
  static constexpr int NLANES = 3;
  static constexpr int NEVENTS = 16;
  cudaStream_t cur[NLANES] = {nullptr, nullptr, nullptr};
  cudaEvent_t evs[NEVENTS];
  int ev_next = 0;
  bool forked[NLANES] = {false, false, false};
  bool multi_lane = true;
  float* scratch_l[NLANES];
  void* splitk_l[NLANES];
  int64_t splitk_bytes = 0;
  int64_t total_bytes = 0;

  int L(int lane) const { return multi_lane ? lane : 0; }
  cudaStream_t S(int lane) const { return cur[L(lane)]; }
  // lane `lane` waits for everything enqueued on lane 0 so far
  int fork(int lane) {
    lane = L(lane);
    if (lane == 0) return GG_OK;
    cudaEvent_t ev = evs[ev_next];
    ev_next = (ev_next + 1) % NEVENTS;
    GG_CUDA_CHECK(cudaEventRecord(ev, cur[0]));
    GG_CUDA_CHECK(cudaStreamWaitEvent(cur[lane], ev, 0));
    forked[lane] = true;
    return GG_OK;
  }
  // lane 0 waits for everything enqueued on `lane`
  int join(int lane) {
    lane = L(lane);
    if (lane == 0 || !forked[lane]) return GG_OK;
    cudaEvent_t ev = evs[ev_next];
    ev_next = (ev_next + 1) % NEVENTS;
    GG_CUDA_CHECK(cudaEventRecord(ev, cur[lane]));
    GG_CUDA_CHECK(cudaStreamWaitEvent(cur[0], ev, 0));
    forked[lane] = false;
    return GG_OK;
  }
  // lane `dst` waits for everything enqueued on lane `src` so far (neither is lane 0's bookkeeping)
  int wait_lane(int dst, int src) {
    dst = L(dst);
    src = L(src);
    if (dst == src) return GG_OK;
    if (src != 0 && !forked[src]) return GG_OK;  // nothing enqueued on `src` in this entry point
    cudaEvent_t ev = evs[ev_next];
    ev_next = (ev_next + 1) % NEVENTS;
    GG_CUDA_CHECK(cudaEventRecord(ev, cur[src]));
    GG_CUDA_CHECK(cudaStreamWaitEvent(cur[dst], ev, 0));
    if (dst != 0) forked[dst] = true;
    return GG_OK;
  }
  // the two halves of wait_lane apart: an event recorded on `src` now, waited for by `dst` later
  int record_lane(int src, cudaEvent_t* ev_out) {
    src = L(src);
    *ev_out = nullptr;
    if (src != 0 && !forked[src]) return GG_OK;
    cudaEvent_t ev = evs[ev_next];
    ev_next = (ev_next + 1) % NEVENTS;
    GG_CUDA_CHECK(cudaEventRecord(ev, cur[src]));
    *ev_out = ev;
    return GG_OK;
  }
  int wait_event(int dst, cudaEvent_t ev) {
    if (!ev) return GG_OK;
    dst = L(dst);
    GG_CUDA_CHECK(cudaStreamWaitEvent(cur[dst], ev, 0));
    if (dst != 0) forked[dst] = true;
    return GG_OK;
  }
  int join_all() {
    for (int l = 1; l < NLANES; ++l) {
      int rc = join(l);
      if (rc) return rc;
    }
    return GG_OK;
  }
  void begin(void* stream) { cur[0] = reinterpret_cast<cudaStream_t>(stream); }

  float* P(int net, int slot) const { return nets[net].off[slot] < 0 ? nullptr : nets[net].params + nets[net].off[slot]; }
  float* Gr(int net, int slot) const { return nets[net].off[slot] < 0 ? nullptr : nets[net].grads + nets[net].off[slot]; }
  Op W(int net, int slot) const { return Op{sh[net].w[slot].p, sh[net].w[slot].ld}; }

  int mm(int lane, int M, int N, int K, Op A, int a_mn, Op B, int b_mn, const Epi& epi, int K2 = 0,
         Op A2 = Op(), Op B2 = Op()) const {
    gg_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.M = M; d.N = N;
    d.nseg = K2 > 0 ? 2 : 1;
    d.seg[0].a = A.p; d.seg[0].b = B.p; d.seg[0].lda = A.ld; d.seg[0].ldb = B.ld; d.seg[0].K = K;
    if (K2 > 0) { d.seg[1].a = A2.p; d.seg[1].b = B2.p; d.seg[1].lda = A2.ld; d.seg[1].ldb = B2.ld; d.seg[1].K = K2; }
    d.a_mn_major = a_mn; d.b_mn_major = b_mn;
    d.epi = epi.e;
    d.workspace = splitk_l[L(lane)]; d.workspace_bytes = splitk_bytes;
    d.impl = cfg.gemm_impl;
    return gemm_dispatch(&d, S(lane));
  }
  // Y = X W^T (+b): X [rows, in], W [out, in]
  int linear(int lane, int rows, int out, int in, Op X, Op Wt, const Epi& epi) const {
    return mm(lane, rows, out, in, X, 0, Wt, 0, epi);
  }
  // dX = dY W: dY [rows, out], W [out, in]
  int dgrad(int lane, int rows, int in, int out, Op dY, Op Wt, const Epi& epi) const {
    return mm(lane, rows, in, out, dY, 0, Wt, 1, epi);
  }
  // Y = X W^T with FP32 operands read in place and multiplied as TF32 (gg_gemm_desc.tf32_operands): X [rows, in]
  // (pitch ldx), W [out, in] (pitch ldw) -- the gradient-penalty microbenchmark's fp32 profiles against the fp32 master
  // weights: no bf16 copy of either
  int linear_tf32(int lane, int rows, int out, int in, const float* X, int64_t ldx, const float* Wt, int64_t ldw,
                  const Epi& epi) const {
    gg_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.M = rows; d.N = out; d.nseg = 1;
    d.seg[0].a = X; d.seg[0].b = Wt; d.seg[0].lda = ldx; d.seg[0].ldb = ldw; d.seg[0].K = in;
    d.epi = epi.e;
    d.workspace = splitk_l[L(lane)]; d.workspace_bytes = splitk_bytes;
    d.impl = cfg.gemm_impl;
    d.tf32_operands = 1;
    return gemm_dispatch(&d, S(lane));
  }
  // Weight / bias gradients feed nothing but the optimizer: they run on side lanes, ordered after
  // everything lane 0 has enqueued so far (their operands), and are joined at the end of the entry point.
  // dW[out, in] = dY^T X, fp32 into the gradient buffer (pitch ldw)
  // They are queued and launched as ONE grouped GEMM / ONE grouped column-sum kernel per flush_grads().
  int wgrad(int rows_out, int in, int rows, Op dY, Op X, float* dW, int64_t ldw) {
    if (!group_grads || cfg.gemm_impl != GG_IMPL_TCGEN05 || wq.size() >= WGRAD_GROUP_MAX) {
      GG_TRY(fork(1));
      return mm(1, rows_out, in, rows, dY, 1, X, 1, Epi().of32(dW, ldw));
    }
    wq.push_back(WgradItem{dY.p, dY.ld, X.p, X.ld, rows_out, in, rows, dW, ldw});
    return GG_OK;
  }
  int bgrad(const bf16* dY, int64_t ld, int64_t rows, int N, float* db) {
    if (!db) return GG_OK;
    if (!group_grads || bq.size() >= COLSUM_GROUP_MAX) {
      GG_TRY(fork(2));
      return k_colsum(dY, 0, ld, rows, N, nullptr, 1.f, db, 0, scratch_l[L(2)], S(2));
    }
    bq.push_back(ColsumItem{dY, ld, rows, N, db});
    return GG_OK;
  }
  // launches what has been queued: ordered after everything lane 0 has enqueued so far
  int flush_grads() {
    // a bias gradient whose dY is also the dY of a queued weight gradient is formed inside the grouped wgrad
    // kernel (ones-column MMA); only the rest (CLS token, in-proj slices without a matching GEMM) needs a column sum
    if (fuse_bias && !wq.empty()) {
      size_t keep = 0;
      for (size_t i = 0; i < bq.size(); ++i) {
        const ColsumItem& b = bq[i];
        bool fused = false;
        for (WgradItem& w : wq) {
          if (!w.bias && w.dy == b.in && w.ld_dy == b.ld && w.K == b.rows && w.M == b.N) {
            w.bias = b.out;
            fused = true;
            break;
          }
        }
        if (!fused) bq[keep++] = b;
      }
      bq.resize(keep);
    }
    if (!wq.empty()) {
      GG_TRY(fork(1));
      GG_TRY(wait_lane(1, 2));  // operands produced on lane 2 (text-side gradients)
      GG_TRY(k_wgrad_group(wq.data(), static_cast<int>(wq.size()), wg_ws, wg_ws_bytes, S(1)));
      wq.clear();
    }
    if (!bq.empty()) {
      GG_TRY(fork(2));
      GG_TRY(k_colsum_group(bq.data(), static_cast<int>(bq.size()), cs_ws, cs_ws_bytes, S(2)));
      bq.clear();
    }
    return GG_OK;
  }
  std::vector<WgradItem> wq;
  std::vector<ColsumItem> bq;
  bool group_grads = true;
  // text-side work on lane 2: measured on cfg3 (ms / train()): fwd+bwd 7.89, bwd only 7.96, fwd only 7.75, none 7.82
  bool text_lane_fwd = true, text_lane_bwd = false;  // GEMMGAN_TEXT_LANE = <fwd><bwd> digits overrides
  bool split_tail_flush = false;  // measured: splitting the last flush costs 0.28 ms / train() (the grouped kernel occupies every SM)
  bool fuse_bias = true;  // GEMMGAN_FUSE_BIAS=0: keep every bias gradient in the grouped column-sum kernel
  // GEMMGAN_LAYER_BITS=1: the fused layer kernels (<= 16 tokens) read precomputed keep bits too. MEASURED SLOWER at cfg3
  // (7.39 vs 7.00 ms per train() on the same box): the fused kernel already draws its masks while its epilogue warps
  // wait for the accumulators, so nothing leaves its critical path, and the extra mask launches take SM time from the
  // tower head. Off by default; results are bitwise the same either way (tests/test_gpu_enc_layer.py).
  bool layer_bits = false;
  // One text token (T == 1: BASELINE configs 1-3): the text2patch attention is a softmax over ONE key, i.e. exactly 1 for
  // every query, so its output is that token's value vector whatever the query is, and no gradient reaches the query or the
  // key (d softmax = 0). The engine then skips the query projection, both attention launches and their backward, and forms
  // c = pv + (v_text Wo^T + bo) with the text-only term computed once on the text lane. GEMMGAN_T1_SHORTCUT=0 runs the
  // general path (same numbers to bf16 rounding; tests compare both).
  bool t1_shortcut = true;
  bool gemm_ln = true;     // GEMMGAN_GEMM_LN=0: out-projection / linear2 and their add + LayerNorm as two launches each
  bool film_fused = true;  // GEMMGAN_FILM_FUSED=0: film_apply_kernel + patch GEMM + token assembly as three launches
  bool gp_tf32 = false;   // GEMMGAN_GP_TF32=1 (gg_engine_gp_step)
  bool attn_bits = true;  // GEMMGAN_ATTN_BITS=0: the 17 .. 320-token attention kernels draw their dropout masks themselves
  bool fused_layer = true;  // GEMMGAN_FUSED_LAYER=0: encoder layers as seven launches instead of enc_layer.cu's one
  // one workspace per flush in flight: flushes of one entry point run back to back on their lane, so
  // they rotate through GROUP_WS_SLOTS regions
  void *wg_ws = nullptr, *cs_ws = nullptr;
  int64_t wg_ws_bytes = 0, cs_ws_bytes = 0;
};

namespace gg {

static bool layer_bits_requested() {
  static const bool on = [] { const char* v = getenv("GEMMGAN_LAYER_BITS"); return v && v[0] == '1'; }();
  return on;
}

static bool slot_matrix_shape(const gg_model_cfg& c, int net, int slot, int* rows, int* cols) {
  const int E = c.E, F = c.ffn, condw = c.variant == GG_VARIANT_VANILLA ? 0 : E;
  if (c.variant == GG_VARIANT_LABEL && slot < GG_P_TR0_W) return false;  // embedding tables are gathered in fp32
  if (slot == GG_P_FILM_W) {  // (the IMG variant keeps its patch-encoder LayerNorm vectors in the FiLM slots)
    if (c.variant == GG_VARIANT_IMG || c.variant == GG_VARIANT_ATTN) return false;
    *rows = 2 * c.Dp; *cols = c.Dt; return true;
  }
  if (slot == GG_P_TEXT_W) { *rows = E; *cols = c.Dt; return true; }
  if (slot == GG_P_PATCH_W) { *rows = E; *cols = c.Dp; return true; }
  if (slot >= GG_P_LAYER0 && slot < GG_P_LAYER0 + 24) {
    const int k = (slot - GG_P_LAYER0) % GG_L_COUNT;
    if (k == GG_L_IN_W) { *rows = 3 * E; *cols = E; return true; }
    if (k == GG_L_OUT_W) { *rows = E; *cols = E; return true; }
    if (k == GG_L_FF1_W) { *rows = F; *cols = E; return true; }
    if (k == GG_L_FF2_W) { *rows = E; *cols = F; return true; }
    return false;
  }
  if (slot == GG_P_P2T_IN_W || slot == GG_P_T2P_IN_W) { *rows = 3 * E; *cols = E; return true; }
  if (slot == GG_P_P2T_OUT_W || slot == GG_P_T2P_OUT_W) { *rows = E; *cols = E; return true; }
  if (slot == GG_P_TR0_W) { *rows = c.H; *cols = (net == GG_NET_GEN ? c.L : c.G) + condw; return true; }
  if (slot == GG_P_TR1_W) { *rows = c.H; *cols = c.H; return true; }
  if (slot == GG_P_FIN_W) { *rows = net == GG_NET_GEN ? c.G : 1; *cols = c.H; return true; }
  return false;
}

static void layout_shadows(gg_engine& e, Arena& ar) {
  const gg_model_cfg& c = e.cfg;
  for (int net = 0; net < 2; ++net) {
    NetShadow& s = e.sh[net];
    s.segs.clear();
    int64_t elems = 0;
    std::vector<std::pair<int, int64_t>> where;  // (slot or -1 for tr0_c, offset)
    for (int slot = 0; slot < GG_NSLOTS; ++slot) {
      int rows, cols;
      if (e.nets[net].off[slot] < 0 || !slot_matrix_shape(c, net, slot, &rows, &cols)) continue;
      if (slot == GG_P_FIN_W && net == GG_NET_DISC) continue;  // used as an fp32 vector
      auto add = [&](int col0, int ncols, WShadow* dst, bool transpose = false) {
        const int64_t ld = round_up64(transpose ? rows : ncols, 8);
        elems = round_up64(elems, 128);
        ShadowSeg sg;
        memset(&sg, 0, sizeof(sg));
        sg.p_off = e.nets[net].off[slot];
        sg.rows = rows; sg.cols = cols; sg.col0 = col0; sg.ncols = ncols;
        sg.s_off = elems; sg.s_ld = ld;
        sg.transpose = transpose ? 1 : 0;
        s.segs.push_back(sg);
        dst->ld = ld;
        dst->p = reinterpret_cast<bf16*>(static_cast<uintptr_t>(elems));  // offset for now
        elems += static_cast<int64_t>(transpose ? ncols : rows) * ld;
      };
      if (slot == GG_P_TR0_W && c.variant != GG_VARIANT_VANILLA) {
        const int first = cols - c.E;
        add(0, first, &s.w[slot]);
        add(first, c.E, &s.tr0_c);
      } else {
        add(0, cols, &s.w[slot]);
      }
      if (slot >= GG_P_LAYER0 && slot < GG_P_LAYER0 + 24) {
        const int k = (slot - GG_P_LAYER0) % GG_L_COUNT;
        if (k == GG_L_FF1_W || k == GG_L_FF2_W) add(0, cols, &s.wt[slot], true);
      }
    }
    s.elems = elems;
    s.base = ar.take<bf16>(elems);
    for (int slot = 0; slot < GG_NSLOTS; ++slot) {
      if (s.w[slot].ld) s.w[slot].p = s.base + reinterpret_cast<uintptr_t>(s.w[slot].p);
      if (s.wt[slot].ld) s.wt[slot].p = s.base + reinterpret_cast<uintptr_t>(s.wt[slot].p);
    }

    if (s.tr0_c.ld) s.tr0_c.p = s.base + reinterpret_cast<uintptr_t>(s.tr0_c.p);
    s.segs_dev = ar.take<ShadowSeg>(static_cast<int64_t>(s.segs.size()) + 1);
  }
}

static void layout_tower(gg_engine& e, Tower& t, int Rmax, Arena& ar) {
  const gg_model_cfg& c = e.cfg;
  const int64_t B = c.B, S = e.S_, E = c.E, F = e.F, P = c.P, T = c.T;
  const int64_t rows = static_cast<int64_t>(Rmax) * B * S;
  t.Rmax = Rmax;
  t.gb = ar.take<float>(B * 2 * c.Dp);
  t.mod = ar.take<bf16>(B * P * c.Dp);
  t.te = ar.take<bf16>(B * T * E);
  for (int i = 0; i < 3; ++i) t.X[i] = ar.take<bf16>(rows * E);
  for (int l = 0; l < 2; ++l) {
    TowerLayer& L = t.L[l];
    L.qkv = ar.take<bf16>(rows * 3 * E);
    L.ao = ar.take<bf16>(rows * E);
    L.z1 = ar.take<bf16>(rows * E);
    L.x1 = ar.take<bf16>(rows * E);
    L.h = ar.take<bf16>(rows * F);
    L.z2 = ar.take<bf16>(rows * E);
    if (c.dropout_p > 0.f && S > 16 && !e.attn)
      L.dbits = ar.take<uint32_t>(dropout_bits_words(static_cast<int64_t>(Rmax) * B * c.n_heads * S * S));
    if (c.dropout_p > 0.f && S <= 16 && !e.attn && layer_bits_requested()) {
      L.lbits[0] = ar.take<uint32_t>(dropout_bits_words(rows * E));
      L.lbits[1] = ar.take<uint32_t>(dropout_bits_words(rows * F));
      L.lbits[2] = ar.take<uint32_t>(dropout_bits_words(rows * E));
    }
    L.mean1 = ar.take<float>(rows);
    L.rstd1 = ar.take<float>(rows);
    L.mean2 = ar.take<float>(rows);
    L.rstd2 = ar.take<float>(rows);
  }
  const int64_t rb = static_cast<int64_t>(Rmax) * B;
  t.qp = ar.take<bf16>(B * E);
  t.kvp = ar.take<bf16>(rows * 2 * E);
  t.ap = ar.take<bf16>(rb * E);
  t.pv = ar.take<bf16>(rb * E);
  t.qt = ar.take<bf16>(rb * E);
  t.kvt = ar.take<bf16>(B * T * 2 * E);
  t.at = ar.take<bf16>(rb * E);
  t.c = ar.take<bf16>(rb * E);
  t.ot = ar.take<float>(B * E);
  t.tmpE = ar.take<bf16>(rows * E);
  if (e.img) {
    t.pe_h = ar.take<bf16>(B * P * E);
    t.pe_ln = ar.take<bf16>(B * P * E);
    t.zero = ar.take<bf16>(B * P * E);
    t.pe_mean = ar.take<float>(B * P);
    t.pe_rstd = ar.take<float>(B * P);
  }
}

static int64_t layout(gg_engine& e, uint8_t* base) {
  const gg_model_cfg& c = e.cfg;
  Arena ar(base);
  const int64_t B = c.B, S = e.S_, E = c.E, F = e.F, P = c.P, T = c.T, H = c.H, Gp = e.Gp;
  layout_shadows(e, ar);
  e.xfr = ar.take<bf16>(2 * B * Gp);
  e.xin = ar.take<bf16>(B * Gp);
  e.zbf = ar.take<bf16>(B * round_up64(c.L, 8));
  e.stats = ar.take<float>(GG_STATS_COUNT);
  e.opt_step[0] = ar.take<float>(4);
  e.opt_step[1] = ar.take<float>(4);
  e.normbuf = ar.take<float>(8);
  e.rng = ar.take<uint64_t>(2);
  if (e.concat) {
    if (e.label) {
      e.labels = ar.take<int64_t>(2 * B);
    } else {
      e.text = ar.take<bf16>(B * T * c.Dt);
    }
    for (int n = 0; n < 2; ++n) {
      e.tw[n].Rmax = 1;
      e.tw[n].c = ar.take<bf16>(B * E);
    }
    e.gs.dc = ar.take<bf16>(2 * B * E);
  } else if (e.cond) {
    e.patches = ar.take<bf16>(B * P * c.Dp);
    e.text = ar.take<bf16>(B * T * c.Dt);
    e.mask_s = ar.take<uint8_t>(B * S);
    e.tpad = ar.take<uint8_t>(B * T);
    {
      const int64_t lmax = S > T ? S : T;
      e.attn_stat = ar.take<float>(2 * 3 * B * c.n_heads * lmax);
    }
    layout_tower(e, e.tw[GG_NET_GEN], 1, ar);
    layout_tower(e, e.tw[GG_NET_DISC], c.dropout_p > 0.f ? 3 : 1, ar);
    GradScratch& g = e.gs;
    const int64_t n = 2 * B, rows = n * S;
    g.dc = ar.take<bf16>(n * E);
    g.dat = ar.take<bf16>(n * E);
    g.dqt = ar.take<bf16>(n * E);
    g.dkvt = ar.take<bf16>(n * T * 2 * E);
    g.dkvt_sum = ar.take<bf16>(B * T * 2 * E);
    g.dp = ar.take<bf16>(n * E);
    g.dap = ar.take<bf16>(n * E);
    g.dqp = ar.take<bf16>(n * E);
    g.dqp_sum = ar.take<bf16>(B * E);
    g.dcs = ar.take<bf16>(B * E);
    g.dkvp = ar.take<bf16>(rows * 2 * E);
    g.ga = ar.take<bf16>(rows * E);
    g.gb = ar.take<bf16>(rows * E);
    g.gao = ar.take<bf16>(rows * E);
    for (int l = 0; l < 2; ++l) {
      LayerGrads& lg = g.L[l];
      lg.gz2 = ar.take<bf16>(rows * E);
      lg.gy2 = ar.take<bf16>(rows * E);
      lg.gz1 = ar.take<bf16>(rows * E);
      lg.gy1 = ar.take<bf16>(rows * E);
      lg.gh = ar.take<bf16>(rows * F);
      lg.gqkv = ar.take<bf16>(rows * 3 * E);
      lg.lnp2 = ar.take<float>(ln_bwd_scratch_floats(rows, static_cast<int>(E)));
      lg.lnp1 = ar.take<float>(ln_bwd_scratch_floats(rows, static_cast<int>(E)));
    }
    g.dte = ar.take<bf16>(B * T * E);
    g.dte0 = ar.take<bf16>(B * E);
    g.dpe = ar.take<bf16>(B * P * E);
    g.dmod = ar.take<bf16>(B * P * c.Dp);
    g.dgb = ar.take<bf16>(B * 2 * c.Dp);
    if (e.attn) {
      e.bn_mean = ar.take<float>(E);
      e.bn_rstd = ar.take<float>(E);
    }
    if (e.img) {
      g.dpe_z = ar.take<bf16>(B * P * E);
      g.dpe_h = ar.take<bf16>(B * P * E);
      g.lnp_pe = ar.take<float>(ln_bwd_scratch_floats(B * P, static_cast<int>(E)));
    }
  }
  TrunkBufs& t = e.tb;
  t.a1x = ar.take<float>(2 * B * H);
  t.a1c = ar.take<float>(3 * B * H);
  t.h2f = ar.take<float>(3 * B * H);
  t.score = ar.take<float>(3 * B);
  t.Mg = ar.take<float>(H * H);
  t.u1f = ar.take<float>(B * H);
  t.y = ar.take<float>(B * H);
  t.norms = ar.take<float>(B);
  t.pen = ar.take<float>(B);
  t.du2f = ar.take<float>(B * H);
  t.roww = ar.take<float>(3 * B);
  t.h1 = ar.take<bf16>(3 * B * H);
  t.h2 = ar.take<bf16>(3 * B * H);
  t.Mgb = ar.take<bf16>(H * H);
  t.u2 = ar.take<bf16>(B * H);
  t.u1b = ar.take<bf16>(B * H);
  t.ru1 = ar.take<bf16>(B * H);
  t.dv1 = ar.take<bf16>(B * H);
  t.Qb = ar.take<bf16>(H * H);
  t.da2 = ar.take<bf16>(2 * B * H);
  t.da1 = ar.take<bf16>(2 * B * H);
  t.hg1 = ar.take<bf16>(B * H);
  t.hg2 = ar.take<bf16>(B * H);
  t.dfake = ar.take<bf16>(B * Gp);
  t.dag2 = ar.take<bf16>(B * H);
  t.dag1 = ar.take<bf16>(B * H);
  // scratch for deterministic column sums / LayerNorm weight grads
  int64_t maxN = c.G;
  if (3 * E > maxN) maxN = 3 * E;
  if (2 * c.Dp > maxN) maxN = 2 * c.Dp;
  if (F > maxN) maxN = F;
  if (H > maxN) maxN = H;
  int64_t scratch_floats = 64 * maxN;
  if (scratch_floats < 296 * 2 * E) scratch_floats = 296 * 2 * E;
  if (scratch_floats < 1024) scratch_floats = 1024;
  {
    int64_t max_elems = 0, max_cols = 0;
    for (int net = 0; net < 2; ++net) {
      int64_t elems = 0, cols = 0;
      for (int slot = 0; slot < GG_NSLOTS; ++slot) {
        int r, cc;
        if (slot_matrix_shape(c, net, slot, &r, &cc)) {
          elems += static_cast<int64_t>(r) * cc;
          cols += 3LL * r;  // bias gradients are column sums of [rows, out] matrices (in_proj parts counted apart)
        }
      }
      if (elems > max_elems) max_elems = elems;
      if (cols > max_cols) max_cols = cols;
    }
    e.wg_ws_bytes = wgrad_group_workspace_bytes(max_elems);
    e.cs_ws_bytes = colsum_group_workspace_bytes(max_cols);
    e.wg_ws = ar.take<uint8_t>(e.wg_ws_bytes);
    e.cs_ws = ar.take<uint8_t>(e.cs_ws_bytes);
  }
  e.splitk_bytes = 96LL << 20;
  for (int l = 0; l < gg_engine::NLANES; ++l) {
    e.scratch_l[l] = ar.take<float>(scratch_floats);
    e.splitk_l[l] = ar.take<uint8_t>(e.splitk_bytes);
  }
  return round_up64(ar.off, 256);
}

static int validate_cfg(const gg_model_cfg& c) {
  GG_REQUIRE(c.variant >= GG_VARIANT_VANILLA && c.variant <= GG_VARIANT_ATTN, "unknown variant %d", c.variant);
  GG_REQUIRE(c.variant != GG_VARIANT_ATTN || (c.dropout_p == 0.f && c.T == 1),
             "the attention variant has no dropout and one text vector per sample");
  GG_REQUIRE(c.B > 0 && c.G > 0 && c.L > 0 && c.H > 0, "bad sizes B=%d G=%d L=%d H=%d", c.B, c.G, c.L, c.H);
  GG_REQUIRE(c.L % 8 == 0 && c.H % 8 == 0, "latent and hidden widths must be multiples of 8");
  if (c.variant == GG_VARIANT_LABEL) {
    GG_REQUIRE(c.E % 16 == 0 && c.E <= 1024, "conditioning width %d must be a multiple of 16 (<= 1024)", c.E);
    GG_REQUIRE(c.Dt >= 1 && c.Dp >= 1, "vocabulary sizes must be positive (Dt=%d Dp=%d)", c.Dt, c.Dp);
  } else if (c.variant != GG_VARIANT_VANILLA) {
    GG_REQUIRE(c.E % 32 == 0 && c.E <= 1024, "embedding width %d must be a multiple of 32 (<= 1024)", c.E);
    GG_REQUIRE(c.n_heads > 0 && c.E % c.n_heads == 0 && (c.E / c.n_heads) % 2 == 0 && c.E / c.n_heads <= 64,
               "unsupported head configuration E=%d heads=%d", c.E, c.n_heads);
    GG_REQUIRE(c.n_layers >= 1 && c.n_layers <= 2, "n_layers must be 1 or 2");
    GG_REQUIRE(c.Dt % 8 == 0 && c.Dp % 8 == 0 && c.ffn % 8 == 0, "feature widths must be multiples of 8");
    GG_REQUIRE(c.P >= 1 && c.P + 1 <= 320 && c.T >= 1 && c.T <= 320, "token counts out of range P=%d T=%d", c.P, c.T);
    GG_REQUIRE(c.dropout_p >= 0.f && c.dropout_p < 1.f, "bad dropout");
  }
  return GG_OK;
}

static void derive(gg_engine& e) {
  const gg_model_cfg& c = e.cfg;
  e.cond = c.variant != GG_VARIANT_VANILLA;
  e.paper = c.variant == GG_VARIANT_PAPER || c.variant == GG_VARIANT_CROSS;
  e.film = c.variant == GG_VARIANT_PAPER || c.variant == GG_VARIANT_FILM;
  e.label = c.variant == GG_VARIANT_LABEL;
  e.concat = c.variant == GG_VARIANT_CONCAT || e.label;
  e.img = c.variant == GG_VARIANT_IMG;
  e.attn = c.variant == GG_VARIANT_ATTN;
  e.S_ = e.cond ? c.P + 1 : 1;
  e.Gp = static_cast<int>(round_up64(c.G, 8));
  e.F = c.ffn;
  e.hd = (e.cond && !e.label) ? c.E / c.n_heads : 0;
}

// the fused encoder-layer kernel covers the reference's layer shape (d_model 256, 4 heads, ffn 512) at <= 16 tokens
static bool cfg_fused_layer_ok(const gg_engine& e, bool any_tokens = false) {
  const gg_model_cfg& c = e.cfg;
  return c.gemm_impl == GG_IMPL_TCGEN05 && c.E == 256 && c.ffn == 512 && c.n_heads == 4 && (any_tokens || e.S_ <= 16);
}

// ------------------------------------------------------------------------------- tower forward
// Runs entirely on lane `ln` (the generator's tower runs next to the critic's on another lane).
// save_reps: replicas (leading row blocks) whose backward tensors are kept; the fused layer kernel writes nothing but
// its output for the others (the generator's tower inside a critic step, the critic's interpolated replica, inference)
static int tower_forward(gg_engine& e, int net, int R, float p, int ln, int save_reps) {
  const gg_model_cfg& c = e.cfg;
  Tower& t = e.tw[net];
  cudaStream_t st = e.S(ln);
  GG_REQUIRE(R <= t.Rmax, "tower replicas %d > allocated %d", R, t.Rmax);
  const int B = c.B, S = e.S_, E = c.E, F = e.F, P = c.P, T = c.T, Dp = c.Dp, Dt = c.Dt;
  const int rows = R * B * S;
  const uint32_t site0 = static_cast<uint32_t>(net) * 64u;
  if (e.label)  // benchmark_generative_model.py:138-157 / :204-224: c = [emb0[y0] | emb1[y1]]
    return k_embed_gather(e.P(net, GG_P_EMB0), e.P(net, GG_P_EMB1), e.labels, e.labels + B, Dt, Dp, t.c, B, E / 2, st);
  if (e.concat)  // conditional_gan_concat.py:135-139 / :182-186: c = encoder(text) (or of the masked mean patch)
    return e.linear(ln, B, E, Dt, Op{e.text, Dt}, e.W(net, GG_P_TEXT_W), Epi().bias(e.P(net, GG_P_TEXT_B)).obf(t.c, E));
  if (e.attn) {
    // conditional_gan_attention.py:113-126 / :154-161: query = text_encoder(text) [B, 1, E], keys / values =
    // patches_encoder(patches) [B, P, E] under the key-padding mask, one MultiheadAttention; BatchNorm1d in the generator
    const Op Wa = e.W(net, GG_P_P2T_IN_W);
    const float* ba = e.P(net, GG_P_P2T_IN_B);
    GG_TRY(e.linear(ln, B, E, Dt, Op{e.text, Dt}, e.W(net, GG_P_TEXT_W), Epi().bias(e.P(net, GG_P_TEXT_B)).obf(t.te, E)));
    GG_TRY(e.linear(ln, B * P, E, Dp, Op{e.patches, Dp}, e.W(net, GG_P_PATCH_W),
                    Epi().bias(e.P(net, GG_P_PATCH_B)).obf(t.X[0], E)));
    GG_TRY(e.linear(ln, B, E, E, Op{t.te, E}, Wa, Epi().bias(ba).obf(t.qp, E)));
    GG_TRY(e.linear(ln, B * P, 2 * E, E, Op{t.X[0], E}, Op{Wa.p + static_cast<int64_t>(E) * Wa.ld, Wa.ld},
                    Epi().bias(ba ? ba + E : nullptr).obf(t.kvp, 2 * E)));
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q = t.qp; a.ldq = E; a.q_mod = B;
    a.k = t.kvp; a.v = t.kvp + E; a.ldkv = 2 * E; a.kv_mod = B;
    a.mask = e.mask_s; a.mask_mod = B;  // [B, P]: the padding flags as they came (no CLS column in this variant)
    a.nb = B; a.H = c.n_heads; a.hd = e.hd; a.Lq = 1; a.Lk = P;
    a.o = t.ap; a.ldo = E;
    GG_TRY(k_attention_fwd(a, st));
    const bool bn = net == GG_NET_GEN;
    GG_TRY(e.linear(ln, B, E, E, Op{t.ap, E}, e.W(net, GG_P_P2T_OUT_W),
                    Epi().bias(e.P(net, GG_P_P2T_OUT_B)).obf(bn ? t.pv : t.c, E)));
    if (bn) {
      GG_REQUIRE(e.bn_run_mean && e.bn_run_var, "gg_engine_set_batchnorm has not been called");
      GG_TRY(k_bn_fwd(t.pv, E, e.P(net, GG_P_BN_W), e.P(net, GG_P_BN_B), e.bn_run_mean, e.bn_run_var, e.bn_momentum,
                      e.bn_eps, e.bn_training, t.c, E, e.bn_mean, e.bn_rstd, B, E, st));
    }
    return GG_OK;
  }
  // FiLM parameters from the text CLS / text vector (:129-134); conditional_gan_cross_attention.py has no FiLM
  // (:128-130): its patch encoder reads the patch embeddings as they are
  const bf16* pin = e.patches;
  // FiLM in the A-operand path of the patch encoder (film_patch.cu): modulation, projection, bias, CLS rows and the copies
  // for the dropout replicas in one launch; the modulated patches are stored only when this pass will be back-propagated
  const bool film_fused = e.film && e.film_fused && c.gemm_impl == GG_IMPL_TCGEN05 && E == 256 && Dp % 64 == 0;
  if (e.film) {
    GG_TRY(e.linear(ln, B, 2 * Dp, Dt, Op{e.text, static_cast<int64_t>(T) * Dt}, e.W(net, GG_P_FILM_W),
                    Epi().bias(e.P(net, GG_P_FILM_B)).act(GG_ACT_FILM).of32(t.gb, 2 * Dp)));
    if (!film_fused) {
      GG_TRY(k_film_apply(e.patches, t.gb, t.mod, B, P, Dp, st));
      pin = t.mod;
    }
  }
  // text side of the paper model (:140, :149-152): token projection, the patch2text query and the text2patch
  // keys / values depend on the text only — they run on lane 2 next to the patch encoder of lane `ln`
  const int tl = e.text_lane_fwd ? 2 : ln;
  // The per-element dropout masks of the fused layer kernels depend on nothing but the step counter: they are drawn here,
  // once, next to the tower head (side lane), and the layer kernels read bits instead of running Philox in their epilogues
  const bool layer_bits = p > 0.f && e.layer_bits && e.fused_layer && cfg_fused_layer_ok(e) && t.L[0].lbits[0] != nullptr;
  cudaEvent_t bits_ready = nullptr;
  if (layer_bits) {
    const int bl = 2;
    GG_TRY(e.wait_lane(bl, ln));
    for (int l = 0; l < c.n_layers; ++l) {
      const uint32_t site = site0 + 8u * l;
      const uint32_t sites[3] = {site + 1, site + 2, site + 3};
      const int64_t n[3] = {static_cast<int64_t>(rows) * E, static_cast<int64_t>(rows) * F, static_cast<int64_t>(rows) * E};
      uint32_t* const outs[3] = {t.L[l].lbits[0], t.L[l].lbits[1], t.L[l].lbits[2]};
      GG_TRY(k_dropout_bits3(e.rng, p, sites, n, outs, e.S(bl)));
    }
    if (e.L(bl) != e.L(ln)) GG_TRY(e.record_lane(bl, &bits_ready));  // waited for just before the first layer
  }
  if (e.paper) {
    const Op Wp = e.W(net, GG_P_P2T_IN_W), Wt = e.W(net, GG_P_T2P_IN_W);
    const float* bp = e.P(net, GG_P_P2T_IN_B);
    const float* bt = e.P(net, GG_P_T2P_IN_B);
    GG_TRY(e.wait_lane(tl, ln));
    GG_TRY(e.linear(tl, B * T, E, Dt, Op{e.text, Dt}, e.W(net, GG_P_TEXT_W),
                    Epi().bias(e.P(net, GG_P_TEXT_B)).obf(t.te, E)));
    GG_TRY(e.linear(tl, B, E, E, Op{t.te, static_cast<int64_t>(T) * E}, Wp, Epi().bias(bp).obf(t.qp, E)));
    GG_TRY(e.linear(tl, B * T, 2 * E, E, Op{t.te, E}, Op{Wt.p + static_cast<int64_t>(E) * Wt.ld, Wt.ld},
                    Epi().bias(bt ? bt + E : nullptr).obf(t.kvt, 2 * E)));
    if (T == 1 && e.t1_shortcut)  // the whole text2patch branch of this step: (v_text Wo^T + bo), once for all replicas
      GG_TRY(e.linear(tl, B, E, E, Op{t.kvt + E, 2 * E}, e.W(net, GG_P_T2P_OUT_W),
                      Epi().bias(e.P(net, GG_P_T2P_OUT_B)).of32(t.ot, E)));
  }
  if (e.img) {
    // conditional_gan_img_transformer.py:111-115: patches_encoder = Linear -> ReLU -> LayerNorm (no dropout: once
    // for the B*P rows, shared by the replicas), then CLS + tokens
    GG_TRY(e.linear(ln, B * P, E, Dp, Op{pin, Dp}, e.W(net, GG_P_PATCH_W),
                    Epi().bias(e.P(net, GG_P_PATCH_B)).act(GG_ACT_LEAKY, 0.f).obf(t.pe_h, E)));
    GG_TRY(k_add_ln_fwd(t.zero, t.pe_h, e.P(net, GG_P_PENC_LN_W), e.P(net, GG_P_PENC_LN_B), t.tmpE, t.pe_ln, t.pe_mean,
                        t.pe_rstd, static_cast<int64_t>(B) * P, E, c.ln_eps, 0.f, e.rng, 0, st));
    GG_TRY(k_assemble_tokens(t.X[0], e.P(net, GG_P_CLS), R, B, S, E, st, t.pe_ln));
  } else if (film_fused) {
    const Op Wpe = e.W(net, GG_P_PATCH_W);
    GG_TRY(k_film_patch(e.patches, t.gb, Wpe.p, Wpe.ld, e.P(net, GG_P_PATCH_B), e.P(net, GG_P_CLS), t.X[0],
                        save_reps > 0 ? t.mod : nullptr, B, P, R, Dp, st));
  } else {
    // patch projection written straight behind the CLS row of replica 0 (:139-142)
    GG_TRY(e.linear(ln, B * P, E, Dp, Op{pin, Dp}, e.W(net, GG_P_PATCH_W),
                    Epi().bias(e.P(net, GG_P_PATCH_B)).obf(t.X[0], E).rowmap(P, S, 1)));
    GG_TRY(k_assemble_tokens(t.X[0], e.P(net, GG_P_CLS), R, B, S, E, st));
  }
  for (int l = 0; l < c.n_layers; ++l) {
    TowerLayer& L = t.L[l];
    const int ls = GG_P_LAYER0 + GG_L_COUNT * l;
    const uint32_t site = site0 + 8u * l;
    if (e.fused_layer && cfg_fused_layer_ok(e)) {
      // the whole post-norm layer as one tcgen05 kernel (enc_layer.cu): same tensors, same dropout streams
      if (l == 0 && bits_ready) GG_TRY(e.wait_event(ln, bits_ready));
      EncLayerParams q;
      memset(&q, 0, sizeof(q));
      q.nb = R * B; q.S = S; q.E = E; q.F = F; q.n_heads = c.n_heads;
      q.save_rows = static_cast<int64_t>(save_reps < R ? save_reps : R) * B * S;
      q.x = t.X[l];
      const Op wi = e.W(net, ls + GG_L_IN_W), wo = e.W(net, ls + GG_L_OUT_W), w1 = e.W(net, ls + GG_L_FF1_W),
               w2 = e.W(net, ls + GG_L_FF2_W);
      q.w_in = wi.p; q.ld_in = wi.ld; q.w_out = wo.p; q.ld_out = wo.ld;
      q.w_ff1 = w1.p; q.ld_ff1 = w1.ld; q.w_ff2 = w2.p; q.ld_ff2 = w2.ld;
      q.b_in = e.P(net, ls + GG_L_IN_B); q.b_out = e.P(net, ls + GG_L_OUT_B);
      q.b_ff1 = e.P(net, ls + GG_L_FF1_B); q.b_ff2 = e.P(net, ls + GG_L_FF2_B);
      q.g1 = e.P(net, ls + GG_L_N1_W); q.be1 = e.P(net, ls + GG_L_N1_B);
      q.g2 = e.P(net, ls + GG_L_N2_W); q.be2 = e.P(net, ls + GG_L_N2_B);
      q.mask = e.mask_s; q.mask_mod = B;
      q.drop_p = p; q.eps = c.ln_eps; q.rng = e.rng; q.site = site;
      q.qkv = L.qkv; q.ao = L.ao; q.z1 = L.z1; q.x1 = L.x1; q.h = L.h; q.z2 = L.z2; q.out = t.X[l + 1];
      q.mean1 = L.mean1; q.rstd1 = L.rstd1; q.mean2 = L.mean2; q.rstd2 = L.rstd2;
      if (p > 0.f && L.lbits[0] && e.layer_bits) { q.dbits1 = L.lbits[0]; q.dbits2 = L.lbits[1]; q.dbits3 = L.lbits[2]; }
      GG_TRY(k_enc_layer_fwd(q, st));
      continue;
    }
    GG_TRY(e.linear(ln, rows, 3 * E, E, Op{t.X[l], E}, e.W(net, ls + GG_L_IN_W),
                    Epi().bias(e.P(net, ls + GG_L_IN_B)).obf(L.qkv, 3 * E)));
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q = L.qkv; a.ldq = 3 * E; a.q_mod = R * B;
    a.k = L.qkv + E; a.v = L.qkv + 2 * E; a.ldkv = 3 * E; a.kv_mod = R * B;
    a.mask = e.mask_s; a.mask_mod = B;
    a.nb = R * B; a.H = c.n_heads; a.hd = e.hd; a.Lq = S; a.Lk = S;
    a.drop_p = p; a.rng = e.rng; a.site = site + 0;
    a.o = L.ao; a.ldo = E;
    if (p > 0.f && L.dbits && e.attn_bits) {  // draw the probability-dropout mask once: forward and backward read bits
      GG_TRY(k_dropout_bits(e.rng, site + 0, p, static_cast<int64_t>(R) * B * c.n_heads * S * S, L.dbits, st));
      a.dbits = L.dbits;
    }
    GG_TRY(k_attention_fwd(a, st));
    // out-projection / linear2 with the residual + dropout + LayerNorm in the GEMM epilogue (gemm_ln.cu), else two launches
    const bool gemm_ln = e.gemm_ln && c.gemm_impl == GG_IMPL_TCGEN05 && E == 256 && F % 64 == 0;
    if (gemm_ln) {
      const Op wo = e.W(net, ls + GG_L_OUT_W);
      GG_TRY(k_gemm_ln(L.ao, E, wo.p, wo.ld, E, e.P(net, ls + GG_L_OUT_B), t.X[l], e.P(net, ls + GG_L_N1_W),
                       e.P(net, ls + GG_L_N1_B), L.z1, L.x1, L.mean1, L.rstd1, rows, c.ln_eps, p, e.rng, site + 1, st));
    } else {
    GG_TRY(e.linear(ln, rows, E, E, Op{L.ao, E}, e.W(net, ls + GG_L_OUT_W),
                    Epi().bias(e.P(net, ls + GG_L_OUT_B)).obf(t.tmpE, E)));
    GG_TRY(k_add_ln_fwd(t.X[l], t.tmpE, e.P(net, ls + GG_L_N1_W), e.P(net, ls + GG_L_N1_B), L.z1, L.x1, L.mean1,
                        L.rstd1, rows, E, c.ln_eps, p, e.rng, site + 1, st));
    }
    GG_TRY(e.linear(ln, rows, F, E, Op{L.x1, E}, e.W(net, ls + GG_L_FF1_W),
                    Epi().bias(e.P(net, ls + GG_L_FF1_B)).act(GG_ACT_LEAKY, 0.f).drop(p, e.rng, site + 2).obf(L.h, F)));
    if (gemm_ln) {
      const Op w2 = e.W(net, ls + GG_L_FF2_W);
      GG_TRY(k_gemm_ln(L.h, F, w2.p, w2.ld, F, e.P(net, ls + GG_L_FF2_B), L.x1, e.P(net, ls + GG_L_N2_W),
                       e.P(net, ls + GG_L_N2_B), L.z2, t.X[l + 1], L.mean2, L.rstd2, rows, c.ln_eps, p, e.rng, site + 3, st));
    } else {
    GG_TRY(e.linear(ln, rows, E, F, Op{L.h, F}, e.W(net, ls + GG_L_FF2_W),
                    Epi().bias(e.P(net, ls + GG_L_FF2_B)).obf(t.tmpE, E)));
    GG_TRY(k_add_ln_fwd(L.x1, t.tmpE, e.P(net, ls + GG_L_N2_W), e.P(net, ls + GG_L_N2_B), L.z2, t.X[l + 1],
                        L.mean2, L.rstd2, rows, E, c.ln_eps, p, e.rng, site + 3, st));
    }
  }
  if (!e.paper) return GG_OK;  // film: conditioning = CLS row of X[n_layers]
  bf16* Xf = t.X[c.n_layers];
  const Op Wp = e.W(net, GG_P_P2T_IN_W), Wt = e.W(net, GG_P_T2P_IN_W);
  const float* bp = e.P(net, GG_P_P2T_IN_B);
  const float* bt = e.P(net, GG_P_T2P_IN_B);
  // patch2text: query = encoded text CLS (lane 2, above), keys/values = encoder output (:149-150)
  GG_TRY(e.wait_lane(ln, tl));
  GG_TRY(e.linear(ln, rows, 2 * E, E, Op{Xf, E}, Op{Wp.p + static_cast<int64_t>(E) * Wp.ld, Wp.ld},
                  Epi().bias(bp ? bp + E : nullptr).obf(t.kvp, 2 * E)));
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.q = t.qp; a.ldq = E; a.q_mod = B;
  a.k = t.kvp; a.v = t.kvp + E; a.ldkv = 2 * E; a.kv_mod = R * B;
  a.mask = e.mask_s; a.mask_mod = B;
  a.nb = R * B; a.H = c.n_heads; a.hd = e.hd; a.Lq = 1; a.Lk = S;
  a.o = t.ap; a.ldo = E;
  GG_TRY(k_attention_fwd(a, st));
  GG_TRY(e.linear(ln, R * B, E, E, Op{t.ap, E}, e.W(net, GG_P_P2T_OUT_W),
                  Epi().bias(e.P(net, GG_P_P2T_OUT_B)).obf(t.pv, E)));
  if (T == 1 && e.t1_shortcut)  // softmax over the one text token = 1: c = pv + (v_text Wo^T + bo) (:151-155)
    return k_add_bcast_replicas(t.pv, t.ot, t.c, R, static_cast<int64_t>(B) * E, st);
  // text2patch: query = that vector, keys/values = encoded text tokens (:151-152)
  GG_TRY(e.linear(ln, R * B, E, E, Op{t.pv, E}, Wt, Epi().bias(bt).obf(t.qt, E)));
  memset(&a, 0, sizeof(a));
  a.q = t.qt; a.ldq = E; a.q_mod = R * B;
  a.k = t.kvt; a.v = t.kvt + E; a.ldkv = 2 * E; a.kv_mod = B;
  a.mask = e.has_tpad ? e.tpad : nullptr; a.mask_mod = B;
  a.nb = R * B; a.H = c.n_heads; a.hd = e.hd; a.Lq = 1; a.Lk = T;
  a.o = t.at; a.ldo = E;
  GG_TRY(k_attention_fwd(a, st));
  // c = text vector + patch vector (:153-155)
  GG_TRY(e.linear(ln, R * B, E, E, Op{t.at, E}, e.W(net, GG_P_T2P_OUT_W),
                  Epi().bias(e.P(net, GG_P_T2P_OUT_B)).res(t.pv, E).obf(t.c, E)));
  return GG_OK;
}

static Op cond_vec(const gg_engine& e, int net) {
  const Tower& t = e.tw[net];
  if (e.paper || e.concat || e.attn) return Op{t.c, e.cfg.E};
  return Op{t.X[e.cfg.n_layers], static_cast<int64_t>(e.S_) * e.cfg.E};
}

// ------------------------------------------------------------------------------ tower backward
// dc: [Rg*B, E] gradient of the loss w.r.t. the conditioning vectors of the first Rg replicas.
// The dependent chain (dgrads, attention / LayerNorm backward) runs on lane 0; every weight- and
// bias-gradient reduction is forked onto the side lanes (e.wgrad / e.bgrad).
// Stages (data parallel: the trainer all-reduces a stage's gradients while the next stages run): 0 = head (the
// cross-attention tail / CLS scatter), 1 .. n_layers = encoder layers from the last to the first, n_layers + 1 =
// embedding tail (CLS, patch encoder, FiLM). Runs the stages in [s0, s1).
static int tower_backward(gg_engine& e, int net, int Rg, float p, const bf16* dc, int s0 = 0, int s1 = 1 << 20) {
  const gg_model_cfg& c = e.cfg;
  Tower& t = e.tw[net];
  GradScratch& g = e.gs;
  cudaStream_t st = e.S(0);
  const int B = c.B, S = e.S_, E = c.E, F = e.F, P = c.P, T = c.T, Dp = c.Dp, Dt = c.Dt;
  const int n = Rg * B, rows = n * S;
  const uint32_t site0 = static_cast<uint32_t>(net) * 64u;
  const float keep_scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  if (e.label) {  // rows of the two tables: sums of the dc rows that carry each label
    if (s0 > 0) return GG_OK;
    return k_embed_grad(dc, e.labels, e.labels + B, Dt, Dp, e.Gr(net, GG_P_EMB0), e.Gr(net, GG_P_EMB1), B, E / 2, st);
  }
  if (e.concat) {
    if (s0 > 0) return GG_OK;
    GG_TRY(e.wgrad(E, Dt, B, Op{dc, E}, Op{e.text, Dt}, e.Gr(net, GG_P_TEXT_W), Dt));
    GG_TRY(e.bgrad(dc, E, B, E, e.Gr(net, GG_P_TEXT_B)));
    return e.flush_grads();
  }
  if (e.attn) {
    if (s0 > 0) return GG_OK;
    GG_REQUIRE(Rg == 1, "the attention variant has one tower pass");
    const Op Wa = e.W(net, GG_P_P2T_IN_W);
    const Op Wa_kv{Wa.p + static_cast<int64_t>(E) * Wa.ld, Wa.ld};
    float* gWa = e.Gr(net, GG_P_P2T_IN_W);
    float* gba = e.Gr(net, GG_P_P2T_IN_B);
    const bf16* dpv = dc;
    if (net == GG_NET_GEN) {  // c = attn_bn(pv)
      GG_TRY(k_bn_bwd(dc, E, t.pv, E, e.bn_mean, e.bn_rstd, e.P(net, GG_P_BN_W), g.dp, E, e.Gr(net, GG_P_BN_W),
                      e.Gr(net, GG_P_BN_B), B, E, st));
      dpv = g.dp;
    }
    // pv = ap Wo^T + bo
    GG_TRY(e.wgrad(E, E, B, Op{dpv, E}, Op{t.ap, E}, e.Gr(net, GG_P_P2T_OUT_W), E));
    GG_TRY(e.bgrad(dpv, E, B, E, e.Gr(net, GG_P_P2T_OUT_B)));
    GG_TRY(e.dgrad(0, B, E, E, Op{dpv, E}, e.W(net, GG_P_P2T_OUT_W), Epi().obf(g.dap, E)));
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q = t.qp; a.ldq = E; a.q_mod = B;
    a.k = t.kvp; a.v = t.kvp + E; a.ldkv = 2 * E; a.kv_mod = B;
    a.mask = e.mask_s; a.mask_mod = B;
    a.nb = B; a.H = c.n_heads; a.hd = e.hd; a.Lq = 1; a.Lk = P;
    a.dout = g.dap; a.lddo = E; a.dq = g.dqp; a.lddq = E;
    a.dk = g.dkvp; a.dv = g.dkvp + E; a.lddkv = 2 * E;
    a.stat = e.attn_stat;
    GG_TRY(k_attention_bwd(a, st));
    // q = te Wq^T + bq ; te = text Wt^T + bt
    GG_TRY(e.wgrad(E, E, B, Op{g.dqp, E}, Op{t.te, E}, gWa, E));
    GG_TRY(e.bgrad(g.dqp, E, B, E, gba));
    GG_TRY(e.dgrad(0, B, E, E, Op{g.dqp, E}, Wa, Epi().obf(g.dte0, E)));
    GG_TRY(e.wgrad(E, Dt, B, Op{g.dte0, E}, Op{e.text, Dt}, e.Gr(net, GG_P_TEXT_W), Dt));
    GG_TRY(e.bgrad(g.dte0, E, B, E, e.Gr(net, GG_P_TEXT_B)));
    // kv = pe Wkv^T + bkv ; pe = patches Wp^T + bp
    GG_TRY(e.wgrad(2 * E, E, B * P, Op{g.dkvp, 2 * E}, Op{t.X[0], E}, gWa + static_cast<int64_t>(E) * E, E));
    GG_TRY(e.bgrad(g.dkvp, 2 * E, B * P, 2 * E, gba ? gba + E : nullptr));
    GG_TRY(e.dgrad(0, B * P, E, 2 * E, Op{g.dkvp, 2 * E}, Wa_kv, Epi().obf(g.dpe, E)));
    GG_TRY(e.wgrad(E, Dp, B * P, Op{g.dpe, E}, Op{e.patches, Dp}, e.Gr(net, GG_P_PATCH_W), Dp));
    GG_TRY(e.bgrad(g.dpe, E, B * P, E, e.Gr(net, GG_P_PATCH_B)));
    return e.flush_grads();
  }
  bf16* Xf = t.X[c.n_layers];
  const bool head = s0 <= 0 && 0 < s1;
  const int tail_stage = c.n_layers + 1;
  if (!head) {
  } else if (e.paper) {
    const Op Wp = e.W(net, GG_P_P2T_IN_W), Wt = e.W(net, GG_P_T2P_IN_W);
    const Op Wp_kv{Wp.p + static_cast<int64_t>(E) * Wp.ld, Wp.ld}, Wt_kv{Wt.p + static_cast<int64_t>(E) * Wt.ld, Wt.ld};
    float* gWp = e.Gr(net, GG_P_P2T_IN_W);
    float* gWt = e.Gr(net, GG_P_T2P_IN_W);
    float* gbp = e.Gr(net, GG_P_P2T_IN_B);
    float* gbt = e.Gr(net, GG_P_T2P_IN_B);
    const bool t1 = T == 1 && e.t1_shortcut;
    const bf16* dp_chain = g.dp;  // gradient w.r.t. pv that the patch2text backward continues from
    AttnArgs a;
    if (t1) {
      // c = pv + (v_text Wo_t^T + bo_t): dp = dc; everything else of the text2patch branch sees dc summed over the
      // replicas and only feeds the text encoder (lane `bl`, off the dependent chain). No gradient reaches the query
      // projection or the key projection: their slices of the in-proj gradient are exactly zero, as in the reference.
      dp_chain = dc;
      const int bl1 = e.text_lane_bwd ? 2 : 0;
      GG_TRY(e.wait_lane(bl1, 0));
      cudaStream_t sb = e.S(bl1);
      const bf16* dcs = dc;
      if (Rg > 1) {
        GG_TRY(k_sum_replicas(dc, g.dcs, Rg, static_cast<int64_t>(B) * E, sb));
        dcs = g.dcs;
      }
      GG_TRY(e.wgrad(E, E, B, Op{dcs, E}, Op{t.kvt + E, 2 * E}, e.Gr(net, GG_P_T2P_OUT_W), E));
      GG_TRY(e.bgrad(dcs, E, B, E, e.Gr(net, GG_P_T2P_OUT_B)));
      GG_CUDA_CHECK(cudaMemsetAsync(g.dkvt_sum, 0, sizeof(bf16) * static_cast<size_t>(B) * 2 * E, sb));
      GG_TRY(e.dgrad(bl1, B, E, E, Op{dcs, E}, e.W(net, GG_P_T2P_OUT_W), Epi().obf(g.dkvt_sum + E, 2 * E)));  // dV
      GG_CUDA_CHECK(cudaMemsetAsync(gWt, 0, sizeof(float) * static_cast<size_t>(E) * E, sb));   // query projection: 0
      if (gbt) GG_CUDA_CHECK(cudaMemsetAsync(gbt, 0, sizeof(float) * static_cast<size_t>(E), sb));
      GG_TRY(e.wgrad(2 * E, E, B, Op{g.dkvt_sum, 2 * E}, Op{t.te, E}, gWt + static_cast<int64_t>(E) * E, E));
      GG_TRY(e.bgrad(g.dkvt_sum, 2 * E, B, 2 * E, gbt ? gbt + E : nullptr));
      GG_TRY(e.dgrad(bl1, B, E, 2 * E, Op{g.dkvt_sum, 2 * E}, Wt_kv, Epi().obf(g.dte, E)));
    }
    if (!t1) {
    // c = at Wo_t^T + bo_t + pv
    GG_TRY(e.wgrad(E, E, n, Op{dc, E}, Op{t.at, E}, e.Gr(net, GG_P_T2P_OUT_W), E));
    GG_TRY(e.bgrad(dc, E, n, E, e.Gr(net, GG_P_T2P_OUT_B)));
    GG_TRY(e.dgrad(0, n, E, E, Op{dc, E}, e.W(net, GG_P_T2P_OUT_W), Epi().obf(g.dat, E)));
    memset(&a, 0, sizeof(a));
    a.q = t.qt; a.ldq = E; a.q_mod = n;
    a.k = t.kvt; a.v = t.kvt + E; a.ldkv = 2 * E; a.kv_mod = B;
    a.mask = e.has_tpad ? e.tpad : nullptr; a.mask_mod = B;
    a.nb = n; a.H = c.n_heads; a.hd = e.hd; a.Lq = 1; a.Lk = T;
    a.dout = g.dat; a.lddo = E; a.dq = g.dqt; a.lddq = E;
    a.dk = g.dkvt; a.dv = g.dkvt + E; a.lddkv = 2 * E;
    a.stat = e.attn_stat;
    GG_TRY(k_attention_bwd(a, st));
    // qt = pv Wq_t^T + bq_t ; dp = dqt Wq_t + dc (residual path)
    GG_TRY(e.wgrad(E, E, n, Op{g.dqt, E}, Op{t.pv, E}, gWt, E));
    GG_TRY(e.bgrad(g.dqt, E, n, E, gbt));
    GG_TRY(e.dgrad(0, n, E, E, Op{g.dqt, E}, Wt, Epi().res(dc, E).obf(g.dp, E)));
    // kvt = te Wkv_t^T + bkv_t (shared by the replicas). Everything that only flows back into the text encoder
    // (dkvt -> dte, dqp -> dte0) leaves the dependent chain: lane 2
    const bf16* dkvt = g.dkvt;
    const int bl = e.text_lane_bwd ? 2 : 0;
    GG_TRY(e.wait_lane(bl, 0));
    if (Rg > 1) {
      GG_TRY(k_sum_replicas(g.dkvt, g.dkvt_sum, Rg, static_cast<int64_t>(B) * T * 2 * E, e.S(bl)));
      dkvt = g.dkvt_sum;
    }
    GG_TRY(e.wgrad(2 * E, E, B * T, Op{dkvt, 2 * E}, Op{t.te, E}, gWt + static_cast<int64_t>(E) * E, E));
    GG_TRY(e.bgrad(dkvt, 2 * E, B * T, 2 * E, gbt ? gbt + E : nullptr));
    GG_TRY(e.dgrad(bl, B * T, E, 2 * E, Op{dkvt, 2 * E}, Wt_kv, Epi().obf(g.dte, E)));
    }  // !t1
    const int bl = e.text_lane_bwd ? 2 : 0;
    // pv = ap Wo_p^T + bo_p
    GG_TRY(e.wgrad(E, E, n, Op{dp_chain, E}, Op{t.ap, E}, e.Gr(net, GG_P_P2T_OUT_W), E));
    GG_TRY(e.bgrad(dp_chain, E, n, E, e.Gr(net, GG_P_P2T_OUT_B)));
    GG_TRY(e.dgrad(0, n, E, E, Op{dp_chain, E}, e.W(net, GG_P_P2T_OUT_W), Epi().obf(g.dap, E)));
    memset(&a, 0, sizeof(a));
    a.q = t.qp; a.ldq = E; a.q_mod = B;
    a.k = t.kvp; a.v = t.kvp + E; a.ldkv = 2 * E; a.kv_mod = n;
    a.mask = e.mask_s; a.mask_mod = B;
    a.nb = n; a.H = c.n_heads; a.hd = e.hd; a.Lq = 1; a.Lk = S;
    a.dout = g.dap; a.lddo = E; a.dq = g.dqp; a.lddq = E;
    a.dk = g.dkvp; a.dv = g.dkvp + E; a.lddkv = 2 * E;
    a.stat = e.attn_stat;
    GG_TRY(k_attention_bwd(a, st));
    // qp = te[:,0] Wq_p^T + bq_p (shared by the replicas)
    const bf16* dqp = g.dqp;
    GG_TRY(e.wait_lane(bl, 0));
    if (Rg > 1) {
      GG_TRY(k_sum_replicas(g.dqp, g.dqp_sum, Rg, static_cast<int64_t>(B) * E, e.S(bl)));
      dqp = g.dqp_sum;
    }
    GG_TRY(e.wgrad(E, E, B, Op{dqp, E}, Op{t.te, static_cast<int64_t>(T) * E}, gWp, E));
    GG_TRY(e.bgrad(dqp, E, B, E, gbp));
    GG_TRY(e.dgrad(bl, B, E, E, Op{dqp, E}, Wp, Epi().obf(g.dte0, E)));
    GG_TRY(k_scatter_add_rows(g.dte, g.dte0, B, T, E, e.S(bl)));
    // kvp = Xf Wkv_p^T + bkv_p
    GG_TRY(e.wgrad(2 * E, E, rows, Op{g.dkvp, 2 * E}, Op{Xf, E}, gWp + static_cast<int64_t>(E) * E, E));
    GG_TRY(e.bgrad(g.dkvp, 2 * E, rows, 2 * E, gbp ? gbp + E : nullptr));
    GG_TRY(e.dgrad(0, rows, E, 2 * E, Op{g.dkvp, 2 * E}, Wp_kv, Epi().obf(g.ga, E)));
    // text encoder
    GG_TRY(e.wgrad(E, Dt, B * T, Op{g.dte, E}, Op{e.text, Dt}, e.Gr(net, GG_P_TEXT_W), Dt));
    GG_TRY(e.bgrad(g.dte, E, B * T, E, e.Gr(net, GG_P_TEXT_B)));
    GG_TRY(e.flush_grads());
  } else {
    GG_TRY(k_scatter_cls(g.ga, dc, n, S, E, st));
  }
  for (int l = c.n_layers - 1; l >= 0; --l) {
    const int stage = c.n_layers - l;
    if (stage < s0 || stage >= s1) continue;
    TowerLayer& L = t.L[l];
    LayerGrads& lg = g.L[l];
    const int ls = GG_P_LAYER0 + GG_L_COUNT * l;
    const uint32_t site = site0 + 8u * l;
    // x_out = LN2(x1 + drop(ff))
    const bf16* dff = p > 0.f ? lg.gy2 : lg.gz2;
    const bool fused_ffn = e.fused_layer && cfg_fused_layer_ok(e, /*any_tokens=*/true);
    if (fused_ffn) {
      // dependent chain: ONE kernel (enc_layer.cu::enc_ffn_bwd_kernel) LayerNorm-2 backward -> dgrad ffn2 -> masks ->
      // dgrad ffn1 -> + residual. The LayerNorm parameter gradients and the bf16 gz / gy tensors the weight-gradient
      // GEMMs read come from the unfused LayerNorm-backward kernel, which now runs next to it on lane 2.
      GG_TRY(e.fork(2));
      GG_TRY(k_add_ln_bwd(g.ga, L.z2, L.mean2, L.rstd2, e.P(net, ls + GG_L_N2_W), lg.gz2, p > 0.f ? lg.gy2 : nullptr,
                          nullptr, nullptr, rows, E, p, e.rng, site + 3, lg.lnp2, e.S(2)));
      GG_TRY(k_ln_bwd_finish(lg.lnp2, rows, E, e.Gr(net, ls + GG_L_N2_W), e.Gr(net, ls + GG_L_N2_B), e.S(2)));
      EncFfnBwdParams q;
      memset(&q, 0, sizeof(q));
      q.rows = rows;
      q.dout = g.ga; q.z2 = L.z2; q.mean2 = L.mean2; q.rstd2 = L.rstd2; q.gamma2 = e.P(net, ls + GG_L_N2_W);
      q.h = L.h;
      q.w2t = e.sh[net].wt[ls + GG_L_FF2_W].p; q.ld_w2t = e.sh[net].wt[ls + GG_L_FF2_W].ld;
      q.w1t = e.sh[net].wt[ls + GG_L_FF1_W].p; q.ld_w1t = e.sh[net].wt[ls + GG_L_FF1_W].ld;
      q.drop_p = p; q.rng = e.rng; q.site = site + 3;
      if (p > 0.f && L.lbits[2] && e.layer_bits && e.fused_layer && cfg_fused_layer_ok(e)) q.dbits = L.lbits[2];
      q.gh = lg.gh; q.gb = g.gb;
      GG_TRY(k_enc_ffn_bwd(q, st));
      GG_TRY(e.wgrad(E, F, rows, Op{dff, E}, Op{L.h, F}, e.Gr(net, ls + GG_L_FF2_W), F));
      GG_TRY(e.bgrad(dff, E, rows, E, e.Gr(net, ls + GG_L_FF2_B)));
      GG_TRY(e.wgrad(F, E, rows, Op{lg.gh, F}, Op{L.x1, E}, e.Gr(net, ls + GG_L_FF1_W), E));
      GG_TRY(e.bgrad(lg.gh, F, rows, F, e.Gr(net, ls + GG_L_FF1_B)));
    } else {
      GG_TRY(k_add_ln_bwd(g.ga, L.z2, L.mean2, L.rstd2, e.P(net, ls + GG_L_N2_W), lg.gz2, p > 0.f ? lg.gy2 : nullptr,
                          nullptr, nullptr, rows, E, p, e.rng, site + 3, lg.lnp2, st));
      GG_TRY(e.fork(2));
      GG_TRY(k_ln_bwd_finish(lg.lnp2, rows, E, e.Gr(net, ls + GG_L_N2_W), e.Gr(net, ls + GG_L_N2_B), e.S(2)));
      GG_TRY(e.wgrad(E, F, rows, Op{dff, E}, Op{L.h, F}, e.Gr(net, ls + GG_L_FF2_W), F));
      GG_TRY(e.bgrad(dff, E, rows, E, e.Gr(net, ls + GG_L_FF2_B)));
      GG_TRY(e.dgrad(0, rows, F, E, Op{dff, E}, e.W(net, ls + GG_L_FF2_W),
                     Epi().mask(L.h, F, keep_scale, 0.f).obf(lg.gh, F)));
      GG_TRY(e.wgrad(F, E, rows, Op{lg.gh, F}, Op{L.x1, E}, e.Gr(net, ls + GG_L_FF1_W), E));
      GG_TRY(e.bgrad(lg.gh, F, rows, F, e.Gr(net, ls + GG_L_FF1_B)));
      GG_TRY(e.dgrad(0, rows, E, F, Op{lg.gh, F}, e.W(net, ls + GG_L_FF1_W), Epi().res(lg.gz2, E).obf(g.gb, E)));
    }
    // x1 = LN1(x_in + drop(sa))
    GG_TRY(k_add_ln_bwd(g.gb, L.z1, L.mean1, L.rstd1, e.P(net, ls + GG_L_N1_W), lg.gz1, p > 0.f ? lg.gy1 : nullptr,
                        nullptr, nullptr, rows, E, p, e.rng, site + 1, lg.lnp1, st));
    GG_TRY(e.fork(2));
    GG_TRY(k_ln_bwd_finish(lg.lnp1, rows, E, e.Gr(net, ls + GG_L_N1_W), e.Gr(net, ls + GG_L_N1_B), e.S(2)));
    const bf16* dsa = p > 0.f ? lg.gy1 : lg.gz1;
    GG_TRY(e.wgrad(E, E, rows, Op{dsa, E}, Op{L.ao, E}, e.Gr(net, ls + GG_L_OUT_W), E));
    GG_TRY(e.bgrad(dsa, E, rows, E, e.Gr(net, ls + GG_L_OUT_B)));
    GG_TRY(e.dgrad(0, rows, E, E, Op{dsa, E}, e.W(net, ls + GG_L_OUT_W), Epi().obf(g.gao, E)));
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q = L.qkv; a.ldq = 3 * E; a.q_mod = n;
    a.k = L.qkv + E; a.v = L.qkv + 2 * E; a.ldkv = 3 * E; a.kv_mod = n;
    a.mask = e.mask_s; a.mask_mod = B;
    a.nb = n; a.H = c.n_heads; a.hd = e.hd; a.Lq = S; a.Lk = S;
    a.drop_p = p; a.rng = e.rng; a.site = site + 0;
    if (p > 0.f && L.dbits && e.attn_bits) a.dbits = L.dbits;  // the forward of this step drew them
    a.o = L.ao; a.ldo = E;  // forward output (kept for the out-projection's weight gradient): delta_i = dO_i . O_i
    a.dout = g.gao; a.lddo = E; a.dq = lg.gqkv; a.lddq = 3 * E;
    a.dk = lg.gqkv + E; a.dv = lg.gqkv + 2 * E; a.lddkv = 3 * E;
    a.stat = e.attn_stat;
    GG_TRY(k_attention_bwd(a, st));
    GG_TRY(e.wgrad(3 * E, E, rows, Op{lg.gqkv, 3 * E}, Op{t.X[l], E}, e.Gr(net, ls + GG_L_IN_W), E));
    GG_TRY(e.bgrad(lg.gqkv, 3 * E, rows, 3 * E, e.Gr(net, ls + GG_L_IN_B)));
    if (fused_ffn) GG_TRY(e.wait_lane(0, 2));  // lane 2's LayerNorm-2 backward has read g.ga: it may be overwritten
    GG_TRY(e.dgrad(0, rows, E, 3 * E, Op{lg.gqkv, 3 * E}, e.W(net, ls + GG_L_IN_W), Epi().res(lg.gz1, E).obf(g.ga, E)));
    GG_TRY(e.flush_grads());
  }
  if (tail_stage < s0 || tail_stage >= s1) return e.flush_grads();
  // X0 = [cls | patch projections], replicas share the projections
  GG_TRY(e.bgrad(g.ga, static_cast<int64_t>(S) * E, n, E, e.Gr(net, GG_P_CLS)));  // d cls = sum over the CLS rows
  GG_TRY(k_unassemble_tokens(g.ga, g.dpe, nullptr, Rg, B, S, E, st));
  if (e.img) {  // LayerNorm and ReLU of the patch encoder, then its Linear's gradients (the patches are inputs)
    GG_TRY(k_add_ln_bwd(g.dpe, t.pe_h, t.pe_mean, t.pe_rstd, e.P(net, GG_P_PENC_LN_W), g.dpe_z, nullptr, nullptr,
                        nullptr, static_cast<int64_t>(B) * P, E, 0.f, e.rng, 0, g.lnp_pe, st));
    GG_TRY(e.fork(2));
    GG_TRY(k_ln_bwd_finish(g.lnp_pe, static_cast<int64_t>(B) * P, E, e.Gr(net, GG_P_PENC_LN_W),
                           e.Gr(net, GG_P_PENC_LN_B), e.S(2)));
    GG_TRY(k_relu_bwd(g.dpe_z, t.pe_h, g.dpe_h, static_cast<int64_t>(B) * P * E, st));
    GG_TRY(e.wgrad(E, Dp, B * P, Op{g.dpe_h, E}, Op{e.patches, Dp}, e.Gr(net, GG_P_PATCH_W), Dp));
    GG_TRY(e.bgrad(g.dpe_h, E, B * P, E, e.Gr(net, GG_P_PATCH_B)));
    return e.flush_grads();
  }
  GG_TRY(e.wgrad(E, Dp, B * P, Op{g.dpe, E}, Op{e.film ? t.mod : e.patches, Dp}, e.Gr(net, GG_P_PATCH_W), Dp));
  GG_TRY(e.bgrad(g.dpe, E, B * P, E, e.Gr(net, GG_P_PATCH_B)));
  if (e.film) {
    if (e.split_tail_flush) GG_TRY(e.flush_grads());  // patch-encoder gradients overlap the FiLM backward  // without FiLM the patch embeddings are a plain input: nothing upstream needs a gradient
    GG_TRY(e.dgrad(0, B * P, Dp, E, Op{g.dpe, E}, e.W(net, GG_P_PATCH_W), Epi().obf(g.dmod, Dp)));
    GG_TRY(k_film_bwd(g.dmod, e.patches, t.gb, g.dgb, B, P, Dp, st));
    GG_TRY(e.wgrad(2 * Dp, Dt, B, Op{g.dgb, 2 * Dp}, Op{e.text, static_cast<int64_t>(T) * Dt},
                   e.Gr(net, GG_P_FILM_W), Dt));
    GG_TRY(e.bgrad(g.dgb, 2 * Dp, B, 2 * Dp, e.Gr(net, GG_P_FILM_B)));
  }
  return e.flush_grads();
}

// --------------------------------------------------------------------------- generator forward
static int gen_forward(gg_engine& e, const float* z, float p, float* out_f32, int ln, bool need_bwd) {
  const gg_model_cfg& c = e.cfg;
  TrunkBufs& t = e.tb;
  cudaStream_t st = e.S(ln);
  const int B = c.B, H = c.H, L = c.L, G = c.G, E = c.E;
  const int Lp = static_cast<int>(round_up64(L, 8));
  GG_TRY(k_cast_f32_bf16(z, L, e.zbf, Lp, B, L, st));
  const int net = GG_NET_GEN;
  Epi e1 = Epi().bias(e.P(net, GG_P_TR0_B)).act(GG_ACT_LEAKY, c.slope).obf(t.hg1, H);
  if (e.cond) {
    GG_TRY(tower_forward(e, net, 1, p, ln, need_bwd ? 1 : 0));
    const Op cv = cond_vec(e, net);
    GG_TRY(e.mm(ln, B, H, L, Op{e.zbf, Lp}, 0, e.W(net, GG_P_TR0_W), 0, e1, E, cv,
                Op{e.sh[net].tr0_c.p, e.sh[net].tr0_c.ld}));
  } else {
    GG_TRY(e.linear(ln, B, H, L, Op{e.zbf, Lp}, e.W(net, GG_P_TR0_W), e1));
  }
  GG_TRY(e.linear(ln, B, H, H, Op{t.hg1, H}, e.W(net, GG_P_TR1_W),
                  Epi().bias(e.P(net, GG_P_TR1_B)).act(GG_ACT_LEAKY, c.slope).obf(t.hg2, H)));
  Epi ef = Epi().bias(e.P(net, GG_P_FIN_B)).obf(e.xfr, e.Gp);
  if (out_f32) ef.of32(out_f32, G);
  GG_TRY(e.linear(ln, B, G, H, Op{t.hg2, H}, e.W(net, GG_P_FIN_W), ef));
  return GG_OK;
}

// critic trunk forward on npass row groups; nx = number of gene matrices in xfr (1: fake, 2: fake+real)
// fake_f32 / real_f32 (both non-null): the first layer reads those fp32 [B, G] tensors directly (TF32) instead of the bf16
// [fake; real] matrix x
// defer_score: the caller runs critic_scores() itself (on a side lane: in the training steps the scores only feed the loss
// statistics, the backward starts from h2's signs)
static int critic_scores(gg_engine& e, int npass, cudaStream_t st) {
  const gg_model_cfg& c = e.cfg;
  return k_rowdot_bias(e.tb.h2f, e.P(GG_NET_DISC, GG_P_FIN_W), e.P(GG_NET_DISC, GG_P_FIN_B), e.tb.score, npass * c.B, c.H, st);
}
static int critic_trunk_forward(gg_engine& e, const bf16* x, int nx, int npass, int R, const float* alpha,
                                const float* fake_f32 = nullptr, const float* real_f32 = nullptr, bool defer_score = false) {
  const gg_model_cfg& c = e.cfg;
  TrunkBufs& t = e.tb;
  cudaStream_t st = e.S(0);
  const int B = c.B, H = c.H, G = c.G, E = c.E, net = GG_NET_DISC;
  if (fake_f32 && real_f32 && H == 256 && !e.gp_tf32) {
    // fp32 profiles read in place, converted on chip, weight k-blocks multicast across a cluster (xw_f32.cu)
    const Op W1 = e.W(net, GG_P_TR0_W);
    GG_TRY(k_xw_f32(fake_f32, real_f32, B, G, W1.p, W1.ld, t.a1x, e.splitk_l[0], e.splitk_bytes, st));
  } else if (fake_f32 && real_f32) {
    const int64_t ldw = static_cast<int64_t>(G) + (e.cond ? E : 0);
    GG_TRY(e.linear_tf32(0, B, H, G, fake_f32, G, e.P(net, GG_P_TR0_W), ldw, Epi().of32(t.a1x, H)));
    GG_TRY(e.linear_tf32(0, B, H, G, real_f32, G, e.P(net, GG_P_TR0_W), ldw,
                         Epi().of32(t.a1x + static_cast<int64_t>(B) * H, H)));
  } else {
    GG_TRY(e.linear(0, nx * B, H, G, Op{x, e.Gp}, e.W(net, GG_P_TR0_W), Epi().of32(t.a1x, H)));
  }
  const float* a1c = nullptr;
  if (e.cond) {
    const Op cv = cond_vec(e, net);
    GG_TRY(e.linear(0, R * B, H, E, cv, Op{e.sh[net].tr0_c.p, e.sh[net].tr0_c.ld},
                    Epi().bias(e.P(net, GG_P_TR0_B)).of32(t.a1c, H)));
    a1c = t.a1c;
  }
  GG_TRY(k_trunk1_combine(t.a1x, a1c, e.P(net, GG_P_TR0_B), alpha, t.h1, B, H, npass, R, c.slope, st));
  GG_TRY(e.linear(0, npass * B, H, H, Op{t.h1, H}, e.W(net, GG_P_TR1_W),
                  Epi().bias(e.P(net, GG_P_TR1_B)).act(GG_ACT_LEAKY, c.slope).obf(t.h2, H).of32(t.h2f, H)));
  if (!defer_score) GG_TRY(critic_scores(e, npass, st));
  return GG_OK;
}

// Critic forward on [fake; real] + interpolated rows and the gradient-penalty value through the Gram
// matrix of W1x. Leaves u2, u1, y, dv1, ru1, norms, pen and the loss stats behind.
// `fake_lane`: lane that is producing the fake rows of xfr (joined before the trunk reads them); the
// critic tower (conditioning only) and the Gram matrix (weights only) do not wait for it.
static int disc_forward_gp(gg_engine& e, int R, float p, const float* alpha, int fake_lane, int gp_lane = 0,
                           int save_reps = 0, const float* fake_f32 = nullptr, const float* real_f32 = nullptr) {
  const gg_model_cfg& c = e.cfg;
  TrunkBufs& t = e.tb;
  const int B = c.B, H = c.H, G = c.G, net = GG_NET_DISC;
  const float inv_b = 1.f / static_cast<float>(B);
  const Op W1x = e.W(net, GG_P_TR0_W), W2 = e.W(net, GG_P_TR1_W);
  GG_TRY(e.fork(2));
  GG_TRY(e.mm(2, H, H, G, W1x, 0, W1x, 0, Epi().of32(t.Mg, H).obf(t.Mgb, H)));
  if (e.cond) GG_TRY(tower_forward(e, net, R, p, 0, save_reps));
  GG_TRY(e.join(fake_lane));
  GG_TRY(critic_trunk_forward(e, e.xfr, 2, 3, R, alpha, fake_f32, real_f32, /*defer_score=*/gp_lane != 0));
  // The penalty's own chain (u2 -> u1 -> y = u1 M -> row norms -> losses) feeds only the loss statistics and the
  // GP weight gradients: in the training step it runs on `gp_lane` next to the backward chain of lane 0.
  if (gp_lane != 0) GG_TRY(e.fork(gp_lane));
  cudaStream_t st = e.S(gp_lane);
  if (gp_lane != 0) GG_TRY(critic_scores(e, 3, st));  // D(fake), D(real), D(x_hat): loss statistics only
  const float* w3 = e.P(net, GG_P_FIN_W);
  const bf16* h1i = t.h1 + static_cast<int64_t>(2) * B * H;
  const bf16* h2i = t.h2 + static_cast<int64_t>(2) * B * H;
  GG_TRY(k_gp_u2(h2i, w3, t.u2, B, H, c.slope, st));
  GG_TRY(e.dgrad(gp_lane, B, H, H, Op{t.u2, H}, W2, Epi().mask(h1i, H, 1.f, c.slope).obf(t.u1b, H).of32(t.u1f, H)));
  if (gp_lane != 0) GG_TRY(e.wait_lane(gp_lane, 2));
  else GG_TRY(e.join(2));
  GG_TRY(e.linear(gp_lane, B, H, H, Op{t.u1b, H}, Op{t.Mgb, H}, Epi().of32(t.y, H)));
  GG_TRY(k_gp_rows(t.y, t.u1f, h1i, t.norms, t.pen, t.ru1, t.dv1, B, H, c.slope, c.gp_weight, inv_b, st));
  GG_TRY(k_disc_losses(t.score, t.pen, e.stats, B, c.gp_weight, inv_b, st));
  return GG_OK;
}

}  // namespace gg

// =============================================================================== C ABI
extern "C" int gg_engine_workspace_bytes(const gg_model_cfg* cfg, int64_t* bytes) {
  GG_REQUIRE(cfg && bytes, "null argument");
  GG_TRY(validate_cfg(*cfg));
  gg_engine tmp;
  tmp.cfg = *cfg;
  for (int n = 0; n < 2; ++n)
    for (int s = 0; s < GG_NSLOTS; ++s) tmp.nets[n].off[s] = 0;  // assume every tensor present (upper bound)
  derive(tmp);
  *bytes = layout(tmp, nullptr);
  return GG_OK;
}

extern "C" int gg_engine_create(const gg_model_cfg* cfg, const gg_net_buffers* gen, const gg_net_buffers* disc,
                                void* workspace, int64_t workspace_bytes, void* stream, gg_engine** out) {
  GG_REQUIRE(cfg && gen && disc && workspace && out, "null argument");
  GG_REQUIRE(gen->step_count && disc->step_count, "net buffers need a step_count");
  GG_TRY(validate_cfg(*cfg));
  int dev = 0;
  GG_CUDA_CHECK(cudaGetDevice(&dev));
  GG_TRY(gg_check_device(dev));
  gg_engine* e = new gg_engine();
  e->cfg = *cfg;
  e->nets[GG_NET_GEN] = *gen;
  e->nets[GG_NET_DISC] = *disc;
  derive(*e);
  GG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  const int64_t need = layout(*e, reinterpret_cast<uint8_t*>(workspace));
  if (need > workspace_bytes) {
    set_error("workspace too small: need %lld bytes, have %lld", (long long)need, (long long)workspace_bytes);
    delete e;
    return GG_ERR_WORKSPACE;
  }
  e->total_bytes = need;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  e->cur[0] = st;
  {
    const char* ml = getenv("GEMMGAN_LANES");
    e->multi_lane = !(ml && ml[0] == '1' && ml[1] == 0);
    const char* gr = getenv("GEMMGAN_GROUP_GRADS");
    e->group_grads = !(gr && gr[0] == '0');
    if (const char* tlv = getenv("GEMMGAN_TEXT_LANE")) {
      e->text_lane_fwd = tlv[0] == '1';
      e->text_lane_bwd = tlv[0] && tlv[1] == '1';
    }
    if (const char* sf = getenv("GEMMGAN_SPLIT_TAIL_FLUSH")) e->split_tail_flush = sf[0] != '0';
    const char* fb = getenv("GEMMGAN_FUSE_BIAS");
    e->fuse_bias = !(fb && fb[0] == '0');
    const char* fl = getenv("GEMMGAN_FUSED_LAYER");
    e->fused_layer = !(fl && fl[0] == '0');
    e->layer_bits = layer_bits_requested();
    const char* gl = getenv("GEMMGAN_GEMM_LN");
    e->gemm_ln = !(gl && gl[0] == '0');
    const char* ff = getenv("GEMMGAN_FILM_FUSED");
    e->film_fused = !(ff && ff[0] == '0');
    const char* t1 = getenv("GEMMGAN_T1_SHORTCUT");
    e->t1_shortcut = !(t1 && t1[0] == '0');
    const char* ab = getenv("GEMMGAN_ATTN_BITS");
    e->attn_bits = !(ab && ab[0] == '0');
  }
  for (int l = 1; l < gg_engine::NLANES; ++l)
    GG_CUDA_CHECK(cudaStreamCreateWithFlags(&e->cur[l], cudaStreamNonBlocking));
  for (int i = 0; i < gg_engine::NEVENTS; ++i)
    GG_CUDA_CHECK(cudaEventCreateWithFlags(&e->evs[i], cudaEventDisableTiming));
  for (int n = 0; n < 2; ++n) {
    NetShadow& s = e->sh[n];
    if (!s.segs.empty())
      GG_CUDA_CHECK(cudaMemcpyAsync(s.segs_dev, s.segs.data(), s.segs.size() * sizeof(ShadowSeg),
                                    cudaMemcpyHostToDevice, st));
  }
  GG_CUDA_CHECK(cudaStreamSynchronize(st));  // segs vectors stay alive, but keep create() simple and safe
  if (e->img)
    for (int n = 0; n < 2; ++n)
      GG_CUDA_CHECK(cudaMemsetAsync(e->tw[n].zero, 0, static_cast<size_t>(cfg->B) * cfg->P * cfg->E * sizeof(bf16), st));
  GG_CUDA_CHECK(cudaMemsetAsync(e->wg_ws, 0, GROUP_COUNTER_BYTES, st));
  GG_CUDA_CHECK(cudaMemsetAsync(e->cs_ws, 0, GROUP_COUNTER_BYTES, st));
  GG_CUDA_CHECK(cudaMemsetAsync(e->stats, 0, GG_STATS_COUNT * sizeof(float), st));
  GG_CUDA_CHECK(cudaMemsetAsync(e->opt_step[0], 0, 4 * sizeof(float), st));
  GG_CUDA_CHECK(cudaMemsetAsync(e->opt_step[1], 0, 4 * sizeof(float), st));
  uint64_t rng0[2] = {cfg->seed, 0};
  GG_CUDA_CHECK(cudaMemcpyAsync(e->rng, rng0, sizeof(rng0), cudaMemcpyHostToDevice, st));
  GG_CUDA_CHECK(cudaStreamSynchronize(st));
  for (int n = 0; n < 2; ++n) GG_TRY(gg_engine_refresh_shadows(e, n, stream));
  *out = e;
  return GG_OK;
}

extern "C" void gg_engine_destroy(gg_engine* e) {
  if (!e) return;
  for (int l = 1; l < gg_engine::NLANES; ++l)
    if (e->cur[l]) cudaStreamDestroy(e->cur[l]);
  for (int i = 0; i < gg_engine::NEVENTS; ++i) cudaEventDestroy(e->evs[i]);
  delete e;
}

extern "C" int gg_engine_set_lanes(gg_engine* e, int enabled) {
  GG_REQUIRE(e, "null engine");
  e->multi_lane = enabled != 0;
  return GG_OK;
}

extern "C" int gg_engine_refresh_shadows(gg_engine* e, int net, void* stream) {
  GG_REQUIRE(e && (net == 0 || net == 1), "bad argument");
  NetShadow& s = e->sh[net];
  return k_refresh_shadows(e->nets[net].params, s.base, s.segs_dev, static_cast<int>(s.segs.size()), 0,
                           reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gg_engine_set_batch(gg_engine* e, const float* genes, const float* patches, const uint8_t* patch_pad,
                                   const float* text, const uint8_t* text_pad, void* stream) {
  GG_REQUIRE(e, "null engine");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const gg_model_cfg& c = e->cfg;
  if (genes)  // real genes -> second half of the [fake; real] matrix
    GG_TRY(k_cast_f32_bf16(genes, c.G, e->xfr + static_cast<int64_t>(c.B) * e->Gp, e->Gp, c.B, c.G, st));
  if (e->label) {
    // labels arrive through gg_engine_set_labels
  } else if (e->concat) {
    GG_REQUIRE(text, "the concat variant needs the conditioning vector (text embedding or masked mean patch)");
    GG_TRY(k_cast_f32_bf16(text, c.Dt, e->text, c.Dt, static_cast<int64_t>(c.B) * c.T, c.Dt, st));
  } else if (e->cond) {
    GG_REQUIRE(patches && (text || e->img), "conditional variants need patches and text");
    GG_TRY(k_cast_f32_bf16(patches, c.Dp, e->patches, c.Dp, static_cast<int64_t>(c.B) * c.P, c.Dp, st));
    if (text) GG_TRY(k_cast_f32_bf16(text, c.Dt, e->text, c.Dt, static_cast<int64_t>(c.B) * c.T, c.Dt, st));
    if (patch_pad && e->attn) {
      GG_CUDA_CHECK(cudaMemcpyAsync(e->mask_s, patch_pad, static_cast<size_t>(c.B) * c.P, cudaMemcpyDeviceToDevice, st));
    } else if (patch_pad) {
      GG_TRY(k_mask_with_cls(patch_pad, e->mask_s, c.B, c.P, st));
    } else {
      GG_CUDA_CHECK(cudaMemsetAsync(e->mask_s, 0, static_cast<size_t>(c.B) * e->S_, st));
    }
    e->has_tpad = text_pad != nullptr && e->paper;
    if (e->has_tpad)
      GG_CUDA_CHECK(cudaMemcpyAsync(e->tpad, text_pad, static_cast<size_t>(c.B) * c.T, cudaMemcpyDeviceToDevice, st));
  }
  return GG_OK;
}

extern "C" int gg_engine_set_batchnorm(gg_engine* e, float* running_mean, float* running_var, float momentum, float eps) {
  GG_REQUIRE(e && running_mean && running_var, "null argument");
  GG_REQUIRE(e->attn, "gg_engine_set_batchnorm is for GG_VARIANT_ATTN engines");
  GG_REQUIRE(momentum >= 0.f && momentum <= 1.f && eps > 0.f, "bad BatchNorm momentum / eps");
  e->bn_run_mean = running_mean;
  e->bn_run_var = running_var;
  e->bn_momentum = momentum;
  e->bn_eps = eps;
  return GG_OK;
}

extern "C" int gg_engine_set_labels(gg_engine* e, const int64_t* labels0, const int64_t* labels1, void* stream) {
  GG_REQUIRE(e && labels0 && labels1, "null argument");
  GG_REQUIRE(e->label, "gg_engine_set_labels is for GG_VARIANT_LABEL engines");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t bytes = sizeof(int64_t) * static_cast<size_t>(e->cfg.B);
  GG_CUDA_CHECK(cudaMemcpyAsync(e->labels, labels0, bytes, cudaMemcpyDeviceToDevice, st));
  GG_CUDA_CHECK(cudaMemcpyAsync(e->labels + e->cfg.B, labels1, bytes, cudaMemcpyDeviceToDevice, st));
  return GG_OK;
}

// phase 0: whole step; 1: forward + trunk backward (trunk gradients final on return); 2: tower backward
static int disc_grads_impl(gg_engine* e, const float* z, const float* alpha, int training, int phase, void* stream) {
  GG_REQUIRE(e && z && alpha, "null argument");
  const bool gen_eval = (training & GG_TRAIN_GEN_EVAL) != 0;  // the generator was left in eval mode (see gemmgan.h)
  training &= 1;
  const bool no_join = (phase & GG_PHASE_NO_JOIN) != 0;  // side lanes stay open: a later call joins them
  phase &= ~GG_PHASE_NO_JOIN;
  GG_REQUIRE((phase >= 0 && phase <= 2) || (phase >= GG_PHASE_STAGE0 && phase < GG_PHASE_STAGE0 + 16), "bad phase %d", phase);
  e->begin(stream);
  cudaStream_t st = e->S(0);
  const gg_model_cfg& c = e->cfg;
  TrunkBufs& t = e->tb;
  const int B = c.B, H = c.H, G = c.G, E = c.E, net = GG_NET_DISC;
  const float p = (training && e->cond && !e->concat) ? c.dropout_p : 0.f;
  const int R = p > 0.f ? 3 : 1;   // independently-dropped tower passes: fake, real, interpolated
  const int Rg = p > 0.f ? 2 : 1;  // replicas that carry gradient (the GP's tower gradient is zero)
  const float inv_b = 1.f / static_cast<float>(B);
  if (phase >= 2) {
    const int s0 = phase == 2 ? 0 : phase - GG_PHASE_STAGE0;
    if (e->cond) GG_TRY(tower_backward(*e, net, Rg, p, e->gs.dc, s0, phase == 2 ? (1 << 20) : s0 + 1));
    return no_join ? GG_OK : e->join_all();
  }
  GG_TRY(k_bump_rng(e->rng, st));
  // ---- forward: G(z) (no graph) on lane 1 next to the critic tower on lane 0; D on fake / real /
  // interpolated rows (:391-408), GP value (:351-374)
  GG_TRY(e->fork(1));
  e->bn_training = gen_eval ? 0 : training;
  GG_TRY(gen_forward(*e, z, gen_eval ? 0.f : p, nullptr, 1, false));      // (no graph: the generator is not trained here)
  GG_TRY(disc_forward_gp(*e, R, p, alpha, 1, 1, Rg));  // the GP chain continues on lane 1
  const Op W1x = e->W(net, GG_P_TR0_W), W2 = e->W(net, GG_P_TR1_W);
  const float* w3 = e->P(net, GG_P_FIN_W);
  const bf16* h2i = t.h2 + static_cast<int64_t>(2) * B * H;
  float* gw3 = e->Gr(net, GG_P_FIN_W);
  // ---- the chain towards the tower backward first (lane 0): loss_real + loss_fake on the fake / real rows
  GG_TRY(k_score_bwd(t.h2, w3, t.da2, t.roww, 2 * B, B, H, c.slope, +1.f, -1.f, inv_b, st));
  GG_TRY(e->dgrad(0, 2 * B, H, H, Op{t.da2, H}, W2, Epi().mask(t.h1, H, 1.f, c.slope).obf(t.da1, H)));
  const int64_t ldw1 = static_cast<int64_t>(G) + (e->cond ? E : 0);
  float* gW1 = e->Gr(net, GG_P_TR0_W);
  const Op da1_f{t.da1, H}, da1_r{t.da1 + static_cast<int64_t>(B) * H, H};
  Op cv, W1c;
  if (e->cond) {
    cv = cond_vec(*e, net);
    W1c = Op{e->sh[net].tr0_c.p, e->sh[net].tr0_c.ld};
    if (Rg == 2) {
      GG_TRY(e->dgrad(0, 2 * B, E, H, da1_f, W1c, Epi().obf(e->gs.dc, E)));
    } else {  // one shared tower pass: both row groups meet the same conditioning vectors
      GG_TRY(e->mm(0, B, E, H, da1_f, 0, W1c, 1, Epi().obf(e->gs.dc, E), H, da1_r, W1c));
    }
  }
  // ---- everything that only feeds the optimizer goes to lane 1 (GEMMs) / lane 2 (column sums)
  GG_TRY(e->fork(1));
  // GP: Q = (r u1)^T u1 ; du2 = dv1 W2^T masked by m2 -> gradient of the GP w.r.t. w3
  GG_TRY(e->mm(1, H, H, B, Op{t.ru1, H}, 1, Op{t.u1b, H}, 1, Epi().obf(t.Qb, H)));
  GG_TRY(e->linear(1, B, H, H, Op{t.dv1, H}, W2, Epi().mask(h2i, H, 1.f, c.slope).of32(t.du2f, H)));
  GG_TRY(k_colsum(t.h2f, 1, H, 2 * B, H, t.roww, 1.f, gw3, 0, e->scratch_l[e->L(1)], e->S(1)));
  GG_TRY(k_colsum(t.du2f, 1, H, B, H, nullptr, 1.f, gw3, 1, e->scratch_l[e->L(1)], e->S(1)));
  GG_TRY(k_fill_f32(e->Gr(net, GG_P_FIN_B), 0.f, 1, e->S(1)));  // d/db3 of mean(D(fake)) - mean(D(real)) is exactly 0
  // dW2 = da2^T h1(fake,real)  +  u2^T dv1 (GP)
  GG_TRY(e->mm(1, H, H, 2 * B, Op{t.da2, H}, 1, Op{t.h1, H}, 1, Epi().of32(e->Gr(net, GG_P_TR1_W), H), B,
               Op{t.u2, H}, Op{t.dv1, H}));
  // dW1[:, :G] = da1^T [fake; real]  +  Q W1x (GP, second K-segment)
  GG_TRY(e->mm(1, H, G, 2 * B, Op{t.da1, H}, 1, Op{e->xfr, e->Gp}, 1, Epi().of32(gW1, ldw1), H, Op{t.Qb, H}, W1x));
  if (e->cond) {
    if (Rg == 2) {
      GG_TRY(e->wgrad(H, E, 2 * B, da1_f, cv, gW1 + G, ldw1));
    } else {
      GG_TRY(e->mm(1, H, E, B, da1_f, 1, cv, 1, Epi().of32(gW1 + G, ldw1), B, da1_r, cv));
    }
  }
  GG_TRY(e->bgrad(t.da2, H, 2 * B, H, e->Gr(net, GG_P_TR1_B)));
  GG_TRY(e->bgrad(t.da1, H, 2 * B, H, e->Gr(net, GG_P_TR0_B)));
  GG_TRY(e->flush_grads());
  if (phase == 0 && e->cond) GG_TRY(tower_backward(*e, net, Rg, p, e->gs.dc));
  return no_join ? GG_OK : e->join_all();
}

extern "C" int gg_engine_disc_grads(gg_engine* e, const float* z, const float* alpha, int training, void* stream) {
  return disc_grads_impl(e, z, alpha, training, 0, stream);
}
extern "C" int gg_engine_disc_grads_phase(gg_engine* e, const float* z, const float* alpha, int training, int phase,
                                          void* stream) {
  GG_REQUIRE(phase != 0, "phase 0 is gg_engine_disc_grads");
  return disc_grads_impl(e, z, alpha, training, phase, stream);
}

// Generator trunk backward from d loss / d fake (t.dfake, bf16 [B, Gp]): every trunk gradient, gs.dc for the tower, and
// (optionally) the gradient w.r.t. the latent input z (fp32 [B, L])
static int gen_trunk_backward(gg_engine& e, float* dz_f32) {
  const gg_model_cfg& c = e.cfg;
  TrunkBufs& t = e.tb;
  const int B = c.B, H = c.H, G = c.G, E = c.E, L = c.L, Gn = GG_NET_GEN;
  const int Lp = static_cast<int>(round_up64(L, 8));
  const Op Wf = e.W(Gn, GG_P_FIN_W), Wg2 = e.W(Gn, GG_P_TR1_W);
  GG_TRY(e.wgrad(G, H, B, Op{t.dfake, e.Gp}, Op{t.hg2, H}, e.Gr(Gn, GG_P_FIN_W), H));
  GG_TRY(e.bgrad(t.dfake, e.Gp, B, G, e.Gr(Gn, GG_P_FIN_B)));
  GG_TRY(e.dgrad(0, B, H, G, Op{t.dfake, e.Gp}, Wf, Epi().mask(t.hg2, H, 1.f, c.slope).obf(t.dag2, H)));
  GG_TRY(e.wgrad(H, H, B, Op{t.dag2, H}, Op{t.hg1, H}, e.Gr(Gn, GG_P_TR1_W), H));
  GG_TRY(e.bgrad(t.dag2, H, B, H, e.Gr(Gn, GG_P_TR1_B)));
  GG_TRY(e.dgrad(0, B, H, H, Op{t.dag2, H}, Wg2, Epi().mask(t.hg1, H, 1.f, c.slope).obf(t.dag1, H)));
  const int64_t ldw = static_cast<int64_t>(L) + (e.cond ? E : 0);
  float* gW1 = e.Gr(Gn, GG_P_TR0_W);
  if (e.cond) {
    const Op Wc{e.sh[Gn].tr0_c.p, e.sh[Gn].tr0_c.ld};
    GG_TRY(e.dgrad(0, B, E, H, Op{t.dag1, H}, Wc, Epi().obf(e.gs.dc, E)));
  }
  if (dz_f32) GG_TRY(e.dgrad(0, B, L, H, Op{t.dag1, H}, e.W(Gn, GG_P_TR0_W), Epi().of32(dz_f32, L)));
  GG_TRY(e.wgrad(H, L, B, Op{t.dag1, H}, Op{e.zbf, Lp}, gW1, ldw));
  GG_TRY(e.bgrad(t.dag1, H, B, H, e.Gr(Gn, GG_P_TR0_B)));
  if (e.cond) {
    const Op cv = cond_vec(e, Gn);
    GG_TRY(e.wgrad(H, E, B, Op{t.dag1, H}, cv, gW1 + L, ldw));
  }
  return e.flush_grads();
}

static int gen_grads_impl(gg_engine* e, const float* z, int training, int phase, void* stream) {
  GG_REQUIRE(e && z, "null argument");
  const bool no_join = (phase & GG_PHASE_NO_JOIN) != 0;
  phase &= ~GG_PHASE_NO_JOIN;
  GG_REQUIRE((phase >= 0 && phase <= 2) || (phase >= GG_PHASE_STAGE0 && phase < GG_PHASE_STAGE0 + 16), "bad phase %d", phase);
  e->begin(stream);
  cudaStream_t st = e->S(0);
  const gg_model_cfg& c = e->cfg;
  TrunkBufs& t = e->tb;
  const int B = c.B, H = c.H, G = c.G, E = c.E, L = c.L;
  const int Lp = static_cast<int>(round_up64(L, 8));
  const float p = (training && e->cond && !e->concat) ? c.dropout_p : 0.f;
  const float inv_b = 1.f / static_cast<float>(B);
  const int D = GG_NET_DISC, Gn = GG_NET_GEN;
  if (phase >= 2) {
    const int s0 = phase == 2 ? 0 : phase - GG_PHASE_STAGE0;
    if (e->cond) GG_TRY(tower_backward(*e, Gn, 1, p, e->gs.dc, s0, phase == 2 ? (1 << 20) : s0 + 1));
    return no_join ? GG_OK : e->join_all();
  }
  GG_TRY(k_bump_rng(e->rng, st));
  // ---- forward: fake = G(z) on lane 0, the critic's tower (conditioning only) next to it on lane 1,
  // then D(fake) (:441-452)
  if (e->cond) {
    GG_TRY(e->fork(1));
    GG_TRY(tower_forward(*e, D, 1, p, 1, 0));  // (the critic is frozen: nothing of its tower is back-propagated)
  }
  GG_REQUIRE(training || !e->attn, "the BatchNorm backward of the attention variant is the training-mode one");
  e->bn_training = training;
  GG_TRY(gen_forward(*e, z, p, nullptr, 0, true));
  GG_TRY(e->join(1));
  GG_TRY(critic_trunk_forward(*e, e->xfr, 1, 1, 1, nullptr, nullptr, nullptr, /*defer_score=*/true));
  GG_TRY(e->fork(1));  // -mean D(G(z)) is a statistic: the backward starts from h2's signs
  GG_TRY(critic_scores(*e, 1, e->S(1)));
  GG_TRY(k_gen_loss(t.score, e->stats, B, inv_b, e->S(1)));
  // ---- critic trunk input gradient (critic weights frozen, its tower is not on G's path)
  const Op W1x = e->W(D, GG_P_TR0_W), W2 = e->W(D, GG_P_TR1_W);
  GG_TRY(k_score_bwd(t.h2, e->P(D, GG_P_FIN_W), t.da2, nullptr, B, B, H, c.slope, -1.f, -1.f, inv_b, st));
  GG_TRY(e->dgrad(0, B, H, H, Op{t.da2, H}, W2, Epi().mask(t.h1, H, 1.f, c.slope).obf(t.da1, H)));
  GG_TRY(e->dgrad(0, B, G, H, Op{t.da1, H}, W1x, Epi().obf(t.dfake, e->Gp)));
  GG_TRY(gen_trunk_backward(*e, nullptr));
  if (e->cond && phase == 0) GG_TRY(tower_backward(*e, Gn, 1, p, e->gs.dc));
  return no_join ? GG_OK : e->join_all();
}

// Makes `stream` (the trainer's communication stream) wait for everything enqueued so far on the engine's side
// lanes — the weight- / bias-gradient work of the stages run with GG_PHASE_NO_JOIN — without joining them into the
// caller's stream: the all-reduce of a finished bucket then starts behind its producers while lane 0 moves on.
extern "C" int gg_engine_lanes_signal(gg_engine* e, void* stream) {
  GG_REQUIRE(e && stream, "null argument");
  cudaStream_t dst = reinterpret_cast<cudaStream_t>(stream);
  for (int l = 1; l < gg_engine::NLANES; ++l) {
    if (!e->forked[e->L(l)] || e->L(l) == 0) continue;
    cudaEvent_t ev = e->evs[e->ev_next];
    e->ev_next = (e->ev_next + 1) % gg_engine::NEVENTS;
    GG_CUDA_CHECK(cudaEventRecord(ev, e->cur[l]));
    GG_CUDA_CHECK(cudaStreamWaitEvent(dst, ev, 0));
  }
  return GG_OK;
}

extern "C" int gg_engine_gen_grads(gg_engine* e, const float* z, int training, void* stream) {
  return gen_grads_impl(e, z, training, 0, stream);
}
extern "C" int gg_engine_gen_grads_phase(gg_engine* e, const float* z, int training, int phase, void* stream) {
  GG_REQUIRE(phase != 0, "phase 0 is gg_engine_gen_grads");
  return gen_grads_impl(e, z, training, phase, stream);
}

extern "C" int gg_engine_optim_step(gg_engine* e, int net, float lr, void* stream) {
  GG_REQUIRE(e && (net == 0 || net == 1), "bad argument");
  e->begin(stream);
  cudaStream_t st = e->S(0);
  const gg_net_buffers& nb = e->nets[net];
  const float max_norm = net == GG_NET_DISC ? e->cfg.clip_d : e->cfg.clip_g;
  const float* coef = nullptr;
  float* statp = e->stats + (net == GG_NET_DISC ? GG_STAT_D_GRAD_NORM : GG_STAT_G_GRAD_NORM);
  if (max_norm > 0.f) {
    GG_TRY(k_grad_norm_clip(nb.grads, nb.n_used, max_norm, statp, e->scratch_l[0], st));
    coef = statp + 1;
  }
  GG_TRY(k_optim_step(e->cfg.optimizer, nb.params, nb.grads, nb.exp_avg, nb.exp_avg_sq, nb.n_used, lr, coef,
                      nb.step_count, st, /*bump_step=*/false));
  NetShadow& sh = e->sh[net];  // bf16 shadows of the new weights + the step-counter increment, one launch
  return k_refresh_shadows(nb.params, sh.base, sh.segs_dev, static_cast<int>(sh.segs.size()), 0, st, nb.step_count);
}

extern "C" int gg_engine_generate(gg_engine* e, const float* z, float* out_f32, int training, void* stream) {
  GG_REQUIRE(e && z && out_f32, "null argument");
  e->begin(stream);
  const float p = (training && e->cond && !e->concat) ? e->cfg.dropout_p : 0.f;
  if (p > 0.f) GG_TRY(k_bump_rng(e->rng, e->S(0)));
  e->bn_training = training;
  return gen_forward(*e, z, p, out_f32, 0, false);
}

// ---- module-level autograd (SURVEY.md section 8 b "who calls it": generator.forward / discriminator.forward as
// free-standing differentiable modules). *_keep = the same forward, keeping what the backward reads; *_backward = the
// first-order backward from an arbitrary upstream gradient (the trainers' fused steps never come through here).
extern "C" int gg_engine_generate_keep(gg_engine* e, const float* z, float* out_f32, int training, void* stream) {
  GG_REQUIRE(e && z && out_f32, "null argument");
  e->begin(stream);
  const float p = (training && e->cond && !e->concat) ? e->cfg.dropout_p : 0.f;
  if (p > 0.f) GG_TRY(k_bump_rng(e->rng, e->S(0)));
  e->bn_training = training;
  e->kept_p[GG_NET_GEN] = p;
  e->kept_training[GG_NET_GEN] = training;
  return gen_forward(*e, z, p, out_f32, 0, true);
}

extern "C" int gg_engine_generate_backward(gg_engine* e, const float* dout_f32, float* dz_f32, void* stream) {
  GG_REQUIRE(e && dout_f32, "null argument");
  GG_REQUIRE(e->kept_p[GG_NET_GEN] >= 0.f, "gg_engine_generate_backward without a gg_engine_generate_keep forward");
  GG_REQUIRE(e->kept_training[GG_NET_GEN] || !e->attn, "the BatchNorm backward is the training-mode one");
  e->begin(stream);
  const gg_model_cfg& c = e->cfg;
  GG_TRY(k_cast_f32_bf16(dout_f32, c.G, e->tb.dfake, e->Gp, c.B, c.G, e->S(0)));
  GG_TRY(gen_trunk_backward(*e, dz_f32));
  if (e->cond) GG_TRY(tower_backward(*e, GG_NET_GEN, 1, e->kept_p[GG_NET_GEN], e->gs.dc));
  return e->join_all();
}

extern "C" int gg_engine_critic_keep(gg_engine* e, const float* genes_f32, float* score_f32, int training, void* stream) {
  GG_REQUIRE(e && genes_f32 && score_f32, "null argument");
  e->begin(stream);
  cudaStream_t st = e->S(0);
  const gg_model_cfg& c = e->cfg;
  const float p = (training && e->cond && !e->concat) ? c.dropout_p : 0.f;
  if (p > 0.f) GG_TRY(k_bump_rng(e->rng, st));
  e->kept_p[GG_NET_DISC] = p;
  GG_TRY(k_cast_f32_bf16(genes_f32, c.G, e->xin, e->Gp, c.B, c.G, st));
  if (e->cond) GG_TRY(tower_forward(*e, GG_NET_DISC, 1, p, 0, 1));
  GG_TRY(critic_trunk_forward(*e, e->xin, 1, 1, 1, nullptr));
  GG_CUDA_CHECK(cudaMemcpyAsync(score_f32, e->tb.score, sizeof(float) * c.B, cudaMemcpyDeviceToDevice, st));
  return GG_OK;
}

extern "C" int gg_engine_critic_backward(gg_engine* e, const float* dscore_f32, float* dgenes_f32, void* stream) {
  GG_REQUIRE(e && dscore_f32, "null argument");
  GG_REQUIRE(e->kept_p[GG_NET_DISC] >= 0.f, "gg_engine_critic_backward without a gg_engine_critic_keep forward");
  e->begin(stream);
  cudaStream_t st = e->S(0);
  const gg_model_cfg& c = e->cfg;
  TrunkBufs& t = e->tb;
  const int B = c.B, H = c.H, G = c.G, E = c.E, net = GG_NET_DISC;
  const Op W1x = e->W(net, GG_P_TR0_W), W2 = e->W(net, GG_P_TR1_W);
  // score = h2 . w3 + b3
  GG_TRY(k_score_bwd_rows(t.h2, e->P(net, GG_P_FIN_W), dscore_f32, t.da2, B, H, c.slope, st));
  GG_TRY(e->dgrad(0, B, H, H, Op{t.da2, H}, W2, Epi().mask(t.h1, H, 1.f, c.slope).obf(t.da1, H)));
  const int64_t ldw1 = static_cast<int64_t>(G) + (e->cond ? E : 0);
  float* gW1 = e->Gr(net, GG_P_TR0_W);
  if (e->cond) {
    const Op W1c{e->sh[net].tr0_c.p, e->sh[net].tr0_c.ld};
    GG_TRY(e->dgrad(0, B, E, H, Op{t.da1, H}, W1c, Epi().obf(e->gs.dc, E)));
  }
  if (dgenes_f32) GG_TRY(e->dgrad(0, B, G, H, Op{t.da1, H}, W1x, Epi().of32(dgenes_f32, G)));
  GG_TRY(e->fork(1));
  GG_TRY(k_colsum(t.h2f, 1, H, B, H, dscore_f32, 1.f, e->Gr(net, GG_P_FIN_W), 0, e->scratch_l[e->L(1)], e->S(1)));
  GG_TRY(k_colsum(dscore_f32, 1, 1, B, 1, nullptr, 1.f, e->Gr(net, GG_P_FIN_B), 0, e->scratch_l[e->L(1)], e->S(1)));
  GG_TRY(e->wgrad(H, H, B, Op{t.da2, H}, Op{t.h1, H}, e->Gr(net, GG_P_TR1_W), H));
  GG_TRY(e->bgrad(t.da2, H, B, H, e->Gr(net, GG_P_TR1_B)));
  GG_TRY(e->wgrad(H, G, B, Op{t.da1, H}, Op{e->xin, e->Gp}, gW1, ldw1));
  GG_TRY(e->bgrad(t.da1, H, B, H, e->Gr(net, GG_P_TR0_B)));
  if (e->cond) GG_TRY(e->wgrad(H, E, B, Op{t.da1, H}, cond_vec(*e, net), gW1 + G, ldw1));
  GG_TRY(e->flush_grads());
  if (e->cond) GG_TRY(tower_backward(*e, net, 1, e->kept_p[net], e->gs.dc));
  return e->join_all();
}

extern "C" int gg_engine_critic(gg_engine* e, const float* genes_f32, float* score_f32, int training, void* stream) {
  GG_REQUIRE(e && genes_f32 && score_f32, "null argument");
  e->begin(stream);
  cudaStream_t st = e->S(0);
  const gg_model_cfg& c = e->cfg;
  const float p = (training && e->cond && !e->concat) ? c.dropout_p : 0.f;
  if (p > 0.f) GG_TRY(k_bump_rng(e->rng, st));
  GG_TRY(k_cast_f32_bf16(genes_f32, c.G, e->xin, e->Gp, c.B, c.G, st));
  if (e->cond) GG_TRY(tower_forward(*e, GG_NET_DISC, 1, p, 0, 0));
  GG_TRY(critic_trunk_forward(*e, e->xin, 1, 1, 1, nullptr));
  GG_CUDA_CHECK(cudaMemcpyAsync(score_f32, e->tb.score, sizeof(float) * c.B, cudaMemcpyDeviceToDevice, st));
  return GG_OK;
}

extern "C" int gg_engine_gradient_penalty(gg_engine* e, const float* real_f32, const float* fake_f32,
                                          const float* alpha, int training, float* gp_out, void* stream) {
  GG_REQUIRE(e && fake_f32 && alpha && gp_out, "null argument");
  e->begin(stream);
  cudaStream_t st = e->S(0);
  const gg_model_cfg& c = e->cfg;
  const float p = (training && e->cond && !e->concat) ? c.dropout_p : 0.f;
  if (p > 0.f) GG_TRY(k_bump_rng(e->rng, st));
  GG_TRY(k_cast_f32_bf16(fake_f32, c.G, e->xfr, e->Gp, c.B, c.G, st));
  if (real_f32)
    GG_TRY(k_cast_f32_bf16(real_f32, c.G, e->xfr + static_cast<int64_t>(c.B) * e->Gp, e->Gp, c.B, c.G, st));
  GG_TRY(disc_forward_gp(*e, p > 0.f ? 3 : 1, p, alpha, 0));
  GG_TRY(e->join_all());
  GG_CUDA_CHECK(cudaMemcpyAsync(gp_out, e->stats + GG_STAT_GP, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return GG_OK;
}

// The gradient penalty alone, value AND gradients (BASELINE.json config 5, SURVEY.md section 8d "GP microbench"):
// x_hat = alpha*real + (1-alpha)*fake, GP = mean((||dD/dx_hat|| - 1)^2) and gp_weight * dGP/d{W1, W2, w3} written
// into the critic's gradient buffer — what autograd computes with a forward, torch.autograd.grad(create_graph)
// and a double backward (reference :351-374 + the GP part of :412). fp32 inputs are cast to bf16 once (8*B*G
// bytes read); no interpolated tensor, no [B,G] gradient tensor: the Gram-matrix formulation.
extern "C" int gg_engine_gp_step(gg_engine* e, const float* real_f32, const float* fake_f32, const float* alpha,
                                 float* gp_out, void* stream) {
  GG_REQUIRE(e && real_f32 && fake_f32 && alpha, "null argument");
  GG_REQUIRE(!e->cond, "gg_engine_gp_step is the unconditional-critic microbenchmark entry point");
  e->begin(stream);
  cudaStream_t st = e->S(0);
  const gg_model_cfg& c = e->cfg;
  TrunkBufs& t = e->tb;
  const int B = c.B, H = c.H, G = c.G, net = GG_NET_DISC;
  // The only [B, G] work of the penalty is critic layer 1 on real and fake. Default (H = 256): xw_f32.cu reads the two fp32
  // tensors in place — 8 B G bytes of HBM traffic in all — converting to bf16 on chip and multicasting the weight k-blocks
  // across a cluster of row tiles (round-2 history: cast pass + bf16 GEMM 1.15 ms at B = 16384, G = 20000; TF32 GEMMs on the
  // fp32 tensors against the fp32 master weights 1.47 ms: every 128-row tile re-streamed the 20 MB fp32 weight matrix from
  // L2 and the L2 -> shared-memory fabric, not HBM, became the bound).
  // GEMMGAN_GP_DIRECT=0: cast both tensors to bf16 first (1.5x the HBM bytes), then the step's bf16 GEMM; GEMMGAN_GP_TF32=1:
  // TF32 GEMMs on the fp32 tensors and fp32 master weights (measured slower: the fp32 weight matrix re-streamed per tile)
  static const bool gp_direct = [] { const char* v = getenv("GEMMGAN_GP_DIRECT"); return !(v && v[0] == '0'); }();
  static const bool gp_tf32 = [] { const char* v = getenv("GEMMGAN_GP_TF32"); return v && v[0] == '1'; }();
  e->gp_tf32 = gp_tf32;
  const bool direct = (gp_tf32 || (gp_direct && H == 256)) && c.gemm_impl == GG_IMPL_TCGEN05 && G % 4 == 0 &&
                      ((reinterpret_cast<uintptr_t>(real_f32) | reinterpret_cast<uintptr_t>(fake_f32)) & 15) == 0;
  if (direct) {
    GG_TRY(disc_forward_gp(*e, 1, 0.f, alpha, 0, 0, 0, fake_f32, real_f32));
  } else {
    GG_TRY(k_cast_f32_bf16(fake_f32, G, e->xfr, e->Gp, B, G, st));
    GG_TRY(k_cast_f32_bf16(real_f32, G, e->xfr + static_cast<int64_t>(B) * e->Gp, e->Gp, B, G, st));
    GG_TRY(disc_forward_gp(*e, 1, 0.f, alpha, 0));
  }
  const Op W1x = e->W(net, GG_P_TR0_W), W2 = e->W(net, GG_P_TR1_W);
  const bf16* h2i = t.h2 + static_cast<int64_t>(2) * B * H;
  float* gw3 = e->Gr(net, GG_P_FIN_W);
  // Q = (r u1)^T u1 ; du2 = dv1 W2^T masked by m2 -> d/dw3 ; dW2 = u2^T dv1 ; dW1 = Q W1x
  GG_TRY(e->mm(0, H, H, B, Op{t.ru1, H}, 1, Op{t.u1b, H}, 1, Epi().obf(t.Qb, H)));
  GG_TRY(e->linear(0, B, H, H, Op{t.dv1, H}, W2, Epi().mask(h2i, H, 1.f, c.slope).of32(t.du2f, H)));
  GG_TRY(k_colsum(t.du2f, 1, H, B, H, nullptr, 1.f, gw3, 0, e->scratch_l[0], st));
  GG_TRY(e->mm(0, H, H, B, Op{t.u2, H}, 1, Op{t.dv1, H}, 1, Epi().of32(e->Gr(net, GG_P_TR1_W), H)));
  GG_TRY(e->mm(0, H, G, H, Op{t.Qb, H}, 0, W1x, 1, Epi().of32(e->Gr(net, GG_P_TR0_W), G)));
  GG_TRY(e->join_all());
  if (gp_out)
    GG_CUDA_CHECK(cudaMemcpyAsync(gp_out, e->stats + GG_STAT_GP, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return GG_OK;
}

// out[b, :] = sum over the non-padded rows p of x[b, p, :] / count_b  (conditional_gan_concat.py:137-138 /
// :184-185 apply the Linear encoder to every patch and then take this masked mean; the encoder is affine, so
// the mean is taken first and the encoder runs once per sample).
extern "C" int gg_xw_f32(const float* x0, const float* x1, int32_t B, int32_t K, const void* w_bf16, int64_t ldw, float* out,
                         void* workspace, int64_t workspace_bytes, void* stream) {
  return k_xw_f32(x0, x1, B, K, reinterpret_cast<const bf16*>(w_bf16), ldw, out, workspace, workspace_bytes,
                  reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gg_film_patch_encode(const void* patches_bf16, const float* gamma_beta, const void* w_bf16, int64_t ldw,
                                    const float* bias, const float* cls, void* x0_bf16, void* mod_bf16, int32_t B, int32_t P,
                                    int32_t R, int32_t Dp, void* stream) {
  return k_film_patch(reinterpret_cast<const bf16*>(patches_bf16), gamma_beta, reinterpret_cast<const bf16*>(w_bf16), ldw, bias,
                      cls, reinterpret_cast<bf16*>(x0_bf16), reinterpret_cast<bf16*>(mod_bf16), B, P, R, Dp,
                      reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gg_gemm_layernorm(const void* a_bf16, int64_t lda, const void* w_bf16, int64_t ldw, int32_t K, const float* bias,
                                 const void* res_bf16, const float* gamma, const float* beta, void* z_bf16, void* out_bf16,
                                 float* mean, float* rstd, int64_t rows, float eps, float drop_p, const uint64_t* rng,
                                 uint32_t site, void* stream) {
  return k_gemm_ln(reinterpret_cast<const bf16*>(a_bf16), lda, reinterpret_cast<const bf16*>(w_bf16), ldw, K, bias,
                   reinterpret_cast<const bf16*>(res_bf16), gamma, beta, reinterpret_cast<bf16*>(z_bf16),
                   reinterpret_cast<bf16*>(out_bf16), mean, rstd, rows, eps, drop_p, rng, site,
                   reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gg_masked_mean_rows(const float* x, const uint8_t* pad, float* out, int B, int P, int D, void* stream) {
  GG_REQUIRE(x && out && B > 0 && P > 0 && D > 0, "bad argument");
  return k_masked_mean_rows(x, pad, out, B, P, D, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gg_gather_rows(const float* src, int64_t ld_src, const int64_t* index, float* dst, int64_t ld_dst,
                              int64_t rows, int32_t cols, void* stream) {
  GG_REQUIRE(src && index && dst && rows >= 0 && cols > 0 && ld_src >= cols && ld_dst >= cols, "gg_gather_rows: bad argument");
  return k_gather_rows(src, ld_src, index, dst, ld_dst, rows, cols, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" float* gg_engine_stats(gg_engine* e) { return e ? e->stats : nullptr; }

extern "C" void* gg_engine_buffer(gg_engine* e, const char* name, int64_t* rows, int64_t* cols, int64_t* ld,
                                  int32_t* is_f32) {
  if (!e || !name) return nullptr;
  const gg_model_cfg& c = e->cfg;
  const std::string s(name);
  void* p = nullptr;
  int64_t r = 0, cc = 0, l = 0;
  int32_t f = 0;
  if (s == "fake_bf16") { p = e->xfr; r = c.B; cc = c.G; l = e->Gp; }
  else if (s == "real_bf16") { p = e->xfr + static_cast<int64_t>(c.B) * e->Gp; r = c.B; cc = c.G; l = e->Gp; }
  else if (s == "score") { p = e->tb.score; r = 3 * c.B; cc = 1; l = 1; f = 1; }
  else if (s == "gp_norms") { p = e->tb.norms; r = c.B; cc = 1; l = 1; f = 1; }
  else if (s == "gp_pen") { p = e->tb.pen; r = c.B; cc = 1; l = 1; f = 1; }
  else if (s == "h1") { p = e->tb.h1; r = 3 * c.B; cc = c.H; l = c.H; }
  else if (s == "h2f") { p = e->tb.h2f; r = 3 * c.B; cc = c.H; l = c.H; f = 1; }
  else if (s == "gram") { p = e->tb.Mg; r = c.H; cc = c.H; l = c.H; f = 1; }
  else if (s == "dfake") { p = e->tb.dfake; r = c.B; cc = c.G; l = e->Gp; }
  else if (s == "da1") { p = e->tb.da1; r = 2 * c.B; cc = c.H; l = c.H; }
  else if (s == "da2") { p = e->tb.da2; r = 2 * c.B; cc = c.H; l = c.H; }
  else if (s == "u2") { p = e->tb.u2; r = c.B; cc = c.H; l = c.H; }
  else if (s == "u1f") { p = e->tb.u1f; r = c.B; cc = c.H; l = c.H; f = 1; }
  else if (s == "y") { p = e->tb.y; r = c.B; cc = c.H; l = c.H; f = 1; }
  else if (s == "dv1") { p = e->tb.dv1; r = c.B; cc = c.H; l = c.H; }
  else if (s == "ru1") { p = e->tb.ru1; r = c.B; cc = c.H; l = c.H; }
  else if (s == "Qb") { p = e->tb.Qb; r = c.H; cc = c.H; l = c.H; }
  else if (s == "a1x") { p = e->tb.a1x; r = 2 * c.B; cc = c.H; l = c.H; f = 1; }
  else if (s == "h2") { p = e->tb.h2; r = 3 * c.B; cc = c.H; l = c.H; }
  else if (e->cond && (s == "cond_gen" || s == "cond_disc")) {
    const int net = s == "cond_gen" ? GG_NET_GEN : GG_NET_DISC;
    const Op cv = cond_vec(*e, net);
    p = const_cast<bf16*>(cv.p); r = static_cast<int64_t>(e->tw[net].Rmax) * c.B; cc = c.E; l = cv.ld;
  } else if (e->cond && s == "film_gb_disc") { p = e->tw[GG_NET_DISC].gb; r = c.B; cc = 2 * c.Dp; l = 2 * c.Dp; f = 1; }
  else if (e->cond && s == "tokens_disc") { p = e->tw[GG_NET_DISC].X[c.n_layers]; r = static_cast<int64_t>(e->tw[GG_NET_DISC].Rmax) * c.B * e->S_; cc = c.E; l = c.E; }
  else if (e->cond && s == "tokens0_disc") { p = e->tw[GG_NET_DISC].X[0]; r = static_cast<int64_t>(e->tw[GG_NET_DISC].Rmax) * c.B * e->S_; cc = c.E; l = c.E; }
  if (rows) *rows = r;
  if (cols) *cols = cc;
  if (ld) *ld = l;
  if (is_f32) *is_f32 = f;
  return p;
}

extern "C" int gg_attention_fwd(const gg_attn_args* a, void* stream) {
  GG_REQUIRE(a && a->q && a->k && a->v && a->o, "null argument");
  return k_attention_fwd(*reinterpret_cast<const AttnArgs*>(a), reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int64_t gg_dropout_bits_words(int64_t n_elems) { return dropout_bits_words(n_elems); }
extern "C" int gg_dropout_bits(const uint64_t* rng, uint32_t site, float p, int64_t n_elems, uint32_t* out, void* stream) {
  return k_dropout_bits(rng, site, p, n_elems, out, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int gg_attention_bwd(const gg_attn_args* a, void* stream) {
  GG_REQUIRE(a && a->q && a->k && a->v && a->dout && a->dq && a->dk && a->dv, "null argument");
  return k_attention_bwd(*reinterpret_cast<const AttnArgs*>(a), reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int64_t gg_wgrad_group_workspace_bytes(int64_t sum_output_elems) {
  return wgrad_group_workspace_bytes(sum_output_elems);
}
extern "C" int gg_wgrad_group(const gg_wgrad_item* items, int n, void* workspace, int64_t workspace_bytes,
                              void* stream) {
  GG_REQUIRE(items && workspace, "null argument");
  return k_wgrad_group(items, n, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int64_t gg_colsum_group_workspace_bytes(int64_t sum_columns) {
  return colsum_group_workspace_bytes(sum_columns);
}
extern "C" int gg_colsum_group(const gg_colsum_item* items, int n, void* workspace, int64_t workspace_bytes,
                               void* stream) {
  GG_REQUIRE(items && workspace, "null argument");
  return k_colsum_group(items, n, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gg_optim_step(int kind, float* p, float* g, float* m, float* v, int64_t n, float lr, float max_norm,
                             float* step_count, float* norm_out2, float* scratch, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float* coef = nullptr;
  if (max_norm > 0.f) {
    GG_REQUIRE(norm_out2 && scratch, "clipping needs norm_out2[2] and scratch[>=592]");
    GG_TRY(k_grad_norm_clip(g, n, max_norm, norm_out2, scratch, st));
    coef = norm_out2 + 1;
  }
  return k_optim_step(kind, p, g, m, v, n, lr, coef, step_count, st);
}
