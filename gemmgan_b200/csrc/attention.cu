// Multi-head attention core for the fusion tower: softmax(q k^T / sqrt(hd) + key_padding_mask) v with
// dropout on the probabilities, forward and backward, one CTA per (batch row, head).
// Sequence lengths here are tiny (S = P+1 <= 257 tokens, T <= 300, single-query cross attention in the
// paper model), so K/V (and Q/dO in the backward) live in shared memory for the whole CTA and the
// probabilities are recomputed in the backward instead of being stored (flash-style, deterministic:
// pass A owns query rows -> dQ, pass B owns key rows -> dK, dV; no atomics).
//
// Replaces F.multi_head_attention_forward / scaled_dot_product_attention and their autograd backward
// used by nn.TransformerEncoderLayer.self_attn and the patch2text / text2patch nn.MultiheadAttention
// modules (src/conditional_gan_cross_attention_with_film.py:114-123, 144-152). As in torch, q is scaled
// by 1/sqrt(head_dim) before q k^T and padded keys get -inf.
#include "host_util.h"
#include "pdl.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace gg {

constexpr int ATT_MAXC = 10;  // keys / queries per lane: sequence length <= 320
constexpr int ATT_WARPS = 4;

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wsum2(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// rows of `n` x hd bf16 from global (pitch ld) into smem with row pitch hd+2 (bank-conflict-free columns)
__device__ __forceinline__ void load_rows(bf16* dst, const bf16* src, int64_t ld, int n, int hd) {
  const int pitch = hd + 2;
  const int half = hd >> 1;
  for (int i = threadIdx.x; i < n * half; i += blockDim.x) {
    const int r = i / half, c = (i % half) * 2;
    *reinterpret_cast<__nv_bfloat162*>(dst + r * pitch + c) =
        *reinterpret_cast<const __nv_bfloat162*>(src + static_cast<int64_t>(r) * ld + c);
  }
}

__device__ __forceinline__ float dot_row(const bf16* a, const bf16* b, int hd) {
  float acc = 0.f;
  for (int d = 0; d < hd; d += 2) {
    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a + d));
    const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(b + d));
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(ATT_WARPS * 32) attention_fwd_kernel(const AttnArgs a) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int hd = a.hd, pitch = hd + 2, Lk = a.Lk, Lq = a.Lq;
  bf16* Ks = reinterpret_cast<bf16*>(smem_att);
  bf16* Vs = Ks + Lk * pitch;
  bf16* Qs = Vs + Lk * pitch;
  float* pbuf = reinterpret_cast<float*>(Qs + ((Lq * pitch + 1) & ~1));
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t kvrow0 = static_cast<int64_t>(b % a.kv_mod) * Lk;
  const int64_t qrow0 = static_cast<int64_t>(b % a.q_mod) * Lq;
  load_rows(Ks, a.k + kvrow0 * a.ldkv + h * hd, a.ldkv, Lk, hd);
  load_rows(Vs, a.v + kvrow0 * a.ldkv + h * hd, a.ldkv, Lk, hd);
  load_rows(Qs, a.q + qrow0 * a.ldq + h * hd, a.ldq, Lq, hd);
  __syncthreads();
  const uint8_t* mk = a.mask ? a.mask + static_cast<int64_t>(b % a.mask_mod) * Lk : nullptr;
  const float scale = rsqrtf(static_cast<float>(hd));
  uint64_t seed = 0, step = 0;
  if (a.drop_p > 0.f) {
    seed = a.rng[0];
    step = a.rng[1];
  }
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  float* pw = pbuf + warp * Lk;
  for (int i = warp; i < Lq; i += ATT_WARPS) {
    float s[ATT_MAXC];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < ATT_MAXC; ++c) {
      const int j = c * 32 + lane;
      s[c] = -INFINITY;
      if (j < Lk && !(mk && mk[j])) s[c] = scale * dot_row(Qs + i * pitch, Ks + j * pitch, hd);
      m = fmaxf(m, s[c]);
    }
    m = wmax(m);
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < ATT_MAXC; ++c) {
      s[c] = (s[c] == -INFINITY) ? 0.f : __expf(s[c] - m);
      l += s[c];
    }
    l = wsum2(l);
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    const uint64_t pbase = ((static_cast<uint64_t>(b) * a.H + h) * Lq + i) * static_cast<uint64_t>(Lk);
#pragma unroll
    for (int c = 0; c < ATT_MAXC; ++c) {
      const int j = c * 32 + lane;
      if (j < Lk) {
        float p = s[c] * inv_l;
        if (a.drop_p > 0.f) p = dropout_keep(seed, step, a.site, pbase + j, a.drop_p) ? p * keep_scale : 0.f;
        pw[j] = p;
      }
    }
    __syncwarp();
    const int d = lane * 2;
    if (d < hd) {
      float o0 = 0.f, o1 = 0.f;
      for (int j = 0; j < Lk; ++j) {
        const float p = pw[j];
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Vs + j * pitch + d));
        o0 = fmaf(p, v.x, o0);
        o1 = fmaf(p, v.y, o1);
      }
      *reinterpret_cast<__nv_bfloat162*>(a.o + (static_cast<int64_t>(b) * Lq + i) * a.ldo + h * hd + d) =
          __floats2bfloat162_rn(o0, o1);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(ATT_WARPS * 32) attention_bwd_kernel(const AttnArgs a) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int hd = a.hd, pitch = hd + 2, Lk = a.Lk, Lq = a.Lq;
  const int Lmax = Lk > Lq ? Lk : Lq;
  bf16* Ks = reinterpret_cast<bf16*>(smem_att);
  bf16* Vs = Ks + Lk * pitch;
  bf16* Qs = Vs + Lk * pitch;
  bf16* dOs = Qs + Lq * pitch;
  float* lse = reinterpret_cast<float*>(dOs + ((Lq * pitch + 1) & ~1));
  float* delta = lse + Lq;
  float* wbuf = delta + Lq;  // [ATT_WARPS][2 * Lmax]
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t kvrow0 = static_cast<int64_t>(b % a.kv_mod) * Lk;
  const int64_t qrow0 = static_cast<int64_t>(b % a.q_mod) * Lq;
  load_rows(Ks, a.k + kvrow0 * a.ldkv + h * hd, a.ldkv, Lk, hd);
  load_rows(Vs, a.v + kvrow0 * a.ldkv + h * hd, a.ldkv, Lk, hd);
  load_rows(Qs, a.q + qrow0 * a.ldq + h * hd, a.ldq, Lq, hd);
  load_rows(dOs, a.dout + (static_cast<int64_t>(b) * Lq) * a.lddo + h * hd, a.lddo, Lq, hd);
  __syncthreads();
  const uint8_t* mk = a.mask ? a.mask + static_cast<int64_t>(b % a.mask_mod) * Lk : nullptr;
  const float scale = rsqrtf(static_cast<float>(hd));
  uint64_t seed = 0, step = 0;
  if (a.drop_p > 0.f) {
    seed = a.rng[0];
    step = a.rng[1];
  }
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  float* w0 = wbuf + warp * 2 * Lmax;
  float* w1 = w0 + Lmax;

  // ---- pass A: one warp per query row -> softmax stats, delta, dQ
  for (int i = warp; i < Lq; i += ATT_WARPS) {
    float s[ATT_MAXC];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < ATT_MAXC; ++c) {
      const int j = c * 32 + lane;
      s[c] = -INFINITY;
      if (j < Lk && !(mk && mk[j])) s[c] = scale * dot_row(Qs + i * pitch, Ks + j * pitch, hd);
      m = fmaxf(m, s[c]);
    }
    m = wmax(m);
    float l = 0.f;
#pragma unroll
    for (int c = 0; c < ATT_MAXC; ++c) {
      s[c] = (s[c] == -INFINITY) ? 0.f : __expf(s[c] - m);
      l += s[c];
    }
    l = wsum2(l);
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    const uint64_t pbase = ((static_cast<uint64_t>(b) * a.H + h) * Lq + i) * static_cast<uint64_t>(Lk);
    float dp[ATT_MAXC];
    float dl = 0.f;
#pragma unroll
    for (int c = 0; c < ATT_MAXC; ++c) {
      const int j = c * 32 + lane;
      dp[c] = 0.f;
      if (j < Lk) {
        s[c] *= inv_l;  // p_ij
        float g = dot_row(dOs + i * pitch, Vs + j * pitch, hd);
        if (a.drop_p > 0.f) g = dropout_keep(seed, step, a.site, pbase + j, a.drop_p) ? g * keep_scale : 0.f;
        dp[c] = g;
        dl = fmaf(s[c], g, dl);
      }
    }
    dl = wsum2(dl);
#pragma unroll
    for (int c = 0; c < ATT_MAXC; ++c) {
      const int j = c * 32 + lane;
      if (j < Lk) w0[j] = s[c] * (dp[c] - dl);  // dS_ij
    }
    if (lane == 0) {
      lse[i] = l > 0.f ? m + __logf(l) : INFINITY;
      delta[i] = dl;
    }
    __syncwarp();
    const int d = lane * 2;
    if (d < hd) {
      float g0 = 0.f, g1 = 0.f;
      for (int j = 0; j < Lk; ++j) {
        const float ds = w0[j];
        const float2 kk = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Ks + j * pitch + d));
        g0 = fmaf(ds, kk.x, g0);
        g1 = fmaf(ds, kk.y, g1);
      }
      *reinterpret_cast<__nv_bfloat162*>(a.dq + (static_cast<int64_t>(b) * Lq + i) * a.lddq + h * hd + d) =
          __floats2bfloat162_rn(g0 * scale, g1 * scale);
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- pass B: one warp per key row -> dK, dV
  for (int j = warp; j < Lk; j += ATT_WARPS) {
    const bool masked = mk && mk[j];
#pragma unroll
    for (int c = 0; c < ATT_MAXC; ++c) {
      const int i = c * 32 + lane;
      if (i < Lq) {
        float ds = 0.f, pd = 0.f;
        if (!masked) {
          const float sc = scale * dot_row(Qs + i * pitch, Ks + j * pitch, hd);
          const float p = __expf(sc - lse[i]);
          float g = dot_row(dOs + i * pitch, Vs + j * pitch, hd);
          pd = p;
          if (a.drop_p > 0.f) {
            const uint64_t pbase = ((static_cast<uint64_t>(b) * a.H + h) * Lq + i) * static_cast<uint64_t>(Lk);
            const bool keep = dropout_keep(seed, step, a.site, pbase + j, a.drop_p);
            g = keep ? g * keep_scale : 0.f;
            pd = keep ? p * keep_scale : 0.f;
          }
          ds = p * (g - delta[i]);
        }
        w0[i] = ds;
        w1[i] = pd;
      }
    }
    __syncwarp();
    const int d = lane * 2;
    if (d < hd) {
      float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
      for (int i = 0; i < Lq; ++i) {
        const float ds = w0[i], pd = w1[i];
        const float2 qq = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Qs + i * pitch + d));
        const float2 dd = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dOs + i * pitch + d));
        k0 = fmaf(ds, qq.x, k0);
        k1 = fmaf(ds, qq.y, k1);
        v0 = fmaf(pd, dd.x, v0);
        v1 = fmaf(pd, dd.y, v1);
      }
      const int64_t row = static_cast<int64_t>(b) * Lk + j;
      *reinterpret_cast<__nv_bfloat162*>(a.dk + row * a.lddkv + h * hd + d) = __floats2bfloat162_rn(k0 * scale, k1 * scale);
      *reinterpret_cast<__nv_bfloat162*>(a.dv + row * a.lddkv + h * hd + d) = __floats2bfloat162_rn(v0, v1);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ short sequences (Lk <= 16)
// Four threads per row, each owning hd/4 head dims; scores live in registers. A warp covers 8 rows, a CTA 32:
// at S = 9 (8 patch tokens + CLS) this keeps every lane busy, where one-CTA-per-(row, head) would idle.
constexpr int SM_MAXL = 16;

// DPT consecutive head dims of one row; runs of 8 use 16-byte accesses (row pitches and head offsets are
// multiples of 8 elements whenever DPT is, and every base pointer is 16-byte aligned).
template <int DPT>
__device__ __forceinline__ void ld_slice(const bf16* p, float* v) {
  if constexpr (DPT % 8 == 0) {
#pragma unroll
    for (int d = 0; d < DPT; d += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p + d));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(h[q]);
        v[d + 2 * q] = f.x;
        v[d + 2 * q + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int d = 0; d < DPT; d += 2) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + d));
      v[d] = f.x;
      v[d + 1] = f.y;
    }
  }
}
template <int DPT>
__device__ __forceinline__ void st_slice(bf16* p, const float* v, float scale) {
  if constexpr (DPT % 8 == 0) {
#pragma unroll
    for (int d = 0; d < DPT; d += 8) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(v[d + 2 * q] * scale, v[d + 2 * q + 1] * scale);
      *reinterpret_cast<uint4*>(p + d) = u;
    }
  } else {
#pragma unroll
    for (int d = 0; d < DPT; d += 2)
      *reinterpret_cast<__nv_bfloat162*>(p + d) = __floats2bfloat162_rn(v[d] * scale, v[d + 1] * scale);
  }
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// MODE 0: forward (writes o). MODE 1: backward pass A (writes dq, lse, delta).
template <int DPT, int MODE>
__global__ void __launch_bounds__(128) attn_small_q_kernel(const AttnArgs a, float* __restrict__ stat) {
  pdl_entry();
  const int hd = a.hd, Lk = a.Lk, Lq = a.Lq;
  const int64_t total = static_cast<int64_t>(a.nb) * a.H * Lq;
  int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 2;
  const int part = threadIdx.x & 3;
  const bool valid = r < total;
  if (!valid) r = total - 1;
  const int i = static_cast<int>(r % Lq);
  const int h = static_cast<int>((r / Lq) % a.H);
  const int b = static_cast<int>(r / (static_cast<int64_t>(Lq) * a.H));
  const int col = h * hd + part * DPT;
  const float scale = rsqrtf(static_cast<float>(hd));
  float q[DPT], g[DPT], acc[DPT];
  ld_slice<DPT>(a.q + (static_cast<int64_t>(b % a.q_mod) * Lq + i) * a.ldq + col, q);
  if (MODE == 1) ld_slice<DPT>(a.dout + (static_cast<int64_t>(b) * Lq + i) * a.lddo + col, g);
#pragma unroll
  for (int d = 0; d < DPT; ++d) acc[d] = 0.f;
  const bf16* kbase = a.k + static_cast<int64_t>(b % a.kv_mod) * Lk * a.ldkv + col;
  const bf16* vbase = a.v + static_cast<int64_t>(b % a.kv_mod) * Lk * a.ldkv + col;
  const uint8_t* mk = a.mask ? a.mask + static_cast<int64_t>(b % a.mask_mod) * Lk : nullptr;
  uint64_t seed = 0, step = 0;
  if (a.drop_p > 0.f) {
    seed = a.rng[0];
    step = a.rng[1];
  }
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  const uint64_t pbase = ((static_cast<uint64_t>(b) * a.H + h) * Lq + i) * static_cast<uint64_t>(Lk);
  float s[SM_MAXL], dp[SM_MAXL];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < SM_MAXL; ++j) {
    s[j] = -INFINITY;
    dp[j] = 0.f;
    if (j < Lk) {
      float kk[DPT];
      ld_slice<DPT>(kbase + static_cast<int64_t>(j) * a.ldkv, kk);
      float t = 0.f;
#pragma unroll
      for (int d = 0; d < DPT; ++d) t = fmaf(q[d], kk[d], t);
      t = quad_sum(t) * scale;
      if (!(mk && mk[j])) s[j] = t;
      m = fmaxf(m, s[j]);
    }
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < SM_MAXL; ++j) {
    s[j] = (s[j] == -INFINITY) ? 0.f : __expf(s[j] - m);
    l += s[j];
  }
  const float inv_l = l > 0.f ? 1.f / l : 0.f;
  float delta = 0.f;
#pragma unroll
  for (int j = 0; j < SM_MAXL; ++j) {
    if (j < Lk) {
      s[j] *= inv_l;  // p_ij
      const bool keep = a.drop_p > 0.f ? dropout_keep(seed, step, a.site, pbase + j, a.drop_p) : true;
      float vv[DPT];
      ld_slice<DPT>(vbase + static_cast<int64_t>(j) * a.ldkv, vv);
      if (MODE == 0) {
        const float pd = keep ? s[j] * keep_scale : 0.f;
#pragma unroll
        for (int d = 0; d < DPT; ++d) acc[d] = fmaf(pd, vv[d], acc[d]);
      } else {
        float t = 0.f;
#pragma unroll
        for (int d = 0; d < DPT; ++d) t = fmaf(g[d], vv[d], t);
        t = quad_sum(t);
        dp[j] = keep ? t * keep_scale : 0.f;
        delta = fmaf(s[j], dp[j], delta);
      }
    }
  }
  if (MODE == 0) {
    if (valid) st_slice<DPT>(a.o + (static_cast<int64_t>(b) * Lq + i) * a.ldo + col, acc, 1.f);
    return;
  }
#pragma unroll
  for (int j = 0; j < SM_MAXL; ++j) {
    if (j < Lk) {
      const float ds = s[j] * (dp[j] - delta);
      float kk[DPT];
      ld_slice<DPT>(kbase + static_cast<int64_t>(j) * a.ldkv, kk);
#pragma unroll
      for (int d = 0; d < DPT; ++d) acc[d] = fmaf(ds, kk[d], acc[d]);
    }
  }
  if (valid) {
    st_slice<DPT>(a.dq + (static_cast<int64_t>(b) * Lq + i) * a.lddq + col, acc, scale);
    if (part == 0) {
      stat[2 * r] = l > 0.f ? m + __logf(l) : INFINITY;
      stat[2 * r + 1] = delta;
    }
  }
}

// backward pass B: one thread quad per key row -> dK, dV (sums over the queries; deterministic)
template <int DPT>
__global__ void __launch_bounds__(128) attn_small_kv_kernel(const AttnArgs a, const float* __restrict__ stat) {
  pdl_entry();
  const int hd = a.hd, Lk = a.Lk, Lq = a.Lq;
  const int64_t total = static_cast<int64_t>(a.nb) * a.H * Lk;
  int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 2;
  const int part = threadIdx.x & 3;
  const bool valid = r < total;
  if (!valid) r = total - 1;
  const int j = static_cast<int>(r % Lk);
  const int h = static_cast<int>((r / Lk) % a.H);
  const int b = static_cast<int>(r / (static_cast<int64_t>(Lk) * a.H));
  const int col = h * hd + part * DPT;
  const float scale = rsqrtf(static_cast<float>(hd));
  float kk[DPT], vv[DPT], dk[DPT], dv[DPT];
  const int64_t kvrow = static_cast<int64_t>(b % a.kv_mod) * Lk + j;
  ld_slice<DPT>(a.k + kvrow * a.ldkv + col, kk);
  ld_slice<DPT>(a.v + kvrow * a.ldkv + col, vv);
#pragma unroll
  for (int d = 0; d < DPT; ++d) dk[d] = dv[d] = 0.f;
  const bool masked = a.mask && a.mask[static_cast<int64_t>(b % a.mask_mod) * Lk + j];
  uint64_t seed = 0, step = 0;
  if (a.drop_p > 0.f) {
    seed = a.rng[0];
    step = a.rng[1];
  }
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  const bf16* qbase = a.q + static_cast<int64_t>(b % a.q_mod) * Lq * a.ldq + col;
  const bf16* gbase = a.dout + static_cast<int64_t>(b) * Lq * a.lddo + col;
  const int64_t srow = (static_cast<int64_t>(b) * a.H + h) * Lq;
  for (int i = 0; i < Lq; ++i) {
    float q[DPT], g[DPT];
    ld_slice<DPT>(qbase + static_cast<int64_t>(i) * a.ldq, q);
    ld_slice<DPT>(gbase + static_cast<int64_t>(i) * a.lddo, g);
    float t = 0.f, u = 0.f;
#pragma unroll
    for (int d = 0; d < DPT; ++d) {
      t = fmaf(q[d], kk[d], t);
      u = fmaf(g[d], vv[d], u);
    }
    t = quad_sum(t) * scale;
    u = quad_sum(u);
    const float p = masked ? 0.f : __expf(t - stat[2 * (srow + i)]);
    const bool keep = a.drop_p > 0.f
                          ? dropout_keep(seed, step, a.site, static_cast<uint64_t>(srow + i) * Lk + j, a.drop_p)
                          : true;
    const float pd = keep ? p * keep_scale : 0.f;
    const float ds = p * ((keep ? u * keep_scale : 0.f) - stat[2 * (srow + i) + 1]);
#pragma unroll
    for (int d = 0; d < DPT; ++d) {
      dk[d] = fmaf(ds, q[d], dk[d]);
      dv[d] = fmaf(pd, g[d], dv[d]);
    }
  }
  if (valid) {
    const int64_t orow = static_cast<int64_t>(b) * Lk + j;
    st_slice<DPT>(a.dk + orow * a.lddkv + col, dk, scale);
    st_slice<DPT>(a.dv + orow * a.lddkv + col, dv, 1.f);
  }
}

// ------------------------------------------------------- short self-attention, head_dim 64 (the tower)
// S = Lq = Lk <= 16 tokens, hd = 64: the encoder layers of the paper / FiLM models at P = 8 patches.
// One WARP per (sequence, head): the 16x16 score tile and the 16x64 outputs are m16n8k16 tensor-core
// fragments (mma.sync, bf16 in / fp32 accumulate — a 9x9 problem is far below the 128-row tcgen05 tile; the
// op is memory bound and this keeps it at a few hundred instructions per sequence). Q / K / V (/ dO) rows of
// the CTA's four groups are staged with 16-byte cp.async (rows >= S zero-filled), fragments come from
// ldmatrix. The backward is one kernel: S and P are recomputed, dP = dO V^T, dS = P o (dP - delta),
// dQ = dS K, and dK = dS^T Q, dV = P^T dO go through a bf16 copy of dS / P in shared memory (ldmatrix.trans);
// no atomics, deterministic.
constexpr int SELF_HD = 64;
constexpr int SELF_PITCH = SELF_HD + 8;   // bf16 elements: 144-byte rows, conflict-free ldmatrix
constexpr int SELF_GROUPS = 4;            // warps = (sequence, head) groups per CTA
constexpr int SELF_THREADS = SELF_GROUPS * 32;
constexpr int SELF_TILE = 16 * SELF_PITCH;  // one 16-row tile (elements)
constexpr int SELF_SP = 24;               // pitch of the 16x16 bf16 dS / P tiles

// The five PTX wrappers below are the only inline assembly the kernels of this file depend on for their results;
// tests/cuda_emu/emu_attention.cpp defines GG_EMULATED_PTX and supplies host versions of them (cp.async as a copy,
// ldmatrix / mma.sync as warp-collective exchanges), so that the kernels themselves run unchanged on the CPU suite.
#ifndef GG_EMULATED_PTX
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 2^x on the SFU (ex2.approx.ftz: one instruction; 2^-inf = 0). The softmaxes below work in the log2 domain:
// score * (log2(e) / sqrt(hd)) + key bias (0, or -inf for padded / out-of-range keys), so that a probability is
// ex2(v - max) with no select around it (__expf cost ~12 issued instructions per element with its range handling).
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
#endif  // GG_EMULATED_PTX
constexpr float SCALE_LOG2E = 0.125f * 1.4426950408889634f;  // log2(e) / sqrt(64)

__device__ __forceinline__ uint32_t pack_bf16(float x, float y) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_add(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---- precomputed dropout keep bits (AttnArgs::dbits). The mid / long self-attention kernels spend most of their issue
// slots on Philox when they draw the masks themselves (a 16 x 16 score tile of one warp meets ~40 eight-element groups
// of the stream, and the backward draws them twice more: cfg2 forward 354 -> 1180 us with dropout on). One pass of
// dropout_bits_kernel draws every group exactly once — bit idx of the mask = element idx of the stream, 1 = kept —
// and forward and backward read 16-bit windows of it (L2-resident: S*S bits per head).
__device__ __forceinline__ uint32_t keep_window16(const uint32_t* __restrict__ bits, uint64_t o) {
  const uint64_t w = o >> 5;
  const uint64_t two = static_cast<uint64_t>(__ldg(bits + w)) | (static_cast<uint64_t>(__ldg(bits + w + 1)) << 32);
  return static_cast<uint32_t>(two >> (static_cast<uint32_t>(o) & 31u)) & 0xFFFFu;
}

__global__ void __launch_bounds__(256)
    dropout_bits_kernel(const uint64_t* __restrict__ rng, uint32_t site, float p, int64_t n_words,
                        uint32_t* __restrict__ out) {
  pdl_entry();
  const uint64_t seed = rng[0], step = rng[1];
  const uint32_t thr = dropout_thr(p);
  for (int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; w < n_words;
       w += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      v |= keep_bits8(dropout_words(seed, step, site, static_cast<uint64_t>(w) * 4 + k), thr) << (8 * k);
    out[w] = v;
  }
}

// C[16 x 16] (2 n-tiles) = A[16 x 64] * B^T with B stored [16 rows (n)][64 (k)]: both row-major tiles in smem
__device__ __forceinline__ void mma_ab_t(float (&c)[2][4], const bf16* A, const bf16* B, int lane) {
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[nt][j] = 0.f;
  const int arow = (lane & 7) + ((lane >> 3) & 1) * 8, acol = (lane >> 4) * 8;
  const int brow = (lane & 7) + (lane >> 4) * 8, bcol = ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < SELF_HD / 16; ++kk) {
    uint32_t a[4], b[4];
    ldsm_x4(a, A + arow * SELF_PITCH + kk * 16 + acol);
    ldsm_x4(b, B + brow * SELF_PITCH + kk * 16 + bcol);
    mma16816(c[0], a, b[0], b[1]);
    mma16816(c[1], a, b[2], b[3]);
  }
}
// O[16 x 64] (8 n-tiles) = A[16 x 16] (fragments) * B with B stored [16 rows (k)][64 (n)] row-major in smem
__device__ __forceinline__ void mma_frag_b(float (&o)[8][4], const uint32_t (&a)[4], const bf16* B, int lane) {
  const int brow = (lane & 7) + ((lane >> 3) & 1) * 8, bcol = (lane >> 4) * 8;
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    uint32_t b[4];
    ldsm_x4_t(b, B + brow * SELF_PITCH + np * 16 + bcol);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[2 * np][j] = o[2 * np + 1][j] = 0.f;
    mma16816(o[2 * np], a, b[0], b[1]);
    mma16816(o[2 * np + 1], a, b[2], b[3]);
  }
}
// fragments (rows g / g+8, cols nt*8 + 2t) -> bf16 rows in a smem tile, then the first S rows -> global
__device__ __forceinline__ void store_tile(bf16* stage, const float (&o)[8][4], float scale, bf16* gdst, int64_t ld,
                                           int S, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(stage + g * SELF_PITCH + nt * 8 + 2 * t) = pack_bf16(o[nt][0] * scale, o[nt][1] * scale);
    *reinterpret_cast<uint32_t*>(stage + (g + 8) * SELF_PITCH + nt * 8 + 2 * t) = pack_bf16(o[nt][2] * scale, o[nt][3] * scale);
  }
  __syncwarp();
  for (int c = lane; c < S * 8; c += 32) {
    const int r = c >> 3, part = c & 7;
    *reinterpret_cast<uint4*>(gdst + static_cast<int64_t>(r) * ld + part * 8) =
        *reinterpret_cast<const uint4*>(stage + r * SELF_PITCH + part * 8);
  }
  __syncwarp();
}

// MODE 0: forward. MODE 1: backward (dq, dk, dv).
template <int MODE>
__global__ void __launch_bounds__(SELF_THREADS) attn_self_small_kernel(const AttnArgs a) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t smem_self[];
  const int S = a.Lq;
  constexpr int NT = MODE == 1 ? 4 : 3;               // staged tensors: Q, K, V (, dO)
  bf16* tiles = reinterpret_cast<bf16*>(smem_self);   // [group][NT + 1 (staging)][16][SELF_PITCH]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ngroups = static_cast<int64_t>(a.nb) * a.H;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * SELF_GROUPS;
  constexpr int PER_GROUP = (NT + 1) * SELF_TILE + (MODE == 1 ? 2 * 16 * SELF_SP : 0);
  // ---- stage: thread = (row, 16-byte part) of every (group, tensor) tile; rows >= S and groups beyond the
  // end are zero-filled. (b, h) advance incrementally: no per-chunk integer division.
  {
    const int row = threadIdx.x >> 3, part = threadIdx.x & 7;
    const uint32_t H = static_cast<uint32_t>(a.H);
    uint32_t bb = static_cast<uint32_t>(g0 / H), hh = static_cast<uint32_t>(g0 % H);
    // row pointers of sequence bb (recomputed only when the group walk crosses into the next sequence: with 4 heads
    // and 4 groups per CTA that is never, and the four heads' slices are 128 bytes apart in the same rows)
    const bf16 *qrow = nullptr, *krow = nullptr, *vrow = nullptr, *grow = nullptr;
    auto row_ptrs = [&]() {
      const int64_t qr = static_cast<int64_t>(a.q_mod >= a.nb ? bb : bb % static_cast<uint32_t>(a.q_mod)) * S + row;
      const int64_t kr = static_cast<int64_t>(a.kv_mod >= a.nb ? bb : bb % static_cast<uint32_t>(a.kv_mod)) * S + row;
      qrow = a.q + qr * a.ldq + part * 8;
      krow = a.k + kr * a.ldkv + part * 8;
      vrow = a.v + kr * a.ldkv + part * 8;
      if (MODE == 1) grow = a.dout + (static_cast<int64_t>(bb) * S + row) * a.lddo + part * 8;
    };
    row_ptrs();
    const bool row_ok = row < S;
    bf16* dst = tiles + row * SELF_PITCH + part * 8;
#pragma unroll
    for (int gi = 0; gi < SELF_GROUPS; ++gi) {
      const bool valid = row_ok && g0 + gi < ngroups;
      const int col = static_cast<int>(hh) * SELF_HD;
      cp_async16(dst, valid ? qrow + col : a.q, valid);
      cp_async16(dst + SELF_TILE, valid ? krow + col : a.q, valid);
      cp_async16(dst + 2 * SELF_TILE, valid ? vrow + col : a.q, valid);
      if (MODE == 1) cp_async16(dst + 3 * SELF_TILE, valid ? grow + col : a.q, valid);
      dst += PER_GROUP;
      if (++hh == H) {
        hh = 0;
        ++bb;
        if (gi + 1 < SELF_GROUPS) row_ptrs();
      }
    }
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  const int64_t gid = g0 + warp;
  if (gid >= ngroups) return;
  const int b = static_cast<int>(gid / a.H), h = static_cast<int>(gid % a.H);
  bf16* Qs = tiles + warp * PER_GROUP;
  bf16* Ks = Qs + SELF_TILE;
  bf16* Vs = Ks + SELF_TILE;
  bf16* Gs = Vs + SELF_TILE;                       // dO (backward only)
  bf16* stage = Qs + NT * SELF_TILE;
  const uint8_t* mk = a.mask ? a.mask + static_cast<int64_t>(b % a.mask_mod) * S : nullptr;
  const float scale = 0.125f;  // 1/sqrt(64)
  const int g = lane >> 2, t = lane & 3;
  // ---- scores and probabilities: thread holds rows g, g+8 x keys {2t, 2t+1, 8+2t, 9+2t}
  float sc[2][4];
  mma_ab_t(sc, Qs, Ks, lane);
  float kbias[4];  // 0, or -inf for a padded / out-of-range key
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int j = (e >> 1) * 8 + 2 * t + (e & 1);
    kbias[e] = (j < S && !(mk && mk[j])) ? 0.f : -INFINITY;
  }
  float p[2][4];  // [row half][key e]
#pragma unroll
  for (int rh = 0; rh < 2; ++rh) {
    float m = -INFINITY;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      p[rh][e] = fmaf(sc[e >> 1][rh * 2 + (e & 1)], SCALE_LOG2E, kbias[e]);
      m = fmaxf(m, p[rh][e]);
    }
    m = quad_max(m);
    m = m == -INFINITY ? 0.f : m;  // every key masked: all probabilities 0
    float l = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      p[rh][e] = fast_ex2(p[rh][e] - m);
      l += p[rh][e];
    }
    l = quad_add(l);
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) p[rh][e] *= inv_l;
  }
  // dropout multipliers (0 or 1/(1-p)) on the probabilities, same element index as every other path
  float mult[2][4];
#pragma unroll
  for (int rh = 0; rh < 2; ++rh)
#pragma unroll
    for (int e = 0; e < 4; ++e) mult[rh][e] = 1.f;
  if (a.drop_p > 0.f) {
    const uint64_t seed = a.rng[0], step = a.rng[1];
    const float keep_scale = 1.f / (1.f - a.drop_p);
    DropoutStream ds(seed, step, a.site, a.drop_p);
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int i = g + rh * 8;
      const uint64_t pbase = ((static_cast<uint64_t>(b) * a.H + h) * S + i) * static_cast<uint64_t>(S);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = (e >> 1) * 8 + 2 * t + (e & 1);
        if (i < S && j < S) mult[rh][e] = ds.keep(pbase + j) ? keep_scale : 0.f;
      }
    }
  }
  // A fragment of the (dropped) probabilities: a0 (row g, keys 2t..), a1 (row g+8), a2 (row g, keys 8+2t..), a3
  uint32_t pa[4];
  pa[0] = pack_bf16(p[0][0] * mult[0][0], p[0][1] * mult[0][1]);
  pa[1] = pack_bf16(p[1][0] * mult[1][0], p[1][1] * mult[1][1]);
  pa[2] = pack_bf16(p[0][2] * mult[0][2], p[0][3] * mult[0][3]);
  pa[3] = pack_bf16(p[1][2] * mult[1][2], p[1][3] * mult[1][3]);
  float o[8][4];
  if (MODE == 0) {
    mma_frag_b(o, pa, Vs, lane);
    store_tile(stage, o, 1.f, a.o + static_cast<int64_t>(b) * S * a.ldo + h * SELF_HD, a.ldo, S, lane);
    return;
  }
  // ---- backward
  float dp[2][4];
  mma_ab_t(dp, Gs, Vs, lane);  // dP = dO V^T
  float ds[2][4];
#pragma unroll
  for (int rh = 0; rh < 2; ++rh) {
    float delta = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d = dp[e >> 1][rh * 2 + (e & 1)] * mult[rh][e];
      ds[rh][e] = d;
      delta = fmaf(p[rh][e], d, delta);
    }
    delta = quad_add(delta);
#pragma unroll
    for (int e = 0; e < 4; ++e) ds[rh][e] = p[rh][e] * (ds[rh][e] - delta);
  }
  uint32_t da[4];
  da[0] = pack_bf16(ds[0][0], ds[0][1]);
  da[1] = pack_bf16(ds[1][0], ds[1][1]);
  da[2] = pack_bf16(ds[0][2], ds[0][3]);
  da[3] = pack_bf16(ds[1][2], ds[1][3]);
  // dQ = dS K * scale
  mma_frag_b(o, da, Ks, lane);
  store_tile(stage, o, scale, a.dq + static_cast<int64_t>(b) * S * a.lddq + h * SELF_HD, a.lddq, S, lane);
  // bf16 copies of dS and P (dropped) for the transposed products
  bf16* dSs = Qs + (NT + 1) * SELF_TILE;
  bf16* Ps = dSs + 16 * SELF_SP;
  *reinterpret_cast<uint32_t*>(dSs + g * SELF_SP + 2 * t) = da[0];
  *reinterpret_cast<uint32_t*>(dSs + (g + 8) * SELF_SP + 2 * t) = da[1];
  *reinterpret_cast<uint32_t*>(dSs + g * SELF_SP + 8 + 2 * t) = da[2];
  *reinterpret_cast<uint32_t*>(dSs + (g + 8) * SELF_SP + 8 + 2 * t) = da[3];
  *reinterpret_cast<uint32_t*>(Ps + g * SELF_SP + 2 * t) = pa[0];
  *reinterpret_cast<uint32_t*>(Ps + (g + 8) * SELF_SP + 2 * t) = pa[1];
  *reinterpret_cast<uint32_t*>(Ps + g * SELF_SP + 8 + 2 * t) = pa[2];
  *reinterpret_cast<uint32_t*>(Ps + (g + 8) * SELF_SP + 8 + 2 * t) = pa[3];
  __syncwarp();
  // A fragments of X^T from the [query][key] tile: a0 = X[q 0-7][k 0-7]^T, a1 = X[q 0-7][k 8-15]^T,
  // a2 = X[q 8-15][k 0-7]^T, a3 = X[q 8-15][k 8-15]^T
  const int trow = (lane & 7) + (lane >> 4) * 8, tcol = ((lane >> 3) & 1) * 8;
  uint32_t ta[4];
  ldsm_x4_t(ta, dSs + trow * SELF_SP + tcol);
  mma_frag_b(o, ta, Qs, lane);  // dK = dS^T Q * scale
  store_tile(stage, o, scale, a.dk + static_cast<int64_t>(b) * S * a.lddkv + h * SELF_HD, a.lddkv, S, lane);
  ldsm_x4_t(ta, Ps + trow * SELF_SP + tcol);
  mma_frag_b(o, ta, Gs, lane);  // dV = P^T dO
  store_tile(stage, o, 1.f, a.dv + static_cast<int64_t>(b) * S * a.lddkv + h * SELF_HD, a.lddkv, S, lane);
}

// ------------------------------------------------- mid-size self-attention, head_dim 64 (17 <= S <= 128)
// The encoder layers at P = 64 patches (BASELINE config 4: S = 65). One CTA per (sequence, head), NT = ceil(S/16)
// tiles of 16 tokens; Q / K / V (/ dO) of the head are staged once with 16-byte cp.async (rows >= S zero-filled).
// Forward: a warp owns 16 query rows: the whole 16 x S score row-block lives in m16n8k16 accumulator fragments
// (softmax in registers, quad shuffles), P V accumulates over the key tiles. Backward: phase 1 (warp = query
// tile) recomputes P, dP = dO V^T, dS = P o (dP - delta), dQ = dS K and leaves bf16 dS / P in shared memory;
// phase 2 (warp = key tile) forms dK = dS^T Q and dV = P^T dO from them through ldmatrix.trans. No atomics.
// (The CUDA-core kernels above took 2.6 ms forward / 10 ms backward per call at 12288 sequences x 65 tokens.)
// A fragments (16 x 64) of a tile held in registers across a walk over B tiles, and C = A * B^T from them
__device__ __forceinline__ void load_a_frags16(uint32_t (&fa)[4][4], const bf16* A, int lane) {
  const int arow = (lane & 7) + ((lane >> 3) & 1) * 8, acol = (lane >> 4) * 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldsm_x4(fa[kk], A + arow * SELF_PITCH + kk * 16 + acol);
}
__device__ __forceinline__ void mma_afrag_bt(float (&c)[2][4], const uint32_t (&fa)[4][4], const bf16* B, int lane) {
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[nt][j] = 0.f;
  const int brow = (lane & 7) + (lane >> 4) * 8, bcol = ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t b[4];
    ldsm_x4(b, B + brow * SELF_PITCH + kk * 16 + bcol);
    mma16816(c[0], fa[kk], b[0], b[1]);
    mma16816(c[1], fa[kk], b[2], b[3]);
  }
}
__device__ __forceinline__ void mma_frag_b_acc(float (&o)[8][4], const uint32_t (&a)[4], const bf16* B, int lane) {
  const int brow = (lane & 7) + ((lane >> 3) & 1) * 8, bcol = (lane >> 4) * 8;
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    uint32_t b[4];
    ldsm_x4_t(b, B + brow * SELF_PITCH + np * 16 + bcol);
    mma16816(o[2 * np], a, b[0], b[1]);
    mma16816(o[2 * np + 1], a, b[2], b[3]);
  }
}


template <int MODE, int NT>
__global__ void __launch_bounds__(NT * 32) attn_self_mid_kernel(const AttnArgs a) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t smem_mid[];
  constexpr int MID_WARPS = NT;  // one warp per 16-token tile (query tile in phase 1, key tile in phase 2)
  constexpr int ROWS = NT * 16;
  constexpr int SP = ROWS + 8;                       // pitch of the bf16 dS / P matrices
  constexpr int NTEN = MODE == 1 ? 4 : 3;            // staged tensors: Q, K, V (, dO)
  const int S = a.Lq;
  bf16* Qs = reinterpret_cast<bf16*>(smem_mid);
  bf16* Ks = Qs + ROWS * SELF_PITCH;
  bf16* Vs = Ks + ROWS * SELF_PITCH;
  bf16* Gs = Vs + ROWS * SELF_PITCH;                 // dO (backward only)
  bf16* stage_all = Qs + NTEN * ROWS * SELF_PITCH;   // one 16-row output staging tile per warp
  bf16* dSs = stage_all + MID_WARPS * SELF_TILE;     // backward only: [ROWS][SP]
  bf16* Ps = dSs + ROWS * SP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = static_cast<int>(blockIdx.x) / a.H, h = static_cast<int>(blockIdx.x) % a.H;
  {
    const int col = h * SELF_HD;
    const int64_t qb = static_cast<int64_t>(a.q_mod >= a.nb ? b : b % a.q_mod) * S;
    const int64_t kb = static_cast<int64_t>(a.kv_mod >= a.nb ? b : b % a.kv_mod) * S;
    for (int idx = threadIdx.x; idx < ROWS * 8; idx += MID_WARPS * 32) {
      const int row = idx >> 3, part = idx & 7;
      const bool valid = row < S;
      bf16* dst = Qs + row * SELF_PITCH + part * 8;
      cp_async16(dst, valid ? a.q + (qb + row) * a.ldq + col + part * 8 : a.q, valid);
      cp_async16(dst + ROWS * SELF_PITCH, valid ? a.k + (kb + row) * a.ldkv + col + part * 8 : a.q, valid);
      cp_async16(dst + 2 * ROWS * SELF_PITCH, valid ? a.v + (kb + row) * a.ldkv + col + part * 8 : a.q, valid);
      if (MODE == 1)
        cp_async16(dst + 3 * ROWS * SELF_PITCH,
                   valid ? a.dout + (static_cast<int64_t>(b) * S + row) * a.lddo + col + part * 8 : a.q, valid);
    }
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  const uint8_t* mk = a.mask ? a.mask + static_cast<int64_t>(b % a.mask_mod) * S : nullptr;
  const float scale = 0.125f;  // 1/sqrt(64)
  const int g = lane >> 2, t = lane & 3;
  bf16* stage = stage_all + warp * SELF_TILE;
  const bool drop = a.drop_p > 0.f;
  uint64_t seed = 0, step = 0;
  if (drop) {
    seed = a.rng[0];
    step = a.rng[1];
  }
  const float keep_scale = drop ? 1.f / (1.f - a.drop_p) : 1.f;
  // key validity of this thread's columns: key j = kt*16 + (e>>1)*8 + 2t + (e&1)
  float kbias[NT][4];  // 0, or -inf for a padded / out-of-range key
#pragma unroll
  for (int kt = 0; kt < NT; ++kt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = kt * 16 + (e >> 1) * 8 + 2 * t + (e & 1);
      kbias[kt][e] = (j < S && !(mk && mk[j])) ? 0.f : -INFINITY;
    }
  for (int qt = warp; qt < NT; qt += MID_WARPS) {
    const bf16* Qt = Qs + qt * SELF_TILE;
    float p[NT][2][4];  // [key tile][row half][e]
    {
      float sc[NT][2][4];
      uint32_t qa[4][4];  // the query tile's A fragments: loaded once, not once per key tile
      load_a_frags16(qa, Qt, lane);
#pragma unroll
      for (int kt = 0; kt < NT; ++kt) mma_afrag_bt(sc[kt], qa, Ks + kt * SELF_TILE, lane);
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        float m = -INFINITY;
#pragma unroll
        for (int kt = 0; kt < NT; ++kt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            p[kt][rh][e] = fmaf(sc[kt][e >> 1][rh * 2 + (e & 1)], SCALE_LOG2E, kbias[kt][e]);
            m = fmaxf(m, p[kt][rh][e]);
          }
        m = quad_max(m);
        m = m == -INFINITY ? 0.f : m;
        float l = 0.f;
#pragma unroll
        for (int kt = 0; kt < NT; ++kt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            p[kt][rh][e] = fast_ex2(p[kt][rh][e] - m);
            l += p[kt][rh][e];
          }
        l = quad_add(l);
        const float inv_l = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
        for (int kt = 0; kt < NT; ++kt)
#pragma unroll
          for (int e = 0; e < 4; ++e) p[kt][rh][e] *= inv_l;
      }
    }
    // dropout multipliers folded into a bit mask (bit kt*8 + rh*4 + e = dropped)
    uint64_t dropped = 0;
    if (drop && a.dbits) {
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int i = qt * 16 + g + rh * 8;
        if (i >= S) continue;
        const uint64_t pbase = ((static_cast<uint64_t>(b) * a.H + h) * S + i) * static_cast<uint64_t>(S);
#pragma unroll
        for (int kt = 0; kt < NT; ++kt) {
          const uint32_t win = keep_window16(a.dbits, pbase + kt * 16);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (!((win >> ((e >> 1) * 8 + 2 * t + (e & 1))) & 1u)) dropped |= 1ull << (kt * 8 + rh * 4 + e);
        }
      }
    } else if (drop) {
      DropoutStream ds(seed, step, a.site, a.drop_p);
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int i = qt * 16 + g + rh * 8;
        const uint64_t pbase = ((static_cast<uint64_t>(b) * a.H + h) * S + i) * static_cast<uint64_t>(S);
#pragma unroll
        for (int kt = 0; kt < NT; ++kt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = kt * 16 + (e >> 1) * 8 + 2 * t + (e & 1);
            if (i < S && j < S && !ds.keep(pbase + j)) dropped |= 1ull << (kt * 8 + rh * 4 + e);
          }
      }
    }
    auto mult = [&](int kt, int rh, int e) { return ((dropped >> (kt * 8 + rh * 4 + e)) & 1ull) ? 0.f : keep_scale; };
    uint32_t pa[NT][4];  // A fragments of the (dropped) probabilities
#pragma unroll
    for (int kt = 0; kt < NT; ++kt) {
      pa[kt][0] = pack_bf16(p[kt][0][0] * mult(kt, 0, 0), p[kt][0][1] * mult(kt, 0, 1));
      pa[kt][1] = pack_bf16(p[kt][1][0] * mult(kt, 1, 0), p[kt][1][1] * mult(kt, 1, 1));
      pa[kt][2] = pack_bf16(p[kt][0][2] * mult(kt, 0, 2), p[kt][0][3] * mult(kt, 0, 3));
      pa[kt][3] = pack_bf16(p[kt][1][2] * mult(kt, 1, 2), p[kt][1][3] * mult(kt, 1, 3));
    }
    const int rows_valid = min(16, S - qt * 16);
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[n][j] = 0.f;
    if (MODE == 0) {
#pragma unroll
      for (int kt = 0; kt < NT; ++kt) mma_frag_b_acc(o, pa[kt], Vs + kt * SELF_TILE, lane);
      store_tile(stage, o, 1.f, a.o + (static_cast<int64_t>(b) * S + qt * 16) * a.ldo + h * SELF_HD, a.ldo, rows_valid,
                 lane);
      continue;
    }
    // ---- backward, phase 1
    const bf16* Gt = Gs + qt * SELF_TILE;
    float delta[2] = {0.f, 0.f};
    float dsv[NT][2][4];
    uint32_t ga[4][4];
    load_a_frags16(ga, Gt, lane);
#pragma unroll
    for (int kt = 0; kt < NT; ++kt) {
      float dp[2][4];
      mma_afrag_bt(dp, ga, Vs + kt * SELF_TILE, lane);  // dP~ = dO V^T
#pragma unroll
      for (int rh = 0; rh < 2; ++rh)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d = dp[e >> 1][rh * 2 + (e & 1)] * mult(kt, rh, e);
          dsv[kt][rh][e] = d;
          delta[rh] = fmaf(p[kt][rh][e], d, delta[rh]);
        }
    }
    delta[0] = quad_add(delta[0]);
    delta[1] = quad_add(delta[1]);
#pragma unroll
    for (int kt = 0; kt < NT; ++kt) {
      uint32_t da[4];
      da[0] = pack_bf16(p[kt][0][0] * (dsv[kt][0][0] - delta[0]), p[kt][0][1] * (dsv[kt][0][1] - delta[0]));
      da[1] = pack_bf16(p[kt][1][0] * (dsv[kt][1][0] - delta[1]), p[kt][1][1] * (dsv[kt][1][1] - delta[1]));
      da[2] = pack_bf16(p[kt][0][2] * (dsv[kt][0][2] - delta[0]), p[kt][0][3] * (dsv[kt][0][3] - delta[0]));
      da[3] = pack_bf16(p[kt][1][2] * (dsv[kt][1][2] - delta[1]), p[kt][1][3] * (dsv[kt][1][3] - delta[1]));
      mma_frag_b_acc(o, da, Ks + kt * SELF_TILE, lane);  // dQ += dS K
      bf16* dst = dSs + (qt * 16) * SP + kt * 16;
      bf16* pst = Ps + (qt * 16) * SP + kt * 16;
      *reinterpret_cast<uint32_t*>(dst + g * SP + 2 * t) = da[0];
      *reinterpret_cast<uint32_t*>(dst + (g + 8) * SP + 2 * t) = da[1];
      *reinterpret_cast<uint32_t*>(dst + g * SP + 8 + 2 * t) = da[2];
      *reinterpret_cast<uint32_t*>(dst + (g + 8) * SP + 8 + 2 * t) = da[3];
      *reinterpret_cast<uint32_t*>(pst + g * SP + 2 * t) = pa[kt][0];
      *reinterpret_cast<uint32_t*>(pst + (g + 8) * SP + 2 * t) = pa[kt][1];
      *reinterpret_cast<uint32_t*>(pst + g * SP + 8 + 2 * t) = pa[kt][2];
      *reinterpret_cast<uint32_t*>(pst + (g + 8) * SP + 8 + 2 * t) = pa[kt][3];
    }
    store_tile(stage, o, scale, a.dq + (static_cast<int64_t>(b) * S + qt * 16) * a.lddq + h * SELF_HD, a.lddq,
               rows_valid, lane);
  }
  if (MODE == 0) return;
  __syncthreads();
  // ---- backward, phase 2: A fragments of X^T from the [query][key] matrices
  const int trow = (lane & 7) + (lane >> 4) * 8, tcol = ((lane >> 3) & 1) * 8;
  for (int kt = warp; kt < NT; kt += MID_WARPS) {
    const int rows_valid = min(16, S - kt * 16);
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[n][j] = 0.f;
#pragma unroll
    for (int qt = 0; qt < NT; ++qt) {
      uint32_t ta[4];
      ldsm_x4_t(ta, dSs + (qt * 16 + trow) * SP + kt * 16 + tcol);
      mma_frag_b_acc(o, ta, Qs + qt * SELF_TILE, lane);  // dK += dS^T Q
    }
    store_tile(stage, o, scale, a.dk + (static_cast<int64_t>(b) * S + kt * 16) * a.lddkv + h * SELF_HD, a.lddkv,
               rows_valid, lane);
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[n][j] = 0.f;
#pragma unroll
    for (int qt = 0; qt < NT; ++qt) {
      uint32_t ta[4];
      ldsm_x4_t(ta, Ps + (qt * 16 + trow) * SP + kt * 16 + tcol);
      mma_frag_b_acc(o, ta, Gs + qt * SELF_TILE, lane);  // dV += P^T dO
    }
    store_tile(stage, o, 1.f, a.dv + (static_cast<int64_t>(b) * S + kt * 16) * a.lddkv + h * SELF_HD, a.lddkv,
               rows_valid, lane);
  }
}

template <int MODE, int NT>
static int launch_mid_nt(const AttnArgs& a, cudaStream_t st) {
  constexpr int ROWS = NT * 16;
  constexpr int MID_WARPS = NT;
  const size_t smem = (static_cast<size_t>((MODE == 1 ? 4 : 3) * ROWS * SELF_PITCH + MID_WARPS * SELF_TILE +
                                           (MODE == 1 ? 2 * ROWS * (ROWS + 8) : 0))) * 2 + 16;
  static bool configured = false;
  if (!configured) {
    GG_CUDA_CHECK(cudaFuncSetAttribute(attn_self_mid_kernel<MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    configured = true;
  }
  launch_k(attn_self_mid_kernel<MODE, NT>, static_cast<unsigned>(a.nb) * a.H, MID_WARPS * 32, smem, st, a);
  GG_LAUNCH_CHECK();
  return GG_OK;
}
template <int MODE>
static int launch_mid(const AttnArgs& a, cudaStream_t st) {
  const int nt = (a.Lq + 15) / 16;
  if (nt <= 2) return launch_mid_nt<MODE, 2>(a, st);
  if (nt <= 3) return launch_mid_nt<MODE, 3>(a, st);
  if (nt <= 4) return launch_mid_nt<MODE, 4>(a, st);
  if (nt <= 5) return launch_mid_nt<MODE, 5>(a, st);
  if (nt <= 6) return launch_mid_nt<MODE, 6>(a, st);
  return launch_mid_nt<MODE, 8>(a, st);
}

// ------------------------------------------------- long self-attention, head_dim 64 (129 <= S <= 320)
// The encoder layers at the scripts' default of 256 patches (BASELINE config 2: S = 257). Same one-CTA-per-(sequence,
// head) staging as the mid-size kernel, but a 16 x S score row-block no longer fits in registers: the key tiles are
// streamed with an online softmax (running row maximum / sum, output rescaled), flash-attention style. Backward,
// everything recomputed from Q, K, V, dO, O: phase 1 (warp = query tile) re-derives the row statistics, forms
// delta_i = dO_i . O_i (valid with dropout: O already carries the mask) and accumulates dQ; phase 2 (warp = key
// tile) walks the query tiles with the transposed products S^T = K Q^T, dP^T = V dO^T straight in accumulator
// fragments, so dK = dS^T Q and dV = P^T dO need no shared-memory copies of dS / P. No atomics, deterministic.
// 1 CTA per SM (shared memory): 16 warps hide the ldmatrix / mma latencies (8 warps: 1.24 / 3.08 ms at cfg2).
// Warp w owns tiles w, w + WARPS, ...: with 16 warps a sequence of 17 .. 19 tiles (S = 257 = 256 patches + CLS is
// the reference's film configuration) would leave one to three warps a second tile while the others idle -- twice
// the time for one more row -- so those lengths run with one warp per tile (17 .. 19 warps; the register cap of the
// launch bound falls to 120 / 112 / 104). 20 tiles (S > 304) stay on 16 warps: 20 staging tiles no longer fit.
constexpr int LONG_WARPS = 16;
constexpr int LONG_MAXT = 20;  // S <= 320

template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) attn_self_long_kernel(const AttnArgs a) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t smem_long[];
  const int S = a.Lq;
  const int NT = (S + 15) / 16;
  const int ROWS = NT * 16;
  constexpr int NTEN = MODE == 1 ? 4 : 3;
  bf16* Qs = reinterpret_cast<bf16*>(smem_long);
  bf16* Ks = Qs + ROWS * SELF_PITCH;
  bf16* Vs = Ks + ROWS * SELF_PITCH;
  bf16* Gs = Vs + ROWS * SELF_PITCH;                   // dO (backward only)
  bf16* stage_all = Qs + NTEN * ROWS * SELF_PITCH;     // one 16-row staging tile per warp
  float* row_m = reinterpret_cast<float*>(stage_all + WARPS * SELF_TILE);  // backward: per query row
  float* row_il = row_m + ROWS;
  float* row_delta = row_il + ROWS;
  uint8_t* kval = reinterpret_cast<uint8_t*>(row_delta + ROWS);                 // key j usable (in range, not padded)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = static_cast<int>(blockIdx.x) / a.H, h = static_cast<int>(blockIdx.x) % a.H;
  const uint8_t* mk = a.mask ? a.mask + static_cast<int64_t>(b % a.mask_mod) * S : nullptr;
  {
    const int col = h * SELF_HD;
    const int64_t qb = static_cast<int64_t>(a.q_mod >= a.nb ? b : b % a.q_mod) * S;
    const int64_t kb = static_cast<int64_t>(a.kv_mod >= a.nb ? b : b % a.kv_mod) * S;
    for (int idx = threadIdx.x; idx < ROWS * 8; idx += WARPS * 32) {
      const int row = idx >> 3, part = idx & 7;
      const bool valid = row < S;
      bf16* dst = Qs + row * SELF_PITCH + part * 8;
      cp_async16(dst, valid ? a.q + (qb + row) * a.ldq + col + part * 8 : a.q, valid);
      cp_async16(dst + ROWS * SELF_PITCH, valid ? a.k + (kb + row) * a.ldkv + col + part * 8 : a.q, valid);
      cp_async16(dst + 2 * ROWS * SELF_PITCH, valid ? a.v + (kb + row) * a.ldkv + col + part * 8 : a.q, valid);
      if (MODE == 1)
        cp_async16(dst + 3 * ROWS * SELF_PITCH,
                   valid ? a.dout + (static_cast<int64_t>(b) * S + row) * a.lddo + col + part * 8 : a.q, valid);
    }
    for (int j = threadIdx.x; j < ROWS; j += WARPS * 32) kval[j] = (j < S && !(mk && mk[j])) ? 1 : 0;
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  const float scale = 0.125f;
  const int g = lane >> 2, t = lane & 3;
  bf16* stage = stage_all + warp * SELF_TILE;
  const bool drop = a.drop_p > 0.f;
  uint64_t seed = 0, step = 0;
  if (drop) {
    seed = a.rng[0];
    step = a.rng[1];
  }
  const float keep_scale = drop ? 1.f / (1.f - a.drop_p) : 1.f;
  const uint64_t head_base = (static_cast<uint64_t>(b) * a.H + h) * static_cast<uint64_t>(S);  // + i -> row of P
  // this thread's 4 key columns of a key tile: j = kt*16 + (e>>1)*8 + 2t + (e&1)
  auto key_of = [&](int kt, int e) { return kt * 16 + (e >> 1) * 8 + 2 * t + (e & 1); };

  for (int qt = warp; qt < NT; qt += WARPS) {
    const bf16* Qt = Qs + qt * SELF_TILE;
    const int rows_valid = min(16, S - qt * 16);
    uint32_t qa[4][4];  // the query tile's A fragments stay in registers for both key walks
    load_a_frags16(qa, Qt, lane);
    // ---- online softmax statistics (and, forward, the output) over the key tiles
    float m_run[2] = {-INFINITY, -INFINITY}, l_thr[2] = {0.f, 0.f};
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[n][j] = 0.f;
    // rows of this thread in the precomputed keep-bit mask (rows >= S read row S - 1: their results are never stored)
    const uint64_t wrow[2] = {(head_base + min(qt * 16 + g, S - 1)) * static_cast<uint64_t>(S),
                              (head_base + min(qt * 16 + g + 8, S - 1)) * static_cast<uint64_t>(S)};
    const bool use_bits = drop && a.dbits != nullptr;
    for (int kt = 0; kt < NT; ++kt) {
      uint32_t win[2] = {0xFFFFu, 0xFFFFu};
      if (MODE == 0 && use_bits) {
        win[0] = keep_window16(a.dbits, wrow[0] + kt * 16);
        win[1] = keep_window16(a.dbits, wrow[1] + kt * 16);
      }
      float sc[2][4];
      mma_afrag_bt(sc, qa, Ks + kt * SELF_TILE, lane);
      float pv[2][4];
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        float tm = -INFINITY;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = kval[key_of(kt, e)] ? sc[e >> 1][rh * 2 + (e & 1)] * SCALE_LOG2E : -INFINITY;
          pv[rh][e] = v;
          tm = fmaxf(tm, v);
        }
        tm = quad_max(tm);
        const float m_new = fmaxf(m_run[rh], tm);  // running maximum, log2 domain
        const float m_use = m_new == -INFINITY ? 0.f : m_new;
        const float corr = fast_ex2(m_run[rh] - m_use);  // m_run = -inf -> 0
        m_run[rh] = m_new;
        float ls = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pv[rh][e] = fast_ex2(pv[rh][e] - m_use);
          ls += pv[rh][e];
        }
        l_thr[rh] = l_thr[rh] * corr + ls;
        if (MODE == 0) {
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            o[n][rh * 2] *= corr;
            o[n][rh * 2 + 1] *= corr;
          }
        }
      }
      if (MODE == 0) {
        if (drop && a.dbits) {
#pragma unroll
          for (int rh = 0; rh < 2; ++rh)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              pv[rh][e] = ((win[rh] >> ((e >> 1) * 8 + 2 * t + (e & 1))) & 1u) ? pv[rh][e] * keep_scale : 0.f;
        } else if (drop) {
          DropoutStream ds(seed, step, a.site, a.drop_p);
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            const int i = qt * 16 + g + rh * 8;
            const uint64_t pbase = (head_base + i) * static_cast<uint64_t>(S);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = key_of(kt, e);
              if (i < S && j < S) pv[rh][e] = ds.keep(pbase + j) ? pv[rh][e] * keep_scale : 0.f;
            }
          }
        }
        uint32_t pa[4];
        pa[0] = pack_bf16(pv[0][0], pv[0][1]);
        pa[1] = pack_bf16(pv[1][0], pv[1][1]);
        pa[2] = pack_bf16(pv[0][2], pv[0][3]);
        pa[3] = pack_bf16(pv[1][2], pv[1][3]);
        mma_frag_b_acc(o, pa, Vs + kt * SELF_TILE, lane);
      }
    }
    float inv_l[2];
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const float l = quad_add(l_thr[rh]);
      inv_l[rh] = l > 0.f ? 1.f / l : 0.f;
    }
    if (MODE == 0) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        o[n][0] *= inv_l[0]; o[n][1] *= inv_l[0];
        o[n][2] *= inv_l[1]; o[n][3] *= inv_l[1];
      }
      store_tile(stage, o, 1.f, a.o + (static_cast<int64_t>(b) * S + qt * 16) * a.ldo + h * SELF_HD, a.ldo, rows_valid,
                 lane);
      continue;
    }
    // ---- backward, phase 1: delta_i = dO_i . O_i, then dQ
    const bf16* Gt = Gs + qt * SELF_TILE;
    {
      const int r = lane >> 1, hf = lane & 1;  // row of the tile, half of the 64 head dims
      float acc = 0.f;
      if (r < rows_valid) {
        const bf16* orow = a.o + (static_cast<int64_t>(b) * S + qt * 16 + r) * a.ldo + h * SELF_HD + hf * 32;
        const bf16* grow = Gt + r * SELF_PITCH + hf * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 ov = __ldg(reinterpret_cast<const uint4*>(orow) + c);
          const uint4 gv = *reinterpret_cast<const uint4*>(grow + c * 8);
          const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&ov);
          const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 fx = __bfloat1622float2(x[q]), fy = __bfloat1622float2(y[q]);
            acc = fmaf(fx.x, fy.x, acc);
            acc = fmaf(fx.y, fy.y, acc);
          }
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (hf == 0) row_delta[qt * 16 + r] = acc;
    }
    if (t == 0) {
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int i = qt * 16 + g + rh * 8;
        row_m[i] = m_run[rh] == -INFINITY ? 0.f : m_run[rh];
        row_il[i] = i < S ? inv_l[rh] : 0.f;  // rows beyond S contribute nothing in phase 2
      }
    }
    __syncwarp();
    const float m_safe[2] = {m_run[0] == -INFINITY ? 0.f : m_run[0], m_run[1] == -INFINITY ? 0.f : m_run[1]};
    const float delta[2] = {row_delta[qt * 16 + g], row_delta[qt * 16 + g + 8]};
    for (int kt = 0; kt < NT; ++kt) {
      uint32_t win[2] = {0xFFFFu, 0xFFFFu};
      if (use_bits) {
        win[0] = keep_window16(a.dbits, wrow[0] + kt * 16);
        win[1] = keep_window16(a.dbits, wrow[1] + kt * 16);
      }
      float sc[2][4], dp[2][4];
      mma_afrag_bt(sc, qa, Ks + kt * SELF_TILE, lane);
      mma_ab_t(dp, Gt, Vs + kt * SELF_TILE, lane);  // dP~ = dO V^T
      float dsv[2][4];
      DropoutStream ds(seed, step, a.site, a.drop_p);
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int i = qt * 16 + g + rh * 8;
        const uint64_t pbase = (head_base + i) * static_cast<uint64_t>(S);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = key_of(kt, e);
          const float pij = kval[j] ? fast_ex2(fmaf(sc[e >> 1][rh * 2 + (e & 1)], SCALE_LOG2E, -m_safe[rh])) * inv_l[rh] : 0.f;
          float mult = 1.f;
          if (use_bits) mult = ((win[rh] >> ((e >> 1) * 8 + 2 * t + (e & 1))) & 1u) ? keep_scale : 0.f;
          else if (drop && i < S && j < S) mult = ds.keep(pbase + j) ? keep_scale : 0.f;
          dsv[rh][e] = pij * (dp[e >> 1][rh * 2 + (e & 1)] * mult - delta[rh]);
        }
      }
      uint32_t da[4];
      da[0] = pack_bf16(dsv[0][0], dsv[0][1]);
      da[1] = pack_bf16(dsv[1][0], dsv[1][1]);
      da[2] = pack_bf16(dsv[0][2], dsv[0][3]);
      da[3] = pack_bf16(dsv[1][2], dsv[1][3]);
      mma_frag_b_acc(o, da, Ks + kt * SELF_TILE, lane);  // dQ += dS K
    }
    store_tile(stage, o, scale, a.dq + (static_cast<int64_t>(b) * S + qt * 16) * a.lddq + h * SELF_HD, a.lddq,
               rows_valid, lane);
  }
  if (MODE == 0) return;
  __syncthreads();
  // ---- backward, phase 2: warp = key tile; rows of the fragments are KEYS, columns are QUERIES
  for (int kt = warp; kt < NT; kt += WARPS) {
    const int rows_valid = min(16, S - kt * 16);
    const bool kv0 = kval[kt * 16 + g] != 0, kv1 = kval[kt * 16 + g + 8] != 0;
    float ok[8][4], ov[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) ok[n][j] = ov[n][j] = 0.f;
    for (int qt = 0; qt < NT; ++qt) {
      float st[2][4], dpt[2][4];
      mma_ab_t(st, Ks + kt * SELF_TILE, Qs + qt * SELF_TILE, lane);   // S^T = K Q^T
      mma_ab_t(dpt, Vs + kt * SELF_TILE, Gs + qt * SELF_TILE, lane);  // dP~^T = V dO^T
      float pT[2][4], dsT[2][4];  // [key row half][e]: query column i = qt*16 + (e>>1)*8 + 2t + (e&1)
      DropoutStream ds(seed, step, a.site, a.drop_p);
      const bool use_bits2 = drop && a.dbits != nullptr;
      uint32_t winq[4] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu};  // keys kt*16 .. +15 of this thread's four query rows
      if (use_bits2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = min(qt * 16 + (e >> 1) * 8 + 2 * t + (e & 1), S - 1);
          winq[e] = keep_window16(a.dbits, (head_base + i) * static_cast<uint64_t>(S) + kt * 16);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = qt * 16 + (e >> 1) * 8 + 2 * t + (e & 1);
        const float mi = row_m[i], il = row_il[i], dl = row_delta[i];
        const uint64_t pbase = (head_base + i) * static_cast<uint64_t>(S);
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          const int j = kt * 16 + g + rh * 8;
          const bool kv = rh == 0 ? kv0 : kv1;
          const float pij = kv ? fast_ex2(fmaf(st[e >> 1][rh * 2 + (e & 1)], SCALE_LOG2E, -mi)) * il : 0.f;
          float mult = 1.f;
          if (use_bits2) mult = ((winq[e] >> (g + rh * 8)) & 1u) ? keep_scale : 0.f;
          else if (drop && i < S && j < S) mult = ds.keep(pbase + j) ? keep_scale : 0.f;
          pT[rh][e] = pij * mult;
          dsT[rh][e] = pij * (dpt[e >> 1][rh * 2 + (e & 1)] * mult - dl);
        }
      }
      uint32_t fa[4];
      fa[0] = pack_bf16(dsT[0][0], dsT[0][1]);
      fa[1] = pack_bf16(dsT[1][0], dsT[1][1]);
      fa[2] = pack_bf16(dsT[0][2], dsT[0][3]);
      fa[3] = pack_bf16(dsT[1][2], dsT[1][3]);
      mma_frag_b_acc(ok, fa, Qs + qt * SELF_TILE, lane);  // dK += dS^T Q
      fa[0] = pack_bf16(pT[0][0], pT[0][1]);
      fa[1] = pack_bf16(pT[1][0], pT[1][1]);
      fa[2] = pack_bf16(pT[0][2], pT[0][3]);
      fa[3] = pack_bf16(pT[1][2], pT[1][3]);
      mma_frag_b_acc(ov, fa, Gs + qt * SELF_TILE, lane);  // dV += P~^T dO
    }
    store_tile(stage, ok, scale, a.dk + (static_cast<int64_t>(b) * S + kt * 16) * a.lddkv + h * SELF_HD, a.lddkv,
               rows_valid, lane);
    store_tile(stage, ov, 1.f, a.dv + (static_cast<int64_t>(b) * S + kt * 16) * a.lddkv + h * SELF_HD, a.lddkv,
               rows_valid, lane);
  }
}

// ---- second form of the long kernel (default): register-blocked against shared-memory bandwidth.
// ncu of the kernel above at cfg2 (profiles/r02_attn_long_ncu_full_summary.txt): backward "Mem Busy" 78 %, tensor pipe 29 %:
// with one 16-row tile per warp every m16n8k16 pair is fed by its own ldmatrix.x4 (12 per 16 MMAs forward, 52 per 64
// backward). Here a warp owns TWO query tiles in the forward and in backward phase 1 — the A fragments of Q (and dO) stay in
// registers for the whole key walk and every K / V fragment load feeds both tiles — and in phase 2 the A fragments of its key
// tile (K, V) are loaded once: 4 ldmatrix per 16 MMAs forward, 24 per 64 backward. Half as many warps (9 for 257 tokens), so
// the register cap of the launch bound doubles.
__device__ __forceinline__ void load_a_frags(uint32_t (&fa)[4][4], const bf16* A, int lane) {
  const int arow = (lane & 7) + ((lane >> 3) & 1) * 8, acol = (lane >> 4) * 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldsm_x4(fa[kk], A + arow * SELF_PITCH + kk * 16 + acol);
}
// C[q] (16 x 16) = A[q] (fragments) * B^T, B = one 16-row tile [16 (n)][64 (k)] in smem, for NQ A tiles at once
template <int NQ>
__device__ __forceinline__ void mma_frags_bt(float (&c)[NQ][2][4], const uint32_t (&fa)[NQ][4][4], const bf16* B, int lane) {
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[q][nt][j] = 0.f;
  const int brow = (lane & 7) + (lane >> 4) * 8, bcol = ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t b[4];
    ldsm_x4(b, B + brow * SELF_PITCH + kk * 16 + bcol);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      mma16816(c[q][0], fa[q][kk], b[0], b[1]);
      mma16816(c[q][1], fa[q][kk], b[2], b[3]);
    }
  }
}
// O[q] (16 x 64) += P[q] (16 x 16 fragments) * B, B = [16 (k)][64 (n)] row-major in smem, for NQ tiles at once
template <int NQ>
__device__ __forceinline__ void mma_frags_b_acc(float (&o)[NQ][8][4], const uint32_t (&pa)[NQ][4], const bf16* B, int lane) {
  const int brow = (lane & 7) + ((lane >> 3) & 1) * 8, bcol = (lane >> 4) * 8;
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    uint32_t b[4];
    ldsm_x4_t(b, B + brow * SELF_PITCH + np * 16 + bcol);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      mma16816(o[q][2 * np], pa[q], b[0], b[1]);
      mma16816(o[q][2 * np + 1], pa[q], b[2], b[3]);
    }
  }
}

template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) attn_self_long2_kernel(const AttnArgs a) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t smem_long[];
  const int S = a.Lq;
  const int NT = (S + 15) / 16;
  const int ROWS = NT * 16;
  constexpr int NTEN = MODE == 1 ? 4 : 3;
  bf16* Qs = reinterpret_cast<bf16*>(smem_long);
  bf16* Ks = Qs + ROWS * SELF_PITCH;
  bf16* Vs = Ks + ROWS * SELF_PITCH;
  bf16* Gs = Vs + ROWS * SELF_PITCH;                   // dO (backward only)
  bf16* stage_all = Qs + NTEN * ROWS * SELF_PITCH;     // one 16-row staging tile per warp
  float* row_m = reinterpret_cast<float*>(stage_all + WARPS * SELF_TILE);  // backward: per query row
  float* row_il = row_m + ROWS;
  float* row_delta = row_il + ROWS;
  uint8_t* kval = reinterpret_cast<uint8_t*>(row_delta + ROWS);                 // key j usable (in range, not padded)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = static_cast<int>(blockIdx.x) / a.H, h = static_cast<int>(blockIdx.x) % a.H;
  const uint8_t* mk = a.mask ? a.mask + static_cast<int64_t>(b % a.mask_mod) * S : nullptr;
  {
    const int col = h * SELF_HD;
    const int64_t qb = static_cast<int64_t>(a.q_mod >= a.nb ? b : b % a.q_mod) * S;
    const int64_t kb = static_cast<int64_t>(a.kv_mod >= a.nb ? b : b % a.kv_mod) * S;
    for (int idx = threadIdx.x; idx < ROWS * 8; idx += WARPS * 32) {
      const int row = idx >> 3, part = idx & 7;
      const bool valid = row < S;
      bf16* dst = Qs + row * SELF_PITCH + part * 8;
      cp_async16(dst, valid ? a.q + (qb + row) * a.ldq + col + part * 8 : a.q, valid);
      cp_async16(dst + ROWS * SELF_PITCH, valid ? a.k + (kb + row) * a.ldkv + col + part * 8 : a.q, valid);
      cp_async16(dst + 2 * ROWS * SELF_PITCH, valid ? a.v + (kb + row) * a.ldkv + col + part * 8 : a.q, valid);
      if (MODE == 1)
        cp_async16(dst + 3 * ROWS * SELF_PITCH,
                   valid ? a.dout + (static_cast<int64_t>(b) * S + row) * a.lddo + col + part * 8 : a.q, valid);
    }
    for (int j = threadIdx.x; j < ROWS; j += WARPS * 32) kval[j] = (j < S && !(mk && mk[j])) ? 1 : 0;
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  const float scale = 0.125f;
  const int g = lane >> 2, t = lane & 3;
  bf16* stage = stage_all + warp * SELF_TILE;
  const bool drop = a.drop_p > 0.f;
  uint64_t seed = 0, step = 0;
  if (drop) {
    seed = a.rng[0];
    step = a.rng[1];
  }
  const bool use_bits = drop && a.dbits != nullptr;
  const float keep_scale = drop ? 1.f / (1.f - a.drop_p) : 1.f;
  const uint64_t head_base = (static_cast<uint64_t>(b) * a.H + h) * static_cast<uint64_t>(S);  // + i -> row of P
  auto key_of = [&](int kt, int e) { return kt * 16 + (e >> 1) * 8 + 2 * t + (e & 1); };
  // keep decision of element (i, j) when the masks are not precomputed
  auto keep_of = [&](int i, int j) {
    return dropout_keep(seed, step, a.site, (head_base + i) * static_cast<uint64_t>(S) + j, a.drop_p);
  };

  for (int qp = warp; qp * 2 < NT; qp += WARPS) {
    const int qt0 = qp * 2;
    const int nq = qt0 + 1 < NT ? 2 : 1;  // the last pair of an odd tile count has one tile (the other: tile qt0 again, unused)
    const int qts[2] = {qt0, nq == 2 ? qt0 + 1 : qt0};
    uint32_t qa[2][4][4];
    load_a_frags(qa[0], Qs + qts[0] * SELF_TILE, lane);
    load_a_frags(qa[1], Qs + qts[1] * SELF_TILE, lane);
    uint64_t wrow[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh)
        wrow[q][rh] = (head_base + min(qts[q] * 16 + g + rh * 8, S - 1)) * static_cast<uint64_t>(S);
    // ---- online softmax statistics (and, forward, the output) over the key tiles
    float m_run[2][2], l_thr[2][2];
    float o[2][8][4];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      m_run[q][0] = m_run[q][1] = -INFINITY;
      l_thr[q][0] = l_thr[q][1] = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[q][n][j] = 0.f;
    }
    for (int kt = 0; kt < NT; ++kt) {
      uint32_t win[2][2] = {{0xFFFFu, 0xFFFFu}, {0xFFFFu, 0xFFFFu}};
      if (MODE == 0 && use_bits) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) win[q][rh] = keep_window16(a.dbits, wrow[q][rh] + kt * 16);
      }
      uint32_t kv4 = 0;  // validity of this thread's four key columns
#pragma unroll
      for (int e = 0; e < 4; ++e) kv4 |= kval[key_of(kt, e)] ? (1u << e) : 0u;
      float sc[2][2][4];
      mma_frags_bt<2>(sc, qa, Ks + kt * SELF_TILE, lane);
      uint32_t pa[2][4];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float pv[2][4];
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          float tm = -INFINITY;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v = ((kv4 >> e) & 1u) ? sc[q][e >> 1][rh * 2 + (e & 1)] * SCALE_LOG2E : -INFINITY;
            pv[rh][e] = v;
            tm = fmaxf(tm, v);
          }
          tm = quad_max(tm);
          const float m_new = fmaxf(m_run[q][rh], tm);  // running maximum, log2 domain
          const float m_use = m_new == -INFINITY ? 0.f : m_new;
          const float corr = fast_ex2(m_run[q][rh] - m_use);  // m_run = -inf -> 0
          m_run[q][rh] = m_new;
          float ls = 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            pv[rh][e] = fast_ex2(pv[rh][e] - m_use);
            ls += pv[rh][e];
          }
          l_thr[q][rh] = l_thr[q][rh] * corr + ls;
          if (MODE == 0) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
              o[q][n][rh * 2] *= corr;
              o[q][n][rh * 2 + 1] *= corr;
            }
          }
        }
        if (MODE == 0) {
          if (use_bits) {
#pragma unroll
            for (int rh = 0; rh < 2; ++rh)
#pragma unroll
              for (int e = 0; e < 4; ++e)
                pv[rh][e] = ((win[q][rh] >> ((e >> 1) * 8 + 2 * t + (e & 1))) & 1u) ? pv[rh][e] * keep_scale : 0.f;
          } else if (drop) {
#pragma unroll
            for (int rh = 0; rh < 2; ++rh) {
              const int i = qts[q] * 16 + g + rh * 8;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = key_of(kt, e);
                if (i < S && j < S) pv[rh][e] = keep_of(i, j) ? pv[rh][e] * keep_scale : 0.f;
              }
            }
          }
          pa[q][0] = pack_bf16(pv[0][0], pv[0][1]);
          pa[q][1] = pack_bf16(pv[1][0], pv[1][1]);
          pa[q][2] = pack_bf16(pv[0][2], pv[0][3]);
          pa[q][3] = pack_bf16(pv[1][2], pv[1][3]);
        }
      }
      if (MODE == 0) mma_frags_b_acc<2>(o, pa, Vs + kt * SELF_TILE, lane);
    }
    float inv_l[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const float l = quad_add(l_thr[q][rh]);
        inv_l[q][rh] = l > 0.f ? 1.f / l : 0.f;
      }
    if (MODE == 0) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q >= nq) break;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          o[q][n][0] *= inv_l[q][0]; o[q][n][1] *= inv_l[q][0];
          o[q][n][2] *= inv_l[q][1]; o[q][n][3] *= inv_l[q][1];
        }
        store_tile(stage, o[q], 1.f, a.o + (static_cast<int64_t>(b) * S + qts[q] * 16) * a.ldo + h * SELF_HD, a.ldo,
                   min(16, S - qts[q] * 16), lane);
      }
      continue;
    }
    // ---- backward, phase 1: delta_i = dO_i . O_i, then dQ for both tiles
    float m_safe[2][2], delta[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int qt = qts[q];
      const int rows_valid = min(16, S - qt * 16);
      const bf16* Gt = Gs + qt * SELF_TILE;
      {
        const int r = lane >> 1, hf = lane & 1;  // row of the tile, half of the 64 head dims
        float acc = 0.f;
        if (r < rows_valid) {
          const bf16* orow = a.o + (static_cast<int64_t>(b) * S + qt * 16 + r) * a.ldo + h * SELF_HD + hf * 32;
          const bf16* grow = Gt + r * SELF_PITCH + hf * 32;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 ov = __ldg(reinterpret_cast<const uint4*>(orow) + c);
            const uint4 gv = *reinterpret_cast<const uint4*>(grow + c * 8);
            const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&ov);
            const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 fx = __bfloat1622float2(x[k]), fy = __bfloat1622float2(y[k]);
              acc = fmaf(fx.x, fy.x, acc);
              acc = fmaf(fx.y, fy.y, acc);
            }
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (hf == 0 && q < nq) row_delta[qt * 16 + r] = acc;
      }
      if (t == 0 && q < nq) {
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          const int i = qt * 16 + g + rh * 8;
          row_m[i] = m_run[q][rh] == -INFINITY ? 0.f : m_run[q][rh];
          row_il[i] = i < S ? inv_l[q][rh] : 0.f;  // rows beyond S contribute nothing in phase 2
        }
      }
      __syncwarp();
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        m_safe[q][rh] = m_run[q][rh] == -INFINITY ? 0.f : m_run[q][rh];
        delta[q][rh] = row_delta[qt * 16 + g + rh * 8];
      }
    }
    // the dO fragments stay in registers too when the launch bound leaves room (<= 8 warps: 255 registers per thread; 9
    // and 10 warps put three warps on one scheduler and cap the kernel at 168)
    constexpr bool HOIST_G = WARPS <= 8;
    uint32_t ga[2][4][4];
    if (HOIST_G) {
      load_a_frags(ga[0], Gs + qts[0] * SELF_TILE, lane);
      load_a_frags(ga[1], Gs + qts[1] * SELF_TILE, lane);
    }
    for (int kt = 0; kt < NT; ++kt) {
      uint32_t win[2][2] = {{0xFFFFu, 0xFFFFu}, {0xFFFFu, 0xFFFFu}};
      if (use_bits) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) win[q][rh] = keep_window16(a.dbits, wrow[q][rh] + kt * 16);
      }
      uint32_t kv4 = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) kv4 |= kval[key_of(kt, e)] ? (1u << e) : 0u;
      float sc[2][2][4], dp[2][2][4];
      mma_frags_bt<2>(sc, qa, Ks + kt * SELF_TILE, lane);
      if (!HOIST_G) {
        load_a_frags(ga[0], Gs + qts[0] * SELF_TILE, lane);
        load_a_frags(ga[1], Gs + qts[1] * SELF_TILE, lane);
      }
      mma_frags_bt<2>(dp, ga, Vs + kt * SELF_TILE, lane);  // dP~ = dO V^T
      uint32_t da[2][4];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float dsv[2][4];
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          const int i = qts[q] * 16 + g + rh * 8;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pij = ((kv4 >> e) & 1u)
                                  ? fast_ex2(fmaf(sc[q][e >> 1][rh * 2 + (e & 1)], SCALE_LOG2E, -m_safe[q][rh])) * inv_l[q][rh]
                                  : 0.f;
            float mult = 1.f;
            if (use_bits) {
              mult = ((win[q][rh] >> ((e >> 1) * 8 + 2 * t + (e & 1))) & 1u) ? keep_scale : 0.f;
            } else if (drop) {
              const int j = key_of(kt, e);
              if (i < S && j < S) mult = keep_of(i, j) ? keep_scale : 0.f;
            }
            dsv[rh][e] = pij * (dp[q][e >> 1][rh * 2 + (e & 1)] * mult - delta[q][rh]);
          }
        }
        da[q][0] = pack_bf16(dsv[0][0], dsv[0][1]);
        da[q][1] = pack_bf16(dsv[1][0], dsv[1][1]);
        da[q][2] = pack_bf16(dsv[0][2], dsv[0][3]);
        da[q][3] = pack_bf16(dsv[1][2], dsv[1][3]);
      }
      mma_frags_b_acc<2>(o, da, Ks + kt * SELF_TILE, lane);  // dQ += dS K
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (q >= nq) break;
      store_tile(stage, o[q], scale, a.dq + (static_cast<int64_t>(b) * S + qts[q] * 16) * a.lddq + h * SELF_HD, a.lddq,
                 min(16, S - qts[q] * 16), lane);
    }
  }
  if (MODE == 0) return;
  __syncthreads();
  // ---- backward, phase 2: warp = key tile (A fragments of K and V held in registers); rows of the score fragments are
  // KEYS, columns are QUERIES
  for (int kt = warp; kt < NT; kt += WARPS) {
    const int rows_valid = min(16, S - kt * 16);
    const bool kv0 = kval[kt * 16 + g] != 0, kv1 = kval[kt * 16 + g + 8] != 0;
    uint32_t ka[1][4][4], va[1][4][4];
    load_a_frags(ka[0], Ks + kt * SELF_TILE, lane);
    load_a_frags(va[0], Vs + kt * SELF_TILE, lane);
    float ok[1][8][4], ov[1][8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) ok[0][n][j] = ov[0][n][j] = 0.f;
    for (int qt = 0; qt < NT; ++qt) {
      uint32_t winq[4] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu};  // keys kt*16 .. +15 of this thread's four query rows
      if (use_bits) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = min(qt * 16 + (e >> 1) * 8 + 2 * t + (e & 1), S - 1);
          winq[e] = keep_window16(a.dbits, (head_base + i) * static_cast<uint64_t>(S) + kt * 16);
        }
      }
      float st[1][2][4], dpt[1][2][4];
      mma_frags_bt<1>(st, ka, Qs + qt * SELF_TILE, lane);   // S^T = K Q^T
      mma_frags_bt<1>(dpt, va, Gs + qt * SELF_TILE, lane);  // dP~^T = V dO^T
      float pT[2][4], dsT[2][4];  // [key row half][e]: query column i = qt*16 + (e>>1)*8 + 2t + (e&1)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = qt * 16 + (e >> 1) * 8 + 2 * t + (e & 1);
        const float mi = row_m[i], il = row_il[i], dl = row_delta[i];
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          const int j = kt * 16 + g + rh * 8;
          const bool kv = rh == 0 ? kv0 : kv1;
          const float pij = kv ? fast_ex2(fmaf(st[0][e >> 1][rh * 2 + (e & 1)], SCALE_LOG2E, -mi)) * il : 0.f;
          float mult = 1.f;
          if (use_bits) mult = ((winq[e] >> (g + rh * 8)) & 1u) ? keep_scale : 0.f;
          else if (drop && i < S && j < S) mult = keep_of(i, j) ? keep_scale : 0.f;
          pT[rh][e] = pij * mult;
          dsT[rh][e] = pij * (dpt[0][e >> 1][rh * 2 + (e & 1)] * mult - dl);
        }
      }
      uint32_t fa[1][4];
      fa[0][0] = pack_bf16(dsT[0][0], dsT[0][1]);
      fa[0][1] = pack_bf16(dsT[1][0], dsT[1][1]);
      fa[0][2] = pack_bf16(dsT[0][2], dsT[0][3]);
      fa[0][3] = pack_bf16(dsT[1][2], dsT[1][3]);
      mma_frags_b_acc<1>(ok, fa, Qs + qt * SELF_TILE, lane);  // dK += dS^T Q
      fa[0][0] = pack_bf16(pT[0][0], pT[0][1]);
      fa[0][1] = pack_bf16(pT[1][0], pT[1][1]);
      fa[0][2] = pack_bf16(pT[0][2], pT[0][3]);
      fa[0][3] = pack_bf16(pT[1][2], pT[1][3]);
      mma_frags_b_acc<1>(ov, fa, Gs + qt * SELF_TILE, lane);  // dV += P~^T dO
    }
    store_tile(stage, ok[0], scale, a.dk + (static_cast<int64_t>(b) * S + kt * 16) * a.lddkv + h * SELF_HD, a.lddkv,
               rows_valid, lane);
    store_tile(stage, ov[0], 1.f, a.dv + (static_cast<int64_t>(b) * S + kt * 16) * a.lddkv + h * SELF_HD, a.lddkv,
               rows_valid, lane);
  }
}

template <int MODE, int WARPS>
static int launch_long2_w(const AttnArgs& a, cudaStream_t st) {
  const int rows = (a.Lq + 15) / 16 * 16;
  const size_t smem = static_cast<size_t>((MODE == 1 ? 4 : 3) * rows * SELF_PITCH + WARPS * SELF_TILE) * 2 +
                      static_cast<size_t>(rows) * (3 * 4 + 1) + 16;
  static size_t configured = 0;
  if (smem > configured) {
    GG_CUDA_CHECK(cudaFuncSetAttribute(attn_self_long2_kernel<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    configured = smem;
  }
  launch_k(attn_self_long2_kernel<MODE, WARPS>, static_cast<unsigned>(a.nb) * a.H, WARPS * 32, smem, st, a);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

template <int MODE, int WARPS>
static int launch_long_w(const AttnArgs& a, cudaStream_t st) {
  const int rows = (a.Lq + 15) / 16 * 16;
  const size_t smem = static_cast<size_t>((MODE == 1 ? 4 : 3) * rows * SELF_PITCH + WARPS * SELF_TILE) * 2 +
                      static_cast<size_t>(rows) * (3 * 4 + 1) + 16;
  static size_t configured = 0;
  if (smem > configured) {
    GG_CUDA_CHECK(cudaFuncSetAttribute(attn_self_long_kernel<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    configured = smem;
  }
  launch_k(attn_self_long_kernel<MODE, WARPS>, static_cast<unsigned>(a.nb) * a.H, WARPS * 32, smem, st, a);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

template <int MODE>
static int launch_long(const AttnArgs& a, cudaStream_t st) {
  GG_REQUIRE(MODE == 0 || a.o != nullptr, "long self-attention backward needs the forward output (a.o)");
  // Which form runs where is measured (tests/gpu_attn_bench.py sweep, 768 sequences, dropout on, us old -> new):
  //   forward   129: 226 -> 179   161: 273 -> 224   193: 359 -> 382   257: 542 -> 628   289: 630 -> 677   320: 773 -> 730
  //   backward  129: 481 -> 567   161: 592 -> 704   225: 940 -> 932   257: 1501 -> 1281  289: 1831 -> 1552  320: 1895 -> 1653
  // (two tiles per warp halve the warps of the CTA: the forward of 13 .. 19 tiles misses them more than it gains from the
  // fragment reuse; the backward, bound by the shared-memory pipe, gains from 15 tiles on).
  // GEMMGAN_ATTN_LONG2=0 / 1 forces the first / second form everywhere.
  static const int force = [] { const char* v = getenv("GEMMGAN_ATTN_LONG2"); return v ? (v[0] == '0' ? 0 : 1) : -1; }();
  const int nt = (a.Lq + 15) / 16;
  const bool long2 = force >= 0 ? force == 1 : (MODE == 0 ? (nt <= 11 || nt == 20) : nt >= 15);
  if (long2) {
    switch (((a.Lq + 15) / 16 + 1) / 2) {  // one warp per pair of query tiles
      case 5: return launch_long2_w<MODE, 5>(a, st);
      case 6: return launch_long2_w<MODE, 6>(a, st);
      case 7: return launch_long2_w<MODE, 7>(a, st);
      case 8: return launch_long2_w<MODE, 8>(a, st);
      case 9: return launch_long2_w<MODE, 9>(a, st);
      default: return launch_long2_w<MODE, 10>(a, st);
    }
  }
  switch ((a.Lq + 15) / 16) {  // one warp per tile where a 17th .. 19th tile would otherwise cost a second round
    case 17: return launch_long_w<MODE, 17>(a, st);
    case 18: return launch_long_w<MODE, 18>(a, st);
    case 19: return launch_long_w<MODE, 19>(a, st);
    default: return launch_long_w<MODE, LONG_WARPS>(a, st);
  }
}

// ------------------------------------------------- single-query cross-attention, head_dim 64, Lk <= 128
// patch2text / text2patch attention of the paper model (Lq = 1; :149-152) at 64 patches / 32 text tokens. One WARP
// per (row, head): lane l scores keys l, l+32, ... from their own 128-byte K rows, softmax by warp shuffles, the
// output accumulates the V rows coalesced (lane = 2 head dims). The backward is the same walk (p recomputed):
// dv_j = p_j dO, dk_j = ds_j q / 8, dq = sum_j ds_j k_j / 8 with ds_j = p_j (dO.v_j - sum_i p_i dO.v_i).
// No dropout on these attentions (nn.MultiheadAttention default, :120-123).
constexpr int Q1_MAXK = 4;  // keys per lane

__device__ __forceinline__ float dot64(const uint4 (&a)[8], const bf16* row) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 kv = __ldg(reinterpret_cast<const uint4*>(row) + c);
    const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a[c]);
    const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&kv);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 fx = __bfloat1622float2(x[t]), fy = __bfloat1622float2(y[t]);
      acc = fmaf(fx.x, fy.x, acc);
      acc = fmaf(fx.y, fy.y, acc);
    }
  }
  return acc;
}

template <int MODE>
__global__ void __launch_bounds__(128) attn_q1_kernel(const AttnArgs a) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (gid >= static_cast<int64_t>(a.nb) * a.H) return;
  const int b = static_cast<int>(gid / a.H), h = static_cast<int>(gid % a.H);
  const int Lk = a.Lk;
  const int64_t qrow = a.q_mod >= a.nb ? b : b % a.q_mod;
  const int64_t kb = static_cast<int64_t>(a.kv_mod >= a.nb ? b : b % a.kv_mod) * Lk;
  const bf16* qp = a.q + qrow * a.ldq + h * 64;
  const bf16* Kp = a.k + kb * a.ldkv + h * 64;
  const bf16* Vp = a.v + kb * a.ldkv + h * 64;
  const uint8_t* mk = a.mask ? a.mask + static_cast<int64_t>(b % a.mask_mod) * Lk : nullptr;
  uint4 qv[8];  // the whole query row (and dO row) in every lane: 128 bytes each
#pragma unroll
  for (int c = 0; c < 8; ++c) qv[c] = __ldg(reinterpret_cast<const uint4*>(qp) + c);
  float sc[Q1_MAXK];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < Q1_MAXK; ++i) {
    const int j = lane + 32 * i;
    sc[i] = -INFINITY;
    if (j < Lk && !(mk && mk[j])) sc[i] = dot64(qv, Kp + static_cast<int64_t>(j) * a.ldkv) * 0.125f;
    m = fmaxf(m, sc[i]);
  }
  m = wmax(m);
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < Q1_MAXK; ++i) {
    sc[i] = sc[i] == -INFINITY ? 0.f : __expf(sc[i] - m);
    l += sc[i];
  }
  l = wsum2(l);
  const float inv_l = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
  for (int i = 0; i < Q1_MAXK; ++i) sc[i] *= inv_l;  // p_j of key lane + 32 i
  if (MODE == 0) {
    float2 acc = make_float2(0.f, 0.f);
    for (int j = 0; j < Lk; ++j) {
      const float pj = __shfl_sync(0xffffffffu, sc[j >> 5], j & 31);
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Vp + static_cast<int64_t>(j) * a.ldkv + 2 * lane));
      acc.x = fmaf(pj, v.x, acc.x);
      acc.y = fmaf(pj, v.y, acc.y);
    }
    *reinterpret_cast<__nv_bfloat162*>(a.o + static_cast<int64_t>(b) * a.ldo + h * 64 + 2 * lane) = __floats2bfloat162_rn(acc.x, acc.y);
    return;
  }
  // ---- backward
  const bf16* gp = a.dout + static_cast<int64_t>(b) * a.lddo + h * 64;
  uint4 gv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) gv[c] = __ldg(reinterpret_cast<const uint4*>(gp) + c);
  float dp[Q1_MAXK];
  float delta = 0.f;
#pragma unroll
  for (int i = 0; i < Q1_MAXK; ++i) {
    const int j = lane + 32 * i;
    dp[i] = (j < Lk && sc[i] != 0.f) ? dot64(gv, Vp + static_cast<int64_t>(j) * a.ldkv) : 0.f;
    delta = fmaf(sc[i], dp[i], delta);
  }
  delta = wsum2(delta);
#pragma unroll
  for (int i = 0; i < Q1_MAXK; ++i) dp[i] = sc[i] * (dp[i] - delta) * 0.125f;  // ds_j / 8
  const float2 g2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gp + 2 * lane));
  const float2 q2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qp + 2 * lane));
  bf16* dKp = a.dk + static_cast<int64_t>(b) * Lk * a.lddkv + h * 64;
  bf16* dVp = a.dv + static_cast<int64_t>(b) * Lk * a.lddkv + h * 64;
  float2 dq = make_float2(0.f, 0.f);
  for (int j = 0; j < Lk; ++j) {
    const float pj = __shfl_sync(0xffffffffu, sc[j >> 5], j & 31);
    const float dsj = __shfl_sync(0xffffffffu, dp[j >> 5], j & 31);
    const float2 kx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Kp + static_cast<int64_t>(j) * a.ldkv + 2 * lane));
    dq.x = fmaf(dsj, kx.x, dq.x);
    dq.y = fmaf(dsj, kx.y, dq.y);
    *reinterpret_cast<__nv_bfloat162*>(dKp + static_cast<int64_t>(j) * a.lddkv + 2 * lane) = __floats2bfloat162_rn(dsj * q2.x, dsj * q2.y);
    *reinterpret_cast<__nv_bfloat162*>(dVp + static_cast<int64_t>(j) * a.lddkv + 2 * lane) = __floats2bfloat162_rn(pj * g2.x, pj * g2.y);
  }
  *reinterpret_cast<__nv_bfloat162*>(a.dq + static_cast<int64_t>(b) * a.lddq + h * 64 + 2 * lane) = __floats2bfloat162_rn(dq.x, dq.y);
}

static bool aligned16(const void* p);
static bool q1_path(const AttnArgs& a) {
  // Also for <= 16 keys (the patch2text attention over 9 tokens at cfg3): one launch forward and ONE backward instead of the
  // generic short-sequence pair (25 + 30 us on the dependent chain); 6.77 -> 6.64 ms per train(). GEMMGAN_Q1_SMALL=0: off.
  static const bool q1_small = [] { const char* v = getenv("GEMMGAN_Q1_SMALL"); return !(v && v[0] == '0'); }();
  if (!(a.Lq == 1 && (a.Lk > SM_MAXL || q1_small) && a.Lk <= 32 * Q1_MAXK && a.hd == 64 && a.drop_p == 0.f)) return false;
  return a.ldq % 8 == 0 && a.ldkv % 8 == 0 && aligned16(a.q) && aligned16(a.k) && aligned16(a.v) &&
         (!a.dout || (a.lddo % 8 == 0 && aligned16(a.dout)));
}
template <int MODE>
static int launch_q1(const AttnArgs& a, cudaStream_t st) {
  const int64_t groups = static_cast<int64_t>(a.nb) * a.H;
  launch_k(attn_q1_kernel<MODE>, static_cast<unsigned>((groups + 3) / 4), 128, 0, st, a);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

static size_t self_smem_bytes(int mode) {
  const int nt = mode == 1 ? 4 : 3;
  return static_cast<size_t>(SELF_GROUPS) * ((nt + 1) * SELF_TILE + (mode == 1 ? 2 * 16 * SELF_SP : 0)) * 2 + 16;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool small_path(const AttnArgs& a) {
  const int dpt = a.hd / 4;
  if (!(a.Lk <= SM_MAXL && a.hd % 8 == 0 && (dpt == 2 || dpt == 4 || dpt == 8 || dpt == 16))) return false;
  if (dpt % 8 == 0) {  // the 16-byte path needs aligned bases and pitches
    const bool ok = a.ldq % 8 == 0 && a.ldkv % 8 == 0 && aligned16(a.q) && aligned16(a.k) && aligned16(a.v) &&
                    (!a.o || (a.ldo % 8 == 0 && aligned16(a.o))) &&
                    (!a.dout || (a.lddo % 8 == 0 && aligned16(a.dout) && a.lddq % 8 == 0 && aligned16(a.dq) &&
                                 a.lddkv % 8 == 0 && aligned16(a.dk) && aligned16(a.dv)));
    if (!ok) return false;
  }
  return true;
}
template <int MODE>
static void launch_small_q(const AttnArgs& a, float* stat, cudaStream_t st) {
  const int64_t threads = static_cast<int64_t>(a.nb) * a.H * a.Lq * 4;
  const unsigned grid = static_cast<unsigned>((threads + 127) / 128);
  switch (a.hd / 4) {
    case 2: launch_k(attn_small_q_kernel<2, MODE>, grid, 128, 0, st, a, stat); break;
    case 4: launch_k(attn_small_q_kernel<4, MODE>, grid, 128, 0, st, a, stat); break;
    case 8: launch_k(attn_small_q_kernel<8, MODE>, grid, 128, 0, st, a, stat); break;
    default: launch_k(attn_small_q_kernel<16, MODE>, grid, 128, 0, st, a, stat); break;
  }
}
static void launch_small_kv(const AttnArgs& a, const float* stat, cudaStream_t st) {
  const int64_t threads = static_cast<int64_t>(a.nb) * a.H * a.Lk * 4;
  const unsigned grid = static_cast<unsigned>((threads + 127) / 128);
  switch (a.hd / 4) {
    case 2: launch_k(attn_small_kv_kernel<2>, grid, 128, 0, st, a, stat); break;
    case 4: launch_k(attn_small_kv_kernel<4>, grid, 128, 0, st, a, stat); break;
    case 8: launch_k(attn_small_kv_kernel<8>, grid, 128, 0, st, a, stat); break;
    default: launch_k(attn_small_kv_kernel<16>, grid, 128, 0, st, a, stat); break;
  }
}

constexpr int MID_MAXL = 128;
static bool self_path(const AttnArgs& a, int maxl = SM_MAXL) {
  if (!(a.Lq == a.Lk && a.Lk <= maxl && a.hd == SELF_HD)) return false;
  const bool ok = a.ldq % 8 == 0 && a.ldkv % 8 == 0 && aligned16(a.q) && aligned16(a.k) && aligned16(a.v) &&
                  (!a.o || (a.ldo % 8 == 0 && aligned16(a.o))) &&
                  (!a.dout || (a.lddo % 8 == 0 && aligned16(a.dout) && a.lddq % 8 == 0 && aligned16(a.dq) &&
                               a.lddkv % 8 == 0 && aligned16(a.dk) && aligned16(a.dv)));
  return ok;
}
template <int MODE>
static int launch_self(const AttnArgs& a, cudaStream_t st) {
  const size_t smem = self_smem_bytes(MODE);
  static bool configured = false;
  if (!configured) {
    GG_CUDA_CHECK(cudaFuncSetAttribute(attn_self_small_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    configured = true;
  }
  const int64_t ngroups = static_cast<int64_t>(a.nb) * a.H;
  const unsigned grid = static_cast<unsigned>((ngroups + SELF_GROUPS - 1) / SELF_GROUPS);
  launch_k(attn_self_small_kernel<MODE>, grid, SELF_THREADS, smem, st, a);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

static int check_args(const AttnArgs& a) {
  GG_REQUIRE(a.hd >= 2 && a.hd <= 64 && a.hd % 2 == 0, "attention head_dim %d unsupported (even, <= 64)", a.hd);
  GG_REQUIRE(a.Lk >= 1 && a.Lk <= 32 * ATT_MAXC && a.Lq >= 1 && a.Lq <= 32 * ATT_MAXC,
             "attention length Lq=%d Lk=%d unsupported (<= %d)", a.Lq, a.Lk, 32 * ATT_MAXC);
  GG_REQUIRE(a.drop_p == 0.f || a.rng, "dropout needs rng state");
  return GG_OK;
}

// the same for up to three sites in one launch (the three per-element dropout sites of an encoder layer pass)
struct DropBitsJobs {
  uint32_t site[3];
  int64_t n_words[3];
  uint32_t* out[3];
};
__global__ void __launch_bounds__(256) dropout_bits3_kernel(const uint64_t* __restrict__ rng, float p, const DropBitsJobs jobs) {
  pdl_entry();
  const uint64_t seed = rng[0], step = rng[1];
  const uint32_t thr = dropout_thr(p);
  const int64_t total = jobs.n_words[0] + jobs.n_words[1] + jobs.n_words[2];
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int j = 0;
    int64_t w = i;
    if (w >= jobs.n_words[0]) { w -= jobs.n_words[0]; j = 1; }
    if (j == 1 && w >= jobs.n_words[1]) { w -= jobs.n_words[1]; j = 2; }
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      v |= keep_bits8(dropout_words(seed, step, jobs.site[j], static_cast<uint64_t>(w) * 4 + k), thr) << (8 * k);
    jobs.out[j][w] = v;
  }
}

int64_t dropout_bits_words(int64_t n_elems) { return (n_elems + 31) / 32 + 4; }
int k_dropout_bits3(const uint64_t* rng, float p, const uint32_t (&site)[3], const int64_t (&n_elems)[3], uint32_t* const (&out)[3],
                    cudaStream_t st) {
  DropBitsJobs jobs;
  int64_t total = 0;
  for (int j = 0; j < 3; ++j) {
    GG_REQUIRE(out[j] && n_elems[j] >= 0, "bad dropout-bits argument");
    jobs.site[j] = site[j];
    jobs.n_words[j] = dropout_bits_words(n_elems[j]);
    jobs.out[j] = out[j];
    total += jobs.n_words[j];
  }
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(dropout_bits3_kernel, static_cast<unsigned>(blocks), 256, 0, st, rng, p, jobs);
  GG_LAUNCH_CHECK();
  return GG_OK;
}
int k_dropout_bits(const uint64_t* rng, uint32_t site, float p, int64_t n_elems, uint32_t* out, cudaStream_t st) {
  GG_REQUIRE(rng && out && n_elems >= 0, "bad dropout-bits argument");
  const int64_t n_words = dropout_bits_words(n_elems);
  int64_t blocks = (n_words + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(dropout_bits_kernel, static_cast<unsigned>(blocks), 256, 0, st, rng, site, p, n_words, out);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

int k_attention_fwd(const AttnArgs& a, cudaStream_t st) {
  int rc = check_args(a);
  if (rc) return rc;
  if (self_path(a)) return launch_self<0>(a, st);
  if (self_path(a, MID_MAXL)) return launch_mid<0>(a, st);
  if (self_path(a, 16 * LONG_MAXT)) return launch_long<0>(a, st);
  if (q1_path(a)) return launch_q1<0>(a, st);
  if (small_path(a)) {
    launch_small_q<0>(a, nullptr, st);
    GG_LAUNCH_CHECK();
    return GG_OK;
  }
  const int pitch = a.hd + 2;
  const size_t smem = (static_cast<size_t>(2 * a.Lk * pitch + ((a.Lq * pitch + 1) & ~1))) * 2 +
                      static_cast<size_t>(ATT_WARPS) * a.Lk * 4 + 16;
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    GG_CUDA_CHECK(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  GG_REQUIRE(smem <= 200 * 1024, "attention working set too large");
  launch_k(attention_fwd_kernel, a.nb * a.H, ATT_WARPS * 32, smem, st, a);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

int k_attention_bwd(const AttnArgs& a, cudaStream_t st) {
  int rc = check_args(a);
  if (rc) return rc;
  if (self_path(a)) return launch_self<1>(a, st);
  if (self_path(a, MID_MAXL)) return launch_mid<1>(a, st);
  if (self_path(a, 16 * LONG_MAXT) && a.o) return launch_long<1>(a, st);  // needs the forward output for delta
  if (q1_path(a)) return launch_q1<1>(a, st);
  if (small_path(a)) {
    GG_REQUIRE(a.stat != nullptr, "short-sequence attention backward needs a stats scratch buffer");
    launch_small_q<1>(a, a.stat, st);
    GG_LAUNCH_CHECK();
    launch_small_kv(a, a.stat, st);
    GG_LAUNCH_CHECK();
    return GG_OK;
  }
  const int pitch = a.hd + 2;
  const int Lmax = a.Lk > a.Lq ? a.Lk : a.Lq;
  const size_t smem = (static_cast<size_t>(2 * a.Lk * pitch + a.Lq * pitch + ((a.Lq * pitch + 1) & ~1))) * 2 +
                      static_cast<size_t>(2 * a.Lq) * 4 + static_cast<size_t>(ATT_WARPS) * 2 * Lmax * 4 + 16;
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    GG_CUDA_CHECK(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = 220 * 1024;
  }
  GG_REQUIRE(smem <= 220 * 1024, "attention backward working set too large");
  launch_k(attention_bwd_kernel, a.nb * a.H, ATT_WARPS * 32, smem, st, a);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg
