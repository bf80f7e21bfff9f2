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
// One thread per (sequence, head, token) row with all 64 head dims in registers; a CTA owns SELF_GROUPS
// (sequence, head) groups whose Q / K / V (/ dO) rows are staged in shared memory with coalesced 16-byte
// loads (row pitch 144 B: conflict-free 16-byte reads). Scores never leave registers. The backward is one
// kernel: pass A (thread = query row) recomputes the probabilities, writes dQ and leaves dS / dropped P in
// shared memory; pass B (thread = key row) forms dK, dV from them (no atomics, deterministic).
constexpr int SELF_HD = 64;
constexpr int SELF_PITCH = SELF_HD + 8;  // bf16 elements
constexpr int SELF_THREADS = 128;

struct SelfGeom {
  int G;       // groups per CTA
  int rows;    // G * S
};
__host__ __device__ inline SelfGeom self_geom(int S) {
  SelfGeom g;
  g.G = SELF_THREADS / S;
  g.rows = g.G * S;
  return g;
}

// cooperative copy of nrows x 64 bf16 rows of up to four tensors into smem (pitch SELF_PITCH) with 16-byte
// cp.async (no register staging: every load of the CTA is in flight at once). off(t, r) = element offset of
// row r of tensor t (head column included), or -1 beyond the last group (zero-filled).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
template <int NT, class OffFn>
__device__ __forceinline__ void self_stage(bf16* const* dst, const bf16* const* src, int nrows, OffFn off) {
  const int n = nrows * 8;
  for (int c = threadIdx.x; c < n; c += SELF_THREADS) {
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int64_t o = off(t, c >> 3);
      cp_async16(dst[t] + (c >> 3) * SELF_PITCH + (c & 7) * 8, src[t] + (o >= 0 ? o : 0) + (c & 7) * 8, o >= 0);
    }
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ void self_row_f32(const bf16* row, float* v) {
#pragma unroll
  for (int d = 0; d < SELF_HD; d += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + d);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = __bfloat1622float2(h[q]);
      v[d + 2 * q] = f.x;
      v[d + 2 * q + 1] = f.y;
    }
  }
}
__device__ __forceinline__ float self_dot(const float* q, const bf16* row) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // four independent chains
#pragma unroll
  for (int d = 0; d < SELF_HD; d += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + d);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __bfloat1622float2(h[t]);
      acc[t] = fmaf(q[d + 2 * t], f.x, acc[t]);
      acc[t] = fmaf(q[d + 2 * t + 1], f.y, acc[t]);
    }
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}
__device__ __forceinline__ void self_axpy(float* acc, float w, const bf16* row) {
#pragma unroll
  for (int d = 0; d < SELF_HD; d += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + d);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __bfloat1622float2(h[t]);
      acc[d + 2 * t] = fmaf(w, f.x, acc[d + 2 * t]);
      acc[d + 2 * t + 1] = fmaf(w, f.y, acc[d + 2 * t + 1]);
    }
  }
}
__device__ __forceinline__ void self_store_row(bf16* dst, const float* v, float scale) {
#pragma unroll
  for (int d = 0; d < SELF_HD; d += 8) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[d + 2 * t] * scale, v[d + 2 * t + 1] * scale);
    *reinterpret_cast<uint4*>(dst + d) = u;
  }
}

// MODE 0: forward. MODE 1: backward (dq, dk, dv).
template <int MODE>
__global__ void __launch_bounds__(SELF_THREADS) attn_self_small_kernel(const AttnArgs a) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t smem_self[];
  const int S = a.Lq;
  const SelfGeom geo = self_geom(S);
  bf16* Qs = reinterpret_cast<bf16*>(smem_self);
  bf16* Ks = Qs + geo.rows * SELF_PITCH;
  bf16* Vs = Ks + geo.rows * SELF_PITCH;
  bf16* Gs = Vs + geo.rows * SELF_PITCH;                                    // dO (backward only)
  float* sc = reinterpret_cast<float*>(Gs + (MODE == 1 ? geo.rows * SELF_PITCH : 0));  // [2][SM_MAXL][threads]
  float* dS = sc + 2 * SM_MAXL * SELF_THREADS;                                          // [G][S][S] x 2
  const int64_t ngroups = static_cast<int64_t>(a.nb) * a.H;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * geo.G;
  auto off = [&](int t, int r) -> int64_t {
    const int64_t gid = g0 + r / S;
    if (gid >= ngroups) return -1;
    const int64_t bb = gid / a.H, col = (gid % a.H) * SELF_HD;
    const int tok = r % S;
    if (t == 0) return ((bb % a.q_mod) * S + tok) * a.ldq + col;
    if (t == 3) return (bb * S + tok) * a.lddo + col;
    return ((bb % a.kv_mod) * S + tok) * a.ldkv + col;
  };
  {
    bf16* dsts[4] = {Qs, Ks, Vs, Gs};
    const bf16* srcs[4] = {a.q, a.k, a.v, a.dout};
    if (MODE == 1) self_stage<4>(dsts, srcs, geo.rows, off);
    else self_stage<3>(dsts, srcs, geo.rows, off);
  }
  __syncthreads();
  const int r = threadIdx.x;
  const int g = r / S, i = r % S;
  const int64_t gid = g0 + g;
  const bool active = r < geo.rows && gid < ngroups;
  const int b = active ? static_cast<int>(gid / a.H) : 0;
  const int h = active ? static_cast<int>(gid % a.H) : 0;
  const uint8_t* mk = (a.mask && active) ? a.mask + static_cast<int64_t>(b % a.mask_mod) * S : nullptr;
  const float scale = 0.125f;  // 1/sqrt(64)
  uint64_t seed = 0, step = 0;
  if (a.drop_p > 0.f) {
    seed = a.rng[0];
    step = a.rng[1];
  }
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  const bf16* Kg = Ks + g * S * SELF_PITCH;
  const bf16* Vg = Vs + g * S * SELF_PITCH;
  // per-thread score rows live in shared memory (column = thread => conflict-free), so the key loops stay
  // rolled: fully unrolled they are ~50 KB of SASS and thrash the instruction cache
  float* pr = sc + threadIdx.x;   // p_j   at pr[j * SELF_THREADS]
  float* pd = pr + SM_MAXL * SELF_THREADS;  // dropped p_j (forward weight) / dP_j (backward)
  if (active) {
    float q[SELF_HD];
    self_row_f32(Qs + r * SELF_PITCH, q);
    float m = -INFINITY;
#pragma unroll 1
    for (int j = 0; j < S; ++j) {
      float t = -INFINITY;
      if (!(mk && mk[j])) t = self_dot(q, Kg + j * SELF_PITCH) * scale;
      pr[j * SELF_THREADS] = t;
      m = fmaxf(m, t);
    }
    float l = 0.f;
#pragma unroll 1
    for (int j = 0; j < S; ++j) {
      const float t = pr[j * SELF_THREADS];
      const float e = (t == -INFINITY) ? 0.f : __expf(t - m);
      pr[j * SELF_THREADS] = e;
      l += e;
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    const uint64_t pbase = ((static_cast<uint64_t>(b) * a.H + h) * S + i) * static_cast<uint64_t>(S);
#pragma unroll 1
    for (int j = 0; j < S; ++j) {
      const float pj = pr[j * SELF_THREADS] * inv_l;
      const bool keep = a.drop_p > 0.f ? dropout_keep(seed, step, a.site, pbase + j, a.drop_p) : true;
      pr[j * SELF_THREADS] = pj;
      pd[j * SELF_THREADS] = keep ? keep_scale : 0.f;  // dropout multiplier
    }
  }
  if (MODE == 0) {
    if (active) {
      float acc[SELF_HD];
#pragma unroll
      for (int d = 0; d < SELF_HD; ++d) acc[d] = 0.f;
#pragma unroll 1
      for (int j = 0; j < S; ++j) self_axpy(acc, pr[j * SELF_THREADS] * pd[j * SELF_THREADS], Vg + j * SELF_PITCH);
      self_store_row(a.o + (static_cast<int64_t>(b) * S + i) * a.ldo + h * SELF_HD, acc, 1.f);
    }
    return;
  }
  // ---- backward pass A: thread = query row
  float* dSg = dS + g * 2 * S * S;
  if (active) {
    float go[SELF_HD];
    self_row_f32(Gs + r * SELF_PITCH, go);
    float delta = 0.f;
#pragma unroll 1
    for (int j = 0; j < S; ++j) {
      const float mult = pd[j * SELF_THREADS];
      const float dpj = self_dot(go, Vg + j * SELF_PITCH) * mult;   // dP_ij through the dropout mask
      delta = fmaf(pr[j * SELF_THREADS], dpj, delta);
      dSg[S * S + i * S + j] = pr[j * SELF_THREADS] * mult;          // dropped probability (for dV)
      pd[j * SELF_THREADS] = dpj;
    }
    float acc[SELF_HD];
#pragma unroll
    for (int d = 0; d < SELF_HD; ++d) acc[d] = 0.f;
#pragma unroll 1
    for (int j = 0; j < S; ++j) {
      const float ds = pr[j * SELF_THREADS] * (pd[j * SELF_THREADS] - delta);
      self_axpy(acc, ds, Kg + j * SELF_PITCH);
      dSg[i * S + j] = ds;
    }
    self_store_row(a.dq + (static_cast<int64_t>(b) * S + i) * a.lddq + h * SELF_HD, acc, scale);
  }
  __syncthreads();
  // ---- pass B: thread = key row j (= i)
  if (active) {
    const int j = i;
    float dk[SELF_HD], dv[SELF_HD];
#pragma unroll
    for (int d = 0; d < SELF_HD; ++d) dk[d] = dv[d] = 0.f;
    const bf16* Qg = Qs + g * S * SELF_PITCH;
    const bf16* Gg = Gs + g * S * SELF_PITCH;
#pragma unroll 1
    for (int ii = 0; ii < S; ++ii) {
      self_axpy(dk, dSg[ii * S + j], Qg + ii * SELF_PITCH);
      self_axpy(dv, dSg[S * S + ii * S + j], Gg + ii * SELF_PITCH);
    }
    const int64_t row = static_cast<int64_t>(b) * S + j;
    self_store_row(a.dk + row * a.lddkv + h * SELF_HD, dk, scale);
    self_store_row(a.dv + row * a.lddkv + h * SELF_HD, dv, 1.f);
  }
}

static size_t self_smem_bytes(int S, int mode) {
  const SelfGeom geo = self_geom(S);
  size_t b = static_cast<size_t>(mode == 1 ? 4 : 3) * geo.rows * SELF_PITCH * 2;
  b += static_cast<size_t>(2) * SM_MAXL * SELF_THREADS * 4;
  if (mode == 1) b += static_cast<size_t>(geo.G) * 2 * S * S * 4;
  return b + 16;
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

static bool self_path(const AttnArgs& a) {
  if (!(a.Lq == a.Lk && a.Lk <= SM_MAXL && a.hd == SELF_HD)) return false;
  const bool ok = a.ldq % 8 == 0 && a.ldkv % 8 == 0 && aligned16(a.q) && aligned16(a.k) && aligned16(a.v) &&
                  (!a.o || (a.ldo % 8 == 0 && aligned16(a.o))) &&
                  (!a.dout || (a.lddo % 8 == 0 && aligned16(a.dout) && a.lddq % 8 == 0 && aligned16(a.dq) &&
                               a.lddkv % 8 == 0 && aligned16(a.dk) && aligned16(a.dv)));
  return ok;
}
template <int MODE>
static int launch_self(const AttnArgs& a, cudaStream_t st) {
  const SelfGeom geo = self_geom(a.Lq);
  const size_t smem = self_smem_bytes(a.Lq, MODE);
  static bool configured = false;
  if (!configured) {
    GG_CUDA_CHECK(cudaFuncSetAttribute(attn_self_small_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(self_smem_bytes(SM_MAXL, MODE))));
    configured = true;
  }
  const int64_t ngroups = static_cast<int64_t>(a.nb) * a.H;
  const unsigned grid = static_cast<unsigned>((ngroups + geo.G - 1) / geo.G);
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

int k_attention_fwd(const AttnArgs& a, cudaStream_t st) {
  int rc = check_args(a);
  if (rc) return rc;
  if (self_path(a)) return launch_self<0>(a, st);
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
