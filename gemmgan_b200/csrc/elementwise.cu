// HBM-bound glue kernels of the training step: casts, FiLM modulation, token assembly, replica
// sums, deterministic column sums (bias gradients), and the small [B,H]-sized pieces of the
// critic trunk / gradient-penalty math. All are vectorised where alignment allows and sized so
// consecutive threads touch consecutive addresses.
//
// Reference call sites replaced (src/conditional_gan_cross_attention_with_film.py):
//   FiLM gamma*x+beta :136, torch.cat(cls, patches) :142, interpolation :358,
//   grad_norm / (norm-1)^2 mean :372-374, D_loss / G_loss :32-46.
#include "host_util.h"
#include "pdl.cuh"
#include "kernels.h"

namespace gg {

static inline unsigned grid_for(int64_t work, int block, int64_t cap = 148LL * 16) {
  int64_t g = (work + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

// ----------------------------------------------------------------------------------- cast
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, int64_t ld_src, bf16* __restrict__ dst,
                                     int64_t ld_dst, int64_t rows, int cols, int vec) {
  pdl_entry();
  if (vec) {
    const int c4 = cols >> 2;
    const int64_t total = rows * c4;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t r = i / c4;
      const int c = static_cast<int>(i % c4) * 4;
      const float4 f = __ldg(reinterpret_cast<const float4*>(src + r * ld_src + c));
      __nv_bfloat162 lo = __floats2bfloat162_rn(f.x, f.y), hi = __floats2bfloat162_rn(f.z, f.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = pk;
    }
  } else {
    const int64_t total = rows * cols;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t r = i / cols;
      const int c = static_cast<int>(i % cols);
      dst[r * ld_dst + c] = __float2bfloat16_rn(src[r * ld_src + c]);
    }
  }
}

int k_cast_f32_bf16(const float* src, int64_t ld_src, bf16* dst, int64_t ld_dst, int64_t rows, int cols,
                    cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return GG_OK;
  const int vec = (cols % 4 == 0) && (ld_src % 4 == 0) && (ld_dst % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
  const int64_t work = vec ? rows * (cols / 4) : rows * cols;
  launch_k(cast_f32_bf16_kernel, grid_for(work, 256), 256, 0, st, src, ld_src, dst, ld_dst, rows, cols, vec);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__global__ void mask_with_cls_kernel(const uint8_t* in, uint8_t* out, int B, int P) {
  pdl_entry();
  const int S = P + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * S) return;
  const int b = i / S, s = i % S;
  out[i] = s == 0 ? 0 : (in[b * P + s - 1] ? 1 : 0);
}
int k_mask_with_cls(const uint8_t* in, uint8_t* out, int B, int P, cudaStream_t st) {
  launch_k(mask_with_cls_kernel, (B * (P + 1) + 255) / 256, 256, 0, st, in, out, B, P);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// ----------------------------------------------------------------------------------- FiLM
// 8 consecutive features per thread: one 16-byte load of the patch row, four 16-byte loads of gamma / beta.
__global__ void __launch_bounds__(256)
    film_apply_kernel(const bf16* __restrict__ patches, const float* __restrict__ gb, bf16* __restrict__ mod, int B,
                      int P, int Dp) {
  pdl_entry();
  const unsigned d8 = static_cast<unsigned>(Dp) >> 3;
  const unsigned rows = static_cast<unsigned>(B) * P;
  const unsigned total = rows * d8;  // < 2^32 for every supported shape (checked on the host)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned bj = i / d8, k = (i - bj * d8) * 8;
    const unsigned b = bj / static_cast<unsigned>(P);
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(patches + static_cast<int64_t>(bj) * Dp + k));
    const float* gp = gb + static_cast<int64_t>(b) * 2 * Dp + k;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp)), g1 = __ldg(reinterpret_cast<const float4*>(gp) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(gp + Dp)), b1 = __ldg(reinterpret_cast<const float4*>(gp + Dp) + 1);
    const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&xv);
    const float2 x0 = __bfloat1622float2(x[0]), x1 = __bfloat1622float2(x[1]), x2 = __bfloat1622float2(x[2]),
                 x3 = __bfloat1622float2(x[3]);
    uint4 ov;
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&ov);
    o[0] = __floats2bfloat162_rn(fmaf(g0.x, x0.x, b0.x), fmaf(g0.y, x0.y, b0.y));
    o[1] = __floats2bfloat162_rn(fmaf(g0.z, x1.x, b0.z), fmaf(g0.w, x1.y, b0.w));
    o[2] = __floats2bfloat162_rn(fmaf(g1.x, x2.x, b1.x), fmaf(g1.y, x2.y, b1.y));
    o[3] = __floats2bfloat162_rn(fmaf(g1.z, x3.x, b1.z), fmaf(g1.w, x3.y, b1.w));
    *reinterpret_cast<uint4*>(mod + static_cast<int64_t>(bj) * Dp + k) = ov;
  }
}
int k_film_apply(const bf16* patches, const float* gb, bf16* mod, int B, int P, int Dp, cudaStream_t st) {
  GG_REQUIRE(Dp % 8 == 0 && static_cast<int64_t>(B) * P * (Dp / 8) < (1LL << 32), "FiLM: unsupported shape");
  launch_k(film_apply_kernel, grid_for(static_cast<int64_t>(B) * P * Dp / 8, 256), 256, 0, st, patches, gb, mod, B, P, Dp);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// thread = (sample, pair of features): walks the P patch rows with 4-byte loads (a warp covers 128 contiguous bytes)
__global__ void __launch_bounds__(256)
    film_bwd_kernel(const bf16* __restrict__ dmod, const bf16* __restrict__ patches, const float* __restrict__ gb,
                    bf16* __restrict__ dgb, int B, int P, int Dp) {
  pdl_entry();
  const unsigned d2 = static_cast<unsigned>(Dp) >> 1;
  const unsigned total = static_cast<unsigned>(B) * d2;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned b = i / d2, k = (i - b * d2) * 2;
    const bf16* dm = dmod + static_cast<int64_t>(b) * P * Dp + k;
    const bf16* pa = patches + static_cast<int64_t>(b) * P * Dp + k;
    float dgx = 0.f, dgy = 0.f, dbx = 0.f, dby = 0.f;
#pragma unroll 4
    for (int j = 0; j < P; ++j) {
      const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dm + static_cast<int64_t>(j) * Dp));
      const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pa + static_cast<int64_t>(j) * Dp));
      dgx = fmaf(d.x, x.x, dgx);
      dgy = fmaf(d.y, x.y, dgy);
      dbx += d.x;
      dby += d.y;
    }
    const float* gp = gb + static_cast<int64_t>(b) * 2 * Dp + k;
    const float2 gamma = *reinterpret_cast<const float2*>(gp), beta = *reinterpret_cast<const float2*>(gp + Dp);
    bf16* o = dgb + static_cast<int64_t>(b) * 2 * Dp + k;
    *reinterpret_cast<__nv_bfloat162*>(o) =
        __floats2bfloat162_rn(dgx * (1.f - gamma.x * gamma.x), dgy * (1.f - gamma.y * gamma.y));
    *reinterpret_cast<__nv_bfloat162*>(o + Dp) =
        __floats2bfloat162_rn(fabsf(beta.x) < 5.0f ? dbx : 0.f, fabsf(beta.y) < 5.0f ? dby : 0.f);
  }
}
int k_film_bwd(const bf16* dmod, const bf16* patches, const float* gb, bf16* dgb, int B, int P, int Dp,
               cudaStream_t st) {
  launch_k(film_bwd_kernel, grid_for(static_cast<int64_t>(B) * Dp / 2, 256), 256, 0, st, dmod, patches, gb, dgb, B, P, Dp);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// ----------------------------------------------------------------------------- token plumbing
// 8 consecutive features (16 bytes) per thread
__global__ void __launch_bounds__(256)
    assemble_tokens_kernel(bf16* x, const float* __restrict__ cls, int R, int B, int S, int E, const bf16* __restrict__ src) {
  pdl_entry();
  const unsigned e8 = static_cast<unsigned>(E) >> 3;
  const unsigned rows_rep = static_cast<unsigned>(B) * S;
  const unsigned total = static_cast<unsigned>(R) * rows_rep * e8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned row = i / e8, e = (i - row * e8) * 8;   // row = (r * B + b) * S + s
    const unsigned rr = row % rows_rep;                      // b * S + s
    const unsigned s = rr % static_cast<unsigned>(S);
    uint4 v;
    if (s == 0) {
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(cls + e)), c1 = __ldg(reinterpret_cast<const float4*>(cls + e) + 1);
      __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&v);
      o[0] = __floats2bfloat162_rn(c0.x, c0.y); o[1] = __floats2bfloat162_rn(c0.z, c0.w);
      o[2] = __floats2bfloat162_rn(c1.x, c1.y); o[3] = __floats2bfloat162_rn(c1.z, c1.w);
    } else if (src) {
      const unsigned b = rr / static_cast<unsigned>(S);
      v = __ldg(reinterpret_cast<const uint4*>(src + (static_cast<int64_t>(b) * (S - 1) + (s - 1)) * E + e));
    } else if (row >= rows_rep) {
      v = *reinterpret_cast<const uint4*>(x + static_cast<int64_t>(rr) * E + e);  // replica 0's row (not a CLS row)
    } else {
      continue;  // replica 0: the patch projection wrote this row already
    }
    *reinterpret_cast<uint4*>(x + static_cast<int64_t>(row) * E + e) = v;
  }
}
int k_assemble_tokens(bf16* x, const float* cls, int R, int B, int S, int E, cudaStream_t st, const bf16* src) {
  GG_REQUIRE(E % 8 == 0 && static_cast<int64_t>(R) * B * S * (E / 8) < (1LL << 32), "token assembly: unsupported shape");
  launch_k(assemble_tokens_kernel, grid_for(static_cast<int64_t>(R) * B * S * E / 8, 256), 256, 0, st, x, cls, R, B, S, E, src);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__global__ void relu_bwd_kernel(const bf16* __restrict__ g, const bf16* __restrict__ h, bf16* __restrict__ out, int64_t n) {
  pdl_entry();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = __bfloat162float(h[i]) > 0.f ? g[i] : __float2bfloat16_rn(0.f);
}
int k_relu_bwd(const bf16* g, const bf16* h, bf16* out, int64_t n, cudaStream_t st) {
  launch_k(relu_bwd_kernel, grid_for(n, 256), 256, 0, st, g, h, out, n);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__global__ void __launch_bounds__(256)
    unassemble_tokens_kernel(const bf16* __restrict__ dx, bf16* __restrict__ dpe, int R, int B, int S, int E) {
  pdl_entry();
  const unsigned P = static_cast<unsigned>(S) - 1;
  const unsigned e8 = static_cast<unsigned>(E) >> 3;
  const unsigned total = static_cast<unsigned>(B) * P * e8;
  const int64_t rep = static_cast<int64_t>(B) * S * E;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned row = i / e8, e = (i - row * e8) * 8;  // row = b * P + j
    const unsigned b = row / P, j = row - b * P;
    const int64_t src = (static_cast<int64_t>(b) * S + 1 + j) * E + e;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < R; ++r) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(dx + r * rep + src));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(h[q]);
        acc[2 * q] += f.x;
        acc[2 * q + 1] += f.y;
      }
    }
    uint4 ov;
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
    *reinterpret_cast<uint4*>(dpe + static_cast<int64_t>(row) * E + e) = ov;
  }
}
int k_unassemble_tokens(const bf16* dx, bf16* dpe, float* /*unused*/, int R, int B, int S, int E,
                        cudaStream_t st) {
  if (S <= 1) return GG_OK;
  launch_k(unassemble_tokens_kernel, grid_for(static_cast<int64_t>(B) * (S - 1) * E / 8, 256), 256, 0, st, dx, dpe, R, B, S, E);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__global__ void sum_replicas_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int R, int64_t n) {
  pdl_entry();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < R; ++r) acc += __bfloat162float(in[r * n + i]);
    out[i] = __float2bfloat16_rn(acc);
  }
}
int k_sum_replicas(const bf16* in, bf16* out, int R, int64_t n, cudaStream_t st) {
  launch_k(sum_replicas_kernel, grid_for(n, 256), 256, 0, st, in, out, R, n);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// out[r * n + i] = in[r * n + i] + add[i]  (bf16 rows + an fp32 row block shared by the R replicas)
__global__ void add_bcast_replicas_kernel(const bf16* __restrict__ in, const float* __restrict__ add, bf16* __restrict__ out,
                                          int R, int64_t n) {
  pdl_entry();
  const int64_t total = static_cast<int64_t>(R) * n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(__bfloat162float(in[i]) + add[i % n]);
}
int k_add_bcast_replicas(const bf16* in, const float* add, bf16* out, int R, int64_t n, cudaStream_t st) {
  launch_k(add_bcast_replicas_kernel, grid_for(static_cast<int64_t>(R) * n, 256), 256, 0, st, in, add, out, R, n);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// Label-conditioned baseline (benchmark_generative_model.py:138-150): the conditioning vector of sample b is the
// concatenation of one row of each embedding table. Tables are the fp32 master parameters (a few KB: L2 resident).
__global__ void embed_gather_kernel(const float* __restrict__ emb0, const float* __restrict__ emb1,
                                    const int64_t* __restrict__ y0, const int64_t* __restrict__ y1, int V0, int V1,
                                    bf16* __restrict__ c, int B, int Eh) {
  pdl_entry();
  const int64_t total = static_cast<int64_t>(B) * 2 * Eh;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / (2 * Eh));
    const int col = static_cast<int>(i % (2 * Eh));
    const bool second = col >= Eh;
    int64_t y = second ? y1[b] : y0[b];
    const int V = second ? V1 : V0;
    y = y < 0 ? 0 : (y >= V ? V - 1 : y);
    const float* tab = second ? emb1 : emb0;
    c[i] = __float2bfloat16_rn(tab[y * Eh + (second ? col - Eh : col)]);
  }
}
int k_embed_gather(const float* emb0, const float* emb1, const int64_t* y0, const int64_t* y1, int V0, int V1, bf16* c,
                   int B, int Eh, cudaStream_t st) {
  launch_k(embed_gather_kernel, grid_for(static_cast<int64_t>(B) * 2 * Eh, 256), 256, 0, st, emb0, emb1, y0, y1, V0, V1,
           c, B, Eh);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// Gradient of the tables: g_t[v, :] = sum over the samples with y_t[b] = v of dc[b, t*Eh : (t+1)*Eh]. One thread per
// table element walks the batch in order (fixed summation order: deterministic, no atomics); the threads of a warp
// share v, so the label read is a broadcast and the dc read is coalesced. Every element is written (0 for unused rows).
__global__ void embed_grad_kernel(const bf16* __restrict__ dc, const int64_t* __restrict__ y0,
                                  const int64_t* __restrict__ y1, int V0, int V1, float* __restrict__ g0,
                                  float* __restrict__ g1, int B, int Eh) {
  pdl_entry();
  const int64_t total = static_cast<int64_t>(V0 + V1) * Eh;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int v = static_cast<int>(i / Eh);
    const int col = static_cast<int>(i % Eh);
    const bool second = v >= V0;
    if (second) v -= V0;
    const int V = second ? V1 : V0;
    const int64_t* y = second ? y1 : y0;
    const bf16* src = dc + (second ? Eh : 0) + col;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      int64_t yb = y[b];
      yb = yb < 0 ? 0 : (yb >= V ? V - 1 : yb);
      if (yb == v) acc += __bfloat162float(src[static_cast<int64_t>(b) * 2 * Eh]);
    }
    (second ? g1 : g0)[static_cast<int64_t>(v) * Eh + col] = acc;
  }
}
int k_embed_grad(const bf16* dc, const int64_t* y0, const int64_t* y1, int V0, int V1, float* g0, float* g1, int B,
                 int Eh, cudaStream_t st) {
  launch_k(embed_grad_kernel, grid_for(static_cast<int64_t>(V0 + V1) * Eh, 128), 128, 0, st, dc, y0, y1, V0, V1, g0, g1,
           B, Eh);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__global__ void scatter_add_rows_kernel(bf16* dst, const bf16* __restrict__ src, int B, int stride_rows, int E) {
  pdl_entry();
  const int64_t total = static_cast<int64_t>(B) * E;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / E;
    const int e = static_cast<int>(i % E);
    const int64_t o = b * stride_rows * E + e;
    dst[o] = __float2bfloat16_rn(__bfloat162float(dst[o]) + __bfloat162float(src[i]));
  }
}
int k_scatter_add_rows(bf16* dst, const bf16* src, int B, int stride_rows, int E, cudaStream_t st) {
  launch_k(scatter_add_rows_kernel, grid_for(static_cast<int64_t>(B) * E, 256), 256, 0, st, dst, src, B, stride_rows, E);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__global__ void scatter_cls_kernel(bf16* __restrict__ dst, const bf16* __restrict__ src, int B, int S, int E) {
  pdl_entry();
  const int64_t total = static_cast<int64_t>(B) * S * E;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % E);
    const int s = static_cast<int>((i / E) % S);
    const int64_t b = i / (static_cast<int64_t>(E) * S);
    dst[i] = s == 0 ? src[b * E + e] : __float2bfloat16_rn(0.f);
  }
}
int k_scatter_cls(bf16* dst, const bf16* src, int B, int S, int E, cudaStream_t st) {
  launch_k(scatter_cls_kernel, grid_for(static_cast<int64_t>(B) * S * E, 256), 256, 0, st, dst, src, B, S, E);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// ------------------------------------------------------------------ deterministic column sums
// stage 1: block (32 columns x 8 row lanes) reduces a row chunk; stage 2 sums the chunk partials in order.
__global__ void colsum_stage1_kernel(const void* __restrict__ in, int in_f32, int64_t ld, int64_t rows, int N,
                                     const float* __restrict__ roww, int64_t rows_per_chunk,
                                     float* __restrict__ partial) {
  pdl_entry();
  __shared__ float sm[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk;
  int64_t r1 = r0 + rows_per_chunk;
  if (r1 > rows) r1 = rows;
  float acc = 0.f;
  if (n < N) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float x = in_f32 ? reinterpret_cast<const float*>(in)[r * ld + n]
                             : __bfloat162float(reinterpret_cast<const bf16*>(in)[r * ld + n]);
      acc = roww ? fmaf(roww[r], x, acc) : acc + x;
    }
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += sm[y][threadIdx.x];
    partial[static_cast<int64_t>(blockIdx.y) * N + n] = t;
  }
}
__global__ void colsum_stage2_kernel(const float* __restrict__ partial, int nchunks, int N, float scale,
                                     float* __restrict__ out, int accumulate) {
  pdl_entry();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float t = 0.f;
  for (int c = 0; c < nchunks; ++c) t += partial[static_cast<int64_t>(c) * N + n];
  t *= scale;
  out[n] = accumulate ? out[n] + t : t;
}
static inline int colsum_chunks(int64_t rows) {
  int64_t c = (rows + 255) / 256;
  if (c > 64) c = 64;
  if (c < 1) c = 1;
  return static_cast<int>(c);
}
int k_colsum(const void* in, int in_f32, int64_t ld, int64_t rows, int N, const float* roww, float scale,
             float* out, int accumulate, float* scratch, cudaStream_t st) {
  GG_REQUIRE(scratch != nullptr, "k_colsum needs scratch");
  const int nchunks = colsum_chunks(rows);
  const int64_t rpc = (rows + nchunks - 1) / nchunks;
  dim3 grid((N + 31) / 32, nchunks), block(32, 8);
  launch_k(colsum_stage1_kernel, grid, block, 0, st, in, in_f32, ld, rows, N, roww, rpc, scratch);
  GG_LAUNCH_CHECK();
  launch_k(colsum_stage2_kernel, (N + 255) / 256, 256, 0, st, scratch, nchunks, N, scale, out, accumulate);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// ------------------------------------------------------------------ grouped column sums
// All bias gradients of one backward pass in one launch: problem p sums the rows of a bf16 matrix
// [rows_p, N_p] (pitch ld_p) into out_p[N_p] (fp32). A work item is (problem, 64-column group, row chunk);
// the chunk partials of a column group are summed in chunk order by the last CTA to arrive (fixed order =>
// bit-reproducible, no float atomics). Thread = 8 consecutive columns (one 16-byte load per row), 32 row
// lanes per CTA.
struct ColsumGroupParams {
  int nprob;
  int total_work;
  struct {
    const bf16* in;
    int64_t ld;
    int rows, N;
    int work0, ngroups, nchunks, rows_per_chunk;
    float* out;
    float* partial;      // [nchunks][N]
    unsigned* counters;  // [ngroups]
  } p[COLSUM_GROUP_MAX];
};

__global__ void __launch_bounds__(256) colsum_group_kernel(const __grid_constant__ ColsumGroupParams P) {
  pdl_entry();
  __shared__ float sm[32][65];
  __shared__ unsigned last_flag;
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  for (int w = blockIdx.x; w < P.total_work; w += gridDim.x) {
    int pi = 0;
    while (pi + 1 < P.nprob && w >= P.p[pi + 1].work0) ++pi;
    const auto& p = P.p[pi];
    const int local = w - p.work0;
    const int chunk = local / p.ngroups, grp = local % p.ngroups;
    const int n0 = grp * 64 + cg * 8;
    const int r0 = chunk * p.rows_per_chunk;
    const int r1 = min(p.rows, r0 + p.rows_per_chunk);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const bool vec = (p.ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0) && n0 + 8 <= p.N;
    if (vec) {
      const bf16* base = p.in + n0;
      int r = r0 + rl;
      for (; r + 96 < r1; r += 128) {  // four independent 16-byte loads in flight per thread
        uint4 u[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) u[t] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<int64_t>(r + 32 * t) * p.ld));
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[t]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __bfloat1622float2(h[q]);
            acc[2 * q] += f.x;
            acc[2 * q + 1] += f.y;
          }
        }
      }
      for (; r < r1; r += 32) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + static_cast<int64_t>(r) * p.ld));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(h[q]);
          acc[2 * q] += f.x;
          acc[2 * q + 1] += f.y;
        }
      }
    } else {
      for (int r = r0 + rl; r < r1; r += 32)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n0 + j < p.N) acc[j] += __bfloat162float(p.in[static_cast<int64_t>(r) * p.ld + n0 + j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[rl][cg * 8 + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 64) {
      float t = 0.f;
#pragma unroll
      for (int y = 0; y < 32; ++y) t += sm[y][threadIdx.x];
      const int n = grp * 64 + threadIdx.x;
      if (n < p.N) {
        if (p.nchunks == 1) p.out[n] = t;
        else __stcg(p.partial + static_cast<int64_t>(chunk) * p.N + n, t);
      }
    }
    if (p.nchunks > 1) {
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(p.counters + grp, 1u);
        const bool last = prev == static_cast<unsigned>(p.nchunks - 1);
        if (last) p.counters[grp] = 0;
        last_flag = last ? 1u : 0u;
        __threadfence();
      }
      __syncthreads();
      if (last_flag && threadIdx.x < 64) {
        const int n = grp * 64 + threadIdx.x;
        if (n < p.N) {
          float t = 0.f;
          for (int c = 0; c < p.nchunks; ++c) t += __ldcg(p.partial + static_cast<int64_t>(c) * p.N + n);
          p.out[n] = t;
        }
      }
    }
    __syncthreads();  // sm / last_flag are reused by the next work item
  }
}

int64_t colsum_group_workspace_bytes(int64_t max_total_columns) {
  return GROUP_COUNTER_BYTES + static_cast<int64_t>(COLSUM_GROUP_MAX_CHUNKS) * max_total_columns * 4 +
         256LL * COLSUM_GROUP_MAX;
}

// The first GROUP_COUNTER_BYTES of `workspace` must be zero-initialised once (arrival counters).
int k_colsum_group(const ColsumItem* items, int n, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (n <= 0) return GG_OK;
  GG_REQUIRE(n <= COLSUM_GROUP_MAX, "too many grouped column sums (%d > %d)", n, COLSUM_GROUP_MAX);
  static thread_local ColsumGroupParams P;
  P.nprob = n;
  int work = 0;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int64_t off = GROUP_COUNTER_BYTES, ctr_off = 0;
  for (int i = 0; i < n; ++i) {
    const ColsumItem& it = items[i];
    GG_REQUIRE(it.in && it.out && it.rows > 0 && it.N > 0, "bad grouped column-sum item %d", i);
    auto& p = P.p[i];
    p.in = reinterpret_cast<const bf16*>(it.in); p.ld = it.ld; p.rows = static_cast<int>(it.rows); p.N = it.N; p.out = it.out;
    p.ngroups = ceil_div(it.N, 64);
    int nch = static_cast<int>((it.rows + 2047) / 2048);
    if (nch > COLSUM_GROUP_MAX_CHUNKS) nch = COLSUM_GROUP_MAX_CHUNKS;
    if (nch < 1) nch = 1;
    p.nchunks = nch;
    p.rows_per_chunk = static_cast<int>((it.rows + nch - 1) / nch);
    p.work0 = work;
    work += p.ngroups * nch;
    p.partial = reinterpret_cast<float*>(ws + off);
    if (nch > 1) off += round_up64(static_cast<int64_t>(nch) * it.N * 4, 256);
    p.counters = reinterpret_cast<unsigned*>(ws + ctr_off);
    ctr_off += static_cast<int64_t>(p.ngroups) * 4;
    GG_REQUIRE(off <= workspace_bytes && ctr_off <= GROUP_COUNTER_BYTES, "grouped column-sum workspace too small");
  }
  P.total_work = work;
  const unsigned grid = static_cast<unsigned>(work < 148 * 8 ? work : 148 * 8);
  launch_k(colsum_group_kernel, grid, 256, 0, st, P);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// ---------------------------------------------------------------------- trunk / GP glue
__global__ void trunk1_combine_kernel(const float* __restrict__ a1x, const float* __restrict__ a1c,
                                      const float* __restrict__ b1, const float* __restrict__ alpha,
                                      bf16* __restrict__ h1, int B, int H, int npass, int R, float slope) {
  pdl_entry();
  const int64_t total = static_cast<int64_t>(npass) * B * H;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int h = static_cast<int>(i % H);
    const int64_t m = i / H;
    const int pass = static_cast<int>(m / B);
    const int b = static_cast<int>(m % B);
    float v;
    if (pass < 2) {
      v = a1x[(static_cast<int64_t>(pass) * B + b) * H + h];
    } else {
      const float al = alpha[b];
      // same association as the reference: alpha*real + (1-alpha)*fake (:358), applied after W1x (linear)
      v = al * a1x[(static_cast<int64_t>(B) + b) * H + h] + (1.f - al) * a1x[static_cast<int64_t>(b) * H + h];
    }
    if (a1c) v += a1c[(static_cast<int64_t>(R > 1 ? pass : 0) * B + b) * H + h];
    else if (b1) v += b1[h];
    v = v > 0.f ? v : slope * v;
    h1[i] = __float2bfloat16_rn(v);
  }
}
int k_trunk1_combine(const float* a1x, const float* a1c, const float* b1, const float* alpha, bf16* h1, int B,
                     int H, int npass, int R, float slope, cudaStream_t st) {
  GG_REQUIRE(npass == 1 || npass == 3, "npass must be 1 or 3");
  launch_k(trunk1_combine_kernel, grid_for(static_cast<int64_t>(npass) * B * H, 256), 256, 0, st, a1x, a1c, b1, alpha, h1, B, H, npass, R, slope);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void rowdot_bias_kernel(const float* __restrict__ h2f, const float* __restrict__ w3,
                                   const float* __restrict__ b3, float* __restrict__ score, int rows, int H) {
  pdl_entry();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc = 0.f;
  for (int h = lane; h < H; h += 32) acc = fmaf(h2f[static_cast<int64_t>(row) * H + h], w3[h], acc);
  acc = warp_sum(acc);
  if (lane == 0) score[row] = acc + (b3 ? b3[0] : 0.f);
}
int k_rowdot_bias(const float* h2f, const float* w3, const float* b3, float* score, int rows, int H,
                  cudaStream_t st) {
  launch_k(rowdot_bias_kernel, (rows + 7) / 8, 256, 0, st, h2f, w3, b3, score, rows, H);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__global__ void gp_u2_kernel(const bf16* __restrict__ h2i, const float* __restrict__ w3, bf16* __restrict__ u2,
                             int B, int H, float slope) {
  pdl_entry();
  const int64_t total = static_cast<int64_t>(B) * H;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i % H);
    const float m = __bfloat162float(h2i[i]) > 0.f ? 1.f : slope;
    u2[i] = __float2bfloat16_rn(m * w3[o]);
  }
}
int k_gp_u2(const bf16* h2i, const float* w3, bf16* u2, int B, int H, float slope, cudaStream_t st) {
  launch_k(gp_u2_kernel, grid_for(static_cast<int64_t>(B) * H, 256), 256, 0, st, h2i, w3, u2, B, H, slope);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// One warp per interpolated sample: ||g_b||^2 = u1_b . (u1_b M) with y = u1 M precomputed.
__global__ void gp_rows_kernel(const float* __restrict__ y, const float* __restrict__ u1f,
                               const bf16* __restrict__ h1i, float* __restrict__ norms, float* __restrict__ pen,
                               bf16* __restrict__ ru1, bf16* __restrict__ dv1, int B, int H, float slope,
                               float gp_weight, float inv_batch) {
  pdl_entry();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const int64_t o = static_cast<int64_t>(row) * H;
  float acc = 0.f;
  for (int h = lane; h < H; h += 32) acc = fmaf(y[o + h], u1f[o + h], acc);
  acc = warp_sum(acc);
  const float n = sqrtf(fmaxf(acc, 0.f));
  // d/du1 of gp_weight * mean_b (n_b - 1)^2  =  r_b * (u1_b M),  r_b = gp_weight * (2/B) * (1 - 1/n_b)
  const float r = n > 0.f ? gp_weight * 2.f * inv_batch * (1.f - 1.f / n) : 0.f;
  if (lane == 0) {
    norms[row] = n;
    pen[row] = (n - 1.f) * (n - 1.f);
  }
  for (int h = lane; h < H; h += 32) {
    ru1[o + h] = __float2bfloat16_rn(r * u1f[o + h]);
    const float m1 = __bfloat162float(h1i[o + h]) > 0.f ? 1.f : slope;
    dv1[o + h] = __float2bfloat16_rn(m1 * r * y[o + h]);
  }
}
int k_gp_rows(const float* y, const float* u1f, const bf16* h1i, float* norms, float* pen, bf16* ru1, bf16* dv1,
              int B, int H, float slope, float gp_weight, float inv_batch, cudaStream_t st) {
  launch_k(gp_rows_kernel, (B + 7) / 8, 256, 0, st, y, u1f, h1i, norms, pen, ru1, dv1, B, H, slope, gp_weight,
                                             inv_batch);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__global__ void score_bwd_kernel(const bf16* __restrict__ h2, const float* __restrict__ w3, bf16* __restrict__ da2,
                                 float* __restrict__ roww, int rows, int B, int H, float slope, float s0, float s1,
                                 float inv_batch) {
  pdl_entry();
  const int64_t total = static_cast<int64_t>(rows) * H;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i % H);
    const int64_t m = i / H;
    const float d = (m < B ? s0 : s1) * inv_batch;
    const float mk = __bfloat162float(h2[i]) > 0.f ? 1.f : slope;
    da2[i] = __float2bfloat16_rn(d * w3[o] * mk);
    if (o == 0 && roww) roww[m] = d;
  }
}
int k_score_bwd(const bf16* h2, const float* w3, bf16* da2, float* roww, int rows, int B, int H, float slope,
                float sign_first, float sign_second, float inv_batch, cudaStream_t st) {
  launch_k(score_bwd_kernel, grid_for(static_cast<int64_t>(rows) * H, 256), 256, 0, st, h2, w3, da2, roww, rows, B, H, slope, sign_first, sign_second, inv_batch);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// da2[m, o] = drow[m] * w3[o] * LeakyReLU'(h2[m, o]): the critic's last layer backward for an arbitrary upstream
// gradient of the scores (module-level autograd: discriminator(x, ...).backward(dscore))
__global__ void score_bwd_rows_kernel(const bf16* __restrict__ h2, const float* __restrict__ w3,
                                      const float* __restrict__ drow, bf16* __restrict__ da2, int rows, int H, float slope) {
  pdl_entry();
  const int64_t total = static_cast<int64_t>(rows) * H;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i % H);
    const float mk = __bfloat162float(h2[i]) > 0.f ? 1.f : slope;
    da2[i] = __float2bfloat16_rn(drow[i / H] * w3[o] * mk);
  }
}
int k_score_bwd_rows(const bf16* h2, const float* w3, const float* drow, bf16* da2, int rows, int H, float slope,
                     cudaStream_t st) {
  launch_k(score_bwd_rows_kernel, grid_for(static_cast<int64_t>(rows) * H, 256), 256, 0, st, h2, w3, drow, da2, rows, H, slope);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__device__ float block_sum_ordered(float v, float* sm) {
  // fixed-order tree: deterministic for a fixed block size
  const int t = threadIdx.x;
  sm[t] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (t < s) sm[t] += sm[t + s];
    __syncthreads();
  }
  const float r = sm[0];
  __syncthreads();
  return r;
}

__global__ void disc_losses_kernel(const float* __restrict__ score, const float* __restrict__ pen,
                                   float* __restrict__ stats, int B, float gp_weight, float inv_batch) {
  pdl_entry();
  __shared__ float sm[256];
  float f = 0.f, r = 0.f, p = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    f += score[b];
    r += score[B + b];
    p += pen[b];
  }
  f = block_sum_ordered(f, sm);
  r = block_sum_ordered(r, sm);
  p = block_sum_ordered(p, sm);
  if (threadIdx.x == 0) {
    const float loss_real = -r * inv_batch, loss_fake = f * inv_batch, gp = p * inv_batch;
    stats[0] = loss_real;
    stats[1] = loss_fake;
    stats[2] = gp;
    stats[3] = loss_real + loss_fake + gp_weight * gp;
  }
}
int k_disc_losses(const float* score, const float* pen, float* stats, int B, float gp_weight, float inv_batch,
                  cudaStream_t st) {
  launch_k(disc_losses_kernel, 1, 256, 0, st, score, pen, stats, B, gp_weight, inv_batch);
  GG_LAUNCH_CHECK();
  return GG_OK;
}
__global__ void gen_loss_kernel(const float* __restrict__ score, float* __restrict__ stats, int B, float inv_batch) {
  pdl_entry();
  __shared__ float sm[256];
  float f = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) f += score[b];
  f = block_sum_ordered(f, sm);
  if (threadIdx.x == 0) stats[4] = -f * inv_batch;
}
int k_gen_loss(const float* score, float* stats, int B, float inv_batch, cudaStream_t st) {
  launch_k(gen_loss_kernel, 1, 256, 0, st, score, stats, B, inv_batch);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__global__ void fill_f32_kernel(float* p, float v, int64_t n) {
  pdl_entry();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}
int k_fill_f32(float* p, float v, int64_t n, cudaStream_t st) {
  if (n <= 0) return GG_OK;
  launch_k(fill_f32_kernel, grid_for(n, 256), 256, 0, st, p, v, n);
  GG_LAUNCH_CHECK();
  return GG_OK;
}
__global__ void bump_rng_kernel(uint64_t* rng) {
  pdl_entry(); rng[1] += 1; }
int k_bump_rng(uint64_t* rng, cudaStream_t st) {
  launch_k(bump_rng_kernel, 1, 1, 0, st, rng);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

__global__ void __launch_bounds__(256)
    masked_mean_rows_kernel(const float* __restrict__ x, const uint8_t* __restrict__ pad, float* __restrict__ out,
                            int P, int D) {
  pdl_entry();
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float* xb = x + static_cast<int64_t>(b) * P * D + d;
  float acc = 0.f;
  int cnt = 0;
  for (int p = 0; p < P; ++p) {
    if (pad && pad[static_cast<int64_t>(b) * P + p]) continue;
    acc += xb[static_cast<int64_t>(p) * D];
    ++cnt;
  }
  out[static_cast<int64_t>(b) * D + d] = acc / static_cast<float>(cnt);  // cnt = 0 -> inf/nan, as in the reference
}
int k_masked_mean_rows(const float* x, const uint8_t* pad, float* out, int B, int P, int D, cudaStream_t st) {
  dim3 grid(static_cast<unsigned>((D + 255) / 256), static_cast<unsigned>(B));
  launch_k(masked_mean_rows_kernel, grid, 256, 0, st, x, pad, out, P, D);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// dst[r, 0:cols] = index[r] >= 0 ? src[index[r], 0:cols] : 0 -- batch assembly on the device: the
// rows of one batch (patch embeddings picked / zero-padded per case, gene profiles, token embeddings) are gathered
// from a dataset that stays resident in HBM instead of being np.load-ed, collated and copied per step
// (src/multi_patch_multi_token_gan_dataloader.py:25-55). One warp per row, 16-byte accesses when aligned. HBM bound.
__global__ void __launch_bounds__(256)
    gather_rows_kernel(const float* __restrict__ src, int64_t ld_src, const int64_t* __restrict__ index,
                       float* __restrict__ dst, int64_t ld_dst, int64_t rows, int cols) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const bool vec = (cols % 4 == 0) && (ld_src % 4 == 0) && (ld_dst % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); r < rows;
       r += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t s = index[r];
    float* d = dst + r * ld_dst;
    if (vec) {
      const float4* sp = s >= 0 ? reinterpret_cast<const float4*>(src + s * ld_src) : nullptr;
      for (int c = lane; c < cols / 4; c += 32)
        reinterpret_cast<float4*>(d)[c] = sp ? __ldg(sp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int c = lane; c < cols; c += 32) d[c] = s >= 0 ? src[s * ld_src + c] : 0.f;
    }
  }
}
int k_gather_rows(const float* src, int64_t ld_src, const int64_t* index, float* dst, int64_t ld_dst, int64_t rows,
                  int cols, cudaStream_t st) {
  if (rows <= 0) return GG_OK;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(gather_rows_kernel, static_cast<unsigned>(blocks), 256, 0, st, src, ld_src, index, dst, ld_dst, rows, cols);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// ---------------------------------------------------------------------------------------- BatchNorm1d over the batch
// conditional_gan_attention.py:108, :126 (generator only): y = (x - mean_B) / sqrt(var_B + eps) * gamma + beta over the
// rows of x [B, E] in training mode (biased batch variance; the running statistics take the UNBIASED one, torch
// nn.BatchNorm1d), the running statistics in eval mode. One block per 32 columns, the 8 warps stride over the rows and
// meet in shared memory in a fixed order (deterministic); x is a few hundred KB and stays in L2 between the passes.
__device__ __forceinline__ float bn_block_sum(float v, float (*red)[33]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  red[w][lane] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][lane];
  return s;
}

__global__ void __launch_bounds__(256)
    bn_fwd_kernel(const bf16* __restrict__ x, int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float* __restrict__ run_mean, float* __restrict__ run_var, float momentum, float eps, int training,
                  bf16* __restrict__ y, int64_t ldy, float* __restrict__ mean_out, float* __restrict__ rstd_out, int B,
                  int E) {
  pdl_entry();
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const bool ok = col < E;
  float mean, rstd;
  if (training) {
    float s = 0.f;
    if (ok)
      for (int r = w; r < B; r += 8) s += __bfloat162float(x[r * ldx + col]);
    mean = bn_block_sum(s, red) / static_cast<float>(B);
    float ss = 0.f;
    if (ok)
      for (int r = w; r < B; r += 8) {
        const float d = __bfloat162float(x[r * ldx + col]) - mean;
        ss += d * d;
      }
    ss = bn_block_sum(ss, red);
    rstd = rsqrtf(ss / static_cast<float>(B) + eps);
    if (ok && w == 0 && run_mean) {
      run_mean[col] = (1.f - momentum) * run_mean[col] + momentum * mean;
      run_var[col] = (1.f - momentum) * run_var[col] + momentum * ss / static_cast<float>(B - 1);
    }
  } else {
    mean = ok ? run_mean[col] : 0.f;
    rstd = ok ? rsqrtf(run_var[col] + eps) : 0.f;
  }
  if (!ok) return;
  if (w == 0) {
    mean_out[col] = mean;
    rstd_out[col] = rstd;
  }
  const float g = gamma[col] * rstd, b = beta[col];
  for (int r = w; r < B; r += 8) y[r * ldy + col] = __float2bfloat16((__bfloat162float(x[r * ldx + col]) - mean) * g + b);
}

// dx = gamma * rstd * (dy - mean_B(dy) - xhat * mean_B(dy * xhat)), dgamma = sum dy * xhat, dbeta = sum dy (training mode)
__global__ void __launch_bounds__(256)
    bn_bwd_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x, int64_t ldx,
                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                  bf16* __restrict__ dx, int64_t lddx, float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int E) {
  pdl_entry();
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const bool ok = col < E;
  const float mu = ok ? mean[col] : 0.f, rs = ok ? rstd[col] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  if (ok)
    for (int r = w; r < B; r += 8) {
      const float d = __bfloat162float(dy[r * lddy + col]);
      s1 += d;
      s2 += d * (__bfloat162float(x[r * ldx + col]) - mu) * rs;
    }
  s1 = bn_block_sum(s1, red);
  s2 = bn_block_sum(s2, red);
  if (!ok) return;
  if (w == 0) {
    dgamma[col] = s2;
    dbeta[col] = s1;
  }
  const float inv_b = 1.f / static_cast<float>(B), g = gamma[col] * rs;
  for (int r = w; r < B; r += 8) {
    const float xh = (__bfloat162float(x[r * ldx + col]) - mu) * rs;
    dx[r * lddx + col] = __float2bfloat16(g * (__bfloat162float(dy[r * lddy + col]) - s1 * inv_b - xh * s2 * inv_b));
  }
}

int k_bn_fwd(const bf16* x, int64_t ldx, const float* gamma, const float* beta, float* run_mean, float* run_var,
             float momentum, float eps, int training, bf16* y, int64_t ldy, float* mean_out, float* rstd_out, int B, int E,
             cudaStream_t st) {
  GG_REQUIRE(!training || B > 1, "BatchNorm1d in training mode needs more than one row per channel (torch raises too)");
  GG_REQUIRE(training || (run_mean && run_var), "eval-mode BatchNorm needs the running statistics");
  launch_k(bn_fwd_kernel, static_cast<unsigned>((E + 31) / 32), 256, 0, st, x, ldx, gamma, beta, run_mean, run_var, momentum,
           eps, training, y, ldy, mean_out, rstd_out, B, E);
  GG_LAUNCH_CHECK();
  return GG_OK;
}
int k_bn_bwd(const bf16* dy, int64_t lddy, const bf16* x, int64_t ldx, const float* mean, const float* rstd,
             const float* gamma, bf16* dx, int64_t lddx, float* dgamma, float* dbeta, int B, int E, cudaStream_t st) {
  launch_k(bn_bwd_kernel, static_cast<unsigned>((E + 31) / 32), 256, 0, st, dy, lddy, x, ldx, mean, rstd, gamma, dx, lddx,
           dgamma, dbeta, B, E);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg
