// Residual-add + dropout + LayerNorm, forward and backward, one warp per token row (fp32 statistics).
// HBM-bound: forward reads x,y and writes z,out (8 B/element in bf16); backward reads dout,z and
// writes dz(,dy). Weight/bias gradients are reduced per block and finished in a fixed order.
//
// Replaces the `x = norm(x + dropout(sublayer(x)))` halves of nn.TransformerEncoderLayer's
// post-norm branch used at src/conditional_gan_cross_attention_with_film.py:114-119,144
// (torch/nn/modules/transformer.py, norm_first=False).
#include "host_util.h"
#include "kernels.h"
#include "philox.cuh"

namespace gg {

constexpr int LN_MAX_PER_LANE = 32;  // E <= 1024

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
    add_ln_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, const float* __restrict__ w,
                      const float* __restrict__ b, bf16* __restrict__ z, bf16* __restrict__ out,
                      float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int E, float eps,
                      float drop_p, const uint64_t* __restrict__ rng, uint32_t site) {
  const int lane = threadIdx.x & 31;
  const int per = E >> 5;
  const float inv_e = 1.f / static_cast<float>(E);
  uint64_t seed = 0, step = 0;
  if (drop_p > 0.f) {
    seed = rng[0];
    step = rng[1];
  }
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * 8) {
    float v[LN_MAX_PER_LANE];
    float s = 0.f;
    const int64_t base = row * E;
#pragma unroll
    for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
      if (k < per) {
        const int e = k * 32 + lane;
        float yv = y ? __bfloat162float(y[base + e]) : 0.f;
        if (drop_p > 0.f) yv = dropout_keep(seed, step, site, static_cast<uint64_t>(base + e), drop_p) ? yv * keep_scale : 0.f;
        v[k] = __bfloat162float(x[base + e]) + yv;
        s += v[k];
      }
    }
    const float mu = wsum(s) * inv_e;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_PER_LANE; ++k)
      if (k < per) q += (v[k] - mu) * (v[k] - mu);
    const float rs = rsqrtf(wsum(q) * inv_e + eps);
#pragma unroll
    for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
      if (k < per) {
        const int e = k * 32 + lane;
        if (z) z[base + e] = __float2bfloat16_rn(v[k]);
        const float o = (v[k] - mu) * rs * w[e] + (b ? b[e] : 0.f);
        out[base + e] = __float2bfloat16_rn(o);
      }
    }
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

int k_add_ln_fwd(const bf16* x, const bf16* y, const float* w, const float* b, bf16* z, bf16* out, float* mean,
                 float* rstd, int64_t rows, int E, float eps, float drop_p, const uint64_t* rng, uint32_t site,
                 cudaStream_t st) {
  GG_REQUIRE(E % 32 == 0 && E <= 32 * LN_MAX_PER_LANE, "LayerNorm width %d unsupported", E);
  GG_REQUIRE(drop_p == 0.f || rng, "dropout needs rng state");
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  add_ln_fwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, y, w, b, z, out, mean, rstd, rows, E, eps,
                                                                  drop_p, rng, site);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

static inline int ln_bwd_blocks(int64_t rows) {
  int64_t blocks = (rows + 31) / 32;
  if (blocks > 296) blocks = 296;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}
int64_t ln_bwd_scratch_floats(int64_t rows, int E) { return static_cast<int64_t>(ln_bwd_blocks(rows)) * 2 * E; }

__global__ void __launch_bounds__(256)
    add_ln_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ z, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ w, bf16* __restrict__ dz,
                      bf16* __restrict__ dy, float* __restrict__ partial, int64_t rows, int E, float drop_p,
                      const uint64_t* __restrict__ rng, uint32_t site) {
  extern __shared__ float sm[];  // [8 warps][2E]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = E >> 5;
  const float inv_e = 1.f / static_cast<float>(E);
  uint64_t seed = 0, step = 0;
  if (drop_p > 0.f) {
    seed = rng[0];
    step = rng[1];
  }
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  float aw[LN_MAX_PER_LANE], ab[LN_MAX_PER_LANE];
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) aw[k] = ab[k] = 0.f;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp; row < rows;
       row += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t base = row * E;
    const float mu = mean[row], rs = rstd[row];
    float g[LN_MAX_PER_LANE], xh[LN_MAX_PER_LANE];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
      if (k < per) {
        const int e = k * 32 + lane;
        const float d = __bfloat162float(dout[base + e]);
        xh[k] = (__bfloat162float(z[base + e]) - mu) * rs;
        g[k] = d * w[e];
        c1 += g[k];
        c2 += g[k] * xh[k];
        aw[k] += d * xh[k];
        ab[k] += d;
      }
    }
    c1 = wsum(c1) * inv_e;
    c2 = wsum(c2) * inv_e;
#pragma unroll
    for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
      if (k < per) {
        const int e = k * 32 + lane;
        const float v = rs * (g[k] - c1 - xh[k] * c2);
        dz[base + e] = __float2bfloat16_rn(v);
        if (dy) {
          const bool keep = drop_p > 0.f ? dropout_keep(seed, step, site, static_cast<uint64_t>(base + e), drop_p) : true;
          dy[base + e] = __float2bfloat16_rn(keep ? v * keep_scale : 0.f);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
    if (k < per) {
      sm[warp * 2 * E + k * 32 + lane] = aw[k];
      sm[warp * 2 * E + E + k * 32 + lane] = ab[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * E; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += sm[wv * 2 * E + i];
    partial[static_cast<int64_t>(blockIdx.x) * 2 * E + i] = t;
  }
}
__global__ void ln_bwd_finish_kernel(const float* __restrict__ partial, int nblocks, int E, float* __restrict__ dw,
                                     float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * E) return;
  float t = 0.f;
  for (int c = 0; c < nblocks; ++c) t += partial[static_cast<int64_t>(c) * 2 * E + i];
  if (i < E) dw[i] = t;
  else if (db) db[i - E] = t;
}

int k_add_ln_bwd(const bf16* dout, const bf16* z, const float* mean, const float* rstd, const float* w, bf16* dz,
                 bf16* dy, float* dw, float* db, int64_t rows, int E, float drop_p, const uint64_t* rng,
                 uint32_t site, float* scratch, cudaStream_t st) {
  GG_REQUIRE(E % 32 == 0 && E <= 32 * LN_MAX_PER_LANE, "LayerNorm width %d unsupported", E);
  const int blocks = ln_bwd_blocks(rows);
  const size_t smem = static_cast<size_t>(8) * 2 * E * sizeof(float);
  if (smem > 48 * 1024) {
    static bool set = false;
    if (!set) {
      GG_CUDA_CHECK(cudaFuncSetAttribute(add_ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      set = true;
    }
  }
  add_ln_bwd_kernel<<<blocks, 256, smem, st>>>(dout, z, mean, rstd, w, dz, dy, scratch, rows, E, drop_p, rng, site);
  GG_LAUNCH_CHECK();
  ln_bwd_finish_kernel<<<(2 * E + 255) / 256, 256, 0, st>>>(scratch, blocks, E, dw, db);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg
