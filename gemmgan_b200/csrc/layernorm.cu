// Residual-add + dropout + LayerNorm, forward and backward, one warp per token row (fp32 statistics).
// HBM-bound: forward reads x,y and writes z,out (8 B/element in bf16); backward reads dout,z and
// writes dz(,dy). Rows whose width is a multiple of 256 use 16-byte accesses (each lane owns runs of 8
// consecutive elements, a warp covers 512 contiguous bytes per access); other widths (multiples of 32,
// <= 256: the reduced-width test models) use a scalar path. Weight/bias gradients are reduced per block
// and finished in a fixed order (deterministic).
//
// Replaces the `x = norm(x + dropout(sublayer(x)))` halves of nn.TransformerEncoderLayer's
// post-norm branch used at src/conditional_gan_cross_attention_with_film.py:114-119,144
// (torch/nn/modules/transformer.py, norm_first=False).
#include "host_util.h"
#include "pdl.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace gg {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Element ownership: VEC: lane owns elements (k*32 + lane)*8 + t, t < 8, k < PER  (E = 256*PER)
//                    !VEC: lane owns elements k*32 + lane, k < per <= PER          (E = 32*per)
template <int PER, bool VEC>
struct RowIO {
  static constexpr int N = VEC ? PER * 8 : PER;
  __device__ static __forceinline__ int elem(int k, int t, int lane) {
    return VEC ? (k * 32 + lane) * 8 + t : k * 32 + lane;
  }
  __device__ static __forceinline__ void load_bf16(const bf16* p, int lane, int per, float* v) {
    if (VEC) {
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const uint4 u = *reinterpret_cast<const uint4*>(p + (k * 32 + lane) * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(h[q]);
          v[k * 8 + 2 * q] = f.x;
          v[k * 8 + 2 * q + 1] = f.y;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < PER; ++k) v[k] = k < per ? __bfloat162float(p[k * 32 + lane]) : 0.f;
    }
  }
  __device__ static __forceinline__ void store_bf16(bf16* p, int lane, int per, const float* v) {
    if (VEC) {
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(v[k * 8 + 2 * q], v[k * 8 + 2 * q + 1]);
        *reinterpret_cast<uint4*>(p + (k * 32 + lane) * 8) = u;
      }
    } else {
#pragma unroll
      for (int k = 0; k < PER; ++k)
        if (k < per) p[k * 32 + lane] = __float2bfloat16_rn(v[k]);
    }
  }
  __device__ static __forceinline__ void load_f32(const float* p, int lane, int per, float* v) {
    if (VEC) {
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p + (k * 32 + lane) * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p + (k * 32 + lane) * 8) + 1);
        v[k * 8 + 0] = a.x; v[k * 8 + 1] = a.y; v[k * 8 + 2] = a.z; v[k * 8 + 3] = a.w;
        v[k * 8 + 4] = b.x; v[k * 8 + 5] = b.y; v[k * 8 + 6] = b.z; v[k * 8 + 7] = b.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < PER; ++k) v[k] = k < per ? __ldg(p + k * 32 + lane) : 0.f;
    }
  }
  // keep[i] for the elements this lane owns (dropout site stream, element index base + elem)
  __device__ static __forceinline__ void keep_mask(uint64_t seed, uint64_t step, uint32_t site, uint64_t base,
                                                   float p, int lane, int per, bool* keep) {
    if (VEC) {
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const uint64_t e0 = base + static_cast<uint64_t>((k * 32 + lane) * 8);  // multiple of 8
        const uint32_t kb = keep_bits8(dropout_words(seed, step, site, e0 >> 3), dropout_thr(p));
#pragma unroll
        for (int t = 0; t < 8; ++t) keep[k * 8 + t] = ((kb >> t) & 1u) != 0;
      }
    } else {
#pragma unroll
      for (int k = 0; k < PER; ++k)
        keep[k] = k < per ? dropout_keep(seed, step, site, base + static_cast<uint64_t>(k * 32 + lane), p) : false;
    }
  }
};

template <int PER, bool VEC>
__global__ void __launch_bounds__(256)
    add_ln_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, const float* __restrict__ w,
                      const float* __restrict__ b, bf16* __restrict__ z, bf16* __restrict__ out,
                      float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int E, float eps,
                      float drop_p, const uint64_t* __restrict__ rng, uint32_t site) {
  pdl_entry();
  using IO = RowIO<PER, VEC>;
  constexpr int N = IO::N;
  const int lane = threadIdx.x & 31;
  const int per = E >> 5;
  const float inv_e = 1.f / static_cast<float>(E);
  uint64_t seed = 0, step = 0;
  if (drop_p > 0.f) {
    seed = rng[0];
    step = rng[1];
  }
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  float wv[N], bv[N];
  IO::load_f32(w, lane, per, wv);
  if (b) IO::load_f32(b, lane, per, bv);
  else {
#pragma unroll
    for (int i = 0; i < N; ++i) bv[i] = 0.f;
  }
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t base = row * E;
    float v[N], yv[N];
    IO::load_bf16(x + base, lane, per, v);
    if (y) {
      IO::load_bf16(y + base, lane, per, yv);
      if (drop_p > 0.f) {
        bool keep[N];
        IO::keep_mask(seed, step, site, static_cast<uint64_t>(base), drop_p, lane, per, keep);
#pragma unroll
        for (int i = 0; i < N; ++i) yv[i] = keep[i] ? yv[i] * keep_scale : 0.f;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] += yv[i];
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) s += v[i];  // lanes beyond `per` hold zeros in the scalar path
    const float mu = wsum(s) * inv_e;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const bool live = VEC || i < per;
      q += live ? (v[i] - mu) * (v[i] - mu) : 0.f;
    }
    const float rs = rsqrtf(wsum(q) * inv_e + eps);
    if (z) IO::store_bf16(z + base, lane, per, v);
    float o[N];
#pragma unroll
    for (int i = 0; i < N; ++i) o[i] = (v[i] - mu) * rs * wv[i] + bv[i];
    IO::store_bf16(out + base, lane, per, o);
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

int k_add_ln_fwd(const bf16* x, const bf16* y, const float* w, const float* b, bf16* z, bf16* out, float* mean,
                 float* rstd, int64_t rows, int E, float eps, float drop_p, const uint64_t* rng, uint32_t site,
                 cudaStream_t st) {
  GG_REQUIRE(E % 32 == 0 && (E % 256 == 0 ? E <= 1024 : E <= 256), "LayerNorm width %d unsupported", E);
  GG_REQUIRE(drop_p == 0.f || rng, "dropout needs rng state");
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const unsigned gb = static_cast<unsigned>(blocks);
#define LN_FWD(PER, VEC) \
  launch_k(add_ln_fwd_kernel<PER, VEC>, gb, 256, 0, st, x, y, w, b, z, out, mean, rstd, rows, E, eps, drop_p, rng, site)
  if (E == 256) LN_FWD(1, true);
  else if (E == 512) LN_FWD(2, true);
  else if (E == 768 || E == 1024) {
    if (E == 1024) LN_FWD(4, true);
    else { set_error("LayerNorm width 768 unsupported"); return GG_ERR_ARG; }
  } else LN_FWD(8, false);
#undef LN_FWD
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

template <int PER, bool VEC>
__global__ void __launch_bounds__(256)
    add_ln_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ z, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ w, bf16* __restrict__ dz,
                      bf16* __restrict__ dy, float* __restrict__ partial, int64_t rows, int E, float drop_p,
                      const uint64_t* __restrict__ rng, uint32_t site) {
  pdl_entry();
  using IO = RowIO<PER, VEC>;
  constexpr int N = IO::N;
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
  float wv[N], aw[N], ab[N];
  IO::load_f32(w, lane, per, wv);
#pragma unroll
  for (int i = 0; i < N; ++i) aw[i] = ab[i] = 0.f;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp; row < rows;
       row += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t base = row * E;
    const float mu = mean[row], rs = rstd[row];
    float d[N], xh[N];
    IO::load_bf16(dout + base, lane, per, d);
    IO::load_bf16(z + base, lane, per, xh);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const bool live = VEC || i < per;
      xh[i] = live ? (xh[i] - mu) * rs : 0.f;
      const float g = d[i] * wv[i];
      c1 += g;
      c2 += g * xh[i];
      aw[i] += d[i] * xh[i];
      ab[i] += d[i];
    }
    c1 = wsum(c1) * inv_e;
    c2 = wsum(c2) * inv_e;
    float v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = rs * (d[i] * wv[i] - c1 - xh[i] * c2);
    IO::store_bf16(dz + base, lane, per, v);
    if (dy) {
      if (drop_p > 0.f) {
        bool keep[N];
        IO::keep_mask(seed, step, site, static_cast<uint64_t>(base), drop_p, lane, per, keep);
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = keep[i] ? v[i] * keep_scale : 0.f;
      }
      IO::store_bf16(dy + base, lane, per, v);
    }
  }
#pragma unroll
  for (int k = 0; k < PER; ++k) {
#pragma unroll
    for (int t = 0; t < (VEC ? 8 : 1); ++t) {
      if (VEC || k < per) {
        const int e = IO::elem(k, t, lane);
        const int i = VEC ? k * 8 + t : k;
        sm[warp * 2 * E + e] = aw[i];
        sm[warp * 2 * E + E + e] = ab[i];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * E; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int wv2 = 0; wv2 < 8; ++wv2) t += sm[wv2 * 2 * E + i];
    partial[static_cast<int64_t>(blockIdx.x) * 2 * E + i] = t;
  }
}
// one warp per output element: lanes sum the block partials strided by 32, then a fixed shuffle tree
// (same order every run => deterministic)
__global__ void __launch_bounds__(256)
    ln_bwd_finish_kernel(const float* __restrict__ partial, int nblocks, int E, float* __restrict__ dw,
                         float* __restrict__ db) {
  pdl_entry();
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= 2 * E) return;
  float t = 0.f;
  for (int c = lane; c < nblocks; c += 32) t += partial[static_cast<int64_t>(c) * 2 * E + i];
  t = wsum(t);
  if (lane == 0) {
    if (i < E) dw[i] = t;
    else if (db) db[i - E] = t;
  }
}

template <int PER, bool VEC>
static int launch_ln_bwd(int blocks, size_t smem, cudaStream_t st, const bf16* dout, const bf16* z, const float* mean,
                         const float* rstd, const float* w, bf16* dz, bf16* dy, float* scratch, int64_t rows, int E,
                         float drop_p, const uint64_t* rng, uint32_t site) {
  if (smem > 48 * 1024) {
    static bool set = false;
    if (!set) {
      GG_CUDA_CHECK(cudaFuncSetAttribute(add_ln_bwd_kernel<PER, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         64 * 1024));
      set = true;
    }
  }
  launch_k(add_ln_bwd_kernel<PER, VEC>, blocks, 256, smem, st, dout, z, mean, rstd, w, dz, dy, scratch, rows, E, drop_p,
                                                         rng, site);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

int k_add_ln_bwd(const bf16* dout, const bf16* z, const float* mean, const float* rstd, const float* w, bf16* dz,
                 bf16* dy, float* dw, float* db, int64_t rows, int E, float drop_p, const uint64_t* rng,
                 uint32_t site, float* scratch, cudaStream_t st) {
  GG_REQUIRE(E % 32 == 0 && (E % 256 == 0 ? (E <= 1024 && E != 768) : E <= 256), "LayerNorm width %d unsupported", E);
  const int blocks = ln_bwd_blocks(rows);
  const size_t smem = static_cast<size_t>(8) * 2 * E * sizeof(float);
  int rc;
  if (E == 256) rc = launch_ln_bwd<1, true>(blocks, smem, st, dout, z, mean, rstd, w, dz, dy, scratch, rows, E, drop_p, rng, site);
  else if (E == 512) rc = launch_ln_bwd<2, true>(blocks, smem, st, dout, z, mean, rstd, w, dz, dy, scratch, rows, E, drop_p, rng, site);
  else if (E == 1024) rc = launch_ln_bwd<4, true>(blocks, smem, st, dout, z, mean, rstd, w, dz, dy, scratch, rows, E, drop_p, rng, site);
  else rc = launch_ln_bwd<8, false>(blocks, smem, st, dout, z, mean, rstd, w, dz, dy, scratch, rows, E, drop_p, rng, site);
  if (rc) return rc;
  if (dw) return k_ln_bwd_finish(scratch, rows, E, dw, db, st);
  return GG_OK;
}

// dw / db from the per-block partials k_add_ln_bwd left in `scratch` (call it with dw = NULL to defer this)
int k_ln_bwd_finish(const float* scratch, int64_t rows, int E, float* dw, float* db, cudaStream_t st) {
  const int blocks = ln_bwd_blocks(rows);
  launch_k(ln_bwd_finish_kernel, (2 * E + 7) / 8, 256, 0, st, scratch, blocks, E, dw, db);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg
