// Shared GEMM epilogue: one implementation of the post-accumulation math, used by the
// tcgen05 kernel (values read from tensor memory, W = 32 columns per thread), the split-K reducer
// (values summed from fp32 partials, W = 4 so that a warp covers 512 contiguous bytes) and the
// CUDA-core check kernel. Works on a run of <= W consecutive columns of one output row in registers.
#pragma once
#include "../../include/gemmgan.h"
#include "philox.cuh"
#include <cuda_bf16.h>

namespace gg {

template <int W>
__device__ __forceinline__ void load_row_chunk(const void* base, int is_f32, int64_t ld, int m,
                                               int n0, int ncols, float* dst) {
  static_assert(W % 4 == 0, "chunk width must be a multiple of 4");
  if (is_f32) {
    const float* p = reinterpret_cast<const float*>(base) + static_cast<int64_t>(m) * ld + n0;
    if (ncols == W && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < W / 4; ++j) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
        dst[4 * j + 0] = t.x; dst[4 * j + 1] = t.y; dst[4 * j + 2] = t.z; dst[4 * j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) dst[j] = (j < ncols) ? __ldg(p + j) : 0.f;
    }
  } else {
    const __nv_bfloat16* p =
        reinterpret_cast<const __nv_bfloat16*>(base) + static_cast<int64_t>(m) * ld + n0;
    if (ncols == W && (reinterpret_cast<uintptr_t>(p) & (W >= 8 ? 15 : 7)) == 0) {
      if (W >= 8) {
#pragma unroll
        for (int j = 0; j < W / 8; ++j) {
          uint4 t = __ldg(reinterpret_cast<const uint4*>(p) + j);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 f = __bfloat1622float2(h[q]);
            dst[8 * j + 2 * q] = f.x;
            dst[8 * j + 2 * q + 1] = f.y;
          }
        }
      } else {
        uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
        float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
        dst[0] = f0.x; dst[1] = f0.y; dst[2] = f1.x; dst[3] = f1.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) dst[j] = (j < ncols) ? __bfloat162float(p[j]) : 0.f;
    }
  }
}

// v[0..ncols) holds the raw accumulators of row m, columns [n0, n0+ncols): everything between the
// accumulator and the stores (scale, bias, pre-add, activation, dropout, mask, residual).
template <int W>
__device__ __forceinline__ void epilogue_math(const gg_epilogue& e, int N, int m, int n0,
                                              int ncols, float* v) {
  float t[W];
  if (e.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] *= e.alpha;
  }
  if (e.bias) {
    load_row_chunk<W>(e.bias, 1, 0, 0, n0, ncols, t);
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] += t[j];
  }
  if (e.pre) {
    load_row_chunk<W>(e.pre, e.pre_f32, e.pre_ld, m, n0, ncols, t);
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] += t[j];
  }
  if (e.act == GG_ACT_LEAKY) {
    const float s = e.slope;
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] = v[j] > 0.f ? v[j] : s * v[j];
  } else if (e.act == GG_ACT_FILM) {
    // tanh(x) = 1 - 2 / (e^{2x} + 1) on the SFU (|abs error| ~1e-7: gamma multiplies bf16 patch embeddings). tanhf()'s
    // branchy accurate path made this epilogue 20 of the 30 us of the FiLM GEMM, the first kernel of every tower pass.
    const int half = N >> 1;
    if (n0 + W <= half) {
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = 1.f - 2.f / (__expf(2.f * fminf(v[j], 40.f)) + 1.f);
    } else if (n0 >= half) {
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = fminf(fmaxf(v[j], -5.0f), 5.0f);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        v[j] = (n0 + j < half) ? 1.f - 2.f / (__expf(2.f * fminf(v[j], 40.f)) + 1.f) : fminf(fmaxf(v[j], -5.0f), 5.0f);
    }
  }
  if (e.drop_p > 0.f) {
    const uint64_t seed = e.rng[0], step = e.rng[1];
    const float keep_scale = 1.0f / (1.0f - e.drop_p);
    const uint64_t base = static_cast<uint64_t>(m) * static_cast<uint64_t>(N) + n0;
    if (W % 8 == 0 && (base & 7) == 0) {  // one Philox call per 8 consecutive elements
      const uint32_t thr = dropout_thr(e.drop_p);
#pragma unroll
      for (int g = 0; g < W / 8; ++g) {
        const uint32_t kb = keep_bits8(dropout_words(seed, step, e.site, (base >> 3) + g), thr);
#pragma unroll
        for (int t = 0; t < 8; ++t) v[8 * g + t] = ((kb >> t) & 1u) ? v[8 * g + t] * keep_scale : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        v[j] = dropout_keep(seed, step, e.site, base + j, e.drop_p) ? v[j] * keep_scale : 0.f;
    }
  }
  if (e.mask) {
    load_row_chunk<W>(e.mask, e.mask_f32, e.mask_ld, m, n0, ncols, t);
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] *= (t[j] > 0.f ? e.mask_pos : e.mask_neg);
  }
  if (e.res) {
    load_row_chunk<W>(e.res, e.res_f32, e.res_ld, m, n0, ncols, t);
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] += t[j];
  }
}

__device__ __forceinline__ int epilogue_out_row(const gg_epilogue& e, int m) {
  return e.row_div > 0 ? (m / e.row_div) * e.row_mul + e.row_add + (m % e.row_div) : m;
}

// Direct global stores of a finished run of <= W columns of row m (either output may be skipped).
template <int W>
__device__ __forceinline__ void epilogue_store(const gg_epilogue& e, int m, int n0, int ncols,
                                               const float* v, bool do_f32, bool do_bf16) {
  const int row = epilogue_out_row(e, m);
  if (do_f32 && e.out_f32) {
    float* p = e.out_f32 + static_cast<int64_t>(row) * e.ld_f32 + n0;
    if (e.accum_f32) {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (j < ncols) p[j] += v[j];
    } else if (ncols == W && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < W / 4; ++j)
        reinterpret_cast<float4*>(p)[j] =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (j < ncols) p[j] = v[j];
    }
  }
  if (do_bf16 && e.out_bf16) {
    __nv_bfloat16* p =
        reinterpret_cast<__nv_bfloat16*>(e.out_bf16) + static_cast<int64_t>(row) * e.ld_bf16 + n0;
    if (ncols == W && (reinterpret_cast<uintptr_t>(p) & (W >= 8 ? 15 : 7)) == 0) {
      if (W >= 8) {
#pragma unroll
        for (int j = 0; j < W / 8; ++j) {
          uint4 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            h[q] = __floats2bfloat162_rn(v[8 * j + 2 * q], v[8 * j + 2 * q + 1]);
          reinterpret_cast<uint4*>(p)[j] = pk;
        }
      } else {
        uint2 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
        h[0] = __floats2bfloat162_rn(v[0], v[1]);
        h[1] = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (j < ncols) p[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

template <int W>
__device__ __forceinline__ void epilogue_chunk(const gg_epilogue& e, int N, int m, int n0,
                                               int ncols, float* v) {
  epilogue_math<W>(e, N, m, n0, ncols, v);
  epilogue_store<W>(e, m, n0, ncols, v, true, true);
}

}  // namespace gg
