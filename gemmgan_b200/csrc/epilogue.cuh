// Shared GEMM epilogue: one implementation of the post-accumulation math, used by the
// tcgen05 kernel (values read from tensor memory), the split-K reducer (values summed from
// fp32 partials) and the CUDA-core check kernel. Works on a run of <= 32 consecutive columns
// of one output row held in registers.
#pragma once
#include "../../include/gemmgan.h"
#include "philox.cuh"
#include <cuda_bf16.h>

namespace gg {

__device__ __forceinline__ void load_row_chunk(const void* base, int is_f32, int64_t ld, int m,
                                               int n0, int ncols, float* dst) {
  if (is_f32) {
    const float* p = reinterpret_cast<const float*>(base) + static_cast<int64_t>(m) * ld + n0;
    if (ncols == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
        dst[4 * j + 0] = t.x; dst[4 * j + 1] = t.y; dst[4 * j + 2] = t.z; dst[4 * j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) dst[j] = (j < ncols) ? __ldg(p + j) : 0.f;
    }
  } else {
    const __nv_bfloat16* p =
        reinterpret_cast<const __nv_bfloat16*>(base) + static_cast<int64_t>(m) * ld + n0;
    if (ncols == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
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
#pragma unroll
      for (int j = 0; j < 32; ++j) dst[j] = (j < ncols) ? __bfloat162float(p[j]) : 0.f;
    }
  }
}

// v[0..ncols) holds the raw accumulators of row m, columns [n0, n0+ncols).
__device__ __forceinline__ void epilogue_chunk(const gg_epilogue& e, int N, int m, int n0,
                                               int ncols, float* v) {
  float t[32];
  if (e.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
  }
  if (e.bias) {
    load_row_chunk(e.bias, 1, 0, 0, n0, ncols, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += t[j];
  }
  if (e.pre) {
    load_row_chunk(e.pre, e.pre_f32, e.pre_ld, m, n0, ncols, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += t[j];
  }
  if (e.act == GG_ACT_LEAKY) {
    const float s = e.slope;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : s * v[j];
  } else if (e.act == GG_ACT_FILM) {
    const int half = N >> 1;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      v[j] = (n0 + j < half) ? tanhf(v[j]) : fminf(fmaxf(v[j], -5.0f), 5.0f);
  }
  if (e.drop_p > 0.f) {
    const uint64_t seed = e.rng[0], step = e.rng[1];
    const float keep_scale = 1.0f / (1.0f - e.drop_p);
    const uint64_t base = static_cast<uint64_t>(m) * static_cast<uint64_t>(N) + n0;
    if ((base & 3) == 0) {  // one Philox call per 4 consecutive elements
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const u32x4 r = dropout_words(seed, step, e.site, (base >> 2) + g);
        v[4 * g + 0] = keep_from_word(r.x, e.drop_p) ? v[4 * g + 0] * keep_scale : 0.f;
        v[4 * g + 1] = keep_from_word(r.y, e.drop_p) ? v[4 * g + 1] * keep_scale : 0.f;
        v[4 * g + 2] = keep_from_word(r.z, e.drop_p) ? v[4 * g + 2] * keep_scale : 0.f;
        v[4 * g + 3] = keep_from_word(r.w, e.drop_p) ? v[4 * g + 3] * keep_scale : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = dropout_keep(seed, step, e.site, base + j, e.drop_p) ? v[j] * keep_scale : 0.f;
    }
  }
  if (e.mask) {
    load_row_chunk(e.mask, e.mask_f32, e.mask_ld, m, n0, ncols, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= (t[j] > 0.f ? e.mask_pos : e.mask_neg);
  }
  if (e.res) {
    load_row_chunk(e.res, e.res_f32, e.res_ld, m, n0, ncols, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += t[j];
  }
  int row = m;
  if (e.row_div > 0) row = (m / e.row_div) * e.row_mul + e.row_add + (m % e.row_div);
  if (e.out_f32) {
    float* p = e.out_f32 + static_cast<int64_t>(row) * e.ld_f32 + n0;
    if (e.accum_f32) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) p[j] += v[j];
    } else if (ncols == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(p)[j] =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) p[j] = v[j];
    }
  }
  if (e.out_bf16) {
    __nv_bfloat16* p =
        reinterpret_cast<__nv_bfloat16*>(e.out_bf16) + static_cast<int64_t>(row) * e.ld_bf16 + n0;
    if (ncols == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          h[q] = __floats2bfloat162_rn(v[8 * j + 2 * q], v[8 * j + 2 * q + 1]);
        reinterpret_cast<uint4*>(p)[j] = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) p[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

}  // namespace gg
