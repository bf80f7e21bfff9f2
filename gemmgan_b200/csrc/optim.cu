// Optimizer side of the step: global gradient norm + clip coefficient, one fused flat-buffer
// update kernel (RMSprop / Adam / AdamW) with 128-bit accesses, and the fp32 -> bf16 weight
// shadow refresh. All HBM-bound; Adam moves 28 B/param (r: p,g,m,v  w: p,m,v), RMSprop 20 B/param.
//
// Replaces torch.optim.{RMSprop,Adam,AdamW}.step() and clip_grad_norm_
// (src/conditional_gan_cross_attention_with_film.py:320-331, :414-415, :457-458).
// Update rules follow torch 2.11 defaults: RMSprop(alpha=.99, eps=1e-8, no momentum, not centered),
// Adam(betas=(.9,.99), eps=1e-8), AdamW(weight_decay=.01); clip: coef = min(1, max/(norm+1e-6)).
#include "host_util.h"
#include "pdl.cuh"
#include "kernels.h"

namespace gg {

constexpr int NORM_BLOCKS = 592;  // 4 per SM; fixed so the reduction order never changes

__global__ void __launch_bounds__(256) sumsq_stage1_kernel(const float* __restrict__ g, int64_t n,
                                                           float* __restrict__ partial) {
  pdl_entry();
  __shared__ float sm[256];
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x;
  int64_t hi = lo + per;
  if (hi > n) hi = n;
  float acc = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
    const float x = g[i];
    acc = fmaf(x, x, acc);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(256) sumsq_stage2_kernel(const float* __restrict__ partial, int nparts,
                                                           float max_norm, float* __restrict__ out) {
  pdl_entry();
  __shared__ double sm[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) acc += static_cast<double>(partial[i]);
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = static_cast<float>(sqrt(sm[0]));
    out[0] = norm;
    float coef = 1.f;
    if (max_norm > 0.f) {
      coef = max_norm / (norm + 1e-6f);
      if (coef > 1.f) coef = 1.f;
    }
    out[1] = coef;
  }
}
int k_grad_norm_clip(const float* g, int64_t n, float max_norm, float* norm_out, float* scratch,
                     cudaStream_t st) {
  launch_k(sumsq_stage1_kernel, NORM_BLOCKS, 256, 0, st, g, n, scratch);
  GG_LAUNCH_CHECK();
  launch_k(sumsq_stage2_kernel, 1, 256, 0, st, scratch, NORM_BLOCKS, max_norm, norm_out);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

template <int KIND>
__device__ __forceinline__ void optim_elem(float& p, float& g, float& m, float& v, float lr, float coef,
                                           float bc1, float rsqrt_bc2) {
  g *= coef;
  if (KIND == GG_OPT_RMSPROP) {
    v = 0.99f * v + 0.01f * g * g;  // mul_(alpha).addcmul_(g, g, 1-alpha)
    p -= lr * g / (sqrtf(v) + 1e-8f);
  } else {
    if (KIND == GG_OPT_ADAMW) p *= (1.f - lr * 0.01f);
    m = m + (g - m) * 0.1f;                 // lerp_(g, 1-beta1), beta1 = 0.9
    v = 0.99f * v + 0.01f * g * g;          // beta2 = 0.99
    const float denom = sqrtf(v) * rsqrt_bc2 + 1e-8f;
    p -= (lr / bc1) * (m / denom);
  }
}

template <int KIND>
__global__ void __launch_bounds__(256)
    optim_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float lr, const float* __restrict__ coef_ptr, const float* __restrict__ step_count) {
  pdl_entry();
  const float coef = coef_ptr ? coef_ptr[0] : 1.f;
  float bc1 = 1.f, rsqrt_bc2 = 1.f;
  if (KIND != GG_OPT_RMSPROP) {
    const float t = step_count[0] + 1.f;
    bc1 = 1.f - powf(0.9f, t);
    rsqrt_bc2 = rsqrtf(1.f - powf(0.99f, t));
  }
  const int64_t n4 = n >> 2;
  const bool write_g = coef_ptr != nullptr;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    float4 G = reinterpret_cast<float4*>(g)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float4 M = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KIND != GG_OPT_RMSPROP) M = reinterpret_cast<float4*>(m)[i];
    optim_elem<KIND>(P.x, G.x, M.x, V.x, lr, coef, bc1, rsqrt_bc2);
    optim_elem<KIND>(P.y, G.y, M.y, V.y, lr, coef, bc1, rsqrt_bc2);
    optim_elem<KIND>(P.z, G.z, M.z, V.z, lr, coef, bc1, rsqrt_bc2);
    optim_elem<KIND>(P.w, G.w, M.w, V.w, lr, coef, bc1, rsqrt_bc2);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(v)[i] = V;
    if (KIND != GG_OPT_RMSPROP) reinterpret_cast<float4*>(m)[i] = M;
    if (write_g) reinterpret_cast<float4*>(g)[i] = G;
  }
  // tail (n % 4 elements)
  const int64_t i = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    float P = p[i], G = g[i], V = v[i], M = (KIND != GG_OPT_RMSPROP) ? m[i] : 0.f;
    optim_elem<KIND>(P, G, M, V, lr, coef, bc1, rsqrt_bc2);
    p[i] = P;
    v[i] = V;
    if (KIND != GG_OPT_RMSPROP) m[i] = M;
    if (write_g) g[i] = G;
  }
}
__global__ void bump_step_kernel(float* step_count) {
  pdl_entry(); step_count[0] += 1.f; }

int k_optim_step(int kind, float* p, float* g, float* m, float* v, int64_t n, float lr, const float* coef_ptr,
                 float* step_count, cudaStream_t st, bool bump_step) {
  GG_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(v) & 15) == 0 && (!m || (reinterpret_cast<uintptr_t>(m) & 15) == 0),
             "optimizer buffers must be 16-byte aligned");
  int64_t blocks = ((n >> 2) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const unsigned gb = static_cast<unsigned>(blocks);
  if (kind == GG_OPT_RMSPROP) launch_k(optim_kernel<GG_OPT_RMSPROP>, gb, 256, 0, st, p, g, m, v, n, lr, coef_ptr, step_count);
  else if (kind == GG_OPT_ADAM) launch_k(optim_kernel<GG_OPT_ADAM>, gb, 256, 0, st, p, g, m, v, n, lr, coef_ptr, step_count);
  else if (kind == GG_OPT_ADAMW) launch_k(optim_kernel<GG_OPT_ADAMW>, gb, 256, 0, st, p, g, m, v, n, lr, coef_ptr, step_count);
  else {
    set_error("unknown optimizer kind %d", kind);
    return GG_ERR_ARG;
  }
  GG_LAUNCH_CHECK();
  if (bump_step) {  // the engine folds the increment into its shadow refresh instead (one launch less per step)
    launch_k(bump_step_kernel, 1, 1, 0, st, step_count);
    GG_LAUNCH_CHECK();
  }
  return GG_OK;
}

// One (segment, 4-column run) per thread iteration: float4 load, 8-byte bf16x4 store, 32-bit index math.
__global__ void __launch_bounds__(256)
    refresh_shadows_kernel(const float* __restrict__ p, bf16* __restrict__ shadow, const ShadowSeg* __restrict__ segs,
                           float* bump) {
  pdl_entry();
  if (bump && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) bump[0] += 1.f;  // optimizer step counter
  const ShadowSeg s = segs[blockIdx.y];
  const float* src = p + s.p_off + s.col0;
  bf16* dst = shadow + s.s_off;
  if (s.transpose) {  // dst[c, r] = src[r, c]: consecutive threads walk r (coalesced bf16 writes; the reads hit L2)
    const int64_t total = static_cast<int64_t>(s.rows) * s.ncols;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t c = i / s.rows;
      const int r = static_cast<int>(i % s.rows);
      dst[c * s.s_ld + r] = __float2bfloat16_rn(src[static_cast<int64_t>(r) * s.cols + c]);
    }
    return;
  }
  const bool vec = (s.ncols % 4 == 0) && (s.cols % 4 == 0) && (s.s_ld % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
  if (vec) {
    const unsigned c4 = static_cast<unsigned>(s.ncols) >> 2;
    const unsigned total = static_cast<unsigned>(s.rows) * c4;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const unsigned r = i / c4, c = (i - r * c4) * 4;
      const float4 f = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(r) * s.cols + c));
      __nv_bfloat162 lo = __floats2bfloat162_rn(f.x, f.y), hi = __floats2bfloat162_rn(f.z, f.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(dst + static_cast<int64_t>(r) * s.s_ld + c) = pk;
    }
  } else {
    const int64_t total = static_cast<int64_t>(s.rows) * s.ncols;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t r = i / s.ncols;
      const int c = static_cast<int>(i % s.ncols);
      dst[r * s.s_ld + c] = __float2bfloat16_rn(src[r * s.cols + c]);
    }
  }
}
int k_refresh_shadows(const float* p, bf16* shadow, const ShadowSeg* segs_dev, int nseg, int /*max_rows*/,
                      cudaStream_t st, float* bump_step) {
  if (nseg <= 0) {
    if (bump_step) {
      launch_k(bump_step_kernel, 1, 1, 0, st, bump_step);
      GG_LAUNCH_CHECK();
    }
    return GG_OK;
  }
  dim3 grid(148, nseg);
  launch_k(refresh_shadows_kernel, grid, 256, 0, st, p, shadow, segs_dev, bump_step);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg
