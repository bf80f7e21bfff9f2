// Counter-based RNG for dropout: Philox4x32-7 (Salmon et al., SC'11: 7 rounds is the fewest that
// passes BigCrush), written out here. One call yields 128 bits = eight 16-bit uniforms, i.e. eight
// consecutive elements per call; the keep test is u16 >= round(p * 65536) (p = 0.1 -> 0.100006).
// A dropout decision is a pure function of (seed, step, call-site, element index), so the
// backward pass regenerates the forward mask instead of storing it.
// (The reference draws dropout from torch's Philox stream inside nn.TransformerEncoderLayer,
//  src/conditional_gan_cross_attention_with_film.py:114-116; stream-identical masks are not
//  reproducible from a custom kernel, so parity is exact with p=0 and statistical with p>0.)
#pragma once
#include <stdint.h>

namespace gg {

struct u32x4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi,
                                                   uint32_t& lo) {
  const uint64_t p = static_cast<uint64_t>(a) * static_cast<uint64_t>(b);
  hi = static_cast<uint32_t>(p >> 32);
  lo = static_cast<uint32_t>(p);
}

__host__ __device__ __forceinline__ u32x4 philox4x32_7(u32x4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo32(0xD2511F53u, c.x, hi0, lo0);
    mulhilo32(0xCD9E8D57u, c.z, hi1, lo1);
    u32x4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

__host__ __device__ __forceinline__ u32x4 dropout_words(uint64_t seed, uint64_t step,
                                                        uint32_t site, uint64_t group) {
  u32x4 c;
  c.x = static_cast<uint32_t>(group);
  c.y = static_cast<uint32_t>(group >> 32);
  c.z = site;
  c.w = static_cast<uint32_t>(step);
  return philox4x32_7(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}

// 16-bit threshold of a dropout probability: an element is dropped when its uniform u16 < thr
__host__ __device__ __forceinline__ uint32_t dropout_thr(float p) {
  return static_cast<uint32_t>(p * 65536.0f + 0.5f);
}

// keep decisions of the 8 consecutive elements of one group (bit j of the result = element j kept)
__host__ __device__ __forceinline__ uint32_t keep_bits8(const u32x4& r, uint32_t thr) {
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) >= thr ? 1u : 0u) | ((r.x >> 16) >= thr ? 2u : 0u);
  m |= ((r.y & 0xFFFFu) >= thr ? 4u : 0u) | ((r.y >> 16) >= thr ? 8u : 0u);
  m |= ((r.z & 0xFFFFu) >= thr ? 16u : 0u) | ((r.z >> 16) >= thr ? 32u : 0u);
  m |= ((r.w & 0xFFFFu) >= thr ? 64u : 0u) | ((r.w >> 16) >= thr ? 128u : 0u);
  return m;
}

// keep decision for element `idx` of the tensor at this site.
__host__ __device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t step, uint32_t site,
                                                      uint64_t idx, float p) {
  const u32x4 r = dropout_words(seed, step, site, idx >> 3);
  const uint32_t sel = static_cast<uint32_t>(idx & 7);
  const uint32_t w = (sel >> 1) == 0 ? r.x : (sel >> 1) == 1 ? r.y : (sel >> 1) == 2 ? r.z : r.w;
  const uint32_t u = (sel & 1) ? (w >> 16) : (w & 0xFFFFu);
  return u >= dropout_thr(p);
}

// Same decision as dropout_keep(), remembering the last group's words: consecutive indices cost one
// Philox call per 8 elements.
struct DropoutStream {
  uint64_t seed, step;
  uint32_t site, thr;
  uint64_t grp = ~0ull;
  u32x4 r{0, 0, 0, 0};
  __host__ __device__ __forceinline__ DropoutStream(uint64_t seed_, uint64_t step_, uint32_t site_, float p)
      : seed(seed_), step(step_), site(site_), thr(dropout_thr(p)) {}
  __host__ __device__ __forceinline__ bool keep(uint64_t idx) {
    const uint64_t gidx = idx >> 3;
    if (gidx != grp) {
      grp = gidx;
      r = dropout_words(seed, step, site, gidx);
    }
    const uint32_t sel = static_cast<uint32_t>(idx & 7);
    const uint32_t w = (sel >> 1) == 0 ? r.x : (sel >> 1) == 1 ? r.y : (sel >> 1) == 2 ? r.z : r.w;
    return ((sel & 1) ? (w >> 16) : (w & 0xFFFFu)) >= thr;
  }
};

}  // namespace gg
