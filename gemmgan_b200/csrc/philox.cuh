// Counter-based RNG for dropout: Philox4x32-10 (Salmon et al., SC'11), written out here.
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

__host__ __device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
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
  return philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}

__host__ __device__ __forceinline__ bool keep_from_word(uint32_t w, float p) {
  // uniform in [0,1) with 24 bits; keep with probability 1-p
  return (static_cast<float>(w >> 8) * (1.0f / 16777216.0f)) >= p;
}

// keep decision for element `idx` of the tensor at this site.
__host__ __device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t step, uint32_t site,
                                                      uint64_t idx, float p) {
  const u32x4 r = dropout_words(seed, step, site, idx >> 2);
  const uint32_t lane = static_cast<uint32_t>(idx & 3);
  const uint32_t w = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
  return keep_from_word(w, p);
}

}  // namespace gg
