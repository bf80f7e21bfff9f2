// Host versions of the five PTX wrappers of gemmgan_b200/csrc/attention.cu (compiled out there by GG_EMULATED_PTX) and
// the dynamic shared-memory arrays its kernels declare. Include after emu.h, before attention.cu. See emu_attention.cpp.
#pragma once

namespace gg {
#define EMU_DYN_SMEM(name) thread_local __attribute__((aligned(16))) uint8_t name[227 * 1024]
EMU_DYN_SMEM(smem_att);
EMU_DYN_SMEM(smem_self);
EMU_DYN_SMEM(smem_mid);
EMU_DYN_SMEM(smem_long);

static inline void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  if (valid) memcpy(smem_dst, gsrc, 16);
  else memset(smem_dst, 0, 16);
}
static inline void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const void* rows[32];
  emu_warp_all_gather(p, rows);
  const int t = emu_lane();
  for (int j = 0; j < 4; ++j) memcpy(&r[j], static_cast<const char*>(rows[j * 8 + t / 4]) + (t % 4) * 4, 4);
}
static inline void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const void* rows[32];
  emu_warp_all_gather(p, rows);
  const int t = emu_lane();
  for (int j = 0; j < 4; ++j) {
    uint16_t lo, hi;
    memcpy(&lo, static_cast<const char*>(rows[j * 8 + 2 * (t % 4)]) + (t / 4) * 2, 2);
    memcpy(&hi, static_cast<const char*>(rows[j * 8 + 2 * (t % 4) + 1]) + (t / 4) * 2, 2);
    r[j] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
  }
}
static inline float emu_bf16_half(uint32_t reg, int half) {
  const uint32_t bits = (half ? (reg >> 16) : (reg & 0xffffu)) << 16;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
static inline void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  struct Frag { uint32_t a[4], b[2]; } mine = {{a[0], a[1], a[2], a[3]}, {b0, b1}}, all[32];
  emu_warp_all_gather(mine, all);
  const int t = emu_lane(), g = t / 4, q = t % 4;
  for (int i = 0; i < 4; ++i) {
    const int row = g + (i >= 2 ? 8 : 0), col = 2 * q + (i & 1);
    float acc = c[i];
    for (int k = 0; k < 16; ++k) {
      const Frag& fa = all[(row % 8) * 4 + (k % 8) / 2];
      const Frag& fb = all[col * 4 + (k % 8) / 2];
      const float av = emu_bf16_half(fa.a[(row >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0)], k % 2);
      const float bv = emu_bf16_half(fb.b[k >= 8 ? 1 : 0], k % 2);
      acc = fmaf(av, bv, acc);
    }
    c[i] = acc;
  }
}
static inline float fast_ex2(float x) { return exp2f(x); }
}  // namespace gg

