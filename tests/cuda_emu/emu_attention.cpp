// Host build of gemmgan_b200/csrc/attention.cu (see emu.h): every attention kernel of the step, including the
// mma.sync / ldmatrix ones, runs unchanged. The file's five PTX wrappers are compiled out (GG_EMULATED_PTX) and
// replaced by the host versions below, which follow the PTX ISA's fragment layouts:
//   cp.async 16 B (zero-fill when the source is invalid)            -> memcpy / memset, complete at once
//   ldmatrix.m8n8.x4[.trans].b16: lane l supplies the address of row l%8 of matrix l/8; thread t receives from matrix j
//       row t/4, columns 2(t%4), 2(t%4)+1           (.trans: rows 2(t%4), 2(t%4)+1 of column t/4)
//   mma.m16n8k16 bf16 (g = t/4, q = t%4):  a0 = A[g][2q..], a1 = A[g+8][2q..], a2 = A[g][2q+8..], a3 = A[g+8][2q+8..];
//       b0 = B[2q..][g], b1 = B[2q+8..][g];  c0,c1 = C[g][2q, 2q+1], c2,c3 = C[g+8][2q, 2q+1]; fp32 accumulation
//   ex2.approx -> exp2f
#define GG_EMULATED_PTX 1
#include "emu.h"

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

#include "../../gemmgan_b200/csrc/attention.cu"

extern "C" int gg_attention_fwd(const gg_attn_args* a, void* stream) {
  GG_REQUIRE(a && a->q && a->k && a->v && a->o, "null argument");
  return gg::k_attention_fwd(*reinterpret_cast<const gg::AttnArgs*>(a), reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int gg_attention_bwd(const gg_attn_args* a, void* stream) {
  GG_REQUIRE(a && a->q && a->k && a->v && a->dout && a->dq && a->dk && a->dv, "null argument");
  return gg::k_attention_bwd(*reinterpret_cast<const gg::AttnArgs*>(a), reinterpret_cast<cudaStream_t>(stream));
}
