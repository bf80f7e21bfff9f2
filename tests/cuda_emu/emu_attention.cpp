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

#include "emu_attention_ptx.h"

#include "../../gemmgan_b200/csrc/attention.cu"

extern "C" int gg_attention_fwd(const gg_attn_args* a, void* stream) {
  GG_REQUIRE(a && a->q && a->k && a->v && a->o, "null argument");
  return gg::k_attention_fwd(*reinterpret_cast<const gg::AttnArgs*>(a), reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int gg_attention_bwd(const gg_attn_args* a, void* stream) {
  GG_REQUIRE(a && a->q && a->k && a->v && a->dout && a->dq && a->dk && a->dv, "null argument");
  return gg::k_attention_bwd(*reinterpret_cast<const gg::AttnArgs*>(a), reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int64_t gg_dropout_bits_words(int64_t n_elems) { return gg::dropout_bits_words(n_elems); }
extern "C" int gg_dropout_bits(const uint64_t* rng, uint32_t site, float p, int64_t n_elems, uint32_t* out, void* stream) {
  return gg::k_dropout_bits(rng, site, p, n_elems, out, reinterpret_cast<cudaStream_t>(stream));
}
