// Host build of gemmgan_b200/csrc/layernorm.cu (see emu.h): the residual-add + dropout + LayerNorm kernels of the
// encoder layers, exported through plain C wrappers of the internal launch API (kernels.h).
#include "emu.h"

namespace gg {
thread_local float sm[64 * 1024 / 4];  // `extern __shared__ float sm[]` of add_ln_bwd_kernel (at most 64 KiB)
}

#include "../../gemmgan_b200/csrc/layernorm.cu"

using gg::bf16;

extern "C" int emu_add_ln_fwd(const bf16* x, const bf16* y, const float* w, const float* b, bf16* z, bf16* out,
                              float* mean, float* rstd, int64_t rows, int E, float eps, float drop_p,
                              const uint64_t* rng, uint32_t site) {
  return gg::k_add_ln_fwd(x, y, w, b, z, out, mean, rstd, rows, E, eps, drop_p, rng, site, nullptr);
}
extern "C" int emu_add_ln_bwd(const bf16* dout, const bf16* z, const float* mean, const float* rstd, const float* w,
                              bf16* dz, bf16* dy, float* dw, float* db, int64_t rows, int E, float drop_p,
                              const uint64_t* rng, uint32_t site, float* scratch) {
  return gg::k_add_ln_bwd(dout, z, mean, rstd, w, dz, dy, dw, db, rows, E, drop_p, rng, site, scratch, nullptr);
}
extern "C" int emu_ln_bwd_finish(const float* scratch, int64_t rows, int E, float* dw, float* db) {
  return gg::k_ln_bwd_finish(scratch, rows, E, dw, db, nullptr);
}
extern "C" int64_t emu_ln_bwd_scratch_floats(int64_t rows, int E) { return gg::ln_bwd_scratch_floats(rows, E); }
