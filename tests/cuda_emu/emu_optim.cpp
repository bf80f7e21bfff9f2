// Host build of gemmgan_b200/csrc/optim.cu (see emu.h). gg_optim_step itself lives in engine.cu next to the tcgen05
// code, which cannot be compiled for the host; its ten lines of glue around the two kernels of optim.cu are restated
// here under the same name and signature (include/gemmgan.h).
#include "emu.h"

#include "../../gemmgan_b200/csrc/optim.cu"

extern "C" int gg_optim_step(int kind, float* p, float* g, float* m, float* v, int64_t n, float lr, float max_norm,
                             float* step_count, float* norm_out2, float* scratch, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float* coef = nullptr;
  if (max_norm > 0.f) {
    GG_REQUIRE(norm_out2 && scratch, "clipping needs norm_out2[2] and scratch[>=592]");
    GG_TRY_RC(gg::k_grad_norm_clip(g, n, max_norm, norm_out2, scratch, st));
    coef = norm_out2 + 1;
  }
  return gg::k_optim_step(kind, p, g, m, v, n, lr, coef, step_count, st, true);
}

// [rows, cols] fp32 parameter block at p + p_off -> bf16 shadow at shadow + s_off (one segment per call)
extern "C" int emu_refresh_shadows(const float* p, __nv_bfloat16* shadow, const gg::ShadowSeg* segs, int nseg,
                                   float* bump) {
  return gg::k_refresh_shadows(p, shadow, segs, nseg, 0, nullptr, bump);
}
