// Host build of the WHOLE engine (see emu.h): engine.cu with every kernel file it sequences, in its all-CUDA-core
// configuration (gg_model_cfg.gemm_impl = GG_IMPL_SIMT_F32: every GEMM of the step runs on the check kernel of gemm.cu
// with the shared epilogue; weight gradients take the per-problem path instead of the grouped tcgen05 kernel). One
// critic step / generator step / optimizer step of any variant then runs on the CPU through the same gg_engine_* C
// ABI and gemmgan_b200/runtime.py, and is compared with the oracle (tests/test_engine_emulated.py).
// Not emulated: the tcgen05 / TMA main loops (gemm_tc_kernel, wgrad_group_kernel) — k_wgrad_group is a stub that
// reports GG_ERR_ARCH — and everything about streams (launches complete synchronously, lanes are tokens).
#define GG_EMULATED_PTX 1
#include <cuda.h>

#include <mutex>
#include <vector>

#include "emu.h"
#include "emu_attention_ptx.h"

namespace gg {
thread_local float sm[64 * 1024 / 4];  // layernorm.cu: add_ln_bwd_kernel
thread_local uint8_t smem_raw[1024];   // gemm.cu: gemm_tc_kernel (declared, never run here)
}  // namespace gg

#include "../../gemmgan_b200/csrc/elementwise.cu"
#include "../../gemmgan_b200/csrc/optim.cu"
#include "../../gemmgan_b200/csrc/layernorm.cu"
#include "../../gemmgan_b200/csrc/attention.cu"
#include "../../gemmgan_b200/csrc/gemm.cu"
#include "../../gemmgan_b200/csrc/evalmetrics.cu"

namespace gg {
int64_t wgrad_group_workspace_bytes(int64_t max_output_elems) { return GROUP_COUNTER_BYTES + 4 * max_output_elems; }
int k_wgrad_group(const WgradItem*, int, void*, int64_t, cudaStream_t) {
  set_error("the grouped weight-gradient kernel is tcgen05 only (not available in the host emulation)");
  return GG_ERR_ARCH;
}
int k_xw_f32(const float*, const float*, int, int, const bf16*, int64_t, float*, void*, int64_t, cudaStream_t) {
  set_error("the fp32-input critic layer-1 kernel is tcgen05 only (not available in the host emulation)");
  return GG_ERR_ARCH;
}
int k_film_patch(const bf16*, const float*, const bf16*, int64_t, const float*, const float*, bf16*, bf16*, int, int, int, int,
                 cudaStream_t) {
  set_error("the fused FiLM patch-encoder kernel is tcgen05 only (not available in the host emulation)");
  return GG_ERR_ARCH;
}
int k_gemm_ln(const bf16*, int64_t, const bf16*, int64_t, int, const float*, const bf16*, const float*, const float*, bf16*, bf16*,
              float*, float*, int64_t, float, float, const uint64_t*, uint32_t, cudaStream_t) {
  set_error("the GEMM + LayerNorm kernel is tcgen05 only (not available in the host emulation)");
  return GG_ERR_ARCH;
}
int k_enc_ffn_bwd(const EncFfnBwdParams&, cudaStream_t) {
  set_error("the fused ffn-backward kernel is tcgen05 only (not available in the host emulation)");
  return GG_ERR_ARCH;
}
int k_enc_layer_fwd(const EncLayerParams&, cudaStream_t) {
  set_error("the fused encoder-layer kernel is tcgen05 only (not available in the host emulation)");
  return GG_ERR_ARCH;
}
}  // namespace gg
extern "C" int gg_encoder_layer_fwd(const gg_enc_layer_params*, void*) { return GG_ERR_ARCH; }
extern "C" int gg_enc_layer_set_trace(void*) { return GG_ERR_ARCH; }
extern "C" int gg_encoder_ffn_bwd(const gg_enc_ffn_bwd_params*, void*) { return GG_ERR_ARCH; }

#include "../../gemmgan_b200/csrc/engine.cu"

// the three entry points of abi.cu that are not already in emu.h
extern "C" int gg_abi_version(void) { return GG_ABI_VERSION; }
extern "C" void gg_launch_count_add(long long n) { gg::g_launch_count.fetch_add(n); }
extern "C" long long gg_launch_count(int reset) { return reset ? gg::g_launch_count.exchange(0) : gg::g_launch_count.load(); }
