// Host build of gemmgan_b200/csrc/gemm.cu (see emu.h). The tcgen05 / TMA kernel of that file cannot run here (its PTX
// wrappers compile to nothing and it is never launched); what runs is the CUDA-core check kernel gemm_simt_kernel
// (GG_IMPL_SIMT_F32) and, with it, the epilogue code of epilogue.cuh that BOTH kernels share: bias, pre-activation add,
// LeakyReLU / FiLM activation, Philox dropout, mask scaling, residual, bf16 / fp32 outputs, accumulation, the output
// row remap, two K segments and both operand majors. Linked against libcudart only for the profiling hooks of gemm.cu
// (events), which are never called.
#include <cuda.h>

#include <mutex>
#include <vector>

#include "emu.h"

namespace gg {
thread_local uint8_t smem_raw[1024];  // `extern __shared__` of gemm_tc_kernel (declared, never run here)
}

#include "../../gemmgan_b200/csrc/gemm.cu"
