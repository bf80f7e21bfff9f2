// Host-side helpers shared by the translation units of libgemmgan_sm100a.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/gemmgan.h"

#include <atomic>

namespace gg {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launch_count;  // kernels launched by this library (all streams)

#define GG_CUDA_CHECK(expr)                                                                \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      gg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,      \
                    __LINE__);                                                             \
      return GG_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define GG_LAUNCH_CHECK()                                                                  \
  do {                                                                                     \
    gg::g_launch_count.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      gg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__,  \
                    __LINE__);                                                             \
      return GG_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define GG_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      gg::set_error(__VA_ARGS__);        \
      return GG_ERR_ARG;                 \
    }                                    \
  } while (0)

// Programmatic dependent launch (PDL): every kernel of this library starts with pdl_entry() (griddepcontrol.wait,
// then griddepcontrol.launch_dependents) and is launched with the programmatic-stream-serialization attribute,
// so the launch latency and the prologue of kernel n+1 (block scheduling, barrier init, TMEM allocation,
// tensor-map prefetch) overlap the body of kernel n instead of following its tail. Nothing is read or written
// before the wait, which orders a kernel after the complete predecessor (and, by induction, after everything
// earlier in the stream). GEMMGAN_PDL=0 launches without the attribute (the instructions are then no-ops).
bool pdl_enabled();

template <class... KArgs, class... Args>
static inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface through GG_LAUNCH_CHECK()
}

// Same, for kernels whose CTAs form clusters of `cluster` (1 = none) along x.
template <class... KArgs, class... Args>
static inline void launch_k_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface through GG_LAUNCH_CHECK()
}

#define GG_TRY_RC(x)        \
  do {                      \
    int _rc = (x);          \
    if (_rc) return _rc;    \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t round_up64(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

}  // namespace gg
