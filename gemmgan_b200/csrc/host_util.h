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

#define GG_TRY_RC(x)        \
  do {                      \
    int _rc = (x);          \
    if (_rc) return _rc;    \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t round_up64(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

}  // namespace gg
