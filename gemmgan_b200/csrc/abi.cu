// Error reporting and device probing for the C ABI (include/gemmgan.h).
#include "host_util.h"

#include <stdlib.h>
#include <string.h>

namespace gg {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};

bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("GEMMGAN_PDL");
    return !(v && v[0] == '0');
  }();
  return on;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace gg

extern "C" const char* gg_last_error(void) { return gg::g_err; }

extern "C" int gg_abi_version(void) { return GG_ABI_VERSION; }

// kernels replayed through a captured CUDA graph are added by the host runtime (it knows how many
// launches each captured graph holds)
extern "C" void gg_launch_count_add(long long n) { gg::g_launch_count.fetch_add(n); }

extern "C" long long gg_launch_count(int reset) {
  return reset ? gg::g_launch_count.exchange(0) : gg::g_launch_count.load();
}

extern "C" int gg_check_device(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    gg::set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    return GG_ERR_CUDA;
  }
  if (prop.major != 10) {
    gg::set_error("device %d is sm_%d%d; libgemmgan_sm100a needs compute capability 10.x (no fallback)",
                  dev, prop.major, prop.minor);
    return GG_ERR_ARCH;
  }
  return GG_OK;
}
