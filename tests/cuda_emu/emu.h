#pragma once
// TEST INFRASTRUCTURE ONLY -- lets g++ compile a gemmgan_b200/csrc/*.cu file UNCHANGED and run its (tensor-core-free)
// kernels on the host, so that the build container (no GPU) can check index arithmetic, reductions, tails and argument
// handling against the oracle / torch, optionally under AddressSanitizer (tests/test_*_emulated.py). Usage: a
// translation unit includes this header and then the .cu file. It is not a CPU path of the product: nothing under
// gemmgan_b200/ can load the result, and it is orders of magnitude too slow to be one.
//
// Execution model: the threads of a CTA are fibers (hand-written switch on x86-64, else ucontext) scheduled round-robin on one host thread, the CTAs of a
// launch are spread over the host's cores. `__shared__` becomes `thread_local` (one copy per host thread = per
// CTA in flight), __syncthreads() a counting barrier that exited threads leave (as on the GPU), __shfl_down_sync an
// exchange through a per-warp slot array. What this cannot show: data races inside a CTA (fibers never run
// concurrently), performance, anything about the real launch.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

using std::max;
using std::min;

namespace emu {
// One CTA = up to 256 fibers (ucontext) scheduled round-robin on ONE host thread; CTAs of a launch are spread over the
// host's cores. A barrier is "count arrivals, yield until the generation changes"; threads that leave the kernel stop
// counting (as exited threads do on the GPU).
constexpr int MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = 64 * 1024;
struct Barrier {
  int live = 0, arrived = 0;
  long gen = 0;
};
struct Warp {
  Barrier bar;
  alignas(8) unsigned char slot[32][32];
};
struct Cta {
  Barrier bar;
  Warp warps[MAX_THREADS / 32];
};
// Context switch between fibers. glibc's swapcontext() makes an rt_sigprocmask system call per switch, and a warp
// shuffle is four switches per lane: on x86-64 the switch is done by hand instead (callee-saved registers, stack
// pointer, MXCSR / x87 control word — what the System V ABI lets a function call preserve). AddressSanitizer has to
// see stack switches through its swapcontext interceptor, so that mode (and every other architecture) keeps ucontext.
#if defined(__x86_64__) && !defined(__SANITIZE_ADDRESS__) && !defined(GG_EMU_UCONTEXT)
#define GG_EMU_FAST_SWITCH 1
struct Context {
  void* sp = nullptr;
};
__attribute__((naked, noinline)) static void switch_context(Context* /*from: rdi*/, Context* /*to: rsi*/) {
  asm volatile(
      "pushq %rbp\n\t"
      "pushq %rbx\n\t"
      "pushq %r12\n\t"
      "pushq %r13\n\t"
      "pushq %r14\n\t"
      "pushq %r15\n\t"
      "subq $8, %rsp\n\t"
      "stmxcsr (%rsp)\n\t"
      "fnstcw 4(%rsp)\n\t"
      "movq %rsp, (%rdi)\n\t"
      "movq (%rsi), %rsp\n\t"
      "ldmxcsr (%rsp)\n\t"
      "fldcw 4(%rsp)\n\t"
      "addq $8, %rsp\n\t"
      "popq %r15\n\t"
      "popq %r14\n\t"
      "popq %r13\n\t"
      "popq %r12\n\t"
      "popq %rbx\n\t"
      "popq %rbp\n\t"
      "ret\n\t");
}
// A fresh fiber: its first switch-in "returns" into `entry` with the stack aligned as after a call instruction.
inline void make_context(Context& c, char* stack, size_t bytes, void (*entry)()) {
  uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~static_cast<uintptr_t>(15);
  uint64_t* sp = reinterpret_cast<uint64_t*>(top);
  *--sp = 0;                                        // return address of `entry` (it never returns)
  *--sp = reinterpret_cast<uint64_t>(entry);        // popped by `ret`
  for (int i = 0; i < 6; ++i) *--sp = 0;            // rbp, rbx, r12 .. r15
  *--sp = 0x1F80ull | (0x037Full << 32);            // default MXCSR, default x87 control word
  c.sp = sp;
}
#else
struct Context {
  ucontext_t uc;
};
inline void switch_context(Context* from, Context* to) { swapcontext(&from->uc, &to->uc); }
inline void make_context(Context& c, char* stack, size_t bytes, void (*entry)()) {
  getcontext(&c.uc);
  c.uc.uc_stack.ss_sp = stack;
  c.uc.uc_stack.ss_size = bytes;
  c.uc.uc_link = nullptr;
  makecontext(&c.uc, entry, 0);
}
#endif
struct Fiber {
  Context ctx;
  bool done = false;
  uint3 tid;
};
struct Worker {  // per host thread
  Context sched;
  std::vector<Fiber> fibers;
  std::vector<char> stacks;
  int cur = 0;
  Cta cta;
  std::function<void()> body;
};
thread_local Worker* t_worker = nullptr;
thread_local uint3 t_thread, t_block;
thread_local dim3 t_block_dim, t_grid_dim;

inline void yield() {
  Worker* w = t_worker;
  switch_context(&w->fibers[w->cur].ctx, &w->sched);
}
inline void barrier_wait(Barrier& b) {
  const long g = b.gen;
  if (++b.arrived == b.live) {
    b.arrived = 0;
    ++b.gen;
    return;
  }
  while (b.gen == g) yield();
}
inline void barrier_leave(Barrier& b) {
  --b.live;
  if (b.live > 0 && b.arrived == b.live) {
    b.arrived = 0;
    ++b.gen;
  }
}
inline void fiber_entry() {
  Worker* w = t_worker;
  w->body();
  const int t = w->cur;
  barrier_leave(w->cta.warps[t / 32].bar);
  barrier_leave(w->cta.bar);
  w->fibers[t].done = true;
  switch_context(&w->fibers[t].ctx, &w->sched);  // never resumed
}

// Runs one CTA of `nthreads` threads on the calling host thread.
inline void run_cta(int nthreads, dim3 block, const std::function<void()>& body) {
  static thread_local Worker worker;
  Worker* w = &worker;
  t_worker = w;
  w->body = body;
  w->fibers.assign(nthreads, Fiber());
  w->stacks.resize(static_cast<size_t>(nthreads) * STACK_BYTES);
  w->cta = Cta();
  w->cta.bar.live = nthreads;
  for (int q = 0; q < (nthreads + 31) / 32; ++q) w->cta.warps[q].bar.live = std::min(32, nthreads - 32 * q);
  for (int t = 0; t < nthreads; ++t) {
    Fiber& f = w->fibers[t];
    f.tid = uint3{static_cast<unsigned>(t) % block.x, (static_cast<unsigned>(t) / block.x) % block.y,
                  static_cast<unsigned>(t) / (block.x * block.y)};
    make_context(f.ctx, w->stacks.data() + static_cast<size_t>(t) * STACK_BYTES, STACK_BYTES, fiber_entry);
  }
  for (int remaining = nthreads; remaining > 0;) {
    for (int t = 0; t < nthreads; ++t) {
      Fiber& f = w->fibers[t];
      if (f.done) continue;
      w->cur = t;
      t_thread = f.tid;
      switch_context(&w->sched, &f.ctx);
      if (f.done) --remaining;
    }
  }
}

template <class... KArgs>
cudaError_t launch(const cudaLaunchConfig_t* cfg, void (*kern)(KArgs...), KArgs... args) {
  const dim3 grid = cfg->gridDim, block = cfg->blockDim;
  const int nthreads = static_cast<int>(block.x * block.y * block.z);
  if (nthreads < 1 || nthreads > MAX_THREADS) return cudaErrorInvalidConfiguration;
  const long n_cta = static_cast<long>(grid.x) * grid.y * grid.z;
  std::atomic<long> next{0};
  auto host_thread = [&] {
    for (long c = next.fetch_add(1); c < n_cta; c = next.fetch_add(1)) {
      t_block = uint3{static_cast<unsigned>(c % grid.x), static_cast<unsigned>((c / grid.x) % grid.y),
                      static_cast<unsigned>(c / (static_cast<long>(grid.x) * grid.y))};
      t_block_dim = block;
      t_grid_dim = grid;
      run_cta(nthreads, block, [&] { kern(args...); });
    }
  };
  const int n_host = static_cast<int>(std::min<long>(n_cta, std::max(1u, std::thread::hardware_concurrency())));
  std::vector<std::thread> hosts;
  for (int i = 1; i < n_host; ++i) hosts.emplace_back(host_thread);
  host_thread();
  for (auto& h : hosts) h.join();
  return cudaSuccess;
}
}  // namespace emu

static inline void __syncthreads() { emu::barrier_wait(emu::t_worker->cta.bar); }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int offset) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu::Worker* w = emu::t_worker;
  const int t = w->cur, lane = t % 32;
  emu::Warp& warp = w->cta.warps[t / 32];
  memcpy(warp.slot[lane], &v, sizeof(T));
  emu::barrier_wait(warp.bar);
  T r = v;
  if (lane + offset < 32) memcpy(&r, warp.slot[lane + offset], sizeof(T));
  emu::barrier_wait(warp.bar);
  return r;
}
template <class T>
static inline T emu_shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu::Worker* w = emu::t_worker;
  const int t = w->cur, lane = t % 32;
  emu::Warp& warp = w->cta.warps[t / 32];
  memcpy(warp.slot[lane], &v, sizeof(T));
  emu::barrier_wait(warp.bar);
  T r = v;
  if (src_lane >= 0 && src_lane < 32) memcpy(&r, warp.slot[src_lane], sizeof(T));
  emu::barrier_wait(warp.bar);
  return r;
}
// every lane contributes `mine`; all lanes receive the 32 contributions (warp-collective instructions are built on this)
template <class T>
static inline void emu_warp_all_gather(const T& mine, T (&all)[32]) {
  static_assert(sizeof(T) <= 32, "gather payload");
  emu::Worker* w = emu::t_worker;
  const int t = w->cur, lane = t % 32;
  emu::Warp& warp = w->cta.warps[t / 32];
  memcpy(warp.slot[lane], &mine, sizeof(T));
  emu::barrier_wait(warp.bar);
  for (int l = 0; l < 32; ++l) memcpy(&all[l], warp.slot[l], sizeof(T));
  emu::barrier_wait(warp.bar);
}
static inline int emu_lane() { return emu::t_worker->cur % 32; }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  emu::Worker* w = emu::t_worker;
  emu::barrier_wait(w->cta.warps[w->cur / 32].bar);
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) { return emu_shfl_from(v, (emu::t_worker->cur % 32) ^ mask); }
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl_from(v, src & 31); }
template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline T __ldcg(const T* p) { return *p; }
template <class T>
static inline void __stcg(T* p, T v) { *p = v; }
template <class T>
static inline T __ldcs(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }
static inline long long clock64() { return 0; }
static inline void __trap() { abort(); }
template <class T>
static inline T atomicMin(T* p, T v) {
  T old = *p;
  while (v < old && !__atomic_compare_exchange(p, &old, &v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T>
static inline T atomicMax(T* p, T v) {
  T old = *p;
  while (old < v && !__atomic_compare_exchange(p, &old, &v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
// CUDA's fast-math intrinsics share their names with glibc-internal aliases that <math.h> declares but libm does not export
extern "C" __attribute__((weak)) float __expf(float x) noexcept { return expf(x); }
extern "C" __attribute__((weak)) float __logf(float x) noexcept { return logf(x); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* p, float v) {
  float old = *p, want;
  do want = old + v;
  while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return old;
}
static inline float __int_as_float(int v) {
  float f;
  memcpy(&f, &v, 4);
  return f;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

static inline cudaError_t emu_get_device(int* d) {
  *d = 0;
  return cudaSuccess;
}
// Streams and events: every emulated launch completes before it returns, so streams are opaque tokens, event record /
// wait are no-ops and the asynchronous copies are plain copies ("device" memory is host memory).
static inline cudaError_t emu_stream_create(cudaStream_t* s, unsigned = 0) {
  static int token;
  *s = reinterpret_cast<cudaStream_t>(&token);
  return cudaSuccess;
}
static inline cudaError_t emu_event_create(cudaEvent_t* e, unsigned = 0) {
  static int token;
  *e = reinterpret_cast<cudaEvent_t>(&token);
  return cudaSuccess;
}
static inline cudaError_t emu_ok(...) { return cudaSuccess; }
static inline cudaError_t emu_memcpy(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
  memmove(dst, src, n);
  return cudaSuccess;
}
static inline cudaError_t emu_memset(void* dst, int v, size_t n, cudaStream_t = nullptr) {
  memset(dst, v, n);
  return cudaSuccess;
}
#define cudaStreamCreateWithFlags emu_stream_create
#define cudaEventCreateWithFlags emu_event_create
#define cudaStreamDestroy emu_ok
#define cudaEventDestroy emu_ok
#define cudaEventRecord emu_ok
#define cudaStreamWaitEvent emu_ok
#define cudaStreamSynchronize emu_ok
#define cudaMemcpyAsync emu_memcpy
#define cudaMemsetAsync emu_memset
// runtime calls made by host_util.h's launch helpers / check macros and by the entry points
#define cudaLaunchKernelEx emu::launch
#define cudaGetLastError() cudaSuccess
#define cudaGetErrorString(e) "emulated"
#define cudaGetDevice emu_get_device

// ---- what abi.cu provides in the real library
#include "../../gemmgan_b200/csrc/host_util.h"
namespace gg {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};
bool pdl_enabled() { return true; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace gg
extern "C" const char* gg_last_error(void) { return gg::g_err; }
extern "C" int gg_check_device(int) { return GG_OK; }

// ---- rebind the CUDA spellings the kernel source uses, then compile it as it is
#undef __shared__
// block-scope `thread_local` is implicitly static: one copy per host thread = per CTA in flight. Dynamic shared memory
// (`extern __shared__ T name[];`) becomes a block-scope extern declaration of a thread_local array that the including
// translation unit defines at namespace scope, with the name and type the .cu uses and the largest size it launches.
#define __shared__ thread_local
#define cudaFuncSetAttribute(...) cudaSuccess
#define __launch_bounds__(...)
#define __noinline__
#define threadIdx emu::t_thread
#define blockIdx emu::t_block
#define blockDim emu::t_block_dim
#define gridDim emu::t_grid_dim
#define asm
#define volatile(...) ((void)0)  // pdl.cuh: asm volatile("griddepcontrol...") -> no-op
