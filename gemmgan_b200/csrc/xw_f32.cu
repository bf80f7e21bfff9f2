// Critic layer 1 on FP32 gene profiles: out[m, 0:256] (fp32) = X[m, 0:K] (fp32, row-major) . W[0:256, 0:K]^T (bf16 shadow),
// for the two [B, G] tensors of WGAN_GP.gradient_penalty (real and fake, src/vanilla_gan_unconditional.py:304-327) read IN
// PLACE — the only O(B * G) work of the penalty (BASELINE config 5: B = 16384, G = 20000).
//
// What bounds it, and the layout that follows:
//  * HBM: the 8 B G bytes of the two fp32 tensors, once. A separate fp32 -> bf16 cast pass (what the step does for its
//    resident [fake; real] matrix) costs 1.5x that again; here the conversion happens on chip: TMA brings 256 x 32 fp32
//    boxes into a raw ring, eight "transform" warps (one thread per tile row) convert them to the bf16, 128B-swizzled,
//    K-major A operand tcgen05.mma reads.
//  * L2 -> SM fabric (~7 TB/s, below the HBM rate): every row tile needs the whole [256, K] weight matrix (10 MB at
//    G = 20000). With 128-row tiles that is as many bytes as X itself (measured: 724 us = 7.2 TB/s of fabric traffic). A CTA
//    therefore owns 256 rows — two M = 128 accumulators (all 512 TMEM columns) fed by ONE weight k-block — which halves it.
//    Sharing the k-block across a cluster by TMA multicast instead was built and measured slower (2 CTAs 937 us, 4 CTAs
//    1243 us against 724 us alone: every weight slot then waits for a cluster-wide release handshake).
//  * 148 SMs, 128 row tiles: the (tile, k-block) sequence is cut into 148 equal contiguous shares (stream-K). A share that
//    covers a whole tile writes it directly; the head and the tail of a share go to partial slots (at most two per CTA)
//    that a small kernel adds up, deterministically.
// Roles per CTA (10 warps): 0 TMA producer, 1 MMA issuer (+ TMEM owner), 2-9 transform, then epilogue (TMEM lane quarter =
// warp % 4, accumulator = (warp - 2) / 4).
#include "host_util.h"
#include "pdl.cuh"
#include "kernels.h"
#include "ptx.cuh"

#include <cuda.h>
#include <cuda_runtime.h>

#include <stdlib.h>

#include <mutex>

namespace gg {

int encode_tma_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer, bool f32);

namespace xw {

constexpr int BM = 256, BN = 256, BK = 64;
constexpr int RAW_STAGES = 3, A_STAGES = 2, B_STAGES = 2;
constexpr int RAW_BYTES = BM * 32 * 4;   // one 256 x 32 fp32 box (128-byte rows, 128B swizzle): half a k-block
constexpr int A_BYTES = BM * BK * 2;     // two 128-row sub-tiles
constexpr int B_BYTES = BN * BK * 2;
constexpr int OFF_RAW = 0;
constexpr int OFF_A = OFF_RAW + RAW_STAGES * RAW_BYTES;
constexpr int OFF_B = OFF_A + A_STAGES * A_BYTES;
constexpr int OFF_BAR = OFF_B + B_STAGES * B_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;  // + alignment slack
constexpr int TWARPS = 8;
constexpr int THREADS = (2 + TWARPS) * 32;
constexpr int TILE_FLOATS = BM * BN;

struct Args {
  int B, K;            // rows per tensor, reduction length
  int tiles_per_x;     // ceil(B / 256)
  int total_kb;        // ceil(K / 64)
  int64_t units;       // 2 * tiles_per_x * total_kb
  float* out;          // [2 * B, 256] fp32
  float* partial;      // [2 * gridDim.x][256][256] fp32: slot 2c = head (or only) partial segment of CTA c, 2c + 1 = tail
};

// unit range of worker c of n: [c * U / n, (c + 1) * U / n)
__host__ __device__ __forceinline__ int64_t share_begin(int64_t c, int64_t n, int64_t units) { return c * units / n; }
// the worker whose share contains unit x
__host__ __device__ __forceinline__ int owner_of(int64_t x, int64_t n, int64_t units) {
  int64_t c = x * n / units;
  while (c + 1 < n && share_begin(c + 1, n, units) <= x) ++c;
  while (c > 0 && share_begin(c, n, units) > x) --c;
  return static_cast<int>(c);
}

__global__ void __launch_bounds__(THREADS, 1)
    xw_f32_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                  const __grid_constant__ CUtensorMap tmW, const Args g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* raw_full = bars;                    // [RAW_STAGES]
  uint64_t* raw_empty = raw_full + RAW_STAGES;  // [RAW_STAGES]
  uint64_t* a_full = raw_empty + RAW_STAGES;    // [A_STAGES]
  uint64_t* a_empty = a_full + A_STAGES;        // [A_STAGES]
  uint64_t* b_full = a_empty + A_STAGES;        // [B_STAGES]
  uint64_t* b_empty = b_full + B_STAGES;        // [B_STAGES]
  uint64_t* acc_full = b_empty + B_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = g.total_kb;
  const int64_t u0 = share_begin(blockIdx.x, gridDim.x, g.units), u1 = share_begin(blockIdx.x + 1, gridDim.x, g.units);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < RAW_STAGES; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], TWARPS);
    }
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(&a_full[s], TWARPS);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, TWARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  pdl_entry();

  // segment of this CTA's share that starts at unit u: (tile, k-block range); every role walks the same segments
  auto segment = [&](int64_t u, int& tile, int& kb_begin, int& kb_end) {
    tile = static_cast<int>(u / T);
    kb_begin = static_cast<int>(u % T);
    const int64_t left = u1 - u;
    kb_end = left < T - kb_begin ? kb_begin + static_cast<int>(left) : T;
  };

  if (warp == 0) {
    if (lane == 0) {
      int rs = 0, bs = 0;
      uint32_t rph = 0, bph = 0;
      for (int64_t u = u0; u < u1;) {
        int tile, kb_begin, kb_end;
        segment(u, tile, kb_begin, kb_end);
        const CUtensorMap* mx = tile >= g.tiles_per_x ? &tmX1 : &tmX0;
        const int row0 = (tile % g.tiles_per_x) * BM;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&b_empty[bs], bph ^ 1);
          mbar_arrive_expect_tx(&b_full[bs], B_BYTES);
          tma_load_2d(smem + OFF_B + bs * B_BYTES, &tmW, &b_full[bs], kb * BK, 0);
          if (++bs == B_STAGES) { bs = 0; bph ^= 1; }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&raw_empty[rs], rph ^ 1);
            mbar_arrive_expect_tx(&raw_full[rs], RAW_BYTES);
            tma_load_2d(smem + OFF_RAW + rs * RAW_BYTES, mx, &raw_full[rs], kb * BK + 32 * h, row0);
            if (++rs == RAW_STAGES) { rs = 0; rph ^= 1; }
          }
        }
        u += kb_end - kb_begin;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int64_t u = u0; u < u1; ++it) {
        int tile, kb_begin, kb_end;
        segment(u, tile, kb_begin, kb_end);
        mbar_wait(acc_empty, (it & 1) ^ 1);  // the epilogue has drained both accumulators
        tc_fence_after_sync();
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&a_full[as], aph);
          mbar_wait(&b_full[bs], bph);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(smem + OFF_A + as * A_BYTES);
          const uint32_t b_base = smem_u32(smem + OFF_B + bs * B_BYTES);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub)
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_bf16(tmem_base + sub * BN, make_smem_desc(a_base + sub * (A_BYTES / 2) + k * 32, 16, 1024),
                          make_smem_desc(b_base + k * 32, 16, 1024), idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          tc_commit(&a_empty[as]);
          tc_commit(&b_empty[bs]);
          if (++as == A_STAGES) { as = 0; aph ^= 1; }
          if (++bs == B_STAGES) { bs = 0; bph ^= 1; }
        }
        tc_commit(acc_full);
        u += kb_end - kb_begin;
      }
    }
  } else {
    // ---- transform (fp32 raw rows -> bf16 K-major 128B-swizzled A operand), then the epilogue of the segment
    // tile row of this thread, in the transform and in the epilogue: accumulator (warp - 2) / 4, TMEM lane quarter warp % 4
    // (a warp may only read the 32 TMEM lanes of its own quarter)
    const int r = ((warp - 2) >> 2) * 128 + (warp & 3) * 32 + lane;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    int rs = 0, as = 0;
    uint32_t rph = 0, aph = 0;
    int it = 0;
    bool head_used = false;
    for (int64_t u = u0; u < u1; ++it) {
      int tile, kb_begin, kb_end;
      segment(u, tile, kb_begin, kb_end);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        uint8_t* dst = smem + OFF_A + as * A_BYTES + (r >> 7) * (A_BYTES / 2) + (r & 127) * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&raw_full[rs], rph);
          if (h == 0) mbar_wait(&a_empty[as], aph ^ 1);
          const uint8_t* src = smem + OFF_RAW + rs * RAW_BYTES + r * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {  // output 16-byte chunk 4 h + c = fp32 chunks 2c, 2c + 1 of this half
            const float4 lo = *reinterpret_cast<const float4*>(src + ((static_cast<uint32_t>(2 * c) ^ sw) << 4));
            const float4 hi = *reinterpret_cast<const float4*>(src + ((static_cast<uint32_t>(2 * c + 1) ^ sw) << 4));
            uint4 o;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(lo.x, lo.y), t1 = __floats2bfloat162_rn(lo.z, lo.w);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(hi.x, hi.y), t3 = __floats2bfloat162_rn(hi.z, hi.w);
            o.x = *reinterpret_cast<uint32_t*>(&t0);
            o.y = *reinterpret_cast<uint32_t*>(&t1);
            o.z = *reinterpret_cast<uint32_t*>(&t2);
            o.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(4 * h + c) ^ sw) << 4)) = o;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&raw_empty[rs]);
          if (++rs == RAW_STAGES) { rs = 0; rph ^= 1; }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
        if (++as == A_STAGES) { as = 0; aph ^= 1; }
      }
      // ---- epilogue: accumulator r / 128, TMEM lanes [32 (warp % 4), +32) -> tile row r
      mbar_wait(acc_full, it & 1);
      tc_fence_after_sync();
      const bool whole = kb_begin == 0 && kb_end == T;
      const int m = (tile % g.tiles_per_x) * BM + r;  // row inside its tensor
      float* orow;
      if (whole) {
        orow = g.out + (static_cast<int64_t>(tile >= g.tiles_per_x ? g.B : 0) + m) * BN;
      } else {  // head (first partial segment of the share) or tail
        orow = g.partial + (static_cast<int64_t>(2 * blockIdx.x + (head_used ? 1 : 0)) * BM + r) * BN;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (r >> 7) * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        if (!whole || m < g.B) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            reinterpret_cast<float4*>(orow + c * 32)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      if (!whole) head_used = true;
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      u += kb_end - kb_begin;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// Tiles that more than one share touched: out rows = sum of the partial slots of the contributing CTAs (in CTA order).
// One block per (tile, 8-row group).
__global__ void __launch_bounds__(256) xw_fixup_kernel(const Args g, int n_workers) {
  pdl_entry();
  const int tile = blockIdx.x / (BM / 8), rg = blockIdx.x % (BM / 8);
  const int T = g.total_kb;
  const int64_t t0 = static_cast<int64_t>(tile) * T, t1 = t0 + T;
  const int c_first = owner_of(t0, n_workers, g.units), c_last = owner_of(t1 - 1, n_workers, g.units);
  if (c_first == c_last) return;  // written whole by one CTA
  const int row = rg * 8 + (threadIdx.x >> 5);
  const int m = (tile % g.tiles_per_x) * BM + row;
  if (m >= g.B) return;
  const int lane = threadIdx.x & 31;
  float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  for (int c = c_first; c <= c_last; ++c) {
    // the tile is the first tile of c's share (head slot) unless c's share began in an earlier tile (then it is its tail)
    const int64_t cb = share_begin(c, n_workers, g.units);
    const bool head = cb >= t0;                       // share starts inside this tile
    // slot: 2c for the first partial segment of the share, 2c + 1 for the second. A share that starts exactly on a tile
    // boundary and covers that tile whole has no head; its only partial segment (the tail) then sits in slot 2c.
    int slot;
    if (head) {
      slot = 2 * c;
    } else {
      const int64_t first_tile_end = (cb / T + 1) * T;           // end of the tile the share starts in
      const bool has_head = (cb % T) != 0 || share_begin(c + 1, n_workers, g.units) < first_tile_end;
      slot = 2 * c + (has_head ? 1 : 0);
    }
    const float4* p = reinterpret_cast<const float4*>(g.partial + (static_cast<int64_t>(slot) * BM + row) * BN);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float4 v = p[lane + 32 * k];
      acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
    }
  }
  float4* o = reinterpret_cast<float4*>(g.out + (static_cast<int64_t>(tile >= g.tiles_per_x ? g.B : 0) + m) * BN);
  o[lane] = acc[0];
  o[lane + 32] = acc[1];
}

}  // namespace xw

int64_t xw_f32_workspace_bytes() { return 2LL * 148 * xw::TILE_FLOATS * 4; }

// out [2 * B, 256] fp32: rows [0, B) = x0 . w^T, rows [B, 2B) = x1 . w^T. x0, x1: fp32 [B, K] (row pitch K, 16-byte aligned,
// K % 4 == 0); w: bf16 [256, K] with pitch ldw. workspace: 2 partial tiles per CTA (xw_f32_workspace_bytes()).
int k_xw_f32(const float* x0, const float* x1, int B, int K, const bf16* w, int64_t ldw, float* out, void* workspace,
             int64_t workspace_bytes, cudaStream_t st) {
  GG_REQUIRE(x0 && x1 && w && out && workspace && B > 0 && K > 0, "bad xw_f32 argument");
  GG_REQUIRE(K % 4 == 0 && ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(x1)) & 15) == 0,
             "xw_f32 needs 16-byte aligned fp32 rows (K %% 4 == 0)");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int num_sms = 0;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(xw::xw_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, xw::SMEM_BYTES);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  });
  GG_CUDA_CHECK(attr_err);
  CUtensorMap mx0, mx1, mw;
  GG_TRY_RC(encode_tma_map(&mx0, x0, K, B, K, xw::BM, true));
  GG_TRY_RC(encode_tma_map(&mx1, x1, K, B, K, xw::BM, true));
  GG_TRY_RC(encode_tma_map(&mw, w, K, xw::BN, ldw, xw::BN, false));
  xw::Args a;
  a.B = B; a.K = K;
  a.tiles_per_x = (B + xw::BM - 1) / xw::BM;
  a.total_kb = (K + xw::BK - 1) / xw::BK;
  a.units = 2LL * a.tiles_per_x * a.total_kb;
  a.out = out;
  a.partial = reinterpret_cast<float*>(workspace);
  int64_t workers = num_sms;
  const int64_t by_ws = workspace_bytes / (2LL * xw::TILE_FLOATS * 4);
  if (workers > by_ws) workers = by_ws;
  if (workers > a.units) workers = a.units;
  GG_REQUIRE(workers >= 1, "xw_f32 workspace too small (%lld bytes)", (long long)workspace_bytes);
  launch_k(xw::xw_f32_kernel, static_cast<unsigned>(workers), xw::THREADS, xw::SMEM_BYTES, st, mx0, mx1, mw, a);
  GG_LAUNCH_CHECK();
  launch_k(xw::xw_fixup_kernel, static_cast<unsigned>(2 * a.tiles_per_x * (xw::BM / 8)), 256, 0, st, a,
           static_cast<int>(workers));
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg
