// Grouped weight-gradient GEMM: every dW_p[out_p, in_p] = dY_p^T X_p of one backward pass in ONE launch.
//
// The hand-written backward of the fusion tower produces ~25 weight gradients per step, each a small
// [out, in] <= [768, 512] output reduced over K = (replicas x batch x tokens) rows. Launched one by one
// (split-K over the machine + a reduce kernel each) they cost more in launch / prologue / tail than in
// math. Here they are one persistent kernel: a work item is (problem, 128x128 output tile, K-split); items
// are ordered longest-first and dealt round-robin to one CTA per SM. Same warp roles and tcgen05 / TMEM /
// TMA pipeline as gemm.cu (both operands MN-major: the forward tensors are read in place, no transposes).
// Problems whose K loop is much longer than the rest are split along K; the partial tiles of a split
// problem are summed in split order by the last CTA to arrive (deterministic, no float atomics).
// Bias gradients ride along: db[m] = sum_k dY[k, m] = (dY^T 1)[m]. For the first column tile of a problem with
// a bias output the MMA thread issues one more N = 16 MMA per K step against a constant all-ones B tile into a
// spare TMEM column block, and the epilogue reads it back as one value per row — the separate column-sum pass
// (one more read of every dY, 1.2 ms of kernel time per cfg3 train() call) disappears.
//
// Replaces autograd's `grad_output.t().mm(input)` for every nn.Linear / in_proj / out_proj of
// src/conditional_gan_cross_attention_with_film.py:108-123, 157-162 executed inside disc_loss.backward() /
// gen_loss.backward() (:412, :455).
#include "host_util.h"
#include "pdl.cuh"
#include "kernels.h"
#include "ptx.cuh"

#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

namespace gg {

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int STAGES = 5;
constexpr int ATOM_BYTES = 64 * BK * 2;
constexpr int A_TILE_BYTES = BM * BK * 2, B_TILE_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int SLOT_BYTES = 4096;
constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;
constexpr int STAGING_BYTES = EPI_WARPS * 2 * SLOT_BYTES;
constexpr int ONES_OFFSET = STAGING_OFFSET + STAGING_BYTES;  // all-ones B operand of the bias-gradient MMAs
constexpr int ONES_BYTES = 1024;
constexpr int BAR_OFFSET = ONES_OFFSET + ONES_BYTES;
constexpr int NUM_BARS = 2 * STAGES + 4;
constexpr int SMEM_BYTES = BAR_OFFSET + NUM_BARS * 8 + 32 + 1024;
static_assert(SMEM_BYTES <= 232448, "grouped wgrad kernel exceeds 227 KB of shared memory");
constexpr int BIAS_N = 16;                 // narrowest MMA for M = 128
constexpr int BIAS_COL0 = 2 * BN;          // TMEM columns [256, 320): two 32-column bias accumulator stages
constexpr int TMEM_COLS = 512;

struct Problem {
  int M, N, K;
  int work0;            // first work item of this problem
  int tiles_m, tiles_n, splits;
  int tma_out;
  float* out;
  int64_t ld;
  float* partial;       // [splits][M][N] fp32 when splits > 1
  unsigned* counters;   // [tiles_m * tiles_n] arrival counters (zero between launches)
  float* bias;          // [M] column sums of dY (bias gradient), or null
  float* bias_partial;  // [splits][M] when splits > 1
  int a3d, b3d;         // operand map is the 3-D [mn / 64][k][64] view: one TMA instruction per tile instead of two
};

struct Params {
  int nprob;
  int total_work;
  Problem p[WGRAD_GROUP_MAX];
  CUtensorMap mapA[WGRAD_GROUP_MAX], mapB[WGRAD_GROUP_MAX], mapO[WGRAD_GROUP_MAX];
};

struct Work {
  int pi, tm, tn, z, kb_begin, kb_end;
};

__device__ __forceinline__ Work decode(const Params& P, int w) {
  int pi = 0;
  while (pi + 1 < P.nprob && w >= P.p[pi + 1].work0) ++pi;
  const Problem& p = P.p[pi];
  const int local = w - p.work0;
  const int tiles = p.tiles_m * p.tiles_n;
  Work r;
  r.pi = pi;
  r.z = local / tiles;
  const int t = local % tiles;
  r.tm = t / p.tiles_n;
  r.tn = t % p.tiles_n;
  const int total_kb = (p.K + BK - 1) / BK;
  r.kb_begin = static_cast<int>(static_cast<int64_t>(total_kb) * r.z / p.splits);
  r.kb_end = static_cast<int>(static_cast<int64_t>(total_kb) * (r.z + 1) / p.splits);
  return r;
}

__device__ __forceinline__ void epi_bar() {  // the 8 epilogue warps only
  asm volatile("bar.sync 1, %0;\n" ::"n"(EPI_WARPS * 32) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) wgrad_group_kernel(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint32_t* last_flag = tmem_holder + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {  // constant all-ones operand (bf16 1.0 = 0x3F80), read by the tensor core through the async proxy
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + ONES_OFFSET);
    for (int i = threadIdx.x - 64; i < ONES_BYTES / 4; i += EPI_WARPS * 32) ones[i] = 0x3F803F80u;
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  // everything above (barriers, TMEM, descriptor prefetch) touched no tensor: it overlaps the previous kernel
  pdl_entry();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int w = blockIdx.x; w < P.total_work; w += gridDim.x) {
        const Work k = decode(P, w);
        const CUtensorMap* ma = &P.mapA[k.pi];
        const CUtensorMap* mb = &P.mapB[k.pi];
        const int m0 = k.tm * BM, n0 = k.tn * BN;
        for (int kb = k.kb_begin; kb < k.kb_end; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* a_dst = smem + s * STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          if (P.p[k.pi].a3d) {
            tma_load_3d(a_dst, ma, &full[s], 0, kb * BK, m0 / 64);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(a_dst + j * ATOM_BYTES, ma, &full[s], m0 + 64 * j, kb * BK);
          }
          if (P.p[k.pi].b3d) {
            tma_load_3d(b_dst, mb, &full[s], 0, kb * BK, n0 / 64);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * ATOM_BYTES, mb, &full[s], n0 + 64 * j, kb * BK);
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, BN, 1, 1);
      const uint32_t idesc_bias = make_idesc_bf16(BM, BIAS_N, 1, 0);  // A = dY tile (MN-major), B = ones (K-major)
      // ones tile without swizzle: 8-row x 16-byte core matrices 256 bytes apart in either direction (all inside
      // the 1 KB block, every element 1.0, so the exact walk is immaterial)
      const uint64_t ones_desc = make_smem_desc_noswizzle(smem_u32(smem + ONES_OFFSET), 256, 256);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int w = blockIdx.x; w < P.total_work; w += gridDim.x, ++it) {
        const Work k = decode(P, w);
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const bool with_bias = k.tn == 0 && P.p[k.pi].bias != nullptr;
        const uint32_t d_bias = tmem_base + BIAS_COL0 + acc * 32;
        for (int kb = k.kb_begin; kb < k.kb_end; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_base = a_base + A_TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t ad = make_smem_desc(a_base + kk * 2048, ATOM_BYTES, 1024);
            const uint64_t bd = make_smem_desc(b_base + kk * 2048, ATOM_BYTES, 1024);
            tc_mma_bf16(d_tmem, ad, bd, idesc, (kb > k.kb_begin || kk > 0) ? 1u : 0u);
            if (with_bias) tc_mma_bf16(d_bias, ad, ones_desc, idesc_bias, (kb > k.kb_begin || kk > 0) ? 1u : 0u);
          }
          tc_commit(&empty[s]);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        tc_commit(&tmem_full[acc]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;
    const int hsel = ew >> 2;
    uint8_t* slots = smem + STAGING_OFFSET + ew * (2 * SLOT_BYTES);
    const uint32_t lane_row = static_cast<uint32_t>(lane) * 128u;
    const uint32_t swz = static_cast<uint32_t>(lane & 7);
    int slot = 0;
    int it = 0;
    for (int w = blockIdx.x; w < P.total_work; w += gridDim.x, ++it) {
      const Work k = decode(P, w);
      const Problem& p = P.p[k.pi];
      const int acc = it & 1;
      const int row0 = k.tm * BM + q * 32;
      const int m = row0 + lane;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after_sync();
      float v[2][32];
#pragma unroll
      for (int c = 0; c < 2; ++c)
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + hsel * 64 + c * 32, v[c]);
      const bool with_bias = k.tn == 0 && p.bias != nullptr && hsel == 0;  // four warps cover the tile's 128 rows
      float bsum = 0.f;
      if (with_bias) {
        float vb[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + BIAS_COL0 + acc * 32, vb);
        tmem_ld_wait();
        bsum = vb[0];  // the BIAS_N columns are identical: column 0 is sum_k dY[k, m]
      }
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);

      bool write_out = true;
      if (p.splits > 1) {
        // publish this split's partial tile, then find out whether this CTA is the last one of the tile
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int n0 = k.tn * BN + hsel * 64 + c * 32;
          const int ncols = min(32, p.N - n0);
          if (m < p.M && ncols > 0) {
            float* dst = p.partial + (static_cast<int64_t>(k.z) * p.M + m) * p.N + n0;
            if (ncols == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                __stcg(reinterpret_cast<float4*>(dst) + j,
                       make_float4(v[c][4 * j], v[c][4 * j + 1], v[c][4 * j + 2], v[c][4 * j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) __stcg(dst + j, v[c][j]);
            }
          }
        }
        if (with_bias && m < p.M) __stcg(p.bias_partial + static_cast<int64_t>(k.z) * p.M + m, bsum);
        __threadfence();
        epi_bar();
        if (ew == 0 && lane == 0) {
          unsigned* ctr = p.counters + k.tm * p.tiles_n + k.tn;
          const unsigned prev = atomicAdd(ctr, 1u);
          const bool last = prev == static_cast<unsigned>(p.splits - 1);
          if (last) *ctr = 0;  // ready for the next launch (graph replays)
          *last_flag = last ? 1u : 0u;
          __threadfence();
        }
        epi_bar();
        write_out = *last_flag != 0;
        if (write_out && with_bias && m < p.M) {
          bsum = 0.f;
          for (int z = 0; z < p.splits; ++z) bsum += __ldcg(p.bias_partial + static_cast<int64_t>(z) * p.M + m);
        }
        if (write_out) {
          // sum the partials in split order (fixed order => bit-reproducible)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int n0 = k.tn * BN + hsel * 64 + c * 32;
            const int ncols = min(32, p.N - n0);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c][j] = 0.f;
            if (m < p.M && ncols > 0) {
              for (int z = 0; z < p.splits; ++z) {
                const float* src = p.partial + (static_cast<int64_t>(z) * p.M + m) * p.N + n0;
                if (ncols == 32 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 t = __ldcg(reinterpret_cast<const float4*>(src) + j);
                    v[c][4 * j] += t.x; v[c][4 * j + 1] += t.y; v[c][4 * j + 2] += t.z; v[c][4 * j + 3] += t.w;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (j < ncols) v[c][j] += __ldcg(src + j);
                }
              }
            }
          }
        }
        epi_bar();  // last_flag is reused by the next tile
      }
      if (!write_out) continue;
      if (with_bias && m < p.M) p.bias[m] = bsum;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int n0 = k.tn * BN + hsel * 64 + c * 32;
        const int ncols = min(32, p.N - n0);
        if (ncols <= 0) continue;
        if (p.tma_out) {
          slot ^= 1;
          uint8_t* fs = slots + slot * SLOT_BYTES;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(fs + lane_row + ((static_cast<uint32_t>(j) ^ swz) << 4)) =
                make_float4(v[c][4 * j], v[c][4 * j + 1], v[c][4 * j + 2], v[c][4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&P.mapO[k.pi], fs, n0, row0);
            tma_store_commit();
          }
        } else if (m < p.M) {
          float* dst = p.out + static_cast<int64_t>(m) * p.ld + n0;
          if (ncols == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(dst)[j] = make_float4(v[c][4 * j], v[c][4 * j + 1], v[c][4 * j + 2], v[c][4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) dst[j] = v[c][j];
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode2d(CUtensorMap* map, const void* ptr, bool f32, int64_t inner, int64_t outer, int64_t ld, int box_outer) {
  EncodeTiledFn fn = encode_fn();
  GG_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable");
  const int esz = f32 ? 4 : 2;
  GG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * esz) % 16 == 0 && ld >= inner,
             "grouped wgrad operand %p (ld %lld) is not 16-byte aligned", ptr, (long long)ld);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return GG_OK;
}

// MN-major operand [K rows, mn columns] (pitch ld) viewed as [mn / 64 blocks][K][64]: box = (64, 64, blocks_per_tile),
// i.e. the 64 x 64 atoms of one tile in one instruction, laid out ATOM_BYTES apart. Needs mn % 64 == 0.
int encode3d(CUtensorMap* map, const void* ptr, int64_t mn, int64_t k_rows, int64_t ld, int blocks_per_tile) {
  EncodeTiledFn fn = encode_fn();
  GG_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable");
  GG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0 && ld >= mn && mn % 64 == 0,
             "grouped wgrad operand %p (ld %lld) does not fit the 3-D view", ptr, (long long)ld);
  cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(k_rows), static_cast<cuuint64_t>(mn / 64)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, 128};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(blocks_per_tile)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
  return GG_OK;
}

}  // namespace

// Workspace layout: [GROUP_COUNTER_BYTES of arrival counters | partial tiles]. Only the counter region has to
// be zero when a launch starts, and every launch leaves it zero.
int64_t wgrad_group_workspace_bytes(int64_t max_output_elems) {
  // partial tiles + (generous) partial bias vectors: a bias has at most as many elements as its matrix has rows
  return GROUP_COUNTER_BYTES + static_cast<int64_t>(WGRAD_GROUP_MAX_SPLITS) * max_output_elems * 4 +
         static_cast<int64_t>(WGRAD_GROUP_MAX_SPLITS) * max_output_elems * 4 / 8 + 512LL * WGRAD_GROUP_MAX;
}

// The first GROUP_COUNTER_BYTES of `workspace` must be zero-initialised once.
int k_wgrad_group(const WgradItem* items, int n, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (n <= 0) return GG_OK;
  GG_REQUIRE(n <= WGRAD_GROUP_MAX, "too many grouped weight gradients (%d > %d)", n, WGRAD_GROUP_MAX);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  GG_CUDA_CHECK(attr_err);
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    GG_CUDA_CHECK(cudaGetDevice(&dev));
    GG_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // longest K loops first (static round-robin then approximates longest-processing-time-first)
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return items[a].K > items[b].K; });
  int64_t total_kb = 0;
  for (int i = 0; i < n; ++i) {
    const WgradItem& it = items[i];
    GG_REQUIRE(it.M > 0 && it.N > 0 && it.K > 0 && it.dy && it.x && it.out, "bad grouped wgrad item %d", i);
    total_kb += static_cast<int64_t>(ceil_div(it.M, BM)) * ceil_div(it.N, BN) * ceil_div(it.K, BK);
  }
  // a tile whose K loop exceeds the per-CTA share is split so that no single item dominates the makespan
  const int64_t share = std::max<int64_t>(8, (total_kb + num_sms - 1) / num_sms);
  static thread_local Params P;  // ~14 KB: keep it off the stack
  P.nprob = n;
  int work = 0;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int64_t ws_off = GROUP_COUNTER_BYTES;
  int64_t ctr_off = 0;
  for (int oi = 0; oi < n; ++oi) {
    const WgradItem& it = items[order[oi]];
    Problem& p = P.p[oi];
    p.M = it.M; p.N = it.N; p.K = it.K;
    p.tiles_m = ceil_div(it.M, BM);
    p.tiles_n = ceil_div(it.N, BN);
    const int kb = ceil_div(it.K, BK);
    int splits = static_cast<int>((kb + share - 1) / share);
    if (splits > WGRAD_GROUP_MAX_SPLITS) splits = WGRAD_GROUP_MAX_SPLITS;
    if (splits > kb) splits = kb;
    if (splits < 1) splits = 1;
    p.splits = splits;
    p.work0 = work;
    work += p.tiles_m * p.tiles_n * splits;
    p.out = it.out;
    p.ld = it.ld;
    p.tma_out = ((reinterpret_cast<uintptr_t>(it.out) & 15) == 0 && it.ld % 4 == 0) ? 1 : 0;
    p.partial = reinterpret_cast<float*>(ws + ws_off);
    if (splits > 1) ws_off = round_up64(ws_off + static_cast<int64_t>(splits) * it.M * it.N * 4, 256);
    p.bias = it.bias;
    p.bias_partial = reinterpret_cast<float*>(ws + ws_off);
    if (it.bias && splits > 1) ws_off = round_up64(ws_off + static_cast<int64_t>(splits) * it.M * 4, 256);
    p.counters = reinterpret_cast<unsigned*>(ws + ctr_off);
    ctr_off += static_cast<int64_t>(p.tiles_m) * p.tiles_n * 4;
    GG_REQUIRE(ws_off <= workspace_bytes && ctr_off <= GROUP_COUNTER_BYTES,
               "grouped wgrad workspace too small (%lld > %lld)", (long long)ws_off, (long long)workspace_bytes);
    static const bool use3d = [] { const char* v = getenv("GEMMGAN_WGRAD_3D"); return !(v && v[0] == '0'); }();
    p.a3d = use3d && it.M % 64 == 0;
    p.b3d = use3d && it.N % 64 == 0;
    if (p.a3d) GG_TRY_RC(encode3d(&P.mapA[oi], it.dy, it.M, it.K, it.ld_dy, BM / 64));
    else GG_TRY_RC(encode2d(&P.mapA[oi], it.dy, false, it.M, it.K, it.ld_dy, BK));
    if (p.b3d) GG_TRY_RC(encode3d(&P.mapB[oi], it.x, it.N, it.K, it.ld_x, BN / 64));
    else GG_TRY_RC(encode2d(&P.mapB[oi], it.x, false, it.N, it.K, it.ld_x, BK));
    if (p.tma_out) GG_TRY_RC(encode2d(&P.mapO[oi], it.out, true, it.N, it.M, it.ld, 32));
    else P.mapO[oi] = P.mapA[oi];
  }
  P.total_work = work;
  // The kernel runs on a side lane next to the dependent chain; GEMMGAN_WGRAD_SMS caps how many SMs it takes.
  static int sm_cap = [] { const char* v = getenv("GEMMGAN_WGRAD_SMS"); return v ? atoi(v) : 0; }();
  int slots = num_sms;
  if (sm_cap > 0 && sm_cap < slots) slots = sm_cap;
  const unsigned grid = static_cast<unsigned>(work < slots ? work : slots);
  {
    double fl = 0.0, by = 0.0;
    for (int i = 0; i < n; ++i) {
      const WgradItem& it = items[i];
      fl += 2.0 * it.M * it.N * it.K;
      by += 2.0 * (static_cast<double>(it.M) + it.N) * it.K + 4.0 * it.M * it.N;
    }
    GG_TRY_RC(prof_wgrad_begin(fl, by, st));
  }
  launch_k(wgrad_group_kernel, grid, THREADS, SMEM_BYTES, st, P);
  GG_LAUNCH_CHECK();
  GG_TRY_RC(prof_wgrad_end(st));
  return GG_OK;
}

}  // namespace gg
