// gg_gemm_bf16: the one GEMM of the training step.
//
//   D[M,N] = epilogue( sum over <=2 K-segments of A_seg * B_seg^T )
//
// sm_100a design: one 128 x BN output tile per CTA, 6 warps —
//   warp 0   TMA producer   (cp.async.bulk.tensor.2d into a STAGES-deep 128B-swizzled smem ring)
//   warp 1   MMA issuer     (one thread issues tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators
//                            live in BN columns of tensor memory; tcgen05.commit frees smem stages)
//   warps 2-5 epilogue      (tcgen05.ld 32x32b -> registers -> fused bias / activation / dropout /
//                            mask / residual -> vectorised global stores)
// Both operands can be K-major (row = M/N index) or MN-major (row = K index), so dgrad and wgrad
// read the forward tensors in place with no transpose pass. Split-K (grid.z) writes fp32 partials
// that a second kernel reduces in a fixed order (deterministic, no float atomics) and then runs
// the same epilogue.
//
// Replaces: nn.Linear forward and autograd's mm backward in
//   src/conditional_gan_cross_attention_with_film.py:56-72,129-162,197-231 (reference).
#include "epilogue.cuh"
#include "host_util.h"
#include "pdl.cuh"
#include "ptx.cuh"

#include <mutex>
#include <vector>

namespace gg {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int ATOM_BYTES = 64 * BK * 2;  // one MN-major 64x64 box

struct GemmArgs {
  int M, N;
  int K0, K1;
  int a_mn, b_mn;
  int splits;
  int tma_out_bf16, tma_out_f32;  // outputs written by TMA stores (alignment permitting)
  float* partial;
  gg_epilogue epi;
};

constexpr int SLOT_BYTES = 4096;   // one 32-row x 128-byte box of the output, 128B-swizzled

// EPIW epilogue warps (4 or 8). The 8-warp configurations own an SM (deep TMA ring, two staging slots per
// warp); the 4-warp "light" configuration fits twice on an SM (2-stage ring, one slot per warp, <= 32 K
// registers), so that for short-K tiles one CTA's load latency is covered by the other CTA's math.
template <int BN, int STAGES, int EPIW>
struct TileCfg {
  static constexpr int THREADS = 64 + EPIW * 32;
  static constexpr int CTAS_PER_SM = EPIW == 4 ? 2 : 1;
  static constexpr int SLOTS_PER_WARP = EPIW == 4 ? 1 : 2;
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  // epilogue: warp w drains TMEM lanes [32*(w%4), +32) (a hardware rule) and COLS_PER_WARP columns.
  // BN = 64 keeps only four epilogue warps busy.
  static constexpr int ACTIVE_EPI_WARPS = (BN == 64 || EPIW == 4) ? 4 : 8;
  static constexpr int COLS_PER_WARP = ACTIVE_EPI_WARPS == 4 ? BN : BN / 2;
  static constexpr int CHUNKS = COLS_PER_WARP / 32;
  static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int STAGING_BYTES = EPIW * SLOTS_PER_WARP * SLOT_BYTES;
  static constexpr int BAR_OFFSET = STAGING_OFFSET + STAGING_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 4;  // full/empty per stage + tmem full/empty x 2
  static constexpr int SMEM_BYTES = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 2 * BN;         // two accumulator stages
};

// Persistent: CTA c processes work items c, c + gridDim.x, ... where a work item is one
// (split, m-tile, n-tile) with the n-tile fastest (CTAs running side by side share the A rows in L2).
// The fp32 accumulator is double-buffered in tensor memory: while the epilogue warps drain stage a,
// the MMA warp already accumulates the next tile into stage a^1, and the TMA ring keeps running across
// tile boundaries.
// Epilogue: thread = accumulator row. Per 32-column chunk: tcgen05.ld -> fused math in registers ->
// 16-byte writes into a 128B-swizzled staging box -> one TMA store per box (bf16: 32 rows x 64 columns,
// fp32: 32 x 32; the hardware clips the M / N tails). Outputs whose pitch or base is not 16-byte
// aligned, row-remapped outputs and split-K partials take direct per-row vector stores instead.
template <int BN, int STAGES, int EPIW>
__global__ void __launch_bounds__(TileCfg<BN, STAGES, EPIW>::THREADS, TileCfg<BN, STAGES, EPIW>::CTAS_PER_SM)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                   const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                   const __grid_constant__ CUtensorMap tmOutB, const __grid_constant__ CUtensorMap tmOutF,
                   const GemmArgs g) {
  using Cfg = TileCfg<BN, STAGES, EPIW>;
  constexpr int SLOTS_PER_WARP = Cfg::SLOTS_PER_WARP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFFSET);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb0 = (g.K0 + BK - 1) / BK;
  const int kb1 = (g.K1 + BK - 1) / BK;
  const int total_kb = kb0 + kb1;
  const int tiles_n = (g.N + BN - 1) / BN;
  const int tiles_m = (g.M + BM - 1) / BM;
  const int64_t total_work = static_cast<int64_t>(tiles_n) * tiles_m * g.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (kb1 > 0) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    if (g.tma_out_bf16) tma_prefetch_desc(&tmOutB);
    if (g.tma_out_f32) tma_prefetch_desc(&tmOutF);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], Cfg::ACTIVE_EPI_WARPS);  // one arrival per active epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
    tmem_relinquish();
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
      for (int64_t w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int tn = static_cast<int>(w % tiles_n);
        const int tm = static_cast<int>((w / tiles_n) % tiles_m);
        const int z = static_cast<int>(w / (static_cast<int64_t>(tiles_n) * tiles_m));
        const int m0 = tm * BM, n0 = tn * BN;
        const int kb_begin = static_cast<int>(static_cast<int64_t>(total_kb) * z / g.splits);
        const int kb_end = static_cast<int>(static_cast<int64_t>(total_kb) * (z + 1) / g.splits);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          const bool second = kb >= kb0;
          const int kloc = (second ? kb - kb0 : kb) * BK;
          const CUtensorMap* ma = second ? &tmA1 : &tmA0;
          const CUtensorMap* mb = second ? &tmB1 : &tmB0;
          uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
          if (!g.a_mn) {
            tma_load_2d(a_dst, ma, &full[s], kloc, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(a_dst + j * ATOM_BYTES, ma, &full[s], m0 + 64 * j, kloc);
          }
          if (!g.b_mn) {
            tma_load_2d(b_dst, mb, &full[s], kloc, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(b_dst + j * ATOM_BYTES, mb, &full[s], n0 + 64 * j, kloc);
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
      const uint32_t idesc = make_idesc_bf16(BM, BN, g.a_mn, g.b_mn);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int64_t w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
        const int z = static_cast<int>(w / (static_cast<int64_t>(tiles_n) * tiles_m));
        const int kb_begin = static_cast<int>(static_cast<int64_t>(total_kb) * z / g.splits);
        const int kb_end = static_cast<int>(static_cast<int64_t>(total_kb) * (z + 1) / g.splits);
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);  // epilogue has drained this stage
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint32_t b_base = a_base + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 bf16 of K = 32 bytes inside the 128 B swizzle row.
            // MN-major: 16 k-rows of 128 B = 2048 bytes; 64-wide MN atoms are ATOM_BYTES apart.
            const uint64_t ad = g.a_mn ? make_smem_desc(a_base + k * 2048, ATOM_BYTES, 1024)
                                       : make_smem_desc(a_base + k * 32, 16, 1024);
            const uint64_t bd = g.b_mn ? make_smem_desc(b_base + k * 2048, ATOM_BYTES, 1024)
                                       : make_smem_desc(b_base + k * 32, 16, 1024);
            tc_mma_bf16(d_tmem, ad, bd, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
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
  } else if (warp - 2 < Cfg::ACTIVE_EPI_WARPS) {
    const int ew = warp - 2;   // 0..7
    const int q = warp & 3;    // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int hsel = Cfg::ACTIVE_EPI_WARPS == 4 ? 0 : (ew >> 2);  // which run of COLS_PER_WARP columns this warp drains
    uint8_t* slots = smem + Cfg::STAGING_OFFSET + ew * (SLOTS_PER_WARP * SLOT_BYTES);
    const uint32_t lane_row = static_cast<uint32_t>(lane) * 128u;
    const uint32_t swz = static_cast<uint32_t>(lane & 7);
    const bool tma_b = g.tma_out_bf16 != 0, tma_f = g.tma_out_f32 != 0;
    const bool dual = tma_b && tma_f && SLOTS_PER_WARP > 1;
    int slot = 0;
    uint8_t* bslot = slots;  // staging box of the bf16 output (spans two chunks)
    int it = 0;
    for (int64_t w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
      const int tn = static_cast<int>(w % tiles_n);
      const int tm = static_cast<int>((w / tiles_n) % tiles_m);
      const int z = static_cast<int>(w / (static_cast<int64_t>(tiles_n) * tiles_m));
      const int acc = it & 1;
      const int row0 = tm * BM + q * 32;
      const int m = row0 + lane;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int c = 0; c < Cfg::CHUNKS; ++c) {
        const int col0 = hsel * Cfg::COLS_PER_WARP + c * 32;
        const int n0 = tn * BN + col0;
        float v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + col0, v);
        tmem_ld_wait();
        if (c == Cfg::CHUNKS - 1) {
          // every tcgen05.ld of this warp has completed: hand the TMEM stage back to the MMA warp
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const int ncols = min(32, g.N - n0);  // <= 0: this chunk lies beyond N (warp-uniform)
        const bool live = m < g.M && ncols > 0;
        if (g.splits > 1) {
          if (live) {
            float* p = g.partial + (static_cast<int64_t>(z) * g.M + m) * g.N + n0;
            if (ncols == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) p[j] = v[j];
            }
          }
          continue;
        }
        if (live) epilogue_math<32>(g.epi, g.N, m, n0, ncols, v);
        if (live && ((g.epi.out_f32 && !tma_f) || (g.epi.out_bf16 && !tma_b)))
          epilogue_store<32>(g.epi, m, n0, ncols, v, !tma_f, !tma_b);
        if (tma_f && ncols > 0) {
          uint8_t* fs;
          if (dual) {  // both outputs: fixed slots (0 = the bf16 box that spans two chunks, 1 = fp32)
            fs = slots + SLOT_BYTES;
            if (lane == 0) tma_store_wait_read<0>();
          } else {
            slot = (slot + 1) % SLOTS_PER_WARP;
            fs = slots + slot * SLOT_BYTES;
            if (lane == 0) tma_store_wait_read<SLOTS_PER_WARP - 1>();
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(fs + lane_row + ((static_cast<uint32_t>(j) ^ swz) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOutF, fs, n0, row0);
            tma_store_commit();
          }
        }
        if (tma_b) {
          const int half = c & 1;
          if (half == 0 && ncols > 0) {
            if (dual) {
              bslot = slots;
              if (lane == 0) tma_store_wait_read<0>();
            } else {
              slot = (slot + 1) % SLOTS_PER_WARP;
              bslot = slots + slot * SLOT_BYTES;
              if (lane == 0) tma_store_wait_read<SLOTS_PER_WARP - 1>();
            }
            __syncwarp();
          }
          if (ncols > 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
              *reinterpret_cast<uint4*>(bslot + lane_row + ((static_cast<uint32_t>(half * 4 + j) ^ swz) << 4)) = pk;
            }
          }
          // the box is complete after its second chunk; a box whose first column lies beyond N is skipped
          if (half == 1 && n0 - 32 < g.N) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmOutB, bslot, n0 - 32, row0);
              tma_store_commit();
            }
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_read<0>();  // smem may go; the writes themselves drain before the grid completes
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// Sums split-K partials in split order (deterministic) and applies the epilogue. One thread per 4
// consecutive columns: a warp reads / writes 512 contiguous bytes per split.
__global__ void __launch_bounds__(256)
    splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N,
                         const gg_epilogue epi) {
  pdl_entry();
  const int chunks = (N + 3) / 4;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(M) * chunks) return;
  const int m = static_cast<int>(idx / chunks);
  const int nc0 = static_cast<int>(idx % chunks) * 4;
  const int ncols = min(4, N - nc0);
  const int64_t stride = static_cast<int64_t>(M) * N;
  const float* p = partial + static_cast<int64_t>(m) * N + nc0;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (ncols == 4 && ((reinterpret_cast<uintptr_t>(p) | (static_cast<uintptr_t>(stride) * 4)) & 15) == 0) {
    int z = 0;
    for (; z + 4 <= splits; z += 4) {  // 4 independent 16-byte loads in flight
      const float4 a = __ldcs(reinterpret_cast<const float4*>(p + (z + 0) * stride));
      const float4 b = __ldcs(reinterpret_cast<const float4*>(p + (z + 1) * stride));
      const float4 c = __ldcs(reinterpret_cast<const float4*>(p + (z + 2) * stride));
      const float4 d = __ldcs(reinterpret_cast<const float4*>(p + (z + 3) * stride));
      v[0] = (((v[0] + a.x) + b.x) + c.x) + d.x;
      v[1] = (((v[1] + a.y) + b.y) + c.y) + d.y;
      v[2] = (((v[2] + a.z) + b.z) + c.z) + d.z;
      v[3] = (((v[3] + a.w) + b.w) + c.w) + d.w;
    }
    for (; z < splits; ++z) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(p + z * stride));
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
    }
  } else {
    for (int z = 0; z < splits; ++z)
      for (int j = 0; j < 4; ++j)
        if (j < ncols) v[j] += p[z * stride + j];
  }
  epilogue_chunk<4>(epi, N, m, nc0, ncols, v);
}

// CUDA-core fp32 check path on the same bf16 operands and the same epilogue (tests only).
__global__ void __launch_bounds__(128)
    gemm_simt_kernel(const __nv_bfloat16* a0, const __nv_bfloat16* b0, int64_t lda0, int64_t ldb0,
                     int K0, const __nv_bfloat16* a1, const __nv_bfloat16* b1, int64_t lda1,
                     int64_t ldb1, int K1, int M, int N, int a_mn, int b_mn, const gg_epilogue epi) {
  pdl_entry();
  const int chunks = (N + 31) / 32;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(M) * chunks) return;
  const int m = static_cast<int>(idx / chunks);
  const int nc0 = static_cast<int>(idx % chunks) * 32;
  const int ncols = min(32, N - nc0);
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = 0.f;
  for (int seg = 0; seg < 2; ++seg) {
    const __nv_bfloat16* a = seg ? a1 : a0;
    const __nv_bfloat16* b = seg ? b1 : b0;
    const int64_t lda = seg ? lda1 : lda0, ldb = seg ? ldb1 : ldb0;
    const int K = seg ? K1 : K0;
    for (int k = 0; k < K; ++k) {
      const float av = __bfloat162float(a_mn ? a[static_cast<int64_t>(k) * lda + m]
                                             : a[static_cast<int64_t>(m) * lda + k]);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < ncols) {
          const int n = nc0 + j;
          const float bv = __bfloat162float(b_mn ? b[static_cast<int64_t>(k) * ldb + n]
                                                 : b[static_cast<int64_t>(n) * ldb + k]);
          v[j] = fmaf(av, bv, v[j]);
        }
      }
    }
  }
  epilogue_chunk<32>(epi, N, m, nc0, ncols, v);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D tensor [outer, inner] of bf16 (or fp32) with row pitch ld (elements); box = {box_inner, box_outer},
// 128B swizzle (box_inner * element size = 128 bytes).
static int encode_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld,
                      int box_outer, bool f32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (driver too old?)");
    return GG_ERR_CUDA;
  }
  const int esz = f32 ? 4 : 2;
  GG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand %p not 16-byte aligned",
             ptr);
  GG_REQUIRE((ld * esz) % 16 == 0 && ld >= inner, "GEMM operand ld=%lld must be a multiple of %d and >= %lld",
             (long long)ld, 16 / esz, (long long)inner);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld", (int)r,
              (long long)inner, (long long)outer, (long long)ld);
    return GG_ERR_CUDA;
  }
  return GG_OK;
}

// Optional live profiler: CUDA-event pairs around every tcgen05 GEMM launch (bench.py roofline leg).
struct GemmRecord {
  int M, N, K0, K1, bn, splits, a_mn, b_mn;
};
struct GemmProfile {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // start/stop pairs
  std::vector<GemmRecord> rec;  // one per pair
  size_t used = 0;
  double flops = 0.0;
  long long launches = 0;
};
static GemmProfile g_prof;

template <int BN, int STAGES, int EPIW>
static int launch_tc(const CUtensorMap* maps, const GemmArgs& args, cudaStream_t stream) {
  using Cfg = TileCfg<BN, STAGES, EPIW>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, EPIW>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  GG_CUDA_CHECK(attr_err);
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    GG_CUDA_CHECK(cudaGetDevice(&dev));
    GG_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t work = static_cast<int64_t>(ceil_div(args.N, BN)) * ceil_div(args.M, BM) * args.splits;
  const int64_t slots = static_cast<int64_t>(num_sms) * Cfg::CTAS_PER_SM;
  dim3 grid(static_cast<unsigned>(work < slots ? work : slots));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_prof.on) {
    if (g_prof.used + 2 > g_prof.ev.size()) {
      cudaEvent_t a, b;
      GG_CUDA_CHECK(cudaEventCreate(&a));
      GG_CUDA_CHECK(cudaEventCreate(&b));
      g_prof.ev.push_back(a);
      g_prof.ev.push_back(b);
    }
    e0 = g_prof.ev[g_prof.used];
    e1 = g_prof.ev[g_prof.used + 1];
    g_prof.used += 2;
    g_prof.flops += 2.0 * args.M * args.N * (static_cast<double>(args.K0) + args.K1);
    g_prof.launches += 1;
    g_prof.rec.push_back(GemmRecord{args.M, args.N, args.K0, args.K1, BN, args.splits, args.a_mn, args.b_mn});
    GG_CUDA_CHECK(cudaEventRecord(e0, stream));
  }
  launch_k(gemm_tc_kernel<BN, STAGES, EPIW>, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, maps[0], maps[1], maps[2],
                                                                           maps[3], maps[4], maps[5], args);
  GG_LAUNCH_CHECK();
  if (e1) GG_CUDA_CHECK(cudaEventRecord(e1, stream));
  return GG_OK;
}

int gemm_dispatch(const gg_gemm_desc* d, cudaStream_t stream) {
  GG_REQUIRE(d != nullptr, "null gemm desc");
  GG_REQUIRE(d->M > 0 && d->N > 0, "bad GEMM shape M=%d N=%d", d->M, d->N);
  GG_REQUIRE(d->nseg == 1 || d->nseg == 2, "nseg must be 1 or 2");
  for (int s = 0; s < d->nseg; ++s)
    GG_REQUIRE(d->seg[s].K > 0 && d->seg[s].a && d->seg[s].b, "bad GEMM segment %d (M=%d N=%d K=%d a=%p b=%p)", s,
               d->M, d->N, d->seg[s].K, d->seg[s].a, d->seg[s].b);
  GG_REQUIRE(d->epi.out_bf16 || d->epi.out_f32, "GEMM has no output");
  GG_REQUIRE(d->epi.drop_p == 0.f || d->epi.rng, "dropout needs an rng state pointer");

  if (d->impl == GG_IMPL_SIMT_F32) {
    const gg_gemm_seg& s0 = d->seg[0];
    const gg_gemm_seg& s1 = d->seg[1];
    const int chunks = ceil_div(d->N, 32);
    const int64_t work = static_cast<int64_t>(d->M) * chunks;
    launch_k(gemm_simt_kernel, static_cast<unsigned>((work + 127) / 128), 128, 0, stream, reinterpret_cast<const __nv_bfloat16*>(s0.a), reinterpret_cast<const __nv_bfloat16*>(s0.b),
        s0.lda, s0.ldb, s0.K, reinterpret_cast<const __nv_bfloat16*>(d->nseg > 1 ? s1.a : nullptr),
        reinterpret_cast<const __nv_bfloat16*>(d->nseg > 1 ? s1.b : nullptr),
        d->nseg > 1 ? s1.lda : 0, d->nseg > 1 ? s1.ldb : 0, d->nseg > 1 ? s1.K : 0, d->M, d->N,
        d->a_mn_major, d->b_mn_major, d->epi);
    GG_LAUNCH_CHECK();
    return GG_OK;
  }
  GG_REQUIRE(d->impl == GG_IMPL_TCGEN05, "unknown GEMM impl %d", d->impl);

  int bn = d->block_n;
  if (bn == 0) bn = d->N <= 64 ? 64 : 128;
  GG_REQUIRE(bn == 64 || bn == 128 || bn == 256, "block_n must be 64, 128 or 256");

  CUtensorMap maps[6];
  for (int s = 0; s < 2; ++s) {
    const gg_gemm_seg& sg = d->seg[s < d->nseg ? s : 0];
    int rc;
    if (!d->a_mn_major) rc = encode_map(&maps[2 * s], sg.a, sg.K, d->M, sg.lda, BM);
    else rc = encode_map(&maps[2 * s], sg.a, d->M, sg.K, sg.lda, BK);
    if (rc) return rc;
    if (!d->b_mn_major) rc = encode_map(&maps[2 * s + 1], sg.b, sg.K, d->N, sg.ldb, bn);
    else rc = encode_map(&maps[2 * s + 1], sg.b, d->N, sg.K, sg.ldb, BK);
    if (rc) return rc;
  }

  GemmArgs args;
  args.M = d->M;
  args.N = d->N;
  args.K0 = d->seg[0].K;
  args.K1 = d->nseg > 1 ? d->seg[1].K : 0;
  args.a_mn = d->a_mn_major;
  args.b_mn = d->b_mn_major;
  args.epi = d->epi;
  args.partial = reinterpret_cast<float*>(d->workspace);

  const int total_kb = ceil_div(args.K0, BK) + ceil_div(args.K1, BK);
  const int tiles = ceil_div(d->M, BM) * ceil_div(d->N, bn);
  int splits = 1;
  if (d->force_splits > 0) {
    splits = d->force_splits;
  } else if (d->workspace && tiles < 100 && total_kb >= 8) {
    splits = 296 / tiles;
    if (splits > total_kb / 2) splits = total_kb / 2;
    if (splits > 32) splits = 32;
  }
  if (splits > total_kb) splits = total_kb;
  if (splits < 1) splits = 1;
  if (splits > 1) {
    const int64_t need = static_cast<int64_t>(splits) * d->M * d->N * 4;
    if (!d->workspace || d->workspace_bytes < need) {
      if (d->force_splits > 0) {
        set_error("split-K workspace too small: need %lld bytes, have %lld", (long long)need,
                  (long long)d->workspace_bytes);
        return GG_ERR_WORKSPACE;
      }
      splits = static_cast<int>(d->workspace_bytes / (static_cast<int64_t>(d->M) * d->N * 4));
      if (splits < 2) splits = 1;
    }
  }
  args.splits = splits;

  // outputs go through TMA stores when their layout allows it (16-byte aligned base and pitch, identity
  // row map, plain overwrite); split-K runs its epilogue in the reduce kernel instead
  const gg_epilogue& ep = d->epi;
  args.tma_out_bf16 = splits == 1 && ep.out_bf16 && ep.row_div <= 0 && ep.ld_bf16 % 8 == 0 &&
                      (reinterpret_cast<uintptr_t>(ep.out_bf16) & 15) == 0;
  args.tma_out_f32 = splits == 1 && ep.out_f32 && ep.row_div <= 0 && !ep.accum_f32 && ep.ld_f32 % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(ep.out_f32) & 15) == 0;
  maps[4] = maps[0];
  maps[5] = maps[0];
  if (args.tma_out_bf16) {
    int rc = encode_map(&maps[4], ep.out_bf16, d->N, d->M, ep.ld_bf16, 32, false);
    if (rc) return rc;
  }
  if (args.tma_out_f32) {
    int rc = encode_map(&maps[5], ep.out_f32, d->N, d->M, ep.ld_f32, 32, true);
    if (rc) return rc;
  }

  // short-K, many-tile GEMMs (the tower's [rows, 256..512] x [N, K] products): two light CTAs per SM
  // (measured on B200: no faster than the one-CTA configuration on these shapes, so it is opt-in only)
  bool light = d->light > 0 && bn == 128 && splits == 1 && !(args.tma_out_bf16 && args.tma_out_f32);
  int rc;
  if (bn == 64) rc = launch_tc<64, 6, 8>(maps, args, stream);
  else if (bn == 128 && light) rc = launch_tc<128, 2, 4>(maps, args, stream);
  else if (bn == 128) rc = launch_tc<128, 4, 8>(maps, args, stream);
  else rc = launch_tc<256, 3, 8>(maps, args, stream);
  if (rc) return rc;

  if (splits > 1) {
    const int64_t work = static_cast<int64_t>(d->M) * ceil_div(d->N, 4);
    launch_k(splitk_reduce_kernel, static_cast<unsigned>((work + 255) / 256), 256, 0, stream, args.partial, splits, d->M, d->N, d->epi);
    GG_LAUNCH_CHECK();
  }
  return GG_OK;
}

}  // namespace gg

extern "C" int gg_gemm_profile_begin(void) {
  gg::g_prof.on = true;
  gg::g_prof.used = 0;
  gg::g_prof.rec.clear();
  gg::g_prof.flops = 0.0;
  gg::g_prof.launches = 0;
  return GG_OK;
}
// Synchronises the device and returns the summed duration (ms), FLOPs (2*M*N*K) and count of the
// tcgen05 GEMM launches issued since gg_gemm_profile_begin().
extern "C" int gg_gemm_profile_end(double* ms, double* flops, long long* launches) {
  using namespace gg;
  g_prof.on = false;
  GG_CUDA_CHECK(cudaDeviceSynchronize());
  double total = 0.0;
  for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
    float t = 0.f;
    GG_CUDA_CHECK(cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]));
    total += t;
  }
  if (ms) *ms = total;
  if (flops) *flops = g_prof.flops;
  if (launches) *launches = g_prof.launches;
  return GG_OK;
}

// Writes one CSV line per GEMM launch of the last profiled region (call after gg_gemm_profile_end).
extern "C" int gg_gemm_profile_dump(const char* path) {
  using namespace gg;
  GG_REQUIRE(path, "null path");
  FILE* f = fopen(path, "w");
  GG_REQUIRE(f, "cannot open %s", path);
  fprintf(f, "idx,M,N,K0,K1,block_n,splits,a_mn,b_mn,us,tflops\n");
  for (size_t i = 0; i + 1 < g_prof.used && i / 2 < g_prof.rec.size(); i += 2) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]) != cudaSuccess) t = 0.f;
    const GemmRecord& r = g_prof.rec[i / 2];
    const double fl = 2.0 * r.M * r.N * (static_cast<double>(r.K0) + r.K1);
    fprintf(f, "%zu,%d,%d,%d,%d,%d,%d,%d,%d,%.2f,%.1f\n", i / 2, r.M, r.N, r.K0, r.K1, r.bn, r.splits, r.a_mn,
            r.b_mn, t * 1e3, t > 0.f ? fl / (t * 1e-3) / 1e12 : 0.0);
  }
  fclose(f);
  return GG_OK;
}

extern "C" int gg_gemm_bf16(const gg_gemm_desc* desc, void* stream) {
  return gg::gemm_dispatch(desc, reinterpret_cast<cudaStream_t>(stream));
}
