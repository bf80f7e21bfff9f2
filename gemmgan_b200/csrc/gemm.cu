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
  int lean;                       // >= 0: feature set of the compact epilogue (see lean_tile_epilogue); -1: generic
  int tf32;                       // operands are fp32 in memory, multiplied as TF32 (K-major only): a k-block is 32 elements
  long long* trace;               // diagnostics: per-role clock64() stamps of CTA `trace_cta` (gg_gemm_set_trace)
  int trace_cta;
  unsigned long long* stamp;      // measurement: {min start, max end} of this launch in %globaltimer ns, or null
  float* partial;
  gg_epilogue epi;
};

constexpr int TRACE_ROLE_STRIDE = 512;  // stamps per role: 0 producer (TMA issued), 1 MMA (stage full), 2 MMA (tile committed),
                                        // 3 epilogue warp 0 (accumulator full), 4 epilogue warp 0 (tile stored)
constexpr int SLOT_BYTES = 4096;   // one 32-row x 128-byte box of the output, 128B-swizzled

// EPIW epilogue warps (4 or 8). The 8-warp configurations own an SM (deep TMA ring, two staging slots per
// warp); the 4-warp "light" configuration fits twice on an SM (2-stage ring, one slot per warp, <= 32 K
// registers), so that for short-K tiles one CTA's load latency is covered by the other CTA's math.
// PAIR: the two CTAs of a cluster (the two SMs of a TPC) work as ONE 256 x BN tile with
// tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and only HALF of the B tile (BN/2 rows),
// the leader CTA issues the MMAs, and each CTA's tensor memory receives its 128 accumulator rows. Per output
// element this halves the operand bytes pulled from L2 into shared memory, which is what bounds these
// skinny (K or N = 256) products (measured: 128x128 tiles move ~8 TB/s of L2->SM traffic, the fabric's limit).
template <int BN, int STAGES, int EPIW, bool PAIR = false>
struct TileCfg {
  static constexpr int THREADS = 64 + EPIW * 32;
  static constexpr int CTAS_PER_SM = EPIW == 4 ? 2 : 1;
  static constexpr int SLOTS_PER_WARP = EPIW == 4 ? 1 : 2;
  static constexpr int BN_LOAD = PAIR ? BN / 2 : BN;  // B rows this CTA stages
  static constexpr int TILE_M = PAIR ? 2 * BM : BM;   // rows of one work item
  static constexpr int B_TILE_BYTES = BN_LOAD * BK * 2;
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
  static_assert(SMEM_BYTES <= 232448, "tile configuration exceeds 227 KB of shared memory");
  static_assert(!PAIR || (EPIW == 8 && BN % 32 == 0 && BN <= 256), "pair configuration: 8 epilogue warps, BN <= 256");
};

// The common epilogues as compact straight-line code: bf16 output through TMA-store boxes with a compile-time
// feature set F (fp32 bias, ReLU / LeakyReLU, dropout, bf16 mask, bf16 residual; 16-byte aligned rows, N a
// multiple of 8). The generic epilogue decides every feature per 32-column chunk at run time: ~330 issued
// instructions per chunk scattered over > 100 KB of code. A per-role clock64() trace (tests/gpu_gemm_trace.py)
// showed the epilogue warps stalled on instruction fetch (L0 I-cache: ~6 KB), 2500 cycles per 128 x 128 tile
// and 6000 for the first tile of a launch, against 1024 cycles of MMA work for K = 256: the tower's short-K
// products were bound by their epilogue, not by the tensor pipe or memory. Each instantiation below is a
// ~1.5 KB loop body; the global loads of a chunk are issued while its tcgen05.ld is in flight.
constexpr int LF_BIAS = 1, LF_ACT = 2, LF_DROP = 4, LF_MASK = 8, LF_RES = 16;

struct LeanArgs {  // passed by value (registers) into the out-of-line epilogue bodies
  const float* bias;
  const __nv_bfloat16* mask;
  const __nv_bfloat16* res;
  const uint64_t* rng;
  int64_t mask_ld, res_ld;
  float mask_pos, mask_neg, slope, drop_p;
  uint32_t site;
  int M, N;
};

template <int F, int CHUNKS, int COLS_PER_WARP, int SLOTS_PER_WARP, bool PAIR>
__device__ __noinline__ int lean_tile_epilogue(const LeanArgs e, const CUtensorMap* tmOutB, uint32_t tmem_acc,
                                               int row0, int n_tile0, int hsel, uint8_t* slots, int slot,
                                               uint64_t* tmem_empty_bar) {
  const LeanArgs& g = e;
  const int lane = threadIdx.x & 31;
  const int m = row0 + lane;
  const int mrow = m < g.M ? m : 0;  // rows beyond M compute on row 0's side inputs; their stores are clipped
  const float* __restrict__ bias = e.bias;
  const __nv_bfloat16* maskp = nullptr;
  const __nv_bfloat16* resp = nullptr;
  if (F & LF_MASK) maskp = e.mask + static_cast<int64_t>(mrow) * e.mask_ld;
  if (F & LF_RES) resp = e.res + static_cast<int64_t>(mrow) * e.res_ld;
  uint64_t seed = 0, step = 0;
  float keep_scale = 1.f;
  uint32_t thr = 0;
  if (F & LF_DROP) {
    seed = e.rng[0];
    step = e.rng[1];
    keep_scale = 1.0f / (1.0f - e.drop_p);
    thr = dropout_thr(e.drop_p);
  }
  const uint32_t lane_row = static_cast<uint32_t>(lane) * 128u;
  const uint32_t swz = static_cast<uint32_t>(lane & 7);
  uint8_t* bslot = slots;
#pragma unroll 1
  for (int c = 0; c < CHUNKS; ++c) {
    const int col0 = hsel * COLS_PER_WARP + c * 32;
    const int n0 = n_tile0 + col0;
    const int ncols = g.N - n0;  // >= 32: full chunk; <= 0: beyond N (warp-uniform); else a multiple of 8
    float v[32];
    tmem_ld_32x32(tmem_acc + col0, v);
    float4 bv[8];
    uint4 mk[4], rs[4];
    if (F & LF_BIAS) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        bv[j] = 4 * j < ncols ? __ldg(reinterpret_cast<const float4*>(bias + n0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (F & LF_MASK) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        mk[j] = 8 * j < ncols ? __ldg(reinterpret_cast<const uint4*>(maskp + n0) + j) : make_uint4(0, 0, 0, 0);
    }
    if (F & LF_RES) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        rs[j] = 8 * j < ncols ? __ldg(reinterpret_cast<const uint4*>(resp + n0) + j) : make_uint4(0, 0, 0, 0);
    }
    tmem_ld_wait();
    if (c == CHUNKS - 1) {
      // every tcgen05.ld of this warp has completed: hand the TMEM stage back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(tmem_empty_bar, 0);
        else mbar_arrive(tmem_empty_bar);
      }
    }
    if (ncols <= 0) continue;
    if (F & LF_BIAS) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[4 * j + 0] += bv[j].x; v[4 * j + 1] += bv[j].y; v[4 * j + 2] += bv[j].z; v[4 * j + 3] += bv[j].w;
      }
    }
    if (F & LF_ACT) {
      const float slope = e.slope;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f) + slope * fminf(v[j], 0.f);
    }
    if (F & LF_DROP) {
      const uint64_t base = (static_cast<uint64_t>(m) * static_cast<uint64_t>(g.N) + n0) >> 3;
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) {
        const uint32_t kb = keep_bits8(dropout_words(seed, step, e.site, base + gq), thr);
#pragma unroll
        for (int t = 0; t < 8; ++t) v[8 * gq + t] = ((kb >> t) & 1u) ? v[8 * gq + t] * keep_scale : 0.f;
      }
    }
    if (F & LF_MASK) {
      const float mp = e.mask_pos, mn = e.mask_neg;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&mk[j]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(h[t]);
          v[8 * j + 2 * t] *= f.x > 0.f ? mp : mn;
          v[8 * j + 2 * t + 1] *= f.y > 0.f ? mp : mn;
        }
      }
    }
    if (F & LF_RES) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rs[j]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(h[t]);
          v[8 * j + 2 * t] += f.x;
          v[8 * j + 2 * t + 1] += f.y;
        }
      }
    }
    // staging: one 32-row x 64-column bf16 box per two chunks
    const int half = c & 1;
    if (half == 0) {
      slot = (slot + 1) % SLOTS_PER_WARP;
      bslot = slots + slot * SLOT_BYTES;
      if (lane == 0) tma_store_wait_read<SLOTS_PER_WARP - 1>();
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 pk;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
      *reinterpret_cast<uint4*>(bslot + lane_row + ((static_cast<uint32_t>(half * 4 + j) ^ swz) << 4)) = pk;
    }
    if (half == 1 || ncols <= 32) {  // box complete (or its second half lies beyond N)
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmOutB, bslot, n0 - 32 * half, row0);
        tma_store_commit();
      }
    }
  }
  return slot;
}

// lean feature sets the engine's launches use (anything else takes the generic epilogue)
template <int CHUNKS, int COLS_PER_WARP, int SLOTS_PER_WARP, bool PAIR>
__device__ __forceinline__ int lean_dispatch(int F, const LeanArgs& g, const CUtensorMap* tmOutB, uint32_t tmem_acc,
                                             int row0, int n_tile0, int hsel, uint8_t* slots, int slot,
                                             uint64_t* tmem_empty_bar) {
#define GG_LEAN_CASE(FF)                                                                                        \
  case FF:                                                                                                      \
    return lean_tile_epilogue<FF, CHUNKS, COLS_PER_WARP, SLOTS_PER_WARP, PAIR>(g, tmOutB, tmem_acc, row0, n_tile0, \
                                                                               hsel, slots, slot, tmem_empty_bar);
  switch (F) {
    GG_LEAN_CASE(0)
    GG_LEAN_CASE(LF_BIAS)
    GG_LEAN_CASE(LF_ACT)
    GG_LEAN_CASE(LF_BIAS | LF_ACT)
    GG_LEAN_CASE(LF_ACT | LF_DROP)
    GG_LEAN_CASE(LF_BIAS | LF_ACT | LF_DROP)
    GG_LEAN_CASE(LF_MASK)
    GG_LEAN_CASE(LF_RES)
    GG_LEAN_CASE(LF_BIAS | LF_RES)
    default: return slot;
  }
#undef GG_LEAN_CASE
}
constexpr bool lean_supported(int F) {
  return F == 0 || F == LF_BIAS || F == LF_ACT || F == (LF_BIAS | LF_ACT) || F == (LF_ACT | LF_DROP) ||
         F == (LF_BIAS | LF_ACT | LF_DROP) || F == LF_MASK || F == LF_RES || F == (LF_BIAS | LF_RES);
}

// Persistent: CTA c processes work items c, c + gridDim.x, ... where a work item is one
// (split, m-tile, n-tile) with the n-tile fastest (CTAs running side by side share the A rows in L2).
// The fp32 accumulator is double-buffered in tensor memory: while the epilogue warps drain stage a,
// the MMA warp already accumulates the next tile into stage a^1, and the TMA ring keeps running across
// tile boundaries.
// Epilogue: thread = accumulator row. Per 32-column chunk: tcgen05.ld -> fused math in registers ->
// 16-byte writes into a 128B-swizzled staging box -> one TMA store per box (bf16: 32 rows x 64 columns,
// fp32: 32 x 32; the hardware clips the M / N tails). Outputs whose pitch or base is not 16-byte
// aligned, row-remapped outputs and split-K partials take direct per-row vector stores instead.
template <int BN, int STAGES, int EPIW, bool PAIR>
__global__ void __launch_bounds__(TileCfg<BN, STAGES, EPIW, PAIR>::THREADS, TileCfg<BN, STAGES, EPIW, PAIR>::CTAS_PER_SM)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                   const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                   const __grid_constant__ CUtensorMap tmOutB, const __grid_constant__ CUtensorMap tmOutF,
                   const GemmArgs g) {
  using Cfg = TileCfg<BN, STAGES, EPIW, PAIR>;
  constexpr int SLOTS_PER_WARP = Cfg::SLOTS_PER_WARP;
  constexpr int TILE_M = Cfg::TILE_M;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFFSET);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int BKe = g.tf32 ? BK / 2 : BK;  // elements per 128-byte k-block row
  const int kb0 = (g.K0 + BKe - 1) / BKe;
  const int kb1 = (g.K1 + BKe - 1) / BKe;
  const int total_kb = kb0 + kb1;
  const int tiles_n = (g.N + BN - 1) / BN;
  const int tiles_m = (g.M + TILE_M - 1) / TILE_M;
  const int64_t total_work = static_cast<int64_t>(tiles_n) * tiles_m * g.splits;
  // PAIR: both CTAs of a cluster walk the same work items; rank 1 owns rows [128, 256) of each tile
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int64_t worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int64_t nworkers = PAIR ? (gridDim.x >> 1) : gridDim.x;

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
      // one arrival per active epilogue warp (of both CTAs in a pair: the leader's MMA thread waits on it)
      mbar_init(&tmem_empty[a], (PAIR ? 2 : 1) * Cfg::ACTIVE_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_pair(tmem_holder, Cfg::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  long long* const tr = (g.trace && static_cast<int>(blockIdx.x) == g.trace_cta) ? g.trace : nullptr;
  if (tr && threadIdx.x == 0) tr[5 * TRACE_ROLE_STRIDE] = clock64();  // kernel-relative origin
  // everything above (barriers, TMEM, descriptor prefetch) touched no tensor: it overlaps the previous kernel
  pdl_entry();
  if (g.stamp && threadIdx.x == 0) atomicMin(g.stamp, globaltimer_ns());  // body start (the predecessor has finished)

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int tn_p = 0;
      for (int64_t w = worker; w < total_work; w += nworkers) {
        const int tn = static_cast<int>(w % tiles_n);
        const int tm = static_cast<int>((w / tiles_n) % tiles_m);
        const int z = static_cast<int>(w / (static_cast<int64_t>(tiles_n) * tiles_m));
        const int m0 = tm * TILE_M + static_cast<int>(rank) * BM;
        const int n0 = tn * BN + static_cast<int>(rank) * Cfg::BN_LOAD;  // PAIR: this CTA's half of the B tile
        const int kb_begin = static_cast<int>(static_cast<int64_t>(total_kb) * z / g.splits);
        const int kb_end = static_cast<int>(static_cast<int64_t>(total_kb) * (z + 1) / g.splits);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          if (tr && tn_p < TRACE_ROLE_STRIDE) tr[0 * TRACE_ROLE_STRIDE + tn_p++] = clock64();
          const bool second = kb >= kb0;
          const int kloc = (second ? kb - kb0 : kb) * BKe;
          const CUtensorMap* ma = second ? &tmA1 : &tmA0;
          const CUtensorMap* mb = second ? &tmB1 : &tmB0;
          uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_TILE_BYTES;
          // PAIR: the leader's barrier counts the bytes of both CTAs' loads of this stage
          if (!PAIR) mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
          else if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
          auto load = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
            if (PAIR) tma_load_2d_pair(dst, m, &full[s], c0, c1);
            else tma_load_2d(dst, m, &full[s], c0, c1);
          };
          if (!g.a_mn) {
            load(a_dst, ma, kloc, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) load(a_dst + j * ATOM_BYTES, ma, m0 + 64 * j, kloc);
          }
          if (!g.b_mn) {
            load(b_dst, mb, kloc, n0);
          } else {
#pragma unroll
            for (int j = 0; j < Cfg::BN_LOAD / 64; ++j) load(b_dst + j * ATOM_BYTES, mb, n0 + 64 * j, kloc);
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // PAIR: the leader CTA issues the MMAs of both
      const uint32_t idesc = g.tf32 ? make_idesc_tf32(TILE_M, BN) : make_idesc_bf16(TILE_M, BN, g.a_mn, g.b_mn);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      int tn_f = 0;
      for (int64_t w = worker; w < total_work; w += nworkers, ++it) {
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
          if (tr && tn_f < TRACE_ROLE_STRIDE) tr[1 * TRACE_ROLE_STRIDE + tn_f++] = clock64();
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
            if (PAIR) tc_mma_bf16_pair(d_tmem, ad, bd, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            else if (g.tf32) tc_mma_tf32(d_tmem, ad, bd, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            else tc_mma_bf16(d_tmem, ad, bd, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          if (PAIR) tc_commit_pair(&empty[s]);  // frees the stage in both CTAs
          else tc_commit(&empty[s]);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        if (PAIR) tc_commit_pair(&tmem_full[acc]);
        else tc_commit(&tmem_full[acc]);
        if (tr && it < TRACE_ROLE_STRIDE) tr[2 * TRACE_ROLE_STRIDE + it] = clock64();
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
    LeanArgs la;
    la.bias = g.epi.bias;
    la.mask = reinterpret_cast<const __nv_bfloat16*>(g.epi.mask);
    la.res = reinterpret_cast<const __nv_bfloat16*>(g.epi.res);
    la.rng = g.epi.rng;
    la.mask_ld = g.epi.mask_ld; la.res_ld = g.epi.res_ld;
    la.mask_pos = g.epi.mask_pos; la.mask_neg = g.epi.mask_neg; la.slope = g.epi.slope; la.drop_p = g.epi.drop_p;
    la.site = g.epi.site;
    la.M = g.M; la.N = g.N;
    for (int64_t w = worker; w < total_work; w += nworkers, ++it) {
      const int tn = static_cast<int>(w % tiles_n);
      const int tm = static_cast<int>((w / tiles_n) % tiles_m);
      const int z = static_cast<int>(w / (static_cast<int64_t>(tiles_n) * tiles_m));
      const int acc = it & 1;
      const int row0 = tm * TILE_M + static_cast<int>(rank) * BM + q * 32;
      const int m = row0 + lane;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after_sync();
      if (tr && ew == 0 && lane == 0 && it < TRACE_ROLE_STRIDE) tr[3 * TRACE_ROLE_STRIDE + it] = clock64();
      if (g.lean >= 0) {
        slot = lean_dispatch<Cfg::CHUNKS, Cfg::COLS_PER_WARP, SLOTS_PER_WARP, PAIR>(
            g.lean, la, &tmOutB, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN, row0, tn * BN, hsel, slots,
            slot, &tmem_empty[acc]);
        if (tr && ew == 0 && lane == 0 && it < TRACE_ROLE_STRIDE) tr[4 * TRACE_ROLE_STRIDE + it] = clock64();
        continue;
      }
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
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(&tmem_empty[acc], 0);  // the leader's MMA thread owns both halves
            else mbar_arrive(&tmem_empty[acc]);
          }
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
  if (g.stamp && threadIdx.x == 64) {  // first epilogue thread: its stores have been issued
    tma_store_wait_all<0>();
    atomicMax(g.stamp + 1, globaltimer_ns());
  }
  if (PAIR) {
    cluster_sync_all();  // neither CTA may leave (or free tensor memory) while the other still signals / reads it
    if (warp == 1) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
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
int encode_tma_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer,
                   bool f32);
static int encode_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld,
                      int box_outer, bool f32 = false) {
  return encode_tma_map(map, ptr, inner, outer, ld, box_outer, f32);
}
int encode_tma_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer,
                   bool f32) {
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
  int out_bytes;  // bytes written per output element (2 = bf16, 4 = fp32, 6 = both)
};
struct GemmProfile {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // start/stop pairs
  std::vector<GemmRecord> rec;  // one per pair
  size_t used = 0;
  double flops = 0.0;
  double bytes = 0.0;  // algorithmic: every operand read once, every output written once
  long long launches = 0;
};
static GemmProfile g_prof;
// second record set for the fused encoder-layer kernel (same begin / end switch): bench.py reports both kernels
struct AuxProfile {
  std::vector<cudaEvent_t> ev;
  size_t used = 0;
  double flops = 0.0, bytes = 0.0;
  long long launches = 0;
};
static AuxProfile g_auxes[2];  // 0: fused encoder-layer kernels (enc_layer.cu), 1: grouped weight-gradient kernel
#define g_aux g_auxes[0]
struct StampRecord {
  double flops, bytes;
};
static std::vector<StampRecord> g_stamp_rec;
static unsigned long long* g_stamp_buf = nullptr;  // device: 2 x capacity uint64 (gg_gemm_set_timer)
static int g_stamp_cap = 0, g_stamp_next = 0;
static long long* g_trace_buf = nullptr;  // device buffer of 6 * TRACE_ROLE_STRIDE stamps, or null
static int g_trace_cta = 0;

// Event record that also works while the stream is being captured into a CUDA graph (an external event-record
// node): the profiled train() call can then be replayed as a graph, so the event pairs bracket the kernels as
// they run in the real step (no host launch gaps between them).
static cudaError_t prof_record(cudaEvent_t ev, cudaStream_t stream) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaError_t rc = cudaStreamIsCapturing(stream, &cs);
  if (rc != cudaSuccess) return rc;
  return cs == cudaStreamCaptureStatusActive ? cudaEventRecordWithFlags(ev, stream, cudaEventRecordExternal)
                                             : cudaEventRecord(ev, stream);
}

// Event pair around a launch of another tensor-core kernel of the library while the live profiler is on:
// prof_aux_begin before the launch (records its algorithmic flops / bytes), prof_aux_end after it.
static int prof_chan_begin(AuxProfile& a, double flops, double bytes, cudaStream_t stream) {
  if (!g_prof.on) return GG_OK;
  if (a.used + 2 > a.ev.size()) {
    cudaEvent_t e0, e1;
    GG_CUDA_CHECK(cudaEventCreate(&e0));
    GG_CUDA_CHECK(cudaEventCreate(&e1));
    a.ev.push_back(e0);
    a.ev.push_back(e1);
  }
  a.flops += flops;
  a.bytes += bytes;
  a.launches += 1;
  GG_CUDA_CHECK(prof_record(a.ev[a.used], stream));
  return GG_OK;
}
static int prof_chan_end(AuxProfile& a, cudaStream_t stream) {
  if (!g_prof.on) return GG_OK;
  GG_CUDA_CHECK(prof_record(a.ev[a.used + 1], stream));
  a.used += 2;
  return GG_OK;
}
int prof_aux_begin(double flops, double bytes, cudaStream_t stream) { return prof_chan_begin(g_auxes[0], flops, bytes, stream); }
int prof_aux_end(cudaStream_t stream) { return prof_chan_end(g_auxes[0], stream); }
int prof_wgrad_begin(double flops, double bytes, cudaStream_t stream) { return prof_chan_begin(g_auxes[1], flops, bytes, stream); }
int prof_wgrad_end(cudaStream_t stream) { return prof_chan_end(g_auxes[1], stream); }

template <int BN, int STAGES, int EPIW, bool PAIR = false>
static int launch_tc(const CUtensorMap* maps, const GemmArgs& args, cudaStream_t stream) {
  using Cfg = TileCfg<BN, STAGES, EPIW, PAIR>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, EPIW, PAIR>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  GG_CUDA_CHECK(attr_err);
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    GG_CUDA_CHECK(cudaGetDevice(&dev));
    GG_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t work = static_cast<int64_t>(ceil_div(args.N, BN)) * ceil_div(args.M, Cfg::TILE_M) * args.splits;
  // workers: CTAs, or CTA pairs (one per TPC)
  const int64_t slots = PAIR ? num_sms / 2 : static_cast<int64_t>(num_sms) * Cfg::CTAS_PER_SM;
  const int64_t workers = work < slots ? work : slots;
  dim3 grid(static_cast<unsigned>(PAIR ? 2 * workers : workers));
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
    const int ob = (args.epi.out_bf16 ? 2 : 0) + (args.epi.out_f32 ? 4 : 0);
    g_prof.bytes += (args.tf32 ? 4.0 : 2.0) * (static_cast<double>(args.M) + args.N) * (static_cast<double>(args.K0) + args.K1) +
                    static_cast<double>(ob) * args.M * args.N;
    g_prof.rec.push_back(GemmRecord{args.M, args.N, args.K0, args.K1, PAIR ? -BN : BN, args.splits, args.a_mn, args.b_mn, ob});
    GG_CUDA_CHECK(prof_record(e0, stream));
  }
  launch_k_cluster(gemm_tc_kernel<BN, STAGES, EPIW, PAIR>, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, PAIR ? 2 : 1,
                   maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], args);
  GG_LAUNCH_CHECK();
  if (e1) GG_CUDA_CHECK(prof_record(e1, stream));
  return GG_OK;
}

static bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GEMMGAN_PAIR");
    v = !(e && e[0] == '0');
  }
  return v != 0;
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
    GG_REQUIRE(!d->tf32_operands, "the CUDA-core check path takes bf16 operands");
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
  const bool tf32 = d->tf32_operands != 0;
  const int BKh = tf32 ? BK / 2 : BK;
  GG_REQUIRE(!tf32 || (!d->a_mn_major && !d->b_mn_major), "fp32 (TF32) operands must be K-major");

  int bn = d->block_n;
  // Tile configuration (measured on B200, tests/gpu_gemm_bench.py -> profiles/r01_gemm_shapes_v6.log). A stage of
  // the TMA ring costs its issuing thread ~550 cycles (2 boxes), a 128 x 128 x 64 block only 256 MMA cycles, so
  // short-K products want the most output per loaded byte that still fills the SMs:
  //  * long K (>= 2048) and N >= 256: CTA pairs (256 x 256 tiles, half of B per CTA; split-K fills the TPCs);
  //  * N >= 512 with enough 128 x 256 tiles for every SM: 256-wide single-CTA tiles;
  //  * otherwise 128-wide tiles (more, smaller tiles for the short / narrow problems).
  bool pair = false;
  const int kb_all = ceil_div(d->seg[0].K, BKh) + (d->nseg > 1 ? ceil_div(d->seg[1].K, BKh) : 0);
  if (d->pair > 0 && !tf32) pair = true;
  else if (d->pair == 0 && !tf32 && pair_enabled() && (bn == 0 || bn == 256) && d->N >= 256 && d->M >= 256 && kb_all >= 32) {
    const int64_t tiles2 = static_cast<int64_t>(ceil_div(d->M, 256)) * ceil_div(d->N, 256);
    pair = tiles2 >= 56 || (d->workspace != nullptr && tiles2 * (kb_all / 4) >= 56);
  }
  if (!pair && bn == 0 && d->N >= 512 && static_cast<int64_t>(ceil_div(d->M, BM)) * ceil_div(d->N, 256) >= 148)
    bn = 256;
  if (pair) bn = 256;
  if (bn == 0) bn = d->N <= 64 ? 64 : 128;
  GG_REQUIRE(bn == 64 || bn == 128 || bn == 256, "block_n must be 64, 128 or 256");
  GG_REQUIRE(!pair || bn == 256, "the CTA-pair configuration uses 256-wide tiles");

  CUtensorMap maps[6];
  for (int s = 0; s < 2; ++s) {
    const gg_gemm_seg& sg = d->seg[s < d->nseg ? s : 0];
    int rc;
    if (!d->a_mn_major) rc = encode_map(&maps[2 * s], sg.a, sg.K, d->M, sg.lda, BM, tf32);
    else rc = encode_map(&maps[2 * s], sg.a, d->M, sg.K, sg.lda, BK);
    if (rc) return rc;
    if (!d->b_mn_major) rc = encode_map(&maps[2 * s + 1], sg.b, sg.K, d->N, sg.ldb, pair ? bn / 2 : bn, tf32);
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
  args.tf32 = tf32 ? 1 : 0;
  args.partial = reinterpret_cast<float*>(d->workspace);
  args.trace = g_trace_buf;
  args.trace_cta = g_trace_cta;
  args.stamp = nullptr;
  if (g_stamp_buf && g_stamp_next < g_stamp_cap) {
    args.stamp = g_stamp_buf + 2 * g_stamp_next;
    g_stamp_rec.push_back(StampRecord{2.0 * d->M * d->N * (static_cast<double>(d->seg[0].K) + (d->nseg > 1 ? d->seg[1].K : 0)),
                                      2.0 * (static_cast<double>(d->M) + d->N) * (static_cast<double>(d->seg[0].K) + (d->nseg > 1 ? d->seg[1].K : 0)) +
                                          static_cast<double>((d->epi.out_bf16 ? 2 : 0) + (d->epi.out_f32 ? 4 : 0)) * d->M * d->N});
    ++g_stamp_next;
  }

  const int total_kb = ceil_div(args.K0, BKh) + ceil_div(args.K1, BKh);
  const int tiles = ceil_div(d->M, pair ? 2 * BM : BM) * ceil_div(d->N, bn);
  int splits = 1;
  if (d->force_splits > 0) {
    splits = d->force_splits;
  } else if (pair) {
    if (d->workspace && tiles < 56 && total_kb >= 8) {  // one wave over the 74 TPCs
      splits = 74 / tiles;
      if (splits > total_kb / 4) splits = total_kb / 4;
      if (splits > 32) splits = 32;
    }
  } else if (d->workspace && tiles < 100 && total_kb >= 32) {
    // (a split costs a second launch — ~6 us on a dependent chain — so K <= 1984 stays in one pass: the short
    // [B, 256] x K <= 1024 products of the towers are latency-, not throughput-bound)
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
  {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    static const bool lean_on = [] { const char* v = getenv("GEMMGAN_LEAN_EPILOGUE"); return !(v && v[0] == '0'); }();
    const bool ok = lean_on && splits == 1 && args.tma_out_bf16 && !ep.out_f32 && !ep.pre && ep.alpha == 1.0f &&
                    (ep.act == GG_ACT_NONE || ep.act == GG_ACT_LEAKY) && d->N % 8 == 0 && (!ep.bias || al16(ep.bias)) &&
                    (!ep.mask || (!ep.mask_f32 && al16(ep.mask) && ep.mask_ld % 8 == 0)) &&
                    (!ep.res || (!ep.res_f32 && al16(ep.res) && ep.res_ld % 8 == 0));
    const int F = (ep.bias ? LF_BIAS : 0) | (ep.act == GG_ACT_LEAKY ? LF_ACT : 0) | (ep.drop_p > 0.f ? LF_DROP : 0) |
                  (ep.mask ? LF_MASK : 0) | (ep.res ? LF_RES : 0);
    args.lean = ok && lean_supported(F) ? F : -1;
  }
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
  if (pair) rc = launch_tc<256, 5, 8, true>(maps, args, stream);
  else if (bn == 64) rc = launch_tc<64, 6, 8>(maps, args, stream);
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

// Diagnostics: CTA `cta` of every following tcgen05 GEMM launch writes clock64() stamps of its producer / MMA /
// epilogue roles into `device_buf` (6 * 512 int64; pass NULL to switch tracing off). Not for timed runs.
extern "C" int gg_gemm_set_trace(void* device_buf, int cta) {
  gg::g_trace_buf = reinterpret_cast<long long*>(device_buf);
  gg::g_trace_cta = cta;
  return GG_OK;
}

// In-kernel timing (supplement to the CUDA-event profile: an event pair adds ~6 us to a ~6 us launch). Every
// following tcgen05 GEMM launch gets a slot of `device_buf` (2 x capacity uint64, the caller presets every slot to
// {UINT64_MAX, 0} before each run): thread 0 of every CTA does atomicMin(start) after griddepcontrol.wait, the first
// epilogue thread atomicMax(end) after its stores — also inside replayed CUDA graphs. NULL switches it off.
extern "C" int gg_gemm_set_timer(void* device_buf, int capacity) {
  gg::g_stamp_buf = reinterpret_cast<unsigned long long*>(device_buf);
  gg::g_stamp_cap = device_buf ? capacity : 0;
  gg::g_stamp_next = 0;
  gg::g_stamp_rec.clear();
  return GG_OK;
}
// Slots handed out since gg_gemm_set_timer; flops[i] / bytes[i] (optional, length >= count) = algorithmic work of slot i.
extern "C" int gg_gemm_timer_slots(double* flops, double* bytes, int n) {
  const int count = static_cast<int>(gg::g_stamp_rec.size());
  for (int i = 0; i < count && i < n; ++i) {
    if (flops) flops[i] = gg::g_stamp_rec[i].flops;
    if (bytes) bytes[i] = gg::g_stamp_rec[i].bytes;
  }
  return count;
}

extern "C" int gg_gemm_profile_begin(void) {
  gg::g_prof.on = true;
  gg::g_prof.used = 0;
  gg::g_prof.rec.clear();
  gg::g_prof.flops = 0.0;
  gg::g_prof.bytes = 0.0;
  gg::g_prof.launches = 0;
  for (gg::AuxProfile& a : gg::g_auxes) {
    a.used = 0;
    a.flops = a.bytes = 0.0;
    a.launches = 0;
  }
  return GG_OK;
}
// The fused encoder-layer launches of the last profiled region (call after gg_gemm_profile_end).
static int aux_profile_read(const gg::AuxProfile& a, double* ms, double* flops, double* bytes, long long* launches) {
  double total = 0.0;
  for (size_t i = 0; i + 1 < a.used; i += 2) {
    float t = 0.f;
    GG_CUDA_CHECK(cudaEventElapsedTime(&t, a.ev[i], a.ev[i + 1]));
    total += t;
  }
  if (ms) *ms = total;
  if (flops) *flops = a.flops;
  if (bytes) *bytes = a.bytes;
  if (launches) *launches = a.launches;
  return GG_OK;
}
extern "C" int gg_enc_layer_profile(double* ms, double* flops, double* bytes, long long* launches) {
  return aux_profile_read(gg::g_auxes[0], ms, flops, bytes, launches);
}
// The grouped weight-gradient launches (wgrad_group.cu) of the last profiled region: summed duration, 2*M*N*K flops and
// algorithmic bytes (each bf16 operand matrix once, each fp32 output once).
extern "C" int gg_wgrad_group_profile(double* ms, double* flops, double* bytes, long long* launches) {
  return aux_profile_read(gg::g_auxes[1], ms, flops, bytes, launches);
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

// Algorithmic bytes (operands read once + outputs written once) of the launches of the last profiled region.
extern "C" double gg_gemm_profile_bytes(void) { return gg::g_prof.bytes; }

// Writes one CSV line per GEMM launch of the last profiled region (call after gg_gemm_profile_end).
extern "C" int gg_gemm_profile_dump(const char* path) {
  using namespace gg;
  GG_REQUIRE(path, "null path");
  FILE* f = fopen(path, "w");
  GG_REQUIRE(f, "cannot open %s", path);
  fprintf(f, "idx,M,N,K0,K1,block_n,splits,a_mn,b_mn,us,tflops,gbs\n");
  for (size_t i = 0; i + 1 < g_prof.used && i / 2 < g_prof.rec.size(); i += 2) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]) != cudaSuccess) t = 0.f;
    const GemmRecord& r = g_prof.rec[i / 2];
    const double fl = 2.0 * r.M * r.N * (static_cast<double>(r.K0) + r.K1);
    const double by = 2.0 * (static_cast<double>(r.M) + r.N) * (static_cast<double>(r.K0) + r.K1) +
                      static_cast<double>(r.out_bytes) * r.M * r.N;
    fprintf(f, "%zu,%d,%d,%d,%d,%d,%d,%d,%d,%.2f,%.1f,%.1f\n", i / 2, r.M, r.N, r.K0, r.K1, r.bn, r.splits, r.a_mn,
            r.b_mn, t * 1e3, t > 0.f ? fl / (t * 1e-3) / 1e12 : 0.0, t > 0.f ? by / (t * 1e-3) / 1e9 : 0.0);
  }
  fclose(f);
  return GG_OK;
}

extern "C" int gg_gemm_bf16(const gg_gemm_desc* desc, void* stream) {
  return gg::gemm_dispatch(desc, reinterpret_cast<cudaStream_t>(stream));
}
