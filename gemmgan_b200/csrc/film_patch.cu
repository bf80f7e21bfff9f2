// FiLM in the A-operand path of the patch encoder: X0 = [cls | (gamma_b * patches + beta_b) Wp^T + bp] in one
// kernel — reference src/conditional_gan_cross_attention_with_film.py:129-142 (`patches = gamma * patches + beta`,
// `patches_encoder`, the CLS token prepended) for all R dropout replicas of the tower at once.
//
// The modulated patches never make a round trip through HBM for the GEMM: TMA brings 128 x 64 bf16 boxes of the RAW patch
// embeddings into a ring, four transform warps (thread = tile row) apply this sample's gamma / beta (fp32, L1 / L2 hits: the
// P rows of a sample share them) and write the bf16, 128B-swizzled, K-major A operand tcgen05.mma reads; the same threads
// store the modulated row to `mod` when the tower will be back-propagated (its weight gradient dWp = dpe^T mod needs it),
// and nothing when it will not (the generator's tower inside a critic step). The epilogue adds the bias and writes each
// output row behind the CLS row of its sequence in every replica; the thread of a sequence's first patch also writes the CLS
// row, so the separate token-assembly pass disappears as well.
// Roles per CTA (6 warps): 0 TMA producer, 1 MMA issuer (+ TMEM owner), 2-5 transform, then epilogue (TMEM lane quarter =
// warp % 4). N = E = 256 (one 256-column accumulator), K = Dp in 64-column blocks.
#include "host_util.h"
#include "pdl.cuh"
#include "kernels.h"
#include "ptx.cuh"

#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

namespace gg {

int encode_tma_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer, bool f32);

namespace fp {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int RAW_STAGES = 3, A_STAGES = 2, B_STAGES = 3;
constexpr int RAW_BYTES = BM * BK * 2;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int OFF_RAW = 0;
constexpr int OFF_A = OFF_RAW + RAW_STAGES * RAW_BYTES;
constexpr int OFF_B = OFF_A + A_STAGES * A_BYTES;
constexpr int OFF_BAR = OFF_B + B_STAGES * B_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int THREADS = 6 * 32;

struct Args {
  int rows;            // B * P patch rows
  int P, S, R, Dp;
  int64_t rep_rows;    // B * S token rows per replica
  const float* gb;     // [B, 2 * Dp]: gamma | beta
  const float* bias;   // [256] or null
  const float* cls;    // [256]
  bf16* x0;            // [R * B * S, 256]
  bf16* mod;           // [B * P, Dp] or null
};

__global__ void __launch_bounds__(THREADS, 1)
    film_patch_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Args g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = raw_full + RAW_STAGES;
  uint64_t* a_full = raw_empty + RAW_STAGES;
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* b_full = a_empty + A_STAGES;
  uint64_t* b_empty = b_full + B_STAGES;
  uint64_t* acc_full = b_empty + B_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = g.Dp / BK;
  const int tiles = (g.rows + BM - 1) / BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < RAW_STAGES; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], 4);
    }
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(&a_full[s], 4);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, BN);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  pdl_entry();

  if (warp == 0) {
    if (lane == 0) {
      int rs = 0, bs = 0;
      uint32_t rph = 0, bph = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&b_empty[bs], bph ^ 1);
          mbar_arrive_expect_tx(&b_full[bs], B_BYTES);
          tma_load_2d(smem + OFF_B + bs * B_BYTES, &tmW, &b_full[bs], kb * BK, 0);
          if (++bs == B_STAGES) { bs = 0; bph ^= 1; }
          mbar_wait(&raw_empty[rs], rph ^ 1);
          mbar_arrive_expect_tx(&raw_full[rs], RAW_BYTES);
          tma_load_2d(smem + OFF_RAW + rs * RAW_BYTES, &tmX, &raw_full[rs], kb * BK, tile * BM);
          if (++rs == RAW_STAGES) { rs = 0; rph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        mbar_wait(acc_empty, (it & 1) ^ 1);
        tc_fence_after_sync();
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&a_full[as], aph);
          mbar_wait(&b_full[bs], bph);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(smem + OFF_A + as * A_BYTES);
          const uint32_t b_base = smem_u32(smem + OFF_B + bs * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma_bf16(tmem_base, make_smem_desc(a_base + k * 32, 16, 1024), make_smem_desc(b_base + k * 32, 16, 1024),
                        idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&a_empty[as]);
          tc_commit(&b_empty[bs]);
          if (++as == A_STAGES) { as = 0; aph ^= 1; }
          if (++bs == B_STAGES) { bs = 0; bph ^= 1; }
        }
        tc_commit(acc_full);
      }
    }
  } else {
    const int r = (warp & 3) * 32 + lane;  // tile row of this thread, in the transform and (TMEM lane) in the epilogue
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    int rs = 0, as = 0;
    uint32_t rph = 0, aph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int m = tile * BM + r;
      const bool valid = m < g.rows;
      const int b = (valid ? m : g.rows - 1) / g.P;
      const float* gam = g.gb + static_cast<int64_t>(b) * 2 * g.Dp;
      bf16* mrow = (g.mod && valid) ? g.mod + static_cast<int64_t>(m) * g.Dp : nullptr;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&raw_full[rs], rph);
        mbar_wait(&a_empty[as], aph ^ 1);
        const uint8_t* src = smem + OFF_RAW + rs * RAW_BYTES + r * 128;
        uint8_t* dst = smem + OFF_A + as * A_BYTES + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // 16-byte chunk c = features kb * 64 + 8 c .. + 7
          const uint4 xv = *reinterpret_cast<const uint4*>(src + ((static_cast<uint32_t>(c) ^ sw) << 4));
          const float* gp = gam + kb * BK + 8 * c;
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp)), g1 = __ldg(reinterpret_cast<const float4*>(gp) + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(gp + g.Dp)),
                       b1 = __ldg(reinterpret_cast<const float4*>(gp + g.Dp) + 1);
          const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&xv);
          const float2 x0 = __bfloat1622float2(x[0]), x1 = __bfloat1622float2(x[1]), x2 = __bfloat1622float2(x[2]),
                       x3 = __bfloat1622float2(x[3]);
          uint4 ov;
          __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&ov);
          o[0] = __floats2bfloat162_rn(fmaf(g0.x, x0.x, b0.x), fmaf(g0.y, x0.y, b0.y));
          o[1] = __floats2bfloat162_rn(fmaf(g0.z, x1.x, b0.z), fmaf(g0.w, x1.y, b0.w));
          o[2] = __floats2bfloat162_rn(fmaf(g1.x, x2.x, b1.x), fmaf(g1.y, x2.y, b1.y));
          o[3] = __floats2bfloat162_rn(fmaf(g1.z, x3.x, b1.z), fmaf(g1.w, x3.y, b1.w));
          if (!valid) ov = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(c) ^ sw) << 4)) = ov;
          if (mrow) *reinterpret_cast<uint4*>(mrow + kb * BK + 8 * c) = ov;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_full[as]);
          mbar_arrive(&raw_empty[rs]);
        }
        if (++rs == RAW_STAGES) { rs = 0; rph ^= 1; }
        if (++as == A_STAGES) { as = 0; aph ^= 1; }
      }
      // ---- epilogue: + bias, bf16, behind the CLS row of the sequence in every replica
      mbar_wait(acc_full, it & 1);
      tc_fence_after_sync();
      const int j = valid ? m % g.P : 0;
      bf16* orow = g.x0 + (static_cast<int64_t>(b) * g.S + 1 + j) * BN;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        if (!valid) continue;
        uint4 ov[4];
        __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(ov);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float b0 = g.bias ? __ldg(g.bias + c * 32 + 2 * q) : 0.f, b1 = g.bias ? __ldg(g.bias + c * 32 + 2 * q + 1) : 0.f;
          o[q] = __floats2bfloat162_rn(v[2 * q] + b0, v[2 * q + 1] + b1);
        }
        for (int rep = 0; rep < g.R; ++rep) {
          uint4* d = reinterpret_cast<uint4*>(orow + rep * g.rep_rows * BN + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) d[q] = ov[q];
        }
        if (j == 0) {  // first patch of the sequence: this thread also writes the CLS row
          uint4 cv[4];
          __nv_bfloat162* co = reinterpret_cast<__nv_bfloat162*>(cv);
#pragma unroll
          for (int q = 0; q < 16; ++q)
            co[q] = __floats2bfloat162_rn(__ldg(g.cls + c * 32 + 2 * q), __ldg(g.cls + c * 32 + 2 * q + 1));
          for (int rep = 0; rep < g.R; ++rep) {
            uint4* d = reinterpret_cast<uint4*>(orow - BN + rep * g.rep_rows * BN + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) d[q] = cv[q];
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, BN);
  }
}

}  // namespace fp

// x0 [R * B * S, 256] (S = P + 1): row (r, b, 0) = cls, row (r, b, 1 + j) = (gamma_b * patches[b, j] + beta_b) w^T + bias for every
// replica r. patches bf16 [B * P, Dp] (Dp % 64 == 0), gb fp32 [B, 2 * Dp], w bf16 [256, Dp] (pitch ldw), mod (optional) bf16
// [B * P, Dp] receives the modulated patches.
int k_film_patch(const bf16* patches, const float* gb, const bf16* w, int64_t ldw, const float* bias, const float* cls, bf16* x0,
                 bf16* mod, int B, int P, int R, int Dp, cudaStream_t st) {
  GG_REQUIRE(patches && gb && w && cls && x0 && B > 0 && P > 0 && R > 0 && Dp > 0, "bad film_patch argument");
  GG_REQUIRE(Dp % fp::BK == 0, "film_patch: the patch feature width must be a multiple of 64");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int num_sms = 0;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(fp::film_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fp::SMEM_BYTES);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  });
  GG_CUDA_CHECK(attr_err);
  const int rows = B * P;
  CUtensorMap mx, mw;
  GG_TRY_RC(encode_tma_map(&mx, patches, Dp, rows, Dp, fp::BM, false));
  GG_TRY_RC(encode_tma_map(&mw, w, Dp, fp::BN, ldw, fp::BN, false));
  fp::Args a;
  a.rows = rows; a.P = P; a.S = P + 1; a.R = R; a.Dp = Dp;
  a.rep_rows = static_cast<int64_t>(B) * (P + 1);
  a.gb = gb; a.bias = bias; a.cls = cls; a.x0 = x0; a.mod = mod;
  const int tiles = (rows + fp::BM - 1) / fp::BM;
  launch_k(fp::film_patch_kernel, static_cast<unsigned>(tiles < num_sms ? tiles : num_sms), fp::THREADS, fp::SMEM_BYTES, st, mx, mw,
           a);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg
