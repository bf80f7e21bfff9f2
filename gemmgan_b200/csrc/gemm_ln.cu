// GEMM with a residual + dropout + LayerNorm epilogue, for the encoder layers that do not fit the fused
// <= 16-token layer kernel (BASELINE configs 2 and 4: 257 / 65 tokens):
//     z = res + dropout(A W^T + b),  out = LayerNorm(z) * gamma + beta
// i.e. `x = norm1(x + dropout1(out_proj(attn)))` and `x = norm2(x + dropout2(linear2(h)))` of torch's post-norm
// nn.TransformerEncoderLayer (src/conditional_gan_film.py:114-119, ...with_film.py:114-119, :144) — one launch instead of a
// GEMM that writes the projection to HBM and an add + LayerNorm kernel that reads it back.
// d_model = 256 = one 256-column accumulator: a thread of the epilogue owns one token row. Sweep 1 reads the accumulator
// (tcgen05.ld), adds bias, applies the Philox keep mask (same site / element indexing as add_ln_fwd_kernel, so the unfused
// backward regenerates it), adds the residual row, accumulates sum / sum of squares, parks z back in tensor memory
// (tcgen05.st) and writes its bf16 copy (the backward's input); sweep 2 normalises from tensor memory. Two accumulators
// (all 512 TMEM columns): the MMAs of the next tile run under the epilogue of this one.
// Roles per CTA (6 warps): 0 TMA producer, 1 MMA issuer (+ TMEM owner), 2-5 epilogue (TMEM lane quarter = warp % 4).
#include "host_util.h"
#include "pdl.cuh"
#include "kernels.h"
#include "philox.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

namespace gg {

int encode_tma_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer, bool f32);

namespace gl {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int OFF_BAR = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int THREADS = 6 * 32;

struct Args {
  int64_t rows;
  int K;
  const float *bias, *gamma, *beta;  // bias / beta may be null
  const bf16* res;                   // [rows, 256]
  bf16 *z, *out;                     // [rows, 256]
  float *mean, *rstd;                // [rows]
  float eps, drop_p;
  const uint64_t* rng;
  uint32_t site;
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__global__ void __launch_bounds__(THREADS, 1)
    gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const Args g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = g.K / BK;
  const int tiles = static_cast<int>((g.rows + BM - 1) / BM);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);
    }
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

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* a_dst = smem + s * STAGE_BYTES;
          tma_load_2d(a_dst, &tmA, &full[s], kb * BK, tile * BM);
          tma_load_2d(a_dst + A_BYTES, &tmW, &full[s], kb * BK, 0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma_bf16(tmem_base + acc * BN, make_smem_desc(a_base + k * 32, 16, 1024), make_smem_desc(b_base + k * 32, 16, 1024),
                        idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(&acc_full[acc]);
      }
    }
  } else {
    const int r = (warp & 3) * 32 + lane;
    const bool drop = g.drop_p > 0.f;
    uint64_t seed = 0, step = 0;
    if (drop) {
      seed = g.rng[0];
      step = g.rng[1];
    }
    const uint32_t thr = dropout_thr(g.drop_p);
    const float keep_scale = drop ? 1.f / (1.f - g.drop_p) : 1.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int64_t m = static_cast<int64_t>(tile) * BM + r;
      const bool valid = m < g.rows;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + acc * BN;
      const bf16* rrow = g.res + m * BN;
      bf16* zrow = g.z + m * BN;
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tc_fence_after_sync();
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        uint4 rv[4];
        if (valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) rv[q] = __ldg(reinterpret_cast<const uint4*>(rrow + c * 32) + q);
        }
        uint32_t keep = 0xFFFFFFFFu;
        if (drop && valid) {
          keep = 0;
          const uint64_t g0 = (static_cast<uint64_t>(m) * BN + c * 32) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) keep |= keep_bits8(dropout_words(seed, step, g.site, g0 + j), thr) << (8 * j);
        }
        tmem_ld_wait();
        if (valid) {
          const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(rv);
          uint4 zo[4];
          __nv_bfloat162* zb = reinterpret_cast<__nv_bfloat162*>(zo);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float2 rr = __bfloat1622float2(rb[q]);
            float y0 = v[2 * q] + (g.bias ? __ldg(g.bias + c * 32 + 2 * q) : 0.f);
            float y1 = v[2 * q + 1] + (g.bias ? __ldg(g.bias + c * 32 + 2 * q + 1) : 0.f);
            y0 = ((keep >> (2 * q)) & 1u) ? y0 * keep_scale : 0.f;
            y1 = ((keep >> (2 * q + 1)) & 1u) ? y1 * keep_scale : 0.f;
            const float z0 = rr.x + y0, z1 = rr.y + y1;
            v[2 * q] = z0;
            v[2 * q + 1] = z1;
            s1 += z0 + z1;
            s2 = fmaf(z0, z0, fmaf(z1, z1, s2));
            zb[q] = __floats2bfloat162_rn(z0, z1);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(zrow + c * 32)[q] = zo[q];
        }
        tmem_st_32x32(taddr + c * 32, v);
      }
      tmem_st_wait();
      const float mu = s1 * (1.f / BN);
      const float var = fmaxf(s2 * (1.f / BN) - mu * mu, 0.f);
      const float rs = rsqrtf(var + g.eps);
      if (valid) {
        g.mean[m] = mu;
        g.rstd[m] = rs;
      }
      bf16* orow = g.out + m * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        if (!valid) continue;
        uint4 oo[4];
        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(oo);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int col = c * 32 + 2 * q;
          const float w0 = __ldg(g.gamma + col), w1 = __ldg(g.gamma + col + 1);
          const float b0 = g.beta ? __ldg(g.beta + col) : 0.f, b1 = g.beta ? __ldg(g.beta + col + 1) : 0.f;
          ob[q] = __floats2bfloat162_rn((v[2 * q] - mu) * rs * w0 + b0, (v[2 * q + 1] - mu) * rs * w1 + b1);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(orow + c * 32)[q] = oo[q];
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace gl

// z = res + dropout(a w^T + bias), out = LayerNorm(z) * gamma + beta over 256 columns. a bf16 [rows, K] (pitch lda, K % 64 == 0),
// w bf16 [256, K] (pitch ldw), res / z / out bf16 [rows, 256], mean / rstd fp32 [rows]; bias / beta may be null. The dropout
// stream is (rng, site) with element index row * 256 + column, as in k_add_ln_fwd.
int k_gemm_ln(const bf16* a, int64_t lda, const bf16* w, int64_t ldw, int K, const float* bias, const bf16* res,
              const float* gamma, const float* beta, bf16* z, bf16* out, float* mean, float* rstd, int64_t rows, float eps,
              float drop_p, const uint64_t* rng, uint32_t site, cudaStream_t st) {
  GG_REQUIRE(a && w && res && gamma && z && out && mean && rstd && rows > 0, "bad gemm_ln argument");
  GG_REQUIRE(K > 0 && K % gl::BK == 0, "gemm_ln: K must be a multiple of 64");
  GG_REQUIRE(drop_p == 0.f || rng, "dropout needs an rng state pointer");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int num_sms = 0;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gl::gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gl::SMEM_BYTES);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  });
  GG_CUDA_CHECK(attr_err);
  CUtensorMap ma, mw;
  GG_TRY_RC(encode_tma_map(&ma, a, K, rows, lda, gl::BM, false));
  GG_TRY_RC(encode_tma_map(&mw, w, K, gl::BN, ldw, gl::BN, false));
  gl::Args g;
  g.rows = rows; g.K = K;
  g.bias = bias; g.gamma = gamma; g.beta = beta; g.res = res; g.z = z; g.out = out; g.mean = mean; g.rstd = rstd;
  g.eps = eps; g.drop_p = drop_p; g.rng = rng; g.site = site;
  const int64_t tiles = (rows + gl::BM - 1) / gl::BM;
  launch_k(gl::gemm_ln_kernel, static_cast<unsigned>(tiles < num_sms ? tiles : num_sms), gl::THREADS, gl::SMEM_BYTES, st, ma, mw, g);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg
