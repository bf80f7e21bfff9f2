// Evaluation metrics of the reference's fit() loop on the GPU: pairwise sample distances,
// k-th nearest-neighbour radii, the PRDC / manifold membership counts, per-gene standardisation, gene-gene Pearson
// correlation and the fused gamma coefficient. The reference runs these on the host with sklearn / numpy
// (src/distribution_distances.py:51-142, src/unsupervised_metrics.py:114-324, src/corr_score.py:43-120) or as
// broadcasted torch expressions that materialise [128, N, G] tensors (src/privacy_evaluator.py:9-66).
//
// All arithmetic is fp32 on the CUDA cores (fp64 for the gamma sums): the L1 distance |x - y| is not a product, and
// nearest-neighbour *comparisons* (d < radius, first / second neighbour ratios) do not survive bf16 operands, so none
// of this is reshaped into a tensor-core GEMM. The tile kernels are register-blocked (64 x 64 outputs per CTA, 4 x 4
// per thread, operands staged through shared memory) and compute bound; the row / column reductions are HBM bound.
#include <math.h>

#include "host_util.h"
#include "kernels.h"
#include "pdl.cuh"

namespace gg {

namespace {

constexpr int ET = 64;         // outputs per CTA along each axis
constexpr int EK = 16;         // reduction slice staged per step
constexpr int EPITCH = ET + 4; // floats; keeps float4 reads aligned and spreads the transposing stores over the banks

int require_sm100() {
  static int cached_dev = -1;
  int dev = 0;
  GG_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev == cached_dev) return GG_OK;
  GG_TRY_RC(gg_check_device(dev));
  cached_dev = dev;
  return GG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// out[i, j] = dist(x[i, :], y[j, :]) for row-major x [n, d], y [m, d]; METRIC 0: sum |a-b|, 1: sum (a-b)^2, 2: sqrt.
template <int METRIC>
__global__ void __launch_bounds__(256)
    pairwise_distance_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ y, int64_t ldy, int n,
                             int m, int d, float* __restrict__ out, int64_t ldo) {
  pdl_entry();
  __shared__ __align__(16) float xs[EK][EPITCH];
  __shared__ __align__(16) float ys[EK][EPITCH];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.y * ET, col0 = blockIdx.x * ET;
  // staging: thread -> (tile row lr, four consecutive features lk .. lk+3)
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int xr = row0 + lr, yr = col0 + lr;
  const float* xp = x + static_cast<int64_t>(min(xr, n - 1)) * ldx;
  const float* yp = y + static_cast<int64_t>(min(yr, m - 1)) * ldy;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < d; k0 += EK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + lk + u;
      // features past d read as 0 on both sides: |0 - 0| adds nothing. Rows past n / m are clamped (never stored).
      xs[lk + u][lr] = k < d ? __ldg(xp + k) : 0.f;
      ys[lk + u][lr] = k < d ? __ldg(yp + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < EK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&ys[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float t = av[i] - bv[j];
          if (METRIC == 0) acc[i][j] += fabsf(t);
          else acc[i][j] = fmaf(t, t, acc[i][j]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c < m) out[static_cast<int64_t>(r) * ldo + c] = METRIC == 2 ? sqrtf(acc[i][j]) : acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// out[i, j] = scale * sum_s a[s, i] * b[s, j] for sample-major a [n, ga], b [n, gb] (np.dot(x_.T, y_) / n).
__global__ void __launch_bounds__(256)
    gene_correlation_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb, int n,
                            int ga, int gb, float scale, float* __restrict__ out, int64_t ldo) {
  pdl_entry();
  __shared__ __align__(16) float as[EK][EPITCH];
  __shared__ __align__(16) float bs[EK][EPITCH];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.y * ET, col0 = blockIdx.x * ET;
  const int lc = tid & 63, ls = tid >> 6;  // staging: thread -> (gene lc of the tile, samples ls, ls+4, ls+8, ls+12)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int s0 = 0; s0 < n; s0 += EK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = s0 + ls + 4 * u;
      as[ls + 4 * u][lc] = (s < n && row0 + lc < ga) ? __ldg(a + static_cast<int64_t>(s) * lda + row0 + lc) : 0.f;
      bs[ls + 4 * u][lc] = (s < n && col0 + lc < gb) ? __ldg(b + static_cast<int64_t>(s) * ldb + col0 + lc) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < EK; ++k) {
      const float4 av4 = *reinterpret_cast<const float4*>(&as[k][ty * 4]);
      const float4 bv4 = *reinterpret_cast<const float4*>(&bs[k][tx * 4]);
      const float av[4] = {av4.x, av4.y, av4.z, av4.w};
      const float bv[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= ga) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c < gb) out[static_cast<int64_t>(r) * ldo + c] = scale * acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused gamma coefficient (corr_score.py:71-120): for every gene pair i < j the two correlations
// cx = sum_s xs[s,i] xs[s,j] / nx and cy = sum_s ys[s,i] ys[s,j] / ny are formed in registers and only their five
// moments leave the CTA (fp64): count, sum cx, sum cy, sum cx^2, sum cy^2, sum cx*cy. Neither [G, G] matrix nor
// the G(G-1)/2-long lists of upper_diag_list are ever written. Tiles below the diagonal exit at once.
__device__ __forceinline__ void tile_gram(const float* __restrict__ a, int64_t lda, int n, int g, int row0, int col0,
                                          float (*as)[EPITCH], float (*bs)[EPITCH], float acc[4][4]) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lc = tid & 63, ls = tid >> 6;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int s0 = 0; s0 < n; s0 += EK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = s0 + ls + 4 * u;
      as[ls + 4 * u][lc] = (s < n && row0 + lc < g) ? __ldg(a + static_cast<int64_t>(s) * lda + row0 + lc) : 0.f;
      bs[ls + 4 * u][lc] = (s < n && col0 + lc < g) ? __ldg(a + static_cast<int64_t>(s) * lda + col0 + lc) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < EK; ++k) {
      const float4 av4 = *reinterpret_cast<const float4*>(&as[k][ty * 4]);
      const float4 bv4 = *reinterpret_cast<const float4*>(&bs[k][tx * 4]);
      const float av[4] = {av4.x, av4.y, av4.z, av4.w};
      const float bv[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
}

constexpr int GAMMA_MOMENTS = 6;

__global__ void __launch_bounds__(256)
    gamma_moments_kernel(const float* __restrict__ xs, int64_t ldx, int nx, const float* __restrict__ ys, int64_t ldy,
                         int ny, int g, double* __restrict__ partials) {
  pdl_entry();
  __shared__ __align__(16) float as[EK][EPITCH];
  __shared__ __align__(16) float bs[EK][EPITCH];
  __shared__ double red[8][GAMMA_MOMENTS];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  double* mine = partials + (static_cast<int64_t>(blockIdx.y) * gridDim.x + blockIdx.x) * GAMMA_MOMENTS;
  if (blockIdx.x < blockIdx.y) {  // strictly below the diagonal: no pair with i < j
    if (tid < GAMMA_MOMENTS) mine[tid] = 0.0;
    return;
  }
  const int row0 = blockIdx.y * ET, col0 = blockIdx.x * ET;
  float cx[4][4], cy[4][4];
  tile_gram(xs, ldx, nx, g, row0, col0, as, bs, cx);
  tile_gram(ys, ldy, ny, g, row0, col0, as, bs, cy);
  const float inx = 1.f / static_cast<float>(nx), iny = 1.f / static_cast<float>(ny);
  double m[GAMMA_MOMENTS] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = row0 + ty * 4 + i, gj = col0 + tx * 4 + j;
      if (gi < gj && gj < g) {
        const double a = static_cast<double>(cx[i][j] * inx), b = static_cast<double>(cy[i][j] * iny);
        m[0] += 1.0; m[1] += a; m[2] += b; m[3] += a * a; m[4] += b * b; m[5] += a * b;
      }
    }
  // fixed-order reduction: lanes by shuffle, then the 8 warps in index order
#pragma unroll
  for (int q = 0; q < GAMMA_MOMENTS; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m[q] += __shfl_down_sync(0xffffffffu, m[q], o);
  if ((tid & 31) == 0)
#pragma unroll
    for (int q = 0; q < GAMMA_MOMENTS; ++q) red[tid >> 5][q] = m[q];
  __syncthreads();
  if (tid < GAMMA_MOMENTS) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][tid];
    mine[tid] = s;
  }
}

// sums[q] = sum over CTAs (index order) of partials[cta, q]; one CTA, fixed order -> deterministic.
__global__ void __launch_bounds__(256) gamma_finish_kernel(const double* __restrict__ partials, int64_t n_cta,
                                                           double* __restrict__ sums) {
  pdl_entry();
  __shared__ double red[256];
  for (int q = 0; q < GAMMA_MOMENTS; ++q) {
    double s = 0.0;
    for (int64_t c = threadIdx.x; c < n_cta; c += 256) s += partials[c * GAMMA_MOMENTS + q];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) sums[q] = red[0];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// out[s, c] = (x[s, c] - mean_c) / std_c (population std, ddof 0); a constant column gives 0 (the reference turns
// its 0/0 NaNs back into x - mean = 0, corr_score.py:55-61). One thread per column, coalesced across columns,
// moments in fp64.
__global__ void __launch_bounds__(256)
    standardize_columns_kernel(const float* __restrict__ x, int64_t ldx, int n, int g, float* __restrict__ out,
                               int64_t ldo) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g) return;
  double s = 0.0;
  for (int r = 0; r < n; ++r) s += static_cast<double>(x[static_cast<int64_t>(r) * ldx + c]);
  const double mean = s / n;
  double v = 0.0;
  for (int r = 0; r < n; ++r) {
    const double t = static_cast<double>(x[static_cast<int64_t>(r) * ldx + c]) - mean;
    v += t * t;
  }
  const double sd = sqrt(v / n);
  const float meanf = static_cast<float>(mean);
  const float inv = sd > 0.0 ? static_cast<float>(1.0 / sd) : 1.f;
  for (int r = 0; r < n; ++r)
    out[static_cast<int64_t>(r) * ldo + c] = (x[static_cast<int64_t>(r) * ldx + c] - meanf) * inv;
}

// ---------------------------------------------------------------------------------------------------------------
// kth[i] = the (k+1)-th smallest entry of row i (sorted[k], duplicates counted), argmin[i] = first index of the
// smallest. One CTA per row: pass p finds the smallest value above the previous one and how often it occurs, so a
// row is read at most k + 1 times (from L1 / L2) and nothing is sorted or written.
struct MinCount {
  float v;
  int cnt;
  int idx;
};
__device__ __forceinline__ MinCount combine(MinCount a, MinCount b) {
  if (b.v < a.v) return b;
  if (a.v < b.v) return a;
  MinCount r;
  r.v = a.v;
  r.cnt = a.cnt + b.cnt;
  r.idx = min(a.idx, b.idx);
  return r;
}

__global__ void __launch_bounds__(256)
    row_kth_smallest_kernel(const float* __restrict__ dist, int64_t ld, int m, int k, float* __restrict__ kth,
                            int32_t* __restrict__ argmin) {
  pdl_entry();
  __shared__ MinCount red[8];
  __shared__ MinCount winner;
  const float* row = dist + static_cast<int64_t>(blockIdx.x) * ld;
  const float inf = __int_as_float(0x7f800000);
  float prev = -inf;
  bool first = true;
  int remaining = k;
  for (;;) {
    MinCount mc;
    mc.v = inf; mc.cnt = 0; mc.idx = 0x7fffffff;
    for (int j = threadIdx.x; j < m; j += 256) {
      const float v = row[j];
      if (first ? true : v > prev) {
        MinCount e;
        e.v = v; e.cnt = 1; e.idx = j;
        mc = combine(mc, e);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      MinCount other;
      other.v = __shfl_down_sync(0xffffffffu, mc.v, o);
      other.cnt = __shfl_down_sync(0xffffffffu, mc.cnt, o);
      other.idx = __shfl_down_sync(0xffffffffu, mc.idx, o);
      mc = combine(mc, other);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mc;
    __syncthreads();
    if (threadIdx.x == 0) {
      MinCount w = red[0];
      for (int q = 1; q < 8; ++q) w = combine(w, red[q]);
      winner = w;
    }
    __syncthreads();
    const MinCount w = winner;
    __syncthreads();  // `winner` / `red` are rewritten by the next pass
    if (first && argmin && threadIdx.x == 0) argmin[blockIdx.x] = w.idx;
    if (w.cnt == 0 || remaining < w.cnt) {  // cnt == 0: fewer than k+1 entries -> +inf
      if (threadIdx.x == 0) kth[blockIdx.x] = w.v;
      return;
    }
    remaining -= w.cnt;
    prev = w.v;
    first = false;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Row side of compute_prdc / ManifoldEstimator.evaluate for dist [n, m] (one CTA per row i):
//   row_any[i]   = any_j dist[i, j] (< or <=) col_radius[j]      (recall / manifold membership)
//   row_min[i]   = min_j dist[i, j], row_argmin[i] = first such j  (coverage / nearest neighbour)
//   row_ratio[i] = max_j col_radius[j] / (dist[i, j] + eps)      (realism score)
// Any output (and col_radius) may be NULL.
__global__ void __launch_bounds__(256)
    row_membership_kernel(const float* __restrict__ dist, int64_t ld, int m, const float* __restrict__ col_radius,
                          int inclusive, float eps, uint8_t* __restrict__ row_any, float* __restrict__ row_min,
                          int32_t* __restrict__ row_argmin, float* __restrict__ row_ratio) {
  pdl_entry();
  __shared__ MinCount red[8];
  __shared__ float red_ratio[8];
  __shared__ int red_any[8];
  const float* row = dist + static_cast<int64_t>(blockIdx.x) * ld;
  const float inf = __int_as_float(0x7f800000);
  MinCount mc;
  mc.v = inf; mc.cnt = 0; mc.idx = 0x7fffffff;
  float ratio = -inf;
  int any = 0;
  for (int j = threadIdx.x; j < m; j += 256) {
    const float v = row[j];
    MinCount e;
    e.v = v; e.cnt = 1; e.idx = j;
    mc = combine(mc, e);
    if (col_radius) {
      const float r = col_radius[j];
      any |= inclusive ? (v <= r) : (v < r);
      ratio = fmaxf(ratio, r / (v + eps));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MinCount other;
    other.v = __shfl_down_sync(0xffffffffu, mc.v, o);
    other.cnt = __shfl_down_sync(0xffffffffu, mc.cnt, o);
    other.idx = __shfl_down_sync(0xffffffffu, mc.idx, o);
    mc = combine(mc, other);
    ratio = fmaxf(ratio, __shfl_down_sync(0xffffffffu, ratio, o));
    any |= __shfl_down_sync(0xffffffffu, any, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = mc;
    red_ratio[threadIdx.x >> 5] = ratio;
    red_any[threadIdx.x >> 5] = any;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) {
      mc = combine(mc, red[q]);
      ratio = fmaxf(ratio, red_ratio[q]);
      any |= red_any[q];
    }
    if (row_any) row_any[blockIdx.x] = any ? 1 : 0;
    if (row_min) row_min[blockIdx.x] = mc.v;
    if (row_argmin) row_argmin[blockIdx.x] = mc.idx;
    if (row_ratio) row_ratio[blockIdx.x] = ratio;
  }
}

// Column side: col_hits[j] += #{ i in this CTA's row chunk : dist[i, j] < row_radius[i] } (precision / density).
// Thread = column (coalesced), blockIdx.y = chunk of rows; integer atomics, so the result is order independent.
constexpr int COL_ROWS_PER_CTA = 128;
__global__ void __launch_bounds__(256)
    col_hits_kernel(const float* __restrict__ dist, int64_t ld, int n, int m, const float* __restrict__ row_radius,
                    int inclusive, int32_t* __restrict__ col_hits) {
  pdl_entry();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int r0 = blockIdx.y * COL_ROWS_PER_CTA;
  const int r1 = min(n, r0 + COL_ROWS_PER_CTA);
  int hits = 0;
  for (int i = r0; i < r1; ++i) {
    const float v = dist[static_cast<int64_t>(i) * ld + j];
    const float r = row_radius[i];
    hits += inclusive ? (v <= r) : (v < r);
  }
  if (hits) atomicAdd(col_hits + j, hits);
}

}  // namespace

}  // namespace gg

using namespace gg;

extern "C" int gg_pairwise_distance(const float* x, int64_t ldx, const float* y, int64_t ldy, int32_t n, int32_t m,
                                    int32_t d, int32_t metric, float* out, int64_t ldo, void* stream) {
  GG_REQUIRE(x && y && out && n > 0 && m > 0 && d > 0, "gg_pairwise_distance: bad argument");
  GG_REQUIRE(ldx >= d && ldy >= d && ldo >= m, "gg_pairwise_distance: leading dimension smaller than the row");
  GG_REQUIRE(metric >= GG_DIST_L1 && metric <= GG_DIST_L2, "gg_pairwise_distance: unknown metric %d", metric);
  GG_TRY_RC(require_sm100());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>(ceil_div(m, ET)), static_cast<unsigned>(ceil_div(n, ET)));
  GG_REQUIRE(grid.y <= 65535u, "gg_pairwise_distance: more than 65535 row tiles; split the rows");
  if (metric == GG_DIST_L1) launch_k(pairwise_distance_kernel<0>, grid, 256, 0, st, x, ldx, y, ldy, n, m, d, out, ldo);
  else if (metric == GG_DIST_SQL2) launch_k(pairwise_distance_kernel<1>, grid, 256, 0, st, x, ldx, y, ldy, n, m, d, out, ldo);
  else launch_k(pairwise_distance_kernel<2>, grid, 256, 0, st, x, ldx, y, ldy, n, m, d, out, ldo);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_row_kth_smallest(const float* dist, int64_t ld, int32_t n, int32_t m, int32_t k, float* kth,
                                   int32_t* argmin, void* stream) {
  GG_REQUIRE(dist && kth && n > 0 && m > 0 && ld >= m, "gg_row_kth_smallest: bad argument");
  GG_REQUIRE(k >= 0 && k < m, "gg_row_kth_smallest: rank %d outside a row of %d", k, m);
  GG_TRY_RC(require_sm100());
  launch_k(row_kth_smallest_kernel, static_cast<unsigned>(n), 256, 0, reinterpret_cast<cudaStream_t>(stream), dist, ld,
           m, k, kth, argmin);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_row_membership(const float* dist, int64_t ld, int32_t n, int32_t m, const float* col_radius,
                                 int32_t inclusive, float eps, uint8_t* row_any, float* row_min, int32_t* row_argmin,
                                 float* row_ratio, void* stream) {
  GG_REQUIRE(dist && n > 0 && m > 0 && ld >= m, "gg_row_membership: bad argument");
  GG_REQUIRE(col_radius || (!row_any && !row_ratio), "gg_row_membership: row_any / row_ratio need col_radius");
  GG_TRY_RC(require_sm100());
  launch_k(row_membership_kernel, static_cast<unsigned>(n), 256, 0, reinterpret_cast<cudaStream_t>(stream), dist, ld, m,
           col_radius, inclusive, eps, row_any, row_min, row_argmin, row_ratio);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_col_hits(const float* dist, int64_t ld, int32_t n, int32_t m, const float* row_radius,
                           int32_t inclusive, int32_t* col_hits, void* stream) {
  GG_REQUIRE(dist && row_radius && col_hits && n > 0 && m > 0 && ld >= m, "gg_col_hits: bad argument");
  GG_TRY_RC(require_sm100());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>(ceil_div(m, 256)), static_cast<unsigned>(ceil_div(n, COL_ROWS_PER_CTA)));
  GG_REQUIRE(grid.y <= 65535u, "gg_col_hits: too many rows; split them");
  launch_k(col_hits_kernel, grid, 256, 0, st, dist, ld, n, m, row_radius, inclusive, col_hits);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_standardize_columns(const float* x, int64_t ldx, int32_t n, int32_t g, float* out, int64_t ldo,
                                      void* stream) {
  GG_REQUIRE(x && out && n > 0 && g > 0 && ldx >= g && ldo >= g, "gg_standardize_columns: bad argument");
  GG_TRY_RC(require_sm100());
  launch_k(standardize_columns_kernel, static_cast<unsigned>(ceil_div(g, 256)), 256, 0,
           reinterpret_cast<cudaStream_t>(stream), x, ldx, n, g, out, ldo);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_gene_correlation(const float* a, int64_t lda, const float* b, int64_t ldb, int32_t n, int32_t ga,
                                   int32_t gb, float* out, int64_t ldo, void* stream) {
  GG_REQUIRE(a && b && out && n > 0 && ga > 0 && gb > 0, "gg_gene_correlation: bad argument");
  GG_REQUIRE(lda >= ga && ldb >= gb && ldo >= gb, "gg_gene_correlation: leading dimension smaller than the row");
  GG_TRY_RC(require_sm100());
  dim3 grid(static_cast<unsigned>(ceil_div(gb, ET)), static_cast<unsigned>(ceil_div(ga, ET)));
  GG_REQUIRE(grid.y <= 65535u, "gg_gene_correlation: too many genes");
  launch_k(gene_correlation_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), a, lda, b, ldb, n, ga, gb,
           1.f / static_cast<float>(n), out, ldo);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int64_t gg_gamma_moments_workspace_bytes(int32_t g) {
  if (g <= 0) return 0;
  const int64_t t = ceil_div(g, ET);
  return t * t * GAMMA_MOMENTS * static_cast<int64_t>(sizeof(double));
}

extern "C" int gg_gamma_moments(const float* xs, int64_t ldx, int32_t nx, const float* ys, int64_t ldy, int32_t ny,
                                int32_t g, void* workspace, int64_t workspace_bytes, double* sums, void* stream) {
  GG_REQUIRE(xs && ys && sums && nx > 0 && ny > 0 && g > 1 && ldx >= g && ldy >= g, "gg_gamma_moments: bad argument");
  if (!workspace || workspace_bytes < gg_gamma_moments_workspace_bytes(g)) {
    set_error("gg_gamma_moments: workspace of %lld bytes, %lld needed", static_cast<long long>(workspace_bytes),
              static_cast<long long>(gg_gamma_moments_workspace_bytes(g)));
    return GG_ERR_WORKSPACE;
  }
  GG_TRY_RC(require_sm100());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int t = ceil_div(g, ET);
  GG_REQUIRE(t <= 65535, "gg_gamma_moments: too many genes");
  dim3 grid(static_cast<unsigned>(t), static_cast<unsigned>(t));
  double* partials = static_cast<double*>(workspace);
  launch_k(gamma_moments_kernel, grid, 256, 0, st, xs, ldx, nx, ys, ldy, ny, g, partials);
  GG_LAUNCH_CHECK();
  launch_k(gamma_finish_kernel, 1, 256, 0, st, static_cast<const double*>(partials), static_cast<int64_t>(t) * t, sums);
  GG_LAUNCH_CHECK();
  return GG_OK;
}
