// gemm_inner for panels of very few vectors (k*m <= 4: dot, the [w x 1] overlaps of the Gram-Schmidt steps,
// reference src/molpro/linalg/itsolv/propose_rspace.h:430-443 and :451-463, the [1 x q] blocks of DIIS).
// With so few vectors there is nothing to share between threads, so the shared-memory pipeline of gemm_inner.cu only adds
// latency: here every thread streams its own row pairs with 128-bit coalesced loads (4 pairs of every vector in flight
// per thread), keeps the K x M sums in registers, and the CTA folds them with warp shuffles. The per-CTA partial sums
// are finished exactly as in gemm_inner.cu (last CTA adds them in CTA order; run-to-run reproducible).
#include <algorithm>

#include "common.cuh"
#include "gi_finalize.cuh"

namespace itsolv {

void fill_finalize(itsolv_ctx* ctx, int grid_bound, int km, GiFinalize* f, bool* host_direct); // gemm_inner.cu
int launch_reduce_partials(itsolv_ctx* ctx, int grid, int km);                             // gemm_inner.cu
int finish_with_peers(itsolv_ctx* ctx, int km, bool* host_direct);                         // gemm_inner.cu

constexpr int kDirectThreads = 256;
constexpr int kDirectUnroll = 4;

struct GdParams {
  const double* x[4];
  const double* y[4];
  GiFinalize fin;
  size_t n;
  int k, m;    // actual panel (<= K, M of the instantiation; missing vectors repeat the first one, results are dropped)
  int same_xy; // 1 x 1 with x == y: load the vector once
};

template <int K, int M>
__device__ __forceinline__ void fma_pair(const double2 (&xv)[K], const double2 (&yv)[M], double (&acc)[K][M]) {
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < M; ++b) {
      acc[a][b] = fma(xv[a].x, yv[b].x, acc[a][b]);
      acc[a][b] = fma(xv[a].y, yv[b].y, acc[a][b]);
    }
}

template <int K, int M, bool VEC, bool SAME>
__global__ void __launch_bounds__(kDirectThreads, 4) gemm_inner_direct_kernel(const __grid_constant__ GdParams p) {
  __shared__ double s_part[kDirectThreads / 32][K * M];
  __shared__ int s_is_last;
  double acc[K][M];
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < M; ++b)
      acc[a][b] = 0.0;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = size_t(gridDim.x) * blockDim.x;
  if (VEC) {
    const size_t npairs = p.n / 2;
    size_t r = tid;
    for (; r + (kDirectUnroll - 1) * nthreads < npairs; r += kDirectUnroll * nthreads) {
      double2 xv[kDirectUnroll][K], yv[kDirectUnroll][M];
#pragma unroll
      for (int u = 0; u < kDirectUnroll; ++u) {
#pragma unroll
        for (int a = 0; a < K; ++a)
          xv[u][a] = reinterpret_cast<const double2*>(p.x[a])[r + u * nthreads];
        if constexpr (!SAME) {
#pragma unroll
          for (int b = 0; b < M; ++b)
            yv[u][b] = reinterpret_cast<const double2*>(p.y[b])[r + u * nthreads];
        }
      }
#pragma unroll
      for (int u = 0; u < kDirectUnroll; ++u) {
        if constexpr (SAME)
          fma_pair<K, K>(xv[u], xv[u], acc);
        else
          fma_pair<K, M>(xv[u], yv[u], acc);
      }
    }
    for (; r < npairs; r += nthreads) {
      double2 xv[K], yv[M];
#pragma unroll
      for (int a = 0; a < K; ++a)
        xv[a] = reinterpret_cast<const double2*>(p.x[a])[r];
#pragma unroll
      for (int b = 0; b < M; ++b)
        yv[b] = SAME ? xv[0] : reinterpret_cast<const double2*>(p.y[b])[r];
      fma_pair<K, M>(xv, yv, acc);
    }
    if ((p.n & 1) && tid == 0) {
      const size_t i = p.n - 1;
#pragma unroll
      for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = 0; b < M; ++b)
          acc[a][b] = fma(p.x[a][i], p.y[b][i], acc[a][b]);
    }
  } else {
    for (size_t i = tid; i < p.n; i += nthreads) {
#pragma unroll
      for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = 0; b < M; ++b)
          acc[a][b] = fma(p.x[a][i], p.y[b][i], acc[a][b]);
    }
  }
  // warp tree, then the warps of the CTA in warp order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < M; ++b) {
      double v = acc[a][b];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, off);
      if (lane == 0)
        s_part[warp][a * M + b] = v;
    }
  __syncthreads();
  if (threadIdx.x < K * M) {
    const int a = threadIdx.x / M, b = threadIdx.x % M;
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < kDirectThreads / 32; ++w)
      sum += s_part[w][threadIdx.x];
    if (a < p.k && b < p.m)
      p.fin.partials[size_t(blockIdx.x) * (p.k * p.m) + a * p.m + b] = sum;
  }
  gi_finalize(p.fin, p.k * p.m, &s_is_last);
}

using GdKernel = void (*)(const GdParams);

template <int K, int M>
static GdKernel gd_pick(bool vec, bool same) {
  if constexpr (K == 1 && M == 1) {
    if (same)
      return vec ? gemm_inner_direct_kernel<1, 1, true, true> : gemm_inner_direct_kernel<1, 1, false, true>;
  }
  return vec ? gemm_inner_direct_kernel<K, M, true, false> : gemm_inner_direct_kernel<K, M, false, false>;
}

int gemm_inner_direct_device(itsolv_ctx* ctx, const double* const* xx, int k, const double* const* yy, int m, size_t n,
                             bool* host_direct, bool* handled) {
  *handled = false;
  if (k * m > 4 || ctx->opt_gi_direct == 2)
    return 0;
  GdParams p;
  bool vec = true;
  for (int a = 0; a < 4; ++a) {
    p.x[a] = xx[a < k ? a : 0];
    p.y[a] = yy[a < m ? a : 0];
    vec = vec && aligned16(p.x[a]) && aligned16(p.y[a]);
  }
  p.n = n;
  p.k = k;
  p.m = m;
  p.same_xy = (k == 1 && m == 1 && xx[0] == yy[0]) ? 1 : 0;
  const int K = k <= 1 ? 1 : (k <= 2 ? 2 : 4), M = m <= 1 ? 1 : (m <= 2 ? 2 : 4);
  GdKernel kernel = nullptr;
  if (K == 1 && M == 1)
    kernel = gd_pick<1, 1>(vec, p.same_xy != 0);
  else if (K == 2 && M == 1)
    kernel = gd_pick<2, 1>(vec, false);
  else if (K == 1 && M == 2)
    kernel = gd_pick<1, 2>(vec, false);
  else if (K == 4 && M == 1)
    kernel = gd_pick<4, 1>(vec, false);
  else if (K == 1 && M == 4)
    kernel = gd_pick<1, 4>(vec, false);
  else if (K == 2 && M == 2)
    kernel = gd_pick<2, 2>(vec, false);
  if (!kernel)
    return 0;
  const int per_sm = ctx->opt_gi_direct_ctas > 0 ? ctx->opt_gi_direct_ctas : 4;
  const size_t units = vec ? n / 2 : n;
  const size_t want = (units + size_t(kDirectThreads) * kDirectUnroll - 1) / (size_t(kDirectThreads) * kDirectUnroll);
  const int grid = int(std::max<size_t>(1, std::min<size_t>(want, size_t(ctx->num_sms) * per_sm)));
  const int km = k * m;
  if (ensure_partials(ctx, size_t(grid) * km))
    return 1;
  fill_finalize(ctx, ctx->num_sms * per_sm, km, &p.fin, host_direct);
  mark_launch(ctx);
  kernel<<<grid, kDirectThreads, 0, ctx->stream>>>(p);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  if (!p.fin.fused) {
    if (launch_reduce_partials(ctx, grid, km))
      return 1;
    if (finish_with_peers(ctx, km, host_direct))
      return 1;
  }
  *handled = true;
  return 0;
}

} // namespace itsolv
