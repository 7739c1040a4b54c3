// One pivot step of the R-R modified Gram-Schmidt (reference src/molpro/linalg/itsolv/propose_rspace.h:451-463) that
// also delivers the inner products the NEXT pivot step needs.
//
// The reference alternates, for pivot i = 0..w-1:  norm_i = sqrt(<r_i, r_i>)          dot
//                                                   r_i *= 1/norm_i                    scal
//                                                   for j > i: ov = <r_i, r_j>; r_j -= ov r_i     (w-i-1) x (dot + axpy)
// i.e. every step reads the vectors once for the products and once more for the update. Here one pass does the update of
// step i (r_i scaled; r_j + (-ov_j) r_i with product and sum rounded separately, as scal and axpy do, so the vectors are
// bit-identical) and, on the values it has just formed, accumulates
//   dots[0]     = <r_i', r_i'>                      (what the reference's closing normalise() asks for)
//   dots[1 + t] = <r_{i+1}', r_{i+1+t}'>, t >= 0    (norm and overlaps of the next pivot with all later vectors)
// so the next step needs no pass of its own for them: w(w+1) vector passes for the whole R-R part instead of 2w(w+1).
// The sums are finished like the Gram kernels' (last CTA adds the per-CTA partials in CTA order, all-reduce over NVLink
// peer memory, result written into mapped host memory).
#include <algorithm>

#include "common.cuh"
#include "gi_finalize.cuh"

namespace itsolv {

void fill_finalize(itsolv_ctx* ctx, int grid_bound, int km, GiFinalize* f, bool* host_direct); // gemm_inner.cu
int launch_reduce_partials(itsolv_ctx* ctx, int grid, int km);                             // gemm_inner.cu
int finish_with_peers(itsolv_ctx* ctx, int km, bool* host_direct);                         // gemm_inner.cu
int finish_result(itsolv_ctx* ctx, int count, double* out, bool host_direct);              // gemm_inner.cu

constexpr int kMfThreads = 256;
constexpr int kMfMaxLater = 16; // later vectors per launch

struct MfParams {
  double* ri;
  double* rj[kMfMaxLater];
  double neg_ov[kMfMaxLater];
  double inv_norm;
  GiFinalize fin;
  size_t n;
  int m;
};

//! U row groups of one thread: all loads first, then the arithmetic, then all stores
template <int M, int U, class RV>
__device__ __forceinline__ void mf_rows(const MfParams& p, const size_t (&r)[U], int nu, double (&dots)[M + 1]) {
  RV v[U], y[U][M > 0 ? M : 1];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (u < nu) {
      v[u] = reinterpret_cast<const RV*>(p.ri)[r[u]];
#pragma unroll
      for (int k = 0; k < M; ++k)
        if (k < p.m)
          y[u][k] = reinterpret_cast<const RV*>(p.rj[k])[r[u]];
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (u < nu) {
      if constexpr (sizeof(RV) == sizeof(double2)) {
        double2& a = reinterpret_cast<double2&>(v[u]);
        a.x = __dmul_rn(a.x, p.inv_norm);
        a.y = __dmul_rn(a.y, p.inv_norm);
        dots[0] = fma(a.x, a.x, dots[0]);
        dots[0] = fma(a.y, a.y, dots[0]);
#pragma unroll
        for (int k = 0; k < M; ++k)
          if (k < p.m) {
            double2& b = reinterpret_cast<double2&>(y[u][k]);
            b.x = __dadd_rn(b.x, __dmul_rn(p.neg_ov[k], a.x));
            b.y = __dadd_rn(b.y, __dmul_rn(p.neg_ov[k], a.y));
            const double2& b0 = reinterpret_cast<const double2&>(y[u][0]);
            dots[1 + k] = fma(b0.x, b.x, dots[1 + k]);
            dots[1 + k] = fma(b0.y, b.y, dots[1 + k]);
          }
      } else {
        double& a = reinterpret_cast<double&>(v[u]);
        a = __dmul_rn(a, p.inv_norm);
        dots[0] = fma(a, a, dots[0]);
#pragma unroll
        for (int k = 0; k < M; ++k)
          if (k < p.m) {
            double& b = reinterpret_cast<double&>(y[u][k]);
            b = __dadd_rn(b, __dmul_rn(p.neg_ov[k], a));
            dots[1 + k] = fma(reinterpret_cast<const double&>(y[u][0]), b, dots[1 + k]);
          }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (u < nu) {
      reinterpret_cast<RV*>(p.ri)[r[u]] = v[u];
#pragma unroll
      for (int k = 0; k < M; ++k)
        if (k < p.m)
          reinterpret_cast<RV*>(p.rj[k])[r[u]] = y[u][k];
    }
  }
}

template <int M, bool VEC>
__global__ void __launch_bounds__(kMfThreads, M <= 4 ? 4 : 2) mgs_step_dots_kernel(const __grid_constant__ MfParams p) {
  constexpr int U = M <= 1 ? 4 : (M <= 4 ? 2 : 1);
  __shared__ double s_part[kMfThreads / 32][M + 1];
  __shared__ int s_is_last;
  double dots[M + 1];
#pragma unroll
  for (int k = 0; k <= M; ++k)
    dots[k] = 0.0;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = size_t(gridDim.x) * blockDim.x;
  if (VEC) {
    const size_t npairs = p.n / 2;
    for (size_t r0 = tid; r0 < npairs; r0 += U * nthreads) {
      size_t r[U];
      int nu = 0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        r[u] = r0 + u * nthreads;
        if (r[u] < npairs)
          nu = u + 1;
      }
      mf_rows<M, U, double2>(p, r, nu, dots);
    }
    if ((p.n & 1) && tid == 0) {
      const size_t r[1] = {p.n - 1};
      mf_rows<M, 1, double>(p, r, 1, dots);
    }
  } else {
    for (size_t r0 = tid; r0 < p.n; r0 += U * nthreads) {
      size_t r[U];
      int nu = 0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        r[u] = r0 + u * nthreads;
        if (r[u] < p.n)
          nu = u + 1;
      }
      mf_rows<M, U, double>(p, r, nu, dots);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k <= M; ++k) {
    double s = dots[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
      s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0)
      s_part[warp][k] = s;
  }
  __syncthreads();
  if (threadIdx.x <= p.m) {
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < kMfThreads / 32; ++w)
      sum += s_part[w][threadIdx.x];
    p.fin.partials[size_t(blockIdx.x) * (p.m + 1) + threadIdx.x] = sum;
  }
  gi_finalize(p.fin, p.m + 1, &s_is_last);
}

using MfKernel = void (*)(const MfParams);
template <int M>
static MfKernel mf_pick_vec(bool vec) {
  return vec ? mgs_step_dots_kernel<M, true> : mgs_step_dots_kernel<M, false>;
}
static MfKernel mf_pick(int mt, bool vec) {
  switch (mt) {
  case 0:
    return mf_pick_vec<0>(vec);
  case 1:
    return mf_pick_vec<1>(vec);
  case 2:
    return mf_pick_vec<2>(vec);
  case 4:
    return mf_pick_vec<4>(vec);
  case 8:
    return mf_pick_vec<8>(vec);
  case 16:
    return mf_pick_vec<16>(vec);
  }
  return nullptr;
}

} // namespace itsolv

using namespace itsolv;

extern "C" {

int itsolv_mgs_step_dots_f64(itsolv_ctx* ctx, double inv_norm, double* ri, const double* ov, double* const* rj, int m,
                             size_t n, double* dots) {
  ITSOLV_REQUIRE(m >= 0 && m <= kMfMaxLater, "itsolv_mgs_step_dots_f64: at most 16 later vectors per call");
  ITSOLV_REQUIRE(ri != nullptr && dots != nullptr && (m == 0 || (ov != nullptr && rj != nullptr)),
                 "itsolv_mgs_step_dots_f64: null argument");
  ctx->counters.n_scal += 1;
  ctx->counters.n_axpy += m;
  ctx->counters.n_dot += m + 1;
  CallScope scope(ctx, OP_BLAS1, 16.0 * double(n) * (m + 1));
  bool direct = false;
  const int km = m + 1;
  if (n == 0) { // an empty shard still takes part in the all-reduce
    ITSOLV_CUDA(cudaMemsetAsync(ctx->d_result, 0, size_t(km) * sizeof(double), ctx->stream));
    if (finish_with_peers(ctx, km, &direct))
      return 1;
  } else {
    MfParams p;
    bool vec = aligned16(ri);
    for (int k = 0; k < kMfMaxLater; ++k) {
      p.rj[k] = k < m ? rj[k] : nullptr;
      p.neg_ov[k] = k < m ? -ov[k] : 0.0;
      if (k < m) {
        ITSOLV_REQUIRE(rj[k] != ri, "itsolv_mgs_step_dots_f64: a target aliases the pivot vector");
        for (int k2 = 0; k2 < k; ++k2)
          ITSOLV_REQUIRE(rj[k] != rj[k2], "itsolv_mgs_step_dots_f64: the same target twice");
        vec = vec && aligned16(rj[k]);
      }
    }
    p.ri = ri;
    p.inv_norm = inv_norm;
    p.n = n;
    p.m = m;
    int mt = 0;
    if (m > 0) {
      mt = 1;
      while (mt < m)
        mt *= 2;
    }
    MfKernel kernel = mf_pick(mt, vec);
    ITSOLV_REQUIRE(kernel != nullptr, "itsolv_mgs_step_dots_f64: vector count not instantiated");
    const int per_sm = mt <= 4 ? 4 : 2;
    const int unroll = mt <= 1 ? 4 : (mt <= 4 ? 2 : 1);
    const size_t units = vec ? n / 2 : n;
    const size_t want = (units + size_t(kMfThreads) * unroll - 1) / (size_t(kMfThreads) * unroll);
    const int grid = int(std::max<size_t>(1, std::min<size_t>(want, size_t(ctx->num_sms) * per_sm)));
    if (ensure_partials(ctx, size_t(grid) * km))
      return 1;
    fill_finalize(ctx, ctx->num_sms * per_sm, km, &p.fin, &direct);
    mark_launch(ctx);
    kernel<<<grid, kMfThreads, 0, ctx->stream>>>(p);
    ITSOLV_CUDA(cudaGetLastError());
    ctx->counters.launches += 1;
    if (!p.fin.fused) {
      if (launch_reduce_partials(ctx, grid, km))
        return 1;
      if (finish_with_peers(ctx, km, &direct))
        return 1;
    }
  }
  scope.stop();
  return finish_result(ctx, km, dots, direct);
}

} // extern "C"
