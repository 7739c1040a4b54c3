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
constexpr int kMgsChainDefault = 1; // option MGS_CHAIN is added to this: > 0 chains the steps on the device (MGS_CHAIN=-1: off)

struct MfParams {
  double* ri;
  double* rj[kMfMaxLater];
  double neg_ov[kMfMaxLater];
  double inv_norm;
  // chained steps: {1/|r_i|, null flag, -ov...} left in device memory by the previous launch's tail (GiFinalize::chain_out)
  // replace inv_norm / neg_ov; with the null flag set nothing is stored (the reference skips a null pivot,
  // propose_rspace.h:451-463) but the products of the untouched vectors are still returned
  const double* chain_in;
  GiFinalize fin;
  size_t n;
  int m;
};

//! the coefficients of one step, in registers
template <int M>
struct MfCoef {
  double inv_norm;
  double neg_ov[M > 0 ? M : 1];
  bool store;
};

//! U row groups of one thread: all loads first, then the arithmetic, then all stores
template <int M, int U, class RV>
__device__ __forceinline__ void mf_rows(const MfParams& p, const MfCoef<M>& c, const size_t (&r)[U], int nu,
                                        double (&dots)[M + 1]) {
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
        a.x = __dmul_rn(a.x, c.inv_norm);
        a.y = __dmul_rn(a.y, c.inv_norm);
        dots[0] = fma(a.x, a.x, dots[0]);
        dots[0] = fma(a.y, a.y, dots[0]);
#pragma unroll
        for (int k = 0; k < M; ++k)
          if (k < p.m) {
            double2& b = reinterpret_cast<double2&>(y[u][k]);
            b.x = __dadd_rn(b.x, __dmul_rn(c.neg_ov[k], a.x));
            b.y = __dadd_rn(b.y, __dmul_rn(c.neg_ov[k], a.y));
            const double2& b0 = reinterpret_cast<const double2&>(y[u][0]);
            dots[1 + k] = fma(b0.x, b.x, dots[1 + k]);
            dots[1 + k] = fma(b0.y, b.y, dots[1 + k]);
          }
      } else {
        double& a = reinterpret_cast<double&>(v[u]);
        a = __dmul_rn(a, c.inv_norm);
        dots[0] = fma(a, a, dots[0]);
#pragma unroll
        for (int k = 0; k < M; ++k)
          if (k < p.m) {
            double& b = reinterpret_cast<double&>(y[u][k]);
            b = __dadd_rn(b, __dmul_rn(c.neg_ov[k], a));
            dots[1 + k] = fma(reinterpret_cast<const double&>(y[u][0]), b, dots[1 + k]);
          }
      }
    }
  }
  if (!c.store)
    return;
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
  MfCoef<M> c;
  c.inv_norm = p.chain_in ? p.chain_in[0] : p.inv_norm;
  c.store = p.chain_in ? p.chain_in[1] == 0.0 : true;
#pragma unroll
  for (int k = 0; k < M; ++k)
    c.neg_ov[k] = k < p.m ? (p.chain_in ? p.chain_in[2 + k] : p.neg_ov[k]) : 0.0;
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
      mf_rows<M, U, double2>(p, c, r, nu, dots);
    }
    if ((p.n & 1) && tid == 0) {
      const size_t r[1] = {p.n - 1};
      mf_rows<M, 1, double>(p, c, r, 1, dots);
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
      mf_rows<M, U, double>(p, c, r, nu, dots);
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

namespace itsolv {

int gemm_inner_device(itsolv_ctx* ctx, const double* const* xx, int k, const double* const* yy, int m, size_t n,
                      bool* host_direct); // gemm_inner.cu
bool gemm_outer_dots_supported(int k, int m); // gemm_outer.cu
int gemm_outer_with_dots(itsolv_ctx* ctx, const double* alpha, int k, int m, const double* const* xx, double* const* yy,
                         size_t n, const double* yscale, int pivot, const int* cols, int ncols, bool* direct); // gemm_outer.cu
int wait_host_result(itsolv_ctx* ctx);   // gemm_inner.cu: waits for the newest sequence word

//! launch one step; the sums are delivered as fill_finalize() decides (no waiting here)
static int mgs_step_launch(itsolv_ctx* ctx, double inv_norm, double* ri, const double* ov, double* const* rj, int m,
                           size_t n, const double* chain_in, bool* direct) {
  MfParams p;
  bool vec = aligned16(ri);
  for (int k = 0; k < kMfMaxLater; ++k) {
    p.rj[k] = k < m ? rj[k] : nullptr;
    p.neg_ov[k] = (k < m && ov) ? -ov[k] : 0.0;
    if (k < m) {
      ITSOLV_REQUIRE(rj[k] != ri, "mgs step: a target aliases the pivot vector");
      for (int k2 = 0; k2 < k; ++k2)
        ITSOLV_REQUIRE(rj[k] != rj[k2], "mgs step: the same target twice");
      vec = vec && aligned16(rj[k]);
    }
  }
  p.ri = ri;
  p.inv_norm = inv_norm;
  p.chain_in = chain_in;
  p.n = n;
  p.m = m;
  int mt = 0;
  if (m > 0) {
    mt = 1;
    while (mt < m)
      mt *= 2;
  }
  MfKernel kernel = mf_pick(mt, vec);
  ITSOLV_REQUIRE(kernel != nullptr, "mgs step: vector count not instantiated");
  const int per_sm = mt <= 4 ? 4 : 2;
  const int unroll = mt <= 1 ? 4 : (mt <= 4 ? 2 : 1);
  const size_t units = vec ? n / 2 : n;
  const size_t want = (units + size_t(kMfThreads) * unroll - 1) / (size_t(kMfThreads) * unroll);
  const int grid = int(std::max<size_t>(1, std::min<size_t>(want, size_t(ctx->num_sms) * per_sm)));
  const int km = m + 1;
  if (ensure_partials(ctx, size_t(grid) * km))
    return 1;
  fill_finalize(ctx, ctx->num_sms * per_sm, km, &p.fin, direct);
  mark_launch(ctx);
  kernel<<<grid, kMfThreads, 0, ctx->stream>>>(p);
  ITSOLV_CUDA(cudaGetLastError());
  ctx->counters.launches += 1;
  if (!p.fin.fused) {
    if (launch_reduce_partials(ctx, grid, km))
      return 1;
    if (finish_with_peers(ctx, km, direct))
      return 1;
  }
  return 0;
}

} // namespace itsolv

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
  } else if (mgs_step_launch(ctx, inv_norm, ri, ov, rj, m, n, nullptr, &direct)) {
    return 1;
  }
  scope.stop();
  return finish_result(ctx, km, dots, direct);
}

int itsolv_mgs_chain_supported(itsolv_ctx* ctx, int w, size_t n) {
  if (ctx->opt_mgs_chain + kMgsChainDefault <= 0 || w < 1 || w > kMfMaxLater + 1 || n == 0)
    return 0;
  GiPeers peers;
  // the sums must be delivered by the kernels themselves (single rank, or exchange buffers of all ranks mapped)
  return (itsolv_comm_size(ctx) == 1 || comm_peers(ctx, &peers)) ? 1 : 0;
}

int itsolv_mgs_chain_f64(itsolv_ctx* ctx, double* const* r, int w, size_t n, double thresh, double* rows) {
  ITSOLV_REQUIRE(itsolv_mgs_chain_supported(ctx, w, n), "itsolv_mgs_chain_f64: not available for this call");
  ITSOLV_REQUIRE(r != nullptr && rows != nullptr, "itsolv_mgs_chain_f64: null argument");
  for (int i = 0; i < w; ++i)
    for (int j = 0; j < i; ++j)
      ITSOLV_REQUIRE(r[i] != r[j], "itsolv_mgs_chain_f64: the same vector twice");
  ctx->counters.n_scal += w;
  ctx->counters.n_axpy += w * (w - 1) / 2;
  ctx->counters.n_dot += w * (w + 1) / 2 + w;
  // device scratch: coefficients of step i at chain + 32 i (upper part of the device result buffer)
  double* chain = ctx->d_result + 12288;
  size_t offset = 0; // of the next block of sums in the host result buffer
  bool direct = false;
  {
    // the first pivot's norm and overlaps with all later vectors: one Gram row, its tail prepares step 0
    CallScope scope(ctx, OP_GEMM_INNER, 8.0 * double(n) * w);
    ctx->result_offset = offset;
    ctx->chain_out = chain;
    ctx->chain_offset = 0;
    ctx->chain_count = w;
    ctx->chain_thresh = thresh;
    const double* x0 = r[0];
    if (gemm_inner_device(ctx, &x0, 1, r, w, n, &direct))
      return 1;
    ITSOLV_REQUIRE(direct, "itsolv_mgs_chain_f64: the Gram row was not delivered by the kernel");
    offset += size_t(w);
  }
  for (int i = 0; i < w; ++i) {
    const int m = w - i - 1;
    CallScope scope(ctx, OP_BLAS1, 16.0 * double(n) * (m + 1));
    ctx->result_offset = offset;
    if (m > 0) { // the sums of this step hold the next pivot's row from index 1 on
      ctx->chain_out = chain + 32 * (i + 1);
      ctx->chain_offset = 1;
      ctx->chain_count = m;
      ctx->chain_thresh = thresh;
    }
    if (mgs_step_launch(ctx, 1.0, r[i], nullptr, r + i + 1, m, n, chain + 32 * i, &direct))
      return 1;
    ITSOLV_REQUIRE(direct, "itsolv_mgs_chain_f64: the sums were not delivered by the kernel");
    offset += size_t(m + 1);
  }
  // one wait for the whole chain: the sequence word of the last launch implies all earlier ones (stream order)
  if (wait_host_result(ctx))
    return 1;
  for (size_t e = 0; e < offset; ++e)
    rows[e] = ctx->h_result[e];
  return 0;
}

int itsolv_project_mgs_chain_supported(itsolv_ctx* ctx, int k, int m, int w, size_t n) {
  return (ctx->opt_project_chain >= 0 && w >= 1 && w <= m && gemm_outer_dots_supported(k, m) &&
          itsolv_mgs_chain_supported(ctx, w, n))
             ? 1
             : 0;
}

int itsolv_project_mgs_chain_f64(itsolv_ctx* ctx, const double* alpha, int k, int m, const double* const* x,
                                 double* const* y, const double* yscale, const int* keep, int w, size_t n, double thresh,
                                 double* rows) {
  ITSOLV_REQUIRE(itsolv_project_mgs_chain_supported(ctx, k, m, w, n), "itsolv_project_mgs_chain_f64: not available for this call");
  ITSOLV_REQUIRE(alpha && x && y && keep && rows, "itsolv_project_mgs_chain_f64: null argument");
  for (int i = 0; i < w; ++i)
    ITSOLV_REQUIRE(keep[i] >= 0 && keep[i] < m && (i == 0 || keep[i] > keep[i - 1]),
                   "itsolv_project_mgs_chain_f64: the kept columns must ascend");
  double* r[kMfMaxLater + 1];
  for (int i = 0; i < w; ++i)
    r[i] = y[keep[i]];
  ctx->counters.n_scal += w;
  ctx->counters.n_axpy += w * (w - 1) / 2;
  ctx->counters.n_dot += w * (w + 1) / 2 + w;
  double* chain = ctx->d_result + 12288;
  size_t offset = 0;
  bool direct = false;
  {
    // the projection of all m new vectors; its tail returns the first pivot's norm and overlaps with the later kept
    // vectors (the Gram row that itsolv_mgs_chain_f64 spends a launch on) and prepares step 0
    ctx->result_offset = offset;
    ctx->chain_out = chain;
    ctx->chain_offset = 0;
    ctx->chain_count = w;
    ctx->chain_thresh = thresh;
    if (gemm_outer_with_dots(ctx, alpha, k, m, x, y, n, yscale, keep[0], keep, w, &direct))
      return 1;
    ITSOLV_REQUIRE(direct, "itsolv_project_mgs_chain_f64: the Gram row was not delivered by the kernel");
    offset += size_t(w);
  }
  for (int i = 0; i < w; ++i) {
    const int later = w - i - 1;
    CallScope scope(ctx, OP_BLAS1, 16.0 * double(n) * (later + 1));
    ctx->result_offset = offset;
    if (later > 0) {
      ctx->chain_out = chain + 32 * (i + 1);
      ctx->chain_offset = 1;
      ctx->chain_count = later;
      ctx->chain_thresh = thresh;
    }
    if (mgs_step_launch(ctx, 1.0, r[i], nullptr, r + i + 1, later, n, chain + 32 * i, &direct))
      return 1;
    ITSOLV_REQUIRE(direct, "itsolv_project_mgs_chain_f64: the sums were not delivered by the kernel");
    offset += size_t(later + 1);
  }
  if (wait_host_result(ctx))
    return 1;
  for (size_t e = 0; e < offset; ++e)
    rows[e] = ctx->h_result[e];
  return 0;
}

} // extern "C"
