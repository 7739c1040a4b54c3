// gemm_outer: subspace -> full-space expansion  y_j += sum_i alpha(i,j) * x_i   (k x m coefficients)
// (contract: reference src/molpro/linalg/array/ArrayHandler.h:195; CPU path array/util/gemm.h:186-203, 258-265, which
// performs k*m separate axpy sweeps = 3km vector passes; here x is read once per 16 output columns and every y is read
// and written once: 8n(k+2m) bytes, or 8n(k+m) with beta_zero when the caller has just zeroed y,
// reference itsolv/IterativeSolverTemplate.h:45-47).
//
// Rows are independent, so lanes map to consecutive row pairs (128-bit coalesced loads of every x_i and y_j); a
// thread keeps the m (<= 16 per pass) running sums of its two rows in registers, streams the k x-vectors through
// them with 4 independent 128-bit loads in flight, and reads alpha (staged once per CTA into shared memory) with
// warp-uniform broadcast loads. The sum over i runs in ascending i as the reference's loop does; each term is one FMA.
#include <algorithm>

#include "common.cuh"
#include "gi_finalize.cuh"

namespace itsolv {

void fill_finalize(itsolv_ctx* ctx, int grid_bound, int km, GiFinalize* f, bool* host_direct); // gemm_inner.cu
int launch_reduce_partials(itsolv_ctx* ctx, int grid, int km);                             // gemm_inner.cu
int finish_with_peers(itsolv_ctx* ctx, int km, bool* host_direct);                         // gemm_inner.cu

constexpr int kGoThreads = 256;
constexpr int kGoMaxDots = 8; // columns of the variant that also returns a row of inner products of what it wrote
constexpr int kGoUnroll = 4;

struct GoParams {
  const double* x[ITSOLV_MAX_PANEL];
  double* y[ITSOLV_MAX_PANEL];
  const double* alpha; // device, k x m row-major
  size_t n;
  int k, m;
  int ld; // leading dimension of alpha in shared memory (m rounded up to a multiple of MJ)
  int beta_zero;
  int scaled; // row k of the staged coefficients holds a factor per y: y_j is multiplied by it (rounded) before the sums
  // DOTS variant (m <= MJ <= kGoMaxDots): <y_pivot, y_col[t]> of the values written, t < ndots, finished like the sums of
  // the Gram kernels (last CTA, peer all-reduce, mapped host memory, optional chain coefficients; gi_finalize.cuh)
  int dot_pivot, ndots;
  int dot_col[kGoMaxDots];
  GiFinalize fin;
};

template <class RV>
struct RowOps;
template <>
struct RowOps<double2> {
  static constexpr int width = 2;
  static __device__ __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ void fma_to(double2& acc, double a, const double2& x) {
    acc.x = fma(a, x.x, acc.x);
    acc.y = fma(a, x.y, acc.y);
  }
  static __device__ __forceinline__ void scale(double2& acc, double s) {
    acc.x = __dmul_rn(acc.x, s);
    acc.y = __dmul_rn(acc.y, s);
  }
  static __device__ __forceinline__ void dot_to(double& d, const double2& a, const double2& b) {
    d = fma(a.x, b.x, d);
    d = fma(a.y, b.y, d);
  }
};
template <>
struct RowOps<double> {
  static constexpr int width = 1;
  static __device__ __forceinline__ double zero() { return 0.0; }
  static __device__ __forceinline__ void fma_to(double& acc, double a, const double& x) { acc = fma(a, x, acc); }
  static __device__ __forceinline__ void scale(double& acc, double s) { acc = __dmul_rn(acc, s); }
  static __device__ __forceinline__ void dot_to(double& d, const double& a, const double& b) { d = fma(a, b, d); }
};

//! MJ consecutive coefficients of one alpha row; the address is warp-uniform (broadcast) and 16-byte aligned for MJ >= 2
template <int MJ>
__device__ __forceinline__ void load_alpha_row(const double* __restrict__ arow, double (&av)[MJ]) {
  if constexpr (MJ >= 2) {
#pragma unroll
    for (int b = 0; b < MJ / 2; ++b) {
      const double2 t = reinterpret_cast<const double2*>(arow)[b];
      av[2 * b] = t.x;
      av[2 * b + 1] = t.y;
    }
  } else {
    av[0] = arow[0];
  }
}

//! one thread, one row group `r` (index in units of RV), all column chunks
template <int MJ, class RV, bool SCALED, bool DOTS>
__device__ __forceinline__ void expand_rows(const GoParams& p, const double* __restrict__ sa, size_t r, double (&dots)[MJ]) {
  using Ops = RowOps<RV>;
  for (int jc = 0; jc < p.m; jc += MJ) {
    RV acc[MJ];
#pragma unroll
    for (int b = 0; b < MJ; ++b) {
      if (!p.beta_zero && jc + b < p.m)
        acc[b] = reinterpret_cast<const RV*>(p.y[jc + b])[r];
      else
        acc[b] = Ops::zero();
    }
    if constexpr (SCALED) { // compile-time: a run-time branch here would separate the loads of y from those of x
      double sv[MJ];
      load_alpha_row<MJ>(sa + size_t(p.k) * p.ld + jc, sv);
#pragma unroll
      for (int b = 0; b < MJ; ++b)
        Ops::scale(acc[b], sv[b]);
    }
    int i = 0;
    for (; i + kGoUnroll <= p.k; i += kGoUnroll) {
      RV xv[kGoUnroll];
#pragma unroll
      for (int u = 0; u < kGoUnroll; ++u)
        xv[u] = reinterpret_cast<const RV*>(p.x[i + u])[r];
#pragma unroll
      for (int u = 0; u < kGoUnroll; ++u) {
        double av[MJ];
        load_alpha_row<MJ>(sa + size_t(i + u) * p.ld + jc, av);
#pragma unroll
        for (int b = 0; b < MJ; ++b)
          Ops::fma_to(acc[b], av[b], xv[u]);
      }
    }
    for (; i < p.k; ++i) {
      const RV xv = reinterpret_cast<const RV*>(p.x[i])[r];
      double av[MJ];
      load_alpha_row<MJ>(sa + size_t(i) * p.ld + jc, av);
#pragma unroll
      for (int b = 0; b < MJ; ++b)
        Ops::fma_to(acc[b], av[b], xv);
    }
#pragma unroll
    for (int b = 0; b < MJ; ++b)
      if (jc + b < p.m)
        reinterpret_cast<RV*>(p.y[jc + b])[r] = acc[b];
    if constexpr (DOTS) { // one column chunk (m <= MJ): products of the pivot column with every column, from the written values
      RV first = acc[0];
#pragma unroll
      for (int b = 1; b < MJ; ++b)
        if (b == p.dot_pivot)
          first = acc[b];
#pragma unroll
      for (int b = 0; b < MJ; ++b)
        Ops::dot_to(dots[b], first, acc[b]);
    }
  }
}

template <int MJ, bool VEC, bool SCALED, bool DOTS>
__global__ void __launch_bounds__(kGoThreads, 2) gemm_outer_kernel(const __grid_constant__ GoParams p) {
  extern __shared__ __align__(16) double sa[]; // k (+1 when scaled) x ld, zero padded columns
  for (int e = threadIdx.x; e < (p.k + p.scaled) * p.ld; e += blockDim.x) {
    const int i = e / p.ld, j = e % p.ld;
    sa[e] = j < p.m ? p.alpha[size_t(i) * p.m + j] : 0.0;
  }
  __syncthreads();
  double dots[MJ];
#pragma unroll
  for (int b = 0; b < MJ; ++b)
    dots[b] = 0.0;
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = size_t(gridDim.x) * blockDim.x;
  if (VEC) {
    const size_t npairs = p.n / 2;
    for (size_t r = tid; r < npairs; r += nthreads)
      expand_rows<MJ, double2, SCALED, DOTS>(p, sa, r, dots);
    if ((p.n & 1) && tid == 0)
      expand_rows<MJ, double, SCALED, DOTS>(p, sa, p.n - 1, dots);
  } else {
    for (size_t r = tid; r < p.n; r += nthreads)
      expand_rows<MJ, double, SCALED, DOTS>(p, sa, r, dots);
  }
  if constexpr (DOTS) {
    __shared__ double s_part[kGoThreads / 32][MJ];
    __shared__ int s_is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int b = 0; b < MJ; ++b) {
      double v = dots[b];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, off);
      if (lane == 0)
        s_part[warp][b] = v;
    }
    __syncthreads();
    if (threadIdx.x < p.ndots) {
      const int col = p.dot_col[threadIdx.x];
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < kGoThreads / 32; ++w)
        sum += s_part[w][col];
      p.fin.partials[size_t(blockIdx.x) * p.ndots + threadIdx.x] = sum;
    }
    gi_finalize(p.fin, p.ndots, &s_is_last);
  }
}

//! request for the DOTS variant: <y_pivot, y_cols[t]> of the written values, delivered as fill_finalize() decides
struct GoDots {
  int pivot, ncols;
  const int* cols;
  bool* direct;
};

using GoKernel = void (*)(const GoParams);
template <int MJ>
static GoKernel go_pick_vec(bool vec, bool scaled) {
  if (scaled)
    return vec ? gemm_outer_kernel<MJ, true, true, false> : gemm_outer_kernel<MJ, false, true, false>;
  return vec ? gemm_outer_kernel<MJ, true, false, false> : gemm_outer_kernel<MJ, false, false, false>;
}
template <int MJ>
static GoKernel go_pick_dots(bool vec, bool scaled) {
  if constexpr (MJ <= kGoMaxDots) {
    if (scaled)
      return vec ? gemm_outer_kernel<MJ, true, true, true> : gemm_outer_kernel<MJ, false, true, true>;
    return vec ? gemm_outer_kernel<MJ, true, false, true> : gemm_outer_kernel<MJ, false, false, true>;
  }
  return nullptr;
}
static GoKernel go_pick(int mj, bool vec, bool scaled, bool dots = false) {
  switch (mj) {
  case 1:
    return dots ? go_pick_dots<1>(vec, scaled) : go_pick_vec<1>(vec, scaled);
  case 2:
    return dots ? go_pick_dots<2>(vec, scaled) : go_pick_vec<2>(vec, scaled);
  case 4:
    return dots ? go_pick_dots<4>(vec, scaled) : go_pick_vec<4>(vec, scaled);
  case 8:
    return dots ? go_pick_dots<8>(vec, scaled) : go_pick_vec<8>(vec, scaled);
  case 16:
    return dots ? nullptr : go_pick_vec<16>(vec, scaled);
  }
  return nullptr;
}

} // namespace itsolv

using namespace itsolv;

extern "C" {

static int gemm_outer_impl(itsolv_ctx* ctx, const double* alpha, int k, int m, const double* const* xx,
                           double* const* yy, size_t n, int beta_zero, const double* yscale, const GoDots* dots = nullptr);

int itsolv_gemm_outer_f64(itsolv_ctx* ctx, const double* alpha, int k, int m, const double* const* xx,
                          double* const* yy, size_t n, int beta_zero) {
  return gemm_outer_impl(ctx, alpha, k, m, xx, yy, n, beta_zero, nullptr);
}

int itsolv_gemm_outer_scaled_f64(itsolv_ctx* ctx, const double* alpha, int k, int m, const double* const* xx,
                                 double* const* yy, size_t n, const double* yscale) {
  ITSOLV_REQUIRE(yscale != nullptr, "itsolv_gemm_outer_scaled_f64: null scale factors");
  if (k <= 0) { // nothing to add: the scaling alone
    ctx->counters.n_gemm_outer++;
    return m > 0 ? itsolv_scal_batch_f64(ctx, yscale, yy, m, n) : 0;
  }
  ctx->counters.n_scal += m > 0 ? m : 0;
  return gemm_outer_impl(ctx, alpha, k, m, xx, yy, n, 0, yscale);
}

static int gemm_outer_impl(itsolv_ctx* ctx, const double* alpha, int k, int m, const double* const* xx,
                           double* const* yy, size_t n, int beta_zero, const double* yscale, const GoDots* dots) {
  ctx->counters.n_gemm_outer++;
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  if (m <= 0)
    return 0;
  if (k <= 0) {
    if (beta_zero)
      for (int j = 0; j < m; ++j)
        if (itsolv_fill_f64(ctx, 0.0, yy[j], n))
          return 1;
    return 0;
  }
  if (n == 0)
    return 0;
  // Aliased operands (a y that is also an x, or the same y twice) have sequential meaning in the reference's loop of
  // axpys; keep that meaning by issuing the axpys one by one.
  bool aliased = false;
  for (int j = 0; j < m && !aliased; ++j) {
    for (int i = 0; i < k; ++i)
      if (xx[i] == yy[j])
        aliased = true;
    for (int j2 = 0; j2 < j; ++j2)
      if (yy[j2] == yy[j])
        aliased = true;
  }
  if (aliased) {
    if (yscale && itsolv_scal_batch_f64(ctx, yscale, yy, m, n))
      return 1;
    if (beta_zero)
      for (int j = 0; j < m; ++j)
        if (itsolv_fill_f64(ctx, 0.0, yy[j], n))
          return 1;
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < m; ++j)
        if (itsolv_axpy_f64(ctx, alpha[size_t(i) * m + j], xx[i], yy[j], n))
          return 1;
    return 0;
  }
  // blocks of at most ITSOLV_MAX_PANEL x ITSOLV_MAX_PANEL coefficients per launch
  for (int j0 = 0; j0 < m; j0 += ITSOLV_MAX_PANEL) {
    const int mb = std::min(ITSOLV_MAX_PANEL, m - j0);
    for (int i0 = 0; i0 < k; i0 += ITSOLV_MAX_PANEL) {
      const int kb = std::min(ITSOLV_MAX_PANEL, k - i0);
      const bool bz = beta_zero && i0 == 0;
      CallScope scope(ctx, OP_GEMM_OUTER, 8.0 * double(n) * (kb + (bz ? 1.0 : 2.0) * mb));
      GoParams p;
      bool vec = true;
      for (int i = 0; i < kb; ++i) {
        p.x[i] = xx[i0 + i];
        vec = vec && aligned16(p.x[i]);
      }
      for (int j = 0; j < mb; ++j) {
        p.y[j] = yy[j0 + j];
        vec = vec && aligned16(p.y[j]);
      }
      char *h = nullptr, *d = nullptr;
      int slot = 0;
      const bool scaled = yscale != nullptr && i0 == 0; // the factor is applied with the first block of x vectors
      const size_t abytes = size_t(kb + (scaled ? 1 : 0)) * mb * sizeof(double);
      if (stage_acquire(ctx, abytes, &h, &d, &slot))
        return 1;
      double* ha = reinterpret_cast<double*>(h);
      for (int i = 0; i < kb; ++i)
        for (int j = 0; j < mb; ++j)
          ha[size_t(i) * mb + j] = alpha[size_t(i0 + i) * m + (j0 + j)];
      if (scaled)
        for (int j = 0; j < mb; ++j)
          ha[size_t(kb) * mb + j] = yscale[j0 + j];
      if (stage_commit(ctx, slot, abytes))
        return 1;
      p.alpha = reinterpret_cast<const double*>(d);
      p.n = n;
      p.k = kb;
      p.m = mb;
      p.beta_zero = bz ? 1 : 0;
      p.scaled = scaled ? 1 : 0;
      int mj = 1;
      while (mj < mb && mj < 16)
        mj *= 2;
      if (ctx->opt_go_cols > 0 && !dots)
        mj = ctx->opt_go_cols;
      p.ld = ((mb + mj - 1) / mj) * mj;
      GoKernel kernel = go_pick(mj, vec, scaled, dots != nullptr);
      ITSOLV_REQUIRE(kernel != nullptr, "gemm_outer: column tile not instantiated");
      const size_t smem = size_t(kb + p.scaled) * p.ld * sizeof(double);
      if (ensure_dynamic_smem(ctx, reinterpret_cast<const void*>(kernel), smem))
        return 1;
      int per_sm = ctx->opt_go_ctas > 0 ? ctx->opt_go_ctas : (mj <= 2 ? 6 : mj == 4 ? 4 : mj == 8 ? 3 : 2);
      if (smem * per_sm > size_t(ctx->max_smem_optin))
        per_sm = std::max<int>(1, int(size_t(ctx->max_smem_optin) / smem));
      const size_t units = vec ? n / 2 : n;
      size_t grid = std::min<size_t>((units + kGoThreads - 1) / kGoThreads, size_t(ctx->num_sms) * per_sm);
      if (grid == 0)
        grid = 1;
      p.dot_pivot = p.ndots = 0;
      if (dots) { // one block (checked by the caller): the row of inner products is finished by this launch's tail
        p.dot_pivot = dots->pivot;
        p.ndots = dots->ncols;
        for (int t = 0; t < dots->ncols; ++t)
          p.dot_col[t] = dots->cols[t];
        if (ensure_partials(ctx, grid * size_t(dots->ncols)))
          return 1;
        fill_finalize(ctx, ctx->num_sms * 6, dots->ncols, &p.fin, dots->direct);
      }
      mark_launch(ctx);
      kernel<<<int(grid), kGoThreads, smem, ctx->stream>>>(p);
      ITSOLV_CUDA(cudaGetLastError());
      ctx->counters.launches += 1;
      if (stage_done(ctx, slot))
        return 1;
      if (dots && !p.fin.fused) {
        if (launch_reduce_partials(ctx, int(grid), dots->ncols))
          return 1;
        if (finish_with_peers(ctx, dots->ncols, dots->direct))
          return 1;
      }
    }
  }
  return 0;
}

} // extern "C"

namespace itsolv {

bool gemm_outer_dots_supported(int k, int m) { return k >= 1 && k <= ITSOLV_MAX_PANEL && m >= 1 && m <= kGoMaxDots; }

//! y_j = (yscale ? yscale_j * y_j : y_j) + sum_i alpha(i,j) x_i for all m columns, and from the written values the row
//! {<y_pivot, y_cols[t]>}: the projection of a working set together with the first Gram row of its R-R Gram-Schmidt
//! (mgs_fused.cu). The operands must not alias (the caller's working set never does).
int gemm_outer_with_dots(itsolv_ctx* ctx, const double* alpha, int k, int m, const double* const* xx, double* const* yy,
                         size_t n, const double* yscale, int pivot, const int* cols, int ncols, bool* direct) {
  ITSOLV_REQUIRE(gemm_outer_dots_supported(k, m) && n > 0, "gemm_outer_with_dots: shape not supported");
  ITSOLV_REQUIRE(ncols >= 1 && ncols <= m && pivot >= 0 && pivot < m, "gemm_outer_with_dots: invalid columns");
  for (int j = 0; j < m; ++j) {
    for (int i = 0; i < k; ++i)
      ITSOLV_REQUIRE(xx[i] != yy[j], "gemm_outer_with_dots: a target is also a source");
    for (int j2 = 0; j2 < j; ++j2)
      ITSOLV_REQUIRE(yy[j2] != yy[j], "gemm_outer_with_dots: the same target twice");
  }
  if (yscale)
    ctx->counters.n_scal += m;
  GoDots dots{pivot, ncols, cols, direct};
  return gemm_outer_impl(ctx, alpha, k, m, xx, yy, n, 0, yscale, &dots);
}

} // namespace itsolv
