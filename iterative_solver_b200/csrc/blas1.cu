// Streaming (BLAS-1) kernels of the handler contract and the Davidson diagonal preconditioner.
// All are HBM-bound with no reuse: 128-bit coalesced accesses, 4 independent row pairs per thread per trip so that
// each SM keeps >= 64 KB of loads in flight, grid = a whole number of CTAs per SM (no tail wave).
// Arithmetic mirrors the reference's CPU loops operation for operation (no FMA contraction) so results are bit-identical
// to ArrayHandlerIterable (reference src/molpro/linalg/array/ArrayHandlerIterable.h:46-74) and to
// precondition_default (reference src/molpro/linalg/itsolv/IterativeSolver.h:46-55).
#include "common.cuh"

namespace itsolv {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

struct PrecondParams {
  double* r[ITSOLV_MAX_PANEL];
  double shift[ITSOLV_MAX_PANEL];
  const double* diag;
  size_t n;
  int w;
};

struct BatchParams {
  double* y[ITSOLV_MAX_PANEL];
  const double* x[ITSOLV_MAX_PANEL];
  double alpha[ITSOLV_MAX_PANEL];
  size_t n;
  int w;
};

struct MgsStepParams {
  double* rj[ITSOLV_MAX_PANEL];
  double ov[ITSOLV_MAX_PANEL];
  double* ri;
  double inv_norm;
  size_t n;
  int m;
};

// ---- generic driver: VEC = all pointers 16-byte aligned -> double2 path for the even prefix, scalar for the last odd row
template <bool VEC, class OpPair, class OpOne>
__device__ __forceinline__ void stream_rows(size_t n, OpPair op2, OpOne op1) {
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = size_t(gridDim.x) * blockDim.x;
  if (VEC) {
    const size_t npairs = n / 2;
    size_t p = tid;
    for (; p + (kUnroll - 1) * nthreads < npairs; p += kUnroll * nthreads) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        op2(p + u * nthreads);
    }
    for (; p < npairs; p += nthreads)
      op2(p);
    if ((n & 1) && tid == 0)
      op1(n - 1);
  } else {
    for (size_t i = tid; i < n; i += nthreads)
      op1(i);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4) fill_kernel(double alpha, double* __restrict__ x, size_t n) {
  stream_rows<VEC>(
      n, [&](size_t p) { reinterpret_cast<double2*>(x)[p] = make_double2(alpha, alpha); },
      [&](size_t i) { x[i] = alpha; });
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4) scal_kernel(double alpha, double* __restrict__ x, size_t n) {
  stream_rows<VEC>(
      n,
      [&](size_t p) {
        double2 v = reinterpret_cast<double2*>(x)[p];
        v.x = __dmul_rn(v.x, alpha);
        v.y = __dmul_rn(v.y, alpha);
        reinterpret_cast<double2*>(x)[p] = v;
      },
      [&](size_t i) { x[i] = __dmul_rn(x[i], alpha); });
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4) copy_kernel(double* __restrict__ dst, const double* __restrict__ src, size_t n) {
  stream_rows<VEC>(
      n, [&](size_t p) { reinterpret_cast<double2*>(dst)[p] = reinterpret_cast<const double2*>(src)[p]; },
      [&](size_t i) { dst[i] = src[i]; });
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4)
    axpy_kernel(double alpha, const double* __restrict__ x, double* __restrict__ y, size_t n) {
  // y + alpha*x with the product rounded before the sum (reference ArrayHandlerIterable.h:71-72)
  stream_rows<VEC>(
      n,
      [&](size_t p) {
        const double2 xv = reinterpret_cast<const double2*>(x)[p];
        double2 yv = reinterpret_cast<double2*>(y)[p];
        yv.x = __dadd_rn(yv.x, __dmul_rn(alpha, xv.x));
        yv.y = __dadd_rn(yv.y, __dmul_rn(alpha, xv.y));
        reinterpret_cast<double2*>(y)[p] = yv;
      },
      [&](size_t i) { y[i] = __dadd_rn(y[i], __dmul_rn(alpha, x[i])); });
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4) shift_kernel(double c, const double* __restrict__ x, double* __restrict__ out, size_t n) {
  stream_rows<VEC>(
      n,
      [&](size_t p) {
        double2 v = reinterpret_cast<const double2*>(x)[p];
        v.x = __dadd_rn(v.x, c);
        v.y = __dadd_rn(v.y, c);
        reinterpret_cast<double2*>(out)[p] = v;
      },
      [&](size_t i) { out[i] = __dadd_rn(x[i], c); });
}

// element-wise members of the reference's DistrArray (array/DistrArray.cpp:79-167), one rounding per operation as there
__device__ __forceinline__ double elementwise_op(int op, double c, double a, double b, double s) {
  switch (op) {
  case ITSOLV_EW_ADD_SCALAR:
    return __dadd_rn(c, s);                                   // c += s            (:81-85)
  case ITSOLV_EW_RECIP:
    return __ddiv_rn(1.0, c);                                 // c = 1 / c         (:91-95)
  case ITSOLV_EW_TIMES_INPLACE:
    return __dmul_rn(c, a);                                   // c *= a            (:97-107)
  case ITSOLV_EW_TIMES:
    return __dmul_rn(a, b);                                   // c = a * b         (:109-122)
  case ITSOLV_EW_DIVIDE:
    return __ddiv_rn(a, __dadd_rn(b, s));                     // c = a / (b + s)   (:164-165)
  case ITSOLV_EW_DIVIDE_NEGATIVE:
    return __ddiv_rn(-a, __dadd_rn(b, s));                    // c = -a / (b + s)  (:161-162)
  case ITSOLV_EW_DIVIDE_APPEND:
    return __dadd_rn(c, __ddiv_rn(a, __dadd_rn(b, s)));       // c += a / (b + s)  (:157-158)
  default:
    return __dsub_rn(c, __ddiv_rn(a, __dadd_rn(b, s)));       // c -= a / (b + s)  (:154-155)
  }
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4) elementwise_kernel(int op, double* __restrict__ c, const double* a,
                                                                  const double* b, double s, size_t n) {
  const bool use_c = op != ITSOLV_EW_TIMES && op != ITSOLV_EW_DIVIDE && op != ITSOLV_EW_DIVIDE_NEGATIVE;
  stream_rows<VEC>(
      n,
      [&](size_t p) {
        double2 cv = use_c ? reinterpret_cast<const double2*>(c)[p] : make_double2(0.0, 0.0);
        const double2 av = a ? reinterpret_cast<const double2*>(a)[p] : make_double2(0.0, 0.0);
        const double2 bv = b ? reinterpret_cast<const double2*>(b)[p] : make_double2(0.0, 0.0);
        cv.x = elementwise_op(op, cv.x, av.x, bv.x, s);
        cv.y = elementwise_op(op, cv.y, av.y, bv.y, s);
        reinterpret_cast<double2*>(c)[p] = cv;
      },
      [&](size_t i) { c[i] = elementwise_op(op, use_c ? c[i] : 0.0, a ? a[i] : 0.0, b ? b[i] : 0.0, s); });
}

// r_k[i] = r_k[i] / ((diag[i] - shift_k) + 1e-15): the diagonal is read once for all w residuals
template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4) precondition_kernel(const __grid_constant__ PrecondParams prm) {
  const double* __restrict__ diag = prm.diag;
  const int w = prm.w;
  stream_rows<VEC>(
      prm.n,
      [&](size_t p) {
        const double2 d = reinterpret_cast<const double2*>(diag)[p];
        for (int k = 0; k < w; ++k) {
          double2* rk = reinterpret_cast<double2*>(prm.r[k]);
          double2 v = rk[p];
          const double s = prm.shift[k];
          v.x = div_rn(v.x, __dadd_rn(__dsub_rn(d.x, s), 1e-15));
          v.y = div_rn(v.y, __dadd_rn(__dsub_rn(d.y, s), 1e-15));
          rk[p] = v;
        }
      },
      [&](size_t i) {
        const double d = diag[i];
        for (int k = 0; k < w; ++k)
          prm.r[k][i] = div_rn(prm.r[k][i], __dadd_rn(__dsub_rn(d, prm.shift[k]), 1e-15));
      });
}

// y_k[i] = y_k[i] * alpha_k for w vectors in one launch (the normalisation of a working set,
// reference itsolv/propose_rspace.h:17-28: w separate scal sweeps)
template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4) scal_batch_kernel(const __grid_constant__ BatchParams prm) {
  const int w = prm.w;
  stream_rows<VEC>(
      prm.n,
      [&](size_t p) {
        for (int k = 0; k < w; ++k) {
          double2* yk = reinterpret_cast<double2*>(prm.y[k]);
          double2 v = yk[p];
          v.x = __dmul_rn(v.x, prm.alpha[k]);
          v.y = __dmul_rn(v.y, prm.alpha[k]);
          yk[p] = v;
        }
      },
      [&](size_t i) {
        for (int k = 0; k < w; ++k)
          prm.y[k][i] = __dmul_rn(prm.y[k][i], prm.alpha[k]);
      });
}

// y_k[i] = alpha_k for w vectors in one launch (zeroing a working set before sparse contributions are added,
// reference itsolv/IterativeSolverTemplate.h:44-46: one fill sweep per vector)
template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4) fill_batch_kernel(const __grid_constant__ BatchParams prm) {
  const int w = prm.w;
  stream_rows<VEC>(
      prm.n,
      [&](size_t p) {
        for (int k = 0; k < w; ++k)
          reinterpret_cast<double2*>(prm.y[k])[p] = make_double2(prm.alpha[k], prm.alpha[k]);
      },
      [&](size_t i) {
        for (int k = 0; k < w; ++k)
          prm.y[k][i] = prm.alpha[k];
      });
}

// y_k[i] = y_k[i] + alpha_k * x_k[i] for w independent pairs in one launch (residual construction,
// reference itsolv/LinearEigensystemDavidson.h:186-192: one axpy sweep per root); product rounded before the sum
template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4) axpy_batch_kernel(const __grid_constant__ BatchParams prm) {
  const int w = prm.w;
  stream_rows<VEC>(
      prm.n,
      [&](size_t p) {
        for (int k = 0; k < w; ++k) {
          const double2 xv = reinterpret_cast<const double2*>(prm.x[k])[p];
          double2* yk = reinterpret_cast<double2*>(prm.y[k]);
          double2 yv = yk[p];
          yv.x = __dadd_rn(yv.x, __dmul_rn(prm.alpha[k], xv.x));
          yv.y = __dadd_rn(yv.y, __dmul_rn(prm.alpha[k], xv.y));
          yk[p] = yv;
        }
      },
      [&](size_t i) {
        for (int k = 0; k < w; ++k)
          prm.y[k][i] = __dadd_rn(prm.y[k][i], __dmul_rn(prm.alpha[k], prm.x[k][i]));
      });
}

// One step of the R-R modified Gram-Schmidt (reference itsolv/propose_rspace.h:451-463): r_i is scaled to unit norm and
// removed from the m later vectors, r_i read once: r_i *= inv_norm; r_j += (-ov_j) * r_i. Element for element the
// arithmetic of the reference's scal followed by its axpys.
template <bool VEC>
__global__ void __launch_bounds__(kThreads, 4) mgs_step_kernel(const __grid_constant__ MgsStepParams prm) {
  const int m = prm.m;
  stream_rows<VEC>(
      prm.n,
      [&](size_t p) {
        double2* ri = reinterpret_cast<double2*>(prm.ri);
        double2 v = ri[p];
        v.x = __dmul_rn(v.x, prm.inv_norm);
        v.y = __dmul_rn(v.y, prm.inv_norm);
        ri[p] = v;
        for (int k = 0; k < m; ++k) {
          double2* rj = reinterpret_cast<double2*>(prm.rj[k]);
          double2 y = rj[p];
          y.x = __dadd_rn(y.x, __dmul_rn(-prm.ov[k], v.x));
          y.y = __dadd_rn(y.y, __dmul_rn(-prm.ov[k], v.y));
          rj[p] = y;
        }
      },
      [&](size_t i) {
        const double v = __dmul_rn(prm.ri[i], prm.inv_norm);
        prm.ri[i] = v;
        for (int k = 0; k < m; ++k)
          prm.rj[k][i] = __dadd_rn(prm.rj[k][i], __dmul_rn(-prm.ov[k], v));
      });
}

static int stream_grid(itsolv_ctx* ctx, size_t n) {
  const int per_sm = ctx->opt_blas1_ctas > 0 ? ctx->opt_blas1_ctas : 4; // 4 x 256 threads, 4 row pairs in flight each
  const size_t want = (n / 2 + size_t(kThreads) * kUnroll - 1) / (size_t(kThreads) * kUnroll);
  size_t grid = size_t(ctx->num_sms) * per_sm;
  if (want < grid)
    grid = want ? want : 1;
  return int(grid);
}

} // namespace itsolv

using namespace itsolv;

#define LAUNCH_STREAM(kernel, vec, ...)                                                                                \
  do {                                                                                                                 \
    const int grid__ = stream_grid(ctx, n);                                                                            \
    mark_launch(ctx);                                                                                                  \
    if (vec)                                                                                                           \
      kernel<true><<<grid__, kThreads, 0, ctx->stream>>>(__VA_ARGS__);                                                 \
    else                                                                                                               \
      kernel<false><<<grid__, kThreads, 0, ctx->stream>>>(__VA_ARGS__);                                                \
    ctx->counters.launches += 1;                                                                                       \
    ITSOLV_CUDA(cudaGetLastError());                                                                                   \
  } while (0)

extern "C" {

int itsolv_fill_f64(itsolv_ctx* ctx, double alpha, double* x, size_t n) {
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  ctx->counters.n_fill++;
  if (n == 0)
    return 0;
  CallScope scope(ctx, OP_BLAS1, 8.0 * n);
  LAUNCH_STREAM(fill_kernel, aligned16(x), alpha, x, n);
  return 0;
}

int itsolv_scal_f64(itsolv_ctx* ctx, double alpha, double* x, size_t n) {
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  ctx->counters.n_scal++;
  if (n == 0)
    return 0;
  CallScope scope(ctx, OP_BLAS1, 16.0 * n);
  LAUNCH_STREAM(scal_kernel, aligned16(x), alpha, x, n);
  return 0;
}

int itsolv_copy_f64(itsolv_ctx* ctx, double* dst, const double* src, size_t n) {
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  ctx->counters.n_copy++;
  if (n == 0 || dst == src)
    return 0;
  CallScope scope(ctx, OP_BLAS1, 16.0 * n);
  LAUNCH_STREAM(copy_kernel, aligned16(dst) && aligned16(src), dst, src, n);
  return 0;
}

int itsolv_axpy_f64(itsolv_ctx* ctx, double alpha, const double* x, double* y, size_t n) {
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  ctx->counters.n_axpy++;
  if (n == 0)
    return 0;
  CallScope scope(ctx, OP_BLAS1, 24.0 * n);
  LAUNCH_STREAM(axpy_kernel, aligned16(x) && aligned16(y), alpha, x, y, n);
  return 0;
}

int itsolv_scal_batch_f64(itsolv_ctx* ctx, const double* alpha, double* const* x, int w, size_t n) {
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  ctx->counters.n_scal += w > 0 ? w : 0;
  if (n == 0 || w <= 0)
    return 0;
  for (int start = 0; start < w; start += ITSOLV_MAX_PANEL) {
    const int cnt = (w - start) < ITSOLV_MAX_PANEL ? (w - start) : ITSOLV_MAX_PANEL;
    BatchParams prm;
    bool vec = true;
    for (int k = 0; k < cnt; ++k) {
      prm.y[k] = x[start + k];
      prm.alpha[k] = alpha[start + k];
      vec = vec && aligned16(prm.y[k]);
    }
    prm.n = n;
    prm.w = cnt;
    CallScope scope(ctx, OP_BLAS1, 16.0 * n * cnt);
    LAUNCH_STREAM(scal_batch_kernel, vec, prm);
  }
  return 0;
}

int itsolv_fill_batch_f64(itsolv_ctx* ctx, const double* alpha, double* const* x, int w, size_t n) {
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  ctx->counters.n_fill += w > 0 ? w : 0;
  if (n == 0 || w <= 0)
    return 0;
  for (int start = 0; start < w; start += ITSOLV_MAX_PANEL) {
    const int cnt = (w - start) < ITSOLV_MAX_PANEL ? (w - start) : ITSOLV_MAX_PANEL;
    BatchParams prm;
    bool vec = true;
    for (int k = 0; k < cnt; ++k) {
      prm.y[k] = x[start + k];
      prm.alpha[k] = alpha[start + k];
      vec = vec && aligned16(prm.y[k]);
    }
    prm.n = n;
    prm.w = cnt;
    CallScope scope(ctx, OP_BLAS1, 8.0 * n * cnt);
    LAUNCH_STREAM(fill_batch_kernel, vec, prm);
  }
  return 0;
}

int itsolv_axpy_batch_f64(itsolv_ctx* ctx, const double* alpha, const double* const* x, double* const* y, int w, size_t n) {
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  ctx->counters.n_axpy += w > 0 ? w : 0;
  if (n == 0 || w <= 0)
    return 0;
  for (int a = 0; a < w; ++a)
    for (int b = 0; b < w; ++b)
      ITSOLV_REQUIRE(x[a] != y[b] && (a == b || y[a] != y[b]), "itsolv_axpy_batch_f64: the pairs must not alias each other");
  for (int start = 0; start < w; start += ITSOLV_MAX_PANEL) {
    const int cnt = (w - start) < ITSOLV_MAX_PANEL ? (w - start) : ITSOLV_MAX_PANEL;
    BatchParams prm;
    bool vec = true;
    for (int k = 0; k < cnt; ++k) {
      prm.y[k] = y[start + k];
      prm.x[k] = x[start + k];
      prm.alpha[k] = alpha[start + k];
      vec = vec && aligned16(prm.y[k]) && aligned16(prm.x[k]);
    }
    prm.n = n;
    prm.w = cnt;
    CallScope scope(ctx, OP_BLAS1, 24.0 * n * cnt);
    LAUNCH_STREAM(axpy_batch_kernel, vec, prm);
  }
  return 0;
}

int itsolv_mgs_step_f64(itsolv_ctx* ctx, double inv_norm, double* ri, const double* ov, double* const* rj, int m, size_t n) {
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  ctx->counters.n_scal += 1;
  ctx->counters.n_axpy += m > 0 ? m : 0;
  if (n == 0)
    return 0;
  ITSOLV_REQUIRE(m >= 0 && m <= ITSOLV_MAX_PANEL, "itsolv_mgs_step_f64: too many vectors");
  MgsStepParams prm;
  bool vec = aligned16(ri);
  for (int k = 0; k < m; ++k) {
    ITSOLV_REQUIRE(rj[k] != ri, "itsolv_mgs_step_f64: a target aliases the pivot vector");
    prm.rj[k] = rj[k];
    prm.ov[k] = ov[k];
    vec = vec && aligned16(rj[k]);
  }
  prm.ri = ri;
  prm.inv_norm = inv_norm;
  prm.n = n;
  prm.m = m;
  CallScope scope(ctx, OP_BLAS1, 16.0 * n * (m + 1));
  LAUNCH_STREAM(mgs_step_kernel, vec, prm);
  return 0;
}

int itsolv_shift_f64(itsolv_ctx* ctx, double c, const double* x, double* out, size_t n) {
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  if (n == 0)
    return 0;
  LAUNCH_STREAM(shift_kernel, aligned16(x) && aligned16(out), c, x, out, n);
  return 0;
}

int itsolv_elementwise_f64(itsolv_ctx* ctx, int op, double* c, const double* a, const double* b, double scalar, size_t n) {
  ++ctx->write_epoch;
  ITSOLV_REQUIRE(op >= ITSOLV_EW_ADD_SCALAR && op <= ITSOLV_EW_DIVIDE_APPEND_NEGATIVE, "itsolv_elementwise_f64: unknown operation");
  const bool need_a = op >= ITSOLV_EW_TIMES_INPLACE, need_b = op >= ITSOLV_EW_TIMES;
  ITSOLV_REQUIRE(c != nullptr && (!need_a || a != nullptr) && (!need_b || b != nullptr), "itsolv_elementwise_f64: null argument");
  if (n == 0)
    return 0;
  if (!need_a)
    a = nullptr;
  if (!need_b)
    b = nullptr;
  CallScope scope(ctx, OP_BLAS1, 8.0 * n * (2 + (need_a ? 1 : 0) + (need_b ? 1 : 0)));
  LAUNCH_STREAM(elementwise_kernel, aligned16(c) && (!a || aligned16(a)) && (!b || aligned16(b)), op, c, a, b, scalar, n);
  return 0;
}

int itsolv_precondition_f64(itsolv_ctx* ctx, double* const* r, int w, const double* diag, const double* shift, size_t n) {
  ++ctx->write_epoch; // before any early return: every rank advances alike, also one with an empty shard
  ctx->counters.n_precondition++;
  if (n == 0 || w == 0)
    return 0;
  ITSOLV_REQUIRE(w > 0, "itsolv_precondition_f64: w < 0");
  for (int start = 0; start < w; start += ITSOLV_MAX_PANEL) {
    const int cnt = (w - start) < ITSOLV_MAX_PANEL ? (w - start) : ITSOLV_MAX_PANEL;
    PrecondParams prm;
    bool vec = aligned16(diag);
    for (int k = 0; k < cnt; ++k) {
      prm.r[k] = r[start + k];
      prm.shift[k] = shift[start + k];
      vec = vec && aligned16(prm.r[k]);
    }
    prm.diag = diag;
    prm.n = n;
    prm.w = cnt;
    CallScope scope(ctx, OP_BLAS1, 8.0 * n * (2.0 * cnt + 1));
    LAUNCH_STREAM(precondition_kernel, vec, prm);
  }
  return 0;
}

} // extern "C"
