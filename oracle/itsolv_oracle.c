/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; nothing under iterative_solver_b200/ links, loads or calls it.
 *
 * Plain-C restatement of the reference's CPU vector path (std::vector containers handled by ArrayHandlerIterable /
 * ArrayHandlerIterableSparse) — the arithmetic the CUDA kernels must reproduce. Every function cites the reference
 * file:line it follows (paths relative to the reference checkout). Compiled with -ffp-contract=off so that a*x+y
 * is two roundings, as in the reference's Release build on x86-64.
 *
 * Pinning: tests/test_oracle.py checks this file (a) against the golden vectors of the reference's own tests
 * (test/itsolv/subspace/test_util.cpp:154-173 modified Gram-Schmidt, test/array/testArrayHandlerIterable.cpp:57-69
 * select_max_dot, test/array/testArrayHandlers.cpp:27-60 sparse axpy/dot, test/array/testGemm.cpp:58-88 gemm == loops
 * of dot/axpy) and (b) bit for bit against the reference's own templates compiled in place (oracle/_ref/libitsolv_ref.so,
 * ref_handler_* entry points of oracle/ref_driver.cpp) wherever /root/reference is available, with the resulting
 * input/output vectors committed under tests/golden/.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* std::inner_product(begin(x), end(x), begin(y), 0.0): reference src/molpro/linalg/array/ArrayHandlerIterable.h:76-82 */
double oracle_dot(size_t n, const double* x, const double* y) {
  double acc = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double t = x[i] * y[i];
    acc = acc + t;
  }
  return acc;
}

/* same sum in extended precision: yardstick for the rounding error of both the oracle and the GPU tree sum */
long double oracle_dot_long(size_t n, const double* x, const double* y) {
  long double acc = 0.0L;
  for (size_t i = 0; i < n; ++i)
    acc += (long double)x[i] * (long double)y[i];
  return acc;
}

/* std::transform(y, x, y, ely + alpha*elx): reference ArrayHandlerIterable.h:65-74 */
void oracle_axpy(size_t n, double alpha, const double* x, double* y) {
  for (size_t i = 0; i < n; ++i) {
    const double t = alpha * x[i];
    y[i] = y[i] + t;
  }
}

/* el *= alpha: reference ArrayHandlerIterable.h:55-58 */
void oracle_scal(size_t n, double alpha, double* x) {
  for (size_t i = 0; i < n; ++i)
    x[i] = x[i] * alpha;
}

/* std::fill: reference ArrayHandlerIterable.h:60-63 */
void oracle_fill(size_t n, double alpha, double* x) {
  for (size_t i = 0; i < n; ++i)
    x[i] = alpha;
}

/* std::copy: reference ArrayHandlerIterable.h:48-53 */
void oracle_copy(size_t n, double* dst, const double* src) { memcpy(dst, src, n * sizeof(double)); }

/* gemm_inner_default: mat(i,j) = dot(xx[i], yy[j]), row-major k x m: reference src/molpro/linalg/array/util/gemm.h:268-279.
 * Vectors are packed row after row (X is k*n doubles). */
void oracle_gemm_inner(int k, int m, size_t n, const double* X, const double* Y, double* out) {
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < m; ++j)
      out[(size_t)i * m + j] = oracle_dot(n, X + (size_t)i * n, Y + (size_t)j * n);
}

/* gemm_outer_default: for ii, for jj: axpy(alphas(ii,jj), xx[ii], yy[jj]): reference util/gemm.h:258-265 */
void oracle_gemm_outer(int k, int m, size_t n, const double* alpha, const double* X, double* Y) {
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < m; ++j)
      oracle_axpy(n, alpha[(size_t)i * m + j], X + (size_t)i * n, Y + (size_t)j * n);
}

/* the CUDA gemm_outer forms each term with one FMA (same ascending-i order): the bit-exact model of that kernel */
void oracle_gemm_outer_fma(int k, int m, size_t n, const double* alpha, const double* X, double* Y) {
  for (int j = 0; j < m; ++j)
    for (size_t r = 0; r < n; ++r) {
      double acc = Y[(size_t)j * n + r];
      for (int i = 0; i < k; ++i)
        acc = fma(alpha[(size_t)i * m + j], X[(size_t)i * n + r], acc);
      Y[(size_t)j * n + r] = acc;
    }
}

/* precondition_default for iterable containers: a[i] = a[i] / (d[i] - shift[k] + 1e-15):
 * reference src/molpro/linalg/itsolv/IterativeSolver.h:46-55 */
void oracle_precondition(int w, size_t n, double* R, const double* shift, const double* diag) {
  for (int k = 0; k < w; ++k)
    for (size_t i = 0; i < n; ++i)
      R[(size_t)k * n + i] = R[(size_t)k * n + i] / (diag[i] - shift[k] + 1e-15);
}

/* select / select_max_dot: the reference pushes (key, index) pairs through a min-heap of size n and pops the smallest
 * pair each time, i.e. it keeps the n LARGEST pairs under lexicographic (key, index) order; key = v or |v| when max,
 * -v or -|v| otherwise; returned value has the sign restored; result ordered by index (std::map):
 * reference src/molpro/linalg/array/util/select.h:28-55, util/select_max_dot.h:22-46 */
typedef struct {
  double key;
  int64_t idx;
  double val;
} sel_t;
static int sel_cmp_desc(const void* a, const void* b) {
  const sel_t *p = (const sel_t*)a, *q = (const sel_t*)b;
  if (p->key != q->key)
    return p->key > q->key ? -1 : 1;
  if (p->idx != q->idx)
    return p->idx > q->idx ? -1 : 1;
  return 0;
}
static int sel_cmp_idx(const void* a, const void* b) {
  const sel_t *p = (const sel_t*)a, *q = (const sel_t*)b;
  return p->idx < q->idx ? -1 : (p->idx > q->idx ? 1 : 0);
}
int oracle_select(size_t nsel, size_t n, const double* x, const double* y, int max, int ignore_sign, int64_t* idx,
                  double* val) {
  sel_t* all = (sel_t*)malloc((n ? n : 1) * sizeof(sel_t));
  for (size_t i = 0; i < n; ++i) {
    double v;
    if (y) { /* select_max_dot: abs(x*y), largest */
      v = fabs(x[i] * y[i]);
      all[i].key = v;
    } else {
      v = ignore_sign ? fabs(x[i]) : x[i];
      all[i].key = max ? v : -v;
    }
    all[i].idx = (int64_t)i;
    all[i].val = v;
  }
  qsort(all, n, sizeof(sel_t), sel_cmp_desc);
  const size_t c = nsel < n ? nsel : n;
  qsort(all, c, sizeof(sel_t), sel_cmp_idx);
  for (size_t i = 0; i < c; ++i) {
    idx[i] = all[i].idx;
    val[i] = all[i].val;
  }
  free(all);
  return (int)c;
}

/* subspace::util::modified_gram_schmidt: reference src/molpro/linalg/itsolv/subspace/gram_schmidt.h:128-145.
 * Returns the number of null vectors, their indices in null_idx. */
int oracle_modified_gram_schmidt(int nvec, size_t n, double* V, double null_thresh, int* null_idx) {
  int nnull = 0;
  for (int i = 0; i < nvec; ++i) {
    double norm = oracle_dot(n, V + (size_t)i * n, V + (size_t)i * n);
    norm = sqrt(fabs(norm));
    if (norm > null_thresh) {
      oracle_scal(n, 1. / norm, V + (size_t)i * n);
      for (int j = i + 1; j < nvec; ++j) {
        const double ov = oracle_dot(n, V + (size_t)i * n, V + (size_t)j * n);
        oracle_axpy(n, -ov, V + (size_t)i * n, V + (size_t)j * n);
      }
    } else {
      null_idx[nnull++] = i;
    }
  }
  return nnull;
}

/* ---- dense x sparse (std::map packed as map_ptr / idx / val): reference ArrayHandlerIterableSparse.h:35-63 ---- */

/* copy(x, map): x = 0; x[idx] = val: reference ArrayHandlerIterableSparse.h:35-40 */
void oracle_sparse_copy(size_t n, double* x, int nnz, const int64_t* idx, const double* val) {
  for (size_t i = 0; i < n; ++i)
    x[i] = 0.0;
  for (int e = 0; e < nnz; ++e)
    x[idx[e]] = val[e];
}
/* dot(x, map): tot += x[idx]*val for idx < n: reference ArrayHandlerIterableSparse.h:53-59 */
double oracle_sparse_dot(size_t n, const double* x, int nnz, const int64_t* idx, const double* val) {
  double tot = 0.0;
  for (int e = 0; e < nnz; ++e)
    if ((size_t)idx[e] < n) {
      const double t = x[idx[e]] * val[e];
      tot = tot + t;
    }
  return tot;
}
/* axpy(alpha, map, y): y[idx] += alpha*val: reference ArrayHandlerIterableSparse.h:46-51 */
void oracle_sparse_axpy(size_t n, double alpha, int nnz, const int64_t* idx, const double* val, double* y) {
  for (int e = 0; e < nnz; ++e)
    if ((size_t)idx[e] < n) {
      const double t = alpha * val[e];
      y[idx[e]] = y[idx[e]] + t;
    }
}
/* gemm_inner_default over (dense, map) pairs: reference ArrayHandlerIterableSparse.h:65-67, util/gemm.h:268-279 */
void oracle_sparse_gemm_inner(int k, int m, size_t n, const double* X, const int32_t* map_ptr, const int64_t* idx,
                              const double* val, double* out) {
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < m; ++j)
      out[(size_t)i * m + j] =
          oracle_sparse_dot(n, X + (size_t)i * n, map_ptr[j + 1] - map_ptr[j], idx + map_ptr[j], val + map_ptr[j]);
}
/* gemm_outer_default over (map, dense) pairs, alphas is nmap x ndense: reference ArrayHandlerIterableSparse.h:61-63,
 * util/gemm.h:258-265 */
void oracle_sparse_gemm_outer(int nmap, int ndense, size_t n, const double* alpha, const int32_t* map_ptr,
                              const int64_t* idx, const double* val, double* Y) {
  for (int i = 0; i < nmap; ++i)
    for (int j = 0; j < ndense; ++j)
      oracle_sparse_axpy(n, alpha[(size_t)i * ndense + j], map_ptr[i + 1] - map_ptr[i], idx + map_ptr[i],
                         val + map_ptr[i], Y + (size_t)j * n);
}

/* util::make_distribution_spread_remainder: reference src/molpro/linalg/array/util/Distribution.h:99-110 */
void oracle_distribution(size_t n, int nchunks, int64_t* borders) {
  const size_t block = n / (size_t)nchunks, extra = n % (size_t)nchunks;
  borders[0] = 0;
  for (int c = 0; c < nchunks; ++c)
    borders[c + 1] = borders[c] + (int64_t)(block + ((size_t)c < extra ? 1 : 0));
}

/* the harness' synthetic banded operator, CPU twin of the CUDA SpMV; not reference code */
void oracle_banded_apply(int64_t n, int b, double eps, const double* v, double* a) {
  for (int64_t i = 0; i < n; ++i) {
    double acc = 0.0;
    const int64_t lo = i - b > 0 ? i - b : 0, hi = i + b < n - 1 ? i + b : n - 1;
    for (int64_t j = lo; j <= hi; ++j) {
      const double aij = i == j ? (double)(i + 1) : eps * (double)(1 + ((i + j) % 7));
      const double t = aij * v[j];
      acc = acc + t;
    }
    a[i] = acc;
  }
}

/* per-op timing helper for bench.py's cpu_baseline "port" leg: op 0 dot, 1 axpy, 3 gemm_inner, 4 gemm_outer */
double oracle_checksum(size_t n, const double* x) {
  double s = 0.0;
  for (size_t i = 0; i < n; ++i)
    s += x[i];
  return s;
}
