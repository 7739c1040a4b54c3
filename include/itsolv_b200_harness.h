/*
 * C ABI of the solver harness: runs the reference's own solver drivers
 * (molpro::linalg::itsolv::LinearEigensystemDavidson / LinearEquationsDavidson / NonLinearEquationsDIIS,
 * reference src/molpro/linalg/itsolv/{LinearEigensystemDavidson,LinearEquationsDavidson,NonLinearEquationsDIIS}.h)
 * through IterativeSolver::solve() (IterativeSolverTemplate.h:322-408) on the synthetic banded operator, with
 * DistrArrayCUDA as the R/Q container and ArrayHandlerCUDA as the handler set.
 *
 * The same two structs are consumed by the CPU oracle build of the reference (oracle/ref_driver.cpp), so a parity
 * test hands one spec to both sides and compares the two results.
 */
#ifndef ITSOLV_B200_HARNESS_H
#define ITSOLV_B200_HARNESS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  ITSOLV_KIND_DAVIDSON = 0, /* LinearEigensystemDavidson */
  ITSOLV_KIND_LINEQ = 1,    /* LinearEquationsDavidson   */
  ITSOLV_KIND_DIIS = 2      /* NonLinearEquationsDIIS    */
};
enum {
  ITSOLV_PROBLEM_BANDED = 0, /* synthetic banded symmetric operator, SURVEY.md section 8(d) */
  ITSOLV_PROBLEM_EXAMPLE = 1 /* dense matrix of the reference's examples/ExampleProblem.h:8 (small n only) */
};
/* Right-hand sides of the LinearEquations cases, b_k = A x_k with a known x_k; u_k(i) = ((i (k+2) + k) mod (2k+5)) / (2k+5) - 1/2.
 *   SCALED (default): x_k(i) = ((k+1) + u_k(i)) / (i+1): b_k has entries of order one in every row, so the relative
 *     residual weighs all rows alike and the reference's solver converges in a size-independent number of iterations
 *     (the right-hand sides of the reference's own test, test/itsolv/test_LinearEquations.cpp:29-33, are A (k+1) 1; on
 *     this operator, whose diagonal grows like i, those are parallel and dominated by the last rows).
 *   LEGACY: x_k(i) = u_k(i): b_k(i) ~ i u_k(i); the reference's solver stagnates on it from n ~ 1e5 on (kept as a test of
 *     identical behaviour on an ill-posed input). */
enum { ITSOLV_RHS_SCALED = 0, ITSOLV_RHS_LEGACY = 1 };
#define ITSOLV_MAX_ROOTS 64

typedef struct itsolv_solve_spec {
  int64_t n;                    /* global vector length */
  int32_t kind;                 /* ITSOLV_KIND_* */
  int32_t problem;              /* ITSOLV_PROBLEM_* */
  int32_t nroots;               /* roots sought, or number of right-hand sides; ignored for DIIS */
  int32_t nbuffers;             /* R vectors handed to solve(); 0 means nroots */
  int32_t half_bandwidth;       /* b: A(i,j) != 0 for |i-j| <= b */
  int32_t hermitian;            /* set_hermiticity() */
  double eps;                   /* off-diagonal scale */
  double convergence_threshold; /* set_convergence_threshold() */
  int32_t max_iter;             /* set_max_iter() */
  int32_t max_size_qspace;      /* <= 0: unlimited */
  int32_t reset_D;              /* <= 0: library default (never) */
  int32_t max_p;                /* P-space size for solve(); 0: none */
  int32_t verbosity;            /* 0..3 as IterativeSolver::set_verbosity(int) */
  int32_t trace;                /* record every dot/gemm_inner result for parity checks */
  int32_t explicit_csr;         /* banded operator: 1 = stored CSR arrays, 0 = entries generated in the kernel */
  int32_t fused;                /* CUDA backend: 1 = fused driver path (Davidson: FusedDavidson.h, 2 = its batched pieces under
                                   the reference's solve() loop; LinearEquations / DIIS: fused X space, FusedEquations.h);
                                   0 = the reference's classes call for call; ignored by the oracle */
  int32_t rhs_kind;             /* LinearEquations right-hand sides b_k = A x_k of the harness (ITSOLV_RHS_*); DIIS: the target t of
                                   the residual r(v) = A (v - t): SCALED t(i) = 1/(i+1), LEGACY t = 1 */
} itsolv_solve_spec;

typedef struct itsolv_solve_result {
  int32_t converged;
  int32_t iterations; /* statistics().iterations */
  int32_t nroots;
  int32_t nwork_final;
  double eigenvalues[ITSOLV_MAX_ROOTS]; /* Davidson only */
  double errors[ITSOLV_MAX_ROOTS];
  double seconds_solve;   /* host wall time of solve() alone */
  double seconds_action;  /* of which inside Problem::action/residual */
  double seconds_precond; /* of which inside Problem::precondition */
  int64_t r_creations, q_creations, p_creations, d_creations;
  /* handler call counts and algorithmic bytes accumulated by the CUDA handlers; zero for the oracle */
  int64_t n_dot, n_axpy, n_scal, n_copy, n_fill, n_gemm_inner, n_gemm_outer;
  double handler_bytes;
  double handler_device_seconds;
  int64_t kernel_launches;
  double device_ms_solve; /* solve() bracketed by CUDA events on the context's stream (0 for the oracle) */
  double bytes_gemm_inner, seconds_gemm_inner, bytes_gemm_outer, seconds_gemm_outer, bytes_blas1, seconds_blas1;
  double bytes_residual, seconds_residual; /* the fused solution/residual/preconditioner kernel */
  int64_t calls_gemm_inner, calls_gemm_outer, calls_blas1, calls_residual;
} itsolv_solve_result;

/* One trace record = one handler call that returned numbers to the host. op: 'd' dot, 'g' gemm_inner. */
typedef struct itsolv_trace_entry {
  int32_t op, rows, cols, reserved;
  int64_t offset; /* into the value array */
} itsolv_trace_entry;

struct itsolv_ctx; /* include/itsolv_b200.h */

/*
 * Run one solve on the context's GPU (all ranks of the context's communicator call it collectively).
 * solutions: optional host buffer, nroots * n_local doubles, receives this rank's rows of each solution vector.
 * Returns 0 on success, non-zero on error (message via itsolv_harness_last_error()).
 */
int itsolv_harness_solve(struct itsolv_ctx* ctx, const itsolv_solve_spec* spec, itsolv_solve_result* result,
                         double* solutions);

/*
 * End-to-end entry: the caller owns the operator on the HOST as CSR (row_ptr[n_local+1] relative to this rank's first
 * row, col[nnz] global column indices within half_bandwidth of the local rows, val[nnz]) plus its diagonal; the call
 * uploads it, solves with the stored-CSR SpMV, and downloads eigenvalues/errors and the solution vectors.
 */
int itsolv_harness_solve_host_csr(struct itsolv_ctx* ctx, const itsolv_solve_spec* spec, const int64_t* row_ptr,
                                  const int32_t* col, const double* val, const double* diag,
                                  itsolv_solve_result* result, double* solutions);
const char* itsolv_harness_last_error(void);

/*
 * The same two calls split so that the operator stays resident in HBM between solves: create uploads (or, with
 * row_ptr == NULL, generates: stored CSR when spec->explicit_csr, else entries computed inside the SpMV kernel) the
 * operator once; solve may then be called repeatedly. All ranks call collectively.
 */
typedef struct itsolv_harness_problem itsolv_harness_problem;
int itsolv_harness_problem_create(struct itsolv_ctx* ctx, const itsolv_solve_spec* spec, const int64_t* row_ptr,
                                  const int32_t* col, const double* val, const double* diag,
                                  itsolv_harness_problem** problem);
int itsolv_harness_problem_solve(itsolv_harness_problem* problem, const itsolv_solve_spec* spec,
                                 itsolv_solve_result* result, double* solutions);
/* As itsolv_harness_problem_solve, but the solution vectors stay on the GPU: *device_solutions receives nroots * n_local
 * doubles taken from the context's pool (itsolv_alloc) AFTER the solver has finished, so they do not add to the solver's
 * high-water mark; the caller releases them with itsolv_free. */
int itsolv_harness_problem_solve_device(itsolv_harness_problem* problem, const itsolv_solve_spec* spec,
                                        itsolv_solve_result* result, double** device_solutions);
void itsolv_harness_problem_destroy(itsolv_harness_problem* problem);

size_t itsolv_harness_trace_entries(void);
size_t itsolv_harness_trace_values(void);
void itsolv_harness_trace_read(itsolv_trace_entry* entries, double* values);

/*
 * The handler contract exercised through the C++ plugin classes (DistrArrayCUDA + ArrayHandlerCUDA /
 * ArrayHandlerCUDASparse) with HOST buffers: vectors are packed row after row (X is k*n doubles, GLOBAL length n;
 * each rank uploads its own shard and, for outputs, writes back its own rows only). These mirror, call for call, the
 * ref_handler_* entry points of the oracle build (oracle/ref_driver.cpp), so a parity test runs the same inputs through both.
 */
int itsolv_handler_blas1(struct itsolv_ctx* ctx, int op /*0 dot 1 axpy 2 scal 3 fill 4 copy*/, size_t n, double alpha,
                         const double* x, double* y, double* result);
int itsolv_handler_gemm_inner(struct itsolv_ctx* ctx, int k, int m, size_t n, const double* X, const double* Y,
                              int y_is_x, double* out);
int itsolv_handler_gemm_outer(struct itsolv_ctx* ctx, int k, int m, size_t n, const double* alpha, const double* X,
                              double* Y);
int itsolv_handler_select(struct itsolv_ctx* ctx, size_t nsel, size_t n, const double* x, const double* y_or_null,
                          int max, int ignore_sign, int64_t* idx, double* val);
/* Problem::precondition of the harness problem (reference IterativeSolver.h:135-137) */
int itsolv_handler_precondition(struct itsolv_ctx* ctx, int w, size_t n, double* r, const double* shift,
                                const double* diag);
/* the reference's own subspace::util::modified_gram_schmidt (subspace/gram_schmidt.h:128-145) run on the CUDA handler */
int itsolv_handler_modified_gram_schmidt(struct itsolv_ctx* ctx, int nvec, size_t n, double* data, double thresh,
                                         int* null_idx);
int itsolv_handler_sparse_copy(struct itsolv_ctx* ctx, size_t n, double* x, int nnz, const int64_t* idx,
                               const double* val);
int itsolv_handler_sparse_gemm_inner(struct itsolv_ctx* ctx, int k, int m, size_t n, const double* X,
                                     const int32_t* map_ptr, const int64_t* idx, const double* val, double* out);
int itsolv_handler_sparse_gemm_outer(struct itsolv_ctx* ctx, int nmap, int ndense, size_t n, const double* alpha,
                                     const int32_t* map_ptr, const int64_t* idx, const double* val, double* Y);
/* The element-wise and sparse members of the container itself (DistrArrayCUDA, mirroring reference array/DistrArray.cpp:79-167,
 * 248-262, 419-465), for the reference's conformance tests (test/array/testDistrArray.h:496-678). c is updated in place.
 * op: 0 add(b) 1 sub(b) 2 add(scalar) 3 sub(scalar) 4 recip 5 times(a) 6 times(a, b) 7 divide(a, b, scalar, flags&1 append,
 * flags&2 negative) 8 axpy(scalar, sparse) 9 dot(sparse) -> *result 10 select_max_dot(n = flags, sparse) -> sel_* (returns
 * the count) 11 zero. Returns -1 on error. */
int itsolv_handler_distr_array(struct itsolv_ctx* ctx, int op, size_t n, double scalar, int flags, const double* a,
                               const double* b, double* c, int nnz, const int64_t* idx, const double* val,
                               double* result, int64_t* sel_idx, double* sel_val);
/* harness operator on host vectors (single rank: whole vector; multi rank: global x in, this rank's rows of y out) */
int itsolv_harness_banded_apply(struct itsolv_ctx* ctx, int64_t n, int b, double eps, int explicit_csr, const double* x,
                                double* y);

/* Host subspace algebra entry points (restated helper, reference helper-implementation.h:318-543), exported for tests */
int itsolv_host_eigenproblem(const double* matrix, const double* metric, size_t dimension, int hermitian,
                             double svd_threshold, double* eigenvalues, double* eigenvectors, size_t* nfound);
/* svd_system (helper-implementation.h:263-296): values[k], vectors[k*ncols .. ] for the nfound entries of the returned list */
int itsolv_host_svd_system(size_t nrows, size_t ncols, const double* m, double threshold, int hermitian,
                           int reduce_to_rank, double* values, double* vectors, size_t* nfound);
/* solve_LinearEquations (:553-617): solution is dimension x nroot column-major; eigenvalues (nroot) only with augmented_hessian > 0 */
int itsolv_host_solve_linear_equations(const double* matrix, const double* metric, const double* rhs, size_t dimension,
                                       size_t nroot, double augmented_hessian, double svd_threshold, double* solution,
                                       double* eigenvalues);
/* solve_DIIS (:619-669) */
int itsolv_host_solve_diis(const double* matrix, size_t dimension, double svd_threshold, double* solution);

#ifdef __cplusplus
}
#endif
#endif
