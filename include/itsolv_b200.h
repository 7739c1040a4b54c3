/*
 * itsolv_b200 — C ABI of the B200 (sm_100a) vector backend for molpro::linalg::itsolv.
 *
 * Every entry point below is what the reference's ArrayHandler contract
 * (reference src/molpro/linalg/array/ArrayHandler.h:184-222) needs from a device: plain pointers and sizes, no C++
 * types, no allocation on the hot path, no CPU fallback. The C++ host layer (iterative_solver_b200/host/
 * ArrayHandlerCUDA.h) is a thin adaptor from the contract's containers to these calls; a foreign host (C, Fortran,
 * Python ctypes) can bind them directly.
 *
 * Conventions
 *   - all vectors are FP64, device resident, LOCAL shards of length n (rows owned by the calling rank);
 *   - "xx"/"yy" are HOST arrays of DEVICE pointers (the Q/R vectors are separate allocations,
 *     reference itsolv/subspace/QSpace.h:157), 1 <= k,m <= ITSOLV_MAX_PANEL per call;
 *   - small matrices (Gram results, alphas, shifts) are HOST, row-major, exactly the layout of the reference's
 *     Matrix<double> (reference itsolv/subspace/Matrix.h:59);
 *   - calls that return numbers to the host (dot, gemm_inner, select) finish the work on the context's stream,
 *     all-reduce over the context's communicator when one is attached (the reference's MPI_Allreduce sites,
 *     array/util/gemm.h:179-182, DistrArray.cpp:134-136) and synchronise; every other call is asynchronous on the stream;
 *   - return value 0 = success; otherwise itsolv_last_error() describes the failure (thread local).
 */
#ifndef ITSOLV_B200_H
#define ITSOLV_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ITSOLV_MAX_PANEL 128     /* vectors per side of one gemm_inner / gemm_outer launch */
#define ITSOLV_UNIQUE_ID_BYTES 128

typedef struct itsolv_ctx itsolv_ctx;

/* ---- context: one per process/GPU. Owns a stream, the partial-sum workspace, pinned result buffers ---- */
int itsolv_ctx_create(int device, itsolv_ctx** ctx);
/* as above but all work is issued on the caller's cudaStream_t (e.g. torch's current stream) */
int itsolv_ctx_create_on_stream(int device, void* cuda_stream, itsolv_ctx** ctx);
void itsolv_ctx_destroy(itsolv_ctx* ctx);
const char* itsolv_last_error(void);
void* itsolv_ctx_stream(itsolv_ctx* ctx);
int itsolv_ctx_device(itsolv_ctx* ctx);
int itsolv_ctx_synchronize(itsolv_ctx* ctx);
/* tuning knob (also read from the environment variable ITSOLV_<NAME> at context creation); returns previous value */
int itsolv_ctx_set_option(itsolv_ctx* ctx, const char* name, int value);

/* ---- accounting: launches of this library's kernels, algorithmic bytes and, when profiling is
 * enabled, device time measured with CUDA events on the context's stream around every call ---- */
typedef struct itsolv_counters {
  int64_t launches;
  int64_t n_dot, n_axpy, n_scal, n_copy, n_fill, n_gemm_inner, n_gemm_outer, n_precondition, n_select, n_sparse;
  double bytes;          /* algorithmic bytes of all calls */
  double device_seconds; /* only while profiling */
  double bytes_gemm_inner, seconds_gemm_inner; /* gemm_inner_kernel launches: gemm_inner and dot (its 1 x 1 case) */
  double bytes_gemm_outer, seconds_gemm_outer; /* gemm_outer_kernel launches */
  double bytes_blas1, seconds_blas1;           /* streaming kernels: axpy, scal, copy, fill, preconditioner, mgs steps */
  double bytes_residual, seconds_residual;     /* davidson_residual_kernel launches (fused driver path) */
  int64_t calls_gemm_inner, calls_gemm_outer, calls_blas1, calls_residual; /* accounted calls (~ launches) per family */
} itsolv_counters;
void itsolv_ctx_counters(itsolv_ctx* ctx, itsolv_counters* out);
void itsolv_ctx_reset_counters(itsolv_ctx* ctx);
void itsolv_ctx_set_profiling(itsolv_ctx* ctx, int enabled);
/* CUDA-event stopwatches on the context's stream; id in [0, ITSOLV_TIMERS) (the solve harness uses id 0) */
#define ITSOLV_TIMERS 4
int itsolv_ctx_timer_start(itsolv_ctx* ctx, int id);
int itsolv_ctx_timer_stop(itsolv_ctx* ctx, int id, double* milliseconds);

/* ---- memory: stream-ordered pool (no cudaMalloc/cudaFree on the hot path; the reference churns 2w Q vectors per
 * iteration, itsolv/subspace/QSpace.h:80-84) ---- */
int itsolv_alloc(itsolv_ctx* ctx, size_t n, double** out);
int itsolv_free(itsolv_ctx* ctx, double* p);
int itsolv_upload(itsolv_ctx* ctx, double* dst_device, const double* src_host, size_t n);   /* synchronous */
int itsolv_download(itsolv_ctx* ctx, double* dst, const double* src_device, size_t n); /* synchronous; dst: host or device memory */
int itsolv_upload_bytes(itsolv_ctx* ctx, void* dst_device, const void* src_host, size_t bytes);  /* synchronous */
/* bytes currently handed out by itsolv_alloc and their high-water mark (memory planning of the large configurations) */
int itsolv_mem_usage(itsolv_ctx* ctx, size_t* live_bytes, size_t* peak_bytes, int reset_peak);
int itsolv_mem_info(itsolv_ctx* ctx, size_t* free_bytes, size_t* total_bytes);
/* hands the pool's unused memory back to the driver (the pool otherwise keeps every freed vector for reuse) */
int itsolv_mem_trim(itsolv_ctx* ctx);

/* ---- communicator: row-sharded vectors, one rank per GPU; replaces the reference's MPI communicator
 * (array/DistrArray.h:100,109). The unique id is created on rank 0 and broadcast by the host (torch.distributed, MPI, ...) ---- */
int itsolv_comm_unique_id(void* id /* ITSOLV_UNIQUE_ID_BYTES */);
int itsolv_comm_init(itsolv_ctx* ctx, int rank, int nranks, const void* id);
/* Peer-memory all-reduce fused into the Gram kernels (one process per GPU on one NVSwitch box): every rank exports the
 * IPC handle of its exchange buffer, the host gathers the handles of all ranks (rank order, ITSOLV_IPC_HANDLE_BYTES
 * each) and every rank imports them. Without this step dot/gemm_inner fall back to ncclAllReduce + copy + sync. */
#define ITSOLV_IPC_HANDLE_BYTES 64
int itsolv_comm_p2p_export(itsolv_ctx* ctx, void* handle);
int itsolv_comm_p2p_disable(itsolv_ctx* ctx); /* back to the ncclAllReduce path (all ranks must agree) */
int itsolv_comm_p2p_import(itsolv_ctx* ctx, const void* handles /* nranks * ITSOLV_IPC_HANDLE_BYTES */);
int itsolv_comm_rank(itsolv_ctx* ctx);
int itsolv_comm_size(itsolv_ctx* ctx);
int itsolv_comm_barrier(itsolv_ctx* ctx);
/* in-place sum / max of a small HOST array over all ranks (through a device staging buffer) */
int itsolv_comm_allreduce_host(itsolv_ctx* ctx, double* values, size_t count, int op_max);
/* the same for all w vectors of a working set at once: out[k][0..b) receives the last b rows of x[k] on rank-1, out[k][b..2b)
 * the first b rows of x[k] on rank+1 (zeros at the ends of the global vector). One launch with stores into the
 * neighbours' exchange buffers over NVLink when they are mapped (itsolv_comm_p2p_import), else w NCCL exchanges. */
int itsolv_comm_halo_exchange_multi(itsolv_ctx* ctx, const double* const* x, int w, size_t n_local, int b, double* out);
/* exchange of `count` doubles with the neighbouring ranks (harness halo): send_lo goes to rank-1, send_hi to rank+1 */
int itsolv_comm_halo_exchange(itsolv_ctx* ctx, const double* send_lo, const double* send_hi, double* recv_lo,
                              double* recv_hi, size_t count);
/* chunk borders of util::make_distribution_spread_remainder (reference array/util/Distribution.h:99-110); borders[nranks+1] */
void itsolv_distribution(size_t n, int nranks, int64_t* borders);

/* ---- BLAS-1 part of the contract (reference ArrayHandler.h:186-190; CPU path ArrayHandlerIterable.h:46-82) ---- */
int itsolv_fill_f64(itsolv_ctx* ctx, double alpha, double* x, size_t n);
int itsolv_scal_f64(itsolv_ctx* ctx, double alpha, double* x, size_t n);
int itsolv_copy_f64(itsolv_ctx* ctx, double* dst, const double* src, size_t n);
/* y[i] = y[i] + alpha*x[i], product and sum rounded separately as the reference's std::transform does */
int itsolv_axpy_f64(itsolv_ctx* ctx, double alpha, const double* x, double* y, size_t n);
int itsolv_dot_f64(itsolv_ctx* ctx, const double* x, const double* y, size_t n, double* result);

/* ---- batched forms used by the fused driver path (iterative_solver_b200/host/FusedDavidson.h): the same arithmetic as
 * w separate calls of the functions above, in one pass / one launch ---- */
/* x_k *= alpha[k] */
int itsolv_scal_batch_f64(itsolv_ctx* ctx, const double* alpha, double* const* x, int w, size_t n);
/* x_k = alpha[k] for w vectors in one launch */
int itsolv_fill_batch_f64(itsolv_ctx* ctx, const double* alpha, double* const* x, int w, size_t n);
/* y_k = y_k + alpha[k]*x_k for w independent pairs (residual construction, LinearEigensystemDavidson.h:186-192) */
int itsolv_axpy_batch_f64(itsolv_ctx* ctx, const double* alpha, const double* const* x, double* const* y, int w, size_t n);
/* one step of the R-R modified Gram-Schmidt (propose_rspace.h:451-463): ri *= inv_norm; rj[k] += (-ov[k])*ri */
int itsolv_mgs_step_f64(itsolv_ctx* ctx, double inv_norm, double* ri, const double* ov, double* const* rj, int m, size_t n);
/* the same step that also returns, from the values it has just written, the inner products the next step needs:
 * dots[0] = <ri', ri'>, dots[1+t] = <rj[0]', rj[t]'> for t in [0, m) (HOST, m+1 values, all-reduced over ranks); m <= 16 */
int itsolv_mgs_step_dots_f64(itsolv_ctx* ctx, double inv_norm, double* ri, const double* ov, double* const* rj, int m,
                             size_t n, double* dots);
/* The whole R-R modified Gram-Schmidt of w vectors (reference itsolv/propose_rspace.h:451-463) as ONE chain of launches
 * without host round trips: a Gram row for the first pivot, then w steps as above, each of which takes its coefficients
 * (1/|r_i|, -<r_i,r_j>/|r_i|) from device memory where the previous launch's tail has left them, computed with the host's
 * arithmetic from the all-reduced sums. A pivot whose norm is <= thresh is left untouched, and so are the later
 * vectors with respect to it. rows (HOST, w + w(w+1)/2 values): {<r_0,r_j> before the chain, j = 0..w-1}, then per step i
 * {<r_i',r_i'>, <r_{i+1}',r_j'> for j = i+1..w-1}: everything the caller needs to repeat the decisions. w <= 17.
 * itsolv_mgs_chain_supported: 1 when the chain can run (option MGS_CHAIN, sums delivered by the kernels themselves). */
int itsolv_mgs_chain_supported(itsolv_ctx* ctx, int w, size_t n);
int itsolv_mgs_chain_f64(itsolv_ctx* ctx, double* const* r, int w, size_t n, double thresh, double* rows);
/* The projection of a working set against the subspace and its R-R Gram-Schmidt as ONE chain: y_j = (yscale ? yscale[j] *
 * y_j : y_j) + sum_i alpha[i*m+j] x_i for all m new vectors (itsolv_gemm_outer_scaled_f64 / itsolv_gemm_outer_f64 bit for
 * bit), whose tail returns the first Gram row of the w vectors that stay (keep[0..w), ascending column indices; the others
 * were found redundant, reference propose_rspace.h:482-512) and starts the pivot steps of itsolv_mgs_chain_f64 on them.
 * rows: as itsolv_mgs_chain_f64. k <= 128, m <= 8. */
int itsolv_project_mgs_chain_supported(itsolv_ctx* ctx, int k, int m, int w, size_t n);
int itsolv_project_mgs_chain_f64(itsolv_ctx* ctx, const double* alpha, int k, int m, const double* const* x,
                                 double* const* y, const double* yscale, const int* keep, int w, size_t n, double thresh,
                                 double* rows);
/* counter that every call which may write a vector advances (and DistrArrayCUDA::data() non-const): results cached
 * on the host side (ArrayHandlerCUDA's primed dots) are valid only while it stands still */
unsigned long long itsolv_ctx_write_epoch(itsolv_ctx* ctx);
void itsolv_ctx_note_write(itsolv_ctx* ctx);

/* ---- panel contractions (reference ArrayHandler.h:195,200; CPU path array/util/gemm.h:157-203, 258-279) ---- */
/* out[i*m+j] = sum_r xx[i][r]*yy[j][r]; each HBM byte of the k+m vectors is read once */
int itsolv_gemm_inner_f64(itsolv_ctx* ctx, const double* const* xx, int k, const double* const* yy, int m, size_t n,
                          double* out);
/* yy[j] += sum_i alpha[i*m+j]*xx[i] (i ascending, as the reference's loop); beta_zero!=0: yy[j] = sum (yy not read) */
int itsolv_gemm_outer_f64(itsolv_ctx* ctx, const double* alpha, int k, int m, const double* const* xx,
                          double* const* yy, size_t n, int beta_zero);

/* yy[j] = yscale[j]*yy[j] + sum_i alpha[i*m+j]*xx[i]: the product yscale*yy is rounded first, then the sum runs as in
 * itsolv_gemm_outer_f64, i.e. the result is bit-identical to itsolv_scal_batch_f64 followed by itsolv_gemm_outer_f64
 * (normalise + Gram-Schmidt projection of the new vectors, reference itsolv/propose_rspace.h:17-28, 430-443) */
int itsolv_gemm_outer_scaled_f64(itsolv_ctx* ctx, const double* alpha, int k, int m, const double* const* xx,
                                 double* const* yy, size_t n, const double* yscale);

/* ---- Davidson diagonal preconditioner (reference itsolv/IterativeSolver.h:46-55):
 * r_k[i] = r_k[i] / (diag[i] - shift[k] + 1e-15), w vectors in one pass over diag ---- */
int itsolv_precondition_f64(itsolv_ctx* ctx, double* const* r, int w, const double* diag, const double* shift, size_t n);

/* ---- fused solution -> residual -> error norm -> preconditioner step of one Davidson iteration
 * (reference itsolv/IterativeSolverTemplate.h:34-65 construct_solution x 2, LinearEigensystemDavidson.h:186-192
 * construct_residual, IterativeSolverTemplate.h:96-102 update_errors, IterativeSolver.h:46-55 precondition_default),
 * one pass over the k subspace vectors q_i and their actions a_i, for m roots:
 *   x_j = sum_i coef[i*m+j]*q[i];  r_j = sum_i coef[i*m+j]*a[i];  r_j = r_j + (-lambda[j])*x_j;  norm2[j] = <r_j, r_j>
 *   out_r[j] = diag ? r_j / (diag - shift[j] + 1e-15) : r_j;       out_x[j] = x_j when out_x != NULL
 * Every operation is rounded as in the separate calls (gemm_outer with beta_zero, axpy, precondition), so the vectors
 * are bit-identical to that sequence; norm2 (HOST, m values, all-reduced over ranks) is taken BEFORE preconditioning,
 * norm2_out[j] = <out_r[j], out_r[j]> of what was written (either may be NULL).
 * 8n(2k + m + 1) bytes instead of 8n(2k + 8m + 1). ---- */
int itsolv_davidson_residual_f64(itsolv_ctx* ctx, const double* coef, int k, int m, const double* const* q,
                                 const double* const* a, const double* lambda, const double* diag, const double* shift,
                                 double* const* out_x, double* const* out_r, size_t n, double* norm2,
                                 double* norm2_out);
/*
 * The same pass for the other residual forms of the reference's solvers:
 *   mode 0   r_j = sum_i c_ij a_i - lambda_j x_j                       (LinearEigensystemDavidson.h:186-192; as above)
 *   mode 1   r_j = (sum_i c_ij a_i - rhs_j) * rscale_j                 (LinearEquationsDavidson.h:173-184: axpy(-1, rhs) then
 *                                                                        scal(1/|rhs|), each operation rounded)
 *   mode 2   r_j = sum_i c_ij a_i                                     (NonLinearEquationsDIIS: solution() of the
 *                                                                        extrapolated parameters and residual)
 *   mode 3   as 2, and what is stored in out_x[j] is x_j - out_r[j]      (its end_iteration, NonLinearEquationsDIIS.h:103-119:
 *                                                                        axpy(-1, preconditioned residual, parameters))
 * accumulate != 0: the expansions start from the present contents of out_x[j] / out_r[j] instead of zero - the P-space
 * parts (IterativeSolverTemplate.h:44-64 puts them first; the caller's apply_p contribution, :210-211) - out_x is then
 * required. With a diagonal the residuals are preconditioned as above (shift = 0 for linear equations).
 */
int itsolv_subspace_residual_f64(itsolv_ctx* ctx, int mode, int accumulate, const double* coef, int k, int m,
                                 const double* const* q, const double* const* a, const double* lambda,
                                 const double* const* rhs, const double* rscale, const double* diag, const double* shift,
                                 double* const* out_x, double* const* out_r, size_t n, double* norm2, double* norm2_out);

/* ---- select / select_max_dot (reference array/util/select.h:28-55, ArrayHandler.h:212,222): the nsel entries that are
 * largest under the reference's (key, index) pair ordering, key = max ? v : -v (|v| when ignore_sign); for
 * select_max_dot pass y != NULL: v = |x[i]*y[i]|, max. Indices are global (global_offset + local). With a communicator
 * the candidates of all ranks are merged, every rank receives the same list. Output sorted by index. ---- */
int itsolv_select_f64(itsolv_ctx* ctx, const double* x, const double* y, size_t n, size_t global_offset, size_t nsel,
                      int max, int ignore_sign, int64_t* indices, double* values, int* nfound);
/* host-side merge used by the call above; exported for tests: candidates (idx,val) from all ranks -> best nsel */
int itsolv_select_merge(const int64_t* idx, const double* val, size_t ncand, size_t nsel, int max, int ignore_sign,
                        int64_t* out_idx, double* out_val);

/* ---- dense x sparse (P-space) ops (reference ArrayHandlerIterableSparse.h:35-63, ArrayHandlerDistrSparse.h:30-65,
 * array/util/gemm.h:207-253). Sparse vectors are std::map<size_t,double> packed CSR-like on the HOST:
 * map_ptr[nmap+1], idx[] GLOBAL indices, val[]; entries outside [global_offset, global_offset+n) are skipped ---- */
/* x = 0; x[idx-global_offset] = val */
int itsolv_sparse_copy_f64(itsolv_ctx* ctx, double* x, size_t n, size_t global_offset, int nnz, const int64_t* idx,
                           const double* val);
/* out[i*nmap+j] = sum_e xx[i][idx_e]*val_e over map j */
int itsolv_sparse_gemm_inner_f64(itsolv_ctx* ctx, const double* const* xx, int k, size_t n, size_t global_offset,
                                 int nmap, const int32_t* map_ptr, const int64_t* idx, const double* val, double* out);
/* yy[j][idx_e] += alpha[i*ndense+j]*val_e for every entry e of map i (alpha is nmap x ndense) */
int itsolv_sparse_gemm_outer_f64(itsolv_ctx* ctx, const double* alpha, int nmap, int ndense, const int32_t* map_ptr,
                                 const int64_t* idx, const double* val, double* const* yy, size_t n,
                                 size_t global_offset);

/* ---- harness operator kernels (the user's Problem::action, reference itsolv/IterativeSolver.h:100; not part of the
 * measured subspace path). Synthetic banded operator of SURVEY.md section 8(d):
 * A(i,i)=i+1, A(i,j)=eps*(1+((i+j) mod 7)) for 0<|i-j|<=b. x_lo/x_hi: b halo rows from the neighbouring shards
 * (NULL at the global ends). ---- */
int itsolv_banded_apply_f64(itsolv_ctx* ctx, int64_t n_global, int64_t row_offset, size_t n, int b, double eps,
                            const double* x, const double* x_lo, const double* x_hi, double* y);
/* stored CSR, 64-bit row pointers, columns global and within b of the local range */
int itsolv_csr_apply_f64(itsolv_ctx* ctx, int64_t n_global, int64_t row_offset, size_t n, int b, const int64_t* row_ptr,
                         const int32_t* col, const double* val, const double* x, const double* x_lo, const double* x_hi,
                         double* y);
/* the same for w vectors in one pass over the matrix (Problem::action receives the whole working set) */
int itsolv_csr_apply_multi_f64(itsolv_ctx* ctx, int64_t n_global, int64_t row_offset, size_t n, int b,
                               const int64_t* row_ptr, const int32_t* col, const double* val, int w,
                               const double* const* x, const double* const* x_lo, const double* const* x_hi,
                               double* const* y);
/* d[i] = row_offset+i+1 ; fills a vector with f(global index): kind 0 = diagonal,
 * 1 = u_k(i), 2 = ((k+1) + u_k(i)) / (i+1) with u_k(i) = ((i (k+2) + k) mod (2k+5)) / (2k+5) - 1/2 (known solutions of the
 * harness' LinearEquations right-hand sides, include/itsolv_b200_harness.h ITSOLV_RHS_*) */
int itsolv_banded_fill_f64(itsolv_ctx* ctx, int kind, int k, int64_t row_offset, size_t n, double* out);
/* P-space part of the action (reference Problem::p_action, itsolv/IterativeSolver.h:160-171; example
 * examples/ExampleProblemDistrArray.h:100-116): actions[k] += sum_p pcoef[k*nP+p] * A * P_p for the banded operator */
int itsolv_banded_p_action_f64(itsolv_ctx* ctx, int64_t n_global, int64_t row_offset, size_t n, int b, double eps,
                               int nact, double* const* actions, int nP, const int32_t* map_ptr, const int64_t* idx,
                               const double* val, const double* pcoef);
/* dense toy operator of the reference's examples/ExampleProblem.h:8 (single rank, small n): y = M x,
 * M(i,j) = i==j ? i+1 : 0.001*((i+j) mod n) */
int itsolv_example_apply_f64(itsolv_ctx* ctx, size_t n, const double* x, double* y);
/* Element-wise members of the reference's DistrArray that the solvers do not call but its conformance tests do
 * (array/DistrArray.cpp:79-167; test/array/testDistrArray.h:496-678): c is updated in place from a, b and a scalar. */
enum {
  ITSOLV_EW_ADD_SCALAR = 0,            /* c += s            add(value)  / sub(value)  */
  ITSOLV_EW_RECIP = 1,                 /* c = 1 / c         recip()                   */
  ITSOLV_EW_TIMES_INPLACE = 2,         /* c *= a            times(y)                  */
  ITSOLV_EW_TIMES = 3,                 /* c = a * b         times(y, z)               */
  ITSOLV_EW_DIVIDE = 4,                /* c = a / (b + s)   divide(y, z, s, false, false) */
  ITSOLV_EW_DIVIDE_NEGATIVE = 5,       /* c = -a / (b + s)  divide(y, z, s, false, true)  */
  ITSOLV_EW_DIVIDE_APPEND = 6,         /* c += a / (b + s)  divide(y, z, s, true, false)  */
  ITSOLV_EW_DIVIDE_APPEND_NEGATIVE = 7 /* c -= a / (b + s)  divide(y, z, s, true, true)   */
};
int itsolv_elementwise_f64(itsolv_ctx* ctx, int op, double* c, const double* a, const double* b, double scalar, size_t n);
/* out[i] = x[i] + c */
int itsolv_shift_f64(itsolv_ctx* ctx, double c, const double* x, double* out, size_t n);
/* out[i] = x[i] - t(row_offset + i): the argument of the harness' non-linear residual r(v) = A (v - t) (DIIS cases);
 * target_kind 0: t(i) = 1 / (i+1) (residual entries of order one in every row), 1: t(i) = 1 (legacy; its rounding floor
 * grows like n^1.5 and passes 1e-8 at n ~ 1e6) */
int itsolv_banded_target_shift_f64(itsolv_ctx* ctx, int target_kind, int64_t row_offset, const double* x, double* out,
                                   size_t n);

#ifdef __cplusplus
}
#endif
#endif
