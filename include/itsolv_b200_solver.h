/*
 * itsolv_b200_solver — flat C interface of the iterative solvers over DEVICE buffers.
 *
 * It takes the place of the reference's C interface (reference src/molpro/linalg/IterativeSolverC.h:6-73, implemented
 * in IterativeSolverCMPI.cpp over DistrArraySpan views of the caller's HOST buffers) for a caller whose vectors live in
 * GPU memory: same call sequence, same argument meaning, same "only the top instance is active" stack of solvers
 * (IterativeSolverCMPI.cpp:59,76), with the CUDA containers (DistrArrayCUDA views of the caller's memory, ArrayHandlerCUDA)
 * underneath and the fused Davidson driver for the eigensolver.
 *
 * Differences from the reference's interface, all forced by where the data lives:
 *   - `parameters`, `action`, `rhs`, ... are DEVICE pointers to `buffer_size` vectors stored one after the other with a
 *     stride of THIS RANK'S shard length (range_end - range_begin as returned by the Initialize call), not the global
 *     length: a rank never holds rows of another rank. Consequently there is no `sync` argument (the reference gathers
 *     the full vectors to every rank when it is set, IterativeSolverCMPI.cpp:109-115).
 *   - the first argument of every Initialize call is the context (include/itsolv_b200.h) that owns the GPU, the stream
 *     and the communicator, instead of a Fortran MPI communicator handle.
 *   - errors do not propagate as C++ exceptions: functions return -1 (or non-zero) and ItsolvB200LastError() describes
 *     the failure.
 * All ranks of the context's communicator call every function collectively, as with the reference's MPI build.
 */
#ifndef ITSOLV_B200_SOLVER_H
#define ITSOLV_B200_SOLVER_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct itsolv_ctx; /* include/itsolv_b200.h */

/* options: blank- or comma-separated key=value pairs; understood: max_size_qspace, reset_D, reset_D_max_Q_size, max_iter,
 * fused (eigensolver: 1 by default = the fused driver, 0 = the reference's class call for call; equation solvers: 0 by
 * default, 1 = fused X space). Unknown keys are an error. */

/* reference IterativeSolverC.h:6-9; n = global length of the vectors */
int ItsolvB200LinearEigensystemInitialize(struct itsolv_ctx* ctx, size_t n, size_t nroot, size_t* range_begin,
                                          size_t* range_end, double thresh, double thresh_value, int hermitian,
                                          int verbosity, const char* options);
/* reference IterativeSolverC.h:11-15; rhs: nroot right-hand sides on the device (stride = shard length) */
int ItsolvB200LinearEquationsInitialize(struct itsolv_ctx* ctx, size_t n, size_t nroot, size_t* range_begin,
                                        size_t* range_end, const double* rhs, double aughes, double thresh,
                                        double thresh_value, int hermitian, int verbosity, const char* options);
/* reference IterativeSolverC.h:17-19 (DIIS) */
int ItsolvB200NonLinearEquationsInitialize(struct itsolv_ctx* ctx, size_t n, size_t* range_begin, size_t* range_end,
                                           double thresh, int verbosity, const char* options);
/* reference IterativeSolverC.h:25 */
int ItsolvB200Finalize(void);

/* reference IterativeSolverC.h:27: takes buffer_size parameter vectors and their actions, returns the size of the working
 * set; on return the first `working set` vectors hold solutions and residuals. -1 on error. */
long ItsolvB200AddVector(size_t buffer_size, double* parameters, double* action);
/* reference IterativeSolverC.h:29 */
int ItsolvB200Solution(int nroot, const int* roots, double* parameters, double* action);
/* reference IterativeSolverC.h:33: takes the (preconditioned) residuals, returns the number of new parameter vectors
 * written to `solution`. -1 on error. */
long ItsolvB200EndIteration(size_t buffer_size, double* solution, double* residual);
/* reference IterativeSolverC.h:35 */
int ItsolvB200EndIterationNeeded(void);
/* reference IterativeSolverC.h:37-40: P space of nP sparse vectors (offsets[nP+1] into indices/coefficients), pp = the
 * nP x nP action matrix; func(p_coefficients, action_device_pointer, update_size, ranges) adds the P-space part of the
 * action to the `update_size` action vectors, ranges = (begin, end) of this rank's rows per vector */
long ItsolvB200AddP(size_t buffer_size, size_t nP, const size_t* offsets, const size_t* indices,
                    const double* coefficients, const double* pp, double* parameters, double* action,
                    void (*func)(const double*, double*, const size_t, const size_t*));
/* reference IterativeSolverC.h:42-46; the arrays are HOST arrays sized by the caller (number of roots / working set) */
int ItsolvB200Errors(double* errors);
int ItsolvB200Eigenvalues(double* eigenvalues);
int ItsolvB200WorkingSetEigenvalues(double* eigenvalues);
/* roots of the current working set (host array of working-set size); returns the size */
long ItsolvB200WorkingSet(int* roots);
/* reference IterativeSolverC.h:48-49: indices (global) of up to maximumNumber elements proposed for the P space from the
 * n_roots solution and residual vectors on the device; returns their number (the solvers of this path propose none,
 * reference IterativeSolverTemplate.h:238-241). -1 on error. */
long ItsolvB200SuggestP(const double* solution, const double* residual, size_t maximumNumber, double threshold,
                        size_t* indices);
/* reference IterativeSolverC.h:51: the active solver's statistics on stdout */
int ItsolvB200PrintStatistics(void);
/* reference IterativeSolverC.h:53-56; HasValues is 1 for Optimize instances only, hence always 0 here */
int ItsolvB200NonLinear(void);
int ItsolvB200HasValues(void);
int ItsolvB200HasEigenvalues(void);
/* reference IterativeSolverC.h:62: the current function value of the active solver (NaN on error) */
double ItsolvB200Value(void);
/* NOT PROVIDED: IterativeSolverOptimizeInitialize and IterativeSolverAddValue (reference IterativeSolverC.h:21-23,31): the
 * BFGS / steepest-descent optimiser is outside the accelerated path; the mpicomm_* helpers
 * (:69-73) have no counterpart because the communicator belongs to the context. */
/* reference IterativeSolverC.h:58-60: the diagonal is kept in a device vector owned by the instance */
int ItsolvB200SetDiagonals(const double* diagonals);
int ItsolvB200Diagonals(double* diagonals);
/* Davidson update of the first `nwork` residuals with the stored diagonal and the working set's eigenvalues
 * (reference precondition_default, itsolv/IterativeSolver.h:46-55; in the reference's Fortran binding the caller does
 * this on the host between AddVector and EndIteration) */
int ItsolvB200PreconditionDefault(size_t nwork, double* residual);
/* reference IterativeSolverC.h:64-67 */
int ItsolvB200Verbosity(void);
int ItsolvB200MaxIter(void);
int ItsolvB200SetMaxIter(int max_iter);
/* iterations counted by the active solver (reference Statistics::iterations) */
long ItsolvB200Iterations(void);
const char* ItsolvB200LastError(void);

#ifdef __cplusplus
}
#endif
#endif
