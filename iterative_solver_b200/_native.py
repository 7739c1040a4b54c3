"""ctypes bindings of the two native libraries. Loading fails loudly when a library is missing: there is no Python or
CPU implementation of any operation in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(_HERE, "lib")

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)
c_void_pp = C.POINTER(C.c_void_p)

MAX_PANEL = 128
UNIQUE_ID_BYTES = 128
IPC_HANDLE_BYTES = 64
MAX_ROOTS = 64


class Counters(C.Structure):
    _fields_ = [("launches", C.c_int64)] + [(n, C.c_int64) for n in (
        "n_dot", "n_axpy", "n_scal", "n_copy", "n_fill", "n_gemm_inner", "n_gemm_outer", "n_precondition", "n_select",
        "n_sparse")] + [(n, C.c_double) for n in (
            "bytes", "device_seconds", "bytes_gemm_inner", "seconds_gemm_inner", "bytes_gemm_outer",
            "seconds_gemm_outer", "bytes_blas1", "seconds_blas1", "bytes_residual", "seconds_residual")] + [
                (n, C.c_int64) for n in ("calls_gemm_inner", "calls_gemm_outer", "calls_blas1", "calls_residual")]


class SolveSpec(C.Structure):
    """itsolv_solve_spec of include/itsolv_b200_harness.h"""
    _fields_ = [("n", C.c_int64), ("kind", C.c_int32), ("problem", C.c_int32), ("nroots", C.c_int32),
                ("nbuffers", C.c_int32), ("half_bandwidth", C.c_int32), ("hermitian", C.c_int32), ("eps", C.c_double),
                ("convergence_threshold", C.c_double), ("max_iter", C.c_int32), ("max_size_qspace", C.c_int32),
                ("reset_D", C.c_int32), ("max_p", C.c_int32), ("verbosity", C.c_int32), ("trace", C.c_int32),
                ("explicit_csr", C.c_int32), ("fused", C.c_int32), ("rhs_kind", C.c_int32)]


class SolveResult(C.Structure):
    """itsolv_solve_result of include/itsolv_b200_harness.h"""
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_int32), ("nroots", C.c_int32), ("nwork_final", C.c_int32),
                ("eigenvalues", C.c_double * MAX_ROOTS), ("errors", C.c_double * MAX_ROOTS),
                ("seconds_solve", C.c_double), ("seconds_action", C.c_double), ("seconds_precond", C.c_double),
                ("r_creations", C.c_int64), ("q_creations", C.c_int64), ("p_creations", C.c_int64),
                ("d_creations", C.c_int64), ("n_dot", C.c_int64), ("n_axpy", C.c_int64), ("n_scal", C.c_int64),
                ("n_copy", C.c_int64), ("n_fill", C.c_int64), ("n_gemm_inner", C.c_int64), ("n_gemm_outer", C.c_int64),
                ("handler_bytes", C.c_double), ("handler_device_seconds", C.c_double), ("kernel_launches", C.c_int64),
                ("device_ms_solve", C.c_double), ("bytes_gemm_inner", C.c_double), ("seconds_gemm_inner", C.c_double),
                ("bytes_gemm_outer", C.c_double), ("seconds_gemm_outer", C.c_double), ("bytes_blas1", C.c_double),
                ("seconds_blas1", C.c_double), ("bytes_residual", C.c_double), ("seconds_residual", C.c_double),
                ("calls_gemm_inner", C.c_int64), ("calls_gemm_outer", C.c_int64), ("calls_blas1", C.c_int64),
                ("calls_residual", C.c_int64)]


class TraceEntry(C.Structure):
    _fields_ = [("op", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32), ("reserved", C.c_int32),
                ("offset", C.c_int64)]


KIND_DAVIDSON, KIND_LINEQ, KIND_DIIS = 0, 1, 2
PROBLEM_BANDED, PROBLEM_EXAMPLE = 0, 1
RHS_SCALED, RHS_LEGACY = 0, 1

# name -> (restype, argtypes); the names are exactly the declarations of include/itsolv_b200.h
KERNEL_API = {
    "itsolv_ctx_create": (C.c_int, [C.c_int, c_void_pp]),
    "itsolv_ctx_create_on_stream": (C.c_int, [C.c_int, C.c_void_p, c_void_pp]),
    "itsolv_ctx_destroy": (None, [C.c_void_p]),
    "itsolv_last_error": (C.c_char_p, []),
    "itsolv_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "itsolv_ctx_device": (C.c_int, [C.c_void_p]),
    "itsolv_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "itsolv_ctx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "itsolv_ctx_counters": (None, [C.c_void_p, C.POINTER(Counters)]),
    "itsolv_ctx_reset_counters": (None, [C.c_void_p]),
    "itsolv_ctx_set_profiling": (None, [C.c_void_p, C.c_int]),
    "itsolv_ctx_timer_start": (C.c_int, [C.c_void_p, C.c_int]),
    "itsolv_ctx_timer_stop": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "itsolv_alloc": (C.c_int, [C.c_void_p, C.c_size_t, c_void_pp]),
    "itsolv_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "itsolv_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "itsolv_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "itsolv_upload_bytes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "itsolv_mem_usage": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_int]),
    "itsolv_mem_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "itsolv_mem_trim": (C.c_int, [C.c_void_p]),
    "itsolv_comm_unique_id": (C.c_int, [C.c_void_p]),
    "itsolv_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "itsolv_comm_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "itsolv_comm_p2p_import": (C.c_int, [C.c_void_p, C.c_void_p]),
    "itsolv_comm_p2p_disable": (C.c_int, [C.c_void_p]),
    "itsolv_comm_rank": (C.c_int, [C.c_void_p]),
    "itsolv_comm_size": (C.c_int, [C.c_void_p]),
    "itsolv_comm_barrier": (C.c_int, [C.c_void_p]),
    "itsolv_comm_allreduce_host": (C.c_int, [C.c_void_p, c_double_p, C.c_size_t, C.c_int]),
    "itsolv_comm_halo_exchange_multi": (C.c_int, [C.c_void_p, c_void_pp, C.c_int, C.c_size_t, C.c_int, C.c_void_p]),
    "itsolv_comm_halo_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "itsolv_distribution": (None, [C.c_size_t, C.c_int, c_int64_p]),
    "itsolv_fill_f64": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_size_t]),
    "itsolv_scal_f64": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_size_t]),
    "itsolv_copy_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "itsolv_axpy_f64": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_size_t]),
    "itsolv_dot_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, c_double_p]),
    "itsolv_scal_batch_f64": (C.c_int, [C.c_void_p, c_double_p, c_void_pp, C.c_int, C.c_size_t]),
    "itsolv_fill_batch_f64": (C.c_int, [C.c_void_p, c_double_p, c_void_pp, C.c_int, C.c_size_t]),
    "itsolv_axpy_batch_f64": (C.c_int, [C.c_void_p, c_double_p, c_void_pp, c_void_pp, C.c_int, C.c_size_t]),
    "itsolv_mgs_step_f64": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, c_double_p, c_void_pp, C.c_int, C.c_size_t]),
    "itsolv_ctx_write_epoch": (C.c_ulonglong, [C.c_void_p]),
    "itsolv_ctx_note_write": (None, [C.c_void_p]),
    "itsolv_gemm_inner_f64": (C.c_int, [C.c_void_p, c_void_pp, C.c_int, c_void_pp, C.c_int, C.c_size_t, c_double_p]),
    "itsolv_gemm_outer_f64": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, c_void_pp, c_void_pp, C.c_size_t,
                                        C.c_int]),
    "itsolv_mgs_step_dots_f64": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, c_double_p, c_void_pp, C.c_int,
                                           C.c_size_t, c_double_p]),
    "itsolv_mgs_chain_supported": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t]),
    "itsolv_mgs_chain_f64": (C.c_int, [C.c_void_p, c_void_pp, C.c_int, C.c_size_t, C.c_double, c_double_p]),
    "itsolv_project_mgs_chain_supported": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t]),
    "itsolv_project_mgs_chain_f64": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, c_void_pp, c_void_pp, c_double_p,
                                               C.POINTER(C.c_int), C.c_int, C.c_size_t, C.c_double, c_double_p]),
    "itsolv_gemm_outer_scaled_f64": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, c_void_pp, c_void_pp,
                                               C.c_size_t, c_double_p]),
    "itsolv_precondition_f64": (C.c_int, [C.c_void_p, c_void_pp, C.c_int, C.c_void_p, c_double_p, C.c_size_t]),
    "itsolv_davidson_residual_f64": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, c_void_pp, c_void_pp, c_double_p,
                                               C.c_void_p, c_double_p, c_void_pp, c_void_pp, C.c_size_t, c_double_p,
                                               c_double_p]),
    "itsolv_subspace_residual_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int, c_void_pp,
                                               c_void_pp, c_double_p, c_void_pp, c_double_p, C.c_void_p, c_double_p,
                                               c_void_pp, c_void_pp, C.c_size_t, c_double_p, c_double_p]),
    "itsolv_select_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                    C.c_int, c_int64_p, c_double_p, C.POINTER(C.c_int)]),
    "itsolv_select_merge": (C.c_int, [c_int64_p, c_double_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, c_int64_p,
                                      c_double_p]),
    "itsolv_sparse_copy_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, c_int64_p,
                                         c_double_p]),
    "itsolv_sparse_gemm_inner_f64": (C.c_int, [C.c_void_p, c_void_pp, C.c_int, C.c_size_t, C.c_size_t, C.c_int,
                                               c_int32_p, c_int64_p, c_double_p, c_double_p]),
    "itsolv_sparse_gemm_outer_f64": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, c_int32_p, c_int64_p,
                                               c_double_p, c_void_pp, C.c_size_t, C.c_size_t]),
    "itsolv_banded_apply_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_size_t, C.c_int, C.c_double,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "itsolv_csr_apply_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "itsolv_csr_apply_multi_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_size_t, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int, c_void_pp, c_void_pp, c_void_pp,
                                             c_void_pp]),
    "itsolv_banded_fill_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_size_t, C.c_void_p]),
    "itsolv_banded_p_action_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_size_t, C.c_int, C.c_double, C.c_int,
                                             c_void_pp, C.c_int, c_int32_p, c_int64_p, c_double_p, c_double_p]),
    "itsolv_example_apply_f64": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "itsolv_shift_f64": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_size_t]),
    "itsolv_elementwise_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_size_t]),
    "itsolv_banded_target_shift_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t]),
}

# declarations of include/itsolv_b200_harness.h
HARNESS_API = {
    "itsolv_harness_solve": (C.c_int, [C.c_void_p, C.POINTER(SolveSpec), C.POINTER(SolveResult), c_double_p]),
    "itsolv_harness_solve_host_csr": (C.c_int, [C.c_void_p, C.POINTER(SolveSpec), c_int64_p, c_int32_p, c_double_p,
                                                c_double_p, C.POINTER(SolveResult), c_double_p]),
    "itsolv_harness_last_error": (C.c_char_p, []),
    "itsolv_harness_problem_create": (C.c_int, [C.c_void_p, C.POINTER(SolveSpec), c_int64_p, c_int32_p, c_double_p,
                                                c_double_p, c_void_pp]),
    "itsolv_harness_problem_solve": (C.c_int, [C.c_void_p, C.POINTER(SolveSpec), C.POINTER(SolveResult), c_double_p]),
    "itsolv_harness_problem_solve_device": (C.c_int, [C.c_void_p, C.POINTER(SolveSpec), C.POINTER(SolveResult),
                                                      c_void_pp]),
    "itsolv_harness_problem_destroy": (None, [C.c_void_p]),
    "itsolv_harness_trace_entries": (C.c_size_t, []),
    "itsolv_harness_trace_values": (C.c_size_t, []),
    "itsolv_harness_trace_read": (None, [C.POINTER(TraceEntry), c_double_p]),
    "itsolv_handler_blas1": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_double, c_double_p, c_double_p,
                                       c_double_p]),
    "itsolv_handler_gemm_inner": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, c_double_p, c_double_p, C.c_int,
                                            c_double_p]),
    "itsolv_handler_gemm_outer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, c_double_p, c_double_p,
                                            c_double_p]),
    "itsolv_handler_select": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, c_double_p, c_double_p, C.c_int, C.c_int,
                                        c_int64_p, c_double_p]),
    "itsolv_handler_precondition": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, c_double_p, c_double_p, c_double_p]),
    "itsolv_handler_modified_gram_schmidt": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, c_double_p, C.c_double,
                                                       C.POINTER(C.c_int)]),
    "itsolv_handler_sparse_copy": (C.c_int, [C.c_void_p, C.c_size_t, c_double_p, C.c_int, c_int64_p, c_double_p]),
    "itsolv_handler_sparse_gemm_inner": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, c_double_p, c_int32_p,
                                                   c_int64_p, c_double_p, c_double_p]),
    "itsolv_handler_sparse_gemm_outer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, c_double_p, c_int32_p,
                                                   c_int64_p, c_double_p, c_double_p]),
    "itsolv_handler_distr_array": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_double, C.c_int, c_double_p, c_double_p,
                                             c_double_p, C.c_int, c_int64_p, c_double_p, c_double_p, c_int64_p, c_double_p]),
    "itsolv_harness_banded_apply": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, c_double_p,
                                              c_double_p]),
    "itsolv_host_eigenproblem": (C.c_int, [c_double_p, c_double_p, C.c_size_t, C.c_int, C.c_double, c_double_p,
                                           c_double_p, C.POINTER(C.c_size_t)]),
    "itsolv_host_svd_system": (C.c_int, [C.c_size_t, C.c_size_t, c_double_p, C.c_double, C.c_int, C.c_int, c_double_p,
                                         c_double_p, C.POINTER(C.c_size_t)]),
    "itsolv_host_solve_linear_equations": (C.c_int, [c_double_p, c_double_p, c_double_p, C.c_size_t, C.c_size_t,
                                                     C.c_double, C.c_double, c_double_p, c_double_p]),
    "itsolv_host_solve_diis": (C.c_int, [c_double_p, C.c_size_t, C.c_double, c_double_p]),
}


# declarations of include/itsolv_b200_solver.h (flat C interface of the solvers over device buffers)
c_size_p = C.POINTER(C.c_size_t)
APPLY_ON_P = C.CFUNCTYPE(None, c_double_p, C.c_void_p, C.c_size_t, c_size_p)
SOLVER_API = {
    "ItsolvB200LinearEigensystemInitialize": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, c_size_p, c_size_p,
                                                        C.c_double, C.c_double, C.c_int, C.c_int, C.c_char_p]),
    "ItsolvB200LinearEquationsInitialize": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, c_size_p, c_size_p,
                                                      C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                                      C.c_char_p]),
    "ItsolvB200NonLinearEquationsInitialize": (C.c_int, [C.c_void_p, C.c_size_t, c_size_p, c_size_p, C.c_double,
                                                         C.c_int, C.c_char_p]),
    "ItsolvB200Finalize": (C.c_int, []),
    "ItsolvB200AddVector": (C.c_long, [C.c_size_t, C.c_void_p, C.c_void_p]),
    "ItsolvB200Solution": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]),
    "ItsolvB200EndIteration": (C.c_long, [C.c_size_t, C.c_void_p, C.c_void_p]),
    "ItsolvB200EndIterationNeeded": (C.c_int, []),
    "ItsolvB200AddP": (C.c_long, [C.c_size_t, C.c_size_t, c_size_p, c_size_p, c_double_p, c_double_p, C.c_void_p,
                                  C.c_void_p, APPLY_ON_P]),
    "ItsolvB200Errors": (C.c_int, [c_double_p]),
    "ItsolvB200Eigenvalues": (C.c_int, [c_double_p]),
    "ItsolvB200WorkingSetEigenvalues": (C.c_int, [c_double_p]),
    "ItsolvB200WorkingSet": (C.c_long, [C.POINTER(C.c_int)]),
    "ItsolvB200SuggestP": (C.c_long, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, c_size_p]),
    "ItsolvB200PrintStatistics": (C.c_int, []),
    "ItsolvB200NonLinear": (C.c_int, []),
    "ItsolvB200HasValues": (C.c_int, []),
    "ItsolvB200Value": (C.c_double, []),
    "ItsolvB200HasEigenvalues": (C.c_int, []),
    "ItsolvB200SetDiagonals": (C.c_int, [C.c_void_p]),
    "ItsolvB200Diagonals": (C.c_int, [C.c_void_p]),
    "ItsolvB200PreconditionDefault": (C.c_int, [C.c_size_t, C.c_void_p]),
    "ItsolvB200Verbosity": (C.c_int, []),
    "ItsolvB200MaxIter": (C.c_int, []),
    "ItsolvB200SetMaxIter": (C.c_int, [C.c_int]),
    "ItsolvB200Iterations": (C.c_long, []),
    "ItsolvB200LastError": (C.c_char_p, []),
}


def _bind(lib: C.CDLL, table: dict) -> None:
    for name, (restype, argtypes) in table.items():
        fn = getattr(lib, name)  # AttributeError = the library does not export what the header declares
        fn.restype = restype
        fn.argtypes = argtypes


_kernels = None
_host = None


def kernels() -> C.CDLL:
    """libitsolv_b200.so (CUDA kernels + C ABI)."""
    global _kernels
    if _kernels is None:
        path = os.path.join(LIBDIR, "libitsolv_b200.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        _bind(lib, KERNEL_API)
        _kernels = lib
    return _kernels


def host() -> C.CDLL:
    """libitsolv_b200_host.so (DistrArrayCUDA/ArrayHandlerCUDA plugged into the reference's solver templates)."""
    global _host
    if _host is None:
        kernels()
        path = os.path.join(LIBDIR, "libitsolv_b200_host.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        _bind(lib, HARNESS_API)
        _bind(lib, SOLVER_API)
        _host = lib
    return _host
