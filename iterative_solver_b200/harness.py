"""Python face of the solve harness (include/itsolv_b200_harness.h): the reference's solvers running on
DistrArrayCUDA + ArrayHandlerCUDA. Used by tests, bench.py and __graft_entry__.smoke()."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .api import BackendError, Context, _dbl


def make_spec(n, kind=N.KIND_DAVIDSON, problem=N.PROBLEM_BANDED, nroots=1, nbuffers=0, half_bandwidth=4, hermitian=1,
              eps=1e-3, convergence_threshold=0.0, max_iter=0, max_size_qspace=0, reset_D=0, max_p=0, verbosity=0,
              trace=0, explicit_csr=0, fused=0, rhs_kind=N.RHS_SCALED) -> N.SolveSpec:
    return N.SolveSpec(n=n, kind=kind, problem=problem, nroots=nroots, nbuffers=nbuffers, half_bandwidth=half_bandwidth,
                       hermitian=hermitian, eps=eps, convergence_threshold=convergence_threshold, max_iter=max_iter,
                       max_size_qspace=max_size_qspace, reset_D=reset_D, max_p=max_p, verbosity=verbosity, trace=trace,
                       explicit_csr=explicit_csr, fused=fused, rhs_kind=rhs_kind)


def _hcheck(rc: int):
    if rc:
        raise BackendError(N.host().itsolv_harness_last_error().decode())


def solve(ctx: Context, spec: N.SolveSpec, want_solutions: bool = False):
    """Run the reference's solve() with the CUDA containers. Returns (result, solutions or None)."""
    lib = N.host()
    res = N.SolveResult()
    sol = None
    if want_solutions:
        borders = np.zeros(ctx.nranks + 1, dtype=np.int64)
        N.kernels().itsolv_distribution(spec.n, ctx.nranks, borders.ctypes.data_as(N.c_int64_p))
        nloc = int(borders[ctx.rank + 1] - borders[ctx.rank])
        nroots = 1 if spec.kind == N.KIND_DIIS else spec.nroots
        sol = np.zeros((nroots, nloc))
    _hcheck(lib.itsolv_harness_solve(ctx.handle, C.byref(spec), C.byref(res), _dbl(sol) if sol is not None else None))
    return res, sol


def banded_csr_host(n: int, b: int, eps: float, lo: int = 0, hi: int | None = None):
    """The synthetic banded operator as host CSR for rows [lo, hi): what a user would own."""
    hi = n if hi is None else hi
    rows = np.arange(lo, hi, dtype=np.int64)
    offs = np.arange(-b, b + 1, dtype=np.int64)
    cols = rows[:, None] + offs[None, :]
    valid = (cols >= 0) & (cols < n)
    vals = np.where(offs[None, :] == 0, (rows[:, None] + 1).astype(np.float64),
                    eps * (1 + ((rows[:, None] + cols) % 7)).astype(np.float64))
    row_ptr = np.zeros(hi - lo + 1, dtype=np.int64)
    np.cumsum(valid.sum(axis=1), out=row_ptr[1:])
    return row_ptr, cols[valid].astype(np.int32), vals[valid], (rows + 1).astype(np.float64)


def solve_host_csr(ctx: Context, spec: N.SolveSpec, row_ptr, col, val, diag, want_solutions: bool = True):
    """End-to-end entry: host CSR in, eigenvalues and solution vectors out (uploads and downloads inside)."""
    lib = N.host()
    res = N.SolveResult()
    nloc = row_ptr.size - 1
    nroots = 1 if spec.kind == N.KIND_DIIS else spec.nroots
    sol = np.zeros((nroots, nloc)) if want_solutions else None
    _hcheck(lib.itsolv_harness_solve_host_csr(ctx.handle, C.byref(spec), row_ptr.ctypes.data_as(N.c_int64_p),
                                              col.ctypes.data_as(N.c_int32_p), _dbl(val), _dbl(diag), C.byref(res),
                                              _dbl(sol) if sol is not None else None))
    return res, sol


class Problem:
    """The harness operator resident in HBM: create once, solve repeatedly (bench.py's device-resident leg)."""

    def __init__(self, ctx: Context, spec: N.SolveSpec, csr=None):
        self.ctx = ctx
        h = C.c_void_p()
        if csr is None:
            _hcheck(N.host().itsolv_harness_problem_create(ctx.handle, C.byref(spec), None, None, None, None, C.byref(h)))
        else:
            row_ptr, col, val, diag = csr
            _hcheck(N.host().itsolv_harness_problem_create(ctx.handle, C.byref(spec), row_ptr.ctypes.data_as(N.c_int64_p),
                                                           col.ctypes.data_as(N.c_int32_p), _dbl(val), _dbl(diag),
                                                           C.byref(h)))
        self.handle = h

    def solve(self, spec: N.SolveSpec, solutions: np.ndarray | int | None = None) -> N.SolveResult:
        """solutions: host array (nroots x n_local), or the address of DEVICE memory of that size (the vectors then stay
        on the GPU), or None"""
        res = N.SolveResult()
        if solutions is None:
            out = None
        elif isinstance(solutions, int):
            out = C.cast(C.c_void_p(solutions), N.c_double_p)
        else:
            out = _dbl(solutions)
        _hcheck(N.host().itsolv_harness_problem_solve(self.handle, C.byref(spec), C.byref(res), out))
        return res

    def solve_device(self, spec: N.SolveSpec):
        """solve and leave the solution vectors on the GPU: returns (result, device address of nroots x n_local doubles
        from the context's pool, to be released with Context.free)"""
        res = N.SolveResult()
        out = C.c_void_p()
        _hcheck(N.host().itsolv_harness_problem_solve_device(self.handle, C.byref(spec), C.byref(res), C.byref(out)))
        return res, int(out.value or 0)

    def close(self):
        if self.handle:
            N.host().itsolv_harness_problem_destroy(self.handle)
            self.handle = None


def read_trace():
    """Every dot / gemm_inner result the handlers returned during the last traced solve, in call order."""
    lib = N.host()
    ne, nv = lib.itsolv_harness_trace_entries(), lib.itsolv_harness_trace_values()
    entries = (N.TraceEntry * max(ne, 1))()
    values = np.zeros(max(nv, 1))
    lib.itsolv_harness_trace_read(entries, _dbl(values))
    return [(chr(entries[i].op), entries[i].rows, entries[i].cols,
             values[entries[i].offset:entries[i].offset + entries[i].rows * entries[i].cols].copy()) for i in range(ne)]


# ---- the handler contract through the C++ plugin classes, host arrays in and out (parity tests)

def handler_dot(ctx, x, y):
    r = C.c_double()
    _hcheck(N.host().itsolv_handler_blas1(ctx.handle, 0, x.size, 0.0, _dbl(x), _dbl(y), C.byref(r)))
    return r.value


def handler_axpy(ctx, alpha, x, y):
    y = y.copy()
    _hcheck(N.host().itsolv_handler_blas1(ctx.handle, 1, x.size, alpha, _dbl(x), _dbl(y), None))
    return y


def handler_scal(ctx, alpha, y):
    y = y.copy()
    _hcheck(N.host().itsolv_handler_blas1(ctx.handle, 2, y.size, alpha, None, _dbl(y), None))
    return y


def handler_fill(ctx, alpha, n):
    y = np.zeros(n)
    _hcheck(N.host().itsolv_handler_blas1(ctx.handle, 3, n, alpha, None, _dbl(y), None))
    return y


def handler_copy(ctx, x):
    y = np.zeros_like(x)
    _hcheck(N.host().itsolv_handler_blas1(ctx.handle, 4, x.size, 0.0, _dbl(x), _dbl(y), None))
    return y


def handler_gemm_inner(ctx, X, Y=None):
    X = np.ascontiguousarray(X)
    k, n = X.shape
    if Y is None:
        out = np.zeros((k, k))
        _hcheck(N.host().itsolv_handler_gemm_inner(ctx.handle, k, k, n, _dbl(X), None, 1, _dbl(out)))
        return out
    Y = np.ascontiguousarray(Y)
    m = Y.shape[0]
    out = np.zeros((k, m))
    _hcheck(N.host().itsolv_handler_gemm_inner(ctx.handle, k, m, n, _dbl(X), _dbl(Y), 0, _dbl(out)))
    return out


def handler_gemm_outer(ctx, alpha, X, Y):
    X = np.ascontiguousarray(X)
    Y = np.ascontiguousarray(Y).copy()
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    _hcheck(N.host().itsolv_handler_gemm_outer(ctx.handle, X.shape[0], Y.shape[0], X.shape[1], _dbl(a), _dbl(X), _dbl(Y)))
    return Y


def handler_select(ctx, x, nsel, max=False, ignore_sign=False, y=None):
    idx = np.zeros(nsel, dtype=np.int64)
    val = np.zeros(nsel)
    c = N.host().itsolv_handler_select(ctx.handle, nsel, x.size, _dbl(x), _dbl(y) if y is not None else None, int(max),
                                       int(ignore_sign), idx.ctypes.data_as(N.c_int64_p), _dbl(val))
    if c < 0:
        raise BackendError(N.host().itsolv_harness_last_error().decode())
    return idx[:c], val[:c]


def handler_precondition(ctx, R, shift, diag):
    R = np.ascontiguousarray(R).copy()
    s = np.ascontiguousarray(shift, dtype=np.float64)
    _hcheck(N.host().itsolv_handler_precondition(ctx.handle, R.shape[0], R.shape[1], _dbl(R), _dbl(s), _dbl(diag)))
    return R


def handler_modified_gram_schmidt(ctx, V, thresh=1e-14):
    V = np.ascontiguousarray(V).copy()
    nulls = (C.c_int * max(1, V.shape[0]))()
    c = N.host().itsolv_handler_modified_gram_schmidt(ctx.handle, V.shape[0], V.shape[1], _dbl(V), thresh, nulls)
    if c < 0:
        raise BackendError(N.host().itsolv_harness_last_error().decode())
    return V, [nulls[i] for i in range(c)]


def handler_distr_array(ctx, op, c, a=None, b=None, scalar=0.0, flags=0, sparse=None):
    """A member of the container itself (DistrArrayCUDA) through include/itsolv_b200_harness.h:
    itsolv_handler_distr_array. Returns (c after the call, scalar result, selection dict)."""
    c = np.ascontiguousarray(c, dtype=np.float64).copy()
    idx = np.asarray(sorted(sparse) if sparse else [], dtype=np.int64)
    val = np.asarray([sparse[k] for k in sorted(sparse)] if sparse else [], dtype=np.float64)
    res = C.c_double()
    sel_i = np.zeros(max(1, int(flags)), dtype=np.int64)
    sel_v = np.zeros(max(1, int(flags)))
    cnt = N.host().itsolv_handler_distr_array(
        ctx.handle, op, c.size, scalar, int(flags), _dbl(a) if a is not None else None, _dbl(b) if b is not None else None,
        _dbl(c), idx.size, idx.ctypes.data_as(N.c_int64_p), _dbl(val), C.byref(res), sel_i.ctypes.data_as(N.c_int64_p),
        _dbl(sel_v))
    if cnt < 0:
        raise BackendError(N.host().itsolv_harness_last_error().decode())
    return c, res.value, {int(sel_i[k]): float(sel_v[k]) for k in range(cnt)}


def pack_maps(maps):
    ptr = np.zeros(len(maps) + 1, dtype=np.int32)
    idx, val = [], []
    for j, m in enumerate(maps):
        for k in sorted(m):
            idx.append(k)
            val.append(m[k])
        ptr[j + 1] = len(idx)
    return ptr, np.asarray(idx, dtype=np.int64).reshape(-1), np.asarray(val, dtype=np.float64).reshape(-1)


def handler_sparse_copy(ctx, x, m):
    x = x.copy()
    ptr, idx, val = pack_maps([m])
    _hcheck(N.host().itsolv_handler_sparse_copy(ctx.handle, x.size, _dbl(x), idx.size, idx.ctypes.data_as(N.c_int64_p),
                                                _dbl(val)))
    return x


def handler_sparse_gemm_inner(ctx, X, maps):
    X = np.ascontiguousarray(X)
    ptr, idx, val = pack_maps(maps)
    out = np.zeros((X.shape[0], len(maps)))
    _hcheck(N.host().itsolv_handler_sparse_gemm_inner(ctx.handle, X.shape[0], len(maps), X.shape[1], _dbl(X),
                                                      ptr.ctypes.data_as(N.c_int32_p), idx.ctypes.data_as(N.c_int64_p),
                                                      _dbl(val), _dbl(out)))
    return out


def handler_sparse_gemm_outer(ctx, alpha, maps, Y):
    Y = np.ascontiguousarray(Y).copy()
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    ptr, idx, val = pack_maps(maps)
    _hcheck(N.host().itsolv_handler_sparse_gemm_outer(ctx.handle, len(maps), Y.shape[0], Y.shape[1], _dbl(a),
                                                      ptr.ctypes.data_as(N.c_int32_p), idx.ctypes.data_as(N.c_int64_p),
                                                      _dbl(val), _dbl(Y)))
    return Y


def harness_banded_apply(ctx, x, b, eps, explicit_csr=False):
    y = np.zeros_like(x)
    _hcheck(N.host().itsolv_harness_banded_apply(ctx.handle, x.size, b, eps, int(explicit_csr), _dbl(x), _dbl(y)))
    return y
