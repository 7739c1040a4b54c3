"""ctypes access to the CPU oracle (oracle/): TEST INFRASTRUCTURE. `c` is the plain-C restatement
(oracle/itsolv_oracle.c), `ref` the reference's own templates compiled in place (oracle/ref_driver.cpp), or None when
oracle/_ref/libitsolv_ref.so has not been built (it needs /root/reference)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

dp = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)


def _d(a):
    return a.ctypes.data_as(dp)


def build(want_ref: bool = True) -> None:
    """make -C oracle: the C restatement always, the reference build when its sources are present."""
    targets = ["_ref/libitsolv_oracle_c.so"]
    if want_ref and os.path.isdir(os.environ.get("ITSOLV_REFERENCE", "/root/reference")):
        targets.append("ref")
    subprocess.run(["make", "-C", ORACLE_DIR] + targets, check=True, capture_output=True)


class COracle:
    def __init__(self, lib):
        self.lib = lib
        lib.oracle_dot.restype = C.c_double
        lib.oracle_dot.argtypes = [C.c_size_t, dp, dp]
        lib.oracle_dot_long.restype = C.c_longdouble
        lib.oracle_dot_long.argtypes = [C.c_size_t, dp, dp]
        lib.oracle_axpy.argtypes = [C.c_size_t, C.c_double, dp, dp]
        lib.oracle_scal.argtypes = [C.c_size_t, C.c_double, dp]
        lib.oracle_gemm_inner.argtypes = [C.c_int, C.c_int, C.c_size_t, dp, dp, dp]
        lib.oracle_gemm_outer.argtypes = [C.c_int, C.c_int, C.c_size_t, dp, dp, dp]
        lib.oracle_gemm_outer_fma.argtypes = [C.c_int, C.c_int, C.c_size_t, dp, dp, dp]
        lib.oracle_precondition.argtypes = [C.c_int, C.c_size_t, dp, dp, dp]
        lib.oracle_select.restype = C.c_int
        lib.oracle_select.argtypes = [C.c_size_t, C.c_size_t, dp, dp, C.c_int, C.c_int, i64p, dp]
        lib.oracle_modified_gram_schmidt.restype = C.c_int
        lib.oracle_modified_gram_schmidt.argtypes = [C.c_int, C.c_size_t, dp, C.c_double, C.POINTER(C.c_int)]
        lib.oracle_sparse_copy.argtypes = [C.c_size_t, dp, C.c_int, i64p, dp]
        lib.oracle_sparse_gemm_inner.argtypes = [C.c_int, C.c_int, C.c_size_t, dp, i32p, i64p, dp, dp]
        lib.oracle_sparse_gemm_outer.argtypes = [C.c_int, C.c_int, C.c_size_t, dp, i32p, i64p, dp, dp]
        lib.oracle_distribution.argtypes = [C.c_size_t, C.c_int, i64p]
        lib.oracle_banded_apply.argtypes = [C.c_int64, C.c_int, C.c_double, dp, dp]

    def dot(self, x, y):
        return self.lib.oracle_dot(x.size, _d(x), _d(y))

    def dot_long(self, x, y):
        return float(self.lib.oracle_dot_long(x.size, _d(x), _d(y)))

    def axpy(self, alpha, x, y):
        y = y.copy()
        self.lib.oracle_axpy(x.size, alpha, _d(x), _d(y))
        return y

    def scal(self, alpha, x):
        x = x.copy()
        self.lib.oracle_scal(x.size, alpha, _d(x))
        return x

    def gemm_inner(self, X, Y):
        X, Y = np.ascontiguousarray(X), np.ascontiguousarray(Y)
        out = np.zeros((X.shape[0], Y.shape[0]))
        self.lib.oracle_gemm_inner(X.shape[0], Y.shape[0], X.shape[1], _d(X), _d(Y), _d(out))
        return out

    def gemm_outer(self, alpha, X, Y, fma=False):
        X, Y = np.ascontiguousarray(X), np.ascontiguousarray(Y).copy()
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        fn = self.lib.oracle_gemm_outer_fma if fma else self.lib.oracle_gemm_outer
        fn(X.shape[0], Y.shape[0], X.shape[1], _d(a), _d(X), _d(Y))
        return Y

    def precondition(self, R, shift, diag):
        R = np.ascontiguousarray(R).copy()
        s = np.ascontiguousarray(shift, dtype=np.float64)
        self.lib.oracle_precondition(R.shape[0], R.shape[1], _d(R), _d(s), _d(diag))
        return R

    def select(self, x, nsel, max=False, ignore_sign=False, y=None):
        idx = np.zeros(nsel, dtype=np.int64)
        val = np.zeros(nsel)
        c = self.lib.oracle_select(nsel, x.size, _d(x), _d(y) if y is not None else None, int(max), int(ignore_sign),
                                   idx.ctypes.data_as(i64p), _d(val))
        return idx[:c], val[:c]

    def modified_gram_schmidt(self, V, thresh=1e-14):
        V = np.ascontiguousarray(V).copy()
        nulls = (C.c_int * max(1, V.shape[0]))()
        c = self.lib.oracle_modified_gram_schmidt(V.shape[0], V.shape[1], _d(V), thresh, nulls)
        return V, [nulls[i] for i in range(c)]

    def sparse_copy(self, x, m):
        from iterative_solver_b200.harness import pack_maps
        x = x.copy()
        ptr, idx, val = pack_maps([m])
        self.lib.oracle_sparse_copy(x.size, _d(x), idx.size, idx.ctypes.data_as(i64p), _d(val))
        return x

    def sparse_gemm_inner(self, X, maps):
        from iterative_solver_b200.harness import pack_maps
        X = np.ascontiguousarray(X)
        ptr, idx, val = pack_maps(maps)
        out = np.zeros((X.shape[0], len(maps)))
        self.lib.oracle_sparse_gemm_inner(X.shape[0], len(maps), X.shape[1], _d(X), ptr.ctypes.data_as(i32p),
                                          idx.ctypes.data_as(i64p), _d(val), _d(out))
        return out

    def sparse_gemm_outer(self, alpha, maps, Y):
        from iterative_solver_b200.harness import pack_maps
        Y = np.ascontiguousarray(Y).copy()
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        ptr, idx, val = pack_maps(maps)
        self.lib.oracle_sparse_gemm_outer(len(maps), Y.shape[0], Y.shape[1], _d(a), ptr.ctypes.data_as(i32p),
                                          idx.ctypes.data_as(i64p), _d(val), _d(Y))
        return Y

    def distribution(self, n, nchunks):
        b = np.zeros(nchunks + 1, dtype=np.int64)
        self.lib.oracle_distribution(n, nchunks, b.ctypes.data_as(i64p))
        return b

    def banded_apply(self, x, b, eps):
        y = np.zeros_like(x)
        self.lib.oracle_banded_apply(x.size, b, eps, _d(x), _d(y))
        return y


class RefOracle:
    """The reference's own ArrayHandlerIterable / solver templates (oracle/ref_driver.cpp)."""

    def __init__(self, lib):
        from iterative_solver_b200 import _native as N
        self.lib = lib
        self.N = N
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_solve.argtypes = [C.POINTER(N.SolveSpec), C.POINTER(N.SolveResult), dp]
        lib.ref_trace_entries.restype = C.c_size_t
        lib.ref_trace_values.restype = C.c_size_t
        lib.ref_trace_read.argtypes = [C.POINTER(N.TraceEntry), dp]
        lib.ref_banded_apply.argtypes = [C.c_int64, C.c_int, C.c_double, dp, dp]
        lib.ref_handler_dot.restype = C.c_double
        lib.ref_handler_dot.argtypes = [C.c_size_t, dp, dp]
        lib.ref_handler_axpy.argtypes = [C.c_size_t, C.c_double, dp, dp]
        lib.ref_handler_scal.argtypes = [C.c_size_t, C.c_double, dp]
        lib.ref_handler_gemm_inner.argtypes = [C.c_int, C.c_int, C.c_size_t, dp, dp, dp]
        lib.ref_handler_gemm_outer.argtypes = [C.c_int, C.c_int, C.c_size_t, dp, dp, dp]
        lib.ref_handler_select.argtypes = [C.c_size_t, C.c_size_t, dp, C.c_int, C.c_int, i64p, dp]
        lib.ref_handler_select_max_dot.argtypes = [C.c_size_t, C.c_size_t, dp, dp, i64p, dp]
        lib.ref_precondition_default.argtypes = [C.c_int, C.c_size_t, dp, dp, dp]
        lib.ref_modified_gram_schmidt.argtypes = [C.c_int, C.c_size_t, dp, C.c_double, C.POINTER(C.c_int)]
        lib.ref_sparse_copy.argtypes = [C.c_size_t, dp, C.c_int, i64p, dp]
        lib.ref_sparse_gemm_inner.argtypes = [C.c_int, C.c_int, C.c_size_t, dp, i32p, i64p, dp, dp]
        lib.ref_sparse_gemm_outer.argtypes = [C.c_int, C.c_int, C.c_size_t, dp, i32p, i64p, dp, dp]
        lib.ref_distribution.argtypes = [C.c_size_t, C.c_int, i64p]
        lib.ref_time_handler_op.restype = C.c_double
        lib.ref_time_handler_op.argtypes = [C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int]
        lib.ref_dense_eigen.argtypes = [C.c_size_t, dp, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, i64p]
        lib.ref_dense_lineq.argtypes = [C.c_size_t, dp, C.c_int, dp, C.c_double, dp, i64p]
        lib.ref_gram_schmidt.argtypes = [C.c_size_t, dp, dp, dp]
        lib.ref_eye_order.argtypes = [C.c_size_t, dp, i64p]
        lib.ref_overlap.argtypes = [C.c_int, C.c_size_t, dp, dp, dp]
        lib.ref_parameter_batches.argtypes = [C.c_size_t, C.c_size_t, i64p, C.c_int]
        lib.ref_max_overlap_with_R.argtypes = [C.c_int, C.c_int, C.c_size_t, dp, dp, i64p]
        lib.ref_dense_diis.argtypes = [C.c_size_t, C.c_double, dp, dp, i64p, C.POINTER(C.c_double)]

    def solve(self, spec, want_solutions=False):
        res = self.N.SolveResult()
        nroots = 1 if spec.kind == self.N.KIND_DIIS else spec.nroots
        sol = np.zeros((nroots, spec.n)) if want_solutions else None
        rc = self.lib.ref_solve(C.byref(spec), C.byref(res), _d(sol) if sol is not None else None)
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return res, sol

    def dense_eigen(self, hmat, nroot, np_=0, hermitian=True, n_working_vectors_max=0):
        """the reference's end-to-end eigensolver test protocol (test/itsolv/test_LinearEigensystem.cpp:245-344) on a dense
        matrix; returns eigenvalues, errors, solutions [nroot x n], (iterations, r_creations, n_iter)"""
        h = np.ascontiguousarray(hmat, dtype=np.float64)
        n = h.shape[0]
        ev, err, sol = np.zeros(nroot), np.zeros(nroot), np.zeros((nroot, n))
        stats = np.zeros(3, dtype=np.int64)
        i64p = C.POINTER(C.c_int64)
        rc = self.lib.ref_dense_eigen(n, _d(h), nroot, np_, int(hermitian), n_working_vectors_max, _d(ev), _d(err), _d(sol),
                                      stats.ctypes.data_as(i64p))
        if rc:
            raise RuntimeError("ref_dense_eigen failed")
        return ev, err, sol, tuple(int(v) for v in stats)

    def gram_schmidt(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        n = s.shape[0]
        t, norms = np.zeros((n, n)), np.zeros(n)
        assert self.lib.ref_gram_schmidt(n, _d(s), _d(t), _d(norms)) == 0
        return t, norms

    def eye_order(self, m):
        m = np.ascontiguousarray(m, dtype=np.float64)
        order = np.zeros(m.shape[0], dtype=np.int64)
        assert self.lib.ref_eye_order(m.shape[0], _d(m), order.ctypes.data_as(C.POINTER(C.c_int64))) == 0
        return order.tolist()

    def overlap(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        k, n = x.shape
        a, b = np.zeros((k, k)), np.zeros((k, k))
        assert self.lib.ref_overlap(k, n, _d(x), _d(a), _d(b)) == 0
        return a, b

    def parameter_batches(self, nsol, nparam):
        pairs = np.zeros(64, dtype=np.int64)
        nb = self.lib.ref_parameter_batches(nsol, nparam, pairs.ctypes.data_as(C.POINTER(C.c_int64)), 32)
        return [(int(pairs[2 * i]), int(pairs[2 * i + 1])) for i in range(nb)]

    def max_overlap_with_R(self, r, q):
        r = np.ascontiguousarray(r, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, r.shape[1])
        out = np.zeros(max(1, q.shape[0]), dtype=np.int64)
        assert self.lib.ref_max_overlap_with_R(r.shape[0], q.shape[0], r.shape[1], _d(r), _d(q),
                                               out.ctypes.data_as(C.POINTER(C.c_int64))) == 0
        return out[:q.shape[0]].tolist()

    def dense_lineq(self, matrix, rhs, threshold=1e-10):
        """the reference's LinearEquations test protocol (test/itsolv/test_LinearEquations.cpp:59-98)"""
        m = np.ascontiguousarray(matrix, dtype=np.float64)
        b = np.ascontiguousarray(rhs, dtype=np.float64)
        sol = np.zeros_like(b)
        stats = np.zeros(2, dtype=np.int64)
        if self.lib.ref_dense_lineq(m.shape[0], _d(m), b.shape[0], _d(b), threshold, _d(sol),
                                    stats.ctypes.data_as(C.POINTER(C.c_int64))):
            raise RuntimeError("ref_dense_lineq failed")
        return sol, int(stats[0]), bool(stats[1])

    def dense_diis(self, n, param):
        """the reference's NonLinearEquations test protocol (test/itsolv/test_NonLinearEquations.cpp:62-110)"""
        sol, res = np.zeros(n), np.zeros(n)
        stats = np.zeros(3, dtype=np.int64)
        err = C.c_double()
        if self.lib.ref_dense_diis(n, param, _d(sol), _d(res), stats.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(err)):
            raise RuntimeError("ref_dense_diis failed")
        return sol, res, tuple(int(v) for v in stats), err.value

    def read_trace(self):
        ne, nv = self.lib.ref_trace_entries(), self.lib.ref_trace_values()
        entries = (self.N.TraceEntry * max(ne, 1))()
        values = np.zeros(max(nv, 1))
        self.lib.ref_trace_read(entries, _d(values))
        return [(chr(entries[i].op), entries[i].rows, entries[i].cols,
                 values[entries[i].offset:entries[i].offset + entries[i].rows * entries[i].cols].copy())
                for i in range(ne)]

    def dot(self, x, y):
        return self.lib.ref_handler_dot(x.size, _d(x), _d(y))

    def axpy(self, alpha, x, y):
        y = y.copy()
        self.lib.ref_handler_axpy(x.size, alpha, _d(x), _d(y))
        return y

    def scal(self, alpha, x):
        x = x.copy()
        self.lib.ref_handler_scal(x.size, alpha, _d(x))
        return x

    def gemm_inner(self, X, Y):
        X, Y = np.ascontiguousarray(X), np.ascontiguousarray(Y)
        out = np.zeros((X.shape[0], Y.shape[0]))
        self.lib.ref_handler_gemm_inner(X.shape[0], Y.shape[0], X.shape[1], _d(X), _d(Y), _d(out))
        return out

    def gemm_outer(self, alpha, X, Y):
        X, Y = np.ascontiguousarray(X), np.ascontiguousarray(Y).copy()
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        self.lib.ref_handler_gemm_outer(X.shape[0], Y.shape[0], X.shape[1], _d(a), _d(X), _d(Y))
        return Y

    def precondition(self, R, shift, diag):
        R = np.ascontiguousarray(R).copy()
        s = np.ascontiguousarray(shift, dtype=np.float64)
        self.lib.ref_precondition_default(R.shape[0], R.shape[1], _d(R), _d(s), _d(diag))
        return R

    def select(self, x, nsel, max=False, ignore_sign=False, y=None):
        idx = np.zeros(nsel, dtype=np.int64)
        val = np.zeros(nsel)
        if y is not None:
            c = self.lib.ref_handler_select_max_dot(nsel, x.size, _d(x), _d(y), idx.ctypes.data_as(i64p), _d(val))
        else:
            c = self.lib.ref_handler_select(nsel, x.size, _d(x), int(max), int(ignore_sign), idx.ctypes.data_as(i64p),
                                            _d(val))
        return idx[:c], val[:c]

    def modified_gram_schmidt(self, V, thresh=1e-14):
        V = np.ascontiguousarray(V).copy()
        nulls = (C.c_int * max(1, V.shape[0]))()
        c = self.lib.ref_modified_gram_schmidt(V.shape[0], V.shape[1], _d(V), thresh, nulls)
        return V, [nulls[i] for i in range(c)]

    def sparse_copy(self, x, m):
        from iterative_solver_b200.harness import pack_maps
        x = x.copy()
        ptr, idx, val = pack_maps([m])
        self.lib.ref_sparse_copy(x.size, _d(x), idx.size, idx.ctypes.data_as(i64p), _d(val))
        return x

    def sparse_gemm_inner(self, X, maps):
        from iterative_solver_b200.harness import pack_maps
        X = np.ascontiguousarray(X)
        ptr, idx, val = pack_maps(maps)
        out = np.zeros((X.shape[0], len(maps)))
        self.lib.ref_sparse_gemm_inner(X.shape[0], len(maps), X.shape[1], _d(X), ptr.ctypes.data_as(i32p),
                                       idx.ctypes.data_as(i64p), _d(val), _d(out))
        return out

    def sparse_gemm_outer(self, alpha, maps, Y):
        from iterative_solver_b200.harness import pack_maps
        Y = np.ascontiguousarray(Y).copy()
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        ptr, idx, val = pack_maps(maps)
        self.lib.ref_sparse_gemm_outer(len(maps), Y.shape[0], Y.shape[1], _d(a), ptr.ctypes.data_as(i32p),
                                       idx.ctypes.data_as(i64p), _d(val), _d(Y))
        return Y

    def distribution(self, n, nchunks):
        b = np.zeros(nchunks + 1, dtype=np.int64)
        self.lib.ref_distribution(n, nchunks, b.ctypes.data_as(i64p))
        return b

    def banded_apply(self, x, b, eps):
        y = np.zeros_like(x)
        self.lib.ref_banded_apply(x.size, b, eps, _d(x), _d(y))
        return y

    def time_op(self, op, n, k, m, reps):
        return self.lib.ref_time_handler_op(op, n, k, m, reps)


_cache = None


def load():
    global _cache
    if _cache is not None:
        return _cache
    c_path = os.path.join(REF_DIR, "libitsolv_oracle_c.so")
    ref_path = os.path.join(REF_DIR, "libitsolv_ref.so")
    if not os.path.exists(c_path) or (not os.path.exists(ref_path) and os.path.isdir("/root/reference")):
        build()
    c = COracle(C.CDLL(c_path))
    ref = RefOracle(C.CDLL(ref_path)) if os.path.exists(ref_path) else None
    # the same reference templates with the PRODUCT's host algebra (oracle/Makefile): a CPU test vehicle
    ph_path = os.path.join(REF_DIR, "libitsolv_ref_producthelper.so")
    if not os.path.exists(ph_path) and os.path.isdir("/root/reference"):
        build()
    ref_product_helper = RefOracle(C.CDLL(ph_path)) if os.path.exists(ph_path) else None
    _cache = SimpleNamespace(c=c, ref=ref, ref_product_helper=ref_product_helper)
    return _cache
