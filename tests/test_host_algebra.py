"""The host subspace algebra (reference itsolv/helper-implementation.h:263-669) of the PRODUCT
(iterative_solver_b200/host/helper_lapack.cpp, through the C entry points of include/itsolv_b200_harness.h) and of the
ORACLE (oracle/helper_literal.cpp) against fixtures generated with numpy/scipy alone
(tests/golden/make_helper_golden.py -> helper_golden.npz). The two restatements share no code, and the fixtures share
none with either, so each is pinned independently. No GPU is needed: the functions run on the host."""
import ctypes as C
import os
import time

import numpy as np
import pytest

import itsolv_oracle_lib
from iterative_solver_b200 import _native as N

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "helper_golden.npz"))
dp = C.POINTER(C.c_double)
szp = C.POINTER(C.c_size_t)


def _d(a):
    return a.ctypes.data_as(dp)


class Algebra:
    """the four entry points of one library, `prefix`_{eigenproblem, svd_system, solve_linear_equations, solve_diis}"""

    def __init__(self, lib, prefix):
        self.f = {}
        sig = {"eigenproblem": [dp, dp, C.c_size_t, C.c_int, C.c_double, dp, dp, szp],
               "svd_system": [C.c_size_t, C.c_size_t, dp, C.c_double, C.c_int, C.c_int, dp, dp, szp],
               "solve_linear_equations": [dp, dp, dp, C.c_size_t, C.c_size_t, C.c_double, C.c_double, dp, dp],
               "solve_diis": [dp, C.c_size_t, C.c_double, dp]}
        for name, args in sig.items():
            fn = getattr(lib, f"{prefix}_{name}")
            fn.restype, fn.argtypes = C.c_int, args
            self.f[name] = fn

    def eigenproblem(self, h, s, n, hermitian, thr=1e-14):
        h, s = np.ascontiguousarray(h, dtype=np.float64), np.ascontiguousarray(s, dtype=np.float64)
        w, v, found = np.zeros(n), np.zeros(n * n), C.c_size_t()
        assert self.f["eigenproblem"](_d(h), _d(s), n, int(hermitian), thr, _d(w), _d(v), C.byref(found)) == 0
        r = found.value
        return w[:r].copy(), v[:n * r].reshape(r, n).T.copy()  # columns = eigenvectors

    def svd_system(self, m, rows, cols, thr, hermitian, reduce_to_rank=False):
        m = np.ascontiguousarray(m, dtype=np.float64)
        vals, vecs, found = np.zeros(cols), np.zeros(cols * cols), C.c_size_t()
        assert self.f["svd_system"](rows, cols, _d(m), thr, int(hermitian), int(reduce_to_rank), _d(vals), _d(vecs),
                                    C.byref(found)) == 0
        k = found.value
        return vals[:k].copy(), vecs[:k * cols].reshape(k, cols).copy()

    def solve_linear_equations(self, a, s, rhs, n, nroot, aug=0.0):
        x, e = np.zeros(n * nroot), np.zeros(nroot)
        a, s, rhs = (np.ascontiguousarray(t, dtype=np.float64) for t in (a, s, rhs))
        assert self.f["solve_linear_equations"](_d(a), _d(s), _d(rhs), n, nroot, aug, 1e-14, _d(x), _d(e)) == 0
        return x, e

    def solve_diis(self, b, n):
        c = np.zeros(n)
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert self.f["solve_diis"](_d(b), n, 1e-10, _d(c)) == 0
        return c


@pytest.fixture(scope="module", params=["product", "oracle"])
def algebra(request):
    if request.param == "product":
        return Algebra(N.host(), "itsolv_host")
    o = itsolv_oracle_lib.load()
    if o.ref is None:
        pytest.skip("oracle/_ref/libitsolv_ref.so is not built on this box")
    return Algebra(o.ref.lib, "ref_host")


@pytest.mark.parametrize("n", [1, 2, 5, 12, 40, 150])
def test_eigenproblem_hermitian(algebra, n):
    h, s = GOLD[f"eig_herm_{n}_H"], GOLD[f"eig_herm_{n}_S"]
    w, v = algebra.eigenproblem(h, s, n, True)
    want_w, want_v = GOLD[f"eig_herm_{n}_w"], GOLD[f"eig_herm_{n}_v"].reshape(n, n).T
    assert w.size == n
    assert np.abs(w - want_w).max() <= 1e-10 * max(1.0, np.abs(want_w).max())
    assert np.abs(v - want_v).max() <= 1e-8 * np.abs(want_v).max(), "eigenvectors (S-normalised, sign convention)"
    sm = s.reshape(n, n)
    assert np.abs(v.T @ sm @ v - np.eye(n)).max() <= 1e-10


@pytest.mark.parametrize("n", [2, 6, 20])
def test_eigenproblem_non_hermitian(algebra, n):
    h, s = GOLD[f"eig_gen_{n}_H"], GOLD[f"eig_gen_{n}_S"]
    w, v = algebra.eigenproblem(h, s, n, False)
    want_w, want_v = GOLD[f"eig_gen_{n}_w"], GOLD[f"eig_gen_{n}_v"].reshape(n, n).T
    assert np.abs(w - want_w).max() <= 1e-9 * np.abs(want_w).max()
    assert np.abs(v - want_v).max() <= 1e-7 * np.abs(want_v).max()
    hm, sm = h.reshape(n, n), s.reshape(n, n).T
    assert np.abs(hm @ v - (sm @ v) * w[None, :]).max() <= 1e-9 * np.abs(hm).max()


@pytest.mark.parametrize("n", [8, 30])
def test_eigenproblem_rank_deficient_metric(algebra, n):
    h, s, thr = GOLD[f"eig_def_{n}_H"], GOLD[f"eig_def_{n}_S"], float(GOLD[f"eig_def_{n}_thr"][0])
    w, v = algebra.eigenproblem(h, s, n, True, thr)
    want = GOLD[f"eig_def_{n}_w"]
    assert w.size == int(GOLD[f"eig_def_{n}_rank"][0]) == want.size
    assert np.abs(np.sort(w) - want).max() <= 1e-8 * max(1.0, np.abs(want).max())
    assert np.all(np.diff(w) >= 0), "ascending order"


@pytest.mark.parametrize("n", [6, 25])
def test_svd_system_hermitian_null_space(algebra, n):
    m = GOLD[f"svd_herm_{n}_M"]
    vals, vecs = algebra.svd_system(m, n, n, 1e-12, True)
    null = int(GOLD[f"svd_herm_{n}_null"][0])
    assert vals.size == null == GOLD[f"svd_herm_{n}_values"].size
    assert np.abs(vals).max() <= 1e-12
    # the null space itself is what is pinned (a basis of a degenerate eigenspace is not unique): projectors agree
    want = GOLD[f"svd_herm_{n}_vectors"].reshape(null, n)
    assert np.abs(vecs.T @ vecs - want.T @ want).max() <= 1e-8
    mm = m.reshape(n, n)
    assert np.abs(mm @ vecs.T).max() <= 1e-11


@pytest.mark.parametrize("shape", ["7x4", "5x5"])
def test_svd_system_general(algebra, shape):
    rows, cols = (int(t) for t in shape.split("x"))
    m, thr = GOLD[f"svd_gen_{shape}_M"], float(GOLD[f"svd_gen_{shape}_thr"][0])
    vals, vecs = algebra.svd_system(m, rows, cols, thr, False)
    want_vals, want_vecs = GOLD[f"svd_gen_{shape}_values"], GOLD[f"svd_gen_{shape}_vectors"].reshape(-1, cols)
    assert vals.size == want_vals.size
    assert np.abs(vals - want_vals).max() <= 1e-12
    for a, b in zip(vecs, want_vecs):  # singular vectors up to sign
        assert min(np.abs(a - b).max(), np.abs(a + b).max()) <= 1e-6


@pytest.mark.parametrize("n,nroot", [(1, 1), (7, 3), (60, 8)])
def test_solve_linear_equations(algebra, n, nroot):
    a, rhs = GOLD[f"lineq_{n}_A"], GOLD[f"lineq_{n}_rhs"]
    x, _ = algebra.solve_linear_equations(a, np.eye(n).ravel(), rhs, n, nroot)
    want = GOLD[f"lineq_{n}_x"]
    assert np.abs(x - want).max() <= 1e-11 * max(1.0, np.abs(want).max())
    # augmented hessian
    s, rhs_cm = GOLD[f"lineq_aug_{n}_S"], GOLD[f"lineq_aug_{n}_rhs"]
    x, e = algebra.solve_linear_equations(a, s, rhs_cm, n, nroot, aug=0.7)
    assert np.abs(e - GOLD[f"lineq_aug_{n}_e"]).max() <= 1e-9 * max(1.0, np.abs(GOLD[f"lineq_aug_{n}_e"]).max())
    assert np.abs(x - GOLD[f"lineq_aug_{n}_x"]).max() <= 1e-8 * max(1.0, np.abs(GOLD[f"lineq_aug_{n}_x"]).max())


@pytest.mark.parametrize("n", [1, 3, 8])
def test_solve_diis(algebra, n):
    c = algebra.solve_diis(GOLD[f"diis_{n}_B"], n)
    want = GOLD[f"diis_{n}_c"]
    assert abs(c.sum() - 1.0) <= 1e-10, "DIIS coefficients sum to one"
    assert np.abs(c - want).max() <= 1e-8 * max(1.0, np.abs(want).max())


def test_product_and_oracle_agree_on_random_subspace_problems():
    """the fast and the literal restatement on the same inputs, shapes as they occur in a solve (k = 4 ... 48)"""
    o = itsolv_oracle_lib.load()
    if o.ref is None:
        pytest.skip("oracle/_ref/libitsolv_ref.so is not built on this box")
    fast, literal = Algebra(N.host(), "itsolv_host"), Algebra(o.ref.lib, "ref_host")
    rng = np.random.default_rng(5)
    for n in (4, 8, 16, 24, 48):
        q, _ = np.linalg.qr(rng.standard_normal((200, n)))
        b = q.T + 0.05 * rng.standard_normal((n, 200))
        s = b @ b.T
        d = np.arange(1, 201, dtype=float)
        h = (b * d) @ b.T
        for herm in (True, False):
            w1, v1 = fast.eigenproblem(h, s.ravel(order="F"), n, herm)
            w2, v2 = literal.eigenproblem(h, s.ravel(order="F"), n, herm)
            assert np.abs(w1 - w2).max() <= 1e-11 * np.abs(w2).max()
            assert np.abs(v1 - v2).max() <= 1e-9 * np.abs(v2).max()


def test_eigenproblem_time_at_the_p_space_size():
    """the k = 560 subspace problem of BASELINE.json configs[4] (P space of 500 + Q); the literal
    restatement needs ~2 s for it"""
    n = 560
    rng = np.random.default_rng(9)
    q, _ = np.linalg.qr(rng.standard_normal((2 * n, n)))
    b = q.T + 0.01 * rng.standard_normal((n, 2 * n))
    s = b @ b.T
    h = (b * np.arange(1, 2 * n + 1, dtype=float)) @ b.T
    fast = Algebra(N.host(), "itsolv_host")
    fast.eigenproblem(h, s.ravel(order="F"), n, True)
    t0 = time.perf_counter()
    w, v = fast.eigenproblem(h, s.ravel(order="F"), n, True)
    dt = time.perf_counter() - t0
    want = np.linalg.eigvalsh(np.linalg.solve(np.linalg.cholesky(s), np.linalg.solve(np.linalg.cholesky(s), h).T))
    assert np.abs(w - want).max() <= 1e-9 * np.abs(want).max()
    assert dt < 0.5, f"eigenproblem at k=560 took {dt:.3f} s"
