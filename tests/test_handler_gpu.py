"""The reference's handler contract exercised through the C++ plugin classes (DistrArrayCUDA + ArrayHandlerCUDA /
ArrayHandlerCUDASparse, iterative_solver_b200/host/) with host buffers in and out, against the reference's own
ArrayHandlerIterable / ArrayHandlerIterableSparse compiled in place (oracle/_ref) — or the C restatement when that
build is absent — and against the golden vectors of the reference's own tests. Tolerances as in test_kernels_gpu.py."""
import numpy as np
import pytest

from iterative_solver_b200 import harness as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cpu(oracle):
    return oracle.ref if oracle.ref is not None else oracle.c


def test_kat_modified_gram_schmidt(ctx):
    """reference test/itsolv/subspace/test_util.cpp:154-173, run on the CUDA handler by the reference's own template"""
    V = np.array([[1., 1., 1., 1.], [1., 1. / 5., 1. / 10., 1. / 15.], [1. / 3., 1. / 6., 1. / 9., 1. / 12.],
                  [1. / 2., 1. / 4., 1. / 6., 1. / 8.]])
    ref = np.array([[0.5, 0.5, 0.5, 0.5],
                    [0.858898629520365, -0.184826287365142, -0.3152919019758303, -0.3587804401793933],
                    [-0.1096025454090415, 0.783967708523809, -0.06699956264207202, -0.6073656004726955]])
    out, nulls = H.handler_modified_gram_schmidt(ctx, V, 1e-14)
    assert nulls == [3]
    assert np.abs(out[:3] - ref).max() <= 1e-14


def test_kat_select_max_dot(ctx):
    """reference test/array/testArrayHandlerIterable.cpp:57-69"""
    x = np.array([1., -2., 1., 0., 3., 0., -4., 1.])
    y = np.ones(8)
    idx, val = H.handler_select(ctx, x, 3, y=y)
    assert dict(zip(idx.tolist(), val.tolist())) == {6: 4.0, 4: 3.0, 1: 2.0}


def test_kat_sparse_axpy_dot(ctx):
    """reference test/array/testArrayHandlers.cpp:27-60"""
    m = {1: 1.0, 3: 2.0, 6: 3.0, 11: 4.0}
    y = np.full(20, 0.5)
    out = H.handler_sparse_gemm_outer(ctx, np.array([[2.0]]), [m], y[None, :])[0]
    want = y.copy()
    for i, v in m.items():
        want[i] += 2.0 * v
    assert np.array_equal(out, want)
    d = H.handler_sparse_gemm_inner(ctx, np.full((1, 20), 0.5), [m])[0, 0]
    assert d == 0.5 + 0.5 * 2. + 0.5 * 3. + 0.5 * 4.


@pytest.mark.parametrize("n", [1, 20, 1001, 30000])
def test_blas1_through_handler(ctx, cpu, n):
    rng = np.random.default_rng(n)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    assert np.array_equal(H.handler_axpy(ctx, -0.3, x, y), cpu.axpy(-0.3, x, y))
    assert np.array_equal(H.handler_scal(ctx, 3.1, y), cpu.scal(3.1, y))
    assert np.array_equal(H.handler_copy(ctx, x), x)
    assert np.array_equal(H.handler_fill(ctx, 7.0, n), np.full(n, 7.0))
    d = H.handler_dot(ctx, x, y)
    assert abs(d - cpu.dot(x, y)) <= 1e-12 * np.linalg.norm(x) * np.linalg.norm(y)


@pytest.mark.parametrize("k,m,n", [(4, 16, 2000), (16, 20, 999), (1, 1, 5), (4, 1, 40001), (8, 8, 1)])
def test_gemm_through_handler(ctx, cpu, k, m, n):
    rng = np.random.default_rng(k + m + n)
    X, Y = rng.standard_normal((k, n)), rng.standard_normal((m, n))
    G = H.handler_gemm_inner(ctx, X, Y)
    want = cpu.gemm_inner(X, Y)
    scale = np.linalg.norm(X, axis=1)[:, None] * np.linalg.norm(Y, axis=1)[None, :]
    assert (np.abs(G - want) <= 1e-12 * scale).all()
    S = H.handler_gemm_inner(ctx, X)  # overlap with itself
    assert np.array_equal(S, S.T)
    alpha = rng.standard_normal((k, m))
    out = H.handler_gemm_outer(ctx, alpha, X, Y)
    ref = cpu.gemm_outer(alpha, X, Y)
    assert (np.abs(out - ref) <= 1e-14 * (np.abs(Y) + np.abs(alpha).T @ np.abs(X))).all()


def test_select_through_handler(ctx, cpu):
    rng = np.random.default_rng(9)
    x = np.round(rng.standard_normal(5000), 1)
    for kw in ({}, {"max": True}, {"max": True, "ignore_sign": True}, {"ignore_sign": True}):
        gi, gv = H.handler_select(ctx, x, 37, **kw)
        wi, wv = cpu.select(x, 37, **kw)
        assert np.array_equal(gi, wi) and np.array_equal(gv, wv)


def test_precondition_through_problem(ctx, cpu):
    rng = np.random.default_rng(10)
    R = rng.standard_normal((4, 3001))
    diag = np.arange(1, 3002, dtype=np.float64)
    shift = np.array([0.99, 1.98, 2.97, 3.96])
    assert np.array_equal(H.handler_precondition(ctx, R, shift, diag), cpu.precondition(R, shift, diag))


def test_sparse_through_handler(ctx, cpu):
    rng = np.random.default_rng(14)
    n = 4000
    X = rng.standard_normal((5, n))
    maps = [{int(i): float(v) for i, v in zip(rng.choice(n, 3, replace=False), rng.standard_normal(3))} for _ in range(7)]
    maps.append({0: 1.0})
    maps.append({n - 1: -2.0, 17: 0.5})
    assert np.array_equal(H.handler_sparse_gemm_inner(ctx, X, maps), cpu.sparse_gemm_inner(X, maps))
    alpha = rng.standard_normal((len(maps), 5))
    assert np.array_equal(H.handler_sparse_gemm_outer(ctx, alpha, maps, X), cpu.sparse_gemm_outer(alpha, maps, X))
    # maps that hit the same index take the sequential path; the reference's accumulation order is kept
    dup = [{5: 1.0, 9: 2.0}, {5: -3.0}, {9: 0.25, 5: 4.0}]
    a2 = rng.standard_normal((3, 2))
    assert np.array_equal(H.handler_sparse_gemm_outer(ctx, a2, dup, X[:2]), cpu.sparse_gemm_outer(a2, dup, X[:2]))
    # copy(dense <- map): zero fill then scatter (initial guess of solve(), reference IterativeSolverTemplate.h:345)
    assert np.array_equal(H.handler_sparse_copy(ctx, X[0], {3: 1.0}), cpu.sparse_copy(X[0], {3: 1.0}))


@pytest.mark.parametrize("explicit_csr", [False, True])
def test_operator_matches_cpu_twin(ctx, oracle, explicit_csr):
    x = np.random.default_rng(15).standard_normal(5003)
    assert np.array_equal(H.harness_banded_apply(ctx, x, 4, 1e-3, explicit_csr), oracle.c.banded_apply(x, 4, 1e-3))


def test_errors_are_raised_not_swallowed(ctx):
    from iterative_solver_b200 import BackendError
    with pytest.raises(BackendError):
        H.handler_select(ctx, np.zeros(3), 5)  # n too large: the handler's error(), as ArrayHandlerIterable.h:96-97
