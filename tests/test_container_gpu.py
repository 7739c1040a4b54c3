"""The reference's conformance tests of a distributed array (test/array/testDistrArray.h:496-678, the typed suite
DistrArrayCollectiveLinAlgF, plus DistrArrayRangeLinAlgF :400-457 and TestDistrArray.select_max_dot :110-131) applied to
DistrArrayCUDA through the C ABI: same operations, same fixture shapes (dim = 30, alpha = 1, beta = 2, a sparse array of
every fifth element), results compared with DoubleEq-style exactness against numpy arithmetic done operation by
operation."""
import numpy as np
import pytest

from iterative_solver_b200 import harness as H

pytestmark = pytest.mark.gpu

DIM, ALPHA, BETA, EVERY = 30, 1.0, 2.0, 5
RANGE_ALPHA = np.arange(DIM, dtype=np.float64)
RANGE_BETA = RANGE_ALPHA * BETA
SPARSE = {int(i): float(RANGE_BETA[i]) for i in range(0, DIM, EVERY)}


def full(v, n=DIM):
    return np.full(n, v)


@pytest.mark.parametrize("n", [DIM, 4097])
def test_add_sub_axpy(ctx, n):
    c, _, _ = H.handler_distr_array(ctx, 0, full(ALPHA, n), b=full(BETA, n))  # add
    assert np.array_equal(c, full(ALPHA + BETA, n))
    c, _, _ = H.handler_distr_array(ctx, 1, full(ALPHA, n), b=full(BETA, n))  # sub
    assert np.array_equal(c, full(ALPHA - BETA, n))
    assert np.array_equal(H.handler_axpy(ctx, -3.0, full(BETA, n), full(ALPHA, n)), full(ALPHA - 3.0 * BETA, n))


def test_scalar_members_and_recip(ctx):
    vals = RANGE_ALPHA + 1.0
    c, _, _ = H.handler_distr_array(ctx, 2, vals, scalar=0.37)  # add(double)
    assert np.array_equal(c, vals + 0.37)
    c, _, _ = H.handler_distr_array(ctx, 3, vals, scalar=0.37)  # sub(double)
    assert np.array_equal(c, vals + (-0.37))
    c, _, _ = H.handler_distr_array(ctx, 4, vals)  # recip
    assert np.array_equal(c, 1.0 / vals)
    c, _, _ = H.handler_distr_array(ctx, 11, vals)  # zero
    assert np.array_equal(c, np.zeros(DIM))
    assert np.array_equal(H.handler_scal(ctx, 2.5, vals), vals * 2.5)


def test_axpy_map_and_dot_map(ctx):
    c, _, _ = H.handler_distr_array(ctx, 8, full(ALPHA), scalar=5.0, sparse=SPARSE)
    want = full(ALPHA)
    for i, v in SPARSE.items():
        want[i] += 5.0 * v
    assert np.array_equal(c, want)
    _, d, _ = H.handler_distr_array(ctx, 9, RANGE_ALPHA, sparse=SPARSE)
    ref = 0.0
    for i in sorted(SPARSE):
        ref += RANGE_ALPHA[i] * SPARSE[i]
    assert d == ref
    assert H.handler_dot(ctx, RANGE_ALPHA.copy(), RANGE_BETA.copy()) == float(np.dot(RANGE_ALPHA, RANGE_BETA))


@pytest.mark.parametrize("n", [DIM, 50001])
def test_times(ctx, n):
    rng = np.random.default_rng(n)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    c, _, _ = H.handler_distr_array(ctx, 6, np.zeros(n), a=a, b=b)  # c = a * b
    assert np.array_equal(c, a * b)
    c, _, _ = H.handler_distr_array(ctx, 5, a, a=b)  # c *= a
    assert np.array_equal(c, a * b)
    c, _, _ = H.handler_distr_array(ctx, 6, full(7.0), a=full(ALPHA), b=full(BETA))
    assert np.array_equal(c, full(ALPHA * BETA))


@pytest.mark.parametrize("append,negative", [(True, True), (True, False), (False, False), (False, True)])
def test_divide(ctx, append, negative):
    """c[i] (=|+=|-=) (-)a[i] / (b[i] + shift), reference array/DistrArray.cpp:140-167"""
    shift = 0.5
    c0 = full(ALPHA)
    c, _, _ = H.handler_distr_array(ctx, 7, c0, a=full(ALPHA), b=full(BETA), scalar=shift,
                                    flags=(1 if append else 0) | (2 if negative else 0))
    q = ALPHA / (BETA + shift)
    want = (ALPHA - q if negative else ALPHA + q) if append else (-q if negative else q)
    assert np.array_equal(c, full(want))
    rng = np.random.default_rng(3)
    a, b, c0 = rng.standard_normal(4099), rng.standard_normal(4099) + 3.0, rng.standard_normal(4099)
    c, _, _ = H.handler_distr_array(ctx, 7, c0, a=a, b=b, scalar=shift, flags=(1 if append else 0) | (2 if negative else 0))
    q = (-a if (negative and not append) else a) / (b + shift)
    want = (c0 - q if negative else c0 + q) if append else q
    assert np.array_equal(c, want)


def test_select_max_dot_with_a_sparse_array(ctx):
    """ArrayHandlerCUDASparse::select_max_dot forwards to the container (as reference ArrayHandlerDistrSparse.h:65-67): the n
    entries of the sparse array with the largest |x[i] * y[i]| (array/util/select_max_dot.h:60-83)"""
    x = RANGE_ALPHA - 12.0
    y = {3: 2.0, 7: -1.0, 11: 4.0, 12: 100.0, 20: 0.5, 29: -0.25}
    _, _, sel = H.handler_distr_array(ctx, 10, x, flags=3, sparse=y)
    prod = sorted(((abs(x[i] * v), i) for i, v in y.items()), reverse=True)[:3]
    assert sel == {i: p for p, i in prod}
    # dense x dense selection of the reference's test (TestDistrArray.select_max_dot): x = iota, y = 1, n = 5
    idx, val = H.handler_select(ctx, RANGE_ALPHA.copy(), 5, y=np.ones(DIM))
    assert list(idx) == [DIM - 5 + i for i in range(5)] and list(val) == [float(DIM - 5 + i) for i in range(5)]
