"""Pins the CPU oracle (TEST INFRASTRUCTURE, oracle/): the plain-C restatement against (a) the golden vectors held by
the reference's own tests, (b) fixtures generated from the reference's own templates (tests/golden/, made by
tests/golden/make_golden.py) and (c) those templates themselves, bit for bit, when oracle/_ref/libitsolv_ref.so is
built. Runs without a GPU."""
import json
import os

import numpy as np
import pytest

from iterative_solver_b200 import _native as N
from iterative_solver_b200 import harness as H

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def inputs(seed=2024, n=4099):
    rng = np.random.default_rng(seed)
    X, Y = rng.standard_normal((4, n)), rng.standard_normal((6, n))
    alpha = rng.standard_normal((4, 6))
    return rng, X, Y, alpha


def test_kat_modified_gram_schmidt(oracle):
    """reference test/itsolv/subspace/test_util.cpp:154-173"""
    V = np.array([[1., 1., 1., 1.], [1., 1. / 5., 1. / 10., 1. / 15.], [1. / 3., 1. / 6., 1. / 9., 1. / 12.],
                  [1. / 2., 1. / 4., 1. / 6., 1. / 8.]])
    ref = np.array([[0.5, 0.5, 0.5, 0.5],
                    [0.858898629520365, -0.184826287365142, -0.3152919019758303, -0.3587804401793933],
                    [-0.1096025454090415, 0.783967708523809, -0.06699956264207202, -0.6073656004726955]])
    out, nulls = oracle.c.modified_gram_schmidt(V, 1e-14)
    assert nulls == [3]
    assert np.abs(out[:3] - ref).max() <= 1e-14


def test_kat_select_max_dot(oracle):
    """reference test/array/testArrayHandlerIterable.cpp:57-69"""
    x = np.array([1., -2., 1., 0., 3., 0., -4., 1.])
    idx, val = oracle.c.select(x, 3, y=np.ones(8))
    assert dict(zip(idx.tolist(), val.tolist())) == {6: 4.0, 4: 3.0, 1: 2.0}


def test_kat_sparse_axpy_dot(oracle):
    """reference test/array/testArrayHandlers.cpp:27-60"""
    m = {1: 1.0, 3: 2.0, 6: 3.0, 11: 4.0}
    y = np.full(20, 0.5)
    out = oracle.c.sparse_gemm_outer(np.array([[2.0]]), [m], y[None, :])[0]
    want = y.copy()
    for i, v in m.items():
        want[i] += 2.0 * v
    assert np.array_equal(out, want)
    assert oracle.c.sparse_gemm_inner(np.full((1, 20), 0.5), [m])[0, 0] == 0.5 + 0.5 * 2. + 0.5 * 3. + 0.5 * 4.


def test_gemm_equals_loops_of_dot_and_axpy(oracle):
    """the property the reference asserts for every handler family (test/array/testGemm.cpp:58-88, 290-326)"""
    _, X, Y, alpha = inputs()
    G = oracle.c.gemm_inner(X, Y)
    for i in range(4):
        for j in range(6):
            assert G[i, j] == oracle.c.dot(X[i].copy(), Y[j].copy())
    out = oracle.c.gemm_outer(alpha, X, Y)
    want = Y.copy()
    for i in range(4):
        for j in range(6):
            want[j] = oracle.c.axpy(alpha[i, j], X[i].copy(), want[j].copy())
    assert np.array_equal(out, want)


def test_c_oracle_against_reference_fixtures(oracle):
    g = np.load(os.path.join(GOLD, "handler_golden.npz"))
    rng, X, Y, alpha = inputs(int(g["seed"]), int(g["n"]))
    n = X.shape[1]
    shift = np.array([0.9, 1.9, 2.9, 3.9])
    diag = np.arange(1, n + 1, dtype=np.float64)
    c = oracle.c
    assert np.array_equal(c.gemm_inner(X, Y), g["gemm_inner"])
    assert np.array_equal(c.gemm_outer(alpha, X, Y), g["gemm_outer"])
    assert np.array_equal(c.axpy(0.37, X[0].copy(), Y[0].copy()), g["axpy"])
    assert np.array_equal(c.scal(-1.7, X[1].copy()), g["scal"])
    assert c.dot(X[0].copy(), Y[0].copy()) == g["dot"][0]
    assert np.array_equal(c.precondition(X, shift, diag), g["precondition"])
    xr = np.round(X[2], 1)
    for name, kw in (("select_min", {}), ("select_max", {"max": True}),
                     ("select_absmax", {"max": True, "ignore_sign": True})):
        i, v = c.select(xr.copy(), 25, **kw)
        assert np.array_equal(i, g[name + "_idx"]) and np.array_equal(v, g[name + "_val"])
    i, v = c.select(xr.copy(), 25, y=np.round(Y[2], 1).copy())
    assert np.array_equal(i, g["select_maxdot_idx"]) and np.array_equal(v, g["select_maxdot_val"])
    V, nulls = c.modified_gram_schmidt(np.vstack([X, X[0] + X[1]]), 1e-10)
    assert np.array_equal(V, g["mgs"]) and nulls == g["mgs_nulls"].tolist()
    maps = [{int(i): float(w) for i, w in zip(rng.choice(n, 3, replace=False), rng.standard_normal(3))} for _ in range(5)]
    assert np.array_equal(c.sparse_gemm_inner(X, maps), g["sparse_gemm_inner"])
    assert np.array_equal(c.sparse_gemm_outer(rng.standard_normal((5, 4)), maps, X), g["sparse_gemm_outer"])
    assert np.array_equal(c.banded_apply(X[3].copy(), 4, 1e-3), g["banded_apply"])
    assert np.array_equal(c.distribution(10, 3), g["distribution_10_3"])
    assert np.array_equal(c.distribution(2_000_000_001, 8), g["distribution_2e9_8"])


def test_c_oracle_against_reference_build(oracle):
    """bit-for-bit against the reference's own ArrayHandlerIterable on fresh random inputs of ragged sizes"""
    if oracle.ref is None:
        pytest.skip("oracle/_ref/libitsolv_ref.so needs /root/reference; covered by the committed fixtures")
    c, r = oracle.c, oracle.ref
    for n in (1, 2, 7, 64, 1001):
        rng = np.random.default_rng(n)
        X, Y = rng.standard_normal((3, n)), rng.standard_normal((5, n))
        alpha = rng.standard_normal((3, 5))
        assert np.array_equal(c.gemm_inner(X, Y), r.gemm_inner(X, Y))
        assert np.array_equal(c.gemm_outer(alpha, X, Y), r.gemm_outer(alpha, X, Y))
        assert np.array_equal(c.precondition(X, [0.1, 0.2, 0.3], np.arange(1., n + 1)),
                              r.precondition(X, [0.1, 0.2, 0.3], np.arange(1., n + 1)))
        xr = np.round(X[0], 1)
        for kw in ({}, {"max": True}, {"max": True, "ignore_sign": True}, {"ignore_sign": True}):
            a, b = c.select(xr.copy(), min(n, 4), **kw), r.select(xr.copy(), min(n, 4), **kw)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.array_equal(c.banded_apply(X[1].copy(), 4, 1e-3), r.banded_apply(X[1].copy(), 4, 1e-3))
    for n, p in ((0, 1), (7, 8), (10, 3), (1000, 7)):
        assert np.array_equal(c.distribution(n, p), r.distribution(n, p))


def test_reference_build_reproduces_its_fixtures(oracle):
    """the solve fixtures are stable outputs of the reference build (and of the LAPACK restatement of its helper)"""
    if oracle.ref is None:
        pytest.skip("oracle/_ref/libitsolv_ref.so needs /root/reference")
    with open(os.path.join(GOLD, "solve_golden.json")) as f:
        golden = json.load(f)
    for name in ("example_davidson_n20_r1", "example_davidson_n20_r2", "example_lineq_n20_r1", "example_diis_n20",
                 "banded_davidson_n30000_r6_qcap8"):
        want = golden[name]
        res, _ = oracle.ref.solve(H.make_spec(**want["spec"]))
        assert res.iterations == want["iterations"]
        assert [res.errors[i] for i in range(res.nroots)] == want["errors"]


def test_config0_example_problem_against_numpy(oracle):
    """BASELINE.json configs[0] (examples/LinearEigensystemExample.cpp): eigenvalues of the ExampleProblem matrix
    from an independent dense solver"""
    with open(os.path.join(GOLD, "solve_golden.json")) as f:
        golden = json.load(f)
    for name, n in (("example_davidson_n20_r2", 20), ("example_davidson_n200_r4_herm", 200)):
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        M = np.where(i == j, i + 1.0, 0.001 * ((i + j) % n))
        ev = np.linalg.eigvalsh(M)
        got = np.array(golden[name]["eigenvalues"])
        assert np.abs(got - ev[:got.size]).max() <= 1e-12
    assert golden["example_davidson_n20_r1"]["iterations"] == 5  # BASELINE.md probe: 5 iterations, 0.999813133574363
    assert abs(golden["example_davidson_n20_r1"]["eigenvalues"][0] - 0.999813133574363) < 1e-14


def _reference_eigen_cases():
    """the matrices of the reference's own end-to-end eigensolver tests (test/itsolv/test_LinearEigensystem.cpp): the
    Hamiltonian files (file_eigen :346-351; phenol is not in the checkout), all-ones matrices with diagonal i*param
    (load_matrix :41-50; n_eigen :353-362, small_eigen :378-386), their non-hermitian variants (nonhermitian_eigen
    :364-376) and the symmetry-blocked ones (symmetry_eigen :388-409)"""
    ham = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hamiltonians.npz"))
    cases = [(name, ham[name]) for name in ("he", "hf", "bh")]

    def ones(n, param=1.0, non_hermiticity=0.0):
        h = np.ones((n, n))
        h[np.diag_indices(n)] = np.arange(n) * param
        if non_hermiticity:
            h[np.tril_indices(n, -1)] *= 1 - non_hermiticity
        return h

    cases.append(("ones100", ones(100)))
    for n in range(1, 5):
        cases.append((f"ones{n}", ones(n)))
    for param in (1.0, 0.1):
        for nh in (0.1, 0.2):
            cases.append((f"ones6_p{param}_nh{nh}", ones(6, param, nh)))
    for n in range(2, 6):
        h = ones(n)
        for i in range(n):
            for j in range(n):
                if (i % 3 == 0) != (j % 3 == 0):
                    h[j, i] = 0
        cases.append((f"sym{n}", h))
    return cases


@pytest.mark.parametrize("helper", ["literal", "product"])
@pytest.mark.parametrize("name,hmat", _reference_eigen_cases(), ids=[c[0] for c in _reference_eigen_cases()])
def test_reference_build_passes_the_references_own_eigensolver_test(oracle, name, hmat, helper):
    """oracle/_ref (the reference's templates + the literal restatement of its Eigen translation unit) driven through the
    protocol and the assertions of the reference's test_eigen (test/itsolv/test_LinearEigensystem.cpp:245-344): errors,
    eigenvalues against a dense diagonalisation, number of R vectors, true residuals, overlap with the dense eigenvectors.
    `product`: the same templates linked with the PRODUCT's host algebra (iterative_solver_b200/host/helper_lapack.cpp) -
    the subspace solver that runs under the CUDA path, under the reference's own test, on the CPU"""
    ref = oracle.ref if helper == "literal" else oracle.ref_product_helper
    if ref is None:
        pytest.skip("oracle/_ref is not built")
    n = hmat.shape[0]
    hermitian = np.abs(hmat - hmat.T).max() < 1e-10
    if hermitian:
        want_ev, want_vec = np.linalg.eigh(hmat)
    else:
        w, v = np.linalg.eig(hmat)
        assert np.abs(w.imag).max() < 1e-12
        order = np.argsort(w.real)
        want_ev, want_vec = w.real[order], v.real[:, order]
    for nroot in range(1, min(n, 28) + 1, max(1, n // 10)):
        for np_ in range(0, min(n, 100) + 1, max(nroot, n // 5)):
            if np_ > 0 and not hermitian:
                break
            if 0 < np_ < nroot:
                continue  # "P space must be empty or at least as large as number of roots sought"
            ev, err, sol, (iterations, r_creations, n_iter) = ref.dense_eigen(hmat, nroot, np_, hermitian)
            where = f"{name}: {nroot} roots, P space {np_}"
            assert (np.abs(err) <= 2e-8).all(), where
            assert np.abs(ev - want_ev[:nroot]).max() <= 2e-9, where
            assert r_creations <= (nroot + 1) * n_iter, where
            for k in range(nroot):
                assert np.linalg.norm(hmat @ sol[k] - ev[k] * sol[k]) <= 1e-8, where
                if hermitian and (k == 0 or want_ev[k] - want_ev[k - 1] > 1e-6) and \
                        (k + 1 == n or want_ev[k + 1] - want_ev[k] > 1e-6):
                    assert abs(abs(sol[k] @ want_vec[:, k]) - 1) <= 1e-8, where


def test_golden_solves_with_the_products_host_algebra(oracle):
    """the golden solves (made with the oracle's literal host algebra) run again on the CPU with the PRODUCT's host
    algebra under the same reference templates: same iteration counts, convergence, creation counters and eigenvalues -
    whatever the fast restatement changes (dgemm products, index sort, symmetric solver, dsyevd) stays below what the
    solver's decisions can see"""
    if oracle.ref_product_helper is None:
        pytest.skip("oracle/_ref is not built")
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solve_golden.json")))
    for name, want in golden.items():
        if want["spec"]["n"] > 100000:
            continue  # seconds each on the CPU; the smaller cases cover every solver and option
        res, _ = oracle.ref_product_helper.solve(H.make_spec(**want["spec"]))
        assert res.iterations == want["iterations"] and res.converged == want["converged"], name
        assert [res.r_creations, res.q_creations, res.p_creations, res.d_creations] == want["creations"], name
        if want.get("eigenvalues"):
            ev = np.array([res.eigenvalues[i] for i in range(len(want["eigenvalues"]))])
            assert np.abs(ev / np.array(want["eigenvalues"]) - 1).max() <= 1e-10, name


@pytest.mark.parametrize("helper", ["literal", "product"])
def test_reference_build_passes_the_references_own_linear_equations_test(oracle, helper):
    """test/itsolv/test_LinearEquations.cpp:59-98 (symmetric_system): matrix(i,j) = i+j+1 (+1 on the diagonal), right-hand
    sides whose solutions are the constant vectors root+1, n = 3..33, up to 13 roots, threshold 1e-10, solutions to 1e-5"""
    ref = oracle.ref if helper == "literal" else oracle.ref_product_helper
    if ref is None:
        pytest.skip("oracle/_ref is not built")
    for n in range(3, 34, 3):
        i = np.arange(n)
        matrix = (i[:, None] + i[None, :] + 1.0) + np.eye(n)
        for nroot in range(1, min(n, 13) + 1):
            expected = np.repeat(np.arange(1.0, nroot + 1)[:, None], n, axis=1)
            rhs = (np.arange(1, nroot + 1)[:, None] * (n * (n + 1) // 2 + i[None, :] * n + 1)).astype(np.float64)
            assert np.array_equal(expected @ matrix, rhs)
            sol, iterations, converged = ref.dense_lineq(matrix, rhs)
            assert np.abs(sol - expected).max() <= 1e-5, (n, nroot, iterations, converged)


@pytest.mark.parametrize("helper", ["literal", "product"])
def test_reference_build_passes_the_references_own_diis_test(oracle, helper):
    """test/itsolv/test_NonLinearEquations.cpp:62-121 (small_quadratic_form): DIIS on a quadratic form, n = 2..50"""
    ref = oracle.ref if helper == "literal" else oracle.ref_product_helper
    if ref is None:
        pytest.skip("oracle/_ref is not built")
    for n in range(2, 51):
        sol, res, (iterations, r_creations, n_iter), error = ref.dense_diis(n, 10.0)
        assert error <= 2e-8, n
        assert r_creations <= 2 * n_iter, n
        assert np.linalg.norm(res) <= 1e-8, n
        assert np.abs(sol - 1.0).max() <= 1e-8, n


def test_kats_of_the_small_subspace_functions(oracle):
    """known answers of test/itsolv/subspace/test_util.cpp for functions of the reference that both the reference's and
    the fused drivers call unmodified: gram_schmidt on an overlap matrix (s_3x3 :124-140, s_4x4_duplicate :142-152),
    eye_order (:84-107), overlap (:26-58), parameter_batches (:175-188)"""
    if oracle.ref is None:
        pytest.skip("oracle/_ref is not built")
    t, norms = oracle.ref.gram_schmidt(np.array([[14, 25, 31], [25, 45, 56], [31, 56, 70]], dtype=float))
    assert np.abs(t - np.array([[1, 0, 0], [-25 / 14, 1, 0], [1, -9 / 5, 1]])).max() <= 1e-14
    assert np.abs(norms - np.sqrt([14, 5 / 14, 1 / 5])).max() <= 1e-13
    t, norms = oracle.ref.gram_schmidt(np.array([[1, 1, 1, 1], [1, 2, 2, 2], [1, 2, 2, 2], [1, 2, 2, 3]], dtype=float))
    assert np.abs(t - np.array([[1, 0, 0, 0], [-1, 1, 0, 0], [0, -1, 1, 0], [0, -1, 0, 1]])).max() <= 1e-14
    assert np.abs(norms - np.array([1, 1, 0, 1])).max() <= 1e-13
    assert oracle.ref.eye_order(np.eye(3)) == [0, 1, 2]
    assert oracle.ref.eye_order(np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=float)) == [2, 0, 1]
    assert oracle.ref.eye_order(np.array([[0.1, 0.5, 0.2], [0.2, 0.1, 0.5], [0.5, 0.2, 0.1]])) == [2, 0, 1]
    alphas = np.array([1.0, 2.0, 3.0])
    a, b = oracle.ref.overlap(np.repeat(alphas[:, None], 5, axis=1))
    want = 5 * alphas[:, None] * alphas[None, :]
    assert np.array_equal(a, want) and np.array_equal(b, want)
    assert oracle.ref.parameter_batches(3, 3) == [(0, 3)]
    assert oracle.ref.parameter_batches(2, 3) == [(0, 2)]
    assert oracle.ref.parameter_batches(9, 3) == [(0, 3), (3, 6), (6, 9)]
    assert oracle.ref.parameter_batches(4, 3) == [(0, 3), (3, 4)]
    # D-space resetter: which Q vector overlaps most with each R vector (test/itsolv/testDSpaceResetter.cpp:44-76)
    eye = np.eye(3)
    assert oracle.ref.max_overlap_with_R(eye, eye) == [2, 1, 0]
    assert oracle.ref.max_overlap_with_R(eye, eye[[0, 2]]) == [1, 0]
    assert oracle.ref.max_overlap_with_R(eye, eye[[1]]) == [0]
    assert oracle.ref.max_overlap_with_R(eye, np.zeros((0, 3))) == []

