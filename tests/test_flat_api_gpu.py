"""The flat C interface of the solvers over DEVICE buffers (include/itsolv_b200_solver.h, the counterpart of the
reference's src/molpro/linalg/IterativeSolverC.h): a caller that owns its vectors in GPU memory drives the solvers step
by step - AddVector / PreconditionDefault / EndIteration / Solution - exactly as the reference's solve() does
(itsolv/IterativeSolverTemplate.h:322-408), and must arrive at the reference's golden results."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from iterative_solver_b200 import _native as N

pytestmark = pytest.mark.gpu
HUGE = float(np.finfo(np.float64).max)
B, EPS = 4, 1e-3

with open(os.path.join(os.path.dirname(__file__), "golden", "solve_golden.json")) as f:
    GOLDEN = json.load(f)


def check(lib, rc):
    assert rc == 0, lib.ItsolvB200LastError().decode()


def apply_operator(ctx, x, y, n):
    for k in range(x.shape[0]):
        ctx.banded_apply(x[k], y[k], n, 0, B, EPS)


def drive(lib, ctx, params, action, nwork, n, precondition=True):
    """the loop of the reference's solve() over the flat interface"""
    nbuf = params.shape[0]
    for _ in range(100):
        if nwork <= 0:
            break
        apply_operator(ctx, params[:nwork], action[:nwork], n)
        nwork = lib.ItsolvB200AddVector(nbuf, params.data_ptr(), action.data_ptr())
        assert nwork >= 0, lib.ItsolvB200LastError().decode()
        while lib.ItsolvB200EndIterationNeeded() == 1:
            if nwork > 0 and precondition:
                check(lib, lib.ItsolvB200PreconditionDefault(nwork, action.data_ptr()))
            nwork = lib.ItsolvB200EndIteration(nbuf, params.data_ptr(), action.data_ptr())
            assert nwork >= 0, lib.ItsolvB200LastError().decode()
    return nwork


@pytest.mark.parametrize("name,options", [("banded_davidson_n100000_r4", b""), ("banded_davidson_n100000_r4", b"fused=0"),
                                          ("banded_davidson_n30000_r6_qcap8", b"max_size_qspace=8")])
def test_davidson_through_the_flat_interface(ctx, name, options):
    want = GOLDEN[name]
    n, nroots = want["spec"]["n"], want["spec"]["nroots"]
    lib = N.host()
    lo, hi = C.c_size_t(), C.c_size_t()
    check(lib, lib.ItsolvB200LinearEigensystemInitialize(ctx.handle, n, nroots, C.byref(lo), C.byref(hi), 1e-8, HUGE, 1, 0,
                                                        options))
    try:
        assert (lo.value, hi.value) == (0, n)
        assert lib.ItsolvB200HasEigenvalues() == 1 and lib.ItsolvB200NonLinear() == 0
        params = torch.zeros((nroots, n), dtype=torch.float64, device="cuda")
        action = torch.zeros_like(params)
        diag = torch.arange(1, n + 1, dtype=torch.float64, device="cuda")
        check(lib, lib.ItsolvB200SetDiagonals(diag.data_ptr()))
        back = torch.zeros_like(diag)
        check(lib, lib.ItsolvB200Diagonals(back.data_ptr()))
        assert torch.equal(back, diag)
        for k in range(nroots):  # the default initial guess: unit vectors on the smallest diagonal elements
            params[k, k] = 1.0
        nwork = drive(lib, ctx, params, action, nroots, n)
        assert nwork == 0
        assert lib.ItsolvB200Iterations() == want["iterations"]
        ev, err = np.zeros(nroots), np.zeros(nroots)
        check(lib, lib.ItsolvB200Eigenvalues(ev.ctypes.data_as(N.c_double_p)))
        check(lib, lib.ItsolvB200Errors(err.ctypes.data_as(N.c_double_p)))
        assert np.abs(ev / np.array(want["eigenvalues"]) - 1).max() <= 1e-10
        assert err.max() <= 1e-8
        roots = (C.c_int * nroots)(*range(nroots))
        check(lib, lib.ItsolvB200Solution(nroots, roots, params.data_ptr(), action.data_ptr()))
        sol = params.cpu().numpy()
        for s, chk in zip(sol, want["solution_checksums"]):
            assert abs(np.sum(s) - chk) <= 1e-7 * max(1.0, np.abs(s).sum())
        assert float(action.norm(dim=1).max()) <= 1e-8  # the residuals A x - lambda x of the solutions
        y = torch.zeros_like(params)
        apply_operator(ctx, params, y, n)
        lam = torch.from_numpy(ev).cuda()[:, None]
        assert float((y - lam * params).norm(dim=1).max()) <= 1e-7
        # the query functions of IterativeSolverC.h:48-62: these solvers propose no P space (IterativeSolverTemplate.h:238)
        # and carry no function value
        indices = (C.c_size_t * 8)()
        assert lib.ItsolvB200SuggestP(params.data_ptr(), action.data_ptr(), 8, 1e-3, indices) == 0
        assert lib.ItsolvB200HasValues() == 0
        assert isinstance(lib.ItsolvB200Value(), float)
        check(lib, lib.ItsolvB200PrintStatistics())
    finally:
        check(lib, lib.ItsolvB200Finalize())


@pytest.mark.parametrize("options", [b"", b"fused=1"])
def test_linear_equations_through_the_flat_interface(ctx, options):
    want = GOLDEN["banded_lineq_n20000_r8"]
    n, nroots = want["spec"]["n"], want["spec"]["nroots"]
    lib = N.host()
    kernels = N.kernels()
    scratch = torch.zeros(n, dtype=torch.float64, device="cuda")
    rhs = torch.zeros((nroots, n), dtype=torch.float64, device="cuda")
    for k in range(nroots):  # the harness's right-hand sides: A applied to the known solutions (ITSOLV_RHS_SCALED)
        assert kernels.itsolv_banded_fill_f64(ctx.handle, 2, k, 0, n, scratch.data_ptr()) == 0
        ctx.banded_apply(scratch, rhs[k], n, 0, B, EPS)
    lo, hi = C.c_size_t(), C.c_size_t()
    check(lib, lib.ItsolvB200LinearEquationsInitialize(ctx.handle, n, nroots, C.byref(lo), C.byref(hi), rhs.data_ptr(), 0.0,
                                                      1e-8, HUGE, 1, 0, options))
    try:
        # initial guess of the golden run: unit vectors on the smallest diagonal elements (what solve() generates with
        # generate_initial_guess = true, reference IterativeSolverTemplate.h:337-350)
        params = torch.zeros_like(rhs)
        for k in range(nroots):
            params[k, k] = 1.0
        action = torch.zeros_like(params)
        diag = torch.arange(1, n + 1, dtype=torch.float64, device="cuda")
        check(lib, lib.ItsolvB200SetDiagonals(diag.data_ptr()))
        nwork = drive(lib, ctx, params, action, nroots, n)
        assert nwork == 0
        assert lib.ItsolvB200Iterations() == want["iterations"]
        roots = (C.c_int * nroots)(*range(nroots))
        check(lib, lib.ItsolvB200Solution(nroots, roots, params.data_ptr(), action.data_ptr()))
        sol = params.cpu().numpy()
        for s, chk in zip(sol, want["solution_checksums"]):
            assert abs(np.sum(s) - chk) <= 1e-6 * max(1.0, abs(chk))
        y = torch.zeros_like(params)
        apply_operator(ctx, params, y, n)
        assert float(((y - rhs).norm(dim=1) / rhs.norm(dim=1)).max()) <= 1e-7
    finally:
        check(lib, lib.ItsolvB200Finalize())


def test_instances_nest_and_errors_are_reported(ctx):
    lib = N.host()
    lo, hi = C.c_size_t(), C.c_size_t()
    assert lib.ItsolvB200LinearEigensystemInitialize(ctx.handle, 100, 2, C.byref(lo), C.byref(hi), 1e-8, HUGE, 1, 0,
                                                    b"no_such_option=1") != 0
    assert b"unknown option" in lib.ItsolvB200LastError()
    check(lib, lib.ItsolvB200LinearEigensystemInitialize(ctx.handle, 100, 2, C.byref(lo), C.byref(hi), 1e-8, HUGE, 1, 0, b""))
    check(lib, lib.ItsolvB200NonLinearEquationsInitialize(ctx.handle, 50, C.byref(lo), C.byref(hi), 1e-8, 0, b""))
    assert lib.ItsolvB200NonLinear() == 1 and lib.ItsolvB200HasEigenvalues() == 0  # the top instance is the active one
    lib.ItsolvB200SetMaxIter(17)
    assert lib.ItsolvB200MaxIter() == 17
    check(lib, lib.ItsolvB200Finalize())
    assert lib.ItsolvB200NonLinear() == 0 and lib.ItsolvB200HasEigenvalues() == 1
    assert lib.ItsolvB200AddVector(1, None, None) == -1  # null buffers
    check(lib, lib.ItsolvB200Finalize())


def _dense_cases():
    ham = np.load(os.path.join(os.path.dirname(__file__), "golden", "hamiltonians.npz"))
    cases = [(name, ham[name]) for name in ("he", "hf", "bh")]
    ones = np.ones((100, 100))
    ones[np.diag_indices(100)] = np.arange(100.0)
    cases.append(("ones100", ones))
    nh = np.ones((6, 6))
    nh[np.diag_indices(6)] = np.arange(6.0)
    nh[np.tril_indices(6, -1)] *= 0.9
    cases.append(("ones6_nh0.1", nh))
    return cases


@pytest.mark.parametrize("name,hmat", _dense_cases(), ids=[c[0] for c in _dense_cases()])
def test_the_references_own_eigensolver_test_on_the_gpu(ctx, name, hmat):
    """The matrices and assertions of the reference's end-to-end eigensolver test (test/itsolv/test_LinearEigensystem.cpp:
    245-344; its Hamiltonian files, tests/golden/hamiltonians.npz) with the CUDA containers and the fused driver underneath,
    driven step by step through the flat interface as the reference's test drives its solver: the operator is the caller's
    (a dense product on the device), unit-vector guess on the lowest diagonal elements, Davidson update, until the working
    set is empty. Errors, eigenvalues against a dense diagonalisation, true residuals, overlap with the dense eigenvectors."""
    lib = N.host()
    n = hmat.shape[0]
    hermitian = bool(np.abs(hmat - hmat.T).max() < 1e-10)
    if hermitian:
        want_ev, want_vec = np.linalg.eigh(hmat)
    else:
        w, v = np.linalg.eig(hmat)
        order = np.argsort(w.real)
        want_ev, want_vec = w.real[order], v.real[:, order]
    Hd = torch.from_numpy(hmat).cuda()
    diag = torch.from_numpy(np.ascontiguousarray(np.diag(hmat))).cuda()
    for nroot in range(1, min(n, 28) + 1, max(1, n // 10)):
        lo, hi = C.c_size_t(), C.c_size_t()
        options = f"max_size_qspace={max(6 * nroot, min(n, 6 * nroot))},reset_D=8".encode()
        check(lib, lib.ItsolvB200LinearEigensystemInitialize(ctx.handle, n, nroot, C.byref(lo), C.byref(hi), 1e-8, HUGE,
                                                            int(hermitian), 0, options))
        try:
            params = torch.zeros((nroot, n), dtype=torch.float64, device="cuda")
            action = torch.zeros_like(params)
            check(lib, lib.ItsolvB200SetDiagonals(diag.data_ptr()))
            for root, at in enumerate(np.argsort(np.diag(hmat), kind="stable")[:nroot]):
                params[root, at] = 1.0
            nwork = nroot
            for _ in range(100):
                if nwork <= 0:
                    break
                action[:nwork] = params[:nwork] @ Hd.T
                nwork = lib.ItsolvB200AddVector(nroot, params.data_ptr(), action.data_ptr())
                assert nwork >= 0, lib.ItsolvB200LastError().decode()
                while lib.ItsolvB200EndIterationNeeded() == 1:
                    if nwork > 0:
                        check(lib, lib.ItsolvB200PreconditionDefault(nwork, action.data_ptr()))
                    nwork = lib.ItsolvB200EndIteration(nroot, params.data_ptr(), action.data_ptr())
                    assert nwork >= 0, lib.ItsolvB200LastError().decode()
            where = f"{name}: {nroot} roots"
            assert nwork == 0, where
            ev, err = np.zeros(nroot), np.zeros(nroot)
            check(lib, lib.ItsolvB200Eigenvalues(ev.ctypes.data_as(N.c_double_p)))
            check(lib, lib.ItsolvB200Errors(err.ctypes.data_as(N.c_double_p)))
            assert (np.abs(err) <= 2e-8).all(), where
            assert np.abs(ev - want_ev[:nroot]).max() <= 2e-9, where
            roots = (C.c_int * nroot)(*range(nroot))
            check(lib, lib.ItsolvB200Solution(nroot, roots, params.data_ptr(), action.data_ptr()))
            sol = params.cpu().numpy()
            for k in range(nroot):
                assert np.linalg.norm(hmat @ sol[k] - ev[k] * sol[k]) <= 1e-8, where
                if hermitian and (k == 0 or want_ev[k] - want_ev[k - 1] > 1e-6) and \
                        (k + 1 == n or want_ev[k + 1] - want_ev[k] > 1e-6):
                    assert abs(abs(sol[k] @ want_vec[:, k]) - 1) <= 1e-8, where
        finally:
            check(lib, lib.ItsolvB200Finalize())


@pytest.mark.parametrize("options", [b"", b"fused=1"])
def test_the_references_own_linear_equations_test_on_the_gpu(ctx, options):
    """test/itsolv/test_LinearEquations.cpp:59-98 (symmetric_system) with the CUDA containers underneath, through the flat
    interface: matrix(i,j) = i+j+1 (+1 on the diagonal), right-hand sides whose solutions are the constant vectors
    root+1, n = 3..33, up to 13 right-hand sides, threshold 1e-10, solutions to 1e-5"""
    lib = N.host()
    for n in range(3, 34, 6):
        i = np.arange(n)
        matrix = (i[:, None] + i[None, :] + 1.0) + np.eye(n)
        Md = torch.from_numpy(matrix).cuda()
        diag = torch.from_numpy(np.ascontiguousarray(np.diag(matrix))).cuda()
        for nroot in sorted({1, 2, min(n, 8), min(n, 13)}):
            expected = np.repeat(np.arange(1.0, nroot + 1)[:, None], n, axis=1)
            rhs = torch.from_numpy(expected @ matrix).cuda().contiguous()
            lo, hi = C.c_size_t(), C.c_size_t()
            check(lib, lib.ItsolvB200LinearEquationsInitialize(ctx.handle, n, nroot, C.byref(lo), C.byref(hi), rhs.data_ptr(),
                                                              0.0, 1e-10, HUGE, 1, 0, options))
            try:
                params = torch.zeros((nroot, n), dtype=torch.float64, device="cuda")
                for k in range(nroot):  # what solve() generates with generate_initial_guess = true
                    params[k, k] = 1.0
                action = torch.zeros_like(params)
                check(lib, lib.ItsolvB200SetDiagonals(diag.data_ptr()))
                nwork = nroot
                for _ in range(200):
                    if nwork <= 0:
                        break
                    action[:nwork] = params[:nwork] @ Md.T
                    nwork = lib.ItsolvB200AddVector(nroot, params.data_ptr(), action.data_ptr())
                    assert nwork >= 0, lib.ItsolvB200LastError().decode()
                    while lib.ItsolvB200EndIterationNeeded() == 1:
                        if nwork > 0:
                            check(lib, lib.ItsolvB200PreconditionDefault(nwork, action.data_ptr()))
                        nwork = lib.ItsolvB200EndIteration(nroot, params.data_ptr(), action.data_ptr())
                        assert nwork >= 0, lib.ItsolvB200LastError().decode()
                assert nwork == 0, (n, nroot)
                roots = (C.c_int * nroot)(*range(nroot))
                check(lib, lib.ItsolvB200Solution(nroot, roots, params.data_ptr(), action.data_ptr()))
                assert np.abs(params.cpu().numpy() - expected).max() <= 1e-5, (n, nroot)
            finally:
                check(lib, lib.ItsolvB200Finalize())


@pytest.mark.parametrize("options", [b"max_size_qspace=6", b"max_size_qspace=6,fused=1"])
def test_the_references_own_diis_test_on_the_gpu(ctx, options):
    """test/itsolv/test_NonLinearEquations.cpp:62-121 (small_quadratic_form) with the CUDA containers underneath, through
    the flat interface: f = (x-1).h.(x-1)/2 with h all ones and diagonal 10 (i+2), start at e_0, update by the diagonal"""
    lib = N.host()
    for n in range(2, 51, 4):
        h = np.ones((n, n))
        h[np.diag_indices(n)] = 10.0 * (np.arange(n) + 2)
        Hd = torch.from_numpy(h).cuda()
        diag = torch.from_numpy(np.ascontiguousarray(np.diag(h))).cuda()
        lo, hi = C.c_size_t(), C.c_size_t()
        check(lib, lib.ItsolvB200NonLinearEquationsInitialize(ctx.handle, n, C.byref(lo), C.byref(hi), 1e-8, 0, options))
        try:
            x = torch.zeros((1, n), dtype=torch.float64, device="cuda")
            g = torch.zeros_like(x)
            x[0, 0] = 1.0
            nwork = 1
            for _ in range(1000):
                if nwork <= 0:
                    break
                g[0] = Hd @ (x[0] - 1.0)
                added = lib.ItsolvB200AddVector(1, x.data_ptr(), g.data_ptr())
                assert added >= 0, lib.ItsolvB200LastError().decode()
                if added:
                    g[0] = g[0] / diag
                nwork = lib.ItsolvB200EndIteration(1, x.data_ptr(), g.data_ptr())
                assert nwork >= 0, lib.ItsolvB200LastError().decode()
            assert nwork == 0, n
            err = np.zeros(1)
            check(lib, lib.ItsolvB200Errors(err.ctypes.data_as(N.c_double_p)))
            assert err[0] <= 2e-8, n
            root = (C.c_int * 1)(0)
            check(lib, lib.ItsolvB200Solution(1, root, x.data_ptr(), g.data_ptr()))
            assert float(g.norm()) <= 1e-8, n
            assert float((x - 1.0).abs().max()) <= 1e-8, n
        finally:
            check(lib, lib.ItsolvB200Finalize())
