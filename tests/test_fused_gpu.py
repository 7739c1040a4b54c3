"""The fused driver path (iterative_solver_b200/host/FusedDavidson.h) against the same
golden results of the reference as the unfused path: identical iteration counts, convergence, creation counters,
eigenvalues to 1e-10 and solution vectors; and its batched kernels bit for bit against the oracle's sequences of
single-vector operations."""
import json
import os

import numpy as np
import pytest
import torch

from iterative_solver_b200 import _native as N
from iterative_solver_b200 import harness as H

pytestmark = pytest.mark.gpu

with open(os.path.join(os.path.dirname(__file__), "golden", "solve_golden.json")) as f:
    GOLDEN = json.load(f)
DAVIDSON = sorted(k for k, v in GOLDEN.items() if v["spec"]["kind"] == N.KIND_DAVIDSON)


def dev_rows(a):
    return [torch.from_numpy(np.ascontiguousarray(a[i])).cuda() for i in range(a.shape[0])]


def host(ts):
    return np.stack([t.cpu().numpy() for t in ts])


@pytest.mark.parametrize("n", [1, 2, 33, 1000, 50001])
def test_batched_kernels_bit_exact(ctx, oracle, n):
    rng = np.random.default_rng(n)
    X, Y = rng.standard_normal((5, n)), rng.standard_normal((5, n))
    alpha = rng.standard_normal(5)
    xs, ys = dev_rows(X), dev_rows(Y)
    ctx.axpy_batch(alpha, xs, ys)
    want = np.stack([oracle.c.axpy(alpha[k], X[k].copy(), Y[k].copy()) for k in range(5)])
    assert np.array_equal(host(ys), want)
    ctx.scal_batch(alpha, xs)
    assert np.array_equal(host(xs), np.stack([oracle.c.scal(alpha[k], X[k].copy()) for k in range(5)]))
    ctx.fill_batch(alpha, xs)
    assert np.array_equal(host(xs), np.repeat(alpha[:, None], n, axis=1))
    # one R-R Gram-Schmidt step == scal of the pivot followed by axpys onto the later vectors
    R = rng.standard_normal((4, n))
    ov = rng.standard_normal(3)
    rs = dev_rows(R)
    ctx.mgs_step(0.731, rs[0], ov, rs[1:])
    pivot = oracle.c.scal(0.731, R[0].copy())
    want = np.stack([pivot] + [oracle.c.axpy(-ov[j], pivot.copy(), R[j + 1].copy()) for j in range(3)])
    assert np.array_equal(host(rs), want)


@pytest.mark.parametrize("n,k,m", [(1, 1, 1), (2, 3, 2), (33, 5, 3), (1000, 4, 4), (4097, 9, 5), (50001, 12, 8),
                                   (20000, 7, 16), (3001, 6, 19),
                                   # the TMA tile kernel of 9..16 roots: odd tail, a second root group of one root, fewer
                                   # vectors than a stage holds, several tiles per CTA, a chunked k loop
                                   (70001, 24, 16), (5000, 13, 9), (8193, 3, 12), (227335, 5, 16), (4096, 40, 16)])
@pytest.mark.parametrize("precondition", [False, True])
def test_davidson_residual_kernel_equals_the_unfused_sequence(ctx, oracle, n, k, m, precondition):
    """itsolv_davidson_residual_f64 against the oracle's sequence of the reference's steps: two expansions (FMA chains
    from zero, as gemm_outer with beta_zero), r -= lambda x as an axpy, the norms, precondition_default. Vectors bit for
    bit; norms to 1e-12 relative (summation order)."""
    rng = np.random.default_rng(1000 * n + 10 * k + m)
    Q, A = rng.standard_normal((k, n)), rng.standard_normal((k, n))
    coef = rng.standard_normal((k, m))
    lam = np.arange(1, m + 1) + 0.25 * rng.standard_normal(m)
    diag = np.arange(1, n + 1, dtype=np.float64)
    want_x = oracle.c.gemm_outer(coef, Q, np.zeros((m, n)), fma=True)
    want_r = oracle.c.gemm_outer(coef, A, np.zeros((m, n)), fma=True)
    want_r = np.stack([oracle.c.axpy(-lam[j], want_x[j], want_r[j]) for j in range(m)])
    want_n2 = np.array([oracle.c.dot(want_r[j], want_r[j]) for j in range(m)])
    want_out = oracle.c.precondition(want_r, lam, diag) if precondition else want_r
    want_n2w = np.array([oracle.c.dot(want_out[j], want_out[j]) for j in range(m)])
    q, a = dev_rows(Q), dev_rows(A)
    for with_x in (False, True):
        out_r = [torch.full((n,), np.nan, dtype=torch.float64, device="cuda") for _ in range(m)]
        out_x = [torch.full((n,), np.nan, dtype=torch.float64, device="cuda") for _ in range(m)] if with_x else None
        n2, n2w = ctx.davidson_residual(coef, q, a, lam, out_r, diag=dev_rows(diag[None])[0] if precondition else None,
                                        out_x=out_x)
        assert np.array_equal(host(out_r), want_out)
        if with_x:
            assert np.array_equal(host(out_x), want_x)
        assert np.abs(n2 - want_n2).max() <= 1e-12 * want_n2.max()
        assert np.abs(n2w - want_n2w).max() <= 1e-12 * want_n2w.max()
    assert np.array_equal(host(q), Q) and np.array_equal(host(a), A), "inputs must not be touched"


@pytest.mark.parametrize("n,m", [(1, 0), (2, 1), (33, 2), (1000, 3), (4097, 4), (50001, 7), (20001, 16)])
def test_mgs_step_dots_equals_step_then_dots(ctx, oracle, n, m):
    """itsolv_mgs_step_dots_f64: vectors bit for bit as scal + axpys; the returned products to 1e-12 of |x||y|"""
    rng = np.random.default_rng(n + m)
    R = rng.standard_normal((m + 1, n))
    ov = rng.standard_normal(m)
    rs = dev_rows(R)
    dots = ctx.mgs_step_dots(0.731, rs[0], ov, rs[1:])
    pivot = oracle.c.scal(0.731, R[0].copy())
    want = np.stack([pivot] + [oracle.c.axpy(-ov[j], pivot.copy(), R[j + 1].copy()) for j in range(m)])
    assert np.array_equal(host(rs), want)
    norms = np.linalg.norm(want, axis=1)
    assert abs(dots[0] - oracle.c.dot(want[0], want[0])) <= 1e-12 * norms[0] ** 2
    for t in range(m):
        assert abs(dots[1 + t] - oracle.c.dot(want[1], want[1 + t])) <= 1e-12 * norms[1] * norms[1 + t]


@pytest.mark.parametrize("n,k,m", [(1, 1, 1), (33, 3, 2), (1000, 5, 4), (50001, 12, 7), (4097, 20, 16), (3001, 4, 19)])
def test_gemm_outer_scaled_equals_scal_then_gemm_outer(ctx, oracle, n, k, m):
    rng = np.random.default_rng(7 * n + k + m)
    X, Y = rng.standard_normal((k, n)), rng.standard_normal((m, n))
    alpha, scale = rng.standard_normal((k, m)), rng.standard_normal(m)
    xs, ys = dev_rows(X), dev_rows(Y)
    ctx.gemm_outer_scaled(alpha, xs, ys, scale)
    scaled = np.stack([oracle.c.scal(scale[j], Y[j].copy()) for j in range(m)])
    assert np.array_equal(host(ys), oracle.c.gemm_outer(alpha, X, scaled, fma=True))


@pytest.mark.parametrize("n,k,m", [(1, 1, 1), (33, 5, 3), (4097, 9, 5), (50001, 12, 8), (20001, 7, 16), (70002, 24, 13)])
@pytest.mark.parametrize("precondition", [False, True])
def test_subspace_residual_linear_equations_form(ctx, oracle, n, k, m, precondition):
    """mode 1 of itsolv_subspace_residual_f64 against the reference's sequence for LinearEquations: two expansions from
    zero, axpy(-1, rhs, r), scal(1/|rhs|, r) (LinearEquationsDavidson.h:173-184), the norms, precondition_default with a
    zero shift. Vectors bit for bit."""
    rng = np.random.default_rng(77 * n + 10 * k + m)
    Q, A, B = rng.standard_normal((k, n)), rng.standard_normal((k, n)), rng.standard_normal((m, n))
    coef = rng.standard_normal((k, m))
    scale = 1.0 / np.sqrt(np.array([oracle.c.dot(B[j], B[j]) for j in range(m)]))
    diag = np.arange(1, n + 1, dtype=np.float64)
    want_x = oracle.c.gemm_outer(coef, Q, np.zeros((m, n)), fma=True)
    want_r = oracle.c.gemm_outer(coef, A, np.zeros((m, n)), fma=True)
    want_r = np.stack([oracle.c.scal(scale[j], oracle.c.axpy(-1.0, B[j], want_r[j])) for j in range(m)])
    want_n2 = np.array([oracle.c.dot(want_r[j], want_r[j]) for j in range(m)])
    want_out = oracle.c.precondition(want_r, np.zeros(m), diag) if precondition else want_r
    want_n2w = np.array([oracle.c.dot(want_out[j], want_out[j]) for j in range(m)])
    out_r = [torch.full((n,), np.nan, dtype=torch.float64, device="cuda") for _ in range(m)]
    out_x = [torch.full((n,), np.nan, dtype=torch.float64, device="cuda") for _ in range(m)]
    n2, n2w = ctx.subspace_residual(coef, dev_rows(Q), dev_rows(A), out_r, rhs=dev_rows(B), rscale=scale,
                                    diag=dev_rows(diag[None])[0] if precondition else None, shift=np.zeros(m), out_x=out_x)
    assert np.array_equal(host(out_r), want_out) and np.array_equal(host(out_x), want_x)
    assert np.abs(n2 - want_n2).max() <= 1e-12 * want_n2.max()
    assert np.abs(n2w - want_n2w).max() <= 1e-12 * want_n2w.max()


@pytest.mark.parametrize("n,k,m", [(1, 1, 1), (33, 5, 1), (4097, 3, 1), (50001, 8, 1), (20000, 6, 2), (70001, 24, 16)])
@pytest.mark.parametrize("step", [False, True])
def test_subspace_residual_diis_form(ctx, oracle, n, k, m, step):
    """modes 2 and 3 of itsolv_subspace_residual_f64 against the reference's DIIS sequence (IterativeSolverTemplate.h:
    191-215, NonLinearEquationsDIIS.h:103-119): two expansions from zero, <r, r>, precondition_default with a zero shift,
    and with the step axpy(-1, preconditioned residual, parameters). Vectors bit for bit."""
    rng = np.random.default_rng(13 * n + 7 * k + m)
    Q, A = rng.standard_normal((k, n)), rng.standard_normal((k, n))
    coef = rng.standard_normal((k, m))
    diag = np.arange(1, n + 1, dtype=np.float64)
    want_x = oracle.c.gemm_outer(coef, Q, np.zeros((m, n)), fma=True)
    want_r = oracle.c.gemm_outer(coef, A, np.zeros((m, n)), fma=True)
    want_n2 = np.array([oracle.c.dot(want_r[j], want_r[j]) for j in range(m)])
    if step:
        want_r = oracle.c.precondition(want_r, np.zeros(m), diag)
        want_x = np.stack([oracle.c.axpy(-1.0, want_r[j], want_x[j]) for j in range(m)])
    out_r = [torch.full((n,), np.nan, dtype=torch.float64, device="cuda") for _ in range(m)]
    out_x = [torch.full((n,), np.nan, dtype=torch.float64, device="cuda") for _ in range(m)]
    n2, _ = ctx.subspace_residual(coef, dev_rows(Q), dev_rows(A), out_r, out_x=out_x, mode=3 if step else 2,
                                  diag=dev_rows(diag[None])[0] if step else None, shift=np.zeros(m))
    assert np.array_equal(host(out_r), want_r) and np.array_equal(host(out_x), want_x)
    assert np.abs(n2 - want_n2).max() <= 1e-12 * want_n2.max()


@pytest.mark.parametrize("n,k,m", [(33, 5, 3), (4097, 9, 4), (50001, 12, 8), (20001, 7, 16)])
def test_subspace_residual_continues_from_the_p_space_parts(ctx, oracle, n, k, m):
    """accumulate: x_j and r_j start from what the output vectors hold (the P-space parts, which the reference adds first
    for the solutions, IterativeSolverTemplate.h:44-57) and the FMA chains continue from there"""
    rng = np.random.default_rng(5 * n + k + m)
    Q, A = rng.standard_normal((k, n)), rng.standard_normal((k, n))
    X0, R0 = rng.standard_normal((m, n)), rng.standard_normal((m, n))
    coef, lam = rng.standard_normal((k, m)), np.arange(1, m + 1) * 0.75
    want_x = oracle.c.gemm_outer(coef, Q, X0, fma=True)
    want_r = oracle.c.gemm_outer(coef, A, R0, fma=True)
    want_r = np.stack([oracle.c.axpy(-lam[j], want_x[j], want_r[j]) for j in range(m)])
    out_x, out_r = dev_rows(X0), dev_rows(R0)
    n2, _ = ctx.subspace_residual(coef, dev_rows(Q), dev_rows(A), out_r, lam=lam, out_x=out_x, accumulate=True)
    assert np.array_equal(host(out_x), want_x) and np.array_equal(host(out_r), want_r)
    want_n2 = np.array([oracle.c.dot(want_r[j], want_r[j]) for j in range(m)])
    assert np.abs(n2 - want_n2).max() <= 1e-12 * want_n2.max()


def test_fused_linear_equations_match_the_reference_run_here_at_2e6_rows(ctx, oracle):
    """LinearEquationsDavidsonFused against the reference's own class on its std::vector path, run in this process:
    n = 2e6, 8 right-hand sides, Q capped at 24 (BASELINE.json configs[2] at 1/100 of its rows)"""
    if oracle.ref is None:
        pytest.skip("oracle/_ref is not built")
    kw = dict(n=2_000_000, kind=N.KIND_LINEQ, nroots=8, hermitian=1, max_size_qspace=24)
    want, wsol = oracle.ref.solve(H.make_spec(**kw), want_solutions=True)
    got, sol = H.solve(ctx, H.make_spec(fused=1, **kw), want_solutions=True)
    plain, _ = H.solve(ctx, H.make_spec(**kw))
    assert want.converged == 1
    assert got.iterations == want.iterations and got.converged == want.converged
    assert [got.r_creations, got.q_creations, got.p_creations, got.d_creations] == \
        [want.r_creations, want.q_creations, want.p_creations, want.d_creations]
    assert got.kernel_launches < 0.5 * plain.kernel_launches
    for k in range(8):
        assert np.abs(sol[k] - wsol[k]).max() <= 1e-6 * max(1.0, np.abs(wsol[k]).max())
        assert got.errors[k] <= 1e-8


def test_davidson_residual_rejects_aliased_outputs(ctx):
    q = [torch.ones(64, dtype=torch.float64, device="cuda") for _ in range(2)]
    a = [torch.ones(64, dtype=torch.float64, device="cuda") for _ in range(2)]
    with pytest.raises(RuntimeError):
        ctx.davidson_residual(np.ones((2, 1)), q, a, [1.0], [q[1]])


@pytest.mark.parametrize("name", DAVIDSON)
def test_fused_solve_matches_reference_golden(ctx, name):
    want = GOLDEN[name]
    spec = H.make_spec(fused=1, **want["spec"])
    res, sol = H.solve(ctx, spec, want_solutions=True)
    assert res.iterations == want["iterations"], "iteration count differs from the reference"
    assert res.converged == want["converged"] and res.nwork_final == want["nwork_final"]
    assert [res.r_creations, res.q_creations, res.p_creations, res.d_creations] == want["creations"]
    ev = np.array([res.eigenvalues[i] for i in range(res.nroots)])
    assert np.all(np.diff(ev) > 0)
    assert np.abs(ev / np.array(want["eigenvalues"]) - 1).max() <= 1e-10
    for s, chk, head in zip(sol, want["solution_checksums"], want["solution_head"]):
        assert abs(np.sum(s) - chk) <= 1e-7 * max(1.0, np.abs(s).sum())
        assert np.abs(s[:8] - np.array(head)).max() <= 1e-7 * max(1.0, np.abs(np.array(head)).max())


def test_fused_path_needs_far_fewer_calls(ctx):
    kw = GOLDEN["banded_davidson_n100000_r4"]["spec"]
    plain, _ = H.solve(ctx, H.make_spec(**kw))
    fused, _ = H.solve(ctx, H.make_spec(fused=1, **kw))
    assert fused.iterations == plain.iterations
    assert fused.kernel_launches < 0.6 * plain.kernel_launches
    assert fused.handler_bytes < 0.7 * plain.handler_bytes
    for i in range(4):
        assert abs(fused.eigenvalues[i] / plain.eigenvalues[i] - 1) <= 1e-12


@pytest.mark.parametrize("kw", [
    dict(n=30000, nroots=16, max_size_qspace=8, nbuffers=8),   # BASELINE.json configs[3] in small: D space + root batches
    dict(n=30000, nroots=16, max_size_qspace=8),
    dict(n=20011, nroots=6, max_size_qspace=6, nbuffers=3),
    dict(n=20011, nroots=8, max_size_qspace=4, reset_D=3),
    dict(n=20011, nroots=5, max_size_qspace=5, nbuffers=2, hermitian=0),
], ids=lambda kw: "_".join(f"{k}{v}" for k, v in kw.items()))
def test_fused_solve_matches_the_reference_run_here(ctx, oracle, kw):
    """memory-capped configurations (Q-space limit -> D space, fewer buffers than roots, D-space resets): the fused
    driver against the reference's own templates run in this process on the same operator"""
    if oracle.ref is None:
        pytest.skip("oracle/_ref is not built")
    kw = dict(kw)
    kw.setdefault("hermitian", 1)
    want, _ = oracle.ref.solve(H.make_spec(kind=N.KIND_DAVIDSON, **kw))
    got, _ = H.solve(ctx, H.make_spec(kind=N.KIND_DAVIDSON, fused=1, **kw))
    assert got.iterations == want.iterations and got.converged == want.converged
    created = [got.r_creations, got.q_creations, got.p_creations, got.d_creations]
    expected = [want.r_creations, want.q_creations, want.p_creations, want.d_creations]
    # identical counters in every configuration, also with 16 roots through 8 buffers, where several new vectors at once
    # lie in the span of the subspace and the choice among them follows the rounding pattern of the overlaps: the fused
    # proposal step then measures them on the normalised vectors, as the reference does (FusedDavidson.h)
    assert created == expected
    # error estimates: against the reference's class run on the same CUDA handlers (fused = 0). The CPU run is not the
    # yardstick for them: where new vectors are dropped as redundant the choice follows the last bits of the overlaps, a
    # GPU tree sum and the CPU's sequential sum differ there, and the residuals of roots that are already converged end
    # anywhere between 1e-14 and 2e-9 (seen with 16 roots through 8 buffers: CPU 2.2e-9, both GPU paths 3.2e-10)
    plain, _ = H.solve(ctx, H.make_spec(kind=N.KIND_DAVIDSON, fused=0, **kw))
    assert [plain.r_creations, plain.q_creations, plain.p_creations, plain.d_creations] == expected
    for i in range(kw["nroots"]):
        assert abs(got.eigenvalues[i] / want.eigenvalues[i] - 1) <= 1e-10
        assert got.errors[i] <= 1e-8 and want.errors[i] <= 1e-8
        # 10 %; below 1e-11 they are the rounding noise of a converged residual
        assert abs(got.errors[i] - plain.errors[i]) <= 0.1 * plain.errors[i] + 1e-11, "error estimates follow the same path"


def test_error_estimates_are_true_residuals_with_a_capped_q_space(ctx):
    """16 roots through 8 buffers with the Q space capped at 8 (BASELINE.json configs[3] in small): the residuals of the
    exported solutions, from an operator application that is independent of the solver's stored actions, are what the
    solver reports. (A D space that reproduces the converged roots only to 1e-7 shows up here and nowhere else: the
    eigenvalues still agree to 1e-14.)"""
    n = 200_000
    spec = H.make_spec(n, kind=N.KIND_DAVIDSON, nroots=16, hermitian=1, max_size_qspace=8, nbuffers=8, fused=1)
    res, sol = H.solve(ctx, spec, want_solutions=True)
    assert res.converged == 1
    for k in range(16):
        x = torch.from_numpy(sol[k]).cuda()
        y = torch.empty_like(x)
        ctx.banded_apply(x, y, n, 0, 4, 1e-3)
        true = float((y - res.eigenvalues[k] * x).norm() / x.norm())
        assert true <= 1e-8 and true <= 3 * res.errors[k] + 1e-12


EQUATIONS = sorted(k for k, v in GOLDEN.items()
                   if v["spec"]["kind"] in (N.KIND_LINEQ, N.KIND_DIIS))


@pytest.mark.parametrize("name", EQUATIONS)
def test_equation_solvers_on_the_fused_x_space_match_reference_golden(ctx, name):
    """LinearEquationsDavidsonFused / NonLinearEquationsDIISFused (host/FusedEquations.h): the reference's solvers with all
    new overlap and action blocks from one Gram launch. Same counts as the reference; solutions as in test_solve_gpu.py."""
    want = GOLDEN[name]
    res, sol = H.solve(ctx, H.make_spec(fused=1, **want["spec"]), want_solutions=True)
    plain, _ = H.solve(ctx, H.make_spec(**want["spec"]))
    assert res.iterations == want["iterations"] and res.converged == want["converged"]
    assert res.nwork_final == want["nwork_final"]
    assert [res.r_creations, res.q_creations, res.p_creations, res.d_creations] == want["creations"]
    assert res.kernel_launches < plain.kernel_launches
    from test_solve_gpu import lineq_vtol
    vtol = lineq_vtol(want["spec"]) if want["spec"]["kind"] == N.KIND_LINEQ else 1e-7
    for s, chk, head in zip(sol, want["solution_checksums"], want["solution_head"]):
        assert abs(np.sum(s) - chk) <= vtol * max(1.0, np.abs(s).sum())
        assert np.abs(s[:8] - np.array(head)).max() <= vtol * max(1.0, np.abs(np.array(head)).max())


def unchained_mgs(ctx, rs, thresh):
    """the R-R Gram-Schmidt as FusedDavidson.h runs it without the chain: one Gram row, then a step per pivot with the
    coefficients formed on the host"""
    w = len(rs)
    row, nulls = None, []
    for i in range(w):
        if row is None:
            row = ctx.gemm_inner([rs[i]], rs[i:])[0]
        norm = np.sqrt(abs(row[0]))
        if norm > thresh:
            dots = ctx.mgs_step_dots(1.0 / norm, rs[i], row[1:] / norm, rs[i + 1:])
            row = dots[1:]
        else:
            nulls.append(i)
            row = None
    return nulls


@pytest.mark.parametrize("n,w", [(1000, 1), (4097, 2), (50001, 4), (20000, 7), (3001, 17)])
def test_mgs_chain_equals_the_steps_one_by_one(ctx, n, w):
    """itsolv_mgs_chain_f64: the coefficients the kernel tails leave on the device are the host's, bit for bit, so the
    vectors after the chain are bit-identical to the step-by-step sequence"""
    rng = np.random.default_rng(n + w)
    R = rng.standard_normal((w, n)) + 0.3 * rng.standard_normal(n)[None, :]  # correlated: the projections matter
    a, b = dev_rows(R), dev_rows(R)
    old = ctx.set_option("MGS_CHAIN", 1)
    try:
        rows = ctx.mgs_chain(a, 1e-10)
    finally:
        ctx.set_option("MGS_CHAIN", old)
    assert unchained_mgs(ctx, b, 1e-10) == []
    assert np.array_equal(host(a), host(b))
    Q = host(a)
    assert np.abs(Q @ Q.T - np.eye(w)).max() <= 1e-12
    assert rows.size == w + w * (w + 1) // 2 and abs(rows[0] - R[0] @ R[0]) <= 1e-12 * (R[0] @ R[0])


@pytest.mark.parametrize("n,k,m,keep,scaled", [(50001, 12, 4, [0, 1, 2, 3], True), (20000, 5, 4, [1, 3], True),
                                               (4097, 3, 1, [0], False), (70002, 24, 8, [0, 2, 3, 5, 6, 7], True),
                                               (33, 2, 3, [2], False), (100000, 128, 6, [0, 1, 2, 3, 4, 5], False)])
def test_projection_with_the_first_gram_row_in_its_tail(ctx, n, k, m, keep, scaled):
    """itsolv_project_mgs_chain_f64 against gemm_outer(_scaled) followed by itsolv_mgs_chain_f64 on the kept vectors: the
    projected vectors that are not kept are bit-identical; the first Gram row comes from a different kernel (summation
    order), so the orthonormalised vectors agree to rounding and the rows to 1e-12"""
    rng = np.random.default_rng(n + k + m)
    X = rng.standard_normal((k, n))
    Y = rng.standard_normal((m, n)) + 0.2 * rng.standard_normal(n)[None, :]
    alpha = 0.05 * rng.standard_normal((k, m))
    scale = rng.uniform(0.5, 2.0, m)
    xs, a, b = dev_rows(X), dev_rows(Y), dev_rows(Y)
    rows = ctx.project_mgs_chain(alpha, xs, a, scale if scaled else None, keep)
    if scaled:
        ctx.gemm_outer_scaled(alpha, xs, b, scale)
    else:
        ctx.gemm_outer(alpha, xs, b)
    want_rows = ctx.mgs_chain([b[i] for i in keep])
    A, B = host(a), host(b)
    dropped = [j for j in range(m) if j not in keep]
    assert np.array_equal(A[dropped], B[dropped])
    assert np.abs(A[keep] - B[keep]).max() <= 1e-13
    assert np.abs(rows - want_rows).max() <= 1e-12 * max(1.0, np.abs(want_rows).max())
    Q = A[keep]
    assert np.abs(Q @ Q.T - np.eye(len(keep))).max() <= 1e-12


def test_mgs_chain_leaves_a_null_pivot_alone(ctx):
    rng = np.random.default_rng(11)
    R = rng.standard_normal((4, 6000))
    R[1] = 0.0  # a null vector in the middle: neither scaled nor projected out of the later ones
    a, b = dev_rows(R), dev_rows(R)
    old = ctx.set_option("MGS_CHAIN", 1)
    try:
        ctx.mgs_chain(a, 1e-10)
    finally:
        ctx.set_option("MGS_CHAIN", old)
    assert unchained_mgs(ctx, b, 1e-10) == [1]
    A, B = host(a), host(b)
    assert np.array_equal(A[1], np.zeros(6000)) and np.array_equal(A[0], B[0])
    assert np.abs(A - B).max() <= 1e-13  # after the null pivot the next row comes from a different kernel: rounding only
    keep = [0, 2, 3]
    assert np.abs(A[keep] @ A[keep].T - np.eye(3)).max() <= 1e-12


@pytest.mark.parametrize("name", DAVIDSON)
def test_fused_solve_with_chained_gram_schmidt_matches_reference_golden(ctx, name):
    want = GOLDEN[name]
    old = ctx.set_option("MGS_CHAIN", 1)
    try:
        res, sol = H.solve(ctx, H.make_spec(fused=1, **want["spec"]), want_solutions=True)
        projected, _ = H.solve(ctx, H.make_spec(fused=1, **want["spec"]))
        # without the first Gram row in the projection's tail: the chain and the steps one by one are the same launches
        # with the same coefficients, bit for bit
        ctx.set_option("PROJECT_CHAIN", -1)
        plain, _ = H.solve(ctx, H.make_spec(fused=1, **want["spec"]))
        ctx.set_option("MGS_CHAIN", -1)  # the steps one by one, coefficients formed on the host
        unchained, _ = H.solve(ctx, H.make_spec(fused=1, **want["spec"]))
        assert unchained.kernel_launches == plain.kernel_launches  # the same launches, only the waiting differs
        assert projected.kernel_launches <= plain.kernel_launches  # one launch fewer per projected working set
    finally:
        ctx.set_option("MGS_CHAIN", old)
        ctx.set_option("PROJECT_CHAIN", 0)
    assert res.iterations == want["iterations"] and res.converged == want["converged"]
    assert [res.r_creations, res.q_creations, res.p_creations, res.d_creations] == want["creations"]
    ev = np.array([res.eigenvalues[i] for i in range(res.nroots)])
    assert np.abs(ev / np.array(want["eigenvalues"]) - 1).max() <= 1e-10
    # bit-identical vectors along the way: the same eigenvalues as without the chain, to the last bit
    assert [plain.eigenvalues[i] for i in range(res.nroots)] == [unchained.eigenvalues[i] for i in range(res.nroots)]
    assert np.abs(ev / np.array([plain.eigenvalues[i] for i in range(res.nroots)]) - 1).max() <= 1e-12
    for s, chk in zip(sol, want["solution_checksums"]):
        assert abs(np.sum(s) - chk) <= 1e-7 * max(1.0, np.abs(s).sum())
