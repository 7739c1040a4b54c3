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
    # one R-R Gram-Schmidt step == scal of the pivot followed by axpys onto the later vectors
    R = rng.standard_normal((4, n))
    ov = rng.standard_normal(3)
    rs = dev_rows(R)
    ctx.mgs_step(0.731, rs[0], ov, rs[1:])
    pivot = oracle.c.scal(0.731, R[0].copy())
    want = np.stack([pivot] + [oracle.c.axpy(-ov[j], pivot.copy(), R[j + 1].copy()) for j in range(3)])
    assert np.array_equal(host(rs), want)


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
