"""Parity of the CUDA kernels, called through the C ABI (include/itsolv_b200.h), with the CPU oracle on the same
seeded inputs. Tolerances: element-wise ops that the kernels evaluate with the reference's own operation order
(axpy, scal, copy, fill, preconditioner, sparse ops, select) are BIT-EXACT; reductions (dot, gemm_inner) differ from the
oracle's sequential sum only by summation order and must agree to 1e-12 relative to |x|.|y| (BASELINE.json north_star:
"Gram/subspace matrices to 1e-12 relative"); gemm_outer uses one FMA per term: bit-exact against the FMA model of the
oracle and within 1e-14 relative of the reference's two-rounding loop."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rows(t):
    """views into one 2-D allocation: for odd n every second row is only 8-byte aligned (scalar / plain-load paths)"""
    return [t[i] for i in range(t.shape[0])]


def dev_rows(a):
    """separate allocations, as the solver's Q/R vectors are: always 16-byte aligned (TMA / 128-bit paths)"""
    return [dev(a[i]) for i in range(a.shape[0])]


def host(ts):
    return np.stack([t.cpu().numpy() for t in ts])


def vectors(rng, k, n):
    return rng.standard_normal((k, n)) * (1.0 + 0.1 * np.arange(k)[:, None])


def gram_close(got, want, X, Y, tol=1e-12):
    scale = np.linalg.norm(X, axis=1)[:, None] * np.linalg.norm(Y, axis=1)[None, :] + 1e-300
    err = np.abs(got - want) / scale
    assert err.max() <= tol, f"max norm-relative error {err.max():.3e}"


SIZES = [1, 2, 3, 15, 16, 17, 31, 63, 64, 65, 257, 1000, 4097, 100003]


@pytest.mark.parametrize("n", SIZES)
def test_blas1_bit_exact(ctx, oracle, n):
    rng = np.random.default_rng(n)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    dx, dy = dev(x), dev(y)
    ctx.axpy(0.37, dx, dy)
    assert np.array_equal(dy.cpu().numpy(), oracle.c.axpy(0.37, x, y))
    ctx.scal(-1.7, dx)
    assert np.array_equal(dx.cpu().numpy(), oracle.c.scal(-1.7, x))
    dz = torch.empty_like(dx)
    ctx.copy(dz, dx)
    assert np.array_equal(dz.cpu().numpy(), dx.cpu().numpy())
    ctx.fill(2.5, dz)
    assert np.array_equal(dz.cpu().numpy(), np.full(n, 2.5))


def test_blas1_unaligned_views(ctx, oracle):
    rng = np.random.default_rng(7)
    x, y = rng.standard_normal(1001), rng.standard_normal(1001)
    dx, dy = dev(x), dev(y)
    # views starting at an odd element are only 8-byte aligned: the scalar path must give the same bits
    ctx.axpy(1.25, dx[1:], dy[1:])
    want = y.copy()
    want[1:] = oracle.c.axpy(1.25, x[1:].copy(), y[1:].copy())
    assert np.array_equal(dy.cpu().numpy(), want)
    got = ctx.dot(dx[1:], dy[1:])
    ref = oracle.c.dot(x[1:].copy(), want[1:].copy())
    assert abs(got - ref) <= 1e-12 * np.linalg.norm(x[1:]) * np.linalg.norm(want[1:])


@pytest.mark.parametrize("n", SIZES + [1 << 20])
def test_dot(ctx, oracle, n):
    rng = np.random.default_rng(100 + n)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    dx, dy = dev(x), dev(y)
    got = ctx.dot(dx, dy)
    scale = np.linalg.norm(x) * np.linalg.norm(y)
    assert abs(got - oracle.c.dot(x, y)) <= 1e-12 * scale
    assert abs(got - oracle.c.dot_long(x, y)) <= 1e-13 * scale  # the tree sum is closer to the exact value than 1e-12
    got_xx = ctx.dot(dx, dx)
    assert abs(got_xx - oracle.c.dot(x, x)) <= 1e-12 * np.linalg.norm(x)**2
    assert ctx.dot(dx, dy) == got  # run-to-run reproducible


SHAPES = [(1, 1), (4, 1), (1, 7), (4, 4), (4, 16), (3, 5), (16, 16), (16, 64), (8, 100), (16, 128), (33, 65), (128, 128)]


@pytest.mark.parametrize("k,m", SHAPES)
@pytest.mark.parametrize("n", [1, 17, 1000, 50001])
def test_gemm_inner(ctx, oracle, k, m, n):
    if k * m >= 4096 and n > 20000:
        n = 20000
    rng = np.random.default_rng(k * 1000 + m + n)
    X, Y = vectors(rng, k, n), vectors(rng, m, n)
    want = oracle.c.gemm_inner(X, Y)
    xs, ys = dev_rows(X), dev_rows(Y)
    got = ctx.gemm_inner(xs, ys)
    gram_close(got, want, X, Y)
    assert np.array_equal(got, ctx.gemm_inner(xs, ys))  # run-to-run reproducible
    dX, dY = dev(X), dev(Y)
    gram_close(ctx.gemm_inner(rows(dX), rows(dY)), want, X, Y)


def test_gemm_inner_equals_loop_of_dots(ctx):
    """the defining property the reference asserts for every handler (test/array/testGemm.cpp:58-88)"""
    rng = np.random.default_rng(5)
    X, Y = vectors(rng, 5, 3001), vectors(rng, 7, 3001)
    dX, dY = dev(X), dev(Y)
    G = ctx.gemm_inner(rows(dX), rows(dY))
    for i in range(5):
        for j in range(7):
            assert abs(G[i, j] - ctx.dot(dX[i], dY[j])) <= 1e-13 * np.linalg.norm(X[i]) * np.linalg.norm(Y[j])


def test_gemm_inner_aliased_and_symmetric(ctx, oracle):
    rng = np.random.default_rng(11)
    X = vectors(rng, 6, 12345)
    dX = dev(X)
    xs = rows(dX)
    G = ctx.gemm_inner(xs, xs)  # overlap(x, x): every vector is loaded once
    gram_close(G, oracle.c.gemm_inner(X, X), X, X)
    assert np.array_equal(G, G.T)
    mixed = ctx.gemm_inner(xs[:3], [xs[1], xs[4], xs[1]])
    gram_close(mixed, oracle.c.gemm_inner(X[:3], X[[1, 4, 1]]), X[:3], X[[1, 4, 1]])


def test_gemm_inner_unaligned(ctx, oracle):
    rng = np.random.default_rng(12)
    X, Y = vectors(rng, 3, 4098), vectors(rng, 4, 4098)
    dX, dY = dev(X), dev(Y)
    got = ctx.gemm_inner([r[1:] for r in rows(dX)], [r[1:] for r in rows(dY)])
    gram_close(got, oracle.c.gemm_inner(X[:, 1:], Y[:, 1:]), X[:, 1:], Y[:, 1:])


def test_gemm_inner_wide_panels(ctx, oracle):
    """more than ITSOLV_MAX_PANEL vectors on a side are processed block by block"""
    rng = np.random.default_rng(13)
    X, Y = vectors(rng, 130, 777), vectors(rng, 3, 777)
    got = ctx.gemm_inner(rows(dev(X)), rows(dev(Y)))
    gram_close(got, oracle.c.gemm_inner(X, Y), X, Y)


@pytest.mark.parametrize("k,m", [(1, 1), (1, 4), (20, 4), (4, 1), (7, 3), (40, 16), (100, 16), (5, 17), (16, 40), (128, 2)])
@pytest.mark.parametrize("n", [1, 2, 33, 1000, 50001])
def test_gemm_outer(ctx, oracle, k, m, n):
    rng = np.random.default_rng(k * 100 + m + n)
    X, Y = vectors(rng, k, n), vectors(rng, m, n)
    alpha = rng.standard_normal((k, m))
    xs, ys = dev_rows(X), dev_rows(Y)
    ctx.gemm_outer(alpha, xs, ys)
    got = host(ys)
    assert np.array_equal(got, oracle.c.gemm_outer(alpha, X, Y, fma=True))
    dX, dY = dev(X), dev(Y)
    ctx.gemm_outer(alpha, rows(dX), rows(dY))  # rows of one allocation: unaligned for odd n
    assert np.array_equal(dY.cpu().numpy(), got)
    ref = oracle.c.gemm_outer(alpha, X, Y)  # the reference's loop of axpys (two roundings per term)
    bound = 1e-14 * (np.abs(Y) + np.abs(alpha).T @ np.abs(X)) + 1e-300
    assert (np.abs(got - ref) <= bound).all()
    # beta_zero: the targets are overwritten, not read (NaN must not leak)
    dZ = torch.full_like(dY, float("nan"))
    ctx.gemm_outer(alpha, rows(dX), rows(dZ), beta_zero=True)
    assert np.array_equal(dZ.cpu().numpy(), oracle.c.gemm_outer(alpha, X, np.zeros_like(Y), fma=True))


def test_gemm_outer_aliased_targets(ctx, oracle):
    """a target that is also a source keeps the reference's sequential axpy meaning"""
    rng = np.random.default_rng(21)
    V = vectors(rng, 3, 1001)
    alpha = rng.standard_normal((2, 2))
    dV = dev(V)
    v = rows(dV)
    ctx.gemm_outer(alpha, [v[0], v[1]], [v[1], v[2]])
    want = V.copy()
    for i, xi in enumerate([0, 1]):
        for j, yj in enumerate([1, 2]):
            want[yj] = oracle.c.axpy(alpha[i, j], want[xi].copy(), want[yj].copy())
    assert np.array_equal(dV.cpu().numpy(), want)


@pytest.mark.parametrize("w", [1, 4, 16])
@pytest.mark.parametrize("n", [1, 2, 1001, 65536])
def test_precondition_bit_exact(ctx, oracle, w, n):
    rng = np.random.default_rng(w + n)
    R = vectors(rng, w, n)
    diag = np.arange(1, n + 1, dtype=np.float64)
    shift = rng.uniform(0.5, 4.5, w)
    rs, dd = dev_rows(R), dev(diag)
    ctx.precondition(rs, dd, shift)
    assert np.array_equal(host(rs), oracle.c.precondition(R, shift, diag))
    dR = dev(R)
    ctx.precondition(rows(dR), dd, shift)
    assert np.array_equal(dR.cpu().numpy(), oracle.c.precondition(R, shift, diag))


@pytest.mark.parametrize("mode", ["min", "max"])
def test_select_when_the_boundary_bucket_exceeds_the_candidate_buffer(ctx, oracle, mode):
    """1.3e6 values that share sign, exponent and the four leading mantissa bits: more candidates after the two leading
    key bytes than the candidate buffer (2^20) holds, so the later passes keep reading the vector; many exact ties"""
    rng = np.random.default_rng(3)
    x = 1.0 + np.round(0.05 * rng.random(1_300_000), 4)
    kw = {"max": True} if mode == "max" else {}
    idx, val = ctx.select(dev(x), 9, **kw)
    want_idx, want_val = oracle.c.select(x, 9, **kw)
    assert np.array_equal(idx, want_idx) and np.array_equal(val, want_val)


@pytest.mark.parametrize("mode", ["min", "max"])
def test_select_on_a_shard_of_a_sorted_diagonal(ctx, oracle, mode):
    """the shard of rank 7 of 8 of d(i) = i + 1: every value shares sign, exponent and leading mantissa bits, so the
    boundary bucket is the whole shard after two key bytes and small only after four (the second gather attempt)"""
    n, offset = 1_500_000, 70_000_000
    x = np.arange(offset + 1, offset + n + 1, dtype=np.float64)
    kw = {"max": True} if mode == "max" else {}
    idx, val = ctx.select(dev(x), 6, global_offset=offset, **kw)
    want_idx, want_val = oracle.c.select(x, 6, **kw)
    assert np.array_equal(idx, want_idx + offset) and np.array_equal(val, want_val)


def test_select_among_more_equal_values_than_the_candidate_buffer_holds(ctx, oracle):
    """1.2e6 equal values below a few larger ones: the boundary bucket never fits, every pass reads the vector and the
    index digits decide (the highest indices survive the reference's heap)"""
    x = np.full(1_200_000, 2.5)
    x[[7, 70_000, 1_100_000]] = [3.0, 4.0, 5.0]
    idx, val = ctx.select(dev(x), 7, max=True)
    want_idx, want_val = oracle.c.select(x, 7, max=True)
    assert np.array_equal(idx, want_idx) and np.array_equal(val, want_val)
    idx, val = ctx.select(dev(x), 5)
    want_idx, want_val = oracle.c.select(x, 5)
    assert np.array_equal(idx, want_idx) and np.array_equal(val, want_val)


def same_bits_or_both_nan(got, want):
    nan = np.isnan(want)
    return np.array_equal(np.isnan(got), nan) and np.array_equal(got[~nan].view(np.uint64), want[~nan].view(np.uint64))


def test_precondition_special_values(ctx, oracle):
    """the quotient follows IEEE division in the corners too: signed zeros (the kernels answer zero numerators
    without the generic division), subnormal and tiny numerators, zero / infinite / NaN denominators"""
    special = np.array([0.0, -0.0, 5e-324, -5e-324, 1e-310, 1e-200, -1e-200, 1e-37, 7.5e-37, 1.0, -1.0, 1e300, np.inf,
                        -np.inf, np.nan])
    dens = np.array([1.0, -1.0, 3.0, -3.0, 0.0, 1e-300, -1e300, np.inf, -np.inf, np.nan, 1e-15, 0.5])
    num, den = [a.ravel() for a in np.meshgrid(special, dens)]
    shift = np.array([0.0, 2.0])
    diag = np.where(np.isfinite(den), den - 1e-15, den)  # (diag - 0) + 1e-15 reproduces most of the denominators
    R = np.stack([num, num])
    with np.errstate(all="ignore"):
        want = oracle.c.precondition(R, shift, diag)
    rs = dev_rows(R)
    ctx.precondition(rs, dev(diag), shift)
    assert same_bits_or_both_nan(host(rs), want)
    # the same corners through the fused residual kernel (x = 1 * 0, r = 1 * num - shift * x, then the preconditioner),
    # expected values from the oracle's sequence of the separate steps
    zero = np.zeros_like(num)
    q, a = [dev(zero)], [dev(num)]
    out = [dev(zero) for _ in range(2)]
    ctx.davidson_residual(np.ones((1, 2)), q, a, shift, out, diag=dev(diag))
    with np.errstate(all="ignore"):
        x = oracle.c.gemm_outer(np.ones((1, 2)), zero[None], np.zeros((2, num.size)), fma=True)
        r = oracle.c.gemm_outer(np.ones((1, 2)), num[None], np.zeros((2, num.size)), fma=True)
        r = np.stack([oracle.c.axpy(-shift[j], x[j], r[j]) for j in range(2)])
        want = oracle.c.precondition(r, shift, diag)
    assert same_bits_or_both_nan(host(out), want)


@pytest.mark.parametrize("n,nsel", [(1, 1), (10, 3), (1000, 4), (1000, 500), (100003, 16), (4096, 4096)])
@pytest.mark.parametrize("mode", ["min", "max", "absmax", "absmin", "maxdot"])
def test_select(ctx, oracle, n, nsel, mode):
    rng = np.random.default_rng(n + nsel)
    x = np.round(rng.standard_normal(n), 2)  # rounding creates many exact ties: the higher index must win
    y = np.round(rng.standard_normal(n), 1)
    dx, dy = dev(x), dev(y)
    kw = dict(min={}, max={"max": True}, absmax={"max": True, "ignore_sign": True}, absmin={"ignore_sign": True},
              maxdot={"y": y})[mode]
    want_idx, want_val = oracle.c.select(x, nsel, **kw)
    if mode == "maxdot":
        kw = {"y": dy}
    idx, val = ctx.select(dx, nsel, **kw)
    assert np.array_equal(idx, want_idx)
    assert np.array_equal(val, want_val)


def test_select_ascending_diagonal(ctx, oracle):
    """the use made by solve(): the nroots smallest diagonal elements (reference IterativeSolverTemplate.h:340-349)"""
    d = np.arange(1, 200001, dtype=np.float64)
    idx, val = ctx.select(dev(d), 4)
    assert idx.tolist() == [0, 1, 2, 3] and val.tolist() == [1.0, 2.0, 3.0, 4.0]
    z = np.zeros(5000)
    idx, _ = ctx.select(dev(z), 3)  # all equal: the three highest indices survive the reference's heap
    assert idx.tolist() == [4997, 4998, 4999]
    assert np.array_equal(idx, oracle.c.select(z, 3)[0])


@pytest.mark.parametrize("n", [9, 1000, 20011])
@pytest.mark.parametrize("b", [0, 1, 4])
def test_banded_apply_bit_exact(ctx, oracle, n, b):
    rng = np.random.default_rng(n + b)
    x = rng.standard_normal(n)
    dx = dev(x)
    dy = torch.empty_like(dx)
    ctx.banded_apply(dx, dy, n, 0, b, 1e-3)
    assert np.array_equal(dy.cpu().numpy(), oracle.c.banded_apply(x, b, 1e-3))


def test_banded_apply_sharded_with_halos(ctx, oracle):
    """two shards with explicit halo rows reproduce the unsharded action bit for bit"""
    n, b = 1001, 4
    x = np.random.default_rng(3).standard_normal(n)
    want = oracle.c.banded_apply(x, b, 1e-3)
    cut = 501
    lo, hi = dev(x[:cut]), dev(x[cut:])
    ylo, yhi = torch.empty_like(lo), torch.empty_like(hi)
    ctx.banded_apply(lo, ylo, n, 0, b, 1e-3, x_hi=dev(x[cut:cut + b]))
    ctx.banded_apply(hi, yhi, n, cut, b, 1e-3, x_lo=dev(x[cut - b:cut]))
    assert np.array_equal(np.concatenate([ylo.cpu().numpy(), yhi.cpu().numpy()]), want)


def test_full_size_properties(ctx):
    """BASELINE.json config[1] size (n = 1e7, 4 roots x 16 Q vectors): size-independent properties."""
    n, k, m = 10_000_000, 4, 16
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(k, n, dtype=torch.float64, device="cuda", generator=g)
    Y = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    xs, ys = rows(X), rows(Y)
    G = ctx.gemm_inner(xs, ys)
    # linearity in the first argument: <2 x0 + x1, y> = 2 <x0,y> + <x1,y>
    z = (2.0 * X[0] + X[1]).contiguous()
    Gz = ctx.gemm_inner([z], ys)
    assert np.allclose(Gz[0], 2 * G[0] + G[1], rtol=0, atol=1e-12 * n)
    # transpose symmetry: gemm_inner(Y, X) == gemm_inner(X, Y)^T to rounding
    GT = ctx.gemm_inner(ys, xs)
    assert np.abs(GT.T - G).max() <= 1e-12 * n
    # against torch's own fp64 matmul (a plain library reference of the same contraction)
    Gt = (X @ Y.T).cpu().numpy()
    assert np.abs(Gt - G).max() <= 1e-12 * n
    # expansion followed by contraction: y_j += sum_i a_ij x_i  =>  <x_l, y_j'> = <x_l, y_j> + sum_i a_ij <x_l, x_i>
    alpha = np.random.default_rng(2).standard_normal((k, m))
    S = ctx.gemm_inner(xs, xs)
    ctx.gemm_outer(alpha, xs, ys)
    G2 = ctx.gemm_inner(xs, ys)
    assert np.abs(G2 - (G + S @ alpha)).max() <= 1e-11 * n


def csr_reference(row_ptr, col, val, x):
    """row sums in ascending entry order, product and sum rounded separately (the CPU twin's arithmetic)"""
    n = row_ptr.size - 1
    y = np.zeros(n)
    length = np.diff(row_ptr)
    for j in range(int(length.max()) if n else 0):
        rows = np.nonzero(length > j)[0]
        e = row_ptr[rows] + j
        y[rows] = y[rows] + val[e] * x[col[e]]
    return y


@pytest.mark.parametrize("n,b,w,kind", [(1, 0, 1, "band"), (255, 4, 1, "band"), (256, 4, 2, "band"), (257, 4, 3, "band"),
                                        (5000, 4, 4, "band"), (5000, 7, 5, "band"), (3000, 4, 8, "band"),
                                        (3000, 4, 9, "band"), (2000, 130, 2, "band"), (4000, 3, 4, "scattered"),
                                        (1500, 12, 3, "long_rows"), (1000, 2, 2, "empty_rows")])
def test_csr_apply_multi_bit_exact(ctx, n, b, w, kind):
    """the stored-CSR operator kernel (shared-memory staged rows and x window; global fallback for columns outside the
    window; several rounds for rows with many entries) against the sequential row sums"""
    rng = np.random.default_rng(n + 17 * b + w)
    rows_cols = []
    for i in range(n):
        if kind == "empty_rows" and i % 3 == 0:
            rows_cols.append(np.zeros(0, dtype=np.int64))
            continue
        lo, hi = max(0, i - b), min(n - 1, i + b)
        c = np.arange(lo, hi + 1)
        if kind == "scattered":  # a few entries far outside the band: read through the global fallback
            c = np.unique(np.concatenate([c, rng.integers(0, n, 3)]))
        if kind == "long_rows" and i % 50 == 0:  # more entries in 256 rows than one staging round holds
            c = np.arange(max(0, i - 10 * b), min(n - 1, i + 10 * b) + 1)
        rows_cols.append(c)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    row_ptr[1:] = np.cumsum([c.size for c in rows_cols])
    col = np.concatenate(rows_cols).astype(np.int32) if n else np.zeros(0, dtype=np.int32)
    val = rng.standard_normal(col.size)
    X = rng.standard_normal((w, n))
    band = 10 * b if kind == "long_rows" else b  # "scattered": the window stays 2b wide, the rest is read globally
    d_rp, d_col, d_val = torch.from_numpy(row_ptr).cuda(), torch.from_numpy(col).cuda(), dev(val)
    xs = dev_rows(X)
    ys = [torch.full((n,), np.nan, dtype=torch.float64, device="cuda") for _ in range(w)]
    ctx.csr_apply_multi(d_rp.data_ptr(), d_col.data_ptr() if col.size else 0, d_val, xs, ys, n, 0, min(band, n))
    want = np.stack([csr_reference(row_ptr, col, val, X[k]) for k in range(w)])
    assert np.array_equal(host(ys), want)


def test_csr_apply_multi_sharded_with_halos(ctx):
    n, b, w, cut = 3000, 5, 3, 1234
    rng = np.random.default_rng(5)
    cols = [np.arange(max(0, i - b), min(n - 1, i + b) + 1) for i in range(n)]
    X = rng.standard_normal((w, n))

    def shard(r0, r1):
        row_ptr = np.zeros(r1 - r0 + 1, dtype=np.int64)
        row_ptr[1:] = np.cumsum([cols[i].size for i in range(r0, r1)])
        col = np.concatenate(cols[r0:r1]).astype(np.int32)
        return row_ptr, col

    full_ptr, full_col = shard(0, n)
    val = rng.standard_normal(full_col.size)
    want = np.stack([csr_reference(full_ptr, full_col, val, X[k]) for k in range(w)])
    for r0, r1 in ((0, cut), (cut, n)):
        row_ptr, col = shard(r0, r1)
        v = val[full_ptr[r0]:full_ptr[r1]]
        d_rp, d_col, d_val = torch.from_numpy(row_ptr).cuda(), torch.from_numpy(col).cuda(), dev(v)
        xs = dev_rows(X[:, r0:r1])
        lo = dev_rows(X[:, r0 - b:r0]) if r0 > 0 else None
        hi = dev_rows(X[:, r1:r1 + b]) if r1 < n else None
        ys = [torch.full((r1 - r0,), np.nan, dtype=torch.float64, device="cuda") for _ in range(w)]
        ctx.csr_apply_multi(d_rp.data_ptr(), d_col.data_ptr(), d_val, xs, ys, n, r0, b, x_lo=lo, x_hi=hi)
        assert np.array_equal(host(ys), want[:, r0:r1])
