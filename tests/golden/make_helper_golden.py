"""Golden vectors for the host subspace algebra (eigenproblem, svd_system, solve_LinearEquations, solve_DIIS of the
reference's itsolv/helper-implementation.h:263-669), generated with numpy/scipy ONLY: no line of
iterative_solver_b200/host/helper_lapack.cpp or oracle/helper_literal.cpp is involved. The reference itself holds no
fixtures at this boundary; these pin both restatements independently of each other.

    python tests/golden/make_helper_golden.py        -> tests/golden/helper_golden.npz

What is stored per case: the inputs exactly as the C++ functions receive them (flat arrays, the reference's storage
conventions) and the outputs in the reference's conventions:
  eigenproblem: eigenvalues ascending; eigenvectors column-major dimension x rank, the sign of each fixed so that its
      largest-magnitude component among the first `rank` is positive (helper-implementation.h:434-440); hermitian:
      normalised x^T S x = 1 by construction of the S^-1/2 transformation; non-hermitian: x^T S x = 1 and largest
      component overall positive (:451-506).
  rank-deficient S: the reference keeps the FIRST `rank` columns of the ascending eigen-decomposition of S
      (`singularValues.head(rank)`, :370) and zeroes the factor of every value <= 1e-14 (:371-372); the expected values
      follow that statement, restated here in numpy.
  svd_system(hermitian): eigenpairs of value <= threshold in descending order (:263-283, :167-195).
  svd_system(general): singular triplets of value < threshold, ascending (:12-32).
  solve_LinearEquations: QR solve of the row-major system (:595-616); augmented hessian: lowest eigenpair of the bordered
      generalised problem (:561-594).
  solve_DIIS: pseudo-inverse solution of the bordered system (:619-669).
"""
import os

import numpy as np
import scipy.linalg as sl

HERE = os.path.dirname(os.path.abspath(__file__))


def fix_sign(x, upto):
    x = x.copy()
    for k in range(x.shape[1]):
        m = np.argmax(np.abs(x[:upto, k]))
        if x[m, k] < 0:
            x[:, k] = -x[:, k]
    return x


def hermitian_case(rng, n, cond=1e2):
    a = rng.standard_normal((n, n))
    q, _ = np.linalg.qr(a)
    s = (q * np.geomspace(1.0, 1.0 / cond, n)) @ q.T
    s = 0.5 * (s + s.T)
    h = rng.standard_normal((n, n))
    h = 0.5 * (h + h.T) + np.diag(np.arange(n, dtype=float))
    w, v = sl.eigh(h, s)  # ascending, v^T s v = 1
    return h, s, w, fix_sign(v, n)


def nonhermitian_case(rng, n):
    a = rng.standard_normal((n, n))
    s = a @ a.T / n + np.eye(n)
    # real, well separated spectrum: H = S X diag(lam) X^-1 with a well conditioned X
    x = np.eye(n) + 0.2 * rng.standard_normal((n, n))
    lam = np.arange(1, n + 1, dtype=float) + 0.1 * rng.standard_normal(n)
    h = s @ x @ np.diag(lam) @ np.linalg.inv(x)
    w, v = sl.eig(h, s)
    assert np.abs(w.imag).max() < 1e-9
    order = np.argsort(w.real, kind="stable")
    w, v = w.real[order], v.real[:, order]
    for k in range(n):  # x^T S x = 1, largest component positive
        v[:, k] /= np.sqrt(v[:, k] @ s @ v[:, k])
        if v[np.argmax(np.abs(v[:, k])), k] < 0:
            v[:, k] = -v[:, k]
    return h, s, w, v


def rank_deficient_case(rng, n, null, threshold):
    # small overall scale: the null eigenvalues come out as rounding noise ~1e-17, safely below the reference's absolute
    # 1e-14 guard (:371), whichever LAPACK/Eigen routine computes them
    b = 0.03 * rng.standard_normal((n, n - null))
    s = b @ b.T
    s = 0.5 * (s + s.T)
    h = rng.standard_normal((n, n))
    h = 0.5 * (h + h.T)
    # the reference's statements (helper-implementation.h:342-374) in numpy
    sv, v = np.linalg.eigh(s)
    rank = int(np.sum(sv >= threshold * sv.max()))
    svmh = np.where(sv[:rank] > 1e-14, 1.0 / np.sqrt(np.abs(sv[:rank])), 0.0)
    hbar = (svmh[:, None] * (v[:, :rank].T @ h @ v[:, :rank])) * svmh[None, :]
    w = np.linalg.eigvalsh(0.5 * (hbar + hbar.T))
    return h, s, w, rank


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    for n in (1, 2, 5, 12, 40, 150):
        h, s, w, v = hermitian_case(rng, n)
        out[f"eig_herm_{n}_H"], out[f"eig_herm_{n}_S"] = h.ravel(), s.ravel(order="F")
        out[f"eig_herm_{n}_w"], out[f"eig_herm_{n}_v"] = w, v.ravel(order="F")
    for n in (2, 6, 20):
        h, s, w, v = nonhermitian_case(rng, n)
        out[f"eig_gen_{n}_H"], out[f"eig_gen_{n}_S"] = h.ravel(), s.ravel(order="F")
        out[f"eig_gen_{n}_w"], out[f"eig_gen_{n}_v"] = w, v.ravel(order="F")
    for n, null in ((8, 2), (30, 5)):
        thr = 1e-10
        h, s, w, rank = rank_deficient_case(rng, n, null, thr)
        out[f"eig_def_{n}_H"], out[f"eig_def_{n}_S"] = h.ravel(), s.ravel(order="F")
        out[f"eig_def_{n}_w"], out[f"eig_def_{n}_rank"], out[f"eig_def_{n}_thr"] = w, np.array([rank]), np.array([thr])
    # svd_system, hermitian: overlap matrix of unit vectors with `null` exact linear dependencies
    for n, null in ((6, 1), (25, 3)):
        b = rng.standard_normal((n - null, 300))
        b = np.vstack([b, rng.standard_normal((null, n - null)) @ b])
        b /= np.linalg.norm(b, axis=1)[:, None]
        m = b @ b.T
        m = 0.5 * (m + m.T)
        w, v = np.linalg.eigh(m)
        keep = [i for i in range(n - 1, -1, -1) if w[i] <= 1e-12]  # descending, values <= threshold
        out[f"svd_herm_{n}_M"], out[f"svd_herm_{n}_values"] = m.ravel(order="F"), w[keep]
        out[f"svd_herm_{n}_vectors"] = v[:, keep].T.copy().ravel()
        out[f"svd_herm_{n}_null"] = np.array([null])
    # svd_system, general rectangular (column-major input)
    for rows, cols in ((7, 4), (5, 5)):
        u, _ = np.linalg.qr(rng.standard_normal((rows, cols)))
        vt, _ = np.linalg.qr(rng.standard_normal((cols, cols)))
        sv = np.geomspace(1.0, 1e-9, cols)
        m = (u * sv) @ vt.T
        _, s_, vh = np.linalg.svd(m, full_matrices=False)
        thr = 1e-3
        keep = [i for i in range(cols - 1, -1, -1) if abs(s_[i]) < thr]  # ascending values
        out[f"svd_gen_{rows}x{cols}_M"], out[f"svd_gen_{rows}x{cols}_values"] = m.ravel(order="F"), s_[keep]
        out[f"svd_gen_{rows}x{cols}_vectors"] = vh[keep, :].copy().ravel()
        out[f"svd_gen_{rows}x{cols}_thr"] = np.array([thr])
    # solve_LinearEquations: matrix and rhs row-major; solution dimension x nroot column-major
    for n, nroot in ((1, 1), (7, 3), (60, 8)):
        a = rng.standard_normal((n, n)) + n * np.eye(n)
        a = 0.5 * (a + a.T)
        rhs = rng.standard_normal((n, nroot))
        out[f"lineq_{n}_A"], out[f"lineq_{n}_rhs"] = a.ravel(), rhs.ravel()
        out[f"lineq_{n}_x"] = np.linalg.solve(a, rhs).ravel(order="F")
        # augmented hessian (alpha = 0.7): lowest eigenpair of [[A, -alpha b], [-alpha b^T, 0]] v = e [[S, 0], [0, 1]] v
        alpha = 0.7
        s = np.eye(n) + 0.05 * (lambda t: t + t.T)(rng.standard_normal((n, n))) / n
        xs, es = [], []
        rhs_cm = rng.standard_normal((n, nroot))  # this branch reads rhs column-major (:570)
        for r in range(nroot):
            big_a = np.zeros((n + 1, n + 1))
            big_s = np.zeros((n + 1, n + 1))
            big_a[:n, :n], big_s[:n, :n] = a, s
            big_a[:n, n] = big_a[n, :n] = -alpha * rhs_cm[:, r]
            big_s[n, n] = 1
            w, v = sl.eigh(big_a, big_s)
            es.append(w[0])
            xs.append(v[:n, 0] / (alpha * v[n, 0]))
        out[f"lineq_aug_{n}_S"], out[f"lineq_aug_{n}_rhs"] = s.ravel(order="F"), rhs_cm.ravel(order="F")
        out[f"lineq_aug_{n}_x"], out[f"lineq_aug_{n}_e"] = np.concatenate(xs), np.array(es)
    # solve_DIIS: B = <r_i, r_j> of a converging residual sequence
    for n in (1, 3, 8):
        r = rng.standard_normal((n, 50)) * np.geomspace(1.0, 1e-3, n)[:, None]
        bm = r @ r.T
        aug = np.zeros((n + 1, n + 1))
        aug[:n, :n] = bm
        aug[n, :n] = aug[:n, n] = -1
        rhs = np.zeros(n + 1)
        rhs[n] = -1
        out[f"diis_{n}_B"] = bm.ravel(order="F")
        out[f"diis_{n}_c"] = (np.linalg.pinv(aug, rcond=1e-300) @ rhs)[:n]
    np.savez_compressed(os.path.join(HERE, "helper_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
