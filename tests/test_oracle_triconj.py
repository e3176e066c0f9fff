"""The tridiagonal / bidiagonal conjugation oracle (oracle/ila_oracle_triconj.c) against the reference's own tests
(test/test_bidiagonalconjugation.jl:105-198): Tridiagonal(U*X) and Tridiagonal(U*X/U) from dense sections for the five
connection-matrix pairs and the Cholesky case, and Y[n,n+1] ≈ 1/2 at n = 100 000."""
import numpy as np

from oracle import banded as ob
from triconj_cases import cholesky_case, dense_conjugation, reference_pairs


def _close(a, b, tol=1e-10):
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def test_reference_pairs_against_dense_sections():
    n = 1000
    for name, U, X, Ud in reference_pairs(n + 3):
        Y, UX = ob.tridiagonal_conjugation(U, X, n, want_ux=True)
        Yd, UXd = dense_conjugation(Ud, X, n + 1)
        # rows 1..n-2 of the finite section are unaffected by the truncation
        assert _close(UX[0][:, : n - 2], UXd[:, : n - 2]), name
        assert _close(Y[0][:, : n - 2], Yd[:, : n - 2]), name


def test_chebyshev_limit_at_100000():
    # test_bidiagonalconjugation.jl:184-190: Y[n,n+1] ≈ 1/2 for T -> U
    n = 100_000
    _, U, X, _ = reference_pairs(n + 2, dense=False, first_only=True)[0]
    Y = ob.tridiagonal_conjugation(U, X, n)
    assert abs(Y[0, 2, n - 1] - 0.5) <= 1e-14 and abs(Y[0, 0, n - 1] - 0.5) <= 1e-14


def test_cholesky_case_against_dense():
    # test_bidiagonalconjugation.jl:159-178: R = cholesky(M).U adaptive, conjugated
    n = 1000
    u, p, q, head, X, M = cholesky_case(n + 8)
    cg = (n + 4) // 2
    o = ob.chol_banded_solve(u, p, q, head=head, b=[1.0], tol=np.inf, colgrowth=cg, ncap=2 * cg + 3 * u + 2, want_factors=True)
    assert o["status"][0] == 0 and o["ncols"][0] >= n + 3
    f = o["factors"][0]
    U = np.stack([f[d:n + 2 + d, u - d] for d in range(3)])          # U[k, k+d] = factors[k+d][u-d]
    Ud = np.linalg.cholesky(M).T
    assert np.abs(U[0] - np.diag(Ud)[: n + 2]).max() <= 1e-13
    Y = ob.tridiagonal_conjugation(U, X, n)
    Yd, _ = dense_conjugation(Ud, X, n + 1)
    assert _close(Y[0][:, : n - 2], Yd[:, : n - 2], 1e-9)


def test_bidiagonal_against_dense():
    # test_bidiagonalconjugation.jl:1-100 shape: B = Bidiagonal(inv(U) X V) for random well-conditioned tridiagonal-supported U, C
    rng = np.random.default_rng(7)
    n, m = 200, 203
    for uplo in ("U", "L"):
        # choose B bidiagonal and U triangular of the matching shape, so that C = U*B is tridiagonal and B is recovered
        dv = 1 + rng.random(m); ev = rng.normal(size=m) * 0.3
        ud = 1 + rng.random(m); uo = rng.normal(size=m) * 0.3
        if uplo == "U":
            B = np.diag(dv) + np.diag(ev[: m - 1], 1)
            Ud = np.diag(ud) + np.diag(uo[: m - 1], 1)
        else:
            B = np.diag(dv) + np.diag(ev[: m - 1], -1)
            Ud = np.diag(ud) + np.diag(uo[: m - 1], -1)
        Cd = Ud @ B
        tri = lambda A: np.stack([np.append(np.diag(A, -1), 0.0), np.diag(A), np.append(np.diag(A, 1), 0.0)])
        d, e = ob.bidiagonal_conjugation(tri(Ud), tri(Cd), n, uplo)
        assert np.abs(d[0] - dv[:n]).max() <= 1e-13 and np.abs(e[0] - ev[:n]).max() <= 1e-13
