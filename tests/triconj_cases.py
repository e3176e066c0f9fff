"""Fixtures of test/test_bidiagonalconjugation.jl:105-178 (the `TridiagonalConjugation` test set) as band arrays,
shared by the oracle tests and the GPU parity tests.  U[d, k] = U[k+1, k+1+d]; X rows = (dl, d, du)."""
import numpy as np


def _dense_upper(bands, N):
    return sum(np.diag(b[: N - d], d) for d, b in enumerate(bands))


def _dense_tri(X, N):
    return np.diag(X[1, :N]) + np.diag(X[0, : N - 1], -1) + np.diag(X[2, : N - 1], 1)


def reference_pairs(m, dense=True, first_only=False):
    """The five (R, X) pairs of test_bidiagonalconjugation.jl:106-128.  Returns a list of (name, U bands (nbands x m),
    X (3 x m), dense U (m x m) including the bands beyond +2)."""
    k = np.arange(1, m + 1, dtype=np.float64)          # 1-based index
    out = []
    # T -> U: R = _BandedMatrix([-1/2...; 0...; 1, 1/2...], 0, 2), X = Tridiagonal([1, 1/2...], 0, 1/2...)
    R0 = np.full(m, 0.5); R0[0] = 1.0
    U = np.stack([R0, np.zeros(m), np.full(m, -0.5)])
    dl = np.full(m, 0.5); dl[0] = 1.0
    X = np.stack([dl, np.zeros(m), np.full(m, 0.5)])
    out.append(("T->U", U, X, _dense_upper(U, m) if dense else None))
    if first_only:
        return out
    # P -> C^(3/2): R[j-2,j] = -1/(2j-1), R[j,j] = 1/(2j-1); X = Tridiagonal(k/(2k-1), 0, k/(2k+1))
    U = np.stack([1 / (2 * k - 1), np.zeros(m), -1 / (2 * k + 3)])
    Xleg = np.stack([k / (2 * k - 1), np.zeros(m), k / (2 * k + 1)])
    out.append(("P->C3/2", U, Xleg, _dense_upper(U, m)))
    # P^(1,0) -> P^(2,0): R[j-1,j] = -(j-1)/(2j), R[j,j] = (j+1)/(2j), with and without the zero third band
    U3 = np.stack([(k + 1) / (2 * k), -k / (2 * (k + 1)), np.zeros(m)])
    X = np.stack([(k + 1) / (2 * k + 1), -1 / ((2 * k - 1) * (2 * k + 1)), k / (2 * k + 1)])
    out.append(("P10->P20 (0,2)", U3, X, _dense_upper(U3, m)))
    out.append(("P10->P20 (0,1)", U3[:2].copy(), X, _dense_upper(U3, m)))
    # P -> C^(5/2): product of two (0,2) banded matrices (bandwidth (0,4)); only its bands 0..2 are read
    kk = np.arange(1, m + 5, dtype=np.float64)
    Ra = _dense_upper(np.stack([3 / (2 * kk + 1), np.zeros(m + 4), -3 / (2 * kk + 5)]), m + 4)
    Rb = _dense_upper(np.stack([1 / (2 * kk - 1), np.zeros(m + 4), -1 / (2 * kk + 3)]), m + 4)
    Rab = (Ra @ Rb)[:m + 2, :m + 2]
    U = np.stack([np.diag(Rab)[:m], np.diag(Rab, 1)[:m], np.diag(Rab, 2)[:m]])
    out.append(("P->C5/2", U, Xleg, Rab[:m, :m]))
    return out


def dense_conjugation(Ud, X, n):
    """Tridiagonal(U*X/U) and Tridiagonal(U*X) from dense N x N sections (test_bidiagonalconjugation.jl:129-136)."""
    N = Ud.shape[0]
    Xd = _dense_tri(X, N)
    UX = Ud @ Xd
    Y = np.linalg.solve(Ud.T, UX.T).T
    pick = lambda M: np.stack([np.diag(M, -1)[:n], np.diag(M)[:n], np.diag(M, 1)[:n]])
    return pick(Y), pick(UX)


def cholesky_case(m):
    """test_bidiagonalconjugation.jl:160-164: M = Symmetric(_BandedMatrix([-1/(2√2) x3, -1/4...; 0...; 0.5, 0.25, 0.5...], 0, 2)),
    X = Tridiagonal([1/√2, 1/2...], 0, [1/√2, 1/2...]).  Returns the generator description (u, p, q, head), X and dense M."""
    s = 1 / (2 * np.sqrt(2.0))
    head = np.array([[-s, -0.25], [0.0, 0.0], [0.5, 0.25]])
    p = np.array([[-0.25, 0.0, 0.5]]); q = np.zeros((1, 3))
    e = np.full(m, 0.5); e[0] = 1 / np.sqrt(2.0)
    X = np.stack([e, np.zeros(m), e])
    d0 = np.full(m, 0.5); d0[1] = 0.25
    d2 = np.full(m, -0.25); d2[0] = -s
    M = np.diag(d0) + np.diag(d2[: m - 2], 2) + np.diag(d2[: m - 2], -2)
    return 2, p, q, head, X, M
