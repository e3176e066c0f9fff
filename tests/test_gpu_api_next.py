"""The Julia-interface mirror for the §8f rows built last (tridiagonal conjugation, block-banded QR), reading like the
reference's tests (test/test_bidiagonalconjugation.jl:105-198, test/test_infqr.jl:270-313)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_tridiagonal_conjugation_mirror(ila):
    from ila_b200.api import BandedMatrix, Fill, Symmetric, Tridiagonal, TridiagonalConjugation, Vcat, Zeros, cholesky
    # T -> U (test_bidiagonalconjugation.jl:108-109, 184-190)
    R = BandedMatrix({0: Vcat([1.0], Fill(0.5)), 1: Zeros(), 2: Fill(-0.5)})
    X = Tridiagonal(Vcat([1.0], Fill(0.5)), Zeros(), Fill(0.5))
    Y = TridiagonalConjugation(R, X)
    n = 100_000
    assert abs(Y[n, n + 1] - 0.5) <= 1e-14 and abs(Y[1, 1]) <= 1e-15 and Y[1, 5] == 0.0
    # Cholesky (test_bidiagonalconjugation.jl:159-178): Y = Tridiagonal(U X inv(U)), U = cholesky(M).U
    s = 1 / (2 * np.sqrt(2.0))
    M = Symmetric(BandedMatrix({0: Vcat([0.5, 0.25], Fill(0.5)), 1: Zeros(), 2: Vcat([-s], Fill(-0.25))}))
    e = Vcat([1 / np.sqrt(2.0)], Fill(0.5))
    Y = TridiagonalConjugation(cholesky(M), Tridiagonal(e, Zeros(), e))
    N = 400
    Md = np.zeros((N, N))
    Md[np.arange(N), np.arange(N)] = 0.5; Md[1, 1] = 0.25
    Md[np.arange(N - 2), np.arange(2, N)] = -0.25; Md[0, 2] = -s
    Md = np.triu(Md) + np.triu(Md, 1).T
    Ud = np.linalg.cholesky(Md).T
    Xd = np.diag(np.zeros(N)) + np.diag(np.r_[1 / np.sqrt(2.0), np.full(N - 2, 0.5)], 1)
    Xd = Xd + Xd.T
    Yd = np.linalg.solve(Ud.T, (Ud @ Xd).T).T
    B = Y.bands(300)
    assert np.abs(B[1] - np.diag(Yd)[:300]).max() <= 1e-10 and np.abs(B[2] - np.diag(Yd, 1)[:300]).max() <= 1e-10
    assert np.abs(B[0] - np.diag(Yd, -1)[:300]).max() <= 1e-10


def test_blockbanded_mirror(ila):
    # test_infqr.jl:270-312
    from ila_b200.api import Eye, Fill, I, KronTrav, Tridiagonal, Zeros
    D = Tridiagonal(Fill(0.5), Zeros(), Fill(0.5))          # Δ = BandedMatrix(1 => Ones/2, -1 => Ones/2)
    bs = lambda K: K * (K - 1) // 2
    x, y = 0.1, 0.2
    th, ph = np.arccos(x), np.arccos(y)
    s50 = np.sin(np.arange(1, 51) * th) / np.sin(th)
    A = KronTrav(D, Eye) - 2 * I
    u = A.solve([1.0])
    assert abs(u[[bs(K) + K - 1 for K in range(1, 51)]] @ s50 - 1 / (x - 2)) <= 1e-14
    B = KronTrav(Eye, D) - 2 * I
    u = B.solve([1.0])
    assert abs(u[[bs(K) for K in range(1, 51)]] @ s50 - 1 / (x - 2)) <= 1e-14

    def gen2d(u, nb):
        S = 0.0
        for K in range(1, nb + 1):
            k = np.arange(1, K + 1)
            S += np.sum(u[bs(K):bs(K + 1)] * np.sin(k * th) / np.sin(th) * np.sin((K - k + 1) * ph) / np.sin(ph))
        return S
    u = (A + B).solve([1.0])
    assert abs(gen2d(u, 50) - 1 / (x + y - 4)) <= 1e-14
    u = (2 * KronTrav(D, Eye) + KronTrav(Eye, D) - 8 * I).solve([1.0])
    assert abs(gen2d(u, 40) - 1 / (2 * x + y - 8)) <= 1e-14


def test_almostbanded_mirror(ila):
    # test_infqr.jl:194-221, 223-248
    from scipy.special import gammaln
    from ila_b200.api import AlmostBanded, BoundaryRow, Poly
    ref = np.exp(-gammaln(np.arange(1000) + 1.0))
    A = AlmostBanded([BoundaryRow()], {0: -1.0, 1: Poly(1, 1)})                     # Vcat(Ones(1,∞), BandedMatrix(0 => -Ones, 1 => 1:∞))
    assert np.abs(A.solve([np.e])[:1000] - ref).max() <= 1e-15
    A = AlmostBanded([BoundaryRow(), BoundaryRow(s=-1)], {0: -1.0, 2: Poly(2, 3, 1)})  # two-bands, B = BandedMatrix(0 => -Ones, 2 => k(k+1))
    assert np.abs(A.solve([np.e, 1 / np.e])[:1000] - ref).max() <= 1e-15
    x = 0.1
    assert abs(A.solve([1.0])[:1000] @ x ** np.arange(1000) - (np.exp(1 - x) * (-1 + np.exp(2 + 2 * x))) / (-1 + np.exp(4))) <= 1e-15
