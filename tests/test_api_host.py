"""Host-side logic of the Julia-interface mirror that needs no GPU: lazy diagonals, operator algebra, generator extraction."""
import numpy as np

import ila_b200  # noqa: F401  (alias module)
from ila_b200.api import (AlmostBanded, BandedMatrix, BoundaryRow, Eye, Fill, I, KronTrav, Poly, Rational, Tridiagonal, Vcat, Zeros,
                          _lazy_values)


def test_lazy_values():
    assert np.array_equal(_lazy_values(Fill(0.5), 4), [0.5] * 4)
    assert np.array_equal(_lazy_values(Vcat([1.0], Fill(0.5)), 4), [1.0, 0.5, 0.5, 0.5])          # Vcat(1.0, Fill(1/2,∞))
    assert np.allclose(_lazy_values(Rational(0, 1, 0, 2, 1), 3), [1 / 2, 1 / 3, 1 / 4])            # 1 ./ (2:∞)
    T = Tridiagonal(Vcat([1.0], Fill(0.5)), Zeros(), Fill(0.5))
    assert T.bands(3).tolist() == [[1.0, 0.5, 0.5], [0.0, 0.0, 0.0], [0.5, 0.5, 0.5]]


def test_krontrav_algebra():
    D = Tridiagonal(Fill(0.5), Zeros(), Fill(0.5))
    A = KronTrav(D, Eye) - 2 * I                        # test_infqr.jl:271
    B = KronTrav(Eye, D) - 2 * I                        # :287
    assert (A.alpha, A.beta, A.lam) == (1.0, 0.0, 2.0) and (B.alpha, B.beta, B.lam) == (0.0, 1.0, 2.0)
    L = A + B                                           # :292
    assert (L.alpha, L.beta, L.lam) == (1.0, 1.0, 4.0)
    M = 2 * KronTrav(D, Eye) + KronTrav(Eye, D) - 8 * I   # :309
    assert (M.alpha, M.beta, M.lam) == (2.0, 1.0, 8.0)
    # the dense section of the oracle agrees with the Kronecker definition for a small case: KronTrav(A,B)[K,J][k,j] = A[k,j] B[K-k+1,J-j+1]
    from oracle import block_qr as bq
    T = D.bands(10)
    Md = bq.dense_operator(T, T, 2.0, 1.0, 8.0, 3)
    Dd = np.diag(np.full(2, 0.5), 1) + np.diag(np.full(2, 0.5), -1)
    Id = np.eye(3)
    ref = np.zeros((6, 6))
    bs = lambda K: K * (K - 1) // 2
    for K in range(1, 4):
        for J in range(1, 4):
            for k in range(1, K + 1):
                for j in range(1, J + 1):
                    a1, b1 = Dd[k - 1, j - 1], Id[K - k, J - j]           # KronTrav(Δ, Eye)
                    a2, b2 = Id[k - 1, j - 1], Dd[K - k, J - j]           # KronTrav(Eye, Δ)
                    ref[bs(K) + k - 1, bs(J) + j - 1] = 2.0 * a1 * b1 + 1.0 * a2 * b2 - 8.0 * (K == J and k == j)
    assert np.array_equal(Md, ref)


def test_almostbanded_description():
    A = AlmostBanded([BoundaryRow(), BoundaryRow(s=-1)], {0: -1.0, 2: Poly(2, 3, 1)})
    assert [r.g for r in A.rows] == [(1.0, 1.0, 0.0, 0.0), (-1, 1.0, 0.0, 0.0)]
    from oracle import almostbanded_qr as ab
    S = ab.dense_section(0, 2, [[2, 3, 1], [0, 0, 0], [-1, 0, 0]], [list(r.g) for r in A.rows], 0.0, 5, 6)
    # Vcat(Ones(1,∞), ((-1).^(0:∞))', BandedMatrix(0 => -Ones(∞), 2 => (1:∞).*(2:∞)))[1:5, 1:6]
    ref = np.array([[1, 1, 1, 1, 1, 1], [1, -1, 1, -1, 1, -1], [-1, 0, 2, 0, 0, 0], [0, -1, 0, 6, 0, 0], [0, 0, -1, 0, 12, 0]], dtype=float)
    assert np.array_equal(S, ref)


def test_banded_generators():
    A = BandedMatrix({0: Vcat([0.5, 0.25], Fill(0.5)), 1: Zeros(), 2: Vcat([-0.3], Fill(-0.25))})
    p, q, r, head = A.generators()
    assert A.bandwidths == (0, 2) and head.shape == (3, 2)
    assert np.array_equal(head, [[-0.3, -0.25], [0.0, 0.0], [0.5, 0.25]]) and np.array_equal(p, [-0.25, 0.0, 0.5])
