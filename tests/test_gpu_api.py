"""The host-side mirror of the reference's Julia interface, exercised the way the reference's own tests read."""
import numpy as np
import pytest
from scipy.special import jv

pytestmark = pytest.mark.gpu


def test_bullhead_readme(ila):
    # README.md:83-85, examples/toeplitz.jl:14-20,71-72
    from ila_b200.api import BandedMatrix, Fill, I, ql, ql_L11, DomainError
    A = BandedMatrix({-3: Fill(7 / 10), -2: Fill(1), 1: Fill(2j)})
    F = ql(A - 5 * I)
    assert F.L[1, 1] == pytest.approx(-4.747700587226856, rel=1e-13)
    with pytest.raises(DomainError):
        ql(A - (0.1 + 0.1j) * I)
    lam = np.array([5, -3 - 0.1j, 2 + 3j])
    L, st = ql_L11(A, lam)
    assert np.all(st == 0) and L[0] == F.L[1, 1]
    # Q[1:10,1:11]*L[1:11,1:10] ≈ A[1:10,1:10]   (test/test_infql.jl:41-67)
    T = (A - 5 * I).finite(10, 10)
    assert np.allclose(F.Q[1:10, 1:11] @ F.L[1:11, 1:10], T, atol=1e-12)


def test_tridiagonal_toeplitz_cases(ila):
    # test/test_infql.jl:12-38
    from ila_b200.api import BandedMatrix, Fill, ql, DomainError
    for (Z, A_, B) in ((2.0, 5.1, 0.5), (2.0, -2.2, 0.5), (0.5, -5.1, 2.0)):
        T = BandedMatrix({-1: Fill(Z), 0: Fill(A_), 1: Fill(B)})
        F = ql(T)
        assert np.allclose(F.Q[1:10, 1:11] @ F.L[1:11, 1:10], T.finite(10), atol=1e-12)
    for (Z, A_, B) in ((0.5, 2.1, 2.0), (0.5, -1.1, 2.0)):
        with pytest.raises(DomainError):
            ql(BandedMatrix({-1: Fill(Z), 0: Fill(A_), 1: Fill(B)}))
    T = BandedMatrix({-1: Fill(2 + 0j), 0: Fill(2.1 + 0.1j), 1: Fill(0.5 + 0j)})
    F = ql(T)
    assert np.allclose(F.Q[1:10, 1:12] @ F.L[1:12, 1:10], T.finite(10), atol=1e-12)


def test_pert_toeplitz_jacobi(ila):
    # test/test_infql.jl:147-153
    from ila_b200.api import BandedMatrix, Fill, Vcat, I, ql
    J = BandedMatrix({0: Vcat([1.0], Fill(0.0)), 1: Fill(0.5), -1: Fill(0.5)})
    F = ql(J - 3.5 * I)
    assert F.L[1, 1] == pytest.approx(-2.4010806467474954, rel=1e-12)


def test_bessel_readme(ila):
    # README.md:41-66, test/test_infqr.jl:93-105
    from ila_b200.api import BandedMatrix, Affine, Ones, Vcat, Zeros, qr, ldiv
    z = 1000.0
    A = BandedMatrix({0: Affine(0.0, -2.0, z), 1: Ones(), -1: Ones()})
    Jv = qr(A).solve(Vcat([jv(1, z)], Zeros()))
    assert np.allclose(Jv[:2000], jv(np.arange(2000), z), atol=1e-12)
    assert np.array_equal(ldiv(A, [jv(1, z)]), Jv)
    # test/test_infqr.jl:70-90: bands (1, 1:∞, 1), b = [1,2,3,0,...]
    A = BandedMatrix({0: Affine(1.0, 1.0), 1: Ones(), -1: Ones()})
    x = ldiv(A, [1.0, 2.0, 3.0])
    r = A.finite(100, 200) @ x[:200]
    assert np.allclose(r, [1, 2, 3] + [0] * 97, atol=1e-14)


def test_cholesky_matches_qr(ila):
    # test/test_infcholesky.jl:5-11: S = Symmetric(BandedMatrix(0 => 1:∞, 1 => Ones(∞)))
    from ila_b200.api import BandedMatrix, Affine, Ones, Symmetric, cholesky, qr
    S = BandedMatrix({0: Affine(1.0, 1.0), 1: Ones(), -1: Ones()})
    b = [1.0, 2.0, 3.0]
    xc = cholesky(Symmetric(S)).solve(b)
    xq = qr(S).solve(b)
    m = min(len(xc), len(xq))
    assert np.allclose(xc[:m], xq[:m], rtol=1e-12, atol=1e-15)


def test_periodic_schrodinger(ila):
    # examples/periodicschrodinger.jl:8-27, test/test_infql.jl:278-287
    from ila_b200.api import BlockTridiagonal, I, ql, ql_L11, NotConvergedError
    c, a, b = [[0., 1.], [0., 0.]], [[-1., 1.], [1., 1.]], [[0., 0.], [1., 0.]]
    A = BlockTridiagonal(([c], c), ([a], a), ([b], b))
    A[1, 1] = 2
    l11 = ql(A - (-0.95) * I).L[1, 1]
    assert isinstance(l11, float) and l11 == pytest.approx(1.8965599461750668, rel=1e-12)
    xx = np.concatenate([np.arange(-0.95, 0.951, 0.05), np.arange(2.25, 4.0001, 0.125 / 4)])
    L, st = ql_L11(A, xx)
    assert np.all(st == 0) and L[0] == l11
    # complex non-selfadjoint case of test/test_periodic.jl:20-26: A - 5im*I
    c2, a2, b2 = [[0, 0.5], [0, 0]], [[0, 2.0], [0.5, 0]], [[0, 0.0], [2.0, 0]]
    Ac = BlockTridiagonal(([c2], c2), ([a2], a2), ([b2], b2)) - 5j * I
    lc = ql(Ac).L[1, 1]
    assert isinstance(lc, complex) and lc == pytest.approx(4.79196327384051, rel=1e-12)
