"""Parity of the CUDA tridiagonal / bidiagonal conjugation (K10) against the C oracle, through the C ABI.
Bar: bit-exact (the kernel's per-row closed form issues the reference's operations in the reference's order)."""
import numpy as np
import pytest

from triconj_cases import cholesky_case, dense_conjugation, reference_pairs

pytestmark = pytest.mark.gpu


def test_reference_pairs_bit_exact(ila):
    from oracle import banded as ob
    n = 1000
    for name, U, X, Ud in reference_pairs(n + 3):
        Yo, UXo = ob.tridiagonal_conjugation(U, X, n, want_ux=True)
        Yg, UXg = ila.tridiagonal_conjugation(U, X, n, want_ux=True)
        assert np.array_equal(Yg, Yo) and np.array_equal(UXg, UXo), name
        Yd, _ = dense_conjugation(Ud, X, n + 1)
        assert np.abs(Yg[0][:, : n - 2] - Yd[:, : n - 2]).max() <= 1e-10 * max(1.0, np.abs(Yd).max()), name


def test_batch_of_scaled_operators_bit_exact_and_large_n(ila):
    from oracle import banded as ob
    rng = np.random.default_rng(11)
    n, batch = 5000, 37
    m = n + 1
    U = rng.normal(size=(batch, 3, m)) * 0.3
    U[:, 0] = 1 + rng.random((batch, m))
    X = rng.normal(size=(batch, 3, m))
    Yo = ob.tridiagonal_conjugation(U, X, n)
    Yg = ila.tridiagonal_conjugation(U, X, n)
    assert np.array_equal(Yg, Yo)
    # shared X, two bands
    Yo = ob.tridiagonal_conjugation(U[:, :2].copy(), X[0], n)
    Yg = ila.tridiagonal_conjugation(U[:, :2].copy(), X[0], n)
    assert np.array_equal(Yg, Yo)
    # n = 100 000: Y[n,n+1] ≈ 1/2 for T -> U (test_bidiagonalconjugation.jl:184-190)
    n = 100_000
    _, U, X, _ = reference_pairs(n + 2, dense=False, first_only=True)[0]
    Yg = ila.tridiagonal_conjugation(U, X, n)
    assert np.array_equal(Yg, ob.tridiagonal_conjugation(U, X, n)) and abs(Yg[0, 2, n - 1] - 0.5) <= 1e-14


def test_cholesky_fused_bit_exact(ila):
    # test_bidiagonalconjugation.jl:159-178 with the factor staying on the device (K5 records -> K10)
    from oracle import banded as ob
    n = 1000
    u, p, q, head, X, M = cholesky_case(n + 8)
    g = ila.chol_tridiagonal_conjugation(u, p, q, X, n, head=head)
    assert g["status"][0] == 0
    cg = (n + 4) // 2
    o = ob.chol_banded_solve(u, p, q, head=head, b=[1.0], tol=np.inf, colgrowth=cg, ncap=2 * cg + 3 * u + 2, want_factors=True)
    f = o["factors"][0]
    U = np.stack([f[d:n + 2 + d, u - d] for d in range(3)])
    assert np.array_equal(g["Y"][0], ob.tridiagonal_conjugation(U, X, n)[0])
    Yd, _ = dense_conjugation(np.linalg.cholesky(M).T, X, n + 1)
    assert np.abs(g["Y"][0][:, : n - 2] - Yd[:, : n - 2]).max() <= 1e-9


def test_cholesky_fused_shift_family(ila):
    # Christoffel-type modification: S_σ = σ I - J (J the Legendre Jacobi matrix in its orthonormal form), 70 shifts, n = 777
    from oracle import banded as ob
    n, m = 777, 790
    k = np.arange(1, m + 1, dtype=np.float64)
    e = k / np.sqrt(4 * k * k - 1)
    X = np.stack([e, np.zeros(m), e])
    sigma = np.linspace(1.05, 3.0, 70)
    # upper bands of σI - J: +1 -> -e_k (explicit head covers all n+4 rows used), 0 -> σ
    hp = n + 6
    head = np.stack([-e[:hp], np.zeros(hp)])
    p = np.array([[0.0, 0.0]]); q = np.zeros((1, 2))
    g = ila.chol_tridiagonal_conjugation(1, p, q, X, n, shift=-sigma, head=head)
    assert np.all(g["status"] == 0)
    cg = (n + 4) // 2
    o = ob.chol_banded_solve(1, np.repeat(p, 70, 0), np.repeat(q, 70, 0), shift=-sigma, head=head, b=[1.0], tol=np.inf, colgrowth=cg,
                             ncap=2 * cg + 3 + 2, want_factors=True)
    for i in range(70):
        f = o["factors"][i]
        U = np.stack([f[d:n + 2 + d, 1 - d] for d in range(2)])
        assert np.array_equal(g["Y"][i], ob.tridiagonal_conjugation(U, X, n)[0]), i
    # symmetric result: Y is the Jacobi matrix of the modified weight
    assert np.abs(g["Y"][:, 0] - g["Y"][:, 2]).max() <= 1e-10
    # a non-positive-definite shift reports status 4 and NaN rows
    g = ila.chol_tridiagonal_conjugation(1, p, q, X, n, shift=-np.array([0.5, 2.0]), head=head)
    assert list(g["status"]) == [4, 0] and np.all(np.isnan(g["Y"][0])) and np.all(np.isfinite(g["Y"][1]))


def test_bidiagonal_bit_exact(ila):
    from oracle import banded as ob
    rng = np.random.default_rng(5)
    n, m, batch = 3000, 3001, 9
    U = rng.normal(size=(batch, 3, m)) * 0.2
    U[:, 1] = 1 + rng.random((batch, m))
    Cm = rng.normal(size=(batch, 3, m))
    for uplo in ("U", "L"):
        do, eo = ob.bidiagonal_conjugation(U, Cm, n, uplo)
        dg, eg = ila.bidiagonal_conjugation(U, Cm, n, uplo)
        assert np.array_equal(dg, do) and np.array_equal(eg, eo), uplo
