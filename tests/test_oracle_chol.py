"""Pins the C oracle of the adaptive banded Cholesky solve (src/infcholesky.jl:42-59,94-140). CPU only."""
import numpy as np
import pytest

from oracle import banded as ob


def _dense_sym(p, q, u, n, shift=0.0):
    A = np.zeros((n, n))
    for rr in range(u + 1):
        d = u - rr
        for t in range(n - d):
            A[t, t + d] = p[rr] + q[rr] * t
            A[t + d, t] = A[t, t + d]
    return A - shift * np.eye(n)


def test_tridiagonal_cholesky_matches_qr_and_finite():
    # test/test_infcholesky.jl:5-11 — cholesky(S)\b ≈ qr(S)\b for S = Symmetric(BandedMatrix(0 => 1:∞, 1 => Ones(∞)))
    oc = ob.chol_banded_solve(1, [[1.0, 1.0]], [[0.0, 1.0]], b=[1.0, 2.0, 3.0], ncap=6000, want_factors=True)
    oq = ob.qr_banded_solve(1, 1, [[1.0, 1.0, 1.0]], [[0.0, 1.0, 0.0]], b=[1.0, 2.0, 3.0], ncap=6000)
    assert oc["status"][0] == 0 and oq["status"][0] == 0
    n = int(min(oc["ncols"][0], oq["ncols"][0]))
    assert np.allclose(oc["x"][0, :n], oq["x"][0, :n], rtol=1e-12, atol=1e-15)
    # U against a dense Cholesky of the finite section (test_infcholesky.jl:59)
    N = 100
    S = _dense_sym([1.0, 1.0], [0.0, 1.0], 1, N + 5)
    Ud = np.linalg.cholesky(S).T
    fac = oc["factors"][0]
    for j in range(N):
        assert np.isclose(fac[j, 1], Ud[j, j], rtol=1e-13)            # diagonal
        if j > 0:
            assert np.isclose(fac[j, 0], Ud[j - 1, j], rtol=1e-12, atol=1e-15)


def test_pentadiagonal_c4_family():
    # BASELINE config C4 family: residual of the solve and U'U = S on the finite section
    for sigma in (0.01, 1.0, 10.0):
        p = [1.0, -4.0, 6.0 + sigma]
        q = [0.0, 0.0, 1e-4]
        o = ob.chol_banded_solve(2, [p], [q], b=[1.0], ncap=30000, want_factors=True)
        assert o["status"][0] == 0
        n = int(o["ncols"][0])
        assert o["nconv"][0] % 1000 == 1 and n == o["nconv"][0] + 1000 + 1     # n = n_conv + COLGROWTH - 1 + u
        N = 400
        S = _dense_sym(p, q, 2, N)
        x = o["x"][0, :N]
        r = S[:N - 4] @ x
        assert abs(r[0] - 1.0) < 1e-12 and np.max(np.abs(r[1:])) < 1e-12
        fac = o["factors"][0]
        Ud = np.linalg.cholesky(S).T
        assert np.allclose(fac[:N - 4, 2], np.diag(Ud)[:N - 4], rtol=1e-12)


def test_not_positive_definite_and_capacity():
    o = ob.chol_banded_solve(1, [[1.0, -1.0]], [[0.0, 0.0]], b=[1.0], ncap=5000)
    assert o["status"][0] == 4 and o["ncols"][0] == 0
    o = ob.chol_banded_solve(2, [[1.0, -4.0, 6.01]], [[0.0, 0.0, 1e-4]], b=[1.0], ncap=900)
    assert o["status"][0] == 3
