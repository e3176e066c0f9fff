"""Pins the C oracle of the adaptive banded QR solve against the reference's known answers. CPU only."""
import numpy as np
import pytest
from scipy.special import jv

from oracle import banded as ob


def _dense_qr_householder(A):
    """Unblocked Householder QR with LinearAlgebra.reflector!/reflectorApply! conventions (qrunblocked)."""
    A = np.array(A, dtype=float, copy=True)
    m, n = A.shape
    tau = np.zeros(min(m, n))
    for k in range(min(m, n)):
        x = A[k:, k]
        normu = np.sqrt(np.sum(x * x))
        if normu == 0:
            continue
        nu = np.copysign(normu, x[0])
        xi1 = x[0] + nu
        x[0] = -nu
        x[1:] /= xi1
        t = xi1 / nu
        tau[k] = t
        v = np.concatenate([[1.0], x[1:]])
        for j in range(k + 1, n):
            w = t * (v @ A[k:, j])
            A[k:, j] -= v * w
    return A, tau


def _band_to_dense(fac, l, u, n):
    """(n, 2l+u+1) column-major band storage with bandwidths (l, l+u) -> dense n+l x n."""
    U = l + u
    D = np.zeros((n + l, n))
    for j in range(n):
        for i in range(max(0, j - U), min(n + l, j + l + 1)):
            D[i, j] = fac[j, U + i - j]
    return D


def test_first_entries_and_finite_section():
    # test/test_infqr.jl:12-41: bands (1, 1:∞, 1); factors[1,1] = -√2, τ[1] = 1+√2/2, factors/τ vs qrunblocked
    o = ob.qr_banded_solve(1, 1, [[1, 1, 1]], [[0, 1, 0]], b=[1., 2, 3], want_factors=True, ncap=4000)
    assert o["status"][0] == 0
    n = int(o["ncols"][0])
    fac = o["factors"][0, :n]
    assert np.isclose(fac[0, 2], -np.sqrt(2), rtol=1e-15)
    assert np.isclose(o["tau"][0, 0], 1 + np.sqrt(2) / 2, rtol=1e-15)
    N = 101
    A = np.diag(np.arange(1.0, N + 2)) + np.diag(np.ones(N), 1) + np.diag(np.ones(N), -1)
    F, tau = _dense_qr_householder(A[:N + 1, :N])
    D = _band_to_dense(fac, 1, 1, 120)
    assert np.allclose(D[:N + 1, :N - 2], F[:, :N - 2], rtol=1e-13, atol=1e-15)   # R and reflectors
    assert np.allclose(o["tau"][0, :N - 2], tau[:N - 2], rtol=1e-13)
    assert np.isclose(D[99, 99], F[99, 99], rtol=1e-13)


def test_residual_and_support():
    # test/test_infqr.jl:87: (A*(A\b))[1:100] ≈ [1,2,3,0...]
    o = ob.qr_banded_solve(1, 1, [[1, 1, 1]], [[0, 1, 0]], b=[1., 2, 3], ncap=4000)
    n = int(o["ncols"][0])
    assert n == 2001 and o["nconv"][0] == 1001            # converges at the first check, datasize = 1001+999+1
    x = o["x"][0]
    N = 200
    A = np.diag(np.arange(1.0, N + 1)) + np.diag(np.ones(N - 1), 1) + np.diag(np.ones(N - 1), -1)
    r = A[:100] @ x[:N]
    assert np.allclose(r, [1, 2, 3] + [0] * 97, atol=1e-14)


@pytest.mark.parametrize("z,n_expect,nconv_expect,tol", [(1000.0, 3001, 2001, 1e-12), (10000.0, 13001, 12001, 2e-11)])
def test_bessel_closed_form(z, n_expect, nconv_expect, tol):
    # test/test_infqr.jl:93-105 and README.md:41-66; conditioning-limited: error ∝ z·eps
    p, q, r = ob.bessel_operator(z)
    o = ob.qr_banded_solve(1, 1, p, q, r, b=[jv(1, z)], ncap=int(1.06 * z + 4000))
    assert o["status"][0] == 0
    assert o["ncols"][0] == n_expect and o["nconv"][0] == nconv_expect     # SURVEY A.1 [probe]
    x = o["x"][0]
    k = np.arange(int(2 * z))
    ref = jv(k, z)
    m = min(len(ref), n_expect)
    assert np.max(np.abs(x[:m] - ref[:m])) <= tol * np.max(np.abs(ref))
    if z == 10000.0:
        assert abs(x[999] - jv(999, z)) < 5e-14          # README.md:62-63 reports -6.8e-15
        assert abs(x[10999] - jv(10999, z)) < 1e-142     # README.md:65-66 reports 3.4e-143 (high relative accuracy)


def test_quad_oracle_brackets_double():
    z = 5000.0
    p, q, r = ob.bessel_operator(z)
    o = ob.qr_banded_solve(1, 1, p, q, r, b=[jv(1, z)], ncap=12000)
    oq = ob.qr_banded_solve(1, 1, p, q, r, b=[jv(1, z)], ncap=12000, quad=True)
    n = int(o["ncols"][0])
    assert n == oq["ncols"][0]
    err = np.max(np.abs(o["x"][0, :n] - oq["x"][0, :n])) / np.max(np.abs(oq["x"][0, :n]))
    assert err < 2 * z * np.finfo(float).eps             # the κ-aware bound of BASELINE.md


def test_five_band_against_finite_solve():
    # test/test_infqr.jl:114-119
    head = np.zeros((5, 3))
    head[0, :] = 1
    head[1, :] = [1, 0, 0]
    head[2, :] = [4, 5, 6]
    head[3, :] = [1, 0, 0]
    head[4, :] = 1
    o = ob.qr_banded_solve(2, 2, [[1, 0, 3, 0, 1]], [[0] * 5], head=head, b=[3., 4, 5], ncap=8000)
    assert o["status"][0] == 0
    N = 2000
    A = (np.diag(np.full(N, 3.0)) + np.diag(np.ones(N - 2), 2) + np.diag(np.ones(N - 2), -2))
    A[0, 0], A[1, 1], A[2, 2] = 4, 5, 6
    A[0, 1] = A[1, 0] = 1
    rhs = np.zeros(N)
    rhs[:3] = [3, 4, 5]
    xs = np.linalg.solve(A, rhs)
    assert np.allclose(o["x"][0, :N], xs, rtol=1e-10, atol=1e-14)


def test_tolerance_keyword_and_ncap():
    # test/test_infqr.jl:330-334 (tolerance=1E-8 stops earlier); ncap too small -> status 3
    z = 1000.0
    p, q, r = ob.bessel_operator(z)
    o1 = ob.qr_banded_solve(1, 1, p, q, r, b=[jv(1, z)], ncap=8000)
    o2 = ob.qr_banded_solve(1, 1, p, q, r, b=[jv(1, z)], ncap=8000, tol=1e-8)
    assert o2["nconv"][0] <= o1["nconv"][0]
    assert np.allclose(o1["x"][0, :1000], o2["x"][0, :1000], atol=1e-8)
    o3 = ob.qr_banded_solve(1, 1, p, q, r, b=[jv(1, z)], ncap=1500)
    assert o3["status"][0] == 3 and o3["ncols"][0] == 0
