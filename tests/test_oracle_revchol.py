"""Pins the reverse-Cholesky oracle (oracle/ila_oracle_revchol.c) with the reference's own tests
(test/test_infreversecholesky.jl).  CPU only."""
import numpy as np

from oracle import banded as ob

FILL = lambda v: [v, 0, 0, 1, 0]                 # value(k) = v
RECIP = lambda c0: [c0, 1, 0, 2, 1]              # value(k) = c0 + 1/(k+2)   (`1 ./ (2:∞)`)


def _residual(la, lb, a, b):
    # (L'L)[i,i] = la_i^2 + lb_i^2,  (L'L)[i,i+1] = lb_i la_{i+1}
    return max(np.abs(la ** 2 + lb ** 2 - a).max(), np.abs(lb[:-1] * la[1:] - b[:-1]).max())


def test_toeplitz_closed_form_and_residual():
    # :18-24  A = SymTridiagonal(Fill(3,∞), Fill(1,∞)); the Toeplitz layout's closed form (infreversecholeskytoeplitz.jl:
    # α² = (a + sqrt(a² - 4b²))/2, β = b/α) gives α = golden ratio; (L'L)[1:1000,1:1000] ≈ A
    o = ob.revchol_tridiag(FILL(3), FILL(1), 1000)
    phi = (1 + 5 ** 0.5) / 2
    assert o["status"][0] == 0 and o["N"][0] == 4000
    assert np.allclose(o["la"][0], phi, rtol=0, atol=4e-16) and np.allclose(o["lb"][0], 1 / phi, rtol=0, atol=4e-16)
    assert _residual(o["la"][0], o["lb"][0], np.full(1000, 3.0), np.ones(1000)) <= 1e-15


def test_another_example_fixtures():
    # :53-59  SymTridiagonal(Ones(∞), 1 ./ (2:∞)) and 5I + SymTridiagonal(1 ./ (2:∞), Ones(∞))
    k = np.arange(1000)
    o = ob.revchol_tridiag(FILL(1), RECIP(0), 1000)
    assert o["status"][0] == 0
    assert _residual(o["la"][0], o["lb"][0], np.ones(1000), 1 / (k + 2.0)) <= 1e-15
    o = ob.revchol_tridiag(RECIP(5), FILL(1), 1000)
    assert o["status"][0] == 0
    assert _residual(o["la"][0], o["lb"][0], 5 + 1 / (k + 2.0), np.ones(1000)) <= 4e-15


def test_perturbed_toeplitz_against_finite_section():
    # :10-13  A = SymTridiagonal([[4,5,6]; Fill(3,∞)], [[2,3]; Fill(1,∞)]): U[1:100,1:100] ≈ reversecholesky(A[1:1000,1:1000])
    o = ob.revchol_tridiag(FILL(3), FILL(1), 100, a_head=[4.0, 5.0, 6.0], b_head=[2.0, 3.0, 1.0])
    N = 1000
    a = np.full(N, 3.0); a[:3] = [4, 5, 6]
    b = np.ones(N); b[:2] = [2, 3]
    la = np.zeros(N); lb = np.zeros(N)
    la[N - 1] = np.sqrt(a[N - 1])
    for i in range(N - 2, -1, -1):
        lb[i] = b[i] / la[i + 1]
        la[i] = np.sqrt(a[i] - lb[i] ** 2)
    assert o["status"][0] == 0
    assert np.allclose(o["la"][0], la[:100], rtol=0, atol=1e-15) and np.allclose(o["lb"][0], lb[:100], rtol=0, atol=1e-15)


def test_status_codes():
    # non-positive pivot (sqrt of a negative number: DomainError upstream) -> 4; slow decay -> N grows; no decay -> 3
    o = ob.revchol_tridiag(FILL(1), FILL(1), 8)
    assert o["status"][0] == 4 and np.all(np.isnan(o["la"][0]))
    o = ob.revchol_tridiag(FILL(2.0005), FILL(1), 16)
    assert o["status"][0] == 0 and o["N"][0] > 64
    o = ob.revchol_tridiag(FILL(2.0), FILL(1), 4)              # a = 2|b|: ν does not decay, never converges
    assert o["status"][0] == 3 and o["N"][0] > (1 << 21) - 1
