"""Pins the oracle of `ql(A - λI) \\ b` for Toeplitz operators (src/infql.jl:266-289, hessenbergq.jl:125-135). CPU only."""
import numpy as np

from oracle import toeplitz_ql as tq
from oracle import banded


def _finite(col, lam, N):
    A = np.zeros((N, N), dtype=np.result_type(col.dtype, type(lam), np.float64))
    u = 1
    for rr, v in enumerate(col):
        d = u - rr
        A += np.diag(np.full(N - abs(d), v), d)
    return A - lam * np.eye(N)


def test_solve_with_ql_matches_qr():
    # test/test_infql.jl:143-145: (qr(A) \ e1)[1:1000] ≈ (ql(A) \ e1)[1:1000], A = (-1 => 2, 0 => 5, 1 => 0.5)
    col = np.array([0.5, 5.0, 2.0])
    x, st = tq.ql_toeplitz_solve(col, np.array([0.0]), np.array([1.0]), 1000)
    assert st[0] == 0
    # adaptive-QR oracle on the same operator: data rows [super, diag, sub] constant
    p = np.array([[0.5, 5.0, 2.0]]); q = np.zeros((1, 3))
    out = banded.qr_banded_solve(1, 1, p, q, None, b=np.array([[1.0]]), ncap=20000)
    n = min(1000, int(out["ncols"][0]))
    xq = np.zeros(1000); xq[:n] = out["x"][0][:n]
    assert np.allclose(x[0], xq, rtol=0, atol=1e-15)
    ref = np.linalg.solve(_finite(col, 0.0, 400), np.eye(400)[:, 0])
    assert np.allclose(x[0][:100], ref[:100], rtol=0, atol=1e-15)


def test_complex_hessenberg_vs_finite_section():
    col = np.array([2j, 0, 0, 1, 0.7])           # Bull-head symbol (README.md:83-85)
    lam = np.array([5.0 + 1j, -3 - 0.1j, 2 + 3j])
    b = np.array([1.0, 0.5j, -0.25])
    x, st = tq.ql_toeplitz_solve(col, lam, b, 40)
    assert np.all(st == 0)
    for i, l in enumerate(lam):
        rhs = np.zeros(600, complex); rhs[:3] = b
        ref = np.linalg.solve(_finite(col, l, 600), rhs)
        assert np.abs(x[i] - ref[:40]).max() <= 1e-13
    # outside the domain of existence: NaN row, status 1
    x, st = tq.ql_toeplitz_solve(col, np.array([0.1 + 0.1j]), b, 5)
    assert st[0] == 1 and np.all(np.isnan(x[0]))


def test_perturbed_jacobi_golden_qtb_and_residual():
    # test/test_infql.jl:147-160: J = (0 => [1, 0, ...], ±1 => 0.5), A = J - 3.5I, b = e1:
    #   (Q'b)[1] ≈ 0.9892996329463546 ;  u = F \ b ;  (A*u)[1:10] ≈ [1; zeros(9)]
    z = 3.5
    head = np.array([[1.0 - z, 0.5], [0.5, -z]])
    a_rev = np.array([0.5, -z, 0.5])
    x, st, qtb = tq.ql_pert_toeplitz_solve(head, a_rev, np.array([1.0]), 40)
    assert st == 0
    assert np.isclose(qtb[0], 0.9892996329463546, rtol=1e-13, atol=0)
    N = 40
    A = np.diag(np.full(N, -z)) + np.diag(np.full(N - 1, 0.5), 1) + np.diag(np.full(N - 1, 0.5), -1)
    A[0, 0] = 1 - z
    r = (A @ x)[:10]
    assert np.allclose(r, np.eye(10)[0], rtol=0, atol=1e-15)


def test_perturbed_hessenberg_vs_finite_section():
    rng = np.random.default_rng(3)
    a = np.array([0.7, 1, 0, -(5.0 + 1j), 2j])         # shifted Bull-head symbol, reversed order [a_-3 .. a_+1]
    m, l, p = 5, 3, 6
    N = 500
    A = np.zeros((N, N), complex)
    for i in range(N):
        for j in range(max(0, i - l), min(N, i + 2)):
            A[i, j] = tq._toeplitz_entry(a, i + 1, j + 1)
    for i in range(p - 1):
        for j in range(max(0, i - l), min(p - 1, i + 2)):
            A[i, j] += 0.3 * (rng.normal() + 1j * rng.normal())
    head = A[:p, :p].copy()
    b = rng.normal(size=8) + 1j * rng.normal(size=8)    # right-hand side longer than the head
    x, st, _ = tq.ql_pert_toeplitz_solve(head, a, b, 40)
    rhs = np.zeros(N, complex); rhs[:8] = b
    ref = np.linalg.solve(A, rhs)
    assert st == 0 and np.abs(x - ref[:40]).max() <= 1e-13 * np.abs(ref).max()
