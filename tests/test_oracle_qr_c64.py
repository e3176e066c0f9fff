"""Pins the complex-eltype adaptive-QR oracle (oracle/ila_oracle_c64.c).  CPU only.  The reference has no complex
adaptive-QR test; the restatement is pinned by its real-valued limit (bit for bit against the real oracle, which the
reference's own known answers pin) and by dense solves of finite sections."""
import numpy as np
from scipy.special import jv

from oracle import banded as ob


def test_real_limit_is_bit_identical_to_the_real_oracle():
    z = np.array([300.0, 57.3, 1000.0])
    p, q, r = ob.bessel_operator(z)
    o = ob.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap=6000)
    oc = ob.qr_banded_solve_c64(1, 1, p, q, r, b=jv(1, z)[:, None], ncap=6000)
    assert np.array_equal(o["ncols"], oc["ncols"]) and np.array_equal(o["nconv"], oc["nconv"])
    assert np.all(oc["status"] == 0)
    assert np.array_equal(o["x"], oc["x"].real) and np.all(oc["x"].imag == 0)
    # 5-band operator of test/test_infqr.jl:114-119 with a head, (l,u) = (2,2)
    p5 = np.array([[1.0, 2.0, 5.0, 2.0, 1.0]]); q5 = np.array([[0.0, 0.0, 0.01, 0.0, 0.0]])
    o = ob.qr_banded_solve(2, 2, p5, q5, None, b=[1.0, 2.0, 3.0], ncap=8000)
    oc = ob.qr_banded_solve_c64(2, 2, p5, q5, None, b=[1.0, 2.0, 3.0], ncap=8000)
    assert np.array_equal(o["ncols"], oc["ncols"]) and np.array_equal(o["x"], oc["x"].real)


def test_complex_shift_against_finite_sections():
    # resolvent of the free Jacobi operator and of a Bessel-type operator at complex shifts
    for lam in (2.0 + 1.5j, -0.3 + 0.2j, 1j):
        pp = np.array([[0.5, 0.0, 0.5]]); qq = np.zeros((1, 3))
        oc = ob.qr_banded_solve_c64(1, 1, pp, qq, None, shift=[lam], b=[1.0, 0.5j], ncap=30000)
        assert oc["status"][0] == 0
        n = int(oc["ncols"][0])
        N = min(max(n, 400), 3000)
        A = np.diag(np.full(N, -lam)) + np.diag(np.full(N - 1, 0.5), 1) + np.diag(np.full(N - 1, 0.5), -1)
        rhs = np.zeros(N, complex); rhs[0] = 1; rhs[1] = 0.5j
        ref = np.linalg.solve(A, rhs)
        m = min(n, 200)
        assert np.abs(oc["x"][0][:m] - ref[:m]).max() <= 1e-13 * np.abs(ref).max()
    # complex generators: diagonal (2 + 0.5i) + (0.01 - 0.02i) k, off-diagonals 1 and 0.7i, (l,u) = (1,2)
    p = np.array([[0.3j, 1.0, 2 + 0.5j, 0.7j]]); q = np.array([[0, 0, 0.01 - 0.02j, 0]])
    oc = ob.qr_banded_solve_c64(1, 2, p, q, None, b=[1.0, -1j, 0.25], ncap=20000)
    assert oc["status"][0] == 0
    n = int(oc["ncols"][0]); N = n + 50
    A = np.zeros((N, N), complex)
    for k in range(N):
        A[k, k] = p[0, 2] + q[0, 2] * k
        if k + 1 < N: A[k, k + 1] = p[0, 1]; A[k + 1, k] = p[0, 3]
        if k + 2 < N: A[k, k + 2] = p[0, 0]
    rhs = np.zeros(N, complex); rhs[:3] = [1.0, -1j, 0.25]
    ref = np.linalg.solve(A, rhs)
    assert np.abs(oc["x"][0][:n] - ref[:n]).max() <= 1e-12 * np.abs(ref).max()
