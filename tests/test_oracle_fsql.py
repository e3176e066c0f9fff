"""Pins the finite-section adaptive QL oracle (oracle/ila_oracle_fsql.c) with the reference's own tests
(test/test_infql.jl:208-271, "Adaptive finite-section-based QL").  CPU only."""
import numpy as np

from oracle import banded as ob
from oracle import toeplitz_ql as tq

C = lambda v: [v, 0, 0, 1, 0]          # value(t) = v


def _dense_L(Lband, n):
    nd = Lband.shape[1]
    L = np.zeros((n, n))
    for i in range(n):
        for c in range(nd):
            if i - c >= 0:
                L[i, i - c] = Lband[i, c]
    return L


def test_basic_properties_golden_vector():
    # :209-222  A = _BandedMatrix(Vcat(2*Ones(1,∞), ((1 ./(1:∞)).+1/4)', Ones(1,∞)./3), ℵ₀, 1, 1); b = [1,2,3,0,...]
    #   (L*b)[1:6] == [0., -5.25, -7.833333333333333, -2.4166666666666666, -1., 0.]
    g = [C(2.0), [0.25, 1, 0, 1, 1], C(1 / 3)]
    o = ob.ql_finite_section(1, 1, g, j=50)
    assert o["status"][0] == 0
    L = _dense_L(o["Lband"][0], 8)
    b = np.zeros(8); b[:3] = [1, 2, 3]
    assert np.array_equal((L @ b)[:6], [0.0, -5.25, -7.833333333333333, -2.4166666666666666, -1.0, 0.0])
    # the finite banded QL under it equals the dense Householder QL of the final section
    N = int(o["N"][0])
    A = np.zeros((N, N))
    for k in range(N):
        A[k, k] = 0.25 + 1 / (k + 1)
        if k + 1 < N:
            A[k, k + 1] = 2; A[k + 1, k] = 1 / 3
    F, t = tq.dense_ql(A)
    ref = np.array([[F[i, i - c] if i - c >= 0 else 0 for c in range(3)] for i in range(50)])
    assert np.array_equal(ref, o["Lband"][0]) and np.array_equal(t[:50], o["tau"][0])


def test_runaway_protection():
    # :223-226  A = _BandedMatrix(Vcat((0:∞)', ((1:∞).+1/4)', Ones(1,∞)./3), ℵ₀, 1, 1): never converges -> the error
    # super-diagonal data row is indexed by column: A[j-1,j] = j-1 (1-based) = t + 1 with t = row (0-based)
    g = [[0, 1, 1, 1, 0], [0, 1.25, 1, 1, 0], C(1 / 3)]
    o = ob.ql_finite_section(1, 1, g, j=50)
    assert o["status"][0] == 3 and np.all(np.isnan(o["Lband"][0]))


def test_explicit_tolerance_and_toeplitz_comparison():
    # :227-256  Tridiagonal([[1,2]; Fill(1)], [[1,2]; Fill(3)], [[1,2]; Fill(1)]) with tol 1e-10, and against the Toeplitz QL
    head = np.array([[1.0, 2.0, 1.0], [1.0, 2.0, 3.0], [1.0, 2.0, 1.0]])   # rows: super A[t,t+1], diag A[t,t], sub A[t+1,t]
    g = [C(1.0), C(3.0), C(1.0)]
    o = ob.ql_finite_section(1, 1, g, j=60, head=head, tol=1e-10)
    assert o["status"][0] == 0
    # compare with the perturbed-Toeplitz QL oracle (ql_hessenberg! path): L[1:.., 1:..] ≈
    A = np.zeros((6, 6))
    d = [1, 2, 3, 3, 3, 3]; e = [1, 2, 1, 1, 1]
    for k in range(6):
        A[k, k] = d[k]
        if k < 5:
            A[k, k + 1] = e[k]; A[k + 1, k] = e[k]
    res = tq.ql_pert_toeplitz(A[:4, :4].copy(), np.array([1.0, 3.0, 1.0]))
    F = res["head_factors"]
    L = _dense_L(o["Lband"][0], 4)
    assert np.allclose(np.abs(np.tril(F)), np.abs(L), rtol=0, atol=1e-9)      # rows agree up to the sign convention of Q


def test_non_tridiagonal_against_finite_section():
    # :262-266  A = _BandedMatrix(Vcat(2*Ones(2,∞), ((1 ./(1:∞)).+4)', Ones(1,∞)./3, Ones(1,∞)./3), ℵ₀, 2, 2):
    # ql(A[1:100,1:100]).Q[1:10,1:10] ≈ Q[1:10,1:10]  ->  here the factors themselves
    g = [C(2.0), C(2.0), [4.0, 1, 0, 1, 1], C(1 / 3), C(1 / 3)]
    o = ob.ql_finite_section(2, 2, g, j=50)
    assert o["status"][0] == 0
    N = 100
    A = np.zeros((N, N))
    for k in range(N):
        A[k, k] = 4 + 1 / (k + 1)
        for dd in (1, 2):
            if k + dd < N:
                A[k, k + dd] = 2.0; A[k + dd, k] = 1 / 3
    F, t = tq.dense_ql(A)
    ref = np.array([[F[i, i - c] if i - c >= 0 else 0 for c in range(5)] for i in range(10)])
    assert np.allclose(ref, o["Lband"][0][:10], rtol=0, atol=1e-13) and np.allclose(t[:10], o["tau"][0][:10], rtol=0, atol=1e-13)


def test_complex_entries_against_dense_ql():
    # :257-262  A = im * Tridiagonal([[1,2]; Fill(1)], [[1,2]; Fill(3)], [[1,2]; Fill(1)]): (Q*L)[1:50,1:50] ≈ A  ->  here the
    # factors against the dense complex Householder QL of the final section (numpy's complex division rounds differently: 1e-14)
    head = 1j * np.array([[1.0, 2.0, 1.0], [1.0, 2.0, 3.0], [1.0, 2.0, 1.0]])
    g = [C(1j), C(3j), C(1j)]
    o = ob.ql_finite_section(1, 1, g, j=51, head=head)
    assert o["status"][0] == 0 and np.iscomplexobj(o["Lband"])
    N = int(o["N"][0])
    A = np.zeros((N, N), complex)
    d = [1, 2] + [3] * (N - 2); e = [1, 2] + [1] * (N - 3)
    for k in range(N):
        A[k, k] = 1j * d[k]
        if k < N - 1:
            A[k, k + 1] = 1j * e[k]; A[k + 1, k] = 1j * e[k]
    F, t = tq.dense_ql(A)
    ref = np.array([[F[i, i - c] if i - c >= 0 else 0 for c in range(3)] for i in range(51)])
    assert np.abs(ref - o["Lband"][0]).max() <= 1e-14 and np.abs(t[:51] - o["tau"][0]).max() <= 1e-14
