"""Pins the numpy oracle of the block-tridiagonal Toeplitz-tail QL (src/infql.jl:166-209). CPU only."""
import numpy as np

from oracle import block_ql as bq
from oracle import toeplitz_ql as tq


def _finite_section(dl, d, du, nb, x, d0=None):
    n = d.shape[0]
    A = np.zeros((nb * n, nb * n))
    for k in range(nb):
        A[k * n:(k + 1) * n, k * n:(k + 1) * n] = (d0 if (k == 0 and d0 is not None) else d) - x * np.eye(n)
        if k < nb - 1:
            A[k * n:(k + 1) * n, (k + 1) * n:(k + 2) * n] = du
            A[(k + 1) * n:(k + 2) * n, k * n:(k + 1) * n] = dl
    return A


def test_periodic_schrodinger_energies():
    # examples/periodicschrodinger.jl:24 energies; SURVEY Appendix B: 7-51 iterations, median 13, all converge;
    # test/test_infql.jl:278-287: L[1,1] is a Float64 at x = -0.95
    xx = np.concatenate([np.arange(-0.95, 0.951, 0.05), np.arange(2.25, 4.0001, 0.125 / 4)])
    res = [bq.periodic_schrodinger(x) for x in xx]
    its = [r[1] for r in res]
    assert all(r[2] == 0 for r in res)
    assert min(its) == 7 and max(its) == 51 and int(np.median(its)) == 13
    assert np.isfinite(res[0][0])


def test_against_finite_section():
    # test/test_periodic.jl:5-18: ∞ factors vs the QL of a 100-block finite section (the fixed point is iterated to 1e-10)
    c = np.array([[0.0, 1.0], [0.0, 0.0]])
    a = np.array([[-1.0, 1.0], [1.0, 1.0]])
    b = np.array([[0.0, 0.0], [1.0, 0.0]])
    a0 = a.copy()
    a0[0, 0] = 2.0
    for x in (-0.95, 0.3, 2.5, 3.9):
        l11 = bq.periodic_schrodinger(x)[0]
        F, _ = tq.dense_ql(_finite_section(c, a, b, 100, x, d0=a0))
        # (the phase of the tail's fixed point is fixed by the iteration's start, so only |L[1,1]| is comparable)
        assert np.isclose(abs(l11), abs(F[0, 0]), rtol=0, atol=5e-9)
    # unperturbed operator of test_periodic.jl:6-8: head == tail
    l11 = bq.blocktri_ql_l11([c], [a], [b], c, a, b)[0]
    F, _ = tq.dense_ql(_finite_section(c, a, b, 100, 0.0))
    assert np.isclose(l11, F[0, 0], atol=5e-9)


def test_degenerate_and_four_by_four_blocks():
    # test/test_periodic.jl:28-34: |L[1,1]| <= 1e-11 (degenerate)
    c2, a2, b2 = np.array([[0, 0.5], [0, 0]]), np.array([[0, 2.0], [0.5, 0]]), np.array([[0, 0.0], [2.0, 0]])
    l11, it, st, _ = bq.blocktri_ql_l11([c2], [a2], [b2], c2, a2, b2)
    assert st == 0 and abs(l11) <= 1e-11
    # test/test_periodic.jl:42-60 ("bi"): 4x4 blocks, Q*L ≈ A  ->  here: L[1,1] against the finite section
    B = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [0, 0, 0, 0], [0, 0, 0, 0]]) / 2
    A0 = np.array([[1, .5, .5, 0], [.5, -1, 0, .5], [.5, 0, -1, 0], [0, .5, 0, 1.0]])
    A = np.array([[1, 0, .5, 0], [0, -1, 0, .5], [.5, 0, -1, 0], [0, .5, 0, 1.0]])
    l11, it, st, _ = bq.blocktri_ql_l11([B], [A0], [B.T.copy()], B, A, B.T.copy())
    assert st == 0
    F, _ = tq.dense_ql(_finite_section(B, A, B.T, 60, 0.0, d0=A0))
    assert np.isclose(l11, F[0, 0], atol=5e-8)


def test_complex_nonselfadjoint_shift():
    # test/test_periodic.jl:20-26: the degenerate blocks shifted by -5im*I (complex eltype: the QL loop runs to k = 1)
    c2, a2, b2 = np.array([[0, 0.5], [0, 0]]), np.array([[0, 2.0], [0.5, 0]]), np.array([[0, 0.0], [2.0, 0]])
    for x in (5j, 2.0 + 4.0j, -0.5 - 3.0j):
        l11, it, st, _ = bq.blocktri_ql_l11([c2], [a2], [b2], c2, a2, b2, shift=x, maxit=5000)
        assert st == 0 and it < 200
        A = _finite_section(c2, a2, b2, 40, 0.0).astype(complex) - x * np.eye(80)
        F, _ = tq.dense_ql(A)
        assert abs(np.imag(l11)) <= 1e-15 * abs(l11)          # the k = 1 reflector makes L[1,1] real
        # (the phase of the tail's fixed point is fixed by the iteration's start, so only |L[1,1]| is comparable)
        assert np.isclose(abs(l11), abs(F[0, 0]), rtol=0, atol=5e-9)


def test_inf_block_factors_against_the_finite_section():
    """test/test_periodic.jl:5-18: the ∞ factors assembled from (factored head, τ, fixed point P, τ tail) — the data the GPU
    path exports — agree with the Householder QL of the 100-block finite section: factors[1:100,1:100], τ[1:100],
    L[1:100,1:100], and Q[1:10,1:12] L[1:12,1:10] = A[1:10,1:10] (the reference compares with ≈, i.e. 1.5e-8; the fixed point
    is iterated to 1e-10)."""
    c = np.array([[0., 1.], [0., 0.]])
    a = np.array([[-1., 1.], [1., 1.]])
    b = np.array([[0., 0.], [1., 0.]])
    L11, it, st, ex = bq.blocktri_ql_l11([c], [a], [b], c, a, b)
    assert st == 0
    F, tau = bq.assemble_inf_factors(ex, 2, ex["N"], 60)
    nb = 100
    A = np.zeros((2 * nb, 2 * nb))
    for k in range(nb):
        A[2 * k:2 * k + 2, 2 * k:2 * k + 2] = a
        if k + 1 < nb:
            A[2 * k:2 * k + 2, 2 * k + 2:2 * k + 4] = b
            A[2 * k + 2:2 * k + 4, 2 * k:2 * k + 2] = c
    Ff, tf = bq.dense_ql_seq(A)
    assert np.abs(Ff[:100, :100] - F[:100, :100]).max() < 1e-9
    assert np.abs(tf[:100] - tau[:100]).max() < 1e-9
    assert np.abs(np.tril(Ff)[:100, :100] - np.tril(F)[:100, :100]).max() < 1e-9
    Q, L = bq.dense_Q_L(F, tau, 2)
    assert np.abs(Q[:10, :12] @ L[:12, :10] - A[:10, :10]).max() < 1e-9
    # complex non-selfadjoint operator of test_periodic.jl:20-26
    c2, a2, b2 = np.array([[0, 0.5], [0, 0]]), np.array([[0, 2.0], [0.5, 0]]), np.array([[0, 0.0], [2.0, 0]])
    L11, it, st, ex = bq.blocktri_ql_l11([c2], [a2], [b2], c2, a2, b2, shift=5j)
    assert st == 0
    F, tau = bq.assemble_inf_factors(ex, 2, ex["N"], 30)
    Q, L = bq.dense_Q_L(F, tau, 2)
    A = np.zeros((20, 20), dtype=complex)
    for k in range(10):
        A[2 * k:2 * k + 2, 2 * k:2 * k + 2] = a2 - 5j * np.eye(2)
        if k + 1 < 10:
            A[2 * k:2 * k + 2, 2 * k + 2:2 * k + 4] = b2
            A[2 * k + 2:2 * k + 4, 2 * k:2 * k + 2] = c2
    assert np.abs(Q[:10, :12] @ L[:12, :10] - A[:10, :10]).max() < 1e-9
