"""Pins the numpy Toeplitz-QL oracle against the reference's own known answers.  CPU only."""
import numpy as np
import pytest

from oracle import toeplitz_ql as tq


def test_golden_householder_constants():
    # test/test_reduceql.jl:41-45: (Z,A,B) = (2, 2.1+0.01im, 0.5)
    sigma = -0.9998783597231515 + 0.01559697910943707j
    v = -0.18889614060709453 + 7.427756729736341e-19j
    tau = 1.9310951421717215 - 7.593434620701001e-18j
    o = tq.ql_toeplitz_tail(np.array([2, 2.1 + 0.01j, 0.5]))
    assert o["status"] == 0
    assert abs(np.conj(1 - o["tau"][0]) - sigma) < 2e-15
    assert abs(o["X"][0, -1] - v) < 2e-15
    assert abs(o["tau"][1] - tau) < 2e-15


def _tri(Z, A, B, n, dtype):
    return (np.diag(np.full(n, A, dtype=dtype)) + np.diag(np.full(n - 1, B, dtype=dtype), 1)
            + np.diag(np.full(n - 1, Z, dtype=dtype), -1))


@pytest.mark.parametrize("ZAB", [(2.0, 5.1, 0.5), (2.0, 2.2, 0.5), (2.0, -2.2, 0.5), (2.0, -5.1, 0.5),
                                 (0.5, 5.1, 2.0), (0.5, -5.1, 2.0)])
def test_real_tridiagonal_matches_finite_section(ZAB):
    # test/test_infql.jl:12-19: L[1:10,1:10] ≈ Ln[1:10,1:10] for the finite section (n = 100_000 there; the
    # corner converges geometrically, 600 is ample)
    Z, A, B = ZAB
    o = tq.ql_toeplitz_tail(np.array([Z, A, B]))
    assert o["status"] == 0
    F, _ = tq.dense_ql(_tri(Z, A, B, 600, np.float64))
    X = o["X"]
    assert np.isclose(o["L11"], F[0, 0], rtol=1e-12)
    # tail column of L = X[2, m:-1:1]
    assert np.allclose(X[1, ::-1], [F[5, 5], F[6, 5], F[7, 5]], rtol=1e-12)
    # first column = [X[1,m-1]; X[2,m-1:-1:1]]
    assert np.allclose([X[0, 1], X[1, 1], X[1, 0]], [F[0, 0], F[1, 0], F[2, 0]], rtol=1e-12)


@pytest.mark.parametrize("ZAB", [(0.5, 2.1, 2.0), (0.5, -1.1, 2.0)])
def test_domain_error_cases(ZAB):
    # test/test_infql.jl:36-38
    Z, A, B = ZAB
    o = tq.ql_toeplitz_tail(np.array([Z, A, B]))
    assert o["status"] != 0


def _inf_ql_corner(o, n):
    """Dense n x n corner of L and (n x n+1) of Q from the oracle's tail data (infqltoeplitz.jl:63-64)."""
    X, tau = o["X"], o["tau"]
    m = X.shape[-1]
    dt = X.dtype
    N = n + m
    L = np.zeros((N, N), dtype=dt)
    first = np.concatenate([[X[0, m - 2]], X[1, m - 2::-1]])
    tailc = X[1, ::-1]
    L[:len(first), 0] = first
    for j in range(1, N):
        for i, val in enumerate(tailc):
            r = j + i
            if r < N:
                L[r, j] = val
    # 2x2 block: QLPackedQ(X[:, m-1:m], tau)
    fac = X[:, m - 2:m]
    q = tq.ql_packed_Q(fac, tau)
    Q = np.eye(N, dtype=dt)
    # LowerHessenbergQ applied to a vector runs q_1 on rows 1:2 first, then q_2 on rows 2:3, ...
    # (src/banded/hessenbergq.jl:111-123); apply the same sweep to the identity
    for k in range(0, N - 1):
        Q[k:k + 2, :] = q @ Q[k:k + 2, :]
    return Q, L


@pytest.mark.parametrize("ZAB", [(2, 5.0j, 0.5), (2, 2.1 + 0.1j, 0.5), (2, 2.1 - 0.1j, 0.5), (2, 1.5 + 0.1j, 0.5),
                                 (2, 1.5 + 0.0j, 0.5), (2, 0.0 + 0.0j, 0.5), (2, -0.1 + 0.1j, 0.5),
                                 (2, -1.1 + 0.1j, 0.5), (2, -0.1 - 0.1j, 0.5)])
def test_complex_tridiagonal_reconstruction(ZAB):
    # test/test_infql.jl:21-34: Q[1:10,1:12]*L[1:12,1:10] ≈ T[1:10,1:10]
    Z, A, B = ZAB
    o = tq.ql_toeplitz_tail(np.array([Z, A, B], dtype=np.complex128))
    assert o["status"] == 0
    Q, L = _inf_ql_corner(o, 40)
    T = _tri(Z, A, B, Q.shape[0], np.complex128)
    assert np.allclose((Q @ L)[:10, :10], T[:10, :10], atol=1e-12)


@pytest.mark.parametrize("a", [[1, 2, 3, 0.5], [1, 2, 3 + 1j, 0.5],
                               [0.4314153362874331, -0.19352887774167749 - 0.3738926065520737j,
                                0.538107104952824 - 0.951620830938543j, 0.5794879759059747 + 0.0j][::-1][::-1]])
def test_hessenberg_reconstruction(a):
    # test/test_infql.jl:41-51: a = [a_{-l}..a_0,a_{+1}] reversed data column; Q*L ≈ T
    a = np.asarray(a)
    o = tq.ql_toeplitz_tail(a)
    assert o["status"] == 0
    Q, L = _inf_ql_corner(o, 40)
    N = Q.shape[0]
    m = len(a)
    T = np.zeros((N, N), dtype=Q.dtype)
    for k, ak in enumerate(a):           # a[k] sits on diagonal d = k - (m-2)  (a[m-1] = super-diagonal)
        d = k - (m - 2)
        T += np.diag(np.full(N - abs(d), ak, dtype=Q.dtype), d)
    assert np.allclose((Q @ L)[:10, :10], T[:10, :10], atol=1e-12)


def test_bullhead_finite_section_and_probe_values():
    # SURVEY Appendix B probe values + |L11| ≈ |ql(A[1:n,1:n]).L[1,1]| (test_reduceql.jl:577 identity, transposed there)
    col = np.array([2j, 0, 0, 1, 0.7])
    lam = np.array([5, -3 - 0.1j, 2 + 3j, -6 + 2j, 0.1 + 0.1j])
    L, st = tq.ql_toeplitz_l11(col, lam)
    assert list(st) == [0, 0, 0, 0, 1]
    assert np.allclose(L[:4].real, [-4.747700587226856, -1.83554289328665, -2.4543911495549953, -5.917786604460436],
                       rtol=1e-13)
    assert np.abs(L[:4].imag).max() < 1e-15
    n = 400
    for lm, l11 in zip(lam[:2], L[:2]):
        A = np.zeros((n, n), dtype=complex)
        for d, v in ((1, 2j), (0, -lm), (-2, 1.0), (-3, 0.7)):
            A += np.diag(np.full(n - abs(d), v, dtype=complex), d)
        F, _ = tq.dense_ql(A)
        assert np.isclose(abs(F[0, 0]), abs(l11), rtol=1e-12)


def test_real_path_householder_sign_rule_is_the_finite_section_limit():
    """The C ABI's `_f64` sign rule (oracle sign_rule="householder").  Upstream the sign of the whole real-typed
    factorization is whatever LAPACK's dgeev returns for the eigenvector (infqltoeplitz.jl:14-15) — (Q, L) or (-Q, -L).
    Independent evidence for the rule: on random real Hessenberg Toeplitz symbols the householder-rule L equals the dense
    finite-section Householder QL's L (sign included — what test/test_infql.jl:14-17 asserts), and the literal-LAPACK
    outcome is always one of the two members of the pair."""
    rng = np.random.default_rng(7)
    checked = lapack_other = 0
    for trial in range(120):
        l = int(rng.integers(1, 5))
        col = rng.normal(size=l + 2)
        if abs(col[0]) < 0.3:
            col[0] = 1.0
        a = col[::-1].copy()
        a[-2] -= rng.normal() * 4
        o = tq.ql_toeplitz_tail(a, sign_rule="householder")
        ol = tq.ql_toeplitz_tail(a, sign_rule="lapack")
        assert o["status"] == ol["status"]
        if o["status"] != 0:
            continue
        m = len(a)
        def section(n):
            A = np.zeros((n, n))
            for i in range(n):
                for j in range(max(0, i - l), min(n, i + 2)):
                    A[i, j] = a[(j - i) + m - 2]
            return tq.dense_ql(A)[0]
        F1, F2 = section(120), section(170)
        if abs(F1[0, 0] - F2[0, 0]) > 1e-9 * abs(F2[0, 0]):
            continue                                                     # finite section not converged yet: skip
        checked += 1
        X = o["X"]
        first_col = np.concatenate([[X[0, m - 2]], X[1, m - 2::-1]])     # [X[1,m-1]; X[2,m-1:-1:1]]
        tail_col = X[1, ::-1]                                            # X[2,m:-1:1]
        assert np.allclose(F2[:m, 0], first_col, rtol=1e-8, atol=1e-10)
        assert np.allclose(F2[3:3 + m, 3], tail_col, rtol=1e-8, atol=1e-10)
        same = np.allclose(ol["X"], X, rtol=1e-12, atol=1e-14)
        # (Q,L) -> (-Q,-L): rows 1 (up to column m-1) and 2 of X negated, v = X[1,m] negated, tau1 -> 2 - tau1
        other = np.allclose(ol["X"], -X, rtol=1e-12, atol=1e-14) and np.isclose(ol["tau"][0], 2 - o["tau"][0]) and \
            np.isclose(ol["tau"][1], o["tau"][1])
        assert same or other
        lapack_other += int(other and not same)
    assert checked >= 60
    assert lapack_other > 0          # the ambiguity is real: LAPACK's sign is not the finite-section one everywhere


def test_hessenberg_q_apply_restatement_is_pinned():
    """test/test_hessenbergq.jl:31-62: LowerHessenbergQ(Fill(q, 2)) = [1 0; 0 q] [q 0; 0 1] column by column, its adjoint,
    and the ∞ case Q[1,1] = cos(0.1) = 0.9950041652780258."""
    c, s = np.cos(0.1), np.sin(0.1)
    q = np.array([[c, s], [s, -c]])
    QM = np.block([[np.eye(1), np.zeros((1, 2))], [np.zeros((2, 1)), q]]) @ np.block([[q, np.zeros((2, 1))], [np.zeros((1, 2)), np.eye(1)]])
    for j in range(3):
        e = np.zeros(3); e[j] = 1.0
        blocks = lambda n: q if n <= 2 else np.eye(2)
        y, _ = tq.hessenberg_q_apply(blocks, e, 3)
        assert np.allclose(y, QM[:, j], atol=1e-15)
        yt, _ = tq.hessenberg_q_apply(blocks, e, 3, adjoint=True)
        assert np.allclose(yt, QM.T[:, j], atol=1e-15)
    y, stop = tq.hessenberg_q_apply(lambda n: q, np.array([1.0]), 8000)
    assert abs(y[0] - 0.9950041652780258) < 1e-15
    assert 0 < stop < 8000 and np.all(y[stop + 1:] == 0)          # the early exit fires once |s|^n underflows
    assert abs(np.sum(y ** 2) - 1.0) < 1e-12                       # an orthogonal column
