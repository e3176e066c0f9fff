"""Resumable factorization state: the reference's AdaptiveQRData / AdaptiveCholeskyFactors carry
`ncols` and continue from it — partialqr!(F, n) (src/infqr.jl:32-48, pinned by test/test_infqr.jl:19-28) and
partialcholesky!(F, n) (src/infcholesky.jl:42-59).  Here: a sweep stopped at a column limit keeps its work and its state;
continuing it gives the SAME BITS as one sweep."""
import numpy as np
import pytest
import torch
from scipy.special import jv

pytestmark = pytest.mark.gpu


def _snapshot(plan):
    torch.cuda.synchronize()
    return dict(ncols=plan.d_ncols.cpu().numpy().copy(), nconv=plan.d_nconv.cpu().numpy().copy(),
                nswept=plan.d_nswept.cpu().numpy().copy(), status=plan.d_status.cpu().numpy().copy(),
                xoff=plan.d_xoff.cpu().numpy().copy(), work=plan.d_work.cpu().numpy().copy())


def _solution(plan):
    plan.run_backsolve()
    torch.cuda.synchronize()
    xoff = plan.d_xoff.cpu().numpy()
    x = plan.d_x.cpu().numpy()
    return [x[xoff[i]:xoff[i + 1]].copy() for i in range(plan.batch)]


def test_adaptive_qr_resumes_bit_exact(ila):
    """Bessel solves whose first column limit is too small: status 3 with the finished columns KEPT, then resumed."""
    from oracle import banded as ob
    zs = np.array([300.0, 900.0, 2500.0, 2700.0, 7000.0])
    p, q, r = ila.bessel_operator(zs)
    b = jv(1, zs)[:, None]
    cap = 12000
    one = ila.QrBandedPlan(1, 1, p, q, r, b=b, ncap=cap)
    one.run_factor()
    ref = _snapshot(one)
    xref = _solution(one)
    assert np.all(ref["status"] == 0)

    two = ila.QrBandedPlan(1, 1, p, q, r, b=b, ncap=cap)
    two.run_factor_resumable(3500, resume=False)
    mid = _snapshot(two)
    # z = 300 and 900 converge below 3500 columns; the others stop at the limit WITH their columns kept (round 1: ncols = 0)
    assert list(mid["status"]) == [0, 0, 3, 3, 3]
    assert np.all(mid["ncols"][2:] == 3000) and np.all(mid["nswept"][2:] == 3000)      # last whole COLGROWTH chunk that fits
    assert np.array_equal(mid["ncols"][:2], ref["ncols"][:2])
    two.run_factor_resumable(5500, resume=True)          # still too small for z = 7000
    mid2 = _snapshot(two)
    assert list(mid2["status"]) == [0, 0, 0, 0, 3] and mid2["ncols"][4] == 5000
    two.run_factor_resumable(cap, resume=True)
    fin = _snapshot(two)
    for k in ("ncols", "nconv", "nswept", "status", "xoff"):
        assert np.array_equal(fin[k], ref[k]), k
    xs = _solution(two)
    for i in range(len(zs)):
        assert np.array_equal(xs[i].view(np.int64), xref[i].view(np.int64)), f"instance {i}: resumed solve differs from one sweep"
    # and both equal the oracle bit for bit
    o = ob.qr_banded_solve(1, 1, p, q, r, b=b, ncap=cap)
    for i in range(len(zs)):
        assert np.array_equal(xs[i], o["x"][i, :int(o["ncols"][i])])


def test_partialqr_in_two_steps_equals_one(ila):
    """test/test_infqr.jl:12-28: partialqr!(F, 3) then partialqr!(F, 6) — factors, tau and the updated trailing columns."""
    from oracle import banded as ob
    pq = dict(p=[[1.0, 1.0, 1.0]], q=[[0.0, 1.0, 0.0]])               # bands (1, 1:∞, 1)
    for (n1, n2) in [(3, 6), (6, 2500), (1000, 2000)]:
        a = ila.QrBandedPlan(1, 1, pq["p"], pq["q"], b=[1.0], ncap=4000, colgrowth=1000)
        a.partialqr(n2)
        fa, ta, _, xo = a.export_factors()
        wa = a.state_window().copy()
        b2 = ila.QrBandedPlan(1, 1, pq["p"], pq["q"], b=[1.0], ncap=4000, colgrowth=1000)
        b2.partialqr(n1)
        f1, t1, _, _ = b2.export_factors()
        assert int(b2.d_ncols.cpu()[0]) == n1
        b2.partialqr(n2, resume=True)
        fb, tb, _, _ = b2.export_factors()
        assert int(b2.d_ncols.cpu()[0]) == n2
        assert np.array_equal(fb.view(np.int64), fa.view(np.int64)) and np.array_equal(tb.view(np.int64), ta.view(np.int64))
        assert np.array_equal(b2.state_window().view(np.int64), wa.view(np.int64))
        # the first n1 columns are final after the first call (what was still moving were the trailing columns n1+1.., i.e. the
        # window — the reference's "extra columns have been modified" state)
        assert np.array_equal(t1, tb[:n1]) and np.array_equal(f1, fb[:n1])
        # against the oracle's factors (dense Householder restatement pinned by -sqrt(2), 1 + sqrt(2)/2)
        o = ob.qr_banded_solve(1, 1, pq["p"], pq["q"], b=[1.0, 2.0, 3.0], ncap=5000, want_factors=True)
        m = min(n2, int(o["ncols"][0]) - 3)
        assert np.array_equal(ta[:m], o["tau"][0, :m])
        assert np.array_equal(fa[:m, 3], o["factors"][0, :m, 3])     # sub-diagonal band row: the reflector entries v
        assert np.array_equal(fa[:m, 2], o["factors"][0, :m, 2])     # diagonal of R
    assert np.isclose(fa[0, 2], -np.sqrt(2), rtol=1e-15) and np.isclose(ta[0], 1 + np.sqrt(2) / 2, rtol=1e-15)
    # the window after n columns is Q_n' A restricted to rows n+1.., i.e. triu!(F.data[1:n+1, 1:n+2]) ≈ Q'A[1:n+1, 1:n+2]
    n = 3
    c = ila.QrBandedPlan(1, 1, pq["p"], pq["q"], b=[1.0], ncap=4000)
    c.partialqr(n)
    W = c.state_window()[0]                                            # W[j][i] = (Q'A)[n+i, n+j]
    A = np.zeros((n + 1, n + 3))
    for i in range(n + 1):
        for j in range(n + 3):
            if abs(i - j) <= 1:
                A[i, j] = 1.0 if i != j else float(i + 1)
    Q, _ = np.linalg.qr(A[:, :n], mode="complete")
    QA = Q.T @ A
    # n reflectors touch rows 0..n: the complement of span(A[:, :n]) in R^(n+1) is one-dimensional, so row n of Q'A is
    # unique up to sign; row n+1 of the window is the untouched operator row
    assert np.allclose(np.abs(W[:3, 0]), np.abs(QA[n, n:n + 3]), rtol=1e-12, atol=1e-14)
    assert list(W[:3, 1]) == [1.0, float(n + 2), 1.0]


def test_adaptive_cholesky_resumes_bit_exact(ila):
    sigma = np.array([0.01, 0.7, 3.0, 9.9])
    p4, q4 = ila.pentadiagonal_operator(sigma, eps=1e-3)
    one = ila.QrBandedPlan(0, 2, p4, q4, b=[1.0], ncap=9000, kind="chol")
    one.run_factor()
    ref = _snapshot(one)
    xref = _solution(one)
    assert np.all(ref["status"] == 0) and ref["ncols"].max() > 2000
    lim = int(ref["ncols"].max()) - 1200
    two = ila.QrBandedPlan(0, 2, p4, q4, b=[1.0], ncap=9000, kind="chol")
    two.run_factor_resumable(lim, resume=False)
    mid = _snapshot(two)
    assert (mid["status"] == 3).any() and np.all(mid["ncols"][mid["status"] == 3] > 0)
    two.run_factor_resumable(9000, resume=True)
    fin = _snapshot(two)
    for k in ("ncols", "nconv", "nswept", "status", "xoff"):
        assert np.array_equal(fin[k], ref[k]), k
    xs = _solution(two)
    for i in range(len(sigma)):
        assert np.array_equal(xs[i].view(np.int64), xref[i].view(np.int64)), f"shift {i}"
