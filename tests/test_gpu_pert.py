"""Parity of the CUDA perturbed-Toeplitz QL (K3 + K4) against the numpy/LAPACK oracle, through the C ABI.
Tolerance: 1e-12 relative on L[1,1] (the tail goes through K3's root finder, so not bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle(head_data, a_col, lam, p):
    from oracle import toeplitz_ql as tq
    dt = np.result_type(np.asarray(head_data), np.asarray(a_col), lam, np.float64)
    H = tq.pert_toeplitz_head(head_data, a_col, p=p).astype(dt) - lam * np.eye(p)
    a_rev = np.asarray(a_col)[::-1].astype(dt)
    a_rev[-2] -= lam
    return tq.ql_pert_toeplitz(H, a_rev)


def test_jacobi_real_path(ila):
    from oracle import toeplitz_ql as tq
    head_data = np.array([[0.5], [1.0], [0.5]])
    a_col = np.array([0.5, 0.0, 0.5])
    head = tq.pert_toeplitz_head(head_data, a_col, p=2)
    lams = np.array([3.5, 2.5, -3.0, 1.5])
    L, st, hf = ila.ql_perttoeplitz_l11(head, a_col, lams, want_head_factors=True)
    for i, lam in enumerate(lams):
        o = _oracle(head_data, a_col, lam, 2)
        assert st[i] == o["status"]
        if o["status"] == 0:
            assert np.isclose(L[i], o["L11"], rtol=1e-12)
            assert np.allclose(np.tril(hf[i]), np.tril(o["head_factors"]), rtol=1e-12, atol=1e-14)
    assert L[0] == pytest.approx(-2.4010806467474954, rel=1e-13)


def test_complex_shifts_and_hessenberg_heads(ila):
    from oracle import toeplitz_ql as tq
    head_data = np.array([[0.5], [2.0], [2.0]])
    a_col = np.array([0.5, 0.0, 2.0])
    head = tq.pert_toeplitz_head(head_data, a_col, p=3)
    lams = np.array([-2.1 - 0.01j, -2.1 + 0.01j, -2.1 + 0.0j, -1.0 + 1j, -3.0 + 1j, -3.0 - 1j])
    L, st = ila.ql_perttoeplitz_l11(head, a_col, lams)
    for i, lam in enumerate(lams):
        o = _oracle(head_data, a_col, lam, 3)
        assert st[i] == 0 and o["status"] == 0
        assert abs(L[i] - o["L11"]) <= 1e-12 * abs(o["L11"])
    # (3,1) bandwidths, 2-column head, a grid of shifts (test_infql.jl:89-99, 112-121)
    head_data = np.array([[0.5, 0.5], [-1, 3], [2, 2], [1, 1], [0.1, 0.1]], dtype=float)
    a_col = np.array([0.5, 3, 2, 1, 0.1])
    head = tq.pert_toeplitz_head(head_data, a_col, p=5)
    x = np.linspace(-6, 8, 24)
    y = np.linspace(-5, 5, 24)
    lams = (x[None, :] + 1j * y[:, None])
    L, st = ila.ql_perttoeplitz_l11(head, a_col, lams)
    nchk = 0
    for idx in np.ndindex(lams.shape):
        o = _oracle(head_data, a_col, lams[idx], 5)
        assert st[idx] == o["status"]
        if o["status"] == 0 and abs(o["tail"]["L11"]) > 0.05:
            assert abs(L[idx] - o["L11"]) <= 1e-11 * max(1.0, abs(o["L11"]))
            nchk += 1
    assert nchk > 300
