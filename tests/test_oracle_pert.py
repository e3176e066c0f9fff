"""Pins the numpy oracle of the perturbed-Toeplitz head `ql_hessenberg!` (src/infql.jl:60-93). CPU only."""
import numpy as np
import pytest

from oracle import toeplitz_ql as tq


def _finite(head_data, a_col, n, lam):
    H = tq.pert_toeplitz_head(head_data, a_col, p=n).astype(np.result_type(head_data, a_col, lam, np.float64))
    return H - lam * np.eye(n)


def _oracle(head_data, a_col, lam, p):
    dt = np.result_type(np.asarray(head_data), np.asarray(a_col), lam, np.float64)
    H = tq.pert_toeplitz_head(head_data, a_col, p=p).astype(dt) - lam * np.eye(p)
    a_rev = np.asarray(a_col)[::-1].astype(dt)
    a_rev[-2] -= lam
    return tq.ql_pert_toeplitz(H, a_rev)


def test_jacobi_minus_shift_constant():
    # test/test_infql.jl:147-153: J - 3.5I, (Q'b)[1] ≈ 0.9892996329463546
    head_data = np.array([[0.5], [1.0], [0.5]])
    a_col = np.array([0.5, 0.0, 0.5])
    o = _oracle(head_data, a_col, 3.5, p=2)
    F, t = tq.dense_ql(_finite(head_data, a_col, 300, 3.5))
    assert o["status"] == 0
    assert np.isclose(o["L11"], F[0, 0], rtol=1e-13)
    assert np.isclose(o["head_factors"][1, 0], F[1, 0], rtol=1e-13)
    Q = tq.ql_packed_Q(F, t)
    assert np.isclose(Q[0, 0], 0.9892996329463546, rtol=1e-13)


@pytest.mark.parametrize("lam", [-2.1 - 0.01j, -2.1 + 0.01j, -2.1 + 0.0j, -1.0 + 1j, -3.0 + 1j, -3.0 - 1j])
def test_pert_tridiagonal_complex_shifts(lam):
    # test/test_infql.jl:70-78: A = Tridiagonal(Fill(2), [2, 0, 0, ...], Fill(0.5)), Q*L ≈ A - λI  ->  L[1,1] vs finite section
    head_data = np.array([[0.5], [2.0], [2.0]])
    a_col = np.array([0.5, 0.0, 2.0])
    o = _oracle(head_data, a_col, lam, p=2)
    assert o["status"] == 0
    F, _ = tq.dense_ql(_finite(head_data, a_col, 400, lam))
    assert np.isclose(abs(o["L11"]), abs(F[0, 0]), rtol=1e-11)


def test_pert_hessenberg_head():
    # test/test_infql.jl:89-99: a = [0.1,1,2,3,0.5] tail with a 2-column head, bandwidths (3,1)
    head_data = np.array([[0.5, 0.5], [-1, 3], [2, 2], [1, 1], [0.1, 0.1]], dtype=float)
    a_col = np.array([0.5, 3, 2, 1, 0.1])
    o = _oracle(head_data, a_col, 0.0, p=4)
    assert o["status"] == 0
    F, _ = tq.dense_ql(_finite(head_data, a_col, 400, 0.0))
    assert np.isclose(o["L11"], F[0, 0], rtol=1e-11)
    assert np.allclose(o["head_factors"][:3, 0], F[:3, 0], rtol=1e-10)
