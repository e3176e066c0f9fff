"""Error behaviour of the C ABI on a GPU box: the reference throws (DomainError, ErrorException("Did not converge"),
PosDefException, ArgumentError); the library returns negative codes for bad calls and per-instance status[] otherwise.
Also the empty / degenerate inputs the reference's tests touch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rc(ila, name, *args):
    from ila_b200 import _lib
    return getattr(_lib.lib(), name)(*args)


def test_bad_arguments_return_codes(ila):
    from ila_b200 import _lib
    P = _lib.ptr
    col = np.array([2j, 0, 0, 1, 0.7])
    lam = np.array([1.0 + 1j]); L = np.empty(1, complex); st = np.empty(1, np.int32)
    # u != 1 is unsupported upstream as well (ql_pruneband undefined): -4
    assert _rc(ila, "ila_ql_toeplitz_l11_c64", 3, 2, P(col), P(lam), 1, 1, P(L), None, P(st), 1) == -4
    # null pointers / negative n: -2
    assert _rc(ila, "ila_ql_toeplitz_l11_c64", 3, 1, P(col), None, 1, 1, P(L), None, P(st), 1) == -2
    assert _rc(ila, "ila_ql_toeplitz_l11_c64", 3, 1, P(col), P(lam), -1, 1, P(L), None, P(st), 1) == -2
    # branch_rank out of range: -2
    assert _rc(ila, "ila_ql_toeplitz_l11_c64", 3, 1, P(col), P(lam), 1, 9, P(L), None, P(st), 1) == -2
    # zero super-diagonal (u would be 0): -2
    col0 = np.array([0j, 1, 1, 1, 1])
    assert _rc(ila, "ila_ql_toeplitz_l11_c64", 3, 1, P(col0), P(lam), 1, 1, P(L), None, P(st), 1) == -2
    assert "super-diagonal" in _lib.lib().ila_last_error().decode()
    # solve: right-hand side longer than 16 entries: -4
    x = np.empty(4, complex); b = np.ones(17, complex)
    assert _rc(ila, "ila_ql_toeplitz_solve_c64", 3, 1, P(col), P(lam), 1, 1, 17, P(b), 4, P(x), P(st), 1) == -4
    # block QL: block size 5: -4
    blk = np.zeros(25)
    e = np.zeros(1); Lr = np.empty(1); it = np.empty(1, np.int32)
    assert _rc(ila, "ila_ql_blocktri_l11_f64", 5, P(blk), P(blk), P(blk), 0, None, 0, None, 0, None, P(e), 1, 1e-10, 10,
               P(Lr), P(it), P(st), 1) == -4


def test_empty_batches_are_no_ops(ila):
    col = np.array([2j, 0, 0, 1, 0.7])
    L, st = ila.ql_toeplitz_l11(col, np.zeros(0, complex))
    assert L.shape == (0,) and st.shape == (0,)
    x, st = ila.ql_toeplitz_solve(col, np.zeros(0, complex), [1.0], 3)
    assert x.shape == (0, 3)
    L, it, st = ila.ql_blocktri_l11(*ila.periodic_schrodinger_blocks(), np.zeros(0))
    assert L.shape == (0,)
    z = ila.ql_toeplitz_absl11_grid(col, np.zeros(0), np.zeros(0))
    assert z.shape == (0, 0)


def test_status_instead_of_exceptions(ila):
    from ila_b200.api import BandedMatrix, DomainError, Fill, I, NotConvergedError, BlockTridiagonal, ql
    # DomainError cases of test/test_infql.jl:36-38 through the mirror API
    for Z, A, B in ((0.5, 2.1, 2.0), (0.5, -1.1, 2.0)):
        T = BandedMatrix({-1: Fill(Z), 0: Fill(A), 1: Fill(B)})
        with pytest.raises(DomainError):
            ql(T)
    # "Did not converge" (src/infql.jl:181) inside the essential spectrum of the periodic Schrödinger operator
    L, it, st = ila.ql_blocktri_l11(*ila.periodic_schrodinger_blocks(), [1.5], maxit=500)
    assert st[0] == 3 and it[0] == 500 and np.isnan(L[0])
    # capacity reached before convergence: status 3, ncols reports the capacity-limited sweep
    z = np.array([5000.0])
    out = ila.qr_banded_solve(1, 1, *ila.bessel_operator(z), b=np.array([[0.1]]), ncap=3000)
    assert out["status"][0] == 3
    # not positive definite: status 4 (PosDefException upstream)
    p = np.array([[1.0, -4.0, 1.0]]); q = np.zeros((1, 3))
    out = ila.chol_banded_solve(2, p, q, b=[1.0], ncap=4000)
    assert out["status"][0] == 4


def test_release_and_reuse(ila):
    from ila_b200 import _lib
    col = np.array([2j, 0, 0, 1, 0.7])
    L0, _ = ila.ql_toeplitz_l11(col, np.array([1 + 1j, 5.0]))
    _lib.check(_lib.lib().ila_release())
    L1, _ = ila.ql_toeplitz_l11(col, np.array([1 + 1j, 5.0]))       # pools are re-created on demand
    assert np.array_equal(L0, L1, equal_nan=True)
