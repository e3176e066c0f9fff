"""Edge cases of the hot path through the C ABI: empty batches, capacity limits, extreme scalings, non-finite inputs, and
concurrent callers.  A result is either right (against the oracle) or flagged in status[] — never silently wrong."""
import threading

import numpy as np
import pytest

from conftest import BULLHEAD_COLUMN, bullhead_grid

pytestmark = pytest.mark.gpu


def _cmp(L, st, Lo, sto, scale, rtol=1e-11):
    assert np.array_equal(st == 0, sto == 0), (st[st != sto][:5], sto[st != sto][:5])
    ok = sto == 0
    if ok.any():
        err = np.abs(L[ok] - Lo[ok])
        assert err.max() <= 4e-12 * scale, err.max() / scale
        big = np.abs(Lo[ok]) >= 0.05 * scale
        if big.any():
            assert (err[big] / np.abs(Lo[ok][big])).max() <= rtol


@pytest.mark.parametrize("scale", [1e-150, 1e-8, 1.0, 1e6, 1e120])
def test_symbol_and_shift_scaling(ila, scale):
    """The factorization is homogeneous: ql(sA - sλ I).L[1,1] = s ql(A - λI).L[1,1].  Symbols and shifts scaled by 1e-150 … 1e120
    (the root finder works on w = μ/β, so the scale should not matter) against the oracle and against the unscaled result."""
    from oracle import toeplitz_ql as tq
    col = np.array(BULLHEAD_COLUMN, dtype=complex)
    lam = bullhead_grid(40, 30)
    L1, st1 = ila.ql_toeplitz_l11(col, lam)
    L, st = ila.ql_toeplitz_l11(col * scale, lam * scale)
    if 1e-100 <= scale <= 1e100:
        # (beyond, |μ|² = abs2.(λ) under/overflows in the reference's own arithmetic — the numpy/LAPACK oracle returns -5.15
        # instead of -4.75 at 1e-140 and DomainError at 1e150; the kernel works on w = μ/β and stays homogeneous)
        Lo, sto = tq.ql_toeplitz_l11(col * scale, lam * scale)
        _cmp(L, st, Lo, sto, 2.0 * scale)
    assert np.array_equal(st, st1)
    ok = st == 0
    assert np.allclose(L[ok] / scale, L1[ok], rtol=1e-11, atol=4e-12)


@pytest.mark.parametrize("beta", [1e-9, 1e-4, 1e4, 1e9])
def test_extreme_superdiagonal(ila, beta):
    """Tiny / huge super-diagonal coefficient β: the normalised roots w = μ/β span 1e±9 (FP32 scout range, FP64 fallback)."""
    from oracle import toeplitz_ql as tq
    rng = np.random.default_rng(3)
    col = np.array([beta * (1 + 0.5j), 0.3, -0.2j, 1.0, 0.7])
    lam = (rng.normal(size=400) + 1j * rng.normal(size=400)) * 3
    L, st = ila.ql_toeplitz_l11(col, lam)
    Lo, sto = tq.ql_toeplitz_l11(col, lam)
    assert not np.any(st == 3)
    _cmp(L, st, Lo, sto, max(abs(col[0]), 1.0), rtol=1e-10)


def test_non_finite_and_huge_shifts(ila):
    """NaN / Inf shifts must come back flagged (status != 0, NaN value), never as a number.  Shifts far outside the symbol's
    range (|λ| up to 1e100: one root of modulus 1e100, the others tiny — beyond the FP32 scout and the FP64 simultaneous
    iteration) are served by the dominant-root rescue (Newton on the reversed polynomial + Pellet's separation test)."""
    col = np.array(BULLHEAD_COLUMN, dtype=complex)
    lam = np.array([np.nan + 0j, np.inf + 0j, 1j * np.inf, complex(np.nan, np.nan), 5.0 + 0j, 1e100 + 0j, -1e120 + 1e120j])
    L, st = ila.ql_toeplitz_l11(col, lam)
    assert np.all(st[:4] != 0) and np.all(np.isnan(L[:4].real))
    assert st[4] == 0 and abs(L[4].real - (-4.747700587226856)) < 1e-12
    from oracle import toeplitz_ql as tq
    Lo, sto = tq.ql_toeplitz_l11(col, lam[5:])
    for k in range(2):
        if sto[k] == 0 and np.isfinite(Lo[k]):
            assert st[5 + k] == 0 and abs(L[5 + k] - Lo[k]) <= 1e-11 * abs(Lo[k])
        else:
            assert st[5 + k] != 0 or not np.isfinite(L[5 + k].real)
    Lr, s8 = ila.ql_toeplitz_l11_real(col, lam)
    assert np.array_equal(s8, st.astype(np.uint8)) and np.array_equal(ila.status_from_nan_payload(Lr)[:4], st[:4])
    # a sweep of magnitudes through the hand-over points scout -> FP64 iteration -> rescue, both branches of the phase
    mags = 10.0 ** np.arange(2, 101, 7)
    lam2 = np.concatenate([mags + 0j, 1j * mags, -(1 + 1j) * mags])
    L2, st2 = ila.ql_toeplitz_l11(col, lam2)
    Lo2, sto2 = tq.ql_toeplitz_l11(col, lam2)
    assert np.all(st2 == 0) and np.all(sto2 == 0)
    assert np.abs(L2 - Lo2).max() <= 1e-11 * np.abs(Lo2).max() and (np.abs(L2 - Lo2) / np.abs(Lo2)).max() <= 1e-11


def test_limits_and_empty_batches(ila):
    # every lower bandwidth the kernel is compiled for, and the first one it is not
    rng = np.random.default_rng(0)
    for l in range(1, 8):
        col = rng.normal(size=l + 2) + 1j * rng.normal(size=l + 2)
        L, st = ila.ql_toeplitz_l11(col, np.array([0.3 + 0.2j, 5.0]))
        assert L.shape == (2,)
    with pytest.raises(ila.IlaError) as e:
        ila.ql_toeplitz_l11(rng.normal(size=10) + 0j, np.array([0.3 + 0.2j]))
    assert e.value.code == -4
    with pytest.raises(ila.IlaError):                                   # u != 1 is ql_pruneband upstream: unsupported, loudly
        ila.lib().ila_ql_toeplitz_l11_c64  # noqa: B018
        from ila_b200 import _lib
        c = np.zeros(5, dtype=complex); lam = np.zeros(1, dtype=complex); L = np.zeros(1, dtype=complex); s = np.zeros(1, dtype=np.int32)
        _lib.check(ila.lib().ila_ql_toeplitz_l11_c64(2, 2, _lib.ptr(c), _lib.ptr(lam), 1, 1, _lib.ptr(L), None, _lib.ptr(s), 1))
    with pytest.raises(ila.IlaError):                                   # zero super-diagonal: u would be 0
        ila.ql_toeplitz_l11(np.array([0.0, 1.0, 2.0], dtype=complex), np.array([0.1 + 0j]))
    with pytest.raises(ila.IlaError):                                   # right-hand sides longer than 16 entries
        ila.ql_toeplitz_solve(BULLHEAD_COLUMN, np.array([5.0 + 0j]), np.ones(17), 4)
    # empty batches everywhere
    z = np.zeros(0)
    assert ila.ql_toeplitz_l11(BULLHEAD_COLUMN, np.zeros(0, dtype=complex))[0].shape == (0,)
    assert ila.ql_toeplitz_l11_real(BULLHEAD_COLUMN, np.zeros(0, dtype=complex))[0].shape == (0,)
    assert ila.ql_toeplitz_solve(BULLHEAD_COLUMN, np.zeros(0, dtype=complex), [1.0], 3)[0].shape == (0, 3)
    assert ila.ql_toeplitz_absl11_grid(BULLHEAD_COLUMN, z, z).shape == (0, 0)
    o = ila.qr_banded_solve(1, 1, np.zeros((0, 3)), np.zeros((0, 3)), b=np.zeros((0, 1)), ncap=100)
    assert len(o["x"]) == 0
    o = ila.chol_banded_solve(2, np.zeros((0, 3)), np.zeros((0, 3)), b=np.zeros((0, 1)), ncap=100)
    assert len(o["x"]) == 0
    assert ila.ql_blocktri_l11(*ila.periodic_schrodinger_blocks(), z)[0].shape == (0,)
    # a single shift, an odd count, a count that is not a multiple of anything
    for n in (1, 3, 1023, 4097):
        lam = bullhead_grid(n, 1).reshape(-1)
        L, st = ila.ql_toeplitz_l11(BULLHEAD_COLUMN, lam)
        Lr, s8 = ila.ql_toeplitz_l11_real(BULLHEAD_COLUMN, lam)
        assert np.array_equal(st.astype(np.uint8), s8)
        assert np.array_equal(np.nan_to_num(L.real), np.nan_to_num(Lr))


def test_concurrent_callers(ila):
    """Host entry points are serialised by the library's lock: eight threads calling different entries at once (ctypes drops
    the GIL) get the same answers as one thread."""
    from scipy.special import jv
    lam = bullhead_grid(200, 160)
    ref_L, ref_st = ila.ql_toeplitz_l11(BULLHEAD_COLUMN, lam)
    z = np.linspace(100.0, 900.0, 33)
    p, q, r = ila.bessel_operator(z)
    ref_x = ila.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap_per=ila.bessel_ncap(z))["xflat"]
    ref_g = ila.ql_toeplitz_absl11_grid(BULLHEAD_COLUMN, np.linspace(-10, 7, 200), np.linspace(-7, 8, 160))
    errs = []

    def work(kind):
        try:
            for _ in range(4):
                if kind == 0:
                    L, st = ila.ql_toeplitz_l11(BULLHEAD_COLUMN, lam)
                    assert np.array_equal(st, ref_st) and np.array_equal(np.nan_to_num(L), np.nan_to_num(ref_L))
                elif kind == 1:
                    x = ila.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap_per=ila.bessel_ncap(z))["xflat"]
                    assert np.array_equal(x, ref_x)
                elif kind == 2:
                    g = ila.ql_toeplitz_absl11_grid(BULLHEAD_COLUMN, np.linspace(-10, 7, 200), np.linspace(-7, 8, 160))
                    assert np.array_equal(g, ref_g)
                else:
                    Lr, s8 = ila.ql_toeplitz_l11_real(BULLHEAD_COLUMN, lam)
                    assert np.array_equal(s8, ref_st.astype(np.uint8))
        except Exception as ex:      # noqa: BLE001
            errs.append(repr(ex))
    th = [threading.Thread(target=work, args=(k % 4,)) for k in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs


def test_adaptive_solves_non_finite_and_scaled_inputs(ila):
    """K1+K2 / K5 edge cases against the C oracle, bit for bit: a NaN right-hand side (the reference's driver meets NaN in its
    convergence test — status 5 here), operators and right-hand sides scaled by 1e±100 (the sweeps are scale-free apart from
    the floatmin tolerance), a zero right-hand side (datasize 0), a right-hand side longer than one COLGROWTH chunk."""
    from oracle import banded as ob
    from scipy.special import jv
    z = np.array([200.0, 800.0])
    p, q, r = ila.bessel_operator(z)
    cases = {
        "nan rhs": dict(b=np.array([[np.nan], [0.3]])),
        "zero rhs": dict(b=np.zeros((2, 1))),
        "tiny rhs": dict(b=jv(1, z)[:, None] * 1e-200),
        "huge rhs": dict(b=jv(1, z)[:, None] * 1e200),
        "long rhs": dict(b=np.cos(np.arange(1500.0))[None, :].repeat(2, 0) * 0.5 ** np.arange(1500)[None, :] + 1e-3),
        "tolerance 1e-8": dict(b=jv(1, z)[:, None], tol=1e-8),
    }
    for name, kw in cases.items():
        g = ila.qr_banded_solve(1, 1, p, q, r, ncap=6000, **kw)
        o = ob.qr_banded_solve(1, 1, p, q, r, ncap=6000, **kw)
        assert np.array_equal(g["status"], o["status"]), name
        assert np.array_equal(g["ncols"], o["ncols"]) and np.array_equal(g["nconv"], o["nconv"]), name
        for i in range(2):
            n = int(o["ncols"][i])
            assert np.array_equal(g["x"][i], o["x"][i, :n], equal_nan=True), (name, i)
    assert ila.qr_banded_solve(1, 1, p, q, r, b=np.array([[np.nan], [0.3]]), ncap=6000)["status"][0] == 5
    for s in (1e-100, 1e100):                      # operator scaled: x scales by 1/s, counts unchanged
        g1 = ila.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap=6000)
        gs = ila.qr_banded_solve(1, 1, p * s, q * s, r, b=jv(1, z)[:, None], ncap=6000)
        os_ = ob.qr_banded_solve(1, 1, p * s, q * s, r, b=jv(1, z)[:, None], ncap=6000)
        assert np.array_equal(gs["status"], os_["status"]) and np.array_equal(gs["ncols"], os_["ncols"])
        for i in range(2):
            assert np.array_equal(gs["x"][i], os_["x"][i, :int(os_["ncols"][i])], equal_nan=True)
    # Cholesky: NaN shift / indefinite / scaled
    p4, q4 = ila.pentadiagonal_operator(np.array([0.5, 2.0]), eps=1e-2)
    for name, kw in {"plain": {}, "scaled rhs": dict(b=[1e-150]), "nan rhs": dict(b=[np.nan])}.items():
        kw = dict(dict(b=[1.0]), **kw)
        g = ila.chol_banded_solve(2, p4, q4, ncap=6000, **kw)
        o = ob.chol_banded_solve(2, p4, q4, ncap=6000, **kw)
        assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["ncols"], o["ncols"]), name
        for i in range(2):
            assert np.array_equal(g["x"][i], o["x"][i, :int(o["ncols"][i])], equal_nan=True), (name, i)
