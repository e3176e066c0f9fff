"""Parity of the CUDA adaptive almost-banded QR solve (K12) against the dense-section oracle, through the C ABI.
Bar: 1e-12 relative to max|x| (the kernel carries the fill as coefficient vectors, the oracle factorizes dense sections);
ncols / nconv / status identical."""
import numpy as np
import pytest
from scipy.special import gammaln

pytestmark = pytest.mark.gpu

ONES, ALT = [1, 1, 0, 0], [-1, 1, 0, 0]
INVFACT = np.exp(-gammaln(np.arange(1000) + 1.0))


def _check(g, i, o):
    assert g["status"][i] == o["status"] and g["ncols"][i] == o["ncols"] and g["nconv"][i] == o["nconv"], (i, g["status"][i], g["ncols"][i], g["nconv"][i], o["status"], o["ncols"], o["nconv"])
    if o["status"] == 0:
        n = o["ncols"]
        assert np.abs(g["x"][i, :n] - o["x"]).max() <= 1e-12 * np.abs(o["x"]).max()
        assert np.all(g["x"][i, n:] == 0)


def test_reference_operators(ila):
    from oracle import almostbanded_qr as ab
    # one-band (test_infqr.jl:194-221)
    g = ila.qr_almostbanded_solve(0, 1, [[1, 1, 0], [-1, 0, 0]], [ONES], [np.e], ncap=3200)
    _check(g, 0, ab.almostbanded_qr_solve(0, 1, [[1, 1, 0], [-1, 0, 0]], [ONES], [np.e], ncap=3200))
    assert np.abs(g["x"][0, :1000] - INVFACT).max() <= 1e-15
    # two-bands (:223-248)
    dg = [[2, 3, 1], [0, 0, 0], [-1, 0, 0]]
    g = ila.qr_almostbanded_solve(0, 2, dg, [ONES, ALT], [[1.0, 0.0], [np.e, 1 / np.e]], ncap=3200)
    _check(g, 0, ab.almostbanded_qr_solve(0, 2, dg, [ONES, ALT], [1.0, 0.0], ncap=3200))
    _check(g, 1, ab.almostbanded_qr_solve(0, 2, dg, [ONES, ALT], [np.e, 1 / np.e], ncap=3200))
    x = 0.1
    assert abs(g["x"][0, :1000] @ x ** np.arange(1000) - (np.exp(1 - x) * (-1 + np.exp(2 + 2 * x))) / (-1 + np.exp(4))) <= 1e-15
    assert np.abs(g["x"][1, :1000] - INVFACT).max() <= 1e-15
    # more bands (:250-266): L*u ≈ [1; 2; 0...]
    dg = [[1, 0, 0], [0, 0, 0], [1, 0, 0], [4, 2, 0], [1, 0, 0], [0, 0, 0], [1, 0, 0]]
    g = ila.qr_almostbanded_solve(1, 5, dg, [ONES, ALT], [1.0, 2.0], ncap=4200)
    o = ab.almostbanded_qr_solve(1, 5, dg, [ONES, ALT], [1.0, 2.0], ncap=4200)
    _check(g, 0, o)
    A = ab.dense_section(1, 5, dg, [ONES, ALT], 0.0, 1000, o["ncols"])
    rhs = np.zeros(1000); rhs[:2] = [1, 2]
    assert np.abs(A @ g["x"][0, :o["ncols"]] - rhs).max() <= 1e-14


def test_shift_family_tolerance_capacity(ila):
    # u' = λ u, u(1) = e^λ: Vcat(Ones, BandedMatrix(0 => -λ, 1 => 1:∞)) \ [e^λ; 0...] = λ^k / k!, 40 values of λ
    from oracle import almostbanded_qr as ab
    lam = np.linspace(-2.0, 3.0, 40)
    dg = [[1, 1, 0], [0, 0, 0]]
    g = ila.qr_almostbanded_solve(0, 1, dg, [ONES], np.exp(lam)[:, None], shift=lam, ncap=3200)
    assert np.all(g["status"] == 0)
    k = np.arange(150)
    for i in (0, 13, 16, 39):
        _check(g, i, ab.almostbanded_qr_solve(0, 1, dg, [ONES], [np.exp(lam[i])], shift=lam[i], ncap=3200))
    for i in range(40):
        ref = np.sign(lam[i]) ** k * np.exp(k * np.log(np.abs(lam[i])) - gammaln(k + 1.0)) if lam[i] != 0 else (k == 0) * 1.0
        assert np.abs(g["x"][i, :150] - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), i
    # looser tolerance and shorter chunks; capacity reached -> status 3
    kw = dict(tol=1e-20, colgrowth=100, ncap=700)
    g = ila.qr_almostbanded_solve(0, 1, dg, [ONES], [np.e], shift=[1.0], **kw)
    _check(g, 0, ab.almostbanded_qr_solve(0, 1, dg, [ONES], [np.e], shift=1.0, **kw))
    g = ila.qr_almostbanded_solve(0, 1, [[1, 1, 0], [-1, 0, 0]], [ONES], [np.e], ncap=900)
    assert g["status"][0] == 3 and g["ncols"][0] == 0


def test_generic_shape_and_quadratic_boundary_row(ila):
    # (lD, uD, r) = (1, 2, 1) is not one of the register-resident shapes: the runtime-shape kernel; boundary row (-1)^j (1 + j/2)
    from oracle import almostbanded_qr as ab
    dg = [[0.3, 0, 0], [1, 1, 0], [-2, 0, 0], [0.4, 0, 0]]
    g = ila.qr_almostbanded_solve(1, 2, dg, [ONES], [1.0], ncap=3300)
    _check(g, 0, ab.almostbanded_qr_solve(1, 2, dg, [ONES], [1.0], ncap=3300))
    bg = [[-1, 1, 0.5, 0]]
    g = ila.qr_almostbanded_solve(1, 2, dg, bg, [[1.0, -0.3], [0.5, 2.0]], shift=[0.7, -0.2], ncap=3300)
    _check(g, 0, ab.almostbanded_qr_solve(1, 2, dg, bg, [1.0, -0.3], shift=0.7, ncap=3300))
    _check(g, 1, ab.almostbanded_qr_solve(1, 2, dg, bg, [0.5, 2.0], shift=-0.2, ncap=3300))
