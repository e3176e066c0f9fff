"""The almost-banded adaptive QR oracle (oracle/almostbanded_qr.py, dense finite sections) against the reference's own tests
(test/test_infqr.jl:193-267): one-band, two-bands and more-bands operators with their closed-form solutions."""
import numpy as np
from scipy.special import gammaln

from oracle import almostbanded_qr as ab

ONES, ALT = [1, 1, 0, 0], [-1, 1, 0, 0]
INVFACT = np.exp(-gammaln(np.arange(1000) + 1.0))


def test_one_band():
    # A = Vcat(Ones(1,∞), BandedMatrix(0 => -Ones(∞), 1 => 1:∞)); (A \ [ℯ; zeros(∞)])[1:1000] ≈ 1/k!   (:194-221)
    o = ab.almostbanded_qr_solve(0, 1, [[1, 1, 0], [-1, 0, 0]], [ONES], [np.e], ncap=3200)
    assert o["status"] == 0 and o["nconv"] == 1001 and o["ncols"] == 2001
    assert np.abs(o["x"][:1000] - INVFACT).max() <= 1e-15
    # factors of the leading block against LAPACK's QR of the section (:203-212, R part; |.| as the signs are a convention)
    A = ab.dense_section(0, 1, [[1, 1, 0], [-1, 0, 0]], [ONES], 0.0, 11, 10)
    assert np.abs(np.abs(o["R"][:10, :10]) - np.abs(np.linalg.qr(A, mode="r"))).max() <= 1e-13


def test_two_bands():
    # B = BandedMatrix(0 => -Ones(∞), 2 => (1:∞).*(2:∞)); A = Vcat(Ones(1,∞), ((-1).^(0:∞))', B)   (:223-248)
    dg = [[2, 3, 1], [0, 0, 0], [-1, 0, 0]]
    o = ab.almostbanded_qr_solve(0, 2, dg, [ONES, ALT], [1.0], ncap=3200)
    x = 0.1
    assert abs(o["x"][:1000] @ x ** np.arange(1000) - (np.exp(1 - x) * (-1 + np.exp(2 + 2 * x))) / (-1 + np.exp(4))) <= 1e-15
    o = ab.almostbanded_qr_solve(0, 2, dg, [ONES, ALT], [np.e, 1 / np.e], ncap=3200)
    assert o["status"] == 0 and np.abs(o["x"][:1000] - INVFACT).max() <= 1e-15


def test_more_bands():
    # L = Vcat(Ones, (-1)^k, BandedMatrix(-1 => Ones, 1 => Ones, 2 => 4:2:∞, 3 => Ones, 5 => Ones)); L*u ≈ [1; 2; 0...]   (:250-266)
    dg = [[1, 0, 0], [0, 0, 0], [1, 0, 0], [4, 2, 0], [1, 0, 0], [0, 0, 0], [1, 0, 0]]
    A = ab.dense_section(1, 5, dg, [ONES, ALT], 0.0, 13, 19)
    assert A[1, 7] == -1                                               # F.data.data[2,8] == -1 (:254)
    o = ab.almostbanded_qr_solve(1, 5, dg, [ONES, ALT], [1.0, 2.0], ncap=4200)
    assert o["status"] == 0
    A = ab.dense_section(1, 5, dg, [ONES, ALT], 0.0, 1000, o["ncols"])
    rhs = np.zeros(1000); rhs[:2] = [1, 2]
    assert np.abs(A @ o["x"] - rhs).max() <= 1e-14
    assert ab.almostbanded_qr_solve(1, 5, dg, [ONES, ALT], [1.0, 2.0], ncap=900)["status"] == 3
