"""The oracle still reproduces the committed golden fixtures (tests/golden/, made by make_golden.py). CPU only.
C-oracle results (QR, Cholesky) and the sequential numpy block QL must match bit for bit; LAPACK-backed ones to 1e-12."""
import os

import numpy as np
from scipy.special import jv

from oracle import banded as ob
from oracle import block_ql as bq
from oracle import toeplitz_ql as tq

G = os.path.join(os.path.dirname(__file__), "golden")


def test_bessel_and_cholesky_fixtures_bit_exact():
    g = np.load(os.path.join(G, "bessel_z1000.npz"))
    p, q, r = ob.bessel_operator(float(g["z"]))
    o = ob.qr_banded_solve(1, 1, p, q, r, b=[jv(1, float(g["z"]))], ncap=8000)
    n = int(g["ncols"])
    assert int(o["ncols"][0]) == n and int(o["nconv"][0]) == int(g["nconv"])
    assert np.array_equal(o["x"][0, :n], g["x"])
    c = np.load(os.path.join(G, "chol_pentadiag.npz"))
    p4 = np.tile(np.array([1.0, -4.0, 6.0]), (4, 1)); p4[:, 2] += c["sigma"]
    q4 = np.zeros((4, 3)); q4[:, 2] = 1e-4
    o = ob.chol_banded_solve(2, p4, q4, b=[1.0], ncap=8000)
    assert np.array_equal(o["ncols"], c["ncols"]) and np.array_equal(o["nconv"], c["nconv"])
    assert np.array_equal(o["x"][:, :c["x"].shape[1]], c["x"])


def test_block_fixture_bit_exact():
    g = np.load(os.path.join(G, "blocktri.npz"))
    res = [bq.periodic_schrodinger(x) for x in g["energies"]]
    assert np.array_equal([r[1] for r in res], g["iters"]) and np.array_equal([r[0] for r in res], g["L11"])


def test_toeplitz_fixtures():
    g = np.load(os.path.join(G, "bullhead_64x64.npz"))
    L, st = tq.ql_toeplitz_l11(np.array([2j, 0, 0, 1, 0.7]), g["lam"])
    assert np.array_equal(st, g["status"])
    ok = st == 0
    assert np.allclose(L[ok], g["L11"][ok], rtol=1e-12, atol=2e-12)
    s = np.load(os.path.join(G, "ql_solve.npz"))
    x, st = tq.ql_toeplitz_solve(s["bull_col"], s["bull_lam"], s["bull_b"], 32)
    assert np.array_equal(st, s["bull_status"])
    ok = st == 0
    assert np.allclose(x[ok], s["bull_x"][ok], rtol=1e-12, atol=1e-13) and np.all(np.isnan(x[~ok]))


def test_triconj_and_blockqr_fixtures():
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from triconj_cases import cholesky_case, reference_pairs
    from oracle import block_qr as bqr
    g = np.load(os.path.join(G, "triconj.npz"))
    pairs = reference_pairs(203)
    assert np.array_equal(ob.tridiagonal_conjugation(pairs[1][1], pairs[1][2], 200)[0], g["Y_P_C32"])
    assert np.array_equal(ob.tridiagonal_conjugation(pairs[2][1], pairs[2][2], 200)[0], g["Y_P10_P20"])
    b = np.load(os.path.join(G, "blockqr.npz"))
    T = bqr.toeplitz_bands(0.0, 0.5, 80)
    for i, c in enumerate(b["cases"]):
        o = bqr.block_qr_solve(T, T, *c)
        assert o["ncols"] == b["ncols"][i] and o["nblocks"] == b["nblocks"][i]
        assert np.abs(o["x"][:300] - b["x"][i]).max() <= 1e-13 and np.abs(o["qtb"][:300] - b["qtb"][i]).max() <= 1e-13


def test_almostbanded_fixture():
    from oracle import almostbanded_qr as abq
    g = np.load(os.path.join(G, "almostbanded.npz"))
    o1 = abq.almostbanded_qr_solve(0, 1, [[1, 1, 0], [-1, 0, 0]], [[1, 1, 0, 0]], [np.e], ncap=3200)
    assert o1["ncols"] == g["ncols"][0] and o1["nconv"] == g["nconv"][0] and np.abs(o1["x"][:400] - g["x1"]).max() <= 1e-15
