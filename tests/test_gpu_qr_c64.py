"""Parity of the complex-eltype CUDA adaptive QR solve (qr_banded_c64.cu) against the C oracle, through the C ABI.
Bar: bit-exact (same unfused operations in the same order), identical ncols / nconv / status."""
import numpy as np
import pytest
from scipy.special import jv

pytestmark = pytest.mark.gpu


def _same(o, g, batch):
    assert np.array_equal(g["status"], o["status"])
    assert np.array_equal(g["ncols"], o["ncols"]) and np.array_equal(g["nconv"], o["nconv"])
    for i in range(batch):
        n = int(o["ncols"][i])
        xg = np.asarray(g["x"][i])
        assert xg.shape[0] == n
        assert np.array_equal(xg.view(np.float64), o["x"][i][:n].copy().view(np.float64)) or \
            np.array_equal(xg, o["x"][i][:n])          # (-0.0 == +0.0 allowed)


def test_complex_shift_batch_bit_exact(ila):
    from oracle import banded as ob
    rng = np.random.default_rng(1)
    lam = np.concatenate([rng.normal(size=40) + 1j * (0.2 + rng.random(40)), [2.0 + 1.5j, 1j, -0.3 + 0.05j]])
    batch = lam.size
    p = np.tile(np.array([0.5, 0.0, 0.5]), (batch, 1)); q = np.zeros((batch, 3))
    b = np.tile(np.array([1.0, 0.5j]), (batch, 1))
    o = ob.qr_banded_solve_c64(1, 1, p, q, None, shift=lam, b=b, ncap=40000)
    g = ila.qr_banded_solve_c64(1, 1, p, q, None, shift=lam, b=b, ncap=40000)
    assert np.all(o["status"] == 0)
    _same(o, g, batch)


def test_real_limit_matches_real_kernel(ila):
    z = np.array([300.0, 57.3, 1000.0, 2500.0])
    p, q, r = ila.bessel_operator(z)
    gr = ila.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap=8000)
    gc = ila.qr_banded_solve_c64(1, 1, p, q, r, b=jv(1, z)[:, None], ncap=8000)
    assert np.array_equal(gr["ncols"], gc["ncols"]) and np.array_equal(gr["nconv"], gc["nconv"])
    for i in range(4):
        assert np.array_equal(np.asarray(gc["x"][i]).real, gr["x"][i]) and np.all(np.asarray(gc["x"][i]).imag == 0)


def test_complex_generators_bandwidths_heads_capacity(ila):
    from oracle import banded as ob
    rng = np.random.default_rng(2)
    for (l, u) in ((1, 2), (2, 1), (2, 2), (3, 1), (1, 3), (3, 3)):
        nd = l + u + 1
        batch = 5
        p = rng.normal(size=(batch, nd)) + 1j * rng.normal(size=(batch, nd))
        p[:, u] += 6.0 + 2.0j                                   # diagonally dominant
        q = np.zeros((batch, nd), complex); q[:, u] = 0.01 * (1 + 0.5j)
        head = (rng.normal(size=(nd, 3)) + 1j * rng.normal(size=(nd, 3))); head[u] += 7.0
        b = rng.normal(size=(batch, 4)) + 1j * rng.normal(size=(batch, 4))
        caps = np.array([6000, 5000, 7000, 6500, 5500])
        o = ob.qr_banded_solve_c64(l, u, p, q, None, head=head, b=b, ncap=7000)
        g = ila.qr_banded_solve_c64(l, u, p, q, None, head=head, b=b, ncap_per=caps)
        assert np.all(o["status"] == 0), (l, u)
        _same(o, g, batch)
    # capacity reached before convergence: status 3, empty solution
    g = ila.qr_banded_solve_c64(1, 1, np.array([[0.5, 0.0, 0.5]]), np.zeros((1, 3)), None, shift=[0.01j], b=[1.0], ncap=1500)
    assert g["status"][0] == 3 and g["ncols"][0] == 0


def test_qr_and_ql_resolvents_agree_at_complex_shifts(ila):
    # test/test_infql.jl:143-145 ((qr(A) \ b)[1:1000] ≈ (ql(A) \ b)[1:1000]) carried to complex shifts through the mirror API
    from ila_b200.api import BandedMatrix, Fill, I, Vcat, Zeros, ql, qr
    A = BandedMatrix({-1: Fill(2.0), 0: Fill(5.0), 1: Fill(0.5)})
    for lam in (0.5 + 0.25j, -1.0 - 2.0j):
        xq = qr(A - lam * I).solve(Vcat([1.0], Zeros()))
        xl = ql(A - lam * I).solve(Vcat([1.0], Zeros()), n=300)
        m = min(300, xq.shape[0])
        assert np.iscomplexobj(xq) and m >= 100
        assert np.abs(xq[:m] - xl[:m]).max() <= 1e-13 * np.abs(xl).max()
