"""Parity of the CUDA adaptive reverse Cholesky (K8) against the C oracle, through the C ABI.
Bar: bit-exact la / lb, identical N and status (IEEE, unfused, the reference's operation order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FILL = lambda v: [v, 0, 0, 1, 0]
RECIP = lambda c0: [c0, 1, 0, 2, 1]


def _same(g, o):
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["N"], o["N"])
    assert np.array_equal(g["la"], o["la"], equal_nan=True) and np.array_equal(g["lb"], o["lb"], equal_nan=True)


def test_reference_fixtures_bit_exact(ila):
    from oracle import banded as ob
    for ga, gb, kw in ((FILL(3), FILL(1), {}), (FILL(1), RECIP(0), {}), (RECIP(5), FILL(1), {}),
                       (FILL(3), FILL(1), dict(a_head=[4.0, 5.0, 6.0], b_head=[2.0, 3.0, 1.0]))):
        o = ob.revchol_tridiag(ga, gb, 1000, **kw)
        g = ila.revchol_tridiag(ga, gb, 1000, **kw)
        assert o["status"][0] == 0
        _same(g, o)
    # (L'L)[1:1000,1:1000] ≈ A  (test/test_infreversecholesky.jl:22-23) on the GPU result itself
    g = ila.revchol_tridiag(FILL(3), FILL(1), 1000)
    la, lb = g["la"][0], g["lb"][0]
    assert np.abs(la ** 2 + lb ** 2 - 3).max() <= 1e-15 and np.abs(lb[:-1] * la[1:] - 1).max() <= 1e-15


def test_shift_family_batch_bit_exact(ila):
    # 300 shifted operators: the decay slows (N grows) as the diagonal approaches 2|b|
    from oracle import banded as ob
    shift = np.linspace(0.0, 0.995, 300)                     # A - shift*I: diagonal 3 - shift -> 2.005
    o = ob.revchol_tridiag(FILL(3), FILL(1), 64, shift=shift)
    g = ila.revchol_tridiag(FILL(3), FILL(1), 64, shift=shift)
    assert np.all(o["status"] == 0) and o["N"].max() > o["N"].min()
    _same(g, o)


def test_status_codes(ila):
    from oracle import banded as ob
    for ga, gb, n, want in ((FILL(1), FILL(1), 8, 4), (FILL(2.0005), FILL(1), 16, 0), (FILL(2.0), FILL(1), 4, 3)):
        g = ila.revchol_tridiag(ga, gb, n)
        _same(g, ob.revchol_tridiag(ga, gb, n))
        assert g["status"][0] == want


def test_mirror_api_reference_cases(ila):
    # test/test_infreversecholesky.jl:5-12, 53-59 through the Julia-interface mirror
    from ila_b200.api import Fill, I, Ones, Rational, SymTridiagonal, Vcat, reversecholesky

    def check(A_dense, F, n):
        L = np.array([[F.L[i, j] for j in range(1, n + 2)] for i in range(1, n + 2)])
        assert np.allclose((L.T @ L)[:n, :n], A_dense[:n, :n], rtol=0, atol=1e-14)

    n = 60
    k = np.arange(n + 1)
    tri = lambda d, e: np.diag(d) + np.diag(e[:-1], 1) + np.diag(e[:-1], -1)
    check(tri(np.full(n + 1, 3.0), np.ones(n + 1)), reversecholesky(SymTridiagonal(Fill(3.0), Fill(1.0))), n)
    a = np.full(n + 1, 3.0); a[:3] = [4, 5, 6]
    b = np.ones(n + 1); b[:2] = [2, 3]
    check(tri(a, b), reversecholesky(SymTridiagonal(Vcat([4.0, 5.0, 6.0], Fill(3.0)), Vcat([2.0, 3.0], Fill(1.0)))), n)
    check(tri(np.ones(n + 1), 1 / (k + 2.0)), reversecholesky(SymTridiagonal(Ones(), Rational(0, 1, 0, 2, 1))), n)
    check(tri(5 + 1 / (k + 2.0), np.ones(n + 1)), reversecholesky(5 * I + SymTridiagonal(Rational(0, 1, 0, 2, 1), Ones())), n)
