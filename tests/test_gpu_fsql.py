"""Parity of the CUDA adaptive finite-section QL (K9) against the C oracle, through the C ABI.
Bar: bit-exact Lband / V / tau, identical N and status (IEEE, unfused, the reference's operation order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

C = lambda v: [v, 0, 0, 1, 0]


def _same(g, o):
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["N"], o["N"])
    for k in ("Lband", "V", "tau"):
        assert np.array_equal(g[k], o[k], equal_nan=True), k


def test_reference_fixtures_bit_exact(ila):
    from oracle import banded as ob
    fixtures = [
        (1, 1, [C(2.0), [0.25, 1, 0, 1, 1], C(1 / 3)], {}),                                  # test_infql.jl:210
        (1, 1, [[0, 1, 1, 1, 0], [0, 1.25, 1, 1, 0], C(1 / 3)], {}),                         # :224 (never converges)
        (1, 1, [C(1.0), C(3.0), C(1.0)],
         dict(head=np.array([[1.0, 2.0, 1.0], [1.0, 2.0, 3.0], [1.0, 2.0, 1.0]]), tol=1e-10, j=60)),   # :228-231
        (2, 2, [C(2.0), C(2.0), [4.0, 1, 0, 1, 1], C(1 / 3), C(1 / 3)], {}),                 # :263
    ]
    for l, u, g, kw in fixtures:
        o = ob.ql_finite_section(l, u, g, **kw)
        gg = ila.ql_finite_section(l, u, g, **kw)
        _same(gg, o)
    # the golden vector of test_infql.jl:219 on the GPU result itself
    gg = ila.ql_finite_section(1, 1, fixtures[0][2])
    Lb = gg["Lband"][0]
    L = np.zeros((8, 8))
    for i in range(8):
        for c in range(3):
            if i - c >= 0:
                L[i, i - c] = Lb[i, c]
    b = np.zeros(8); b[:3] = [1, 2, 3]
    assert np.array_equal((L @ b)[:6], [0.0, -5.25, -7.833333333333333, -2.4166666666666666, -1.0, 0.0])


def test_shift_family_and_bandwidths_bit_exact(ila):
    # L[1,1](λ) of a Jacobi-type operator without a Toeplitz tail over many shifts; other bandwidths
    from oracle import banded as ob
    lam = np.linspace(-6.0, -3.0, 200)
    g = [C(1.0), [0.0, 1, 0, 1, 1], C(1.0)]                  # diag 1/(k+1), off-diagonals 1
    o = ob.ql_finite_section(1, 1, g, shift=lam)
    gg = ila.ql_finite_section(1, 1, g, shift=lam)
    assert np.all(o["status"] == 0)
    _same(gg, o)
    rng = np.random.default_rng(4)
    for (l, u) in ((1, 2), (2, 1), (3, 1), (1, 3)):
        nd = l + u + 1
        g = [C(float(v)) for v in rng.normal(size=nd)]
        g[u] = [6.0, 1, 0, 2, 1]                               # 6 + 1/(k+2): dominant diagonal
        o = ob.ql_finite_section(l, u, g, shift=[0.0, 0.5, -0.5])
        gg = ila.ql_finite_section(l, u, g, shift=[0.0, 0.5, -0.5])
        assert np.all(o["status"] == 0), (l, u)
        _same(gg, o)


def test_mirror_api_reference_case(ila):
    # test/test_infql.jl:209-222 through the Julia-interface mirror: ql(A); (L*b)[1:6]; and the runaway protection :223-226
    from ila_b200.api import Affine, BandedMatrix, Fill, NotConvergedError, Rational, ql_adaptive
    A = BandedMatrix({1: Fill(2.0), 0: Rational(0.25, 1, 0, 1, 1), -1: Fill(1 / 3)})
    F = ql_adaptive(A)
    b = [1.0, 2.0, 3.0]
    Lb = [sum(F.L[i, jj] * b[jj - 1] for jj in range(1, 4)) for i in range(1, 7)]
    assert Lb == [0.0, -5.25, -7.833333333333333, -2.4166666666666666, -1.0, 0.0]
    qtb = F.Qt_mul(b)
    assert list(qtb[:2]) == [-0.0, -1.0]                       # (Q'*b)[1:2] == [-0., -1.]  (test_infql.jl:218)
    B = BandedMatrix({1: Affine(1.0, 1.0), 0: Affine(1.25, 1.0), -1: Fill(1 / 3)})
    with pytest.raises(NotConvergedError):
        ql_adaptive(B)


def test_complex_eltype_bit_exact(ila):
    # test/test_infql.jl:257-262 and a complex-shift family; same unfused operations as the oracle: bit for bit
    from oracle import banded as ob
    head = 1j * np.array([[1.0, 2.0, 1.0], [1.0, 2.0, 3.0], [1.0, 2.0, 1.0]])
    g = [C(1j), C(3j), C(1j)]
    _same(ila.ql_finite_section(1, 1, g, j=51, head=head), ob.ql_finite_section(1, 1, g, j=51, head=head))
    lam = np.linspace(-2.0, 2.0, 50) + 1.5j
    g = [C(1.0), [0.0, 1, 0, 1, 1], C(1.0)]
    o = ob.ql_finite_section(1, 1, g, shift=lam)
    gg = ila.ql_finite_section(1, 1, g, shift=lam)
    assert (o["status"] == 0).sum() >= 45 and np.iscomplexobj(gg["Lband"])     # (one shift sits where the sections do not settle)
    _same(gg, o)
