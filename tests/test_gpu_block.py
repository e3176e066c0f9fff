"""Parity of the CUDA block-tridiagonal Toeplitz-tail QL (K6) against the numpy oracle, through the C ABI.
Bar: bit-exact L[1,1] and identical iteration counts (same operation order, no FMA, IEEE sqrt/div)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_periodic_schrodinger_bit_exact(ila):
    from oracle import block_ql as bq
    xx = np.concatenate([np.arange(-0.95, 0.951, 0.05), np.arange(2.25, 4.0001, 0.125 / 4)])
    blocks = ila.periodic_schrodinger_blocks()
    L, it, st = ila.ql_blocktri_l11(*blocks, xx)
    ref = [bq.periodic_schrodinger(x) for x in xx]
    assert np.all(st == 0)
    assert np.array_equal(it, [r[1] for r in ref])
    assert np.array_equal(L, [r[0] for r in ref])
    assert it.min() == 7 and it.max() == 51


def test_block_sizes_and_heads_bit_exact(ila):
    from oracle import block_ql as bq
    # degenerate 2x2 (test_periodic.jl:28-34)
    c2, a2, b2 = np.array([[0, 0.5], [0, 0]]), np.array([[0, 2.0], [0.5, 0]]), np.array([[0, 0.0], [2.0, 0]])
    L, it, st = ila.ql_blocktri_l11([c2], [a2], [b2], c2, a2, b2, [0.0])
    r = bq.blocktri_ql_l11([c2], [a2], [b2], c2, a2, b2)
    assert st[0] == 0 and it[0] == r[1] and L[0] == r[0] and abs(L[0]) <= 1e-11
    # 4x4 blocks (test_periodic.jl:42-60)
    B = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [0, 0, 0, 0], [0, 0, 0, 0]]) / 2
    A0 = np.array([[1, .5, .5, 0], [.5, -1, 0, .5], [.5, 0, -1, 0], [0, .5, 0, 1.0]])
    A = np.array([[1, 0, .5, 0], [0, -1, 0, .5], [.5, 0, -1, 0], [0, .5, 0, 1.0]])
    xs = np.array([0.0, 2.5, -2.5, 3.0])
    L, it, st = ila.ql_blocktri_l11([B], [A0], [B.T.copy()], B, A, B.T.copy(), xs)
    for i, x in enumerate(xs):
        r = bq.blocktri_ql_l11([B], [A0], [B.T.copy()], B, A, B.T.copy(), shift=x)
        assert st[i] == r[2]
        if r[2] == 0:
            assert it[i] == r[1] and L[i] == r[0]
    # scalar blocks (n = 1) with a longer head: a tridiagonal operator seen as 1x1 blocks
    rng = np.random.default_rng(5)
    dl = [np.array([[0.7]]), np.array([[1.1]])]
    d = [np.array([[v]]) for v in rng.normal(size=3)]
    du = [np.array([[0.4]])]
    c, a, b = np.array([[1.0]]), np.array([[0.2]]), np.array([[0.5]])
    xs = np.array([3.0, -3.5, 4.0 ])
    L, it, st = ila.ql_blocktri_l11(dl, d, du, c, a, b, xs)
    for i, x in enumerate(xs):
        r = bq.blocktri_ql_l11(dl, d, du, c, a, b, shift=x)
        assert st[i] == 0 and it[i] == r[1] and L[i] == r[0]


def test_not_converged_inside_spectrum(ila):
    # inside the essential spectrum the iteration never settles: "Did not converge" (src/infql.jl:181) -> status 3
    blocks = ila.periodic_schrodinger_blocks()
    L, it, st = ila.ql_blocktri_l11(*blocks, [1.5], maxit=2000)
    assert st[0] == 3 and np.isnan(L[0]) and it[0] == 2000


def test_c5_full_size_properties(ila):
    """BASELINE config C5 (512 energies in [-0.95,0.95] ∪ [2.25,4.0]): determinism, gap symmetry, finite-section check."""
    from oracle import toeplitz_ql as tq
    xs = np.concatenate([np.linspace(-0.95, 0.95, 256), np.linspace(2.25, 4.0, 256)])
    blocks = ila.periodic_schrodinger_blocks()
    L, it, st = ila.ql_blocktri_l11(*blocks, xs)
    L2, it2, st2 = ila.ql_blocktri_l11(*blocks, xs)
    assert np.all(st == 0) and np.array_equal(L, L2) and np.array_equal(it, it2)
    assert it.max() < 200
    dl, d, du = blocks[3], blocks[4], blocks[5]
    for i in (0, 255, 300, 511):
        nb, n = 120, 2
        A = np.zeros((nb * n, nb * n))
        for k in range(nb):
            A[k * n:(k + 1) * n, k * n:(k + 1) * n] = d - xs[i] * np.eye(n)
            if k < nb - 1:
                A[k * n:(k + 1) * n, (k + 1) * n:(k + 2) * n] = du
                A[(k + 1) * n:(k + 2) * n, k * n:(k + 1) * n] = dl
        A[0, 0] = 2.0 - xs[i]
        F, _ = tq.dense_ql(A)
        assert np.isclose(L[i], F[0, 0], atol=5e-9)


def test_complex_eltype_vs_oracle(ila):
    """Complex eltype (test/test_periodic.jl:20-26, `A - 5im*I`; complex blocks): the QL loop runs to k = 1.  Bar: same
    status and iteration count, L[1,1] within 1e-12 relative (complex multiply/divide orders differ at rounding level)."""
    from oracle import block_ql as bq
    c2, a2, b2 = np.array([[0, 0.5], [0, 0]]), np.array([[0, 2.0], [0.5, 0]]), np.array([[0, 0.0], [2.0, 0]])
    xs = np.array([5j, 2.0 + 4.0j, -0.5 - 3.0j, 3.0 + 0.5j, -4.0 + 0.0j])
    L, it, st = ila.ql_blocktri_l11([c2], [a2], [b2], c2, a2, b2, xs, maxit=5000)
    assert L.dtype == np.complex128
    for i, x in enumerate(xs):
        r = bq.blocktri_ql_l11([c2], [a2], [b2], c2, a2, b2, shift=x, maxit=5000)
        assert st[i] == r[2]
        if r[2] == 0:
            assert it[i] == r[1]
            assert abs(L[i] - r[0]) <= 1e-12 * abs(r[0])
    # complex blocks with a perturbed head, real energies
    rng = np.random.default_rng(11)
    c = np.array([[0, 1.0], [0, 0]]) * (1 + 0.3j)
    a = np.array([[-1.0, 1.0 + 0.2j], [1.0 - 0.2j, 1.0]])
    b = np.array([[0, 0], [1.0 - 0.3j, 0]])
    a0 = a + 0.1 * (rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))
    xs = np.array([4.0, -4.5, 3.5 + 1j])
    L, it, st = ila.ql_blocktri_l11([c], [a0], [b], c, a, b, xs, maxit=5000)
    for i, x in enumerate(xs):
        r = bq.blocktri_ql_l11([c], [a0], [b], c, a, b, shift=x, maxit=5000)
        assert st[i] == r[2] == 0 and it[i] == r[1]
        assert abs(L[i] - r[0]) <= 1e-12 * abs(r[0])


def test_block_factors_tau_and_qtb_vs_oracle(ila):
    """Row a12 in full: head factors, head τ, fixed point P, τ tail (src/infql.jl:199-204) and Q'b (:213-243) for the
    periodic Schrödinger operator over energies — real eltype bit-exact against the oracle; then the reference's own checks
    (test/test_periodic.jl:12-18) through the interface mirror."""
    from oracle import block_ql as bq
    blocks = ila.periodic_schrodinger_blocks()
    xs = np.concatenate([np.linspace(-0.95, 0.95, 9), np.linspace(2.25, 4.0, 8)])
    rhs = np.array([1.0, -2.0, 0.5])
    o = ila.ql_blocktri_factors(*blocks, xs, rhs=rhs, nout=8)
    assert np.all(o["status"] == 0)
    for i, x in enumerate(xs):
        L11, it, st, ex = bq.blocktri_ql_l11(*blocks, shift=x)
        assert st == 0 and it == o["iters"][i]
        assert np.array_equal(o["L11"][i], L11)
        assert np.array_equal(o["tail"][i], ex["tail_X"]), f"fixed point P differs at x = {x}"
        assert np.array_equal(o["tau_tail"][i], ex["tail_tau"])
        assert np.array_equal(o["head"][i], ex["head_factors"])
        assert np.array_equal(o["head_tau"][i], ex["head_tau"])
        F, tau = bq.assemble_inf_factors(ex, 2, ex["N"], 12)
        q = bq.qt_times_b(F, tau, rhs, 2, 8)
        assert np.array_equal(o["qtb"][i], q), f"Q'b differs at x = {x}"


def test_block_ql_object_like_test_periodic(ila):
    """test/test_periodic.jl:5-18 through the interface mirror: factors / τ / L against the 100-block finite section,
    Q[1:10,1:12] L[1:12,1:10] = A[1:10,1:10] (real, complex non-selfadjoint, degenerate and its adjoint-like case)."""
    from ila_b200.api import BlockTridiagonal, I, ql
    from oracle import block_ql as bq
    c = np.array([[0., 1.], [0., 0.]]); a = np.array([[-1., 1.], [1., 1.]]); b = np.array([[0., 0.], [1., 0.]])
    A = BlockTridiagonal(([c], c), ([a], a), ([b], b))
    F = ql(A)
    nb = 100
    Ad = np.zeros((2 * nb, 2 * nb))
    for k in range(nb):
        Ad[2 * k:2 * k + 2, 2 * k:2 * k + 2] = a
        if k + 1 < nb:
            Ad[2 * k:2 * k + 2, 2 * k + 2:2 * k + 4] = b
            Ad[2 * k + 2:2 * k + 4, 2 * k:2 * k + 2] = c
    Ff, tf = bq.dense_ql_seq(Ad)
    assert np.abs(F.factors[1:100, 1:100] - Ff[:100, :100]).max() < 1e-9        # Q̃.factors[1:100,1:100] ≈ Q.factors[1:100,1:100]
    assert np.abs(F.tau[1:100] - tf[:100]).max() < 1e-9                          # Q̃.τ[1:100] ≈ Q.τ[1:100]
    assert np.abs(F.L[1:100, 1:100] - np.tril(Ff)[:100, :100]).max() < 1e-9     # L ≈ L̃
    assert np.abs(F.Q[1:10, 1:12] @ F.L[1:12, 1:10] - Ad[:10, :10]).max() < 1e-9
    # complex non-selfadjoint (test_periodic.jl:20-26) and the degenerate real operator (:28-31)
    c2, a2, b2 = np.array([[0, 0.5], [0, 0]]), np.array([[0, 2.0], [0.5, 0]]), np.array([[0, 0.0], [2.0, 0]])
    for shift in (5j, 0.0):
        A2 = BlockTridiagonal(([c2], c2), ([a2], a2), ([b2], b2)) - shift * I
        F2 = ql(A2)
        Ad2 = np.zeros((12, 12), dtype=complex)
        for k in range(6):
            Ad2[2 * k:2 * k + 2, 2 * k:2 * k + 2] = a2 - shift * np.eye(2)
            if k + 1 < 6:
                Ad2[2 * k:2 * k + 2, 2 * k + 2:2 * k + 4] = b2
                Ad2[2 * k + 2:2 * k + 4, 2 * k:2 * k + 2] = c2
        assert np.abs(F2.Q[1:10, 1:12] @ F2.L[1:12, 1:10] - Ad2[:10, :10]).max() < 1e-9
    assert abs(F2.L[1, 1]) <= 1e-11                                              # degenerate


def test_c5_full_batch_bit_exact_vs_oracle(ila):
    """BASELINE config C5 at full size: all 512 energies of the periodic Schrödinger operator (256 in (-0.95, 0.95), 256 in
    (2.25, 4)) against the oracle: L[1,1], iteration counts and statuses identical, fixed point and τ bit for bit."""
    import json, os
    from oracle import block_ql as bq
    xs = np.concatenate([np.linspace(-0.95, 0.95, 256), np.linspace(2.25, 4.0, 256)])
    blocks = ila.periodic_schrodinger_blocks()
    o = ila.ql_blocktri_factors(*blocks, xs)
    bad = dict(status=0, iters=0, L11=0, tail=0, tau=0, head=0)
    for i, x in enumerate(xs):
        L11, it, st, ex = bq.blocktri_ql_l11(*blocks, shift=x)
        bad["status"] += int(st != o["status"][i])
        if st != 0:
            continue
        bad["iters"] += int(it != o["iters"][i])
        bad["L11"] += int(L11 != o["L11"][i])
        bad["tail"] += int(not np.array_equal(o["tail"][i], ex["tail_X"]))
        bad["tau"] += int(not (np.array_equal(o["tau_tail"][i], ex["tail_tau"]) and np.array_equal(o["head_tau"][i], ex["head_tau"])))
        bad["head"] += int(not np.array_equal(o["head"][i], ex["head_factors"]))
    rep = dict(config="C5 512 periodic-Schrödinger energies", mismatches=bad, iterations_min_max=[int(o["iters"].min()), int(o["iters"].max())],
               converged=int((o["status"] == 0).sum()))
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        json.dump(rep, open(os.path.join(root, "gpurun_out", "r02_parity_c5_full.json"), "w"), indent=1)
    except OSError:
        pass
    assert all(v == 0 for v in bad.values()), rep
    assert rep["converged"] == 512, rep
