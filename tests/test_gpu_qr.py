"""Parity of the CUDA adaptive banded QR solve (K1+K2) against the C oracle, through the C ABI.

Bar: BIT-EXACT against the oracle (the kernel issues the reference's operations in the reference's order with no
fused multiply-add; IEEE sqrt/div), plus closed-form and residual checks with the conditioning-aware bound."""
import numpy as np
import pytest
from scipy.special import jv

pytestmark = pytest.mark.gpu


def _compare(g, o, batch):
    assert np.array_equal(g["status"], o["status"])
    assert np.array_equal(g["ncols"], o["ncols"])
    assert np.array_equal(g["nconv"], o["nconv"])
    for i in range(batch):
        n = int(o["ncols"][i])
        assert np.array_equal(g["x"][i], o["x"][i, :n]), f"instance {i} not bit-exact"


def test_bessel_batch_bit_exact(ila):
    from oracle import banded as ob
    zs = np.concatenate([[10.0, 100.0, 1000.0, 10000.0], np.linspace(1e3, 2e4, 60)])
    p, q, r = ila.bessel_operator(zs)
    b = jv(1, zs)[:, None]
    caps = ila.bessel_ncap(zs)
    g = ila.qr_banded_solve(1, 1, p, q, r, b=b, ncap_per=caps)
    o = ob.qr_banded_solve(1, 1, p, q, r, b=b, ncap=int(caps.max()))
    # the oracle ran with the largest capacity; statuses must agree because every instance converges
    assert np.all(o["status"] == 0)
    _compare(g, o, len(zs))
    # README.md:41-66 / test_infqr.jl:93-105
    i = 3
    x = g["x"][i]
    ref = jv(np.arange(len(x)), 1e4)
    assert np.max(np.abs(x - ref)) <= 2e-11 * np.max(np.abs(ref))
    assert abs(x[999] - jv(999, 1e4)) < 5e-14


def test_factors_export_bit_exact(ila):
    from oracle import banded as ob
    # test/test_infqr.jl:12-41
    args = dict(b=[1., 2, 3], ncap=5000, want_factors=True)
    g = ila.qr_banded_solve(1, 1, [[1, 1, 1]], [[0, 1, 0]], **args)
    o = ob.qr_banded_solve(1, 1, [[1, 1, 1]], [[0, 1, 0]], **args)
    n = int(o["ncols"][0])
    assert np.array_equal(g["x"][0], o["x"][0, :n])
    assert np.array_equal(g["factors"][0], o["factors"][0, :n])
    assert np.array_equal(g["tau"][0], o["tau"][0, :n])
    assert np.array_equal(g["qtb"][0], o["qtb"][0, :n])
    assert np.isclose(g["factors"][0][0, 2], -np.sqrt(2), rtol=1e-15)
    assert np.isclose(g["tau"][0][0], 1 + np.sqrt(2) / 2, rtol=1e-15)


def test_five_band_and_other_bandwidths(ila):
    from oracle import banded as ob
    head = np.zeros((5, 3))
    head[0, :] = 1
    head[1, :] = [1, 0, 0]
    head[2, :] = [4, 5, 6]
    head[3, :] = [1, 0, 0]
    head[4, :] = 1
    kw = dict(head=head, b=[3., 4, 5], ncap=8000)
    g = ila.qr_banded_solve(2, 2, [[1, 0, 3, 0, 1]], [[0] * 5], **kw)
    o = ob.qr_banded_solve(2, 2, [[1, 0, 3, 0, 1]], [[0] * 5], **kw)
    _compare(g, o, 1)
    rng = np.random.default_rng(7)
    for (l, u) in [(1, 2), (2, 1), (3, 1), (1, 3), (3, 3), (1, 0)]:
        nd = l + u + 1
        batch = 5
        p = rng.normal(size=(batch, nd))
        p[:, u] += 6.0 * np.sign(p[:, u])                 # diagonally dominant main diagonal
        q = np.zeros((batch, nd))
        q[:, u] = 0.01 * np.sign(p[:, u])
        shift = rng.normal(size=batch) * 0.1
        b = rng.normal(size=(batch, 4))
        g = ila.qr_banded_solve(l, u, p, q, shift=shift, b=b, ncap=9000)
        o = ob.qr_banded_solve(l, u, p, q, shift=shift, b=b, ncap=9000)
        assert np.all(o["status"] == 0)
        _compare(g, o, batch)


def test_tolerance_ncap_and_ragged_capacity(ila):
    from oracle import banded as ob
    z = np.array([1000.0, 1500.0])
    p, q, r = ila.bessel_operator(z)
    b = jv(1, z)[:, None]
    g = ila.qr_banded_solve(1, 1, p, q, r, b=b, ncap=8000, tol=1e-8)
    o = ob.qr_banded_solve(1, 1, p, q, r, b=b, ncap=8000, tol=1e-8)
    _compare(g, o, 2)
    g = ila.qr_banded_solve(1, 1, p, q, r, b=b, ncap=1500)         # capacity reached before convergence
    assert list(g["status"]) == [3, 3] and list(g["ncols"]) == [0, 0]
    with pytest.raises(ila.IlaError) as e:
        ila.qr_banded_solve(1, 1, p, q, r, b=b, ncap=8000, xcap=100)
    assert e.value.code == -5


def test_c2_full_size_properties(ila):
    """BASELINE config C2 at full size (4096 z in [1e3,1e5]): residual of the recurrence and closed form on samples."""
    zs = 1e3 + (1e5 - 1e3) * np.arange(4096) / 4095
    p, q, r = ila.bessel_operator(zs)
    b = jv(1, zs)[:, None]
    caps = ila.bessel_ncap(zs)
    g = ila.qr_banded_solve(1, 1, p, q, r, b=b, ncap_per=caps)
    assert np.all(g["status"] == 0)
    assert np.all(g["ncols"] == g["nconv"] + 1000)            # n = n_conv + COLGROWTH - 1 + l
    assert np.all(np.diff(g["nconv"]) >= 0)                   # monotone in z
    for i in (0, 1000, 4095):
        z = zs[i]
        x = g["x"][i]
        k = np.arange(1, len(x) - 1)
        # three-term recurrence residual J_{k-1} - (2k/z) J_k + J_{k+1} = 0, and x[1] = J_1(z)
        res = x[:-2] - (2.0 * k / z) * x[1:-1] + x[2:]
        assert np.max(np.abs(res)) <= 1e-13 * max(1.0, np.max(np.abs(x)))
        assert x[1] == pytest.approx(jv(1, z), rel=1e-12)
        ref = jv(np.arange(len(x)), z)
        assert np.max(np.abs(x - ref)) <= 4 * z * np.finfo(float).eps * np.max(np.abs(ref)) + 1e-13


def test_c2_full_batch_bit_exact_vs_oracle(ila):
    """BASELINE config C2 at full size: all 4096 solves (z up to 1e5, 2.2e8 columns) against the C oracle BIT for bit, and
    the κ-aware comparison with the __float128 oracle (back-substitution amplifies rounding ∝ z)."""
    import json, os
    import parity_full
    rep = parity_full.c2_full_report(ila)
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        json.dump(rep, open(os.path.join(root, "gpurun_out", "r02_parity_c2_full.json"), "w"), indent=1)
    except OSError:
        pass
    assert rep["all_converged"], rep
    assert rep["bit_exact"], rep
    assert rep["max_rel_err_over_z_eps"] <= 4.0, rep


def test_output_paths_identical(ila):
    """The back-substitution delivers x either straight into the ragged output (small batches) or through the interleaved
    stream + compaction (large batches): both must give the same bits, including the zero region up to datasize."""
    from ila_b200 import _lib
    from scipy.special import jv
    z = np.linspace(50.0, 3000.0, 77)
    p, q, r = ila.bessel_operator(z)
    outs = []
    for mode in (3, 4, 2):
        _lib.check(_lib.lib().ila_qr_set_arith_mode(mode))
        o = ila.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap_per=ila.bessel_ncap(z))
        outs.append((o["xflat"].copy(), o["ncols"].copy(), o["status"].copy()))
    _lib.check(_lib.lib().ila_qr_set_arith_mode(2))
    for k in (1, 2):
        assert np.array_equal(outs[0][1], outs[k][1]) and np.array_equal(outs[0][2], outs[k][2])
        assert np.array_equal(outs[0][0].view(np.int64), outs[k][0].view(np.int64))


def test_bulk_and_ldgsts_loaders_identical(ila):
    """The back-substitution ring is filled either by per-lane 16-byte cp.async (LDGSTS) or by one cp.async.bulk per group
    completing on an mbarrier (default): same bits, for plain records and for records that carry reflectors."""
    from ila_b200 import _lib
    from scipy.special import jv
    z = np.linspace(50.0, 6000.0, 70)
    p, q, r = ila.bessel_operator(z)
    outs = {}
    try:
        for mode in (5, 6):
            _lib.check(_lib.lib().ila_qr_set_arith_mode(mode))
            for wf in (False, True):
                o = ila.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap_per=ila.bessel_ncap(z), want_factors=wf)
                outs[(mode, wf)] = (o["xflat"].copy(), o["ncols"].copy(), o["status"].copy())
            o5 = ila.qr_banded_solve(2, 2, [[1, 0, 3, 0, 1]] * 40, [[0] * 5] * 40, shift=np.linspace(-0.5, 0.5, 40), b=[3., 4, 5], ncap=9000)
            outs[(mode, "5band")] = (o5["xflat"].copy(), o5["ncols"].copy(), o5["status"].copy())
    finally:
        _lib.check(_lib.lib().ila_qr_set_arith_mode(6))
    for key in (False, True, "5band"):
        a, b = outs[(5, key)], outs[(6, key)]
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert np.array_equal(a[0].view(np.int64), b[0].view(np.int64)), key


def _solve_both_sweeps(ila, *args, **kw):
    """The same solve with the one-warp forward sweep (mode 7) and the scout / verifier / clerk sweep forced (mode 11)."""
    from ila_b200 import _lib
    outs = []
    try:
        for mode in (7, 11):
            _lib.check(_lib.lib().ila_qr_set_arith_mode(mode))
            o = ila.qr_banded_solve(*args, **kw)
            outs.append(o)
    finally:
        _lib.check(_lib.lib().ila_qr_set_arith_mode(8))
    a, b = outs
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["ncols"], b["ncols"]) and np.array_equal(a["nconv"], b["nconv"])
    assert np.array_equal(a["xflat"].view(np.int64), b["xflat"].view(np.int64))
    return b


def test_three_warp_sweep_regimes_bit_exact(ila):
    """The l = 1 forward sweep on three warps (candidates from a shortened chain, proved correctly rounded before use) against
    the one-warp sweep and the oracle, bit for bit, in the regimes that stress the proofs: candidates that are powers of two
    (tau = 1 when the pivot entry vanishes, tau = 2 when the subdiagonal is negligible, normu = |b|), operands outside the
    tests' exponent boxes (every column takes the IEEE path, the verifier goes on alone), subdiagonals outside the sweep's box
    (the group falls back to one warp), chunk boundaries inside every other group, ragged groups, capacity and NaN exits."""
    from oracle import banded as ob
    rng = np.random.default_rng(11)

    def check(l, u, p, q, r=None, **kw):
        g = _solve_both_sweeps(ila, l, u, p, q, r, **kw)
        okw = {k: v for k, v in kw.items() if k != "ncap_per"}
        o = ob.qr_banded_solve(l, u, p, q, r, **okw)
        _compare(g, o, len(np.atleast_2d(p)))
        return g

    # zero main diagonal at the start (pivot entry 0: tau = 1, x1 = +-1), slowly growing diagonal
    batch = 37
    p = np.zeros((batch, 3)); p[:, 0] = 1.0; p[:, 2] = 1.0
    q = np.zeros((batch, 3)); q[:, 1] = np.linspace(0.002, 0.02, batch)
    g = check(1, 1, p, q, b=[1.0, 0.5], ncap=60000)
    assert np.all(g["status"] == 0)
    # subdiagonal negligible against a fast-growing diagonal (normu = |a|, tau = 2 exactly), scaled subdiagonals inside the box
    for sub in (1e-9, 1e-60, 3e70):
        p = np.tile([[0.7 * sub, 2.0, sub]], (5, 1)); p[:, 1] = sub * np.array([1.0, -3.0, 0.25, 1e3, -1e-3])
        q = np.zeros((5, 3)); q[:, 1] = sub * np.array([1e4, 2.0, 1e8, -5.0, 0.5])
        check(1, 1, p, q, b=[sub, -sub, 2 * sub], ncap=20000, tol=1e-300)
    # diagonal so large that s leaves the proofs' box (2^900): IEEE columns throughout; subdiagonal outside the sweep's box
    p = np.array([[1.0, 1e200, 1.0], [1.0, -3e180, 2.0]]); q = np.array([[0.0, 1e198, 0.0], [0.0, 1e179, 0.0]])
    check(1, 1, p, q, b=[1.0], ncap=6000, tol=1e-280)
    p = np.array([[1.0, 2.0, 1e-150], [1.0, 3.0, 1e150]]); q = np.array([[0.0, 0.3, 0.0], [0.0, -0.2, 0.0]])
    check(1, 1, p, q, b=[1.0, 1.0], ncap=8000, tol=1e-250)
    # other upper bandwidths, denominators, shifts, a right-hand side longer than a group, tiny chunks (a chunk boundary in every
    # other group of 8 columns), 33 and 64 instances
    for (u, batch, cgw, nb) in [(0, 33, 16, 1), (2, 64, 24, 11), (3, 5, 1000, 3), (1, 33, 12, 20)]:
        nd = u + 2
        p = rng.normal(size=(batch, nd)); p[:, u] += 5.0 * np.sign(p[:, u])
        p[:, u + 1] = rng.choice([-1.0, 0.5, 2.0, 1.0], size=batch)
        q = np.zeros((batch, nd)); q[:, u] = 0.03 * np.sign(p[:, u]) * rng.uniform(0.5, 2.0, size=batch)
        r = np.ones((batch, nd)); r[:, u] = rng.uniform(0.5, 3.0, size=batch)
        check(1, u, p, q, r, shift=rng.normal(size=batch) * 0.2, b=rng.normal(size=(batch, nb)), ncap=12000, colgrowth=cgw, tol=1e-30)
    # capacity reached (status 3, nothing returned) next to instances that converge; NaN in an operator
    z = np.array([200.0, 5000.0, 300.0])
    p, q, r = ila.bessel_operator(z)
    g = _solve_both_sweeps(ila, 1, 1, p, q, r, b=jv(1, z)[:, None], ncap=3000)
    assert list(g["status"]) == [0, 3, 0] and g["ncols"][1] == 0
    o = ob.qr_banded_solve(1, 1, p[[0, 2]], q[[0, 2]], r[[0, 2]], b=jv(1, z[[0, 2]])[:, None], ncap=3000)
    assert np.array_equal(g["x"][0], o["x"][0, :int(o["ncols"][0])]) and np.array_equal(g["x"][2], o["x"][1, :int(o["ncols"][1])])
    p2 = p.copy(); p2[1, 1] = np.nan
    g = _solve_both_sweeps(ila, 1, 1, p2, q, r, b=jv(1, z)[:, None], ncap=9000)
    assert g["status"][1] != 0 and g["status"][0] == 0 and g["status"][2] == 0


def test_three_warp_sweep_randomized_against_one_warp(ila):
    """Random l = 1 operators (bandwidths, scales over six decades, slopes of either sign, denominators, shifts, right-hand
    sides, chunk lengths, tolerances, batch sizes that leave ragged groups): the three-warp sweep must reproduce the one-warp
    sweep bit for bit — solution, ncols, nconv, status — whatever happens on the way (re-seeds, IEEE columns, fallback groups,
    instances that hit the capacity)."""
    from oracle import banded as ob
    rng = np.random.default_rng(2024)
    statuses = set()
    for trial in range(40):
        u = int(rng.integers(0, 4))
        nd = u + 2
        batch = int(rng.integers(1, 71))
        scale = 10.0 ** rng.uniform(-3, 3)
        p = rng.normal(size=(batch, nd)) * scale
        p[:, u] += 4.0 * scale * np.sign(p[:, u])
        p[:, u + 1] = scale * rng.choice([-2.0, -1.0, 0.25, 1.0, 3.0], size=batch) * 10.0 ** rng.uniform(-2, 2)
        q = np.zeros((batch, nd))
        q[:, u] = scale * rng.choice([-1.0, 1.0], size=batch) * 10.0 ** rng.uniform(-3, 0, size=batch)
        r = np.ones((batch, nd))
        if rng.random() < 0.5:
            r[:, u] = rng.uniform(0.3, 4.0, size=batch)
        shift = rng.normal(size=batch) * scale if rng.random() < 0.7 else None
        nb = int(rng.integers(1, 6))
        b = rng.normal(size=(batch, nb)) * 10.0 ** rng.uniform(-5, 5)
        kw = dict(shift=shift, b=b, ncap=int(rng.choice([700, 4000, 12000])), colgrowth=int(rng.choice([8, 50, 1000])),
                  tol=float(rng.choice([1e-300, 1e-40, 1e-12])))
        g = _solve_both_sweeps(ila, 1, u, p, q, r, **kw)
        statuses.update(int(s) for s in g["status"])
        if trial % 8 == 0:      # and the oracle, on a few of them
            o = ob.qr_banded_solve(1, u, p, q, r, **kw)
            assert np.array_equal(g["status"], o["status"])
            ok = o["status"] == 0
            assert np.array_equal(g["ncols"][ok], o["ncols"][ok]) and np.array_equal(g["nconv"][ok], o["nconv"][ok])
            for i in np.nonzero(ok)[0]:
                assert np.array_equal(g["x"][i], o["x"][i, :int(o["ncols"][i])])
    assert 0 in statuses


def test_three_warp_sweep_recovers_from_misrounded_candidates(ila):
    """Fault injection (mode 1000 + n: the scout corrupts one candidate by an ulp every n columns): the verifier must reject
    it, take the IEEE column for that lane, re-seed the scout — and, when faults come too often, go on alone.  Whatever the
    rate, solution, ncols, nconv and status must equal the one-warp sweep's bit for bit."""
    from ila_b200 import _lib
    lib = _lib.lib()
    z = np.concatenate([np.linspace(40.0, 2500.0, 37), [6000.0]])
    p, q, r = ila.bessel_operator(z)
    kw = dict(b=jv(1, z)[:, None], ncap_per=ila.bessel_ncap(z))
    rng = np.random.default_rng(5)
    p3 = rng.normal(size=(9, 5)); p3[:, 3] += 5.0; p3[:, 4] = 1.5
    q3 = np.zeros((9, 5)); q3[:, 3] = 0.02
    try:
        _lib.check(lib.ila_qr_set_arith_mode(7))
        ref = ila.qr_banded_solve(1, 1, p, q, r, **kw)
        ref3 = ila.qr_banded_solve(1, 3, p3, q3, b=rng.normal(size=(9, 2)), ncap=9000, tol=1e-200)
        rng = np.random.default_rng(5); rng.normal(size=(9, 5))
        _lib.check(lib.ila_qr_set_arith_mode(11))
        for every in (4001, 257, 31, 3, 1):      # rare -> every column (the verifier ends up alone)
            _lib.check(lib.ila_qr_set_arith_mode(1000 + every))
            g = ila.qr_banded_solve(1, 1, p, q, r, **kw)
            assert np.array_equal(g["status"], ref["status"]) and np.array_equal(g["ncols"], ref["ncols"]) and np.array_equal(g["nconv"], ref["nconv"])
            assert np.array_equal(g["xflat"].view(np.int64), ref["xflat"].view(np.int64)), every
            rng2 = np.random.default_rng(5); rng2.normal(size=(9, 5))
            g3 = ila.qr_banded_solve(1, 3, p3, q3, b=rng2.normal(size=(9, 2)), ncap=9000, tol=1e-200)
            assert np.array_equal(g3["ncols"], ref3["ncols"]) and np.array_equal(g3["xflat"].view(np.int64), ref3["xflat"].view(np.int64)), every
    finally:
        _lib.check(lib.ila_qr_set_arith_mode(1000))
        _lib.check(lib.ila_qr_set_arith_mode(8))
