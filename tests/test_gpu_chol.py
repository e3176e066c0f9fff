"""Parity of the CUDA adaptive banded Cholesky solve (K5 + K2) against the C oracle, through the C ABI.  Bar: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(g, o, batch, factors=False):
    assert np.array_equal(g["status"], o["status"])
    assert np.array_equal(g["ncols"], o["ncols"])
    assert np.array_equal(g["nconv"], o["nconv"])
    for i in range(batch):
        n = int(o["ncols"][i])
        assert np.array_equal(g["x"][i], o["x"][i, :n]), f"instance {i}: solution not bit-exact"
        if factors:
            assert np.array_equal(g["factors"][i], o["factors"][i, :n]), f"instance {i}: U not bit-exact"
            assert np.array_equal(g["y"][i], o["y"][i, :n]), f"instance {i}: L\\b not bit-exact"


def test_tridiagonal_bit_exact(ila):
    # test/test_infcholesky.jl:5-11
    from oracle import banded as ob
    kw = dict(b=[1.0, 2.0, 3.0], ncap=6000, want_factors=True)
    g = ila.chol_banded_solve(1, [[1.0, 1.0]], [[0.0, 1.0]], **kw)
    o = ob.chol_banded_solve(1, [[1.0, 1.0]], [[0.0, 1.0]], **kw)
    _compare(g, o, 1, factors=True)


def test_pentadiagonal_shift_family_bit_exact(ila):
    # BASELINE config C4 (sub-sampled): 64 shifts, several chunk boundaries each
    from oracle import banded as ob
    sigma = 0.01 + (10 - 0.01) * np.arange(0, 2048, 32) / 2047
    p, q = ila.pentadiagonal_operator(sigma)
    g = ila.chol_banded_solve(2, p, q, b=[1.0], ncap=12000)
    o = ob.chol_banded_solve(2, p, q, b=[1.0], ncap=12000)
    assert np.all(o["status"] == 0)
    _compare(g, o, len(sigma))
    g = ila.chol_banded_solve(2, p[:3], q[:3], b=[1.0, -0.5], ncap=12000, want_factors=True)
    o = ob.chol_banded_solve(2, p[:3], q[:3], b=[1.0, -0.5], ncap=12000, want_factors=True)
    _compare(g, o, 3, factors=True)


def test_other_bandwidths_shift_head_colgrowth(ila):
    from oracle import banded as ob
    rng = np.random.default_rng(3)
    for u in (1, 2, 3, 4):
        nd = u + 1
        batch = 6
        p = rng.normal(size=(batch, nd)) * 0.3
        p[:, u] = 4.0 + rng.random(batch) * 2          # diagonally dominant => SPD
        q = np.zeros((batch, nd))
        q[:, u] = 0.02
        shift = rng.random(batch) * 0.5
        head = np.abs(rng.normal(size=(nd, 3))) * 0.1
        head[u, :] += 5.0
        b = rng.normal(size=(batch, 3))
        kw = dict(shift=shift, head=head, b=b, ncap=9000, colgrowth=300)
        g = ila.chol_banded_solve(u, p, q, **kw)
        o = ob.chol_banded_solve(u, p, q, **kw)
        assert np.all(o["status"] == 0)
        _compare(g, o, batch)


def test_status_codes(ila):
    g = ila.chol_banded_solve(1, [[1.0, -1.0]], [[0.0, 0.0]], b=[1.0], ncap=5000)
    assert g["status"][0] == 4 and g["ncols"][0] == 0
    g = ila.chol_banded_solve(2, [[1.0, -4.0, 6.01]], [[0.0, 0.0, 1e-4]], b=[1.0], ncap=900)
    assert g["status"][0] == 3


def test_c4_full_size_properties(ila):
    """BASELINE config C4 at full size (2048 shifts): residual and cross-method agreement cholesky(S)\\b ≈ qr(S)\\b
    (test/test_infcholesky.jl:7,11,51,57)."""
    sigma = 0.01 + (10 - 0.01) * np.arange(2048) / 2047
    p, q = ila.pentadiagonal_operator(sigma)
    g = ila.chol_banded_solve(2, p, q, b=[1.0], ncap=12000)
    assert np.all(g["status"] == 0)
    assert np.all(g["ncols"] == g["nconv"] + 1001)
    for i in (0, 700, 2047):
        x = g["x"][i]
        k = np.arange(2, 300)
        d = p[i, 2] + q[i, 2] * k
        res = x[k - 2] - 4 * x[k - 1] + d * x[k] - 4 * x[k + 1] + x[k + 2]
        assert np.max(np.abs(res)) < 1e-13
    # the same operators through the QR path (full 5-band description)
    pq = np.stack([p[:, 0], p[:, 1], p[:, 2], p[:, 1], p[:, 0]], axis=1)[::256]
    qq = np.zeros_like(pq)
    qq[:, 2] = q[::256, 2]
    gq = ila.qr_banded_solve(2, 2, pq, qq, b=[1.0], ncap=12000)
    for j, i in enumerate(range(0, 2048, 256)):
        m = min(len(g["x"][i]), len(gq["x"][j]))
        assert np.allclose(g["x"][i][:m], gq["x"][j][:m], rtol=1e-9, atol=1e-14)


def test_c4_full_batch_bit_exact_vs_oracle(ila):
    """BASELINE config C4 at full size: all 2048 pentadiagonal shifts against the C oracle bit for bit (status, ncols, nconv, x),
    and cholesky(S) \\ b == qr(S) \\ b on every 64th shift (test/test_infcholesky.jl:5-11)."""
    import json, os
    from oracle import banded as ob
    sigma = 0.01 + (10 - 0.01) * np.arange(2048) / 2047
    p, q = ila.pentadiagonal_operator(sigma)
    g = ila.chol_banded_solve(2, p, q, b=[1.0], ncap=12000)
    mism = dict(status=0, ncols=0, nconv=0, x_entries=0)
    cols = 0
    for lo in range(0, 2048, 256):
        hi = lo + 256
        o = ob.chol_banded_solve(2, p[lo:hi], q[lo:hi], b=[1.0], ncap=12000, nthreads=os.cpu_count() or 1)
        mism["status"] += int((g["status"][lo:hi] != o["status"]).sum())
        mism["ncols"] += int((g["ncols"][lo:hi] != o["ncols"]).sum())
        mism["nconv"] += int((g["nconv"][lo:hi] != o["nconv"]).sum())
        for i in range(lo, hi):
            n = int(o["ncols"][i - lo])
            cols += n
            mism["x_entries"] += int((g["x"][i].view(np.int64) != o["x"][i - lo, :n].view(np.int64)).sum()) if len(g["x"][i]) == n else n
    # full symmetric band description for the QR entry (data rows +2, +1, 0, -1, -2)
    p5 = np.stack([p[::64, 0], p[::64, 1], p[::64, 2], p[::64, 1], p[::64, 0]], axis=1)
    q5 = np.zeros_like(p5)
    q5[:, 2] = q[::64, 2]
    r = ila.qr_banded_solve(2, 2, p5, q5, b=[1.0], ncap=12000)
    worst = 0.0
    for k, i in enumerate(range(0, 2048, 64)):
        m = min(len(g["x"][i]), len(r["x"][k]))
        worst = max(worst, float(np.abs(g["x"][i][:m] - r["x"][k][:m]).max() / np.abs(r["x"][k][:m]).max()))
    rep = dict(config="C4 2048 pentadiagonal Cholesky shifts", columns_compared=cols, mismatches_vs_c_oracle=mism,
               all_ok=bool(np.all(g["status"] == 0)), chol_vs_qr_max_rel=worst)
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        json.dump(rep, open(os.path.join(root, "gpurun_out", "r02_parity_c4_full.json"), "w"), indent=1)
    except OSError:
        pass
    assert rep["all_ok"] and all(v == 0 for v in mism.values()), rep
    assert worst <= 1e-12, rep
