"""Parity of the CUDA adaptive block-banded QR solve (K11) against the numpy oracle, through the C ABI.
Bar: 1e-12 relative to max|x| on the solution and on Q'b (Householder QR with a different summation order: the oracle
uses BLAS dot products on a dense global matrix, the kernel sequential shared-memory sweeps); ncols / nblocks / status equal."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(params=[0, 1], ids=["blocked", "one-reflector"], autouse=True)
def _mode(request, ila):
    """Every test runs with four reflectors per pass (default) and with one reflector per pass (validation mode)."""
    assert ila.lib().ila_qr_blockbanded_set_mode(request.param) == 0
    yield
    ila.lib().ila_qr_blockbanded_set_mode(0)


def _T(m=80):
    from oracle import block_qr as bq
    return bq.toeplitz_bands(0.0, 0.5, m)


def _check(g, i, o):
    assert g["status"][i] == o["status"] and g["nblocks"][i] == o["nblocks"] and g["ncols"][i] == o["ncols"], (i, g["status"][i], g["nblocks"][i], g["ncols"][i], o["status"], o["nblocks"], o["ncols"])
    n = o["ncols"]
    if o["status"] == 0:
        assert np.abs(g["x"][i, :n] - o["x"]).max() <= TOL * np.abs(o["x"]).max()
        assert np.all(g["x"][i, n:] == 0)


def test_reference_operators_and_golden_vector(ila):
    from oracle import block_qr as bq
    T = _T()
    al = [1.0, 0.0, 1.0, 2.0]; be = [0.0, 1.0, 1.0, 1.0]; lam = [2.0, 2.0, 4.0, 8.0]
    g = ila.qr_blockbanded_solve(T, T, al, be, lam, want_qtb=True)
    for i in range(4):
        o = bq.block_qr_solve(T, T, al[i], be[i], lam[i])
        _check(g, i, o)
        m = len(o["qtb"])
        assert np.abs(g["qtb"][i, :m] - o["qtb"]).max() <= TOL
    gold = [-0.9701425001453321, 0, -0.23386170701251197, 0, 0, -0.06193705069863463]      # test_infqr.jl:276
    assert np.abs(g["qtb"][0, :6] - gold).max() <= 1e-15
    # (A*u)[1:10] ≈ e1 and the generating function 1/(x-2) on the GPU result itself (test_infqr.jl:279-285)
    u = g["x"][0]
    A = bq.dense_operator(T, T, 1.0, 0.0, 2.0, 12)
    e1 = np.zeros(10); e1[0] = 1
    assert np.abs((A @ u[:A.shape[0]])[:10] - e1).max() <= 1e-14
    x = 0.1; th = np.arccos(x)
    assert abs(u[[bq.block_start(K) + K - 1 for K in range(1, 51)]] @ (np.sin(np.arange(1, 51) * th) / np.sin(th)) - 1 / (x - 2)) <= 1e-14


def test_shift_family_and_rhs_prefix(ila):
    # 2-D resolvent family (X + Y - lam I) \ b with a three-entry right-hand side, non-Toeplitz 1-D operators
    from oracle import block_qr as bq
    m = 80
    k = np.arange(1, m + 1, dtype=np.float64)
    e = k / np.sqrt(4 * k * k - 1)                      # Legendre Jacobi matrix
    T1 = np.stack([e, np.zeros(m), e])
    T2 = np.stack([np.full(m, 0.5), 0.1 / k, np.full(m, 0.5)])
    lam = np.array([2.6, 3.0, 4.0, 6.0, -3.5, 12.0])
    b = np.array([1.0, -2.0, 0.5])
    g = ila.qr_blockbanded_solve(T1, T2, 1.0, 1.0, lam, b=b)
    for i in range(len(lam)):
        _check(g, i, bq.block_qr_solve(T1, T2, 1.0, 1.0, lam[i], b=b))
    assert len(set(g["nblocks"])) > 1


def test_tolerance_colgrowth_capacity(ila):
    from oracle import block_qr as bq
    T = _T()
    for kw in (dict(tol=1e-8, colgrowth=100), dict(maxblocks=40), dict(maxblocks=30, tol=1e-10, colgrowth=50), dict(maxblocks=57)):
        g = ila.qr_blockbanded_solve(T, T, 1.0, 1.0, [4.0, 5.0], **kw)
        for i, lam in enumerate((4.0, 5.0)):
            o = bq.block_qr_solve(T, T, 1.0, 1.0, lam, **kw)
            _check(g, i, o)
    g = ila.qr_blockbanded_solve(T, T, 1.0, 1.0, [4.0], maxblocks=40)
    assert g["status"][0] == 3 and g["ncols"][0] == 0


def test_batch_larger_than_the_sm_count(ila):
    # 300 shifts (more CTAs than SMs); spot-check against the oracle and a residual property on every instance
    from oracle import block_qr as bq
    T = _T()
    lam = np.linspace(4.0, 9.0, 300)
    g = ila.qr_blockbanded_solve(T, T, 1.0, 1.0, lam, tol=1e-20)
    assert np.all(g["status"] == 0)
    for i in (0, 150, 299):
        _check(g, i, bq.block_qr_solve(T, T, 1.0, 1.0, lam[i], tol=1e-20))
    A0 = bq.dense_operator(T, T, 1.0, 1.0, 0.0, 8)
    n = A0.shape[0]
    e1 = np.zeros(10); e1[0] = 1
    for i in range(300):
        r = ((A0 - lam[i] * np.eye(n)) @ g["x"][i, :n])[:10] - e1
        assert np.abs(r).max() <= 1e-13, i


def test_long_rhs_prefix_and_asymmetric_bands(ila):
    # b spans four blocks (SB = 4: the adaptive test may only fire past it), T1 / T2 not symmetric, negative weights
    from oracle import block_qr as bq
    m = 80
    k = np.arange(1, m + 1, dtype=np.float64)
    T1 = np.stack([0.3 + 0.1 / k, 0.05 * np.cos(k), 0.6 - 0.1 / k])
    T2 = np.stack([np.full(m, 0.45), -0.2 / k, np.full(m, 0.55)])
    b = np.array([1.0, 0.0, -0.5, 0.25, 0.0, 0.0, 2.0])
    al = np.array([1.0, -1.0, 0.7]); be = np.array([1.0, 0.5, -1.3]); lam = np.array([3.5, -3.0, 4.2])
    g = ila.qr_blockbanded_solve(T1, T2, al, be, lam, b=b, tol=1e-25)
    for i in range(3):
        _check(g, i, bq.block_qr_solve(T1, T2, al[i], be[i], lam[i], b=b, tol=1e-25))
