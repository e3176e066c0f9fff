"""The adaptive block-banded QR oracle (oracle/block_qr.py) against the reference's own block-banded tests
(test/test_infqr.jl:270-313): the golden vectors (F.Q'e1)[1:6] and (F.Q e1)[1:6], |factors| against a dense Householder QR
of the finite section, (A u)[1:10] = e1 and the generating-function identities 1/(x-2), 1/(x+y-4), 1/(2x+y-8)."""
import numpy as np

from oracle import block_qr as bq

T = bq.toeplitz_bands(0.0, 0.5, 80)        # Δ = BandedMatrix(1 => Ones/2, -1 => Ones/2)


def _gen2d(u, nb, x, y):
    th, ph = np.arccos(x), np.arccos(y)
    S = 0.0
    for K in range(1, nb + 1):
        blk = u[bq.block_start(K):bq.block_start(K + 1)]
        k = np.arange(1, K + 1)
        S += np.sum(blk * np.sin(k * th) / np.sin(th) * np.sin((K - k + 1) * ph) / np.sin(ph))
    return S


def test_golden_vectors_and_factors():
    o = bq.block_qr_solve(T, T, 1.0, 0.0, 2.0)                       # A = KronTrav(Δ - 2I, Eye)
    assert o["status"] == 0
    gold = [-0.9701425001453321, 0, -0.23386170701251197, 0, 0, -0.06193705069863463]      # test_infqr.jl:276
    assert np.abs(o["qtb"][:6] - gold).max() <= 1e-15
    F = o["F"]
    # (F.Q*e1)[1:6] (test_infqr.jl:277): H_1 e1 = e1 - tau_1 [1; v]
    q1 = np.zeros(6); q1[0] = 1.0
    v = np.concatenate([[1.0], F.M[1:3, 0]])
    q1[:3] -= F.tau[0] * v
    assert np.abs(q1 - [-0.9701425001453321, 0, 0.24253562503633297, 0, 0, 0]).max() <= 1e-15
    # abs.(F.factors[1:15,1:10]) ≈ abs.(qr(A[1:15,1:10]).factors) (:274) — R part against LAPACK's QR of the section
    A = bq.dense_operator(T, T, 1.0, 0.0, 2.0, 6)
    R = np.linalg.qr(A[:15, :10], mode="r")
    assert np.abs(np.abs(np.triu(F.M[:10, :10])) - np.abs(R)).max() <= 1e-14


def test_solutions_of_the_reference_operators():
    x, y = 0.1, 0.2
    th = np.arccos(x)
    s50 = np.sin(np.arange(1, 51) * th) / np.sin(th)
    # A = KronTrav(Δ - 2I, Eye): (A*u)[1:10] ≈ e1, dot(u[Block(K)][K], U_{K-1}(x)) ≈ 1/(x-2)  (:279-285)
    o = bq.block_qr_solve(T, T, 1.0, 0.0, 2.0)
    u = o["x"]
    A = bq.dense_operator(T, T, 1.0, 0.0, 2.0, 12)
    e1 = np.zeros(10); e1[0] = 1
    assert np.abs((A @ u[:A.shape[0]])[:10] - e1).max() <= 1e-14
    assert abs(u[[bq.block_start(K) + K - 1 for K in range(1, 51)]] @ s50 - 1 / (x - 2)) <= 1e-14
    # B = KronTrav(Eye, Δ - 2I): u[Block(K)][1]  (:287-290)
    u = bq.block_qr_solve(T, T, 0.0, 1.0, 2.0)["x"]
    assert abs(u[[bq.block_start(K) for K in range(1, 51)]] @ s50 - 1 / (x - 2)) <= 1e-14
    # L = A + B (:292-304) and 2X + Y - 8II (:306-312)
    o = bq.block_qr_solve(T, T, 1.0, 1.0, 4.0)
    assert abs(_gen2d(o["x"], 50, x, y) - 1 / (x + y - 4)) <= 1e-14
    A = bq.dense_operator(T, T, 1.0, 1.0, 4.0, 12)
    assert np.abs((A @ o["x"][:A.shape[0]])[:10] - e1).max() <= 1e-14
    o = bq.block_qr_solve(T, T, 2.0, 1.0, 8.0)
    assert o["ncols"] >= bq.block_start(41) and abs(_gen2d(o["x"], 40, x, y) - 1 / (2 * x + y - 8)) <= 1e-14


def test_capacity_and_chunk_cadence():
    o = bq.block_qr_solve(T, T, 1.0, 1.0, 4.0)
    assert o["nblocks"] == 56 and o["ncols"] == bq.block_start(52)            # = 1326 (51 blocks); chunks end at blocks 24, 35, 43, 50, 56
    o = bq.block_qr_solve(T, T, 1.0, 1.0, 4.0, maxblocks=40)
    assert o["status"] == 3
    o = bq.block_qr_solve(T, T, 1.0, 1.0, 4.0, tol=1e-8, colgrowth=100)
    assert o["status"] == 0 and o["nblocks"] < 30
