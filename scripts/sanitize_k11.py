"""Tiny K10 / K11 / K12 runs for compute-sanitizer (memcheck / racecheck).  Not a test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ila_b200
T = np.stack([np.full(80, 0.5), np.zeros(80), np.full(80, 0.5)])
for mode in (0, 1):
    ila_b200.lib().ila_qr_blockbanded_set_mode(mode)
    o = ila_b200.qr_blockbanded_solve(T, T, [1.0, 2.0], [1.0, 1.0], [4.0, 8.0], tol=1e-10, colgrowth=60, maxblocks=30)
    print("K11 mode", mode, o["status"], o["nblocks"], o["ncols"])
rng = np.random.default_rng(0)
U = rng.normal(size=(3, 3, 301)) * 0.3; U[:, 0] = 1 + rng.random((3, 301)); X = rng.normal(size=(3, 3, 301))
print("K10", float(np.abs(ila_b200.tridiagonal_conjugation(U, X, 300)).max()))
k = np.arange(1, 121, dtype=np.float64); e = k / np.sqrt(4 * k * k - 1)
o = ila_b200.chol_tridiagonal_conjugation(1, np.zeros((1, 2)), np.zeros((1, 2)), np.stack([e, np.zeros(120), e]), 100, shift=-np.array([1.5, 2.0]), head=np.stack([-e[:110], np.zeros(110)]))
print("K5->K10", o["status"])
o = ila_b200.qr_almostbanded_solve(0, 1, [[1, 1, 0], [-1, 0, 0]], [[1, 1, 0, 0]], [np.e], tol=1e-12, colgrowth=50, ncap=400)
print("K12", o["status"], o["ncols"])
o = ila_b200.qr_almostbanded_solve(1, 2, [[0.3, 0, 0], [1, 1, 0], [-2, 0, 0], [0.4, 0, 0]], [[1, 1, 0, 0]], [1.0], tol=1e-12, colgrowth=50, ncap=400)
print("K12 generic", o["status"], o["ncols"])
