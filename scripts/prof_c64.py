"""Complex QR run (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ila_b200
lam = np.linspace(-2, 2, 512) + 0.2j
p = np.tile(np.array([0.5, 0.0, 0.5]), (512, 1)); q = np.zeros((512, 3))
for _ in range(2):
    g = ila_b200.qr_banded_solve_c64(1, 1, p, q, None, shift=lam, b=[1.0], ncap=20000)
print("ok", int(g["ncols"].sum()), int((g["status"] != 0).sum()))
