"""Complex-eltype adaptive QR: wall time of the host C-ABI call (not a test). Run under gpurun."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ila_b200
for batch, im in ((512, 0.2), (4096, 0.2), (4096, 0.05)):
    lam = np.linspace(-2, 2, batch) + 1j * im
    p = np.tile(np.array([0.5, 0.0, 0.5]), (batch, 1)); q = np.zeros((batch, 3))
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        g = ila_b200.qr_banded_solve_c64(1, 1, p, q, None, shift=lam, b=[1.0], ncap=60000 if im < 0.1 else 20000)
        ts.append(time.perf_counter() - t0)
    cols = int(g["ncols"].sum())
    print(f"batch {batch} Im={im}: {min(ts)*1e3:.2f} ms, columns {cols} ({cols/batch:.0f}/instance), {cols/min(ts)/1e9:.3f} Gcol/s, bad {int((g['status']!=0).sum())}")
