import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ila_b200
from oracle import block_qr as bq
T = bq.toeplitz_bands(0.0, 0.5, 80)
al = [1.0, 0.0, 1.0, 2.0]; be = [0.0, 1.0, 1.0, 1.0]; lam = [2.0, 2.0, 4.0, 8.0]
for mode in (0, 1):
    ila_b200.lib().ila_qr_blockbanded_set_mode(mode)
    g = ila_b200.qr_blockbanded_solve(T, T, al, be, lam, want_qtb=True)
    for i in range(4):
        o = bq.block_qr_solve(T, T, al[i], be[i], lam[i])
        n = o["ncols"]; m = len(o["qtb"])
        ex = np.abs(g["x"][i, :n] - o["x"]).max() if g["ncols"][i] == n else -1
        print(mode, i, g["status"][i], g["nblocks"][i], g["ncols"][i], o["nblocks"], n, "x err", ex, "qtb err", np.abs(g["qtb"][i, :m] - o["qtb"]).max(),
              "first bad qtb idx", np.argmax(np.abs(g["qtb"][i, :m] - o["qtb"]) > 1e-12))
