import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ila_b200
col = np.array([2j, 0, 0, 1, 0.7])
lam = ila_b200.sharding.grid_rows(1024, 4096, 1, 4)
L, st = ila_b200.ql_toeplitz_l11(col, lam)
c = L.imag
ok = st == 0
f32 = np.floor(c % 100); f64 = np.floor(c / 100) % 100; flags = np.floor(c / 10000)
cost = np.where(ok, f32 + 3 * f64, 0)
print("f64 fallbacks:", int((ok & (f64 > 0)).sum()), "lost:", int((ok & (flags >= 4)).sum()), "max f32", f32[ok].max(), "max f64", f64[ok].max())
runs = cost.reshape(-1, 8).sum(1)
top = np.argsort(-runs)[:6]
for r in top:
    i = r * 8
    print("run", i // 1024, i % 1024, lam[i], "cost", runs[r], "codes", np.round(c[i:i + 8], 3))
print("typical run cost", np.median(runs))
