import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ila_b200
col = np.array([2j, 0, 0, 1, 0.7])
x = np.linspace(-10, 7, 1024); y = -0.92822265625
for s in (984, 992):
    L, st = ila_b200.ql_toeplitz_l11(col, x[s:s + 8] + 1j * y)
    print(s, "codes (f32 sweeps + 100*f64 sweeps + 10000*flags[1 warm,2 full,4 lost]):", np.round(L.imag, 4))
# whole grid statistics with the same offsets as the slow slab
yy = -7 + (np.arange(1024) + 0.5) * (15.0 / 1024)
lam = (x[None, :] + 1j * yy[:, None]).reshape(-1)
L, st = ila_b200.ql_toeplitz_l11(col, lam)
c = L.imag[st == 0]
f32 = np.floor(c % 100); f64 = np.floor(c / 100) % 100; flags = np.floor(c / 10000)
print("points with f64 fallback:", int((f64 > 0).sum()), "max f32 sweeps", f32.max(), "max f64 sweeps", f64.max(), "lost:", int((flags >= 4).sum()))
idx = np.nonzero(st == 0)[0][np.argsort(-(f32 + f64))[:10]]
for i in idx: print(i // 1024, i % 1024, lam[i], L.imag[i])
