"""Diagnostic (not a test): GPU vs oracle error distributions and quick timings. Run under gpurun."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ila_b200
from oracle import toeplitz_ql as tq
from oracle import banded as ob
from scipy.special import jv

def bullhead(nx, ny):
    x = np.linspace(-10, 7, nx); y = np.linspace(-7, 8, ny)
    return (x[None, :] + 1j * y[:, None])

col = np.array([2j, 0, 0, 1, 0.7])
print("devices", ila_b200.lib().ila_device_count(), "fp64 peak TF", ila_b200.fp64_peak_tflops())
for nx in (256, 1024):
    lam = bullhead(nx, nx)
    t = time.time(); L, st = ila_b200.ql_toeplitz_l11(col, lam); t1 = time.time() - t
    t = time.time(); L, st = ila_b200.ql_toeplitz_l11(col, lam); t2 = time.time() - t
    t = time.time(); Lo, sto = tq.ql_toeplitz_l11(col, lam); t3 = time.time() - t
    print(f"grid {nx}: gpu first {t1*1e3:.1f} ms, second {t2*1e3:.1f} ms, oracle {t3:.2f} s")
    print(" status mismatches", int((st != sto).sum()), "of", st.size, " noql frac", float((sto == 1).mean()))
    ok = (st == 0) & (sto == 0)
    err = np.abs(L[ok] - Lo[ok]); rel = err / np.abs(Lo[ok])
    print(" max imag gpu", np.abs(L[ok].imag).max(), "oracle", np.abs(Lo[ok].imag).max())
    for qq in (0.5, 0.9, 0.99, 0.999, 0.9999, 1.0):
        print(f"  q{qq}: abs {np.quantile(err, qq):.3e} rel {np.quantile(rel, qq):.3e}")
    bad = np.argsort(rel)[-5:]
    lam_ok = lam[ok]
    for i in bad:
        print("   worst", lam_ok[i], L[ok][i], Lo[ok][i], rel[i])
    mm = np.argwhere(st != sto)[:5]
    for ij in mm:
        print("   mismatch at", lam[tuple(ij)], st[tuple(ij)], sto[tuple(ij)], L[tuple(ij)], Lo[tuple(ij)])

# real path: the six reference cases
for (Z, A, B) in ((2.0, 5.1, 0.5), (2.0, 2.2, 0.5), (2.0, -2.2, 0.5), (2.0, -5.1, 0.5), (0.5, 5.1, 2.0), (0.5, -5.1, 2.0), (0.5, 2.1, 2.0)):
    L, st, tail = ila_b200.ql_toeplitz_l11(np.array([B, A, Z]), np.array([0.0]), want_tail=True)
    o = tq.ql_toeplitz_tail(np.array([Z, A, B]))
    print("real", (Z, A, B), L, st, o["L11"], o["status"], np.abs(tail[0][:3] - o["X"][0]).max(), np.abs(tail[0][3:6] - o["X"][1]).max(), tail[0][6:], o["tau"])

# QR: bit-exactness vs oracle
for zs in ([10.0, 100.0, 1000.0, 2500.0], np.linspace(1e3, 1e4, 64)):
    zs = np.asarray(zs, dtype=np.float64)
    p, q, r = ila_b200.bessel_operator(zs)
    b = jv(1, zs)[:, None]
    ncap = int(1.06 * zs.max() + 3000)
    t = time.time(); g = ila_b200.qr_banded_solve(1, 1, p, q, r, b=b, ncap=ncap); tg = time.time() - t
    t = time.time(); g = ila_b200.qr_banded_solve(1, 1, p, q, r, b=b, ncap=ncap); tg2 = time.time() - t
    t = time.time(); o = ob.qr_banded_solve(1, 1, p, q, r, b=b, ncap=ncap); to = time.time() - t
    print(f"QR batch {len(zs)}: gpu {tg*1e3:.1f}/{tg2*1e3:.1f} ms oracle {to*1e3:.1f} ms; ncols eq {np.array_equal(g['ncols'], o['ncols'])} nconv eq {np.array_equal(g['nconv'], o['nconv'])} status {g['status'][:4]} {o['status'][:4]}")
    nbit = 0; mx = 0.0
    for i in range(len(zs)):
        n = int(o["ncols"][i])
        xe = o["x"][i, :n]; xg = g["x"][i]
        if len(xg) == n:
            nbit += int(np.array_equal(xe, xg))
            mx = max(mx, float(np.abs(xe - xg).max() / np.abs(xe).max()))
    print("  bit-exact instances", nbit, "of", len(zs), "max rel diff", mx)
    k = np.arange(int(zs[-1]) + 500)
    ref = jv(k, zs[-1]); xg = g["x"][-1]
    m = min(len(ref), len(xg))
    print("  vs besselj (last z)", np.abs(xg[:m] - ref[:m]).max() / np.abs(ref).max())
# factors export
g = ila_b200.qr_banded_solve(1, 1, [[1, 1, 1]], [[0, 1, 0]], b=[1., 2, 3], ncap=5000, want_factors=True)
o = ob.qr_banded_solve(1, 1, [[1, 1, 1]], [[0, 1, 0]], b=[1., 2, 3], ncap=5000, want_factors=True)
n = int(o["ncols"][0])
print("factors eq", np.array_equal(g["factors"][0], o["factors"][0, :n]), "tau eq", np.array_equal(g["tau"][0], o["tau"][0, :n]), "qtb eq", np.array_equal(g["qtb"][0], o["qtb"][0, :n]), g["factors"][0][0], g["tau"][0][0])
# 5-band
head = np.zeros((5, 3)); head[0, :] = 1; head[1, :] = [1, 0, 0]; head[2, :] = [4, 5, 6]; head[3, :] = [1, 0, 0]; head[4, :] = 1
pp = [[1, 0, 3, 0, 1]]; qq = [[0, 0, 0, 0, 0]]
g = ila_b200.qr_banded_solve(2, 2, pp, qq, head=head, b=[3., 4, 5], ncap=8000)
o = ob.qr_banded_solve(2, 2, pp, qq, head=head, b=[3., 4, 5], ncap=8000)
n = int(o["ncols"][0])
print("5-band: status", g["status"], o["status"], "ncols", g["ncols"], o["ncols"], "bit-exact", np.array_equal(g["x"][0], o["x"][0, :n]), g["x"][0][:4])
