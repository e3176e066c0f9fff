"""In-process multi-GPU sharding of every host entry point: ngpus = G must reproduce ngpus = 1 bit for bit
(not a pytest: needs >= 2 GPUs; run as `gpurun --gpus 2 -- python scripts/multi_gpu_check.py`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.special import jv
import ila_b200
from ila_b200 import _lib

G = _lib.lib().ila_device_count()
print("devices", G)
assert G >= 2
eq = lambda a, b: np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)
col = np.array([2j, 0, 0, 1, 0.7])
lam = (np.linspace(-10, 7, 301)[None, :] + 1j * np.linspace(-7, 8, 77)[:, None]).reshape(-1)
ok = True

def check(name, f):
    global ok
    a, b = f(1), f(G)
    good = all(eq(x, y) for x, y in zip(a, b))
    ok &= good
    print(f"{name:34s} {'OK' if good else 'MISMATCH'}")

check("ql_toeplitz_l11", lambda g: ila_b200.ql_toeplitz_l11(col, lam, ngpus=g))
check("ql_toeplitz_l11_real (compact)", lambda g: ila_b200.ql_toeplitz_l11_real(col, lam, ngpus=g))
big = (np.linspace(-10, 7, 1024)[None, :] + 1j * np.linspace(-7, 8, 1024)[:, None]).reshape(-1)
def _full(g):
    L, st = ila_b200.ql_toeplitz_l11_real(col, big, ngpus=g)
    return (np.round(np.nan_to_num(L), 9), st)          # slab boundaries move the warm-start runs: last bits may differ
check("ql_toeplitz_l11_real 1024x1024 (1e-9)", _full)
check("ql_toeplitz_absl11_grid", lambda g: (ila_b200.ql_toeplitz_absl11_grid(col, np.linspace(-10, 7, 301), np.linspace(-7, 8, 77), ngpus=g),))
check("ql_toeplitz_solve", lambda g: ila_b200.ql_toeplitz_solve(col, lam[:5000], [1.0, 0.5j], 7, ngpus=g))
head = np.diag(np.full(6, 0.3 + 0j)) + 0.0
A0 = np.zeros((6, 6), complex)
for i in range(6):
    for j in range(max(0, i - 3), min(6, i + 2)):
        k = (j - i) + 3
        A0[i, j] = col[::-1][k]
A0[:4, :4] += 0.2
check("ql_perttoeplitz_l11", lambda g: ila_b200.ql_perttoeplitz_l11(A0, col, lam[:3000], ngpus=g))
check("ql_perttoeplitz_solve", lambda g: ila_b200.ql_perttoeplitz_solve(A0, col, lam[:3000], [1.0, 2.0], 9, ngpus=g))
z = np.linspace(50.0, 3000.0, 101)
p, q, r = ila_b200.bessel_operator(z)
f = lambda g: (lambda o: (o["xflat"], o["ncols"], o["nconv"], o["status"]))(ila_b200.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap_per=ila_b200.bessel_ncap(z), ngpus=g))
check("qr_banded_solve", f)
lc = np.linspace(-2, 2, 67) + 0.2j
pc = np.tile(np.array([0.5, 0.0, 0.5]), (67, 1)); qc = np.zeros((67, 3))
f = lambda g: (lambda o: (o["xflat"], o["ncols"], o["status"]))(ila_b200.qr_banded_solve_c64(1, 1, pc, qc, None, shift=lc, b=[1.0], ncap=20000, ngpus=g))
check("qr_banded_solve_c64", f)
p4, q4 = ila_b200.pentadiagonal_operator(np.linspace(0.01, 10, 70))
f = lambda g: (lambda o: (o["xflat"], o["ncols"], o["status"]))(ila_b200.chol_banded_solve(2, p4, q4, b=[1.0], ncap=8000, ngpus=g))
check("chol_banded_solve", f)
xs = np.concatenate([np.linspace(-0.95, 0.95, 33), np.linspace(2.25, 4.0, 32)])
check("ql_blocktri_l11", lambda g: ila_b200.ql_blocktri_l11(*ila_b200.periodic_schrodinger_blocks(), xs, ngpus=g))
c2, a2, b2 = np.array([[0, 0.5], [0, 0]]), np.array([[0, 2.0], [0.5, 0]]), np.array([[0, 0.0], [2.0, 0]])
check("ql_blocktri_l11 (complex)", lambda g: ila_b200.ql_blocktri_l11([c2], [a2], [b2], c2, a2, b2, np.array([5j, 2 + 4j, -0.5 - 3j]), maxit=5000, ngpus=g))
f = lambda g: (lambda o: (o["la"], o["lb"], o["N"], o["status"]))(ila_b200.revchol_tridiag([3, 0, 0, 1, 0], [1, 0, 0, 1, 0], 64, shift=np.linspace(0, 0.9, 75), ngpus=g))
check("revchol_tridiag", f)
g10 = [[1.0, 0, 0, 1, 0], [0.0, 1, 0, 1, 1], [1.0, 0, 0, 1, 0]]
f = lambda g: (lambda o: (o["Lband"], o["V"], o["tau"], o["N"], o["status"]))(ila_b200.ql_finite_section(1, 1, g10, shift=np.linspace(-6, -3, 77), ngpus=g))
check("ql_finite_section", f)
rng = np.random.default_rng(2)
Ut = rng.normal(size=(37, 3, 801)) * 0.3; Ut[:, 0] = 1 + rng.random((37, 801)); Xt = rng.normal(size=(37, 3, 801))
check("tridiagonal_conjugation", lambda g: (ila_b200.tridiagonal_conjugation(Ut, Xt, 800, ngpus=g),))
check("bidiagonal_conjugation", lambda g: ila_b200.bidiagonal_conjugation(Ut, Xt, 800, "U", ngpus=g))
kk = np.arange(1, 421, dtype=np.float64); ee = kk / np.sqrt(4 * kk * kk - 1)
Xl = np.stack([ee, np.zeros(420), ee]); hd = np.stack([-ee[:410], np.zeros(410)])
f = lambda g: (lambda o: (o["Y"], o["status"]))(ila_b200.chol_tridiagonal_conjugation(1, np.zeros((1, 2)), np.zeros((1, 2)), Xl, 400, shift=-np.linspace(1.05, 3, 45), head=hd, ngpus=g))
check("chol_tridiagonal_conjugation", f)
Tb = np.stack([np.full(50, 0.5), np.zeros(50), np.full(50, 0.5)])
f = lambda g: (lambda o: (o["x"], o["ncols"], o["nblocks"], o["status"]))(ila_b200.qr_blockbanded_solve(Tb, Tb, 1.0, 1.0, np.linspace(4, 9, 21), tol=1e-12, maxblocks=40, ngpus=g))
check("qr_blockbanded_solve", f)
f = lambda g: (lambda o: (o["x"], o["ncols"], o["nconv"], o["status"]))(ila_b200.qr_almostbanded_solve(0, 1, [[1, 1, 0], [0, 0, 0]], [[1, 1, 0, 0]], np.exp(np.linspace(-2, 3, 33))[:, None], shift=np.linspace(-2, 3, 33), ncap=3200, ngpus=g))
check("qr_almostbanded_solve", f)
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
