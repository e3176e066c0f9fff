"""Tiny runs of every hot kernel for compute-sanitizer (memcheck / racecheck / initcheck).  Not a test: run as
`compute-sanitizer --tool memcheck python scripts/sanitize_all.py` under gpurun; the log goes to profiles/."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.special import jv
import ila_b200
from ila_b200 import _lib
col = np.array([2j, 0, 0, 1, 0.7])
lam = (np.linspace(-10, 7, 64)[None, :] + 1j * np.linspace(-7, 8, 40)[:, None])
L, st = ila_b200.ql_toeplitz_l11(col, lam); print("K3", int((st != 0).sum()))
L, st8 = ila_b200.ql_toeplitz_l11_real(col, lam); print("K3 compact", int((st8 != 0).sum()))
L, st, tail = ila_b200.ql_toeplitz_l11(np.array([0.5, 5.1, 2.0]), np.linspace(-1, 1, 50), want_tail=True); print("K3 real tail", st.sum())
x, st = ila_b200.ql_toeplitz_solve(col, lam.reshape(-1)[:300], [1.0, 0.5j], 9); print("K7", int((st != 0).sum()))
y, ns, st = ila_b200.ql_toeplitz_applyq(col, lam.reshape(-1)[:300], [1.0], 50); print("Qb", int((st != 0).sum()))
z = np.array([30.0, 100.0, 400.0]); p, q, r = ila_b200.bessel_operator(z)
o = ila_b200.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap_per=ila_b200.bessel_ncap(z), want_factors=True); print("K1+K2", o["status"], o["ncols"])
for mode in (5, 6):
    _lib.check(_lib.lib().ila_qr_set_arith_mode(mode))
    o = ila_b200.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap_per=ila_b200.bessel_ncap(z)); print("K2 loader", mode, o["status"])
_lib.check(_lib.lib().ila_qr_set_arith_mode(6))
plan = ila_b200.QrBandedPlan(1, 1, p, q, r, b=jv(1, z)[:, None], ncap=4000)
plan.run_factor_resumable(1500, False); plan.run_factor_resumable(4000, True); plan.run_backsolve(); print("resume", plan.d_status.cpu().numpy())
p4, q4 = ila_b200.pentadiagonal_operator(np.array([0.1, 3.0]), eps=1e-2)
o = ila_b200.chol_banded_solve(2, p4, q4, b=[1.0], ncap=6000); print("K5", o["status"], o["ncols"])
o = ila_b200.ql_blocktri_factors(*ila_b200.periodic_schrodinger_blocks(), np.array([-0.5, 3.0]), rhs=[1.0, 2.0], nout=7); print("K6", o["status"])
T = np.stack([np.full(80, 0.5), np.zeros(80), np.full(80, 0.5)])
for mode in (0, 1):
    ila_b200.lib().ila_qr_blockbanded_set_mode(mode)
    o = ila_b200.qr_blockbanded_solve(T, T, [1.0, 2.0], [1.0, 1.0], [4.0, 8.0], tol=1e-10, colgrowth=60, maxblocks=30)
    print("K11 mode", mode, o["status"], o["nblocks"], o["ncols"])
ila_b200.lib().ila_qr_blockbanded_set_mode(0)
rng = np.random.default_rng(0)
U = rng.normal(size=(3, 3, 301)) * 0.3; U[:, 0] = 1 + rng.random((3, 301)); X = rng.normal(size=(3, 3, 301))
print("K10", float(np.abs(ila_b200.tridiagonal_conjugation(U, X, 300)).max()))
o = ila_b200.qr_almostbanded_solve(0, 1, [[1, 1, 0], [-1, 0, 0]], [[1, 1, 0, 0]], [np.e], tol=1e-12, colgrowth=50, ncap=400); print("K12", o["status"], o["ncols"])
o = ila_b200.revchol_tridiag([3, 0, 0, 1, 0], [1, 0, 0, 1, 0], 32, shift=np.linspace(0, 0.9, 5)); print("K8", o["status"])
o = ila_b200.ql_finite_section(1, 1, [[1.0, 0, 0, 1, 0], [0.0, 1, 0, 1, 1], [1.0, 0, 0, 1, 0]], shift=np.linspace(-6, -3, 5)); print("K9", o["status"])
print("done")
