"""Generates the committed golden fixtures from the oracle (the reference itself cannot run here: no Julia).

    python tests/golden/make_golden.py

bullhead_64x64.npz : Bull-head L[1,1]/status on a 64x64 sub-grid of BASELINE config C3 (numpy/LAPACK oracle)
bessel_z1000.npz   : A \\ [besselj(1,z); 0...] for z = 1000 (C oracle), with ncols/nconv
chol_pentadiag.npz : BASELINE config C4 operator at 4 shifts: adaptive Cholesky solve (C oracle)
blocktri.npz       : periodic Schrödinger block QL at 16 energies (C5) + the complex `-5im*I` case (numpy oracle)
ql_solve.npz       : ql(A - λI) \\ b for the tridiagonal symbol of test/test_infql.jl:143-145 and the Bull-head symbol
"""
import os
import sys

import numpy as np
from scipy.special import jv

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import banded as ob          # noqa: E402
from oracle import toeplitz_ql as tq     # noqa: E402

x = np.linspace(-10.0, 7.0, 64)
y = np.linspace(-7.0, 8.0, 64)
lam = x[None, :] + 1j * y[:, None]
L11, status = tq.ql_toeplitz_l11(np.array([2j, 0, 0, 1, 0.7]), lam)
np.savez_compressed(os.path.join(HERE, "bullhead_64x64.npz"), lam=lam, L11=L11, status=status)

z = 1000.0
p, q, r = ob.bessel_operator(z)
o = ob.qr_banded_solve(1, 1, p, q, r, b=[jv(1, z)], ncap=8000)
n = int(o["ncols"][0])
np.savez_compressed(os.path.join(HERE, "bessel_z1000.npz"), z=z, b=jv(1, z), x=o["x"][0, :n], ncols=n,
                    nconv=int(o["nconv"][0]))

sigma = np.array([0.01, 1.0, 5.0, 10.0])
p4 = np.tile(np.array([1.0, -4.0, 6.0]), (4, 1)); p4[:, 2] += sigma
q4 = np.zeros((4, 3)); q4[:, 2] = 1e-4
o = ob.chol_banded_solve(2, p4, q4, b=[1.0], ncap=8000)
nmax = int(o["ncols"].max())
np.savez_compressed(os.path.join(HERE, "chol_pentadiag.npz"), sigma=sigma, x=o["x"][:, :nmax], ncols=o["ncols"],
                    nconv=o["nconv"], status=o["status"])

from oracle import block_ql as bq        # noqa: E402
xs = np.concatenate([np.linspace(-0.95, 0.95, 8), np.linspace(2.25, 4.0, 8)])
res = [bq.periodic_schrodinger(x) for x in xs]
c2, a2, b2 = np.array([[0, 0.5], [0, 0]]), np.array([[0, 2.0], [0.5, 0]]), np.array([[0, 0.0], [2.0, 0]])
xc = np.array([5j, 2.0 + 4.0j, -0.5 - 3.0j])
resc = [bq.blocktri_ql_l11([c2], [a2], [b2], c2, a2, b2, shift=x, maxit=5000) for x in xc]
np.savez_compressed(os.path.join(HERE, "blocktri.npz"), energies=xs, L11=np.array([r[0] for r in res]),
                    iters=np.array([r[1] for r in res]), status=np.array([r[2] for r in res]),
                    c_energies=xc, c_L11=np.array([r[0] for r in resc]), c_iters=np.array([r[1] for r in resc]))

col3 = np.array([0.5, 5.0, 2.0])
x3, s3 = tq.ql_toeplitz_solve(col3, np.array([0.0, 1.0, -2.5]), np.array([1.0]), 64)
colb = np.array([2j, 0, 0, 1, 0.7])
lamb = np.array([5.0 + 1j, -3 - 0.1j, 2 + 3j, 0.1 + 0.1j])
xb, sb = tq.ql_toeplitz_solve(colb, lamb, np.array([1.0, 0.5j, -0.25]), 32)
np.savez_compressed(os.path.join(HERE, "ql_solve.npz"), tri_col=col3, tri_lam=np.array([0.0, 1.0, -2.5]), tri_x=x3,
                    tri_status=s3, bull_col=colb, bull_lam=lamb, bull_b=np.array([1.0, 0.5j, -0.25]), bull_x=xb,
                    bull_status=sb)
print("wrote fixtures")

# triconj.npz : Tridiagonal(U*X/U) for the P -> C^(3/2) and P^(1,0) -> P^(2,0) pairs of test_bidiagonalconjugation.jl:111-120
#               (n = 200) and the Cholesky case (:159-164, n = 100)
# blockqr.npz : adaptive block-banded QR solves of test_infqr.jl:270-312 (x prefixes, Q'b prefixes, counts)
sys.path.insert(0, os.path.dirname(HERE))
from triconj_cases import cholesky_case, reference_pairs   # noqa: E402
from oracle import block_qr as bqr                          # noqa: E402
pairs = reference_pairs(203)
Y1 = ob.tridiagonal_conjugation(pairs[1][1], pairs[1][2], 200)[0]
Y2 = ob.tridiagonal_conjugation(pairs[2][1], pairs[2][2], 200)[0]
u_, p_, q_, head_, X_, M_ = cholesky_case(110)
cg_ = (100 + 4) // 2
oc = ob.chol_banded_solve(u_, p_, q_, head=head_, b=[1.0], tol=np.inf, colgrowth=cg_, ncap=2 * cg_ + 3 * u_ + 2, want_factors=True)
Uc = np.stack([oc["factors"][0][d:102 + d, u_ - d] for d in range(3)])
Y3 = ob.tridiagonal_conjugation(Uc, X_, 100)[0]
np.savez_compressed(os.path.join(HERE, "triconj.npz"), Y_P_C32=Y1, Y_P10_P20=Y2, Y_chol=Y3)
T_ = bqr.toeplitz_bands(0.0, 0.5, 80)
cases = [(1.0, 0.0, 2.0), (0.0, 1.0, 2.0), (1.0, 1.0, 4.0), (2.0, 1.0, 8.0)]
outs = [bqr.block_qr_solve(T_, T_, *c) for c in cases]
np.savez_compressed(os.path.join(HERE, "blockqr.npz"), cases=np.array(cases), ncols=np.array([o["ncols"] for o in outs]),
                    nblocks=np.array([o["nblocks"] for o in outs]), x=np.stack([o["x"][:300] for o in outs]),
                    qtb=np.stack([o["qtb"][:300] for o in outs]))
print("wrote triconj / blockqr fixtures")

# almostbanded.npz : adaptive QR solves of the almost-banded operators of test_infqr.jl:194-266 (x prefixes, counts)
from oracle import almostbanded_qr as abq                   # noqa: E402
ONES_, ALT_ = [1, 1, 0, 0], [-1, 1, 0, 0]
o1 = abq.almostbanded_qr_solve(0, 1, [[1, 1, 0], [-1, 0, 0]], [ONES_], [np.e], ncap=3200)
o2 = abq.almostbanded_qr_solve(0, 2, [[2, 3, 1], [0, 0, 0], [-1, 0, 0]], [ONES_, ALT_], [np.e, 1 / np.e], ncap=3200)
dg5 = [[1, 0, 0], [0, 0, 0], [1, 0, 0], [4, 2, 0], [1, 0, 0], [0, 0, 0], [1, 0, 0]]
o3 = abq.almostbanded_qr_solve(1, 5, dg5, [ONES_, ALT_], [1.0, 2.0], ncap=4200)
np.savez_compressed(os.path.join(HERE, "almostbanded.npz"), x1=o1["x"][:400], x2=o2["x"][:400], x3=o3["x"][:400],
                    ncols=np.array([o1["ncols"], o2["ncols"], o3["ncols"]]), nconv=np.array([o1["nconv"], o2["nconv"], o3["nconv"]]))
print("wrote almostbanded fixture")
