"""Generates the committed golden fixtures from the oracle (the reference itself cannot run here: no Julia).

    python tests/golden/make_golden.py

bullhead_64x64.npz : Bull-head L[1,1]/status on a 64x64 sub-grid of BASELINE config C3 (numpy/LAPACK oracle)
bessel_z1000.npz   : A \\ [besselj(1,z); 0...] for z = 1000 (C oracle), with ncols/nconv
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
print("wrote fixtures")
