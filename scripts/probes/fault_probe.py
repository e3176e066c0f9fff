import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scipy.special import jv
import ila_b200
from ila_b200 import _lib
lib = _lib.lib()
z = np.linspace(40.0, 2500.0, 32)
p, q, r = ila_b200.bessel_operator(z)
for every in (0, 4001, 257, 3, 1):
    _lib.check(lib.ila_qr_set_arith_mode(11)); _lib.check(lib.ila_qr_set_arith_mode(1000 + every))
    print("fault every", every, flush=True)
    g = ila_b200.qr_banded_solve(1, 1, p, q, r, b=jv(1, z)[:, None], ncap_per=ila_b200.bessel_ncap(z))
