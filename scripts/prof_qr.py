"""Small QR run for ncu (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ila_b200
from scipy.special import jv
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
zs = np.full(nb, 3000.0)
p, q, r = ila_b200.bessel_operator(zs)
plan = ila_b200.QrBandedPlan(1, 1, p, q, r, b=jv(1, zs)[:, None], ncap_per=ila_b200.bessel_ncap(zs))
for _ in range(2):
    plan.run()
torch.cuda.synchronize()
print("ok", int(plan.d_nswept.sum().item()))
