"""C2 workload run (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ila_b200
from scipy.special import jv
zs = 1e3 + (1e5 - 1e3) * np.arange(4096) / 4095
p, q, r = ila_b200.bessel_operator(zs)
plan = ila_b200.QrBandedPlan(1, 1, p, q, r, b=jv(1, zs)[:, None], ncap_per=ila_b200.bessel_ncap(zs))
for _ in range(2):
    plan.run()
torch.cuda.synchronize()
print("ok", int(plan.d_nswept.sum().item()), int((plan.d_status != 0).sum().item()))
