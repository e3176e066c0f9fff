"""C4 workload run (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ila_b200
sigma = 0.01 + (10 - 0.01) * np.arange(2048) / 2047
p4, q4 = ila_b200.pentadiagonal_operator(sigma)
plan = ila_b200.QrBandedPlan(0, 2, p4, q4, b=[1.0], ncap=8000, kind="chol")
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
print("ok", int(plan.d_nswept.sum().item()), int((plan.d_status != 0).sum().item()), plan.d_nconv[:4].tolist(), plan.d_nconv[-4:].tolist())
