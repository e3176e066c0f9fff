"""C4 (2048 pentadiagonal Cholesky shifts): forward / back-substitution device time."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import ila_b200
sigma = 0.01 + (10 - 0.01) * np.arange(2048) / 2047
p4, q4 = ila_b200.pentadiagonal_operator(sigma)
plan = ila_b200.QrBandedPlan(0, 2, p4, q4, b=[1.0], ncap=8000, kind="chol")
def t(fn, n=7):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
f = t(plan.run_factor); b = t(plan.run_backsolve); a = t(plan.run)
cols = int(plan.d_nswept.max().item())
print(f"C4 forward {f:.3f} ms ({f*1e-3*1.965e9/cols:.0f} cycles/column of the longest, {cols} columns)  back-substitution {b:.3f} ms  run {a:.3f} ms  bad {int((plan.d_status != 0).sum())}")
