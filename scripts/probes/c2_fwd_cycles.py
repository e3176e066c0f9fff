"""C2 forward sweep / back-substitution: device ms per launch and cycles per column of the longest instance (z = 1e5).
ILA_B200_SO selects the library (the probe build with -DILA_QR_PROBE runs the dependent chain only and never converges:
its column count is the capacity)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from scipy.special import jv
import ila_b200
zs = 1e3 + (1e5 - 1e3) * np.arange(4096) / 4095
p, q, r = ila_b200.bessel_operator(zs)
caps = ila_b200.bessel_ncap(zs)
plan = ila_b200.QrBandedPlan(1, 1, p, q, r, b=jv(1, zs)[:, None], ncap_per=caps)
def t(fn, n=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
f = t(plan.run_factor)
nsw = plan.d_nswept.cpu().numpy()
st = plan.d_status.cpu().numpy()
ncol = int(nsw.max()) if nsw is not None and nsw.max() > 0 else int(caps.max())
if st is not None and (st != 0).all(): ncol = int(caps.max())
clk = 1.965e9
print(f"forward {f:.3f} ms  longest {ncol} columns  {f*1e-3*clk/ncol:.1f} cycles/column (at 1965 MHz)   status0 {None if st is None else int((st==0).sum())}")
if st is not None and (st == 0).all():
    b = t(plan.run_backsolve)
    print(f"backward {b:.3f} ms  {b*1e-3*clk/ncol:.1f} cycles/row")
# the same back-substitution through the interleaved stream + compaction pass
from ila_b200 import _lib
_lib.check(_lib.lib().ila_qr_set_arith_mode(3))
b3 = t(plan.run_backsolve)
_lib.check(_lib.lib().ila_qr_set_arith_mode(2))
print(f"backward (interleaved + compaction) {b3:.3f} ms")

import json
json.dump({"forward_ms": f, "backward_ms": b, "backward_interleaved_ms": b3, "columns": ncol}, open(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "gpurun_out", "c2_cycles.json"), "w"))
