"""C2 back-substitution with the two ring loaders (LDGSTS per lane / cp.async.bulk + mbarrier): device time per launch."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.special import jv
import ila_b200
from ila_b200 import _lib
lib = _lib.lib()
out = {}
for name, zs in (("C2 4096 z in [1e3,1e5]", 1e3 + (1e5 - 1e3) * np.arange(4096) / 4095), ("32768 x z=2000", np.full(32768, 2000.0))):
    p, q, r = ila_b200.bessel_operator(zs)
    plan = ila_b200.QrBandedPlan(1, 1, p, q, r, b=jv(1, zs)[:, None], ncap_per=ila_b200.bessel_ncap(zs))
    plan.run_factor(); torch.cuda.synchronize()
    for mode, label in ((5, "ldgsts"), (6, "bulk")):
        _lib.check(lib.ila_qr_set_arith_mode(mode))
        for _ in range(2): plan.run_backsolve()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); plan.run_backsolve(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        out[f"{name} / {label}"] = float(np.median(ts))
        print(f"{name:28s} {label:7s} back-substitution {np.median(ts):8.3f} ms", flush=True)
    # forward for context
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); plan.run_factor(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    out[f"{name} / forward"] = float(np.median(ts))
    print(f"{name:28s} forward sweep {np.median(ts):8.3f} ms")
    del plan
_lib.check(lib.ila_qr_set_arith_mode(6))
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_c2_loader.json"), "w"), indent=1)
