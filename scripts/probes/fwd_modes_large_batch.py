"""Forward sweep, one warp per group (mode 7) against three warps per group (mode 11: forced), on batches of different sizes: the three-warp
sweep is built for the latency-bound regime (about one group per SM); with many groups per SM its polling warps compete with
other groups' working warps."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from scipy.special import jv
import ila_b200
from ila_b200 import _lib
lib = _lib.lib()
def t(fn, n=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
out = {}
for name, zs in (("4096 z in [1e3,1e5]", 1e3 + (1e5 - 1e3) * np.arange(4096) / 4095), ("8192 x z=8000", np.full(8192, 8000.0)),
                 ("16384 x z=4000", np.full(16384, 4000.0)), ("32768 x z=2000", np.full(32768, 2000.0))):
    p, q, r = ila_b200.bessel_operator(zs)
    plan = ila_b200.QrBandedPlan(1, 1, p, q, r, b=jv(1, zs)[:, None], ncap_per=ila_b200.bessel_ncap(zs))
    for mode in (7, 11):
        _lib.check(lib.ila_qr_set_arith_mode(mode))
        out[f"{name} / mode {mode}"] = t(plan.run_factor)
        print(f"{name:24s} mode {mode}: forward {out[f'{name} / mode {mode}']:.3f} ms", flush=True)
    del plan
_lib.check(lib.ila_qr_set_arith_mode(8))
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "gpurun_out", "r02_fwd_modes_batch.json"), "w"), indent=1)
