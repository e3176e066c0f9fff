"""Randomized stress of the three-warp forward sweep against the one-warp sweep (bit for bit): many random l = 1 configurations,
with and without fault injection.  Not a test (minutes); writes gpurun_out/r02_stress_qr_sweeps.txt.
   python scripts/stress_qr_sweeps.py [trials]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ila_b200
from ila_b200 import _lib
lib = _lib.lib()
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rng = np.random.default_rng(77)
t0 = time.time()
bad = 0
cols = 0
stat = {}
for trial in range(trials):
    u = int(rng.integers(0, 4)); nd = u + 2
    batch = int(rng.integers(1, 130))
    scale = 10.0 ** rng.uniform(-4, 4)
    p = rng.normal(size=(batch, nd)) * scale
    p[:, u] += rng.choice([0.0, 2.0, 6.0]) * scale * np.sign(p[:, u])
    p[:, u + 1] = scale * rng.choice([-2.0, -1.0, 0.25, 1.0, 3.0, 1e-6, 1e5], size=batch)
    q = np.zeros((batch, nd)); q[:, u] = scale * rng.choice([-1.0, 1.0], size=batch) * 10.0 ** rng.uniform(-4, 1, size=batch)
    r = np.ones((batch, nd))
    if rng.random() < 0.5: r[:, u] = rng.uniform(0.2, 5.0, size=batch)
    shift = rng.normal(size=batch) * scale if rng.random() < 0.7 else None
    nb = int(rng.integers(1, 8))
    b = rng.normal(size=(batch, nb)) * 10.0 ** rng.uniform(-6, 6)
    kw = dict(shift=shift, b=b, ncap=int(rng.choice([600, 5000, 20000])), colgrowth=int(rng.choice([4, 9, 64, 1000])), tol=float(rng.choice([1e-300, 1e-60, 1e-10])))
    fault = int(rng.choice([0, 0, 0, 1, 7, 100, 3001]))
    outs = []
    for mode in (7, 11):
        _lib.check(lib.ila_qr_set_arith_mode(mode))
        _lib.check(lib.ila_qr_set_arith_mode(1000 + (fault if mode == 11 else 0)))
        outs.append(ila_b200.qr_banded_solve(1, u, p, q, r, **kw))
    a, g = outs
    same = (np.array_equal(a["status"], g["status"]) and np.array_equal(a["ncols"], g["ncols"]) and np.array_equal(a["nconv"], g["nconv"])
            and np.array_equal(a["xflat"].view(np.int64), g["xflat"].view(np.int64)))
    bad += 0 if same else 1
    cols += int(a["ncols"].sum())
    for s_ in a["status"]: stat[int(s_)] = stat.get(int(s_), 0) + 1
    if not same: print("MISMATCH trial", trial, u, batch, fault, kw["ncap"], kw["colgrowth"], flush=True)
_lib.check(lib.ila_qr_set_arith_mode(1000)); _lib.check(lib.ila_qr_set_arith_mode(8))
msg = f"{trials} random configurations, {cols} solution entries compared bit for bit, statuses {stat}: {bad} mismatches ({time.time() - t0:.0f} s)"
print(msg)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/r02_stress_qr_sweeps.txt", "w").write(msg + "\n")
sys.exit(1 if bad else 0)
