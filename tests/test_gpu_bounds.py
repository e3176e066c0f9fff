"""compute-sanitizer substitute (the tool is closed on the GPU pool): every hot kernel on small cases under the
bounds-checked debug build of the library (`-DILA_BOUNDS`, libila_b200_dbg.so).  The kernels with non-trivial index
arithmetic — K1/K2 group-interleaved records and ring slots, K5 records, K11's circular shared-memory panel, R slab and
back-substitution scratch — address their buffers through checked pointers that RECORD an out-of-range index instead of
performing it; the run must finish with zero recorded failures, and a deliberately undersized workspace must be caught."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECK = r'''
import ctypes, os, sys
sys.path.insert(0, os.environ["ILA_ROOT"])
__file__ = os.path.join(os.environ["ILA_ROOT"], "scripts", "sanitize_all.py")
exec(open(__file__).read())
from ila_b200 import _lib
assert _lib.SO_PATH.endswith("libila_b200_dbg.so")
c, s, i, e = (ctypes.c_int64(0) for _ in range(4))
_lib.check(_lib.lib().ila_debug_bounds_report(ctypes.byref(c), ctypes.byref(s), ctypes.byref(i), ctypes.byref(e)))
print("BOUNDS", c.value, s.value, i.value, e.value)
assert c.value == 0, (c.value, s.value, i.value, e.value)
# negative control: a QR workspace laid out for 600 columns while the sweep is allowed 4000 -> record stores beyond the group
import numpy as np, torch
from scipy.special import jv
z = np.array([900.0]); p, q, r = ila_b200.bessel_operator(z)
plan = ila_b200.QrBandedPlan(1, 1, p, q, r, b=jv(1, z)[:, None], ncap=600)
plan.ncap = 4000             # (the checked pointer redirects the stray stores to element 0: nothing is actually written out of range)
plan.run_factor(); torch.cuda.synchronize()
_lib.check(_lib.lib().ila_debug_bounds_report(ctypes.byref(c), ctypes.byref(s), ctypes.byref(i), ctypes.byref(e)))
print("NEGATIVE", c.value, s.value, i.value, e.value)
assert c.value > 0 and s.value in (1201, 1205)   # the record pointer of the one-warp sweep / of the three-warp sweep's clerk
print("bounds ok")
'''


def test_all_kernels_under_the_bounds_checked_build():
    so = os.path.join(ROOT, "infinitelinearalgebra.jl_b200", "libila_b200_dbg.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "infinitelinearalgebra.jl_b200", "csrc"), "-j4", "-s", "debug"])
    env = dict(os.environ, ILA_B200_DEBUG_LIB="1", ILA_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-c", CHECK], capture_output=True, text=True, timeout=900, env=env)
    out = r.stdout + r.stderr
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        open(os.path.join(ROOT, "gpurun_out", "r02_bounds_check.txt"), "w").write(out)
    except OSError:
        pass
    assert r.returncode == 0 and "bounds ok" in out, out[-3000:]
