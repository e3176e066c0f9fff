import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ila():
    import ila_b200
    return ila_b200


def bullhead_grid(nx, ny):
    """λ grid of BASELINE config C3 (window of examples/toeplitz.jl:72)."""
    import numpy as np
    x = np.linspace(-10.0, 7.0, nx)
    y = np.linspace(-7.0, 8.0, ny)
    return x[None, :] + 1j * y[:, None]


BULLHEAD_COLUMN = [2j, 0.0, 0.0, 1.0, 0.7]   # data column of BandedMatrix(-3=>7/10, -2=>1, 1=>2im): +1, 0, -1, -2, -3


def pytest_sessionfinish(session, exitstatus):
    """With ILA_B200_DEBUG_LIB=1 the whole suite runs on the bounds-checked debug build (csrc `make debug`): report what the
    checked pointers recorded and fail the session if anything was out of range."""
    if os.environ.get("ILA_B200_DEBUG_LIB") != "1":
        return
    try:
        import ctypes
        from ila_b200 import _lib
        if not _lib.SO_PATH.endswith("libila_b200_dbg.so"):
            return
        c, s, i, e = (ctypes.c_int64(0) for _ in range(4))
        _lib.lib().ila_debug_bounds_report(ctypes.byref(c), ctypes.byref(s), ctypes.byref(i), ctypes.byref(e))
        msg = f"bounds-checked build: {c.value} out-of-range accesses recorded (first: site {s.value}, index {i.value}, extent {e.value})"
        print("\n" + msg)
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            open(os.path.join(ROOT, "gpurun_out", "r02_bounds_full_suite.txt"), "w").write(msg + "\n")
        except OSError:
            pass
        if c.value != 0:
            session.exitstatus = 1
    except Exception as exc:   # the report is a diagnostic: never mask the suite's own result
        print("\nbounds report unavailable:", exc)
