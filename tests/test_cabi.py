"""The C-ABI library builds, loads and exports every symbol include/ila_b200.h declares (no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ila_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ila_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(ila):
    syms = _declared_symbols()
    assert len(syms) >= 15
    L = ctypes.CDLL(ila.SO_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ila_b200.h but not exported"
    # the ctypes table covers the header and nothing else
    from ila_b200 import _lib
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_no_cpu_fallback(ila):
    L = ila.lib()
    assert L.ila_version() >= 100
    if L.ila_device_count() == 0:
        with pytest.raises(ila.IlaError) as e:
            ila.ql_toeplitz_l11([2j, 0, 0, 1, 0.7], np.array([5.0 + 0j]))
        assert e.value.code == -3
        with pytest.raises(ila.IlaError):
            ila.qr_banded_solve(1, 1, [[1, 1, 1]], [[0, 1, 0]], b=[1.0], ncap=3000)


def test_product_path_never_imports_oracle():
    """The shipped package must not reference oracle/ (prompt ③): grep the package sources."""
    pkg = os.path.join(ROOT, "infinitelinearalgebra.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "libila_oracle" not in src, f


def test_work_layout_host_helper(ila):
    import ctypes
    from ila_b200 import _lib
    caps = np.array([5, 8, 13] + [3] * 31, dtype=np.int64)          # 34 instances -> 2 groups of 32
    gbase = np.zeros(3, dtype=np.int64)
    wb, xb = ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check(ila.lib().ila_qr_banded_work_layout(1, 1, 34, 0, _lib.ptr(caps), _lib.ptr(gbase), ctypes.byref(wb),
                                                   ctypes.byref(xb)))
    assert list(gbase) == [0, 16, 24]                               # group capacities rounded up to 8 rows
    assert xb.value == 24 * 32 * 8 and wb.value == 24 * 32 * 8 * 6  # 6 doubles per record with reflectors (l=u=1)


def test_new_entries_refuse_to_run_without_a_gpu(ila):
    """K10 / K11 / K12 host entries: ILA_ERR_NOGPU (-3) on a box without a CUDA device — never a CPU fallback."""
    if ila.lib().ila_device_count() != 0:
        pytest.skip("a CUDA device is visible")
    U = np.ones((1, 3, 12)); X = np.ones((3, 12))
    T = np.stack([np.full(80, 0.5), np.zeros(80), np.full(80, 0.5)])
    calls = [lambda: ila.tridiagonal_conjugation(U, X, 10),
             lambda: ila.bidiagonal_conjugation(U, U, 10),
             lambda: ila.chol_tridiagonal_conjugation(1, [[0.0, 2.0]], [[0.0, 0.0]], X, 10),
             lambda: ila.qr_blockbanded_solve(T, T, 1.0, 1.0, [4.0]),
             lambda: ila.qr_almostbanded_solve(0, 1, [[1, 1, 0], [-1, 0, 0]], [[1, 1, 0, 0]], [np.e])]
    for f in calls:
        with pytest.raises(ila.IlaError) as e:
            f()
        assert e.value.code == -3
