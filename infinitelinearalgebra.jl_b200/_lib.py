"""ctypes binding of libila_b200.so (the C ABI declared in include/ila_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C infinitelinearalgebra.jl_b200/csrc`.
There is no CPU fallback: if the shared object is missing, loading fails loudly; if no CUDA device is
visible, every compute entry returns ILA_ERR_NOGPU and the wrappers raise `IlaError`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# ILA_B200_DEBUG_LIB=1 selects the bounds-checked debug build (`make -C csrc debug`; the compute-sanitizer substitute)
SO_PATH = os.path.join(_HERE, "libila_b200_dbg.so" if os.environ.get("ILA_B200_DEBUG_LIB") == "1" else "libila_b200.so")
if os.environ.get("ILA_B200_SO"):        # explicit override (profiling builds)
    SO_PATH = os.environ["ILA_B200_SO"]
CSRC = os.path.join(_HERE, "csrc")

ILA_OK, ILA_NO_QL, ILA_NOT_REAL, ILA_NOT_CONVERGED, ILA_NOT_POSDEF, ILA_NAN = range(6)
ILA_ERR_CUDA, ILA_ERR_ARG, ILA_ERR_NOGPU, ILA_ERR_UNSUPPORTED, ILA_ERR_CAPACITY = -1, -2, -3, -4, -5


class IlaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libila_b200 error {code}: {msg}")
        self.code = code


class c64(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_sp = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); kept in one table so tests can check every declared symbol is exported
SIGNATURES = {
    "ila_version": (C.c_int, []),
    "ila_device_count": (C.c_int, []),
    "ila_last_error": (C.c_char_p, []),
    "ila_set_device": (C.c_int, [C.c_int]),
    "ila_get_device": (C.c_int, []),
    "ila_kernel_launches": (C.c_int64, []),
    "ila_fp64_peak_probe": (C.c_int, [_dp, _dp]),
    "ila_fp64_latency_probe": (C.c_int, [_vp]),
    "ila_ql_toeplitz_l11_c64": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp, C.c_int]),
    "ila_ql_toeplitz_l11_f64": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp, C.c_int]),
    "ila_ql_toeplitz_l11_c64_dev": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp, _vp]),
    "ila_ql_toeplitz_l11_f64_dev": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp, _vp]),
    "ila_ql_toeplitz_absl11_grid_c64": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int64, C.c_int, _vp, C.c_int]),
    "ila_ql_toeplitz_absl11_grid_c64_dev": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int64, C.c_int, _vp, _vp]),
    "ila_ql_perttoeplitz_l11_c64": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp, C.c_int]),
    "ila_ql_perttoeplitz_l11_f64": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp, C.c_int]),
    "ila_ql_blocktri_l11_f64": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int64,
                                          C.c_double, C.c_int, _vp, _vp, _vp, C.c_int]),
    "ila_ql_blocktri_l11_f64_dev": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp,
                                              C.c_int64, C.c_double, C.c_int, _vp, _vp, _vp, _vp]),
    "ila_ql_blocktri_l11_c64": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int64,
                                          C.c_double, C.c_int, _vp, _vp, _vp, C.c_int]),
    "ila_ql_blocktri_l11_c64_dev": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp,
                                              C.c_int64, C.c_double, C.c_int, _vp, _vp, _vp, _vp]),
    "ila_ql_toeplitz_solve_c64": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp, C.c_int]),
    "ila_ql_toeplitz_solve_f64": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp, C.c_int]),
    "ila_ql_toeplitz_applyq_c64": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp, _vp, C.c_int]),
    "ila_ql_toeplitz_applyq_f64": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp, _vp, C.c_int]),
    "ila_ql_toeplitz_solve_c64_dev": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "ila_ql_toeplitz_solve_f64_dev": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "ila_ql_perttoeplitz_solve_c64": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp, C.c_int]),
    "ila_ql_perttoeplitz_solve_f64": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp, C.c_int]),
    "ila_qr_banded_solve_c64": (C.c_int, [C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp,
                                          C.c_double, C.c_int, C.c_int64, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int]),
    "ila_revchol_tridiag_f64": (C.c_int, [C.c_int64, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int]),
    "ila_triconj_bands_f64": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, C.c_int64, C.c_int, _vp, _vp, C.c_int]),
    "ila_triconj_bands_f64_dev": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, C.c_int64, C.c_int, _vp, _vp, _vp]),
    "ila_triconj_records_f64_dev": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, _vp, _vp]),
    "ila_chol_triconj_f64": (C.c_int, [C.c_int, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int64, _vp, C.c_int64, C.c_int,
                                       _vp, _vp, C.c_int]),
    "ila_biconj_bands_f64": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _vp, _vp, C.c_int64, _vp, _vp, C.c_int]),
    "ila_qr_blockbanded_solve_f64": (C.c_int, [C.c_int64, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, _vp, C.c_double, C.c_int,
                                               C.c_int, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int]),
    "ila_qr_blockbanded_work_doubles": (C.c_int, [C.c_int, _vp]),
    "ila_qr_blockbanded_set_mode": (C.c_int, [C.c_int]),
    "ila_qr_blockbanded_solve_f64_dev": (C.c_int, [C.c_int64, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, _vp, C.c_double, C.c_int,
                                                   C.c_int, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "ila_qr_almostbanded_solve_f64": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, C.c_int, _vp, C.c_double, C.c_int,
                                                C.c_int64, _vp, _vp, _vp, _vp, C.c_int]),
    "ila_ql_blocktri_factors_f64": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int64,
                                              C.c_double, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int64, _vp, _vp, _vp,
                                              C.c_int]),
    "ila_ql_blocktri_factors_c64": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int64,
                                              C.c_double, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int64, _vp, _vp, _vp,
                                              C.c_int]),
    "ila_release": (C.c_int, []),
    "ila_ql_finite_section_f64": (C.c_int, [C.c_int, C.c_int, C.c_int64, _vp, _vp, C.c_int, _vp, C.c_int64, C.c_double, C.c_int64,
                                            _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "ila_ql_finite_section_c64": (C.c_int, [C.c_int, C.c_int, C.c_int64, _vp, _vp, C.c_int, _vp, C.c_int64, C.c_double, C.c_int64,
                                            _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "ila_ql_set_shifts_per_thread": (C.c_int, [C.c_int]),
    "ila_ql_set_max_sweeps": (C.c_int, [C.c_int, C.c_int]),
    "ila_ql_set_cold_start": (C.c_int, [C.c_int]),
    "ila_debug_bounds_report": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ila_ql_toeplitz_l11re_c64": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, C.c_int]),
    "ila_ql_toeplitz_l11re_c64_dev": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp]),
    "ila_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "ila_host_free": (C.c_int, [_vp]),
    "ila_host_register": (C.c_int, [_vp, C.c_int64]),
    "ila_host_unregister": (C.c_int, [_vp]),
    "ila_set_host_copy_mode": (C.c_int, [C.c_int]),
    "ila_set_host_chunking": (C.c_int, [C.c_int64, C.c_int]),
    "ila_qr_banded_solve_f64": (C.c_int, [C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp,
                                          C.c_double, C.c_int, C.c_int64, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp,
                                          _vp, _vp, _vp, C.c_int]),
    "ila_qr_set_arith_mode": (C.c_int, [C.c_int]),
    "ila_chol_banded_solve_f64": (C.c_int, [C.c_int, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_double,
                                            C.c_int, C.c_int64, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "ila_chol_banded_factor_f64_dev": (C.c_int, [C.c_int, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp,
                                                 C.c_double, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                 _vp, C.c_int, _vp]),
    "ila_qr_banded_state_doubles": (C.c_int, [C.c_int, C.c_int]),
    "ila_chol_banded_state_doubles": (C.c_int, [C.c_int]),
    "ila_qr_banded_factor_resume_f64_dev": (C.c_int, [C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int,
                                                      _vp, C.c_double, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                      _vp, C.c_int, _vp, C.c_int, _vp]),
    "ila_chol_banded_factor_resume_f64_dev": (C.c_int, [C.c_int, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp,
                                                        C.c_double, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                        _vp, C.c_int, _vp, C.c_int, _vp]),
    "ila_qr_banded_partialqr_f64_dev": (C.c_int, [C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int64,
                                                  C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "ila_qr_banded_export_f64_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ila_qr_banded_work_layout": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp]),
    "ila_qr_banded_factor_f64_dev": (C.c_int, [C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int,
                                               _vp, C.c_double, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                               _vp, C.c_int, _vp]),
    "ila_qr_banded_backsolve_f64_dev": (C.c_int, [C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                  C.c_int64, C.c_int, _vp]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a with the Makefile in csrc/ (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j4"]
    if not verbose:
        cmd.append("-s")
    subprocess.check_call(cmd)
    return SO_PATH


def lib():
    """Load libila_b200.so; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise FileNotFoundError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise IlaError(rc, lib().ila_last_error().decode("utf-8", "replace"))


def ptr(arr):
    """Raw address of a numpy array / torch tensor / None as c_void_p."""
    if arr is None:
        return None
    if hasattr(arr, "data_ptr"):
        return C.c_void_p(arr.data_ptr())
    return C.c_void_p(arr.ctypes.data)
