"""infinitelinearalgebra.jl_b200 — B200-native adaptive-factorization hot path of InfiniteLinearAlgebra.jl.

The directory name contains a dot, so import it through the top-level alias module:

    import ila_b200                                  # loads this package as `ila_b200`
    L11, status = ila_b200.ql_toeplitz_l11(col, lambdas)

Contents: `csrc/` (hand-written sm_100a CUDA + the C ABI of include/ila_b200.h), `_lib` (ctypes binding),
`batched` (batched numpy/torch-pointer entry points), `api` (host-side mirror of the reference's Julia interface).
"""
from . import _lib, api, sharding
from ._lib import IlaError, SO_PATH, build, lib
from .batched import (FLOATMIN, QrBandedPlan, bessel_ncap, bessel_operator, chol_banded_solve, pentadiagonal_operator, fp64_peak_tflops, periodic_schrodinger_blocks, ql_blocktri_l11, ql_blocktri_factors, ql_perttoeplitz_l11, ql_perttoeplitz_solve, ql_toeplitz_absl11_grid, ql_toeplitz_l11, ql_toeplitz_l11_real, status_from_nan_payload, ql_toeplitz_solve, ql_toeplitz_applyq,
                      ql_toeplitz_l11_dev, qr_banded_solve, qr_banded_solve_c64, revchol_tridiag, ql_finite_section, tridiagonal_conjugation,
                      chol_tridiagonal_conjugation, bidiagonal_conjugation, qr_blockbanded_solve, qr_almostbanded_solve)

__all__ = ["IlaError", "SO_PATH", "build", "lib", "FLOATMIN", "QrBandedPlan", "bessel_ncap", "bessel_operator", "chol_banded_solve", "pentadiagonal_operator", "fp64_peak_tflops", "periodic_schrodinger_blocks", "ql_blocktri_l11", "ql_blocktri_factors", "ql_perttoeplitz_l11", "ql_perttoeplitz_solve", "ql_toeplitz_absl11_grid", "ql_toeplitz_solve", "ql_toeplitz_applyq",
           "ql_toeplitz_l11", "ql_toeplitz_l11_real", "status_from_nan_payload", "ql_toeplitz_l11_dev", "qr_banded_solve", "qr_banded_solve_c64", "revchol_tridiag", "ql_finite_section", "tridiagonal_conjugation",
           "chol_tridiagonal_conjugation", "bidiagonal_conjugation", "qr_blockbanded_solve", "qr_almostbanded_solve"]
