"""Randomised parity sweep of K3 / K7 against the oracle for every supported lower bandwidth (not a pytest; run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ila_b200
from oracle import toeplitz_ql as tq

rng = np.random.default_rng(2026)
worst = 0.0
bad = 0
for l in range(1, 8):
    for trial in range(6):
        cplx = trial % 2 == 0
        col = rng.normal(size=l + 2) + (1j * rng.normal(size=l + 2) if cplx else 0)
        col[0] = col[0] if abs(col[0]) > 0.3 else 1.0
        lam = (rng.normal(size=400) * 3 + 1j * rng.normal(size=400) * 3) if cplx else rng.normal(size=400) * 4
        for branch in (1, 2) if l >= 1 else (1,):
            Lg, sg = ila_b200.ql_toeplitz_l11(col, lam, branch_rank=branch)
            # the C ABI's documented sign rule for the real-typed path ("householder" = the finite-section limit)
            Lo, so = tq.ql_toeplitz_l11(col, lam, branch_rank=branch, sign_rule="householder")
            mism = int((sg != so).sum())
            ok = (sg == 0) & (so == 0)
            scale = np.abs(col[0])
            a, b = Lg[ok], Lo[ok]                      # signed comparison on both paths
            flips = int((np.sign(Lg[ok].real) != np.sign(Lo[ok].real)).sum())
            err = np.abs(a - b)
            rel = err / np.maximum(np.abs(Lo[ok]), 0.05 * scale)
            w = float(rel.max()) if ok.any() else 0.0
            worst = max(worst, w)
            if mism or w > 1e-11 or flips:
                bad += (mism > 0 or w > 1e-11)
                print(f"l={l} trial={trial} cplx={cplx} branch={branch}: status mismatches {mism} (of {lam.size}), worst rel {w:.2e}, sign flips {flips} of {int(ok.sum())}")
print("worst relative error", worst, "cases flagged", bad)
