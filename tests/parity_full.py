"""Full-size parity reports for the BASELINE configs (C3: 1024 x 1024 Bull-head grid, C2: 4096 Bessel solves) — every
unit compared against the oracle, not a sample.  Used by the `-m gpu` tests (assertions) and by
scripts/parity_report.py (prints the summaries kept under profiles/).  Test infrastructure: imports oracle/."""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

BULLHEAD_COLUMN = np.array([2j, 0.0, 0.0, 1.0, 0.7])


def _oracle_chunk(args):
    from oracle import toeplitz_ql as tq
    col, lam, kw = args
    return tq.ql_toeplitz_l11(col, lam, **kw)


def oracle_l11_parallel(col, lam, **kw):
    """The numpy/LAPACK oracle over a flat shift array on all host cores (worker processes around zgeev)."""
    lam = np.asarray(lam)
    flat = lam.reshape(-1)
    nproc = max(1, min(os.cpu_count() or 1, 32))
    if flat.size < 20000 or nproc == 1:
        L, st = _oracle_chunk((np.asarray(col), flat, kw))
        return L.reshape(lam.shape), st.reshape(lam.shape)
    chunks = np.array_split(flat, nproc * 4)
    with mp.get_context("fork").Pool(nproc) as pool:
        res = pool.map(_oracle_chunk, [(np.asarray(col), c, kw) for c in chunks])
    L = np.concatenate([r[0] for r in res]).reshape(lam.shape)
    st = np.concatenate([r[1] for r in res]).reshape(lam.shape)
    return L, st


def c3_full_report(ila, nx=1024, ny=1024):
    """ql(A - λI).L[1,1] on the full C3 grid, GPU (C ABI, host entry) against the oracle at ALL nx*ny shifts.
    Bars (north star 1e-12): statuses identical; |ΔL11| <= 2e-12 |β| everywhere; relative 1e-12 where |L11| >= 0.05 |β|
    (L11 = -sqrt(|μ|²-|β|²) loses relative accuracy like 1/|L11|² next to the symbol curve, in LAPACK as much as here)."""
    x = np.linspace(-10.0, 7.0, nx)
    y = np.linspace(-7.0, 8.0, ny)
    lam = x[None, :] + 1j * y[:, None]
    t0 = time.perf_counter()
    L, st = ila.ql_toeplitz_l11(BULLHEAD_COLUMN, lam)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    Lo, sto = oracle_l11_parallel(BULLHEAD_COLUMN, lam)
    t_cpu = time.perf_counter() - t0
    beta = 2.0
    ok = sto == 0
    err = np.abs(L[ok] - Lo[ok])
    big = np.abs(Lo[ok]) >= 0.05 * beta
    rel = err[big] / np.abs(Lo[ok][big])
    rep = {
        "config": f"C3 Bull-head {nx}x{ny}", "points": int(lam.size), "status_mismatches": int((st != sto).sum()),
        "no_ql_points": int((sto == 1).sum()), "not_converged_points": int((st == 3).sum()),
        "max_abs_err": float(err.max()), "abs_bar": 2e-12 * beta,
        "max_rel_err_where_L11_ge_0.05beta": float(rel.max()), "rel_bar": 1e-12,
        "rel_err_quantiles_all_ok_points": {q: float(np.quantile(err / np.abs(Lo[ok]), float(q))) for q in ("0.5", "0.99", "0.9999", "1.0")},
        "min_abs_L11": float(np.abs(Lo[ok]).min()), "imag_all_zero": bool(np.all(L[ok].imag == 0)),
        "nan_where_status_nonzero": bool(np.all(np.isnan(L[~ok].real))) if (~ok).any() else True,
        "gpu_call_s": t_gpu, "oracle_s": t_cpu, "oracle_procs": int(min(os.cpu_count() or 1, 32)),
    }
    rep["pass"] = bool(rep["status_mismatches"] == 0 and rep["max_abs_err"] <= rep["abs_bar"] and
                       rep["max_rel_err_where_L11_ge_0.05beta"] <= 1e-12 and rep["imag_all_zero"] and
                       rep["nan_where_status_nonzero"] and rep["not_converged_points"] == 0)
    return rep


def c2_z():
    return 1e3 + (1e5 - 1e3) * np.arange(4096) / 4095


def c2_full_report(ila, quad_every=64, chunk=128):
    """The full C2 batch (4096 Bessel-recurrence solves, z in [1e3, 1e5]) against the C oracle: status, ncols, nconv and the
    whole solution vector BIT for bit; and against the __float128 oracle on every `quad_every`-th z with the
    conditioning-aware bound of BASELINE.md §3 (back-substitution amplifies rounding ∝ z: 4 z ε)."""
    from scipy.special import jv
    from oracle import banded as ob
    zs = c2_z()
    p, q, r = ila.bessel_operator(zs)
    b = jv(1, zs)[:, None]
    caps = ila.bessel_ncap(zs)
    t0 = time.perf_counter()
    g = ila.qr_banded_solve(1, 1, p, q, r, b=b, ncap_per=caps)
    t_gpu = time.perf_counter() - t0
    nthreads = os.cpu_count() or 1
    mism = {"status": 0, "ncols": 0, "nconv": 0, "x_instances": 0, "x_entries": 0}
    cols = 0
    t0 = time.perf_counter()
    for lo in range(0, len(zs), chunk):
        hi = min(lo + chunk, len(zs))
        o = ob.qr_banded_solve(1, 1, p[lo:hi], q[lo:hi], r[lo:hi], b=b[lo:hi], ncap=int(caps[lo:hi].max()), nthreads=nthreads)
        mism["status"] += int((g["status"][lo:hi] != o["status"]).sum())
        mism["ncols"] += int((g["ncols"][lo:hi] != o["ncols"]).sum())
        mism["nconv"] += int((g["nconv"][lo:hi] != o["nconv"]).sum())
        for i in range(lo, hi):
            n = int(o["ncols"][i - lo])
            cols += n
            gx = g["x"][i]
            if len(gx) != n:
                mism["x_instances"] += 1
                continue
            d = int((gx.view(np.int64) != o["x"][i - lo, :n].view(np.int64)).sum())
            mism["x_entries"] += d
            mism["x_instances"] += 1 if d else 0
        del o
    t_cpu = time.perf_counter() - t0
    # conditioning-aware report against the __float128 oracle
    idx = np.arange(0, len(zs), quad_every)
    if idx[-1] != len(zs) - 1:
        idx = np.append(idx, len(zs) - 1)
    t0 = time.perf_counter()
    oq = ob.qr_banded_solve(1, 1, p[idx], q[idx], r[idx], b=b[idx], ncap=int(caps[idx].max()), quad=True, nthreads=nthreads)
    t_quad = time.perf_counter() - t0
    eps = np.finfo(float).eps
    kap = []
    for k, i in enumerate(idx):
        n = int(g["ncols"][i])
        xq = oq["x"][k, :n]
        e = np.max(np.abs(g["x"][i] - xq)) / np.max(np.abs(xq))
        kap.append((float(zs[i]), float(e), float(e / (zs[i] * eps))))
    kap = np.array(kap)
    rep = {
        "config": "C2 4096 Bessel solves, z in [1e3,1e5]", "instances": int(len(zs)), "columns_compared": int(cols),
        "mismatches_vs_c_oracle": mism, "bit_exact": bool(all(v == 0 for v in mism.values())),
        "all_converged": bool(np.all(g["status"] == 0)),
        "quad_instances": int(len(idx)), "max_rel_err_vs_float128": float(kap[:, 1].max()),
        "max_rel_err_over_z_eps": float(kap[:, 2].max()), "kappa_bar_z_eps_multiples": 4.0,
        "rel_err_vs_float128_at_z": {f"{kap[j, 0]:.0f}": float(kap[j, 1]) for j in (0, len(kap) // 4, len(kap) // 2, len(kap) - 1)},
        "gpu_call_s": t_gpu, "c_oracle_s": t_cpu, "float128_oracle_s": t_quad, "oracle_threads": int(nthreads),
    }
    rep["pass"] = bool(rep["bit_exact"] and rep["all_converged"] and rep["max_rel_err_over_z_eps"] <= 4.0)
    return rep
