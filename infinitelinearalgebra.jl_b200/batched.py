"""Batched host-side entry points over the C ABI (numpy in / numpy out), plus device-resident plans.

These are the "batched form over a vector of shifts" of the north star: thin argument marshalling around
`ila_ql_toeplitz_l11_*` and `ila_qr_banded_solve_f64`.  All arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib, ptr

FLOATMIN = float(np.finfo(np.float64).tiny)


# ------------------------------------------------------------------------------------------------
# K3: Toeplitz-tail QL
# ------------------------------------------------------------------------------------------------

def ql_toeplitz_l11(data_column, lambdas, branch_rank: int = 1, want_tail: bool = False, ngpus: int = 1,
                    real_path: bool | None = None, out=None, status_out=None):
    """L[1,1] of ql(A - λI) for the pure Toeplitz A (bandwidths (l,1)) with BandedMatrices data column
    `data_column` = [a_{+1}, a_0, a_{-1}, …, a_{-l}], over all shifts `lambdas`
    (reference: src/banded/infqltoeplitz.jl:56-65,191 and `.L[1,1]`, src/infql.jl:96).

    Returns (L11, status[, tail]); entries with status != 0 are NaN (the reference throws DomainError there).
    """
    col = np.asarray(data_column)
    lam = np.asarray(lambdas)
    if real_path is None:
        real_path = not (np.iscomplexobj(col) or np.iscomplexobj(lam))
    dt = np.float64 if real_path else np.complex128
    col = np.ascontiguousarray(col, dtype=dt)
    shape = lam.shape
    lam = np.ascontiguousarray(lam, dtype=dt).reshape(-1)
    n = lam.size
    l = col.size - 2
    m = l + 2
    L11 = out if out is not None else np.empty(n, dtype=dt)
    status = status_out if status_out is not None else np.empty(n, dtype=np.int32)
    tail = np.empty((n, 2 * m + 2), dtype=dt) if want_tail else None
    fn = lib().ila_ql_toeplitz_l11_f64 if real_path else lib().ila_ql_toeplitz_l11_c64
    check(fn(l, 1, ptr(col), ptr(lam), n, branch_rank, ptr(L11), ptr(tail), ptr(status), ngpus))
    res = (L11.reshape(shape), status.reshape(shape))
    if want_tail:
        res = res + (tail.reshape(shape + (2 * m + 2,)),)
    return res


def ql_toeplitz_l11_real(data_column, lambdas, branch_rank: int = 1, ngpus: int = 1, want_status: bool = True, out=None,
                         status_out=None):
    """Compact complex-eltype form (`ila_ql_toeplitz_l11re_c64`): L[1,1] of ql(A - λI) is REAL for ComplexF64 input, so
    8 B (+ 1 B status) per shift come back instead of 16 + 4.  Returns (L11 float64 — NaN where status != 0 with the
    status in the NaN payload —, status uint8 or None)."""
    col = np.ascontiguousarray(np.asarray(data_column), dtype=np.complex128)
    lam = np.asarray(lambdas)
    shape = lam.shape
    lam = np.ascontiguousarray(lam, dtype=np.complex128).reshape(-1)
    n = lam.size
    L11 = out if out is not None else np.empty(n, dtype=np.float64)
    status = status_out if status_out is not None else (np.empty(n, dtype=np.uint8) if want_status else None)
    check(lib().ila_ql_toeplitz_l11re_c64(col.size - 2, 1, ptr(col), ptr(lam), n, branch_rank, ptr(L11), ptr(status), ngpus))
    return L11.reshape(shape), (status.reshape(shape) if status is not None else None)


def status_from_nan_payload(L11re):
    """Status codes carried by the NaN payloads of `ql_toeplitz_l11_real`'s result (0 where the entry is a number)."""
    bits = np.ascontiguousarray(L11re, dtype=np.float64).view(np.uint64)
    return np.where(np.isnan(L11re), (bits & np.uint64(0xFF)).astype(np.int32), 0).astype(np.int32)


def ql_toeplitz_solve(data_column, lambdas, b, nout: int, branch_rank: int = 1, ngpus: int = 1,
                      real_path: bool | None = None):
    """First `nout` entries of x = ql(A - λI) \\ b over all shifts `lambdas`, for the pure Toeplitz A of
    `ql_toeplitz_l11` and a right-hand side with len(b) <= 16 leading entries (reference: `F \\ b` →
    ldiv!(F.L, lmul!(F.Q', b)), src/infql.jl:266-289, src/banded/hessenbergq.jl:125-135; test/test_infql.jl:143-145).
    Returns (x [shape lambdas.shape + (nout,)], status); rows with status != 0 are NaN."""
    col = np.asarray(data_column)
    lam = np.asarray(lambdas)
    bb = np.asarray(b)
    if real_path is None:
        real_path = not (np.iscomplexobj(col) or np.iscomplexobj(lam) or np.iscomplexobj(bb))
    dt = np.float64 if real_path else np.complex128
    col = np.ascontiguousarray(col, dtype=dt)
    bb = np.ascontiguousarray(bb, dtype=dt).reshape(-1)
    shape = lam.shape
    lam = np.ascontiguousarray(lam, dtype=dt).reshape(-1)
    n = lam.size
    x = np.empty((int(nout), n), dtype=dt)            # column-major n x nout
    status = np.empty(n, dtype=np.int32)
    fn = lib().ila_ql_toeplitz_solve_f64 if real_path else lib().ila_ql_toeplitz_solve_c64
    check(fn(col.size - 2, 1, ptr(col), ptr(lam), n, branch_rank, bb.size, ptr(bb), int(nout), ptr(x), ptr(status), ngpus))
    return np.ascontiguousarray(x.T).reshape(shape + (int(nout),)), status.reshape(shape)


def ql_toeplitz_applyq(data_column, lambdas, b, nout: int, adjoint: bool = False, branch_rank: int = 1, ngpus: int = 1,
                       real_path: bool | None = None):
    """y = Q b (adjoint=False) or Q'b (adjoint=True), first `nout` entries, Q the ∞ Hessenberg Q of ql(A - λI) for every λ
    (reference: LowerHessenbergQ / UpperHessenbergQ apply, src/banded/hessenbergq.jl:111-135).  Returns (y [λ-shape + (nout,)],
    nstop [the n at which the reference's early exit fires, 0 if beyond nout; None for the adjoint], status)."""
    col = np.asarray(data_column)
    lam = np.asarray(lambdas)
    bb = np.asarray(b)
    if real_path is None:
        real_path = not (np.iscomplexobj(col) or np.iscomplexobj(lam) or np.iscomplexobj(bb))
    dt = np.float64 if real_path else np.complex128
    col = np.ascontiguousarray(col, dtype=dt)
    bb = np.ascontiguousarray(bb, dtype=dt).reshape(-1)
    shape = lam.shape
    lam = np.ascontiguousarray(lam, dtype=dt).reshape(-1)
    n = lam.size
    y = np.empty((int(nout), n), dtype=dt)
    status = np.empty(n, dtype=np.int32)
    nstop = None if adjoint else np.empty(n, dtype=np.int64)
    fn = lib().ila_ql_toeplitz_applyq_f64 if real_path else lib().ila_ql_toeplitz_applyq_c64
    check(fn(col.size - 2, 1, ptr(col), ptr(lam), n, branch_rank, 1 if adjoint else 0, bb.size, ptr(bb), int(nout), ptr(y),
             ptr(nstop), ptr(status), ngpus))
    return (np.ascontiguousarray(y.T).reshape(shape + (int(nout),)), None if nstop is None else nstop.reshape(shape),
            status.reshape(shape))


def ql_toeplitz_absl11_grid(data_column, x, y, branch_rank: int = 1, ngpus: int = 1, out=None):
    """The reference example's λ-grid loop (examples/toeplitz.jl:14-33): z[k, j] = abs(ql(A - (x[j] + i y[k]) I).L[1,1]),
    -1.0 where the QL factorization does not exist.  Shifts are formed on device."""
    col = np.ascontiguousarray(np.asarray(data_column), dtype=np.complex128)
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    z = out if out is not None else np.empty((y.size, x.size), dtype=np.float64)
    check(lib().ila_ql_toeplitz_absl11_grid_c64(col.size - 2, 1, ptr(col), ptr(x), x.size, ptr(y), y.size, branch_rank,
                                                ptr(z), ngpus))
    return z


def ql_toeplitz_l11_dev(data_column, d_lambda, d_L11, d_status, branch_rank: int = 1, d_tail=None, stream=None,
                        real_path: bool = False):
    """Device-resident form: `d_*` are torch CUDA tensors (complex128 / float64 / int32) on the current device."""
    dt = np.float64 if real_path else np.complex128
    col = np.ascontiguousarray(np.asarray(data_column), dtype=dt)
    l = col.size - 2
    n = d_lambda.numel()
    fn = lib().ila_ql_toeplitz_l11_f64_dev if real_path else lib().ila_ql_toeplitz_l11_c64_dev
    s = C.c_void_p(stream) if stream else None
    check(fn(l, 1, ptr(col), ptr(d_lambda), n, branch_rank, ptr(d_L11), ptr(d_tail), ptr(d_status), s))


def ql_perttoeplitz_l11(head, data_column, lambdas, branch_rank: int = 1, want_head_factors: bool = False,
                        ngpus: int = 1, real_path: bool | None = None):
    """L[1,1] of ql(A - λI) for a perturbed Toeplitz A (bandwidths (l,1)): `head` is the dense p x p top-left corner
    of A, `data_column` the Toeplitz tail column [a_{+1}, a_0, ..., a_{-l}] (reference src/infql.jl:60-93).
    Returns (L11, status[, head_factors (..., p, p)])."""
    col = np.asarray(data_column)
    head = np.asarray(head)
    lam = np.asarray(lambdas)
    if real_path is None:
        real_path = not (np.iscomplexobj(col) or np.iscomplexobj(lam) or np.iscomplexobj(head))
    dt = np.float64 if real_path else np.complex128
    col = np.ascontiguousarray(col, dtype=dt)
    p = head.shape[0]
    hflat = np.ascontiguousarray(np.asarray(head, dtype=dt).reshape(-1, order="F"))
    shape = lam.shape
    lam = np.ascontiguousarray(lam, dtype=dt).reshape(-1)
    n = lam.size
    l = col.size - 2
    L11 = np.empty(n, dtype=dt)
    status = np.empty(n, dtype=np.int32)
    hf = np.empty((n, p * p), dtype=dt) if want_head_factors else None
    fn = lib().ila_ql_perttoeplitz_l11_f64 if real_path else lib().ila_ql_perttoeplitz_l11_c64
    check(fn(l, 1, p, ptr(hflat), ptr(col), ptr(lam), n, branch_rank, ptr(L11), ptr(hf), ptr(status), ngpus))
    res = (L11.reshape(shape), status.reshape(shape))
    if want_head_factors:
        res = res + (hf.reshape(shape + (p, p)).swapaxes(-1, -2),)     # column-major -> [i, j]
    return res


def ql_perttoeplitz_solve(head, data_column, lambdas, b, nout: int, branch_rank: int = 1, ngpus: int = 1,
                          real_path: bool | None = None):
    """First `nout` entries of x = ql(A - λI) \\ b for the perturbed Toeplitz A of `ql_perttoeplitz_l11`
    (reference: ql_hessenberg! assembly, src/infql.jl:70-93, and F \\ b, src/infql.jl:266-289; test/test_infql.jl:147-160).
    Returns (x [shape lambdas.shape + (nout,)], status)."""
    col = np.asarray(data_column)
    head = np.asarray(head)
    lam = np.asarray(lambdas)
    bb = np.asarray(b)
    if real_path is None:
        real_path = not (np.iscomplexobj(col) or np.iscomplexobj(lam) or np.iscomplexobj(head) or np.iscomplexobj(bb))
    dt = np.float64 if real_path else np.complex128
    col = np.ascontiguousarray(col, dtype=dt)
    bb = np.ascontiguousarray(bb, dtype=dt).reshape(-1)
    p = head.shape[0]
    hflat = np.ascontiguousarray(np.asarray(head, dtype=dt).reshape(-1, order="F"))
    shape = lam.shape
    lam = np.ascontiguousarray(lam, dtype=dt).reshape(-1)
    n = lam.size
    x = np.empty((int(nout), n), dtype=dt)            # column-major n x nout
    status = np.empty(n, dtype=np.int32)
    fn = lib().ila_ql_perttoeplitz_solve_f64 if real_path else lib().ila_ql_perttoeplitz_solve_c64
    check(fn(col.size - 2, 1, p, ptr(hflat), ptr(col), ptr(lam), n, branch_rank, bb.size, ptr(bb), int(nout), ptr(x),
             ptr(status), ngpus))
    return np.ascontiguousarray(x.T).reshape(shape + (int(nout),)), status.reshape(shape)


# ------------------------------------------------------------------------------------------------
# K6: block-tridiagonal Toeplitz-tail QL
# ------------------------------------------------------------------------------------------------

def _blocks(lst, n, dt=np.float64):
    if lst is None or len(lst) == 0:
        return None, 0
    arr = np.stack([np.asfortranarray(np.asarray(m, dtype=dt)).reshape(-1, order="F") for m in lst])
    assert arr.shape[1] == n * n
    return np.ascontiguousarray(arr), len(lst)


def ql_blocktri_l11(dl_head, d_head, du_head, c, a, b, energies, atol=1e-10, maxit=1_000_000, ngpus=1):
    """L[1,1] of ql(A - x*I) over `energies` for the block-tridiagonal perturbed Toeplitz operator
    A = BlockTridiagonal(Vcat(dl_head, Fill(c,∞)), Vcat(d_head, Fill(a,∞)), Vcat(du_head, Fill(b,∞)))
    (reference src/infql.jl:166-209; examples/periodicschrodinger.jl:13-27).  Real blocks and energies run the real
    kernel (QL loop down to k = 2, bit-exact against the oracle); any complex block or energy (`A - 5im*I`,
    test/test_periodic.jl:20-26) runs the complex one (loop down to k = 1, complex L[1,1]).
    Returns (L11, iters, status)."""
    mats = [c, a, b] + list(dl_head or []) + list(d_head or []) + list(du_head or [])
    cplx = np.iscomplexobj(np.asarray(energies)) or any(np.iscomplexobj(np.asarray(m)) for m in mats)
    dt = np.complex128 if cplx else np.float64
    c = np.asarray(c, dtype=dt)
    n = c.shape[0]
    flat = lambda m: np.ascontiguousarray(np.asarray(m, dtype=dt).reshape(-1, order="F"))
    cc, aa, bb = flat(c), flat(a), flat(b)
    dlh, ndl = _blocks(dl_head, n, dt)
    dh, nd = _blocks(d_head, n, dt)
    duh, ndu = _blocks(du_head, n, dt)
    e = np.ascontiguousarray(np.asarray(energies, dtype=dt).reshape(-1))
    L11 = np.empty(e.size, dtype=dt)
    iters = np.empty(e.size, dtype=np.int32)
    status = np.empty(e.size, dtype=np.int32)
    fn = lib().ila_ql_blocktri_l11_c64 if cplx else lib().ila_ql_blocktri_l11_f64
    check(fn(n, ptr(cc), ptr(aa), ptr(bb), ndl, ptr(dlh), nd, ptr(dh), ndu, ptr(duh), ptr(e),
             e.size, float(atol), int(maxit), ptr(L11), ptr(iters), ptr(status), ngpus))
    shape = np.asarray(energies).shape
    return L11.reshape(shape), iters.reshape(shape), status.reshape(shape)


def ql_blocktri_factors(dl_head, d_head, du_head, c, a, b, energies, atol=1e-10, maxit=1_000_000, ngpus=1, rhs=None,
                        nout=None):
    """The whole factorization of ql(A - x*I) per energy — what the reference's QL object holds (src/infql.jl:199-204) — and,
    with `rhs`, (Q'rhs)[1:nout] (lmul!(Q', b), src/infql.jl:213-243).  Returns a dict: L11, tail (…, 2n, 3n) = P,
    tau_tail (…, n), head (…, Nn, Nn) factored head, head_tau (…, Nn), qtb (…, nout) or None, iters, status, n, N."""
    mats = [c, a, b] + list(dl_head or []) + list(d_head or []) + list(du_head or [])
    cplx = np.iscomplexobj(np.asarray(energies)) or any(np.iscomplexobj(np.asarray(m)) for m in mats) or \
        (rhs is not None and np.iscomplexobj(np.asarray(rhs)))
    dt = np.complex128 if cplx else np.float64
    c = np.asarray(c, dtype=dt)
    n = c.shape[0]
    flat = lambda m: np.ascontiguousarray(np.asarray(m, dtype=dt).reshape(-1, order="F"))
    cc, aa, bb = flat(c), flat(a), flat(b)
    dlh, ndl = _blocks(dl_head, n, dt)
    dh, nd = _blocks(d_head, n, dt)
    duh, ndu = _blocks(du_head, n, dt)
    N = max(ndu + 1, nd, ndl)
    hn = N * n
    shape = np.asarray(energies).shape
    e = np.ascontiguousarray(np.asarray(energies, dtype=dt).reshape(-1))
    m = e.size
    L11 = np.empty(m, dtype=dt)
    tail = np.empty((m, 3 * n, 2 * n), dtype=dt)            # column-major 2n x 3n per energy
    tau_tail = np.empty((m, n), dtype=dt)
    head = np.empty((m, hn, hn), dtype=dt)
    head_tau = np.empty((m, hn), dtype=dt)
    iters = np.empty(m, dtype=np.int32)
    status = np.empty(m, dtype=np.int32)
    qtb = rb = None
    nb = 0
    if rhs is not None:
        rb = np.ascontiguousarray(np.asarray(rhs, dtype=dt).reshape(-1))
        nb = rb.size
        nout = int(nout if nout is not None else nb + 2 * n - 1)
        qtb = np.empty((m, nout), dtype=dt)
    fn = lib().ila_ql_blocktri_factors_c64 if cplx else lib().ila_ql_blocktri_factors_f64
    check(fn(n, ptr(cc), ptr(aa), ptr(bb), ndl, ptr(dlh), nd, ptr(dh), ndu, ptr(duh), ptr(e), m, float(atol), int(maxit),
             ptr(L11), ptr(tail), ptr(tau_tail), ptr(head), ptr(head_tau), nb, ptr(rb), int(nout or 0), ptr(qtb), ptr(iters),
             ptr(status), ngpus))
    return dict(L11=L11.reshape(shape), tail=np.swapaxes(tail, 1, 2).reshape(shape + (2 * n, 3 * n)),
                tau_tail=tau_tail.reshape(shape + (n,)), head=np.swapaxes(head, 1, 2).reshape(shape + (hn, hn)),
                head_tau=head_tau.reshape(shape + (hn,)), qtb=None if qtb is None else qtb.reshape(shape + (nout,)),
                iters=iters.reshape(shape), status=status.reshape(shape), n=n, N=N)


def periodic_schrodinger_blocks():
    """Blocks of examples/periodicschrodinger.jl:8-17: (dl_head, d_head, du_head, c, a, b)."""
    c = np.array([[0.0, 1.0], [0.0, 0.0]])
    a = np.array([[-1.0, 1.0], [1.0, 1.0]])
    b = np.array([[0.0, 0.0], [1.0, 0.0]])
    a0 = a.copy()
    a0[0, 0] = 2.0
    return [c], [a0], [b], c, a, b


# ------------------------------------------------------------------------------------------------
# K1+K2: adaptive banded QR solve
# ------------------------------------------------------------------------------------------------

def _op_arrays(nd, batch, p, q, r, shift, head):
    p = np.ascontiguousarray(np.broadcast_to(np.asarray(p, dtype=np.float64), (batch, nd)))
    q = np.ascontiguousarray(np.broadcast_to(np.asarray(q, dtype=np.float64), (batch, nd)))
    r = None if r is None else np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64), (batch, nd)))
    shift = None if shift is None else np.ascontiguousarray(
        np.broadcast_to(np.asarray(shift, dtype=np.float64), (batch,)))
    hp = 0
    if head is not None:
        head = np.ascontiguousarray(np.asarray(head, dtype=np.float64))
        if head.shape[0] != nd:
            raise ValueError("head must have l+u+1 rows")
        hp = head.shape[1]
    return p, q, r, shift, hp, head


def qr_banded_solve(l, u, p, q, r=None, shift=None, head=None, b=(1.0,), tol=FLOATMIN, colgrowth=1000,
                    ncap=200_000, ncap_per=None, batch=None, want_factors=False, ngpus=1, xcap=None, x_out=None):
    """Batched `A_i \\ b_i` by adaptive banded QR (reference src/infqr.jl:236-274,350-355,370-374).

    Operator description: see include/ila_b200.h (generators per diagonal, BandedMatrices data-row order).
    Returns dict(x=list of per-instance solution prefixes (views), xflat, xoff, ncols, nconv, status
    [, factors, tau, qtb as per-instance lists]).
    """
    nd = l + u + 1
    if batch is None:
        batch = int(np.asarray(p).reshape(-1, nd).shape[0])
    p, q, r, shift, hp, head = _op_arrays(nd, batch, p, q, r, shift, head)
    b = np.asarray(b, dtype=np.float64)
    if b.ndim == 1:
        b = np.broadcast_to(b, (batch, b.shape[0]))
    b = np.ascontiguousarray(b)
    nb = b.shape[1]
    caps = None
    if ncap_per is not None:
        caps = np.ascontiguousarray(np.asarray(ncap_per, dtype=np.int64))
    if xcap is None:
        xcap = int(caps.sum()) if caps is not None else int(batch) * int(ncap)
    x = x_out if x_out is not None else np.empty(xcap, dtype=np.float64)
    xcap = x.size
    xoff = np.zeros(batch + 1, dtype=np.int64)
    ncols = np.zeros(batch, dtype=np.int64)
    nconv = np.zeros(batch, dtype=np.int64)
    status = np.zeros(batch, dtype=np.int32)
    LD = 2 * l + u + 1
    fac = np.empty(xcap * LD) if want_factors else None
    tau = np.empty(xcap) if want_factors else None
    qtb = np.empty(xcap) if want_factors else None
    check(lib().ila_qr_banded_solve_f64(l, u, batch, ptr(p), ptr(q), ptr(r), ptr(shift), hp, ptr(head), nb, ptr(b),
                                        float(tol), int(colgrowth), int(ncap), ptr(caps), ptr(x), xcap, ptr(xoff),
                                        ptr(ncols), ptr(nconv), ptr(fac), ptr(tau), ptr(qtb), ptr(status), ngpus))
    out = dict(xflat=x[:xoff[batch]], xoff=xoff, ncols=ncols, nconv=nconv, status=status,
               x=[x[xoff[i]:xoff[i + 1]] for i in range(batch)])
    if want_factors:
        out["factors"] = [fac[xoff[i] * LD:xoff[i + 1] * LD].reshape(-1, LD) for i in range(batch)]
        out["tau"] = [tau[xoff[i]:xoff[i + 1]] for i in range(batch)]
        out["qtb"] = [qtb[xoff[i]:xoff[i + 1]] for i in range(batch)]
    return out


def qr_banded_solve_c64(l, u, p, q, r=None, shift=None, head=None, b=(1.0,), tol=FLOATMIN, colgrowth=1000,
                        ncap=200_000, ncap_per=None, batch=None, ngpus=1, xcap=None):
    """Complex-eltype `A_i \\ b_i` by adaptive banded QR (same reference path as `qr_banded_solve`; p, q, shift, head, b
    complex, r real): e.g. `(J - λ_i I) \\ b` for complex shifts.  Returns dict(x, xflat, xoff, ncols, nconv, status)."""
    nd = l + u + 1
    if batch is None:
        batch = int(np.asarray(p).reshape(-1, nd).shape[0])
    cz = np.complex128
    p = np.ascontiguousarray(np.broadcast_to(np.asarray(p, dtype=cz), (batch, nd)))
    q = np.ascontiguousarray(np.broadcast_to(np.asarray(q, dtype=cz), (batch, nd)))
    r = None if r is None else np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64), (batch, nd)))
    shift = None if shift is None else np.ascontiguousarray(np.broadcast_to(np.asarray(shift, dtype=cz), (batch,)))
    hp = 0
    if head is not None:
        head = np.ascontiguousarray(np.asarray(head, dtype=cz))
        assert head.shape[0] == nd
        hp = head.shape[1]
    b = np.asarray(b, dtype=cz)
    if b.ndim == 1:
        b = np.broadcast_to(b, (batch, b.shape[0]))
    b = np.ascontiguousarray(b)
    nb = b.shape[1]
    caps = None if ncap_per is None else np.ascontiguousarray(np.asarray(ncap_per, dtype=np.int64))
    if xcap is None:
        xcap = int(caps.sum()) if caps is not None else int(batch) * int(ncap)
    x = np.empty(xcap, dtype=cz)
    xoff = np.zeros(batch + 1, dtype=np.int64)
    ncols = np.zeros(batch, dtype=np.int64)
    nconv = np.zeros(batch, dtype=np.int64)
    status = np.zeros(batch, dtype=np.int32)
    check(lib().ila_qr_banded_solve_c64(l, u, batch, ptr(p), ptr(q), ptr(r), ptr(shift), hp, ptr(head), nb, ptr(b),
                                        float(tol), int(colgrowth), int(ncap), ptr(caps), ptr(x), xcap, ptr(xoff),
                                        ptr(ncols), ptr(nconv), ptr(status), ngpus))
    return dict(xflat=x[:xoff[batch]], xoff=xoff, ncols=ncols, nconv=nconv, status=status,
                x=[x[xoff[i]:xoff[i + 1]] for i in range(batch)])


def chol_banded_solve(u, p, q, r=None, shift=None, head=None, b=(1.0,), tol=FLOATMIN, colgrowth=1000, ncap=20_000,
                      ncap_per=None, batch=None, want_factors=False, ngpus=1, xcap=None):
    """Batched `cholesky(S_i) \\ b_i` by adaptive banded Cholesky (reference src/infcholesky.jl:42-59,94-140).

    The symmetric operators are described by their UPPER bands (data rows 0..u <-> diagonals +u..0), generators as for
    `qr_banded_solve`.  Returns dict(x, xflat, xoff, ncols, nconv, status[, factors (ncols x (u+1)), y]).
    """
    nd = u + 1
    if batch is None:
        batch = int(np.asarray(p).reshape(-1, nd).shape[0])
    p, q, r, shift, hp, head = _op_arrays(nd, batch, p, q, r, shift, head)
    b = np.asarray(b, dtype=np.float64)
    if b.ndim == 1:
        b = np.broadcast_to(b, (batch, b.shape[0]))
    b = np.ascontiguousarray(b)
    nb = b.shape[1]
    caps = None if ncap_per is None else np.ascontiguousarray(np.asarray(ncap_per, dtype=np.int64))
    if xcap is None:
        xcap = int(caps.sum()) if caps is not None else int(batch) * int(ncap)
    x = np.empty(xcap, dtype=np.float64)
    xoff = np.zeros(batch + 1, dtype=np.int64)
    ncols = np.zeros(batch, dtype=np.int64)
    nconv = np.zeros(batch, dtype=np.int64)
    status = np.zeros(batch, dtype=np.int32)
    fac = np.empty(xcap * nd) if want_factors else None
    y = np.empty(xcap) if want_factors else None
    check(lib().ila_chol_banded_solve_f64(u, batch, ptr(p), ptr(q), ptr(r), ptr(shift), hp, ptr(head), nb, ptr(b),
                                          float(tol), int(colgrowth), int(ncap), ptr(caps), ptr(x), xcap, ptr(xoff),
                                          ptr(ncols), ptr(nconv), ptr(fac), ptr(y), ptr(status), ngpus))
    out = dict(xflat=x[:xoff[batch]], xoff=xoff, ncols=ncols, nconv=nconv, status=status,
               x=[x[xoff[i]:xoff[i + 1]] for i in range(batch)])
    if want_factors:
        out["factors"] = [fac[xoff[i] * nd:xoff[i + 1] * nd].reshape(-1, nd) for i in range(batch)]
        out["y"] = [y[xoff[i]:xoff[i + 1]] for i in range(batch)]
    return out


def revchol_tridiag(ga, gb, n, shift=None, a_head=None, b_head=None, batch=None, ngpus=1):
    """Adaptive reverse Cholesky A = L'L of ∞ symmetric tridiagonal operators, batched
    (reference src/banded/infreversecholeskytridiagonal.jl:66-141; test/test_infreversecholesky.jl).
    ga, gb: (batch x) 5 generator coefficients (c0, p, q, r, s) of the diagonal / off-diagonal:
    value(k) = c0 + (p + q k)/(r + s k); `shift` is subtracted from the diagonal; optional shared explicit heads.
    Returns dict(la, lb (batch x n: diagonal and sub-diagonal of L), N, status)."""
    ga = np.asarray(ga, dtype=np.float64).reshape(-1, 5)
    gb = np.asarray(gb, dtype=np.float64).reshape(-1, 5)
    if batch is None:
        batch = max(ga.shape[0], gb.shape[0], 1 if shift is None else np.asarray(shift).size)
    ga = np.ascontiguousarray(np.broadcast_to(ga, (batch, 5)))
    gb = np.ascontiguousarray(np.broadcast_to(gb, (batch, 5)))
    shift = None if shift is None else np.ascontiguousarray(np.broadcast_to(np.asarray(shift, dtype=np.float64), (batch,)))
    hp = 0
    if a_head is not None:
        a_head = np.ascontiguousarray(np.asarray(a_head, dtype=np.float64))
        b_head = np.ascontiguousarray(np.asarray(b_head, dtype=np.float64))
        assert a_head.shape == b_head.shape
        hp = a_head.shape[0]
    la = np.empty((batch, int(n))); lb = np.empty((batch, int(n)))
    N = np.zeros(batch, dtype=np.int64); status = np.zeros(batch, dtype=np.int32)
    check(lib().ila_revchol_tridiag_f64(batch, ptr(ga), ptr(gb), ptr(shift), hp, ptr(a_head), ptr(b_head), int(n), ptr(la),
                                        ptr(lb), ptr(N), ptr(status), ngpus))
    return dict(la=la, lb=lb, N=N, status=status)


def ql_finite_section(l, u, g, j=50, shift=None, head=None, tol=np.finfo(np.float64).eps, maxN=10000, batch=None, ngpus=1):
    """Adaptive finite-section QL of ∞ banded operators without a Toeplitz tail, batched (reference src/infql.jl:364-462:
    `ql(A::InfBandedMatrix, tol)`; test/test_infql.jl:208-271).  g: (batch x) (l+u+1) x 5 generator coefficients
    (c0, p, q, r, s) per diagonal in data-row order (+u first): value(t) = c0 + (p + q t)/(r + s t), t = min(row, col)
    0-based; `shift` is subtracted from the main diagonal.  Complex g / shift / head select the complex eltype.
    Returns dict(Lband (batch x j x (l+u+1): Lband[k, c] = L[k, k-c]), V (batch x j x u), tau (batch x j), N, status);
    L[1,1] of operator i is Lband[i, 0, 0]."""
    nd = l + u + 1
    cplx = any(v is not None and np.iscomplexobj(np.asarray(v)) for v in (g, shift, head))
    dt = np.complex128 if cplx else np.float64
    g = np.asarray(g, dtype=dt).reshape(-1, nd, 5)
    if batch is None:
        batch = max(g.shape[0], 1 if shift is None else np.asarray(shift).size)
    g = np.ascontiguousarray(np.broadcast_to(g, (batch, nd, 5)))
    shift = None if shift is None else np.ascontiguousarray(np.broadcast_to(np.asarray(shift, dtype=dt), (batch,)))
    hp = 0
    if head is not None:
        head = np.ascontiguousarray(np.asarray(head, dtype=dt))
        assert head.shape[0] == nd
        hp = head.shape[1]
    j = int(j)
    Lband = np.empty((batch, j, nd), dtype=dt); V = np.empty((batch, j, u), dtype=dt); tau = np.empty((batch, j), dtype=dt)
    N = np.zeros(batch, dtype=np.int64); status = np.zeros(batch, dtype=np.int32)
    fn = lib().ila_ql_finite_section_c64 if cplx else lib().ila_ql_finite_section_f64
    check(fn(l, u, batch, ptr(g), ptr(shift), hp, ptr(head), j, float(tol), int(maxN), ptr(Lband), ptr(V), ptr(tau), ptr(N),
             ptr(status), ngpus))
    return dict(Lband=Lband, V=V, tau=tau, N=N, status=status)


def tridiagonal_conjugation(U, X, n, want_ux=False, ngpus=1):
    """Y = Tridiagonal(U*X/U), batched (reference src/banded/tridiagonalconjugation.jl: TridiagonalConjugationData +
    resizedata!(data, n)).  U: (batch x) nbands x m bands, U[d, k] = U[k+1, k+1+d] (nbands 2 or 3); X: 3 x m (shared) or
    batch x 3 x m with rows dl, d, du; m >= n+1.  Returns Y (batch x 3 x n: dl, d, du) [and UX = Tridiagonal(U*X)]."""
    U = np.ascontiguousarray(np.asarray(U, dtype=np.float64))
    if U.ndim == 2:
        U = U[None]
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    batch, nbands, uld = U.shape
    shared = X.ndim == 2
    assert shared or X.shape[0] == batch
    n = int(n)
    Y = np.empty((batch, 3, n)); UX = np.empty((batch, 3, n)) if want_ux else None
    check(lib().ila_triconj_bands_f64(nbands, batch, n, ptr(U), uld, ptr(X), X.shape[-1], 1 if shared else 0, ptr(Y), ptr(UX), ngpus))
    return (Y, UX) if want_ux else Y


def chol_tridiagonal_conjugation(u, p, q, X, n, r=None, shift=None, head=None, batch=None, ngpus=1):
    """TridiagonalConjugation(cholesky(S_i).U, X) for a batch of symmetric positive definite banded operators (generators as
    in `chol_banded_solve`), the factor staying on the device between the Cholesky sweep (K5) and the conjugation (K10).
    Returns dict(Y (batch x 3 x n), status)."""
    nd = u + 1
    if batch is None:
        batch = max(np.atleast_2d(np.asarray(p)).shape[0], 1 if shift is None else np.asarray(shift).size)
    p, q, r, shift, hp, head = _op_arrays(nd, batch, p, q, r, shift, head)
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    shared = X.ndim == 2
    assert shared or X.shape[0] == batch
    n = int(n)
    Y = np.empty((batch, 3, n)); status = np.zeros(batch, dtype=np.int32)
    check(lib().ila_chol_triconj_f64(u, batch, ptr(p), ptr(q), ptr(r), ptr(shift), hp, ptr(head), n, ptr(X), X.shape[-1],
                                     1 if shared else 0, ptr(Y), ptr(status), ngpus))
    return dict(Y=Y, status=status)


def bidiagonal_conjugation(U, Cm, n, uplo="U", ngpus=1):
    """Bidiagonal part of inv(U)*X*V from the tridiagonal parts of U and C = X*V (reference
    src/banded/bidiagonalconjugation.jl:19-60).  U, Cm: (batch x) 3 x m (sub, main, super), m >= n+1.  Returns dv, ev."""
    U = np.ascontiguousarray(np.asarray(U, dtype=np.float64)); Cm = np.ascontiguousarray(np.asarray(Cm, dtype=np.float64))
    if U.ndim == 2:
        U = U[None]; Cm = Cm[None]
    batch, _, ld = U.shape
    assert Cm.shape == U.shape
    n = int(n)
    dv = np.empty((batch, n)); ev = np.empty((batch, n))
    check(lib().ila_biconj_bands_f64(1 if uplo in ("U", ":U") else 0, batch, n, ptr(U), ptr(Cm), ld, ptr(dv), ptr(ev), ngpus))
    return dv, ev


def qr_blockbanded_solve(T1, T2, alpha, beta, lam, b=(1.0,), tol=1e-30, colgrowth=300, maxblocks=64, want_qtb=False, ngpus=1):
    """Adaptive block-banded QR solve, batched over (alpha, beta, lam) (reference src/infqr.jl:23-28, 68-86, 303-337, 358-366;
    test/test_infqr.jl:270-313):  (alpha KronTrav(T1, I) + beta KronTrav(I, T2) - lam I) \\ [b; 0; ...]  with T1, T2 tridiagonal
    (3 x m bands dl, d, du, m >= maxblocks+1) and block sizes 1, 2, 3, ...  Returns dict(x (batch x maxblocks(maxblocks+1)/2,
    zero padded), ncols, nblocks, status[, qtb])."""
    alpha = np.atleast_1d(np.asarray(alpha, dtype=np.float64)); beta = np.atleast_1d(np.asarray(beta, dtype=np.float64))
    lam = np.atleast_1d(np.asarray(lam, dtype=np.float64))
    batch = max(alpha.size, beta.size, lam.size)
    alpha = np.ascontiguousarray(np.broadcast_to(alpha, (batch,))); beta = np.ascontiguousarray(np.broadcast_to(beta, (batch,)))
    lam = np.ascontiguousarray(np.broadcast_to(lam, (batch,)))
    T1 = np.ascontiguousarray(np.asarray(T1, dtype=np.float64)); T2 = np.ascontiguousarray(np.asarray(T2, dtype=np.float64))
    assert T1.shape == T2.shape and T1.shape[0] == 3
    b = np.asarray(b, dtype=np.float64)
    if b.ndim == 1:
        b = np.broadcast_to(b, (batch, b.shape[0]))
    b = np.ascontiguousarray(b)
    xld = maxblocks * (maxblocks + 1) // 2
    x = np.empty((batch, xld)); qtb = np.empty((batch, xld)) if want_qtb else None
    ncols = np.zeros(batch, dtype=np.int64); nblocks = np.zeros(batch, dtype=np.int32); status = np.zeros(batch, dtype=np.int32)
    check(lib().ila_qr_blockbanded_solve_f64(batch, ptr(alpha), ptr(beta), ptr(lam), ptr(T1), ptr(T2), T1.shape[1], b.shape[1], ptr(b),
                                             float(tol), int(colgrowth), int(maxblocks), ptr(x), xld, ptr(ncols), ptr(nblocks),
                                             ptr(qtb), ptr(status), ngpus))
    out = dict(x=x, ncols=ncols, nblocks=nblocks, status=status)
    if want_qtb:
        out["qtb"] = qtb
    return out


def qr_almostbanded_solve(lD, uD, dg, bg, b, shift=None, tol=FLOATMIN, colgrowth=1000, ncap=6000, batch=None, ngpus=1):
    """Adaptive QR solve of ∞ almost-banded operators Vcat(B, D - shift I), batched (reference src/infqr.jl:15-21, 50-66, 236-274,
    350-355; test/test_infqr.jl:193-267).  dg: (batch x) (lD+uD+1) x 3 polynomial generators of D's diagonals (+uD first, value(t) =
    g0 + g1 t + g2 t^2 in the row index of D); bg: (batch x) r x 4 boundary rows s^j (a + b j + c j^2).
    Returns dict(x (batch x ncap, zero padded), ncols, nconv, status)."""
    nd = lD + uD + 1
    dg = np.asarray(dg, dtype=np.float64).reshape(-1, nd, 3)
    bg = np.asarray(bg, dtype=np.float64)
    bg = bg.reshape((-1,) + bg.shape[-2:]) if bg.ndim >= 2 else bg.reshape(1, 1, 4)
    r = bg.shape[1]
    b = np.asarray(b, dtype=np.float64)
    if batch is None:
        batch = max(dg.shape[0], bg.shape[0], 1 if shift is None else np.asarray(shift).size, b.shape[0] if b.ndim == 2 else 1)
    dg = np.ascontiguousarray(np.broadcast_to(dg, (batch, nd, 3)))
    bg = np.ascontiguousarray(np.broadcast_to(bg, (batch, r, 4)))
    if b.ndim == 1:
        b = np.broadcast_to(b, (batch, b.shape[0]))
    b = np.ascontiguousarray(b)
    shift = None if shift is None else np.ascontiguousarray(np.broadcast_to(np.asarray(shift, dtype=np.float64), (batch,)))
    ncap = int(ncap)
    x = np.empty((batch, ncap))
    ncols = np.zeros(batch, dtype=np.int64); nconv = np.zeros(batch, dtype=np.int64); status = np.zeros(batch, dtype=np.int32)
    check(lib().ila_qr_almostbanded_solve_f64(lD, uD, r, batch, ptr(dg), ptr(bg), ptr(shift), b.shape[1], ptr(b), float(tol),
                                              int(colgrowth), ncap, ptr(x), ptr(ncols), ptr(nconv), ptr(status), ngpus))
    return dict(x=x, ncols=ncols, nconv=nconv, status=status)


def pentadiagonal_operator(sigma, eps=1e-4):
    """BASELINE config C4: S_σ = pentadiag(diag 6+σ+ε·k, ±1: -4, ±2: 1), upper bands in data-row order (+2, +1, 0)."""
    sigma = np.atleast_1d(np.asarray(sigma, dtype=np.float64))
    batch = sigma.shape[0]
    p = np.tile(np.array([1.0, -4.0, 6.0]), (batch, 1))
    p[:, 2] += sigma
    q = np.zeros((batch, 3))
    q[:, 2] = eps
    return p, q


class QrBandedPlan:
    """Device-resident two-phase adaptive QR solve (torch tensors as plain device memory).

    All inputs live in HBM before `run()`; `run()` enqueues forward sweep, offset scan and back-substitution on
    the current torch stream without any host synchronisation — the timed region of bench.py's `value`.
    """

    def __init__(self, l, u, p, q, r=None, shift=None, head=None, b=(1.0,), tol=FLOATMIN, colgrowth=1000,
                 ncap=200_000, ncap_per=None, device=None, xcap=None, want_reflectors=False, kind="qr"):
        import torch
        self.torch = torch
        self.kind = kind            # "qr" (K1+K2) or "chol" (K5+K2: l must be 0, generators of the upper bands)
        if kind == "chol" and l != 0:
            raise ValueError("Cholesky plans take l = 0 and the upper-band generators")
        nd = l + u + 1
        batch = int(np.asarray(p).reshape(-1, nd).shape[0])
        p, q, r, shift, hp, head = _op_arrays(nd, batch, p, q, r, shift, head)
        b = np.asarray(b, dtype=np.float64)
        if b.ndim == 1:
            b = np.broadcast_to(b, (batch, b.shape[0]))
        b = np.ascontiguousarray(b)
        self.l, self.u, self.batch, self.hp, self.nb = l, u, batch, hp, b.shape[1]
        self.tol, self.colgrowth = float(tol), int(colgrowth)
        self.refl = 1 if want_reflectors else 0
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        caps = None if ncap_per is None else np.ascontiguousarray(np.asarray(ncap_per, dtype=np.int64))
        gbase = np.zeros((batch + 31) // 32 + 1, dtype=np.int64)
        wbytes = C.c_int64(0)
        xbytes = C.c_int64(0)
        check(lib().ila_qr_banded_work_layout(l, u, batch, int(ncap), ptr(caps), ptr(gbase), C.byref(wbytes),
                                              C.byref(xbytes)))
        t = lambda a: None if a is None else torch.from_numpy(np.array(a, copy=True)).to(dev)
        self.d_p, self.d_q, self.d_r, self.d_shift, self.d_head, self.d_b = t(p), t(q), t(r), t(shift), t(head), t(b)
        self.ncap = int(ncap)
        self.d_caps = t(caps)
        self.d_gbase = t(gbase)
        self.d_work = torch.empty(wbytes.value // 8, dtype=torch.float64, device=dev)
        self.d_xi = torch.empty(xbytes.value // 8, dtype=torch.float64, device=dev)
        self.d_ncols = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.d_nconv = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.d_nswept = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.d_xoff = torch.zeros(batch + 1, dtype=torch.int64, device=dev)
        self.d_status = torch.zeros(batch, dtype=torch.int32, device=dev)
        if xcap is None:
            xcap = int(caps.sum()) if caps is not None else batch * int(ncap)
        self.xcap = int(xcap)
        self.d_x = torch.empty(self.xcap, dtype=torch.float64, device=dev)
        # resumable state (the reference's AdaptiveQRData / AdaptiveCholeskyFactors: ncols + the trailing updated columns)
        sd = lib().ila_chol_banded_state_doubles(u) if kind == "chol" else lib().ila_qr_banded_state_doubles(l, u)
        self.state_doubles = int(sd)
        self.d_state = torch.zeros(batch * self.state_doubles, dtype=torch.float64, device=dev)

    def _stream(self):
        lib().ila_set_device(self.device.index or 0)
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def run(self):
        self.run_factor()
        self.run_backsolve()
        return 4   # kernels launched: forward, scan, back-substitution, compaction

    def run_factor_resumable(self, limit: int, resume: bool):
        """Adaptive sweep with the column limit `limit` (<= the capacity the workspace was laid out for): an instance that
        reaches it keeps its finished columns and its state (status 3, ncols = columns done); `resume=True` continues every
        such instance from there — `partialqr!(F, n)` / `partialcholesky!(F, n)` called again with a larger n
        (src/infqr.jl:32-48, src/infcholesky.jl:42-59).  Two calls give the same bits as one."""
        s = self._stream()
        if self.d_caps is not None:
            raise ValueError("resumable sweeps use the uniform column limit")
        if int(limit) > self.ncap:
            raise ValueError("limit exceeds the capacity the workspace was laid out for")
        if self.kind == "chol":
            check(lib().ila_chol_banded_factor_resume_f64_dev(
                self.u, self.batch, ptr(self.d_p), ptr(self.d_q), ptr(self.d_r), ptr(self.d_shift), self.hp,
                ptr(self.d_head), self.nb, ptr(self.d_b), self.tol, self.colgrowth, int(limit), None,
                ptr(self.d_gbase), ptr(self.d_work), ptr(self.d_ncols), ptr(self.d_nconv), ptr(self.d_nswept),
                ptr(self.d_xoff), ptr(self.d_status), 0, ptr(self.d_state), 1 if resume else 0, s))
        else:
            check(lib().ila_qr_banded_factor_resume_f64_dev(
                self.l, self.u, self.batch, ptr(self.d_p), ptr(self.d_q), ptr(self.d_r), ptr(self.d_shift), self.hp,
                ptr(self.d_head), self.nb, ptr(self.d_b), self.tol, self.colgrowth, int(limit), None,
                ptr(self.d_gbase), ptr(self.d_work), ptr(self.d_ncols), ptr(self.d_nconv), ptr(self.d_nswept),
                ptr(self.d_xoff), ptr(self.d_status), self.refl, ptr(self.d_state), 1 if resume else 0, s))
        return 2

    def partialqr(self, n: int, resume: bool = False):
        """`partialqr!(F, n)` (src/infqr.jl:32-48) for every operator of the plan: factorize through column n, reflectors and
        tau stored in the records; no right-hand side.  `resume=True` continues from the state an earlier call left."""
        if self.kind != "qr":
            raise ValueError("partialqr is the QR plan's")
        if int(n) > self.ncap:
            raise ValueError("n exceeds the capacity the workspace was laid out for")
        s = self._stream()
        self.refl = 1
        check(lib().ila_qr_banded_partialqr_f64_dev(
            self.l, self.u, self.batch, ptr(self.d_p), ptr(self.d_q), ptr(self.d_r), ptr(self.d_shift), self.hp,
            ptr(self.d_head), int(n), self.colgrowth, ptr(self.d_gbase), ptr(self.d_work), ptr(self.d_state),
            1 if resume else 0, ptr(self.d_ncols), ptr(self.d_nswept), ptr(self.d_xoff), ptr(self.d_status), s))
        return 2

    def export_factors(self):
        """Band storage of the factors, tau and Q'b of what has been swept so far (ragged at d_xoff): returns
        (factors [total x (2l+u+1)], tau [total], qtb [total], xoff [batch+1]) as numpy arrays."""
        torch = self.torch
        s = self._stream()
        xoff = self.d_xoff.cpu().numpy()
        total = int(xoff[-1])
        LD = 2 * self.l + self.u + 1
        d_f = torch.zeros(max(total, 1) * LD, dtype=torch.float64, device=self.device)
        d_t = torch.zeros(max(total, 1), dtype=torch.float64, device=self.device)
        d_q = torch.zeros(max(total, 1), dtype=torch.float64, device=self.device)
        check(lib().ila_qr_banded_export_f64_dev(self.l, self.u, self.refl, self.batch, ptr(self.d_gbase), ptr(self.d_work),
                                                 ptr(self.d_ncols), ptr(self.d_nswept), ptr(self.d_xoff), ptr(d_f),
                                                 ptr(d_t) if self.refl else None, ptr(d_q), s))
        torch.cuda.synchronize(self.device)
        return (d_f.cpu().numpy()[:total * LD].reshape(total, LD), d_t.cpu().numpy()[:total], d_q.cpu().numpy()[:total], xoff)

    def state_window(self):
        """The saved trailing window per instance: W[j][i] = (Q_n' A)[n+i, n+j], j = 0..l+u, i = 0..l (QR plans)."""
        st = self.d_state.cpu().numpy().reshape(self.batch, self.state_doubles)
        NC, NR = self.l + self.u + 1, self.l + 1
        return st[:, :NC * NR].reshape(self.batch, NC, NR)

    def run_factor(self):
        s = self._stream()
        if self.kind == "chol":
            check(lib().ila_chol_banded_factor_f64_dev(
                self.u, self.batch, ptr(self.d_p), ptr(self.d_q), ptr(self.d_r), ptr(self.d_shift), self.hp,
                ptr(self.d_head), self.nb, ptr(self.d_b), self.tol, self.colgrowth, self.ncap, ptr(self.d_caps),
                ptr(self.d_gbase), ptr(self.d_work), ptr(self.d_ncols), ptr(self.d_nconv), ptr(self.d_nswept),
                ptr(self.d_xoff), ptr(self.d_status), 0, s))
            return 2
        check(lib().ila_qr_banded_factor_f64_dev(
            self.l, self.u, self.batch, ptr(self.d_p), ptr(self.d_q), ptr(self.d_r), ptr(self.d_shift), self.hp,
            ptr(self.d_head), self.nb, ptr(self.d_b), self.tol, self.colgrowth, self.ncap, ptr(self.d_caps),
            ptr(self.d_gbase), ptr(self.d_work), ptr(self.d_ncols), ptr(self.d_nconv), ptr(self.d_nswept),
            ptr(self.d_xoff), ptr(self.d_status), self.refl, s))
        return 2

    def run_backsolve(self):
        s = self._stream()
        check(lib().ila_qr_banded_backsolve_f64_dev(
            self.l, self.u, self.batch, ptr(self.d_gbase), ptr(self.d_work), ptr(self.d_xi), ptr(self.d_ncols),
            ptr(self.d_nswept), ptr(self.d_xoff), ptr(self.d_status), ptr(self.d_x), self.xcap, self.refl, s))
        return 2


def bessel_operator(z):
    """Generator description of A = BandedMatrix(0 => -2*(0:∞)/z, 1 => Ones(∞), -1 => Ones(∞))
    (README.md:41, test/test_infqr.jl:95); data-row order: diagonal +1, 0, -1."""
    z = np.atleast_1d(np.asarray(z, dtype=np.float64))
    batch = z.shape[0]
    p = np.tile(np.array([1.0, 0.0, 1.0]), (batch, 1))
    q = np.tile(np.array([0.0, -2.0, 0.0]), (batch, 1))
    r = np.ones((batch, 3))
    r[:, 1] = z
    return p, q, r


def bessel_ncap(z, colgrowth=1000):
    """Column capacity that lets the Bessel solve converge: |(Q'b)_k| drops below floatmin at
    k ≈ z + 83 z^(1/3) (Debye asymptotics of J_k(z), k > z); the reference then factorizes up to two more
    COLGROWTH chunks."""
    z = np.atleast_1d(np.asarray(z, dtype=np.float64))
    return np.ceil(z + 90.0 * np.cbrt(z) + 2 * colgrowth + 100).astype(np.int64)


def fp64_peak_tflops():
    t = C.c_double(0.0)
    ms = C.c_double(0.0)
    check(lib().ila_fp64_peak_probe(C.byref(t), C.byref(ms)))
    return t.value
