"""ctypes front-end of the C oracle (oracle/ila_oracle.c) — TEST INFRASTRUCTURE ONLY.

Adaptive banded QR solve (reference src/infqr.jl:8-13,32-48,236-274,350-355) and adaptive banded
Cholesky solve (src/infcholesky.jl:42-59,94-140), restated in plain C with the reference's operation
order.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libila_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_sp = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile."""
    src = [os.path.join(_HERE, f) for f in ("ila_oracle.c", "ila_oracle_impl.h", "ila_oracle_c64.c", "ila_oracle_revchol.c", "ila_oracle_fsql.c", "ila_oracle_triconj.c", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def _prep_op(nd, batch, p, q, r, shift, head):
    p = np.ascontiguousarray(np.broadcast_to(np.asarray(p, dtype=np.float64), (batch, nd)))
    q = np.ascontiguousarray(np.broadcast_to(np.asarray(q, dtype=np.float64), (batch, nd)))
    r = None if r is None else np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64), (batch, nd)))
    shift = None if shift is None else np.ascontiguousarray(np.broadcast_to(np.asarray(shift, dtype=np.float64), (batch,)))
    hp = 0
    if head is not None:
        head = np.ascontiguousarray(np.asarray(head, dtype=np.float64))
        assert head.shape[0] == nd
        hp = head.shape[1]
    return p, q, r, shift, hp, head


def qr_banded_solve(l, u, p, q, r=None, shift=None, head=None, b=(1.0,), tol=np.finfo(np.float64).tiny,
                    colgrowth=1000, ncap=200_000, batch=None, want_factors=False, quad=False, nthreads=0):
    """Batched adaptive QR solve through the oracle; returns dict(x, ncols, nconv, status[, factors, tau, qtb])."""
    nd = l + u + 1
    if batch is None:
        batch = int(np.asarray(p).reshape(-1, nd).shape[0])
    p, q, r, shift, hp, head = _prep_op(nd, batch, p, q, r, shift, head)
    b = np.asarray(b, dtype=np.float64)
    if b.ndim == 1:
        b = np.broadcast_to(b, (batch, b.shape[0]))
    b = np.ascontiguousarray(b)
    nb = b.shape[1]
    ldx = ncap
    x = np.zeros((batch, ldx))
    ncols = np.zeros(batch, dtype=np.int64)
    nconv = np.zeros(batch, dtype=np.int64)
    status = np.zeros(batch, dtype=np.int32)
    out = dict(x=x, ncols=ncols, nconv=nconv, status=status)
    L = lib()
    if quad:
        L.ila_oracle_qr_banded_solve_f128(
            C.c_int(l), C.c_int(u), C.c_int64(batch), _ptr(p, _dp), _ptr(q, _dp), _ptr(r, _dp), _ptr(shift, _dp),
            C.c_int(hp), _ptr(head, _dp), C.c_int(nb), _ptr(b, _dp), C.c_double(tol), C.c_int(colgrowth),
            C.c_int64(ncap), C.c_int64(ldx), _ptr(x, _dp), _ptr(ncols, _ip), _ptr(nconv, _ip), _ptr(status, _sp),
            C.c_int(nthreads))
        return out
    fac = tau = qtb = None
    if want_factors:
        fac = np.zeros((batch, ldx, 2 * l + u + 1))
        tau = np.zeros((batch, ldx))
        qtb = np.zeros((batch, ldx))
        out.update(factors=fac, tau=tau, qtb=qtb)
    L.ila_oracle_qr_banded_solve_f64(
        C.c_int(l), C.c_int(u), C.c_int64(batch), _ptr(p, _dp), _ptr(q, _dp), _ptr(r, _dp), _ptr(shift, _dp),
        C.c_int(hp), _ptr(head, _dp), C.c_int(nb), _ptr(b, _dp), C.c_double(tol), C.c_int(colgrowth),
        C.c_int64(ncap), C.c_int64(ldx), _ptr(x, _dp), _ptr(ncols, _ip), _ptr(nconv, _ip), _ptr(fac, _dp),
        _ptr(tau, _dp), _ptr(qtb, _dp), _ptr(status, _sp), C.c_int(nthreads))
    return out


def qr_banded_solve_c64(l, u, p, q, r=None, shift=None, head=None, b=(1.0,), tol=np.finfo(np.float64).tiny,
                        colgrowth=1000, ncap=200_000, batch=None, nthreads=0):
    """Complex-eltype adaptive QR solve (oracle/ila_oracle_c64.c): p, q, shift, head, b complex; r real.
    Returns dict(x (batch x ncap complex, zero padded), ncols, nconv, status)."""
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
    x = np.zeros((batch, ncap), dtype=cz)
    ncols = np.zeros(batch, dtype=np.int64)
    nconv = np.zeros(batch, dtype=np.int64)
    status = np.zeros(batch, dtype=np.int32)
    vp = C.c_void_p
    pv = lambda a: None if a is None else vp(a.ctypes.data)
    L = lib()
    L.ila_oracle_qr_banded_solve_c64.restype = C.c_int
    rc = L.ila_oracle_qr_banded_solve_c64(
        C.c_int(l), C.c_int(u), C.c_int64(batch), pv(p), pv(q), pv(r), pv(shift), C.c_int(hp), pv(head), C.c_int(nb), pv(b),
        C.c_double(tol), C.c_int(colgrowth), C.c_int64(ncap), pv(x), C.c_int64(ncap), _ptr(ncols, _ip), _ptr(nconv, _ip),
        _ptr(status, _sp), C.c_int(nthreads))
    assert rc == 0
    return dict(x=x, ncols=ncols, nconv=nconv, status=status)


def revchol_tridiag(ga, gb, n, shift=None, a_head=None, b_head=None, batch=None, nthreads=0):
    """Adaptive reverse Cholesky A = L'L of ∞ symmetric tridiagonal operators (oracle/ila_oracle_revchol.c).
    ga, gb: batch x 5 generator coefficients (c0, p, q, r, s): value(k) = c0 + (p + q k)/(r + s k).
    Returns dict(la (batch x n diagonal of L), lb (batch x n sub-diagonal of L), N, status)."""
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
    la = np.zeros((batch, n)); lb = np.zeros((batch, n))
    N = np.zeros(batch, dtype=np.int64); status = np.zeros(batch, dtype=np.int32)
    L = lib()
    L.ila_oracle_revchol_tridiag_f64(C.c_int64(batch), _ptr(ga, _dp), _ptr(gb, _dp), _ptr(shift, _dp), C.c_int(hp),
                                     _ptr(a_head, _dp), _ptr(b_head, _dp), C.c_int64(n), _ptr(la, _dp), _ptr(lb, _dp),
                                     _ptr(N, _ip), _ptr(status, _sp), C.c_int(nthreads))
    return dict(la=la, lb=lb, N=N, status=status)


def ql_finite_section(l, u, g, j=50, shift=None, head=None, tol=np.finfo(np.float64).eps, maxN=10000, batch=None,
                      nthreads=0):
    """Adaptive finite-section QL of ∞ banded operators (oracle/ila_oracle_fsql.c; reference src/infql.jl:364-462).
    g: (batch x) (l+u+1) x 5 generator coefficients (c0, p, q, r, s) per diagonal, data-row order (+u first):
    value(t) = c0 + (p + q t)/(r + s t), t = min(row, col) 0-based.  Complex g / shift / head select the complex eltype.
    Returns dict(Lband (batch x j x (l+u+1): Lband[k, c] = L[k, k-c]), V (batch x j x u reflectors), tau (batch x j), N, status)."""
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
    Lband = np.zeros((batch, j, nd), dtype=dt); V = np.zeros((batch, j, u), dtype=dt); tau = np.zeros((batch, j), dtype=dt)
    N = np.zeros(batch, dtype=np.int64); status = np.zeros(batch, dtype=np.int32)
    L = lib()
    fn = L.ila_oracle_ql_finite_section_c64 if cplx else L.ila_oracle_ql_finite_section_f64
    fn.restype = C.c_int
    pv = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
    rc = fn(C.c_int(l), C.c_int(u), C.c_int64(batch), pv(g), pv(shift), C.c_int(hp), pv(head), C.c_int64(j), C.c_double(tol),
            C.c_int64(maxN), pv(Lband), pv(V), pv(tau), _ptr(N, _ip), _ptr(status, _sp), C.c_int(nthreads))
    assert rc == 0, rc
    return dict(Lband=Lband, V=V, tau=tau, N=N, status=status)


def chol_banded_solve(u, p, q, r=None, shift=None, head=None, b=(1.0,), tol=np.finfo(np.float64).tiny,
                      colgrowth=1000, ncap=20_000, batch=None, want_factors=False, nthreads=0):
    """Batched adaptive Cholesky solve through the oracle (upper bands, data-row order +u … 0)."""
    nd = u + 1
    if batch is None:
        batch = int(np.asarray(p).reshape(-1, nd).shape[0])
    p, q, r, shift, hp, head = _prep_op(nd, batch, p, q, r, shift, head)
    b = np.asarray(b, dtype=np.float64)
    if b.ndim == 1:
        b = np.broadcast_to(b, (batch, b.shape[0]))
    b = np.ascontiguousarray(b)
    nb = b.shape[1]
    ldx = ncap
    x = np.zeros((batch, ldx))
    ncols = np.zeros(batch, dtype=np.int64)
    nconv = np.zeros(batch, dtype=np.int64)
    status = np.zeros(batch, dtype=np.int32)
    out = dict(x=x, ncols=ncols, nconv=nconv, status=status)
    fac = y = None
    if want_factors:
        fac = np.zeros((batch, ldx, u + 1))
        y = np.zeros((batch, ldx))
        out.update(factors=fac, y=y)
    lib().ila_oracle_chol_banded_solve_f64(
        C.c_int(u), C.c_int64(batch), _ptr(p, _dp), _ptr(q, _dp), _ptr(r, _dp), _ptr(shift, _dp), C.c_int(hp),
        _ptr(head, _dp), C.c_int(nb), _ptr(b, _dp), C.c_double(tol), C.c_int(colgrowth), C.c_int64(ncap),
        C.c_int64(ldx), _ptr(x, _dp), _ptr(ncols, _ip), _ptr(nconv, _ip), _ptr(fac, _dp), _ptr(y, _dp),
        _ptr(status, _sp), C.c_int(nthreads))
    return out


def tridiagonal_conjugation(U, X, n, want_ux=False, nthreads=0):
    """Y = Tridiagonal(U*X/U) by the reference's two sequential sweeps (oracle/ila_oracle_triconj.c;
    src/banded/tridiagonalconjugation.jl:10-215, 260-282).  U: (batch x) nbands x m bands (U[d, k] = U[k+1, k+1+d]),
    X: 3 x m shared or batch x 3 x m (dl, d, du), m >= n+1.  Returns Y (batch x 3 x n) [and UX]."""
    U = np.ascontiguousarray(np.asarray(U, dtype=np.float64))
    if U.ndim == 2:
        U = U[None]
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    batch, nbands, uld = U.shape
    shared = X.ndim == 2
    xld = X.shape[-1]
    assert uld >= n + 1 and xld >= n + 1 and nbands in (2, 3)
    Y = np.zeros((batch, 3, n)); UX = np.zeros((batch, 3, n)) if want_ux else None
    lib().ila_oracle_triconj_f64(C.c_int(nbands), C.c_int64(batch), C.c_int64(n), _ptr(U, _dp), C.c_int64(uld), _ptr(X, _dp),
                                 C.c_int64(xld), C.c_int64(0 if shared else 3 * xld), _ptr(Y, _dp), _ptr(UX, _dp), C.c_int(nthreads))
    return (Y, UX) if want_ux else Y


def bidiagonal_conjugation(U, Cm, n, uplo="U", nthreads=0):
    """Bidiagonal part of inv(U)*C (C = X*V) from the tridiagonal parts of U and C (src/banded/bidiagonalconjugation.jl:37-60).
    U, Cm: (batch x) 3 x m (sub, main, super), m >= n+1.  Returns dv, ev (batch x n)."""
    U = np.ascontiguousarray(np.asarray(U, dtype=np.float64)); Cm = np.ascontiguousarray(np.asarray(Cm, dtype=np.float64))
    if U.ndim == 2:
        U = U[None]; Cm = Cm[None]
    batch, _, ld = U.shape
    assert Cm.shape == U.shape and ld >= n + 1
    dv = np.zeros((batch, n)); ev = np.zeros((batch, n))
    lib().ila_oracle_biconj_f64(C.c_int(1 if uplo == "U" else 0), C.c_int64(batch), C.c_int64(n), _ptr(U, _dp), C.c_int64(ld),
                                _ptr(Cm, _dp), _ptr(dv, _dp), _ptr(ev, _dp), C.c_int(nthreads))
    return dv, ev


def num_threads() -> int:
    return int(lib().ila_oracle_num_threads())


# -- operator descriptions of the reference's fixtures --------------------------------------------

def bessel_operator(z):
    """A = BandedMatrix(0 => -2*(0:∞)/z, 1 => Ones(∞), -1 => Ones(∞)) (README.md:41, test_infqr.jl:95).
    Rows of (p,q,r) follow BandedMatrices data-row order: diagonal +1, 0, -1."""
    z = np.atleast_1d(np.asarray(z, dtype=np.float64))
    batch = z.shape[0]
    p = np.tile(np.array([1.0, 0.0, 1.0]), (batch, 1))
    q = np.tile(np.array([0.0, -2.0, 0.0]), (batch, 1))
    r = np.ones((batch, 3))
    r[:, 1] = z
    return p, q, r
