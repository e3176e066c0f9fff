"""CPU oracle for the adaptive QR solve of ∞ ALMOST-BANDED operators — TEST INFRASTRUCTURE ONLY.

A = Vcat(B, D): r dense rows B (boundary conditions / functionals) on top of a banded operator D (bandwidths (lD, uD)) —
the ultraspherical-spectral-method shape of test/test_infqr.jl:193-267.  As a matrix A has almost-bandwidths
(l, u) = (lD + r, uD - r) and rank-r fill above the band.

Restates, on a DENSE finite section (no semiseparable representation at all — the kernel keeps the fill as coefficient
vectors times the boundary rows), the reference path
    AdaptiveQRData(::AbstractAlmostBandedLayout, A)         src/infqr.jl:15-21
    partialqr!(F, n) -> _almostbanded_qr!                   src/infqr.jl:50-66   [SemiseparableMatrices ≥ 0.4, un-vendored:
                                                            Householder QR, reflector k over rows k..k+l, LAPACK conventions]
    Q'b adaptive driver (COLGROWTH = 1000, tol floatmin)    src/infqr.jl:236-274 (shared with the banded layout)
    ldiv!(R, B)                                             src/infqr.jl:350-355
A Householder reflector of column k only involves rows k..k+l, so the leading N x N part of the factors of the ∞ operator
equals that of the (N + l) x N section: the oracle factorizes such a section densely.

Operator description (covers every almost-banded fixture of the reference's tests):
    D's diagonals, data-row order +uD first: value(t) = g0 + g1 t + g2 t^2, t = 0-based ROW of D (Ones, 1:∞, (1:∞).*(2:∞), 4:2:∞)
    boundary rows: B_m(j) = s^j (a + b j + c j^2), s = ±1, j = 0-based column (Ones, (-1)^k)
    shift: subtracted from D's main diagonal (batched families D - λ I).
"""
from __future__ import annotations

import numpy as np

FLOATMIN = np.finfo(np.float64).tiny


def boundary_row(bg, n):
    s, a, b, c = bg
    j = np.arange(n, dtype=np.float64)
    sign = np.where(np.arange(n) % 2 == 0, 1.0, s) if s < 0 else np.ones(n)
    return sign * (a + b * j + c * j * j)


def dense_section(lD, uD, dg, bg, shift, nrows, ncols):
    """A[0:nrows, 0:ncols] of Vcat(B, D - shift I)."""
    dg = np.asarray(dg, dtype=np.float64).reshape(lD + uD + 1, 3)
    bg = np.asarray(bg, dtype=np.float64).reshape(-1, 4)
    r = bg.shape[0]
    A = np.zeros((nrows, ncols))
    for m in range(min(r, nrows)):
        A[m] = boundary_row(bg[m], ncols)
    for rr, d in enumerate(range(uD, -lD - 1, -1)):
        g0, g1, g2 = dg[rr]
        for t in range(max(0, -d), nrows - r):          # row t of D, column t + d
            j = t + d
            if j >= ncols:
                break
            v = g0 + g1 * t + g2 * t * t
            if d == 0:
                v -= shift
            A[r + t, j] = v
    return A


def almostbanded_qr_solve(lD, uD, dg, bg, b, shift=0.0, tol=FLOATMIN, colgrowth=1000, ncap=6000):
    """Vcat(B, D - shift I) \\ [b; 0; ...] by the reference's adaptive QR.  Returns dict(x, ncols, nconv, status, qtb, R (dense))."""
    bg = np.asarray(bg, dtype=np.float64).reshape(-1, 4)
    r = bg.shape[0]
    l, u = lD + r, uD - r
    b = np.asarray(b, dtype=np.float64)
    nb = len(b)
    N = int(ncap)
    M = dense_section(lD, uD, dg, bg, shift, N + l + 1, N)
    tau = np.zeros(N)
    Bv = np.zeros(N + l + 1)
    Bv[:nb] = b
    state = dict(ncols=0)

    def partialqr(n):
        for k in range(state["ncols"], n):
            x = M[k:k + l + 1, k]
            alpha = x[0]
            xnorm = np.sqrt(x[1:] @ x[1:])
            if xnorm == 0.0:
                tau[k] = 0.0
                continue
            beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
            tau[k] = (beta - alpha) / beta
            v = x[1:] * (1.0 / (alpha - beta))
            M[k, k] = beta
            M[k + 1:k + l + 1, k] = v
            Cm = M[k:k + l + 1, k + 1:]
            w = Cm[0] + v @ Cm[1:]
            Cm[0] -= tau[k] * w
            Cm[1:] -= np.outer(v, tau[k] * w)
        state["ncols"] = max(state["ncols"], n)

    status, nconv, n = 0, 0, nb
    jfirst, jlast = 0, colgrowth - 1
    while True:
        cs_max = jlast + l
        if cs_max + 1 + l + u + 2 > N:
            status = 3
            break
        partialqr(jlast + 1)
        n = max(n, cs_max + 1)
        if jfirst + 1 > nb:
            mx = np.abs(Bv[jfirst:jfirst + l + 1]).max()
            if np.isnan(mx):
                status = 5
                nconv = jfirst + 1
                break
            if mx <= tol:
                nconv = jfirst + 1
                break
        partialqr(cs_max + 1)
        for k in range(jfirst, jlast + 1):
            if tau[k] == 0.0:
                continue
            v = M[k + 1:k + l + 1, k]
            w = tau[k] * (Bv[k] + v @ Bv[k + 1:k + l + 1])
            Bv[k] -= w
            Bv[k + 1:k + l + 1] -= v * w
        jfirst, jlast = jlast + 1, jlast + colgrowth
    out = dict(status=status, nconv=nconv)
    if status == 3:
        out.update(x=np.zeros(0), ncols=0, qtb=np.zeros(0))
        return out
    if not np.any(np.abs(Bv[:n]) > tol):
        n = 0
    partialqr(n)
    R = np.triu(M[:n, :n])
    qtb = Bv[:n].copy()
    x = np.zeros(n)
    for i in range(n - 1, -1, -1):
        x[i] = (qtb[i] - R[i, i + 1:] @ x[i + 1:]) / R[i, i]
    out.update(x=x, ncols=n, qtb=qtb, R=R)
    return out
