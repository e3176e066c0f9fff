"""CPU oracle for the adaptive block-banded QR solve — TEST INFRASTRUCTURE ONLY.

Restates, on a dense global matrix (nothing like the kernel's panel scheme), the reference path for ∞ block-banded
operators with block sizes 1, 2, 3, … (the `KronTrav` / triangular-traversal operators of 2-D problems):

    AdaptiveQRData(::AbstractBlockLayout, A)               src/infqr.jl:23-28    block bandwidths (l, l+u)
    partialqr!(F, N::Block{1}) -> _blockbanded_qr!         src/infqr.jl:68-86    [BlockBandedMatrices ≥ 0.13: per block column
                                                            a dense Householder QR of the rows of blocks J..J+l (LAPACK geqrf
                                                            conventions: beta = -sign(alpha)‖x‖, tau = (beta-alpha)/beta, tau = 0
                                                            when the sub-column vanishes), applied to block columns J+1..J+l+u]
    Q'b adaptive driver, COLGROWTH = 300, tolerance 1e-30   src/infqr.jl:303-337
    resizedata_chop!(::BlockedVector, tol)                  src/infqr.jl:221-232  keep up to the end of the block of the last |c| > tol
    ldiv!(R, B::BlockedArray)                               src/infqr.jl:358-366  back-substitution on the leading n x n blocks

Operator family (every block-banded fixture of test/test_infqr.jl:270-313):
    L = alpha * KronTrav(T1, I) + beta * KronTrav(I, T2) - lam * I,      T1, T2 tridiagonal (bands dl, d, du),
with KronTrav(A, B)[Block(K, J)][k, j] = A[k, j] * B[K-k+1, J-j+1] (LazyBandedMatrices, an un-vendored dependency; the
convention is pinned by the golden vectors (F.Q'e1)[1:6] and (F.Q e1)[1:6] of test_infqr.jl:276-277 and by
u[Block(K)][K] / u[Block(K)][1] in :284-290).  Block K holds K entries; block bandwidths (1, 1).
"""
from __future__ import annotations

import numpy as np

COLGROWTH = 300
TOL = 1e-30


def block_start(K):
    """0-based global index of the first entry of block K (1-based)."""
    return K * (K - 1) // 2


def findblock(i):
    """Block (1-based) that holds the 1-based global index i."""
    K = int((np.sqrt(8.0 * i + 1) - 1) / 2)
    while K * (K + 1) // 2 < i:
        K += 1
    while K > 1 and (K - 1) * K // 2 >= i:
        K -= 1
    return K


def toeplitz_bands(d, e, m):
    """Symmetric tridiagonal Toeplitz 1-D operator as bands (dl, d, du) of length m."""
    return np.stack([np.full(m, float(e)), np.full(m, float(d)), np.full(m, float(e))])


def dense_operator(T1, T2, alpha, beta, lam, nblocks):
    """Leading nblocks x nblocks blocks of L as a dense matrix.  T1, T2: 3 x m bands (dl[k] = T[k+2,k+1], d, du[k] = T[k+1,k+2])."""
    N = block_start(nblocks + 1)
    A = np.zeros((N, N))
    for K in range(1, nblocks + 1):
        rK = block_start(K)
        for k in range(1, K + 1):
            r = rK + k - 1
            # diagonal block
            A[r, r] = alpha * T1[1, k - 1] + beta * T2[1, K - k] - lam
            if K + 1 <= nblocks:
                cJ = block_start(K + 1)
                A[r, cJ + k] += alpha * T1[2, k - 1]            # (k, k+1): T1[k, k+1]
                A[r, cJ + k - 1] += beta * T2[2, K - k]         # (k, k):   T2[K-k+1, K-k+2]
            if K - 1 >= 1:
                cJ = block_start(K - 1)
                if k - 1 >= 1:
                    A[r, cJ + k - 2] += alpha * T1[0, k - 2]    # (k, k-1): T1[k, k-1]
                if k <= K - 1:
                    A[r, cJ + k - 1] += beta * T2[0, K - k - 1]  # (k, k):   T2[K-k+1, K-k]
    return A


class _Factor:
    """Dense storage of the growing factorization (R above the diagonal, reflectors below, tau)."""

    def __init__(self, T1, T2, alpha, beta, lam, maxblocks):
        self.l, self.u = 1, 1
        self.maxblocks = maxblocks
        self.M = dense_operator(T1, T2, alpha, beta, lam, maxblocks + 3)
        self.tau = np.zeros(self.M.shape[0])
        self.nbc = 0           # block columns factorized

    def partialqr(self, N):
        M, tau = self.M, self.tau
        for J in range(self.nbc + 1, N + 1):
            rend = block_start(J + self.l + 1)                 # exclusive: rows through block J+l
            cend = block_start(J + self.l + self.u + 1)        # exclusive: columns through block J+l+u
            for c in range(block_start(J), block_start(J + 1)):
                x = M[c:rend, c]
                alpha = x[0]
                xnorm = np.sqrt(np.dot(x[1:], x[1:]))
                if xnorm == 0.0:
                    tau[c] = 0.0
                    continue
                beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
                tau[c] = (beta - alpha) / beta
                v = x[1:] * (1.0 / (alpha - beta))
                M[c, c] = beta
                M[c + 1:rend, c] = v
                C = M[c:rend, c + 1:cend]
                w = C[0] + v @ C[1:]
                C[0] -= tau[c] * w
                C[1:] -= np.outer(v, tau[c] * w)
        self.nbc = max(self.nbc, N)

    def apply_qt(self, bvec, J0, J1):
        """b <- Q_N' b with the reflectors of block columns J0..J1 (rows through block J1+l)."""
        M, tau = self.M, self.tau
        for J in range(J0, J1 + 1):
            rend = block_start(J + self.l + 1)
            for c in range(block_start(J), block_start(J + 1)):
                if tau[c] == 0.0:
                    continue
                v = M[c + 1:rend, c]
                w = bvec[c] + v @ bvec[c + 1:rend]
                bvec[c] -= tau[c] * w
                bvec[c + 1:rend] -= v * (tau[c] * w)


def block_qr_solve(T1, T2, alpha, beta, lam, b=(1.0,), tol=TOL, colgrowth=COLGROWTH, maxblocks=64):
    """(alpha KronTrav(T1,I) + beta KronTrav(I,T2) - lam I) \\ [b; 0; ...] by the reference's adaptive block-banded QR.
    Returns dict(x (length ncols), ncols, nblocks (block columns whose reflectors reached b), qtb (Q'b before the chop),
    status (0 ok, 3 maxblocks reached), F (the dense factor storage, for tests))."""
    F = _Factor(T1, T2, alpha, beta, lam, maxblocks)
    l = F.l
    b = np.asarray(b, dtype=np.float64)
    c = np.zeros(F.M.shape[0])
    c[: len(b)] = b
    SB = findblock(len(b))
    J0, J1 = 1, findblock(colgrowth)
    status, done = 0, 0
    while True:
        lo, hi = block_start(J0), block_start(J0 + l + 1)
        mx = np.abs(c[lo:hi]).max()
        if np.isnan(mx):
            status = 5
            break
        if J0 > SB and mx <= tol:
            break
        if J1 + l > maxblocks:                           # capacity (the reference has none): last chunk clamped, then status 3
            J1 = maxblocks - l
            if J1 < J0:
                status = 3
                break
        F.partialqr(J1 + l)
        F.apply_qt(c, J0, J1)
        done = J1
        last_jr = block_start(J1 + 1)                    # 1-based last column of JR
        J0, J1 = J1 + 1, findblock(last_jr + colgrowth)
    out = dict(status=status, nblocks=done, qtb=c[: block_start(done + l + 1)].copy(), F=F)
    if status != 0:
        out.update(x=np.zeros(0), ncols=0)
        return out
    data = c[: block_start(done + l + 1)]
    nz = np.nonzero(np.abs(data) > tol)[0]
    if len(nz) == 0:
        out.update(x=np.zeros(0), ncols=0)
        return out
    K = findblock(int(nz[-1]) + 1)
    n2 = block_start(K + 1)
    F.partialqr(K)
    R = np.triu(F.M[:n2, :n2])
    x = np.zeros(n2)
    for i in range(n2 - 1, -1, -1):                     # back-substitution, row-oriented
        x[i] = (data[i] - R[i, i + 1:n2] @ x[i + 1:n2]) / R[i, i]
    out.update(x=x, ncols=n2)
    return out
