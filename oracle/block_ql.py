"""CPU oracle for the block-tridiagonal Toeplitz-tail QL (TEST INFRASTRUCTURE — never imported by the product path).

numpy restatement of `blocktailiterate` and `_blocktripert_ql` (reference src/infql.jl:166-209): the matrix fixed point
(d, e) of the block QL tail by direct iteration of a dense 2n x 3n Householder QL, then the finite block head.
The dense QL is `toeplitz_ql.dense_ql` (loop of src/infql.jl:10-22 with the `(T<:Real)` bound of examples/jacobi.jl:63).

Pinned by: test/test_periodic.jl:5-18 (factors / tau / L against a 100-block finite section), :28-34 (degenerate
|L[1,1]| <= 1e-11), test/test_infql.jl:278-287 (L[1,1] real at x = -0.95), SURVEY Appendix B (7-51 iterations on the
example's energies) — see tests/test_oracle_block.py.
"""
from __future__ import annotations

import numpy as np

from .toeplitz_ql import _reflector, _reflector_apply, dense_ql

STATUS_OK = 0
STATUS_NOT_CONVERGED = 3


def _norm_seq(x):
    """LinearAlgebra.norm of a short vector (generic_norm2): sqrt of the sequential sum of abs2."""
    s = 0.0
    first = True
    for v in x:
        t = (v.real * v.real + v.imag * v.imag) if np.iscomplexobj(x) else v * v
        s = t if first else s + t
        first = False
    return np.sqrt(s)


def dense_ql_seq(A):
    """Householder QL with explicitly sequential sums (the order of the Julia loops, src/infql.jl:10-22 and
    reflector!/reflectorApply!); real eltype skips k = 1 (examples/jacobi.jl:63).  Returns (factors, tau)."""
    A = np.array(A, copy=True)
    cplx = np.iscomplexobj(A)
    m, n = A.shape
    mn = min(m, n)
    tau = np.zeros(mn, dtype=A.dtype)
    for k in range(mn, (1 if cplx else 2) - 1, -1):
        mu, nu = m - mn + k, n - mn + k
        x = A[:mu, nu - 1][::-1].copy()          # view(A, mu:-1:1, nu)
        xi1 = x[0]
        normu = _norm_seq(x)
        if normu == 0:
            t = 0.0
        else:
            v = np.copysign(normu, xi1.real)
            xi1 = xi1 + v
            x[0] = -v
            for i in range(1, mu):
                x[i] = x[i] / xi1
            t = xi1 / v
        A[:mu, nu - 1] = x[::-1]
        tau[k - 1] = t
        for j in range(nu - 1):                   # reflectorApply!
            vA = A[mu - 1, j]
            for i in range(1, mu):
                vA = vA + np.conj(x[i]) * A[mu - 1 - i, j]
            vA = np.conj(t) * vA
            A[mu - 1, j] = A[mu - 1, j] - vA
            for i in range(1, mu):
                A[mu - 1 - i, j] = A[mu - 1 - i, j] - x[i] * vA
    return A, tau


def _packed_q_apply(factors, tau, B):
    """QLPackedQ(factors, tau) * B for square packed factors (lmul!: H_1 first, then H_2, ...; sequential sums)."""
    B = np.array(B, copy=True)
    n = factors.shape[0]
    for k in range(1, n + 1):
        for j in range(B.shape[1]):
            vB = B[k - 1, j]
            for i in range(k - 1):
                vB = vB + np.conj(factors[i, k - 1]) * B[i, j]
            vB = tau[k - 1] * vB
            B[k - 1, j] = B[k - 1, j] - vB
            for i in range(k - 1):
                B[i, j] = B[i, j] - factors[i, k - 1] * vB
    return B


def _fro_seq(M):
    return _norm_seq(M.reshape(-1, order="F"))


def blocktailiterate(c, a, b, d=None, e=None, atol=1e-10, maxit=1_000_000):
    """src/infql.jl:166-182.  Returns (X (2n x 3n), tau_tail (n), iterations, status)."""
    c, a, b = (np.asarray(m) for m in (c, a, b))
    d = c if d is None else np.asarray(d)
    e = a if e is None else np.asarray(e)
    n = c.shape[0]
    dt = np.result_type(c, a, b, d, e, np.float64)
    z = np.zeros((n, n), dtype=dt)
    for it in range(1, maxit + 1):
        X = np.block([[c, a, b], [z, d, e]]).astype(dt)
        F, tau = dense_ql_seq(X)
        # F.L = tril(factors, n): L[1:n,1:n] is full, L[1:n,n+1:2n] lower triangular
        dn = F[:n, :n].copy()
        en = np.tril(F[:n, n:2 * n])
        fq, tq = F[:n, n:2 * n], tau[:n]
        dn = _packed_q_apply(fq, tq, dn)     # undo last rotation (src/infql.jl:174)
        en = _packed_q_apply(fq, tq, en)
        if _fro_seq(dn - d) <= atol and _fro_seq(en - e) <= atol:
            F[:n, :n] = dn
            F[:n, n:2 * n] = en
            return F, tau[n:2 * n], it, STATUS_OK
        d, e = dn, en
    return None, None, maxit, STATUS_NOT_CONVERGED


def blocktri_ql_l11(dl_head, d_head, du_head, c, a, b, shift=0.0, atol=1e-10, maxit=1_000_000):
    """L[1,1] (and iteration count) of ql(A - shift*I) for the block-tridiagonal perturbed Toeplitz operator
    A = BlockTridiagonal(Vcat(dl_head, Fill(c,∞)), Vcat(d_head, Fill(a,∞)), Vcat(du_head, Fill(b,∞)))
    (src/infql.jl:189-209; dl = sub-diagonal blocks A[Block(k+1,k)], du = super-diagonal blocks A[Block(k,k+1)]).
    """
    c, a, b = (np.asarray(m, dtype=np.result_type(m, np.float64)) for m in (c, a, b))
    n = c.shape[0]
    dl_head = [np.asarray(m) for m in dl_head]
    d_head = [np.asarray(m) for m in d_head]
    du_head = [np.asarray(m) for m in du_head]
    dt = np.result_type(c, a, b, shift, np.float64, *[m for m in dl_head + d_head + du_head])
    I = np.eye(n, dtype=dt)

    def DL(k):  # A[Block(k+1,k)], k 1-based
        return (dl_head[k - 1] if k <= len(dl_head) else c).astype(dt)

    def D(k):
        return (d_head[k - 1] if k <= len(d_head) else a).astype(dt) - shift * I

    def DU(k):  # A[Block(k,k+1)]
        return (du_head[k - 1] if k <= len(du_head) else b).astype(dt)

    N = max(len(du_head) + 1, len(d_head), len(dl_head))
    cc, aa, bb = DL(N), D(N), DU(N - 1) if N >= 2 else DU(1)
    d0, e0 = DU(2), D(3)                     # ql(A) = _blocktripert_ql(A, A[Block(2,3)], A[Block(3,3)])
    X, tau_tail, iters, st = blocktailiterate(cc, aa, bb, d0, e0, atol, maxit)
    if st != STATUS_OK:
        return np.nan, iters, st, None
    # finite head: N x N blocks, last block row replaced by P[Block(1), Block.(1:2)] = (d~, e~)
    H = np.zeros((n * N, n * N), dtype=dt)
    for k in range(1, N + 1):
        H[(k - 1) * n:k * n, (k - 1) * n:k * n] = D(k)
        if k < N:
            H[(k - 1) * n:k * n, k * n:(k + 1) * n] = DU(k)
            H[k * n:(k + 1) * n, (k - 1) * n:k * n] = DL(k)
    H[(N - 1) * n:, :] = 0
    if N >= 2:
        H[(N - 1) * n:, (N - 2) * n:(N - 1) * n] = X[:n, :n]
    H[(N - 1) * n:, (N - 1) * n:] = X[:n, n:2 * n]
    F, tau = dense_ql_seq(H)
    return F[0, 0], iters, STATUS_OK, dict(head_factors=F, head_tau=tau, tail_X=X, tail_tau=tau_tail, n=n, N=N)


def periodic_schrodinger(x):
    """examples/periodicschrodinger.jl:8-17 at energy x."""
    c = np.array([[0.0, 1.0], [0.0, 0.0]])
    a = np.array([[-1.0, 1.0], [1.0, 1.0]])
    b = np.array([[0.0, 0.0], [1.0, 0.0]])
    a0 = a.copy()
    a0[0, 0] = 2.0
    return blocktri_ql_l11([c], [a0], [b], c, a, b, shift=x)


# ----------------------------------------------------------------------------------------------------------------
# the whole ∞ factorization object (src/infql.jl:197-204) and lmul!(Q', b) (src/infql.jl:213-243), on a leading section
# ----------------------------------------------------------------------------------------------------------------

def assemble_inf_factors(ex, n, N, nblocks):
    """Leading (nblocks*n)^2 section of the ∞ `factors` and the first nblocks*n entries of `τ` of
    QL(_BlockSkylineMatrix(Vcat(BB.data, Fill(tail column))), Vcat(F.τ, Fill(τ))) (src/infql.jl:199-204) from
    ex = dict(head_factors, head_tau, tail_X, tail_tau) — block bandwidths (2, 1)."""
    H, ht, X, tt = ex["head_factors"], ex["head_tau"], ex["tail_X"], ex["tail_tau"]
    dt = np.result_type(H, X)
    M = nblocks * n
    F = np.zeros((M + 3 * n, M + n), dtype=dt)
    hn = N * n
    F[:hn, :hn] = H
    P = lambda I, J: X[(I - 1) * n:I * n, (J - 1) * n:J * n]          # P[Block(I, J)], 1-based
    blk = lambda I, J: (slice((I - 1) * n, I * n), slice((J - 1) * n, J * n))
    if N >= 2:
        F[blk(N + 1, N - 1)] = P(2, 1)                                 # BB[Block(N+1), Block.(N-1:N)] = P[Block(2), Block.(1:2)]
    F[blk(N + 1, N)] = P(2, 2)
    F[blk(N + 2, N)] = P(2, 1)                                         # BB[Block(N+2), Block(N)] = P[Block(2), Block(1)]
    for J in range(N + 1, nblocks + 1):                                # Toeplitz tail columns (:203)
        F[blk(J - 1, J)] = P(1, 3)
        F[blk(J, J)] = P(2, 3)
        F[blk(J + 1, J)] = P(2, 2)
        F[blk(J + 2, J)] = P(2, 1)
    tau = np.concatenate([ht, np.tile(tt, nblocks - N + 1)])[:M]
    return F[:M, :M], tau


def qt_times_b(F, tau, b, n, nout):
    """lmul!(Q', b) of src/infql.jl:213-243 on a section of the ∞ factors: k = last(colsupport(b)) + u : -1 : 1 with
    u = 2n - 1 entries above the diagonal (the reference hard-codes 2u_block + 1 = 3: the same for its n = 2 tests)."""
    u = 2 * n - 1
    nb = len(b)
    dt = np.result_type(F, np.asarray(b), np.float64)
    t = np.zeros(F.shape[0], dtype=dt)
    t[:nb] = b
    for k in range(nb + u, 0, -1):
        lo = max(1, k - u)
        vB = t[k - 1]
        for i in range(lo, k):
            vB = vB + np.conj(F[i - 1, k - 1]) * t[i - 1]
        vB = np.conj(tau[k - 1]) * vB
        t[k - 1] = t[k - 1] - vB
        for i in range(lo, k):
            t[i - 1] = t[i - 1] - F[i - 1, k - 1] * vB
    return t[:nout]


def dense_Q_L(F, tau, n):
    """Dense Q (leading rows) and L of a section of the ∞ factors: L = lower triangle, Q column i = conj(Q' e_i)."""
    M = F.shape[0]
    L = np.tril(F)
    m = M - 3 * n                                                       # rows whose reflectors all lie inside the section
    Q = np.zeros((m, m), dtype=np.result_type(F, np.float64))
    for i in range(m):
        e = np.zeros(i + 1)
        e[i] = 1.0
        Q[i, :] = np.conj(qt_times_b(F, tau, e, n, m))
    return Q, L
