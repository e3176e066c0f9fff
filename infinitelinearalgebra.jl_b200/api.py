"""Host-side mirror of the reference's Julia interface for the hot path (names, argument meaning, error behaviour).

Julia is not available in this image, so this thin Python layer plays the role of the `ccall` shim
(`julia/ILAB200.jl`): it only describes operators symbolically and marshals them to the C ABI; every factorization
runs in the CUDA library.  The spelling follows the reference so that tests read like its own:

    A = BandedMatrix({-3: Fill(7/10), -2: Fill(1), 1: Fill(2j)})          # README.md:83-85
    ql(A - 5*I).L[1, 1]                                                    # src/banded/infqltoeplitz.jl:191
    ql_L11(A, lambdas)                                                     # batched form over a vector of shifts
    A = BandedMatrix({0: Affine(0, -2, z), 1: Ones(), -1: Ones()})        # README.md:41
    J = qr(A) / Vcat([besselj1], Zeros())    ( `A \\ b` is spelled  qr(A).solve(b)  or  ldiv(A, b) )
    cholesky(S).solve(b)                                                   # src/infcholesky.jl:71-75
    ql(BlockTridiagonal(...) - x*I).L[1, 1]                                # src/infql.jl:207

Indices in `.L[i, j]` / `.Q[i, j]` are 1-based like the reference.
"""
from __future__ import annotations

import numpy as np

from . import batched as _b


class DomainError(ValueError):
    """The reference's DomainError (src/banded/infqltoeplitz.jl:12-13,23)."""


class NotConvergedError(RuntimeError):
    """`error("Did not converge")` (src/infql.jl:181) / capacity reached."""


class PosDefException(ArithmeticError):
    pass


_QL_MSG = ("QL factorization does not exist. This could indicate that the operator is not Fredholm or that the "
           "dimension of the kernel exceeds that of the co-kernel. Try again with the adjoint.")
_QL_REAL_MSG = ("Real-valued QL factorization does not exist. Try ql(complex(A)) to see if a complex-valued QL "
                "factorization exists.")


def _raise_status(st, what):
    st = int(st)
    if st == 1:
        raise DomainError(_QL_MSG)
    if st == 2:
        raise DomainError(_QL_REAL_MSG)
    if st == 3:
        raise NotConvergedError(f"{what}: did not converge within the capacity / iteration limit")
    if st == 4:
        raise PosDefException(f"{what}: matrix is not positive definite")
    if st == 5:
        raise ArithmeticError("Not-a-number encounted")


# ---- lazy infinite vectors (the little of InfiniteArrays / FillArrays the fixtures need) -------------------

class Affine:
    """Entry k (k = 0, 1, ...) = (p + q*k)/r  — covers Fill, 1:∞, -2(0:∞)/z, 2 .+ ε(1:∞)."""

    def __init__(self, p, q=0.0, r=1.0):
        self.p, self.q, self.r = p, q, r

    def value(self, k):
        return (self.p + self.q * k) / self.r


class Rational:
    """Entry k (k = 0, 1, ...) = c0 + (p + q*k)/(r + s*k) — e.g. `1 ./ (2:∞)` = Rational(0, 1, 0, 2, 1)."""

    def __init__(self, c0=0.0, p=0.0, q=0.0, r=1.0, s=0.0):
        self.c = (c0, p, q, r, s)


def Fill(v):
    return Affine(v, 0.0, 1.0)


def Ones():
    return Fill(1.0)


def Zeros():
    return Fill(0.0)


class Vcat:
    """Vcat(head_values, tail) — finite head followed by a lazy tail (Affine), positions counted from the start."""

    def __init__(self, head, tail):
        self.head = list(np.atleast_1d(head)) if not isinstance(head, list) else head
        self.tail = tail


class UniformScaling:
    def __init__(self, lam=1.0):
        self.lam = lam

    def __rmul__(self, s):
        return UniformScaling(self.lam * s)

    __mul__ = __rmul__


I = UniformScaling(1.0)


class BandedMatrix:
    """∞ banded operator given by its diagonals `{d: diag}` (d = col - row, as in `BandedMatrix(d => v)`)."""

    def __init__(self, diags: dict, shift=0.0):
        self.diags = {int(d): (v if isinstance(v, (Affine, Rational, Vcat)) else Fill(v)) for d, v in diags.items()}
        self.shift = shift

    @property
    def bandwidths(self):
        return (max(0, -min(self.diags)), max(0, max(self.diags)))

    def __sub__(self, other):
        if isinstance(other, UniformScaling):
            return BandedMatrix(self.diags, self.shift + other.lam)
        return NotImplemented

    def __add__(self, other):
        if isinstance(other, UniformScaling):
            return BandedMatrix(self.diags, self.shift - other.lam)
        return NotImplemented

    def _diag(self, d):
        return self.diags.get(d, Fill(0.0))

    def entry(self, i, j):
        """A[i, j], 0-based."""
        d = j - i
        if d not in self.diags:
            return -self.shift if d == 0 else 0.0
        t = min(i, j)
        v = self.diags[d]
        val = (v.head[t] if t < len(v.head) else v.tail.value(t)) if isinstance(v, Vcat) else v.value(t)
        return val - self.shift if d == 0 else val

    def finite(self, n, m=None):
        m = n if m is None else m
        cplx = np.iscomplexobj(np.array([self.entry(0, 0), self.shift] + [self.entry(max(0, -d), max(0, d)) for d in self.diags]))
        A = np.zeros((n, m), dtype=complex if cplx else float)
        l, u = self.bandwidths
        for i in range(n):
            for j in range(max(0, i - l), min(m, i + u + 1)):
                A[i, j] = self.entry(i, j)
        return A

    # -- descriptions for the C ABI ------------------------------------------------------------------
    def is_toeplitz(self):
        return all(isinstance(v, Affine) and v.q == 0 for v in self.diags.values())

    def tail_column(self):
        """BandedMatrices data column of the Toeplitz tail: [a_{+u}, ..., a_0, ..., a_{-l}] (no shift)."""
        l, u = self.bandwidths
        col = []
        for d in range(u, -l - 1, -1):
            v = self._diag(d)
            t = v.tail if isinstance(v, Vcat) else v
            if t.q != 0:
                raise ValueError("operator does not have a Toeplitz tail")
            col.append(t.p / t.r)
        return np.array(col)

    def head_rows(self):
        """Smallest p such that row p of A (1-based) is already pure Toeplitz and p >= l + 1."""
        l, _ = self.bandwidths
        p = l + 1
        for d, v in self.diags.items():
            if isinstance(v, Vcat):
                p = max(p, len(v.head) + 1 + max(0, -d))
        return p

    def head_width(self):
        return max([len(v.head) for v in self.diags.values() if isinstance(v, Vcat)] + [0])

    def generators(self):
        """(p, q, r, head) rows in data-row order d = u, ..., -l for the QR / Cholesky entries."""
        l, u = self.bandwidths
        hp = self.head_width()
        p, q, r = [], [], []
        head = np.zeros((l + u + 1, hp)) if hp else None
        for rr, d in enumerate(range(u, -l - 1, -1)):
            v = self._diag(d)
            t = v.tail if isinstance(v, Vcat) else v
            p.append(t.p), q.append(t.q), r.append(t.r)
            for k in range(hp):
                head[rr, k] = (v.head[k] if (isinstance(v, Vcat) and k < len(v.head)) else t.value(k))
        return np.array(p, float), np.array(q, float), np.array(r, float), head


class Symmetric:
    def __init__(self, A: BandedMatrix):
        self.parent = A


class _OneBased:
    def __init__(self, fn):
        self._fn = fn

    def __getitem__(self, ij):
        i, j = ij
        if isinstance(i, slice) or isinstance(j, slice):
            ri = range(i.start or 1, i.stop + 1) if isinstance(i, slice) else [i]
            rj = range(j.start or 1, j.stop + 1) if isinstance(j, slice) else [j]
            return np.array([[self._fn(a, b) for b in rj] for a in ri])
        return self._fn(i, j)


# ---- ql ----------------------------------------------------------------------------------------------------

class QLHessenberg:
    """Result of ql on a (perturbed) Toeplitz operator with u == 1 (src/banded/hessenbergq.jl:16-72)."""

    def __init__(self, l, X, tau, head_factors=None, head_l11=None, source=None):
        self.l, self.X, self.tau = l, X, tau
        self._head, self._head_l11 = head_factors, head_l11
        self._source = source                         # (data column, shift, branch) of a pure Toeplitz operator
        self.L = _OneBased(self._L)
        self.Q = _OneBased(self._Q)

    def _L(self, i, j):
        if self._head is not None:
            p = self._head.shape[0]
            if i <= p and j <= p:
                return self._head[i - 1, j - 1] if i >= j else 0.0
            raise IndexError("only the p x p head of a perturbed operator's L is materialised")
        m = self.X.shape[1]
        if i < j or i - j > m - 1:
            return 0.0
        if j == 1:                                   # first column [X[1,m-1]; X[2,m-1:-1:1]]  (infqltoeplitz.jl:63)
            return self.X[0, m - 2] if i == 1 else self.X[1, m - i]
        return self.X[1, m - 1 - (i - j)]            # tail column X[2, m:-1:1]

    def solve(self, b, n=1000):
        """First n entries of `F \\ b` (ldiv!(F.L, lmul!(F.Q', b)), src/infql.jl:266-289) for a pure Toeplitz operator."""
        bb = np.atleast_1d(np.asarray(b.head if isinstance(b, Vcat) else b))
        if self._source is None:
            raise NotImplementedError("F \\ b needs the operator the factorization came from")
        if isinstance(self._source[0], str) and self._source[0] == "pert":
            _, head, col, lam, branch = self._source
            x, st = _b.ql_perttoeplitz_solve(head, col, np.array([lam]), bb, int(n), branch_rank=branch)
            _raise_status(st[0], "ql(A) \\ b")
            return x[0]
        col, lam, branch = self._source
        x, st = _b.ql_toeplitz_solve(col, np.array([lam]), bb, int(n), branch_rank=branch)
        _raise_status(st[0], "ql(A) \\ b")
        return x[0]

    def _q2(self):
        m = self.X.shape[1]
        v = self.X[0, m - 1]
        H1 = np.diag([1 - self.tau[0], 1])
        vv = np.array([v, 1])
        H2 = np.eye(2) - self.tau[1] * np.outer(vv, np.conj(vv))
        return H2 @ H1

    def apply_Q(self, b, n=1000, adjoint=False):
        """First n entries of `F.Q * b` (LowerHessenbergQ apply with its early exit, src/banded/hessenbergq.jl:111-123) or
        `F.Q' * b` (:125-135) for a pure Toeplitz operator — on the device."""
        bb = np.atleast_1d(np.asarray(b.head if isinstance(b, Vcat) else b))
        if self._source is None or (isinstance(self._source[0], str) and self._source[0] == "pert"):
            raise NotImplementedError("Q*b on the device needs a pure Toeplitz operator")
        col, lam, branch = self._source
        y, nstop, st = _b.ql_toeplitz_applyq(col, np.array([lam]), bb, int(n), adjoint=adjoint, branch_rank=branch)
        _raise_status(st[0], "ql(A).Q * b")
        return y[0]

    def _Q(self, i, j):
        pure = self._source is not None and not (isinstance(self._source[0], str) and self._source[0] == "pert")
        if pure and j <= 16 and self._head is None:
            e = np.zeros(j)
            e[j - 1] = 1.0
            return self.apply_Q(e, n=max(i, j + 1))[i - 1]         # Q[i, j] = (Q e_j)[i], on the device
        # perturbed operators (head blocks differ from the tail block): entry read-out from the 2 x 2 blocks
        n = max(i, j) + 2
        q = self._q2()
        Q = np.eye(n, dtype=q.dtype)
        for k in range(n - 1):
            Q[k:k + 2, :] = q @ Q[k:k + 2, :]
        return Q[i - 1, j - 1]


class _BlockQL:
    """Result of ql on a block-tridiagonal perturbed Toeplitz operator: the reference's
    QL(_BlockSkylineMatrix(Vcat(head data, Fill(tail column))), Vcat(F.τ, Fill(τ))) (src/infql.jl:199-204), laid out from the
    records the library returns (factored head, its τ, the fixed point P, τ tail).  `.factors[i, j]`, `.tau[k]`, `.L[i, j]`
    are index arithmetic on those records (no floating-point work here); `.Q[i, j]` = conj((Q'e_i)[j]) runs the library's
    block Q'b (lmul!(Q', b), src/infql.jl:213-243; getindex :245-246) for i <= 16."""

    def __init__(self, rec, args):
        self._rec, self._args = rec, args
        self.n, self.N = rec["n"], rec["N"]
        self._l11 = rec["L11"]
        self.L = _OneBased(self._L)
        self.Q = _OneBased(self._Q)
        self.factors = _OneBased(self._F)
        self.tau = _OneBased1(self._tau)
        self._qrows = {}

    def _P(self, I, J):
        n = self.n
        return self._rec["tail"][(I - 1) * n:I * n, (J - 1) * n:J * n]

    def _F(self, i, j):
        n, N = self.n, self.N
        hn = N * n
        I, J = (i - 1) // n + 1, (j - 1) // n + 1              # block indices
        r, c = (i - 1) % n, (j - 1) % n
        if i <= hn and j <= hn:
            return self._rec["head"][i - 1, j - 1]
        if J <= N:                                              # rows below the head (src/infql.jl:199-200)
            if I == N + 1 and J == N - 1:
                return self._P(2, 1)[r, c]
            if I == N + 1 and J == N:
                return self._P(2, 2)[r, c]
            if I == N + 2 and J == N:
                return self._P(2, 1)[r, c]
            return 0.0
        d = I - J                                               # Toeplitz tail column (:203)
        if d == -1:
            return self._P(1, 3)[r, c]
        if d == 0:
            return self._P(2, 3)[r, c]
        if d == 1:
            return self._P(2, 2)[r, c]
        if d == 2:
            return self._P(2, 1)[r, c]
        return 0.0

    def _tau(self, k):
        hn = self.N * self.n
        return self._rec["head_tau"][k - 1] if k <= hn else self._rec["tau_tail"][(k - 1) % self.n]

    def _L(self, i, j):
        return self._F(i, j) if i >= j else 0.0

    def _Q(self, i, j):
        if i > 16:
            raise IndexError("Q[i, j] is served by the library's Q'b for i <= 16")
        if i not in self._qrows:
            e = np.zeros(i)
            e[i - 1] = 1.0
            dlh, dh, duh, c, a, b, shift = self._args
            o = _b.ql_blocktri_factors(dlh, dh, duh, c, a, b, [shift], rhs=e, nout=i + 2 * self.n - 1)
            self._qrows[i] = np.conj(o["qtb"][0])
        row = self._qrows[i]
        return row[j - 1] if j <= row.size else 0.0


class _OneBased1:
    def __init__(self, fn):
        self._fn = fn

    def __getitem__(self, k):
        if isinstance(k, slice):
            return np.array([self._fn(a) for a in range(k.start or 1, k.stop + 1)])
        return self._fn(k)


class BlockTridiagonal:
    """BlockTridiagonal(Vcat(dl_head, Fill(c)), Vcat(d_head, Fill(a)), Vcat(du_head, Fill(b))) — each argument a
    `(head_list, tail_block)` pair (examples/periodicschrodinger.jl:8-10)."""

    def __init__(self, dl, d, du, shift=0.0, overrides=None):
        self.dl, self.d, self.du, self.shift = dl, d, du, shift
        self.overrides = dict(overrides or {})

    def __sub__(self, other):
        if isinstance(other, UniformScaling):
            return BlockTridiagonal(self.dl, self.d, self.du, self.shift + other.lam, self.overrides)
        return NotImplemented

    def __setitem__(self, ij, v):      # A[1,1] = 2 (1-based scalar perturbation inside the first diagonal block)
        self.overrides[ij] = v

    def _heads(self):
        dt = complex if any(np.iscomplexobj(np.asarray(m)) for m in list(self.d[0]) + [self.d[1]]) else float
        d_head = [np.array(m, dtype=dt, copy=True) for m in self.d[0]]
        for (i, j), v in self.overrides.items():
            n = np.asarray(self.d[1]).shape[0]
            if i > n or j > n:
                raise IndexError("only perturbations of the first diagonal block are supported")
            if not d_head:
                d_head = [np.array(self.d[1], dtype=dt, copy=True)]
            d_head[0][i - 1, j - 1] = v
        return list(self.dl[0]), d_head, list(self.du[0])


def ql(A, branch=1):
    """ql(A; branch) for ∞ Toeplitz / perturbed Toeplitz (u == 1) and block-tridiagonal perturbed Toeplitz operators."""
    if isinstance(A, BlockTridiagonal):
        dlh, dh, duh = A._heads()
        o = _b.ql_blocktri_factors(dlh, dh, duh, A.dl[1], A.d[1], A.du[1], [A.shift])
        _raise_status(o["status"][0], "ql(BlockTridiagonal)")
        rec = {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in o.items() if k != "qtb"}
        rec["L11"] = complex(rec["L11"]) if np.iscomplexobj(o["L11"]) else float(rec["L11"])
        return _BlockQL(rec, (dlh, dh, duh, A.dl[1], A.d[1], A.du[1], A.shift))
    l, u = A.bandwidths
    if u != 1:
        raise NotImplementedError("only u == 1 (ql_pruneband is undefined upstream)")
    col = A.tail_column()
    lam = A.shift
    cplx = np.iscomplexobj(col) or np.iscomplexobj(np.asarray(lam))
    hp = A.head_width()
    if hp == 0:
        L, st, tail = _b.ql_toeplitz_l11(col, np.array([lam]), branch_rank=branch, want_tail=True,
                                         real_path=not cplx)
        _raise_status(st[0], "ql")
        m = l + 2
        t = tail[0]
        return QLHessenberg(l, np.stack([t[:m], t[m:2 * m]]), t[2 * m:2 * m + 2], source=(col, lam, branch))
    p = A.head_rows()
    head = A.finite(p) + lam * np.eye(p)             # the C entry applies the shift itself
    L, st, hf = _b.ql_perttoeplitz_l11(head, col, np.array([lam]), branch_rank=branch, want_head_factors=True,
                                       real_path=not (cplx or np.iscomplexobj(head)))
    _raise_status(st[0], "ql")
    return QLHessenberg(l, None, None, head_factors=hf[0], head_l11=L[0], source=("pert", head, col, lam, branch))


def ql_L11(A, lambdas, branch=1, ngpus=1):
    """Batched form over a vector of shifts: L11[i] = ql(A - lambdas[i]*I).L[1,1]; NaN + status where it does not exist."""
    lam = np.asarray(lambdas) + A.shift
    if isinstance(A, BlockTridiagonal):
        dlh, dh, duh = A._heads()
        L, it, st = _b.ql_blocktri_l11(dlh, dh, duh, A.dl[1], A.d[1], A.du[1], lam, ngpus=ngpus)
        return L, st
    col = A.tail_column()
    if A.head_width() == 0:
        return _b.ql_toeplitz_l11(col, lam, branch_rank=branch, ngpus=ngpus)
    p = A.head_rows()
    head = A.finite(p) + A.shift * np.eye(p)
    return _b.ql_perttoeplitz_l11(head, col, lam, branch_rank=branch, ngpus=ngpus)


# ---- qr / cholesky ------------------------------------------------------------------------------------------

class _Solved(np.ndarray):
    """Solution prefix of an ∞ vector (zero beyond its length), 0-based numpy indexing."""


class AdaptiveQR:
    def __init__(self, A: BandedMatrix):
        self.A = A

    def solve(self, b, tolerance=_b.FLOATMIN, ncap=None):
        p, q, r, head = self.A.generators()
        l, u = self.A.bandwidths
        bb = np.atleast_1d(np.asarray(b.head if isinstance(b, Vcat) else b))
        ncap = int(ncap) if ncap else 400_000
        cplx = any(np.iscomplexobj(np.asarray(v)) for v in (p, q, bb, self.A.shift) + (() if head is None else (head,)))
        solver = _b.qr_banded_solve_c64 if cplx else _b.qr_banded_solve      # complex eltype (e.g. A - λI, λ complex)
        if not cplx:
            bb = bb.astype(float)
        out = solver(l, u, p[None], q[None], r[None], shift=[self.A.shift] if self.A.shift else None,
                     head=head, b=bb[None], tol=tolerance, ncap=ncap)
        _raise_status(out["status"][0], "qr(A) \\ b")
        return out["x"][0]

    __truediv__ = None

    def ldiv(self, b, **kw):
        return self.solve(b, **kw)


def qr(A: BandedMatrix):
    """qr(A) for an ∞ banded operator → adaptive QR (src/infqr.jl:163-174); `.solve(b)` is `F \\ b`."""
    return AdaptiveQR(A)


def ldiv(A, b, **kw):
    """`A \\ b` = `qr(A) \\ b` (factorize, src/infqr.jl:173-174,377)."""
    return qr(A).solve(b, **kw)


class AdaptiveCholesky:
    def __init__(self, S):
        self.A = S.parent if isinstance(S, Symmetric) else S

    def solve(self, b, ncap=None):
        p, q, r, head = self.A.generators()
        l, u = self.A.bandwidths
        sel = slice(0, u + 1)                         # upper bands only
        bb = np.atleast_1d(np.asarray(b.head if isinstance(b, Vcat) else b, dtype=float))
        out = _b.chol_banded_solve(u, p[None, sel], q[None, sel], r[None, sel],
                                   shift=[self.A.shift] if self.A.shift else None,
                                   head=None if head is None else head[sel], b=bb[None], ncap=int(ncap) if ncap else 400_000)
        _raise_status(out["status"][0], "cholesky(S) \\ b")
        return out["x"][0]


def cholesky(S):
    """cholesky(S) for a symmetric positive definite ∞ banded operator → adaptive Cholesky (src/infcholesky.jl:71-75)."""
    return AdaptiveCholesky(S)


# ---- reversecholesky (symmetric tridiagonal) ------------------------------------------------------------------

def _gen5(v):
    if isinstance(v, Rational):
        return list(v.c)
    if isinstance(v, Affine):
        return [0.0, v.p, v.q, v.r, 0.0] if (v.q != 0.0 or v.r != 1.0) else [v.p, 0.0, 0.0, 1.0, 0.0]
    raise TypeError("SymTridiagonal entries must be Fill / Affine / Rational, optionally behind a Vcat head")


class SymTridiagonal:
    """SymTridiagonal(dv, ev) with lazy ∞ diagonals (test/test_infreversecholesky.jl); `A + c*I` / `A - c*I` shift it."""

    def __init__(self, dv, ev, shift=0.0):
        self.dv, self.ev, self.shift = dv, ev, shift

    def __sub__(self, other):
        return SymTridiagonal(self.dv, self.ev, self.shift + other.lam)

    def __add__(self, other):
        return SymTridiagonal(self.dv, self.ev, self.shift - other.lam)

    __radd__ = __add__


class ReverseCholesky:
    """reversecholesky(A): A = U*U' = L'*L with L lower bidiagonal (src/banded/infreversecholeskytridiagonal.jl:66-74).
    `.L[i, j]` / `.U[i, j]` (1-based) expand the factor adaptively like the reference's getindex (:53-64)."""

    def __init__(self, A):
        self.A = A
        self._n = 0
        self._la = self._lb = None
        self.L = _OneBased(self._L)
        self.U = _OneBased(lambda i, j: self._L(j, i))
        self._expand(16)

    def _expand(self, n):
        if n <= self._n:
            return
        dv, ev = self.A.dv, self.A.ev
        ah = list(dv.head) if isinstance(dv, Vcat) else []
        bh = list(ev.head) if isinstance(ev, Vcat) else []
        dt, et = (dv.tail if isinstance(dv, Vcat) else dv), (ev.tail if isinstance(ev, Vcat) else ev)
        ga, gb = _gen5(dt), _gen5(et)
        hp = max(len(ah), len(bh))
        val = lambda g, k: g[0] + (g[1] + g[2] * k) / (g[3] + g[4] * k)
        # the lazy tails count positions from the end of their own head (Vcat), the kernel from 0: shift the generators
        def shifted(g, off):
            c0, p, q, r, s = g
            return [c0, p - q * off, q, r - s * off, s]
        a_head = [ah[k] if k < len(ah) else val(shifted(ga, len(ah)), k) for k in range(hp)] if hp else None
        b_head = [bh[k] if k < len(bh) else val(shifted(gb, len(bh)), k) for k in range(hp)] if hp else None
        out = _b.revchol_tridiag(shifted(ga, len(ah)), shifted(gb, len(bh)), int(n), shift=[self.A.shift] if self.A.shift else None,
                                 a_head=a_head, b_head=b_head)
        _raise_status(out["status"][0], "reversecholesky(A)")
        self._la, self._lb, self._n = out["la"][0], out["lb"][0], int(n)

    def _L(self, i, j):
        if j > i or i - j > 1:
            return 0.0
        self._expand(max(i, j) if max(i, j) <= self._n else 2 * max(i, j))
        return self._la[i - 1] if i == j else self._lb[j - 1]


def reversecholesky(A: SymTridiagonal):
    return ReverseCholesky(A)


# ---- adaptive finite-section QL (operators without a Toeplitz tail) -------------------------------------------

class FiniteSectionQL:
    """ql(A, tol) for an ∞ banded operator by adaptive finite sections (src/infql.jl:385-462): the leading j x j part of
    the factors; `.L[i, j]` (1-based), `.tau[k]`."""

    def __init__(self, out, l, u):
        self._Lband, self._V, self.tau, self.N = out["Lband"][0], out["V"][0], out["tau"][0], int(out["N"][0])
        self.l, self.u = l, u
        self.L = _OneBased(self._L)

    def _L(self, i, j):
        c = i - j
        if c < 0 or c > self.l + self.u:
            return 0.0
        if i > self._Lband.shape[0]:
            raise IndexError("only the leading block is materialised; pass a larger j to ql_adaptive")
        return self._Lband[i - 1, c]

    def Qt_mul(self, b):
        """(Q' * b)[1:len] for a finitely supported b (src/infql.jl:497-528: reflectors k = last(colsupport(b))+u down to 1,
        each touching rows max(1,k-u)..k).  Host glue over the leading block's packed reflectors."""
        bb = np.array(b.head if isinstance(b, Vcat) else b, dtype=float)
        u, nz = self.u, len(bb)
        top = nz + u
        if top > self._V.shape[0]:
            raise IndexError("right-hand side reaches beyond the materialised block; pass a larger j to ql_adaptive")
        B = np.zeros(top)
        B[:nz] = bb
        for k in range(top, 0, -1):                      # 1-based column k; V[k-1, r-1] = F[k-r, k]
            lo = max(1, k - u)
            vB = B[k - 1]
            for i in range(lo, k):
                vB = vB + self._V[k - 1, k - i - 1] * B[i - 1]
            vB = self.tau[k - 1] * vB
            B[k - 1] = B[k - 1] - vB
            for i in range(lo, k):
                B[i - 1] = B[i - 1] - self._V[k - 1, k - i - 1] * vB
        return B


def ql_adaptive(A: BandedMatrix, tol=np.finfo(float).eps, j=50, maxN=10000):
    """ql(A, tol) (src/infql.jl:385) for a banded operator whose diagonals are Fill / Affine / Rational (+ Vcat heads)."""
    l, u = A.bandwidths
    gens, heads = [], []
    for d in range(u, -l - 1, -1):
        v = A.diags.get(d, Fill(0.0))
        h = list(v.head) if isinstance(v, Vcat) else []
        t = v.tail if isinstance(v, Vcat) else v
        c0, p, q, r, s = _gen5(t)
        off = len(h)
        gens.append([c0, p - q * off, q, r - s * off, s])      # the lazy tail counts from the end of its own head
        heads.append(h)
    hp = max(len(h) for h in heads)
    head = None
    if hp:
        val = lambda g, k: g[0] + (g[1] + g[2] * k) / (g[3] + g[4] * k)
        head = np.array([[heads[rr][k] if k < len(heads[rr]) else val(gens[rr], k) for k in range(hp)] for rr in range(l + u + 1)])
    out = _b.ql_finite_section(l, u, gens, j=j, shift=[A.shift] if A.shift else None, head=head, tol=tol, maxN=maxN)
    if int(out["status"][0]) == 3:
        raise NotConvergedError("Reached max. iterations in adaptive QL without convergence to desired tolerance.")
    return FiniteSectionQL(out, l, u)


# ---- tridiagonal / bidiagonal conjugation --------------------------------------------------------------------

def _lazy_values(v, m):
    """First m entries of a lazy ∞ vector (Fill / Affine / Rational, optionally behind a Vcat head) or of an array."""
    if isinstance(v, Vcat):
        h = np.asarray(v.head, dtype=float)
        return np.concatenate([h, _lazy_values(v.tail, max(m - len(h), 0))])[:m]
    k = np.arange(m, dtype=float)
    if isinstance(v, Rational):
        c0, p, q, r, s = v.c
        return c0 + (p + q * k) / (r + s * k)
    if isinstance(v, Affine):
        return (v.p + v.q * k) / v.r + 0.0 * k
    a = np.asarray(v, dtype=float)
    if a.shape[0] < m:
        raise ValueError("explicit diagonal shorter than the requested size")
    return a[:m]


class Tridiagonal:
    """Tridiagonal(dl, d, du) with lazy ∞ diagonals (LazyBandedMatrices.Tridiagonal in test_bidiagonalconjugation.jl)."""

    def __init__(self, dl, d, du):
        self.dl, self.d, self.du = dl, d, du

    def bands(self, m):
        return np.stack([_lazy_values(self.dl, m), _lazy_values(self.d, m), _lazy_values(self.du, m)])


class TridiagonalConjugation:
    """TridiagonalConjugation(R, X): the ∞ tridiagonal matrix Tridiagonal(R*X/R), filled lazily like the reference's
    TridiagonalConjugationData (src/banded/tridiagonalconjugation.jl:224-303).  R: an upper triangular BandedMatrix
    (bandwidths (0, 1) or (0, 2); wider upper bands do not enter the tridiagonal part) or `cholesky(S)` of a symmetric
    positive definite banded operator (the factor then stays on the device).  Indexing `Y[i, j]` is 1-based."""

    def __init__(self, R, X: Tridiagonal):
        self.R, self.X = R, X
        self._n = 0
        self._Y = None

    def resizedata(self, n):
        if n <= self._n:
            return self
        n = max(int(n), 2 * self._n, 100)
        Xb = self.X.bands(n + 1)
        if isinstance(self.R, AdaptiveCholesky):
            A = self.R.A
            p, q, r, head = A.generators()
            _, u = A.bandwidths
            sel = slice(0, u + 1)
            out = _b.chol_tridiagonal_conjugation(u, p[None, sel], q[None, sel], Xb, n, r=r[None, sel],
                                                  shift=[A.shift] if A.shift else None, head=None if head is None else head[sel])
            _raise_status(out["status"][0], "TridiagonalConjugation(cholesky(S).U, X)")
            self._Y = out["Y"][0]
        else:
            Rb = self.R
            U = np.stack([_lazy_values(Rb._diag(d) if d in Rb.diags else Fill(0.0), n + 1) for d in (0, 1, 2)])
            if Rb.shift:
                U[0] -= Rb.shift
            self._Y = _b.tridiagonal_conjugation(U, Xb, n)[0]
        self._n = n
        return self

    def __getitem__(self, ij):
        i, j = ij
        if abs(i - j) > 1:
            return 0.0
        self.resizedata(max(i, j) + 1)
        return self._Y[1 + (j - i), min(i, j) - 1]

    def bands(self, n):
        """(dl, d, du)[:, :n]."""
        self.resizedata(n)
        return self._Y[:, :n]


def SymTridiagonalConjugation(R, X):
    """Same data, read as a symmetric tridiagonal matrix (diagonal and super-diagonal; tridiagonalconjugation.jl:300-303)."""
    return TridiagonalConjugation(R, X)


def BidiagonalConjugation(U_bands, C_bands, n, uplo="U"):
    """BidiagonalConjugation(U, X, V, uplo) from the tridiagonal parts of U and C = X*V (3 x m arrays: sub-, main,
    super-diagonal): returns (dv, ev) of Bidiagonal(inv(U) X V) (src/banded/bidiagonalconjugation.jl:19-60)."""
    dv, ev = _b.bidiagonal_conjugation(U_bands, C_bands, n, uplo)
    return dv[0], ev[0]


# ---- block-banded operators (KronTrav) -------------------------------------------------------------------------

class KronTravSum:
    """alpha*KronTrav(T1, Eye) + beta*KronTrav(Eye, T2) - lam*I with T1, T2 ∞ tridiagonal (`Tridiagonal` of lazy diagonals):
    the block-banded operators of test/test_infqr.jl:270-313 (block sizes 1, 2, 3, ...).  `A - c*I` / `A + B` compose."""

    def __init__(self, T1, T2, alpha=0.0, beta=0.0, lam=0.0):
        self.T1, self.T2, self.alpha, self.beta, self.lam = T1, T2, alpha, beta, lam

    def __sub__(self, other):
        if isinstance(other, UniformScaling):
            return KronTravSum(self.T1, self.T2, self.alpha, self.beta, self.lam + other.lam)
        return NotImplemented

    def __add__(self, other):
        if isinstance(other, KronTravSum):
            T1 = self.T1 if self.alpha != 0 else other.T1
            T2 = self.T2 if self.beta != 0 else other.T2
            return KronTravSum(T1, T2, self.alpha + other.alpha, self.beta + other.beta, self.lam + other.lam)
        if isinstance(other, UniformScaling):
            return KronTravSum(self.T1, self.T2, self.alpha, self.beta, self.lam - other.lam)
        return NotImplemented

    def __rmul__(self, s):
        return KronTravSum(self.T1, self.T2, s * self.alpha, s * self.beta, s * self.lam)

    def solve(self, b, tolerance=1e-30, maxblocks=64):
        """`A \\ b` = qr(A) \\ b by the adaptive block-banded QR (src/infqr.jl:303-337, 358-366)."""
        bb = np.atleast_1d(np.asarray(b.head if isinstance(b, Vcat) else b, dtype=float))
        m = maxblocks + 2
        out = _b.qr_blockbanded_solve(self.T1.bands(m), self.T2.bands(m), self.alpha, self.beta, self.lam, b=bb,
                                      tol=tolerance, maxblocks=maxblocks)
        _raise_status(out["status"][0], "qr(A) \\ b (block-banded)")
        return out["x"][0][: int(out["ncols"][0])]


def KronTrav(A, B):
    """KronTrav(T, Eye) or KronTrav(Eye, T) (LazyBandedMatrices; test_infqr.jl:271, 287): one factor must be `Eye`."""
    if B is Eye and A is not Eye:
        return KronTravSum(A, A, alpha=1.0)
    if A is Eye and B is not Eye:
        return KronTravSum(B, B, beta=1.0)
    raise TypeError("KronTrav: exactly one factor must be Eye")


class _EyeType:
    pass


Eye = _EyeType()


# ---- almost-banded operators: Vcat(boundary rows, BandedMatrix) ------------------------------------------------

class Poly:
    """Entry t (t = 0, 1, ...) = g0 + g1*t + g2*t^2 — `1:∞` = Poly(1, 1), `(1:∞).*(2:∞)` = Poly(2, 3, 1), `4:2:∞` = Poly(4, 2)."""

    def __init__(self, g0=0.0, g1=0.0, g2=0.0):
        self.g = (g0, g1, g2)


class BoundaryRow:
    """Dense ∞ row s^j (a + b j + c j^2): `Ones(1,∞)` = BoundaryRow(), `((-1).^(0:∞))'` = BoundaryRow(s=-1)."""

    def __init__(self, s=1.0, a=1.0, b=0.0, c=0.0):
        self.g = (s, a, b, c)


class AlmostBanded:
    """Vcat(rows..., D): dense boundary rows on top of an ∞ banded operator given by `{d: Poly | number}` with the entry of
    diagonal d counted along the ROWS of D (test/test_infqr.jl:193-267).  `A - c*I` is not defined upstream for this layout; a
    shifted family is spelled through `shift`, subtracted from D's main diagonal."""

    def __init__(self, rows, diags: dict, shift=0.0):
        self.rows, self.diags, self.shift = rows, diags, shift

    def solve(self, b, tolerance=_b.FLOATMIN, ncap=6000):
        """`A \\ b` = qr(A) \\ b (src/infqr.jl:15-21, 50-66, 236-274, 350-355)."""
        lD, uD = max(0, -min(self.diags)), max(0, max(self.diags))
        dg = []
        for d in range(uD, -lD - 1, -1):
            v = self.diags.get(d, 0.0)
            dg.append(list(v.g) if isinstance(v, Poly) else [float(v), 0.0, 0.0])
        bb = np.atleast_1d(np.asarray(b.head if isinstance(b, Vcat) else b, dtype=float))
        out = _b.qr_almostbanded_solve(lD, uD, dg, [list(r.g) for r in self.rows], bb, shift=[self.shift] if self.shift else None,
                                       tol=tolerance, ncap=ncap)
        _raise_status(out["status"][0], "qr(A) \\ b (almost-banded)")
        return out["x"][0][: int(out["ncols"][0])]
