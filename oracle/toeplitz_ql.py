"""CPU oracle for the Toeplitz-tail QL (TEST INFRASTRUCTURE — never imported by the product path).

numpy restatement of the reference's `tail_de`, `ql_X!` and `ql_hessenberg(::InfToeplitz)`
(reference: src/banded/infqltoeplitz.jl:7-65) plus the perturbed-Toeplitz head
`ql_hessenberg!` (src/infql.jl:60-93).  The eigen-decomposition goes through
`numpy.linalg.eig`, i.e. LAPACK `?geev` — the same routine family Julia's `eigen`
(`geevx`) reaches — and mirrors Julia's default `sortby` ordering of the eigenvalues
(lexicographic on (real, imag)) before `branch` is applied.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this.

Parity status: pinned by the reference's known answers for this path
(test/test_reduceql.jl:41-45 constants, test/test_infql.jl:12-67 reconstruction identities,
test/test_infql.jl:147-153 `(Q'b)[1]`) — see tests/test_oracle_toeplitz.py.
The *sign* of the real-typed path follows LAPACK's eigenvector sign (infqltoeplitz.jl:14-15): the reference returns
either (Q, L) or (-Q, -L) depending on which sign `dgeev` hands back for V[:, j].  `sign_rule="lapack"` (default)
restates that literally (numpy's LAPACK standing in for Julia's).  `sign_rule="householder"` picks, of the two, the
factorization for which `ql_X!` needs no sign flip — which is the limit of the finite-section Householder QL, i.e. the
L the reference's own tests assert (`L[1:10,1:10] ≈ ql(A[1:n,1:n]).L[1:10,1:10]`, test/test_infql.jl:14-17); the two
rules agree in every case those tests hold, and tests/test_oracle_toeplitz.py checks the householder rule against
dense finite sections on random real symbols.  The C ABI's `_f64` entries are specified with the householder rule.
"""
from __future__ import annotations

import numpy as np

STATUS_OK = 0
STATUS_NO_QL = 1          # DomainError: |mu|^2 < |a_u|^2 (infqltoeplitz.jl:13,23)
STATUS_NOT_REAL = 2       # DomainError: dominant eigenvalue not real (infqltoeplitz.jl:12)


def _csign(z):
    """Julia's sign(z) for complex z: z/|z| (0 for 0)."""
    z = np.asarray(z)
    az = np.abs(z)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(az == 0, 0, z / np.where(az == 0, 1, az))


def companion(a):
    """C of infqltoeplitz.jl:9/20 for a batch of coefficient rows a[..., m].

    First column (a[m-1], ..., a[1]) (1-based), then -a[m]*I_{m-2} stacked over a zero row.
    """
    a = np.asarray(a)
    m = a.shape[-1]
    n = m - 1
    C = np.zeros(a.shape[:-1] + (n, n), dtype=a.dtype)
    C[..., :, 0] = a[..., m - 2::-1]
    for i in range(n - 1):
        C[..., i, i + 1] = -a[..., m - 1]
    return C


def _julia_sorted_eig(C):
    """eigen(C) with Julia's default ordering: eigenvalues sorted by (real, imag)."""
    lam, V = np.linalg.eig(C)
    # lexsort: last key is primary
    order = np.lexsort((lam.imag, lam.real), axis=-1) if lam.ndim == 1 else \
        np.stack([np.lexsort((li.imag, li.real)) for li in lam.reshape(-1, lam.shape[-1])]).reshape(lam.shape)
    lam = np.take_along_axis(lam, order, axis=-1)
    V = np.take_along_axis(V, order[..., None, :], axis=-1)
    return lam, V


def _select(n2, branch_rank):
    """branch: rank 1 = findmax (first maximum); rank k = k-th largest |mu|^2
    (examples/toeplitz.jl:22-27: sortperm(n2)[end-k+1])."""
    if branch_rank == 1:
        return np.argmax(n2, axis=-1)
    order = np.argsort(n2, axis=-1, kind="stable")
    return order[..., -branch_rank]


def tail_de(a, branch_rank=1, real_path=None):
    """Batched `tail_de` (infqltoeplitz.jl:7-27).

    a: (..., m) coefficients [a_{-l}, ..., a_0, a_{+1}] (i.e. reverse of the Toeplitz data column).
    Returns (de (..., m-1), status (...), mu (...)).  Rows with status != 0 hold NaN.
    real_path: True → the `T<:Real` method; default: a is real dtype.
    """
    a = np.asarray(a)
    if real_path is None:
        real_path = not np.iscomplexobj(a)
    batch_shape = a.shape[:-1]
    m = a.shape[-1]
    C = companion(a)
    lam, V = _julia_sorted_eig(C)
    n2 = lam.real ** 2 + lam.imag ** 2
    j = _select(n2, branch_rank)
    n2j = np.take_along_axis(n2, j[..., None], axis=-1)[..., 0]
    muj = np.take_along_axis(lam, j[..., None], axis=-1)[..., 0]
    Vj = np.take_along_axis(V, j[..., None, None], axis=-1)[..., 0]       # (..., m-1)
    au = a[..., m - 1]
    status = np.zeros(batch_shape, dtype=np.int32)
    if real_path:
        notreal = np.iscomplexobj(lam) and (muj.imag != 0)
        if np.iscomplexobj(lam):
            status = np.where(muj.imag != 0, STATUS_NOT_REAL, status)
        au2 = np.real(au) ** 2
        status = np.where((status == 0) & ~(n2j >= au2), STATUS_NO_QL, status)
        with np.errstate(invalid="ignore", divide="ignore"):
            c = np.sqrt((n2j - au2) / np.real(Vj[..., 0]) ** 2)
            de = c[..., None] * np.real(Vj[..., ::-1])
        de = np.where((status == 0)[..., None], de, np.nan)
        return de, status.astype(np.int32), muj
    au2 = au.real ** 2 + au.imag ** 2 if np.iscomplexobj(au) else au ** 2
    status = np.where(~(n2j >= au2), STATUS_NO_QL, status)
    with np.errstate(invalid="ignore", divide="ignore"):
        v1 = Vj[..., 0]
        c_abs = np.sqrt((n2j - au2) / (v1.real ** 2 + v1.imag ** 2))
        c_sgn = -_csign(muj) / _csign(v1 * a[..., m - 2] - Vj[..., 1] * au)
        de = (c_sgn * c_abs)[..., None] * Vj[..., ::-1]
    de = np.where((status == 0)[..., None], de, np.nan)
    return de, status.astype(np.int32), muj


def _reflector(x):
    """LinearAlgebra.reflector! on the leading axis of x (shape (n, ...)); returns tau, mutates x.

    xi1 = x[0]; normu = ||x||; nu = copysign(normu, real(xi1)); xi1 += nu; x[0] = -nu; x[1:] /= xi1;
    tau = xi1/nu  (0 when normu == 0).
    """
    xi1 = x[0].copy()
    normu = np.sqrt(np.sum(x.real ** 2 + x.imag ** 2, axis=0)) if np.iscomplexobj(x) \
        else np.sqrt(np.sum(x * x, axis=0))
    zero = normu == 0
    nu = np.copysign(normu, xi1.real)
    nu_safe = np.where(zero, 1, nu)
    xi1n = xi1 + nu
    xi1n_safe = np.where(zero, 1, xi1n)
    x[0] = np.where(zero, x[0], -nu)
    if x.shape[0] > 1:
        x[1:] = np.where(zero, x[1:], x[1:] / xi1n_safe)
    tau = np.where(zero, 0, xi1n / nu_safe)
    return tau


def _reflector_apply(x, tau, A):
    """LinearAlgebra.reflectorApply!(x, tau, A): A (n, ncols, ...) <- (I - conj(tau) v v')A, v=[1;x[1:]]."""
    vA = A[0].copy()
    for i in range(1, x.shape[0]):
        vA = vA + np.conj(x[i])[None] * A[i]
    vA = np.conj(tau)[None] * vA
    A[0] = A[0] - vA
    for i in range(1, x.shape[0]):
        A[i] = A[i] - x[i][None] * vA


def ql_X(X, return_flip=False):
    """`ql_X!` (infqltoeplitz.jl:31-39) on a batch X[..., 2, m]; returns (X_factored, tau[..., 2]) (+ the mask of entries
    whose sign normalisation (:34-37) fired, with return_flip).

    `ql!` of a 2×m matrix (MatrixFactorizations; the loop is restated in src/infql.jl:10-22 and
    examples/jacobi.jl:61-74): k = 2: reflector on (X[2,m], X[1,m]) applied to columns 1..m-1;
    k = 1 (complex only): length-1 reflector on X[1,m-1] applied to X[1,1:m-2]; then the sign fix.
    """
    X = np.array(X, copy=True)
    cplx = np.iscomplexobj(X)
    m = X.shape[-1]
    s = np.sign(X[..., 1, m - 1].real)
    tau = np.zeros(X.shape[:-2] + (2,), dtype=X.dtype)
    # k = 2 : x = (X[2,m], X[1,m]); rows reversed (view(A, 2:-1:1, m))
    Xt = np.moveaxis(X, (-2, -1), (0, 1))        # (2, m, ...)
    x = np.stack([Xt[1, m - 1], Xt[0, m - 1]])
    t2 = _reflector(x)
    Xt[1, m - 1], Xt[0, m - 1] = x[0], x[1]
    A = np.stack([Xt[1, :m - 1], Xt[0, :m - 1]])   # (2, m-1, ...)
    _reflector_apply(x, t2, A)
    Xt[1, :m - 1], Xt[0, :m - 1] = A[0], A[1]
    tau[..., 1] = t2
    if cplx:
        x1 = Xt[0, m - 2][None].copy()
        t1 = _reflector(x1)
        Xt[0, m - 2] = x1[0]
        if m > 2:
            A1 = Xt[0, :m - 2][None].copy()
            _reflector_apply(x1, t1, A1)
            Xt[0, :m - 2] = A1[0]
        tau[..., 0] = t1
    X = np.moveaxis(Xt, (0, 1), (-2, -1))
    flip = s != np.sign(X[..., 0, m - 2].real)
    tau[..., 0] = np.where(flip, 2 - tau[..., 0], tau[..., 0])
    X[..., 0, :m - 1] = np.where(flip[..., None], -X[..., 0, :m - 1], X[..., 0, :m - 1])
    if return_flip:
        return X, tau, flip
    return X, tau


def ql_toeplitz_tail(a, branch_rank=1, real_path=None, sign_rule="lapack"):
    """`ql_hessenberg(::InfToeplitz)` data (infqltoeplitz.jl:56-65) for a batch of symbols.

    sign_rule (real-typed path only, see the module docstring): "lapack" = the eigenvector sign LAPACK returns,
    "householder" = the sign for which ql_X! does not flip (= the finite-section Householder QL limit).

    a: (..., m) = reverse(A.data.args[1]) = [a_{-l}, ..., a_0, a_{+1}].
    Returns dict with L11 (...), X (..., 2, m) factored, tau (..., 2), de, status, mu.
      first L column  = [X[1,m-1]; X[2,m-1:-1:1]]   (band rows u+1 ..)
      tail  L column  = X[2,m:-1:1]
      2×2 Q block     = QLPackedQ(X[:, m-1:m], tau)
    """
    a = np.asarray(a)
    de, status, mu = tail_de(a, branch_rank, real_path)
    m = a.shape[-1]
    dt = de.dtype if np.iscomplexobj(de) or np.iscomplexobj(a) else np.float64
    dt = np.result_type(a.dtype, de.dtype)
    X = np.zeros(a.shape[:-1] + (2, m), dtype=dt)
    X[..., 0, :] = a
    X[..., 1, 1:] = de
    if real_path is None:
        real_path = not np.iscomplexobj(a)
    if sign_rule not in ("lapack", "householder"):
        raise ValueError(sign_rule)
    with np.errstate(invalid="ignore", divide="ignore"):
        Xf, tau, flip = ql_X(X, return_flip=True)
        if real_path and sign_rule == "householder" and np.any(flip):
            # V -> -V: de -> -de; row 2 of X changes sign, s changes sign, and the normalisation no longer fires
            de = np.where(np.asarray(flip)[..., None], -de, de)
            X[..., 1, 1:] = de
            Xf, tau, flip2 = ql_X(X, return_flip=True)
            assert not np.any(flip2 & (status == 0))
    L11 = Xf[..., 0, m - 2]
    return dict(L11=L11, X=Xf, tau=tau, de=de, status=status, mu=mu)


def ql_toeplitz_l11(coeffs_data_column, lambdas, branch_rank=1, real_path=None, sign_rule="lapack"):
    """L[1,1] of ql(A - lambda*I) for the pure Toeplitz A whose BandedMatrices data column is
    `coeffs_data_column` = [a_{+u}, ..., a_0, ..., a_{-l}] with u == 1 (infqltoeplitz.jl:56-65, 191).

    lambdas: array of shifts.  Returns (L11, status).
    """
    col = np.asarray(coeffs_data_column)
    lambdas = np.asarray(lambdas)
    dt = np.result_type(col.dtype, lambdas.dtype)
    a = np.empty(lambdas.shape + (col.shape[0],), dtype=dt)
    a[...] = col[::-1]
    m = col.shape[0]
    a[..., m - 2] = col[1] - lambdas          # diagonal coefficient a_0 - lambda
    out = ql_toeplitz_tail(a, branch_rank, real_path, sign_rule)
    return out["L11"], out["status"]


# ----------------------------------------------------------------------------------------------
# helpers used by the tests (dense finite sections; restated from src/infql.jl:10-22)
# ----------------------------------------------------------------------------------------------

def dense_ql(A):
    """Householder QL of a dense square/tall matrix, loop of src/infql.jl:10-22 with
    the `(T<:Real)` lower bound of examples/jacobi.jl:63 (the convention MatrixFactorizations.ql! uses).

    Returns (factors, tau): L in the lower triangle (for m == n), reflector vectors above.
    """
    A = np.array(A, copy=True)
    cplx = np.iscomplexobj(A)
    m, n = A.shape
    mn = min(m, n)
    tau = np.zeros(mn, dtype=A.dtype)
    kmin = 1 if cplx else 2
    for k in range(mn, kmin - 1, -1):
        mu = m - mn + k
        nu = n - mn + k
        x = A[mu - 1::-1, nu - 1].copy() if mu - 1 >= 0 else None
        x = A[:mu, nu - 1][::-1].copy()
        t = _reflector(x)
        A[:mu, nu - 1] = x[::-1]
        tau[k - 1] = t
        if nu - 1 > 0:
            B = A[:mu, :nu - 1][::-1].copy()
            _reflector_apply(x, np.asarray(t), B)
            A[:mu, :nu - 1] = B[::-1]
    return A, tau


def ql_packed_Q(factors, tau):
    """Dense Q of a packed QL (square factors): Q = H_n ... H_1?  Built by applying to the identity.

    Q = prod_k H_k with H_k = I - tau_k v_k v_k', v_k = [factors[0:k-1, nu]; 1; 0...]. A = Q L.
    """
    m, n = factors.shape
    mn = min(m, n)
    Q = np.eye(m, dtype=factors.dtype)
    # A = Q L  with  Q' = H_1' ... H_mn'  applied in order k = mn, ..., 1 to A; so Q = H_mn H_{mn-1} ... H_1
    for k in range(1, mn + 1):
        mu = m - mn + k
        nu = n - mn + k
        v = np.zeros(m, dtype=factors.dtype)
        v[:mu - 1] = factors[:mu - 1, nu - 1]
        v[mu - 1] = 1
        H = np.eye(m, dtype=factors.dtype) - tau[k - 1] * np.outer(v, np.conj(v))
        Q = H @ Q
    return Q


def ql_toeplitz_solve(coeffs_data_column, lambdas, b, nout, branch_rank=1, real_path=None):
    """First `nout` entries of  x = ql(A - λI) \\ b  for the pure Toeplitz A (u == 1) and a finitely supported b.

    Reference: `ldiv!(F, b) = ldiv!(F.L, lmul!(F.Q', b))` (src/infql.jl:266-267) with
      Q  = LowerHessenbergQ(Fill(q, ∞)), q = QLPackedQ(X, τ) (2×2)      (infqltoeplitz.jl:63-64, hessenbergq.jl:86-135):
           Q' = UpperHessenbergQ(adjoint.(q)) applied for n = nz, nz-1, ..., 1 to b[n:n+1];
      L  = lower-triangular band: first column [X[1,m-1]; X[2,m-1:-1:1]], every other column X[2,m:-1:1]
           (infqltoeplitz.jl:63), solved by the column-oriented forward substitution of src/infql.jl:269-289.
    Returns (x [shape lambdas.shape + (nout,)], status).  Pinned by test/test_infql.jl:143-145 (vs qr(A) \\ b).
    """
    col = np.asarray(coeffs_data_column)
    lambdas = np.asarray(lambdas)
    b = np.asarray(b)
    dt = np.result_type(col.dtype, lambdas.dtype, b.dtype, np.float64)
    m = col.shape[0]
    flat = lambdas.reshape(-1)
    a = np.empty((flat.size, m), dtype=np.result_type(col.dtype, lambdas.dtype))
    a[...] = col[::-1]
    a[:, m - 2] = col[1] - flat
    out = ql_toeplitz_tail(a, branch_rank, real_path)
    X, tau, status = out["X"], out["tau"], out["status"]
    dt = np.result_type(dt, X.dtype)
    nb = b.shape[0]
    x = np.full((flat.size, nout), np.nan, dtype=dt)
    for i in range(flat.size):
        if status[i] != 0:
            continue
        q = ql_packed_Q(X[i][:, m - 2:m], tau[i])
        qh = q.conj().T
        t = np.zeros(max(nb + 1, nout), dtype=dt)
        t[:nb] = b
        for n in range(nb, 0, -1):                     # UpperHessenbergQ: n = nz:-1:1
            t[n - 1:n + 1] = qh @ t[n - 1:n + 1]
        t = t[:nout].copy() if t.size >= nout else np.concatenate([t, np.zeros(nout - t.size, dtype=dt)])
        sub = X[i][1, m - 2::-1]                       # L[j+i, j] = X[2, m-i], i = 1..m-1
        for j in range(nout):
            d = X[i][0, m - 2] if j == 0 else X[i][1, m - 1]
            t[j] = t[j] / d
            hi = min(nout, j + m)
            t[j + 1:hi] -= sub[:hi - j - 1] * t[j]
        x[i] = t
    return x.reshape(lambdas.shape + (nout,)), status.reshape(lambdas.shape)


def hessenberg_q_apply(q_blocks, b, nout, adjoint=False):
    """`Q*b` for Q = LowerHessenbergQ(q) (materialize!, src/banded/hessenbergq.jl:111-123: n = 1, 2, ...: x[n:n+1] = q[n] x[n:n+1],
    early exit once n > nz and norm(t) <= 10 floatmin) or, with adjoint=True, `Q'*b` = UpperHessenbergQ(adjoint.(q)) applied for
    n = nz, ..., 1 (:125-135).  q_blocks(n) returns the n-th 2x2 block (1-based).  Returns (first nout entries, n at exit or 0)."""
    b = np.asarray(b)
    dt = np.result_type(b.dtype, q_blocks(1).dtype, np.float64)
    nz = b.shape[0]
    x = np.zeros(max(nout, nz) + 2, dtype=dt)
    x[:nz] = b
    if adjoint:
        for n in range(nz, 0, -1):
            x[n - 1:n + 1] = q_blocks(n).conj().T @ x[n - 1:n + 1]
        return x[:nout].copy(), 0
    stop = 0
    tiny = 10 * np.finfo(np.float64).tiny
    for n in range(1, nout + 1):
        t = q_blocks(n) @ x[n - 1:n + 1]
        x[n - 1:n + 1] = t
        # LinearAlgebra.norm (generic_norm2) rescales by the largest entry when the squares would underflow
        if n > nz and np.hypot(np.hypot(t[0].real, t[0].imag), np.hypot(t[1].real, t[1].imag)) <= tiny:
            stop = n
            break
    return x[:nout].copy(), stop


def ql_toeplitz_applyq(coeffs_data_column, lambdas, b, nout, adjoint=False, branch_rank=1, real_path=None):
    """Q b / Q'b for the ∞ Hessenberg Q of ql(A - λI), every block equal to q = QLPackedQ(X[:, m-1:m], τ)
    (infqltoeplitz.jl:63-64).  Returns (y [λ-shape + (nout,)], nstop, status)."""
    col = np.asarray(coeffs_data_column)
    lambdas = np.asarray(lambdas)
    m = col.shape[0]
    flat = lambdas.reshape(-1)
    a = np.empty((flat.size, m), dtype=np.result_type(col.dtype, lambdas.dtype))
    a[...] = col[::-1]
    a[:, m - 2] = col[1] - flat
    out = ql_toeplitz_tail(a, branch_rank, real_path, sign_rule="householder")
    X, tau, status = out["X"], out["tau"], out["status"]
    dt = np.result_type(X.dtype, np.asarray(b).dtype)
    y = np.full((flat.size, nout), np.nan, dtype=dt)
    ns = np.zeros(flat.size, dtype=np.int64)
    for i in range(flat.size):
        if status[i] != 0:
            continue
        q = ql_packed_Q(X[i][:, m - 2:m], tau[i])
        y[i], ns[i] = hessenberg_q_apply(lambda n: q, b, nout, adjoint)
    return y.reshape(lambdas.shape + (nout,)), ns.reshape(lambdas.shape), status.reshape(lambdas.shape)


# ----------------------------------------------------------------------------------------------
# perturbed Toeplitz head: ql_hessenberg!(B)  (src/infql.jl:60-93), u == 1
# ----------------------------------------------------------------------------------------------

def _toeplitz_entry(a_rev, i, j):
    """T[i,j] (1-based) of the Toeplitz tail with reversed coefficient vector a_rev = [a_{-l},...,a_0,a_{+1}]."""
    m = len(a_rev)
    k = (j - i) + m - 2
    return a_rev[k] if 0 <= k < m else 0.0


def ql_pert_toeplitz(head, a_rev, real_path=None, branch_rank=1):
    """`ql_hessenberg!` restated for a perturbed Toeplitz operator with bandwidths (l, 1).

    head : p x p top-left corner of A (dense), p >= l + 1; row p of A must already be pure Toeplitz
    a_rev: tail symbol [a_{-l}, ..., a_0, a_{+1}] (shift already applied to head and a_rev by the caller)
    Returns dict(L11, head_factors (p x p, L in the lower triangle), head_tau, tail=..., status).
    """
    head = np.array(head, copy=True)
    a_rev = np.asarray(a_rev)
    if real_path is None:
        real_path = not (np.iscomplexobj(head) or np.iscomplexobj(a_rev))
    dt = np.float64 if real_path else np.complex128
    head = head.astype(dt)
    a_rev = a_rev.astype(dt)
    p = head.shape[0]
    m = len(a_rev)
    lp = m - 1                                   # l' = l + u: lower bandwidth after widening (infql.jl:60)
    tail = ql_toeplitz_tail(a_rev, branch_rank, real_path)
    if tail["status"] != 0:
        return dict(L11=np.nan, status=int(tail["status"]), tail=tail)
    X, tau = tail["X"], tail["tau"]
    # 2x2 block q = QLPackedQ(X[:, m-1:m], tau) = H_2 H_1  (infqltoeplitz.jl:64)
    q = ql_packed_Q(X[:, m - 2:m].astype(dt), tau.astype(dt))
    # (Q∞')[1, j] = conj(q11 q21^(j-1))   (LowerHessenbergQ with every block equal to q, hessenbergq.jl:111-135,162)
    qrow = np.array([np.conj(q[0, 0] * q[1, 0] ** (j - 1)) for j in range(1, lp)], dtype=dt)
    Bt = head.copy()
    Bt[p - 1, p - 1] = tail["L11"]
    if lp > 1:
        Tblk = np.array([[_toeplitz_entry(a_rev, lp - 1 + ii, jj) for jj in range(1, lp)] for ii in range(1, lp)], dtype=dt)
        Bt[p - 1, p - lp:p - 1] = qrow @ Tblk     # infql.jl:82
    F, ftau = dense_ql(Bt)                        # infql.jl:83 (banded ql!: zeros outside the band stay zero)
    return dict(L11=F[0, 0], head_factors=F, head_tau=ftau, tail=tail, status=0)


def _hessenberg_q_dense(blocks, K):
    """Leading rows of Q = ... q_3 q_2 q_1 (LowerHessenbergQ, hessenbergq.jl:111-123): dense (K+2) x (K+2), rows < K+1 final."""
    dt = np.result_type(*[b.dtype for b in blocks])
    Q = np.eye(K + 2, dtype=dt)
    for n in range(K + 1):
        Q[n:n + 2, :] = blocks[n] @ Q[n:n + 2, :]
    return Q


def ql_pert_toeplitz_solve(head, a_rev, b, nout, real_path=None, branch_rank=1):
    """First `nout` entries of x = ql(A) \\ b for the perturbed Toeplitz A of `ql_pert_toeplitz` (u == 1).

    Reference: ql_hessenberg! (src/infql.jl:70-93) assembles
      L: columns 1..p from the factored head B~ (p x p) extended by the l' rows
         B~[p+1:p+l', p-l'+1:p] = (Q∞')[2:l'+1, 1:l'+1] T[l':2l', 1:l']            (:87)
         and the Toeplitz tail columns [X[2,m:-1:1]] after them                    (:91-92);
      Q: Vcat(LowerHessenbergQ(F.Q).q, F∞.q) — p-1 head blocks (hessenbergq.jl:182-190) then the constant tail block;
    and F \\ b = ldiv!(F.L, lmul!(F.Q', b)) (src/infql.jl:266-289).  Pinned by test/test_infql.jl:147-160
    ((Q'b)[1] ≈ 0.9892996329463546, (A*(F\\b))[1:10] ≈ e1).  Returns (x, status, qtb) with qtb = Q'b (leading part).
    """
    res = ql_pert_toeplitz(head, a_rev, real_path, branch_rank)
    if res["status"] != 0:
        return np.full(nout, np.nan), res["status"], None
    F, ftau, tail = res["head_factors"], res["head_tau"], res["tail"]
    dt = F.dtype
    a_rev = np.asarray(a_rev).astype(dt)
    p = F.shape[0]
    m = len(a_rev)
    lp = m - 1
    X, tau = tail["X"].astype(dt), tail["tau"].astype(dt)
    qinf = ql_packed_Q(X[:, m - 2:m], tau)
    # head blocks: q_1 = QLPackedQ(F[1:2,1:2], tau[1:2]); q_j = QLPackedQ(F[j:j+1,j:j+1], [0, tau[j+1]]), j = 2..p-1
    hq = []
    if p >= 2:
        hq.append(ql_packed_Q(F[0:2, 0:2], ftau[0:2]))
        for j in range(2, p):
            hq.append(ql_packed_Q(F[j - 1:j + 1, j - 1:j + 1], np.array([0, ftau[j]], dtype=dt)))
    blk = lambda n: hq[n - 1] if n <= p - 1 else qinf                    # 1-based block index
    b = np.asarray(b, dtype=dt)
    nb = b.shape[0]
    size = max(nb + 1, nout, p + lp) + lp + 1
    t = np.zeros(size, dtype=dt)
    t[:nb] = b
    for n in range(nb, 0, -1):
        t[n - 1:n + 1] = blk(n).conj().T @ t[n - 1:n + 1]
    qtb = t[:nb + 1].copy()
    # extension block E (l' x l'): rows p+1..p+l', columns p-l'+1..p
    Qd = _hessenberg_q_dense([qinf] * (lp + 2), lp + 1)
    Qadj = Qd.conj().T
    Tblk = np.array([[_toeplitz_entry(a_rev, lp + ii, 1 + jj) for jj in range(lp)] for ii in range(lp + 1)], dtype=dt)
    E = Qadj[1:lp + 1, 0:lp + 1] @ Tblk
    sub = X[1, m - 2::-1]                                                 # tail column below its diagonal
    for j in range(nout):
        if j < p:
            t[j] = t[j] / F[j, j]
            t[j + 1:p] -= F[j + 1:p, j] * t[j]
            c = j - (p - lp)
            if c >= 0:
                t[p:p + lp] -= E[:, c] * t[j]
        else:
            t[j] = t[j] / X[1, m - 1]
            t[j + 1:j + 1 + lp] -= sub[:lp] * t[j]
    return t[:nout].copy(), 0, qtb


def pert_toeplitz_head(head_data, a_col, p=None):
    """Dense p x p top-left corner from BandedMatrices data (l+2) x p_head (rows: +1, 0, -1, ..., -l), continued by
    the Toeplitz column a_col when p exceeds the head width."""
    head_data = np.asarray(head_data)
    a_col = np.asarray(a_col)
    nrow, ph = head_data.shape
    if p is None:
        p = max(ph, nrow - 1)
    dt = np.result_type(head_data.dtype, a_col.dtype)
    H = np.zeros((p, p), dtype=dt)
    for j in range(p):
        col = head_data[:, j] if j < ph else a_col
        for r in range(nrow):
            i = j + r - 1          # data row r <-> diagonal d = 1 - r, row index i = j - d
            if 0 <= i < p:
                H[i, j] = col[r]
    return H
