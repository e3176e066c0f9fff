/* CPU oracle, type-generic body (TEST INFRASTRUCTURE — never linked into the product library).
 *
 * Included twice by ila_oracle.c: once with REAL = double (suffix _f64) and once with
 * REAL = __float128 (suffix _f128, the conditioning-aware reference of SURVEY §0.6).
 *
 * Every function cites the reference lines it restates; paths are relative to the
 * InfiniteLinearAlgebra.jl checkout.  Arithmetic order is the reference's (and the un-vendored
 * BandedMatrices / ArrayLayouts / LAPACK loops it calls, restated from their published
 * algorithms); compile with -ffp-contract=off so that no multiply-add is fused.
 */

#ifndef REAL
#error "define REAL, SQRT, FABS, COPYSIGN, SUF before including"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* ---------------------------------------------------------------------------------------------
 * Operator generator shared by the QR and Cholesky oracles (mirrors include/ila_b200.h):
 * diagonal d = col-row in [-l, u]; data row rr = u-d (BandedMatrices storage order);
 * position along the diagonal t0 = min(row, col) (0-based) as in `BandedMatrix(d => v)`;
 *   value = head[rr*hp + t0]              if t0 < hp
 *         = (p[rr] + q[rr]*t0) / r[rr]    otherwise       (r == NULL means 1)
 * and `shift` is subtracted on the main diagonal (A - shift*I).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int l, u;
    const double *p, *q, *r;
    double shift;
    int hp;
    const double *head;
} FN(op_t);

static inline REAL FN(op_entry)(const FN(op_t) * A, int64_t row, int64_t col)
{ /* 0-based row/col */
    int64_t d = col - row;
    if (d > A->u || d < -A->l) return (REAL)0;
    int rr = (int)(A->u - d);
    int64_t t0 = d >= 0 ? row : col;
    REAL v;
    if (t0 < A->hp) {
        v = (REAL)A->head[(int64_t)rr * A->hp + t0];
    } else {
        REAL den = A->r ? (REAL)A->r[rr] : (REAL)1;
        v = ((REAL)A->p[rr] + (REAL)A->q[rr] * (REAL)t0) / den;
    }
    if (d == 0) v = v - (REAL)A->shift;
    return v;
}

/* LinearAlgebra.norm for a short vector (generic_norm2): sqrt of the running sum of squares, with the
 * scaled branch when maxabs^2 under/overflows.  [dependency: LinearAlgebra stdlib, julia >= 1.10] */
static inline REAL FN(norm2)(const REAL *x, int n)
{
    REAL maxabs = 0;
    for (int i = 0; i < n; i++) {
        REAL a = FABS(x[i]);
        if (a > maxabs || a != a) maxabs = a;
    }
    if (maxabs == 0 || maxabs != maxabs || maxabs > (REAL)1.7e308) return maxabs;
    REAL mm = maxabs * maxabs;
    REAL chk = (REAL)n * maxabs * maxabs;
    if (chk <= (REAL)1.7e308 && mm != 0) {
        REAL s = x[0] * x[0];
        for (int i = 1; i < n; i++) s = s + x[i] * x[i];
        return SQRT(s);
    }
    REAL inv = (REAL)1 / maxabs;
    REAL t0 = FABS(x[0]) * inv;
    REAL s = t0 * t0;
    for (int i = 1; i < n; i++) {
        REAL t = FABS(x[i]) * inv;
        s = s + t * t;
    }
    return maxabs * SQRT(s);
}

/* reflector!(x) — SURVEY A.1; the same steps are spelled out at src/infql.jl:16-18. */
static inline REAL FN(reflector)(REAL *x, int n)
{
    REAL xi1 = x[0];
    REAL normu = FN(norm2)(x, n);
    if (normu == 0) return (REAL)0;
    REAL nu = COPYSIGN(normu, xi1);
    xi1 = xi1 + nu;
    x[0] = -nu;
    for (int i = 1; i < n; i++) x[i] = x[i] / xi1;
    return xi1 / nu;
}

/* reflectorApply!(x, tau, A[:, j]) on one column held in a[0..n-1] — src/infql.jl:19, SURVEY A.1. */
static inline void FN(reflector_apply)(const REAL *x, REAL tau, REAL *a, int n)
{
    REAL vA = a[0];
    for (int i = 1; i < n; i++) vA = vA + x[i] * a[i];
    vA = tau * vA;
    a[0] = a[0] - vA;
    for (int i = 1; i < n; i++) a[i] = a[i] - x[i] * vA;
}

/* ---------------------------------------------------------------------------------------------
 * Adaptive banded QR solve, one instance.
 *   reference: AdaptiveQRData banded ctor   src/infqr.jl:8-13  (cache bandwidths (l, l+u))
 *              partialqr!                    src/infqr.jl:32-48 (-> BandedMatrices._banded_qr!)
 *              Q'b adaptive driver           src/infqr.jl:236-274 (COLGROWTH chunks, tolerance test)
 *              resizedata_chop!              src/infqr.jl:207-219 (never shrinks unless all <= tol)
 *              ldiv!(R, B)                   src/infqr.jl:350-355 (tbsv 'U','N','N' over datasize)
 * Outputs: x[0..*n_out), n_out = B.datasize after the driver, nconv = first(jr) at the break,
 *          optional factors ((2l+u+1) x n_out column-major band storage, R above / v below),
 *          tau[n_out], qtb[n_out] (= Q'b before the back-substitution).
 * status: 0 ok, 3 ncap reached before convergence, 5 NaN met.
 * ------------------------------------------------------------------------------------------- */
static int FN(qr_banded_solve_one)(const FN(op_t) * A, int nb, const double *b, double tol, int colgrowth,
                                   int64_t ncap, REAL *x, int64_t *n_out, int64_t *nconv_out, REAL *factors,
                                   REAL *tau_out, REAL *qtb_out)
{
    const int l = A->l, u = A->u, U = l + u, ld = 2 * l + u + 1;
    REAL *W = (REAL *)calloc((size_t)ld * (size_t)(ncap + U + 1), sizeof(REAL));
    REAL *tau = (REAL *)calloc((size_t)ncap + 1, sizeof(REAL));
    REAL *B = (REAL *)calloc((size_t)ncap + l + 1, sizeof(REAL));
    int64_t nfill = 0, ncols = 0; /* columns filled from A, columns factorized */
    int status = 0;
    for (int i = 0; i < nb && i < ncap; i++) B[i] = (REAL)b[i];
#define WW(i, j) W[(size_t)(U + (i) - (j)) + (size_t)(j) * ld]

    /* partialqr!(F, n): factorize columns ncols..n-1 (0-based), reflectors hit the next U columns */
#define PARTIALQR(nreq)                                                                         \
    do {                                                                                        \
        int64_t n_ = (nreq);                                                                    \
        int64_t need_ = n_ + U;                                                                 \
        for (; nfill < need_; nfill++) {                                                        \
            int64_t c_ = nfill;                                                                 \
            int64_t lo_ = c_ - u < 0 ? 0 : c_ - u;                                              \
            for (int64_t i_ = lo_; i_ <= c_ + l; i_++) WW(i_, c_) = FN(op_entry)(A, i_, c_);    \
        }                                                                                       \
        for (; ncols < n_; ncols++) {                                                           \
            int64_t k_ = ncols;                                                                 \
            REAL *xk_ = &WW(k_, k_);                                                            \
            REAL t_ = l > 0 ? FN(reflector)(xk_, l + 1) : (REAL)0;                              \
            tau[k_] = t_;                                                                       \
            if (l > 0)                                                                          \
                for (int j_ = 1; j_ <= U; j_++) FN(reflector_apply)(xk_, t_, &WW(k_, k_ + j_), l + 1); \
        }                                                                                       \
    } while (0)

    int64_t n = nb; /* B.datasize */
    int64_t nconv = 0;
    if (l == 0) { /* infqr.jl:250-252 diagonal special case: Q = I */
        n = nb;
    } else {
        int64_t jfirst = 0, jlast = colgrowth - 1; /* jr, 0-based inclusive */
        for (;;) {
            int64_t cs_max = jlast + l;            /* last(colsupport(factors, last(jr))) */
            if (cs_max + 1 + U + 1 > ncap) { status = 3; break; }
            PARTIALQR(jlast + 1);                  /* colsupport(A.factors,last(jr)) -> partialqr! (infqr.jl:117-120,259) */
            if (cs_max + 1 > n) n = cs_max + 1;    /* resizedata!(B, cs_max) (infqr.jl:262) */
            if (jfirst + 1 > nb) {                 /* j > sB */
                REAL mx = 0;
                int isnan_ = 0;
                for (int64_t i = jfirst; i <= jfirst + l; i++) {
                    REAL a = FABS(B[i]);
                    if (a != a) isnan_ = 1;
                    if (a > mx) mx = a;
                }
                if (isnan_) { status = 5; nconv = jfirst + 1; break; }
                if (mx <= (REAL)tol) { nconv = jfirst + 1; break; }
            }
            PARTIALQR(cs_max + 1);                 /* infqr.jl:267 */
            for (int64_t k = jfirst; k <= jlast; k++) { /* lmul!(Q_N', view(Bdata,kr)) (infqr.jl:268-269) */
                REAL *xk = &WW(k, k);
                REAL vB = B[k];
                for (int i = 1; i <= l; i++) vB = vB + xk[i] * B[k + i];
                vB = tau[k] * vB;
                B[k] = B[k] - vB;
                for (int i = 1; i <= l; i++) B[k + i] = B[k + i] - xk[i] * vB;
            }
            jfirst = jlast + 1;
            jlast = jlast + colgrowth;
        }
        /* resizedata_chop! (infqr.jl:207-219): datasize kept unless every entry <= tol */
        int any = 0;
        for (int64_t k = n - 1; k >= 0; k--)
            if (FABS(B[k]) > (REAL)tol) { any = 1; break; }
        if (!any) n = 0;
    }
    if (status == 0 || status == 5) {
        /* ldiv!(R, B) (infqr.jl:350-355): partialqr!(n); tbsv 'U','N','N' bandwidth U on n x n */
        if (n + U + 1 > ncap) status = 3;
        else {
            PARTIALQR(n);
            if (qtb_out) for (int64_t k = 0; k < n; k++) qtb_out[k] = B[k];
            for (int64_t j = n - 1; j >= 0; j--) {
                REAL xj = B[j] / WW(j, j);
                B[j] = xj;
                int64_t lo = j - U < 0 ? 0 : j - U;
                for (int64_t i = lo; i < j; i++) B[i] = B[i] - xj * WW(i, j);
            }
        }
    }
    if (status == 3) n = 0;
    for (int64_t k = 0; k < n; k++) x[k] = B[k];
    if (factors)
        for (int64_t j = 0; j < n; j++)
            for (int r = 0; r < ld; r++) factors[(size_t)j * ld + r] = W[(size_t)j * ld + r];
    if (tau_out) for (int64_t k = 0; k < n; k++) tau_out[k] = tau[k];
    *n_out = n;
    *nconv_out = nconv;
    free(W); free(tau); free(B);
#undef WW
#undef PARTIALQR
    return status;
}

/* ---------------------------------------------------------------------------------------------
 * Adaptive banded Cholesky solve, one instance (S symmetric, upper bands 0..u given by the operator).
 *   reference: partialcholesky!              src/infcholesky.jl:42-59
 *                banded_chol! -> LAPACK pbtrf('U') -> unblocked pbtf2 for small bandwidth:
 *                ajj = sqrt(a_jj); row j scaled by (1/ajj) [dscal]; trailing kn x kn -= x x' [dsyr 'U']
 *              chunk hand-off                src/infcholesky.jl:51-55 (U1' \ B, next block -= B'B)
 *              adaptive L \ b                src/infcholesky.jl:94-133 (COLGROWTH = 1000, tol = floatmin;
 *                per chunk tbsv 'U','T','N': dot over the band, subtract, divide; coupling by muladd!)
 *              U \ b                         src/infcholesky.jl:135-140 (tbsv 'U','N','N')
 * status: 0 ok, 3 ncap, 4 not positive definite (the reference discards pbtrf's info and produces NaN;
 *         the oracle keeps going exactly as LAPACK would not — it stops and flags).
 * ------------------------------------------------------------------------------------------- */
static int FN(chol_banded_solve_one)(const FN(op_t) * A, int nb, const double *b, double tol, int colgrowth,
                                     int64_t ncap, REAL *x, int64_t *n_out, int64_t *nconv_out, REAL *factors,
                                     REAL *y_out)
{
    const int u = A->u, ld = u + 1;
    REAL *W = (REAL *)calloc((size_t)ld * (size_t)(ncap + 2 * u + 2), sizeof(REAL));
    REAL *B = (REAL *)calloc((size_t)ncap + 2 * u + 2, sizeof(REAL));
    int64_t nfill = 0, ncols = 0;
    int status = 0;
    for (int i = 0; i < nb && i < ncap; i++) B[i] = (REAL)b[i];
    /* upper band storage: W(i,j) for j-u <= i <= j at row u+i-j */
#define WW(i, j) W[(size_t)(u + (i) - (j)) + (size_t)(j) * ld]
#define PARTIALCHOL(nreq)                                                                        \
    do {                                                                                         \
        int64_t n_ = (nreq);                                                                     \
        if (n_ > ncols) {                                                                        \
            for (; nfill < n_ + u; nfill++) {                                                    \
                int64_t c_ = nfill;                                                              \
                int64_t lo_ = c_ - u < 0 ? 0 : c_ - u;                                           \
                for (int64_t i_ = lo_; i_ <= c_; i_++) WW(i_, c_) = FN(op_entry)(A, i_, c_);     \
            }                                                                                    \
            int64_t k0_ = ncols;                                                                 \
            for (int64_t j_ = k0_; j_ < n_ && status == 0; j_++) { /* pbtf2 'U' on block k0_..n_-1 */ \
                REAL ajj_ = WW(j_, j_);                                                          \
                if (!(ajj_ > 0)) { status = 4; break; }                                          \
                ajj_ = SQRT(ajj_);                                                               \
                WW(j_, j_) = ajj_;                                                               \
                int64_t kn_ = n_ - 1 - j_ < u ? n_ - 1 - j_ : u;                                 \
                if (kn_ > 0) {                                                                   \
                    REAL rinv_ = (REAL)1 / ajj_;                                                 \
                    for (int64_t c_ = j_ + 1; c_ <= j_ + kn_; c_++) WW(j_, c_) = rinv_ * WW(j_, c_); \
                    for (int64_t c_ = j_ + 1; c_ <= j_ + kn_; c_++) { /* dsyr upper, alpha = -1 */ \
                        REAL xc_ = WW(j_, c_);                                                   \
                        if (xc_ != 0) {                                                          \
                            REAL t_ = -xc_;                                                      \
                            for (int64_t i_ = j_ + 1; i_ <= c_; i_++)                            \
                                WW(i_, c_) = WW(i_, c_) + WW(j_, i_) * t_;                       \
                        }                                                                        \
                    }                                                                            \
                }                                                                                \
            }                                                                                    \
            if (status == 0 && u > 0) {                                                          \
                /* hand-off (infcholesky.jl:51-55): kr2 = max(n-u+1,kr[1]):n ; B = data[kr2, n+1:n+u] */ \
                int64_t r0_ = n_ - u > k0_ ? n_ - u : k0_;                                       \
                for (int64_t c_ = n_; c_ < n_ + u; c_++) {     /* ldiv!(U1', B): forward substitution */ \
                    for (int64_t i_ = r0_; i_ < n_; i_++) {                                      \
                        if (i_ < c_ - u) continue;                                               \
                        REAL s_ = WW(i_, c_);                                                    \
                        for (int64_t k_ = r0_; k_ < i_; k_++)                                    \
                            if (k_ >= c_ - u && k_ >= i_ - u) s_ = s_ - WW(k_, i_) * WW(k_, c_); \
                        WW(i_, c_) = s_ / WW(i_, i_);                                            \
                    }                                                                            \
                }                                                                                \
                for (int64_t c_ = n_; c_ < n_ + u; c_++)       /* next block -= B'B */           \
                    for (int64_t a_ = n_; a_ <= c_; a_++) {                                      \
                        REAL s_ = 0;                                                             \
                        int first_ = 1;                                                          \
                        for (int64_t i_ = r0_; i_ < n_; i_++) {                                  \
                            if (i_ < c_ - u || i_ < a_ - u) continue;                            \
                            REAL t_ = WW(i_, a_) * WW(i_, c_);                                   \
                            if (first_) { s_ = t_; first_ = 0; } else s_ = s_ + t_;              \
                        }                                                                        \
                        if (!first_) WW(a_, c_) = WW(a_, c_) - s_;                               \
                    }                                                                            \
            }                                                                                    \
            ncols = n_;                                                                          \
        }                                                                                        \
    } while (0)

    int64_t n = nb, nconv = 0;
    int64_t jfirst = 0, jlast = colgrowth - 1;
    for (;;) {
        int64_t cs_max = jlast + u;
        if (cs_max + 1 + 2 * u + 1 > ncap) { status = 3; break; }
        PARTIALCHOL(jlast + 1 + u); /* colsupport(A,last(jr)) -> partialcholesky!(last(jr)+u) (infcholesky.jl:81-84,116) */
        if (status) break;
        if (cs_max + 1 > n) n = cs_max + 1;
        if (jfirst + 1 > nb) {
            REAL mx = 0;
            for (int64_t i = jfirst; i <= jfirst + u; i++) {
                REAL a = FABS(B[i]);
                if (a > mx) mx = a;
            }
            if (mx <= (REAL)tol) { nconv = jfirst + 1; break; }
        }
        /* ldiv!(U_N', view(B.data, jr)): tbsv 'U','T','N' on the chunk (dot, subtract, divide) */
        for (int64_t i = jfirst; i <= jlast; i++) {
            int64_t lo = i - u < jfirst ? jfirst : i - u;
            if (lo < i) {
                REAL s = WW(lo, i) * B[lo];
                for (int64_t k = lo + 1; k < i; k++) s = s + WW(k, i) * B[k];
                B[i] = B[i] - s;
            }
            B[i] = B[i] / WW(i, i);
        }
        /* muladd!(-1, U[jr1,jr2]', b[jr1], 1, b[jr2]) (infcholesky.jl:126-128) */
        for (int64_t c = jlast + 1; c <= jlast + u; c++) {
            REAL s = 0;
            int first = 1;
            for (int64_t k = jlast - u + 1; k <= jlast; k++) {
                if (k < 0 || k < c - u) continue;
                REAL t = WW(k, c) * B[k];
                if (first) { s = t; first = 0; } else s = s + t;
            }
            if (!first) B[c] = B[c] - s;
        }
        jfirst = jlast + 1;
        jlast = jlast + colgrowth;
    }
    if (status == 0) {
        if (n + 2 * u + 1 > ncap) status = 3;
        else {
            PARTIALCHOL(n);
            if (status == 0) {
                if (y_out) for (int64_t k = 0; k < n; k++) y_out[k] = B[k];
                for (int64_t j = n - 1; j >= 0; j--) {
                    REAL xj = B[j] / WW(j, j);
                    B[j] = xj;
                    int64_t lo = j - u < 0 ? 0 : j - u;
                    for (int64_t i = lo; i < j; i++) B[i] = B[i] - xj * WW(i, j);
                }
            }
        }
    }
    if (status != 0) n = 0;
    for (int64_t k = 0; k < n; k++) x[k] = B[k];
    if (factors)
        for (int64_t j = 0; j < n; j++)
            for (int r = 0; r < ld; r++) factors[(size_t)j * ld + r] = W[(size_t)j * ld + r];
    *n_out = n;
    *nconv_out = nconv;
    free(W); free(B);
#undef WW
#undef PARTIALCHOL
    return status;
}

#undef CAT_
#undef CAT
#undef FN
