/* ila_oracle_c64.c — CPU oracle for the adaptive banded QR solve with a COMPLEX eltype.
 *
 * TEST INFRASTRUCTURE (same rules as ila_oracle.c: only tests/, smoke() and bench.py's CPU legs may load it).
 *
 * Restates, for ComplexF64, the same reference path as qr_banded_solve_one in ila_oracle_impl.h:
 *   AdaptiveQRData banded ctor   src/infqr.jl:8-13,  partialqr! src/infqr.jl:32-48 (-> BandedMatrices._banded_qr!),
 *   Q'b adaptive driver          src/infqr.jl:236-274, resizedata_chop! :207-219, ldiv!(R,B) :350-355,
 * with LinearAlgebra.reflector! / reflectorApply! for complex vectors (xi1 += copysign(norm, real(xi1)), x[2:end] ./= xi1,
 * tau = xi1/nu; vA = A[1,j] + sum conj(x_i) A[i,j]; vA = conj(tau) vA) — the same statements src/infql.jl:17-19 uses.
 *
 * Complex arithmetic is spelled out (no _Complex operators, no libgcc __muldc3/__divdc3): products are the four-multiply
 * form, quotients Smith's algorithm, nothing is fused (-ffp-contract=off).  The CUDA kernels issue the same operations in
 * the same order, so the comparison in tests/test_gpu_qr_c64.py is bit for bit.
 *
 * Parity status: the reference has NO complex adaptive-QR test (test/test_infqr.jl is real throughout); the restatement is
 * pinned by its real-valued limit (imaginary parts zero reproduce ila_oracle_qr_banded_solve_f64 bit for bit) and by
 * dense solves of finite sections — tests/test_oracle_qr_c64.py.  "parity unpinned" against reference golden vectors.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } c64;

static inline c64 mkc(double re, double im) { c64 z = {re, im}; return z; }
static inline c64 cadd(c64 a, c64 b) { return mkc(a.re + b.re, a.im + b.im); }
static inline c64 csub(c64 a, c64 b) { return mkc(a.re - b.re, a.im - b.im); }
static inline c64 cmul(c64 a, c64 b) { return mkc(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static inline c64 cconj(c64 a) { return mkc(a.re, -a.im); }
static inline double cabs2(c64 a) { return a.re * a.re + a.im * a.im; }
static inline double cabs1(c64 a) { return sqrt(cabs2(a)); }
static inline c64 cdiv(c64 a, c64 b)
{   /* Smith */
    if (fabs(b.re) >= fabs(b.im)) {
        double r = b.im / b.re, den = b.re + b.im * r;
        return mkc((a.re + a.im * r) / den, (a.im - a.re * r) / den);
    }
    double r = b.re / b.im, den = b.re * r + b.im;
    return mkc((a.re * r + a.im) / den, (a.im * r - a.re) / den);
}

typedef struct {
    int l, u;
    const c64 *p, *q;     /* l+u+1 generator coefficients, data-row order (diagonal +u first) */
    const double *r;      /* real denominators or NULL */
    c64 shift;
    int hp;
    const c64 *head;      /* (l+u+1) x hp explicit head, row-major */
} opc_t;

static inline c64 opc_entry(const opc_t *A, int64_t row, int64_t col)
{
    int64_t d = col - row;
    if (d > A->u || d < -A->l) return mkc(0, 0);
    int rr = A->u - (int)d;
    int64_t t0 = row < col ? row : col;
    c64 v;
    if (t0 < A->hp) v = A->head[(int64_t)rr * A->hp + t0];
    else {
        c64 qt = mkc(A->q[rr].re * (double)t0, A->q[rr].im * (double)t0);
        v = cadd(A->p[rr], qt);
        if (A->r && A->r[rr] != 1.0) v = mkc(v.re / A->r[rr], v.im / A->r[rr]);
    }
    if (d == 0) v = csub(v, A->shift);
    return v;
}

static inline c64 reflector_c(c64 *x, int n)
{
    double s = cabs2(x[0]);
    for (int i = 1; i < n; i++) s = s + cabs2(x[i]);
    double normu = sqrt(s);
    if (normu == 0) return mkc(0, 0);
    c64 xi1 = x[0];
    double nu = copysign(normu, xi1.re);
    xi1.re = xi1.re + nu;
    x[0] = mkc(-nu, 0);
    for (int i = 1; i < n; i++) x[i] = cdiv(x[i], xi1);
    return mkc(xi1.re / nu, xi1.im / nu);
}

static inline void reflector_apply_c(const c64 *x, c64 tau, c64 *a, int n)
{
    c64 vA = a[0];
    for (int i = 1; i < n; i++) vA = cadd(vA, cmul(cconj(x[i]), a[i]));
    vA = cmul(cconj(tau), vA);
    a[0] = csub(a[0], vA);
    for (int i = 1; i < n; i++) a[i] = csub(a[i], cmul(x[i], vA));
}

static int qr_banded_solve_one_c64(const opc_t *A, int nb, const c64 *b, double tol, int colgrowth, int64_t ncap, c64 *x,
                                   int64_t *n_out, int64_t *nconv_out)
{
    const int l = A->l, u = A->u, U = l + u, ld = 2 * l + u + 1;
    c64 *W = (c64 *)calloc((size_t)ld * (size_t)(ncap + U + 1), sizeof(c64));
    c64 *tau = (c64 *)calloc((size_t)ncap + 1, sizeof(c64));
    c64 *B = (c64 *)calloc((size_t)ncap + l + 1, sizeof(c64));
    int64_t nfill = 0, ncols = 0;
    int status = 0;
    for (int i = 0; i < nb && i < ncap; i++) B[i] = b[i];
#define WW(i, j) W[(size_t)(U + (i) - (j)) + (size_t)(j) * ld]
#define PARTIALQR(nreq)                                                                         \
    do {                                                                                        \
        int64_t n_ = (nreq);                                                                    \
        int64_t need_ = n_ + U;                                                                 \
        for (; nfill < need_; nfill++) {                                                        \
            int64_t c_ = nfill;                                                                 \
            int64_t lo_ = c_ - u < 0 ? 0 : c_ - u;                                              \
            for (int64_t i_ = lo_; i_ <= c_ + l; i_++) WW(i_, c_) = opc_entry(A, i_, c_);       \
        }                                                                                       \
        for (; ncols < n_; ncols++) {                                                           \
            int64_t k_ = ncols;                                                                 \
            c64 *xk_ = &WW(k_, k_);                                                             \
            c64 t_ = reflector_c(xk_, l + 1);                                                   \
            tau[k_] = t_;                                                                       \
            for (int j_ = 1; j_ <= U; j_++) reflector_apply_c(xk_, t_, &WW(k_, k_ + j_), l + 1); \
        }                                                                                       \
    } while (0)

    int64_t n = nb, nconv = 0;
    int64_t jfirst = 0, jlast = colgrowth - 1;
    for (;;) {
        int64_t cs_max = jlast + l;
        if (cs_max + 1 + U + 1 > ncap) { status = 3; break; }
        PARTIALQR(jlast + 1);
        if (cs_max + 1 > n) n = cs_max + 1;
        if (jfirst + 1 > nb) {
            double mx = 0;
            int isnan_ = 0;
            for (int64_t i = jfirst; i <= jfirst + l; i++) {
                double a = cabs1(B[i]);
                if (a != a) isnan_ = 1;
                if (a > mx) mx = a;
            }
            if (isnan_) { status = 5; nconv = jfirst + 1; break; }
            if (mx <= tol) { nconv = jfirst + 1; break; }
        }
        PARTIALQR(cs_max + 1);
        for (int64_t k = jfirst; k <= jlast; k++) {
            reflector_apply_c(&WW(k, k), tau[k], &B[k], l + 1);
        }
        jfirst = jlast + 1;
        jlast = jlast + colgrowth;
    }
    if (status == 0) {
        int any = 0;
        for (int64_t k = n - 1; k >= 0; k--)
            if (cabs1(B[k]) > tol) { any = 1; break; }
        if (!any) n = 0;
        if (n + U + 1 > ncap) status = 3;
        else {
            PARTIALQR(n);
            for (int64_t j = n - 1; j >= 0; j--) {
                c64 xj = cdiv(B[j], WW(j, j));
                B[j] = xj;
                int64_t lo = j - U < 0 ? 0 : j - U;
                for (int64_t i = lo; i < j; i++) B[i] = csub(B[i], cmul(xj, WW(i, j)));
            }
        }
    }
    if (status != 0) n = 0;
    for (int64_t k = 0; k < n; k++) x[k] = B[k];
    *n_out = n;
    *nconv_out = nconv;
    free(W); free(tau); free(B);
#undef WW
#undef PARTIALQR
    return status;
}

/* Batched entry.  p, q: batch x (l+u+1) complex; r: batch x (l+u+1) real or NULL; shift: batch complex or NULL;
 * head: (l+u+1) x hp complex shared; b: batch x nb complex; x: batch x ldx complex (zero padded). */
int ila_oracle_qr_banded_solve_c64(int l, int u, int64_t batch, const double *p, const double *q, const double *r,
                                   const double *shift, int hp, const double *head, int nb, const double *b, double tol,
                                   int colgrowth, int64_t ncap, double *x, int64_t ldx, int64_t *ncols, int64_t *nconv,
                                   int32_t *status, int nthreads)
{
    const int nd = l + u + 1;
    if (l < 1) return -4;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic)
    for (int64_t i = 0; i < batch; i++) {
        opc_t A;
        A.l = l; A.u = u;
        A.p = (const c64 *)p + i * nd;
        A.q = (const c64 *)q + i * nd;
        A.r = r ? r + i * nd : NULL;
        A.shift = shift ? ((const c64 *)shift)[i] : mkc(0, 0);
        A.hp = hp;
        A.head = (const c64 *)head;
        c64 *xi = (c64 *)x + i * ldx;
        memset(xi, 0, sizeof(c64) * (size_t)ldx);
        int64_t n = 0, nc = 0;
        c64 *tmp = (c64 *)calloc((size_t)ncap + 1, sizeof(c64));
        int st = qr_banded_solve_one_c64(&A, nb, (const c64 *)b + i * nb, tol, colgrowth, ncap, tmp, &n, &nc);
        for (int64_t k = 0; k < n && k < ldx; k++) xi[k] = tmp[k];
        free(tmp);
        ncols[i] = n;
        nconv[i] = nc;
        status[i] = st;
    }
    return 0;
}
