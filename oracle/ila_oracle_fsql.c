/* ila_oracle_fsql.c — CPU oracle for the adaptive finite-section QL of ∞ banded operators (real eltype).
 *
 * TEST INFRASTRUCTURE (same rules as ila_oracle.c).
 *
 * Restates the "experimental adaptive finite section QL" of src/infql.jl:364-462 (InfiniteLinearAlgebra.jl v0.10.3):
 *   ql(A::InfBandedMatrix, tol = eps)                 :385-388
 *   initialadaptiveQLblock / cache_filldata!          :391-411, 440-458
 *       j = block size (50 initially), checkinds = max(1, j - l - u)
 *       Ls = ql(A[checkinds:N, checkinds:N]).L[2:j-checkinds+1, 2:j-checkinds+1], N = j
 *       while Lerr > tol:  Ll = same block from A[checkinds:2N, checkinds:2N];  Lerr = norm(Ll - Ls, 2) (Frobenius);
 *                          N >= maxN -> error("Reached max. iterations ...");  Ls = Ll;  N = 2N
 *       F = ql(A[1:N÷2, 1:N÷2]);  factors[1:j,1:j], τ[1:j]
 *   the finite banded ql! is MatrixFactorizations' Householder QL (restated at src/infql.jl:10-22; real eltype stops at k = 2).
 * The QL of a banded section runs from its last column to its first; only rows k-u..k of columns k-l-u..k are live, so
 * the sweep keeps that (u+1) x (l+u+1) window and nothing of size N: this file and the CUDA kernel
 * (csrc/ql_finite_section.cu) do the same operations in the same order (IEEE, unfused) and agree bit for bit.
 *
 * Operator entries: diagonal d = col - row in [-l, u] has data row rr = u - d; value(t) = c0 + (p + q t)/(r + s t) with
 * t = min(row, col) (0-based) from the 5 coefficients g[rr*5 .. rr*5+4]; an optional explicit head replaces t < hp;
 * `shift` is subtracted from the main diagonal.
 * Outputs per operator: Lband[k*(l+u+1) + c] = L[k, k-c] (0-based row k < j, c = 0..l+u; 0 outside the matrix),
 * V[k*u + (r-1)] = reflector entry of column k at row k-r (r = 1..u), tau[k], N_used (size of the final section).
 * status: 0 ok, 3 "Reached max. iterations in adaptive QL without convergence to desired tolerance".
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXB 8   /* l + u + 1 <= 8, u + 1 <= 8 */

typedef struct {
    int l, u, hp;
    const double *g;      /* (l+u+1) x 5 */
    const double *head;   /* (l+u+1) x hp or NULL */
    double shift;
} fsop_t;

static inline double fs_entry(const fsop_t *A, int64_t row, int64_t col)
{
    const int64_t d = col - row;
    if (d > A->u || d < -A->l || row < 0 || col < 0) return 0.0;
    const int rr = A->u - (int)d;
    const int64_t t = row < col ? row : col;
    double v;
    if (t < A->hp) v = A->head[(int64_t)rr * A->hp + t];
    else {
        const double *g = A->g + rr * 5;
        v = g[0] + (g[1] + g[2] * (double)t) / (g[3] + g[4] * (double)t);
    }
    if (d == 0) v = v - A->shift;
    return v;
}

/* Householder QL sweep of the section A[lo..hi, lo..hi] (0-based inclusive) from column hi down to column kstop.
 * For every processed column k <= krec the row k of L, the reflector and tau are handed to `rec`. */
typedef void (*fs_rec_fn)(void *ctx, int64_t k, const double *Lrow /* l+u+1 */, const double *v /* u */, double tau);

static void fs_sweep(const fsop_t *A, int64_t lo, int64_t hi, int64_t kstop, int64_t krec, fs_rec_fn rec, void *ctx)
{
    const int l = A->l, u = A->u, NC = l + u + 1, NR = u + 1;
    double W[MAXB][MAXB];   /* W[c][r] = current A[k-r, k-c] */
    int64_t k = hi;
    for (int c = 0; c < NC; c++)
        for (int r = 0; r < NR; r++) {
            const int64_t row = k - r, col = k - c;
            W[c][r] = (row >= lo && col >= lo) ? fs_entry(A, row, col) : 0.0;
        }
    for (; k >= kstop; k--) {
        double x[MAXB], tau = 0.0;
        for (int r = 0; r < NR; r++) x[r] = W[0][r];
        if (k > lo) {   /* real eltype: the QL loop stops at the section's second column (examples/jacobi.jl:63 bound) */
            double s = x[0] * x[0];
            for (int r = 1; r < NR; r++) s = s + x[r] * x[r];
            const double normu = sqrt(s);
            if (normu != 0.0) {
                double xi1 = x[0];
                const double nu = copysign(normu, xi1);
                xi1 = xi1 + nu;
                x[0] = -nu;
                for (int r = 1; r < NR; r++) x[r] = x[r] / xi1;
                tau = xi1 / nu;
            }
            for (int r = 0; r < NR; r++) W[0][r] = x[r];
            for (int c = 1; c < NC; c++) {
                double vA = W[c][0];
                for (int r = 1; r < NR; r++) vA = vA + x[r] * W[c][r];
                vA = tau * vA;
                W[c][0] = W[c][0] - vA;
                for (int r = 1; r < NR; r++) W[c][r] = W[c][r] - x[r] * vA;
            }
        }
        if (k <= krec) {
            double Lrow[MAXB];
            for (int c = 0; c < NC; c++) Lrow[c] = W[c][0];
            rec(ctx, k, Lrow, x + 1, tau);
        }
        /* slide to column k-1 */
        for (int c = 0; c + 1 < NC; c++)
            for (int r = 0; r + 1 < NR; r++) W[c][r] = W[c + 1][r + 1];
        for (int r = 0; r + 1 < NR; r++) W[NC - 1][r] = 0.0;
        for (int c = 0; c < NC; c++) {
            const int64_t row = k - 1 - u, col = k - 1 - c;
            W[c][u] = (row >= lo && col >= lo) ? fs_entry(A, row, col) : 0.0;
        }
    }
}

typedef struct { int64_t base; int m, NC; double blk[MAXB][MAXB]; } blk_ctx;
static void rec_block(void *ctx_, int64_t k, const double *Lrow, const double *v, double tau)
{
    (void)v; (void)tau;
    blk_ctx *c = (blk_ctx *)ctx_;
    const int64_t i = k - c->base;          /* block row 0..m-1 */
    if (i < 0 || i >= c->m) return;
    for (int j = 0; j < c->m; j++) {
        const int64_t cc = i - j;           /* L[k, k-cc] */
        c->blk[i][j] = (cc >= 0 && cc < c->NC) ? Lrow[cc] : 0.0;
    }
}

typedef struct { int64_t j; int NC, u; double *Lband, *V, *tau; } out_ctx;
static void rec_out(void *ctx_, int64_t k, const double *Lrow, const double *v, double tau)
{
    out_ctx *c = (out_ctx *)ctx_;
    if (k >= c->j) return;
    for (int cc = 0; cc < c->NC; cc++) c->Lband[k * c->NC + cc] = Lrow[cc];
    for (int r = 0; r < c->u; r++) c->V[k * c->u + r] = v[r];
    c->tau[k] = tau;
}

static int fsql_one(const fsop_t *A, int64_t j, double tol, int64_t maxN, double *Lband, double *V, double *tau, int64_t *N_used)
{
    const int l = A->l, u = A->u, NC = l + u + 1;
    const int64_t check1 = (j - l - u > 1) ? j - l - u : 1;   /* 1-based checkinds */
    const int64_t lo = check1 - 1;                             /* 0-based first row of the test sections */
    const int m = (int)(j - check1);                           /* block rows: section rows 2..j-checkinds+1 */
    blk_ctx Ls, Ll;
    Ls.base = Ll.base = lo + 1; Ls.m = Ll.m = m; Ls.NC = Ll.NC = NC;
    memset(Ls.blk, 0, sizeof(Ls.blk)); memset(Ll.blk, 0, sizeof(Ll.blk));
    int64_t N = j;
    fs_sweep(A, lo, N - 1, lo + 1, j - 1, rec_block, &Ls);
    double Lerr = 1.0;
    int status = 0;
    while (Lerr > tol) {
        fs_sweep(A, lo, 2 * N - 1, lo + 1, j - 1, rec_block, &Ll);
        double s = 0.0;
        for (int cj = 0; cj < m; cj++)          /* column-major, like norm(::Matrix) */
            for (int ri = 0; ri < m; ri++) {
                const double dlt = Ll.blk[ri][cj] - Ls.blk[ri][cj];
                s = s + dlt * dlt;
            }
        Lerr = sqrt(s);
        if (N >= maxN) { status = 3; break; }
        memcpy(Ls.blk, Ll.blk, sizeof(Ls.blk));
        N = 2 * N;
    }
    if (status == 0) {
        out_ctx oc = {j, NC, u, Lband, V, tau};
        fs_sweep(A, 0, N / 2 - 1, 0, j - 1, rec_out, &oc);
        *N_used = N / 2;
    } else {
        for (int64_t k = 0; k < j * NC; k++) Lband[k] = NAN;
        for (int64_t k = 0; k < j * u; k++) V[k] = NAN;
        for (int64_t k = 0; k < j; k++) tau[k] = NAN;
        *N_used = N;
    }
    return status;
}

int ila_oracle_ql_finite_section_f64(int l, int u, int64_t batch, const double *g, const double *shift, int hp,
                                     const double *head, int64_t j, double tol, int64_t maxN, double *Lband, double *V,
                                     double *tau, int64_t *N_used, int32_t *status, int nthreads)
{
    if (l < 0 || u < 1 || l + u + 1 > MAXB || j < l + u + 2) return -4;
    const int nd = l + u + 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic)
    for (int64_t i = 0; i < batch; i++) {
        fsop_t A = {l, u, hp, g + i * nd * 5, head, shift ? shift[i] : 0.0};
        status[i] = fsql_one(&A, j, tol, maxN, Lband + i * j * nd, V + i * j * u, tau + i * j, N_used + i);
    }
    return 0;
}

/* ================= complex eltype (test/test_infql.jl:257-262): the QL loop runs down to the section's first column,
 * inner products conjugate the reflector, tau is complex; generator coefficients c0, p, q complex, r, s real (the real
 * parts of the 4th and 5th coefficient are used). ================= */
typedef struct { double re, im; } c64;
static inline c64 zmk(double re, double im) { c64 z = {re, im}; return z; }
static inline c64 zadd_(c64 a, c64 b) { return zmk(a.re + b.re, a.im + b.im); }
static inline c64 zsub_(c64 a, c64 b) { return zmk(a.re - b.re, a.im - b.im); }
static inline c64 zmul_(c64 a, c64 b) { return zmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static inline c64 zconj_(c64 a) { return zmk(a.re, -a.im); }
static inline double zabs2_(c64 a) { return a.re * a.re + a.im * a.im; }
static inline c64 zdiv_(c64 a, c64 b)
{
    if (fabs(b.re) >= fabs(b.im)) {
        double r = b.im / b.re, den = b.re + b.im * r;
        return zmk((a.re + a.im * r) / den, (a.im - a.re * r) / den);
    }
    double r = b.re / b.im, den = b.re * r + b.im;
    return zmk((a.re * r + a.im) / den, (a.im * r - a.re) / den);
}

typedef struct {
    int l, u, hp;
    const c64 *g;
    const c64 *head;
    c64 shift;
} fsopc_t;

static inline c64 fsc_entry(const fsopc_t *A, int64_t row, int64_t col)
{
    const int64_t d = col - row;
    if (d > A->u || d < -A->l || row < 0 || col < 0) return zmk(0, 0);
    const int rr = A->u - (int)d;
    const int64_t t = row < col ? row : col;
    c64 v;
    if (t < A->hp) v = A->head[(int64_t)rr * A->hp + t];
    else {
        const c64 *g = A->g + rr * 5;
        const c64 num = zadd_(g[1], zmk(g[2].re * (double)t, g[2].im * (double)t));
        const double den = g[3].re + g[4].re * (double)t;
        v = zadd_(g[0], zmk(num.re / den, num.im / den));
    }
    if (d == 0) v = zsub_(v, A->shift);
    return v;
}

typedef void (*fsc_rec_fn)(void *ctx, int64_t k, const c64 *Lrow, const c64 *v, c64 tau);

static void fsc_sweep(const fsopc_t *A, int64_t lo, int64_t hi, int64_t kstop, int64_t krec, fsc_rec_fn rec, void *ctx)
{
    const int l = A->l, u = A->u, NC = l + u + 1, NR = u + 1;
    c64 W[MAXB][MAXB];
    int64_t k = hi;
    for (int c = 0; c < NC; c++)
        for (int r = 0; r < NR; r++) {
            const int64_t row = k - r, col = k - c;
            W[c][r] = (row >= lo && col >= lo) ? fsc_entry(A, row, col) : zmk(0, 0);
        }
    for (; k >= kstop; k--) {
        c64 x[MAXB], tau = zmk(0, 0);
        for (int r = 0; r < NR; r++) x[r] = W[0][r];
        {
            double s = zabs2_(x[0]);
            for (int r = 1; r < NR; r++) s = s + zabs2_(x[r]);
            const double normu = sqrt(s);
            if (normu != 0.0) {
                c64 xi1 = x[0];
                const double nu = copysign(normu, xi1.re);
                xi1.re = xi1.re + nu;
                x[0] = zmk(-nu, 0);
                for (int r = 1; r < NR; r++) x[r] = zdiv_(x[r], xi1);
                tau = zmk(xi1.re / nu, xi1.im / nu);
            }
            const c64 tauc = zconj_(tau);
            for (int r = 0; r < NR; r++) W[0][r] = x[r];
            for (int c = 1; c < NC; c++) {
                c64 vA = W[c][0];
                for (int r = 1; r < NR; r++) vA = zadd_(vA, zmul_(zconj_(x[r]), W[c][r]));
                vA = zmul_(tauc, vA);
                W[c][0] = zsub_(W[c][0], vA);
                for (int r = 1; r < NR; r++) W[c][r] = zsub_(W[c][r], zmul_(x[r], vA));
            }
        }
        if (k <= krec) {
            c64 Lrow[MAXB];
            for (int c = 0; c < NC; c++) Lrow[c] = W[c][0];
            rec(ctx, k, Lrow, x + 1, tau);
        }
        for (int c = 0; c + 1 < NC; c++)
            for (int r = 0; r + 1 < NR; r++) W[c][r] = W[c + 1][r + 1];
        for (int r = 0; r + 1 < NR; r++) W[NC - 1][r] = zmk(0, 0);
        for (int c = 0; c < NC; c++) {
            const int64_t row = k - 1 - u, col = k - 1 - c;
            W[c][u] = (row >= lo && col >= lo) ? fsc_entry(A, row, col) : zmk(0, 0);
        }
    }
}

typedef struct { int64_t base; int m, NC; c64 blk[MAXB][MAXB]; } blkc_ctx;
static void recc_block(void *ctx_, int64_t k, const c64 *Lrow, const c64 *v, c64 tau)
{
    (void)v; (void)tau;
    blkc_ctx *c = (blkc_ctx *)ctx_;
    const int64_t i = k - c->base;
    if (i < 0 || i >= c->m) return;
    for (int j = 0; j < c->m; j++) {
        const int64_t cc = i - j;
        c->blk[i][j] = (cc >= 0 && cc < c->NC) ? Lrow[cc] : zmk(0, 0);
    }
}
typedef struct { int64_t j; int NC, u; c64 *Lband, *V, *tau; } outc_ctx;
static void recc_out(void *ctx_, int64_t k, const c64 *Lrow, const c64 *v, c64 tau)
{
    outc_ctx *c = (outc_ctx *)ctx_;
    if (k >= c->j) return;
    for (int cc = 0; cc < c->NC; cc++) c->Lband[k * c->NC + cc] = Lrow[cc];
    for (int r = 0; r < c->u; r++) c->V[k * c->u + r] = v[r];
    c->tau[k] = tau;
}

static int fsqlc_one(const fsopc_t *A, int64_t j, double tol, int64_t maxN, c64 *Lband, c64 *V, c64 *tau, int64_t *N_used)
{
    const int l = A->l, u = A->u, NC = l + u + 1;
    const int64_t check1 = (j - l - u > 1) ? j - l - u : 1;
    const int64_t lo = check1 - 1;
    const int m = (int)(j - check1);
    blkc_ctx Ls, Ll;
    Ls.base = Ll.base = lo + 1; Ls.m = Ll.m = m; Ls.NC = Ll.NC = NC;
    memset(Ls.blk, 0, sizeof(Ls.blk)); memset(Ll.blk, 0, sizeof(Ll.blk));
    int64_t N = j;
    fsc_sweep(A, lo, N - 1, lo + 1, j - 1, recc_block, &Ls);
    double Lerr = 1.0;
    int status = 0;
    while (Lerr > tol) {
        fsc_sweep(A, lo, 2 * N - 1, lo + 1, j - 1, recc_block, &Ll);
        double s = 0.0;
        for (int cj = 0; cj < m; cj++)
            for (int ri = 0; ri < m; ri++) s = s + zabs2_(zsub_(Ll.blk[ri][cj], Ls.blk[ri][cj]));
        Lerr = sqrt(s);
        if (N >= maxN) { status = 3; break; }
        memcpy(Ls.blk, Ll.blk, sizeof(Ls.blk));
        N = 2 * N;
    }
    if (status == 0) {
        outc_ctx oc = {j, NC, u, Lband, V, tau};
        fsc_sweep(A, 0, N / 2 - 1, 0, j - 1, recc_out, &oc);
        *N_used = N / 2;
    } else {
        for (int64_t k = 0; k < j * NC; k++) Lband[k] = zmk(NAN, NAN);
        for (int64_t k = 0; k < j * u; k++) V[k] = zmk(NAN, NAN);
        for (int64_t k = 0; k < j; k++) tau[k] = zmk(NAN, NAN);
        *N_used = N;
    }
    return status;
}

int ila_oracle_ql_finite_section_c64(int l, int u, int64_t batch, const double *g, const double *shift, int hp,
                                     const double *head, int64_t j, double tol, int64_t maxN, double *Lband, double *V,
                                     double *tau, int64_t *N_used, int32_t *status, int nthreads)
{
    if (l < 0 || u < 1 || l + u + 1 > MAXB || j < l + u + 2) return -4;
    const int nd = l + u + 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic)
    for (int64_t i = 0; i < batch; i++) {
        fsopc_t A = {l, u, hp, (const c64 *)g + i * nd * 5, (const c64 *)head, shift ? ((const c64 *)shift)[i] : zmk(0, 0)};
        status[i] = fsqlc_one(&A, j, tol, maxN, (c64 *)Lband + i * j * nd, (c64 *)V + i * j * u, (c64 *)tau + i * j, N_used + i);
    }
    return 0;
}
