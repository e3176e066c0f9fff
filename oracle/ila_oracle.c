/* ila_oracle.c — CPU oracle for the adaptive banded QR / Cholesky solves.
 *
 * TEST INFRASTRUCTURE.  This file is a plain-C restatement of the reference's algorithms
 * (InfiniteLinearAlgebra.jl v0.10.3; citations in ila_oracle_impl.h).  It is compiled into
 * oracle/_build/libila_oracle.so by oracle/Makefile and may only be loaded by tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs.  The product library
 * (infinitelinearalgebra.jl_b200/csrc) never links or calls it.
 *
 * Parity status: the reference cannot run in this image (no Julia).  The restatement is
 * pinned by the reference's own known answers for the path — test/test_infqr.jl:13,34-41
 * (factors[1,1] = -sqrt 2, tau[1] = 1+sqrt(2)/2), :19-28 (factors / tau / trailing columns against a dense
 * Householder QR of the finite section), :87 (residual), :93-105 (Bessel J against closed form),
 * :114-119 (5-band against a finite solve), test/test_infcholesky.jl:5-11,54-60 — see
 * tests/test_oracle_qr.py and tests/test_oracle_chol.py.  Whether the reference's BLAS tail loops
 * contract multiply-adds is not knowable here: the oracle never fuses (-ffp-contract=off).
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REAL double
#define SQRT sqrt
#define FABS fabs
#define COPYSIGN copysign
#define SUF _f64
#include "ila_oracle_impl.h"
#undef REAL
#undef SQRT
#undef FABS
#undef COPYSIGN
#undef SUF

#define REAL __float128
#define SQRT sqrtq
#define FABS fabsq
#define COPYSIGN copysignq
#define SUF _f128
#include "ila_oracle_impl.h"
#undef REAL
#undef SQRT
#undef FABS
#undef COPYSIGN
#undef SUF

int ila_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Batched adaptive QR solve.  Layouts mirror include/ila_b200.h (ila_qr_banded_solve_f64):
 *   p,q,r : batch x (l+u+1) generator coefficients (r may be NULL), shift: batch (may be NULL)
 *   head  : (l+u+1) x hp shared explicit head (may be NULL), b: batch x nb
 *   x     : batch x ldx (instance-major, zero padded), ncols / nconv / status: batch
 *   factors_opt: batch x ldx x (2l+u+1), tau_opt / qtb_opt: batch x ldx (may be NULL)
 */
int ila_oracle_qr_banded_solve_f64(int l, int u, int64_t batch, const double *p, const double *q, const double *r,
                                   const double *shift, int hp, const double *head, int nb, const double *b,
                                   double tol, int colgrowth, int64_t ncap, int64_t ldx, double *x, int64_t *ncols,
                                   int64_t *nconv, double *factors_opt, double *tau_opt, double *qtb_opt,
                                   int32_t *status, int nthreads)
{
    const int nd = l + u + 1, ld = 2 * l + u + 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < batch; i++) {
        op_t_f64 A = {l, u, p + i * nd, q + i * nd, r ? r + i * nd : NULL, shift ? shift[i] : 0.0, hp, head};
        double *xi = (double *)calloc((size_t)ncap + 1, sizeof(double));
        double *fi = factors_opt ? (double *)calloc((size_t)ld * (ncap + 1), sizeof(double)) : NULL;
        double *ti = tau_opt ? (double *)calloc((size_t)ncap + 1, sizeof(double)) : NULL;
        double *qi = qtb_opt ? (double *)calloc((size_t)ncap + 1, sizeof(double)) : NULL;
        int64_t n = 0, nc = 0;
        int st = qr_banded_solve_one_f64(&A, nb, b + i * nb, tol, colgrowth, ncap, xi, &n, &nc, fi, ti, qi);
        int64_t ncopy = n < ldx ? n : ldx;
        memset(x + i * ldx, 0, sizeof(double) * (size_t)ldx);
        memcpy(x + i * ldx, xi, sizeof(double) * (size_t)ncopy);
        if (fi) { memset(factors_opt + i * ldx * ld, 0, sizeof(double) * ldx * ld); memcpy(factors_opt + i * ldx * ld, fi, sizeof(double) * ncopy * ld); }
        if (ti) { memset(tau_opt + i * ldx, 0, sizeof(double) * ldx); memcpy(tau_opt + i * ldx, ti, sizeof(double) * ncopy); }
        if (qi) { memset(qtb_opt + i * ldx, 0, sizeof(double) * ldx); memcpy(qtb_opt + i * ldx, qi, sizeof(double) * ncopy); }
        ncols[i] = n;
        nconv[i] = nc;
        status[i] = st;
        free(xi); free(fi); free(ti); free(qi);
    }
    return 0;
}

/* Same solve carried in __float128 (inputs/outputs double): the conditioning-aware reference. */
int ila_oracle_qr_banded_solve_f128(int l, int u, int64_t batch, const double *p, const double *q, const double *r,
                                    const double *shift, int hp, const double *head, int nb, const double *b,
                                    double tol, int colgrowth, int64_t ncap, int64_t ldx, double *x, int64_t *ncols,
                                    int64_t *nconv, int32_t *status, int nthreads)
{
    const int nd = l + u + 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < batch; i++) {
        op_t_f128 A = {l, u, p + i * nd, q + i * nd, r ? r + i * nd : NULL, shift ? shift[i] : 0.0, hp, head};
        __float128 *xi = (__float128 *)calloc((size_t)ncap + 1, sizeof(__float128));
        int64_t n = 0, nc = 0;
        int st = qr_banded_solve_one_f128(&A, nb, b + i * nb, tol, colgrowth, ncap, xi, &n, &nc, NULL, NULL, NULL);
        int64_t ncopy = n < ldx ? n : ldx;
        memset(x + i * ldx, 0, sizeof(double) * (size_t)ldx);
        for (int64_t k = 0; k < ncopy; k++) x[i * ldx + k] = (double)xi[k];
        ncols[i] = n;
        nconv[i] = nc;
        status[i] = st;
        free(xi);
    }
    return 0;
}

/* Batched adaptive Cholesky solve (upper bands only: p,q,r are batch x (u+1), row 0 = diagonal +u … row u = main). */
int ila_oracle_chol_banded_solve_f64(int u, int64_t batch, const double *p, const double *q, const double *r,
                                     const double *shift, int hp, const double *head, int nb, const double *b,
                                     double tol, int colgrowth, int64_t ncap, int64_t ldx, double *x, int64_t *ncols,
                                     int64_t *nconv, double *factors_opt, double *y_opt, int32_t *status, int nthreads)
{
    const int nd = u + 1, ld = u + 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < batch; i++) {
        /* symmetric operator described by its upper bands: l = 0 in the generator */
        op_t_f64 A = {0, u, p + i * nd, q + i * nd, r ? r + i * nd : NULL, shift ? shift[i] : 0.0, hp, head};
        double *xi = (double *)calloc((size_t)ncap + 1, sizeof(double));
        double *fi = factors_opt ? (double *)calloc((size_t)ld * (ncap + 1), sizeof(double)) : NULL;
        double *yi = y_opt ? (double *)calloc((size_t)ncap + 1, sizeof(double)) : NULL;
        int64_t n = 0, nc = 0;
        int st = chol_banded_solve_one_f64(&A, nb, b + i * nb, tol, colgrowth, ncap, xi, &n, &nc, fi, yi);
        int64_t ncopy = n < ldx ? n : ldx;
        memset(x + i * ldx, 0, sizeof(double) * (size_t)ldx);
        memcpy(x + i * ldx, xi, sizeof(double) * (size_t)ncopy);
        if (fi) { memset(factors_opt + i * ldx * ld, 0, sizeof(double) * ldx * ld); memcpy(factors_opt + i * ldx * ld, fi, sizeof(double) * ncopy * ld); }
        if (yi) { memset(y_opt + i * ldx, 0, sizeof(double) * ldx); memcpy(y_opt + i * ldx, yi, sizeof(double) * ncopy); }
        ncols[i] = n;
        nconv[i] = nc;
        status[i] = st;
        free(xi); free(fi); free(yi);
    }
    return 0;
}
