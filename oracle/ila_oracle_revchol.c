/* ila_oracle_revchol.c — CPU oracle for the adaptive reverse Cholesky factorization of ∞ symmetric tridiagonal operators.
 *
 * TEST INFRASTRUCTURE (same rules as ila_oracle.c).
 *
 * Restates src/banded/infreversecholeskytridiagonal.jl (InfiniteLinearAlgebra.jl v0.10.3):
 *   reversecholesky_layout(::SymTridiagonalLayout, ...)  :66-74   tol = eps, factor initialised at n = 2^4
 *   _expand_factor! / __expand_factor!                    :76-79, 132-141  N := 2n when N < n, then
 *                                                         `while !has_converged && N <= MAX: N := 2N; _finite_revchol!`
 *                                                         (the first test runs on unwritten storage; it is taken as "no")
 *   _finite_revchol!                                      :118-129 la[N] = sqrt(a[N]); lb[i] = b[i]/la[i+1]; la[i] = sqrt(a[i] - lb[i]^2)
 *   compute_ξ / has_converged                             :81-110  ν = 1/la[N]; ν *= -(lb[i]/la[i]) ...; ξ <= eps
 *   MAX_TRIDIAG_CHOL_N = 2^21 - 1                         :1
 * A = L'L with L lower bidiagonal (diagonal la, sub-diagonal lb); the leading n x n block of L is returned.
 * One backward sweep per N computes the factor and ξ together (ξ only needs running values); operation order per
 * entry as in the reference, nothing fused.
 *
 * Operator entries: value(k) = c0 + (p + q k)/(r + s k), k = 0-based index, from the 5 generator coefficients
 * (c0, p, q, r, s) of the diagonal and of the off-diagonal; an optional explicit head replaces the first hp entries;
 * `shift` is subtracted from the diagonal.  Covers the reference's fixtures: Fill(3)/Fill(1), Ones / 1 ./ (2:∞),
 * 5 .+ 1 ./ (2:∞) / Ones, finite heads (test/test_infreversecholesky.jl:5-12, 18-24, 53-59).
 * status: 0 ok, 3 not converged before N exceeded MAX (the reference warns and returns), 4 non-positive pivot
 * (sqrt of a negative number: DomainError upstream).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_TRIDIAG_CHOL_N ((1 << 21) - 1)

static inline double gen_value(const double *g, int64_t k)
{
    const double num = g[1] + g[2] * (double)k;
    const double den = g[3] + g[4] * (double)k;
    return g[0] + num / den;
}

static int revchol_one(const double *ga, const double *gb, double shift, int hp, const double *a_head, const double *b_head,
                       int64_t n, double *la_out, double *lb_out, int64_t *N_out)
{
    const double eps = 2.220446049250313e-16;
    int64_t N = 2 * n;
    int status = 0;
#define AV(k) (((k) < hp ? a_head[k] : gen_value(ga, (k))) - shift)
#define BV(k) ((k) < hp ? b_head[k] : gen_value(gb, (k)))
    for (;;) {
        if (N > MAX_TRIDIAG_CHOL_N) { status = 3; break; }
        N *= 2;
        /* i = N (1-based) */
        double aN = AV(N - 1);
        if (!(aN >= 0.0)) { status = 4; break; }
        double la_next = sqrt(aN);                 /* la[i+1] of the recurrence */
        double nu = 1.0 / la_next;
        double xi = 0.0;
        int bad = 0;
        if (n == N) xi = (la_next * nu) * (la_next * nu);
        for (int64_t i = N - 1; i >= 1; i--) {     /* 1-based row i */
            const double lb = BV(i - 1) / la_next;
            const double t = AV(i - 1) - lb * lb;
            if (!(t >= 0.0)) { bad = 1; break; }
            const double la = sqrt(t);
            if (i > n) {
                nu = nu * (-(lb / la));
            } else if (i == n) {
                nu = nu * (-(lb / la));
                xi = (la * nu) * (la * nu);
            } else {
                double xp = lb * nu;
                nu = nu * (-(lb / la));
                xp = xp + la * nu;
                xi = xi + xp * xp;
            }
            if (i <= n) { la_out[i - 1] = la; lb_out[i - 1] = lb; }
            la_next = la;
        }
        if (bad) { status = 4; break; }
        const double bN = BV(N - 1);
        const double scale = (bN == 0.0) ? 1.0 : fabs(bN);
        if (scale * sqrt(xi) <= eps) break;
    }
#undef AV
#undef BV
    if (status == 4)
        for (int64_t k = 0; k < n; k++) { la_out[k] = NAN; lb_out[k] = NAN; }
    *N_out = N;
    return status;
}

int ila_oracle_revchol_tridiag_f64(int64_t batch, const double *ga, const double *gb, const double *shift, int hp,
                                   const double *a_head, const double *b_head, int64_t n, double *la, double *lb,
                                   int64_t *N_used, int32_t *status, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic)
    for (int64_t i = 0; i < batch; i++) {
        status[i] = revchol_one(ga + i * 5, gb + i * 5, shift ? shift[i] : 0.0, hp, a_head, b_head, n, la + i * n,
                                lb + i * n, N_used + i);
    }
    return 0;
}
