/* ila_oracle_triconj.c — CPU oracle for the tridiagonal conjugation  Y = Tridiagonal(U*X/U)  (U upper triangular banded,
 * X tridiagonal) and the bidiagonal conjugation  B = Bidiagonal(inv(U)*X*V).
 *
 * TEST INFRASTRUCTURE (same rules as ila_oracle.c): only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may call it.
 *
 * Restates src/banded/tridiagonalconjugation.jl (InfiniteLinearAlgebra.jl v0.10.3) with the reference's loop structure —
 * the in-place, sequential form, NOT the per-row closed form of the CUDA kernel — so that agreement of the two is a
 * check of the kernel's restructuring:
 *   initiate_upper_mul_tri_triview!   :42-58    first row of Tridiagonal(U*X)
 *   main_upper_mul_tri_triview!       :75-91    rows kr: UXdl[k-1], UXd[k], UXdu[k]
 *   initiate_tri_mul_invupper_triview!:149-163  Yd[1] = a/R11, Ydu[1] = b - a*R12/R11
 *   main_tri_mul_invupper_triview!    :179-197  rows kr (Ydu[k-1] is finished by row k)
 *   resizedata!(::TridiagonalConjugationData, n) :260-282  rows 2:n+1 of both sweeps
 * and src/banded/bidiagonalconjugation.jl:
 *   _compute_column_up! / _compute_column_lo!    :37-60
 * Products and sums are separate IEEE operations (compile with -ffp-contract=off), `a*b/c` is `(a*b)/c`, `a*b + c*d + e*f`
 * is `((a*b) + (c*d)) + (e*f)` as Julia parses them.
 *
 * Band layout (one operator): U[d*uld + k] = U[k+1, k+1+d] (0-based k, d = 0..nbands-1, nbands = 2 or 3);
 * X[0*xld + k] = X[k+2, k+1] (dl), X[1*xld + k] = X[k+1, k+1] (d), X[2*xld + k] = X[k+1, k+2] (du);
 * outputs Y / UX: [0*n + k] = dl, [1*n + k] = d, [2*n + k] = du, k < n.  Bands must hold n + 1 entries.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void triconj_one(int nbands, int64_t n, const double *U, int64_t uld, const double *X, int64_t xld, double *Y, double *UXo)
{
    /* 1-based accessors as in the reference */
#define R0(k) U[(k) - 1]
#define R1(k) U[uld + (k) - 1]
#define R2(k) (nbands > 2 ? U[2 * uld + (k) - 1] : 0.0)
#define XDL(k) X[(k) - 1]
#define XD(k) X[xld + (k) - 1]
#define XDU(k) X[2 * xld + (k) - 1]
    const int64_t m = n + 1;                 /* rows 1..n+1 of UX are touched */
    double *UXdl = (double *)malloc(sizeof(double) * (m + 1));
    double *UXd = (double *)malloc(sizeof(double) * (m + 1));
    double *UXdu = (double *)malloc(sizeof(double) * (m + 1));
    double *Ydl = (double *)malloc(sizeof(double) * (m + 1));
    double *Yd = (double *)malloc(sizeof(double) * (m + 1));
    double *Ydu = (double *)malloc(sizeof(double) * (m + 1));
    /* initiate_upper_mul_tri_triview! */
    double ak = XD(1), ck = XDL(1), bk, ckm1;
    {
        const double Ukk = R0(1), Ukk1 = R1(1), Ukk2 = R2(1);
        UXd[1] = Ukk * ak + Ukk1 * ck;
        bk = XDU(1); ak = XD(2); ckm1 = ck; ck = XDL(2);
        UXdu[1] = (Ukk * bk + Ukk1 * ak) + Ukk2 * ck;
    }
    /* main_upper_mul_tri_triview!, kr = 2:n+1 (the last row only needs UXdl[n]; its d/du entries are computed when the
     * bands are long enough, as upstream, but never returned) */
    for (int64_t k = 2; k <= n + 1; k++) {
        const double Ukk = R0(k);
        UXdl[k - 1] = Ukk * ckm1;
        if (k <= n) {
            const double Ukk1 = R1(k), Ukk2 = R2(k);
            UXd[k] = Ukk * ak + Ukk1 * ck;
            bk = XDU(k); ak = XD(k + 1); ckm1 = ck; ck = XDL(k + 1);
            UXdu[k] = (Ukk * bk + Ukk1 * ak) + Ukk2 * ck;
        }
    }
    /* initiate_tri_mul_invupper_triview! (R = U, X = UX) */
    double Rkk = R0(1), Rkk1 = R1(1);
    Yd[1] = UXd[1] / Rkk;
    Ydu[1] = UXdu[1] - UXd[1] * Rkk1 / Rkk;
    /* main_tri_mul_invupper_triview!, kr = 2:n+1 */
    for (int64_t k = 2; k <= n + 1; k++) {
        const double cm = UXdl[k - 1];
        Ydl[k - 1] = cm / Rkk;
        if (k <= n) {
            const double a = UXd[k], b = UXdu[k];
            Yd[k] = a - cm * Rkk1 / Rkk;
            Ydu[k] = cm / Rkk;
            const double Rm1k = Rkk1, Rm1k1 = R2(k - 1);
            Rkk = R0(k); Rkk1 = R1(k);
            Yd[k] /= Rkk;
            Ydu[k - 1] /= Rkk;
            Ydu[k] *= Rm1k * Rkk1 / Rkk - Rm1k1;
            Ydu[k] += b - a * Rkk1 / Rkk;
        } else {
            Rkk = R0(k);
            Ydu[k - 1] /= Rkk;
        }
    }
    for (int64_t k = 0; k < n; k++) {
        Y[k] = Ydl[k + 1]; Y[n + k] = Yd[k + 1]; Y[2 * n + k] = Ydu[k + 1];
        if (UXo) { UXo[k] = UXdl[k + 1]; UXo[n + k] = UXd[k + 1]; UXo[2 * n + k] = UXdu[k + 1]; }
    }
    free(UXdl); free(UXd); free(UXdu); free(Ydl); free(Yd); free(Ydu);
#undef R0
#undef R1
#undef R2
#undef XDL
#undef XD
#undef XDU
}

/* batch form: U batch x nbands x uld (ustride = nbands*uld), X shared (xstride == 0) or per operator */
void ila_oracle_triconj_f64(int nbands, int64_t batch, int64_t n, const double *U, int64_t uld, const double *X, int64_t xld,
                            int64_t xstride, double *Y, double *UXo, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < batch; i++)
        triconj_one(nbands, n, U + i * nbands * uld, uld, X + i * xstride, xld, Y + i * 3 * n, UXo ? UXo + i * 3 * n : 0);
}

/* ---- bidiagonal conjugation (src/banded/bidiagonalconjugation.jl:37-60) -----------------------------------------------
 * U and C = X*V are given by their 2x2-relevant bands: U[(d+1)*ld + k] for d = -1,0,1: U[k+2,k+1], U[k+1,k+1], U[k+1,k+2];
 * C likewise.  uplo 'U': dv[i] = (u11 c22 - u21 c12)/det, ev[i-1] = (u22 c12 - u12 c22)/det with the 2x2 blocks at
 * rows/cols (i-1, i); uplo 'L': dv[i] = (u22 c11 - u12 c21)/det, ev[i] = (u11 c21 - u21 c11)/det at rows/cols (i, i+1).
 * `inv(x)` is 1/x; products by the inverse as upstream.  Outputs dv[n], ev[n] ('U': ev[k] = B[k+1,k+2]). */
void ila_oracle_biconj_f64(int upper, int64_t batch, int64_t n, const double *U, int64_t ld, const double *Cm, double *dv,
                           double *ev, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < batch; b++) {
        const double *Ub = U + b * 3 * ld, *Cb = Cm + b * 3 * ld;
        double *d = dv + b * n, *e = ev + b * n;
#define UL(k) Ub[(k) - 1]              /* U[k+1,k] */
#define UD(k) Ub[ld + (k) - 1]         /* U[k,k] */
#define UU(k) Ub[2 * ld + (k) - 1]     /* U[k,k+1] */
#define CL(k) Cb[(k) - 1]
#define CD(k) Cb[ld + (k) - 1]
#define CU(k) Cb[2 * ld + (k) - 1]
        if (upper) {
            /* columns 1..n+1 (ev[n] needs column n+1) */
            for (int64_t i = 1; i <= n + 1; i++) {
                if (i == 1) { d[0] = CD(1) / UD(1); continue; }
                const double u11 = UD(i - 1), u21 = UL(i - 1), u12 = UU(i - 1), u22 = UD(i);
                const double c12 = CU(i - 1), c22 = CD(i);
                const double inv = 1.0 / (u11 * u22 - u12 * u21);
                if (i <= n) d[i - 1] = inv * (u11 * c22 - u21 * c12);
                e[i - 2] = inv * (u22 * c12 - u12 * c22);
            }
        } else {
            for (int64_t i = 1; i <= n; i++) {
                const double u11 = UD(i), u21 = UL(i), u12 = UU(i), u22 = UD(i + 1);
                const double c11 = CD(i), c21 = CL(i);
                const double inv = 1.0 / (u11 * u22 - u12 * u21);
                d[i - 1] = inv * (u22 * c11 - u12 * c21);
                e[i - 1] = inv * (u11 * c21 - u21 * c11);
            }
        }
#undef UL
#undef UD
#undef UU
#undef CL
#undef CD
#undef CU
    }
}
