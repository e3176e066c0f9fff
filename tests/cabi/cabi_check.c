/* cabi_check.c — plain C (gcc) against include/ila_b200.h and libila_b200.so: proves that the header a `ccall` / cgo / FFI
 * binding is written from matches the library, with no hand-written signature table in between (tests/test_cabi_c.py
 * compiles and runs it).  Exit code 0 = every check passed; it prints one line per check.
 *
 * Without a CUDA device every compute entry must return ILA_ERR_NOGPU (there is no CPU fallback); with one, the values
 * are checked against constants the reference holds / scipy gives:
 *   K3  ql(A - λI).L[1,1] of the Bull-head symbol at the probe shifts of SURVEY Appendix B (finite-section values)
 *   K1+K2  the Bessel recurrence solve of README.md:41-66 (z = 100): x[k] = J_k(100)
 *   K5+K2  cholesky(S) \ e1 for a pentadiagonal S against the QR solve of the same operator (test_infcholesky.jl:5-11)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ila_b200.h"

static int failures = 0;
#define CHECK(cond, ...)                                  \
    do {                                                  \
        if (cond) printf("ok    ");                       \
        else { printf("FAIL  "); failures++; }            \
        printf(__VA_ARGS__);                              \
        printf("\n");                                     \
    } while (0)

int main(void)
{
    printf("libila_b200 version %d, %d CUDA device(s)\n", ila_version(), ila_device_count());
    CHECK(ila_version() >= 100, "ila_version");
    const int have_gpu = ila_device_count() > 0;

    /* ---- K3: Bull-head, BandedMatrices data column (+1, 0, -1, -2, -3) = (2i, 0, 0, 1, 0.7) ---- */
    const ila_c64 col[5] = {{0, 2}, {0, 0}, {0, 0}, {1, 0}, {0.7, 0}};
    const ila_c64 lam[5] = {{5, 0}, {-3, -0.1}, {2, 3}, {-6, 2}, {0.1, 0.1}};
    const double want[4] = {-4.747700587226856, -1.83554289328665, -2.4543911495549953, -5.917786604460436};
    ila_c64 L11[5];
    int32_t st[5];
    double Lre[5];
    uint8_t st8[5];
    int rc = ila_ql_toeplitz_l11_c64(3, 1, col, lam, 5, 1, L11, NULL, st, 1);
    int rc2 = ila_ql_toeplitz_l11re_c64(3, 1, col, lam, 5, 1, Lre, st8, 1);
    if (!have_gpu) {
        CHECK(rc == ILA_ERR_NOGPU && rc2 == ILA_ERR_NOGPU, "K3 without a GPU: ILA_ERR_NOGPU (%d, %d), \"%s\"", rc, rc2, ila_last_error());
    } else {
        CHECK(rc == 0 && rc2 == 0, "K3 calls return 0 (%d, %d) %s", rc, rc2, rc || rc2 ? ila_last_error() : "");
        for (int i = 0; i < 4; i++) {
            CHECK(st[i] == ILA_OK && fabs(L11[i].re - want[i]) <= 1e-13 * fabs(want[i]) && L11[i].im == 0.0,
                  "K3 L[1,1] at shift %d: %.16g (want %.16g)", i, L11[i].re, want[i]);
            CHECK(st8[i] == ILA_OK && Lre[i] == L11[i].re, "K3 compact form agrees at shift %d", i);
        }
        CHECK(st[4] == ILA_NO_QL && isnan(L11[4].re) && st8[4] == ILA_NO_QL && isnan(Lre[4]), "K3 DomainError shift -> status 1, NaN");
        uint64_t bits;
        memcpy(&bits, &Lre[4], 8);
        CHECK((bits & 0xff) == ILA_NO_QL, "K3 NaN payload carries the status");
    }

    /* ---- K1+K2: A \ [J_1(z); 0...], A = BandedMatrix(0 => -2(0:∞)/z, ±1 => 1), z = 100 ---- */
    {
        const double z = 100.0, j1 = -0.07714535201411216, j0 = 0.019985850304223122, j5 = -0.07419573696451393;
        const double p[3] = {1, 0, 1}, q[3] = {0, -2, 0}, r[3] = {1, z, 1}, b[1] = {j1};
        enum { NCAP = 4000 };
        double *x = malloc(sizeof(double) * NCAP);
        int64_t xoff[2], ncols[1], nconv[1];
        int32_t s1[1];
        rc = ila_qr_banded_solve_f64(1, 1, 1, p, q, r, NULL, 0, NULL, 1, b, 2.2250738585072014e-308, 1000, NCAP, NULL, x, NCAP,
                                     xoff, ncols, nconv, NULL, NULL, NULL, s1, 1);
        if (!have_gpu) CHECK(rc == ILA_ERR_NOGPU, "K1+K2 without a GPU: ILA_ERR_NOGPU (%d)", rc);
        else {
            CHECK(rc == 0 && s1[0] == ILA_OK, "K1+K2 call (%d, status %d) %s", rc, s1[0], rc ? ila_last_error() : "");
            CHECK(ncols[0] == nconv[0] + 1000 && xoff[1] == ncols[0], "K1+K2 datasize = n_conv + COLGROWTH - 1 + l (%lld, %lld)",
                  (long long)ncols[0], (long long)nconv[0]);
            CHECK(fabs(x[0] - j0) < 1e-14 && fabs(x[1] - j1) < 1e-14 && fabs(x[5] - j5) < 1e-14, "Bessel: x[0..5] = J_k(100): %.16g %.16g %.16g",
                  x[0], x[1], x[5]);
        }
        free(x);
    }

    /* ---- K5+K2 against K1+K2: S = pentadiag(6.5 + 1e-3 k; -4; 1), S \ e1 by Cholesky and by QR ---- */
    {
        const double pc[3] = {1, -4, 6.5}, qc[3] = {0, 0, 1e-3}, b[1] = {1.0};
        const double pq[5] = {1, -4, 6.5, -4, 1}, qq[5] = {0, 0, 1e-3, 0, 0};
        enum { NCAP = 9000 };
        double *xc = malloc(sizeof(double) * NCAP), *xq = malloc(sizeof(double) * NCAP);
        int64_t xoff[2], nc[1], nv[1], nc2[1], nv2[1];
        int32_t s1[1], s2[1];
        rc = ila_chol_banded_solve_f64(2, 1, pc, qc, NULL, NULL, 0, NULL, 1, b, 2.2250738585072014e-308, 1000, NCAP, NULL, xc, NCAP,
                                       xoff, nc, nv, NULL, NULL, s1, 1);
        rc2 = ila_qr_banded_solve_f64(2, 2, 1, pq, qq, NULL, NULL, 0, NULL, 1, b, 2.2250738585072014e-308, 1000, NCAP, NULL, xq, NCAP,
                                      xoff, nc2, nv2, NULL, NULL, NULL, s2, 1);
        if (!have_gpu) CHECK(rc == ILA_ERR_NOGPU && rc2 == ILA_ERR_NOGPU, "K5 without a GPU: ILA_ERR_NOGPU (%d, %d)", rc, rc2);
        else {
            CHECK(rc == 0 && rc2 == 0 && s1[0] == 0 && s2[0] == 0, "K5+K2 / K1+K2 calls (%d %d, status %d %d)", rc, rc2, s1[0], s2[0]);
            double mx = 0, sc = 0;
            const int64_t m = nc[0] < nc2[0] ? nc[0] : nc2[0];
            for (int64_t i = 0; i < m; i++) {
                mx = fmax(mx, fabs(xc[i] - xq[i]));
                sc = fmax(sc, fabs(xq[i]));
            }
            CHECK(m > 1000 && mx <= 1e-12 * sc, "cholesky(S) \\ e1 == qr(S) \\ e1 over %lld entries: max diff %.3g (scale %.3g)", (long long)m, mx, sc);
        }
        free(xc);
        free(xq);
    }

    /* ---- argument checking and error text are part of the ABI ---- */
    rc = ila_ql_toeplitz_l11_c64(3, 1, NULL, lam, 5, 1, L11, NULL, st, 1);
    CHECK(rc == ILA_ERR_ARG && strlen(ila_last_error()) > 0, "NULL symbol -> ILA_ERR_ARG (%d), \"%s\"", rc, ila_last_error());
    CHECK(ila_qr_banded_state_doubles(1, 1) == 12 && ila_chol_banded_state_doubles(2) == 34, "state sizes: qr(1,1) = %d, chol(2) = %d",
          ila_qr_banded_state_doubles(1, 1), ila_chol_banded_state_doubles(2));
    printf(failures ? "%d check(s) FAILED\n" : "all checks passed\n", failures);
    return failures ? 1 : 0;
}
