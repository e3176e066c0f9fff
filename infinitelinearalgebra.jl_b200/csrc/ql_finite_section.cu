// ql_finite_section.cu — K9: adaptive finite-section QL of ∞ banded operators (real eltype), one thread per operator
// (batched over shifts / parameters): ql(A - λI) for operators WITHOUT a Toeplitz tail.
//
// Reference path replaced (InfiniteLinearAlgebra.jl v0.10.3, "experimental adaptive finite section QL", src/infql.jl:364-462):
//   ql(A::InfBandedMatrix, tol = eps)                       :385-388
//   initialadaptiveQLblock / cache_filldata!                :391-411, 440-458
//     j = block size (50), checkinds = max(1, j-l-u); the block L[2:j-checkinds+1, 2:j-checkinds+1] of
//     ql(A[checkinds:N, checkinds:N]) is compared (Frobenius norm) between N and 2N, N doubling from j until the
//     difference is <= tol (error("Reached max. iterations ...") once N >= maxN); then
//     F = ql(A[1:N÷2, 1:N÷2]) and factors[1:j,1:j], τ[1:j] are kept.
//   The finite banded ql! is the Householder QL restated at src/infql.jl:10-22 (real eltype: k = n..2).
// A QL sweep runs from the section's last column to its first and only rows k-u..k of columns k-l-u..k are live, so the
// kernel keeps that (u+1) x (l+u+1) window in registers — nothing of size N is stored — generates the entering row from
// the operator's generators, and emits rows of L only where they are wanted (the test block / the leading j rows).
// IEEE unfused arithmetic in the reference's order: L, the reflectors, τ and the N at which the test fires equal the
// oracle's (oracle/ila_oracle_fsql.c) bit for bit.
#include "ila_common.cuh"
#include "ila_kernels.cuh"
#include "ila_exact.cuh"
#include "ila_unfused.cuh"

namespace ila {

template <typename T, int L, int U>
__global__ void __launch_bounds__(32) ql_finite_section_kernel(FsqlArgsT<T> A)
{
    const int64_t inst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= A.batch) return;
    constexpr int NC = L + U + 1, NR = U + 1, MB = L + U;   // MB: largest test block
    // generator coefficients: c0, p, q of the scalar type, the denominator coefficients r, s real
    T g0[NC], g1[NC], g2[NC];
    double g3[NC], g4[NC];
#pragma unroll
    for (int rr = 0; rr < NC; rr++) {
        const T *gg = A.g + (inst * NC + rr) * 5;
        g0[rr] = gg[0]; g1[rr] = gg[1]; g2[rr] = gg[2];
        g3[rr] = zreal(gg[3]); g4[rr] = zreal(gg[4]);
    }
    const T shift = A.shift ? A.shift[inst] : zfrom<T>(0.0);
    const int hp = A.hp;
    const int64_t j = A.j;

    auto entry = [&](int64_t row, int64_t col, int rr) -> T {   // rr = U - (col - row), inside the band by construction
        if (row < 0 || col < 0) return zfrom<T>(0.0);
        const int64_t t = row < col ? row : col;
        T v;
        if (t < hp) v = A.head[(int64_t)rr * hp + t];
        else {
            const T num = zadd(g1[rr], zmulr(g2[rr], (double)t));
            const double den = __dadd_rn(g3[rr], __dmul_rn(g4[rr], (double)t));
            v = zadd(g0[rr], zdivr(num, den));
        }
        if (rr == U) v = zsub(v, shift);
        return v;
    };

    T blkS[MB][MB], blkL[MB][MB];
    T *Lband = A.Lband + inst * j * NC, *V = A.V + inst * j * U, *tau_out = A.tau + inst * j;

    // mode 0: fill blkS, 1: fill blkL (test block rows base..base+m-1), 2: write the leading j rows
    auto sweep = [&](int64_t lo, int64_t hi, int64_t kstop, int64_t krec, int mode, int64_t base, int m) {
        T W[NC][NR];
        int64_t k = hi;
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
            for (int r = 0; r < NR; r++) {
                const int64_t row = k - r, col = k - c;
                const int d = r - c;   // col - row
                W[c][r] = (row >= lo && col >= lo && d >= -L && d <= U) ? entry(row, col, U - d) : zfrom<T>(0.0);
            }
        for (; k >= kstop; k--) {
            T x[NR], tau = zfrom<T>(0.0);
#pragma unroll
            for (int r = 0; r < NR; r++) x[r] = W[0][r];
            if (k > lo || z_is_cplx<T>::value) {   // real eltype: the QL loop stops at the section's second column
                double s = zabs2(x[0]);
#pragma unroll
                for (int r = 1; r < NR; r++) s = __dadd_rn(s, zabs2(x[r]));
                const double normu = esqrt(s);
                if (normu != 0.0) {
                    T xi1 = x[0];
                    const double nu = copysign(normu, zreal(xi1));
                    xi1 = zaddr(xi1, nu);
                    x[0] = zfrom<T>(-nu);
#pragma unroll
                    for (int r = 1; r < NR; r++) x[r] = zdiv(x[r], xi1);
                    tau = zdivr(xi1, nu);
                }
                const T tauc = zconj(tau);
#pragma unroll
                for (int r = 0; r < NR; r++) W[0][r] = x[r];
#pragma unroll
                for (int c = 1; c < NC; c++) {
                    T vA = W[c][0];
#pragma unroll
                    for (int r = 1; r < NR; r++) vA = zadd(vA, zmul(zconj(x[r]), W[c][r]));
                    vA = zmul(tauc, vA);
                    W[c][0] = zsub(W[c][0], vA);
#pragma unroll
                    for (int r = 1; r < NR; r++) W[c][r] = zsub(W[c][r], zmul(x[r], vA));
                }
            }
            if (k <= krec) {
                if (mode == 2) {
                    if (k < j) {
#pragma unroll
                        for (int c = 0; c < NC; c++) Lband[k * NC + c] = W[c][0];
#pragma unroll
                        for (int r = 1; r < NR; r++) V[k * U + (r - 1)] = x[r];
                        tau_out[k] = tau;
                    }
                } else {
                    const int i = (int)(k - base);
                    if (i >= 0 && i < m) {
#pragma unroll
                        for (int jj = 0; jj < MB; jj++) {
                            const int cc = i - jj;
                            T v = zfrom<T>(0.0);
#pragma unroll
                            for (int c = 0; c < NC; c++)
                                if (c == cc) v = W[c][0];
                            if (jj < m) {
                                // (static indexing keeps the blocks in registers)
#pragma unroll
                                for (int ii = 0; ii < MB; ii++)
                                    if (ii == i) { if (mode == 0) blkS[ii][jj] = v; else blkL[ii][jj] = v; }
                            }
                        }
                    }
                }
            }
            // slide to column k-1
#pragma unroll
            for (int c = 0; c + 1 < NC; c++)
#pragma unroll
                for (int r = 0; r + 1 < NR; r++) W[c][r] = W[c + 1][r + 1];
#pragma unroll
            for (int r = 0; r + 1 < NR; r++) W[NC - 1][r] = zfrom<T>(0.0);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int64_t row = k - 1 - U, col = k - 1 - c;
                W[c][U] = (row >= lo && col >= lo) ? entry(row, col, c) : zfrom<T>(0.0);   // col - row = U - c -> data row c
            }
        }
    };

    const int64_t check1 = (j - L - U > 1) ? j - L - U : 1;
    const int64_t lo = check1 - 1;
    const int m = (int)(j - check1);
#pragma unroll
    for (int a = 0; a < MB; a++)
#pragma unroll
        for (int b = 0; b < MB; b++) { blkS[a][b] = zfrom<T>(0.0); blkL[a][b] = zfrom<T>(0.0); }
    int64_t N = j;
    sweep(lo, N - 1, lo + 1, j - 1, 0, lo + 1, m);
    double Lerr = 1.0;
    int32_t st = ILA_OK;
    while (Lerr > A.tol) {
        sweep(lo, 2 * N - 1, lo + 1, j - 1, 1, lo + 1, m);
        double s = 0.0;
#pragma unroll
        for (int cj = 0; cj < MB; cj++)
#pragma unroll
            for (int ri = 0; ri < MB; ri++)
                if (cj < m && ri < m) {
                    const T dlt = zsub(blkL[ri][cj], blkS[ri][cj]);
                    s = __dadd_rn(s, zabs2(dlt));
                }
        Lerr = esqrt(s);
        if (N >= A.maxN) { st = ILA_NOT_CONVERGED; break; }
#pragma unroll
        for (int a = 0; a < MB; a++)
#pragma unroll
            for (int b = 0; b < MB; b++) blkS[a][b] = blkL[a][b];
        N = 2 * N;
    }
    if (st == ILA_OK) {
        sweep(0, N / 2 - 1, 0, j - 1, 2, 0, 0);
        A.N_used[inst] = N / 2;
    } else {
        const double qn = __longlong_as_double(0x7ff8000000000000LL);
        T qnan = zfrom<T>(qn);
        if (z_is_cplx<T>::value) qnan = zaddr(zmul(qnan, qnan), 0.0);   // NaN in both parts
        for (int64_t k = 0; k < j * NC; k++) Lband[k] = qnan;
        for (int64_t k = 0; k < j * U; k++) V[k] = qnan;
        for (int64_t k = 0; k < j; k++) tau_out[k] = qnan;
        A.N_used[inst] = N;
    }
    A.status[inst] = st;
}

template <typename T, int L, int U>
static int launch_fsql_lu(const FsqlArgsT<T> &a, cudaStream_t s)
{
    const unsigned blocks = (unsigned)((a.batch + 31) / 32);
    ql_finite_section_kernel<T, L, U><<<blocks, 32, 0, s>>>(a);
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <typename T>
static int launch_fsql_t(int l, int u, const FsqlArgsT<T> &a, cudaStream_t s)
{
    if (a.batch <= 0) return 0;
    if (l == 1 && u == 1) return launch_fsql_lu<T, 1, 1>(a, s);
    if (l == 1 && u == 2) return launch_fsql_lu<T, 1, 2>(a, s);
    if (l == 2 && u == 1) return launch_fsql_lu<T, 2, 1>(a, s);
    if (l == 2 && u == 2) return launch_fsql_lu<T, 2, 2>(a, s);
    if (l == 3 && u == 1) return launch_fsql_lu<T, 3, 1>(a, s);
    if (l == 1 && u == 3) return launch_fsql_lu<T, 1, 3>(a, s);
    return set_error(ILA_ERR_UNSUPPORTED, "ql_finite_section: bandwidths (l,u) not instantiated");
}

int launch_ql_finite_section(int l, int u, const FsqlArgs &a, cudaStream_t s) { return launch_fsql_t<double>(l, u, a, s); }
int launch_ql_finite_section_c(int l, int u, const FsqlArgsC &a, cudaStream_t s) { return launch_fsql_t<cplx>(l, u, a, s); }

} // namespace ila
