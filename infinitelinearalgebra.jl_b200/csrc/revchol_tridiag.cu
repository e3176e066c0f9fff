// revchol_tridiag.cu — K8: adaptive reverse Cholesky A = L'L of ∞ symmetric tridiagonal operators, one thread per
// operator (batched over shifts / parameters).
//
// Reference path replaced (InfiniteLinearAlgebra.jl v0.10.3, src/banded/infreversecholeskytridiagonal.jl):
//   reversecholesky_layout(::SymTridiagonalLayout, ...)   :66-74    tol = eps
//   __expand_factor!                                       :132-141  N := 2n, then while !converged && N <= MAX: N := 2N,
//   _finite_revchol!                                       :118-129  la[N] = sqrt(a[N]); lb[i] = b[i]/la[i+1]; la[i] = sqrt(a[i]-lb[i]^2)
//   compute_ξ / has_converged                              :81-110   ν = 1/la[N]; ν *= -(lb[i]/la[i]); ξ = Σ (...)²; scale·sqrt(ξ) <= eps
//   MAX_TRIDIAG_CHOL_N = 2^21 - 1                          :1
// The factor of size N and its convergence measure ξ come out of ONE backward sweep (ξ only needs running values), so
// nothing of size N is stored: the sweep keeps la[i+1], ν, ξ in registers and writes the leading n entries of L.  All
// arithmetic is IEEE and unfused in the reference's order, so la, lb and the N at which the test fires equal the
// oracle's (oracle/ila_oracle_revchol.c) bit for bit.
#include "ila_common.cuh"
#include "ila_kernels.cuh"
#include "ila_exact.cuh"

namespace ila {

constexpr int kMaxTridiagCholN = (1 << 21) - 1;


__global__ void __launch_bounds__(32) revchol_tridiag_kernel(RevCholArgs A)
{
    const int64_t inst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= A.batch) return;
    double ga[5], gb[5];
#pragma unroll
    for (int c = 0; c < 5; c++) { ga[c] = A.ga[inst * 5 + c]; gb[c] = A.gb[inst * 5 + c]; }
    const double shift = A.shift ? A.shift[inst] : 0.0;
    const int hp = A.hp;
    const int64_t n = A.n;
    double *la_out = A.la + inst * n, *lb_out = A.lb + inst * n;
    auto gen = [&](const double (&g)[5], int64_t k) -> double {
        const double num = __dadd_rn(g[1], __dmul_rn(g[2], (double)k));
        const double den = __dadd_rn(g[3], __dmul_rn(g[4], (double)k));
        return __dadd_rn(g[0], ediv(num, den));
    };
    auto av = [&](int64_t k) -> double { return __dsub_rn(k < hp ? A.a_head[k] : gen(ga, k), shift); };
    auto bv = [&](int64_t k) -> double { return k < hp ? A.b_head[k] : gen(gb, k); };
    const double eps = 2.220446049250313e-16;
    int64_t N = 2 * n;
    int32_t st = ILA_OK;
    for (;;) {
        if (N > kMaxTridiagCholN) { st = ILA_NOT_CONVERGED; break; }
        N *= 2;
        const double aN = av(N - 1);
        if (!(aN >= 0.0)) { st = ILA_NOT_POSDEF; break; }
        double la_next = esqrt(aN);
        double nu = ediv(1.0, la_next);
        double xi = 0.0;
        bool bad = false;
        for (int64_t i = N - 1; i >= 1; i--) {     // 1-based row i
            const double lb = ediv(bv(i - 1), la_next);
            const double t = __dsub_rn(av(i - 1), __dmul_rn(lb, lb));
            if (!(t >= 0.0)) { bad = true; break; }
            const double la = esqrt(t);
            const double ratio = -ediv(lb, la);
            if (i > n) {
                nu = __dmul_rn(nu, ratio);
            } else if (i == n) {
                nu = __dmul_rn(nu, ratio);
                const double w = __dmul_rn(la, nu);
                xi = __dmul_rn(w, w);
            } else {
                double xp = __dmul_rn(lb, nu);
                nu = __dmul_rn(nu, ratio);
                xp = __dadd_rn(xp, __dmul_rn(la, nu));
                xi = __dadd_rn(xi, __dmul_rn(xp, xp));
            }
            if (i <= n) { la_out[i - 1] = la; lb_out[i - 1] = lb; }
            la_next = la;
        }
        if (bad) { st = ILA_NOT_POSDEF; break; }
        const double bN = bv(N - 1);
        const double scale = (bN == 0.0) ? 1.0 : fabs(bN);
        if (__dmul_rn(scale, esqrt(xi)) <= eps) break;
    }
    if (st == ILA_NOT_POSDEF) {
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        for (int64_t k = 0; k < n; k++) { la_out[k] = qnan; lb_out[k] = qnan; }
    }
    A.N_used[inst] = N;
    A.status[inst] = st;
}

int launch_revchol_tridiag(const RevCholArgs &a, cudaStream_t s)
{
    if (a.batch <= 0) return 0;
    const unsigned blocks = (unsigned)((a.batch + 31) / 32);
    revchol_tridiag_kernel<<<blocks, 32, 0, s>>>(a);
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace ila
