// triconj.cu — K10: tridiagonal conjugation Y = Tridiagonal(U*X/U) (and the bidiagonal conjugation), the consumers that
// read the R / U band data of the adaptive factorizations directly.
//
// Reference path replaced (InfiniteLinearAlgebra.jl v0.10.3):
//   upper_mul_tri_triview! (initiate / main / finalize)        src/banded/tridiagonalconjugation.jl:10-111
//   tri_mul_invupper_triview! (initiate / main / finalize)     src/banded/tridiagonalconjugation.jl:114-215
//   TridiagonalConjugationData / resizedata!                   src/banded/tridiagonalconjugation.jl:224-282
//   BidiagonalConjugationData, _compute_column_up!/_lo!        src/banded/bidiagonalconjugation.jl:19-60
//
// The reference runs two sequential in-place sweeps (row k finishes Ydu[k-1]).  Unrolling the in-place updates gives a
// closed form per row that only needs rows k-1, k, k+1 of U and X:
//     a = U[k,k] X[k,k] + U[k,k+1] X[k+1,k]                                   (UXd[k])
//     b = (U[k,k] X[k,k+1] + U[k,k+1] X[k+1,k+1]) + U[k,k+2] X[k+2,k+1]       (UXdu[k])
//     c = U[k,k] X[k,k-1]                                                     (UXdl[k-1])
//     Y[k+1,k] = (U[k+1,k+1] X[k+1,k]) / U[k,k]
//     Y[k,k]   = (a - (c U[k-1,k]) / U[k-1,k-1]) / U[k,k]                      (k = 1: a / U[1,1])
//     Y[k,k+1] = ((c / U[k-1,k-1]) ((U[k-1,k] U[k,k+1]) / U[k,k] - U[k-1,k+1]) + (b - (a U[k,k+1]) / U[k,k])) / U[k+1,k+1]
// with every product, sum and quotient a separate IEEE operation in the reference's order, so the result equals the
// sequential oracle (oracle/ila_oracle_triconj.c) bit for bit — and every row is independent: one thread per (operator,
// row), no recurrence, an HBM-streaming kernel (48-72 B per row against 7 quotients over 3 denominators, the refined
// reciprocal of a denominator shared between its quotients).
//
// Two sources for U: (0) explicit band arrays, operator-major (U[i][d][k], coalesced along k); (1) the group-interleaved
// factor records K1 / K5 leave in device memory (record k of lane ℓ of group g: chunk c at double2 index
// ((gbase[g]+k)*nch + c)*32 + ℓ, first entries R[k,k..k+nbands-1]) — the conjugation then runs straight behind the
// factorization with no export pass; the 32 lanes of a warp read one record row (512 contiguous bytes per chunk) and the
// block transposes its 32x32 tile of Y through shared memory so that the stores are 256-byte runs along k.
#include "ila_common.cuh"
#include "ila_kernels.cuh"
#include "ila_unfused.cuh"

namespace ila {

struct TcRow {
    double dl, d, du, uxdl, uxd, uxdu;
};

// a / b with the refined reciprocal of b shared between the quotients of one row (same bits as ediv: div_refined takes the
// reciprocal as an argument; outside the exponent boxes the IEEE intrinsic)
struct Den {
    double b, r;
    bool ok;
    __device__ __forceinline__ explicit Den(double b_) : b(b_), r(rcp_refined(b_)), ok(inbox_den(b_)) {}
    __device__ __forceinline__ double div(double a) const
    {
        if (ok && (inbox(a) || a == 0.0)) return div_refined(a, b, r);
        return __ddiv_rn(a, b);
    }
};

// closed-form row k (1-based) from accessors R0,R1,R2 (rows of U) and the X bands; 7 quotients over 3 denominators
template <class FR0, class FR1, class FR2, class FX0, class FX1, class FX2>
__device__ __forceinline__ TcRow triconj_row(int64_t k, FR0 R0, FR1 R1, FR2 R2, FX0 Xdl, FX1 Xd, FX2 Xdu)
{
    const double Ukk = R0(k), Uk1 = R1(k), Uk2 = R2(k), Unn = R0(k + 1);
    const double xdl_k = Xdl(k), xd_k = Xd(k);
    const double a = zadd(zmul(Ukk, xd_k), zmul(Uk1, xdl_k));
    const double b = zadd(zadd(zmul(Ukk, Xdu(k)), zmul(Uk1, Xd(k + 1))), zmul(Uk2, Xdl(k + 1)));
    const double cn = zmul(Unn, xdl_k);
    const Den dkk(Ukk), dnn(Unn);
    TcRow r;
    r.uxdl = cn; r.uxd = a; r.uxdu = b;
    r.dl = dkk.div(cn);
    const double tail = zsub(b, dkk.div(zmul(a, Uk1)));
    double pre;
    if (k == 1) {
        r.d = dkk.div(a);
        pre = tail;
    } else {
        const double Rp = R0(k - 1), Rpn = R1(k - 1), Rpnn = R2(k - 1);
        const Den dp(Rp);
        const double c = zmul(Ukk, Xdl(k - 1));
        const double t = dp.div(c);
        r.d = dkk.div(zsub(a, dp.div(zmul(c, Rpn))));
        pre = zadd(zmul(t, zsub(dkk.div(zmul(Rpn, Uk1)), Rpnn)), tail);
    }
    r.du = dnn.div(pre);
    return r;
}

// ---- source 0: explicit bands, one thread per (operator, row) -----------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256) triconj_bands_kernel(TriConjArgs A)
{
    const int64_t i = blockIdx.y;
    const int64_t n = A.n;
    const double *U = A.U + i * (int64_t)NB * A.uld;
    const double *X = A.X + i * A.xstride;
    const int64_t uld = A.uld, xld = A.xld;
    auto R0 = [&](int64_t k) { return __ldg(U + k - 1); };
    auto R1 = [&](int64_t k) { return __ldg(U + uld + k - 1); };
    auto R2 = [&](int64_t k) { return NB > 2 ? __ldg(U + 2 * uld + k - 1) : 0.0; };
    auto Xdl = [&](int64_t k) { return __ldg(X + k - 1); };
    auto Xd = [&](int64_t k) { return __ldg(X + xld + k - 1); };
    auto Xdu = [&](int64_t k) { return __ldg(X + 2 * xld + k - 1); };
    double *Y = A.Y + i * 3 * n;
    double *UX = A.UX ? A.UX + i * 3 * n : nullptr;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; k <= n; k += (int64_t)gridDim.x * blockDim.x) {
        const TcRow r = triconj_row(k, R0, R1, R2, Xdl, Xd, Xdu);
        __stcs(Y + k - 1, r.dl); __stcs(Y + n + k - 1, r.d); __stcs(Y + 2 * n + k - 1, r.du);
        if (UX) { __stcs(UX + k - 1, r.uxdl); __stcs(UX + n + k - 1, r.uxd); __stcs(UX + 2 * n + k - 1, r.uxdu); }
    }
}

// ---- source 1: K1 / K5 factor records, block = 32 lanes (operators of one group) x 32 rows ------------------------
template <int NB>
__global__ void __launch_bounds__(1024) triconj_records_kernel(TriConjArgs A)
{
    __shared__ double tile[3][32][33];
    const int lane = threadIdx.x, ty = threadIdx.y;
    const int64_t g = blockIdx.y;
    const int64_t inst = g * 32 + lane;
    const int64_t n = A.n;
    const int64_t k = (int64_t)blockIdx.x * 32 + ty + 1;     // 1-based row
    const int nch = A.nch;
    const double2 *rec = A.work + A.gbase[g] * nch * 32 + lane;
    // element e of record row (0-based): chunk e/2, .x or .y
    auto elem = [&](int64_t row, int e) -> double {
        const double2 v = __ldg(rec + (row * nch + (e >> 1)) * 32);
        return (e & 1) ? v.y : v.x;
    };
    const bool live = inst < A.batch && k <= n && (!A.status || A.status[inst] == ILA_OK);
    TcRow r{};
    if (live) {
        const double *X = A.X + inst * A.xstride;
        const int64_t xld = A.xld;
        auto R0 = [&](int64_t kk) { return elem(kk - 1, 0); };
        auto R1 = [&](int64_t kk) { return elem(kk - 1, 1); };
        auto R2 = [&](int64_t kk) { return NB > 2 ? elem(kk - 1, 2) : 0.0; };
        auto Xdl = [&](int64_t kk) { return __ldg(X + kk - 1); };
        auto Xd = [&](int64_t kk) { return __ldg(X + xld + kk - 1); };
        auto Xdu = [&](int64_t kk) { return __ldg(X + 2 * xld + kk - 1); };
        r = triconj_row(k, R0, R1, R2, Xdl, Xd, Xdu);
    } else if (inst < A.batch && k <= n) {
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        r.dl = r.d = r.du = qnan;
    }
    tile[0][ty][lane] = r.dl; tile[1][ty][lane] = r.d; tile[2][ty][lane] = r.du;
    __syncthreads();
    // transposed write-out: thread (lane, ty) stores row k0 + lane of operator g*32 + ty
    const int64_t oi = g * 32 + ty, ok = (int64_t)blockIdx.x * 32 + lane;
    if (oi < A.batch && ok < n) {
        double *Y = A.Y + oi * 3 * n;
        __stcs(Y + ok, tile[0][lane][ty]); __stcs(Y + n + ok, tile[1][lane][ty]); __stcs(Y + 2 * n + ok, tile[2][lane][ty]);
    }
}

int launch_triconj(const TriConjArgs &a, cudaStream_t s)
{
    if (a.batch <= 0 || a.n <= 0) return 0;
    if (a.nbands != 2 && a.nbands != 3) return set_error(ILA_ERR_UNSUPPORTED, "triconj: nbands must be 2 or 3");
    if (a.work) {
        const dim3 block(32, 32), grid((unsigned)((a.n + 31) / 32), (unsigned)((a.batch + 31) / 32));
        if (grid.y > 65535u) return set_error(ILA_ERR_UNSUPPORTED, "triconj: at most 65535*32 operators per call");
        if (a.nbands == 3) triconj_records_kernel<3><<<grid, block, 0, s>>>(a);
        else triconj_records_kernel<2><<<grid, block, 0, s>>>(a);
    } else {
        if (a.batch > 65535) return set_error(ILA_ERR_UNSUPPORTED, "triconj: at most 65535 operators per call");
        // rows split so that the grid is a few waves of 148 SMs x 8 resident blocks
        int64_t bx = (a.n + 255) / 256;
        const int64_t want = (148 * 8 * 4 + a.batch - 1) / a.batch;
        if (bx > want) bx = want;
        if (bx < 1) bx = 1;
        const dim3 grid((unsigned)bx, (unsigned)a.batch);
        if (a.nbands == 3) triconj_bands_kernel<3><<<grid, 256, 0, s>>>(a);
        else triconj_bands_kernel<2><<<grid, 256, 0, s>>>(a);
    }
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ---- bidiagonal conjugation: one thread per (operator, column) ----------------------------------------------------
__global__ void __launch_bounds__(256) biconj_kernel(BiConjArgs A)
{
    const int64_t b = blockIdx.y, n = A.n, ld = A.ld;
    const double *U = A.U + b * 3 * ld, *Cm = A.C + b * 3 * ld;
    double *dv = A.dv + b * n, *ev = A.ev + b * n;
    auto UL = [&](int64_t k) { return __ldg(U + k - 1); };
    auto UD = [&](int64_t k) { return __ldg(U + ld + k - 1); };
    auto UU = [&](int64_t k) { return __ldg(U + 2 * ld + k - 1); };
    auto CL = [&](int64_t k) { return __ldg(Cm + k - 1); };
    auto CD = [&](int64_t k) { return __ldg(Cm + ld + k - 1); };
    auto CU = [&](int64_t k) { return __ldg(Cm + 2 * ld + k - 1); };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i <= n + 1; i += (int64_t)gridDim.x * blockDim.x) {
        if (A.upper) {
            if (i == 1) { dv[0] = ediv(CD(1), UD(1)); continue; }
            const double u11 = UD(i - 1), u21 = UL(i - 1), u12 = UU(i - 1), u22 = UD(i);
            const double c12 = CU(i - 1), c22 = CD(i);
            const double inv = ediv(1.0, zsub(zmul(u11, u22), zmul(u12, u21)));
            if (i <= n) dv[i - 1] = zmul(inv, zsub(zmul(u11, c22), zmul(u21, c12)));
            ev[i - 2] = zmul(inv, zsub(zmul(u22, c12), zmul(u12, c22)));
        } else if (i <= n) {
            const double u11 = UD(i), u21 = UL(i), u12 = UU(i), u22 = UD(i + 1);
            const double c11 = CD(i), c21 = CL(i);
            const double inv = ediv(1.0, zsub(zmul(u11, u22), zmul(u12, u21)));
            dv[i - 1] = zmul(inv, zsub(zmul(u22, c11), zmul(u12, c21)));
            ev[i - 1] = zmul(inv, zsub(zmul(u11, c21), zmul(u21, c11)));
        }
    }
}

int launch_biconj(const BiConjArgs &a, cudaStream_t s)
{
    if (a.batch <= 0 || a.n <= 0) return 0;
    if (a.batch > 65535) return set_error(ILA_ERR_UNSUPPORTED, "biconj: at most 65535 operators per call");
    int64_t bx = (a.n + 1 + 255) / 256;
    const int64_t want = (148 * 8 * 4 + a.batch - 1) / a.batch;
    if (bx > want) bx = want;
    if (bx < 1) bx = 1;
    biconj_kernel<<<dim3((unsigned)bx, (unsigned)a.batch), 256, 0, s>>>(a);
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace ila
