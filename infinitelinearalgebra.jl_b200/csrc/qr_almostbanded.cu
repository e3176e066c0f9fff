// qr_almostbanded.cu — K12: adaptive QR solve of ∞ ALMOST-BANDED operators A = Vcat(B, D) (r dense boundary rows on top
// of a banded operator), one thread per operator, batched over shifts / parameters (second half).
//
// Reference path replaced (InfiniteLinearAlgebra.jl v0.10.3):
//   AdaptiveQRData(::AbstractAlmostBandedLayout, A)        src/infqr.jl:15-21   AlmostBandedMatrix, bandwidths (l, l+u), rank r
//   partialqr!(F, n) -> _almostbanded_qr!                  src/infqr.jl:50-66   [SemiseparableMatrices]
//   Q'b adaptive driver (COLGROWTH = 1000, tol floatmin)   src/infqr.jl:236-274
//   ldiv!(R, B)                                            src/infqr.jl:350-355
// With l = lD + r, u = uD - r the Householder reflector of column k acts on rows k..k+l.  Above the band the working rows
// are combinations of the boundary rows: row i beyond column k + l + u equals sum_m C[i][m] B_m(j), and a reflector maps
// the coefficient vectors C[i][.] exactly like the explicit entries.  So the state of the sweep is an (l+1) x (l+u+1)
// explicit window, (l+1) x r coefficients and the (l+1) window of Q'b — nothing of size n per column but one record
//   [R[k,k..k+l+u], C_k[0..r-1], (Q'b)_k]
// streamed to memory; the back-substitution keeps the r tail sums S_m = sum_{j beyond the window} B_m(j) x_j running, so the
// rank-r fill costs O(r) per row instead of O(n).  Convergence is tested on device at the reference's cadence; after the
// test passes the sweep factorizes l more columns (the rows that still carry the last, sub-tolerance entries of Q'b) and
// the solution is zero beyond them.  The common shapes are compiled with (l, W, r) as template constants (windows in registers),
// any other supported shape runs on small local arrays; plain IEEE arithmetic (parity
// bar 1e-12: the reference's kernel lives in an un-vendored dependency, its operation order is not pinned).
#include "ila_common.cuh"
#include "ila_kernels.cuh"

namespace ila {

namespace {
constexpr int kAbMaxL = 4;      // l = lD + r
constexpr int kAbMaxW = 12;     // l + u + 1
constexpr int kAbMaxR = 2;

__device__ __forceinline__ double ab_boundary(const double *bg, int64_t j)
{   // B_m(j) = s^j (a + b j + c j^2), s = +-1
    const double jj = (double)j;
    const double v = bg[1] + bg[2] * jj + bg[3] * jj * jj;
    return (bg[0] < 0.0 && (j & 1)) ? -v : v;
}
} // namespace

// TL, TW, TR: compile-time l, window width and r of the common shapes (windows in registers, loops unrolled); 0 = runtime
// values on local arrays (any supported shape)
template <int TL, int TW, int TR>
__global__ void __launch_bounds__(32) qr_almostbanded_kernel(AlmostBandedArgs A)
{
    const int64_t inst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= A.batch) return;
    const int lD = A.lD, uD = A.uD, r = TR ? TR : A.r, l = TL ? TL : lD + r, W = TW ? TW : lD + uD + 1, nd = lD + uD + 1;
    const double *dg = A.dg + inst * nd * 3;
    const double *bg = A.bg + inst * r * 4;
    const double shift = A.shift ? A.shift[inst] : 0.0;
    const double *bi = A.b + inst * A.nb;
    const int nb = A.nb, cg = A.colgrowth;
    const int64_t ncap = A.ncap;
    const int REC = W + r + 1;
    // records of the 32 operators of a warp interleaved: element e of record k of lane ℓ at ((k*REC + e)*32 + ℓ), so that the
    // forward sweep stores and the back-substitution loads 256 contiguous bytes per element (same idea as K1's layout)
    double *rec = A.work + (int64_t)blockIdx.x * 32 * A.work_stride + threadIdx.x;
    double *xo = A.x + inst * ncap;

    // A[R, j] of the banded part (R >= r): row t = R - r of D, diagonal d = j - t
    auto band = [&](int64_t R, int64_t j) -> double {
        const int64_t t = R - r, d = j - t;
        if (t < 0 || j < 0 || d > uD || d < -lD) return 0.0;
        const double *g = dg + (uD - d) * 3;
        const double tt = (double)t;
        double v = g[0] + g[1] * tt + g[2] * tt * tt;
        if (d == 0) v -= shift;
        return v;
    };

    double E[(TL ? TL : kAbMaxL) + 1][TW ? TW : kAbMaxW], C[(TL ? TL : kAbMaxL) + 1][TR ? TR : kAbMaxR], bq[(TL ? TL : kAbMaxL) + 1],
        v[(TL ? TL : kAbMaxL) + 1];
    auto init_row = [&](int i, int64_t R, int64_t k) {       // window row i <- global row R, window columns k..k+W-1
#pragma unroll
        for (int m = 0; m < r; m++) C[i][m] = (R == m) ? 1.0 : 0.0;
#pragma unroll
        for (int c = 0; c < W; c++) E[i][c] = (R < r) ? ab_boundary(bg + R * 4, k + c) : band(R, k + c);
        bq[i] = R < nb ? bi[R] : 0.0;
    };
#pragma unroll
    for (int i = 0; i <= l; i++) init_row(i, i, 0);

    int32_t st = ILA_OK;
    int64_t nconv = 0, kfinal = 0;          // kfinal: number of final R rows
    int64_t jfirst = 0;
    int64_t kstop = -1;                     // after convergence: factorize (without touching b) up to column kstop
    bool converged = false;
    for (int64_t k = 0;; k++) {
        if (!converged && k == jfirst) {
            const int64_t cs_max = jfirst + cg - 1 + l;
            if (cs_max + 1 + W + 1 > ncap) { st = ILA_NOT_CONVERGED; break; }
            if (jfirst + 1 > nb) {
                double mx = 0.0;
                bool isnan_ = false;
#pragma unroll
                for (int i = 0; i <= l; i++) { const double a = fabs(bq[i]); isnan_ |= (a != a); mx = fmax(mx, a); }
                if (isnan_) { st = ILA_NAN; break; }
                if (mx <= A.tol) { converged = true; nconv = jfirst + 1; kstop = jfirst + l - 1; }
            }
            jfirst += cg;
        }
        if (converged && k > kstop) { kfinal = k; break; }
        // ---- Householder reflector of column k over window rows 0..l
        const double alpha = E[0][0];
        double ss = 0.0;
#pragma unroll
        for (int i = 1; i <= l; i++) ss += E[i][0] * E[i][0];
        double tau = 0.0, beta = alpha;
        if (ss != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + ss), alpha);
            tau = (beta - alpha) / beta;
            const double inv = 1.0 / (alpha - beta);
#pragma unroll
            for (int i = 1; i <= l; i++) v[i] = E[i][0] * inv;
#pragma unroll
            for (int c = 1; c < W; c++) {
                double w = E[0][c];
#pragma unroll
                for (int i = 1; i <= l; i++) w += v[i] * E[i][c];
                w *= tau;
                E[0][c] -= w;
#pragma unroll
                for (int i = 1; i <= l; i++) E[i][c] -= v[i] * w;
            }
#pragma unroll
            for (int m = 0; m < r; m++) {
                double w = C[0][m];
#pragma unroll
                for (int i = 1; i <= l; i++) w += v[i] * C[i][m];
                w *= tau;
                C[0][m] -= w;
#pragma unroll
                for (int i = 1; i <= l; i++) C[i][m] -= v[i] * w;
            }
            if (!converged) {
                double w = bq[0];
#pragma unroll
                for (int i = 1; i <= l; i++) w += v[i] * bq[i];
                w *= tau;
                bq[0] -= w;
#pragma unroll
                for (int i = 1; i <= l; i++) bq[i] -= v[i] * w;
            }
        }
        // ---- row k is final
        {
            double *dst = rec + k * REC * 32;
            dst[0] = beta;
#pragma unroll
            for (int c = 1; c < W; c++) dst[c * 32] = E[0][c];
#pragma unroll
            for (int m = 0; m < r; m++) dst[(W + m) * 32] = C[0][m];
            dst[(W + r) * 32] = bq[0];
        }
        // ---- slide the window: rows up, columns left; the entering column is pure fill, the entering row pure band
#pragma unroll
        for (int i = 0; i < l; i++) {
#pragma unroll
            for (int c = 0; c + 1 < W; c++) E[i][c] = E[i + 1][c + 1];
#pragma unroll
            for (int m = 0; m < r; m++) C[i][m] = C[i + 1][m];
            bq[i] = bq[i + 1];
            double f = 0.0;
#pragma unroll
            for (int m = 0; m < r; m++) f += C[i][m] * ab_boundary(bg + m * 4, k + W);
            E[i][W - 1] = f;
        }
        init_row(l, k + 1 + l, k + 1);
    }

    int64_t n = 0;
    if (st == ILA_OK) {
        n = nconv - 1 + cg + l;                    // the reference's datasize
        // ---- back-substitution over the kfinal final rows; x is zero beyond them
        double S[kAbMaxR] = {0.0, 0.0};
        double xw[TW ? TW : kAbMaxW];               // xw[c] = x[k + 1 + c]
#pragma unroll
        for (int c = 0; c < W; c++) xw[c] = 0.0;
        bool any = false;
        for (int64_t k = kfinal - 1; k >= 0; k--) {
            const double *src = rec + k * REC * 32;
            // column k + W leaves the window of row k+1 into the tail of row k
            const int64_t jout = k + W;
            if (jout < kfinal) {
                const double xj = xw[W - 1];
#pragma unroll
                for (int m = 0; m < r; m++) S[m] += ab_boundary(bg + m * 4, jout) * xj;
            }
            double acc = src[(W + r) * 32];
            any |= fabs(acc) > A.tol;
#pragma unroll
            for (int c = 1; c < W; c++) acc -= src[c * 32] * xw[c - 1];
#pragma unroll
            for (int m = 0; m < r; m++) acc -= src[(W + m) * 32] * S[m];
            const double xk = acc / src[0];
#pragma unroll
            for (int c = W - 1; c >= 1; c--) xw[c] = xw[c - 1];
            xw[0] = xk;
            xo[k] = xk;
        }
        if (!any) {                                 // resizedata_chop!: every entry of Q'b at or below the tolerance
            n = 0;
            for (int64_t i = 0; i < kfinal; i++) xo[i] = 0.0;
        }
    }                                               // x was zero-filled by the launcher: nothing to write on failure
    A.ncols[inst] = n;
    A.nconv[inst] = nconv;
    A.status[inst] = st;
}

int launch_qr_almostbanded(const AlmostBandedArgs &a, cudaStream_t s)
{
    if (a.batch <= 0) return 0;
    const int l = a.lD + a.r, W = a.lD + a.uD + 1;
    if (a.r < 1 || a.r > kAbMaxR || l > kAbMaxL || W > kAbMaxW || a.uD - a.r + l < 0 || a.lD < 0 || a.uD < 0)
        return set_error(ILA_ERR_UNSUPPORTED, "qr_almostbanded: need 1 <= r <= 2, lD + r <= 4, lD + uD + 1 <= 12");
    const unsigned blocks = (unsigned)((a.batch + 31) / 32);
    ILA_CUDA_CHECK(cudaMemsetAsync(a.x, 0, sizeof(double) * (size_t)a.batch * (size_t)a.ncap, s));     // zero padding of the solutions
    if (l == 1 && W == 2 && a.r == 1) qr_almostbanded_kernel<1, 2, 1><<<blocks, 32, 0, s>>>(a);            // test_infqr.jl:194 (one band)
    else if (l == 2 && W == 3 && a.r == 2) qr_almostbanded_kernel<2, 3, 2><<<blocks, 32, 0, s>>>(a);       // :223 (two bands)
    else if (l == 2 && W == 3 && a.r == 1) qr_almostbanded_kernel<2, 3, 1><<<blocks, 32, 0, s>>>(a);
    else if (l == 3 && W == 4 && a.r == 2) qr_almostbanded_kernel<3, 4, 2><<<blocks, 32, 0, s>>>(a);
    else if (l == 3 && W == 7 && a.r == 2) qr_almostbanded_kernel<3, 7, 2><<<blocks, 32, 0, s>>>(a);       // :250 (more bands)
    else qr_almostbanded_kernel<0, 0, 0><<<blocks, 32, 0, s>>>(a);
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace ila
