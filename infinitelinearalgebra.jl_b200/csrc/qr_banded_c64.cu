// qr_banded_c64.cu — K1 + K2 for a COMPLEX eltype: adaptive banded Householder QR sweep, Q'b, convergence test and the
// banded back-substitution for A \ b with A an ∞ banded operator with ComplexF64 entries (e.g. a real operator minus a
// complex shift, the resolvent set-up `(A - λI) \ b`).
//
// Reference path replaced (InfiniteLinearAlgebra.jl v0.10.3), same as qr_banded.cu:
//   AdaptiveQRData banded ctor src/infqr.jl:8-13; partialqr! :32-48 (-> BandedMatrices._banded_qr!); the adaptive Q'b
//   driver :236-274 (COLGROWTH chunks, `maximum(abs, b[j:j+l]) <= tolerance`); resizedata_chop! :207-219; ldiv!(R,B)
//   :350-355 — with LinearAlgebra.reflector!/reflectorApply! in their complex form (ξ1 += copysign(‖x‖, real ξ1),
//   x[2:end] ./= ξ1, τ = ξ1/ν;  vA = A[1,j] + Σ conj(x_i) A[i,j],  vA = conj(τ) vA).
//
// Design: the real kernels' structure without their branch-free exact-division machinery (complex quotients are
// Smith's algorithm on IEEE divisions; every product is the four-multiply form, nothing is fused), one thread per
// instance, 32 instances in lock step, window in registers, records [R[k,k..k+l+u], (Q'b)_k] — one complex = one
// 16-byte chunk — in the same group-interleaved layout ((gbase[g]+k)*REC + c)*32 + lane, so the sweep stores and the
// back-substitution loads 512 contiguous bytes per chunk.  The oracle (oracle/ila_oracle_c64.c) issues the same
// operations in the same order: results are compared bit for bit.
#include "ila_common.cuh"
#include "ila_kernels.cuh"
#include "ila_exact.cuh"
#include "ila_unfused.cuh"

namespace ila {


template <int L, int U>
__global__ void __launch_bounds__(32) qr_forward_c64_kernel(QrC64FwdArgs A)
{
    const int64_t inst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= A.batch) return;
    constexpr int UP = L + U, NC = UP + 1, NR = L + 1, ND = L + U + 1;
    constexpr int REC = NC + 1;   // complex scalars (16-byte chunks) per record
    cplx gp[ND], gq[ND];
    double gr[ND];
    const cplx shift = A.shift ? A.shift[inst] : mk(0.0, 0.0);
#pragma unroll
    for (int rr = 0; rr < ND; rr++) {
        gp[rr] = A.p[inst * ND + rr];
        gq[rr] = A.q[inst * ND + rr];
        gr[rr] = A.r ? A.r[inst * ND + rr] : 1.0;
    }
    const cplx *bi = A.b + inst * A.nb;
    const int lane = threadIdx.x;
    const int ncap = (int)(A.ncap_per ? A.ncap_per[inst] : A.ncap);
    double2 *rec = A.work + A.gbase[blockIdx.x] * REC * 32 + lane;
    const int cg = A.colgrowth, hp = A.hp, nb = A.nb;

    auto entry = [&](int row, int col, int rr) -> cplx {
        const int t = (col >= row) ? row : col;
        cplx v;
        if (t < hp) v = A.head[(int64_t)rr * hp + t];
        else {
            const cplx qt = mk(__dmul_rn(gq[rr].re, (double)t), __dmul_rn(gq[rr].im, (double)t));
            v = zadd(gp[rr], qt);
            if (gr[rr] != 1.0) v = mk(ediv(v.re, gr[rr]), ediv(v.im, gr[rr]));
        }
        if (rr == U) v = zsub(v, shift);
        return v;
    };

    // window W[j][i] = A[k+i, k+j], j = 0..UP, i = 0..L  (zero where col-row > u: R's fill-in region)
    cplx W[NC][NR], bw[NR];
#pragma unroll
    for (int j = 0; j < NC; j++)
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const int d = j - i;
            W[j][i] = (d <= U) ? entry(i, j, U - d) : mk(0.0, 0.0);
        }
#pragma unroll
    for (int i = 0; i < NR; i++) bw[i] = (i < nb) ? bi[i] : mk(0.0, 0.0);

    int32_t st = ILA_OK;
    int nconv = 0, n = nb, kstop = -1;
    bool converged = false, anybig = false;
    int next_chunk = 0;
    int k = 0;
    for (;;) {
        if (k == next_chunk) {   // first(jr) of a new COLGROWTH chunk (infqr.jl:257-266)
            const int jlast = k + cg - 1, cs_max = jlast + L;
            if (!converged) {
                if (cs_max + 1 + UP + 1 > ncap) { st = ILA_NOT_CONVERGED; break; }
                if (cs_max + 1 > n) n = cs_max + 1;
                if (k + 1 > nb) {
                    double mx = 0.0;
                    bool nan = false;
#pragma unroll
                    for (int i = 0; i < NR; i++) {
                        const double a = esqrt(zabs2(bw[i]));
                        nan |= (a != a);
                        mx = (a > mx) ? a : mx;
                    }
                    if (nan) { st = ILA_NAN; nconv = k + 1; break; }
                    if (mx <= A.tol) {
                        converged = true;
                        nconv = k + 1;
                        kstop = k + L;   // rows beyond k+L carry (Q'b) == 0: not needed for the solve
                    }
                }
            }
            next_chunk += cg;
        }
        int kend = next_chunk;
        if (converged && kstop + 1 < kend) kend = kstop + 1;
        const bool apply_b = !converged;
        double2 *dst = rec + (int64_t)k * REC * 32;
        for (; k < kend; k++, dst += REC * 32) {
            // reflector!(x = W[0][:])
            cplx tau = mk(0.0, 0.0);
            {
                double s = zabs2(W[0][0]);
#pragma unroll
                for (int i = 1; i < NR; i++) s = __dadd_rn(s, zabs2(W[0][i]));
                const double normu = esqrt(s);
                if (normu != 0.0) {
                    cplx xi1 = W[0][0];
                    const double nu = copysign(normu, xi1.re);
                    xi1.re = __dadd_rn(xi1.re, nu);
                    W[0][0] = mk(-nu, 0.0);
#pragma unroll
                    for (int i = 1; i < NR; i++) W[0][i] = zdiv(W[0][i], xi1);
                    tau = mk(ediv(xi1.re, nu), ediv(xi1.im, nu));
                }
            }
            const cplx tauc = zconj(tau);
            // reflectorApply! on the next l+u columns, then on Q'b
#pragma unroll
            for (int j = 1; j < NC; j++) {
                cplx vA = W[j][0];
#pragma unroll
                for (int i = 1; i < NR; i++) vA = zadd(vA, zmul(zconj(W[0][i]), W[j][i]));
                vA = zmul(tauc, vA);
                W[j][0] = zsub(W[j][0], vA);
#pragma unroll
                for (int i = 1; i < NR; i++) W[j][i] = zsub(W[j][i], zmul(W[0][i], vA));
            }
            if (apply_b) {
                cplx vB = bw[0];
#pragma unroll
                for (int i = 1; i < NR; i++) vB = zadd(vB, zmul(zconj(W[0][i]), bw[i]));
                vB = zmul(tauc, vB);
                bw[0] = zsub(bw[0], vB);
#pragma unroll
                for (int i = 1; i < NR; i++) bw[i] = zsub(bw[i], zmul(W[0][i], vB));
            }
            anybig |= esqrt(zabs2(bw[0])) > A.tol;
            // stream out row k of R and (Q'b)_k
#pragma unroll
            for (int j = 0; j < NC; j++) dst[j * 32] = make_double2(W[j][0].re, W[j][0].im);
            dst[NC * 32] = make_double2(bw[0].re, bw[0].im);
            // slide the window
#pragma unroll
            for (int j = 0; j < UP; j++)
#pragma unroll
                for (int i = 0; i < L; i++) W[j][i] = W[j + 1][i + 1];
#pragma unroll
            for (int i = 0; i < L; i++) W[UP][i] = mk(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < L; i++) bw[i] = bw[i + 1];
            const int row = k + 1 + L;
#pragma unroll
            for (int j = 0; j < NC; j++) W[j][L] = entry(row, k + 1 + j, UP - j);
            bw[L] = (row < nb) ? bi[row] : mk(0.0, 0.0);
        }
        if (converged && k > kstop) break;
    }
    int nsw = 0;
    if (st == ILA_OK) {
        if (!anybig) n = 0;   // resizedata_chop!: every entry <= tol -> datasize 0
        nsw = (n == 0) ? 0 : kstop + 1;
    } else {
        n = 0;
    }
    A.ncols[inst] = n;
    A.nconv[inst] = nconv;
    A.nswept[inst] = nsw;
    A.status[inst] = st;
}

// tbsv 'U','N','N' order per row: b_r -= x_j R[r,j] for j = r+UP down to r+1, then x_r = b_r / R[r,r]; x is written
// straight into the ragged instance-major output (zero between nswept and ncols).
template <int L, int U>
__global__ void __launch_bounds__(32) qr_backsolve_c64_kernel(int64_t batch, const int64_t *__restrict__ gbase,
                                                              const double2 *__restrict__ work,
                                                              const int64_t *__restrict__ ncols,
                                                              const int64_t *__restrict__ nswept,
                                                              const int64_t *__restrict__ xoff, double2 *__restrict__ x)
{
    const int64_t inst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= batch) return;
    constexpr int UP = L + U, NC = UP + 1, REC = NC + 1;
    const int lane = threadIdx.x;
    const int64_t nc = ncols[inst], nsw = nc > 0 ? nswept[inst] : 0;
    const double2 *rec = work + gbase[blockIdx.x] * REC * 32 + lane;
    double2 *xo = x + xoff[inst];
    for (int64_t r = nsw; r < nc; r++) xo[r] = make_double2(0.0, 0.0);
    cplx xw[UP + 1];
#pragma unroll
    for (int j = 0; j <= UP; j++) xw[j] = mk(0.0, 0.0);
    // records are fetched PF rows ahead of their use (register ring): the DRAM latency of a row overlaps the
    // recurrence of the PF rows before it
    constexpr int PF = 4;
    double2 ring[PF][REC];
#pragma unroll
    for (int p = 0; p < PF; p++) {
        const int64_t rr = nsw - 1 - p;
#pragma unroll
        for (int c = 0; c < REC; c++) ring[p][c] = (rr >= 0) ? rec[(rr * REC + c) * 32] : make_double2(0.0, 0.0);
    }
    for (int64_t r0 = nsw - 1; r0 >= 0; r0 -= PF) {
#pragma unroll
        for (int p = 0; p < PF; p++) {
            const int64_t r = r0 - p;
            if (r < 0) break;
            cplx v[REC];
#pragma unroll
            for (int c = 0; c < REC; c++) v[c] = mk(ring[p][c].x, ring[p][c].y);
            const int64_t rn = r - PF;        // refill this slot with the row PF below
#pragma unroll
            for (int c = 0; c < REC; c++) ring[p][c] = (rn >= 0) ? rec[(rn * REC + c) * 32] : make_double2(0.0, 0.0);
            cplx t = v[NC];
#pragma unroll
            for (int j = UP; j >= 1; j--) t = zsub(t, zmul(xw[j], v[j]));
            const cplx xr = zdiv(t, v[0]);
            xo[r] = make_double2(xr.re, xr.im);
#pragma unroll
            for (int j = UP; j >= 2; j--) xw[j] = xw[j - 1];
            xw[1] = xr;
        }
    }
}

template <int L, int U>
static int launch_c64_lu(const QrC64FwdArgs &a, cudaStream_t s)
{
    const unsigned blocks = (unsigned)((a.batch + 31) / 32);
    qr_forward_c64_kernel<L, U><<<blocks, 32, 0, s>>>(a);
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}
template <int L, int U>
static int launch_c64_bwd_lu(int64_t batch, const int64_t *gbase, const double2 *work, const int64_t *ncols,
                             const int64_t *nswept, const int64_t *xoff, double2 *x, cudaStream_t s)
{
    const unsigned blocks = (unsigned)((batch + 31) / 32);
    qr_backsolve_c64_kernel<L, U><<<blocks, 32, 0, s>>>(batch, gbase, work, ncols, nswept, xoff, x);
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

#define ILA_QRC_DISPATCH(FN, l, u, ...)                                                        \
    do {                                                                                       \
        if (l == 1 && u == 1) return FN<1, 1>(__VA_ARGS__);                                    \
        if (l == 1 && u == 2) return FN<1, 2>(__VA_ARGS__);                                    \
        if (l == 2 && u == 1) return FN<2, 1>(__VA_ARGS__);                                    \
        if (l == 2 && u == 2) return FN<2, 2>(__VA_ARGS__);                                    \
        if (l == 3 && u == 1) return FN<3, 1>(__VA_ARGS__);                                    \
        if (l == 1 && u == 3) return FN<1, 3>(__VA_ARGS__);                                    \
        if (l == 3 && u == 3) return FN<3, 3>(__VA_ARGS__);                                    \
        return set_error(ILA_ERR_UNSUPPORTED, "qr_banded_c64: bandwidths (l,u) not instantiated"); \
    } while (0)

int qr_c64_record_chunks(int l, int u) { return l + u + 2; }

int launch_qr_forward_c64(int l, int u, const QrC64FwdArgs &a, cudaStream_t s)
{
    if (a.batch <= 0) return 0;
    ILA_QRC_DISPATCH(launch_c64_lu, l, u, a, s);
}

int launch_qr_backsolve_c64(int l, int u, int64_t batch, const int64_t *gbase, const double2 *work, const int64_t *ncols,
                            const int64_t *nswept, const int64_t *xoff, double2 *x, cudaStream_t s)
{
    if (batch <= 0) return 0;
    ILA_QRC_DISPATCH(launch_c64_bwd_lu, l, u, batch, gbase, work, ncols, nswept, xoff, x, s);
}

} // namespace ila
