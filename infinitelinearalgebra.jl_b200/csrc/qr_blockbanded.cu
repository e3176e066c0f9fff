// qr_blockbanded.cu — K11: adaptive block-banded QR solve for ∞ operators with block sizes 1, 2, 3, … (the KronTrav /
// triangular-traversal operators of 2-D problems), one CTA per operator.
//
// Reference path replaced (InfiniteLinearAlgebra.jl v0.10.3):
//   AdaptiveQRData(::AbstractBlockLayout, A)                 src/infqr.jl:23-28    block bandwidths (l, l+u)
//   partialqr!(F, N::Block{1}) -> _blockbanded_qr!           src/infqr.jl:68-86    dense Householder QR of the rows of blocks
//                                                            J..J+l per block column, applied to block columns J+1..J+l+u
//   Q'b adaptive driver (COLGROWTH = 300, tolerance 1e-30)   src/infqr.jl:303-337
//   resizedata_chop!(::BlockedVector, tol)                   src/infqr.jl:221-232
//   ldiv!(R, B::BlockedArray)                                src/infqr.jl:358-366
// Operator family: L_i = alpha_i KronTrav(T1, I) + beta_i KronTrav(I, T2) - lam_i I with T1, T2 tridiagonal (shared bands),
// KronTrav(A, B)[Block(K,J)][k,j] = A[k,j] B[K-k+1,J-j+1]: block bandwidths (1,1), diagonal blocks diagonal, off-diagonal
// blocks with two bands — every block-banded fixture of test/test_infqr.jl:270-313; the batch axis is (alpha, beta, lam),
// e.g. the shifts of a 2-D resolvent.
//
// B200-first design (a different kernel class from K1: the work per operator is O(N^4) dense panel arithmetic, not a
// scalar recurrence).  One CTA owns one operator.  The active panel — rows of blocks J, J+1, columns of blocks J..J+2 plus
// the right-hand side as one more column, (2J+1) x (3J+4) doubles — lives in shared memory (up to 197 KB: the reason for
// one CTA per SM) with CIRCULAR row and column indexing, so that moving from block column J to J+1 neither copies nor
// shifts anything: the J+1 fresh rows of block J+2 are generated from (T1, T2, alpha, beta, lam) straight into the slots
// the finished rows leave.
//   Default path (BLOCKED): panels of four pivot columns.  Warp 0 holds a panel in registers (row i in lane i mod 32) and
// factorizes it with one packed shuffle all-reduce per pivot column (norm, dot products with the columns to the right and
// Gram entries with the reflectors to the left together), division-free reflector scalars; V (32-byte rows), the four taus
// and the six Gram entries go to shared memory, double buffered.  Twelve warps apply a whole panel to the trailing columns
// (16 column pairs per warp, the two lanes of a pair on alternate rows): one pass for the four dot products, the Gram
// recurrence for the coefficients of the four sequential reflections, one pass for col -= sum w_k v_k — the trailing panel
// is read twice and written once per FOUR reflectors.  Meanwhile warp 0 and a helper warp apply the panel to the next
// panel's columns (two each, meeting at a two-warp named barrier) and warp 0 factorizes them: one block barrier per panel.
//   Validation path (ila_qr_blockbanded_set_mode(1)): one reflector per pass, 16 column pairs per warp, the pair that
// updates the next pivot column forms its scalars on the way (one barrier per column).
// Q'b is the extra column, so the adaptive test (max|Q'b| over the blocks J..J+l at the COLGROWTH cadence) is a
// shared-memory reduction — no host round trip.  Finished R rows stream to the operator's slab of global memory
// (coalesced, L2-resident: <= 2.1 MB) and the back-substitution reads them back block row by block row: warp-per-row dot
// products for the part right of the diagonal block, and ONE warp with the diagonal block in shared memory and the
// right-hand side in registers for the triangular part (reciprocals of the diagonal formed in parallel, shuffle broadcast
// of each new solution entry, no block-wide barrier inside a block row).
#include "ila_common.cuh"
#include "ila_kernels.cuh"
#include "ila_exact.cuh"

namespace ila {

namespace {

__device__ __forceinline__ int bstart(int K) { return K * (K - 1) / 2; }            // 0-based first index of block K (1-based)
__device__ __forceinline__ int findblock(int i)                                      // block holding the 1-based index i
{
    int K = (int)((sqrt(8.0 * (double)i + 1.0) - 1.0) * 0.5);
    while (K * (K + 1) / 2 < i) K++;
    while (K > 1 && (K - 1) * K / 2 >= i) K--;
    return K;
}

constexpr int kThreads = 448;          // 14 warps x 16 column pairs = 224 >= 3*64 + 1 panel columns
constexpr int kWarps = kThreads / 32;

struct Panel {
    ILA_BP(double) P; // HMAX x LD, row-major; column slot wmax is the right-hand side (bounds-checked in the debug build)
    int hmask, wmax, ld;
    __device__ __forceinline__ int rs(int g) const { return g & hmask; }
    __device__ __forceinline__ int cs(int c) const { return c % wmax; }
    __device__ __forceinline__ double &at(int g, int c) { return P[rs(g) * ld + cs(c)]; }
    __device__ __forceinline__ double &rhs(int g) { return P[rs(g) * ld + wmax]; }
};

// LAPACK dlarfg from the pivot alpha and the sum of squares below it: (tau, 1/(alpha - beta), beta); tau = 0 <=> H = I
__device__ __forceinline__ void reflector_scalars(double alpha, double ss, double &tau, double &inv, double &beta)
{
    tau = 0.0; inv = 0.0; beta = alpha;
    if (ss != 0.0) {
        beta = -copysign(esqrt(fma(alpha, alpha, ss)), alpha);
        tau = ediv(beta - alpha, beta);
        inv = ediv(1.0, alpha - beta);
    }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// all-reduce of N values (N = 4 or 16) over the 32 lanes with HALVING payloads: at the stage with lane offset o a lane keeps
// the half of its values selected by its bit o and sends the other half, so N/2 + N/4 + ... + 1 shuffles bring every value
// to one lane group (plus log2(32/N) single-value stages); N more shuffles broadcast the sums back.  16 values: 32 shuffles
// instead of 80.
template <int N>
__device__ __forceinline__ void packed_allreduce(double (&v)[N], int lane)
{
    static_assert(N == 4 || N == 8 || N == 16, "packed_allreduce: N must be 4, 8 or 16");
    int o = 16;
    if constexpr (N >= 16) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const double keep = up ? v[i + 8] : v[i], send = up ? v[i] : v[i + 8];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
        o >>= 1;
    }
    if constexpr (N >= 8) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double keep = up ? v[i + 4] : v[i], send = up ? v[i] : v[i + 4];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
        o >>= 1;
    }
    {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double keep = up ? v[i + 2] : v[i], send = up ? v[i] : v[i + 2];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
        o >>= 1;
    }
    {
        const bool up = lane & o;
        const double keep = up ? v[1] : v[0], send = up ? v[0] : v[1];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        o >>= 1;
    }
    for (; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
    const double sum = v[0];
#pragma unroll
    for (int j = 0; j < N; j++) v[j] = __shfl_sync(0xffffffffu, sum, j * (32 / N));
}

// ---- blocked variant: four reflectors per pass over the trailing panel ---------------------------------------------
// Panel layout in registers of ONE warp: row i (relative to the panel's first pivot row) lives in lane i & 31, chunk i >> 5
// (m <= 127 rows = 4 chunks), the (at most) four pivot columns side by side.  V[i][0..3] (explicit 1 on the diagonal, 0
// above it) goes to shared memory as 32-byte rows, with the taus and the Gram entries g_kl = v_k . v_l that let a trailing
// column apply all four reflectors from ONE set of dot products:
//     w0 = tau0 d0,  w1 = tau1 (d1 - w0 g10),  w2 = tau2 (d2 - w0 g20 - w1 g21),  w3 = tau3 (d3 - w0 g30 - w1 g31 - w2 g32),
//     col -= w0 v0 + w1 v1 + w2 v2 + w3 v3                      (d_k = v_k . col on the column BEFORE the panel).
// reflector scalars without divisions on the dependent chain: y ~ 1/sqrt(alpha^2 + ss) by MUFU + two Newton steps,
// norm = s2*y with one correction, tau = 1 + |alpha|/norm, 1/(alpha - beta) = sign(alpha)/(|alpha| + norm)
__device__ __forceinline__ void reflector_scalars_fast(double alpha, double ss, double &tau, double &inv, double &beta)
{
    tau = 0.0; inv = 0.0; beta = alpha;
    if (ss != 0.0) {
        const double s2 = fma(alpha, alpha, ss);
        if (inbox_sq(s2)) {
            double y;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s2));
            const double h = 0.5 * s2;
            y = fma(y, fma(-h * y, y, 0.5), y);
            y = fma(y, fma(-h * y, y, 0.5), y);
            double nrm = s2 * y;
            nrm = fma(fma(-nrm, nrm, s2), 0.5 * y, nrm);
            const double aa = fabs(alpha), den = aa + nrm;
            tau = fma(aa, y, 1.0);
            tau = fma(fma(-(tau - 1.0), nrm, aa), y, tau);         // one correction of |alpha|/norm
            beta = -copysign(nrm, alpha);
            inv = copysign(rcp_refined(den), alpha);
        } else {
            reflector_scalars(alpha, ss, tau, inv, beta);
        }
    }
}

__device__ __forceinline__ void panel_factor(Panel &pn, double *Vbuf, double *psb, int cp, int kb, int m, int lane)
{
    const int ld = pn.ld;
    double a[4][4];
    int rowoff[4];
    bool valid[4];
    int cslot[4];
    {
        int sl = pn.cs(cp);
#pragma unroll
        for (int k = 0; k < 4; k++) { cslot[k] = sl; if (++sl == pn.wmax) sl = 0; }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int i = lane + 32 * q;
        valid[q] = i < m;
        rowoff[q] = pn.rs(cp + i) * ld;
#pragma unroll
        for (int k = 0; k < 4; k++) a[q][k] = (valid[q] && k < kb) ? pn.P[rowoff[q] + cslot[k]] : 0.0;
    }
    double tau[4] = {0.0, 0.0, 0.0, 0.0};
    double g[4][4];                                   // g[k][l], l < k: v_k . v_l
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int l = 0; l < 4; l++) g[k][l] = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k < kb) {
            // ONE reduction round per pivot column: its sum of squares below the diagonal and its (unscaled) dot products with
            // the columns to its right (the reflector is applied to them) and with the finished reflectors to its left (Gram)
            double red[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = lane + 32 * q;
                if (i > k && valid[q]) {
#pragma unroll
                    for (int l = 0; l < 4; l++) red[l] = fma(a[q][k], a[q][l], red[l]);
                }
            }
            double row_k[4];                          // row k of the panel (lane k, chunk 0)
#pragma unroll
            for (int l = 0; l < 4; l++) row_k[l] = __shfl_sync(0xffffffffu, a[0][l], k);
            packed_allreduce<4>(red, lane);
            double inv, beta;
            reflector_scalars_fast(row_k[k], red[k], tau[k], inv, beta);
            // rows above the diagonal of this column are final R entries; the diagonal becomes beta
            if (lane < k) pn.P[rowoff[0] + cslot[k]] = a[0][k];
            if (lane == k) pn.P[rowoff[0] + cslot[k]] = beta;
            const bool on = tau[k] != 0.0;
#pragma unroll
            for (int l = 0; l < k; l++) g[k][l] = on ? fma(inv, red[l], row_k[l]) : 0.0;      // v_k . v_l = v_l[k] + inv * sum
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = lane + 32 * q;
                a[q][k] = (i > k && valid[q]) ? a[q][k] * inv : ((i == k && on) ? 1.0 : 0.0);
            }
#pragma unroll
            for (int l = k + 1; l < 4; l++) {
                const double w = tau[k] * fma(inv, red[l], row_k[l]);
#pragma unroll
                for (int q = 0; q < 4; q++) a[q][l] = fma(-w, a[q][k], a[q][l]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) a[q][k] = 0.0;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int i = lane + 32 * q;
        if (valid[q]) {
            double2 *dst = reinterpret_cast<double2 *>(Vbuf + i * 4);
            dst[0] = make_double2(a[q][0], a[q][1]);
            dst[1] = make_double2(a[q][2], a[q][3]);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) psb[k] = tau[k];
        psb[4] = g[1][0]; psb[5] = g[2][0]; psb[6] = g[2][1]; psb[7] = g[3][0]; psb[8] = g[3][1]; psb[9] = g[3][2];
    }
}

// the current panel's four reflectors applied to two of the (at most four) columns of the next panel by one warp (warp 0
// takes columns 0-1, the helper warp columns 2-3): all eight dot products in one packed reduction round, then the same Gram
// recurrence as the trailing columns
__device__ __forceinline__ void panel_apply2(Panel &pn, const double *Vbuf, const double *psb, int cp, int m, int cn, int ncol, int lane)
{
    const int ld = pn.ld;
    double a[4][2], v[4][4];
    int rowoff[4];
    bool valid[4];
    int cslot[2];
    {
        int sl = pn.cs(cn);
        cslot[0] = sl;
        if (++sl == pn.wmax) sl = 0;
        cslot[1] = sl;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int i = lane + 32 * q;
        valid[q] = i < m;
        rowoff[q] = pn.rs(cp + i) * ld;
#pragma unroll
        for (int k = 0; k < 2; k++) a[q][k] = (valid[q] && k < ncol) ? pn.P[rowoff[q] + cslot[k]] : 0.0;
        if (valid[q]) {
            const double2 *src = reinterpret_cast<const double2 *>(Vbuf + i * 4);
            const double2 x0 = src[0], x1 = src[1];
            v[q][0] = x0.x; v[q][1] = x0.y; v[q][2] = x1.x; v[q][3] = x1.y;
        } else {
            v[q][0] = v[q][1] = v[q][2] = v[q][3] = 0.0;
        }
    }
    double df[8];                                     // df[2k + l] = v_k . a_l
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int l = 0; l < 2; l++) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < 4; q++) t = fma(v[q][k], a[q][l], t);
            df[2 * k + l] = t;
        }
    packed_allreduce<8>(df, lane);
    const double t0 = psb[0], t1 = psb[1], t2 = psb[2], t3 = psb[3];
    const double g10 = psb[4], g20 = psb[5], g21 = psb[6], g30 = psb[7], g31 = psb[8], g32 = psb[9];
#pragma unroll
    for (int l = 0; l < 2; l++) {
        const double w0 = t0 * df[l];
        const double w1 = t1 * fma(-w0, g10, df[2 + l]);
        const double w2 = t2 * fma(-w1, g21, fma(-w0, g20, df[4 + l]));
        const double w3 = t3 * fma(-w2, g32, fma(-w1, g31, fma(-w0, g30, df[6 + l])));
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double x = a[q][l];
            x = fma(-w0, v[q][0], x); x = fma(-w1, v[q][1], x); x = fma(-w2, v[q][2], x); x = fma(-w3, v[q][3], x);
            a[q][l] = x;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
        for (int k = 0; k < 2; k++)
            if (valid[q] && k < ncol) pn.P[rowoff[q] + cslot[k]] = a[q][k];
    }
}

} // namespace

template <bool BLOCKED>
__global__ void __launch_bounds__(kThreads, 1) qr_blockbanded_kernel(BlockQrArgs A)
{
    extern __shared__ double smem[];
    __shared__ double sc[4];
    __shared__ int si[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t inst = blockIdx.x;
    const int hmax = A.hmax, wmax = A.wmax, ld = wmax + 1;
    Panel pn{ILA_MKBP(double, smem, (size_t)hmax * ld, 1101), hmax - 1, wmax, ld};
    const double al = A.alpha[inst], be = A.beta[inst], lam = A.lam[inst];
    // the 1-D operators' bands staged in shared memory (behind the panel, the V buffers and the panel scalars): the row
    // generator reads them once per fresh row
    constexpr int tld = 72;
    ILA_BP(double) T1 = ILA_MKBP(double, smem + (size_t)hmax * ld + 2 * (size_t)hmax * 4 + 32, 3 * tld, 1108);
    ILA_BP(double) T2 = ILA_MKBP(double, smem + (size_t)hmax * ld + 2 * (size_t)hmax * 4 + 32 + 3 * tld, 3 * tld, 1109);
    for (int i = tid; i < 3 * tld; i += kThreads) {
        const int band = i / tld, k = i % tld;
        T1[i] = k <= A.maxblocks ? A.T1[band * A.tld + k] : 0.0;
        T2[i] = k <= A.maxblocks ? A.T2[band * A.tld + k] : 0.0;
    }
    __syncthreads();
    const double *bi = A.b + inst * A.nb;
    const int nb = A.nb, cg = A.colgrowth, maxblocks = A.maxblocks;
    ILA_BP(double) Rw = ILA_MKBP(double, A.work + inst * A.work_stride, A.work_stride, 1102);
    ILA_BP(double) qtb = ILA_MKBP(double, A.qtb + inst * A.xld, A.xld, 1103);

    // rows of block K (1-based) into the panel (one warp per row): zero the columns of blocks Klo..K+1, then the (at most
    // five) entries of the row and its right-hand-side entry
    auto load_block_rows = [&](int K, int Klo) {
        const int c_lo = bstart(Klo), w = bstart(K + 2) - c_lo, r0 = bstart(K), s_lo = pn.cs(c_lo);
        for (int k = warp; k < K; k += kWarps) {
            ILA_BP(double) row = pn.P + pn.rs(r0 + k) * ld;
            for (int ci = lane; ci < w; ci += 32) {
                int sl = s_lo + ci;
                if (sl >= wmax) sl -= wmax;
                row[sl] = 0.0;
            }
            __syncwarp();
            if (lane == 0) {
                const int kk = k + 1, g = r0 + k;
                row[pn.cs(g)] = al * T1[tld + kk - 1] + be * T2[tld + K - kk] - lam;
                const int cn = bstart(K + 1);
                row[pn.cs(cn + kk)] += al * T1[2 * tld + kk - 1];            // (kk, kk+1): T1[kk, kk+1]
                row[pn.cs(cn + kk - 1)] += be * T2[2 * tld + K - kk];        // (kk, kk):   T2[K-kk+1, K-kk+2]
                if (K > 1) {
                    const int cp = bstart(K - 1);
                    if (kk >= 2) row[pn.cs(cp + kk - 2)] += al * T1[kk - 2];            // (kk, kk-1): T1[kk, kk-1]
                    if (kk <= K - 1) row[pn.cs(cp + kk - 1)] += be * T2[K - kk - 1];    // (kk, kk):   T2[K-kk+1, K-kk]
                }
                row[wmax] = g < nb ? bi[g] : 0.0;
            }
        }
    };

#ifdef ILA_K11_PROFILE
    long long tp_a = 0, tp_n = 0, tp_j = 0, tp_c = 0, tp_bs = 0, tp_t0, tp_w0 = 0, tp_w1 = 0, tp_a2 = 0, tp_bar = 0, tp_pf = 0, tp_sync = 0;
    int tp_np = 0;
#define K11_T0() tp_t0 = clock64()
#define K11_ACC(v) do { const long long t_ = clock64(); v += t_ - tp_t0; tp_t0 = t_; } while (0)
#else
#define K11_T0()
#define K11_ACC(v)
#endif
    int J0 = 1, J1 = findblock(cg), done = 0;
    const int SB = findblock(nb);
    int32_t st = ILA_OK;
    load_block_rows(1, 1);
    __syncthreads();

    // column pair of this thread: 16 pairs per warp, the two lanes of a pair take alternate rows (a 64-bit shared-memory
    // request is served per half-warp, so the two halves cost nothing extra)
    const int tp = warp * 16 + (lane & 15), half = lane >> 4;

    for (;;) {
        // ---- adaptive test at the chunk start (infqr.jl:316-322): the rows of block J0+1 still hold raw b (zero past SB)
        if (warp == 0) {
            double mx = 0.0;
            bool nan = false;
            const int r0 = bstart(J0);
            for (int k = lane; k < J0; k += 32) {
                const double v = pn.rhs(r0 + k);
                nan |= (v != v);
                mx = fmax(mx, fabs(v));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            nan = __any_sync(0xffffffffu, nan);
            if (lane == 0) si[0] = nan ? 2 : ((J0 > SB && mx <= A.tol) ? 1 : 0);
        }
        __syncthreads();
        const int flag = si[0];
        __syncthreads();
        if (flag == 2) { st = ILA_NAN; break; }
        if (flag == 1) break;
        if (J1 + 1 > maxblocks) {
            J1 = maxblocks - 1;
            if (J1 < J0) { st = ILA_NOT_CONVERGED; break; }
        }

        for (int J = J0; J <= J1; J++) {
            const int c0 = bstart(J), nrows = 2 * J + 1, ncols = 3 * J + 3;
            // (a) fresh rows of block J+1; block column J+2 of the rows of block J starts at zero
            K11_T0();
            {
                const int s2 = pn.cs(bstart(J + 2)), w2 = J + 2;
                for (int k = warp; k < J; k += kWarps) {
                    ILA_BP(double) row = pn.P + pn.rs(c0 + k) * ld;
                    for (int ci = lane; ci < w2; ci += 32) {
                        int sl = s2 + ci;
                        if (sl >= wmax) sl -= wmax;
                        row[sl] = 0.0;
                    }
                }
            }
            load_block_rows(J + 1, J);
            __syncthreads();
            K11_ACC(tp_a);
            if constexpr (BLOCKED) {
            // (b) blocked: panels of four pivot columns.  Warp 0 is the panel warp: while the other 13 warps apply panel p to
            // the columns behind the NEXT panel, it applies panel p to the next panel's columns and factorizes them (registers
            // + shuffles), so the trailing update — the shared-memory-bound part — never waits for a panel factorization
            // except at the first panel of a block column.
            double *Vp = smem + (size_t)hmax * ld;            // [2][hmax][4]
            double *ps = Vp + 2 * (size_t)hmax * 4;           // [2][16]
            const int cend = c0 + ncols, rend = c0 + nrows;
            if (warp == 0) panel_factor(pn, Vp, ps, c0, min(4, J), nrows, lane);
            __syncthreads();
            K11_ACC(tp_n);
            int buf = 0;
            for (int j0 = 0; j0 < J; j0 += 4) {
                const int cp = c0 + j0, kb = min(4, J - j0), m = rend - cp;
                const int kbn = min(4, J - j0 - kb);
                const double *Vb = Vp + (size_t)buf * hmax * 4;
                const double *psb = ps + buf * 16;
#ifdef ILA_K11_PROFILE
                const long long tw0 = clock64();
#endif
                if (warp == 0 || warp == kWarps - 1) {
                    // panel warps: warp 0 and the helper each apply the current panel to two columns of the next one, meet at a
                    // two-warp named barrier, then warp 0 factorizes the four columns
                    if (kbn > 0) {
                        const int coff = warp == 0 ? 0 : 2;
#ifdef ILA_K11_PROFILE
                        const long long ta = clock64();
#endif
                        if (kbn > coff) panel_apply2(pn, Vb, psb, cp, m, cp + kb + coff, min(2, kbn - coff), lane);
#ifdef ILA_K11_PROFILE
                        const long long tb = clock64();
#endif
                        asm volatile("bar.sync 1, 64;" ::: "memory");
#ifdef ILA_K11_PROFILE
                        const long long tc = clock64();
#endif
                        if (warp == 0)
                            panel_factor(pn, Vp + (size_t)(buf ^ 1) * hmax * 4, ps + (buf ^ 1) * 16, cp + kb, kbn, m - kb, lane);
#ifdef ILA_K11_PROFILE
                        if (warp == 0 && lane == 0) { tp_a2 += tb - ta; tp_bar += tc - tb; tp_pf += clock64() - tc; tp_np++; }
#endif
                    }
                } else {
                    const int tq = (warp - 1) * 16 + (lane & 15);      // 12 trailing warps x 16 pairs = 192 >= 3*63 + 4 - 8
                    const int first = cp + kb + kbn, ntrail = cend - first;
                    const bool act = tq <= ntrail;
                    int scol = pn.cs(first) + tq;
                    if (scol >= wmax) scol -= wmax;
                    ILA_BP(double) col = pn.P + (tq < ntrail ? scol : wmax);
                    const int cnt = act ? (m - half + 1) >> 1 : 0;       // rows i = half, half+2, ... < m
                    const int hm = pn.hmask;
                    const double2 *__restrict__ V2 = reinterpret_cast<const double2 *>(Vb);
                    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
                    {
                        int r = pn.rs(cp + half), vi = half, k = 0;
                        for (; k + 2 <= cnt; k += 2) {
                            const int oa = r * ld, ob = ((r + 2) & hm) * ld;
                            r = (r + 4) & hm;
                            const double qa = col[oa], qb = col[ob];
                            const double2 a0 = V2[vi * 2], a1 = V2[vi * 2 + 1], b0 = V2[vi * 2 + 4], b1 = V2[vi * 2 + 5];
                            vi += 4;
                            d0 = fma(a0.x, qa, d0); d1 = fma(a0.y, qa, d1); d2 = fma(a1.x, qa, d2); d3 = fma(a1.y, qa, d3);
                            d0 = fma(b0.x, qb, d0); d1 = fma(b0.y, qb, d1); d2 = fma(b1.x, qb, d2); d3 = fma(b1.y, qb, d3);
                        }
                        if (k < cnt) {
                            const double qa = col[r * ld];
                            const double2 a0 = V2[vi * 2], a1 = V2[vi * 2 + 1];
                            d0 = fma(a0.x, qa, d0); d1 = fma(a0.y, qa, d1); d2 = fma(a1.x, qa, d2); d3 = fma(a1.y, qa, d3);
                        }
                    }
                    d0 += __shfl_xor_sync(0xffffffffu, d0, 16); d1 += __shfl_xor_sync(0xffffffffu, d1, 16);
                    d2 += __shfl_xor_sync(0xffffffffu, d2, 16); d3 += __shfl_xor_sync(0xffffffffu, d3, 16);
                    const double w0 = psb[0] * d0;
                    const double w1 = psb[1] * fma(-w0, psb[4], d1);
                    const double w2 = psb[2] * fma(-w1, psb[6], fma(-w0, psb[5], d2));
                    const double w3 = psb[3] * fma(-w2, psb[9], fma(-w1, psb[8], fma(-w0, psb[7], d3)));
                    {
                        int r = pn.rs(cp + half), vi = half, k = 0;
                        for (; k + 2 <= cnt; k += 2) {
                            const int oa = r * ld, ob = ((r + 2) & hm) * ld;
                            r = (r + 4) & hm;
                            double qa = col[oa], qb = col[ob];
                            const double2 a0 = V2[vi * 2], a1 = V2[vi * 2 + 1], b0 = V2[vi * 2 + 4], b1 = V2[vi * 2 + 5];
                            vi += 4;
                            qa = fma(-w0, a0.x, qa); qa = fma(-w1, a0.y, qa); qa = fma(-w2, a1.x, qa); qa = fma(-w3, a1.y, qa);
                            qb = fma(-w0, b0.x, qb); qb = fma(-w1, b0.y, qb); qb = fma(-w2, b1.x, qb); qb = fma(-w3, b1.y, qb);
                            col[oa] = qa; col[ob] = qb;
                        }
                        if (k < cnt) {
                            const int oa = r * ld;
                            double qa = col[oa];
                            const double2 a0 = V2[vi * 2], a1 = V2[vi * 2 + 1];
                            qa = fma(-w0, a0.x, qa); qa = fma(-w1, a0.y, qa); qa = fma(-w2, a1.x, qa); qa = fma(-w3, a1.y, qa);
                            col[oa] = qa;
                        }
                    }
                }
#ifdef ILA_K11_PROFILE
                if (lane == 0 && warp == 0) tp_w0 += clock64() - tw0;
                if (lane == 0 && warp == 1) tp_w1 += clock64() - tw0;
                const long long ts0 = clock64();
#endif
                __syncthreads();
#ifdef ILA_K11_PROFILE
                if (lane == 0 && warp == 0) tp_sync += clock64() - ts0;
#endif
                buf ^= 1;
            }
            } else {
            // (b) Householder QR of block column J.  The scalars of column j+1 are formed by the pair that updates it (the sum
            // of squares rides on its update pass), so only the first column of a block column needs a reduction of its own
            // and there is ONE barrier per column.
            if (warp == 0) {
                const int csc = pn.cs(c0);
                double ss = 0.0;
                for (int i = 1 + lane; i < nrows; i += 32) {
                    const double v = pn.P[pn.rs(c0 + i) * ld + csc];
                    ss = fma(v, v, ss);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                if (lane == 0) {
                    double tau, inv, beta;
                    reflector_scalars(pn.P[pn.rs(c0) * ld + csc], ss, tau, inv, beta);
                    pn.P[pn.rs(c0) * ld + csc] = beta;
                    sc[0] = tau; sc[1] = inv;
                }
            }
            __syncthreads();
            K11_ACC(tp_n);
            int csc = pn.cs(c0);                              // slot of the pivot column, advanced with j
            for (int j = 0; j < J; j++) {
                const int c = c0 + j, m = nrows - j;          // pivot column, rows c .. c+m-1
                const double tau = sc[(j & 1) * 2], inv = sc[(j & 1) * 2 + 1];
                const int ntrail = ncols - j - 1;             // trailing panel columns; pair ntrail owns the right-hand side
                const bool act = tp <= ntrail;
                int scol = csc + 1 + tp;
                if (scol >= wmax) scol -= wmax;
                ILA_BP(const double) piv = pn.P + csc;
                ILA_BP(double) col = pn.P + (tp < ntrail ? scol : wmax);
                const int cnt = (act && m > half) ? (m - half) >> 1 : 0;      // this lane's rows: i = 1+half, 3+half, ...
                const int hm = pn.hmask;
                const int rd = pn.rs(c) * ld;
                const int rfirst = pn.rs(c + 1 + half);
                double dotA = 0.0, dotB = 0.0;
                {
                    int r = rfirst, k = 0;
                    for (; k + 4 <= cnt; k += 4) {
                        const int o0 = r * ld, o1 = ((r + 2) & hm) * ld, o2 = ((r + 4) & hm) * ld, o3 = ((r + 6) & hm) * ld;
                        r = (r + 8) & hm;
                        const double p0 = piv[o0], p1 = piv[o1], p2 = piv[o2], p3 = piv[o3];
                        const double q0 = col[o0], q1 = col[o1], q2 = col[o2], q3 = col[o3];
                        dotA = fma(p0, q0, dotA); dotB = fma(p1, q1, dotB);
                        dotA = fma(p2, q2, dotA); dotB = fma(p3, q3, dotB);
                    }
                    for (; k < cnt; k++) {
                        dotA = fma(piv[r * ld], col[r * ld], dotA);
                        r = (r + 2) & hm;
                    }
                }
                double dot = dotA + dotB;
                dot += __shfl_xor_sync(0xffffffffu, dot, 16);
                const double cd = act ? col[rd] : 0.0;
                const double w = tau * fma(inv, dot, cd);
                const double winv = -(w * inv);
                if (act && half == 0) col[rd] = cd - w;
                double ss = 0.0;
                {
                    int r = rfirst, k = 0;
                    const bool first_is_diag = half == 0;         // row i = 1 is the next pivot's diagonal: not part of its norm
                    for (; k + 4 <= cnt; k += 4) {
                        const int o0 = r * ld, o1 = ((r + 2) & hm) * ld, o2 = ((r + 4) & hm) * ld, o3 = ((r + 6) & hm) * ld;
                        r = (r + 8) & hm;
                        const double p0 = piv[o0], p1 = piv[o1], p2 = piv[o2], p3 = piv[o3];
                        const double q0 = col[o0], q1 = col[o1], q2 = col[o2], q3 = col[o3];
                        const double v0 = fma(p0, winv, q0), v1 = fma(p1, winv, q1), v2 = fma(p2, winv, q2), v3 = fma(p3, winv, q3);
                        col[o0] = v0; col[o1] = v1; col[o2] = v2; col[o3] = v3;
                        if (warp == 0) {
                            if (!(first_is_diag && k == 0)) ss = fma(v0, v0, ss);
                            ss = fma(v1, v1, ss); ss = fma(v2, v2, ss); ss = fma(v3, v3, ss);
                        }
                    }
                    for (; k < cnt; k++) {
                        const double v = fma(piv[r * ld], winv, col[r * ld]);
                        col[r * ld] = v;
                        if (warp == 0 && !(first_is_diag && k == 0)) ss = fma(v, v, ss);
                        r = (r + 2) & hm;
                    }
                }
                if (warp == 0 && j + 1 < J) {
                    ss += __shfl_xor_sync(0xffffffffu, ss, 16);
                    __syncwarp();
                    if (lane == 0) {                          // pair 0 owns the next pivot column c+1
                        double t2, i2, b2;
                        const int rn = pn.rs(c + 1) * ld;
                        reflector_scalars(col[rn], ss, t2, i2, b2);
                        col[rn] = b2;
                        sc[((j + 1) & 1) * 2] = t2; sc[((j + 1) & 1) * 2 + 1] = i2;
                    }
                }
                if (++csc == wmax) csc = 0;
                __syncthreads();
            }
            }
            // (c) rows of block J are final: R[J, J..J+2] and (Q'b)[J] leave the panel (one warp per row, coalesced)
            K11_ACC(tp_j);
            {
                ILA_BP(double) dst = Rw + (int64_t)(J - 1) * J * (J + 1);
                const int s0 = pn.cs(c0);
                for (int k = warp; k < J; k += kWarps) {
                    ILA_BP(const double) row = pn.P + pn.rs(c0 + k) * ld;
                    for (int ci = lane; ci < ncols; ci += 32) {
                        int sl = s0 + ci;
                        if (sl >= wmax) sl -= wmax;
                        dst[k * ncols + ci] = row[sl];
                    }
                    if (lane == 0) qtb[c0 + k] = row[wmax];
                }
            }
            __syncthreads();
            K11_ACC(tp_c);
        }
        done = J1;
        J0 = J1 + 1;
        J1 = findblock(bstart(J1 + 1) + cg);
    }

    int n2 = 0;
    K11_T0();
    if (st == ILA_OK && done > 0) {
        // ---- chop (infqr.jl:221-232): Q'b lives in qtb (blocks 1..done) and in the panel (block done+1)
        const int r0 = bstart(done + 1);
        if (tid < done + 1) qtb[r0 + tid] = pn.rhs(r0 + tid);
        __syncthreads();
        const int len = bstart(done + 2);
        int last = -1;
        for (int i = tid; i < len; i += kThreads)
            if (fabs(qtb[i]) > A.tol) last = i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        if (tid == 0) si[1] = -1;
        __syncthreads();
        if (lane == 0) atomicMax(&si[1], last);
        __syncthreads();
        last = si[1];
        if (last >= 0) {
            const int K = findblock(last + 1);
            n2 = bstart(K + 1);
            // ---- back-substitution on the leading K x K blocks (infqr.jl:358-366)
            ILA_BP(double) Ts = ILA_MKBP(double, smem, 64 * 65, 1104);                   // diagonal block, stride ldt (odd)
            ILA_BP(double) cs_ = ILA_MKBP(double, smem + 64 * 65, 64, 1105);              // right-hand side of the current block row
            ILA_BP(double) xs = ILA_MKBP(double, smem + 64 * 65 + 64, maxblocks * (maxblocks + 1) / 2 + 8, 1106);   // solution
            for (int Kc = K; Kc >= 1; Kc--) {
                const int s = bstart(Kc), wrow = 3 * Kc + 3, ldt = Kc | 1;
                const int wlim = min(wrow, n2 - s);
                ILA_BP(const double) src = Rw + (int64_t)(Kc - 1) * Kc * (Kc + 1);
                for (int k = warp; k < Kc; k += kWarps) {
                    ILA_BP(const double) row = src + (int64_t)k * wrow;
                    double acc = 0.0;
                    for (int ci = lane; ci < wlim; ci += 32) {
                        const double v = row[ci];
                        if (ci < Kc) Ts[k * ldt + ci] = v;
                        else acc = fma(v, xs[s + ci], acc);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane == 0) cs_[k] = qtb[s + k] - acc;
                }
                __syncthreads();
                if (warp == 0) {
                    // rows lane and lane+32 of the block in registers; x_k broadcast from its owner
                    double c_lo = lane < Kc ? cs_[lane] : 0.0, c_hi = lane + 32 < Kc ? cs_[lane + 32] : 0.0;
                    // reciprocals of the diagonal off the dependent chain (x_k = c_k * (1/R_kk): one rounding more than the
                    // reference's division, inside the 1e-12 bar)
                    const double rd_lo = lane < Kc ? 1.0 / Ts[lane * ldt + lane] : 0.0;
                    const double rd_hi = lane + 32 < Kc ? 1.0 / Ts[(lane + 32) * ldt + lane + 32] : 0.0;
                    for (int k = Kc - 1; k >= 0; k--) {
                        const double xk = __shfl_sync(0xffffffffu, k >= 32 ? c_hi * rd_hi : c_lo * rd_lo, k & 31);
                        if (lane < k) c_lo = fma(-Ts[lane * ldt + k], xk, c_lo);
                        if (lane + 32 < k) c_hi = fma(-Ts[(lane + 32) * ldt + k], xk, c_hi);
                        if (lane == 0) xs[s + k] = xk;
                    }
                }
                __syncthreads();
            }
            ILA_BP(double) xo = ILA_MKBP(double, A.x + inst * A.xld, A.xld, 1107);
            for (int i = tid; i < A.xld; i += kThreads) xo[i] = i < n2 ? xs[i] : 0.0;
        }
    }
    if (n2 == 0) {
        double *xo = A.x + inst * A.xld;
        for (int i = tid; i < A.xld; i += kThreads) xo[i] = 0.0;
    }
#ifdef ILA_K11_PROFILE
    K11_ACC(tp_bs);
    if (tid == 0) {
        double *xo = A.x + inst * A.xld + A.xld - 5;
        xo[0] = (double)tp_a; xo[1] = (double)tp_n; xo[2] = (double)tp_j; xo[3] = (double)tp_c; xo[4] = (double)tp_bs;
        xo[-1] = (double)tp_w0;
        xo[-3] = (double)tp_a2; xo[-4] = (double)tp_bar; xo[-5] = (double)tp_pf; xo[-6] = (double)tp_sync; xo[-7] = (double)tp_np;
    }
    if (tid == 32) A.x[inst * A.xld + A.xld - 7] = (double)tp_w1;
#endif
    if (tid == 0) {
        A.ncols[inst] = n2;
        A.nblocks[inst] = done;
        A.status[inst] = st;
    }
}

static int g_blockqr_mode = 0;     // 0 = blocked reflectors (default), 1 = one reflector per pass (validation)
void qr_blockbanded_set_mode(int mode) { g_blockqr_mode = mode; }

int qr_blockbanded_smem_bytes(int maxblocks, int *hmax_out, int *wmax_out)
{
    int hmax = 8;
    while (hmax < 2 * maxblocks - 1) hmax <<= 1;
    const int wmax = 3 * maxblocks;
    if (hmax_out) *hmax_out = hmax;
    if (wmax_out) *wmax_out = wmax;
    size_t need = ((size_t)hmax * (wmax + 1) + 2 * (size_t)hmax * 4 + 32 + 6 * 72) * sizeof(double);
    const size_t backsub = (size_t)(64 * 65 + 64 + maxblocks * (maxblocks + 1) / 2 + 8) * sizeof(double);
    if (need < backsub) need = backsub;
    return (int)need;
}

int launch_qr_blockbanded(BlockQrArgs a, cudaStream_t s)
{
    if (a.batch <= 0) return 0;
    if (a.maxblocks < 3 || a.maxblocks > 64) return set_error(ILA_ERR_UNSUPPORTED, "qr_blockbanded: maxblocks must be in 3..64 (the panel lives in shared memory)");
    const int bytes = qr_blockbanded_smem_bytes(a.maxblocks, &a.hmax, &a.wmax);
    int dev = 0;
    ILA_CUDA_CHECK(cudaGetDevice(&dev));
    static int opted[16] = {};
    if (bytes > 48 * 1024 && (dev >= 16 || opted[dev] < bytes)) {        // the > 48 KB opt-in is per device
        ILA_CUDA_CHECK(cudaFuncSetAttribute(qr_blockbanded_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        ILA_CUDA_CHECK(cudaFuncSetAttribute(qr_blockbanded_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if (dev < 16) opted[dev] = bytes;
    }
    if (g_blockqr_mode == 1) qr_blockbanded_kernel<false><<<(unsigned)a.batch, kThreads, bytes, s>>>(a);
    else qr_blockbanded_kernel<true><<<(unsigned)a.batch, kThreads, bytes, s>>>(a);
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

ILA_BOUNDS_REPORT_FN(qr_blockbanded_bounds_report)

} // namespace ila
