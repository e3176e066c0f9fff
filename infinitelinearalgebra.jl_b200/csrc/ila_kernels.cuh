// ila_kernels.cuh — argument structures and launcher prototypes shared by the kernel files and api.cu (one definition
// each: api.cu used to restate them).
#pragma once
#include "ila_common.cuh"

namespace ila {

// from ql_toeplitz.cu
int launch_ql_toeplitz(bool real_path, int l, int u, const double *a_host, int branch_rank, const double *d_lambda,
                       int64_t n, double *d_L11, double *d_tail, int32_t *d_status, cudaStream_t stream);

void tz_set_spt(int v);
void tz_set_seeds(int v);
void tz_set_max_sweeps(int fp64_sweeps, int scout_sweeps);
int launch_ql_toeplitz_l11re(int l, int u, const double *a_host, int branch_rank, const double *d_lambda, int64_t n,
                             double *d_L11re, uint8_t *d_status8, cudaStream_t stream);
int launch_ql_toeplitz_solve(bool real_path, int l, int nb, const double *b_host, int64_t n, const double *d_tail,
                             const int32_t *d_status, int64_t nout, double *d_x, cudaStream_t stream);
int launch_ql_toeplitz_solve_fused(bool real_path, int l, int u, const double *a_host, int branch_rank, const double *d_lambda,
                                   int64_t n, int nb, const double *b_host, int64_t nout, double *d_L11, double *d_x,
                                   int32_t *d_status, cudaStream_t stream, int mode = 0, int64_t *d_nstop = nullptr);
int launch_ql_toeplitz_grid(int l, int u, const double *a_host, int branch_rank, const double *d_gx, int nx,
                            const double *d_gy, int64_t ny, double *d_absout, cudaStream_t stream);
int launch_ql_pert_head(bool real_path, int l, int p, const double *a_host, const double *head_host, const double *d_lambda,
                        int64_t n, const double *d_tail, const int32_t *d_status, double *d_L11, double *d_head_factors,
                        cudaStream_t stream);
int launch_ql_pert_head_ex(bool real_path, int l, int p, const double *a_host, const double *head_host, const double *d_lambda,
                           int64_t n, const double *d_tail, const int32_t *d_status, double *d_L11, double *d_head_factors,
                           int nb, const double *b_host, int64_t nout, double *d_x, cudaStream_t stream);

// from qr_banded.cu
struct QrFwdArgs {
    int64_t batch;
    const double *p, *q, *r, *shift, *head, *b;
    int hp, nb, colgrowth;
    double tol;
    const int64_t *gbase;
    const int64_t *ncap_per;
    int64_t ncap;
    double2 *work;
    int64_t *ncols, *nconv, *nswept;
    int32_t *status;
    double *state = nullptr;   // optional resumable state, qr_state_doubles(l,u) per instance (see qr_forward_kernel)
    int resume = 0;            // continue from `state` instead of column 0
    int factor_only = 0;       // partialqr!(F, ncap): sweep to column ncap, no right-hand side / convergence test
    int fault_every = 0;       // test hook of the three-warp sweep: the scout corrupts a candidate by one ulp every so many columns
};
int qr_state_doubles(int l, int u);
struct QrBwdArgs {
    int64_t batch;
    const int64_t *gbase;
    const double2 *work;
    double *xi;
    const int64_t *ncols, *nswept;
    const int32_t *status;
    double *xdirect;
    const int64_t *xoff;
};
int qr_record_doubles(int l, int u, bool refl);
void qr_set_arith_mode(int mode);
void qr_set_bulk_loader(int v);
void qr_set_fwd_duo(int v);
void qr_set_fwd_fault(int every);
int qr_get_arith_mode();
int launch_qr_forward(int l, int u, const QrFwdArgs &a, bool refl, cudaStream_t s);
int launch_qr_scan(const int64_t *d_ncols, int64_t batch, int64_t *d_xoff, cudaStream_t s);
int launch_qr_backsolve(int l, int u, const QrBwdArgs &a, bool refl, cudaStream_t s);
int launch_qr_compact(int64_t batch, const int64_t *d_gbase, const double *d_xi, const int64_t *d_ncols,
                      const int64_t *d_nswept, const int64_t *d_xoff, double *d_x, int64_t xcap, cudaStream_t s);
int launch_qr_export(int l, int u, int has_refl, int64_t batch, const int64_t *d_gbase, const double2 *d_work, const int64_t *d_ncols,
                     const int64_t *d_nswept, const int64_t *d_xoff, double *d_factors, double *d_tau, double *d_qtb,
                     cudaStream_t s);

// from qr_banded_c64.cu
struct QrC64FwdArgs {
    int64_t batch;
    const cplx *p, *q;
    const double *r;
    const cplx *shift, *head, *b;
    int hp, nb, colgrowth;
    double tol;
    const int64_t *gbase;
    const int64_t *ncap_per;
    int64_t ncap;
    double2 *work;
    int64_t *ncols, *nconv, *nswept;
    int32_t *status;
};
int qr_c64_record_chunks(int l, int u);
int launch_qr_forward_c64(int l, int u, const QrC64FwdArgs &a, cudaStream_t s);
int launch_qr_backsolve_c64(int l, int u, int64_t batch, const int64_t *gbase, const double2 *work, const int64_t *ncols,
                            const int64_t *nswept, const int64_t *xoff, double2 *x, cudaStream_t s);

// from revchol_tridiag.cu
struct RevCholArgs {
    int64_t batch, n;
    const double *ga, *gb;
    const double *shift;
    int hp;
    const double *a_head, *b_head;
    double *la, *lb;
    int64_t *N_used;
    int32_t *status;
};
int launch_revchol_tridiag(const RevCholArgs &a, cudaStream_t s);

// from triconj.cu
struct TriConjArgs {
    int64_t batch, n;
    int nbands;               // 2 or 3 leading entries of every row of U are used (U[k,k], U[k,k+1], U[k,k+2])
    // source 0: explicit bands U[i][d][k] (uld entries per band)
    const double *U;
    int64_t uld;
    // source 1 (work != NULL): K1 / K5 factor records, nch 16-byte chunks per record
    const int64_t *gbase;
    const double2 *work;
    int nch;
    const int32_t *status;    // per-operator status of the factorization (rows of failed operators are NaN) or NULL
    const double *X;          // [dl; d; du], xld entries per band; xstride = 0 (shared) or 3*xld
    int64_t xld, xstride;
    double *Y, *UX;           // batch x 3 x n (dl, d, du); UX optional (source 0 only)
};
int launch_triconj(const TriConjArgs &a, cudaStream_t s);
struct BiConjArgs {
    int64_t batch, n, ld;
    int upper;
    const double *U, *C;      // batch x 3 x ld: sub-, main, super-diagonal of U and of C = X*V
    double *dv, *ev;          // batch x n
};
int launch_biconj(const BiConjArgs &a, cudaStream_t s);

// from qr_blockbanded.cu
struct BlockQrArgs {
    int64_t batch;
    const double *alpha, *beta, *lam;     // per operator: L = alpha KronTrav(T1,I) + beta KronTrav(I,T2) - lam I
    const double *T1, *T2;                // shared 3 x tld bands (dl, d, du)
    int64_t tld;
    const double *b;                      // batch x nb
    int nb;
    double tol;
    int colgrowth, maxblocks, hmax, wmax;
    double *work;                         // batch x work_stride: R block rows
    int64_t work_stride;
    double *qtb, *x;                      // batch x xld
    int64_t xld;
    int64_t *ncols;
    int32_t *nblocks, *status;
};
int qr_blockbanded_smem_bytes(int maxblocks, int *hmax_out, int *wmax_out);
int launch_qr_blockbanded(BlockQrArgs a, cudaStream_t s);
void qr_blockbanded_set_mode(int mode);

// from qr_almostbanded.cu
struct AlmostBandedArgs {
    int64_t batch;
    int lD, uD, r;
    const double *dg;          // batch x (lD+uD+1) x 3: diagonals of D, +uD first, value(t) = g0 + g1 t + g2 t^2 (t = row of D)
    const double *bg;          // batch x r x 4: boundary rows B_m(j) = s^j (a + b j + c j^2)
    const double *shift;       // batch or NULL (subtracted from D's main diagonal)
    const double *b;           // batch x nb
    int nb, colgrowth;
    double tol;
    int64_t ncap;
    double *work;              // ceil(batch/32) x 32 x work_stride: records, interleaved per warp
    int64_t work_stride;       // ncap * (lD+uD+1 + r + 1)
    double *x;                 // batch x ncap
    int64_t *ncols, *nconv;
    int32_t *status;
};
int launch_qr_almostbanded(const AlmostBandedArgs &a, cudaStream_t s);

// from ql_finite_section.cu
template <typename T>
struct FsqlArgsT {
    int64_t batch, j, maxN;
    const T *g;               // batch x (l+u+1) x 5 generator coefficients (c0, p, q, r, s), data-row order (r, s: real parts used)
    const T *shift;           // batch or NULL
    int hp;
    const T *head;            // (l+u+1) x hp shared explicit head or NULL
    double tol;
    T *Lband, *V, *tau;       // batch x j x (l+u+1), batch x j x u, batch x j
    int64_t *N_used;
    int32_t *status;
};
using FsqlArgs = FsqlArgsT<double>;
using FsqlArgsC = FsqlArgsT<cplx>;
int launch_ql_finite_section(int l, int u, const FsqlArgs &a, cudaStream_t s);
int launch_ql_finite_section_c(int l, int u, const FsqlArgsC &a, cudaStream_t s);

// from chol_banded.cu
struct CholFwdArgs {
    int64_t batch;
    const double *p, *q, *r, *shift, *head, *b;
    int hp, nb, colgrowth, full_sweep;
    double tol;
    const int64_t *gbase;
    const int64_t *ncap_per;
    int64_t ncap;
    double2 *work;
    int64_t *ncols, *nconv, *nswept;
    int32_t *status;
    double *state = nullptr;   // optional resumable state (AdaptiveCholeskyFactors.ncols + the trailing block), chol_state_doubles(u) per instance
    int resume = 0;
};
int chol_state_doubles(int u);
int chol_banded_bounds_report(unsigned long long *out4);
int qr_banded_bounds_report(unsigned long long *out4);
int qr_blockbanded_bounds_report(unsigned long long *out4);
int launch_chol_forward(int u, const CholFwdArgs &a, cudaStream_t s);

// from ql_blocktri.cu
constexpr int kBtMaxN = 4;
constexpr int kBtMaxHead = 4;
template <typename T>
struct BlockTriParamsT {
    int n, n_dl, n_d, n_du, maxit;
    double atol;
    T c[kBtMaxN * kBtMaxN], a[kBtMaxN * kBtMaxN], b[kBtMaxN * kBtMaxN];
    T dl_head[kBtMaxHead][kBtMaxN * kBtMaxN];
    T d_head[kBtMaxHead][kBtMaxN * kBtMaxN];
    T du_head[kBtMaxHead][kBtMaxN * kBtMaxN];
};
// optional outputs of K6: the whole factorization (what the reference's QL object holds, src/infql.jl:203-204) and Q'b
constexpr int kBtMaxB = 16;
template <typename T>
struct BlockTriOutT {
    T *tail = nullptr;       // batch x (2n x 3n), column-major: the fixed point P after ql! with P[1:n,1:2n] = (d~, e~)  (:175-176)
    T *tau_tail = nullptr;   // batch x n: F.τ[n+1:2n]
    T *head = nullptr;       // batch x (Nn x Nn), column-major: the factored head (L in the lower triangle, reflectors above)
    T *head_tau = nullptr;   // batch x Nn
    T *qtb = nullptr;        // batch x nout: (Q'b)[1:nout] for the finitely supported right-hand side b  (:213-243)
    int nb = 0, nout = 0;
    T b[kBtMaxB];
};
using BlockTriParams = BlockTriParamsT<double>;
using BlockTriParamsC = BlockTriParamsT<cplx>;
int launch_ql_blocktri(const BlockTriParams &P, const double *d_energy, int64_t batch, double *d_L11, int32_t *d_iters,
                       int32_t *d_status, const BlockTriOutT<double> &O, cudaStream_t s);
int launch_ql_blocktri_c(const BlockTriParamsC &P, const cplx *d_energy, int64_t batch, cplx *d_L11, int32_t *d_iters,
                         int32_t *d_status, const BlockTriOutT<cplx> &O, cudaStream_t s);
static inline int launch_ql_blocktri_any(const BlockTriParams &P, const double *e, int64_t batch, double *L, int32_t *it,
                                         int32_t *st, cudaStream_t s, const BlockTriOutT<double> &O = BlockTriOutT<double>{})
{ return launch_ql_blocktri(P, e, batch, L, it, st, O, s); }
static inline int launch_ql_blocktri_any(const BlockTriParamsC &P, const cplx *e, int64_t batch, cplx *L, int32_t *it,
                                         int32_t *st, cudaStream_t s, const BlockTriOutT<cplx> &O = BlockTriOutT<cplx>{})
{ return launch_ql_blocktri_c(P, e, batch, L, it, st, O, s); }

} // namespace ila
