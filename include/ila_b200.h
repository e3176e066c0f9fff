/* ila_b200.h — C ABI of the B200-native adaptive-factorization hot path.
 *
 * The reference (InfiniteLinearAlgebra.jl v0.10.3) is pure Julia and has NO FFI; its "boundary" is
 * multiple dispatch.  Each entry point below replaces the narrow waist named beside it
 * and is what a `ccall` shim would bind (see INTEGRATION.md and julia/ILAB200.jl).  All citations are
 * relative to the reference checkout.
 *
 * Conventions
 *   - plain pointers and sizes; no torch / C++ types.  Host entry points take HOST pointers, own all
 *     device memory and streams, are synchronous, and shard the batch over `ngpus` devices of this
 *     process (ngpus <= 0: all visible devices; under one-process-per-GPU launches pass 1).
 *   - `_dev` entry points take DEVICE pointers valid on the current CUDA device and enqueue on
 *     `stream` (a cudaStream_t passed as void*; NULL = default stream) without synchronising.
 *   - return value: 0 on success, negative ILA_ERR_* on a library error (ila_last_error() has text).
 *     Per-instance outcomes (the reference's exceptions) go to status[]:
 *       0 ILA_OK
 *       1 ILA_NO_QL          DomainError "QL factorization does not exist"   (src/banded/infqltoeplitz.jl:13,23)
 *       2 ILA_NOT_REAL       DomainError "Real-valued QL ... does not exist" (src/banded/infqltoeplitz.jl:12)
 *       3 ILA_NOT_CONVERGED  ncap reached / "Did not converge"               (src/infql.jl:181)
 *       4 ILA_NOT_POSDEF     pbtrf info > 0 (the reference discards it,        src/infcholesky.jl:49)
 *       5 ILA_NAN            "Not-a-number encounted"                        (src/infqr.jl:324)
 *   - there is no CPU fallback: without a CUDA device every compute entry returns ILA_ERR_NOGPU.
 */
#ifndef ILA_B200_H
#define ILA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } ila_c64;

enum { ILA_OK = 0, ILA_NO_QL = 1, ILA_NOT_REAL = 2, ILA_NOT_CONVERGED = 3, ILA_NOT_POSDEF = 4, ILA_NAN = 5 };
enum { ILA_ERR_CUDA = -1, ILA_ERR_ARG = -2, ILA_ERR_NOGPU = -3, ILA_ERR_UNSUPPORTED = -4, ILA_ERR_CAPACITY = -5 };

/* ---- library ------------------------------------------------------------------------------- */
int ila_version(void);                 /* 10000*major + 100*minor + patch */
int ila_device_count(void);            /* number of visible CUDA devices (0 without a GPU) */
const char *ila_last_error(void);      /* text of the last error on the calling thread */
int ila_set_device(int dev);           /* select the CUDA device single-GPU calls (ngpus == 1, `_dev`) run on */
int ila_get_device(void);
int64_t ila_kernel_launches(void);     /* number of kernels this library launched so far (process-wide) */
int ila_release(void);                 /* frees the per-device buffer pools / streams the host entry points keep (re-created on demand) */
/* DFMA throughput probe: runs a register-resident FP64 FMA loop on the current device and returns the
 * achieved TFLOP/s (2 flop per FMA) — the FP64 roofline denominator (MEASURED_PEAKS.json has none). */
int ila_fp64_peak_probe(double *tflops_out, double *ms_out);
/* dependent-chain latencies on one warp, in SM cycles per op: out4[0] DFMA, [1] DMUL+DADD pair, [2] DSQRT+DADD,
 * [3] DDIV+DADD — what bounds the serial recurrences of K1/K2/K5. */
int ila_fp64_latency_probe(double *out4);

/* ---- K3: Toeplitz-tail QL, u == 1 -----------------------------------------------------------
 * L11[i] = ql(A - lambda[i]*I).L[1,1] for the pure Toeplitz operator A with bandwidths (l, 1).
 * Replaces, per shift: `A - λ*I` + ql(::InfToeplitz) -> _inf_ql -> ql_hessenberg(::InfToeplitz)
 * (src/banded/infqltoeplitz.jl:56-65,182-191) = tail_de (:7-27) + ql_X! (:31-39) + L[1,1] read (src/infql.jl:96).
 *   a          l+u+1 coefficients in BandedMatrices data-column order: a_{+1}, a_0, a_{-1}, ..., a_{-l}
 *   lambda     n shifts;  branch_rank 1 = `findmax` (default), k = k-th largest |mu|^2 (examples/toeplitz.jl:22-27)
 *   L11        n results (NaN where status != 0)
 *   tail_opt   NULL or n x (2m+2) values, m = l+2:  X[1,1..m], X[2,1..m], tau[1], tau[2]  after ql_X!
 *              (first L column = [X[1,m-1]; X[2,m-1:-1:1]], tail L column = X[2,m:-1:1], 2x2 Q = QLPackedQ(X[:,m-1:m],tau))
 * _f64 is the `T<:Real` method (tail_de :7-16): real dominant eigenvalue required (else ILA_NOT_REAL).
 */
int ila_ql_toeplitz_l11_c64(int l, int u, const ila_c64 *a, const ila_c64 *lambda, int64_t n, int branch_rank,
                            ila_c64 *L11, ila_c64 *tail_opt, int32_t *status, int ngpus);
int ila_ql_toeplitz_l11_f64(int l, int u, const double *a, const double *lambda, int64_t n, int branch_rank,
                            double *L11, double *tail_opt, int32_t *status, int ngpus);
/* Compact form for complex symbols / shifts.  L[1,1] of a ComplexF64 factorization is REAL: ql_X!'s k = 1 reflector
 * leaves X[1,m-1] = -nu (src/banded/infqltoeplitz.jl:33 -> ql!, the `(T<:Real)` bound of examples/jacobi.jl:63) and the
 * sign fix (:34-37) only negates it — so the imaginary part the _c64 form returns is a literal 0.  This form returns
 *   L11re       n doubles; where status != 0 a quiet NaN whose low mantissa bits hold the status code
 *               (bits = 0x7ff8000000000000 | status)
 *   status8_opt NULL, or n bytes with the same codes
 * 16 B in, 8 (+1) B out per shift instead of 16 + 20: the end-to-end call is PCIe-bound. */
int ila_ql_toeplitz_l11re_c64(int l, int u, const ila_c64 *a, const ila_c64 *lambda, int64_t n, int branch_rank,
                              double *L11re, uint8_t *status8_opt, int ngpus);
int ila_ql_toeplitz_l11re_c64_dev(int l, int u, const ila_c64 *a, const ila_c64 *d_lambda, int64_t n, int branch_rank,
                                  double *d_L11re, uint8_t *d_status8_opt, void *stream);

/* Host buffers.  The K3 host entry points accept ANY host pointer.  Pinned memory (ila_host_alloc, ila_host_register,
 * cudaHostAlloc, a torch pinned tensor) is copied by DMA straight from/to the caller's arrays, chunked over several
 * streams.  Ordinary pageable memory (a Julia Vector passed by ccall) goes through the library's own pinned staging
 * ring with worker threads (mode 0 = automatic); ila_set_host_copy_mode(1) forces plain cudaMemcpyAsync on the caller's
 * pointers, (2) forces the staging ring.  ila_host_register pins an existing allocation in place (unregister it before
 * freeing it). */
int ila_host_alloc(void **ptr, int64_t bytes);
int ila_host_free(void *ptr);
int ila_host_register(void *ptr, int64_t bytes);
int ila_host_unregister(void *ptr);
int ila_set_host_copy_mode(int mode);
/* Chunk schedule of the pinned-buffer pipeline of the K3 host call: chunk = remaining / divisor, never below
 * min_chunk_shifts, the last chunk takes the remainder (0, 0 = defaults: 64 Ki shifts, divisor 3). */
int ila_set_host_chunking(int64_t min_chunk_shifts, int divisor);

/* device-resident forms: `a` stays a HOST pointer (l+u+1 values, passed by value to the kernel) */
int ila_ql_toeplitz_l11_c64_dev(int l, int u, const ila_c64 *a, const ila_c64 *d_lambda, int64_t n, int branch_rank,
                                ila_c64 *d_L11, ila_c64 *d_tail_opt, int32_t *d_status, void *stream);
int ila_ql_toeplitz_l11_f64_dev(int l, int u, const double *a, const double *d_lambda, int64_t n, int branch_rank,
                                double *d_L11, double *d_tail_opt, int32_t *d_status, void *stream);

/* Grid form = the λ loop of the reference's examples (examples/toeplitz.jl:14-33: `ℓ11.(Ref(A), x' .+ y.*im)`):
 * out[k*nx + j] = abs(ql(A - (x[j] + i*y[k])*I).L[1,1]), or -1.0 where the factorization does not exist (the example's
 * `catch` branch).  Shifts are formed on device; 8 B/shift come back. */
int ila_ql_toeplitz_absl11_grid_c64(int l, int u, const ila_c64 *a, const double *x, int nx, const double *y, int64_t ny,
                                    int branch_rank, double *out, int ngpus);
int ila_ql_toeplitz_absl11_grid_c64_dev(int l, int u, const ila_c64 *a, const double *d_x, int nx, const double *d_y, int64_t ny,
                                        int branch_rank, double *d_out, void *stream);

/* ---- K3+K4: perturbed Toeplitz, u == 1:  L11[i] = ql(A - lambda[i]*I).L[1,1],  A = finite head + Toeplitz tail ----
 * Replaces ql(::PertToeplitz) -> ql_hessenberg!(B) (src/infql.jl:60-93): tail fixed point (K3), last head row rebuilt
 * from row 1 of Q∞' (:81-82), finite Householder QL of the p x p head (:83).
 *   head   p x p top-left corner of A (column-major, WITHOUT the shift), l+1 <= p <= 8; row p of A must already be pure
 *          Toeplitz (the reference takes p from InfiniteArrays' PertToeplitz constructor — SURVEY A.3)
 *   a      l+u+1 tail coefficients in BandedMatrices data-column order (as for ila_ql_toeplitz_l11_*)
 *   head_factors_opt  NULL or n x (p x p): the factored head (L in the lower triangle, reflectors above)
 */
int ila_ql_perttoeplitz_l11_c64(int l, int u, int p, const ila_c64 *head, const ila_c64 *a, const ila_c64 *lambda, int64_t n,
                                int branch_rank, ila_c64 *L11, ila_c64 *head_factors_opt, int32_t *status, int ngpus);
int ila_ql_perttoeplitz_l11_f64(int l, int u, int p, const double *head, const double *a, const double *lambda, int64_t n,
                                int branch_rank, double *L11, double *head_factors_opt, int32_t *status, int ngpus);

/* Shifts per thread of the K3 kernel: consecutive shifts of one thread warm-start the root finder from the previous
 * shift's roots (any doubt falls back to a cold simultaneous solve, so results do not depend on it).
 * 0 = automatic (default), 1 = every shift cold. */
int ila_ql_set_shifts_per_thread(int spt);
/* Sweep limits of the K3 root finder (0 = defaults: 48 FP64 Aberth sweeps, 16 FP32 scout sweeps).  A shift whose
 * iteration uses up its sweeps without reaching a root to rounding (backward error test) gets status
 * ILA_NOT_CONVERGED and NaN — never a silent value.  Lowering the limits is a test hook that forces that outcome. */
int ila_ql_set_max_sweeps(int fp64_sweeps, int scout_sweeps);
/* Cold start of a run's first shift: 1 (default) = closed-form FP32 seeds where the degree has them (l = 1: quadratic
 * formula, l = 3: Ferrari's resolvent) polished by the usual sweeps; 0 = circle start + Aberth sweeps for every degree.
 * Results do not depend on it beyond the last bits of the selected root's Newton refinement (validation knob). */
int ila_ql_set_cold_start(int closed_form_seeds);
/* compute-sanitizer substitute.  `make -C csrc debug` builds libila_b200_dbg.so with -DILA_BOUNDS: the kernels whose index
 * arithmetic is not trivial (K1/K2 group-interleaved records and ring slots, K5 records, K11 circular shared-memory panel,
 * R slab, back-substitution scratch) then address their buffers through bounds-checked pointers that record every
 * out-of-range index instead of performing it.  This returns (and clears) what they recorded: count (-1 in the release
 * build: not instrumented), and site id / index / extent of the first failure. */
int ila_debug_bounds_report(int64_t *count, int64_t *site, int64_t *index, int64_t *extent);

/* ---- K1+K2: adaptive banded QR solve  x_i = A_i \ b_i ----------------------------------------
 * Replaces qr(A) -> adaptiveqr (src/infqr.jl:163-174) and `\` = ldiv!(F.R, F.Q'*b) (src/infqr.jl:370-374):
 * AdaptiveQRData/partialqr! (:2-13,32-48) -> _banded_qr!, the Q'b adaptive driver (:236-274) and the
 * back-substitution (:350-355), fused per instance with on-device convergence detection.
 *
 * Operator i (bandwidths l,u shared by the batch) is described by generators, one per diagonal, in
 * BandedMatrices data-row order rr = 0..l+u  <->  diagonal d = u - rr (d = col - row):
 *     entry at position t (0-based, along the diagonal, as in `BandedMatrix(d => v)`)
 *         = head[rr*hp + t]                          for t < hp      (shared explicit head, may be NULL/hp = 0)
 *         = (p[i][rr] + q[i][rr]*t) / r[i][rr]       otherwise       (r == NULL means 1)
 *     and shift[i] (NULL = 0) is subtracted on the main diagonal (A_i - shift_i*I).
 *   b          batch x nb (instance-major): the non-zero prefix of the right-hand sides
 *   tol        convergence tolerance on max|(Q'b)[j : j+l]| (reference default floatmin, infqr.jl:236)
 *   colgrowth  the reference's COLGROWTH = 1000 (infqr.jl:238): the test runs at columns 1, 1+cg, 1+2cg, ...
 *   ncap       per-instance column capacity (uniform) — or ncap_per[batch] when non-NULL
 *   x, xoff    ragged output: x[xoff[i] .. xoff[i+1]) = solution prefix of instance i, length ncols[i]
 *              (= the reference's B.datasize: n_conv + colgrowth - 1 + l); xcap = capacity of x in doubles.
 *              If the total exceeds xcap the call returns ILA_ERR_CAPACITY and xoff[batch] holds the need.
 *   nconv      1-based column at which the convergence test passed (0 if it never ran / l == 0)
 *   factors_opt NULL or ragged (2l+u+1) x ncols[i] column-major band storage (R above, reflectors below;
 *              bandwidths (l, l+u), src/infqr.jl:11) at offset (2l+u+1)*xoff[i]; tau_opt, qtb_opt ragged like x.
 */
int ila_qr_banded_solve_f64(int l, int u, int64_t batch, const double *p, const double *q, const double *r,
                            const double *shift, int hp, const double *head, int nb, const double *b, double tol,
                            int colgrowth, int64_t ncap, const int64_t *ncap_per, double *x, int64_t xcap,
                            int64_t *xoff, int64_t *ncols, int64_t *nconv, double *factors_opt, double *tau_opt,
                            double *qtb_opt, int32_t *status, int ngpus);

/* Arithmetic of the sweeps: 0 (default) = branch-free division / square-root sequences that reproduce the IEEE
 * results bit for bit inside an exponent box (redo with intrinsics outside it); 1 = plain __ddiv_rn/__dsqrt_rn
 * (validation mode: both must give identical bits).  Modes 2 / 3 / 4 select how the back-substitution delivers x: auto
 * (direct ragged stores when the batch is at most two warps per SM, coalesced interleaved stream + compaction otherwise) /
 * always compaction / always direct — all three must give identical bits too.  Modes 5 / 6 select the loader of the
 * back-substitution ring: per-lane 16-byte cp.async (LDGSTS) / one cp.async.bulk per group completing on an mbarrier (TMA;
 * default) — same bits.  Modes 7 / 8 / 11 select the forward sweep of l = 1
 * operators (plain solves: no factor export, no resumable state, no head table): one warp per group of 32 instances / three
 * warps per group when the batch is at most one group per SM, the latency-bound regime (default) / three warps whatever the
 * batch size — a scout running a shortened dependent chain, a verifier that proves every candidate of the scout correctly
 * rounded by exact residual tests before it is used, a clerk for Q'b, the adaptive test and the records — same bits again.
 * Modes 1000 + n are a test hook: the scout corrupts one candidate by an ulp every n columns (n = 0: off), which the verifier
 * must catch, repair with the IEEE column and answer by re-seeding the scout (or, when that happens too often, by going on
 * alone) — the results must not change. */
int ila_qr_set_arith_mode(int mode);

/* device-resident two-phase form (no host round trip between the phases; the caller owns all buffers):
 *   phase 1 `_factor_dev`: forward sweep + Q'b + convergence; writes d_ncols, d_nconv, d_nswept, d_status, d_xoff (scan).
 *   phase 2 `_backsolve_dev`: back-substitution + compaction into the ragged d_x using d_xoff.
 * Workspace: instances are swept in groups of 32 (one warp); `ila_qr_banded_work_layout` (host helper) fills
 * gbase_host[(batch+31)/32 + 1] (group bases, in rows) and reports the bytes of d_work (records, coalesced
 * group-interleaved layout, see csrc/qr_banded.cu) and d_xi (interleaved solution).  d_ncap_per may be NULL (uniform ncap). */
int ila_qr_banded_work_layout(int l, int u, int64_t batch, int64_t ncap, const int64_t *ncap_per, int64_t *gbase_host,
                              int64_t *work_bytes, int64_t *xi_bytes);
int ila_qr_banded_factor_f64_dev(int l, int u, int64_t batch, const double *d_p, const double *d_q, const double *d_r,
                                 const double *d_shift, int hp, const double *d_head, int nb, const double *d_b,
                                 double tol, int colgrowth, int64_t ncap, const int64_t *d_ncap_per,
                                 const int64_t *d_gbase, void *d_work, int64_t *d_ncols, int64_t *d_nconv,
                                 int64_t *d_nswept, int64_t *d_xoff, int32_t *d_status, int want_reflectors, void *stream);
/* Resumable factorization state — the reference's AdaptiveQRData (src/infqr.jl:2-13): `ncols` plus the trailing columns
 * that already carry the reflector updates, from which partialqr!(F, n) continues (:32-48; test/test_infqr.jl:19-28).
 * Here: d_state, ila_qr_banded_state_doubles(l,u) doubles per instance (the register window, the Q'b window, counters).
 *   ila_qr_banded_factor_resume_f64_dev  = ila_qr_banded_factor_f64_dev that, when the column limit `ncap` is reached,
 *       keeps what it has: status 3, ncols = nswept = columns finished, state saved.  Calling it again with a larger
 *       ncap (work buffer laid out for the larger capacity from the start: ila_qr_banded_work_layout) and resume = 1
 *       continues from there; instances that had finished are left untouched.  Two calls give the same bits as one.
 *   ila_qr_banded_partialqr_f64_dev      = partialqr!(F, n): factorize through column n (no right-hand side, no
 *       convergence test; reflectors and tau are stored), resume = 1 continues from the state of an earlier call.  The
 *       window left in d_state, W[j][i] = (Q_n' A)[n+i, n+j], is the reference's "extra columns have been modified".
 *   ila_qr_banded_export_f64_dev         = records -> BandedMatrices band storage ((2l+u+1) x ncols, bandwidths (l, l+u)),
 *       tau, Q'b; ragged at d_xoff like the host entry's factors_opt. */
int ila_qr_banded_state_doubles(int l, int u);
int ila_qr_banded_factor_resume_f64_dev(int l, int u, int64_t batch, const double *d_p, const double *d_q, const double *d_r,
                                        const double *d_shift, int hp, const double *d_head, int nb, const double *d_b,
                                        double tol, int colgrowth, int64_t ncap, const int64_t *d_ncap_per,
                                        const int64_t *d_gbase, void *d_work, int64_t *d_ncols, int64_t *d_nconv,
                                        int64_t *d_nswept, int64_t *d_xoff, int32_t *d_status, int want_reflectors,
                                        double *d_state, int resume, void *stream);
int ila_qr_banded_partialqr_f64_dev(int l, int u, int64_t batch, const double *d_p, const double *d_q, const double *d_r,
                                    const double *d_shift, int hp, const double *d_head, int64_t n, int colgrowth,
                                    const int64_t *d_gbase, void *d_work, double *d_state, int resume, int64_t *d_ncols,
                                    int64_t *d_nswept, int64_t *d_xoff, int32_t *d_status, void *stream);
int ila_qr_banded_export_f64_dev(int l, int u, int want_reflectors, int64_t batch, const int64_t *d_gbase, const void *d_work,
                                 const int64_t *d_ncols, const int64_t *d_nswept, const int64_t *d_xoff, double *d_factors,
                                 double *d_tau, double *d_qtb, void *stream);
int ila_qr_banded_backsolve_f64_dev(int l, int u, int64_t batch, const int64_t *d_gbase, const void *d_work, double *d_xi,
                                    const int64_t *d_ncols, const int64_t *d_nswept, const int64_t *d_xoff,
                                    const int32_t *d_status, double *d_x, int64_t xcap, int want_reflectors, void *stream);

/* ---- K5+K2: adaptive banded Cholesky solve  x_i = S_i \ b_i,  S_i symmetric positive definite, banded --------------
 * Replaces cholesky(S) -> adaptivecholesky (src/infcholesky.jl:71-75), partialcholesky! (:42-59, banded_chol! = LAPACK
 * pbtrf 'U' per COLGROWTH block + the U1'\B / B'B hand-off), the adaptive L \ b (:94-133, tolerance floatmin, coupling
 * between chunks) and U \ b (:135-140), fused per instance with on-device convergence detection.
 * The operator is described by the generators of its UPPER bands only (data rows 0..u <-> diagonals +u..0, same
 * generator form as the QR entry; position t = row index).  Outputs as for the QR entry; factors_opt is the ragged
 * (u+1) x ncols[i] column-major band storage of U, y_opt = L \ b before the back-substitution. */
int ila_chol_banded_solve_f64(int u, int64_t batch, const double *p, const double *q, const double *r,
                              const double *shift, int hp, const double *head, int nb, const double *b, double tol,
                              int colgrowth, int64_t ncap, const int64_t *ncap_per, double *x, int64_t xcap,
                              int64_t *xoff, int64_t *ncols, int64_t *nconv, double *factors_opt, double *y_opt,
                              int32_t *status, int ngpus);
/* device-resident phase 1 (phase 2 is ila_qr_banded_backsolve_f64_dev with l = 0: same record layout);
 * workspace from ila_qr_banded_work_layout(0, u, ...).  full_sweep != 0 factorizes all ncols columns (factor export). */
int ila_chol_banded_factor_f64_dev(int u, int64_t batch, const double *d_p, const double *d_q, const double *d_r,
                                   const double *d_shift, int hp, const double *d_head, int nb, const double *d_b,
                                   double tol, int colgrowth, int64_t ncap, const int64_t *d_ncap_per,
                                   const int64_t *d_gbase, void *d_work, int64_t *d_ncols, int64_t *d_nconv,
                                   int64_t *d_nswept, int64_t *d_xoff, int32_t *d_status, int full_sweep, void *stream);

/* Resumable form — AdaptiveCholeskyFactors.ncols and the Schur-updated next block partialcholesky! continues from
 * (src/infcholesky.jl:42-59) plus the forward-substitution carry (:126-128): d_state holds
 * ila_chol_banded_state_doubles(u) doubles per instance; semantics as ila_qr_banded_factor_resume_f64_dev. */
int ila_chol_banded_state_doubles(int u);
int ila_chol_banded_factor_resume_f64_dev(int u, int64_t batch, const double *d_p, const double *d_q, const double *d_r,
                                          const double *d_shift, int hp, const double *d_head, int nb, const double *d_b,
                                          double tol, int colgrowth, int64_t ncap, const int64_t *d_ncap_per,
                                          const int64_t *d_gbase, void *d_work, int64_t *d_ncols, int64_t *d_nconv,
                                          int64_t *d_nswept, int64_t *d_xoff, int32_t *d_status, int full_sweep,
                                          double *d_state, int resume, void *stream);

/* Complex eltype of the adaptive QR solve (e.g. a real operator minus a complex shift: the resolvent (A - λI) \ b by QR).
 * p, q, shift, head, b and x are complex; r (the per-diagonal denominators) is real; everything else as in
 * ila_qr_banded_solve_f64 (no factor export).  Same reference path with LinearAlgebra.reflector!/reflectorApply! in their
 * complex form; complex products are the four-multiply form and quotients Smith's algorithm, unfused, so the result equals
 * the C oracle (oracle/ila_oracle_c64.c) bit for bit.  Supported (l,u): (1,1),(1,2),(2,1),(2,2),(3,1),(1,3),(3,3). */
int ila_qr_banded_solve_c64(int l, int u, int64_t batch, const ila_c64 *p, const ila_c64 *q, const double *r,
                            const ila_c64 *shift, int hp, const ila_c64 *head, int nb, const ila_c64 *b, double tol,
                            int colgrowth, int64_t ncap, const int64_t *ncap_per, ila_c64 *x, int64_t xcap, int64_t *xoff,
                            int64_t *ncols, int64_t *nconv, int32_t *status, int ngpus);

/* ---- K7: solve with the Toeplitz QL,  x[i, 0:nout] = (ql(A - lambda[i]*I) \ b)[1:nout] -------------------------------
 * Pure Toeplitz A (u == 1) as in K3; b has nb <= 16 leading entries, zero beyond.  Replaces
 * ldiv!(F, b) = ldiv!(F.L, lmul!(F.Q', b)) (src/infql.jl:266-267): UpperHessenbergQ apply for n = nb..1
 * (src/banded/hessenbergq.jl:125-135) and the column-oriented forward substitution with the ∞ lower band L
 * (src/infql.jl:269-289; the reference stops at exact zeros, here the caller states nout).  x is COLUMN-major
 * n x nout (x[i + j*n]); rows with status != 0 are NaN.  x[i, 0] for b = e1 is the resolvent entry
 * e1'(A - lambda[i] I)^{-1} e1.  One fused launch: the K3 kernel solves from its registers.  The _dev forms need
 * d_work of n scalars (it receives L[1,1]). */
int ila_ql_toeplitz_solve_c64(int l, int u, const ila_c64 *a, const ila_c64 *lambda, int64_t n, int branch_rank, int nb,
                              const ila_c64 *b, int64_t nout, ila_c64 *x, int32_t *status, int ngpus);
int ila_ql_toeplitz_solve_f64(int l, int u, const double *a, const double *lambda, int64_t n, int branch_rank, int nb,
                              const double *b, int64_t nout, double *x, int32_t *status, int ngpus);
int ila_ql_toeplitz_solve_c64_dev(int l, int u, const ila_c64 *a, const ila_c64 *d_lambda, int64_t n, int branch_rank, int nb,
                                  const ila_c64 *b, int64_t nout, ila_c64 *d_work, ila_c64 *d_x, int32_t *d_status, void *stream);
int ila_ql_toeplitz_solve_f64_dev(int l, int u, const double *a, const double *d_lambda, int64_t n, int branch_rank, int nb,
                                  const double *b, int64_t nout, double *d_work, double *d_x, int32_t *d_status, void *stream);

/* Q b and Q'b alone for the ∞ Hessenberg Q of ql(A - lambda[i] I), Q = LowerHessenbergQ(Fill(q, ∞)), q the 2 x 2 block
 * QLPackedQ(X[:, m-1:m], tau) (src/banded/infqltoeplitz.jl:63-64):
 *   adjoint == 0:  y = Q b, materialize!(::MatLmulVec{HessenbergQLayout{'L'}}) (src/banded/hessenbergq.jl:111-123): for
 *                  n = 1, 2, ...: x[n:n+1] <- q x[n:n+1], stopping once n > nb and norm(x[n:n+1]) <= 10 floatmin; nstop_opt[i]
 *                  receives that n (0 when it did not happen within the nout entries asked for);
 *   adjoint != 0:  y = Q'b, UpperHessenbergQ(adjoint.(q)) for n = nb, ..., 1 (:125-135), support nb + 1.
 * b: nb <= 16 leading entries; y COLUMN-major n x nout, NaN rows where status != 0. */
int ila_ql_toeplitz_applyq_c64(int l, int u, const ila_c64 *a, const ila_c64 *lambda, int64_t n, int branch_rank, int adjoint,
                               int nb, const ila_c64 *b, int64_t nout, ila_c64 *y, int64_t *nstop_opt, int32_t *status, int ngpus);
int ila_ql_toeplitz_applyq_f64(int l, int u, const double *a, const double *lambda, int64_t n, int branch_rank, int adjoint,
                               int nb, const double *b, int64_t nout, double *y, int64_t *nstop_opt, int32_t *status, int ngpus);

/* Perturbed-Toeplitz form of the solve: head / a / p as in ila_ql_perttoeplitz_l11_*.  Replaces the assembly of
 * ql_hessenberg! (src/infql.jl:70-93: head factors, the l' extension rows (Q∞')[2:l'+1,1:l'+1] T[l':2l',1:l'], Toeplitz tail
 * columns; Q = Vcat(LowerHessenbergQ(F.Q).q, F∞.q), src/banded/hessenbergq.jl:182-190) followed by F \ b
 * (src/infql.jl:266-289).  Pinned by test/test_infql.jl:147-160. */
int ila_ql_perttoeplitz_solve_c64(int l, int u, int p, const ila_c64 *head, const ila_c64 *a, const ila_c64 *lambda, int64_t n,
                                  int branch_rank, int nb, const ila_c64 *b, int64_t nout, ila_c64 *x, int32_t *status, int ngpus);
int ila_ql_perttoeplitz_solve_f64(int l, int u, int p, const double *head, const double *a, const double *lambda, int64_t n,
                                  int branch_rank, int nb, const double *b, int64_t nout, double *x, int32_t *status, int ngpus);

/* ---- K6: block-tridiagonal perturbed-Toeplitz QL,  L11[i] = ql(A - energy[i]*I).L[1,1] ----------------------------
 * A = BlockTridiagonal(Vcat(dl_head, Fill(c,∞)), Vcat(d_head, Fill(a,∞)), Vcat(du_head, Fill(b,∞))) with n x n blocks
 * (column-major; dl = A[Block(k+1,k)], du = A[Block(k,k+1)]; head tables hold n_dl / n_d / n_du blocks, n <= 4,
 * at most 4 head block columns).  Replaces ql(::BlockTriPertToeplitz) -> _blocktripert_ql (src/infql.jl:189-209):
 * blocktailiterate (:166-182; dense 2n x 3n Householder QL per iteration, stop when (d,e) move by <= atol, reference
 * atol = 1e-10, maxit = 1_000_000) and the finite block head QL (:193-198).  iters[i] = fixed-point iterations taken;
 * status 3 = "Did not converge" (src/infql.jl:181), e.g. inside the essential spectrum. */
int ila_ql_blocktri_l11_f64(int n, const double *c, const double *a, const double *b, int n_dl, const double *dl_head,
                            int n_d, const double *d_head, int n_du, const double *du_head, const double *energy,
                            int64_t batch, double atol, int maxit, double *L11, int32_t *iters, int32_t *status, int ngpus);
int ila_ql_blocktri_l11_f64_dev(int n, const double *c, const double *a, const double *b, int n_dl, const double *dl_head,
                                int n_d, const double *d_head, int n_du, const double *du_head, const double *d_energy,
                                int64_t batch, double atol, int maxit, double *d_L11, int32_t *d_iters, int32_t *d_status,
                                void *stream);

/* Complex eltype (blocks and/or energies complex, e.g. `A - 5im*I`, test/test_periodic.jl:20-26): the QL loop runs down
 * to k = 1 and L[1,1] is complex.  Parity bar 1e-12 against the oracle (see csrc/ql_blocktri.cu header). */
int ila_ql_blocktri_l11_c64(int n, const ila_c64 *c, const ila_c64 *a, const ila_c64 *b, int n_dl, const ila_c64 *dl_head,
                            int n_d, const ila_c64 *d_head, int n_du, const ila_c64 *du_head, const ila_c64 *energy,
                            int64_t batch, double atol, int maxit, ila_c64 *L11, int32_t *iters, int32_t *status, int ngpus);
int ila_ql_blocktri_l11_c64_dev(int n, const ila_c64 *c, const ila_c64 *a, const ila_c64 *b, int n_dl, const ila_c64 *dl_head,
                                int n_d, const ila_c64 *d_head, int n_du, const ila_c64 *du_head, const ila_c64 *d_energy,
                                int64_t batch, double atol, int maxit, ila_c64 *d_L11, int32_t *d_iters, int32_t *d_status,
                                void *stream);

/* The WHOLE block factorization and Q'b — what the reference's QL object holds (src/infql.jl:199-204) and what
 * lmul!(Q', b) computes (src/infql.jl:213-243); compared in test/test_periodic.jl:12-18.  Same operator description.  With
 * N = max(n_du + 1, n_d, n_dl) head block columns (src/infql.jl:190):
 *   tail_opt      batch x (2n x 3n), column-major: P = the panel [c a b; 0 d e] after ql! with P[1:n,1:2n] = (d~, e~)
 *                 (:175-176).  Tail block column of the ∞ factors = [P[1:n,2n+1:3n]; P[n+1:2n,2n+1:3n]; P[n+1:2n,n+1:2n];
 *                 P[n+1:2n,1:n]] (:203); the rows below the head are P[n+1:2n,1:2n] and P[n+1:2n,1:n] (:199-200)
 *   tau_tail_opt  batch x n: τ of every tail block column (F.τ[n+1:2n], :176)
 *   head_opt      batch x (Nn x Nn), column-major: ql! of the head with its last block row replaced by (d~, e~) (:197-198)
 *   head_tau_opt  batch x Nn
 *   qtb_opt       batch x nout: (Q'b)[1:nout] for b = rhs[0..nb) (nb <= 16, nout <= 16 + 2n + 1), shared by the batch
 * Real eltype: bit-exact against the oracle; NaN where status != 0. */
int ila_ql_blocktri_factors_f64(int n, const double *c, const double *a, const double *b, int n_dl, const double *dl_head,
                                int n_d, const double *d_head, int n_du, const double *du_head, const double *energy,
                                int64_t batch, double atol, int maxit, double *L11, double *tail_opt, double *tau_tail_opt,
                                double *head_opt, double *head_tau_opt, int nb, const double *rhs, int64_t nout, double *qtb_opt,
                                int32_t *iters, int32_t *status, int ngpus);
int ila_ql_blocktri_factors_c64(int n, const ila_c64 *c, const ila_c64 *a, const ila_c64 *b, int n_dl, const ila_c64 *dl_head,
                                int n_d, const ila_c64 *d_head, int n_du, const ila_c64 *du_head, const ila_c64 *energy,
                                int64_t batch, double atol, int maxit, ila_c64 *L11, ila_c64 *tail_opt, ila_c64 *tau_tail_opt,
                                ila_c64 *head_opt, ila_c64 *head_tau_opt, int nb, const ila_c64 *rhs, int64_t nout,
                                ila_c64 *qtb_opt, int32_t *iters, int32_t *status, int ngpus);

/* ---- K8: adaptive reverse Cholesky  A = L'L  of ∞ symmetric tridiagonal operators ------------------
 * Replaces reversecholesky(::SymTridiagonal) -> LazySymTridiagonalReverseCholeskyFactors
 * (src/banded/infreversecholeskytridiagonal.jl:66-141): finite reverse Cholesky of the N x N section (:118-129), convergence
 * measure ξ (:81-110, tolerance eps), N doubled from 4n until ξ <= eps or N > 2^21 - 1 (status 3; the reference warns).
 * Operator i: diagonal value(k) = c0 + (p + q k)/(r + s k) from ga[i*5 .. i*5+4] = (c0,p,q,r,s), minus shift[i];
 * off-diagonal the same from gb; the first hp entries of both come from a_head / b_head when hp > 0 (shared).
 * Output: la[i*n + k] = L[k+1,k+1], lb[i*n + k] = L[k+2,k+1] for k < n (0-based), N_used[i] the section size at which
 * the test fired.  status 4 = non-positive pivot (DomainError upstream), rows NaN. */
int ila_revchol_tridiag_f64(int64_t batch, const double *ga, const double *gb, const double *shift, int hp,
                            const double *a_head, const double *b_head, int64_t n, double *la, double *lb, int64_t *N_used,
                            int32_t *status, int ngpus);

/* ---- K9: adaptive finite-section QL of ∞ banded operators WITHOUT a Toeplitz tail (real eltype) -------
 * Replaces ql(A::InfBandedMatrix, tol) (src/infql.jl:385-388) = initialadaptiveQLblock / cache_filldata! (:391-411,
 * 440-458): the block L[2:j-checkinds+1, 2:j-checkinds+1] of ql(A[checkinds:N, checkinds:N]) is compared (Frobenius)
 * between N and 2N, N doubling from j until the difference is <= tol (status 3 = "Reached max. iterations in adaptive QL
 * without convergence to desired tolerance" once N >= maxN; reference: maxN = 10000 for j = 50); the result is the
 * leading j x j part of ql(A[1:N/2, 1:N/2]).  Operator i: diagonal d = col - row (data row rr = u - d, +u first) has
 * value(t) = c0 + (p + q t)/(r + s t), t = min(row, col) 0-based, from g[((i*(l+u+1) + rr)*5 ..+4] = (c0,p,q,r,s);
 * optional shared explicit head (l+u+1) x hp; shift[i] subtracted from the main diagonal.
 * Outputs: Lband[(i*j + k)*(l+u+1) + c] = L[k, k-c]; V[(i*j + k)*u + r-1] = reflector entry of column k at row k-r;
 * tau[i*j + k]; N_used[i] = size of the final section.  Supported (l,u): (1,1),(1,2),(2,1),(2,2),(3,1),(1,3). */
int ila_ql_finite_section_f64(int l, int u, int64_t batch, const double *g, const double *shift, int hp, const double *head,
                              int64_t j, double tol, int64_t maxN, double *Lband, double *V, double *tau, int64_t *N_used,
                              int32_t *status, int ngpus);
/* Complex eltype (test/test_infql.jl:257-262): coefficients c0, p, q, the head, the shift and the outputs are complex; the
 * real parts of r, s are used; the QL loop then runs down to the section's first column. */
int ila_ql_finite_section_c64(int l, int u, int64_t batch, const ila_c64 *g, const ila_c64 *shift, int hp, const ila_c64 *head,
                              int64_t j, double tol, int64_t maxN, ila_c64 *Lband, ila_c64 *V, ila_c64 *tau, int64_t *N_used,
                              int32_t *status, int ngpus);

/* ---- K10: tridiagonal conjugation  Y = Tridiagonal(U*X/U)  and bidiagonal conjugation ---------------------
 * Replaces upper_mul_tri_triview! / tri_mul_invupper_triview! and TridiagonalConjugationData + resizedata!
 * (src/banded/tridiagonalconjugation.jl:10-111, 114-215, 224-282): U upper triangular banded (the R of qr(A), the U of
 * cholesky(S), or a connection matrix), X tridiagonal (a Jacobi matrix); the Jacobi matrix of the modified weight is the
 * tridiagonal part of U*X*inv(U).  Every row is computed independently from rows k-1, k, k+1 of U with the reference's
 * operation order (bit-identical to its two sequential sweeps).
 *   U      batch x nbands x uld: U[(i*nbands + d)*uld + k] = U_i[k+1, k+1+d] (0-based k; nbands = 2 or 3), uld >= n+1
 *   X      3 x xld (x_shared != 0) or batch x 3 x xld: rows dl (X[k+2,k+1]), d (X[k+1,k+1]), du (X[k+1,k+2]), xld >= n+1
 *   Y      batch x 3 x n: rows dl, d, du of Y (what resizedata!(data, n) makes valid); UX_opt the same for Tridiagonal(U*X). */
int ila_triconj_bands_f64(int nbands, int64_t batch, int64_t n, const double *U, int64_t uld, const double *X, int64_t xld,
                          int x_shared, double *Y, double *UX_opt, int ngpus);
int ila_triconj_bands_f64_dev(int nbands, int64_t batch, int64_t n, const double *d_U, int64_t uld, const double *d_X,
                              int64_t xld, int x_shared, double *d_Y, double *d_UX_opt, void *stream);
/* U read straight from the device-resident factor records of ila_qr_banded_factor_f64_dev / ila_chol_banded_factor_f64_dev
 * (nch = 16-byte chunks per record: qr records ((l+u+1)+1 [+ l+1 reflector entries] rounded up to even)/2); the factorization
 * must have swept at least n+1 columns.  Rows of operators whose d_status is not 0 are NaN (d_status may be NULL). */
int ila_triconj_records_f64_dev(int nbands, int nch, int64_t batch, int64_t n, const int64_t *d_gbase, const void *d_work,
                                const int32_t *d_status, const double *d_X, int64_t xld, int x_shared, double *d_Y, void *stream);
/* cholesky(S_i).U -> Y_i in one call, U never leaves the device (TridiagonalConjugation(cholesky(S).U, X), e.g. S = W(X) a
 * polynomial weight modification or S = shift_i*I - X): operator description as in ila_chol_banded_solve_f64.  The
 * Cholesky blocking is two blocks of ceil((n+3)/2)+... columns (upstream the partition follows the caller's access
 * history — 100 columns at construction, tridiagonalconjugation.jl:244 — so bits at block edges are not defined upstream);
 * status 4 = not positive definite (rows NaN). */
int ila_chol_triconj_f64(int u, int64_t batch, const double *p, const double *q, const double *r, const double *shift,
                         int hp, const double *head, int64_t n, const double *X, int64_t xld, int x_shared, double *Y,
                         int32_t *status, int ngpus);
/* BidiagonalConjugation(U, X, V, uplo) (src/banded/bidiagonalconjugation.jl:19-60): bidiagonal part of inv(U)*X*V from the
 * tridiagonal parts of U and C = X*V: U, C batch x 3 x ld with rows sub-diagonal ([k+2,k+1]), diagonal, super-diagonal
 * ([k+1,k+2]), ld >= n+1.  upper != 0: dv[k] = B[k+1,k+1], ev[k] = B[k+1,k+2]; else ev[k] = B[k+2,k+1]; both batch x n. */
int ila_biconj_bands_f64(int upper, int64_t batch, int64_t n, const double *U, const double *C, int64_t ld, double *dv,
                         double *ev, int ngpus);

/* ---- K11: adaptive block-banded QR solve for ∞ operators with block sizes 1, 2, 3, ... ----------------------
 * Replaces qr(A) \ b for block-banded A: AdaptiveQRData(::AbstractBlockLayout) (src/infqr.jl:23-28), partialqr!(F, ::Block)
 * -> _blockbanded_qr! (:68-86), the Q'b driver with COLGROWTH = 300 and tolerance 1e-30 (:303-337), resizedata_chop!
 * (:221-232) and the block back-substitution (:358-366).
 * Operator i:  L_i = alpha[i] KronTrav(T1, I) + beta[i] KronTrav(I, T2) - lam[i] I,  T1, T2 tridiagonal, shared by the batch
 * as 3 x tld bands (rows dl: T[k+2,k+1], d: T[k+1,k+1], du: T[k+1,k+2]; tld >= maxblocks+1); block K holds K entries and
 * KronTrav(A,B)[Block(K,J)][k,j] = A[k,j] B[K-k+1,J-j+1] — the operators of test/test_infqr.jl:270-313 (Δ ⊗ I - 2I, I ⊗ Δ - 2I,
 * their sum, 2X + Y - 8I).
 *   b        batch x nb non-zero prefixes of the right-hand sides; tol, colgrowth: the reference's 1e-30 and 300
 *   maxblocks  capacity in blocks, 3..64 (the active panel lives in shared memory); status 3 when it is reached
 *   x        batch x xld, zero padded (xld >= maxblocks(maxblocks+1)/2); ncols[i] = length of the solution (the reference's
 *            datasize after the chop); nblocks[i] = block columns whose reflectors were applied to b
 *   qtb_opt  NULL or batch x xld: Q'b before the chop ((F.Q'*b)[1:6] of test_infqr.jl:276). */
int ila_qr_blockbanded_solve_f64(int64_t batch, const double *alpha, const double *beta, const double *lam, const double *T1,
                                 const double *T2, int64_t tld, int nb, const double *b, double tol, int colgrowth,
                                 int maxblocks, double *x, int64_t xld, int64_t *ncols, int32_t *nblocks, double *qtb_opt,
                                 int32_t *status, int ngpus);
/* device-resident form: the caller owns every buffer; d_work holds batch x ila_qr_blockbanded_work_doubles(maxblocks) doubles
 * (the R block rows), d_qtb and d_x batch x xld. */
int ila_qr_blockbanded_work_doubles(int maxblocks, int64_t *per_operator);
/* 0 (default): four reflectors per pass over the trailing panel; 1: one reflector per pass (validation: same results to rounding) */
int ila_qr_blockbanded_set_mode(int mode);
int ila_qr_blockbanded_solve_f64_dev(int64_t batch, const double *d_alpha, const double *d_beta, const double *d_lam,
                                     const double *d_T1, const double *d_T2, int64_t tld, int nb, const double *d_b, double tol,
                                     int colgrowth, int maxblocks, double *d_work, double *d_qtb, double *d_x, int64_t xld,
                                     int64_t *d_ncols, int32_t *d_nblocks, int32_t *d_status, void *stream);

/* ---- K12: adaptive QR solve of ∞ ALMOST-BANDED operators  A = Vcat(B, D)   ------------------------------------
 * r dense rows B (boundary conditions) on top of a banded operator D with bandwidths (lD, uD): the ultraspherical-spectral-
 * method shape of test/test_infqr.jl:193-267.  Replaces AdaptiveQRData(::AbstractAlmostBandedLayout) (src/infqr.jl:15-21),
 * partialqr! -> _almostbanded_qr! (:50-66), the Q'b driver (:236-274) and ldiv!(R, B) (:350-355); the rank-r fill above the
 * band is carried as r coefficients per row (see csrc/qr_almostbanded.cu).
 *   dg   batch x (lD+uD+1) x 3, diagonals of D in data-row order (+uD first): value(t) = g0 + g1 t + g2 t^2, t = 0-based ROW of D
 *   bg   batch x r x 4: B_m(j) = s^j (a + b j + c j^2), s = +1 / -1, j = 0-based column      (Ones: 1,1,0,0; (-1)^k: -1,1,0,0)
 *   shift batch or NULL, subtracted from D's main diagonal;  b: batch x nb;  tol, colgrowth: floatmin and 1000 upstream
 *   x    batch x ncap, zero padded; ncols[i] = the reference's datasize (nconv - 1 + colgrowth + lD + r); status 3 = ncap reached.
 * Supported: 1 <= r <= 2, lD + r <= 4, lD + uD + 1 <= 12. */
int ila_qr_almostbanded_solve_f64(int lD, int uD, int r, int64_t batch, const double *dg, const double *bg, const double *shift,
                                  int nb, const double *b, double tol, int colgrowth, int64_t ncap, double *x, int64_t *ncols,
                                  int64_t *nconv, int32_t *status, int ngpus);

#ifdef __cplusplus
}
#endif
#endif /* ILA_B200_H */
