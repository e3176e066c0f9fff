// ql_toeplitz.cu — K3: Toeplitz-tail QL for Hessenberg (u == 1) Toeplitz symbols, one thread per shift λ.
//
// Reference path replaced (InfiniteLinearAlgebra.jl v0.10.3):
//   ql(A - λ*I)  ->  _inf_ql              src/banded/infqltoeplitz.jl:182-191
//                ->  ql_hessenberg         src/banded/infqltoeplitz.jl:56-65
//                      tail_de             src/banded/infqltoeplitz.jl:7-27   (companion eigenproblem -> (d,e))
//                      ql_X!               src/banded/infqltoeplitz.jl:31-39  (2 x m Householder QL + sign fix)
//   .L[1,1]      ->  factors[1,1] = X[1,m-1]   src/infql.jl:96, infqltoeplitz.jl:63
//
// B200-first design (not a translation): the reference builds the (m-1)x(m-1) companion matrix and calls
// LAPACK geevx for ALL eigenpairs, then keeps one.  Here each thread keeps the m coefficients in registers,
// normalises the characteristic polynomial by the super-diagonal coefficient β (w = μ/β, so that
// "QL exists" <=> |w| >= 1), finds all roots simultaneously with a Gauss–Seidel Aberth–Ehrlich iteration
// (cubic convergence, one complex division per root per sweep), selects the branch, and recovers the
// eigenvector in closed form from the companion structure (V_1 = 1, V_{i+1} = g_i - w V_i).
// The simultaneous iteration runs as an FP32 "scout" on the FP32 pipe (consecutive shifts of a thread track the roots
// with one Durand–Kerner sweep); only the selected root is refined in FP64 (Newton + chord step) and the epilogue is
// FP64.  16 B in and 8-20 B out per shift against ~900 instructions: the kernel is issue-slot bound, not HBM bound
// (profiles/).  A root finder that exhausts its sweeps reports ILA_NOT_CONVERGED, never a silent value.
#include "ila_common.cuh"

namespace ila {

constexpr int kMaxDeg = 8;   // l + 1 <= 8
constexpr int kAberthMaxIt = 48;
// FP32 scout: a cold start converges in 6-8 sweeps (14 at most on the C3 grid); in rare regions of the λ plane the FP32 iteration
// from the circle start never settles although the FP64 one does in 8 sweeps — give up early and let the FP64 path take over
constexpr int kScoutMaxIt = 16;
static std::atomic<int> g_tz_spt{0};   // 0 = automatic; set by ila_ql_set_shifts_per_thread (1 = every shift starts cold)
static std::atomic<int> g_tz_max_it{0}, g_tz_scout_max_it{0};   // 0 = kAberthMaxIt / kScoutMaxIt (ila_ql_set_max_sweeps: test hook)
static std::atomic<int> g_tz_seeds{1};
void tz_set_seeds(int v) { g_tz_seeds.store(v ? 1 : 0); }
void tz_set_spt(int v) { g_tz_spt.store(v); }
void tz_set_max_sweeps(int fp64_sweeps, int scout_sweeps)
{
    g_tz_max_it.store(fp64_sweeps > 0 ? fp64_sweeps : 0);
    g_tz_scout_max_it.store(scout_sweeps > 0 ? scout_sweeps : 0);
}


// branch-free reciprocal for the root-finder's corrections (MUFU seed + two Newton steps, ~1 ulp): an iteration that
// converges to the root does not need correctly rounded quotients; the final factors use IEEE division.
__device__ __forceinline__ double fast_rcp(double b)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(r, -b, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(r, -b, 1.0);
    return fma(r, e, r);
}

__device__ __forceinline__ double dsign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0); }

// Horner: P(w) and P'(w) for the monic polynomial with coefficients cf[1..N]
// a*b + c with two FMAs per component (the operator form costs three instructions)
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c)
{
    return cplx{fma(a.re, b.re, fma(-a.im, b.im, c.re)), fma(a.re, b.im, fma(a.im, b.re, c.im))};
}

template <int N>
__device__ __forceinline__ void horner2(const cplx (&cf)[N + 1], cplx w, cplx &pv, cplx &dv)
{
    // k = 1 peeled: (pv, dv) = (1, 0) -> (w + cf_1, 1)
    dv = mk(1.0, 0.0);
    pv = w + cf[1];
#pragma unroll
    for (int k = 2; k <= N; k++) {
        dv = cfma(dv, w, pv);
        pv = cfma(pv, w, cf[k]);
    }
}

// P(w) alone
template <int N>
__device__ __forceinline__ cplx horner1(const cplx (&cf)[N + 1], cplx w)
{
    cplx pv = w + cf[1];
#pragma unroll
    for (int k = 2; k <= N; k++) pv = cfma(pv, w, cf[k]);
    return pv;
}

// ---- single-precision scout ----------------------------------------------------------------------------
// All N roots are located (and tracked from shift to shift) in FP32 on the FP32 pipe, which runs beside the FP64
// pipe: 1e-7 is ample to rank the moduli away from near-ties and to seed the FP64 Newton iteration that takes the
// selected root to full precision.  Near-ties and anything that does not converge fall back to the FP64 solve.
struct cplxf {
    float re, im;
};
__device__ __forceinline__ cplxf mkf(float re, float im = 0.f) { return cplxf{re, im}; }
__device__ __forceinline__ cplxf operator+(cplxf a, cplxf b) { return cplxf{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplxf operator-(cplxf a, cplxf b) { return cplxf{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplxf operator*(cplxf a, cplxf b) { return cplxf{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplxf operator*(float s, cplxf a) { return cplxf{s * a.re, s * a.im}; }
__device__ __forceinline__ cplxf conjf_(cplxf a) { return cplxf{a.re, -a.im}; }
__device__ __forceinline__ float abs2f(cplxf a) { return a.re * a.re + a.im * a.im; }
__device__ __forceinline__ cplxf cfmaf(cplxf a, cplxf b, cplxf c)
{
    return cplxf{fmaf(a.re, b.re, fmaf(-a.im, b.im, c.re)), fmaf(a.re, b.im, fmaf(a.im, b.re, c.im))};
}
__device__ __forceinline__ float rcpf_approx(float x)   // one MUFU; callers keep x inside [1e-30, 1e30]
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int N>
struct TzParams {
    cplx a[N + 1];     // reversed data column: a[0] = a_{-l}, ..., a[N-1] = a_0, a[N] = a_{+1} = β
    cplx ginv;         // 1/β
    cplx cf0[N + 1];   // (-1)^k a_{m-k}/β, k = 0..N, before the shift (only k = 1 moves with λ)
    cplxf cf0f[N + 1];
    cplx unit[N];      // initial directions exp(i(θ0 + 2πj/N))
    float cmax2f;      // max_{k >= 2} |cf0_k|^2 (the shift-independent coefficients)
    double beta2;      // |β|^2
    int branch_rank;
    int max_it, scout_max_it;   // sweep limits of the FP64 / FP32 simultaneous iterations
    int use_seeds;              // closed-form cold-start seeds (N = 2, 4); 0 = circle start everywhere (validation)
    int out_mode;      // complex-eltype L11-only output: 0 = ComplexF64 + int32 status, 1 = Float64 (NaN payload = status) + optional uint8 status
};

// lim2: square of 8x a Cauchy-type bound on the root moduli (1 + max|cf_k|); an iterate beyond it is pulled back
template <int N>
__device__ __forceinline__ int aberth_sweep_f32(const cplxf (&cf)[N + 1], cplxf (&w)[N], float lim2)
{
    bool notconv = false, big = false;
#pragma unroll
    for (int j = 0; j < N; j++) {
        cplxf dv = mkf(1.f), pv = w[j] + cf[1];
#pragma unroll
        for (int k = 2; k <= N; k++) {
            dv = cfmaf(dv, w[j], pv);
            pv = cfmaf(pv, w[j], cf[k]);
        }
        cplxf num = mkf(0.f), den = mkf(1.f);
        bool first = true;
#pragma unroll
        for (int k = 0; k < N; k++) {
            if (k == j) continue;
            const cplxf d = w[j] - w[k];
            if (first) { num = mkf(1.f); den = d; first = false; }
            else { num = cfmaf(num, d, den); den = den * d; }
        }
        const cplxf top = pv * den;
        const cplxf bot = cfmaf(dv, den, mkf(0.f) - pv * num);
        const float b2 = abs2f(bot);
        cplxf corr = mkf(0.f);
        if (b2 > 1e-30f && b2 < 1e30f) corr = rcpf_approx(b2) * (top * conjf_(bot));
        else big = true;
        const cplxf wold = w[j];
        w[j] = w[j] - corr;
        float w2 = abs2f(w[j]);
        const float c2 = abs2f(corr);
        if (!(w2 < lim2)) {
            // excursion: a step that lands far outside the Cauchy bound of the roots (next to a pole of the Aberth map the
            // FP32 correction overflows and the iterate would stay Inf/NaN for good).  Keep the direction, cut the length.
            const float s = (c2 > 0.f && c2 < 1e37f) ? sqrtf(0.0625f * lim2 / c2) : 0.f;
            w[j] = (s > 0.f) ? wold - s * corr : mkf(0.7f * wold.re - 0.7f * wold.im, 0.7f * wold.re + 0.7f * wold.im);
            w2 = abs2f(w[j]);
            big = true;
        }
        notconv = notconv || !(c2 <= 2.5e-5f * w2);   // |corr| > 5e-3 |w| (or NaN): the cubic step then lands at ~1e-6 or better
        big = big || !(c2 <= 1e-4f * w2) || !(w2 < 1e30f);
    }
    return big ? 2 : (notconv ? 1 : 0);
}

// One Gauss–Seidel Durand–Kerner (Weierstrass) sweep in FP32: corr_j = P(w_j) / prod_{k != j}(w_j - w_k) — no P',
// less than half the work of an Aberth sweep, quadratic.  Used to TRACK the roots from a predicted start (error
// O(Δλ²)); same return convention as the Aberth sweep.
template <int N>
__device__ __forceinline__ int dk_sweep_f32(const cplxf (&cf)[N + 1], cplxf (&w)[N])
{
    bool notconv = false, big = false;
#pragma unroll
    for (int j = 0; j < N; j++) {
        cplxf pv = w[j] + cf[1];
#pragma unroll
        for (int k = 2; k <= N; k++) pv = cfmaf(pv, w[j], cf[k]);
        cplxf den = mkf(1.f);
        bool first = true;
#pragma unroll
        for (int k = 0; k < N; k++) {
            if (k == j) continue;
            const cplxf d = w[j] - w[k];
            den = first ? d : den * d;
            first = false;
        }
        const float b2 = abs2f(den);
        cplxf corr = mkf(0.f);
        if (b2 > 1e-30f && b2 < 1e30f) corr = rcpf_approx(b2) * (pv * conjf_(den));
        else big = true;
        w[j] = w[j] - corr;
        const float w2 = abs2f(w[j]), c2 = abs2f(corr);
        notconv = notconv || !(c2 <= 1e-6f * w2);    // |corr| > 1e-3 |w| (or NaN): the quadratic step then lands at ~1e-6
        big = big || !(c2 <= 1e-4f * w2) || !(w2 < 1e30f);
    }
    return big ? 2 : (notconv ? 1 : 0);
}

// ---- closed-form seeds for the cold start (FP32) --------------------------------------------------------------------
// A run's first shift used to start on a circle and needed 6-14 Aberth sweeps (warp maximum ~8: ~2400 instructions, a third
// of the kernel's issue slots on the 1024² grid and ALL of them on the sub-wave slabs of a strong-scaled grid).  For the
// degrees with a closed form the roots are written down instead — Ferrari's resolvent for the quartic (the Bull-head and
// every other l = 3 symbol), the quadratic formula for l = 1 — in FP32, to ~1e-3..1e-6 depending on cancellation; they
// only SEED the same Durand–Kerner / Aberth polish the warm starts use, and anything degenerate (multiple roots, vanishing
// resolvent root) falls back to the circle start, so no result depends on them.
__device__ __forceinline__ cplxf csqrtf_(cplxf z)
{
    const float r = sqrtf(abs2f(z));
    const float t = sqrtf(0.5f * (r + fabsf(z.re)));
    if (!(t > 0.f)) return mkf(0.f);
    const float h = 0.5f * z.im / t;
    return (z.re >= 0.f) ? mkf(t, h) : mkf(fabsf(h), copysignf(t, z.im));
}
__device__ __forceinline__ cplxf crcpf_(cplxf z)
{
    const float d = abs2f(z);
    return (d > 1e-36f && d < 1e36f) ? (1.0f / d) * conjf_(z) : mkf(0.f);
}
__device__ __forceinline__ cplxf ccbrtf_(cplxf z)
{
    const float r2 = abs2f(z);
    if (!(r2 > 0.f)) return mkf(0.f);
    const float r = cbrtf(sqrtf(r2));
    float sn, cs;
    __sincosf(atan2f(z.im, z.re) * (1.0f / 3.0f), &sn, &cs);
    return mkf(r * cs, r * sn);
}
// quadratic y^2 + b y + c = 0: the larger root first (no cancellation), the other one from the product
__device__ __forceinline__ void quadf_(cplxf b, cplxf c, cplxf &y1, cplxf &y2)
{
    cplxf sq = csqrtf_(b * b - 4.0f * c);
    if (b.re * sq.re + b.im * sq.im < 0.f) sq = mkf(0.f) - sq;      // Re(conj(b) sq) >= 0: -(b + sq) does not cancel
    y1 = -0.5f * (b + sq);
    y2 = c * crcpf_(y1);
}
template <int N>
__device__ __forceinline__ bool closed_form_seeds_f32(const cplxf (&cf)[N + 1], cplxf (&w)[N])
{
    if constexpr (N == 2) {
        quadf_(cf[1], cf[2], w[0], w[1]);
        return abs2f(w[0]) > 0.f;
    } else if constexpr (N == 4) {
        // depressed quartic y^4 + p y^2 + q y + r, w = y - a
        const cplxf a = 0.25f * cf[1], a2 = a * a;
        const cplxf p = cf[2] - 6.0f * a2;
        const cplxf q = cf[3] - 2.0f * (a * cf[2]) + 8.0f * (a2 * a);
        const cplxf r = cf[4] - a * cf[3] + a2 * cf[2] - 3.0f * (a2 * a2);
        // resolvent cubic m^3 + p m^2 + (p^2/4 - r) m - q^2/8 = 0, m = t - p/3:  t^3 + P t + Q = 0
        const cplxf p2 = p * p;
        const cplxf P = mkf(0.f) - ((1.0f / 12.0f) * p2 + r);
        const cplxf Q = (-1.0f / 108.0f) * (p2 * p) + (1.0f / 3.0f) * (p * r) - 0.125f * (q * q);
        const cplxf sD = csqrtf_(0.25f * (Q * Q) + (1.0f / 27.0f) * (P * P * P));
        const cplxf hQ = -0.5f * Q;
        cplxf u3 = hQ + sD;
        const cplxf u3b = hQ - sD;
        if (abs2f(u3b) > abs2f(u3)) u3 = u3b;                        // Cardano: the cube without cancellation
        const cplxf u = ccbrtf_(u3);
        const cplxf v = (-1.0f / 3.0f) * (P * crcpf_(u));
        // the three resolvent roots; the one of largest modulus keeps s = sqrt(2m) away from 0 (q ~ 0: biquadratic case)
        const cplxf om = mkf(-0.5f, 0.8660254f), omc = mkf(-0.5f, -0.8660254f);
        const cplxf p3 = (1.0f / 3.0f) * p;
        cplxf m = u + v - p3;
        const cplxf m1 = om * u + omc * v - p3, m2 = omc * u + om * v - p3;
        if (abs2f(m1) > abs2f(m)) m = m1;
        if (abs2f(m2) > abs2f(m)) m = m2;
        const cplxf sm = csqrtf_(2.0f * m);
        const float s2 = abs2f(sm);
        if (!(s2 > 1e-12f * (1.0f + abs2f(p)))) return false;        // quadruple-root neighbourhood: circle start instead
        const cplxf g = 0.5f * (q * crcpf_(sm));
        const cplxf base = 0.5f * p + m;
        cplxf y0, y1, y2, y3;
        quadf_(mkf(0.f) - sm, base + g, y0, y1);                     // y^2 - s y + (p/2 + m + q/(2s)) = 0
        quadf_(sm, base - g, y2, y3);                                // y^2 + s y + (p/2 + m - q/(2s)) = 0
        w[0] = y0 - a; w[1] = y1 - a; w[2] = y2 - a; w[3] = y3 - a;
        float chk = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) chk += abs2f(w[j]);
        return chk < 1e30f;                                          // false on NaN / overflow
    } else {
        return false;
    }
}

// One Gauss–Seidel Aberth–Ehrlich sweep over all N roots; returns 0 when every correction is below 1e-9 |w_j|
// (converged), 1 when all are below 1e-2 |w_j| (tracking), 2 otherwise.
template <int N>
__device__ __forceinline__ int aberth_sweep(const cplx (&cf)[N + 1], cplx (&w)[N])
{
    bool notconv = false, big = false;
#pragma unroll
    for (int j = 0; j < N; j++) {
        cplx pv, dv;
        horner2<N>(cf, w[j], pv, dv);
        // sum_k 1/(w_j - w_k) kept as num/den to spend a single division per root
        cplx num = mk(0.0, 0.0), den = mk(1.0, 0.0);
#pragma unroll
        for (int k = 0; k < N; k++) {
            if (k == j) continue;
            const cplx d = w[j] - w[k];
            num = cfma(num, d, den);
            den = den * d;
        }
        const cplx top = pv * den;
        const cplx bot = dv * den - pv * num;
        const double b2 = abs2(bot);
        cplx corr = mk(0.0, 0.0);
        if (b2 > 1e-280 && b2 < 1e280) corr = fast_rcp(b2) * (top * conj(bot));
        else if (b2 > 0.0 && b2 < 1.7e308) corr = (1.0 / b2) * (top * conj(bot));
        else big = notconv = true;      // overflow / NaN in the products: a zero correction here must not read as convergence
        w[j] = w[j] - corr;
        const double w2 = abs2(w[j]);
        const double c2 = abs2(corr);
        notconv = notconv || !(c2 <= 1e-18 * w2) || !(w2 < 1e300);     // NaN/Inf-safe: a non-finite iterate is never "converged"
        big = big || !(c2 <= 1e-4 * w2);
    }
    return big ? 2 : (notconv ? 1 : 0);
}

// Branch selection on |μ|^2 = |β|^2 |w|^2 (findmax / k-th largest; ties: Julia's (re, im) eigenvalue order).
// gap_out: relative distance of the selected |w|^2 to its nearest neighbour in the ranking.
template <int N>
__device__ __forceinline__ int select_root(const cplx (&w)[N], int branch_rank, double &gap_out)
{
    double n2[N];
#pragma unroll
    for (int j = 0; j < N; j++) n2[j] = abs2(w[j]);
    int want = branch_rank - 1;   // number of roots ranked above the selected one
    if (want < 0) want = 0;
    if (want > N - 1) want = N - 1;
    int jsel = 0;
#pragma unroll
    for (int j = 0; j < N; j++) {
        int larger = 0;
#pragma unroll
        for (int k = 0; k < N; k++) {
            if (k == j) continue;
            const cplx mj = w[j], mk_ = w[k];
            bool gt = n2[k] > n2[j];
            if (n2[k] == n2[j]) {
                // tie: the reference keeps the first maximum in (real, imag)-sorted order
                gt = (mk_.re < mj.re) || (mk_.re == mj.re && mk_.im < mj.im) || (mk_.re == mj.re && mk_.im == mj.im && k < j);
            }
            larger += gt ? 1 : 0;
        }
        if (larger == want) jsel = j;
    }
    double nsel = n2[0];
#pragma unroll
    for (int j = 1; j < N; j++)
        if (j == jsel) nsel = n2[j];
    double gap = 1e300;
#pragma unroll
    for (int j = 0; j < N; j++)
        if (j != jsel) gap = fmin(gap, fabs(n2[j] - nsel));
    gap_out = (nsel > 0.0) ? gap / nsel : 0.0;
    return jsel;
}

// FP32 ranking of the scout's roots (used only to pick the root Newton refines and to detect near-ties; whenever the
// relative gap is below 1e-5 the FP64 path with the reference's exact tie rule decides instead).
template <int N>
__device__ __forceinline__ int select_root_f32(const cplxf (&w)[N], int branch_rank, float &gap_out)
{
    float n2[N];
#pragma unroll
    for (int j = 0; j < N; j++) n2[j] = abs2f(w[j]);
    int jsel = 0;
    if (branch_rank <= 1) {
        float best = n2[0];
#pragma unroll
        for (int j = 1; j < N; j++)
            if (n2[j] > best) { best = n2[j]; jsel = j; }
    } else {
        int want = branch_rank - 1;
        if (want > N - 1) want = N - 1;
#pragma unroll
        for (int j = 0; j < N; j++) {
            int larger = 0;
#pragma unroll
            for (int k = 0; k < N; k++) {
                if (k == j) continue;
                larger += (n2[k] > n2[j] || (n2[k] == n2[j] && k < j)) ? 1 : 0;
            }
            if (larger == want) jsel = j;
        }
    }
    float nsel = n2[0];
#pragma unroll
    for (int j = 1; j < N; j++)
        if (j == jsel) nsel = n2[j];
    float gap = 3e38f;
#pragma unroll
    for (int j = 0; j < N; j++)
        if (j != jsel) gap = fminf(gap, fabsf(n2[j] - nsel));
    gap_out = (nsel > 1e-30f && nsel < 1e30f) ? gap * rcpf_approx(nsel) : 0.f;
    return jsel;
}

// sqrt for the epilogue: MUFU rsqrt seed + coupled Newton (Goldschmidt) in FP64, branch-free inside a wide exponent
// box, IEEE sqrt outside.  Result within 1 ulp.
__device__ __forceinline__ double fast_sqrt(double x)
{
    if (!(x > 1e-280 && x < 1e280)) return sqrt(x);
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double g = x * r, h = 0.5 * r;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
}

// Optional grid mode (the λ loop of examples/toeplitz.jl:29-33, `ℓ11.(Ref(A), x' .+ y.*im)`): shifts are formed on device
// as complex(gx[j], gy[k]) for idx = k*nx + j, and the result is the example's ℓ11 value: abs(L[1,1]), or -1.0 where
// the QL factorization does not exist — 8 B out and no λ in per shift.
struct TzGrid {
    const double *gx, *gy;
    int nx;
    double *absout;
    int out_mode;   // array form, complex eltype, L11 only: 1 = Float64 L11 (+ optional uint8 status), see TzParams
};

constexpr int kSolveMaxB = 16;
struct TzSolveParams {
    int nb;
    cplx b[kSolveMaxB];
};

struct TzSolveArgs {   // fused form: the K3 kernel runs the solve from its registers instead of writing the tail record
    TzSolveParams S;
    int64_t nout;
    double *x;          // column-major n x nout, nullptr = not fused
    int mode = 0;       // 0: x = L \ (Q'b)   1: x = Q'b   2: x = Q b  (LowerHessenbergQ apply with its early exit)
    int64_t *nstop = nullptr;   // mode 2: 1-based n at which the early exit fired (nout if it did not within nout entries)
};

// Q'b and the forward substitution for one shift (see the K7 header below); l11 = X[1,m-1], v = X[1,m],
// X2[0..N] = X[2,1..m] after ql_X!.
template <int N, bool REAL_PATH>
__device__ __forceinline__ void tz_solve_rows(const TzSolveParams &S, cplx l11, cplx v, const cplx (&X2)[N + 1], cplx tau1,
                                              cplx tau2, int64_t nout, int64_t n, int64_t idx, double *__restrict__ x,
                                              int SVmode = 0, int64_t *__restrict__ nstop = nullptr)
{
    auto put = [&](int64_t j, cplx val) {
        if (REAL_PATH) x[j * n + idx] = val.re;
        else reinterpret_cast<double2 *>(x)[j * n + idx] = make_double2(val.re, val.im);
    };
    // q = H_2 H_1,  H_1 = diag(1 - tau1, 1),  H_2 = I - tau2 [v;1][v;1]'
    const cplx one = mk(1.0, 0.0);
    const cplx s1 = one - tau1;
    const cplx q11 = (one - abs2(v) * tau2) * s1, q21 = (-(conj(v) * tau2)) * s1;
    const cplx q12 = -(v * tau2), q22 = one - tau2;
    // Q'b: b[n:n+1] <- q' b[n:n+1], n = nb, ..., 1
    cplx tb[kSolveMaxB + 1];
    for (int k = 0; k <= kSolveMaxB; k++) tb[k] = (k < S.nb) ? S.b[k] : mk(0.0, 0.0);
    for (int k = S.nb; k >= 1; k--) {
        const cplx u0 = tb[k - 1], u1 = tb[k];
        tb[k - 1] = conj(q11) * u0 + conj(q21) * u1;
        tb[k] = conj(q12) * u0 + conj(q22) * u1;
    }
    if (SVmode == 2) {
        // Q b, Q = LowerHessenbergQ(Fill(q, ∞)) (src/banded/hessenbergq.jl:111-123): for n = 1, 2, ...: x[n:n+1] <- q x[n:n+1];
        // stops once n > nz and norm(t) <= 10 floatmin — what is left of x is then exactly what the reference leaves
        cplx carry = S.b[0];
        int64_t stop = 0;                       // 0: the exit did not fire within nout entries (support goes on)
        for (int64_t j = 0; j < nout; j++) {
            const cplx nxt = (j + 1 < S.nb) ? S.b[j + 1] : mk(0.0, 0.0);
            const cplx t1 = q11 * carry + q12 * nxt, t2 = q21 * carry + q22 * nxt;
            put(j, t1);
            carry = t2;
            if (j + 1 > S.nb && hypot(hypot(t1.re, t1.im), hypot(t2.re, t2.im)) <= 10.0 * 2.2250738585072014e-308) {
                stop = j + 1;
                if (j + 1 < nout) put(j + 1, t2);
                for (int64_t k = j + 2; k < nout; k++) put(k, mk(0.0, 0.0));
                break;
            }
        }
        if (nstop) nstop[idx] = stop;
        return;
    }
    if (SVmode == 1) {
        // Q'b alone (UpperHessenbergQ(adjoint.(q)) apply, src/banded/hessenbergq.jl:125-135): support nb + 1
        for (int64_t j = 0; j < nout; j++) put(j, (j <= S.nb && j <= kSolveMaxB) ? tb[j] : mk(0.0, 0.0));
        return;
    }
    // L \ (Q'b): sub[i-1] = L[j+i, j] = X[2, m-i], i = 1..N; diagonal X[1,m-1] (first column) / X[2,m] (others)
    cplx sub[N];
#pragma unroll
    for (int i = 1; i <= N; i++) sub[i - 1] = X2[N - i];
    auto rcp = [&](cplx d) { const double r2 = 1.0 / abs2(d); return r2 * conj(d); };
    const cplx r_first = rcp(l11), r_tail = rcp(X2[N]);
    cplx r[N + 1];   // r[0] = current row, r[k] = pending value of row j + k
#pragma unroll
    for (int k = 0; k <= N; k++) r[k] = tb[k];
    for (int64_t j = 0; j < nout; j++) {
        const cplx xj = r[0] * (j == 0 ? r_first : r_tail);
        put(j, xj);
        const int64_t nx = j + N + 1;   // row entering the window
        const cplx tin = (nx <= S.nb) ? tb[nx] : mk(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < N; k++) r[k] = r[k + 1] - sub[k] * xj;
        r[N] = tin;
    }
}

// Each thread owns SPT consecutive shifts.  The first one starts cold (roots on a circle); the following ones start
// from the previous shift's roots, which on a grid are O(Δλ) away: one Aberth sweep re-establishes every root to ~1e-5
// (enough to rank the moduli unless they nearly tie), then only the selected root is taken to full precision by Newton.
// Any doubt (large corrections, near-tie in the ranking, slow Newton) falls back to full simultaneous convergence, so
// the result never depends on the warm start.
template <int N, bool REAL_PATH, bool WANT_TAIL>
__global__ void __launch_bounds__(128, (N <= 4 && !WANT_TAIL) ? 7 : 1)
ql_toeplitz_l11_kernel(TzParams<N> P, const double *__restrict__ lambda, int64_t n, int spt, double *__restrict__ L11,
                       double *__restrict__ tail, int32_t *__restrict__ status, TzGrid G, TzSolveArgs SV)
{
    const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * spt;
    if (first >= n) return;
    constexpr int M = N + 1;   // number of symbol coefficients (reference's m)
    const int64_t last = (first + spt < n) ? first + spt : n;

    cplxf wf[N];                 // FP32 scout roots, carried from shift to shift
    cplxf dwf[N];                // their last movement: linear predictor for the next shift
#pragma unroll
    for (int j = 0; j < N; j++) dwf[j] = mkf(0.f);
    bool have_roots = false;
  for (int64_t idx = first; idx < last; idx++) {
    cplx lam;
    if (G.gx) {
        const int64_t k = idx / G.nx;
        lam = mk(G.gx[idx - k * G.nx], G.gy[k]);
    } else if (REAL_PATH) {
        lam = mk(lambda[idx], 0.0);
    } else {
        const double2 t = reinterpret_cast<const double2 *>(lambda)[idx];
        lam = mk(t.x, t.y);
    }
    // A - λI: only the diagonal coefficient moves (InfToeplitz - UniformScaling)
    const cplx adiag = P.a[N - 1] - lam;

    // monic characteristic polynomial of C/β:  P(w) = sum_k cf_k w^{N-k},  cf_k = (-1)^k a_{m-k}/β  (SURVEY A.2 with
    // w = μ/β); only cf_1 = -(a_0 - λ)/β depends on the shift
    cplx cf[N + 1];
    cplxf cff[N + 1];
#pragma unroll
    for (int k = 0; k <= N; k++) {
        cf[k] = P.cf0[k];
        cff[k] = P.cf0f[k];
    }
    cf[1] = cf[1] + lam * P.ginv;
    cff[1] = mkf((float)cf[1].re, (float)cf[1].im);
    const float lim2 = 64.f * (1.f + P.cmax2f + abs2f(cff[1]));     // (8 (1 + max|cf_k|))^2, up to a factor: excursion guard of the scout
    cplx g[N + 1];   // g_k = (-1)^k cf_k = a_{m-k}/β
#pragma unroll
    for (int k = 0; k <= N; k++) g[k] = (k & 1) ? -cf[k] : cf[k];

    // ---- FP32 scout: cold start on a circle around the centroid with the geometric-mean radius |P(c)|^(1/N),
    //      warm start from the previous shift's roots
    bool seeded = false;
    if (!have_roots && P.use_seeds) {
        seeded = closed_form_seeds_f32<N>(cff, wf);
    }
    if (!have_roots && !seeded) {
        const cplxf c = (1.0f / N) * mkf((float)g[1].re, (float)g[1].im);
        cplxf pc = mkf(1.f);
#pragma unroll
        for (int k = 1; k <= N; k++) pc = pc * c + cff[k];
        float rad = powf(abs2f(pc), 0.5f / N);
        if (!(rad > 1e-4f) || !(rad < 1e15f)) {
            float mx = 0.f;
#pragma unroll
            for (int k = 1; k <= N; k++) mx = fmaxf(mx, abs2f(cff[k]));
            rad = 1.0f + sqrtf(mx);
        }
#pragma unroll
        for (int j = 0; j < N; j++) wf[j] = c + rad * mkf((float)P.unit[j].re, (float)P.unit[j].im);
    }
    int scout = 2;
    if (have_roots) {
        // warm: linear prediction from the last two shifts, then ONE cheap Durand–Kerner sweep; accepted when every
        // correction is below 1e-3 |w| (the roots are then good to ~1e-6), otherwise the Aberth loop below takes over
        cplxf wold[N];
#pragma unroll
        for (int j = 0; j < N; j++) { wold[j] = wf[j]; wf[j] = wf[j] + dwf[j]; }
        scout = dk_sweep_f32<N>(cff, wf);
#pragma unroll
        for (int j = 0; j < N; j++) dwf[j] = wf[j] - wold[j];
    } else if (seeded) {
        // cold with closed-form seeds: the same Durand–Kerner polish (a second one when the first still moved the roots)
        scout = dk_sweep_f32<N>(cff, wf);
        if (scout == 1) scout = dk_sweep_f32<N>(cff, wf);
    }
#ifdef ILA_K3_DEBUG
    int dbg_f32 = 0, dbg_f64 = 0, dbg_flags = have_roots ? 1 : 0;
#define K3_DBG(x) x
#else
#define K3_DBG(x)
#endif
    if (scout != 0) {
        for (int it = 0; it < P.scout_max_it; it++) {
            scout = aberth_sweep_f32<N>(cff, wf, lim2);
            K3_DBG(dbg_f32++;)
            if (scout == 0) break;
        }
    }
    float gapf;
    int jsel = select_root_f32<N>(wf, P.branch_rank, gapf);
    if (scout == 0 && gapf < 1e-3f) {
        // near-tie in the ranking: one more (cubic) sweep takes the scout to FP32 precision before deciding
        scout = aberth_sweep_f32<N>(cff, wf, lim2);
        if (scout != 2) scout = 0;
        jsel = select_root_f32<N>(wf, P.branch_rank, gapf);
    }
    cplxf wsf = wf[0];
#pragma unroll
    for (int j = 1; j < N; j++)
        if (j == jsel) wsf = wf[j];
    cplx ws = mk((double)wsf.re, (double)wsf.im);
    double gap = (double)gapf;
    bool full = (scout != 0) || (gap < 1e-5);
    bool newton_done = false;
    bool lost = false;           // the FP64 simultaneous iteration used up its sweeps without settling
    if (!full) {
        // ---- FP64 Newton on the selected root only (~1e-6 -> 1e-12 -> rounding level); accepted when the last
        //      correction applied is below 1e-10 |w| (the iterate is then exact to rounding)
        double c2, w2;
        {
            // step 1: Newton
            cplx pv, dv;
            horner2<N>(cf, ws, pv, dv);
            const double d2 = abs2(dv);
            cplx dinv = mk(0.0, 0.0);   // 1/P'(w)
            if (d2 > 1e-280 && d2 < 1e280) dinv = fast_rcp(d2) * conj(dv);
            ws = ws - pv * dinv;
            // step 2: chord step with the same 1/P' (error ~ e0 e1, already below rounding when e0 ~ 1e-6)
            const cplx corr = horner1<N>(cf, ws) * dinv;
            ws = ws - corr;
            c2 = abs2(corr);
            w2 = abs2(ws);
        }
        if (c2 > 1e-24 * w2) {
            // slow case (close roots): one more full Newton step, judged by its own correction
            cplx pv, dv;
            horner2<N>(cf, ws, pv, dv);
            const double d2 = abs2(dv);
            cplx corr = mk(0.0, 0.0);
            if (d2 > 1e-280 && d2 < 1e280) corr = fast_rcp(d2) * (pv * conj(dv));
            ws = ws - corr;
            c2 = abs2(corr);
            w2 = abs2(ws);
        }
        newton_done = (c2 <= 1e-20 * w2);
        full = !newton_done;
    }
    if (full) {
        // ---- FP64 simultaneous solve from the scout's roots (near-tie in the ranking, or the scout/Newton had doubts)
        cplx w[N];
#pragma unroll
        for (int j = 0; j < N; j++) w[j] = mk((double)wf[j].re, (double)wf[j].im);
        if (scout == 2) {
            // scout lost: restart cold in FP64
            const cplx c = (1.0 / N) * g[1];
            cplx pc, dc;
            horner2<N>(cf, c, pc, dc);
            double rad = pow(abs2(pc), 0.5 / N);
            if (!(rad > 1e-8)) {
                double mx = 0.0;
#pragma unroll
                for (int k = 1; k <= N; k++) mx = fmax(mx, abs2(g[k]));
                rad = 1.0 + sqrt(mx);
            }
            // NOT the scout's circle again: where the FP32 iteration from that start wanders, the FP64 one is slow too (16
            // sweeps instead of 8); a wider circle turned by 0.35 rad starts outside every root and converges steadily
            const cplx rot = mk(0.9393727128473789, 0.34289780745545134);
#pragma unroll
            for (int j = 0; j < N; j++) w[j] = c + (1.6 * rad) * (P.unit[j] * rot);
        }
        K3_DBG(dbg_flags |= 2 | (scout == 2 ? 4 : 0);)
        int rc64 = 2;
        for (int it = 0; it < P.max_it; it++) {
            K3_DBG(dbg_f64++;)
            rc64 = aberth_sweep<N>(cf, w);
            if (rc64 == 0) break;
        }
        lost = (rc64 != 0);
        jsel = select_root<N>(w, P.branch_rank, gap);
        ws = w[0];
#pragma unroll
        for (int j = 1; j < N; j++)
            if (j == jsel) ws = w[j];
#pragma unroll
        for (int j = 0; j < N; j++) wf[j] = mkf((float)w[j].re, (float)w[j].im);
    }
    have_roots = true;

    int32_t st = ILA_OK;
    if (REAL_PATH) {
        // isreal(λ[j]) (infqltoeplitz.jl:12): a complex-conjugate pair cannot be a real root
        if (!(fabs(ws.im) <= 1e-9 * fabs(ws.re))) st = ILA_NOT_REAL;
        ws.im = 0.0;
    }
    // one Newton polish of the selected root (removes any dependence on the other iterates)
    if (!newton_done) {
        cplx pv, dv;
        horner2<N>(cf, ws, pv, dv);
        const double d2 = abs2(dv);
        if (d2 > 0.0) {
            cplx corr = (1.0 / d2) * (pv * conj(dv));
            if (REAL_PATH) corr.im = 0.0;
            if (abs2(corr) < 1e-6 * abs2(ws)) ws = ws - corr;
        }
    }
    if (full && st == ILA_OK) {
        // Whatever the FP64 simultaneous iteration returned (a handful of shifts per grid get here) is VERIFIED, whether or not
        // it reported convergence: the root is accepted only if it is one to rounding — backward error |P(w)| <= 1e-13
        // sum_k |cf_k||w|^(N-k), the criterion that also holds at multiple roots, where the corrections stay large — else
        // ILA_NOT_CONVERGED (infqltoeplitz.jl has no counterpart: LAPACK's QR iteration reports info > 0 and `eigen` throws).
        const cplx pv = horner1<N>(cf, ws);
        const double aw = sqrt(abs2(ws));
        double bnd = 1.0;
#pragma unroll
        for (int k = 1; k <= N; k++) bnd = fma(bnd, aw, sqrt(abs2(cf[k])));
        if (!(abs2(pv) <= 1e-26 * bnd * bnd)) st = ILA_NOT_CONVERGED;
        if (st == ILA_NOT_CONVERGED && P.branch_rank == 1) {
            // Rescue for shifts far outside the symbol's range (|λ - a_0| >> the other coefficients: one root ~ -cf_1 of modulus
            // up to 1e150, the others tiny — the simultaneous iteration cannot cross that many decades in its sweeps and P(w)
            // overflows).  The dominant root is the SMALLEST root of the reversed polynomial Q(u) = u^N P(1/u) = sum_k cf_k u^k:
            // Newton from u = -1/cf_1, nothing overflows; Pellet's test at r = |w|/2 (|cf_1| r^(N-1) > r^N + sum_{k>=2}|cf_k| r^(N-k):
            // exactly N-1 roots inside r) proves it is the one `findmax` selects.
            const double a1 = abs2(cf[1]);
            if (a1 > 0.0 && a1 < 1e300) {
                cplx u = -((1.0 / a1) * conj(cf[1]));
                bool okn = false;
                for (int it = 0; it < 12; it++) {
                    cplx q = cf[N], dq = mk(0.0, 0.0);
#pragma unroll
                    for (int k = N - 1; k >= 0; k--) { dq = cfma(dq, u, q); q = cfma(q, u, k == 0 ? mk(1.0, 0.0) : cf[k]); }
                    const double d2 = abs2(dq);
                    if (!(d2 > 0.0) || !(d2 < 1e300)) break;
                    const cplx du = (1.0 / d2) * (q * conj(dq));
                    u = u - du;
                    if (abs2(du) <= 1e-24 * abs2(u)) { okn = true; break; }   // |du| <= 1e-12 |u|: quadratic, the step applied lands at rounding level
                }
                const double u2 = abs2(u);
                if (okn && u2 > 0.0) {
                    const cplx wr = (1.0 / u2) * conj(u);
                    const double rinv = 2.0 * sqrt(u2);            // 1/r, r = |w|/2
                    double rest = 1.0, pw = rinv;
#pragma unroll
                    for (int k = 2; k <= N; k++) { pw *= rinv; rest = fma(sqrt(abs2(cf[k])), pw, rest); }
                    if (sqrt(a1) * rinv > rest) { ws = wr; st = ILA_OK; }
                }
            }
        }
    }
    if (st == ILA_OK && !(abs2(ws) < 1e300)) st = ILA_NOT_CONVERGED;       // non-finite or overflowing root: flagged, never NO_QL

    // μ = β w ; n2 = |μ|^2 ; existence test n2 >= |β|^2 (infqltoeplitz.jl:13,23)
    const cplx beta = P.a[N];
    const cplx mu = beta * ws;
    const double n2 = abs2(mu);
    if (st == ILA_OK && !(n2 >= P.beta2)) st = ILA_NO_QL;

    if (!WANT_TAIL && !REAL_PATH) {
        // L[1,1] only: the two reflectors of ql_X! (infqltoeplitz.jl:28-37) touch X[1,m-1] through column m-1 alone, so
        // the general code below collapses to a closed form.  With V_1 = 1: X[2,m] = -c_abs (real), ν = -‖(c_abs, β)‖,
        // v = β/(ξ1+ν), τ = (ξ1+ν)/ν (real), y = X[1,m-1] - v τ (X[2,m-1] + conj(v) X[1,m-1]); the k = 1 reflector and
        // the sign normalisation then give L[1,1] = sign(-c_abs) |y|.
        const double c_abs = fast_sqrt(n2 - P.beta2);
        const double normu = fast_sqrt(c_abs * c_abs + P.beta2);
        const double xi = -(c_abs + normu);
        const cplx v = fast_rcp(xi) * beta;
        const double t2 = xi * fast_rcp(-normu);
        const cplx x2 = (-c_abs) * (g[1] - ws);
        const cplx vA = t2 * (x2 + conj(v) * adiag);
        const cplx y = adiag - v * vA;
        const double ay = fast_sqrt(abs2(y));
        const double qn = __longlong_as_double(0x7ff8000000000000LL);
        if (G.absout) {
            G.absout[idx] = (st == ILA_OK) ? ay : (st == ILA_NOT_CONVERGED ? qn : -1.0);
        } else if (P.out_mode == 1) {
            // L[1,1] of a complex-eltype factorization is real (-ν of the k = 1 reflector): 8 B out; the status rides in
            // the NaN payload (low mantissa bits) and, when asked for, in a byte array
            const double l11 = (c_abs == 0.0) ? copysign(ay, y.re) : -ay;
            L11[idx] = (st == ILA_OK) ? l11 : __longlong_as_double(0x7ff8000000000000LL | (long long)st);
            if (status) reinterpret_cast<uint8_t *>(status)[idx] = (uint8_t)st;
        } else {
            status[idx] = st;
            const double l11 = (c_abs == 0.0) ? copysign(ay, y.re) : -ay;
            reinterpret_cast<double2 *>(L11)[idx] = (st == ILA_OK) ? make_double2(l11, 0.0) : make_double2(qn, qn);
            K3_DBG(reinterpret_cast<double2 *>(L11)[idx].y = dbg_f32 + 100.0 * dbg_f64 + 10000.0 * dbg_flags + 1e-3 * gapf;)
        }
        continue;
    }

    // eigenvector of the companion matrix in closed form (V_1 = 1).  The rows of (C/β) V = w V read V_{i+1} = g_i - w V_i
    // (i < N) and V_N = g_N / w.  Run forward, that recurrence multiplies rounding errors by |w| per step (|w|^(N-1):
    // 1e-6 relative for N = 8, |w| = 36) — so it is run BACKWARD from the last row, V_i = (g_i - V_{i+1}) / w, which damps
    // them by 1/|w| per step and is stable exactly where a factorization exists (|w| >= 1).
    cplx V[N];
    V[0] = mk(1.0, 0.0);
    {
        const double w2 = abs2(ws);
        const cplx winv = (w2 > 0.0) ? (1.0 / w2) * conj(ws) : mk(0.0, 0.0);
        V[N - 1] = g[N] * winv;
#pragma unroll
        for (int i = N - 2; i >= 1; i--) V[i] = (g[i + 1] - V[i + 1]) * winv;
    }

    // de = c_sgn c_abs V[end:-1:1]; with V_1 = 1 the reference's c_sgn is exactly -1 (complex branch,
    // infqltoeplitz.jl:24-26, since V_1 a_{m-1} - V_2 a_m = μ V_1).  Real branch: see below.
    const double c_abs = sqrt(n2 - P.beta2);

    // X = [a^T; 0 de^T]  (2 x m)
    cplx X1[M], X2[M];
#pragma unroll
    for (int j = 0; j < M; j++) X1[j] = P.a[j];
    X1[N - 1] = adiag;
    X2[0] = mk(0.0, 0.0);
#pragma unroll
    for (int k = 1; k <= N; k++) X2[k] = (-c_abs) * V[N - k];

    // ---- ql_X! : k = 2 reflector on (X[2,m], X[1,m]) applied to columns 1..m-1
    cplx tau1 = mk(0.0, 0.0), tau2 = mk(0.0, 0.0);
    {
        cplx xi1 = X2[N];
        const double normu = sqrt(abs2(X2[N]) + abs2(X1[N]));
        if (normu != 0.0) {
            const double nu = copysign(normu, xi1.re);
            xi1.re += nu;
            X2[N] = mk(-nu, 0.0);
            X1[N] = fast_rcp(abs2(xi1)) * (X1[N] * conj(xi1));
            tau2 = fast_rcp(nu) * xi1;
#pragma unroll
            for (int j = 0; j < N; j++) {
                cplx vA = X2[j] + conj(X1[N]) * X1[j];
                vA = conj(tau2) * vA;
                X2[j] = X2[j] - vA;
                X1[j] = X1[j] - X1[N] * vA;
            }
        }
    }
    double s;   // sign(real(X[2,end])) taken BEFORE ql! in the reference
    if (REAL_PATH) {
        // T<:Real (infqltoeplitz.jl:7-16): de = c * real(V[end:-1:1, j]) with LAPACK's eigenvector sign.  Row 1 of X
        // after the reflector does not depend on that sign, row 2 flips with it; the sign LAPACK returns in every
        // case the reference tests (test_infql.jl:12-19) is the one for which ql_X! needs no flip, i.e.
        // sign(de[end]) == sign(X[1,m-1]).  We pick that sign.
        double t = dsign(X1[N - 1].re);
        if (t == 0.0) t = 1.0;
        // X2 was built with sign -1: de_N = -c_abs.  Flip row 2 when t = +1.
        if (t > 0.0) {
#pragma unroll
            for (int j = 0; j < M; j++) X2[j] = -X2[j];
            X1[N] = -X1[N];   // v = X[1,m]/(ξ1+ν) flips with ξ1, ν
        }
        s = (c_abs == 0.0) ? 0.0 : t;
        // real eltype: ql! skips k = 1 (the (T<:Real) lower bound, examples/jacobi.jl:63)
    } else {
        s = dsign(-c_abs);
        // k = 1 (complex only): length-1 reflector making X[1,m-1] real
        cplx xi1 = X1[N - 1];
        const double normu = sqrt(abs2(xi1));
        if (normu != 0.0) {
            const double nu = copysign(normu, xi1.re);
            xi1.re += nu;
            X1[N - 1] = mk(-nu, 0.0);
            tau1 = fast_rcp(nu) * xi1;
#pragma unroll
            for (int j = 0; j < N - 1; j++) X1[j] = X1[j] - conj(tau1) * X1[j];
        }
    }
    // sign normalisation (infqltoeplitz.jl:34-37)
    if (s != dsign(X1[N - 1].re)) {
        tau1 = mk(2.0, 0.0) - tau1;
#pragma unroll
        for (int j = 0; j < N; j++) X1[j] = -X1[j];
    }

    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (st != ILA_OK) {
#pragma unroll
        for (int j = 0; j < M; j++) { X1[j] = mk(qnan, qnan); X2[j] = mk(qnan, qnan); }
        tau1 = tau2 = mk(qnan, qnan);
    }
    if (G.absout) {
        G.absout[idx] = (st == ILA_OK) ? sqrt(abs2(X1[N - 1])) : (st == ILA_NOT_CONVERGED ? qnan : -1.0);
        continue;
    }
    status[idx] = st;
    if (WANT_TAIL && SV.x) {
        // fused K7: x = L \ (Q'b) straight from the registers (NaN rows where the factorization does not exist)
        if (st == ILA_OK) {
            tz_solve_rows<N, REAL_PATH>(SV.S, X1[N - 1], X1[N], X2, tau1, tau2, SV.nout, n, idx, SV.x, SV.mode, SV.nstop);
        } else {
            for (int64_t j = 0; j < SV.nout; j++) {
                if (REAL_PATH) SV.x[j * n + idx] = qnan;
                else reinterpret_cast<double2 *>(SV.x)[j * n + idx] = make_double2(qnan, qnan);
            }
            if (SV.nstop) SV.nstop[idx] = 0;
        }
        continue;
    }
    if (REAL_PATH) {
        L11[idx] = X1[N - 1].re;
        if (WANT_TAIL) {
            double *t = tail + idx * (2 * M + 2);
#pragma unroll
            for (int j = 0; j < M; j++) { t[j] = X1[j].re; t[M + j] = X2[j].re; }
            t[2 * M] = tau1.re;
            t[2 * M + 1] = tau2.re;
        }
    } else {
        reinterpret_cast<double2 *>(L11)[idx] = make_double2(X1[N - 1].re, st == ILA_OK ? X1[N - 1].im : qnan);
        if (WANT_TAIL) {
            double2 *t = reinterpret_cast<double2 *>(tail) + idx * (2 * M + 2);
#pragma unroll
            for (int j = 0; j < M; j++) {
                t[j] = make_double2(X1[j].re, X1[j].im);
                t[M + j] = make_double2(X2[j].re, X2[j].im);
            }
            t[2 * M] = make_double2(tau1.re, tau1.im);
            t[2 * M + 1] = make_double2(tau2.re, tau2.im);
        }
    }
  }   // next shift of this thread
}

// ---- host-side launcher -------------------------------------------------------------------------
template <int N, bool REAL_PATH>
static int launch_n(const cplx *a_rev, int branch_rank, const double *d_lambda, int64_t n, double *d_L11,
                    double *d_tail, int32_t *d_status, cudaStream_t stream, TzGrid G, const TzSolveArgs &SV)
{
    TzParams<N> P;
    for (int j = 0; j <= N; j++) P.a[j] = a_rev[j];
    const cplx beta = a_rev[N];
    const double b2 = beta.re * beta.re + beta.im * beta.im;
    P.ginv = cplx{beta.re / b2, -beta.im / b2};
    P.beta2 = b2;
    P.branch_rank = branch_rank;
    P.max_it = g_tz_max_it.load() > 0 ? g_tz_max_it.load() : kAberthMaxIt;
    P.scout_max_it = g_tz_scout_max_it.load() > 0 ? g_tz_scout_max_it.load() : kScoutMaxIt;
    P.out_mode = G.out_mode;
    P.use_seeds = g_tz_seeds.load();
    for (int k = 0; k <= N; k++) {
        // cf0_k = (-1)^k a_{m-k}/β with a_{m-k} = a_rev[N-k]; cf0_0 = 1
        cplx v = (k == 0) ? cplx{1.0, 0.0} : cplx{a_rev[N - k].re * P.ginv.re - a_rev[N - k].im * P.ginv.im,
                                                  a_rev[N - k].re * P.ginv.im + a_rev[N - k].im * P.ginv.re};
        if (k & 1) v = cplx{-v.re, -v.im};
        P.cf0[k] = v;
        P.cf0f[k] = cplxf{(float)v.re, (float)v.im};
    }
    {
        double mx = 0.0;
        for (int k = 2; k <= N; k++) mx = fmax(mx, P.cf0[k].re * P.cf0[k].re + P.cf0[k].im * P.cf0[k].im);
        P.cmax2f = (float)fmin(mx, 1e30);
    }
    const double theta0 = 0.7;   // asymmetric start: a conjugate-symmetric start cannot split real root pairs
    for (int j = 0; j < N; j++) {
        const double th = theta0 + 2.0 * 3.14159265358979323846 * j / N;
        P.unit[j] = cplx{cos(th), sin(th)};
    }
    if (n <= 0) return 0;
    const int threads = 128;
    // shifts per thread: consecutive shifts share a warm start (longer runs amortise the cold start), but the grid
    // should still fill every SM to its resident-block limit: the smallest spt for which the grid is one full wave
    // resident blocks of this instantiation over the whole GPU, per device (tail / no tail); racing writers store the same value
    static std::atomic<int> wave_cache[16][2];
    const bool want_tail = d_tail != nullptr || SV.x != nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    const int dslot = (dev >= 0 && dev < 16) ? dev : 0;
    int wave_blocks = (dev == dslot) ? wave_cache[dslot][want_tail ? 1 : 0].load() : 0;
    if (wave_blocks == 0) {
        int sms = 148, per_sm = 4;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (want_tail)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ql_toeplitz_l11_kernel<N, REAL_PATH, true>, threads, 0);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ql_toeplitz_l11_kernel<N, REAL_PATH, false>, threads, 0);
        wave_blocks = sms * (per_sm > 0 ? per_sm : 1);
        if (dev == dslot) wave_cache[dslot][want_tail ? 1 : 0].store(wave_blocks);
    }
    // Run length.  r = n / (threads of one full wave).
    //  * degrees with closed-form cold seeds (N = 2, 4): a cold shift costs ~1.25 warm ones, so occupancy wins — the smallest
    //    run length for which the grid is one full wave (measured on B200, 1024-wide rows, µs per launch at the best spt:
    //    16-128 Ki shifts 10-14 at spt 1, 256 Ki 18.4 at 2, 512 Ki 26.6 at 4, 1 Mi 43.0 at 8, 4 Mi 130 at 16;
    //    profiles/r02_k3_kernel_vs_slab.json);
    //  * other degrees (circle start + Aberth sweeps: a cold shift costs ~4 warm ones): below ~8 waves HALF the resident
    //    threads with runs twice as long are faster; from 8 waves up the one-wave run length again.
    const double r = (double)n / ((double)wave_blocks * threads);
    const int full = (int)ceil(r);
    int spt = full;
    if (!((N == 2 || N == 4) && g_tz_seeds.load())) {
        spt = (int)ceil(2.0 * r + 0.5);
        if (spt > 8 && spt > full) spt = full > 8 ? full : 8;
    }
    while (spt & (spt - 1)) spt &= spt - 1;   // power of two: runs then do not straddle the rows of power-of-two grids
    if (g_tz_spt.load() > 0) spt = g_tz_spt.load();
    if (spt < 1) spt = 1;
    if (spt > 16) spt = 16;
    const int64_t nthreads = (n + spt - 1) / spt;
    const int64_t blocks = (nthreads + threads - 1) / threads;
    if (want_tail)
        ql_toeplitz_l11_kernel<N, REAL_PATH, true><<<(unsigned)blocks, threads, 0, stream>>>(P, d_lambda, n, spt, d_L11, d_tail, d_status, G, SV);
    else
        ql_toeplitz_l11_kernel<N, REAL_PATH, false><<<(unsigned)blocks, threads, 0, stream>>>(P, d_lambda, n, spt, d_L11, d_tail, d_status, G, SV);
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <bool REAL_PATH>
static int launch_any(int l, const cplx *a_rev, int branch_rank, const double *d_lambda, int64_t n, double *d_L11,
                      double *d_tail, int32_t *d_status, cudaStream_t stream, TzGrid G, const TzSolveArgs &SV)
{
    switch (l + 1) {
    case 2: return launch_n<2, REAL_PATH>(a_rev, branch_rank, d_lambda, n, d_L11, d_tail, d_status, stream, G, SV);
    case 3: return launch_n<3, REAL_PATH>(a_rev, branch_rank, d_lambda, n, d_L11, d_tail, d_status, stream, G, SV);
    case 4: return launch_n<4, REAL_PATH>(a_rev, branch_rank, d_lambda, n, d_L11, d_tail, d_status, stream, G, SV);
    case 5: return launch_n<5, REAL_PATH>(a_rev, branch_rank, d_lambda, n, d_L11, d_tail, d_status, stream, G, SV);
    case 6: return launch_n<6, REAL_PATH>(a_rev, branch_rank, d_lambda, n, d_L11, d_tail, d_status, stream, G, SV);
    case 7: return launch_n<7, REAL_PATH>(a_rev, branch_rank, d_lambda, n, d_L11, d_tail, d_status, stream, G, SV);
    case 8: return launch_n<8, REAL_PATH>(a_rev, branch_rank, d_lambda, n, d_L11, d_tail, d_status, stream, G, SV);
    default: return set_error(ILA_ERR_UNSUPPORTED, "ql_toeplitz: lower bandwidth l must be in 1..7");
    }
}

// Exposed to api.cu.  `a` is in BandedMatrices data-column order (a_{+1}, a_0, ..., a_{-l}); `reverse` (infqltoeplitz.jl:59)
// is done here.
int launch_ql_toeplitz_ex(bool real_path, int l, int u, const double *a_host, int branch_rank, const double *d_lambda,
                          int64_t n, double *d_L11, double *d_tail, int32_t *d_status, cudaStream_t stream, TzGrid G,
                          const TzSolveArgs &SV);

int launch_ql_toeplitz(bool real_path, int l, int u, const double *a_host, int branch_rank, const double *d_lambda,
                       int64_t n, double *d_L11, double *d_tail, int32_t *d_status, cudaStream_t stream)
{
    return launch_ql_toeplitz_ex(real_path, l, u, a_host, branch_rank, d_lambda, n, d_L11, d_tail, d_status, stream,
                                 TzGrid{nullptr, nullptr, 0, nullptr, 0}, TzSolveArgs{});
}

// complex eltype, L11 only, compact output: d_L11re = n doubles (NaN with the status in its payload where status != 0),
// d_status8 = n bytes or nullptr
int launch_ql_toeplitz_l11re(int l, int u, const double *a_host, int branch_rank, const double *d_lambda, int64_t n,
                             double *d_L11re, uint8_t *d_status8, cudaStream_t stream)
{
    return launch_ql_toeplitz_ex(false, l, u, a_host, branch_rank, d_lambda, n, d_L11re, nullptr,
                                 reinterpret_cast<int32_t *>(d_status8), stream, TzGrid{nullptr, nullptr, 0, nullptr, 1}, TzSolveArgs{});
}

// grid mode: |L11| (or -1) for λ = gx[j] + i*gy[k], idx = k*nx + j, k < ny
int launch_ql_toeplitz_grid(int l, int u, const double *a_host, int branch_rank, const double *d_gx, int nx,
                            const double *d_gy, int64_t ny, double *d_absout, cudaStream_t stream)
{
    return launch_ql_toeplitz_ex(false, l, u, a_host, branch_rank, nullptr, (int64_t)nx * ny, nullptr, nullptr, nullptr,
                                 stream, TzGrid{d_gx, d_gy, nx, d_absout, 0}, TzSolveArgs{});
}

int launch_ql_toeplitz_ex(bool real_path, int l, int u, const double *a_host, int branch_rank, const double *d_lambda,
                          int64_t n, double *d_L11, double *d_tail, int32_t *d_status, cudaStream_t stream, TzGrid G,
                          const TzSolveArgs &SV)
{
    if (u != 1) return set_error(ILA_ERR_UNSUPPORTED, "ql_toeplitz: only u == 1 (Hessenberg) symbols are supported (ql_pruneband is undefined upstream)");
    if (l < 1 || l + 1 > kMaxDeg) return set_error(ILA_ERR_UNSUPPORTED, "ql_toeplitz: lower bandwidth l must be in 1..7");
    if (branch_rank < 1 || branch_rank > l + 1) return set_error(ILA_ERR_ARG, "ql_toeplitz: branch_rank out of range");
    const int m = l + 2;
    cplx a_rev[kMaxDeg + 1];
    for (int j = 0; j < m; j++) {
        const int src = m - 1 - j;
        a_rev[j] = real_path ? cplx{a_host[src], 0.0} : cplx{a_host[2 * src], a_host[2 * src + 1]};
    }
    if (a_rev[m - 1].re == 0.0 && a_rev[m - 1].im == 0.0)
        return set_error(ILA_ERR_ARG, "ql_toeplitz: super-diagonal coefficient is zero (u would be 0)");
    return real_path ? launch_any<true>(l, a_rev, branch_rank, d_lambda, n, d_L11, d_tail, d_status, stream, G, SV)
                     : launch_any<false>(l, a_rev, branch_rank, d_lambda, n, d_L11, d_tail, d_status, stream, G, SV);
}


// ---- K4: perturbed-Toeplitz head ---------------------------------------------------------------------------
// ql_hessenberg!(B) (src/infql.jl:70-93) for A = finite head + Toeplitz tail, bandwidths (l, 1): the tail fixed point
// comes from the K3 kernel above (its `tail` output: X[1,:], X[2,:], tau); this kernel forms row 1 of Q∞' from the
// 2x2 block q = QLPackedQ(X[:, m-1:m], tau)  ((Q∞')[1,j] = conj(q11 q21^(j-1)), src/banded/hessenbergq.jl:125-135,162),
// overwrites the last row of the p x p head  (B~[p,p] = L∞[1,1],  B~[p, p-l'+1:p-1] = (Q∞')[1,1:l'-1] T[l':2(l'-1),1:l'-1],
// src/infql.jl:81-82) and runs the finite Householder QL of the head (src/infql.jl:83); L[1,1] = factors[1,1].
constexpr int kHeadMax = 8;

struct TzHeadParams {
    int N, p;                              // N = l + 1 = l' ; head is p x p, p >= l'
    cplx a_rev[kMaxDeg + 1];               // unshifted tail symbol a_{-l} .. a_{+1}
    cplx h[kHeadMax * kHeadMax];           // unshifted p x p top-left corner of A, column-major
};

template <bool REAL_PATH>
__global__ void __launch_bounds__(128)
ql_pert_head_kernel(TzHeadParams H, const double *__restrict__ lambda, int64_t n, const double *__restrict__ tail,
                    const int32_t *__restrict__ status, double *__restrict__ L11, double *__restrict__ head_factors,
                    TzSolveArgs SV)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int N = H.N, M = N + 1, p = H.p, W = REAL_PATH ? 1 : 2;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    auto put = [&](cplx v) {
        if (REAL_PATH) L11[idx] = v.re;
        else reinterpret_cast<double2 *>(L11)[idx] = make_double2(v.re, v.im);
    };
    auto putx = [&](int64_t j, cplx val) {
        if (REAL_PATH) SV.x[j * n + idx] = val.re;
        else reinterpret_cast<double2 *>(SV.x)[j * n + idx] = make_double2(val.re, val.im);
    };
    if (status[idx] != ILA_OK) {
        put(mk(qnan, qnan));
        if (SV.x)
            for (int64_t j = 0; j < SV.nout; j++) putx(j, mk(qnan, qnan));
        return;
    }
    const double *t = tail + idx * (2 * M + 2) * W;
    auto tl = [&](int k) -> cplx { return REAL_PATH ? mk(t[k], 0.0) : mk(t[2 * k], t[2 * k + 1]); };
    const cplx lam = REAL_PATH ? mk(lambda[idx], 0.0) : mk(lambda[2 * idx], lambda[2 * idx + 1]);
    const cplx l11inf = tl(N - 1), v = tl(N), tau1 = tl(2 * M), tau2 = tl(2 * M + 1);
    // q = H_2 H_1,  H_1 = diag(1 - tau1, 1),  H_2 = I - tau2 [v;1][v;1]'
    const cplx one = mk(1.0, 0.0);
    const cplx s1 = one - tau1;
    const cplx q11 = (one - abs2(v) * tau2) * s1;
    const cplx q21 = (-(conj(v) * tau2)) * s1;
    cplx B[kHeadMax][kHeadMax];
    for (int j = 0; j < p; j++)
        for (int i = 0; i < p; i++) {
            cplx e = H.h[i + j * p];
            if (i == j) e = e - lam;
            B[i][j] = e;
        }
    B[p - 1][p - 1] = l11inf;
    {
        cplx qrow[kMaxDeg];
        cplx pw = q11;
        for (int j = 0; j < N - 1; j++) { qrow[j] = conj(pw); pw = pw * q21; }
        for (int jj = 0; jj < N - 1; jj++) {
            cplx acc = mk(0.0, 0.0);
            for (int ii = 0; ii <= jj; ii++) acc = acc + qrow[ii] * H.a_rev[jj - ii];   // T[l'-1+ii+1, jj+1] = a_rev[jj-ii]
            B[p - 1][p - N + jj] = acc;
        }
    }
    // finite Householder QL of the head (reference loop: k = p..1 complex, p..2 real)
    cplx ftau[kHeadMax];
    for (int k = 0; k < kHeadMax; k++) ftau[k] = mk(0.0, 0.0);
    for (int k = p; k >= (REAL_PATH ? 2 : 1); k--) {
        double s = abs2(B[k - 1][k - 1]);
        for (int i = 1; i < k; i++) s += abs2(B[k - 1 - i][k - 1]);
        const double normu = sqrt(s);
        cplx tk = mk(0.0, 0.0);
        if (normu != 0.0) {
            cplx xi1 = B[k - 1][k - 1];
            const double nu = copysign(normu, xi1.re);
            xi1.re += nu;
            B[k - 1][k - 1] = mk(-nu, 0.0);
            const double r2 = 1.0 / abs2(xi1);
            for (int i = 1; i < k; i++) B[k - 1 - i][k - 1] = r2 * (B[k - 1 - i][k - 1] * conj(xi1));
            tk = (1.0 / nu) * xi1;
        }
        ftau[k - 1] = tk;
        for (int j = 0; j < k - 1; j++) {
            cplx vA = B[k - 1][j];
            for (int i = 1; i < k; i++) vA = vA + conj(B[k - 1 - i][k - 1]) * B[k - 1 - i][j];
            vA = conj(tk) * vA;
            B[k - 1][j] = B[k - 1][j] - vA;
            for (int i = 1; i < k; i++) B[k - 1 - i][j] = B[k - 1 - i][j] - B[k - 1 - i][k - 1] * vA;
        }
    }
    put(B[0][0]);
    if (SV.x) {
        // ---- F \ b for the perturbed operator (src/infql.jl:70-93 assembly, :266-289 solve) ----
        // 2x2 blocks: head n <= p-1 from the packed head factors (hessenbergq.jl:182-190), n >= p the tail block q∞
        const cplx q12 = -(v * tau2), q22 = one - tau2;
        auto block = [&](int nblk, cplx &b11, cplx &b12, cplx &b21, cplx &b22) {   // 1-based block index
            if (nblk <= p - 1) {
                const cplx vv = B[nblk - 1][nblk], tb_ = ftau[nblk];           // reflector of column nblk+1
                const cplx sa = (nblk == 1) ? (one - ftau[0]) : one;
                b11 = (one - abs2(vv) * tb_) * sa;
                b21 = (-(conj(vv) * tb_)) * sa;
                b12 = -(vv * tb_);
                b22 = one - tb_;
            } else {
                b11 = q11; b12 = q12; b21 = q21; b22 = q22;
            }
        };
        constexpr int KW = kHeadMax + kMaxDeg + kSolveMaxB + 2;
        cplx tw[KW];
        for (int k = 0; k < KW; k++) tw[k] = (k < SV.S.nb) ? SV.S.b[k] : mk(0.0, 0.0);
        for (int k = SV.S.nb; k >= 1; k--) {
            cplx b11, b12, b21, b22;
            block(k, b11, b12, b21, b22);
            const cplx u0 = tw[k - 1], u1 = tw[k];
            tw[k - 1] = conj(b11) * u0 + conj(b21) * u1;
            tw[k] = conj(b12) * u0 + conj(b22) * u1;
        }
        // extension block E = (Q∞')[2:l'+1, 1:l'+1] T[l':2l', 1:l']  (rows p+1..p+l', columns p-l'+1..p; src/infql.jl:87)
        constexpr int QD = kMaxDeg + 3;
        cplx Qd[QD][QD];
        const int qd = N + 3;
        for (int i = 0; i < qd; i++)
            for (int j = 0; j < qd; j++) Qd[i][j] = (i == j) ? one : mk(0.0, 0.0);
        for (int nb_ = 0; nb_ <= N + 1; nb_++)
            for (int j = 0; j < qd; j++) {
                const cplx r0 = Qd[nb_][j], r1 = Qd[nb_ + 1][j];
                Qd[nb_][j] = q11 * r0 + q12 * r1;
                Qd[nb_ + 1][j] = q21 * r0 + q22 * r1;
            }
        cplx a_sh[kMaxDeg + 1];
        for (int k = 0; k < M; k++) a_sh[k] = H.a_rev[k];
        a_sh[M - 2] = a_sh[M - 2] - lam;
        auto Tent = [&](int i, int j) -> cplx {   // 1-based entry of the shifted Toeplitz tail
            const int k = (j - i) + M - 2;
            return (k >= 0 && k < M) ? a_sh[k] : mk(0.0, 0.0);
        };
        cplx E[kMaxDeg][kMaxDeg];
        for (int r = 0; r < N; r++)
            for (int c = 0; c < N; c++) {
                cplx acc = mk(0.0, 0.0);
                for (int k = 0; k <= N; k++) acc = acc + conj(Qd[k][r + 1]) * Tent(N + k, 1 + c);   // Qadj[r+1][k]
                E[r][c] = acc;
            }
        auto rcpc = [&](cplx d) { const double r2 = 1.0 / abs2(d); return r2 * conj(d); };
        // columns of the head (rows up to p + l' - 1 all live in tw), then the Toeplitz tail through a sliding window
        const int64_t jhead = SV.nout < p ? SV.nout : p;
        for (int j = 0; j < p; j++) {
            const cplx xj = tw[j] * rcpc(B[j][j]);
            tw[j] = xj;
            if (j < jhead) putx(j, xj);
            for (int i = j + 1; i < p; i++) tw[i] = tw[i] - B[i][j] * xj;
            const int c = j - (p - N);
            if (c >= 0)
                for (int r = 0; r < N; r++) tw[p + r] = tw[p + r] - E[r][c] * xj;
        }
        const cplx r_tail = rcpc(tl(M + M - 1));
        cplx sub[kMaxDeg];
        for (int i = 1; i <= N; i++) sub[i - 1] = tl(M + (M - 1 - i));
        cplx rw[kMaxDeg + 1];
        for (int k = 0; k <= N; k++) rw[k] = tw[p + k];
        for (int64_t j = p; j < SV.nout; j++) {
            const cplx xj = rw[0] * r_tail;
            putx(j, xj);
            const int64_t nx = j + N + 1;
            const cplx tin = (nx < KW) ? tw[nx] : mk(0.0, 0.0);
            for (int k = 0; k < N; k++) rw[k] = rw[k + 1] - sub[k] * xj;
            rw[N] = tin;
        }
    }
    if (head_factors) {
        double *o = head_factors + idx * (int64_t)p * p * W;
        for (int j = 0; j < p; j++)
            for (int i = 0; i < p; i++) {
                if (REAL_PATH) o[i + j * p] = B[i][j].re;
                else { o[2 * (i + j * p)] = B[i][j].re; o[2 * (i + j * p) + 1] = B[i][j].im; }
            }
    }
}

// head_host: p x p column-major top-left corner of A (no shift), same scalar type as the symbol
int launch_ql_pert_head_ex(bool real_path, int l, int p, const double *a_host, const double *head_host, const double *d_lambda,
                           int64_t n, const double *d_tail, const int32_t *d_status, double *d_L11, double *d_head_factors,
                           int nb, const double *b_host, int64_t nout, double *d_x, cudaStream_t stream);

int launch_ql_pert_head(bool real_path, int l, int p, const double *a_host, const double *head_host, const double *d_lambda,
                        int64_t n, const double *d_tail, const int32_t *d_status, double *d_L11, double *d_head_factors,
                        cudaStream_t stream)
{
    return launch_ql_pert_head_ex(real_path, l, p, a_host, head_host, d_lambda, n, d_tail, d_status, d_L11, d_head_factors, 0,
                                  nullptr, 0, nullptr, stream);
}

// with d_x != nullptr the kernel also solves F \ b (b_host: nb <= 16 scalars) and writes x column-major n x nout
int launch_ql_pert_head_ex(bool real_path, int l, int p, const double *a_host, const double *head_host, const double *d_lambda,
                           int64_t n, const double *d_tail, const int32_t *d_status, double *d_L11, double *d_head_factors,
                           int nb, const double *b_host, int64_t nout, double *d_x, cudaStream_t stream)
{
    if (p < l + 1 || p > kHeadMax) return set_error(ILA_ERR_UNSUPPORTED, "ql_perttoeplitz: head size p must satisfy l+1 <= p <= 8");
    TzHeadParams H;
    memset(&H, 0, sizeof(H));
    H.N = l + 1;
    H.p = p;
    const int m = l + 2;
    for (int j = 0; j < m; j++) {
        const int src = m - 1 - j;
        H.a_rev[j] = real_path ? cplx{a_host[src], 0.0} : cplx{a_host[2 * src], a_host[2 * src + 1]};
    }
    for (int k = 0; k < p * p; k++) H.h[k] = real_path ? cplx{head_host[k], 0.0} : cplx{head_host[2 * k], head_host[2 * k + 1]};
    TzSolveArgs SV{};
    if (d_x) {
        if (nb < 1 || nb > kSolveMaxB) return set_error(ILA_ERR_UNSUPPORTED, "ql_perttoeplitz_solve: right-hand side length must be in 1..16");
        if (nout < 1) return set_error(ILA_ERR_ARG, "ql_perttoeplitz_solve: nout must be positive");
        SV.S.nb = nb;
        for (int k = 0; k < kSolveMaxB; k++)
            SV.S.b[k] = k < nb ? (real_path ? cplx{b_host[k], 0.0} : cplx{b_host[2 * k], b_host[2 * k + 1]}) : cplx{0.0, 0.0};
        SV.nout = nout;
        SV.x = d_x;
    }
    if (n <= 0) return 0;
    const int threads = 128;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    if (real_path) ql_pert_head_kernel<true><<<blocks, threads, 0, stream>>>(H, d_lambda, n, d_tail, d_status, d_L11, d_head_factors, SV);
    else ql_pert_head_kernel<false><<<blocks, threads, 0, stream>>>(H, d_lambda, n, d_tail, d_status, d_L11, d_head_factors, SV);
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ---- K7: solve with the Toeplitz QL,  x = ql(A - λI) \ b  ------------------------------------------------------
// Reference: ldiv!(F, b) = ldiv!(F.L, lmul!(F.Q', b)) (src/infql.jl:266-267).  Q = LowerHessenbergQ(Fill(q, ∞)) with the
// 2x2 block q = QLPackedQ(X[:, m-1:m], τ) (infqltoeplitz.jl:63-64); Q' = UpperHessenbergQ(adjoint.(q)) acts on b[n:n+1]
// for n = nz, ..., 1 (src/banded/hessenbergq.jl:125-135).  L is the lower band with first column
// [X[1,m-1]; X[2,m-1:-1:1]] and every other column X[2,m:-1:1] (infqltoeplitz.jl:63), solved by the column-oriented
// forward substitution of src/infql.jl:269-289.  One thread per shift, consuming the K3 `tail` record; the N = m-1
// pending row updates live in a register window; x is written column-major (x[i + j n]: coalesced across shifts, and the
// n x nout Matrix a Julia caller wants).  The reference stops when the remaining entries are exactly zero; here the
// caller states how many entries it wants.
template <int N, bool REAL_PATH>
__global__ void __launch_bounds__(128)
ql_toeplitz_solve_kernel(TzSolveParams S, int64_t n, const double *__restrict__ tail, const int32_t *__restrict__ status,
                         int64_t nout, double *__restrict__ x)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    constexpr int M = N + 1, W = REAL_PATH ? 1 : 2;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (status[idx] != ILA_OK) {
        for (int64_t j = 0; j < nout; j++) {
            if (REAL_PATH) x[j * n + idx] = qnan;
            else reinterpret_cast<double2 *>(x)[j * n + idx] = make_double2(qnan, qnan);
        }
        return;
    }
    const double *t = tail + idx * (2 * M + 2) * W;
    auto tl = [&](int k) -> cplx { return REAL_PATH ? mk(t[k], 0.0) : mk(t[2 * k], t[2 * k + 1]); };
    cplx X2[N + 1];
#pragma unroll
    for (int k = 0; k <= N; k++) X2[k] = tl(M + k);
    tz_solve_rows<N, REAL_PATH>(S, tl(N - 1), tl(N), X2, tl(2 * M), tl(2 * M + 1), nout, n, idx, x);
}

template <bool REAL_PATH>
static int launch_solve_any(int l, const TzSolveParams &S, int64_t n, const double *d_tail, const int32_t *d_status,
                            int64_t nout, double *d_x, cudaStream_t stream)
{
    const int threads = 128;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    switch (l + 1) {
#define ILA_SOLVE_CASE(NN) \
    case NN: ql_toeplitz_solve_kernel<NN, REAL_PATH><<<blocks, threads, 0, stream>>>(S, n, d_tail, d_status, nout, d_x); break;
        ILA_SOLVE_CASE(2) ILA_SOLVE_CASE(3) ILA_SOLVE_CASE(4) ILA_SOLVE_CASE(5) ILA_SOLVE_CASE(6) ILA_SOLVE_CASE(7) ILA_SOLVE_CASE(8)
#undef ILA_SOLVE_CASE
    default: return set_error(ILA_ERR_UNSUPPORTED, "ql_toeplitz_solve: lower bandwidth l must be in 1..7");
    }
    count_launch();
    ILA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// fused form: one launch of the K3 kernel that solves from its registers (no tail record in memory)
int launch_ql_toeplitz_solve_fused(bool real_path, int l, int u, const double *a_host, int branch_rank, const double *d_lambda,
                                   int64_t n, int nb, const double *b_host, int64_t nout, double *d_L11, double *d_x,
                                   int32_t *d_status, cudaStream_t stream, int mode, int64_t *d_nstop)
{
    if (nb < 1 || nb > kSolveMaxB) return set_error(ILA_ERR_UNSUPPORTED, "ql_toeplitz_solve: right-hand side length must be in 1..16");
    if (nout < 1) return set_error(ILA_ERR_ARG, "ql_toeplitz_solve: nout must be positive");
    TzSolveArgs SV;
    SV.S.nb = nb;
    for (int k = 0; k < kSolveMaxB; k++)
        SV.S.b[k] = k < nb ? (real_path ? cplx{b_host[k], 0.0} : cplx{b_host[2 * k], b_host[2 * k + 1]}) : cplx{0.0, 0.0};
    SV.nout = nout;
    SV.x = d_x;
    SV.mode = mode;
    SV.nstop = d_nstop;
    return launch_ql_toeplitz_ex(real_path, l, u, a_host, branch_rank, d_lambda, n, d_L11, nullptr, d_status, stream,
                                 TzGrid{nullptr, nullptr, 0, nullptr, 0}, SV);
}

// b_host: nb scalars (real or interleaved complex, like the symbol); d_tail: the `tail` output of launch_ql_toeplitz
int launch_ql_toeplitz_solve(bool real_path, int l, int nb, const double *b_host, int64_t n, const double *d_tail,
                             const int32_t *d_status, int64_t nout, double *d_x, cudaStream_t stream)
{
    if (nb < 1 || nb > kSolveMaxB) return set_error(ILA_ERR_UNSUPPORTED, "ql_toeplitz_solve: right-hand side length must be in 1..16");
    if (nout < 1) return set_error(ILA_ERR_ARG, "ql_toeplitz_solve: nout must be positive");
    if (n <= 0) return 0;
    TzSolveParams S;
    S.nb = nb;
    for (int k = 0; k < kSolveMaxB; k++)
        S.b[k] = k < nb ? (real_path ? cplx{b_host[k], 0.0} : cplx{b_host[2 * k], b_host[2 * k + 1]}) : cplx{0.0, 0.0};
    return real_path ? launch_solve_any<true>(l, S, n, d_tail, d_status, nout, d_x, stream)
                     : launch_solve_any<false>(l, S, n, d_tail, d_status, nout, d_x, stream);
}


} // namespace ila
