// ila_unfused.cuh — unfused IEEE scalar layer shared by the bit-exact kernels: every product / sum is a separate
// correctly rounded operation (no FMA contraction), quotients and square roots are IEEE (ediv / esqrt of ila_exact.cuh: the exact fast sequences in range, the intrinsics outside); complex products are
// the four-multiply form and complex quotients Smith's algorithm — the operation order of the C oracles.
#pragma once
#include "ila_common.cuh"
#include "ila_exact.cuh"

namespace ila {

__device__ __forceinline__ cplx zmul(cplx a, cplx b)
{
    return cplx{__dsub_rn(__dmul_rn(a.re, b.re), __dmul_rn(a.im, b.im)), __dadd_rn(__dmul_rn(a.re, b.im), __dmul_rn(a.im, b.re))};
}
__device__ __forceinline__ cplx zadd(cplx a, cplx b) { return cplx{__dadd_rn(a.re, b.re), __dadd_rn(a.im, b.im)}; }
__device__ __forceinline__ cplx zsub(cplx a, cplx b) { return cplx{__dsub_rn(a.re, b.re), __dsub_rn(a.im, b.im)}; }
__device__ __forceinline__ cplx zconj(cplx a) { return cplx{a.re, -a.im}; }
__device__ __forceinline__ double zabs2(cplx a) { return __dadd_rn(__dmul_rn(a.re, a.re), __dmul_rn(a.im, a.im)); }
__device__ __forceinline__ cplx zdiv(cplx a, cplx b)
{   // Smith's algorithm, branch-free: when |b.im| > |b.re| both operands are multiplied by -i first (a/b is unchanged and
    // every intermediate equals the textbook second branch bit for bit: r -> -r exactly, sums commute), so the 32 lanes of
    // a warp never diverge on the operand shape
    const bool sw = fabs(b.re) < fabs(b.im);
    const cplx aa = sw ? cplx{a.im, -a.re} : a;
    const cplx bb = sw ? cplx{b.im, -b.re} : b;
    const double r = ediv(bb.im, bb.re), den = __dadd_rn(bb.re, __dmul_rn(bb.im, r));
    return cplx{ediv(__dadd_rn(aa.re, __dmul_rn(aa.im, r)), den), ediv(__dsub_rn(aa.im, __dmul_rn(aa.re, r)), den)};
}

// real overloads with the same names, so that kernels templated on the scalar type read the same
__device__ __forceinline__ double zmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double zadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double zsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double zconj(double a) { return a; }
__device__ __forceinline__ double zabs2(double a) { return __dmul_rn(a, a); }
__device__ __forceinline__ double zdiv(double a, double b) { return ediv(a, b); }
__device__ __forceinline__ double zreal(double a) { return a; }
__device__ __forceinline__ double zreal(cplx a) { return a.re; }
__device__ __forceinline__ double zaddr(double a, double v) { return __dadd_rn(a, v); }          // a + (real) v
__device__ __forceinline__ cplx zaddr(cplx a, double v) { return cplx{__dadd_rn(a.re, v), a.im}; }
__device__ __forceinline__ double zdivr(double a, double v) { return ediv(a, v); }          // a / (real) v
__device__ __forceinline__ cplx zdivr(cplx a, double v) { return cplx{ediv(a.re, v), ediv(a.im, v)}; }
__device__ __forceinline__ double zmulr(double a, double v) { return __dmul_rn(a, v); }          // a * (real) v
__device__ __forceinline__ cplx zmulr(cplx a, double v) { return cplx{__dmul_rn(a.re, v), __dmul_rn(a.im, v)}; }
template <typename T> __device__ __forceinline__ T zfrom(double v);
template <> __device__ __forceinline__ double zfrom<double>(double v) { return v; }
template <> __device__ __forceinline__ cplx zfrom<cplx>(double v) { return cplx{v, 0.0}; }
template <typename T> struct z_is_cplx { static constexpr bool value = false; };
template <> struct z_is_cplx<cplx> { static constexpr bool value = true; };

} // namespace ila
