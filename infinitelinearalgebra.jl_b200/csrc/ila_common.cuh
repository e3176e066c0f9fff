// ila_common.cuh — shared device/host helpers of libila_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/ila_b200.h"

namespace ila {

// ---- error plumbing ---------------------------------------------------------------------------
extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launches;

inline int set_error(int code, const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(g_last_error, sizeof(g_last_error), fmt, a, b);
    return code;
}

#define ILA_CUDA_CHECK(expr)                                                                      \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess) {                                                               \
            snprintf(ila::g_last_error, sizeof(ila::g_last_error), "%s:%d: %s -> %s", __FILE__,   \
                     __LINE__, #expr, cudaGetErrorString(err__));                                 \
            return ILA_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- complex FP64 arithmetic (struct-of-two-doubles, register resident) -------------------------
struct cplx {
    double re, im;
};
__host__ __device__ __forceinline__ cplx mk(double re, double im = 0.0) { return cplx{re, im}; }
__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
__host__ __device__ __forceinline__ cplx operator-(cplx a) { return cplx{-a.re, -a.im}; }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b)
{
    return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__host__ __device__ __forceinline__ cplx operator*(double s, cplx a) { return cplx{s * a.re, s * a.im}; }
__host__ __device__ __forceinline__ cplx conj(cplx a) { return cplx{a.re, -a.im}; }
__host__ __device__ __forceinline__ double abs2(cplx a) { return a.re * a.re + a.im * a.im; }
__host__ __device__ __forceinline__ cplx operator/(cplx a, cplx b)
{
    // textbook division; operands on this path are O(1)-scaled (roots normalised by the super-diagonal)
    double d = 1.0 / (b.re * b.re + b.im * b.im);
    return cplx{(a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d};
}
__host__ __device__ __forceinline__ cplx operator/(cplx a, double s) { return cplx{a.re / s, a.im / s}; }

// ---- bounds-checked pointers (debug build, -DILA_BOUNDS; `make debug` -> libila_b200_dbg.so) -------------------------------
// compute-sanitizer is not available on the GPU pool, so the kernels with non-trivial index arithmetic (circular shared-memory
// panel of K11, group-interleaved records of K1/K2/K5/K12, ring slots of the back-substitution) address their buffers
// through ILA_BP(T): a plain T* in the release build, a (pointer, extent, site) triple in the debug build whose operator[]
// records every out-of-range index (count, first site, first index, extent) in a per-translation-unit device counter that
// ila_debug_bounds_report() collects.  tests/test_gpu_bounds.py runs every kernel on small cases under the debug library.
#ifdef ILA_BOUNDS
static __device__ unsigned long long g_ila_bounds[4];
__device__ __noinline__ static void ila_bounds_fail(int site, long long i, long long n)
{
    if (atomicAdd(&g_ila_bounds[0], 1ULL) == 0ULL) {
        g_ila_bounds[1] = (unsigned long long)site;
        g_ila_bounds[2] = (unsigned long long)i;
        g_ila_bounds[3] = (unsigned long long)n;
    }
}
template <class T>
struct bptr {
    T *p;
    long long lo, hi;   // valid indices relative to p: lo <= i < hi (pointer arithmetic moves the window with the pointer)
    int site;
    __device__ __forceinline__ T &operator[](long long i) const
    {
        if (i < lo || i >= hi) { ila_bounds_fail(site, i, hi); return p[lo]; }
        return p[i];
    }
    __device__ __forceinline__ bptr operator+(long long k) const { return bptr{p + k, lo - k, hi - k, site}; }
    __device__ __forceinline__ bptr operator-(long long k) const { return bptr{p - k, lo + k, hi + k, site}; }
    __device__ __forceinline__ bptr &operator+=(long long k) { p += k; lo -= k; hi -= k; return *this; }
    __device__ __forceinline__ operator bptr<const T>() const { return bptr<const T>{p, lo, hi, site}; }
    __device__ __forceinline__ T *raw() const { return p; }
};
#define ILA_BP(T) ila::bptr<T>
#define ILA_MKBP(T, ptr, extent, site) ila::bptr<T>{(ptr), 0LL, (long long)(extent), (site)}
#define ILA_RAW(bp) ((bp).raw())
#define ILA_BOUNDS_REPORT_FN(name)                                                                   \
    int name(unsigned long long *out4)                                                               \
    {                                                                                                \
        ILA_CUDA_CHECK(cudaMemcpyFromSymbol(out4, ila::g_ila_bounds, 4 * sizeof(unsigned long long))); \
        unsigned long long zero[4] = {0, 0, 0, 0};                                                   \
        ILA_CUDA_CHECK(cudaMemcpyToSymbol(ila::g_ila_bounds, zero, sizeof(zero)));                   \
        return 1;                                                                                    \
    }
#else
#define ILA_BP(T) T *
#define ILA_MKBP(T, ptr, extent, site) (ptr)
#define ILA_RAW(bp) (bp)
#define ILA_BOUNDS_REPORT_FN(name) \
    int name(unsigned long long *) { return 0; }
#endif

} // namespace ila
