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

} // namespace ila
