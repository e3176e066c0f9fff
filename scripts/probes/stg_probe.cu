// stg_probe.cu — what does one warp pay per global store instruction when it is the only warp of its SM sub-partition?
//   scattered: lane l writes x[l * stride + it]  (32 lines per instruction)      coalesced: x[it * 32 + l]
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stg_probe stg_probe.cu && ./stg_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *x, long long stride, int iters, int per, long long *cyc, int blocksper)
{
    const int lane = threadIdx.x & 31;
    double v = 1.0 + lane;
    double *base = x + (long long)blockIdx.x * 32 * stride;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 1
        for (int p = 0; p < per; p++) {
            const long long r = (long long)it * per + p;
            if (stride == 1) base[r * 32 + lane] = v;
            else base[lane * stride + r] = v;
            v += 1.0;
        }
        // a little FP64 work between the bursts, like a row group
        for (int q = 0; q < 8; q++) v = fma(v, 1.0000001, 0.5);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (v == 12345.678) x[0] = v;
}
int main()
{
    double *x; long long *c, h;
    const long long stride = 1 << 17;           // 1 MiB apart
    cudaMalloc(&x, 148ull * 32 * stride * 8); cudaMalloc(&c, 8);
    for (int blocks : {1, 128}) for (int per : {8, 16}) for (long long st : {stride, 1LL}) {
        const int iters = 4000;
        k<<<blocks, 32>>>(x, st, iters, per, c, 0); cudaDeviceSynchronize();
        k<<<blocks, 32>>>(x, st, iters, per, c, 0); cudaDeviceSynchronize();
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("blocks %3d  %2d stores per burst  %-9s: %.1f cycles per store instruction (incl. %d FP64 ops per burst)\n", blocks, per, st == 1 ? "coalesced" : "scattered", (double)h / iters / per, 8);
    }
    return 0;
}
