// fp64_issue_probe.cu — how fast can ONE warp issue FP64 work on B200, and does a partially active warp issue faster?
// (The adaptive-QR sweeps run one warp per SM: their time per column is the warp's issue time, not the chip's throughput.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue_probe fp64_issue_probe.cu && ./fp64_issue_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void probe(double *out, long long *cyc, int active, int iters)
{
    const int lane = threadIdx.x & 31;
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = 1.0 + 1e-9 * (lane + i);
    const double m = 1.0000001, c = 1e-9;
    long long t0 = 0, t1 = 0;
    if (lane < active) {
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
        }
        t1 = clock64();
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[threadIdx.x] = s;
    if (lane == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}

template <int ILP>
__global__ void probe_f32(float *out, long long *cyc, int active, int iters)
{
    const int lane = threadIdx.x & 31;
    float a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = 1.0f + 1e-5f * (lane + i);
    const float m = 1.00001f, c = 1e-7f;
    long long t0 = 0, t1 = 0;
    if (lane < active) {
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) a[i] = fmaf(a[i], m, c);
        }
        t1 = clock64();
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[threadIdx.x] = s;
    if (lane == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}

int main()
{
    double *d;
    float *f;
    long long *c, h[8];
    cudaMalloc(&d, 4096);
    cudaMalloc(&f, 4096);
    cudaMalloc(&c, 64);
    const int iters = 4096;
    printf("one warp, FP64 FMA: cycles per warp-instruction (ILP = independent chains)\n");
    for (int active : {32, 16, 8, 1}) {
        probe<1><<<1, 32>>>(d, c, active, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost); double l1 = (double)h[0] / iters;
        probe<4><<<1, 32>>>(d, c, active, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost); double l4 = (double)h[0] / iters / 4;
        probe<8><<<1, 32>>>(d, c, active, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost); double l8 = (double)h[0] / iters / 8;
        probe<16><<<1, 32>>>(d, c, active, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost); double l16 = (double)h[0] / iters / 16;
        printf("  active lanes %2d: ILP1 %.2f (latency)  ILP4 %.2f  ILP8 %.2f  ILP16 %.2f\n", active, l1, l4, l8, l16);
    }
    printf("one warp, FP32 FMA:\n");
    for (int active : {32, 16}) {
        probe_f32<1><<<1, 32>>>(f, c, active, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost); double l1 = (double)h[0] / iters;
        probe_f32<8><<<1, 32>>>(f, c, active, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost); double l8 = (double)h[0] / iters / 8;
        printf("  active lanes %2d: ILP1 %.2f (latency)  ILP8 %.2f\n", active, l1, l8);
    }
    printf("four warps of one block (one per SM sub-partition), FP64 ILP8, cycles per warp-instruction per warp:\n");
    probe<8><<<1, 128>>>(d, c, 32, iters); cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
    printf("  %.2f %.2f %.2f %.2f\n", (double)h[0] / iters / 8, (double)h[1] / iters / 8, (double)h[2] / iters / 8, (double)h[3] / iters / 8);
    printf("eight warps of one block, FP64 ILP8:\n");
    probe<8><<<1, 256>>>(d, c, 32, iters); cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 8; i++) printf("  %.2f", (double)h[i] / iters / 8);
    printf("\n");
    return 0;
}
