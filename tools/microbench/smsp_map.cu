// Which warps of a CTA share an SM sub-partition (and therefore an FP64 pipe)? One CTA of 8 warps on one SM; warps 0 and j
// run the same DFMA burst, the others exit. If the pair shares a sub-partition the burst takes twice as long.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smsp_map smsp_map.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(int j, int iters, double *out, long long *cyc)
{
    const int w = threadIdx.x >> 5;
    if (w != 0 && w != j) return;
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    const double m = 1.0000001, c = 1e-9;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[w] = t1 - t0;
}

int main()
{
    double *out;
    long long *cyc, h[8];
    cudaMalloc(&out, 256 * 8);
    cudaMalloc(&cyc, 64);
    const int iters = 20000;
    for (int j = 0; j < 8; ++j) {
        cudaMemset(cyc, 0, 64);
        k<<<1, 256>>>(j, iters, out, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("warps 0 and %d: %.3f cycles per DFMA (warp 0), %.3f (warp %d)\n", j, (double)h[0] / (8.0 * iters), (double)h[j] / (8.0 * iters), j);
    }
    return 0;
}
