// Microbenchmarks that informed the kernel design (run on a B200 with: nvcc -arch=sm_100a -O3 fp64_mma.cu && ./a.out)
//  1. DFMA throughput (independent chains)            2. DMMA m8n8k4 (FP64 tensor) throughput
//  3. DFMA + DMMA issued together (do they share a pipe?)   4. dependent-chain latencies: DFMA, SHFL+DADD
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters)
{
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], 0.999999, 1e-9);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 1.2345) out[0] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double *out, int iters)
{
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 0.25;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345) out[0] = s;
}

__global__ void k_both(double *out, int iters)
{
    double c[4][2], f[8];
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int i = 0; i < 8; ++i) f[i] = threadIdx.x + i;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 0.25;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fma(f[i], 0.999999, 1e-9);
        }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    for (int i = 0; i < 8; ++i) s += f[i];
    if (s == 1.2345) out[0] = s;
}

__global__ void k_lat(long long *out)
{
    double a = threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) a = fma(a, 0.999999, 1e-9);
    long long t1 = clock64();
    double b = a;
#pragma unroll
    for (int i = 0; i < 256; ++i) b = b + __shfl_xor_sync(0xffffffffu, b, 1);
    long long t2 = clock64();
    double c0 = b, c1 = a;
#pragma unroll
    for (int i = 0; i < 256; ++i) dmma(c0, c1, 1.0, 0.25);
    long long t3 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0); out[1] = (t2 - t1); out[2] = (t3 - t2); out[3] = (long long)(b + c0 + c1); }
}

template <class F> float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    double *out; cudaMalloc(&out, 64);
    long long *lo; cudaMalloc(&lo, 64);
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    float ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
    printf("DFMA : %.2f TFLOP/s\n", 2.0 * 64 * iters * blocks * threads / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dmma<<<blocks, threads>>>(out, iters); });
    printf("DMMA : %.2f TFLOP/s (m8n8k4, 32 per loop trip per warp)\n", 2.0 * 256 * 32 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_both<<<blocks, threads>>>(out, iters); });
    const double fl = (2.0 * 256 * 16 * (threads / 32) + 2.0 * 32 * threads) * iters * (double)blocks;
    printf("both : %.2f TFLOP/s combined (DMMA share %.0f%%)\n", fl / (ms * 1e-3) / 1e12, 100.0 * 2.0 * 256 * 16 * (threads / 32) / (2.0 * 256 * 16 * (threads / 32) + 2.0 * 32 * threads));
    k_lat<<<1, 32>>>(lo);
    long long h[4]; cudaMemcpy(h, lo, 32, cudaMemcpyDeviceToHost);
    printf("latency (cycles per dependent op): DFMA %.1f, SHFL+DADD %.1f, DMMA %.1f\n", h[0] / 256.0, h[1] / 256.0, h[2] / 256.0);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
