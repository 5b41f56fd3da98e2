// va_util.cu -- small support kernels: deterministic row reduction, the seeded synthetic-input generator and the two
// microbenchmarks that give the roofline denominators (FP64 DFMA peak, HBM copy bandwidth) on the device at hand.
#include <algorithm>

#include "va_common.cuh"

namespace {

// ---- out[k] (+)= sum_g in[g*stride + k] ---------------------------------------------------------------------------
// One thread per output element, rows summed in index order: deterministic for a given G. Reads are coalesced in k.
__global__ void k_reduce_rows(const double *__restrict__ in, int64_t G, int64_t stride, int64_t n, double *__restrict__ out,
                              int accumulate)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int64_t g = 0;
    for (; g + 3 < G; g += 4) {
        s0 += in[g * stride + k];
        s1 += in[(g + 1) * stride + k];
        s2 += in[(g + 2) * stride + k];
        s3 += in[(g + 3) * stride + k];
    }
    for (; g < G; ++g) s0 += in[g * stride + k];
    const double s = (s0 + s1) + (s2 + s3);
    out[k] = accumulate ? out[k] + s : s;
}

// ---- counter-based generator, bit-identical to oracle/va_oracle.c:vo_synth_params ---------------------------------
__host__ __device__ inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ inline double u01(uint64_t seed, uint64_t stream, uint64_t b, uint64_t k)
{
    uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ULL * (stream + 1));
    h = mix64(h ^ (b * 0xD1342543DE82EF95ULL));
    h = mix64(h + k);
    return (double)(h >> 11) * 0x1.0p-53;
}
// explicit _rn intrinsics: no FMA contraction, same bits as the host generator
__device__ inline double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ inline double add(double a, double b) { return __dadd_rn(a, b); }

__global__ void k_synth_small(int system, uint64_t seed, int64_t b0, int64_t B, double *__restrict__ p, double *__restrict__ x0)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint64_t bb = (uint64_t)(b0 + b);
    if (system == VA_SYS_HARMONIC) {
        const double s = add(mul(2.0, u01(seed, 0, bb, 0)), -1.0);
        p[b] = mul(0.151, add(1.0, mul(0.5, s)));
        if (x0) { x0[2 * b] = 0.0; x0[2 * b + 1] = 1.0; }
    } else {
        const double u1 = u01(seed, 1, bb, 0), u2 = u01(seed, 1, bb, 1);
        const double mu = ldexp(add(1.0, u2), (int)mul(10.0, u1));
        p[b] = mu;
        if (x0) {
            x0[2 * b] = 2.0;
            // -2/3 + 10/(81 mu) - 292/(2187 mu mu), left to right as in reference examples/VanDerPol/main.cpp:63
            const double t1 = __ddiv_rn(10.0, mul(81.0, mu));
            const double t2 = __ddiv_rn(292.0, mul(mul(2187.0, mu), mu));
            x0[2 * b + 1] = add(add(-2.0 / 3.0, t1), -t2);
        }
    }
}

// one thread per parameter entry; [B][n*n+n]
__global__ void k_synth_glv(int n, uint64_t seed, int64_t b0, int64_t B, double scale, double sqrt3, double *__restrict__ p,
                            double *__restrict__ x0)
{
    const int64_t npar = (int64_t)n * n + n;
    const int64_t total = B * npar;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / npar;
        const int k = (int)(e - b * npar);
        const uint64_t bb = (uint64_t)(b0 + b);
        double v;
        if (k < n) {
            v = mul(0.1, add(1.0, mul(0.1, add(mul(2.0, u01(seed, 2, bb, (uint64_t)k)), -1.0))));
            if (x0) x0[b * n + k] = 0.1;
        } else {
            const int ij = k - n, i = ij / n, j = ij - i * n;
            const uint64_t base = (uint64_t)n + (uint64_t)ij * 5;
            const double u = u01(seed, 2, bb, base);
            if (i == j) {
                v = mul(-10.0, add(1.0, mul(0.1, add(mul(2.0, u), -1.0))));
            } else if (u < 0.5) {
                const double z = add(add(add(add(u01(seed, 2, bb, base + 1), u01(seed, 2, bb, base + 2)), u01(seed, 2, bb, base + 3)),
                                         u01(seed, 2, bb, base + 4)), -2.0);
                v = mul(mul(z, sqrt3), scale);
            } else {
                v = 0.0;
            }
        }
        p[e] = v;
    }
}

// ---- FP64 peak: independent DFMA chains, 8 per thread ---------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) out[0] = s; // never true; keeps the chains alive
}

__global__ void k_copy(const double4 *__restrict__ in, double4 *__restrict__ out, int64_t n4)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

// Stage 1 of the row reduction for tall inputs: block c sums rows [c*rpb, (c+1)*rpb) into part[c][k]. Thread = (k lane,
// row lane): 32 consecutive k per warp (coalesced), 8 row lanes per block combined in a fixed order -> deterministic.
__global__ void __launch_bounds__(256) k_reduce_rows_stage(const double *__restrict__ in, int64_t G, int64_t stride, int64_t n,
                                                           double *__restrict__ part, int64_t rpb)
{
    __shared__ double sm[8][33];
    const int kl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * rpb, r1 = r0 + rpb < G ? r0 + rpb : G;
    for (int64_t kb = (int64_t)blockIdx.y * 32; kb < n; kb += (int64_t)gridDim.y * 32) {
        const int64_t k = kb + kl;
        double s = 0.0;
        if (k < n)
            for (int64_t r = r0 + rl; r < r1; r += 8) s += in[r * stride + k];
        sm[rl][kl] = s;
        __syncthreads();
        if (rl == 0 && k < n) {
            double t = sm[0][kl];
#pragma unroll
            for (int q = 1; q < 8; ++q) t += sm[q][kl];
            part[(int64_t)blockIdx.x * n + k] = t;
        }
        __syncthreads();
    }
}

} // namespace

cudaError_t va_reduce_rows(const double *in, int64_t G, int64_t stride, int64_t n, double *out, int accumulate, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const int threads = 128;
    if (G > 4096) {
        // tall: two stages through a scratch buffer (kept for the life of the process, grown on demand)
        static double *scratch_dev[16] = {nullptr};
        static size_t scratch_elems_dev[16] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        double *&scratch = scratch_dev[dev & 15];
        size_t &scratch_elems = scratch_elems_dev[dev & 15];
        const int64_t chunks = 1024, rpb = (G + chunks - 1) / chunks;
        const int64_t used = (G + rpb - 1) / rpb;
        if ((size_t)(used * n) > scratch_elems) {
            if (scratch) cudaFree(scratch);
            scratch = nullptr;
            scratch_elems = 0;
            cudaError_t e = cudaMalloc(&scratch, (size_t)(used * n) * 8);
            if (e != cudaSuccess) return e;
            scratch_elems = (size_t)(used * n);
        }
        const unsigned gy = (unsigned)std::min<int64_t>((n + 31) / 32, 64);
        k_reduce_rows_stage<<<dim3((unsigned)used, gy), 256, 0, st>>>(in, G, stride, n, scratch, rpb);
        k_reduce_rows<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(scratch, used, n, n, out, accumulate);
        return cudaGetLastError();
    }
    k_reduce_rows<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(in, G, stride, n, out, accumulate);
    return cudaGetLastError();
}

cudaError_t va_synth_launch(int system, int n, uint64_t seed, int64_t b0, int64_t B, double *params, double *x0, cudaStream_t st)
{
    if (B <= 0) return cudaSuccess;
    if (system == VA_SYS_GLV) {
        const double scale = sqrt(10.0 / (double)n), sqrt3 = sqrt(3.0);
        k_synth_glv<<<148 * 16, 256, 0, st>>>(n, seed, b0, B, scale, sqrt3, params, x0);
    } else if (system == VA_SYS_HARMONIC || system == VA_SYS_VANDERPOL) {
        k_synth_small<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(system, seed, b0, B, params, x0);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

extern "C" int va_measure_fp64_peak(int32_t device, double *tflops)
{
    if (cudaSetDevice(device) != cudaSuccess) return VA_E_CUDA;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    double *out = nullptr;
    if (cudaMalloc(&out, 8) != cudaSuccess) return VA_E_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, blocks = sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_dfma_peak<<<blocks, threads>>>(out, iters, 1.0 + rep);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return VA_E_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return VA_OK;
}

extern "C" int va_measure_hbm_copy(int32_t device, double *gbytes_per_s)
{
    if (cudaSetDevice(device) != cudaSuccess) return VA_E_CUDA;
    const int64_t bytes = (int64_t)2 << 30; // 2 GiB each way, far beyond L2
    double4 *src = nullptr, *dst = nullptr;
    if (cudaMalloc(&src, bytes) != cudaSuccess) return VA_E_NOMEM;
    if (cudaMalloc(&dst, bytes) != cudaSuccess) { cudaFree(src); return VA_E_NOMEM; }
    cudaMemset(src, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_copy<<<148 * 16, 512>>>(src, dst, bytes / 32);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gbs = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && gbs > best) best = gbs;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(src);
    cudaFree(dst);
    *gbytes_per_s = best;
    return VA_OK;
}
