// va_glv_stream.cu -- Generalized Lotka-Volterra for ANY number of species: the interaction matrix does not fit one SM's
// register file (N = 256: 512 KB), so it is streamed from L2/HBM for every matrix-vector product; vectors live in shared
// memory and the gradient accumulator Abar in global memory (read-modify-write, coalesced). One persistent 256-thread
// CTA per trajectory, forward then reverse, same store-stages checkpoint blocks as the register kernel.
//
// Same algorithm as va_glv_wide.cu (reference lib/include/detail/runge_kutta.hpp:76-118 forward,
// detail/backpropagation.hpp:83-158 reverse, odeint controller), different data placement:
//   y = A x      : one warp per row, lanes stride over the columns (coalesced 8 B loads), shuffle reduction;
//   A^T v, Abar  : one thread per column j, loop over rows i: A[i][j] coalesced across the CTA, and in the same pass
//                  Abar[i][j] += v_i x_j (Abar is the caller's dJ/dalpha row of this trajectory, or the CTA's partial-sum
//                  row in summed mode).
// Bytes per accepted step: 6 * 8 N^2 (forward reads of A) + 6 * 24 N^2 (reverse: A read, Abar read+write) = 192 N^2 B,
// i.e. 12.6 MB at N = 256 -> this family is L2/HBM-bandwidth bound by construction (0.25 flop/B). It exists so that
// every N is served with the reference's results; the cluster/DSMEM kernel that keeps A distributed over several SMs is
// the next step for N = 256.
#include "va_glv_common.cuh"

namespace {

constexpr int SNT = 256; // threads per CTA

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// RECOMPUTE selects the checkpoint policy (north_star item 4): false = STORE_STAGES (stage states and slopes of every
// accepted step are kept: 16 s N B per step), true = the reference's policy (only (t_n, x_n) is kept: 8 N B per step, the
// stages are recomputed in the reverse sweep with s extra matrix-vector products, detail/backpropagation.hpp:24-64).
template <class Tab, bool ADAPTIVE, bool RECOMPUTE>
__global__ void __launch_bounds__(SNT) k_glv_stream(const __grid_constant__ VaGlvWideArgs a)
{
    constexpr int S = Tab::S, SADJ = Tab::SADJ;
    constexpr int SE = Tab::FSAL ? S - 1 : S;
    extern __shared__ double sm[];
    const int n = a.n, npar = n * n + n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int BLK = RECOMPUTE ? 8 + n : 8 + 2 * SADJ * n;
    // shared vectors: x, xs (stage state), K[S], W[SADJ+1], v, gx, red
    double *x = sm, *xs = x + n, *K = xs + n, *W = K + S * n, *v = W + (SADJ + 1) * n, *gx = v + n, *g0 = gx + n, *red = g0 + n;
    double *rXG = red + 64; // RECOMPUTE: recomputed stage states / slopes of the current step, [2][SADJ][n]
    __shared__ double s_scalar[4];
    double *const slab = a.slab + (int64_t)blockIdx.x * a.slab_stride;

    for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
        const double *pb = a.params + b * npar;
        const double *A = pb + n;
        // g_out[i] = r_i + (A xin)_i
        auto matvec = [&](const double *xin, double *g_out) {
            __syncthreads();
            for (int i = warp; i < n; i += SNT / 32) {
                const double *row = A + (int64_t)i * n;
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) acc = fma(__ldg(row + j), xin[j], acc);
                acc = warp_sum(acc);
                if (lane == 0) g_out[i] = __ldg(pb + i) + acc;
            }
            __syncthreads();
        };
        for (int i = tid; i < n; i += SNT) x[i] = a.x0[b * n + i];
        double t = a.ti, dt = a.dt0;
        const double tf = a.tf;
        int nck = 0, rejects = 0, status = 0, trials = 0;
        matvec(x, g0);
        for (int i = tid; i < n; i += SNT) K[i] = x[i] * g0[i];
        bool active = ADAPTIVE ? va_less_with_sign(t, tf, dt) : va_less_eq_with_sign(t + dt, tf, dt);
        bool fresh = true;
        while (active) {
            double *blk = slab + (int64_t)nck * BLK;
            if (fresh) {
                if (nck >= a.cap) { status |= VA_TRAJ_CKPT_OVERFLOW; break; }
                __syncthreads();
                for (int i = tid; i < n; i += SNT) {
                    blk[8 + i] = x[i];
                    if (!RECOMPUTE) blk[8 + SADJ * n + i] = g0[i];
                }
                if (tid == 0) blk[0] = t;
                if (ADAPTIVE && va_less_with_sign(tf, t + dt, dt)) dt = tf - t;
                trials = 0;
                fresh = false;
            }
#pragma unroll
            for (int m = 1; m < SE; ++m) {
                __syncthreads();
                for (int i = tid; i < n; i += SNT) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < m; ++j)
                        if (Tab::a(m, j) != 0.0) acc = fma(Tab::a(m, j), K[j * n + i], acc);
                    xs[i] = fma(dt, acc, x[i]);
                }
                matvec(xs, gx);
                for (int i = tid; i < n; i += SNT) {
                    K[m * n + i] = xs[i] * gx[i];
                    if (!RECOMPUTE && m < SADJ) { blk[8 + m * n + i] = xs[i]; blk[8 + (SADJ + m) * n + i] = gx[i]; }
                }
            }
            __syncthreads();
            for (int i = tid; i < n; i += SNT) { // xs <- xnew
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < SE; ++j)
                    if (Tab::b(j) != 0.0) acc = fma(Tab::b(j), K[j * n + i], acc);
                xs[i] = fma(dt, acc, x[i]);
            }
            if (Tab::FSAL) {
                matvec(xs, gx);
                for (int i = tid; i < n; i += SNT) K[(S - 1) * n + i] = xs[i] * gx[i];
            }
            bool accept = true;
            double err = 0.0;
            if (ADAPTIVE) {
                __syncthreads();
                double e = 0.0;
                for (int i = tid; i < n; i += SNT) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < S; ++j)
                        if (Tab::db(j) != 0.0) acc = fma(Tab::db(j), K[j * n + i], acc);
                    e = fmax(e, fabs(dt * acc) / (a.eps_abs + a.eps_rel * (fabs(x[i]) + fabs(dt) * fabs(K[i]))));
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, d));
                if (lane == 0) red[warp] = e;
                __syncthreads();
#pragma unroll
                for (int w = 0; w < SNT / 32; ++w) err = fmax(err, red[w]);
                accept = !(err > 1.0);
            }
            if (!accept) {
                dt *= fmax(0.9 * inv_root<(Tab::ERROR_ORDER > 1 ? Tab::ERROR_ORDER - 1 : 1)>(err), 0.2);
                ++rejects;
                if (++trials >= 500) { status |= VA_TRAJ_NO_PROGRESS; break; }
            } else {
                __syncthreads();
                for (int i = tid; i < n; i += SNT) x[i] = xs[i];
                ++nck;
                if (ADAPTIVE) {
                    t += dt;
                    if (err < 0.5) {
                        constexpr int P = Tab::STEPPER_ORDER;
                        double floor_ = 1.0;
#pragma unroll
                        for (int k = 0; k < P; ++k) floor_ *= 0.2;
                        dt *= (err <= floor_) ? 4.5 : 9.0 / 10.0 * inv_root<P>(err);
                    }
                    active = va_less_with_sign(t, tf, dt);
                } else {
                    t = a.ti + (double)nck * dt;
                    active = va_less_eq_with_sign(t + dt, tf, dt);
                }
                fresh = true;
                if (Tab::FSAL) {
                    __syncthreads();
                    for (int i = tid; i < n; i += SNT) { g0[i] = gx[i]; K[i] = K[(S - 1) * n + i]; }
                } else if (active) {
                    matvec(x, g0);
                    for (int i = tid; i < n; i += SNT) K[i] = x[i] * g0[i];
                }
            }
        }
        __syncthreads();
        const int T = nck;
        if (tid == 0) slab[(int64_t)T * BLK] = t;
        int bad = 0;
        for (int i = tid; i < n; i += SNT) bad |= !isfinite(x[i]);
        if (bad) status |= VA_TRAJ_NONFINITE;
        status = __syncthreads_or(status);
        const bool failed = status & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS);
        for (int i = tid; i < n; i += SNT) a.x_final[b * n + i] = failed ? nan("") : x[i];
        if (tid == 0) {
            if (a.n_accept) a.n_accept[b] = T;
            if (a.n_reject) a.n_reject[b] = rejects;
            if (a.status) a.status[b] = status;
        }
        const double t_final = t;

        // ------------------------------------------ reverse sweep ------------------------------------------------------
        for (int o = 0; o < a.n_out; ++o) {
            double *lam_io = a.lambda + (b * a.n_out + o) * n;
            const bool sum_mode = a.reduce == VA_REDUCE_SUM;
            double *gbar = sum_mode ? a.partial + (int64_t)blockIdx.x * npar : a.mu + (b * a.n_out + o) * npar;
            if (failed) {
                for (int i = tid; i < n; i += SNT) lam_io[i] = nan("");
                if (!sum_mode)
                    for (int k = tid; k < npar; k += SNT) gbar[k] = nan("");
                continue;
            }
            if (!sum_mode || (b == blockIdx.x && o == 0))
                for (int k = tid; k < npar; k += SNT) gbar[k] = 0.0; // first use of this accumulator
            double *lam = xs; // reuse
            __syncthreads();
            for (int i = tid; i < n; i += SNT)
                lam[i] = a.objective == VA_OBJ_SUM ? 1.0 : a.objective == VA_OBJ_HALF_NORM2 ? x[i] : lam_io[i];
            double t_hi = t_final;
            for (int step = T - 1; step >= 0; --step) {
                const double *blk = slab + (int64_t)step * BLK;
                const double dt_s = t_hi - blk[0];
                t_hi = blk[0];
                if (RECOMPUTE) {
                    // stage recompute from x_n with dt = t_{n+1} - t_n: X_m = x_n + dt sum_j a_mj K_j, g_m = r + A X_m
                    double *rX = rXG, *rG = rXG + SADJ * n;
#pragma unroll
                    for (int m = 0; m < SADJ; ++m) {
                        __syncthreads();
                        for (int i = tid; i < n; i += SNT) {
                            double acc = 0.0;
#pragma unroll
                            for (int j = 0; j < m; ++j)
                                if (Tab::a(m, j) != 0.0) acc = fma(Tab::a(m, j), K[j * n + i], acc);
                            rX[m * n + i] = fma(dt_s, acc, blk[8 + i]);
                        }
                        matvec(rX + m * n, rG + m * n);
                        for (int i = tid; i < n; i += SNT) K[m * n + i] = rX[m * n + i] * rG[m * n + i];
                    }
                }
                __syncthreads();
                for (int i = tid; i < n; i += SNT) {
                    W[i] = lam[i];
#pragma unroll
                    for (int m = 1; m <= SADJ; ++m) W[m * n + i] = (Tab::b(m - 1) * dt_s) * lam[i];
                }
#pragma unroll
                for (int m = SADJ; m >= 1; --m) {
                    const double *X = RECOMPUTE ? rXG + (m - 1) * n : blk + 8 + (m - 1) * n;
                    const double *G = RECOMPUTE ? rXG + (SADJ + m - 1) * n : blk + 8 + (SADJ + m - 1) * n;
                    __syncthreads();
                    for (int i = tid; i < n; i += SNT) v[i] = W[m * n + i] * X[i];
                    __syncthreads();
                    // thread per column j: (A^T v)_j and Abar[:, j] += v x_j in one pass over the rows
                    for (int j = tid; j < n; j += SNT) {
                        const double xj = X[j];
                        double acc = 0.0;
                        double *gA = gbar + n + j;
                        const double *Aj = A + j;
#pragma unroll 4
                        for (int i = 0; i < n; ++i) {
                            const double vi = v[i];
                            acc = fma(__ldg(Aj + (int64_t)i * n), vi, acc);
                            gA[(int64_t)i * n] = fma(vi, xj, gA[(int64_t)i * n]);
                        }
                        const double gxj = fma(W[m * n + j], G[j], acc);
                        gbar[j] += v[j];
                        W[j] += gxj;
#pragma unroll
                        for (int k = 1; k < m; ++k)
                            if (Tab::a(m - 1, k - 1) != 0.0) W[k * n + j] = fma(gxj * Tab::a(m - 1, k - 1), dt_s, W[k * n + j]);
                    }
                }
                __syncthreads();
                for (int i = tid; i < n; i += SNT) lam[i] = W[i];
            }
            __syncthreads();
            for (int i = tid; i < n; i += SNT) lam_io[i] = lam[i];
        }
        __syncthreads();
    }
    (void)s_scalar;
}

template <class Tab, bool ADAPTIVE, bool RECOMPUTE>
cudaError_t launch2(const VaGlvWideArgs &a, cudaStream_t st, size_t smem)
{
    cudaError_t e = cudaFuncSetAttribute(k_glv_stream<Tab, ADAPTIVE, RECOMPUTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_glv_stream<Tab, ADAPTIVE, RECOMPUTE><<<a.grid, SNT, smem, st>>>(a);
    return cudaGetLastError();
}
template <class Tab, bool ADAPTIVE>
cudaError_t launch(const VaGlvWideArgs &a, cudaStream_t st, size_t smem)
{
    return a.recompute ? launch2<Tab, ADAPTIVE, true>(a, st, smem) : launch2<Tab, ADAPTIVE, false>(a, st, smem);
}

int stages_of(int stepper, int *sadj)
{
    switch (stepper) {
    case VA_RK_EULER: *sadj = TabEuler::SADJ; return TabEuler::S;
    case VA_RK_RK4: *sadj = TabRK4::SADJ; return TabRK4::S;
    case VA_RK_CK54: *sadj = TabCK54::SADJ; return TabCK54::S;
    case VA_RK_DOPRI5: *sadj = TabDOPRI5::SADJ; return TabDOPRI5::S;
    case VA_RK_RKF78: *sadj = TabRKF78::SADJ; return TabRKF78::S;
    }
    *sadj = 0;
    return 0;
}

} // namespace

size_t va_glv_stream_smem(int n, int stepper, int recompute)
{
    int sadj = 0;
    const int s = stages_of(stepper, &sadj);
    return (size_t)(6 + s + sadj + 1 + (recompute ? 2 * sadj : 0)) * n * 8 + 64 * 8 + 64 * 8;
}

bool va_glv_stream_supported(int n, int stepper, int adaptive)
{
    if (n < 1) return false;
    if (va_glv_stream_smem(n, stepper, 1) > 200 * 1024) return false; // vectors must fit shared memory (cash_karp54: N up to ~800)
    // every tableau the reference's ButcherTable knows (ButcherTable.hpp:50-246) + dopri5; error steppers run controlled
    // (make_controlled<...>) or fixed-step (the stepper_tag overload, detail/runge_kutta.hpp:38-72, which ignores the estimate)
    if (stepper == VA_RK_EULER || stepper == VA_RK_RK4) return !adaptive;
    return stepper == VA_RK_CK54 || stepper == VA_RK_DOPRI5 || stepper == VA_RK_RKF78;
}

int va_glv_stream_block_doubles(int n, int stepper, int recompute)
{
    int sadj = 0;
    stages_of(stepper, &sadj);
    return recompute ? 8 + n : 8 + 2 * sadj * n;
}

cudaError_t va_glv_stream_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const size_t smem = va_glv_stream_smem(a.n, a.stepper, a.recompute);
    switch (a.stepper) {
    case VA_RK_EULER: return launch<TabEuler, false>(a, st, smem);
    case VA_RK_RK4: return launch<TabRK4, false>(a, st, smem);
    case VA_RK_CK54: return a.adaptive ? launch<TabCK54, true>(a, st, smem) : launch<TabCK54, false>(a, st, smem);
    case VA_RK_DOPRI5: return a.adaptive ? launch<TabDOPRI5, true>(a, st, smem) : launch<TabDOPRI5, false>(a, st, smem);
    case VA_RK_RKF78: return a.adaptive ? launch<TabRKF78, true>(a, st, smem) : launch<TabRKF78, false>(a, st, smem);
    }
    return cudaErrorInvalidValue;
}
