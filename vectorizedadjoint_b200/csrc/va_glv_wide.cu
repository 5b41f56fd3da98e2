// va_glv_wide.cu -- CTA-per-trajectory forward + discrete-adjoint kernel for the Generalized Lotka-Volterra system
//                    with up to 64 species (the headline path: GLV N = 64, one million parameter sets).
//
// f_i = x_i (r_i + (A x)_i), parameters [r, A row-major] (reference examples/GeneralizedLotkaVolterra/main.cpp:105-119).
// The reference evaluates f and its vector-Jacobian products through AADC-recorded AVX kernels (lib/include/AadData.hpp
// :175-210, :291-330), feeding all N^2+N parameters into the workspace for every call. Here the interaction matrix of a
// trajectory is loaded ONCE into the register file of a 256-thread CTA and stays there:
//
//   forward  (reference lib/include/detail/runge_kutta.hpp:76-118 + odeint controlled stepper):
//       thread (i, q) owns a quarter of row i of A (16 FP64 registers); A x is 16 DFMA + two shuffle-adds per thread.
//   backward (reference lib/include/detail/backpropagation.hpp:83-158, 231-254):
//       thread (j, q) owns a quarter of COLUMN j of A and of the gradient accumulator Abar (16 + 16 registers);
//       A^T v and the rank-1 update Abar += v x^T share the same 16 broadcast loads of v; nothing but v crosses warps.
//
// A persistent CTA integrates trajectory after trajectory (static stride over the batch), forward then backward, and
// keeps its checkpoints in a private slab that is reused for every trajectory and therefore stays in L2.
// Checkpoint policy: STORE_STAGES -- for every accepted step the stage states X_m and g_m = r + A X_m are kept, so the
// reverse sweep needs no stage recompute (the reference recomputes them with s extra RHS calls per step,
// detail/backpropagation.hpp:24-64); per step and trajectory that is 2*s*64*8 B of L2-resident traffic against
// 6*4096 saved DFMA.
//
// Accept/reject logic is odeint's (same error norm, same step-size rules). The matrix-vector products use FMA and a
// tree reduction, so stage values differ from the scalar reference by round-off; accepted-step counts can therefore
// flip only when the error estimate is within ~1e-8 relative of a threshold (documented in DESIGN.md).
#include "va_common.cuh"

namespace {

constexpr int NP = 64;   // padded species count
constexpr int NT = 256;  // threads per CTA: 4 per row (forward) / per column (backward)

// ---- compile-time tableaux: zero weights vanish from the unrolled code --------------------------------------------
struct TabRK4 {
    static constexpr int S = 4, SADJ = 4, STEPPER_ORDER = 4, ERROR_ORDER = 0;
    static constexpr bool FSAL = false, HAS_ERR = false;
    __host__ __device__ static constexpr double a(int m, int j)
    {
        return (m == 1 && j == 0) ? 0.5 : (m == 2 && j == 1) ? 0.5 : (m == 3 && j == 2) ? 1.0 : 0.0;
    }
    __host__ __device__ static constexpr double b(int j) { return (j == 0 || j == 3) ? 1.0 / 6 : 1.0 / 3; }
    __host__ __device__ static constexpr double db(int) { return 0.0; }
};
struct TabCK54 {
    static constexpr int S = 6, SADJ = 6, STEPPER_ORDER = 5, ERROR_ORDER = 4;
    static constexpr bool FSAL = false, HAS_ERR = true;
    __host__ __device__ static constexpr double a(int m, int j)
    {
        constexpr double t[6][5] = {{0, 0, 0, 0, 0},
                                    {1.0 / 5, 0, 0, 0, 0},
                                    {3.0 / 40, 9.0 / 40, 0, 0, 0},
                                    {3.0 / 10, -9.0 / 10, 6.0 / 5, 0, 0},
                                    {-11.0 / 54, 5.0 / 2, -70.0 / 27, 35.0 / 27, 0},
                                    {1631.0 / 55296, 175.0 / 512, 575.0 / 13824, 44275.0 / 110592, 253.0 / 4096}};
        return t[m][j];
    }
    __host__ __device__ static constexpr double b(int j)
    {
        constexpr double t[6] = {37.0 / 378, 0, 250.0 / 621, 125.0 / 594, 0, 512.0 / 1771};
        return t[j];
    }
    __host__ __device__ static constexpr double db(int j)
    {
        constexpr double t[6] = {37.0 / 378 - 2825.0 / 27648, 0, 250.0 / 621 - 18575.0 / 48384, 125.0 / 594 - 13525.0 / 55296,
                                 0.0 - 277.0 / 14336, 512.0 / 1771 - 1.0 / 4};
        return t[j];
    }
};
struct TabDOPRI5 {
    static constexpr int S = 7, SADJ = 6, STEPPER_ORDER = 5, ERROR_ORDER = 4;
    static constexpr bool FSAL = true, HAS_ERR = true;
    __host__ __device__ static constexpr double a(int m, int j)
    {
        constexpr double t[7][6] = {{0, 0, 0, 0, 0, 0},
                                    {1.0 / 5, 0, 0, 0, 0, 0},
                                    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
                                    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
                                    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
                                    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
                                    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
        return t[m][j];
    }
    __host__ __device__ static constexpr double b(int j)
    {
        constexpr double t[7] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 0};
        return t[j];
    }
    __host__ __device__ static constexpr double db(int j)
    {
        constexpr double t[7] = {35.0 / 384 - 5179.0 / 57600, 0, 500.0 / 1113 - 7571.0 / 16695, 125.0 / 192 - 393.0 / 640,
                                 -2187.0 / 6784 - (-92097.0 / 339200), 11.0 / 84 - 187.0 / 2100, -1.0 / 40};
        return t[j];
    }
};

// e^(-1/P) for e > 0: float seed + Newton on y^-P = e (quadratic), accurate to a few ulp; replaces pow() in
// odeint's default_step_adjuster on this path (all 256 threads evaluate it redundantly, so it has to be short).
template <int P>
__device__ __forceinline__ double inv_root(double e)
{
    if (e > 1e30) return 0.0;
    double y = (double)__powf((float)e, -1.0f / (float)P);
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        double yp = y;
#pragma unroll
        for (int k = 1; k < P; ++k) yp *= y;
        y = fma(y * (1.0 / P), fma(-e, yp, 1.0), y);
    }
    return y;
}

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

struct SlabView {
    double *tt; // [cap+2]
    double *sx; // [cap*SADJ][NP] stage states X_m of accepted steps
    double *sg; // [cap*SADJ][NP] g_m = r + A X_m
};

template <class Tab, bool ADAPTIVE, bool EXACT64>
__global__ void __launch_bounds__(NT, 2) k_glv_wide(const __grid_constant__ VaGlvWideArgs a)
{
    constexpr int S = Tab::S, SADJ = Tab::SADJ;
    constexpr int SE = Tab::FSAL ? S - 1 : S; // stages evaluated through an intermediate state
    __shared__ __align__(16) double xs[2][NP];
    __shared__ double red[8];

    const int tid = threadIdx.x;
    const int q = tid & 3;     // quarter
    const int rc = tid >> 2;   // row (forward) / column (backward) owned by this thread
    const int lane = tid & 31, warp = tid >> 5;
    const int n = a.n;
    const int npar = n * n + n;
    const int cap = a.cap;

    SlabView sl;
    {
        double *base = a.slab + (int64_t)blockIdx.x * a.slab_stride;
        const int tt_len = (cap + 2 + 15) & ~15;
        sl.tt = base;
        sl.sx = base + tt_len;
        sl.sg = sl.sx + (int64_t)cap * SADJ * NP;
    }

    // gradient accumulators (backward layout); persistent across trajectories when the caller wants the sum
    double Abar[16], rbar = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) Abar[k] = 0.0;

    int call = 0; // parity of the xs double buffer

    for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
        const double *pb = a.params + b * npar;

        // ================================ forward sweep =====================================
        double Ar[16];
        double r_i = 0.0, x = 0.0;
        {
            const int i = rc;
            if (EXACT64) {
                const double2 *row = reinterpret_cast<const double2 *>(pb + NP + i * NP);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double2 v = __ldg(row + 4 * k + q);
                    Ar[2 * k] = v.x;
                    Ar[2 * k + 1] = v.y;
                }
                r_i = __ldg(pb + i);
                x = __ldg(a.x0 + b * NP + i);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = 8 * k + 2 * q + e;
                        Ar[2 * k + e] = (i < n && col < n) ? __ldg(pb + n + i * n + col) : 0.0;
                    }
                }
                if (i < n) { r_i = __ldg(pb + i); x = __ldg(a.x0 + b * n + i); }
            }
        }

        // g = r_i + (A X)_i for the row of this thread; X is this row's entry of the stage state
        auto matvec = [&](double X) -> double {
            double *buf = xs[call & 1];
            ++call;
            if (q == 0) buf[rc] = X;
            __syncthreads();
            const double2 *xv = reinterpret_cast<const double2 *>(buf);
            double acc0 = (q == 0) ? r_i : 0.0, acc1 = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double2 xx = xv[4 * k + q];
                acc0 = fma(Ar[2 * k], xx.x, acc0);
                acc1 = fma(Ar[2 * k + 1], xx.y, acc1);
            }
            double g = acc0 + acc1;
            g += shfl_xor_d(g, 1);
            g += shfl_xor_d(g, 2);
            return g;
        };

        double t = a.ti, dt = a.dt0;
        const double tf = a.tf;
        int nck = 0, rejects = 0, status = 0;
        double K[S];
        double g0 = matvec(x);
        K[0] = x * g0;

        auto store_stage = [&](int m, double X, double g) {
            if (q == 0) {
                const int64_t o = ((int64_t)nck * SADJ + m) * NP + rc;
                sl.sx[o] = X;
                sl.sg[o] = g;
            }
        };

        bool active = ADAPTIVE ? va_less_with_sign(t, tf, dt) : va_less_eq_with_sign(t + dt, tf, dt);
        bool fresh = true;
        int trials = 0;
        while (active) {
            if (fresh) {
                if (nck >= cap) { status |= VA_TRAJ_CKPT_OVERFLOW; break; }
                store_stage(0, x, g0);
                if (tid == 0) sl.tt[nck] = t;
                if (ADAPTIVE && va_less_with_sign(tf, t + dt, dt)) dt = tf - t;
                trials = 0;
                fresh = false;
            }
            // stages
#pragma unroll
            for (int m = 1; m < SE; ++m) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < m; ++j)
                    if (Tab::a(m, j) != 0.0) acc = fma(Tab::a(m, j), K[j], acc);
                const double X = fma(dt, acc, x);
                const double g = matvec(X);
                K[m] = X * g;
                if (m < SADJ) store_stage(m, X, g);
            }
            double xnew;
            {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < SE; ++j)
                    if (Tab::b(j) != 0.0) acc = fma(Tab::b(j), K[j], acc);
                xnew = fma(dt, acc, x);
            }
            double g_last = 0.0;
            if (Tab::FSAL) {
                g_last = matvec(xnew);
                K[S - 1] = xnew * g_last;
            }
            bool accept = true;
            double err = 0.0;
            if (ADAPTIVE) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < S; ++j)
                    if (Tab::db(j) != 0.0) acc = fma(Tab::db(j), K[j], acc);
                const double xerr = dt * acc;
                // default_error_checker::error, max norm over species
                double e = fabs(xerr) / (a.eps_abs + a.eps_rel * (fabs(x) + fabs(dt) * fabs(K[0])));
                e = fmax(e, shfl_xor_d(e, 4));
                e = fmax(e, shfl_xor_d(e, 8));
                e = fmax(e, shfl_xor_d(e, 16));
                if (lane == 0) red[warp] = e;
                __syncthreads();
#pragma unroll
                for (int w = 0; w < 8; ++w) err = fmax(err, red[w]);
                accept = !(err > 1.0);
            }
            if (!accept) {
                // default_step_adjuster::decrease_step
                dt *= fmax(0.9 * inv_root<(Tab::ERROR_ORDER > 1 ? Tab::ERROR_ORDER - 1 : 1)>(err), 0.2);
                ++rejects;
                if (++trials >= 500) { status |= VA_TRAJ_NO_PROGRESS; break; }
            } else {
                x = xnew;
                ++nck;
                if (ADAPTIVE) {
                    t += dt;
                    // default_step_adjuster::increase_step
                    if (err < 0.5) {
                        constexpr int P = Tab::STEPPER_ORDER;
                        double floor_ = 1.0;
#pragma unroll
                        for (int k = 0; k < P; ++k) floor_ *= 0.2; // 5^-P
                        err = fmax(floor_, err);
                        dt *= 9.0 / 10.0 * inv_root<P>(err);
                    }
                    active = va_less_with_sign(t, tf, dt);
                } else {
                    t = a.ti + (double)nck * dt; // detail/runge_kutta.hpp:64
                    active = va_less_eq_with_sign(t + dt, tf, dt);
                }
                fresh = true;
                if (Tab::FSAL) {
                    g0 = g_last;
                    K[0] = K[S - 1];
                } else if (active) {
                    g0 = matvec(x);
                    K[0] = x * g0;
                }
            }
        }
        const int T = nck;
        if (tid == 0) sl.tt[T] = t;
        if (!isfinite(x)) status |= VA_TRAJ_NONFINITE;
        status = __syncthreads_or(status); // also orders the slab writes before the reverse sweep reads them
        const bool failed = status & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS);
        if (q == 0 && rc < n) a.x_final[b * n + rc] = failed ? nan("") : x;
        if (tid == 0) {
            if (a.n_accept) a.n_accept[b] = T;
            if (a.n_reject) a.n_reject[b] = rejects;
            if (a.status) a.status[b] = status;
        }
        // x(tf) by column for the seeds: exchange through shared memory (row owner -> column owner is the same index)
        const double x_tf = x; // thread (rc, q): forward row rc == backward column rc

        // ================================ reverse sweep =====================================
        const int j = rc;
        double Ac[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = 8 * k + 2 * q + e;
                if (EXACT64) Ac[2 * k + e] = __ldg(pb + NP + row * NP + j);
                else Ac[2 * k + e] = (row < n && j < n) ? __ldg(pb + n + row * n + j) : 0.0;
            }
        }

        for (int o = 0; o < a.n_out; ++o) {
            double *lam_io = a.lambda + (b * a.n_out + o) * n;
            double *mu_o = a.mu + (a.reduce == VA_REDUCE_SUM ? (int64_t)o : (b * a.n_out + o)) * npar;
            if (failed) {
                if (q == 0 && j < n) lam_io[j] = nan("");
                if (a.reduce == VA_REDUCE_NONE)
                    for (int k = tid; k < npar; k += NT) mu_o[k] = nan("");
                continue;
            }
            double lam;
            if (a.objective == VA_OBJ_SUM) lam = (j < n) ? 1.0 : 0.0;
            else if (a.objective == VA_OBJ_HALF_NORM2) lam = x_tf;
            else lam = (j < n) ? lam_io[j] : 0.0;

            if (a.reduce == VA_REDUCE_NONE) {
#pragma unroll
                for (int k = 0; k < 16; ++k) Abar[k] = 0.0;
                rbar = 0.0;
            }

            // flattened (step, stage) index, descending; two-deep register prefetch of X and g from the slab
            int idx = T * SADJ - 1;
            double Xn0 = 0, gn0 = 0, Xn1 = 0, gn1 = 0;
            if (idx >= 0) { Xn0 = sl.sx[(int64_t)idx * NP + j]; gn0 = sl.sg[(int64_t)idx * NP + j]; }
            if (idx >= 1) { Xn1 = sl.sx[(int64_t)(idx - 1) * NP + j]; gn1 = sl.sg[(int64_t)(idx - 1) * NP + j]; }
            double t_hi = (T > 0) ? sl.tt[T] : 0.0;
            for (int step = T - 1; step >= 0; --step) {
                const double t_lo = sl.tt[step];
                const double dt_s = t_hi - t_lo; // StateStorage::GetDt
                t_hi = t_lo;
                double W[SADJ + 1];
                W[0] = lam;
#pragma unroll
                for (int m = 1; m <= SADJ; ++m) W[m] = (Tab::b(m - 1) * dt_s) * lam;
#pragma unroll
                for (int m = SADJ; m >= 1; --m) {
                    const double X = Xn0, g = gn0;
                    Xn0 = Xn1; gn0 = gn1;
                    if (idx >= 2) {
                        Xn1 = sl.sx[(int64_t)(idx - 2) * NP + j];
                        gn1 = sl.sg[(int64_t)(idx - 2) * NP + j];
                    }
                    --idx;
                    const bool live = (m == SADJ) || true; // every stage runs a VJP (detail/backpropagation.hpp:201-223)
                    (void)live;
                    const double v = W[m] * X;
                    double *buf = xs[call & 1];
                    ++call;
                    if (q == 0) buf[j] = v;
                    __syncthreads();
                    const double2 *vv2 = reinterpret_cast<const double2 *>(buf);
                    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const double2 vv = vv2[4 * k + q];
                        acc0 = fma(Ac[2 * k], vv.x, acc0);
                        acc1 = fma(Ac[2 * k + 1], vv.y, acc1);
                        Abar[2 * k] = fma(vv.x, X, Abar[2 * k]);
                        Abar[2 * k + 1] = fma(vv.y, X, Abar[2 * k + 1]);
                    }
                    double sum = acc0 + acc1;
                    sum += shfl_xor_d(sum, 1);
                    sum += shfl_xor_d(sum, 2);
                    const double gx = fma(W[m], g, sum);
                    rbar += v;
                    W[0] += gx;
#pragma unroll
                    for (int k = 1; k < m; ++k)
                        if (Tab::a(m - 1, k - 1) != 0.0) W[k] = fma(gx * Tab::a(m - 1, k - 1), dt_s, W[k]);
                }
                lam = W[0];
            }
            if (q == 0 && j < n) lam_io[j] = lam;
            if (a.reduce == VA_REDUCE_NONE) {
                if (q == 0 && j < n) mu_o[j] = rbar;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int row = 8 * k + 2 * q + e;
                        if (row < n && j < n) mu_o[n + row * n + j] = Abar[2 * k + e];
                    }
                }
            }
        }
        __syncthreads(); // slab is reused by the next trajectory
    }

    if (a.reduce == VA_REDUCE_SUM) {
        double *part = a.partial + (int64_t)blockIdx.x * npar;
        const int j = rc;
        if (q == 0 && j < n) part[j] = rbar;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = 8 * k + 2 * q + e;
                if (row < n && j < n) part[n + row * n + j] = Abar[2 * k + e];
            }
        }
    }
}

template <class Tab, bool ADAPTIVE>
cudaError_t launch(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.n == NP) k_glv_wide<Tab, ADAPTIVE, true><<<a.grid, NT, 0, st>>>(a);
    else k_glv_wide<Tab, ADAPTIVE, false><<<a.grid, NT, 0, st>>>(a);
    return cudaGetLastError();
}

template <class Tab, bool ADAPTIVE>
cudaError_t occupancy(int n, int *ctas_per_sm)
{
    if (n == NP) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_glv_wide<Tab, ADAPTIVE, true>, NT, 0);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_glv_wide<Tab, ADAPTIVE, false>, NT, 0);
}

int sadj_of(int stepper)
{
    switch (stepper) {
    case VA_RK_RK4: return TabRK4::SADJ;
    case VA_RK_CK54: return TabCK54::SADJ;
    case VA_RK_DOPRI5: return TabDOPRI5::SADJ;
    }
    return 0;
}

} // namespace

bool va_glv_wide_supported(int n, int stepper, int adaptive)
{
    if (n < 1 || n > NP) return false;
    if (stepper == VA_RK_RK4) return !adaptive;
    if (stepper == VA_RK_CK54 || stepper == VA_RK_DOPRI5) return adaptive != 0;
    return false;
}

int64_t va_glv_wide_slab_doubles(int n, int stepper, int cap)
{
    (void)n;
    const int64_t tt_len = (cap + 2 + 15) & ~15;
    return tt_len + 2 * (int64_t)cap * sadj_of(stepper) * NP;
}

cudaError_t va_glv_wide_config(int n, int stepper, int device, int *grid, int *ctas_per_sm, int *threads)
{
    int sms = 0;
    cudaError_t err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err != cudaSuccess) return err;
    int occ = 0;
    switch (stepper) {
    case VA_RK_RK4: err = occupancy<TabRK4, false>(n, &occ); break;
    case VA_RK_CK54: err = occupancy<TabCK54, true>(n, &occ); break;
    case VA_RK_DOPRI5: err = occupancy<TabDOPRI5, true>(n, &occ); break;
    default: return cudaErrorInvalidValue;
    }
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    *ctas_per_sm = occ;
    *grid = sms * occ;
    *threads = NT;
    return cudaSuccess;
}

cudaError_t va_glv_wide_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    switch (a.stepper) {
    case VA_RK_RK4: return launch<TabRK4, false>(a, st);
    case VA_RK_CK54: return launch<TabCK54, true>(a, st);
    case VA_RK_DOPRI5: return launch<TabDOPRI5, true>(a, st);
    }
    return cudaErrorInvalidValue;
}
