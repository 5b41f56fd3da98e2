// va_scalar.cu -- thread-per-trajectory kernels for small systems (HarmonicOscillator, VanDerPol: N = 2, Npar = 1).
//
// One GPU lane integrates one parameter set ("one trajectory per lane", mirroring the reference's SIMD lanes but on
// the parameter-set axis). State, stage slopes and stage adjoints live in registers; checkpoints go to a global arena
// with the trajectory index fastest, so every lane-parallel access is coalesced.
//
// Forward: reference lib/include/detail/runge_kutta.hpp:38-72 (fixed step) and :76-118 (adaptive) with odeint's
//          controlled_runge_kutta::try_step, default_error_checker, default_step_adjuster and failed_step_checker
//          restated per lane. Arithmetic order of the stage / solution / error sums is odeint's
//          ((1*x + (a_m0 dt) k_0) + (a_m1 dt) k_1) + ...; this file is compiled with -fmad=false so that no FMA
//          contraction changes an accept/reject decision.
// Reverse: reference lib/include/detail/backpropagation.hpp:24-64 (stage recompute from the stored x_n),
//          :160-229 (one-step adjoint), :256-278 (loop over accepted steps), :280-348 (seeds -> lambda, mu).
//          The AADC vector-Jacobian product (lib/include/AadData.hpp:332-373) is a hand-written device functor.
#include "va_common.cuh"
#include "va_pow.h" // glibc-exact pow(): the step-size controller must reproduce the reference bit for bit

namespace {

// ---- device functors: rhs and vjp of the example systems --------------------------------------------------------
struct SysHarmonic { // reference examples/HarmonicOscillator/main.cpp:20-29 (k = 1)
    static constexpr int N = 2, NPAR = 1;
    __device__ static void rhs(const double *x, const double *p, double, double *dx)
    {
        const double k = 1.0;
        dx[0] = x[1];
        dx[1] = -k * x[0] - p[0] * x[1];
    }
    __device__ static void vjp(const double *x, const double *p, double, const double *w, double *gx, double *gp)
    {
        const double k = 1.0;
        gx[0] = -k * w[1];
        gx[1] = w[0] - p[0] * w[1];
        gp[0] += -x[1] * w[1];
    }
};
struct SysVanDerPol { // reference examples/VanDerPol/main.cpp:38-43
    static constexpr int N = 2, NPAR = 1;
    __device__ static void rhs(const double *x, const double *p, double, double *dx)
    {
        dx[0] = x[1];
        dx[1] = p[0] * ((1.0 - x[0] * x[0]) * x[1] - x[0]);
    }
    __device__ static void vjp(const double *x, const double *p, double, const double *w, double *gx, double *gp)
    {
        const double mu = p[0];
        gx[0] = mu * (-2.0 * x[0] * x[1] - 1.0) * w[1];
        gx[1] = w[0] + mu * (1.0 - x[0] * x[0]) * w[1];
        gp[0] += ((1.0 - x[0] * x[0]) * x[1] - x[0]) * w[1];
    }
};

// One explicit RK step in odeint's arithmetic order. K[0] = f(x,t) on entry.
template <class Sys, int S, bool FSAL, bool WITH_ERR>
__device__ __forceinline__ void rk_step(const VaTableau &tab, const double *x, const double *p, double t, double dt,
                                        double (&K)[S][Sys::N], double *xnew, double *xerr)
{
    constexpr int N = Sys::N;
    constexpr int SE = FSAL ? S - 1 : S;
    double xt[N];
#pragma unroll
    for (int m = 1; m < SE; ++m) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double acc = x[i];
#pragma unroll
            for (int j = 0; j < m; ++j) {
                const double a = tab.a[m * VA_MAX_STAGES + j];
                if (a != 0.0) acc = acc + (a * dt) * K[j][i]; // a zero weight contributes an exact zero in odeint's sum
            }
            xt[i] = acc;
        }
        Sys::rhs(xt, p, t + dt * tab.c[m], K[m]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double acc = x[i];
#pragma unroll
        for (int j = 0; j < SE; ++j) {
            const double b = tab.b[j];
            if (b != 0.0) acc = acc + (b * dt) * K[j][i];
        }
        xnew[i] = acc;
    }
    if (FSAL) Sys::rhs(xnew, p, t + dt, K[S - 1]);
    if (WITH_ERR) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double acc = 0.0;
            bool first = true;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const double d = tab.db[j];
                if (d != 0.0) {
                    const double term = (dt * d) * K[j][i];
                    acc = first ? term : acc + term;
                    first = false;
                }
            }
            xerr[i] = acc;
        }
    }
}

template <class Sys, int S, bool FSAL, bool ADAPTIVE>
__global__ void __launch_bounds__(128) k_scalar_forward(const __grid_constant__ VaScalarArgs a)
{
    constexpr int N = Sys::N, NPAR = Sys::NPAR;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const VaTableau &tab = a.tab;
    double x[N], p[NPAR], K[S][N], xnew[N], xerr[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = a.x0[b * N + i];
#pragma unroll
    for (int k = 0; k < NPAR; ++k) p[k] = a.params[b * NPAR + k];

    const int64_t bs = a.arena_stride;
    double t = a.ti, dt = a.dt0;
    const double tf = a.tf;
    int nck = 0, count = 0, rejects = 0, status = 0;

    auto push = [&]() -> bool {
        if (nck > a.cap) { status |= VA_TRAJ_CKPT_OVERFLOW; return false; }
        a.ck_t[(int64_t)nck * bs + b] = t;
#pragma unroll
        for (int i = 0; i < N; ++i) a.ck_x[((int64_t)nck * N + i) * bs + b] = x[i];
        ++nck;
        return true;
    };

    if (!ADAPTIVE) {
        // detail/runge_kutta.hpp:51-71: t = ti + step*dt (no accumulation of dt)
        while (va_less_eq_with_sign(t + dt, tf, dt)) {
            if (!push()) break;
            Sys::rhs(x, p, t, K[0]);
            rk_step<Sys, S, FSAL, false>(tab, x, p, t, dt, K, xnew, xerr);
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = xnew[i];
            ++count;
            t = a.ti + (double)count * dt;
        }
        if (!status) push();
    } else {
        // detail/runge_kutta.hpp:92-117, one attempt (try_step) per loop trip so that lanes stay converged
        bool active = va_less_with_sign(t, tf, dt);
        bool fresh = true, first_call = true;
        int trials = 0;
        while (active) {
            if (fresh) {
                if (!push()) break;
                if (va_less_with_sign(tf, t + dt, dt)) dt = tf - t;
                trials = 0;
                fresh = false;
            }
            if (!FSAL || first_call) { Sys::rhs(x, p, t, K[0]); first_call = false; }
            rk_step<Sys, S, FSAL, true>(tab, x, p, t, dt, K, xnew, xerr);
            // default_error_checker::error (a_x = a_dxdt = 1), max norm
            double err = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double e = fabs(xerr[i]) / (a.eps_abs + a.eps_rel * (fabs(x[i]) + fabs(dt) * fabs(K[0][i])));
                err = fmax(err, e);
            }
            if (err > 1.0) {
                // default_step_adjuster::decrease_step
                dt *= fmax(0.9 * va_pow(err, -1.0 / ((double)tab.error_order - 1.0)), 0.2);
                ++rejects;
                if (++trials >= 500) { status |= VA_TRAJ_NO_PROGRESS; break; } // failed_step_checker
            } else {
                t += dt;
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = xnew[i];
                if (FSAL) {
#pragma unroll
                    for (int i = 0; i < N; ++i) K[0][i] = K[S - 1][i];
                }
                // default_step_adjuster::increase_step
                if (err < 0.5) {
                    err = fmax(va_pow(5.0, -(double)tab.stepper_order), err);
                    dt *= 9.0 / 10.0 * va_pow(err, -1.0 / (double)tab.stepper_order);
                }
                ++count;
                fresh = true;
                active = va_less_with_sign(t, tf, dt);
            }
        }
        if (!status) push();
    }
    bool finite = true;
#pragma unroll
    for (int i = 0; i < N; ++i) finite = finite && isfinite(x[i]);
    if (!finite) status |= VA_TRAJ_NONFINITE;
#pragma unroll
    for (int i = 0; i < N; ++i) a.x_final[b * N + i] = status & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS) ? nan("") : x[i];
    a.n_accept[b] = count;
    a.n_reject[b] = rejects;
    a.status[b] = status;
}

// One thread per (trajectory, cost function).
template <class Sys, int S>
__global__ void __launch_bounds__(128) k_scalar_adjoint(const __grid_constant__ VaScalarArgs a)
{
    constexpr int N = Sys::N, NPAR = Sys::NPAR;
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = a.B * a.n_out;
    if (w >= total) return;
    // cost function index slowest: neighbouring lanes read neighbouring trajectories' checkpoints
    const int o = (int)(w / a.B);
    const int64_t b = w - (int64_t)o * a.B;
    const VaTableau &tab = a.tab;
    const int64_t bs = a.arena_stride;
    double *lam_io = a.lambda + (b * a.n_out + o) * N;
    double *mu_out = a.mu + (b * a.n_out + o) * NPAR;

    double p[NPAR], lam[N], mu[NPAR];
#pragma unroll
    for (int k = 0; k < NPAR; ++k) { p[k] = a.params[b * NPAR + k]; mu[k] = 0.0; }
    if (a.status[b] & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS)) {
#pragma unroll
        for (int i = 0; i < N; ++i) lam_io[i] = nan("");
#pragma unroll
        for (int k = 0; k < NPAR; ++k) mu_out[k] = nan("");
        return;
    }
    const int T = a.n_accept[b];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (a.objective == VA_OBJ_SUM) lam[i] = 1.0;
        else if (a.objective == VA_OBJ_HALF_NORM2) lam[i] = a.x_final[b * N + i];
        else lam[i] = lam_io[i];
    }
    double K[S][N], W[S + 1][N], u[N], xm[N], gx[N];
    double t_next = a.ck_t[(int64_t)T * bs + b];
    for (int n = T - 1; n >= 0; --n) {
        const double time = a.ck_t[(int64_t)n * bs + b];
        const double dt = t_next - time; // StateStorage::GetDt: difference of stored times
        t_next = time;
#pragma unroll
        for (int i = 0; i < N; ++i) u[i] = a.ck_x[((int64_t)n * N + i) * bs + b];
        // stage recompute, detail/backpropagation.hpp:37-52 (time passed unchanged to every stage, :48)
#pragma unroll
        for (int m = 0; m < S; ++m) {
#pragma unroll
            for (int i = 0; i < N; ++i) xm[i] = u[i];
#pragma unroll
            for (int j = 0; j < m; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) xm[i] += dt * tab.a[m * VA_MAX_STAGES + j] * K[j][i];
            Sys::rhs(xm, p, time, K[m]);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            W[0][i] = lam[i];
#pragma unroll
            for (int m = 1; m <= S; ++m) W[m][i] = tab.b[m - 1] * dt * lam[i];
        }
#pragma unroll
        for (int m = S; m > 0; --m) {
#pragma unroll
            for (int i = 0; i < N; ++i) xm[i] = u[i];
#pragma unroll
            for (int k = 1; k < m; ++k)
#pragma unroll
                for (int i = 0; i < N; ++i) xm[i] += dt * tab.a[(m - 1) * VA_MAX_STAGES + (k - 1)] * K[k - 1][i];
            Sys::vjp(xm, p, time, W[m], gx, mu);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                W[0][i] += gx[i];
#pragma unroll
                for (int k = 1; k < m; ++k) W[k][i] += gx[i] * tab.a[(m - 1) * VA_MAX_STAGES + (k - 1)] * dt;
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) lam[i] = W[0][i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) lam_io[i] = lam[i];
#pragma unroll
    for (int k = 0; k < NPAR; ++k) mu_out[k] = mu[k];
}

template <class Sys, int S, bool FSAL>
cudaError_t launch_forward(const VaScalarArgs &a, cudaStream_t st)
{
    const int threads = 128;
    const unsigned blocks = (unsigned)((a.B + threads - 1) / threads);
    if (a.adaptive) k_scalar_forward<Sys, S, FSAL, true><<<blocks, threads, 0, st>>>(a);
    else k_scalar_forward<Sys, S, FSAL, false><<<blocks, threads, 0, st>>>(a);
    return cudaGetLastError();
}
template <class Sys, int S>
cudaError_t launch_adjoint(const VaScalarArgs &a, cudaStream_t st)
{
    const int threads = 128;
    const unsigned blocks = (unsigned)((a.B * a.n_out + threads - 1) / threads);
    k_scalar_adjoint<Sys, S><<<blocks, threads, 0, st>>>(a);
    return cudaGetLastError();
}

template <class Sys>
cudaError_t dispatch_forward(const VaScalarArgs &a, cudaStream_t st)
{
    switch (a.stepper) {
    case VA_RK_EULER: return launch_forward<Sys, 1, false>(a, st);
    case VA_RK_RK4: return launch_forward<Sys, 4, false>(a, st);
    case VA_RK_CK54: return launch_forward<Sys, 6, false>(a, st);
    case VA_RK_DOPRI5: return launch_forward<Sys, 7, true>(a, st);
    case VA_RK_RKF78: return launch_forward<Sys, 13, false>(a, st);
    }
    return cudaErrorInvalidValue;
}
template <class Sys>
cudaError_t dispatch_adjoint(const VaScalarArgs &a, cudaStream_t st)
{
    switch (a.stepper) {
    case VA_RK_EULER: return launch_adjoint<Sys, 1>(a, st);
    case VA_RK_RK4: return launch_adjoint<Sys, 4>(a, st);
    case VA_RK_CK54: return launch_adjoint<Sys, 6>(a, st);
    case VA_RK_DOPRI5: return launch_adjoint<Sys, 6>(a, st); // the FSAL stage carries no adjoint weight
    case VA_RK_RKF78: return launch_adjoint<Sys, 13>(a, st);
    }
    return cudaErrorInvalidValue;
}

} // namespace

cudaError_t va_scalar_forward(const VaScalarArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    switch (a.system) {
    case VA_SYS_HARMONIC: return dispatch_forward<SysHarmonic>(a, st);
    case VA_SYS_VANDERPOL: return dispatch_forward<SysVanDerPol>(a, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t va_scalar_adjoint(const VaScalarArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    switch (a.system) {
    case VA_SYS_HARMONIC: return dispatch_adjoint<SysHarmonic>(a, st);
    case VA_SYS_VANDERPOL: return dispatch_adjoint<SysVanDerPol>(a, st);
    }
    return cudaErrorInvalidValue;
}
