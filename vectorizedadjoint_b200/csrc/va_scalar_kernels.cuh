// va_scalar_kernels.cuh -- thread-per-trajectory forward / reverse sweeps, generic in the system functor.
//
// Included by va_scalar.cu (ahead-of-time instantiations for the built-in systems) AND handed to NVRTC together with a
// generated functor (va_tape.hpp -> Tape::cuda_source) for recorded systems (va_jit.cpp). Keep it free of host-only code.
//
// One GPU lane integrates one parameter set ("one trajectory per lane", mirroring the reference's SIMD lanes but on
// the parameter-set axis). State, stage slopes and stage adjoints live in registers; checkpoints go to a global arena
// with the trajectory index fastest, so every lane-parallel access is coalesced.
//
// Forward: reference lib/include/detail/runge_kutta.hpp:38-72 (fixed step) and :76-118 (adaptive) with odeint's
//          controlled_runge_kutta::try_step, default_error_checker, default_step_adjuster and failed_step_checker
//          restated per lane. Arithmetic order of the stage / solution / error sums is odeint's
//          ((1*x + (a_m0 dt) k_0) + (a_m1 dt) k_1) + ...; compiled with -fmad=false so that no FMA contraction changes
//          an accept/reject decision; pow() is glibc's bit for bit (va_pow.h).
// Reverse: reference lib/include/detail/backpropagation.hpp:24-64 (stage recompute from the stored x_n),
//          :160-229 (one-step adjoint), :256-278 (loop over accepted steps), :280-348 (seeds -> lambda, mu).
//          The AADC vector-Jacobian product (lib/include/AadData.hpp:332-373) is the functor's vjp().
#pragma once
#include "va_types.h"
#include "va_pow.h"

// Parameter placement. Up to VA_REG_PARAMS parameters a lane keeps its parameter set (and, in the reverse sweep, its gradient
// accumulator) in registers. Wider recorded systems (a user-written Lotka-Volterra variant with 16 species has 272) read the
// parameters where they lie -- params[b][k], through L1/L2 -- and accumulate dJ/dalpha directly in the caller's mu row, so the
// thread-per-trajectory kernels have no parameter-count limit (reference: AadData::Record is size-agnostic,
// lib/include/AadData.hpp:124-171).
#ifndef VA_REG_PARAMS
#define VA_REG_PARAMS 64
#endif

// Checkpoint store (the reference's StateStorage, lib/include/StateStorage.hpp:4-42) in HBM. Two layouts:
//   ck_layout 0 (fixed step): t[n][b], x[n][i][b] -- trajectory index fastest. All lanes are at the same step n, so every
//       access of a warp is one coalesced segment (harmonic oscillator: 57 % of HBM peak).
//   ck_layout 1 (adaptive):   rec[b][n] = {t, x_0..x_{N-1}} -- one contiguous record stream per trajectory. With adaptive
//       steps and dynamically scheduled lanes, neighbouring lanes are at unrelated (n, b); a lane then walks its own
//       stream sequentially (one 32 B sector per record for two-state systems) instead of touching one sector per scalar.
// Record stride of layout 1: (t, x_0..x_{N-1}) padded to a multiple of four doubles, so that a record never straddles a 32-byte
// sector and, for two-state systems, is ONE 256-bit store / load (STG.256 / LDG.256, sm_100). With three separate 8-byte
// stores per step the forward kernel stalled on its store queue (profiles/r02/vdp_forward_ncu_full_before.txt).
#define VA_CK_REC(N) ((((N) + 1) + 3) & ~3)
template <int N>
__device__ __forceinline__ void ck_store(const VaScalarArgs &a, int64_t b, int n, double t, const double *x)
{
    if (a.ck_layout == 0) {
        a.ck_t[(int64_t)n * a.arena_stride + b] = t;
#pragma unroll
        for (int i = 0; i < N; ++i) a.ck_x[((int64_t)n * N + i) * a.arena_stride + b] = x[i];
    } else {
        constexpr int REC = VA_CK_REC(N);
        double *rec = a.ck_t + (b * (int64_t)(a.cap + 1) + n) * REC;
        if (REC == 4) {
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(rec), "d"(t), "d"(x[0]), "d"(x[N > 1 ? 1 : 0]), "d"(x[N > 2 ? 2 : 0])
                         : "memory");
        } else {
            rec[0] = t;
#pragma unroll
            for (int i = 0; i < N; ++i) rec[1 + i] = x[i];
        }
    }
}
template <int N>
__device__ __forceinline__ double ck_time(const VaScalarArgs &a, int64_t b, int n)
{
    return a.ck_layout == 0 ? a.ck_t[(int64_t)n * a.arena_stride + b] : a.ck_t[(b * (int64_t)(a.cap + 1) + n) * VA_CK_REC(N)];
}
// time and state of checkpoint n in one access where the record allows it
template <int N>
__device__ __forceinline__ double ck_load(const VaScalarArgs &a, int64_t b, int n, double *x)
{
    if (a.ck_layout == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = a.ck_x[((int64_t)n * N + i) * a.arena_stride + b];
        return a.ck_t[(int64_t)n * a.arena_stride + b];
    }
    constexpr int REC = VA_CK_REC(N);
    const double *rec = a.ck_t + (b * (int64_t)(a.cap + 1) + n) * REC;
    if (REC == 4) {
        double r0, r1, r2, r3;
        asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r0), "=d"(r1), "=d"(r2), "=d"(r3) : "l"(rec) : "memory");
        x[0] = r1;
        if (N > 1) x[N > 1 ? 1 : 0] = r2;
        if (N > 2) x[N > 2 ? 2 : 0] = r3;
        return r0;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = rec[1 + i];
    return rec[0];
}

// Which tableau entries are non-zero, known at compile time. A kernel instantiation is one stepper -- (S, FSAL) identifies it:
// 1 euler, 4 rk4, 6 cash_karp54, 7 + FSAL dopri5, 13 fehlberg78 -- so the zero weights of its tableau are dropped from the
// unrolled sums by the compiler (same arithmetic on the non-zero terms: odeint sums exact zeros there). The VALUES still come
// from the kernel argument (va_tableau.cpp). With run-time tests `if (a != 0.0)` a quarter of the forward kernel's instructions
// were DSETP / FSEL pairs (profiles/r02/vdp_forward_ncu_full_after.txt).
__device__ constexpr bool va_nz_a(int S, int m, int j)
{
    if (j >= m) return false;
    if (S == 4) return j == m - 1;                            // rk4: sub-diagonal only
    if (S == 7) return !(m == 6 && j == 1);                   // dopri5: the FSAL row is b, and b_1 = 0
    if (S == 13) {                                            // fehlberg78
        switch (m) {
        case 1: case 2: return true;
        case 3: return j != 1;
        case 4: return j != 1;
        case 5: return j == 0 || j >= 3;
        case 6: return j == 0 || j >= 3;
        case 7: return j == 0 || j >= 4;
        case 8: return j == 0 || j >= 3;
        case 9: return j == 0 || j >= 3;
        case 10: return j == 0 || j >= 3;
        case 11: return j == 0 || (j >= 5 && j <= 9);
        case 12: return j == 0 || (j >= 3 && j != 10);
        }
    }
    return true;                                              // euler (no entries), cash_karp54 (full lower triangle)
}
__device__ constexpr bool va_nz_b(int S, int j)
{
    if (S == 6) return j != 1 && j != 4;
    if (S == 7) return j != 1 && j != 6;
    if (S == 13) return (j >= 5 && j <= 9) || j >= 11;
    return true;
}
__device__ constexpr bool va_nz_db(int S, int j)
{
    if (S == 6 || S == 7) return j != 1;
    if (S == 13) return j == 0 || j >= 10;
    return false;
}

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
                if (va_nz_a(S, m, j)) acc = acc + (tab.a[m * VA_MAX_STAGES + j] * dt) * K[j][i]; // zero weights add exact zeros in odeint's sum
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
            if (va_nz_b(S, j)) acc = acc + (tab.b[j] * dt) * K[j][i];
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
                if (va_nz_db(S, j)) {
                    const double term = (dt * tab.db[j]) * K[j][i];
                    acc = first ? term : acc + term;
                    first = false;
                }
            }
            xerr[i] = acc;
        }
    }
}

// Divergence control. Per-trajectory work varies widely (Van der Pol over three decades of mu: 6...600 accepted steps,
// bursts of rejections), so lanes are NOT bound to trajectories: the grid is persistent, and a lane that finishes its
// trajectory immediately fetches the next one from a global counter (work_counter). Every loop trip is one controller
// attempt (one try_step) for every lane that has work, so a warp stays converged on the expensive code and idles only at
// the very end of the batch. Checkpoints are addressed by trajectory index, so the reverse kernel can do the same.
template <class Sys, int S, bool FSAL, bool ADAPTIVE>
__device__ __forceinline__ void scalar_forward_body(const VaScalarArgs &a)
{
    constexpr int N = Sys::N, NPAR = Sys::NPAR;
    constexpr bool PGLOBAL = NPAR > VA_REG_PARAMS;
    const VaTableau &tab = a.tab;
    const double tf = a.tf;
    double x[N], preg[PGLOBAL ? 1 : NPAR], K[S][N], xnew[N], xerr[N];
    const double *p = preg;
    double t = 0.0, dt = 0.0;
    int64_t b = -1;
    int nck = 0, count = 0, rejects = 0, status = 0, trials = 0;
    bool active = false, fresh = true, first_call = true;
    // the controller's pow() tables (5 KB) staged in shared memory: every attempt gathers one entry of each per lane
    __shared__ va_pow_logtab s_logtab[ADAPTIVE ? 128 : 1];
    __shared__ uint64_t s_exptab[ADAPTIVE ? 256 : 1];
    if (ADAPTIVE) {
        for (int k = threadIdx.x; k < 128; k += blockDim.x) s_logtab[k] = va_pow_log_tab_d[k];
        for (int k = threadIdx.x; k < 256; k += blockDim.x) s_exptab[k] = va_pow_exp_tab_d[k];
        __syncthreads();
    }

    auto push = [&]() -> bool {
        if (nck > a.cap) { status |= VA_TRAJ_CKPT_OVERFLOW; return false; }
        ck_store<N>(a, b, nck, t, x);
        ++nck;
        return true;
    };
    auto finalize = [&]() {
        if (!status) push(); // the closing (t, x) entry: detail/runge_kutta.hpp:68-69, 113-115
        bool finite = true;
#pragma unroll
        for (int i = 0; i < N; ++i) finite = finite && isfinite(x[i]);
        if (!finite) status |= VA_TRAJ_NONFINITE;
#pragma unroll
        for (int i = 0; i < N; ++i) a.x_final[b * N + i] = status & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS) ? nan("") : x[i];
        a.n_accept[b] = count;
        a.n_reject[b] = rejects;
        a.status[b] = status;
        active = false;
    };

    for (;;) {
        if (!active) {
            b = (int64_t)atomicAdd(a.work_counter, 1ULL);
            if (b >= a.B) break;
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = a.x0[b * N + i];
            if (PGLOBAL) p = a.params + b * NPAR;
            else {
#pragma unroll
                for (int k = 0; k < NPAR; ++k) preg[k] = a.params[b * NPAR + k];
            }
            t = a.ti; dt = a.dt0;
            nck = count = rejects = status = trials = 0;
            fresh = first_call = true;
            active = ADAPTIVE ? va_less_with_sign(t, tf, dt) : va_less_eq_with_sign(t + dt, tf, dt);
            if (!active) { finalize(); continue; }
        }
        if (!ADAPTIVE) {
            // detail/runge_kutta.hpp:51-71: t = ti + step*dt (no accumulation of dt)
            if (!push()) { finalize(); continue; }
            Sys::rhs(x, p, t, K[0]);
            rk_step<Sys, S, FSAL, false>(tab, x, p, t, dt, K, xnew, xerr);
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = xnew[i];
            ++count;
            t = a.ti + (double)count * dt;
            if (!va_less_eq_with_sign(t + dt, tf, dt)) finalize();
        } else {
            // detail/runge_kutta.hpp:92-117, one attempt (try_step) per loop trip
            if (fresh) {
                if (!push()) { finalize(); continue; }
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
            // default_step_adjuster: decrease_step (err > 1) and increase_step (err < 0.5) both need one pow(); a single
            // call site with lane-dependent arguments keeps the warp converged through the ~60-instruction routine
            const bool reject = err > 1.0;
            const bool grow = err < 0.5;
            double factor = 1.0;
            if (reject || grow) {
                const double base = reject ? err : fmax(tab.growth_floor, err); // growth_floor = pow(5.0, -stepper_order)
                const double expo = reject ? -1.0 / ((double)tab.error_order - 1.0) : -1.0 / (double)tab.stepper_order;
                const double pw = va_pow_t(base, expo, s_logtab, s_exptab);
                factor = reject ? fmax(0.9 * pw, 0.2) : 9.0 / 10.0 * pw;
            }
            if (reject) {
                dt *= factor;
                ++rejects;
                if (++trials >= 500) { status |= VA_TRAJ_NO_PROGRESS; finalize(); } // failed_step_checker
            } else {
                t += dt;
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = xnew[i];
                if (FSAL) {
#pragma unroll
                    for (int i = 0; i < N; ++i) K[0][i] = K[S - 1][i];
                }
                if (grow) dt *= factor;
                ++count;
                fresh = true;
                if (!va_less_with_sign(t, tf, dt)) finalize();
            }
        }
    }
}

// One work item per (trajectory, cost function), fetched dynamically like in the forward kernel; one accepted step of
// the reverse recursion per loop trip.
template <class Sys, int S>
__device__ __forceinline__ void scalar_adjoint_body(const VaScalarArgs &a)
{
    constexpr int N = Sys::N, NPAR = Sys::NPAR;
    constexpr bool PGLOBAL = NPAR > VA_REG_PARAMS;
    // S counts the stages that carry weight in the adjoint: 6 is cash_karp54 or dopri5 without its FSAL stage -- both have a full
    // lower triangle there, so the a-mask of "6 stages" is right for either (the b weights are taken at run time)
    constexpr int SFULL = S;
    const int64_t total = a.B * a.n_out;
    const VaTableau &tab = a.tab;
    double preg[PGLOBAL ? 1 : NPAR], lam[N], mureg[PGLOBAL ? 1 : NPAR];
    const double *p = preg;
    double *mu = mureg;
    double K[S][N], W[S + 1][N], u[N], xm[N], gx[N];
    double t_next = 0.0;
    int64_t b = 0;
    double *lam_io = nullptr, *mu_out = nullptr;
    int n = -1;
    bool have = false;
    for (;;) {
        if (n < 0) {
            if (have) { // write back the finished work item
#pragma unroll
                for (int i = 0; i < N; ++i) lam_io[i] = lam[i];
                if (!PGLOBAL) {
#pragma unroll
                    for (int k = 0; k < NPAR; ++k) mu_out[k] = mureg[k];
                }
                have = false;
            }
            const int64_t w = (int64_t)atomicAdd(a.work_counter + 1, 1ULL);
            if (w >= total) break;
            // cost function index slowest: neighbouring items are neighbouring trajectories' checkpoints
            const int o = (int)(w / a.B);
            b = w - (int64_t)o * a.B;
            lam_io = a.lambda + (b * a.n_out + o) * N;
            mu_out = a.mu + (b * a.n_out + o) * NPAR;
            if (PGLOBAL) {
                p = a.params + b * NPAR;
                mu = mu_out; // the gradient accumulates in the caller's row
                for (int k = 0; k < NPAR; ++k) mu_out[k] = 0.0;
            } else {
#pragma unroll
                for (int k = 0; k < NPAR; ++k) { preg[k] = a.params[b * NPAR + k]; mureg[k] = 0.0; }
            }
            if (a.status[b] & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS)) {
#pragma unroll
                for (int i = 0; i < N; ++i) lam_io[i] = nan("");
#pragma unroll
                for (int k = 0; k < NPAR; ++k) mu_out[k] = nan("");
                continue;
            }
            const int T = a.n_accept[b];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (a.objective == VA_OBJ_SUM) lam[i] = 1.0;
                else if (a.objective == VA_OBJ_HALF_NORM2) lam[i] = a.x_final[b * N + i];
                else lam[i] = lam_io[i];
            }
            t_next = ck_time<N>(a, b, T);
            n = T - 1;
            have = true;
            if (n < 0) continue;
        }
        const double time = ck_load<N>(a, b, n, u);
        const double dt = t_next - time; // StateStorage::GetDt: difference of stored times
        t_next = time;
        // stage recompute, detail/backpropagation.hpp:37-52. The reference passes t_n to every stage (:48) and indexes
        // c(m) off by one in the VJP (:127); harmless there because its examples are autonomous. Here stage m is
        // evaluated at t_n + c_m dt, so recorded non-autonomous systems differentiate correctly.
#pragma unroll
        for (int m = 0; m < S; ++m) {
#pragma unroll
            for (int i = 0; i < N; ++i) xm[i] = u[i];
#pragma unroll
            for (int j = 0; j < m; ++j)
                if (va_nz_a(SFULL, m, j)) {
                    // explicit fma: the reverse sweep takes no accept/reject decision, so unlike the forward sweep it need not
                    // reproduce odeint's rounding; one FP64 instruction per term instead of three (the kernel is FP64-pipe bound)
                    const double c = dt * tab.a[m * VA_MAX_STAGES + j];
#pragma unroll
                    for (int i = 0; i < N; ++i) xm[i] = fma(c, K[j][i], xm[i]);
                }
            Sys::rhs(xm, p, time + tab.c[m] * dt, K[m]);
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
                if (va_nz_a(SFULL, m - 1, k - 1)) {
                    const double c = dt * tab.a[(m - 1) * VA_MAX_STAGES + (k - 1)];
#pragma unroll
                    for (int i = 0; i < N; ++i) xm[i] = fma(c, K[k - 1][i], xm[i]);
                }
            Sys::vjp(xm, p, time + tab.c[m - 1] * dt, W[m], gx, mu);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                W[0][i] += gx[i];
#pragma unroll
                for (int k = 1; k < m; ++k)
                    if (va_nz_a(SFULL, m - 1, k - 1)) W[k][i] = fma(gx[i], tab.a[(m - 1) * VA_MAX_STAGES + (k - 1)] * dt, W[k][i]);
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) lam[i] = W[0][i];
        --n;
    }
}

#ifndef VA_SCALAR_FWD_MINB
#define VA_SCALAR_FWD_MINB 1 // resident CTAs per SM the forward kernel of the <= 7-stage steppers is sized for (experiment knob)
#endif
template <class Sys, int S, bool FSAL, bool ADAPTIVE>
__global__ void __launch_bounds__(128, (S <= 7 ? VA_SCALAR_FWD_MINB : 1)) k_scalar_forward(const __grid_constant__ VaScalarArgs a)
{
    scalar_forward_body<Sys, S, FSAL, ADAPTIVE>(a);
}
template <class Sys, int S>
__global__ void __launch_bounds__(128) k_scalar_adjoint(const __grid_constant__ VaScalarArgs a)
{
    scalar_adjoint_body<Sys, S>(a);
}
