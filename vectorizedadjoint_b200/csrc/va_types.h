// va_types.h -- plain structs shared by host code, ahead-of-time kernels and the run-time compiled (NVRTC) kernels of the
// tape path. Must stay free of host-only includes: it is one of the headers handed to NVRTC (va_jit.cpp).
#pragma once
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "va_engine.h"

#define VA_MAX_STAGES 13

// Butcher tableau in the form the reference's ButcherTable exposes it (dense row-major a, b, c;
// reference lib/include/ButcherTable.hpp:254-263) plus what odeint's controller needs (db, orders, FSAL).
// Passed to kernels BY VALUE as a __grid_constant__ parameter: it then sits in the constant bank, and two
// engines with different steppers can run concurrently (no shared __constant__ symbol).
struct VaTableau {
    int s;          // stages of the forward step
    int s_adj;      // stages carrying weight in the adjoint (dopri5: the FSAL stage has b = 0 and no successors)
    int order, stepper_order, error_order;
    int fsal, has_error, pad_;
    double a[VA_MAX_STAGES * VA_MAX_STAGES];
    double b[VA_MAX_STAGES];
    double db[VA_MAX_STAGES];
    double c[VA_MAX_STAGES];
    double growth_floor; // pow(5.0, -stepper_order), the error floor of default_step_adjuster::increase_step, evaluated once on
                         // the host with the system pow() (what the reference evaluates at every accepted step)
};

// odeint util/detail/less_with_sign.hpp, as used by reference lib/include/detail/runge_kutta.hpp:55,93,98
__host__ __device__ inline bool va_less_with_sign(double t1, double t2, double dt)
{
    const double eps = 2.220446049250313e-16;
    return dt > 0 ? (t2 - t1 > eps) : (t1 - t2 > eps);
}
__host__ __device__ inline bool va_less_eq_with_sign(double t1, double t2, double dt)
{
    const double eps = 2.220446049250313e-16;
    return dt > 0 ? (t1 - t2 <= eps) : (t2 - t1 <= eps);
}

// ---- launch descriptors (engine -> kernel TUs) ---------------------------------------------------------------------

// Thread-per-trajectory family (va_scalar.cu): checkpoints in a global arena, trajectory index fastest:
//   ck_t[n * arena_stride + b],  ck_x[(n * N + i) * arena_stride + b]
struct VaScalarArgs {
    int system, stepper, adaptive, n_out;
    VaTableau tab;
    double eps_abs, eps_rel, ti, tf, dt0;
    int64_t B;            // trajectories in this launch
    int64_t arena_stride; // >= B
    int cap;              // accepted-step capacity (arena holds cap+1 entries per trajectory)
    int objective;
    const double *x0, *params;      // [B][N], [B][NPAR]
    double *x_final;                // [B][N]
    double *lambda;                 // [B][n_out][N]
    double *mu;                     // [B][n_out][NPAR]
    int32_t *n_accept, *n_reject, *status; // [B] (engine-owned or caller's)
    double *ck_t, *ck_x;
    int ck_layout;                    // 0: t[n][b], x[n][i][b] (fixed step); 1: per-trajectory records {t, x} in ck_t (adaptive)
    unsigned long long *work_counter; // [2]: next trajectory (forward kernel), next work item (reverse kernel); zeroed per launch
    int grid_limit;                   // resident CTAs the persistent grid may use
};

