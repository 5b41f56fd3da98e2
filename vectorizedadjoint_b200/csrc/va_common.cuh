// va_common.cuh -- types shared by the engine (va_engine.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/va_engine.h"

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
};

int va_tableau_host(int kind, VaTableau *tb); // va_tableau.cpp

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
};
cudaError_t va_scalar_forward(const VaScalarArgs &a, cudaStream_t st);
cudaError_t va_scalar_adjoint(const VaScalarArgs &a, cudaStream_t st);

// CTA-per-trajectory GLV family (va_glv_wide.cu). One persistent CTA integrates trajectory after trajectory, forward
// then backward, with its private checkpoint slab (L2-resident for typical step counts).
struct VaGlvWideArgs {
    int n;                // species (<= 64 for the NP=64 kernel; padded with inert species)
    int stepper, adaptive, n_out, objective, reduce;
    double eps_abs, eps_rel, ti, tf, dt0;
    int64_t B;
    int cap;              // accepted-step capacity of the slab
    const double *x0, *params;
    double *x_final, *lambda, *mu;
    int32_t *n_accept, *n_reject, *status;
    double *slab;         // grid * slab_stride doubles
    int64_t slab_stride;
    double *partial;      // VA_REDUCE_SUM: [grid][n_par] per-CTA partial sums (reduced by va_reduce_rows)
    int grid;
    struct { double a[7][6], b[7], db[7]; } coef; // tableau values, filled by the launcher
};
bool va_glv_wide_supported(int n, int stepper, int adaptive);
int64_t va_glv_wide_slab_doubles(int n, int stepper, int cap);
int va_glv_wide_block_doubles(int stepper); // step block: [8-double header (t_n) | X_0..X_{s-1} | g_0..g_{s-1}], 64 wide
cudaError_t va_glv_wide_config(int n, int stepper, int device, int *grid, int *ctas_per_sm, int *threads);
cudaError_t va_glv_wide_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st);

// out[k] (+)= sum_{g<G} in[g*stride + k], deterministic order
cudaError_t va_reduce_rows(const double *in, int64_t G, int64_t stride, int64_t n, double *out, int accumulate, cudaStream_t st);

// synthetic inputs / microbenchmarks (va_util.cu)
cudaError_t va_synth_launch(int system, int n, uint64_t seed, int64_t b0, int64_t B, double *params, double *x0, cudaStream_t st);
