// va_scalar.cu -- ahead-of-time instantiations of the thread-per-trajectory kernels (va_scalar_kernels.cuh) for the
// built-in small systems: HarmonicOscillator and VanDerPol (N = 2, Npar = 1), all steppers.
#include <algorithm>

#include "va_common.cuh"
#include "va_scalar_kernels.cuh"

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

template <class Sys, int S, bool FSAL>
cudaError_t launch_forward(const VaScalarArgs &a, cudaStream_t st)
{
    const int threads = 128;
    const unsigned blocks = (unsigned)std::min<int64_t>((a.B + threads - 1) / threads, a.grid_limit > 0 ? a.grid_limit : 148 * 4);
    if (a.adaptive) k_scalar_forward<Sys, S, FSAL, true><<<blocks, threads, 0, st>>>(a);
    else k_scalar_forward<Sys, S, FSAL, false><<<blocks, threads, 0, st>>>(a);
    return cudaGetLastError();
}
template <class Sys, int S>
cudaError_t launch_adjoint(const VaScalarArgs &a, cudaStream_t st)
{
    const int threads = 128;
    const unsigned blocks = (unsigned)std::min<int64_t>((a.B * a.n_out + threads - 1) / threads, a.grid_limit > 0 ? a.grid_limit : 148 * 4);
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
