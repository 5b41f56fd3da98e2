// va_glv_common.cuh -- pieces shared by the GLV kernel families (va_glv_wide.cu: N <= 64, matrix in registers;
// va_glv_stream.cu: any N, matrix streamed from L2/HBM): compile-time tableaux and the controller's inverse root.
#pragma once
#include "va_common.cuh"

namespace {

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

// explicit Euler (reference lib/include/ButcherTable.hpp:50-65): one stage, b = (1)
struct TabEuler {
    static constexpr int S = 1, SADJ = 1, STEPPER_ORDER = 1, ERROR_ORDER = 0;
    static constexpr bool FSAL = false, HAS_ERR = false;
    __host__ __device__ static constexpr double a(int, int) { return 0.0; }
    __host__ __device__ static constexpr double b(int) { return 1.0; }
    __host__ __device__ static constexpr double db(int) { return 0.0; }
};
// Fehlberg 7(8), 13 stages (reference lib/include/ButcherTable.hpp:191-246; odeint rk78_coefficients_*): orders 8 / 8 / 7
struct TabRKF78 {
    static constexpr int S = 13, SADJ = 13, STEPPER_ORDER = 8, ERROR_ORDER = 7;
    static constexpr bool FSAL = false, HAS_ERR = true;
    __host__ __device__ static constexpr double a(int m, int j)
    {
        constexpr double t[13][12] = {
            {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
            {2.0 / 27, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
            {1.0 / 36, 1.0 / 12, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
            {1.0 / 24, 0, 1.0 / 8, 0, 0, 0, 0, 0, 0, 0, 0, 0},
            {5.0 / 12, 0, -25.0 / 16, 25.0 / 16, 0, 0, 0, 0, 0, 0, 0, 0},
            {1.0 / 20, 0, 0, 1.0 / 4, 1.0 / 5, 0, 0, 0, 0, 0, 0, 0},
            {-25.0 / 108, 0, 0, 125.0 / 108, -65.0 / 27, 125.0 / 54, 0, 0, 0, 0, 0, 0},
            {31.0 / 300, 0, 0, 0, 61.0 / 225, -2.0 / 9, 13.0 / 900, 0, 0, 0, 0, 0},
            {2.0, 0, 0, -53.0 / 6, 704.0 / 45, -107.0 / 9, 67.0 / 90, 3.0, 0, 0, 0, 0},
            {-91.0 / 108, 0, 0, 23.0 / 108, -976.0 / 135, 311.0 / 54, -19.0 / 60, 17.0 / 6, -1.0 / 12, 0, 0, 0},
            {2383.0 / 4100, 0, 0, -341.0 / 164, 4496.0 / 1025, -301.0 / 82, 2133.0 / 4100, 45.0 / 82, 45.0 / 164, 18.0 / 41, 0, 0},
            {3.0 / 205, 0, 0, 0, 0, -6.0 / 41, -3.0 / 205, -3.0 / 41, 3.0 / 41, 6.0 / 41, 0, 0},
            {-1777.0 / 4100, 0, 0, -341.0 / 164, 4496.0 / 1025, -289.0 / 82, 2193.0 / 4100, 51.0 / 82, 33.0 / 164, 12.0 / 41, 0, 1.0}};
        return t[m][j];
    }
    __host__ __device__ static constexpr double b(int j)
    {
        constexpr double t[13] = {0, 0, 0, 0, 0, 34.0 / 105, 9.0 / 35, 9.0 / 35, 9.0 / 280, 9.0 / 280, 0, 41.0 / 840, 41.0 / 840};
        return t[j];
    }
    __host__ __device__ static constexpr double db(int j)
    {
        constexpr double t[13] = {0.0 - 41.0 / 840, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.0 - 41.0 / 840, 41.0 / 840, 41.0 / 840};
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


} // namespace
