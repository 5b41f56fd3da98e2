// CPU check of the tape recorder behind recordDriverRHSFunction (vectorizedadjoint_b200/include/va_tape.hpp).
#include <cmath>
#include <cstdio>
#include <iostream>

#include "lib.hpp"

struct HO {
    double k = 1.0;
    template <class T> void operator()(const std::vector<T> &r, std::vector<T> &d, const std::vector<T> &mu, const T) const
    {
        d[0] = r[1];
        d[1] = -k * r[0] - mu[0] * r[1];
    }
};
struct HO_k2 { // same structure, different constant: must NOT be taken for the built-in functor
    double k = 2.0;
    template <class T> void operator()(const std::vector<T> &r, std::vector<T> &d, const std::vector<T> &mu, const T) const
    {
        d[0] = r[1];
        d[1] = -k * r[0] - mu[0] * r[1];
    }
};
struct VdP {
    template <class T> void operator()(const std::vector<T> &x, std::vector<T> &d, const std::vector<T> &mu, const T) const
    {
        d[0] = x[1];
        d[1] = mu[0] * ((1.0 - x[0] * x[0]) * x[1] - x[0]);
    }
};
struct GLV {
    template <class T> void operator()(const std::vector<T> &x, std::vector<T> &d, const std::vector<T> &p, T)
    {
        int N = x.size();
        for (int i = 0; i < N; i++) {
            T sum = 0.0;
            for (int j = 0; j < N; j++) sum += p[N * (i + 1) + j] * x[j];
            d[i] = x[i] * (p[i] + sum);
        }
    }
};
struct GLV_reordered { // mathematically the same RHS written differently: still the built-in functor
    template <class T> void operator()(const std::vector<T> &x, std::vector<T> &d, const std::vector<T> &p, T)
    {
        int N = x.size();
        for (int i = 0; i < N; i++) {
            T acc = p[i] * x[i];
            for (int j = N - 1; j >= 0; j--) acc += x[i] * x[j] * p[N * (i + 1) + j];
            d[i] = acc;
        }
    }
};
struct Pendulum {
    template <class T> void operator()(const std::vector<T> &x, std::vector<T> &d, const std::vector<T> &p, const T t) const
    {
        using std::sin;
        d[0] = x[1];
        d[1] = -p[0] * sin(x[0]) - p[1] * x[1] + exp(-t) / (1.0 + x[0] * x[0]);
    }
};

// every elementary function and the comparison / select surface (AADC idouble.h:660-741, ibool.h:19-28)
struct Zoo {
    template <class T> void operator()(const std::vector<T> &x, std::vector<T> &d, const std::vector<T> &p, const T t) const
    {
        using namespace std;
        using va::iIf;
        d[0] = tan(x[0]) + asin(x[1] * 0.5) - acos(p[0] * x[0]) + atan(x[2] * p[1]) + sinh(x[1]) * cosh(x[0]);
        d[1] = log10(2.0 + x[0] * x[0]) + log2(1.5 + p[1] * p[1]) + exp2(x[2] * 0.3) + cbrt(1.0 + x[1] * x[1]) + erf(p[0] * x[2]) + fabs(x[0] - 0.9);
        d[2] = atan2(x[0] + p[0], 1.0 + x[1] * x[1]) + fmod(3.7 * x[2] + p[1], 1.3) + fmin(x[0] * x[1], p[0]) + fmax(x[2], x[1] * p[1]) +
               iIf(x[0] * p[1] < x[1] + t, x[0] * x[0], -x[1]) + iIf(x[2] >= p[0], p[1] * x[2], x[2] * x[2] * t);
        // == != on active values and && || ! != on recorded conditions (AADC ibool.h:81-190)
        d[0] += iIf((x[0] < 0.5) && !(x[1] > 0.3), p[0] * x[0] * x[1], x[2] * x[2]);
        d[1] += iIf((x[2] > 0.8) || (x[0] == x[0] * 1.0 && x[1] != x[1] + 1.0 && x[0] > 0.9), x[1] * p[1], -x[0] * x[2]);
        d[2] += iIf((x[0] < 0.5) != (x[1] < 0.2), p[1] * x[0], p[0] * x[1]) + iIf(true && (x[1] <= 0.27), 1.0 * x[2], 2.0 * x[2]);
    }
};
// coincides with the built-in harmonic oscillator wherever |r| < 5, differs beyond: must stay on the generated path
struct HO_clamped {
    template <class T> void operator()(const std::vector<T> &r, std::vector<T> &d, const std::vector<T> &mu, const T) const
    {
        using std::fmax;
        using std::fmin;
        d[0] = r[1];
        d[1] = -1.0 * fmax(fmin(r[0], T(5.0)), T(-5.0)) - mu[0] * r[1];
    }
};
struct HO_switch { // same, written with a recorded condition
    template <class T> void operator()(const std::vector<T> &r, std::vector<T> &d, const std::vector<T> &mu, const T) const
    {
        using va::iIf;
        d[0] = r[1];
        d[1] = iIf(r[0] < 100.0, -1.0 * r[0], T(-100.0)) - mu[0] * r[1];
    }
};

int main(int argc, char **argv)
{
    int fails = 0;
    auto expect = [&](const char *name, int got, int want) {
        std::printf("%-16s kind %d (want %d)\n", name, got, want);
        fails += got != want;
    };
    expect("harmonic", va::identify(va::record(HO(), 2, 1)), va::SYS_HARMONIC);
    expect("harmonic k=2", va::identify(va::record(HO_k2(), 2, 1)), va::SYS_TAPE);
    expect("vanderpol", va::identify(va::record(VdP(), 2, 1)), va::SYS_VANDERPOL);
    expect("glv N=5", va::identify(va::record(GLV(), 5, 30)), va::SYS_GLV);
    expect("glv N=64", va::identify(va::record(GLV(), 64, 4160)), va::SYS_GLV);
    expect("glv reordered", va::identify(va::record(GLV_reordered(), 7, 56)), va::SYS_GLV);
    expect("pendulum", va::identify(va::record(Pendulum(), 2, 2)), va::SYS_TAPE);
    expect("harmonic clamped", va::identify(va::record(HO_clamped(), 2, 1)), va::SYS_TAPE);
    expect("harmonic switch", va::identify(va::record(HO_switch(), 2, 1)), va::SYS_TAPE);
    // tape evaluation == direct evaluation; generated vjp source is straight-line CUDA
    va::Tape tp = va::record(Pendulum(), 2, 2);
    std::vector<double> x = {0.4, -0.2}, p = {1.3, 0.05}, f(2), g(2), work;
    tp.eval(x.data(), p.data(), 0.7, f.data(), work);
    Pendulum()(x, g, p, 0.7);
    fails += !(f[0] == g[0] && std::fabs(f[1] - g[1]) < 1e-15);
    const std::string src = tp.cuda_source("SysPendulum");
    fails += src.find("__device__ static void vjp") == std::string::npos || src.find("sin(") == std::string::npos;
    // tape -> CUDA -> NVRTC (sm_100a): the generated functor compiles inside the engine's thread-per-trajectory kernels,
    // for a fixed-step and for controlled steppers (no device needed for the compile step)
    char log[4096];
    const std::string user = tp.cuda_source("VaUserSys");
    for (int stepper : {VA_RK_RK4, VA_RK_DOPRI5, VA_RK_RKF78}) {
        const int rc = va_tape_compile_check(user.c_str(), stepper, log, sizeof(log));
        std::printf("nvrtc stepper %d: rc %d %s\n", stepper, rc, rc ? va_last_error() : "");
        fails += rc != VA_OK;
    }
    fails += va_tape_compile_check("struct VaUserSys { this is not CUDA };", VA_RK_RK4, log, sizeof(log)) != VA_E_NVRTC;
    // the wider operation set: tape evaluation == direct evaluation; the functor compiles under NVRTC; its generated source is
    // written out for the host-side derivative check of tests/test_dropin_cpu.py
    {
        va::Tape tz = va::record(Zoo(), 3, 2);
        fails += va::identify(tz) != va::SYS_TAPE;
        std::vector<double> xz = {0.31, 0.44, 0.52}, pz = {0.8, 0.6}, fz(3), gz(3);
        tz.eval(xz.data(), pz.data(), 0.25, fz.data(), work);
        Zoo()(xz, gz, pz, 0.25);
        for (int i = 0; i < 3; ++i) fails += !(std::fabs(fz[i] - gz[i]) <= 1e-15 * std::fabs(gz[i]));
        // the other branch of both selects
        std::vector<double> xy = {0.95, 0.1, 0.85};
        tz.eval(xy.data(), pz.data(), 0.25, fz.data(), work);
        Zoo()(xy, gz, pz, 0.25);
        for (int i = 0; i < 3; ++i) fails += !(std::fabs(fz[i] - gz[i]) <= 1e-15 * std::fabs(gz[i]));
        const std::string zoo = tz.cuda_source("VaUserSys");
        const int rc = va_tape_compile_check(zoo.c_str(), VA_RK_DOPRI5, log, sizeof(log));
        std::printf("nvrtc zoo: rc %d %s\n", rc, rc ? va_last_error() : "");
        fails += rc != VA_OK;
        if (argc > 1) {
            FILE *fo = std::fopen(argv[1], "w");
            std::fputs(zoo.c_str(), fo);
            std::fclose(fo);
        }
    }
    // Driver surface without a device: preconditions are reported on stdout, the call returns (reference behaviour)
    vectorizedadjoint::Driver driver(2, 1, 1);
    std::vector<double> mu = {0.1};
    vectorizedadjoint::adjointSolve(driver, mu); // prints "Must call setCostGradients() first!"
    vectorizedadjoint::recordDriverRHSFunction(driver, HO());
    std::vector<double> u = {0.5, 0.25}, du(2);
    driver.Rhs(u, du, mu, 0.0);
    fails += !(du[0] == 0.25 && std::fabs(du[1] - (-0.5 - 0.1 * 0.25)) < 1e-16);
    std::printf("%s\n", fails ? "FAILED" : "tape ok");
    return fails;
}
