// tape_systems.hpp -- two ODE systems the engine has NO hand-written device code for, written once against the system concept
// of the reference (a functor templated on the scalar type, reference doc/source/harmonicOscillator.rst:85) so that the SAME
// source is recorded by this repo's tape (va::adouble -> CUDA rhs/vjp -> NVRTC) in examples/pendulum.cpp, examples/switched.cpp
// and the GPU tests, and by the reference's AADC (idouble) in oracle/ref_driver.cpp, which is where the golden fixtures for
// the tape path come from (tests/golden/make_goldens.py). Conditions on active values go through iIf, found by argument-
// dependent lookup for either active type (va::iIf / AADC's ::iIf) and below for plain double.
#ifndef VA_B200_TAPE_SYSTEMS_HPP
#define VA_B200_TAPE_SYSTEMS_HPP

#include <algorithm>
#include <cmath>
#include <vector>

namespace tape_systems
{

inline double iIf(bool c, double a, double b) { return c ? a : b; }

// damped pendulum with an explicitly time-dependent drive; omega is a member, not differentiated (like k in the reference's
// harmonic oscillator)
struct DrivenPendulum {
    double omega = 1.7;
    template <typename T>
    void operator()(const std::vector<T> &x, std::vector<T> &dxdt, const std::vector<T> &p, const T t) const
    {
        using namespace std;
        dxdt[0] = x[1];
        dxdt[1] = -p[0] * sin(x[0]) - p[1] * x[1] + p[2] * cos(omega * t) / (1.0 + x[0] * x[0]);
    }
};

// relay-driven, saturating oscillator: restoring force that saturates (erf), a one-sided damper (iIf on the velocity), a soft
// floor (max) and a drive whose phase depends on the state (atan2); piecewise right-hand side, explicit time dependence
struct Switched {
    double tscale = 1.0; // 0 makes the system autonomous (the reference's reverse sweep is only right for autonomous systems:
                         // it evaluates every stage at t_n, reference lib/include/detail/backpropagation.hpp:48,127)
    template <typename T>
    void operator()(const std::vector<T> &x, std::vector<T> &dxdt, const std::vector<T> &p, const T t) const
    {
        using namespace std;
        const T damper = iIf(x[1] > 0.0, p[1] * x[1], 0.25 * p[1] * x[1]);
        dxdt[0] = x[1];
        dxdt[1] = -p[0] * erf(x[0]) - damper + p[2] * cos(tscale * t + atan2(x[1], 1.0 + x[0] * x[0])) + 0.1 * cbrt(1.0 + x[0] * x[0]) * max(x[0], T(-0.05));
    }
};

// a Lotka-Volterra variant the engine has no built-in functor for: the reference's GLV right-hand side
// (examples/GeneralizedLotkaVolterra/main.cpp:105-119) plus a saturating harvest term. 16 species -> 272 parameters: wider than
// the per-lane register budget of the thread-per-trajectory kernels, so the parameters are read in place
struct HarvestedLotkaVolterra {
    template <typename T>
    void operator()(const std::vector<T> &x, std::vector<T> &dxdt, const std::vector<T> &alpha, T) const
    {
        const int N = (int)x.size();
        for (int i = 0; i < N; i++) {
            T sum = 0.0;
            for (int j = 0; j < N; j++) sum += alpha[N * (i + 1) + j] * x[j];
            dxdt[i] = x[i] * (alpha[i] + sum) - 0.05 * x[i] / (1.0 + x[i]);
        }
    }
};

} // namespace tape_systems

#endif
