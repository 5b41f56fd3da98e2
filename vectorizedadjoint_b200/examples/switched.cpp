// A recorded system that needs the wider operation surface of the tape (AADC idouble / ibool equivalents): a relay-driven,
// saturating oscillator with a piecewise right-hand side written with va::iIf on active values, erf / cbrt / atan2 / fmax
// terms and an explicit time dependence. Tape -> CUDA rhs/vjp -> NVRTC -> thread-per-trajectory kernels; the program checks
// the adjoint gradient against central finite differences of the forward map and prints "switched ok".
#include <boost/numeric/odeint.hpp>
#include <cmath>
#include <cstdio>
#include <iostream>

#include "lib.hpp"
#include "tape_systems.hpp"

using namespace boost::numeric::odeint;
using namespace vectorizedadjoint;

using tape_systems::Switched; // tape_systems.hpp: the same source the reference's AADC records for the golden fixtures

typedef runge_kutta4<std::vector<double>> fixed_type;

static double J_of(std::vector<double> p, std::vector<double> x0)
{
    Driver d(2, 1, 3);
    d.max_steps = 512;
    runge_kutta(fixed_type(), Switched(), x0, p, 0.0, 3.0, 0.01, d);
    return x0[0] - 0.5 * x0[1];
}

int main()
{
    const int N = 2, Npar = 3;
    const std::vector<double> p = {1.1, 0.4, 0.7}, x0 = {0.6, 0.3}; // the velocity changes sign during [0, 3]: both damper branches run
    int fails = 0;
    Driver driver(N, 1, Npar);
    driver.max_steps = 512;
    std::vector<double> x = x0;
    const size_t steps = runge_kutta(fixed_type(), Switched(), x, p, 0.0, 3.0, 0.01, driver);
    auto lambda = std::vector<std::vector<double>>(1, std::vector<double>{1.0, -0.5});
    auto mu = std::vector<std::vector<double>>(1, std::vector<double>(Npar, 0.0));
    setCostGradients(driver, lambda, mu);
    constructDriverButcherTableau(driver, fixed_type());
    recordDriverRHSFunction(driver, Switched());
    adjointSolve(driver, p);
    std::printf("rk4: %zu steps, x(tf) = [%.15g, %.15g]\n", steps, x[0], x[1]);
    bool vneg = false, vpos = false;
    for (int n = 0; n < driver.GetT(); ++n) {
        std::vector<double> u(N);
        driver.GetState(u, n);
        vneg |= u[1] < 0.0;
        vpos |= u[1] > 0.0;
    }
    fails += !(vneg && vpos);
    const double h = 1e-6;
    for (int k = 0; k < Npar; ++k) {
        std::vector<double> pp = p, pm = p;
        pp[k] += h;
        pm[k] -= h;
        const double fd = (J_of(pp, x0) - J_of(pm, x0)) / (2 * h);
        std::printf("  dJ/dp%d: adjoint %.12g  finite differences %.12g\n", k, mu[0][k], fd);
        fails += !(std::fabs(fd - mu[0][k]) <= 2e-6 * std::fabs(fd) + 1e-8);
    }
    for (int i = 0; i < N; ++i) {
        std::vector<double> xp = x0, xm = x0;
        xp[i] += h;
        xm[i] -= h;
        const double fd = (J_of(p, xp) - J_of(p, xm)) / (2 * h);
        std::printf("  dJ/dx%d(0): adjoint %.12g  finite differences %.12g\n", i, lambda[0][i], fd);
        fails += !(std::fabs(fd - lambda[0][i]) <= 2e-6 * std::fabs(fd) + 1e-8);
    }
    std::printf("%s\n", fails ? "switched FAILED" : "switched ok");
    return fails;
}
