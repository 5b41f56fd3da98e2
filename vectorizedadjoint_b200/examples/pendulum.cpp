// A system the engine has NO built-in device code for: a damped pendulum with an explicitly time-dependent drive.
// recordDriverRHSFunction records the functor, the tape is turned into CUDA rhs/vjp device functors and compiled at
// run time (NVRTC) -- the B200 replacement of the reference's AADC JIT. The program checks the adjoint gradient against
// central finite differences of the forward map and prints "pendulum ok".
#include <boost/numeric/odeint.hpp>
#include <cmath>
#include <cstdio>
#include <iostream>

#include "lib.hpp"
#include "tape_systems.hpp"

using namespace boost::numeric::odeint;
using namespace vectorizedadjoint;

using tape_systems::DrivenPendulum; // tape_systems.hpp: the same source the reference's AADC records for the golden fixtures

typedef runge_kutta4<std::vector<double>> fixed_type;
typedef runge_kutta_dopri5<std::vector<double>> err_type;

static std::vector<double> forward_fixed(std::vector<double> p)
{
    Driver d(2, 1, 3);
    d.max_steps = 512;
    std::vector<double> x = {0.4, -0.2};
    runge_kutta(fixed_type(), DrivenPendulum(), x, p, 0.0, 2.0, 0.01, d);
    return x;
}

int main()
{
    const int N = 2, Npar = 3;
    std::vector<double> p = {1.3, 0.15, 0.8};
    int fails = 0;
    // fixed-step RK4: J = x_0(tf) + 2 x_1(tf)
    {
        Driver driver(N, 1, Npar);
        driver.max_steps = 512;
        std::vector<double> x = {0.4, -0.2};
        const size_t steps = runge_kutta(fixed_type(), DrivenPendulum(), x, p, 0.0, 2.0, 0.01, driver);
        auto lambda = std::vector<std::vector<double>>(1, std::vector<double>{1.0, 2.0});
        auto mu = std::vector<std::vector<double>>(1, std::vector<double>(Npar, 0.0));
        setCostGradients(driver, lambda, mu);
        constructDriverButcherTableau(driver, fixed_type());
        recordDriverRHSFunction(driver, DrivenPendulum());
        adjointSolve(driver, p);
        std::printf("rk4: %zu steps, x(tf) = [%.15g, %.15g]\n", steps, x[0], x[1]);
        for (int k = 0; k < Npar; ++k) {
            const double h = 1e-6;
            std::vector<double> pp = p, pm = p;
            pp[k] += h;
            pm[k] -= h;
            const std::vector<double> xp = forward_fixed(pp), xm = forward_fixed(pm);
            const double fd = ((xp[0] + 2 * xp[1]) - (xm[0] + 2 * xm[1])) / (2 * h);
            std::printf("  dJ/dp%d: adjoint %.12g  finite differences %.12g\n", k, mu[0][k], fd);
            fails += !(std::fabs(fd - mu[0][k]) <= 1e-7 * std::fabs(fd) + 1e-9);
        }
        // the stored trajectory is available like in the reference
        fails += !(driver.GetT() == (int)steps + 1 && driver.GetTime(0) == 0.0);
    }
    // controlled Dormand-Prince: full sensitivity matrix d x(tf) / d p
    {
        Driver driver(N, N, Npar);
        std::vector<double> x = {0.4, -0.2};
        const size_t steps = runge_kutta(make_controlled<err_type>(1e-10, 1e-10), DrivenPendulum(), x, p, 0.0, 2.0, 0.01, driver);
        constructDriverButcherTableau(driver, err_type());
        recordDriverRHSFunction(driver, DrivenPendulum());
        auto jac = computeSensitivityMatrix(driver, p);
        std::printf("dopri5: %zu steps, d x(tf)/d p = [[%.10g, %.10g, %.10g], [%.10g, %.10g, %.10g]]\n", steps, jac[0][0], jac[0][1], jac[0][2],
                    jac[1][0], jac[1][1], jac[1][2]);
        fails += !(jac.size() == 2 && std::isfinite(jac[1][2]) && steps > 5);
    }
    std::printf("%s\n", fails ? "pendulum FAILED" : "pendulum ok");
    return fails;
}
