// Client of the drop-in API: sensitivity of the Van der Pol oscillator's final state to its stiffness parameter.
//   x0' = x1,  x1' = mu ((1 - x0^2) x1 - x0),  mu = 1000 (stiff regime; the initial point lies on the limit cycle),
//   t in [0, 0.5], controlled stepper with atol = rtol = <tolerance>.
// usage: vanderpol <tolerance> [rkf78|ck54|dopri5]      (default: Fehlberg 7(8), the stepper of the reference's example)
// Prints step count, x(tf) and d x_i(tf) / d mu in the reference example's format (compared digit for digit with the
// reference's output in tests/test_gpu_dropin_examples.py), then the sensitivity from computeSensitivityMatrix as a cross-check.
#include <boost/numeric/odeint.hpp>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <string>

#include "lib.hpp"

namespace ode = boost::numeric::odeint;
namespace vad = vectorizedadjoint;
typedef std::vector<double> vec;

struct VanDerPolRhs {
    template <typename T>
    void operator()(const std::vector<T> &x, std::vector<T> &f, const std::vector<T> &mu, const T) const
    {
        f[0] = x[1];
        f[1] = mu[0] * ((1.0 - x[0] * x[0]) * x[1] - x[0]);
    }
};

static vec limit_cycle_start(double mu) { return {2.0, -2.0 / 3.0 + 10.0 / (81.0 * mu) - 292.0 / (2187.0 * mu * mu)}; }

template <class ErrorStepper>
static int sensitivities(double tol)
{
    const vec mu = {1e3};
    vec x = limit_cycle_start(mu[0]);
    std::cout << "Initial conditions: x0 = [" << x[0] << ", " << x[1] << "]" << std::endl;

    vad::Driver driver(2, 2, 1); // two cost functions: J_0 = x_0(tf), J_1 = x_1(tf)
    driver.max_steps = 4096;
    const size_t accepted = vad::runge_kutta(ode::make_controlled<ErrorStepper>(tol, tol), VanDerPolRhs(), x, mu, 0.0, 0.5, 0.001, driver);
    std::cout << "Number of steps: " << accepted << std::endl;
    std::cout << "Solution: x = [" << x[0] << ", " << x[1] << "]" << std::endl;

    std::vector<vec> seeds = {{1.0, 0.0}, {0.0, 1.0}}, dmu(2, vec(1, 0.0));
    vad::setCostGradients(driver, seeds, dmu);
    vad::constructDriverButcherTableau(driver, ErrorStepper());
    vad::recordDriverRHSFunction(driver, VanDerPolRhs());
    vad::adjointSolve(driver, mu);
    for (size_t i = 0; i < dmu.size(); i++) std::cout << "mu[" << i << "][0] = " << dmu[i][0] << " " << std::endl;

    // the same numbers through computeSensitivityMatrix (identity seeds built by the library)
    const std::vector<vec> jac = vad::computeSensitivityMatrix(driver, mu);
    int bad = 0;
    for (size_t i = 0; i < jac.size(); i++) bad += !(std::fabs(jac[i][0] - dmu[i][0]) <= 1e-12 * std::fabs(dmu[i][0]));
    std::printf("sensitivity matrix: [%.10g, %.10g] %s\n", jac[0][0], jac[1][0], bad ? "MISMATCH" : "(matches)");
    return bad;
}

int main(int argc, char *argv[])
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <tolerance> [rkf78|ck54|dopri5]\n", argv[0]);
        return 1;
    }
    const double tol = std::stod(argv[1]);
    const std::string stepper = argc > 2 ? argv[2] : "rkf78";
    if (stepper == "ck54") return sensitivities<ode::runge_kutta_cash_karp54<vec>>(tol);
    if (stepper == "dopri5") return sensitivities<ode::runge_kutta_dopri5<vec>>(tol);
    return sensitivities<ode::runge_kutta_fehlberg78<vec>>(tol);
}
