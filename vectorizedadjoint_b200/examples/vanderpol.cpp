// Van der Pol oscillator, mu = 1e3 (the PETSc ex20adj problem the reference uses): d x(tf) / d mu with a controlled
// Fehlberg 7(8) stepper. usage: vanderpol <tolerance> [ck54|dopri5|rkf78]
#include <boost/numeric/odeint.hpp>
#include <iostream>
#include <string>

#include "lib.hpp"

using namespace boost::numeric::odeint;
using namespace vectorizedadjoint;

struct VdP {
    template <typename T>
    void operator()(const std::vector<T> &x, std::vector<T> &dxdt, const std::vector<T> &mu, const T) const
    {
        dxdt[0] = x[1];
        dxdt[1] = mu[0] * ((1.0 - x[0] * x[0]) * x[1] - x[0]);
    }
};

template <class Stepper>
int run(double tol)
{
    std::vector<double> mu = {1e3};
    std::vector<double> x0 = {2.0, -2.0 / 3.0 + 10.0 / (81.0 * mu[0]) - 292.0 / (2187.0 * mu[0] * mu[0])};
    std::cout << "Initial conditions: x0 = [" << x0[0] << ", " << x0[1] << "]" << std::endl;
    const int N = 2, Npar = 1;
    Stepper stepper;
    VdP vdp;
    Driver driver(N, N, Npar);
    driver.max_steps = 4096;
    const size_t steps = runge_kutta(make_controlled<Stepper>(tol, tol), vdp, x0, mu, 0.0, 0.5, 0.001, driver);
    std::cout << "Number of steps: " << steps << std::endl;
    std::cout << "Solution: x = [" << x0[0] << ", " << x0[1] << "]" << std::endl;
    auto lambda = std::vector<std::vector<double>>(N, std::vector<double>(N, 0.0));
    lambda[0][0] = 1.0;
    lambda[1][1] = 1.0;
    auto muadj = std::vector<std::vector<double>>(N, std::vector<double>(Npar, 0.0));
    setCostGradients(driver, lambda, muadj);
    constructDriverButcherTableau(driver, stepper);
    recordDriverRHSFunction(driver, vdp);
    adjointSolve(driver, mu);
    for (int i = 0; i < N; i++) {
        for (int j = 0; j < Npar; j++) std::cout << "mu[" << i << "][" << j << "] = " << muadj[i][j] << " ";
        std::cout << std::endl;
    }
    return 0;
}

int main(int argc, char *argv[])
{
    if (argc < 2) {
        std::cerr << "Usage: " << argv[0] << " <tolerance> [ck54|dopri5|rkf78]" << std::endl;
        return 1;
    }
    const double tol = std::stod(argv[1]);
    const std::string which = argc > 2 ? argv[2] : "rkf78";
    typedef std::vector<double> S;
    if (which == "ck54") return run<runge_kutta_cash_karp54<S>>(tol);
    if (which == "dopri5") return run<runge_kutta_dopri5<S>>(tol);
    return run<runge_kutta_fehlberg78<S>>(tol);
}
