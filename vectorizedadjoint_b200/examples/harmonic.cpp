// Damped harmonic oscillator: dE/dx0, dE/dv0, dE/dmu of E = |r(tf)|^2 / 2 with fixed-step RK4 (dt = 0.01, tf = 10).
// Client of the drop-in API; prints what the reference's examples/HarmonicOscillator prints.
#include <boost/numeric/odeint.hpp>
#include <iostream>

#include "lib.hpp"

using namespace boost::numeric::odeint;
using namespace vectorizedadjoint;

struct Oscillator {
    double stiffness = 1.0;
    template <typename T>
    void operator()(const std::vector<T> &r, std::vector<T> &drdt, const std::vector<T> &mu, const T) const
    {
        drdt[0] = r[1];
        drdt[1] = -stiffness * r[0] - mu[0] * r[1];
    }
};

int main()
{
    typedef runge_kutta4<std::vector<double>> stepper_type;
    const int Nin = 2, Nout = 1, Npar = 1;
    Driver driver(Nin, Nout, Npar);
    driver.max_steps = 1024;
    std::vector<double> mu = {0.151}, r = {0.0, 1.0};
    stepper_type stepper;
    Oscillator sys;
    const size_t steps = runge_kutta(stepper, sys, r, mu, 0.0, 10.0, 0.01, driver);
    std::cout << "Number of steps: " << steps << std::endl;
    std::cout << "Solution: r = [" << r[0] << ", " << r[1] << "]" << std::endl;
    auto lambda = std::vector<std::vector<double>>(Nout, std::vector<double>(Nin));
    lambda[0][0] = r[0];
    lambda[0][1] = r[1];
    auto muadj = std::vector<std::vector<double>>(Nout, std::vector<double>(Npar, 0.0));
    setCostGradients(driver, lambda, muadj);
    constructDriverButcherTableau(driver, stepper);
    recordDriverRHSFunction(driver, sys);
    adjointSolve(driver, mu);
    std::cout << "dEdmr:" << lambda[0][0] << std::endl;
    std::cout << "dEdmv:" << lambda[0][1] << std::endl;
    std::cout << "dEdmu:" << muadj[0][0] << std::endl;
    return 0;
}
