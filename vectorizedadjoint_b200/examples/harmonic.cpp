// Client of the drop-in API: energy gradient of a damped harmonic oscillator.
//   r'' = -k r - mu r',  r(0) = 0, r'(0) = 1,  E = |(r, r')(tf)|^2 / 2,  fixed-step RK4 with dt = 0.01 up to tf = 10.
// Prints the lines the reference's HarmonicOscillator example prints (tests/test_gpu_dropin_examples.py compares them with
// the reference's own output) and then verifies the three adjoint derivatives against central finite differences.
#include <boost/numeric/odeint.hpp>
#include <cmath>
#include <cstdio>
#include <iostream>

#include "lib.hpp"

namespace ode = boost::numeric::odeint;
namespace vad = vectorizedadjoint;

struct DampedSpring {
    double k = 1.0; // not differentiated
    template <typename T>
    void operator()(const std::vector<T> &u, std::vector<T> &du, const std::vector<T> &damping, const T) const
    {
        du[0] = u[1];
        du[1] = -k * u[0] - damping[0] * u[1];
    }
};

typedef ode::runge_kutta4<std::vector<double>> rk4_type;
static const double kT0 = 0.0, kT1 = 10.0, kDt = 0.01;

// E(tf) for given initial state and damping (one forward sweep on the device)
static double energy(double r0, double v0, double mu)
{
    vad::Driver d(2, 1, 1);
    d.max_steps = 1024;
    std::vector<double> u = {r0, v0}, par = {mu};
    vad::runge_kutta(rk4_type(), DampedSpring(), u, par, kT0, kT1, kDt, d);
    return 0.5 * (u[0] * u[0] + u[1] * u[1]);
}

int main()
{
    const double mu0 = 0.151, r0 = 0.0, v0 = 1.0;
    vad::Driver driver(/*Nin*/ 2, /*Nout*/ 1, /*Npar*/ 1);
    driver.max_steps = 1024;
    std::vector<double> state = {r0, v0}, damping = {mu0};
    const size_t n_steps = vad::runge_kutta(rk4_type(), DampedSpring(), state, damping, kT0, kT1, kDt, driver);
    std::cout << "Number of steps: " << n_steps << std::endl;
    std::cout << "Solution: r = [" << state[0] << ", " << state[1] << "]" << std::endl;

    // dE/d(state at tf) = state at tf: the seed of the reverse sweep; mu accumulates dE/d(damping)
    std::vector<std::vector<double>> seed(1, state), dmu(1, std::vector<double>(1, 0.0));
    vad::setCostGradients(driver, seed, dmu);
    vad::constructDriverButcherTableau(driver, rk4_type());
    vad::recordDriverRHSFunction(driver, DampedSpring());
    vad::adjointSolve(driver, damping);
    std::cout << "dEdmr:" << seed[0][0] << std::endl;
    std::cout << "dEdmv:" << seed[0][1] << std::endl;
    std::cout << "dEdmu:" << dmu[0][0] << std::endl;

    // central finite differences of the forward map
    const double h = 1e-6;
    const double fd[3] = {(energy(r0 + h, v0, mu0) - energy(r0 - h, v0, mu0)) / (2 * h), (energy(r0, v0 + h, mu0) - energy(r0, v0 - h, mu0)) / (2 * h),
                          (energy(r0, v0, mu0 + h) - energy(r0, v0, mu0 - h)) / (2 * h)};
    const double ad[3] = {seed[0][0], seed[0][1], dmu[0][0]};
    int bad = 0;
    for (int i = 0; i < 3; ++i) {
        std::printf("check %d: adjoint % .10e  finite differences % .10e\n", i, ad[i], fd[i]);
        bad += !(std::fabs(ad[i] - fd[i]) <= 1e-7 * std::fabs(fd[i]) + 1e-10);
    }
    std::printf("%s\n", bad ? "harmonic FAILED" : "harmonic ok");
    return bad;
}
