// Generalized Lotka-Volterra, N species: sensitivities of x(tf) w.r.t. x0 and all N^2+N parameters (N cost functions,
// identity seeds) with a controlled Cash-Karp 5(4) stepper. usage: lotka <tolerance> <N> [seed]
// Parameters are the seeded synthetic sets of the benchmark (no data files needed); prints timings like the reference's
// example and, in addition, checksums of the results.
#include <boost/numeric/odeint.hpp>
#include <chrono>
#include <cmath>
#include <iostream>
#include <random>

#include "lib.hpp"

using namespace boost::numeric::odeint;
using namespace vectorizedadjoint;

struct GLV {
    template <typename T>
    void operator()(const std::vector<T> &x, std::vector<T> &dxdt, const std::vector<T> &par, T)
    {
        const int N = static_cast<int>(x.size());
        for (int i = 0; i < N; i++) {
            T s = 0.0;
            for (int j = 0; j < N; j++) s += par[N * (i + 1) + j] * x[j];
            dxdt[i] = x[i] * (par[i] + s);
        }
    }
};

int main(int argc, char *argv[])
{
    if (argc < 3) {
        std::cerr << "Usage: " << argv[0] << " <tolerance> <N> [seed]" << std::endl;
        return 1;
    }
    const double tol = std::stod(argv[1]);
    const int N = std::stoi(argv[2]);
    const unsigned seed = argc > 3 ? std::stoul(argv[3]) : 1234u;
    std::cout << "Tolerance: " << tol << std::endl << "N: " << N << std::endl;
    const int Npar = N * N + N;
    std::mt19937_64 gen(seed);
    std::normal_distribution<double> z(0.0, 1.0);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    std::vector<double> alphas(Npar), x0(N, 0.1);
    for (int i = 0; i < N; i++) alphas[i] = 0.1;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) alphas[N * (i + 1) + j] = (i == j) ? -10.0 : (u(gen) < 0.5 ? z(gen) * std::sqrt(10.0 / N) : 0.0);
    typedef runge_kutta_cash_karp54<std::vector<double>> stepper_type;
    stepper_type stepper;
    GLV glv;
    Driver driver(N, N, Npar);
    auto t0 = std::chrono::high_resolution_clock::now();
    const size_t steps = runge_kutta(make_controlled<stepper_type>(tol, tol), glv, x0, alphas, 0.0, 10.0, 1e-3, driver);
    auto t1 = std::chrono::high_resolution_clock::now();
    auto lambda = std::vector<std::vector<double>>(N, std::vector<double>(N, 0.0));
    auto muD = std::vector<std::vector<double>>(N, std::vector<double>(Npar, 0.0));
    for (int i = 0; i < N; i++) lambda[i][i] = 1.0;
    constructDriverButcherTableau(driver, stepper);
    recordDriverRHSFunction(driver, glv);
    setCostGradients(driver, lambda, muD);
    auto t2 = std::chrono::high_resolution_clock::now();
    adjointSolve(driver, alphas);
    auto t3 = std::chrono::high_resolution_clock::now();
    double sx = 0, sl = 0, sm = 0;
    for (int i = 0; i < N; i++) {
        sx += x0[i];
        for (int j = 0; j < N; j++) sl += lambda[i][j];
        for (int k = 0; k < Npar; k++) sm += muD[i][k];
    }
    std::cout << "Number of steps: " << steps << std::endl;
    std::cout << "Time forward integration: " << std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count() << " microseconds" << std::endl;
    std::cout << "Time adjoint integration: " << std::chrono::duration_cast<std::chrono::microseconds>(t3 - t2).count() << " microseconds" << std::endl;
    std::cout.precision(15);
    std::cout << "checksums: sum x(tf) = " << sx << ", sum dx(tf)/dx0 = " << sl << ", sum dx(tf)/dalpha = " << sm << std::endl;
    return 0;
}
