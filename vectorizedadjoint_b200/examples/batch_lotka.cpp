// Batched use of the engine from C++: the reference's Lotka-Volterra example (examples/GeneralizedLotkaVolterra/main.cpp:105-119,
// same functor, same stepper, J_i = x_i(tf) for every species: Nout = N, i.e. the full sensitivity matrix) for B parameter sets in
// ONE call, checked against B single-trajectory Driver runs of the drop-in API. Prints "batch_lotka ok".
#include <boost/numeric/odeint.hpp>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <random>

#include "BatchDriver.hpp"
#include "lib.hpp"

using namespace boost::numeric::odeint;
using namespace vectorizedadjoint;

struct LotkaVolterra {
    template <typename T>
    void operator()(const std::vector<T> &x, std::vector<T> &dxdt, const std::vector<T> &alpha, T)
    {
        const int N = (int)x.size();
        for (int i = 0; i < N; i++) {
            T sum = 0.0;
            for (int j = 0; j < N; j++) sum += alpha[N * (i + 1) + j] * x[j];
            dxdt[i] = x[i] * (alpha[i] + sum);
        }
    }
};
typedef runge_kutta_cash_karp54<std::vector<double>> err_type;

int main(int argc, char **argv)
{
    const int N = 8, Npar = N * N + N, B = argc > 1 ? std::atoi(argv[1]) : 300;
    const double tol = 1e-8, ti = 0.0, tf = 10.0, dt = 1e-3;
    std::mt19937_64 rng(7);
    std::normal_distribution<double> nd;
    std::uniform_real_distribution<double> ud(-1.0, 1.0);
    std::vector<double> alphas((size_t)B * Npar), x0((size_t)B * N, 0.1);
    for (int b = 0; b < B; ++b) {
        double *a = &alphas[(size_t)b * Npar];
        for (int i = 0; i < N; ++i) a[i] = 0.1 * (1.0 + 0.1 * ud(rng));
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) a[N * (i + 1) + j] = i == j ? -10.0 * (1.0 + 0.1 * ud(rng)) : (ud(rng) > 0 ? nd(rng) * std::sqrt(10.0 / N) : 0.0);
    }
    // B parameter sets in one call: identity seeds -> lambda = d x(tf) / d x(ti), mu = d x(tf) / d alpha
    BatchDriver batch(make_controlled<err_type>(tol, tol), LotkaVolterra(), N, N, Npar);
    std::vector<double> lambda((size_t)B * N * N, 0.0), mu, xb = x0;
    for (int b = 0; b < B; ++b)
        for (int o = 0; o < N; ++o) lambda[((size_t)b * N + o) * N + o] = 1.0;
    const auto t0 = std::chrono::steady_clock::now();
    const std::vector<int> steps = batch.forwardAdjoint(xb, alphas, ti, tf, dt, lambda, mu);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::printf("batch of %d parameter sets, N = %d, Nout = %d: %.2f ms, steps %d..%d\n", B, N, N, ms, *std::min_element(steps.begin(), steps.end()),
                *std::max_element(steps.begin(), steps.end()));
    // the same through the single-trajectory drop-in API, for a few of them
    int fails = 0;
    double worst = 0.0;
    for (int b : {0, 1, B / 2, B - 1}) {
        Driver driver(N, N, Npar);
        std::vector<double> x(x0.begin() + (size_t)b * N, x0.begin() + (size_t)(b + 1) * N);
        const std::vector<double> al(alphas.begin() + (size_t)b * Npar, alphas.begin() + (size_t)(b + 1) * Npar);
        const size_t st = runge_kutta(make_controlled<err_type>(tol, tol), LotkaVolterra(), x, al, ti, tf, dt, driver);
        constructDriverButcherTableau(driver, err_type());
        recordDriverRHSFunction(driver, LotkaVolterra());
        const auto jac = computeSensitivityMatrix(driver, al);
        fails += (int)st != steps[(size_t)b];
        for (int i = 0; i < N; ++i) fails += !(x[i] == xb[(size_t)b * N + i]);
        double scale = 0.0;
        for (int o = 0; o < N; ++o)
            for (int k = 0; k < Npar; ++k) scale = std::fmax(scale, std::fabs(jac[o][k]));
        for (int o = 0; o < N; ++o)
            for (int k = 0; k < Npar; ++k) worst = std::fmax(worst, std::fabs(jac[o][k] - mu[((size_t)b * N + o) * Npar + k]) / scale);
    }
    std::printf("sensitivities vs single-trajectory Driver: worst relative difference %.3e\n", worst);
    fails += !(worst < 1e-11);
    std::printf("%s\n", fails ? "batch_lotka FAILED" : "batch_lotka ok");
    return fails;
}
