// Several GPUs from C++: the Lotka-Volterra batch of batch_lotka.cpp sharded over G GPUs inside ONE call (BatchDriver with a
// device list -> va_engine_desc.devices), summed-objective gradient combined by the library's single ncclAllReduce. The reference
// has nothing of the kind (its AAD workspace is single-threaded: reference lib/include/AadData.hpp:32). Checked against the same
// batch on one GPU: per-set results bit-identical, summed gradient equal to summation-order round-off. Prints "multi_lotka ok".
//   multi_lotka [G = 2] [B = 2000] [N = 16]
#include <boost/numeric/odeint.hpp>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "BatchDriver.hpp"

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
    const int G = argc > 1 ? std::atoi(argv[1]) : 2, B = argc > 2 ? std::atoi(argv[2]) : 2000, N = argc > 3 ? std::atoi(argv[3]) : 16;
    const int Npar = N * N + N;
    const double tol = 1e-8, ti = 0.0, tf = 10.0, dt = 1e-3;
    std::mt19937_64 rng(11);
    std::normal_distribution<double> nd;
    std::uniform_real_distribution<double> ud(-1.0, 1.0);
    std::vector<double> alphas((size_t)B * Npar), x0((size_t)B * N, 0.1);
    for (int b = 0; b < B; ++b) {
        double *a = &alphas[(size_t)b * Npar];
        for (int i = 0; i < N; ++i) a[i] = 0.1 * (1.0 + 0.1 * ud(rng));
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) a[N * (i + 1) + j] = i == j ? -10.0 * (1.0 + 0.1 * ud(rng)) : (ud(rng) > 0 ? nd(rng) * std::sqrt(10.0 / N) : 0.0);
    }
    auto seeds = [&] { return std::vector<double>((size_t)B * N, 1.0); }; // J = sum_i x_i(tf)

    std::vector<double> x1 = x0, lam1 = seeds(), mu1, xs1 = x0, lams1 = seeds(), mus1;
    std::vector<int> st1;
    {
        BatchDriver one(make_controlled<err_type>(tol, tol), LotkaVolterra(), N, 1, Npar, 0);
        st1 = one.forwardAdjoint(x1, alphas, ti, tf, dt, lam1, mu1);
        one.forwardAdjointSummed(xs1, alphas, ti, tf, dt, lams1, mus1);
    }
    std::vector<int> devices;
    for (int g = 0; g < G; ++g) devices.push_back(g);
    std::vector<double> xg = x0, lamg = seeds(), mug, xsg = x0, lamsg = seeds(), musg;
    std::vector<int> stg;
    try {
        BatchDriver many(make_controlled<err_type>(tol, tol), LotkaVolterra(), N, 1, Npar, devices);
        stg = many.forwardAdjoint(xg, alphas, ti, tf, dt, lamg, mug);
        many.forwardAdjointSummed(xsg, alphas, ti, tf, dt, lamsg, musg);
        va_engine_info info;
        va_engine_get_info(many.handle(), &info);
        std::printf("%d GPUs (%s), %d parameter sets, N = %d: NCCL %d, %lld all-reduce call(s) on %d ranks\n", info.n_devices, info.device_name, B, N,
                    info.nccl_version, (long long)info.collectives, info.comm_world);
    } catch (const std::exception &ex) {
        std::printf("multi_lotka: cannot use %d GPUs here: %s\n", G, ex.what());
        return 77;
    }
    int fails = 0;
    fails += stg != st1;
    fails += xg != x1;     // bit-identical per parameter set, wherever it ran
    fails += lamg != lam1;
    fails += mug != mu1;
    double scale = 0.0, worst = 0.0;
    for (double v : mus1) scale = std::fmax(scale, std::fabs(v));
    for (size_t k = 0; k < mus1.size(); ++k) worst = std::fmax(worst, std::fabs(musg[k] - mus1[k]) / scale);
    std::printf("summed gradient, %d GPUs vs 1 GPU: worst relative difference %.3e\n", G, worst);
    fails += !(worst < 1e-12);
    std::printf("%s\n", fails ? "multi_lotka FAILED" : "multi_lotka ok");
    return fails;
}
