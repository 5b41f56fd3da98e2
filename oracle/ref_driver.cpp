// ref_driver.cpp -- batch driver around the UNMODIFIED reference (lib/include + AADC), ORACLE / CPU BASELINE ONLY.
//
// Compiled by oracle/Makefile against /root/reference/lib/include, /root/reference/aadc/include and the Boost
// stand-in in oracle/shim into oracle/_ref/libva_ref.so (git-ignored). It drives the reference's own public API
// exactly the way its examples do (examples/*/main.cpp): Driver -> runge_kutta -> setCostGradients ->
// constructDriverButcherTableau -> recordDriverRHSFunction -> adjointSolve, once per parameter set (the reference has
// no batch API). One Driver per thread, RHS recorded once per thread (recording is
// serialised: AADC keeps process-global recording state), sweeps run concurrently.
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <boost/numeric/odeint.hpp>

#include "lib.hpp" // the reference's umbrella header (found through -I/root/reference/lib/include)
#include <aadc/ibool.h> // comparisons / iIf / max on the reference's active type (used by the recorded example systems)

#include "tape_systems.hpp" // vectorizedadjoint_b200/examples: the two systems the product serves through its tape -> CUDA path

using namespace vectorizedadjoint;
namespace odeint = boost::numeric::odeint;

namespace {

// User-side system functors, same maths and operation order as the reference examples' functors.
struct HarmonicSys {
    double k = 1.0;
    template <class T>
    void operator()(const std::vector<T> &r, std::vector<T> &drdt, const std::vector<T> &mu, const T) const
    {
        drdt[0] = r[1];
        drdt[1] = -k * r[0] - mu[0] * r[1];
    }
};
struct VanDerPolSys {
    template <class T>
    void operator()(const std::vector<T> &x, std::vector<T> &dxdt, const std::vector<T> &mu, const T) const
    {
        dxdt[0] = x[1];
        dxdt[1] = mu[0] * ((1.0 - x[0] * x[0]) * x[1] - x[0]);
    }
};
struct GlvSys {
    template <class T>
    void operator()(const std::vector<T> &x, std::vector<T> &dxdt, const std::vector<T> &p, T)
    {
        const int N = static_cast<int>(x.size());
        for (int i = 0; i < N; i++) {
            T sum = 0.0;
            for (int j = 0; j < N; j++) sum += p[N * (i + 1) + j] * x[j];
            dxdt[i] = x[i] * (p[i] + sum);
        }
    }
};

// autonomous variants of the two recorded example systems: the reference's reverse sweep evaluates every stage at t_n
// (detail/backpropagation.hpp:48,127), so only autonomous systems give reference gradients that are right
struct PendulumAutonomous : tape_systems::DrivenPendulum { PendulumAutonomous() { omega = 0.0; } };
struct SwitchedAutonomous : tape_systems::Switched { SwitchedAutonomous() { tscale = 0.0; } };

std::mutex g_record_mutex;

struct Job {
    int n, npar, nout, objective;
    double eps_abs, eps_rel, ti, tf, dt0;
    long b0, b1;
    const double *x0, *p;
    double *x_final, *lambda, *mu;
    int32_t *n_accept;
};

template <class Stepper, class System, bool Adaptive>
void run_range(const Job &j)
{
    typedef std::vector<double> state_type;
    Driver driver(j.n, j.nout, j.npar);
    Stepper stepper;
    System system;
    constructDriverButcherTableau(driver, stepper);
    {
        std::lock_guard<std::mutex> g(g_record_mutex);
        recordDriverRHSFunction(driver, system);
    }
    auto lambda = std::vector<std::vector<double>>(j.nout, std::vector<double>(j.n));
    auto mu = std::vector<std::vector<double>>(j.nout, std::vector<double>(j.npar));
    setCostGradients(driver, lambda, mu);
    state_type x(j.n), alphas(j.npar);
    for (long b = j.b0; b < j.b1; ++b) {
        std::memcpy(x.data(), j.x0 + (size_t)b * j.n, sizeof(double) * j.n);
        std::memcpy(alphas.data(), j.p + (size_t)b * j.npar, sizeof(double) * j.npar);
        size_t steps;
        if constexpr (Adaptive) {
            steps = runge_kutta(odeint::make_controlled<Stepper>(j.eps_abs, j.eps_rel), system, x, alphas, j.ti, j.tf, j.dt0, driver);
        } else {
            driver.p_states->Clear(); // the fixed-step overload never clears (detail/runge_kutta.hpp:38-72)
            steps = runge_kutta(stepper, system, x, alphas, j.ti, j.tf, j.dt0, driver);
        }
        if (j.n_accept) j.n_accept[b] = static_cast<int32_t>(steps);
        std::memcpy(j.x_final + (size_t)b * j.n, x.data(), sizeof(double) * j.n);
        for (int o = 0; o < j.nout; ++o) {
            double *lam_io = j.lambda + ((size_t)b * j.nout + o) * j.n;
            for (int i = 0; i < j.n; ++i) {
                if (j.objective == 1) lambda[o][i] = 1.0;
                else if (j.objective == 2) lambda[o][i] = x[i];
                else lambda[o][i] = lam_io[i];
            }
            std::fill(mu[o].begin(), mu[o].end(), 0.0);
        }
        adjointSolve(driver, alphas);
        for (int o = 0; o < j.nout; ++o) {
            std::memcpy(j.lambda + ((size_t)b * j.nout + o) * j.n, lambda[o].data(), sizeof(double) * j.n);
            std::memcpy(j.mu + ((size_t)b * j.nout + o) * j.npar, mu[o].data(), sizeof(double) * j.npar);
        }
    }
}

template <class System>
int dispatch_stepper(int stepper, const Job &j)
{
    typedef std::vector<double> S;
    switch (stepper) {
    case 0: run_range<odeint::euler<S>, System, false>(j); return 0;
    case 1: run_range<odeint::runge_kutta4<S>, System, false>(j); return 0;
    case 2: run_range<odeint::runge_kutta_cash_karp54<S>, System, true>(j); return 0;
    case 4: run_range<odeint::runge_kutta_fehlberg78<S>, System, true>(j); return 0;
    // error steppers used un-controlled: the stepper_tag overload (detail/runge_kutta.hpp:38-72) takes them through tag inheritance
    case 12: run_range<odeint::runge_kutta_cash_karp54<S>, System, false>(j); return 0;
    case 14: run_range<odeint::runge_kutta_fehlberg78<S>, System, false>(j); return 0;
    default: return -1; // dopri5 is not supported by the reference's ButcherTable (ButcherTable.hpp:247-250)
    }
}

} // namespace

extern "C" {

// sys: 0 harmonic, 1 van der pol, 2 GLV, 3 driven pendulum, 4 switched oscillator (tape_systems.hpp; n = 2, npar = 3),
//      5 / 6 the same two made autonomous (omega = 0 / tscale = 0), 7 harvested Lotka-Volterra (npar = n*n + n).
// stepper: 0 euler, 1 rk4 (fixed step), 2 ck54, 4 rkf78 (adaptive), 12 ck54, 14 rkf78 (fixed step).
// objective: 0 seeds given in lambda_inout, 1 seed = 1 (J = sum x_i(tf)), 2 seed = x(tf) (J = |x(tf)|^2/2).
// lambda_inout [B][nout][n], mu_out [B][nout][npar] (overwritten), x_final [B][n], n_accept [B] or NULL.
int va_ref_forward_adjoint_batch(int sys, int n, int npar, int nout, int stepper, double eps_abs, double eps_rel, long B,
                                 const double *x0, const double *p, double ti, double tf, double dt0, int objective,
                                 double *x_final, double *lambda_inout, double *mu_out, int32_t *n_accept, int threads)
{
    if (threads < 1) threads = 1;
    std::vector<Job> jobs;
    const long chunk = (B + threads - 1) / threads;
    for (int k = 0; k < threads; ++k) {
        long b0 = k * chunk, b1 = std::min(B, b0 + chunk);
        if (b0 >= B) break;
        jobs.push_back(Job{n, npar, nout, objective, eps_abs, eps_rel, ti, tf, dt0, b0, b1, x0, p, x_final, lambda_inout, mu_out, n_accept});
    }
    std::vector<int> rc(jobs.size(), 0);
    auto work = [&](size_t k) {
        switch (sys) {
        case 0: rc[k] = dispatch_stepper<HarmonicSys>(stepper, jobs[k]); break;
        case 1: rc[k] = dispatch_stepper<VanDerPolSys>(stepper, jobs[k]); break;
        case 2: rc[k] = dispatch_stepper<GlvSys>(stepper, jobs[k]); break;
        case 3: rc[k] = dispatch_stepper<tape_systems::DrivenPendulum>(stepper, jobs[k]); break;
        case 4: rc[k] = dispatch_stepper<tape_systems::Switched>(stepper, jobs[k]); break;
        case 5: rc[k] = dispatch_stepper<PendulumAutonomous>(stepper, jobs[k]); break;
        case 6: rc[k] = dispatch_stepper<SwitchedAutonomous>(stepper, jobs[k]); break;
        case 7: rc[k] = dispatch_stepper<tape_systems::HarvestedLotkaVolterra>(stepper, jobs[k]); break;
        default: rc[k] = -2;
        }
    };
    if (jobs.size() == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (size_t k = 0; k < jobs.size(); ++k) th.emplace_back(work, k);
        for (auto &t : th) t.join();
    }
    for (int r : rc)
        if (r) return r;
    return 0;
}

const char *va_ref_describe()
{
    return "unmodified reference lib/include + AADC AVX2 (libaadc.so) + Boost.Odeint stand-in (oracle/shim), g++ -O3 -mavx2";
}
}
