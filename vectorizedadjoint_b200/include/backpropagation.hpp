// backpropagation.hpp -- adjointSolve / computeSensitivityMatrix[NoSIMD] of the reference
// (lib/include/backpropagation.hpp:18-87, recursion in detail/backpropagation.hpp:24-434), executed by the B200 engine.
// Contract kept: lambda[o] is OVERWRITTEN with dJ_o/dx(t0), mu[o] is INCREMENTED by dJ_o/dalpha
// (detail/backpropagation.hpp:313,319); precondition failures are reported on stdout and the call returns normally.
#ifndef VA_B200_BACKPROPAGATION_HPP
#define VA_B200_BACKPROPAGATION_HPP

#include "runge_kutta.hpp"

namespace vectorizedadjoint
{
namespace detail
{

inline void check_forward(Driver &driver)
{
    if (!driver.engine) throw std::runtime_error("Must call runge_kutta() first: there is no forward sweep to differentiate!");
    if (driver.p_butcher->stepper_id != driver.fwd_stepper)
        throw std::runtime_error("The Butcher tableau given to constructDriverButcherTableau() is not the stepper of the forward sweep!");
}

template <class State>
void adjointSolve(Driver &driver, const State &parameters)
{
    const int Nout = driver.GetNout(), Npar = driver.GetNpar(), Nin = driver.GetNin();
    (void)parameters; // the engine keeps the parameter set of the forward sweep
    check_forward(driver);
    std::vector<double> lam((size_t)Nout * Nin), mu((size_t)Nout * Npar);
    for (int o = 0; o < Nout; ++o)
        for (int i = 0; i < Nin; ++i) lam[(size_t)o * Nin + i] = (*driver.p_lambda)[o][i];
    va_batch_args a{};
    a.batch = 1; a.objective = VA_OBJ_SEED; a.reduce = VA_REDUCE_NONE; a.mem = VA_MEM_HOST; a.lambda = lam.data(); a.mu = mu.data();
    detail_runge_kutta::check(va_adjoint_batch(driver.engine.get(), &a), "va_adjoint_batch");
    for (int o = 0; o < Nout; ++o) {
        for (int i = 0; i < Nin; ++i) (*driver.p_lambda)[o][i] = lam[(size_t)o * Nin + i];
        for (int k = 0; k < Npar; ++k) (*driver.p_mu)[o][k] += mu[(size_t)o * Npar + k];
    }
}

// identity seeds -> jacobian[Nin][Npar] = d x(tf) / d alpha. One fused forward+adjoint call with Nin cost functions.
template <class State>
std::vector<std::vector<double>> computeSensitivityMatrix(Driver &driver, const State &parameters)
{
    const int Npar = driver.GetNpar(), Nin = driver.GetNin();
    check_forward(driver);
    va_engine_desc d{};
    d.system = driver.fwd_system; d.n_state = Nin; d.n_par = Npar; d.n_out = Nin; d.stepper = driver.fwd_stepper; d.adaptive = driver.fwd_adaptive;
    d.eps_abs = driver.fwd_eps_abs; d.eps_rel = driver.fwd_eps_rel; d.device = driver.device; d.max_steps = driver.max_steps;
    d.tape_cuda_src = driver.fwd_system == va::SYS_TAPE ? driver.fwd_tape_src.c_str() : nullptr;
    va_engine *e = nullptr;
    detail_runge_kutta::check(va_engine_create(&d, &e), "va_engine_create");
    std::unique_ptr<va_engine, EngineDeleter> guard(e);
    std::vector<double> lam((size_t)Nin * Nin, 0.0), mu((size_t)Nin * Npar), xf(Nin);
    for (int i = 0; i < Nin; ++i) lam[(size_t)i * Nin + i] = 1.0;
    va_batch_args a{};
    a.batch = 1; a.x0 = driver.fwd_x0.data(); a.params = parameters.data(); a.ti = driver.fwd_ti; a.tf = driver.fwd_tf; a.dt0 = driver.fwd_dt0;
    a.objective = VA_OBJ_SEED; a.reduce = VA_REDUCE_NONE; a.mem = VA_MEM_HOST; a.x_final = xf.data(); a.lambda = lam.data(); a.mu = mu.data();
    detail_runge_kutta::check(va_forward_adjoint_batch(e, &a), "va_forward_adjoint_batch");
    std::vector<std::vector<double>> jacobian(Nin, std::vector<double>(Npar, 0.0));
    for (int i = 0; i < Nin; ++i)
        for (int k = 0; k < Npar; ++k) jacobian[i][k] = mu[(size_t)i * Npar + k];
    return jacobian;
}

} // namespace detail

// Reverse sweep for every cost function registered with setCostGradients: on return lambda[o] = dJ_o/dx(t0) and
// mu[o] += dJ_o/dalpha
template <class State>
void adjointSolve(Driver &driver, const State &parameters)
{
    try {
        if (!(driver.p_lambda && driver.p_mu)) throw std::runtime_error("Must call setCostGradients() first!");
        if (!driver.p_butcher) throw std::runtime_error("Must call constructDriverButcherTableau() to set Butcher Tableau!");
        if (!driver.p_aad_data)
            throw std::runtime_error("Must call recordDriverRHSFunction() to record the RHS with automatic differentiation!");
        detail::adjointSolve(driver, parameters);
    } catch (std::exception &e) {
        std::cout << e.what() << std::endl;
    }
}

// d x(tf) / d alpha as an Nin x Npar matrix (one reverse sweep per state component in the reference; here all
// components are seeds of one batched reverse sweep)
template <class State>
auto computeSensitivityMatrix(Driver &driver, const State &parameters)
{
    try {
        if (!driver.p_butcher) throw std::runtime_error("Must call constructDriverButcherTableau() to set Butcher Tableau!");
        if (!driver.p_aad_data)
            throw std::runtime_error("Must call recordDriverRHSFunction() to record the RHS with automatic differentiation!");
        return detail::computeSensitivityMatrix(driver, parameters);
    } catch (std::exception &e) {
        std::cout << e.what() << std::endl;
        return std::vector<std::vector<double>>(0, std::vector<double>(0));
    }
}

template <class State>
auto computeSensitivityMatrixNoSIMD(Driver &driver, const State &parameters)
{
    return computeSensitivityMatrix(driver, parameters);
}

} // end namespace vectorizedadjoint

#endif
