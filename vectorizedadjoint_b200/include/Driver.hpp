// Driver.hpp -- the reference's Driver (lib/include/Driver.hpp:15-114) re-implemented on top of the B200 engine's C-ABI.
//
// Same public surface as the reference: Driver(Nin, Nout, Npar); GetNin/GetNout/GetNpar/GetT/GetTime/GetDt/GetState/Rhs;
// public p_lambda / p_mu (borrowed pointers to the caller's vectors, reference Driver.hpp:22-23, 110-111); the free
// functions constructDriverButcherTableau, recordDriverRHSFunction, setCostGradients, delete_driver_handle.
// What changed underneath: the checkpoint store (StateStorage) lives on the GPU and is mirrored to the host after the
// forward sweep; the Butcher tableau is a stepper id resolved inside the engine; the AADC recording is replaced by a tape
// (va_tape.hpp) that selects -- or generates -- the CUDA device functors for f and its vector-Jacobian product.
#ifndef VA_B200_DRIVER_HPP
#define VA_B200_DRIVER_HPP

#include <cassert>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <boost/numeric/odeint.hpp>

#include "va_engine.h"
#include "va_tape.hpp"

namespace odeint = boost::numeric::odeint;
typedef va::adouble idouble; // the active type the reference's examples name (AADC's idouble there)

namespace vectorizedadjoint
{

// what ButcherTable holds in the reference: here only which stepper, the coefficients live in the engine
struct ButcherTable {
    int stepper_id;
};

// what AadData holds in the reference: the recorded RHS
struct AadData {
    va::Tape tape;
    int system_kind;
};

// host mirror of the device checkpoint store (reference lib/include/StateStorage.hpp)
struct StateStorage {
    std::vector<double> time;
    std::vector<double> states; // [T+1][Nin]
    int nin = 0;
    void Clear() { time.clear(); states.clear(); }
    int GetT() const { return static_cast<int>(time.size()); }
};

struct EngineDeleter {
    void operator()(va_engine *e) const { va_engine_destroy(e); }
};

class Driver
{
  public:
    std::unique_ptr<AadData> p_aad_data;
    std::unique_ptr<StateStorage> p_states;
    std::unique_ptr<ButcherTable> p_butcher;

    std::vector<std::vector<double>> *p_lambda = nullptr;
    std::vector<std::vector<double>> *p_mu = nullptr;

    // engine state of the last forward sweep (B = 1: one Driver = one trajectory, as in the reference)
    std::unique_ptr<va_engine, EngineDeleter> engine;
    int fwd_system = -1, fwd_stepper = -1, fwd_adaptive = 0;
    double fwd_eps_abs = 0, fwd_eps_rel = 0, fwd_ti = 0, fwd_tf = 0, fwd_dt0 = 0;
    std::vector<double> fwd_x0;
    std::string fwd_tape_src; // CUDA source of the recorded functor (tape path)
    int device = 0;
    int max_steps = 0; // 0: engine default

  private:
    int m_Nin;
    int m_Nout;
    int m_Npar;

  public:
    Driver(int Nin, int Nout, int Npar) : p_states(std::make_unique<StateStorage>()), m_Nin(Nin), m_Nout(Nout), m_Npar(Npar) {}

    int GetNin() const { return m_Nin; }
    int GetNout() const { return m_Nout; }
    int GetNpar() const { return m_Npar; }

    template <typename State>
    void GetState(State &u, int n) const
    {
        for (size_t i = 0; i < u.size(); i++) u[i] = p_states->states[(size_t)n * m_Nin + i];
    }
    int GetT() const { return p_states->GetT(); }
    double GetDt(int n) const { return p_states->time[n + 1] - p_states->time[n]; }
    double GetTime(int n) const { return p_states->time[n]; }

    // f(u, alphas, time) from the recorded tape (the reference runs the AADC forward kernel, Driver.hpp:69-78)
    template <typename State, typename Time>
    void Rhs(const State &u, State &dudt, const State &alphas, const Time &time)
    {
        if (!p_aad_data) throw std::runtime_error("Must call recordDriverRHSFunction() to record the RHS with automatic differentiation!");
        std::vector<double> work;
        p_aad_data->tape.eval(u.data(), alphas.data(), static_cast<double>(time), dudt.data(), work);
    }
};

inline void delete_driver_handle(void *ptr)
{
    Driver *p_driver = static_cast<Driver *>(ptr);
    delete p_driver;
}

namespace detail
{
template <class Stepper>
void butcher_init(Driver &driver, Stepper, odeint::explicit_controlled_stepper_tag)
{
    // reference ButcherTable.hpp:28-36: message, tableau left unset
    std::cerr << "To create Driver, we need to supply an error stepper or a regular stepper, not a controlled stepper!" << std::endl;
    driver.p_butcher.reset();
}
template <class Stepper>
void butcher_init(Driver &driver, Stepper, odeint::stepper_tag)
{
    driver.p_butcher = std::make_unique<ButcherTable>(ButcherTable{Stepper::va_stepper_id});
}
} // namespace detail

template <class Stepper>
void constructDriverButcherTableau(Driver &driver, Stepper stepper)
{
    typedef typename odeint::unwrap_reference<Stepper>::type::stepper_category stepper_category;
    detail::butcher_init(driver, stepper, stepper_category());
}

template <typename System>
void recordDriverRHSFunction(Driver &driver, System system)
{
    auto data = std::make_unique<AadData>();
    data->tape = va::record(system, driver.GetNin(), driver.GetNpar());
    data->system_kind = va::identify(data->tape);
    driver.p_aad_data = std::move(data);
}

// Hands the driver the caller's seed vectors: lambda[o] = dJ_o/dx(tf) on entry, mu[o] = the accumulator dJ_o/dalpha is added to
// (borrowed pointers, reference Driver.hpp:103-114)
inline void setCostGradients(Driver &driver, std::vector<std::vector<double>> &lambda, std::vector<std::vector<double>> &mu)
{
    assert((int)lambda.size() == driver.GetNout());
    assert((int)lambda[0].size() == driver.GetNin());
    assert((int)mu.size() == driver.GetNout());
    assert((int)mu[0].size() == driver.GetNpar());
    driver.p_lambda = &lambda;
    driver.p_mu = &mu;
}

} // namespace vectorizedadjoint

#endif
