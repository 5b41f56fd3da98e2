// runge_kutta.hpp -- vectorizedadjoint::runge_kutta(stepper, system, x0, alphas, ti, tf, dt, driver[, observer]),
// the forward sweep entry of the reference (lib/include/runge_kutta.hpp:17-59, loops in detail/runge_kutta.hpp:38-118),
// executed by the B200 engine (va_forward_batch, B = 1). x0 is overwritten with x(tf); every accepted (t_n, x_n) is kept
// for the reverse sweep and mirrored into the Driver; the return value is the number of accepted steps.
#ifndef VA_B200_RUNGE_KUTTA_HPP
#define VA_B200_RUNGE_KUTTA_HPP

#include "Driver.hpp"

namespace vectorizedadjoint
{
namespace detail_runge_kutta
{

inline void check(int rc, const char *what)
{
    if (rc != VA_OK) throw std::runtime_error(std::string(what) + ": " + va_last_error());
}

template <class System, class State, class Time, class Observer>
size_t run_forward(int stepper_id, int adaptive, double eps_abs, double eps_rel, System system, State &start_state, const State &alphas,
                   Time start_time, const Time end_time, Time dt, Driver &driver, Observer observer)
{
    const int n = driver.GetNin(), npar = driver.GetNpar();
    // which device functor: record the functor once (cheap, B = 1) unless the caller already did
    int kind;
    std::string tape_src;
    if (driver.p_aad_data) {
        kind = driver.p_aad_data->system_kind;
        if (kind == va::SYS_TAPE) tape_src = driver.p_aad_data->tape.cuda_source("VaUserSys");
    } else {
        const va::Tape tape = va::record(system, n, npar);
        kind = va::identify(tape);
        if (kind == va::SYS_TAPE) tape_src = tape.cuda_source("VaUserSys");
    }
    // a recorded system gets its own run-time compiled kernels (tape -> CUDA -> NVRTC), never reused across functors
    const bool reuse = kind != va::SYS_TAPE && driver.engine && driver.fwd_system == kind && driver.fwd_stepper == stepper_id && driver.fwd_adaptive == adaptive &&
                       driver.fwd_eps_abs == eps_abs && driver.fwd_eps_rel == eps_rel;
    if (!reuse) {
        va_engine_desc d{};
        d.system = kind; d.n_state = n; d.n_par = npar; d.n_out = driver.GetNout(); d.stepper = stepper_id; d.adaptive = adaptive;
        d.eps_abs = eps_abs; d.eps_rel = eps_rel; d.device = driver.device; d.max_steps = driver.max_steps;
        d.tape_cuda_src = kind == va::SYS_TAPE ? tape_src.c_str() : nullptr;
        va_engine *e = nullptr;
        check(va_engine_create(&d, &e), "va_engine_create");
        driver.engine.reset(e);
        driver.fwd_tape_src = tape_src;
        driver.fwd_system = kind; driver.fwd_stepper = stepper_id; driver.fwd_adaptive = adaptive;
        driver.fwd_eps_abs = eps_abs; driver.fwd_eps_rel = eps_rel;
    }
    driver.fwd_ti = start_time; driver.fwd_tf = end_time; driver.fwd_dt0 = dt;
    driver.fwd_x0.assign(start_state.begin(), start_state.end());

    std::vector<double> x_final(n);
    int32_t n_accept = 0, n_reject = 0, status = 0;
    va_batch_args a{};
    a.batch = 1; a.x0 = driver.fwd_x0.data(); a.params = alphas.data(); a.ti = start_time; a.tf = end_time; a.dt0 = dt;
    a.objective = VA_OBJ_SEED; a.reduce = VA_REDUCE_NONE; a.mem = VA_MEM_HOST; a.x_final = x_final.data();
    a.n_accept = &n_accept; a.n_reject = &n_reject; a.status = &status;
    check(va_forward_batch(driver.engine.get(), &a), "va_forward_batch");
    if (status & VA_TRAJ_NO_PROGRESS) // odeint::no_progress_error, thrown by failed_step_checker in the reference
        throw std::runtime_error("Max number of iterations exceeded (500). A new step size was not found.");
    if (status & VA_TRAJ_CKPT_OVERFLOW)
        throw std::runtime_error("checkpoint capacity exceeded: raise Driver::max_steps");

    // host mirror of the checkpoints (Driver::GetT / GetTime / GetState)
    StateStorage &st = *driver.p_states;
    st.Clear();
    st.nin = n;
    int32_t count = 0;
    check(va_get_checkpoints(driver.engine.get(), 0, 0, nullptr, nullptr, &count), "va_get_checkpoints");
    st.time.resize(count);
    st.states.resize((size_t)count * n);
    check(va_get_checkpoints(driver.engine.get(), 0, count, st.time.data(), st.states.data(), &count), "va_get_checkpoints");

    // the observer sees the same sequence of (x, t) calls as in the reference loop, after the sweep
    State xs(n);
    for (int k = 0; k < count; ++k) {
        for (int i = 0; i < n; ++i) xs[i] = st.states[(size_t)k * n + i];
        observer(xs, static_cast<Time>(st.time[k]));
    }
    for (int i = 0; i < n; ++i) start_state[i] = x_final[i];
    return static_cast<size_t>(n_accept);
}

// tag dispatch: fixed-step steppers (stepper_tag)
template <class Stepper, class System, class State, class Time, class Observer>
size_t runge_kutta(Stepper, System system, State &start_state, const State &alphas, Time start_time, const Time end_time, Time dt,
                   Driver &driver, Observer observer, odeint::stepper_tag)
{
    return run_forward(Stepper::va_stepper_id, 0, 0.0, 0.0, system, start_state, alphas, start_time, end_time, dt, driver, observer);
}

// tag dispatch: make_controlled<...> steppers (controlled_stepper_tag)
template <class Stepper, class System, class State, class Time, class Observer>
size_t runge_kutta(Stepper stepper, System system, State &start_state, const State &alphas, Time start_time, const Time end_time, Time dt,
                   Driver &driver, Observer observer, odeint::controlled_stepper_tag)
{
    return run_forward(Stepper::va_stepper_id, 1, stepper.eps_abs, stepper.eps_rel, system, start_state, alphas, start_time, end_time, dt,
                       driver, observer);
}

} // namespace detail_runge_kutta

// public entry, observer called with every accepted (x, t)
template <class Stepper, class System, class State, class Time, class Observer>
size_t runge_kutta(Stepper stepper, System system, State &start_state, const State &alphas, Time start_time, const Time end_time, Time dt,
                   Driver &driver, Observer observer)
{
    typedef typename odeint::unwrap_reference<Stepper>::type::stepper_category stepper_category;
    return detail_runge_kutta::runge_kutta(stepper, system, start_state, alphas, start_time, end_time, dt, driver, observer,
                                           stepper_category());
}

// public entry, no observer
template <class Stepper, class System, class State, class Time>
size_t runge_kutta(Stepper stepper, System system, State &start_state, const State &alphas, Time start_time, const Time end_time, Time dt,
                   Driver &driver)
{
    typedef typename odeint::unwrap_reference<Stepper>::type::stepper_category stepper_category;
    return detail_runge_kutta::runge_kutta(stepper, system, start_state, alphas, start_time, end_time, dt, driver,
                                           boost::numeric::odeint::null_observer(), stepper_category());
}

} // namespace vectorizedadjoint
#endif
