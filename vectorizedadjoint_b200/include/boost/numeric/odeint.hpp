// boost/numeric/odeint.hpp -- stepper TAGS for the B200 engine's drop-in headers (not Boost).
//
// The reference passes odeint steppers by value to vectorizedadjoint::runge_kutta / constructDriverButcherTableau
// (reference lib/include/runge_kutta.hpp:47-59, Driver.hpp:87-93) and identifies them by type
// (ButcherTable.hpp:50,66,141,191). On the B200 path the arithmetic of those steppers runs inside CUDA kernels, so
// the types below only carry WHICH stepper was chosen (va_stepper id) and, for make_controlled<>, the tolerances.
// A translation unit that includes the real Boost.Odeint must not include this header.
#ifndef VA_B200_ODEINT_TAGS_HPP
#define VA_B200_ODEINT_TAGS_HPP

#include <vector>

namespace boost { namespace numeric { namespace odeint {

struct stepper_tag {};
struct error_stepper_tag : stepper_tag {};
struct explicit_error_stepper_tag : error_stepper_tag {};
struct explicit_error_stepper_fsal_tag : error_stepper_tag {};
struct controlled_stepper_tag {};
struct explicit_controlled_stepper_tag : controlled_stepper_tag {};
struct explicit_controlled_stepper_fsal_tag : controlled_stepper_tag {};

template <class T> struct unwrap_reference { typedef T type; };

struct null_observer {
    template <class State, class Time> void operator()(const State &, Time) const {}
};

// ids are include/va_engine.h's va_stepper values
#define VA_B200_STEPPER(NAME, ID, CATEGORY)                                              \
    template <class State, class Value = double, class Deriv = State, class Time = Value> \
    struct NAME {                                                                        \
        typedef State state_type;                                                        \
        typedef Value value_type;                                                        \
        typedef Time time_type;                                                          \
        typedef CATEGORY stepper_category;                                               \
        static constexpr int va_stepper_id = ID;                                         \
    };
VA_B200_STEPPER(euler, 0, stepper_tag)
VA_B200_STEPPER(runge_kutta4, 1, stepper_tag)
VA_B200_STEPPER(runge_kutta4_classic, 1, stepper_tag)
VA_B200_STEPPER(runge_kutta_cash_karp54, 2, explicit_error_stepper_tag)
VA_B200_STEPPER(runge_kutta_dopri5, 3, explicit_error_stepper_fsal_tag)
VA_B200_STEPPER(runge_kutta_fehlberg78, 4, explicit_error_stepper_tag)
#undef VA_B200_STEPPER

template <class ErrorStepper>
struct controlled_runge_kutta {
    typedef typename ErrorStepper::state_type state_type;
    typedef ErrorStepper stepper_type;
    typedef explicit_controlled_stepper_tag stepper_category;
    static constexpr int va_stepper_id = ErrorStepper::va_stepper_id;
    double eps_abs, eps_rel;
    controlled_runge_kutta(double abs_error = 1e-6, double rel_error = 1e-6) : eps_abs(abs_error), eps_rel(rel_error) {}
};

template <class Stepper>
inline controlled_runge_kutta<Stepper> make_controlled(double abs_error, double rel_error, const Stepper & = Stepper())
{
    return controlled_runge_kutta<Stepper>(abs_error, rel_error);
}

}}} // namespace boost::numeric::odeint

#endif
