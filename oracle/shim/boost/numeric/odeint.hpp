// Stand-in for the slice of Boost.Odeint 1.74 that the reference (lib/include + the three examples) uses.
// ORACLE BUILD ONLY. Boost is not installed in the build container and is not vendored by the reference, so the
// published odeint algorithm is restated here: explicit_generic_rk arithmetic order, controlled_runge_kutta::try_step,
// default_error_checker, default_step_adjuster, failed_step_checker, less_with_sign. With this header on the include
// path the UNMODIFIED reference headers compile; oracle/Makefile gates the result on the reference's own printed
// outputs (tests/golden/reference_goldens.json).
#ifndef VA_SHIM_BOOST_ODEINT_HPP
#define VA_SHIM_BOOST_ODEINT_HPP

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <tuple>
#include <typeinfo>
#include <vector>

#include <boost/numeric/ublas/shim_matrix.hpp>

namespace boost {
template <class T, std::size_t N>
using array = std::array<T, N>;
template <class E>
[[noreturn]] inline void throw_exception(const E &e) { throw e; }
template <class A, class B>
struct is_same : std::is_same<A, B> {};

namespace numeric { namespace odeint {

// --- categories -------------------------------------------------------------------------------------------------
struct stepper_tag {};
struct error_stepper_tag : stepper_tag {};
struct explicit_error_stepper_tag : error_stepper_tag {};
struct explicit_error_stepper_fsal_tag : error_stepper_tag {};
struct controlled_stepper_tag {};
struct explicit_controlled_stepper_tag : controlled_stepper_tag {};
struct explicit_controlled_stepper_fsal_tag : controlled_stepper_tag {};

enum controlled_step_result { success, fail };

template <class T> struct unwrap_reference { typedef T type; };
template <class T> struct unit_value_type { typedef T type; };

struct null_observer {
    template <class State, class Time> void operator()(const State &, Time) const {}
};

struct no_progress_error : std::runtime_error {
    no_progress_error() : std::runtime_error("Max number of iterations exceeded (500). A new step size was not found.") {}
};

// counts every call; the call after 500 earlier ones throws
class failed_step_checker
{
    int m_max_steps = 500, m_steps = 0;

  public:
    void reset() { m_steps = 0; }
    void operator()()
    {
        if (m_steps++ >= m_max_steps) throw no_progress_error();
    }
};

namespace detail {
template <class T> inline bool less_with_sign(T t1, T t2, T dt)
{
    if (dt > 0) return t2 - t1 > std::numeric_limits<T>::epsilon();
    return t1 - t2 > std::numeric_limits<T>::epsilon();
}
template <class T> inline bool less_eq_with_sign(T t1, T t2, T dt)
{
    if (dt > 0) return t1 - t2 <= std::numeric_limits<T>::epsilon();
    return t2 - t1 <= std::numeric_limits<T>::epsilon();
}
} // namespace detail

// --- coefficient tables -------------------------------------------------------------------------------------------
#define VA_COEF(NAME, N, ...)                                                        \
    template <class Value = double> struct NAME : boost::array<Value, N> {           \
        NAME() { const Value v[N] = {__VA_ARGS__}; for (std::size_t i = 0; i < N; ++i) (*this)[i] = v[i]; } \
    };
#define Q(a, b) (static_cast<Value>(a) / static_cast<Value>(b))

VA_COEF(rk4_coefficients_a1, 1, Q(1, 2))
VA_COEF(rk4_coefficients_a2, 2, Value(0), Q(1, 2))
VA_COEF(rk4_coefficients_a3, 3, Value(0), Value(0), Value(1))
VA_COEF(rk4_coefficients_b, 4, Q(1, 6), Q(1, 3), Q(1, 3), Q(1, 6))
VA_COEF(rk4_coefficients_c, 4, Value(0), Q(1, 2), Q(1, 2), Value(1))

VA_COEF(rk54_ck_coefficients_a1, 1, Q(1, 5))
VA_COEF(rk54_ck_coefficients_a2, 2, Q(3, 40), Q(9, 40))
VA_COEF(rk54_ck_coefficients_a3, 3, Q(3, 10), Q(-9, 10), Q(6, 5))
VA_COEF(rk54_ck_coefficients_a4, 4, Q(-11, 54), Q(5, 2), Q(-70, 27), Q(35, 27))
VA_COEF(rk54_ck_coefficients_a5, 5, Q(1631, 55296), Q(175, 512), Q(575, 13824), Q(44275, 110592), Q(253, 4096))
VA_COEF(rk54_ck_coefficients_b, 6, Q(37, 378), Value(0), Q(250, 621), Q(125, 594), Value(0), Q(512, 1771))
VA_COEF(rk54_ck_coefficients_db, 6, Q(37, 378) - Q(2825, 27648), Value(0) - Value(0), Q(250, 621) - Q(18575, 48384),
        Q(125, 594) - Q(13525, 55296), Value(0) - Q(277, 14336), Q(512, 1771) - Q(1, 4))
VA_COEF(rk54_ck_coefficients_c, 6, Value(0), Q(1, 5), Q(3, 10), Q(3, 5), Value(1), Q(7, 8))

VA_COEF(rk78_coefficients_a1, 1, Q(2, 27))
VA_COEF(rk78_coefficients_a2, 2, Q(1, 36), Q(1, 12))
VA_COEF(rk78_coefficients_a3, 3, Q(1, 24), Value(0), Q(1, 8))
VA_COEF(rk78_coefficients_a4, 4, Q(5, 12), Value(0), Q(-25, 16), Q(25, 16))
VA_COEF(rk78_coefficients_a5, 5, Q(1, 20), Value(0), Value(0), Q(1, 4), Q(1, 5))
VA_COEF(rk78_coefficients_a6, 6, Q(-25, 108), Value(0), Value(0), Q(125, 108), Q(-65, 27), Q(125, 54))
VA_COEF(rk78_coefficients_a7, 7, Q(31, 300), Value(0), Value(0), Value(0), Q(61, 225), Q(-2, 9), Q(13, 900))
VA_COEF(rk78_coefficients_a8, 8, Value(2), Value(0), Value(0), Q(-53, 6), Q(704, 45), Q(-107, 9), Q(67, 90), Value(3))
VA_COEF(rk78_coefficients_a9, 9, Q(-91, 108), Value(0), Value(0), Q(23, 108), Q(-976, 135), Q(311, 54), Q(-19, 60), Q(17, 6),
        Q(-1, 12))
VA_COEF(rk78_coefficients_a10, 10, Q(2383, 4100), Value(0), Value(0), Q(-341, 164), Q(4496, 1025), Q(-301, 82),
        Q(2133, 4100), Q(45, 82), Q(45, 164), Q(18, 41))
VA_COEF(rk78_coefficients_a11, 11, Q(3, 205), Value(0), Value(0), Value(0), Value(0), Q(-6, 41), Q(-3, 205), Q(-3, 41),
        Q(3, 41), Q(6, 41), Value(0))
VA_COEF(rk78_coefficients_a12, 12, Q(-1777, 4100), Value(0), Value(0), Q(-341, 164), Q(4496, 1025), Q(-289, 82),
        Q(2193, 4100), Q(51, 82), Q(33, 164), Q(12, 41), Value(0), Value(1))
VA_COEF(rk78_coefficients_b, 13, Value(0), Value(0), Value(0), Value(0), Value(0), Q(34, 105), Q(9, 35), Q(9, 35), Q(9, 280),
        Q(9, 280), Value(0), Q(41, 840), Q(41, 840))
VA_COEF(rk78_coefficients_db, 13, Value(0) - Q(41, 840), Value(0), Value(0), Value(0), Value(0), Value(0), Value(0), Value(0),
        Value(0), Value(0), Value(0) - Q(41, 840), Q(41, 840), Q(41, 840))
VA_COEF(rk78_coefficients_c, 13, Value(0), Q(2, 27), Q(1, 9), Q(1, 6), Q(5, 12), Q(1, 2), Q(5, 6), Q(1, 6), Q(2, 3), Q(1, 3),
        Value(1), Value(0), Value(1))
#undef Q
#undef VA_COEF

// --- explicit generic RK ----------------------------------------------------------------------------------------
// x_m = 1*x + (a_m0*dt)*k_0 + (a_m1*dt)*k_1 + ...  summed left to right, exactly like odeint's scale_sumN functors.
template <class State, std::size_t S>
class rk_core
{
  public:
    typedef State state_type;
    typedef double value_type;
    typedef double time_type;

  protected:
    std::array<std::array<double, S>, S> m_a{};
    std::array<double, S> m_b{}, m_db{}, m_c{};
    State m_k[S], m_xt;
    bool m_has_err = false;

    template <class Arr> void set_row(std::size_t m, const Arr &r)
    {
        for (std::size_t j = 0; j < r.size(); ++j) m_a[m][j] = r[j];
    }
    void resize(std::size_t n)
    {
        if (m_xt.size() != n) {
            m_xt.resize(n);
            for (auto &k : m_k) k.resize(n);
        }
    }

  public:
    // dxdt = f(x,t) supplied by the caller (this is k_0)
    template <class System>
    void do_step(System system, const State &x, const State &dxdt, double t, State &out, double dt, State *xerr)
    {
        const std::size_t n = x.size();
        resize(n);
        m_k[0] = dxdt;
        for (std::size_t m = 1; m < S; ++m) {
            for (std::size_t i = 0; i < n; ++i) {
                double acc = 1.0 * x[i];
                for (std::size_t j = 0; j < m; ++j) acc = acc + (m_a[m][j] * dt) * m_k[j][i];
                m_xt[i] = acc;
            }
            system(m_xt, m_k[m], t + dt * m_c[m]);
        }
        if (out.size() != n) out.resize(n);
        for (std::size_t i = 0; i < n; ++i) {
            double acc = 1.0 * x[i];
            for (std::size_t j = 0; j < S; ++j) acc = acc + (m_b[j] * dt) * m_k[j][i];
            m_xt[i] = acc;
        }
        if (xerr) {
            if (xerr->size() != n) xerr->resize(n);
            for (std::size_t i = 0; i < n; ++i) {
                double acc = (dt * m_db[0]) * m_k[0][i];
                for (std::size_t j = 1; j < S; ++j) acc = acc + (dt * m_db[j]) * m_k[j][i];
                (*xerr)[i] = acc;
            }
        }
        for (std::size_t i = 0; i < n; ++i) out[i] = m_xt[i];
    }
    template <class System>
    void do_step(System system, State &x, double t, double dt)
    {
        State dxdt(x.size());
        system(x, dxdt, t);
        do_step(system, x, dxdt, t, x, dt, nullptr);
    }
};

template <class State, class Value = double, class Deriv = State, class Time = Value>
class euler : public rk_core<State, 1>
{
  public:
    typedef stepper_tag stepper_category;
    euler() { this->m_b[0] = 1.0; }
};

template <class State, class Value = double, class Deriv = State, class Time = Value>
class runge_kutta4 : public rk_core<State, 4>
{
  public:
    typedef stepper_tag stepper_category;
    runge_kutta4()
    {
        this->set_row(1, rk4_coefficients_a1<double>());
        this->set_row(2, rk4_coefficients_a2<double>());
        this->set_row(3, rk4_coefficients_a3<double>());
        this->m_b = rk4_coefficients_b<double>();
        this->m_c = rk4_coefficients_c<double>();
    }
};

template <class State, class Value = double, class Deriv = State, class Time = Value>
class runge_kutta4_classic : public runge_kutta4<State, Value, Deriv, Time>
{
};

template <class State, class Value = double, class Deriv = State, class Time = Value>
class runge_kutta_cash_karp54 : public rk_core<State, 6>
{
  public:
    typedef explicit_error_stepper_tag stepper_category;
    static const unsigned short order_value = 5, stepper_order_value = 5, error_order_value = 4;
    runge_kutta_cash_karp54()
    {
        this->set_row(1, rk54_ck_coefficients_a1<double>());
        this->set_row(2, rk54_ck_coefficients_a2<double>());
        this->set_row(3, rk54_ck_coefficients_a3<double>());
        this->set_row(4, rk54_ck_coefficients_a4<double>());
        this->set_row(5, rk54_ck_coefficients_a5<double>());
        this->m_b = rk54_ck_coefficients_b<double>();
        this->m_db = rk54_ck_coefficients_db<double>();
        this->m_c = rk54_ck_coefficients_c<double>();
    }
};

template <class State, class Value = double, class Deriv = State, class Time = Value>
class runge_kutta_fehlberg78 : public rk_core<State, 13>
{
  public:
    typedef explicit_error_stepper_tag stepper_category;
    static const unsigned short order_value = 8, stepper_order_value = 8, error_order_value = 7;
    runge_kutta_fehlberg78()
    {
        this->set_row(1, rk78_coefficients_a1<double>());
        this->set_row(2, rk78_coefficients_a2<double>());
        this->set_row(3, rk78_coefficients_a3<double>());
        this->set_row(4, rk78_coefficients_a4<double>());
        this->set_row(5, rk78_coefficients_a5<double>());
        this->set_row(6, rk78_coefficients_a6<double>());
        this->set_row(7, rk78_coefficients_a7<double>());
        this->set_row(8, rk78_coefficients_a8<double>());
        this->set_row(9, rk78_coefficients_a9<double>());
        this->set_row(10, rk78_coefficients_a10<double>());
        this->set_row(11, rk78_coefficients_a11<double>());
        this->set_row(12, rk78_coefficients_a12<double>());
        this->m_b = rk78_coefficients_b<double>();
        this->m_db = rk78_coefficients_db<double>();
        this->m_c = rk78_coefficients_c<double>();
    }
};

// --- controlled_runge_kutta (non-FSAL) -----------------------------------------------------------------------------
template <class ErrorStepper>
class controlled_runge_kutta
{
  public:
    typedef explicit_controlled_stepper_tag stepper_category;
    typedef typename ErrorStepper::state_type state_type;
    typedef ErrorStepper stepper_type;

    controlled_runge_kutta(double eps_abs = 1e-6, double eps_rel = 1e-6, double a_x = 1.0, double a_dxdt = 1.0)
        : m_eps_abs(eps_abs), m_eps_rel(eps_rel), m_a_x(a_x), m_a_dxdt(a_dxdt)
    {
    }

    template <class System>
    controlled_step_result try_step(System system, state_type &x, double &t, double &dt)
    {
        const std::size_t n = x.size();
        if (m_dxdt.size() != n) { m_dxdt.resize(n); m_xnew.resize(n); m_xerr.resize(n); }
        system(x, m_dxdt, t);
        m_stepper.do_step(system, x, m_dxdt, t, m_xnew, dt, &m_xerr);
        // default_error_checker::error : max_i |xerr_i| / (eps_abs + eps_rel (a_x |x_i| + a_dxdt |dt| |dxdt_i|))
        const double a_dxdt_dt = m_a_dxdt * std::abs(dt);
        double max_rel_err = 0.0;
        for (std::size_t i = 0; i < n; ++i) {
            const double e = std::abs(m_xerr[i]) / (m_eps_abs + m_eps_rel * (m_a_x * std::abs(x[i]) + a_dxdt_dt * std::abs(m_dxdt[i])));
            max_rel_err = std::max(max_rel_err, e);
        }
        if (max_rel_err > 1.0) {
            // default_step_adjuster::decrease_step
            dt *= std::max(9.0 / 10.0 * std::pow(max_rel_err, -1.0 / (static_cast<double>(ErrorStepper::error_order_value) - 1.0)),
                           1.0 / 5.0);
            return fail;
        }
        t += dt;
        // default_step_adjuster::increase_step
        if (max_rel_err < 0.5) {
            double error = std::max(std::pow(5.0, -static_cast<double>(ErrorStepper::stepper_order_value)), max_rel_err);
            dt *= 9.0 / 10.0 * std::pow(error, -1.0 / static_cast<double>(ErrorStepper::stepper_order_value));
        }
        x = m_xnew;
        return success;
    }

  private:
    ErrorStepper m_stepper;
    double m_eps_abs, m_eps_rel, m_a_x, m_a_dxdt;
    state_type m_dxdt, m_xnew, m_xerr;
};

template <class Stepper>
inline controlled_runge_kutta<Stepper> make_controlled(double abs_error, double rel_error, const Stepper & = Stepper())
{
    return controlled_runge_kutta<Stepper>(abs_error, rel_error);
}

}} // namespace numeric::odeint
} // namespace boost
#endif
