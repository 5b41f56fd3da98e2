// va_tape.hpp -- records the user's templated right-hand-side functor on an active scalar type.
//
// Replaces the recording half of the reference's AadData (lib/include/AadData.hpp:124-171: AADC idouble, startRecording /
// markAsInput / markAsOutput / stopRecording). The hook is the same: the system functor is a template on the scalar type
//     template <class T> void operator()(const std::vector<T>& x, std::vector<T>& dxdt, const std::vector<T>& p, const T t)
// (reference doc/source/harmonicOscillator.rst:85). Calling it once with va::adouble yields a straight-line tape of the
// RHS. The tape is used to (1) evaluate f on the host (Driver::Rhs), (2) identify the system: if the tape computes
// one of the engine's built-in device functors (harmonic oscillator, Van der Pol, generalized Lotka-Volterra) the
// hand-written sm_100a kernels are selected; (3) emit CUDA source for rhs / vjp device functors (tape -> CUDA).
#ifndef VA_B200_TAPE_HPP
#define VA_B200_TAPE_HPP

#include <cmath>
#include <cstdint>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace va {

// The operation set follows the surface of AADC's active scalar (reference aadc/include/aadc/idouble.h:242-631 arithmetic,
// :660-741 elementary functions; ibool.h:19-28 comparisons and iIf).
enum OpCode : uint8_t { OP_INPUT_X, OP_INPUT_P, OP_INPUT_T, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SIN, OP_COS, OP_EXP,
                        OP_LOG, OP_SQRT, OP_TANH, OP_POW,
                        OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN, OP_SINH, OP_COSH, OP_LOG10, OP_LOG2, OP_EXP2, OP_CBRT, OP_ERF, OP_FABS,
                        OP_ATAN2, OP_FMOD, OP_FMIN, OP_FMAX,
                        OP_LT, OP_LE, OP_GT, OP_GE, // comparisons: value 1.0 / 0.0, no derivative
                        OP_SELECT,                  // s ? a : b  (iIf)
                        OP_EQ, OP_NE,               // == != on active values (ibool.h:172-190)
                        OP_AND, OP_OR, OP_XOR, OP_NOT }; // && || != ! on recorded conditions (ibool.h:81-127)

struct Node {
    OpCode op;
    int32_t a, b;   // operand node ids (or input index for OP_INPUT_*)
    double c;       // constant value (OP_CONST)
    int32_t s = -1; // OP_SELECT: node id of the condition
};

class Tape
{
  public:
    std::vector<Node> nodes;
    std::vector<int32_t> outputs; // node id of dxdt[i]
    int n_x = 0, n_p = 0;

    int32_t push(OpCode op, int32_t a = -1, int32_t b = -1, double c = 0.0)
    {
        nodes.push_back(Node{op, a, b, c});
        return static_cast<int32_t>(nodes.size()) - 1;
    }

    // f(x, p, t) on the host
    void eval(const double *x, const double *p, double t, double *dxdt, std::vector<double> &work) const
    {
        work.resize(nodes.size());
        for (size_t k = 0; k < nodes.size(); ++k) {
            const Node &n = nodes[k];
            double v = 0.0;
            switch (n.op) {
            case OP_INPUT_X: v = x[n.a]; break;
            case OP_INPUT_P: v = p[n.a]; break;
            case OP_INPUT_T: v = t; break;
            case OP_CONST: v = n.c; break;
            case OP_ADD: v = work[n.a] + work[n.b]; break;
            case OP_SUB: v = work[n.a] - work[n.b]; break;
            case OP_MUL: v = work[n.a] * work[n.b]; break;
            case OP_DIV: v = work[n.a] / work[n.b]; break;
            case OP_NEG: v = -work[n.a]; break;
            case OP_SIN: v = std::sin(work[n.a]); break;
            case OP_COS: v = std::cos(work[n.a]); break;
            case OP_EXP: v = std::exp(work[n.a]); break;
            case OP_LOG: v = std::log(work[n.a]); break;
            case OP_SQRT: v = std::sqrt(work[n.a]); break;
            case OP_TANH: v = std::tanh(work[n.a]); break;
            case OP_POW: v = std::pow(work[n.a], work[n.b]); break;
            case OP_TAN: v = std::tan(work[n.a]); break;
            case OP_ASIN: v = std::asin(work[n.a]); break;
            case OP_ACOS: v = std::acos(work[n.a]); break;
            case OP_ATAN: v = std::atan(work[n.a]); break;
            case OP_SINH: v = std::sinh(work[n.a]); break;
            case OP_COSH: v = std::cosh(work[n.a]); break;
            case OP_LOG10: v = std::log10(work[n.a]); break;
            case OP_LOG2: v = std::log2(work[n.a]); break;
            case OP_EXP2: v = std::exp2(work[n.a]); break;
            case OP_CBRT: v = std::cbrt(work[n.a]); break;
            case OP_ERF: v = std::erf(work[n.a]); break;
            case OP_FABS: v = std::fabs(work[n.a]); break;
            case OP_ATAN2: v = std::atan2(work[n.a], work[n.b]); break;
            case OP_FMOD: v = std::fmod(work[n.a], work[n.b]); break;
            case OP_FMIN: v = work[n.a] < work[n.b] ? work[n.a] : work[n.b]; break;
            case OP_FMAX: v = work[n.a] > work[n.b] ? work[n.a] : work[n.b]; break;
            case OP_LT: v = work[n.a] < work[n.b] ? 1.0 : 0.0; break;
            case OP_LE: v = work[n.a] <= work[n.b] ? 1.0 : 0.0; break;
            case OP_GT: v = work[n.a] > work[n.b] ? 1.0 : 0.0; break;
            case OP_GE: v = work[n.a] >= work[n.b] ? 1.0 : 0.0; break;
            case OP_SELECT: v = work[n.s] != 0.0 ? work[n.a] : work[n.b]; break;
            case OP_EQ: v = work[n.a] == work[n.b] ? 1.0 : 0.0; break;
            case OP_NE: v = work[n.a] != work[n.b] ? 1.0 : 0.0; break;
            case OP_AND: v = (work[n.a] != 0.0 && work[n.b] != 0.0) ? 1.0 : 0.0; break;
            case OP_OR: v = (work[n.a] != 0.0 || work[n.b] != 0.0) ? 1.0 : 0.0; break;
            case OP_XOR: v = ((work[n.a] != 0.0) != (work[n.b] != 0.0)) ? 1.0 : 0.0; break;
            case OP_NOT: v = work[n.a] != 0.0 ? 0.0 : 1.0; break;
            }
            work[k] = v;
        }
        for (size_t i = 0; i < outputs.size(); ++i) dxdt[i] = work[outputs[i]];
    }

    // CUDA source of the device functors: straight-line rhs and reverse-mode vjp (tape -> CUDA generator).
    std::string cuda_source(const std::string &name) const;
};

inline Tape *&active_tape()
{
    static thread_local Tape *t = nullptr;
    return t;
}

// Active scalar. Outside a recording it behaves like a plain double (node == -1).
class adouble
{
  public:
    double val = 0.0;
    int32_t node = -1;

    adouble() = default;
    adouble(double v) : val(v) {}
    adouble(int v) : val(v) {}

    int32_t id() const
    {
        if (node >= 0) return node;
        Tape *t = active_tape();
        if (!t) throw std::logic_error("va::adouble used in an expression outside a recording");
        return t->push(OP_CONST, -1, -1, val);
    }
    static adouble make(OpCode op, double v, int32_t a, int32_t b = -1)
    {
        adouble r;
        r.val = v;
        r.node = active_tape()->push(op, a, b);
        return r;
    }
    static bool rec(const adouble &a, const adouble &b) { return active_tape() && (a.node >= 0 || b.node >= 0); }
    static bool rec(const adouble &a) { return active_tape() && a.node >= 0; }

    adouble &operator+=(const adouble &o) { return *this = *this + o; }
    adouble &operator-=(const adouble &o) { return *this = *this - o; }
    adouble &operator*=(const adouble &o) { return *this = *this * o; }
    adouble &operator/=(const adouble &o) { return *this = *this / o; }

    friend adouble operator+(const adouble &a, const adouble &b)
    {
        return rec(a, b) ? make(OP_ADD, a.val + b.val, a.id(), b.id()) : adouble(a.val + b.val);
    }
    friend adouble operator-(const adouble &a, const adouble &b)
    {
        return rec(a, b) ? make(OP_SUB, a.val - b.val, a.id(), b.id()) : adouble(a.val - b.val);
    }
    friend adouble operator*(const adouble &a, const adouble &b)
    {
        return rec(a, b) ? make(OP_MUL, a.val * b.val, a.id(), b.id()) : adouble(a.val * b.val);
    }
    friend adouble operator/(const adouble &a, const adouble &b)
    {
        return rec(a, b) ? make(OP_DIV, a.val / b.val, a.id(), b.id()) : adouble(a.val / b.val);
    }
    friend adouble operator-(const adouble &a) { return rec(a) ? make(OP_NEG, -a.val, a.id()) : adouble(-a.val); }
    friend adouble operator+(const adouble &a) { return a; }
};

#define VA_TAPE_UNARY(FN, OP)                                                                      \
    inline adouble FN(const adouble &a)                                                            \
    {                                                                                              \
        return adouble::rec(a) ? adouble::make(OP, std::FN(a.val), a.id()) : adouble(std::FN(a.val)); \
    }
VA_TAPE_UNARY(sin, OP_SIN)
VA_TAPE_UNARY(cos, OP_COS)
VA_TAPE_UNARY(exp, OP_EXP)
VA_TAPE_UNARY(log, OP_LOG)
VA_TAPE_UNARY(sqrt, OP_SQRT)
VA_TAPE_UNARY(tanh, OP_TANH)
VA_TAPE_UNARY(tan, OP_TAN)
VA_TAPE_UNARY(asin, OP_ASIN)
VA_TAPE_UNARY(acos, OP_ACOS)
VA_TAPE_UNARY(atan, OP_ATAN)
VA_TAPE_UNARY(sinh, OP_SINH)
VA_TAPE_UNARY(cosh, OP_COSH)
VA_TAPE_UNARY(log10, OP_LOG10)
VA_TAPE_UNARY(log2, OP_LOG2)
VA_TAPE_UNARY(exp2, OP_EXP2)
VA_TAPE_UNARY(cbrt, OP_CBRT)
VA_TAPE_UNARY(erf, OP_ERF)
VA_TAPE_UNARY(fabs, OP_FABS)
#undef VA_TAPE_UNARY
inline adouble abs(const adouble &a) { return fabs(a); }
#define VA_TAPE_BINARY(FN, OP, EXPR)                                                                   \
    inline adouble FN(const adouble &a, const adouble &b)                                              \
    {                                                                                                  \
        return adouble::rec(a, b) ? adouble::make(OP, EXPR, a.id(), b.id()) : adouble(EXPR);           \
    }
VA_TAPE_BINARY(pow, OP_POW, std::pow(a.val, b.val))
VA_TAPE_BINARY(atan2, OP_ATAN2, std::atan2(a.val, b.val))
VA_TAPE_BINARY(fmod, OP_FMOD, std::fmod(a.val, b.val))
VA_TAPE_BINARY(fmin, OP_FMIN, (a.val < b.val ? a.val : b.val))
VA_TAPE_BINARY(fmax, OP_FMAX, (a.val > b.val ? a.val : b.val))
#undef VA_TAPE_BINARY
inline adouble min(const adouble &a, const adouble &b) { return fmin(a, b); }
inline adouble max(const adouble &a, const adouble &b) { return fmax(a, b); }

// Recorded comparison (AADC's ibool): it does NOT convert to bool -- a branch on an active value would be frozen into the
// tape at its recording-time outcome -- and is consumed by iIf(condition, a, b), which records a select.
class abool
{
  public:
    bool val = false;
    int32_t node = -1;
    abool() = default;
    abool(bool v) : val(v) {}
};
#define VA_TAPE_COMPARE(SYM, OP)                                                                       \
    inline abool operator SYM(const adouble &a, const adouble &b)                                      \
    {                                                                                                  \
        abool r(a.val SYM b.val);                                                                      \
        if (adouble::rec(a, b)) r.node = adouble::make(OP, r.val ? 1.0 : 0.0, a.id(), b.id()).node;    \
        return r;                                                                                      \
    }
VA_TAPE_COMPARE(<, OP_LT)
VA_TAPE_COMPARE(<=, OP_LE)
VA_TAPE_COMPARE(>, OP_GT)
VA_TAPE_COMPARE(>=, OP_GE)
VA_TAPE_COMPARE(==, OP_EQ)
VA_TAPE_COMPARE(!=, OP_NE)
#undef VA_TAPE_COMPARE
// Logic on recorded conditions (AADC ibool.h:81-127: && and || -- there inside namespace aadcBoolOps --, != as exclusive or,
// and !). A passive operand (plain bool) becomes a constant node. Both operands are always evaluated: a recorded condition
// cannot short-circuit.
inline int32_t abool_id(const abool &c) { return c.node >= 0 ? c.node : active_tape()->push(OP_CONST, -1, -1, c.val ? 1.0 : 0.0); }
#define VA_TAPE_LOGIC(SYM, OP, EXPR)                                                                   \
    inline abool operator SYM(const abool &a, const abool &b)                                          \
    {                                                                                                  \
        abool r(EXPR);                                                                                 \
        if (active_tape() && (a.node >= 0 || b.node >= 0)) {                                           \
            const int32_t ia = abool_id(a), ib = abool_id(b);                                          \
            r.node = active_tape()->push(OP, ia, ib);                                                  \
        }                                                                                              \
        return r;                                                                                      \
    }                                                                                                  \
    inline abool operator SYM(bool a, const abool &b) { return abool(a) SYM b; }                       \
    inline abool operator SYM(const abool &a, bool b) { return a SYM abool(b); }
VA_TAPE_LOGIC(&&, OP_AND, a.val && b.val)
VA_TAPE_LOGIC(||, OP_OR, a.val || b.val)
VA_TAPE_LOGIC(!=, OP_XOR, a.val != b.val)
#undef VA_TAPE_LOGIC
inline abool operator!(const abool &a)
{
    abool r(!a.val);
    if (active_tape() && a.node >= 0) r.node = active_tape()->push(OP_NOT, a.node);
    return r;
}
namespace aadcBoolOps {} // source compatibility: `using namespace aadcBoolOps;` in a functor written for AADC is harmless here
inline adouble iIf(const abool &c, const adouble &a, const adouble &b)
{
    if (c.node < 0) return c.val ? a : b; // condition on passive values: an ordinary branch
    adouble r = adouble::make(OP_SELECT, c.val ? a.val : b.val, a.id(), b.id());
    active_tape()->nodes[(size_t)r.node].s = c.node;
    return r;
}
inline double iIf(bool c, double a, double b) { return c ? a : b; } // the same functor body instantiated with T = double

// Record system(x, dxdt, p, t) once. Mirrors AadData::Record (reference lib/include/AadData.hpp:124-171).
template <class System>
Tape record(System system, int n_x, int n_p)
{
    Tape tape;
    tape.n_x = n_x;
    tape.n_p = n_p;
    Tape *prev = active_tape();
    active_tape() = &tape;
    try {
        std::vector<adouble> x(n_x), p(n_p), f(n_x);
        adouble t;
        for (int i = 0; i < n_x; ++i) { x[i].val = 0.3 + 0.01 * i; x[i].node = tape.push(OP_INPUT_X, i); }
        for (int k = 0; k < n_p; ++k) { p[k].val = 1.0; p[k].node = tape.push(OP_INPUT_P, k); }
        t.node = tape.push(OP_INPUT_T);
        system(x, f, p, t);
        tape.outputs.resize(n_x);
        for (int i = 0; i < n_x; ++i) tape.outputs[i] = f[i].id();
    } catch (...) {
        active_tape() = prev;
        throw;
    }
    active_tape() = prev;
    return tape;
}

// ---- built-in systems the engine has hand-written device functors for -----------------------------------------------
enum SystemKind { SYS_HARMONIC = 0, SYS_VANDERPOL = 1, SYS_GLV = 2, SYS_TAPE = 3 };

inline void builtin_rhs(int kind, int n, const double *x, const double *p, double *f)
{
    if (kind == SYS_HARMONIC) {
        f[0] = x[1];
        f[1] = -1.0 * x[0] - p[0] * x[1];
    } else if (kind == SYS_VANDERPOL) {
        f[0] = x[1];
        f[1] = p[0] * ((1.0 - x[0] * x[0]) * x[1] - x[0]);
    } else {
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s += p[n * (i + 1) + j] * x[j];
            f[i] = x[i] * (p[i] + s);
        }
    }
}

// Which built-in functor does the tape compute? A tape is mapped to a hand-written device functor only when that is PROVABLY
// the same function: (1) structurally, the tape must be a polynomial in (x, p, t) -- inputs, constants, + - * and negation
// only; any division, elementary function, min / max / fabs / fmod, comparison or select node keeps it on the generated
// (SYS_TAPE) path, so a branchy functor that merely coincides with a built-in at the sample points is never swapped;
// (2) numerically, the polynomial must agree with the built-in formula (also a polynomial, total degree <= 4) at 12
// pseudo-random points spread over four decades -- two different polynomials of that degree agree at a random point with
// probability zero (Schwartz-Zippel). Robust against a different operation order in the user's functor.
inline int identify(const Tape &tape)
{
    const int n = tape.n_x, np = tape.n_p;
    for (const Node &nd : tape.nodes)
        switch (nd.op) {
        case OP_INPUT_X: case OP_INPUT_P: case OP_INPUT_T: case OP_CONST: case OP_ADD: case OP_SUB: case OP_MUL: case OP_NEG: break;
        default: return SYS_TAPE;
        }
    std::vector<int> candidates;
    if (n == 2 && np == 1) { candidates.push_back(SYS_HARMONIC); candidates.push_back(SYS_VANDERPOL); }
    if (np == n * n + n) candidates.push_back(SYS_GLV);
    std::vector<double> x(n), p(np), f(n), g(n), work;
    for (int kind : candidates) {
        bool same = true;
        uint64_t s = 0x243F6A8885A308D3ULL;
        for (int trial = 0; trial < 12 && same; ++trial) {
            auto u = [&]() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(s >> 11) * 0x1.0p-53 * 2.0 - 1.0; };
            const double sx = std::pow(10.0, (trial % 4) - 2), sp = std::pow(10.0, ((trial / 4) % 3) - 1); // 1e-2..10, 0.1..10
            for (auto &v : x) v = sx * u();
            for (auto &v : p) v = sp * 2.0 * u();
            tape.eval(x.data(), p.data(), 0.37 * trial - 1.5, f.data(), work);
            builtin_rhs(kind, n, x.data(), p.data(), g.data());
            double scale = 0.0;
            for (int i = 0; i < n; ++i) scale = std::fmax(scale, std::fabs(g[i]));
            for (int i = 0; i < n; ++i) same = same && std::fabs(f[i] - g[i]) <= 1e-12 * (scale + std::fabs(g[i])) + 1e-300;
        }
        if (same) return kind;
    }
    return SYS_TAPE;
}

// ---- tape -> CUDA ----------------------------------------------------------------------------------------------------
inline std::string Tape::cuda_source(const std::string &name) const
{
    std::ostringstream o;
    o.precision(17);
    auto v = [](int32_t k) { return "v" + std::to_string(k); };
    auto d = [](int32_t k) { return "d" + std::to_string(k); };
    o << "// generated by va::Tape::cuda_source -- rhs and vjp of a recorded system (" << nodes.size() << " nodes)\n";
    o << "struct " << name << " {\n  static constexpr int N = " << n_x << ", NPAR = " << n_p << ";\n";
    auto forward = [&]() {
        for (size_t k = 0; k < nodes.size(); ++k) {
            const Node &n = nodes[k];
            o << "    const double " << v((int32_t)k) << " = ";
            switch (n.op) {
            case OP_INPUT_X: o << "x[" << n.a << "]"; break;
            case OP_INPUT_P: o << "p[" << n.a << "]"; break;
            case OP_INPUT_T: o << "t"; break;
            case OP_CONST: o << std::hexfloat << n.c << std::defaultfloat; break;
            case OP_ADD: o << v(n.a) << " + " << v(n.b); break;
            case OP_SUB: o << v(n.a) << " - " << v(n.b); break;
            case OP_MUL: o << v(n.a) << " * " << v(n.b); break;
            case OP_DIV: o << v(n.a) << " / " << v(n.b); break;
            case OP_NEG: o << "-" << v(n.a); break;
            case OP_SIN: o << "sin(" << v(n.a) << ")"; break;
            case OP_COS: o << "cos(" << v(n.a) << ")"; break;
            case OP_EXP: o << "exp(" << v(n.a) << ")"; break;
            case OP_LOG: o << "log(" << v(n.a) << ")"; break;
            case OP_SQRT: o << "sqrt(" << v(n.a) << ")"; break;
            case OP_TANH: o << "tanh(" << v(n.a) << ")"; break;
            case OP_POW: o << "pow(" << v(n.a) << ", " << v(n.b) << ")"; break;
            case OP_TAN: o << "tan(" << v(n.a) << ")"; break;
            case OP_ASIN: o << "asin(" << v(n.a) << ")"; break;
            case OP_ACOS: o << "acos(" << v(n.a) << ")"; break;
            case OP_ATAN: o << "atan(" << v(n.a) << ")"; break;
            case OP_SINH: o << "sinh(" << v(n.a) << ")"; break;
            case OP_COSH: o << "cosh(" << v(n.a) << ")"; break;
            case OP_LOG10: o << "log10(" << v(n.a) << ")"; break;
            case OP_LOG2: o << "log2(" << v(n.a) << ")"; break;
            case OP_EXP2: o << "exp2(" << v(n.a) << ")"; break;
            case OP_CBRT: o << "cbrt(" << v(n.a) << ")"; break;
            case OP_ERF: o << "erf(" << v(n.a) << ")"; break;
            case OP_FABS: o << "fabs(" << v(n.a) << ")"; break;
            case OP_ATAN2: o << "atan2(" << v(n.a) << ", " << v(n.b) << ")"; break;
            case OP_FMOD: o << "fmod(" << v(n.a) << ", " << v(n.b) << ")"; break;
            case OP_FMIN: o << "(" << v(n.a) << " < " << v(n.b) << " ? " << v(n.a) << " : " << v(n.b) << ")"; break;
            case OP_FMAX: o << "(" << v(n.a) << " > " << v(n.b) << " ? " << v(n.a) << " : " << v(n.b) << ")"; break;
            case OP_LT: o << "(" << v(n.a) << " < " << v(n.b) << " ? 1.0 : 0.0)"; break;
            case OP_LE: o << "(" << v(n.a) << " <= " << v(n.b) << " ? 1.0 : 0.0)"; break;
            case OP_GT: o << "(" << v(n.a) << " > " << v(n.b) << " ? 1.0 : 0.0)"; break;
            case OP_GE: o << "(" << v(n.a) << " >= " << v(n.b) << " ? 1.0 : 0.0)"; break;
            case OP_SELECT: o << "(" << v(n.s) << " != 0.0 ? " << v(n.a) << " : " << v(n.b) << ")"; break;
            case OP_EQ: o << "(" << v(n.a) << " == " << v(n.b) << " ? 1.0 : 0.0)"; break;
            case OP_NE: o << "(" << v(n.a) << " != " << v(n.b) << " ? 1.0 : 0.0)"; break;
            case OP_AND: o << "((" << v(n.a) << " != 0.0 && " << v(n.b) << " != 0.0) ? 1.0 : 0.0)"; break;
            case OP_OR: o << "((" << v(n.a) << " != 0.0 || " << v(n.b) << " != 0.0) ? 1.0 : 0.0)"; break;
            case OP_XOR: o << "(((" << v(n.a) << " != 0.0) != (" << v(n.b) << " != 0.0)) ? 1.0 : 0.0)"; break;
            case OP_NOT: o << "(" << v(n.a) << " != 0.0 ? 0.0 : 1.0)"; break;
            }
            o << ";\n";
        }
    };
    // a long tape is compiled once, as a real function, instead of being inlined into every stage of the stepper (the kernels
    // call rhs / vjp s times each): compile time of a 40-species Lotka-Volterra variant drops from minutes to seconds
    const char *inl = nodes.size() > 1500 ? "__noinline__ " : "";
    o << "  " << inl << "__device__ static void rhs(const double *x, const double *p, double t, double *dx) {\n";
    forward();
    for (size_t i = 0; i < outputs.size(); ++i) o << "    dx[" << i << "] = " << v(outputs[i]) << ";\n";
    o << "  }\n";
    o << "  " << inl << "__device__ static void vjp(const double *x, const double *p, double t, const double *w, double *gx, double *gp) {\n";
    forward();
    for (size_t k = 0; k < nodes.size(); ++k) o << "    double " << d((int32_t)k) << " = 0.0;\n";
    for (size_t i = 0; i < outputs.size(); ++i) o << "    " << d(outputs[i]) << " += w[" << i << "];\n";
    for (int32_t k = (int32_t)nodes.size() - 1; k >= 0; --k) {
        const Node &n = nodes[k];
        switch (n.op) {
        case OP_ADD: o << "    " << d(n.a) << " += " << d(k) << "; " << d(n.b) << " += " << d(k) << ";\n"; break;
        case OP_SUB: o << "    " << d(n.a) << " += " << d(k) << "; " << d(n.b) << " -= " << d(k) << ";\n"; break;
        case OP_MUL: o << "    " << d(n.a) << " += " << d(k) << " * " << v(n.b) << "; " << d(n.b) << " += " << d(k) << " * " << v(n.a) << ";\n"; break;
        case OP_DIV: o << "    " << d(n.a) << " += " << d(k) << " / " << v(n.b) << "; " << d(n.b) << " -= " << d(k) << " * " << v(k) << " / " << v(n.b) << ";\n"; break;
        case OP_NEG: o << "    " << d(n.a) << " -= " << d(k) << ";\n"; break;
        case OP_SIN: o << "    " << d(n.a) << " += " << d(k) << " * cos(" << v(n.a) << ");\n"; break;
        case OP_COS: o << "    " << d(n.a) << " -= " << d(k) << " * sin(" << v(n.a) << ");\n"; break;
        case OP_EXP: o << "    " << d(n.a) << " += " << d(k) << " * " << v(k) << ";\n"; break;
        case OP_LOG: o << "    " << d(n.a) << " += " << d(k) << " / " << v(n.a) << ";\n"; break;
        case OP_SQRT: o << "    " << d(n.a) << " += " << d(k) << " * 0.5 / " << v(k) << ";\n"; break;
        case OP_TANH: o << "    " << d(n.a) << " += " << d(k) << " * (1.0 - " << v(k) << " * " << v(k) << ");\n"; break;
        case OP_POW:
            o << "    " << d(n.a) << " += " << d(k) << " * " << v(n.b) << " * pow(" << v(n.a) << ", " << v(n.b) << " - 1.0); " << d(n.b)
              << " += " << d(k) << " * " << v(k) << " * log(" << v(n.a) << ");\n";
            break;
        case OP_TAN: o << "    " << d(n.a) << " += " << d(k) << " * (1.0 + " << v(k) << " * " << v(k) << ");\n"; break;
        case OP_ASIN: o << "    " << d(n.a) << " += " << d(k) << " / sqrt(1.0 - " << v(n.a) << " * " << v(n.a) << ");\n"; break;
        case OP_ACOS: o << "    " << d(n.a) << " -= " << d(k) << " / sqrt(1.0 - " << v(n.a) << " * " << v(n.a) << ");\n"; break;
        case OP_ATAN: o << "    " << d(n.a) << " += " << d(k) << " / (1.0 + " << v(n.a) << " * " << v(n.a) << ");\n"; break;
        case OP_SINH: o << "    " << d(n.a) << " += " << d(k) << " * cosh(" << v(n.a) << ");\n"; break;
        case OP_COSH: o << "    " << d(n.a) << " += " << d(k) << " * sinh(" << v(n.a) << ");\n"; break;
        case OP_LOG10: o << "    " << d(n.a) << " += " << d(k) << " / (" << v(n.a) << " * 2.302585092994046);\n"; break;
        case OP_LOG2: o << "    " << d(n.a) << " += " << d(k) << " / (" << v(n.a) << " * 0.6931471805599453);\n"; break;
        case OP_EXP2: o << "    " << d(n.a) << " += " << d(k) << " * " << v(k) << " * 0.6931471805599453;\n"; break;
        case OP_CBRT: o << "    " << d(n.a) << " += " << d(k) << " / (3.0 * " << v(k) << " * " << v(k) << ");\n"; break;
        case OP_ERF: o << "    " << d(n.a) << " += " << d(k) << " * 1.1283791670955126 * exp(-" << v(n.a) << " * " << v(n.a) << ");\n"; break;
        case OP_FABS: o << "    " << d(n.a) << " += " << d(k) << " * copysign(1.0, " << v(n.a) << ");\n"; break;
        case OP_ATAN2:
            o << "    { const double q = " << v(n.a) << " * " << v(n.a) << " + " << v(n.b) << " * " << v(n.b) << "; " << d(n.a) << " += " << d(k)
              << " * " << v(n.b) << " / q; " << d(n.b) << " -= " << d(k) << " * " << v(n.a) << " / q; }\n";
            break;
        case OP_FMOD: o << "    " << d(n.a) << " += " << d(k) << "; " << d(n.b) << " -= " << d(k) << " * trunc(" << v(n.a) << " / " << v(n.b) << ");\n"; break;
        case OP_FMIN: o << "    if (" << v(n.a) << " < " << v(n.b) << ") " << d(n.a) << " += " << d(k) << "; else " << d(n.b) << " += " << d(k) << ";\n"; break;
        case OP_FMAX: o << "    if (" << v(n.a) << " > " << v(n.b) << ") " << d(n.a) << " += " << d(k) << "; else " << d(n.b) << " += " << d(k) << ";\n"; break;
        case OP_SELECT: o << "    if (" << v(n.s) << " != 0.0) " << d(n.a) << " += " << d(k) << "; else " << d(n.b) << " += " << d(k) << ";\n"; break;
        case OP_INPUT_X: o << "    gx[" << n.a << "] = " << d(k) << ";\n"; break;
        case OP_INPUT_P: o << "    gp[" << n.a << "] += " << d(k) << ";\n"; break;
        default: break;
        }
    }
    o << "  }\n};\n";
    return o.str();
}

} // namespace va

#endif
