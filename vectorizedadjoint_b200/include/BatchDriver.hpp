// BatchDriver.hpp -- the batch axis the reference does not have, in the reference's own vocabulary.
//
// The reference integrates ONE parameter set per Driver (lib/include/Driver.hpp:15-79; its SIMD lanes carry adjoint seeds,
// detail/backpropagation.hpp:289-321). BatchDriver integrates B parameter sets per call on one GPU -- or, given a device
// list, sharded over several GPUs inside the call, the summed gradient combined by one NCCL all-reduce -- through the C-ABI
// (include/va_engine.h): same system functor concept (a template on the scalar type, recorded once like
// recordDriverRHSFunction does, Driver.hpp:95-100), same stepper objects (make_controlled<stepper>(abs, rel) or a fixed-step
// stepper, runge_kutta.hpp:47-59), same meaning of lambda (in: dJ/dx(tf) per cost function, out: dJ/dx(ti)) and mu
// (dJ/dalpha) as setCostGradients + adjointSolve (Driver.hpp:103-114, backpropagation.hpp:18-43) -- with one deliberate
// difference: mu is OVERWRITTEN, not incremented (there is no caller-owned accumulator per parameter set).
// Layout: what a caller of the reference would hold per parameter set, concatenated: x0 [B][Nin], alphas [B][Npar],
// lambda [B][Nout][Nin], mu [B][Nout][Npar].
#ifndef VA_B200_BATCH_DRIVER_HPP
#define VA_B200_BATCH_DRIVER_HPP

#include <stdexcept>
#include <string>
#include <vector>

#include <boost/numeric/odeint.hpp>

#include "va_engine.h"
#include "va_tape.hpp"

namespace vectorizedadjoint
{

class BatchDriver
{
  public:
    // Stepper: a fixed-step stepper object (runge_kutta4<State>()) or make_controlled<error_stepper>(abs, rel)
    template <class Stepper, class System>
    BatchDriver(Stepper stepper, System system, int Nin, int Nout, int Npar, int device = 0, int max_steps = 0,
                int ckpt_policy = VA_CKPT_AUTO)
        : BatchDriver(stepper, system, Nin, Nout, Npar, std::vector<int>{device}, max_steps, ckpt_policy)
    {
    }
    // devices: CUDA ordinals the batch is sharded over (contiguous ranges of parameter sets, no exchange during integration)
    template <class Stepper, class System>
    BatchDriver(Stepper stepper, System system, int Nin, int Nout, int Npar, const std::vector<int> &devices, int max_steps = 0,
                int ckpt_policy = VA_CKPT_AUTO)
        : nin_(Nin), nout_(Nout), npar_(Npar)
    {
        if (devices.empty()) throw std::invalid_argument("BatchDriver: empty device list");
        va_engine_desc d{};
        const std::vector<int32_t> devs(devices.begin(), devices.end());
        d.n_devices = static_cast<int32_t>(devs.size());
        d.devices = devs.data();
        const va::Tape tape = va::record(system, Nin, Npar);
        d.system = va::identify(tape);
        std::string src;
        if (d.system == va::SYS_TAPE) {
            src = tape.cuda_source("VaUserSys");
            d.tape_cuda_src = src.c_str();
        }
        d.n_state = Nin;
        d.n_par = Npar;
        d.n_out = Nout;
        d.stepper = Stepper::va_stepper_id;
        set_controller(d, stepper, typename Stepper::stepper_category());
        d.device = devs[0];
        d.max_steps = max_steps;
        d.ckpt_policy = ckpt_policy;
        if (va_engine_create(&d, &e_) != VA_OK) throw std::runtime_error(va_last_error());
    }
    ~BatchDriver() { va_engine_destroy(e_); }
    BatchDriver(const BatchDriver &) = delete;
    BatchDriver &operator=(const BatchDriver &) = delete;

    int GetNin() const { return nin_; }
    int GetNout() const { return nout_; }
    int GetNpar() const { return npar_; }

    // forward sweep + adjoint sweep for B parameter sets. x0 is overwritten with x(tf) (as runge_kutta does), lambda with
    // dJ/dx(ti), mu with dJ/dalpha. Returns the accepted-step counts; throws if any trajectory failed (checkpoint store
    // overflow, 500 consecutive rejections = odeint's no_progress_error, non-finite state).
    std::vector<int> forwardAdjoint(std::vector<double> &x0, const std::vector<double> &alphas, double ti, double tf, double dt,
                                    std::vector<double> &lambda, std::vector<double> &mu)
    {
        return run(x0, alphas, ti, tf, dt, lambda, mu, VA_REDUCE_NONE);
    }
    // the same with the objective summed over the batch: mu [Nout][Npar] = sum over parameter sets of dJ/dalpha (reduced on the
    // GPU; with several devices one ncclAllReduce inside the call)
    std::vector<int> forwardAdjointSummed(std::vector<double> &x0, const std::vector<double> &alphas, double ti, double tf, double dt,
                                          std::vector<double> &lambda, std::vector<double> &mu_sum)
    {
        return run(x0, alphas, ti, tf, dt, lambda, mu_sum, VA_REDUCE_SUM);
    }

    va_engine *handle() const { return e_; }

  private:
    std::vector<int> run(std::vector<double> &x0, const std::vector<double> &alphas, double ti, double tf, double dt,
                         std::vector<double> &lambda, std::vector<double> &mu, int reduce)
    {
        const long B = static_cast<long>(x0.size()) / nin_;
        if ((long)x0.size() != B * nin_ || (long)alphas.size() != B * npar_ || (long)lambda.size() != B * nout_ * nin_)
            throw std::invalid_argument("BatchDriver::forwardAdjoint: x0 [B][Nin], alphas [B][Npar], lambda [B][Nout][Nin]");
        mu.assign((reduce == VA_REDUCE_SUM ? (size_t)1 : (size_t)B) * nout_ * npar_, 0.0);
        std::vector<double> xf(x0.size());
        std::vector<int32_t> acc((size_t)B), status((size_t)B);
        va_batch_args a{};
        a.batch = B;
        a.x0 = x0.data();
        a.params = alphas.data();
        a.ti = ti;
        a.tf = tf;
        a.dt0 = dt;
        a.objective = VA_OBJ_SEED;
        a.reduce = reduce;
        a.mem = VA_MEM_HOST;
        a.x_final = xf.data();
        a.lambda = lambda.data();
        a.mu = mu.data();
        a.n_accept = acc.data();
        a.status = status.data();
        if (va_forward_adjoint_batch(e_, &a) != VA_OK) throw std::runtime_error(va_last_error());
        for (long b = 0; b < B; ++b)
            if (status[(size_t)b] != VA_TRAJ_OK)
                throw std::runtime_error("BatchDriver: parameter set " + std::to_string(b) + " failed with status " + std::to_string(status[(size_t)b]));
        x0.swap(xf);
        return std::vector<int>(acc.begin(), acc.end());
    }

    template <class Stepper>
    static void set_controller(va_engine_desc &d, const Stepper &s, boost::numeric::odeint::controlled_stepper_tag)
    {
        d.adaptive = 1;
        d.eps_abs = s.eps_abs;
        d.eps_rel = s.eps_rel;
    }
    template <class Stepper>
    static void set_controller(va_engine_desc &d, const Stepper &, boost::numeric::odeint::stepper_tag)
    {
        d.adaptive = 0;
    }
    va_engine *e_ = nullptr;
    int nin_, nout_, npar_;
};

} // namespace vectorizedadjoint

#endif
