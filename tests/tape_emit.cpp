// tape_emit <system> -- prints the CUDA rhs/vjp source (va::Tape::cuda_source("VaUserSys")) that recordDriverRHSFunction generates for
// one of the recorded example systems; the Python GPU tests hand it to va_engine_create (VA_SYS_TAPE) to run BATCHES of parameter
// sets through the run-time compiled kernels and compare them with fixtures made by the reference's AADC from the same functor source
// (vectorizedadjoint_b200/examples/tape_systems.hpp, tests/golden/make_goldens.py).
#include <cstdio>
#include <cstring>
#include <string>

#include "lib.hpp"
#include "tape_systems.hpp"

using tape_systems::HarvestedLotkaVolterra;

int main(int argc, char **argv)
{
    const std::string which = argc > 1 ? argv[1] : "";
    va::Tape tape;
    if (which == "pendulum") tape = va::record(tape_systems::DrivenPendulum(), 2, 3);
    else if (which == "pendulum_autonomous") { tape_systems::DrivenPendulum s; s.omega = 0.0; tape = va::record(s, 2, 3); }
    else if (which == "switched") tape = va::record(tape_systems::Switched(), 2, 3);
    else if (which == "switched_autonomous") { tape_systems::Switched s; s.tscale = 0.0; tape = va::record(s, 2, 3); }
    else if (which == "harvested_glv16") tape = va::record(HarvestedLotkaVolterra(), 16, 272);
    else if (which == "harvested_glv40") tape = va::record(HarvestedLotkaVolterra(), 40, 1640);
    else { std::fprintf(stderr, "usage: tape_emit pendulum|pendulum_autonomous|switched|switched_autonomous|harvested_glv16|harvested_glv40\n"); return 2; }
    if (va::identify(tape) != va::SYS_TAPE) { std::fprintf(stderr, "unexpectedly identified as a built-in system\n"); return 3; }
    std::fputs(tape.cuda_source("VaUserSys").c_str(), stdout);
    return 0;
}
