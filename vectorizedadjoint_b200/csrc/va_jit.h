// va_jit.h -- run-time compiled kernels for recorded systems (VA_SYS_TAPE), see va_jit.cpp.
#pragma once
#include <string>

#include "va_types.h"

struct VaJitModule;
// user_src must define `struct VaUserSys` with N, NPAR, rhs(), vjp() (va::Tape::cuda_source("VaUserSys")).
int va_jit_compile(const char *user_src, int stages, int fsal, int stages_adj, VaJitModule **out, std::string &log);
void va_jit_destroy(VaJitModule *m);
// which: 0 forward (adaptive or fixed by args.adaptive), 1 adjoint. Returns a CUresult (0 = success).
int va_jit_launch(VaJitModule *m, int which, const VaScalarArgs &args, int64_t work_items, void *stream);
