// va_pow.h -- pow(x, y) for x > 0 evaluated with the algorithm and tables of the GNU C Library's pow()
// (sysdeps/ieee754/dbl-64/e_pow.c by Szabolcs Nagy: 128-entry log table with a degree-7 tail polynomial in
// double-double, 128-entry exp2 table with a degree-5 polynomial), restated for host and device.
//
// Why it exists: odeint's default_step_adjuster (reference lib/include/detail/runge_kutta.hpp:105 -> try_step) calls
// pow() after every attempted step. For stiff problems the accept/reject sequence is sensitive to the last bit of the
// new step size, and the discrete adjoint differentiates that very sequence. CUDA's pow() differs from glibc's in the
// last bit for a few percent of arguments, which changed 5 of the reference's 9 printed Van der Pol step counts; with
// this routine the device reproduces all of them. tests/test_pow_cpu.py compares the HOST build of this file with the
// system pow() (bit for bit) on millions of arguments; the device executes the same IEEE operations (explicit fma()).
//
// Provenance / licence: the algorithm and the numeric tables (va_pow_tables.h, dumped from this image's libm.so.6 by
// tools/gen_pow_tables.py) originate in the GNU C Library, sysdeps/ieee754/dbl-64/{e_pow.c, e_pow_log_data.c, e_exp_data.c},
// Copyright (C) Free Software Foundation / Arm Ltd., licensed LGPL-2.1-or-later. This file is a restatement written for
// this project, not a copy of glibc source; the tables are glibc's data. Redistribution of this file and va_pow_tables.h
// is therefore under LGPL-2.1-or-later terms.
//
// VA_POW_FMA selects the variant glibc picks at run time on an FMA-capable x86-64 (ifunc __pow_fma); without it the
// Dekker-split variant of the generic build is used.
#pragma once
#ifndef __CUDACC_RTC__
#include <math.h>
#include <stdint.h>
#endif

#include "va_pow_tables.h"

#if defined(__CUDACC__)
#define VA_POW_HD __host__ __device__ __forceinline__
#else
#define VA_POW_HD static inline
#endif

#ifndef VA_POW_FMA
#define VA_POW_FMA 1
#endif

typedef struct { double invc, logc, logctail; } va_pow_logtab;

#if defined(__CUDA_ARCH__)
#define VA_POW_CONST __device__ const
#else
#define VA_POW_CONST static const
#endif
#if defined(__CUDACC__)
__device__ const va_pow_logtab va_pow_log_tab_d[128] = VA_POW_LOG_TAB;
__device__ const uint64_t va_pow_exp_tab_d[256] = VA_EXP_TAB;
#endif
#ifndef __CUDACC_RTC__
static const va_pow_logtab va_pow_log_tab_h[128] = VA_POW_LOG_TAB;
static const uint64_t va_pow_exp_tab_h[256] = VA_EXP_TAB;
#endif

VA_POW_HD uint64_t va_asuint64(double f)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(f);
#else
    union { double f; uint64_t i; } u = {f};
    return u.i;
#endif
}
VA_POW_HD double va_asdouble(uint64_t i)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)i);
#else
    union { uint64_t i; double f; } u = {i};
    return u.f;
#endif
}

// x: positive, finite, normal. y: finite with 2^-65 <= |y| < 2^63 and |y log x| < 512. Otherwise use the library pow.
// T, E: the two tables, wherever the caller keeps them (the kernels stage them in shared memory: the per-lane table gathers
// were the top stall of the Van der Pol forward kernel when they went to global memory, profiles/r02/vdp_forward_ncu_full_before.txt).
VA_POW_HD double va_pow_pos_t(double x, double y, const va_pow_logtab *T, const uint64_t *E)
{
    const double A[7] = VA_POW_LOG_POLY;
    const double C[4] = VA_EXP_POLY; // C2..C5
    const double Ln2hi = VA_POW_LN2HI, Ln2lo = VA_POW_LN2LO;
    // ---- log_inline: log(x) = hi + lo ---------------------------------------------------------------------------
    const uint64_t ix = va_asuint64(x);
    const uint64_t tmp = ix - 0x3fe6955500000000ULL;
    const int i = (int)((tmp >> (52 - 7)) % 128);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & (0xfffULL << 52));
    const double z = va_asdouble(iz);
    const double kd = (double)k;
    const double invc = T[i].invc, logc = T[i].logc, logctail = T[i].logctail;
#if VA_POW_FMA
    const double r = fma(z, invc, -1.0);
#else
    const double zhi = va_asdouble((iz + (1ULL << 31)) & (~0ULL << 32));
    const double zlo = z - zhi;
    const double rhi = zhi * invc - 1.0;
    const double rlo = zlo * invc;
    const double r = rhi + rlo;
#endif
    const double t1 = kd * Ln2hi + logc;
    const double t2 = t1 + r;
    const double lo1 = kd * Ln2lo + logctail;
    const double lo2 = t1 - t2 + r;
    const double ar = A[0] * r;
    const double ar2 = r * ar;
    const double ar3 = r * ar2;
#if VA_POW_FMA
    const double hi = t2 + ar2;
    const double lo3 = fma(ar, r, -ar2);
    const double lo4 = t2 - hi + ar2;
#else
    const double arhi = A[0] * rhi;
    const double arhi2 = rhi * arhi;
    const double hi = t2 + arhi2;
    const double lo3 = rlo * (ar + arhi);
    const double lo4 = t2 - hi + arhi2;
#endif
    const double p = ar3 * (A[1] + r * A[2] + ar2 * (A[3] + r * A[4] + ar2 * (A[5] + r * A[6])));
    const double lo = lo1 + lo2 + lo3 + lo4 + p;
    const double lhi_ = hi + lo;
    const double ltail = hi - lhi_ + lo;
    // ---- y * log(x) = ehi + elo -----------------------------------------------------------------------------------
#if VA_POW_FMA
    const double ehi = y * lhi_;
    const double elo = y * ltail + fma(y, lhi_, -ehi);
#else
    const double yhi = va_asdouble(va_asuint64(y) & (~0ULL << 27));
    const double ylo = y - yhi;
    const double lhi = va_asdouble(va_asuint64(lhi_) & (~0ULL << 27));
    const double llo = lhi_ - lhi + ltail;
    const double ehi = yhi * lhi;
    const double elo = ylo * lhi + y * llo;
#endif
    // ---- exp_inline(ehi, elo) ---------------------------------------------------------------------------------------
    const uint32_t abstop = (uint32_t)(va_asuint64(ehi) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x408u - 0x3c9u) { // |ehi| < 2^-54 or >= 512
        if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + ehi;
        return pow(x, y); // overflow / underflow range: not reached by the step-size controller
    }
    const double zz = VA_EXP_INVLN2N * ehi;
    double kd2 = zz + VA_EXP_SHIFT;
    const uint64_t ki = va_asuint64(kd2);
    kd2 -= VA_EXP_SHIFT;
    double rr = ehi + kd2 * VA_EXP_NEGLN2HIN + kd2 * VA_EXP_NEGLN2LON;
    rr += elo;
    const uint64_t idx = 2 * (ki % 128);
    const uint64_t top = ki << (52 - 7);
    const double tail = va_asdouble(E[idx]);
    const uint64_t sbits = E[idx + 1] + top;
    const double r2 = rr * rr;
#if VA_POW_FMA
    // the FMA build of glibc is compiled with floating-point contraction: these two expressions are fused there
    const double tmp2 = fma(r2 * r2, fma(rr, C[3], C[2]), fma(r2, fma(rr, C[1], C[0]), tail + rr));
#else
    const double tmp2 = tail + rr + r2 * (C[0] + rr * C[1]) + r2 * r2 * (C[2] + rr * C[3]);
#endif
    const double scale = va_asdouble(sbits);
#if VA_POW_FMA
    return fma(scale, tmp2, scale);
#else
    return scale + scale * tmp2;
#endif
}

VA_POW_HD double va_pow_pos(double x, double y)
{
#if defined(__CUDA_ARCH__)
    return va_pow_pos_t(x, y, va_pow_log_tab_d, va_pow_exp_tab_d);
#else
    return va_pow_pos_t(x, y, va_pow_log_tab_h, va_pow_exp_tab_h);
#endif
}

// pow for the step-size controller: falls back to the library for arguments outside the fast path
VA_POW_HD double va_pow_t(double x, double y, const va_pow_logtab *T, const uint64_t *E)
{
    const uint64_t ix = va_asuint64(x);
    const uint32_t topx = (uint32_t)(ix >> 52);
    const uint32_t topy = (uint32_t)(va_asuint64(y) >> 52) & 0x7ff;
    if (topx - 0x001u >= 0x7ffu - 0x001u || topy - 0x3beu >= 0x43eu - 0x3beu) return pow(x, y);
    return va_pow_pos_t(x, y, T, E);
}
VA_POW_HD double va_pow(double x, double y)
{
    const uint64_t ix = va_asuint64(x);
    const uint32_t topx = (uint32_t)(ix >> 52);
    const uint32_t topy = (uint32_t)(va_asuint64(y) >> 52) & 0x7ff;
    if (topx - 0x001u >= 0x7ffu - 0x001u || topy - 0x3beu >= 0x43eu - 0x3beu) return pow(x, y);
    return va_pow_pos(x, y);
}
