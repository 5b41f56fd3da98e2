/* Host check of vectorizedadjoint_b200/csrc/va_pow.h against the system pow(): bit-for-bit on the arguments the
 * step-size controller produces (exponents -1/3, -1/5, -1/6, -1/8, -order). Built and run by tests/test_pow_cpu.py. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "va_pow.h"

static uint64_t s = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

int main(int argc, char **argv)
{
    const long per = argc > 1 ? atol(argv[1]) : 2000000;
    const double ys[] = {-1.0 / 3, -1.0 / 5, -1.0 / 6, -1.0 / 8, -1.0 / 4, -1.0 / 7, -5.0, -8.0};
    long bad = 0, n = 0;
    for (int yi = 0; yi < 8; yi++) {
        const double y = ys[yi];
        for (long it = 0; it < per; it++) {
            const double u = (double)(rnd() >> 11) * 0x1.0p-53;
            double x;
            switch (it & 3) {
            case 0: x = 1.0 + u * 9; break;                               /* rejected steps: err slightly above 1 */
            case 1: x = ldexp(1.0 + u, (int)(rnd() % 40)); break;          /* badly rejected steps */
            case 2: x = 0.00032 + u * 0.5; break;                          /* accepted steps: 5^-5 <= err < 0.5 */
            default: x = ldexp(1.0 + u, -(int)(rnd() % 30));
            }
            if (yi >= 6) x = 5.0;
            n++;
            if (va_pow(x, y) != pow(x, y)) {
                if (bad < 5) printf("mismatch x=%a y=%a mine=%a libm=%a\n", x, y, va_pow(x, y), pow(x, y));
                bad++;
            }
        }
    }
    printf("%ld mismatches of %ld\n", bad, n);
    return bad != 0;
}
