// va_tableau.cpp -- Butcher tableaux of the supported steppers (host side; uploaded to kernels by value).
//
// Replaces reference lib/include/ButcherTable.hpp (which copies them out of Boost.Odeint's coefficient classes by
// typeid match: euler :50, runge_kutta4[_classic] :66, runge_kutta_cash_karp54 :141, runge_kutta_fehlberg78 :191) and
// adds Dormand-Prince 5(4), which the reference's reverse pass does not know (ButcherTable.hpp:247-250).
// Every coefficient is the correctly rounded quotient of two small integers, as in odeint; the embedded-error weights
// are formed as (rounded b) - (rounded b-hat), also as odeint does.
#include <cmath>
#include <cstring>

#include "va_common.cuh"

namespace {
struct Frac { int m, j; double num, den; };

void fill_a(VaTableau *tb, const Frac *f, int count)
{
    for (int k = 0; k < count; ++k) tb->a[f[k].m * VA_MAX_STAGES + f[k].j] = f[k].num / f[k].den;
}
} // namespace

static int fill(int kind, VaTableau *tb);

int va_tableau_host(int kind, VaTableau *tb)
{
    const int rc = fill(kind, tb);
    if (rc == 0) tb->growth_floor = std::pow(5.0, -(double)tb->stepper_order);
    return rc;
}

static int fill(int kind, VaTableau *tb)
{
    std::memset(tb, 0, sizeof(*tb));
    double *b = tb->b, *db = tb->db, *c = tb->c;
    switch (kind) {
    case VA_RK_EULER:
        tb->s = tb->s_adj = 1; tb->order = tb->stepper_order = 1;
        b[0] = 1.0;
        return 0;
    case VA_RK_RK4: {
        tb->s = tb->s_adj = 4; tb->order = tb->stepper_order = 4;
        static const Frac A[] = {{1, 0, 1, 2}, {2, 1, 1, 2}, {3, 2, 1, 1}};
        fill_a(tb, A, 3);
        b[0] = 1.0 / 6; b[1] = 1.0 / 3; b[2] = 1.0 / 3; b[3] = 1.0 / 6;
        c[1] = 0.5; c[2] = 0.5; c[3] = 1.0;
        return 0;
    }
    case VA_RK_CK54: {
        tb->s = tb->s_adj = 6; tb->order = tb->stepper_order = 5; tb->error_order = 4; tb->has_error = 1;
        static const Frac A[] = {{1, 0, 1, 5},
                                 {2, 0, 3, 40}, {2, 1, 9, 40},
                                 {3, 0, 3, 10}, {3, 1, -9, 10}, {3, 2, 6, 5},
                                 {4, 0, -11, 54}, {4, 1, 5, 2}, {4, 2, -70, 27}, {4, 3, 35, 27},
                                 {5, 0, 1631, 55296}, {5, 1, 175, 512}, {5, 2, 575, 13824}, {5, 3, 44275, 110592}, {5, 4, 253, 4096}};
        fill_a(tb, A, 15);
        const double bh[6] = {2825.0 / 27648, 0.0, 18575.0 / 48384, 13525.0 / 55296, 277.0 / 14336, 1.0 / 4};
        b[0] = 37.0 / 378; b[2] = 250.0 / 621; b[3] = 125.0 / 594; b[5] = 512.0 / 1771;
        for (int i = 0; i < 6; ++i) db[i] = b[i] - bh[i];
        c[1] = 1.0 / 5; c[2] = 3.0 / 10; c[3] = 3.0 / 5; c[4] = 1.0; c[5] = 7.0 / 8;
        return 0;
    }
    case VA_RK_DOPRI5: {
        tb->s = 7; tb->s_adj = 6; tb->order = tb->stepper_order = 5; tb->error_order = 4; tb->has_error = 1; tb->fsal = 1;
        static const Frac A[] = {{1, 0, 1, 5},
                                 {2, 0, 3, 40}, {2, 1, 9, 40},
                                 {3, 0, 44, 45}, {3, 1, -56, 15}, {3, 2, 32, 9},
                                 {4, 0, 19372, 6561}, {4, 1, -25360, 2187}, {4, 2, 64448, 6561}, {4, 3, -212, 729},
                                 {5, 0, 9017, 3168}, {5, 1, -355, 33}, {5, 2, 46732, 5247}, {5, 3, 49, 176}, {5, 4, -5103, 18656}};
        fill_a(tb, A, 15);
        b[0] = 35.0 / 384; b[2] = 500.0 / 1113; b[3] = 125.0 / 192; b[4] = -2187.0 / 6784; b[5] = 11.0 / 84;
        for (int j = 0; j < 6; ++j) tb->a[6 * VA_MAX_STAGES + j] = b[j];
        db[0] = b[0] - 5179.0 / 57600; db[2] = b[2] - 7571.0 / 16695; db[3] = b[3] - 393.0 / 640;
        db[4] = b[4] - (-92097.0 / 339200); db[5] = b[5] - 187.0 / 2100; db[6] = -1.0 / 40;
        c[1] = 1.0 / 5; c[2] = 3.0 / 10; c[3] = 4.0 / 5; c[4] = 8.0 / 9; c[5] = 1.0; c[6] = 1.0;
        return 0;
    }
    case VA_RK_RKF78: {
        tb->s = tb->s_adj = 13; tb->order = tb->stepper_order = 8; tb->error_order = 7; tb->has_error = 1;
        static const Frac A[] = {
            {1, 0, 2, 27},
            {2, 0, 1, 36}, {2, 1, 1, 12},
            {3, 0, 1, 24}, {3, 2, 1, 8},
            {4, 0, 5, 12}, {4, 2, -25, 16}, {4, 3, 25, 16},
            {5, 0, 1, 20}, {5, 3, 1, 4}, {5, 4, 1, 5},
            {6, 0, -25, 108}, {6, 3, 125, 108}, {6, 4, -65, 27}, {6, 5, 125, 54},
            {7, 0, 31, 300}, {7, 4, 61, 225}, {7, 5, -2, 9}, {7, 6, 13, 900},
            {8, 0, 2, 1}, {8, 3, -53, 6}, {8, 4, 704, 45}, {8, 5, -107, 9}, {8, 6, 67, 90}, {8, 7, 3, 1},
            {9, 0, -91, 108}, {9, 3, 23, 108}, {9, 4, -976, 135}, {9, 5, 311, 54}, {9, 6, -19, 60}, {9, 7, 17, 6}, {9, 8, -1, 12},
            {10, 0, 2383, 4100}, {10, 3, -341, 164}, {10, 4, 4496, 1025}, {10, 5, -301, 82}, {10, 6, 2133, 4100}, {10, 7, 45, 82},
            {10, 8, 45, 164}, {10, 9, 18, 41},
            {11, 0, 3, 205}, {11, 5, -6, 41}, {11, 6, -3, 205}, {11, 7, -3, 41}, {11, 8, 3, 41}, {11, 9, 6, 41},
            {12, 0, -1777, 4100}, {12, 3, -341, 164}, {12, 4, 4496, 1025}, {12, 5, -289, 82}, {12, 6, 2193, 4100}, {12, 7, 51, 82},
            {12, 8, 33, 164}, {12, 9, 12, 41}, {12, 11, 1, 1}};
        fill_a(tb, A, (int)(sizeof(A) / sizeof(A[0])));
        b[5] = 34.0 / 105; b[6] = 9.0 / 35; b[7] = 9.0 / 35; b[8] = 9.0 / 280; b[9] = 9.0 / 280; b[11] = 41.0 / 840; b[12] = 41.0 / 840;
        db[0] = 0.0 - 41.0 / 840; db[10] = 0.0 - 41.0 / 840; db[11] = 41.0 / 840; db[12] = 41.0 / 840;
        c[1] = 2.0 / 27; c[2] = 1.0 / 9; c[3] = 1.0 / 6; c[4] = 5.0 / 12; c[5] = 1.0 / 2; c[6] = 5.0 / 6; c[7] = 1.0 / 6;
        c[8] = 2.0 / 3; c[9] = 1.0 / 3; c[10] = 1.0; c[11] = 0.0; c[12] = 1.0;
        return 0;
    }
    default:
        return -1;
    }
}
