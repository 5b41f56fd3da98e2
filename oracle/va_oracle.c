/*
 * va_oracle.c -- CPU restatement of the VectorizedAdjoint hot path. TEST INFRASTRUCTURE, see va_oracle.h.
 *
 * Build with FP contraction OFF (the reference is built with -mavx2 only, i.e. no FMA anywhere in the
 * odeint arithmetic): gcc -O2 -ffp-contract=off -fno-fast-math.
 *
 * Sections
 *   1. Butcher tableaux      : Boost.Odeint 1.74 rk4 / rk54_ck / rk78 coefficient classes as read by
 *                              reference lib/include/ButcherTable.hpp:66-246, plus dopri5 (extension).
 *   2. Example systems + VJPs: reference examples/{HarmonicOscillator/main.cpp:20-29, VanDerPol/main.cpp:38-43,
 *                              GeneralizedLotkaVolterra/main.cpp:105-119}; the VJP replaces
 *                              reference lib/include/AadData.hpp:291-373 (AADC reverse kernel).
 *   3. Forward sweep         : reference lib/include/detail/runge_kutta.hpp:38-72 (fixed) and :76-118 (adaptive),
 *                              with odeint's controlled_runge_kutta::try_step / default_error_checker /
 *                              default_step_adjuster / explicit_generic_rk restated (Boost is not vendored).
 *   4. Reverse sweep         : reference lib/include/detail/backpropagation.hpp:24-64 (stage recompute),
 *                              :66-81 (intermediate state), :160-229 (one-step adjoint), :256-278 (loop).
 *   5. Batch driver + synthetic inputs.
 */
#include "va_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MS VO_MAX_STAGES

/* ------------------------------------------------------------------------------------------------ */
/* 1. tableaux                                                                                      */
/* ------------------------------------------------------------------------------------------------ */

static void tb_clear(vo_tableau *tb) { memset(tb, 0, sizeof(*tb)); }
#define A_(m, j) a[(m) * MS + (j)]

int vo_tableau_get(int kind, vo_tableau *tb)
{
    tb_clear(tb);
    double *a = tb->a, *b = tb->b, *c = tb->c, *db = tb->db;
    switch (kind) {
    case VO_RK_EULER: /* ButcherTable.hpp:50-65 */
        tb->s = tb->s_adj = 1; tb->order = tb->stepper_order = 1; tb->error_order = 0;
        b[0] = 1.0; c[0] = 0.0;
        return 0;
    case VO_RK_RK4: /* odeint rk4_coefficients_{a1,a2,a3,b,c}; ButcherTable.hpp:66-120 */
        tb->s = tb->s_adj = 4; tb->order = tb->stepper_order = 4; tb->error_order = 0;
        A_(1, 0) = 1.0 / 2.0;
        A_(2, 0) = 0.0; A_(2, 1) = 1.0 / 2.0;
        A_(3, 0) = 0.0; A_(3, 1) = 0.0; A_(3, 2) = 1.0;
        b[0] = 1.0 / 6.0; b[1] = 1.0 / 3.0; b[2] = 1.0 / 3.0; b[3] = 1.0 / 6.0;
        c[0] = 0.0; c[1] = 1.0 / 2.0; c[2] = 1.0 / 2.0; c[3] = 1.0;
        return 0;
    case VO_RK_CK54: { /* odeint rk54_ck_coefficients_*; ButcherTable.hpp:141-190 */
        tb->s = tb->s_adj = 6; tb->order = 5; tb->stepper_order = 5; tb->error_order = 4; tb->has_error = 1;
        A_(1, 0) = 1.0 / 5.0;
        A_(2, 0) = 3.0 / 40.0; A_(2, 1) = 9.0 / 40.0;
        A_(3, 0) = 3.0 / 10.0; A_(3, 1) = -9.0 / 10.0; A_(3, 2) = 6.0 / 5.0;
        A_(4, 0) = -11.0 / 54.0; A_(4, 1) = 5.0 / 2.0; A_(4, 2) = -70.0 / 27.0; A_(4, 3) = 35.0 / 27.0;
        A_(5, 0) = 1631.0 / 55296.0; A_(5, 1) = 175.0 / 512.0; A_(5, 2) = 575.0 / 13824.0;
        A_(5, 3) = 44275.0 / 110592.0; A_(5, 4) = 253.0 / 4096.0;
        b[0] = 37.0 / 378.0; b[1] = 0.0; b[2] = 250.0 / 621.0; b[3] = 125.0 / 594.0; b[4] = 0.0; b[5] = 512.0 / 1771.0;
        /* odeint forms db as (rounded b) - (rounded b-hat) */
        db[0] = b[0] - 2825.0 / 27648.0; db[1] = b[1] - 0.0; db[2] = b[2] - 18575.0 / 48384.0;
        db[3] = b[3] - 13525.0 / 55296.0; db[4] = b[4] - 277.0 / 14336.0; db[5] = b[5] - 1.0 / 4.0;
        c[0] = 0.0; c[1] = 1.0 / 5.0; c[2] = 3.0 / 10.0; c[3] = 3.0 / 5.0; c[4] = 1.0; c[5] = 7.0 / 8.0;
        return 0;
    }
    case VO_RK_DOPRI5: { /* odeint runge_kutta_dopri5 (hand-coded there); extension, not in ButcherTable.hpp */
        tb->s = 7; tb->s_adj = 6; tb->order = 5; tb->stepper_order = 5; tb->error_order = 4; tb->has_error = 1; tb->fsal = 1;
        A_(1, 0) = 1.0 / 5.0;
        A_(2, 0) = 3.0 / 40.0; A_(2, 1) = 9.0 / 40.0;
        A_(3, 0) = 44.0 / 45.0; A_(3, 1) = -56.0 / 15.0; A_(3, 2) = 32.0 / 9.0;
        A_(4, 0) = 19372.0 / 6561.0; A_(4, 1) = -25360.0 / 2187.0; A_(4, 2) = 64448.0 / 6561.0; A_(4, 3) = -212.0 / 729.0;
        A_(5, 0) = 9017.0 / 3168.0; A_(5, 1) = -355.0 / 33.0; A_(5, 2) = 46732.0 / 5247.0; A_(5, 3) = 49.0 / 176.0;
        A_(5, 4) = -5103.0 / 18656.0;
        b[0] = 35.0 / 384.0; b[1] = 0.0; b[2] = 500.0 / 1113.0; b[3] = 125.0 / 192.0; b[4] = -2187.0 / 6784.0;
        b[5] = 11.0 / 84.0; b[6] = 0.0;
        for (int j = 0; j < 6; ++j) A_(6, j) = b[j];
        db[0] = b[0] - 5179.0 / 57600.0; db[1] = 0.0; db[2] = b[2] - 7571.0 / 16695.0; db[3] = b[3] - 393.0 / 640.0;
        db[4] = b[4] - (-92097.0 / 339200.0); db[5] = b[5] - 187.0 / 2100.0; db[6] = -1.0 / 40.0;
        c[0] = 0.0; c[1] = 1.0 / 5.0; c[2] = 3.0 / 10.0; c[3] = 4.0 / 5.0; c[4] = 8.0 / 9.0; c[5] = 1.0; c[6] = 1.0;
        return 0;
    }
    case VO_RK_RKF78: { /* odeint rk78_coefficients_*; ButcherTable.hpp:191-246 */
        tb->s = tb->s_adj = 13; tb->order = 8; tb->stepper_order = 8; tb->error_order = 7; tb->has_error = 1;
        A_(1, 0) = 2.0 / 27.0;
        A_(2, 0) = 1.0 / 36.0; A_(2, 1) = 1.0 / 12.0;
        A_(3, 0) = 1.0 / 24.0; A_(3, 2) = 1.0 / 8.0;
        A_(4, 0) = 5.0 / 12.0; A_(4, 2) = -25.0 / 16.0; A_(4, 3) = 25.0 / 16.0;
        A_(5, 0) = 1.0 / 20.0; A_(5, 3) = 1.0 / 4.0; A_(5, 4) = 1.0 / 5.0;
        A_(6, 0) = -25.0 / 108.0; A_(6, 3) = 125.0 / 108.0; A_(6, 4) = -65.0 / 27.0; A_(6, 5) = 125.0 / 54.0;
        A_(7, 0) = 31.0 / 300.0; A_(7, 4) = 61.0 / 225.0; A_(7, 5) = -2.0 / 9.0; A_(7, 6) = 13.0 / 900.0;
        A_(8, 0) = 2.0; A_(8, 3) = -53.0 / 6.0; A_(8, 4) = 704.0 / 45.0; A_(8, 5) = -107.0 / 9.0; A_(8, 6) = 67.0 / 90.0;
        A_(8, 7) = 3.0;
        A_(9, 0) = -91.0 / 108.0; A_(9, 3) = 23.0 / 108.0; A_(9, 4) = -976.0 / 135.0; A_(9, 5) = 311.0 / 54.0;
        A_(9, 6) = -19.0 / 60.0; A_(9, 7) = 17.0 / 6.0; A_(9, 8) = -1.0 / 12.0;
        A_(10, 0) = 2383.0 / 4100.0; A_(10, 3) = -341.0 / 164.0; A_(10, 4) = 4496.0 / 1025.0; A_(10, 5) = -301.0 / 82.0;
        A_(10, 6) = 2133.0 / 4100.0; A_(10, 7) = 45.0 / 82.0; A_(10, 8) = 45.0 / 164.0; A_(10, 9) = 18.0 / 41.0;
        A_(11, 0) = 3.0 / 205.0; A_(11, 5) = -6.0 / 41.0; A_(11, 6) = -3.0 / 205.0; A_(11, 7) = -3.0 / 41.0;
        A_(11, 8) = 3.0 / 41.0; A_(11, 9) = 6.0 / 41.0;
        A_(12, 0) = -1777.0 / 4100.0; A_(12, 3) = -341.0 / 164.0; A_(12, 4) = 4496.0 / 1025.0; A_(12, 5) = -289.0 / 82.0;
        A_(12, 6) = 2193.0 / 4100.0; A_(12, 7) = 51.0 / 82.0; A_(12, 8) = 33.0 / 164.0; A_(12, 9) = 12.0 / 41.0;
        A_(12, 11) = 1.0;
        b[5] = 34.0 / 105.0; b[6] = 9.0 / 35.0; b[7] = 9.0 / 35.0; b[8] = 9.0 / 280.0; b[9] = 9.0 / 280.0;
        b[11] = 41.0 / 840.0; b[12] = 41.0 / 840.0;
        db[0] = 0.0 - 41.0 / 840.0; db[10] = 0.0 - 41.0 / 840.0; db[11] = 41.0 / 840.0; db[12] = 41.0 / 840.0;
        c[1] = 2.0 / 27.0; c[2] = 1.0 / 9.0; c[3] = 1.0 / 6.0; c[4] = 5.0 / 12.0; c[5] = 1.0 / 2.0; c[6] = 5.0 / 6.0;
        c[7] = 1.0 / 6.0; c[8] = 2.0 / 3.0; c[9] = 1.0 / 3.0; c[10] = 1.0; c[11] = 0.0; c[12] = 1.0;
        return 0;
    }
    default:
        return -1;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* 2. example systems                                                                               */
/* ------------------------------------------------------------------------------------------------ */

void vo_rhs(int sys, int n, const double *x, const double *p, double t, double *dxdt)
{
    (void)t;
    switch (sys) {
    case VO_SYS_HARMONIC: { /* HarmonicOscillator/main.cpp:27-28, k = 1.0 */
        const double k = 1.0;
        dxdt[0] = x[1];
        dxdt[1] = -k * x[0] - p[0] * x[1];
        break;
    }
    case VO_SYS_VANDERPOL: /* VanDerPol/main.cpp:41-42 */
        dxdt[0] = x[1];
        dxdt[1] = p[0] * ((1.0 - x[0] * x[0]) * x[1] - x[0]);
        break;
    case VO_SYS_GLV: /* GeneralizedLotkaVolterra/main.cpp:109-118: p = [r(n), A(n x n) row-major] */
        for (int i = 0; i < n; ++i) {
            double sum = 0.0;
            const double *Ai = p + (size_t)n * (i + 1);
            for (int j = 0; j < n; ++j) sum += Ai[j] * x[j];
            dxdt[i] = x[i] * (p[i] + sum);
        }
        break;
    }
}

void vo_vjp(int sys, int n, const double *x, const double *p, double t, const double *w, double *gx, double *gp)
{
    (void)t;
    switch (sys) {
    case VO_SYS_HARMONIC: {
        const double k = 1.0;
        gx[0] = -k * w[1];
        gx[1] = w[0] - p[0] * w[1];
        gp[0] += -x[1] * w[1];
        break;
    }
    case VO_SYS_VANDERPOL: {
        const double mu = p[0];
        gx[0] = mu * (-2.0 * x[0] * x[1] - 1.0) * w[1];
        gx[1] = w[0] + mu * (1.0 - x[0] * x[0]) * w[1];
        gp[0] += ((1.0 - x[0] * x[0]) * x[1] - x[0]) * w[1];
        break;
    }
    case VO_SYS_GLV: {
        /* f_i = x_i (r_i + s_i), s = A x.  v = w o x.  gx_k = w_k (r_k + s_k) + sum_i v_i A_ik ; g_r = v ; g_A = v x^T */
        for (int k = 0; k < n; ++k) gx[k] = 0.0;
        for (int i = 0; i < n; ++i) {
            const double *Ai = p + (size_t)n * (i + 1);
            double *gAi = gp + (size_t)n * (i + 1);
            double sum = 0.0;
            for (int j = 0; j < n; ++j) sum += Ai[j] * x[j];
            const double v = w[i] * x[i];
            gx[i] += w[i] * (p[i] + sum);
            for (int j = 0; j < n; ++j) {
                gx[j] += v * Ai[j];
                gAi[j] += v * x[j];
            }
            gp[i] += v;
        }
        break;
    }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* 3. forward sweep                                                                                 */
/* ------------------------------------------------------------------------------------------------ */

/* odeint detail::less_with_sign / less_eq_with_sign (util/detail/less_with_sign.hpp) */
static int less_with_sign(double t1, double t2, double dt)
{
    if (dt > 0) return t2 - t1 > DBL_EPSILON;
    return t1 - t2 > DBL_EPSILON;
}
static int less_eq_with_sign(double t1, double t2, double dt)
{
    if (dt > 0) return t1 - t2 <= DBL_EPSILON;
    return t2 - t1 <= DBL_EPSILON;
}

typedef struct {
    int sys, n;
    const vo_tableau *tb;
    const double *p;
    double *K;    /* [s][n] stage derivatives */
    double *xt;   /* [n] stage state */
    double *xnew; /* [n] */
    double *xerr; /* [n] */
    double *dxdt_new; /* [n] FSAL */
} stepper_ws;

/* explicit_generic_rk::do_step_impl arithmetic: x_m = 1*x + (a_m0 dt) k_0 + (a_m1 dt) k_1 + ...   left to right;
 * K[0] must hold dxdt(x,t) on entry.  Writes xnew and (if has_error) xerr. For FSAL tableaux K[s-1] = f(xnew). */
static void rk_step(stepper_ws *w, const double *x, double t, double dt)
{
    const vo_tableau *tb = w->tb;
    const int n = w->n, s = tb->s;
    const int s_explicit = tb->fsal ? s - 1 : s; /* stages computed through x_m; FSAL last stage is f(xnew) */
    for (int m = 1; m < s_explicit; ++m) {
        for (int i = 0; i < n; ++i) {
            double acc = 1.0 * x[i];
            for (int j = 0; j < m; ++j) acc = acc + (tb->a[m * MS + j] * dt) * w->K[j * n + i];
            w->xt[i] = acc;
        }
        vo_rhs(w->sys, n, w->xt, w->p, t + dt * tb->c[m], w->K + (size_t)m * n);
    }
    for (int i = 0; i < n; ++i) {
        double acc = 1.0 * x[i];
        for (int j = 0; j < s_explicit; ++j) {
            if (tb->fsal && tb->b[j] == 0.0) continue; /* odeint's hand-coded dopri5 skips the zero weight */
            acc = acc + (tb->b[j] * dt) * w->K[j * n + i];
        }
        w->xnew[i] = acc;
    }
    if (tb->fsal) vo_rhs(w->sys, n, w->xnew, w->p, t + dt, w->K + (size_t)(s - 1) * n);
    if (tb->has_error) {
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            int first = 1;
            for (int j = 0; j < s; ++j) {
                if (tb->fsal && tb->db[j] == 0.0) continue;
                const double term = (dt * tb->db[j]) * w->K[j * n + i];
                acc = first ? term : acc + term;
                first = 0;
            }
            w->xerr[i] = acc;
        }
    }
}

long vo_forward(int sys, int n, const vo_tableau *tb, int adaptive, double eps_abs, double eps_rel,
                double *x, const double *p, double ti, double tf, double dt0,
                double *ck_t, double *ck_x, long ck_cap, long *n_reject)
{
    stepper_ws w;
    w.sys = sys; w.n = n; w.tb = tb; w.p = p;
    double *buf = (double *)malloc(sizeof(double) * (size_t)n * (tb->s + 4));
    w.K = buf; w.xt = buf + (size_t)n * tb->s; w.xnew = w.xt + n; w.xerr = w.xnew + n; w.dxdt_new = w.xerr + n;
    long count = 0, rejects = 0, nck = 0, ret = 0;
    double t = ti, dt = dt0;

#define PUSH_CK()                                                        \
    do {                                                                 \
        if (nck >= ck_cap) { ret = -1; goto done; }                      \
        ck_t[nck] = t;                                                   \
        memcpy(ck_x + (size_t)nck * n, x, sizeof(double) * n);           \
        ++nck;                                                           \
    } while (0)

    if (!adaptive) {
        /* detail/runge_kutta.hpp:51-71 */
        long step = 0;
        while (less_eq_with_sign(t + dt, tf, dt)) {
            PUSH_CK();
            vo_rhs(sys, n, x, p, t, w.K);
            rk_step(&w, x, t, dt);
            memcpy(x, w.xnew, sizeof(double) * n);
            ++step;
            t = ti + (double)step * dt;
        }
        PUSH_CK();
        count = step;
    } else {
        /* detail/runge_kutta.hpp:92-117 + controlled_runge_kutta::try_step */
        int first_call = 1;
        while (less_with_sign(t, tf, dt)) {
            PUSH_CK();
            if (less_with_sign(tf, t + dt, dt)) dt = tf - t;
            int trials = 0;
            for (;;) {
                if (!tb->fsal || first_call) { vo_rhs(sys, n, x, p, t, w.K); first_call = 0; }
                rk_step(&w, x, t, dt);
                /* default_error_checker::error, a_x = a_dxdt = 1 */
                double err = 0.0;
                for (int i = 0; i < n; ++i) {
                    const double e = fabs(w.xerr[i]) / (eps_abs + eps_rel * (fabs(x[i]) + fabs(dt) * fabs(w.K[i])));
                    err = fmax(err, e); /* max-norm */
                }
                int fail;
                if (err > 1.0) {
                    /* default_step_adjuster::decrease_step */
                    dt *= fmax(0.9 * pow(err, -1.0 / ((double)tb->error_order - 1.0)), 0.2);
                    fail = 1;
                    ++rejects;
                } else {
                    t += dt;
                    memcpy(x, w.xnew, sizeof(double) * n);
                    if (tb->fsal) memcpy(w.K, w.K + (size_t)(tb->s - 1) * n, sizeof(double) * n);
                    /* default_step_adjuster::increase_step */
                    if (err < 0.5) {
                        err = fmax(pow(5.0, -(double)tb->stepper_order), err);
                        dt *= 9.0 / 10.0 * pow(err, -1.0 / (double)tb->stepper_order);
                    }
                    fail = 0;
                }
                /* failed_step_checker: counts every call, throws when the count reaches 500 */
                if (++trials >= 500 && fail) { ret = -2; goto done; }
                if (!fail) break;
            }
            ++count;
        }
        PUSH_CK();
    }
done:
    free(buf);
    if (n_reject) *n_reject = rejects;
    return ret < 0 ? ret : count;
}

/* ------------------------------------------------------------------------------------------------ */
/* 4. reverse sweep                                                                                 */
/* ------------------------------------------------------------------------------------------------ */

void vo_adjoint(int sys, int n, int npar, const vo_tableau *tb, long T, const double *ck_t, const double *ck_x,
                const double *p, double *lambda, double *mu)
{
    const int s = tb->s_adj;
    double *buf = (double *)malloc(sizeof(double) * ((size_t)n * (2 * s + 4) + (size_t)npar));
    double *K = buf;                       /* [s][n]   */
    double *W = K + (size_t)n * s;         /* [s+1][n] */
    double *xm = W + (size_t)n * (s + 1);  /* [n]      */
    double *gx = xm + n;                   /* [n]      */
    double *alphabar = gx + n;             /* [npar]   */
    (void)npar;
    memset(alphabar, 0, sizeof(double) * (size_t)npar);

    for (long step = T - 1; step >= 0; --step) {
        const double time = ck_t[step];
        const double dt = ck_t[step + 1] - ck_t[step]; /* StateStorage.hpp:22 */
        const double *u = ck_x + (size_t)step * n;
        /* compute_intermediate_states: detail/backpropagation.hpp:37-52 */
        for (int m = 0; m < s; ++m) {
            for (int i = 0; i < n; ++i) xm[i] = u[i];
            for (int j = 0; j < m; ++j)
                for (int i = 0; i < n; ++i) xm[i] += dt * tb->a[m * MS + j] * K[j * n + i];
            vo_rhs(sys, n, xm, p, time, K + (size_t)m * n);
        }
        /* seeds: :183-188 */
        for (int i = 0; i < n; ++i) {
            W[i] = lambda[i];
            for (int m = 1; m <= s; ++m) W[m * n + i] = tb->b[m - 1] * dt * lambda[i];
        }
        /* stages s..1: :201-223 */
        for (int m = s; m > 0; --m) {
            for (int i = 0; i < n; ++i) xm[i] = u[i];
            for (int k = 1; k < m; ++k)
                for (int i = 0; i < n; ++i) xm[i] += dt * tb->a[(m - 1) * MS + (k - 1)] * K[(k - 1) * n + i];
            vo_vjp(sys, n, xm, p, time, W + (size_t)m * n, gx, alphabar);
            for (int i = 0; i < n; ++i) {
                W[i] += gx[i];
                for (int k = 1; k < m; ++k) W[k * n + i] += gx[i] * tb->a[(m - 1) * MS + (k - 1)] * dt;
            }
        }
        for (int i = 0; i < n; ++i) lambda[i] = W[i];
    }
    for (int k = 0; k < npar; ++k) mu[k] += alphabar[k];
    free(buf);
}

/* ------------------------------------------------------------------------------------------------ */
/* 5. batch driver + synthetic inputs                                                               */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    int sys, n, npar, adaptive, objective;
    const vo_tableau *tb;
    double eps_abs, eps_rel, ti, tf, dt0;
    long b0, b1, ck_cap;
    const double *x0, *p;
    double *x_final, *lambda, *mu;
    int32_t *n_accept, *n_reject, *status;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const int n = j->n, npar = j->npar;
    double *ck_t = (double *)malloc(sizeof(double) * (size_t)j->ck_cap);
    double *ck_x = (double *)malloc(sizeof(double) * (size_t)j->ck_cap * n);
    double *x = (double *)malloc(sizeof(double) * n);
    for (long b = j->b0; b < j->b1; ++b) {
        memcpy(x, j->x0 + (size_t)b * n, sizeof(double) * n);
        long rej = 0;
        long T = vo_forward(j->sys, n, j->tb, j->adaptive, j->eps_abs, j->eps_rel, x, j->p + (size_t)b * npar,
                            j->ti, j->tf, j->dt0, ck_t, ck_x, j->ck_cap, &rej);
        if (j->n_reject) j->n_reject[b] = (int32_t)rej;
        double *lam = j->lambda + (size_t)b * n;
        double *mu = j->mu + (size_t)b * npar;
        memset(mu, 0, sizeof(double) * (size_t)npar);
        if (T < 0) {
            if (j->status) j->status[b] = (T == -1) ? 1 : 2;
            if (j->n_accept) j->n_accept[b] = 0;
            for (int i = 0; i < n; ++i) { j->x_final[(size_t)b * n + i] = NAN; lam[i] = NAN; }
            continue;
        }
        if (j->status) j->status[b] = 0;
        if (j->n_accept) j->n_accept[b] = (int32_t)T;
        memcpy(j->x_final + (size_t)b * n, x, sizeof(double) * n);
        if (j->objective == VO_OBJ_SUM) for (int i = 0; i < n; ++i) lam[i] = 1.0;
        else if (j->objective == VO_OBJ_HALF_NORM2) for (int i = 0; i < n; ++i) lam[i] = x[i];
        vo_adjoint(j->sys, n, npar, j->tb, T, ck_t, ck_x, j->p + (size_t)b * npar, lam, mu);
    }
    free(ck_t); free(ck_x); free(x);
    return NULL;
}

int vo_forward_adjoint_batch(int sys, int n, int npar, int stepper, int adaptive, double eps_abs, double eps_rel,
                             long B, const double *x0, const double *p, double ti, double tf, double dt0,
                             int objective, long ck_cap,
                             double *x_final, double *lambda_inout, double *mu_out,
                             int32_t *n_accept, int32_t *n_reject, int32_t *status, int threads)
{
    vo_tableau tb;
    if (vo_tableau_get(stepper, &tb) != 0) return -1;
    if (adaptive && !tb.has_error) return -2;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    batch_job jobs[256];
    pthread_t th[256];
    const long chunk = (B + threads - 1) / threads;
    int used = 0;
    for (int k = 0; k < threads; ++k) {
        long b0 = k * chunk, b1 = b0 + chunk;
        if (b0 >= B) break;
        if (b1 > B) b1 = B;
        batch_job *j = &jobs[used];
        j->sys = sys; j->n = n; j->npar = npar; j->adaptive = adaptive; j->objective = objective; j->tb = &tb;
        j->eps_abs = eps_abs; j->eps_rel = eps_rel; j->ti = ti; j->tf = tf; j->dt0 = dt0;
        j->b0 = b0; j->b1 = b1; j->ck_cap = ck_cap; j->x0 = x0; j->p = p;
        j->x_final = x_final; j->lambda = lambda_inout; j->mu = mu_out;
        j->n_accept = n_accept; j->n_reject = n_reject; j->status = status;
        ++used;
    }
    if (used == 1) {
        batch_worker(&jobs[0]);
    } else {
        for (int k = 0; k < used; ++k) pthread_create(&th[k], NULL, batch_worker, &jobs[k]);
        for (int k = 0; k < used; ++k) pthread_join(th[k], NULL);
    }
    return 0;
}

/* splitmix64 finaliser used as a counter-based generator; pure integer ops + one exact int->double scaling,
 * so host and device produce identical bits (csrc/va_synth.cuh holds the device twin). */
static uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

double vo_u01(uint64_t seed, uint64_t stream, uint64_t b, uint64_t k)
{
    uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ULL * (stream + 1));
    h = mix64(h ^ (b * 0xD1342543DE82EF95ULL));
    h = mix64(h + k);
    return (double)(h >> 11) * 0x1.0p-53;
}

/* Synthetic parameter sets (SURVEY.md section 8d recipe, made bit-reproducible: no transcendental functions).
 *  HARMONIC : mu = 0.151 (1 + 0.5 s),           s = 2u-1
 *  VANDERPOL: mu = 2^floor(10 u1) (1 + u2)      in [1, 1024)  ("stiff-ish" sweep over three decades)
 *  GLV(n)   : r_i = 0.1 (1 + 0.1 s), A_ii = -10 (1 + 0.1 s), A_ij = m z sqrt(10/n), m ~ Bernoulli(1/2),
 *             z = (u1+u2+u3+u4-2) sqrt(3)  (Irwin-Hall approximation of N(0,1)) */
void vo_synth_params(int sys, int n, uint64_t seed, long b0, long B, double *p)
{
    if (sys == VO_SYS_HARMONIC) {
        for (long b = 0; b < B; ++b) {
            const double s = 2.0 * vo_u01(seed, 0, (uint64_t)(b0 + b), 0) - 1.0;
            p[b] = 0.151 * (1.0 + 0.5 * s);
        }
    } else if (sys == VO_SYS_VANDERPOL) {
        for (long b = 0; b < B; ++b) {
            const double u1 = vo_u01(seed, 1, (uint64_t)(b0 + b), 0), u2 = vo_u01(seed, 1, (uint64_t)(b0 + b), 1);
            p[b] = ldexp(1.0 + u2, (int)(10.0 * u1));
        }
    } else {
        const size_t npar = (size_t)n * n + n;
        const double scale = sqrt(10.0 / (double)n), sqrt3 = sqrt(3.0);
        for (long b = 0; b < B; ++b) {
            double *pb = p + (size_t)b * npar;
            const uint64_t bb = (uint64_t)(b0 + b);
            for (int i = 0; i < n; ++i) pb[i] = 0.1 * (1.0 + 0.1 * (2.0 * vo_u01(seed, 2, bb, (uint64_t)i) - 1.0));
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    const uint64_t base = (uint64_t)n + ((uint64_t)i * n + j) * 5;
                    const double u = vo_u01(seed, 2, bb, base);
                    double v;
                    if (i == j) {
                        v = -10.0 * (1.0 + 0.1 * (2.0 * u - 1.0));
                    } else if (u < 0.5) {
                        const double z = (((vo_u01(seed, 2, bb, base + 1) + vo_u01(seed, 2, bb, base + 2)) +
                                           vo_u01(seed, 2, bb, base + 3)) + vo_u01(seed, 2, bb, base + 4)) - 2.0;
                        v = (z * sqrt3) * scale;
                    } else {
                        v = 0.0;
                    }
                    pb[(size_t)n * (i + 1) + j] = v;
                }
        }
    }
}

void vo_synth_x0(int sys, int n, const double *p, long B, double *x0)
{
    for (long b = 0; b < B; ++b) {
        if (sys == VO_SYS_HARMONIC) { /* HarmonicOscillator/main.cpp:49 */
            x0[2 * b] = 0.0; x0[2 * b + 1] = 1.0;
        } else if (sys == VO_SYS_VANDERPOL) { /* VanDerPol/main.cpp:63 */
            const double mu = p[b];
            x0[2 * b] = 2.0;
            x0[2 * b + 1] = -2.0 / 3.0 + 10.0 / (81.0 * mu) - 292.0 / (2187.0 * mu * mu);
        } else { /* GeneralizedLotkaVolterra/main.cpp:156-158 */
            for (int i = 0; i < n; ++i) x0[(size_t)b * n + i] = 0.1;
        }
    }
}
