/*
 * va_oracle.h -- CPU restatement of the VectorizedAdjoint hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle: a plain-C, scalar restatement of
 *   - the forward sweep   (reference lib/include/detail/runge_kutta.hpp:38-118 on top of
 *                          Boost.Odeint 1.74 controlled_runge_kutta / explicit_generic_rk, which is NOT
 *                          vendored in the reference; its published algorithm is restated here),
 *   - the checkpoint store (reference lib/include/StateStorage.hpp:20-42),
 *   - the reverse sweep    (reference lib/include/detail/backpropagation.hpp:24-158, 231-348),
 *   - the three example right-hand sides and their vector-Jacobian products
 *     (reference examples/{HarmonicOscillator,VanDerPol,GeneralizedLotkaVolterra}/main.cpp).
 *
 * Pinning: tests/test_oracle_golden.py checks this file against the 17-digit golden vectors produced by the
 * UNMODIFIED reference headers + AADC (oracle/_ref, built by oracle/Makefile from /root/reference) and against
 * the printed outputs of the reference's prebuilt example binaries (tests/golden/reference_goldens.json).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this
 * library. The product path (vectorizedadjoint_b200/csrc) never links, loads or calls it.
 */
#ifndef VA_ORACLE_H
#define VA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* numeric values are shared with include/va_engine.h on purpose (same enumerators, separate headers) */
enum { VO_SYS_HARMONIC = 0, VO_SYS_VANDERPOL = 1, VO_SYS_GLV = 2 };
enum { VO_RK_EULER = 0, VO_RK_RK4 = 1, VO_RK_CK54 = 2, VO_RK_DOPRI5 = 3, VO_RK_RKF78 = 4 };
enum { VO_OBJ_SEED = 0, VO_OBJ_SUM = 1, VO_OBJ_HALF_NORM2 = 2 };

#define VO_MAX_STAGES 13

typedef struct {
    int s;             /* stages used by the forward step                                   */
    int s_adj;         /* stages that carry weight in the adjoint (dopri5: 6, FSAL stage has b=0) */
    int order, stepper_order, error_order;
    int fsal;          /* first-same-as-last (dopri5)                                        */
    int has_error;     /* error estimate available (can be wrapped by the controller)        */
    double a[VO_MAX_STAGES * VO_MAX_STAGES]; /* row-major, a[m*VO_MAX_STAGES + j], j<m       */
    double b[VO_MAX_STAGES];
    double db[VO_MAX_STAGES];
    double c[VO_MAX_STAGES];
} vo_tableau;

int vo_tableau_get(int kind, vo_tableau *tb);

/* f(x, p, t) for the built-in systems; n = state size */
void vo_rhs(int sys, int n, const double *x, const double *p, double t, double *dxdt);
/* gx = w^T df/dx (overwritten), gp += w^T df/dp (accumulated) */
void vo_vjp(int sys, int n, const double *x, const double *p, double t, const double *w, double *gx, double *gp);

/* Forward sweep for one parameter set.
 *  x[n]       in: x(ti)   out: x(tf)
 *  ck_t/ck_x  checkpoint store, capacity ck_cap entries of (t, x[n]); T+1 entries are written
 *  returns the number of accepted steps T (>=0), -1 on checkpoint overflow, -2 on no-progress (500 rejections) */
long vo_forward(int sys, int n, const vo_tableau *tb, int adaptive, double eps_abs, double eps_rel,
                double *x, const double *p, double ti, double tf, double dt0,
                double *ck_t, double *ck_x, long ck_cap, long *n_reject);

/* Reverse sweep for one seed. lambda[n] in: dJ/dx(tf), out: dJ/dx(ti).  mu[npar] += dJ/dp. */
void vo_adjoint(int sys, int n, int npar, const vo_tableau *tb, long T, const double *ck_t, const double *ck_x,
                const double *p, double *lambda, double *mu);

/* Batched forward+adjoint over B parameter sets (AoS host layout: x0[b*n+i], p[b*npar+k]).
 * objective: VO_OBJ_SEED -> lambda_inout holds the seeds; SUM -> seed = 1; HALF_NORM2 -> seed = x(tf).
 * mu_out[b*npar+k] is overwritten with dJ_b/dp (not accumulated). threads >= 1 (pthreads).
 * status[b]: 0 ok, 1 checkpoint overflow, 2 no progress. */
int vo_forward_adjoint_batch(int sys, int n, int npar, int stepper, int adaptive, double eps_abs, double eps_rel,
                             long B, const double *x0, const double *p, double ti, double tf, double dt0,
                             int objective, long ck_cap,
                             double *x_final, double *lambda_inout, double *mu_out,
                             int32_t *n_accept, int32_t *n_reject, int32_t *status, int threads);

/* Counter-based synthetic inputs, bit-identical to the device generator (csrc/va_synth.cuh). */
double vo_u01(uint64_t seed, uint64_t stream, uint64_t b, uint64_t k);
void vo_synth_params(int sys, int n, uint64_t seed, long b0, long B, double *p /* [B][npar] */);
void vo_synth_x0(int sys, int n, const double *p, long B, double *x0 /* [B][n] */);

#ifdef __cplusplus
}
#endif
#endif
