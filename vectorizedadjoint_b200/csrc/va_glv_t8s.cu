// va_glv_t8s.cu -- GLV forward + discrete-adjoint kernel for 33..64 species, third generation of the headline path (GLV N = 64,
//                   one million parameter sets): the sweeps and the gradient accumulation run on DIFFERENT warps of the CTA.
//
// f_i = x_i (r_i + (A x)_i), parameters [r, A row-major] (reference examples/GeneralizedLotkaVolterra/main.cpp:105-119).
// Same algorithm and the same arithmetic as va_glv_t8.cu (reference lib/include/detail/runge_kutta.hpp:76-118 forward + odeint
// controlled stepper; detail/backpropagation.hpp:83-158, 231-254 reverse) -- every sum is formed in the same order, so the two
// kernels return the same bits. What changed is who does what, chosen from the ncu profile of the second generation
// (profiles/r02/glv64_t8_ncu_full_128k.txt): its forward sweep and state-adjoint sweep are chains of matrix-vector products, each a
// ~155-cycle DFMA burst inside a ~370-cycle exchange -> load -> reduce chain, and with 252 registers per thread (the 64-entry
// matrix tile, later the 64-entry accumulator tile) only two warps fit an SM sub-partition: the FP64 pipe idles ~60 % of the
// sweeps and only the accumulation phase fills it. A third warp per sub-partition needs registers the 252-register design does
// not have. Here:
//   * warps 0..7 ("sweep warps", two per sub-partition, 64 threads per trajectory, four trajectories per CTA as before) run
//     phase 1 (forward sweep, A tile in registers) and phase 2 (state adjoint, A^T tile in the same registers) and NEVER hold an
//     accumulator tile: they need ~200 registers;
//   * warps 8..11 ("accumulate warps", one per sub-partition) run phase 3, Abar += v_m X_{m-1}^T, for all four trajectory slots
//     of the CTA, 32 accumulators per thread (8 x 4 tile, 128 threads per trajectory): 104 registers;
//   * the CTA's registers are re-partitioned at run time with setmaxnreg (2 x 200 + 104 = 504 per lane of a sub-partition);
//   * hand-over through the slab: a slot that finished phase 2 posts a job (slot, steps, slab half) in a shared-memory queue
//     (mbarrier per entry) and goes on to its NEXT trajectory at once; the accumulate warps stream the step blocks by TMA through
//     their own ring (full / empty mbarriers, no block-level barrier) and fill the FP64 pipe while the sweep warps wait for
//     exchanges. A slot's slab has two halves used alternately, so the forward sweep of trajectory i + 1 does not overwrite the
//     blocks trajectory i's accumulation is still reading; a slot waits (rare) for the job of trajectory i - 1 before it starts
//     trajectory i + 1.
// Step blocks as in va_glv_t8.cu: [8-double header (t_n) | X_0..X_{s-1} | g_0..g_{s-1} | v_1..v_s]; with one seed per trajectory
// v aliases g.
#include <cstdlib>

#include "va_glv_common.cuh"
#include "va_tma.cuh"

#ifndef VA_T8S_REGS_SWEEP
#define VA_T8S_REGS_SWEEP 200 // registers per sweep thread after setmaxnreg
#endif
#ifndef VA_T8S_REGS_ACC
#define VA_T8S_REGS_ACC 104 // registers per accumulate thread after setmaxnreg
#endif
#ifndef VA_T8S_NB
#define VA_T8S_NB 4 // step-block buffers per sweep slot (phase 2)
#endif
#ifndef VA_T8S_NBA
#define VA_T8S_NBA 4 // step-block buffers of the accumulate warps
#endif
#ifndef VA_T8S_ACC_POLICY
#define VA_T8S_ACC_POLICY 0 // L2 policy of the accumulate warps' block reads (the last use of a block): 0 = evict_last, 1 = evict_first
#endif
// setmaxnreg moves registers inside the pool the CTA was given at launch: 384 threads x 168 registers (the launch bound) = 504 per
// lane of a sub-partition -- NOT the 512 of the register file. A split that needs more never gets its last increase and hangs.
static_assert(2 * VA_T8S_REGS_SWEEP + VA_T8S_REGS_ACC <= 504 && VA_T8S_REGS_SWEEP % 8 == 0 && VA_T8S_REGS_ACC % 8 == 0, "register split");

namespace {

constexpr int NP = 64;  // padded species count
constexpr int NTT = 64; // threads per trajectory in the sweeps
constexpr int SLOTS = 4;
constexpr int NSW = NTT * SLOTS; // sweep threads
constexpr int NAC = 128;         // accumulate threads
constexpr int NT = NSW + NAC;
constexpr int HDR = 8; // doubles in a step-block header (hdr[0] = t_n)
constexpr int NB = VA_T8S_NB, NBA = VA_T8S_NBA;
constexpr int RT = 4, CT = 16; // matrix tile of a sweep thread: 4 rows x 16 columns
constexpr int QN = 16;         // job queue entries (a slot has at most two jobs + its exit token outstanding)

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

using va_tma::bulk_g2s;
using va_tma::fence_proxy_async;
using va_tma::ldg_hint;
using va_tma::mbar_expect_tx;
using va_tma::mbar_init;
using va_tma::mbar_wait;
using va_tma::policy_evict_first;
using va_tma::policy_evict_last;
using va_tma::smem_u32;
__device__ __forceinline__ void st_hint(double *p, double v, uint64_t policy) { va_tma::st_hint_relaxed(p, v, policy); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// max over the warp of non-negative doubles (or NaN, which orders above everything): two 32-bit redux operations
__device__ __forceinline__ double warp_max_nonneg(double e)
{
    const unsigned hi = (unsigned)__double2hiint(e), lo = (unsigned)__double2loint(e);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}

// e^(-1/P), the controller's step-size root (see va_glv_t8.cu)
template <int P>
__device__ __forceinline__ double inv_root_short(double e)
{
    e = fmin(e, 1e30);
    double y = (double)__powf((float)e, -1.0f / (float)P);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double y2 = y * y;
        const double yp = P == 1 ? y : P == 2 ? y2 : P == 3 ? y2 * y : P == 4 ? y2 * y2 : P == 5 ? (y2 * y2) * y : (y2 * y2) * y2;
        static_assert(P >= 1 && P <= 6, "unsupported order");
        y = fma(y * (1.0 / P), fma(-e, yp, 1.0), y);
    }
    return y;
}

struct SharedState {
    double xs[SLOTS][2][NP]; // operand vector of the current matrix-vector product (double buffered), per slot
    double red[SLOTS][2];
    int st[SLOTS][2];
    uint64_t mbar[SLOTS][NB]; // phase 2: step block landed
    uint64_t abar[NBA];       // accumulate ring: step block landed
    uint64_t ebar[NBA];       // accumulate ring: buffer released by the 128 accumulate threads
    uint64_t qbar[QN];        // job queue: entry written
    int qe[QN][8];            // {slot | exit flag, steps, slab half, seed, trajectory lo, trajectory hi}
    int q_tail;
    int done[SLOTS]; // jobs of this slot the accumulate warps have finished
};

template <class Tab, bool ADAPTIVE, bool EXACT>
__global__ void __launch_bounds__(NT, 1) k_glv_t8s(const __grid_constant__ VaGlvWideArgs a)
{
    constexpr int S = Tab::S, SADJ = Tab::SADJ;
    constexpr int SE = Tab::FSAL ? S - 1 : S; // stages evaluated through an intermediate state
    extern __shared__ __align__(128) double xg_all[]; // SLOTS * NB sweep buffers, then NBA accumulate buffers, of a.blk_doubles
    __shared__ __align__(16) SharedState sh;

    const int wc = threadIdx.x >> 5;
    const int n = a.n;
    const int npar = n * n + n;
    const int cap = a.cap;
    const int blk = a.blk_doubles;
    const int voff = blk - SADJ * NP;                    // v section of a step block
    const uint32_t xg_bytes = (HDR + 2 * SADJ * NP) * 8; // header, X and g
    const bool vsep = voff != HDR + SADJ * NP;           // v has its own section (several seeds per trajectory)
    const int64_t half_stride = (int64_t)(cap + 1) * blk;
    const uint64_t keep = policy_evict_last();

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
#pragma unroll
            for (int i = 0; i < NB; ++i) mbar_init(&sh.mbar[s][i], 1);
            sh.done[s] = 0;
        }
#pragma unroll
        for (int i = 0; i < NBA; ++i) {
            mbar_init(&sh.abar[i], 1);
            mbar_init(&sh.ebar[i], NAC);
        }
#pragma unroll
        for (int i = 0; i < QN; ++i) mbar_init(&sh.qbar[i], 1);
        sh.q_tail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // partial-sum rows (summed objective): every thread zeroes entries it or its role later adds to
    if (a.reduce == VA_REDUCE_SUM && a.n_out > 0) {
        if (wc < NSW / 32) {
            const int slot = ((wc >> 2) << 1) | ((wc >> 1) & 1), own = (wc & 1) * 32 + (threadIdx.x & 31);
            if (own < n) a.partial[((int64_t)blockIdx.x * SLOTS + slot) * npar + own] = 0.0;
        } else {
            const int at = threadIdx.x - NSW;
            for (int s = 0; s < SLOTS; ++s) {
                double *const part = a.partial + ((int64_t)blockIdx.x * SLOTS + s) * npar + n;
                for (int k = at; k < n * n; k += NAC) part[k] = 0.0;
            }
        }
    }
    __syncthreads(); // the only block-wide barrier: from here on the two roles meet through mbarriers only

    if (wc < NSW / 32) {
        // =====================================================================================================================
        // sweep warps: phases 1 and 2
        // =====================================================================================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(VA_T8S_REGS_SWEEP));
        // Warp w of the CTA runs on SM sub-partition w % 4. Slot s takes the warp pair {2 s, 2 s + 1}: the two warps of a trajectory
        // sit on different sub-partitions, and every sub-partition hosts one warp of two different trajectories (+ one accumulate warp).
        const int slot = ((wc >> 2) << 1) | ((wc >> 1) & 1);
        const int warp = wc & 1;                        // warp inside the trajectory
        const int tid = warp * 32 + (threadIdx.x & 31); // thread inside the trajectory
        const int g = tid & (RT - 1);                   // lane inside the reduction group (RT lanes)
        const int hi = tid / RT;                        // reduction group
        const int own = tid;                            // vector component this thread owns after a reduction
        const int64_t gslot = (int64_t)blockIdx.x * SLOTS + slot; // global slot: owns one slab and one partial-sum row
        double *const slab0 = a.slab + gslot * a.slab_stride;
        double *const xg = xg_all + (size_t)slot * NB * blk;
        double(*xs)[NP] = sh.xs[slot];
        double *red = sh.red[slot];
        uint64_t *mbar = sh.mbar[slot];
        const double tf = a.tf;
        double *const part = a.partial + gslot * npar;

        auto slot_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory"); };
        auto slot_or = [&](int v) -> int {
            v = __reduce_or_sync(0xffffffffu, v);
            if ((tid & 31) == 0) sh.st[slot][warp] = v;
            slot_sync();
            return sh.st[slot][0] | sh.st[slot][1];
        };
        // the accumulate warps have finished `target` jobs of this slot (satisfied on arrival in the common case). The counter is
        // read and written with shared-memory atomics only: a flag protocol, not a data race
        auto wait_done = [&](int target) {
            while (atomicAdd(&sh.done[slot], 0) < target) __nanosleep(200);
        };
        auto post = [&](int slot_flags, int T, int half, int o, int64_t b) {
            const int tk = atomicAdd(&sh.q_tail, 1);
            int *e = sh.qe[tk % QN];
            e[0] = slot_flags; e[1] = T; e[2] = half; e[3] = o;
            e[4] = (int)(uint32_t)(b & 0xffffffffll); e[5] = (int)(b >> 32);
            mbar_arrive(&sh.qbar[tk % QN]); // release: the entry is visible to whoever observes the phase
        };
        auto PO = [&](int j) { return g + RT * j; };
        auto FG = [&](int e) { return 2 * PO(e >> 1) + (e & 1); };
        uint32_t mbar_parity = 0; // bit i: parity of the next completion of mbar[i]

        // y_own = sum_c M[k][c] xin[FG(c)] summed over the group of four lanes; see va_glv_t8.cu (register row k <-> tile row k ^ g,
        // one round of three independent shuffles; y = c1 * sum + c0 is what the NEXT product needs)
        auto matvec = [&](const double(&M)[RT][CT], double X, int p, auto &&after_sync, auto &&extra, const double &c1, const double &c0,
                          double &y) -> double {
            xs[p][own] = X;
            slot_sync();
            after_sync();
            const double2 *xv = reinterpret_cast<const double2 *>(xs[p]);
            double s[RT];
#pragma unroll
            for (int j = 0; j < CT / 2; ++j) {
                const double2 v = xv[PO(j)];
                if (j == 0) {
#pragma unroll
                    for (int k = 0; k < RT; ++k) s[k] = M[k][0] * v.x;
                } else {
#pragma unroll
                    for (int k = 0; k < RT; ++k) s[k] = fma(M[k][2 * j], v.x, s[k]);
                }
#pragma unroll
                for (int k = 0; k < RT; ++k) s[k] = fma(M[k][2 * j + 1], v.y, s[k]);
            }
            extra();
            double t[RT];
            t[0] = s[0];
#pragma unroll
            for (int j = 1; j < RT; ++j) t[j] = shx(s[j], j);
            const double q = t[2] + t[3];
            y = fma(c1, q, fma(c1, t[1], fma(c1, t[0], c0)));
            return (t[0] + t[1]) + q;
        };
        auto nop = [] {};
        const double zero = 0.0;
        double ydummy;
        int posted = 0;             // jobs this slot has posted
        int posted_at0 = 0, posted_at1 = 0; // ... when slab half 0 / 1 was last handed over
        int traj = 0;               // trajectories this slot has started

        for (int64_t b = gslot; b < a.B; b += (int64_t)gridDim.x * SLOTS, ++traj) {
            const int half = traj & 1;
            double *const slab = slab0 + half * half_stride;
            const double *pb = a.params + b * npar;
            double M[RT][CT];
            // ================================ phase 1: forward sweep =====================================
            // tile rows RT hi + (k ^ g), columns FG(c)
#pragma unroll
            for (int k = 0; k < RT; ++k) {
                const int row = RT * hi + (k ^ g);
                if (EXACT) {
                    const double2 *src = reinterpret_cast<const double2 *>(pb + NP + row * NP);
#pragma unroll
                    for (int j = 0; j < CT / 2; ++j) {
                        const double2 v = __ldg(src + PO(j));
                        M[k][2 * j] = v.x;
                        M[k][2 * j + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const int col = FG(c);
                        // padded entries: the load goes to a valid address and is discarded
                        const bool in = row < n && col < n;
                        const double v = __ldg(pb + n + (in ? row * n + col : 0));
                        M[k][c] = in ? v : 0.0;
                    }
                }
            }
            double r_own = 0.0, x = 0.0;
            {
                const int oi = own < n ? own : 0; // padded lanes read a valid address and discard it
                const double rv = __ldg(pb + oi), xv0 = __ldg(a.x0 + b * n + oi);
                if (own < n) { r_own = rv; x = xv0; }
            }

            double t = a.ti, dt = a.dt0, K[S], g0;
            int nck = 0, rejects = 0, status = 0, trials = 0;
            bool fresh = true;
            {
                const double sum = matvec(M, x, 0, nop, nop, zero, zero, ydummy);
                g0 = r_own + sum;
                K[0] = x * g0;
            }
            bool act = ADAPTIVE ? va_less_with_sign(t, tf, dt) : va_less_eq_with_sign(t + dt, tf, dt);
            double *sp = slab + HDR + own; // this thread's column in the current step block (advanced on acceptance)
            if (a.skip_forward) {
                // split API (va_forward_batch then va_adjoint_batch): the step blocks of this very trajectory are still in the slab
                act = false;
                nck = a.n_accept[b];
                status = a.status[b];
                sp += (int64_t)nck * blk;
                t = sp[-HDR - own];
                const double xf = a.x_final[b * n + (own < n ? own : 0)];
                x = own < n ? xf : 0.0;
            } else {
                wait_done(half ? posted_at1 : posted_at0); // the accumulation of the trajectory that used this half of the slab has read its blocks
            }

            while (act) {
                if (fresh) {
                    if (nck >= cap) { status |= VA_TRAJ_CKPT_OVERFLOW; break; }
                    st_hint(sp, x, keep);
                    st_hint(sp + SADJ * NP, g0, keep);
                    if (tid == 0) st_hint(sp - HDR, t, keep); // own == 0: header of the current block
                    if (ADAPTIVE && va_less_with_sign(tf, t + dt, dt)) dt = tf - t;
                    trials = 0;
                    fresh = false;
                }
                // Stage m produces K_m = X_m (r + A X_m). The state of the NEXT stage (or the new solution after the last
                // one), Y = x + dt sum_{j<=m} c_j K_j, is split so that only ONE DFMA follows the reduction:
                //   Y = fma(c1, sum, base),  c1 = dt c_m X_m,  base = x + dt sum_{j<m} c_j K_j + c1 r   (all known early).
                double X = fma(dt * a.coef.a[1][0], K[0], x);
                double perr = 0.0; // sum_{j<SE-1} db_j K_j
#pragma unroll
                for (int m = 1; m < SE; ++m) {
                    const bool last = (m == SE - 1);
                    double c1 = 0.0, base = 0.0;
                    double Y;
                    const double sum = matvec(M, X, m & 1, nop, [&] {
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < m; ++j) {
                            const double cz = last ? Tab::b(j) : Tab::a(m + 1, j);
                            if (cz != 0.0) acc = fma(last ? a.coef.b[j] : a.coef.a[m + 1][j], K[j], acc);
                        }
                        const double cm = last ? Tab::b(m) : Tab::a(m + 1, m);
                        c1 = (cm != 0.0) ? (dt * (last ? a.coef.b[m] : a.coef.a[m + 1][m])) * X : 0.0;
                        base = fma(c1, r_own, fma(dt, acc, x));
                        if (last && ADAPTIVE) {
#pragma unroll
                            for (int j = 0; j < m; ++j)
                                if (Tab::db(j) != 0.0) perr = fma(a.coef.db[j], K[j], perr);
                        }
                    }, c1, base, Y);
                    const double gg = r_own + sum;
                    K[m] = X * gg;
                    if (m < SADJ) { st_hint(sp + m * NP, X, keep); st_hint(sp + (SADJ + m) * NP, gg, keep); }
                    X = Y;
                }
                // X = new solution. f(xnew) is evaluated now: the FSAL stage of dopri5, and for the other steppers the first slope
                // of the next step (speculative: discarded if the step is rejected). default_error_checker::error, max norm over
                // species: each warp's maximum crosses to the other warp together with the operands of this product.
                constexpr bool ERR_EARLY = ADAPTIVE && !(Tab::FSAL && Tab::db(S - 1) != 0.0);
                auto err_local = [&]() {
                    double acc = perr;
#pragma unroll
                    for (int j = SE - 1; j < S; ++j)
                        if (Tab::db(j) != 0.0) acc = fma(a.coef.db[j], K[j], acc);
                    const double xerr = dt * acc;
                    double e = fabs(xerr) / (a.eps_abs + a.eps_rel * (fabs(x) + fabs(dt) * fabs(K[0])));
                    if (!(own < n)) e = 0.0;
                    e = warp_max_nonneg(e);
                    if ((tid & 31) == 0) red[warp] = e;
                };
                if (ERR_EARLY) err_local();
                const double gl = r_own + matvec(M, X, SE & 1, nop, nop, zero, zero, ydummy);
                const double Kl = X * gl;
                if (Tab::FSAL) K[S - 1] = Kl;
                double err = 0.0;
                if (ADAPTIVE) {
                    if (!ERR_EARLY) {
                        err_local();
                        slot_sync();
                    }
                    const long long e0 = __double_as_longlong(red[0]), e1 = __double_as_longlong(red[1]);
                    err = __longlong_as_double(e0 > e1 ? e0 : e1);
                }
                const bool accept = !ADAPTIVE || !(err > 1.0);
                if (!accept) {
                    // default_step_adjuster::decrease_step
                    dt *= fmax(0.9 * inv_root_short<(Tab::ERROR_ORDER > 1 ? Tab::ERROR_ORDER - 1 : 1)>(err), 0.2);
                    ++rejects;
                    if (++trials >= 500) { status |= VA_TRAJ_NO_PROGRESS; break; }
                } else {
                    x = X;
                    ++nck;
                    sp += blk;
                    if (ADAPTIVE) {
                        t += dt;
                        // default_step_adjuster::increase_step
                        if (err < 0.5) {
                            constexpr int P = Tab::STEPPER_ORDER;
                            double floor_ = 1.0;
#pragma unroll
                            for (int k = 0; k < P; ++k) floor_ *= 0.2; // 5^-P
                            // err <= 5^-P: the growth factor is exactly 0.9 * 5 (pow(5^-P, -1/P) == 5 in glibc as well)
                            dt *= (err <= floor_) ? 4.5 : 9.0 / 10.0 * inv_root_short<P>(err);
                        }
                        act = va_less_with_sign(t, tf, dt);
                    } else {
                        t = a.ti + (double)nck * dt; // detail/runge_kutta.hpp:64
                        act = va_less_eq_with_sign(t + dt, tf, dt);
                    }
                    fresh = true;
                    g0 = gl;
                    K[0] = Kl;
                }
            }
            // close the trajectory: final time, status, x(tf)
            const int T = nck;
            if (tid == 0 && !a.skip_forward) st_hint(sp - HDR, t, keep); // header of block T carries the final time
            if (own < n && !isfinite(x)) status |= VA_TRAJ_NONFINITE;
            fence_proxy_async(); // generic-proxy slab writes -> visible to the TMA reads of the reverse sweep
            status = slot_or(status);
            const bool failed = status & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS);
            const double x_tf = x, t_final = t;
            if (own < n && !a.skip_forward) a.x_final[b * n + own] = failed ? nan("") : x;
            if (tid == 0 && !a.skip_forward) {
                if (a.n_accept) a.n_accept[b] = T;
                if (a.n_reject) a.n_reject[b] = rejects;
                if (a.status) a.status[b] = status;
            }
            if (a.n_out <= 0) continue;

            // ================================ phase 2: adjoint of the state =====================================
            for (int o = 0; o < a.n_out; ++o) {
                double *lam_io = a.lambda + (b * a.n_out + o) * n;
                double *mu_o = a.mu + (a.reduce == VA_REDUCE_SUM ? (int64_t)o : (b * a.n_out + o)) * npar;
                if (failed) {
                    if (own < n) lam_io[own] = nan("");
                    if (a.reduce == VA_REDUCE_NONE)
                        for (int k = tid; k < npar; k += NTT) mu_o[k] = nan("");
                    continue;
                }
                const uint64_t drop = policy_evict_first(); // last read of this parameter set by this seed
                // transposed tile: M[k][c] = A[FG(c)][RT hi + (k ^ g)]; the owned component stays `own`. A is re-read (L2 hit).
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    const int row = FG(c);
#pragma unroll
                    for (int k = 0; k < RT; ++k) {
                        const int col = RT * hi + (k ^ g);
                        if (EXACT) M[k][c] = ldg_hint(pb + NP + row * NP + col, drop);
                        else {
                            const bool in = row < n && col < n;
                            const double v = ldg_hint(pb + n + (in ? row * n + col : 0), drop);
                            M[k][c] = in ? v : 0.0;
                        }
                    }
                }
                double lam;
                if (a.objective == VA_OBJ_SUM) lam = (own < n) ? 1.0 : 0.0;
                else if (a.objective == VA_OBJ_HALF_NORM2) lam = x_tf;
                else lam = (own < n) ? lam_io[own < n ? own : 0] : 0.0;
                double rbar = 0.0;

                // stream the step blocks back, newest first: iteration it <-> step T-1-it, buffer it % NB, NB-1 blocks ahead
                auto issue2 = [&](int it) {
                    if (it < T) {
                        const int bi = it % NB;
                        mbar_expect_tx(&mbar[bi], xg_bytes);
                        bulk_g2s(xg + bi * blk, slab + (int64_t)(T - 1 - it) * blk, xg_bytes, &mbar[bi], keep);
                    }
                };
                slot_sync(); // every thread is past its reads of the buffers (previous seed / trajectory)
                if (tid == 0)
                    for (int it = 0; it < NB - 1; ++it) issue2(it);
                double t_hi = t_final;
                for (int it = 0; it < T; ++it) {
                    const int step = T - 1 - it, bi = it % NB;
                    mbar_wait(&mbar[bi], (mbar_parity >> bi) & 1);
                    mbar_parity ^= 1u << bi;
                    const double *bs = xg + bi * blk;
                    double *gv = slab + (int64_t)step * blk + voff + own; // v_1..v_s of this step, this thread's column
                    const double t_lo = bs[0];
                    const double dt_s = t_hi - t_lo; // StateStorage::GetDt: difference of the stored times
                    t_hi = t_lo;
                    double W[SADJ + 1];
                    W[0] = lam;
#pragma unroll
                    for (int m = 1; m <= SADJ; ++m) W[m] = Tab::b(m - 1) != 0.0 ? (a.coef.b[m - 1] * dt_s) * lam : 0.0;
                    // v = w_m o X_{m-1} for the stage about to be processed; later stages get it from the previous one
                    double v = W[SADJ] * bs[HDR + (SADJ - 1) * NP + own];
#pragma unroll
                    for (int m = SADJ; m >= 1; --m) {
                        st_hint(gv + (m - 1) * NP, v, keep);
                        double wg = 0.0, c1 = 0.0, c2 = 0.0;
                        double v_next;
                        const double sum = matvec(
                            M, v, m & 1,
                            [&] {
                                // every thread is past its reads of the previous iteration's buffer: refill it
                                if (m == SADJ && tid == 0) issue2(it + NB - 1);
                            },
                            [&] {
                                // gx = (A^T v)_own + w_m g_{m-1}. The next stage's v = (w_{m-1} + gx a dt) X_{m-2} is arranged
                                // as fma(sum, c1, c2) with c1, c2 known before the reduction returns
                                wg = W[m] * bs[HDR + (SADJ + m - 1) * NP + own];
                                if (m > 1) {
                                    const double Xn = bs[HDR + (m - 2) * NP + own];
                                    if (Tab::a(m - 1, m - 2) != 0.0) c1 = (a.coef.a[m - 1][m - 2] * dt_s) * Xn;
                                    c2 = fma(wg, c1, W[m - 1] * Xn);
                                }
                            }, c1, c2, v_next);
                        const double gx = sum + wg;
                        const double gxd = gx * dt_s;
                        rbar += v;
                        W[0] += gx;
#pragma unroll
                        for (int k = 1; k < m; ++k)
                            if (Tab::a(m - 1, k - 1) != 0.0) W[k] = fma(gxd, a.coef.a[m - 1][k - 1], W[k]);
                        v = v_next;
                    }
                    lam = W[0];
                }
                if (own < n) {
                    lam_io[own] = lam;
                    if (a.reduce == VA_REDUCE_NONE) mu_o[own] = rbar;
                    else atomicAdd(part + own, rbar);
                }
                // ---- hand the step blocks (X_{m-1}, v_m) to the accumulate warps and go on
                fence_proxy_async(); // the v sections were written through the generic proxy
                slot_sync();
                if (tid == 0) post(slot, T, half, o, b);
                ++posted;
                if (a.n_out > 1) wait_done(posted); // the next seed overwrites the v sections
            }
            if (half) posted_at1 = posted;
            else posted_at0 = posted;
        }
        // exit token: the accumulate warps leave after the last job of every slot
        if (tid == 0) post(slot | 0x100, 0, 0, 0, 0);
    } else {
        // =====================================================================================================================
        // accumulate warps: phase 3 for every slot of the CTA. Abar[i][j] += v_m[i] X_{m-1}[j] over all steps (newest first) and
        // stages (s..1) -- the order of va_glv_t8.cu. Thread (h8, g16) keeps rows 2 (h8 + 8 j) + {0, 1}, j < 4, and columns
        // 2 (g16 + 16 j) + {0, 1}, j < 2: eight v operands (four LDS.128, two distinct addresses per warp) and four X operands
        // (two LDS.128 over 256 contiguous bytes) feed 32 DFMAs.
        // =====================================================================================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(VA_T8S_REGS_ACC));
        const int at = threadIdx.x - NSW;
        const int h8 = at >> 4, g16 = at & 15;
        double *const ring = xg_all + (size_t)SLOTS * NB * blk;
        auto ROW = [&](int k) { return 2 * (h8 + 8 * (k >> 1)) + (k & 1); };
        auto COL = [&](int c) { return 2 * (g16 + 16 * (c >> 1)) + (c & 1); };
        const uint64_t last_use = VA_T8S_ACC_POLICY ? policy_evict_first() : keep;
        uint32_t n_issued = 0, n_used = 0; // step blocks requested (thread 0) / consumed (every thread) since the kernel started
        auto issue3 = [&](const double *src) {
            const uint32_t k = n_issued++;
            const int bj = k % NBA;
            if (k >= NBA) mbar_wait(&sh.ebar[bj], ((k / NBA) - 1) & 1); // every accumulate thread has released the block that was there
            double *dst = ring + (size_t)bj * blk;
            if (vsep) {
                mbar_expect_tx(&sh.abar[bj], (HDR + 2 * SADJ * NP) * 8);
                bulk_g2s(dst, src, (HDR + SADJ * NP) * 8, &sh.abar[bj], last_use);
                bulk_g2s(dst + voff, src + voff, SADJ * NP * 8, &sh.abar[bj], last_use);
            } else {
                mbar_expect_tx(&sh.abar[bj], xg_bytes);
                bulk_g2s(dst, src, xg_bytes, &sh.abar[bj], last_use);
            }
        };
        int exits = 0;
        for (uint32_t q = 0;; ++q) {
            const int e = q % QN;
            mbar_wait(&sh.qbar[e], (q / QN) & 1);
            const int slot_f = sh.qe[e][0], T = sh.qe[e][1], half = sh.qe[e][2], o = sh.qe[e][3];
            const int64_t b = (int64_t)(uint32_t)sh.qe[e][4] | ((int64_t)sh.qe[e][5] << 32);
            if (slot_f & 0x100) {
                if (++exits == SLOTS) break;
                continue;
            }
            const int slot = slot_f;
            const int64_t gslot = (int64_t)blockIdx.x * SLOTS + slot;
            const double *slab = a.slab + gslot * a.slab_stride + half * half_stride;
            if (at == 0)
                for (int it = 0; it < NBA - 1 && it < T; ++it) issue3(slab + (int64_t)(T - 1 - it) * blk);
            double Ab[8][4];
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int c = 0; c < 4; ++c) Ab[k][c] = 0.0;
            for (int it = 0; it < T; ++it) {
                if (at == 0 && it + NBA - 1 < T) issue3(slab + (int64_t)(T - NBA - it) * blk);
                const uint32_t ku = n_used++;
                const int bi = ku % NBA;
                mbar_wait(&sh.abar[bi], (ku / NBA) & 1);
                const double *bs = ring + (size_t)bi * blk;
#pragma unroll
                for (int m = SADJ; m >= 1; --m) {
                    const double2 *vv = reinterpret_cast<const double2 *>(bs + voff + (m - 1) * NP) + h8;
                    const double2 *xx = reinterpret_cast<const double2 *>(bs + HDR + (m - 1) * NP) + g16;
                    double vr[8], xc[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double2 p = vv[8 * j];
                        vr[2 * j] = p.x; vr[2 * j + 1] = p.y;
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const double2 p = xx[16 * j];
                        xc[2 * j] = p.x; xc[2 * j + 1] = p.y;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)
#pragma unroll
                        for (int c = 0; c < 4; ++c) Ab[k][c] = fma(vr[k], xc[c], Ab[k][c]);
                }
                mbar_arrive(&sh.ebar[bi]); // this thread is done with the buffer (released when all 128 have arrived)
            }
            // every block of the job has landed (thread 0 waited for each): the slot may overwrite this half of its slab
            if (at == 0) atomicAdd(&sh.done[slot], 1);
            if (a.reduce == VA_REDUCE_NONE) {
                double *mu_o = a.mu + (b * a.n_out + o) * npar;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int row = ROW(k);
                    if (EXACT) {
                        double2 *dst = reinterpret_cast<double2 *>(mu_o + NP + row * NP) + g16;
#pragma unroll
                        for (int j = 0; j < 2; ++j) dst[16 * j] = make_double2(Ab[k][2 * j], Ab[k][2 * j + 1]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int col = COL(c);
                            if (row < n && col < n) mu_o[n + row * n + col] = Ab[k][c];
                        }
                    }
                }
            } else {
                // summed objective: fire-and-forget FP64 reductions into the slot's partial-sum row (one writer per address,
                // jobs of a slot in trajectory order -> deterministic)
                double *const part = a.partial + gslot * npar;
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int row = ROW(k), col = COL(c);
                        if (row < n && col < n) atomicAdd(part + n + row * n + col, Ab[k][c]);
                    }
            }
        }
    }
}

template <class Tab, bool ADAPTIVE, bool EXACT>
cudaError_t launch_k(const VaGlvWideArgs &a, cudaStream_t st)
{
    const size_t smem = (size_t)(SLOTS * NB + NBA) * a.blk_doubles * 8;
    cudaError_t e = cudaFuncSetAttribute(k_glv_t8s<Tab, ADAPTIVE, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_glv_t8s<Tab, ADAPTIVE, EXACT><<<a.grid, NT, smem, st>>>(a);
    return cudaGetLastError();
}

template <class Tab, bool ADAPTIVE>
cudaError_t launch(const VaGlvWideArgs &a_in, cudaStream_t st)
{
    VaGlvWideArgs a = a_in;
    // tableau values travel in the kernel arguments (constant bank); the compile-time copy in Tab:: only decides which terms exist
    for (int m = 0; m < Tab::S; ++m) {
        for (int j = 0; j < m && j < 6; ++j) a.coef.a[m][j] = Tab::a(m, j);
        a.coef.b[m] = Tab::b(m);
        a.coef.db[m] = Tab::db(m);
    }
    return a.n == NP ? launch_k<Tab, ADAPTIVE, true>(a, st) : launch_k<Tab, ADAPTIVE, false>(a, st);
}

int sadj_of(int stepper)
{
    switch (stepper) {
    case VA_RK_RK4: return TabRK4::SADJ;
    case VA_RK_CK54: return TabCK54::SADJ;
    case VA_RK_DOPRI5: return TabDOPRI5::SADJ;
    }
    return 0;
}

} // namespace

bool va_glv_t8s_supported(int n, int stepper, int adaptive)
{
    if (n <= 32 || n > NP) return false;
    if (stepper == VA_RK_RK4) return !adaptive;
    if (stepper == VA_RK_CK54 || stepper == VA_RK_DOPRI5) return adaptive != 0;
    return false;
}

// step block: [8-double header | X_0..X_{s-1} | g_0..g_{s-1} | v_1..v_s]; with one seed per trajectory v aliases g
int va_glv_t8s_block_doubles(int stepper, int n_out) { return HDR + (n_out > 1 ? 3 : 2) * sadj_of(stepper) * NP; }

// a slot's slab has two halves of (cap + 1) blocks, used by its trajectories alternately
int64_t va_glv_t8s_slab_doubles(int stepper, int n_out, int cap) { return 2 * (int64_t)(cap + 1) * va_glv_t8s_block_doubles(stepper, n_out); }

cudaError_t va_glv_t8s_config(int n, int stepper, int n_out, int device, int *grid, int *ctas_per_sm, int *threads, int *slots_per_cta)
{
    (void)n;
    int sms = 0;
    cudaError_t err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err != cudaSuccess) return err;
    int smem_max = 0;
    err = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (err != cudaSuccess) return err;
    const size_t smem = (size_t)(SLOTS * NB + NBA) * va_glv_t8s_block_doubles(stepper, n_out) * 8;
    if (smem + sizeof(SharedState) + 1024 > (size_t)smem_max) return cudaErrorInvalidConfiguration;
    *ctas_per_sm = 1; // the register split (setmaxnreg) is sized for one 384-thread CTA per SM
    *grid = sms;
    *threads = NT;
    *slots_per_cta = SLOTS;
    return cudaSuccess;
}

cudaError_t va_glv_t8s_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    switch (a.stepper) {
    case VA_RK_RK4: return launch<TabRK4, false>(a, st);
    case VA_RK_CK54: return launch<TabCK54, true>(a, st);
    case VA_RK_DOPRI5: return launch<TabDOPRI5, true>(a, st);
    }
    return cudaErrorInvalidValue;
}
