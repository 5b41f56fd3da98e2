// va_glv_wide.cu -- CTA-per-trajectory forward + discrete-adjoint kernel for the Generalized Lotka-Volterra system
//                    with up to 64 species (the headline path: GLV N = 64, one million parameter sets).
//
// f_i = x_i (r_i + (A x)_i), parameters [r, A row-major] (reference examples/GeneralizedLotkaVolterra/main.cpp:105-119).
// The reference evaluates f and its vector-Jacobian products through AADC-recorded AVX kernels (lib/include/AadData.hpp
// :175-210, :291-330), feeding all N^2+N parameters into the workspace for every call. Here the interaction matrix of a
// trajectory is loaded into the register file of a 256-thread CTA and stays there for the whole sweep.
//
// Thread <-> data map. The 64x64 matrix is cut into 16x16 tiles of 4x4 (16 FP64 registers per thread). One tile
// coordinate is the lane inside a 16-lane group (g), the other the group index (h). Along g a tile takes the entries
// F(g,k) = 2g + (k&1) + 32(k>>1), so that the 16 lanes of a group read their vector operands as two conflict-free
// LDS.128; along h it takes G(h,k) = 4h + k. With A in registers the scarce resource is the shared-memory / shuffle data path
// (128 B/clk/SM, a quarter of what 64 DFMA/clk would need if every DFMA pulled a fresh 8-byte operand), so the map is
// chosen to minimise operand traffic: a 4x4 tile needs 4 vector entries for 16 DFMA, and the partial sums are combined
// by a recursive-halving exchange inside 16-lane groups (5 double shuffles, no shared memory):
//   forward  (reference lib/include/detail/runge_kutta.hpp:76-118 + odeint controlled stepper):
//       p = tid/16, q = tid%16; y = A x: 16 DFMA, reduce over q inside the warp.
//   backward (reference lib/include/detail/backpropagation.hpp:83-158, 231-254):
//       p = tid%16, q = tid/16 (transposed assignment, A re-read from L2); A^T v: 16 DFMA, reduce over p inside the
//       warp; the rank-1 gradient update Abar += v x^T (16 DFMA) reuses the same operands; Abar lives in 16 registers.
//   After either reduction lane g of a 16-lane group holds the entry 4*(tid/16) + g/4, so the SAME four lanes own a
//   vector component in both sweeps and carry its scalar recurrences (stage state, slopes, stage adjoints).
//
// A persistent CTA integrates trajectory after trajectory (static stride over the batch), forward then backward, and
// keeps its checkpoints in a private slab that is reused for every trajectory and therefore stays in L2.
// Checkpoint policy: STORE_STAGES -- for every accepted step the stage states X_m and g_m = r + A X_m are kept, so the
// reverse sweep needs no stage recompute (the reference recomputes them with s extra RHS calls per step,
// detail/backpropagation.hpp:24-64). The reverse sweep streams one step block (6 KB) at a time from the slab into
// shared memory with TMA bulk copies (cp.async.bulk + mbarrier), double buffered, one step ahead.
//
// Accept/reject logic is odeint's (same error norm, same step-size rules). The matrix-vector products use FMA and a
// tree reduction, so stage values differ from the scalar reference by round-off; accepted-step counts can therefore
// flip only when the error estimate is within ~1e-8 relative of a threshold (documented in DESIGN.md).
#include <cstdlib>

#include "va_glv_common.cuh"

#ifndef VA_GLV_LG
#define VA_GLV_LG 8
#endif
#ifndef VA_GLV_PAIR
#define VA_GLV_PAIR 1 // trajectories integrated forward together per slot
#endif
#ifndef VA_GLV_MINB
#define VA_GLV_MINB 2 // resident CTAs per SM the register allocation is sized for
#endif

namespace {

#ifndef VA_GLV_MINB16
#define VA_GLV_MINB16 5 // resident 128-thread CTAs per SM the NP = 16 instantiation is sized for (build knob GLV_MINB16)
#endif
constexpr int NPMAX = 64; // largest padded species count of this kernel family
constexpr int HDR = 8;   // doubles in a step-block header (hdr[0] = t_n)

template <class Tab, int NP>
constexpr int block_doubles() { return HDR + 2 * Tab::SADJ * NP; }

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Combine the four partial sums of a tile over the LG lanes of a reduction group by recursive halving (2+1 double
// shuffles that halve the data, then log2(LG)-2 butterflies), no selects. Precondition: the tile is held PERMUTED --
// lane g keeps, in register k, the partial sum for tile entry R^k with R = g / (LG/4) (the permutation is applied once
// per trajectory when the tile is loaded). Lane g returns the complete sum for tile entry R; the LG/4 lanes that share
// R end up with bit-identical values (they take control decisions together).
template <int LG>
__device__ __forceinline__ double reduce_group(double s0, double s1, double s2, double s3)
{
    double k0 = s0 + shfl_xor_d(s2, LG / 2);
    const double k1 = s1 + shfl_xor_d(s3, LG / 2);
    if (LG == 8) {
        // the last halving step and the last butterfly in ONE exchange round: the three other lanes of the quad hold
        // this lane's entry in k0 (lane^1) or k1 (lanes ^2, ^3). Symmetric add tree -> lanes g and g^1 agree bit for bit.
        const double a1 = shfl_xor_d(k0, 1), a2 = shfl_xor_d(k1, 2), a3 = shfl_xor_d(k1, 3);
        return (k0 + a1) + (a2 + a3);
    }
    k0 += shfl_xor_d(k1, LG / 4);
#pragma unroll
    for (int d = LG / 8; d >= 1; d >>= 1) k0 += shfl_xor_d(k0, d);
    return k0;
}

// conditional swap used to permute a freshly loaded tile (once per trajectory and sweep)
__device__ __forceinline__ void cswap(bool c, double &u, double &v)
{
    const double t = c ? v : u;
    v = c ? u : v;
    u = t;
}

// ---- mbarrier + TMA bulk copy (global -> shared) ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// NP = padded species count (16, 32 or 64); LG = lanes per reduction group. A trajectory is integrated by NTT = (NP/4) LG
// threads (NP = 64, LG = 8: 128 threads with 4x8 / 8x4 tiles; NP = 16, LG = 8: ONE WARP with 4x2 / 2x4 tiles), and a CTA
// hosts TPC = 128 / NTT independent trajectories ("slots"), each with its own shared buffers, barrier and slab.
template <class Tab, bool ADAPTIVE, bool EXACT, int LG, int NP>
__global__ void __launch_bounds__(128, NP == 64 ? VA_GLV_MINB : NP == 32 ? 3 : VA_GLV_MINB16) k_glv_wide(const __grid_constant__ VaGlvWideArgs a)
{
    constexpr int NTT = (NP / 4) * LG; // threads per trajectory
    constexpr int NT = 128;            // threads per CTA
    constexpr int TPC = NT / NTT;      // trajectories (slots) per CTA
    constexpr int TG = NP / LG;        // tile extent along the lane (g) direction
    constexpr int NW = NTT / 32;       // warps per trajectory
    static_assert(NTT >= 32 && NT % NTT == 0 && TG >= 2 && TG % 2 == 0, "unsupported tile geometry");
    constexpr int S = Tab::S, SADJ = Tab::SADJ;
    constexpr int SE = Tab::FSAL ? S - 1 : S; // stages evaluated through an intermediate state
    constexpr int BLK = block_doubles<Tab, NP>();
    __shared__ __align__(16) double xs_all[TPC][VA_GLV_PAIR][2][NP]; // stage state (forward, per paired trajectory) / v = w o x (backward)
    __shared__ __align__(128) double xg_all[TPC][2][BLK]; // step blocks [hdr | X_0..X_{s-1} | g_0..g_{s-1}] streamed back by TMA
    __shared__ double red_all[TPC][VA_GLV_PAIR * (NW > 0 ? NW : 1)];
    __shared__ __align__(8) uint64_t mbar_all[TPC][2];
    __shared__ int red_i_all[TPC][NW > 0 ? NW : 1];

    const int slot = threadIdx.x / NTT;
    const int tid = threadIdx.x % NTT; // thread index inside the trajectory group
    double(*xs)[2][NP] = xs_all[slot];
    double(*xg)[BLK] = xg_all[slot];
    double *red = red_all[slot];
    uint64_t *mbar = mbar_all[slot];
    const int g = tid & (LG - 1);    // lane inside the reduction group (column tile forward, row tile backward)
    const int hi = tid / LG;         // the other tile coordinate, 0..NP/4-1 (row tile forward, column tile backward)
    const int R = g / (LG / 4);      // tile entry (of 4) this lane ends up with after a reduction
    const int own = 4 * hi + R;      // vector component owned by this lane (LG/4 redundant lanes per component)
    const bool writer = (g & (LG / 4 - 1)) == 0;
    const bool sw_lo = g & (LG / 4), sw_hi = g & (LG / 2); // permutation bits of R
    const int lane = tid & 31, warp = tid >> 5;
    const int n = a.n;
    const int npar = n * n + n;
    const int cap = a.cap;
    const int64_t gslot = (int64_t)blockIdx.x * TPC + slot; // global slot: owns one slab and one partial-sum row

    // barrier over the threads of ONE trajectory: a warp sync when a trajectory is a single warp, else a named barrier
    auto traj_sync = [&]() {
        if (NTT == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(NTT) : "memory");
    };
    // OR over the threads of one trajectory
    auto traj_or = [&](int v) -> int {
        v = __reduce_or_sync(0xffffffffu, v);
        if (NW > 1) {
            if (lane == 0) red_i_all[slot][warp] = v;
            traj_sync();
            v = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) v |= red_i_all[slot][w];
            traj_sync();
        }
        return v;
    };

    // g-direction entries of a tile: FG(e) = 2g + (e&1) + 2 LG (e>>1): the LG lanes of a group read their operands
    // as TG/2 conflict-free LDS.128 (double2 index g + LG j)
    auto FG = [&](int e) { return 2 * g + (e & 1) + 2 * LG * (e >> 1); };

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t mbar_parity = 0; // bit b: parity of the next completion of mbar[b]

    // gradient accumulators (backward tile: TG rows x 4 columns); persistent across trajectories in summed mode
    double Abar[TG][4], rbar = 0.0;
#pragma unroll
    for (int r = 0; r < TG; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) Abar[r][c] = 0.0;

    // NQ trajectories are integrated FORWARD together, interleaved in the same warps: their matrix tiles sit side by side
    // in the register file (the gradient accumulator is not live yet), every stage does NQ independent matrix-vector
    // products behind ONE barrier, and the compiler overlaps the shuffle/LDS latency of one trajectory with the DFMAs of
    // the other. The reverse sweeps then run one after the other (each needs A and Abar: 128 registers).
    constexpr int NQ = VA_GLV_PAIR;
    const int64_t wave = (int64_t)gridDim.x * TPC; // trajectories taken per wave of slots
    for (int64_t b0 = gslot; b0 < a.B; b0 += wave * NQ) {
        int64_t bq[NQ];
        bool ex[NQ];
        double Af[NQ][4][TG];
        double r_own[NQ], x[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            bq[q] = b0 + q * wave;
            ex[q] = bq[q] < a.B;
            r_own[q] = 0.0;
            x[q] = 0.0;
            const double *pb = a.params + (ex[q] ? bq[q] : b0) * npar;
            // forward tile: 4 rows 4hi + k (held permuted: register k <- row R^k), TG columns FG(c)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int row = 4 * hi + r;
                if (EXACT) {
                    const double2 *src = reinterpret_cast<const double2 *>(pb + NP + row * NP + 2 * g);
#pragma unroll
                    for (int j = 0; j < TG / 2; ++j) {
                        const double2 v = __ldg(src + LG * j);
                        Af[q][r][2 * j] = v.x;
                        Af[q][r][2 * j + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < TG; ++c) {
                        const int col = FG(c);
                        {
                            // padded entries load a valid address and are discarded (a guarded load may become load + select)
                            const bool in = row < n && col < n;
                            const double v = __ldg(pb + n + (in ? row * n + col : 0));
                            Af[q][r][c] = in ? v : 0.0;
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < TG; ++c) {
                cswap(sw_lo, Af[q][0][c], Af[q][1][c]);
                cswap(sw_lo, Af[q][2][c], Af[q][3][c]);
                cswap(sw_hi, Af[q][0][c], Af[q][2][c]);
                cswap(sw_hi, Af[q][1][c], Af[q][3][c]);
            }
            {
                const int oi = own < n ? own : 0; // padded lanes read a valid address and discard it
                const double rv = __ldg(pb + oi), xv0 = __ldg(a.x0 + (ex[q] ? bq[q] : b0) * n + oi);
                if (ex[q] && own < n) { r_own[q] = rv; x[q] = xv0; }
            }
        }

        // sum[q] = (A_q X_q)_own for the stage states whose own-components are X[q]. `extra` runs between the operand
        // loads and the reductions: work that does not depend on the results is issued there, off the critical path.
        auto matvec = [&](const double(&X)[NQ], int m, double(&sum)[NQ], auto &&extra) { // m: compile-time after unrolling
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (writer) xs[q][m & 1][own] = X[q];
            traj_sync();
            double s[NQ][4];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const double2 *xv = reinterpret_cast<const double2 *>(xs[q][m & 1]) + g;
                double xc[TG];
#pragma unroll
                for (int j = 0; j < TG / 2; ++j) {
                    const double2 v = xv[LG * j];
                    xc[2 * j] = v.x;
                    xc[2 * j + 1] = v.y;
                }
                if (NQ == 1) {
                    // a single trajectory: two accumulator chains per tile row keep the FP64 pipe busy from one warp
                    double u[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) { s[q][r] = Af[q][r][0] * xc[0]; u[r] = Af[q][r][TG / 2] * xc[TG / 2]; }
#pragma unroll
                    for (int c = 1; c < TG / 2; ++c)
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            s[q][r] = fma(Af[q][r][c], xc[c], s[q][r]);
                            u[r] = fma(Af[q][r][TG / 2 + c], xc[TG / 2 + c], u[r]);
                        }
#pragma unroll
                    for (int r = 0; r < 4; ++r) s[q][r] += u[r];
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) s[q][r] = Af[q][r][0] * xc[0];
#pragma unroll
                    for (int c = 1; c < TG; ++c)
#pragma unroll
                        for (int r = 0; r < 4; ++r) s[q][r] = fma(Af[q][r][c], xc[c], s[q][r]);
                }
            }
            extra();
#pragma unroll
            for (int q = 0; q < NQ; ++q) sum[q] = reduce_group<LG>(s[q][0], s[q][1], s[q][2], s[q][3]);
        };

        const double tf = a.tf;
        double t[NQ], dt[NQ], K[NQ][S], g0[NQ];
        int nck[NQ], rejects[NQ], status[NQ], trials[NQ];
        bool act[NQ], fresh[NQ];
        double *sp[NQ]; // this lane's column in the current step block of trajectory q (advanced on acceptance)
        {
            double sum[NQ];
            matvec(x, 0, sum, [] {});
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                g0[q] = r_own[q] + sum[q];
                K[q][0] = x[q] * g0[q];
                t[q] = a.ti;
                dt[q] = a.dt0;
                nck[q] = rejects[q] = status[q] = trials[q] = 0;
                fresh[q] = true;
                act[q] = ex[q] && (ADAPTIVE ? va_less_with_sign(t[q], tf, dt[q]) : va_less_eq_with_sign(t[q] + dt[q], tf, dt[q]));
                sp[q] = a.slab + (gslot * NQ + q) * a.slab_stride + HDR + own;
            }
        }
        auto store_stage = [&](int q, int m, double X, double gg) {
            if (writer && act[q]) {
                sp[q][m * NP] = X;
                sp[q][(SADJ + m) * NP] = gg;
            }
        };
        auto any_active = [&]() {
            bool r = false;
#pragma unroll
            for (int q = 0; q < NQ; ++q) r = r || act[q];
            return r;
        };

        while (any_active()) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (act[q] && fresh[q]) {
                    if (nck[q] >= cap) {
                        status[q] |= VA_TRAJ_CKPT_OVERFLOW;
                        act[q] = false;
                    } else {
                        store_stage(q, 0, x[q], g0[q]);
                        if (tid == 0) sp[q][-HDR] = t[q]; // own == 0 for thread 0: header of the current block
                        if (ADAPTIVE && va_less_with_sign(tf, t[q] + dt[q], dt[q])) dt[q] = tf - t[q];
                        trials[q] = 0;
                        fresh[q] = false;
                    }
                }
            }
            if (!any_active()) break;
            // Stage m produces K_m = X_m (r + A X_m). The state of the NEXT stage (or the new solution after the last
            // one), Y = x + dt sum_{j<=m} c_j K_j, is split so that only ONE DFMA follows the reduction:
            //   Y = fma(c1, sum, base),  c1 = dt c_m X_m,  base = x + dt sum_{j<m} c_j K_j + c1 r   (all known early).
            double X[NQ], perr[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                X[q] = fma(dt[q] * a.coef.a[1][0], K[q][0], x[q]);
                perr[q] = 0.0; // sum_{j<SE-1} db_j K_j
            }
#pragma unroll
            for (int m = 1; m < SE; ++m) {
                const bool last = (m == SE - 1);
                double c1[NQ], base[NQ], sum[NQ];
                matvec(X, m, sum, [&] {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < m; ++j) {
                            const double cz = last ? Tab::b(j) : Tab::a(m + 1, j);
                            if (cz != 0.0) acc = fma(last ? a.coef.b[j] : a.coef.a[m + 1][j], K[q][j], acc);
                        }
                        const double cm = last ? Tab::b(m) : Tab::a(m + 1, m);
                        c1[q] = (cm != 0.0) ? (dt[q] * (last ? a.coef.b[m] : a.coef.a[m + 1][m])) * X[q] : 0.0;
                        base[q] = fma(c1[q], r_own[q], fma(dt[q], acc, x[q]));
                        if (last && ADAPTIVE) {
#pragma unroll
                            for (int j = 0; j < m; ++j)
                                if (Tab::db(j) != 0.0) perr[q] = fma(a.coef.db[j], K[q][j], perr[q]);
                        }
                    }
                });
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const double Y = fma(c1[q], sum[q], base[q]);
                    const double gg = r_own[q] + sum[q];
                    K[q][m] = X[q] * gg;
                    if (m < SADJ) store_stage(q, m, X[q], gg);
                    X[q] = Y;
                }
            }
            // X = new solution. f(xnew) is evaluated now for every trajectory: it is the FSAL stage of dopri5, and for the
            // other steppers the first slope of the next step (speculative: discarded if the step is rejected).
            double gl[NQ], Kl[NQ];
            {
                double sum[NQ];
                matvec(X, SE & 1 ? 1 : 0, sum, [] {});
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    gl[q] = r_own[q] + sum[q];
                    Kl[q] = X[q] * gl[q];
                    if (Tab::FSAL) K[q][S - 1] = Kl[q];
                }
            }
            double err[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) err[q] = 0.0;
            if (ADAPTIVE) {
                double e[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    double acc = perr[q];
#pragma unroll
                    for (int j = SE - 1; j < S; ++j)
                        if (Tab::db(j) != 0.0) acc = fma(a.coef.db[j], K[q][j], acc);
                    const double xerr = dt[q] * acc;
                    // default_error_checker::error, max norm over species
                    e[q] = fabs(xerr) / (a.eps_abs + a.eps_rel * (fabs(x[q]) + fabs(dt[q]) * fabs(K[q][0])));
#pragma unroll
                    for (int d = 16; d >= LG / 4; d >>= 1) e[q] = fmax(e[q], shfl_xor_d(e[q], d));
                }
                if (NW > 1) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                        if (lane == 0) red[q * NW + warp] = e[q];
                    traj_sync();
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
#pragma unroll
                        for (int w = 0; w < NW; ++w) err[q] = fmax(err[q], red[q * NW + w]);
                } else {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) err[q] = e[q];
                }
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (!act[q]) continue;
                const bool accept = !ADAPTIVE || !(err[q] > 1.0);
                if (!accept) {
                    // default_step_adjuster::decrease_step
                    dt[q] *= fmax(0.9 * inv_root<(Tab::ERROR_ORDER > 1 ? Tab::ERROR_ORDER - 1 : 1)>(err[q]), 0.2);
                    ++rejects[q];
                    if (++trials[q] >= 500) { status[q] |= VA_TRAJ_NO_PROGRESS; act[q] = false; }
                } else {
                    x[q] = X[q];
                    ++nck[q];
                    sp[q] += BLK;
                    if (ADAPTIVE) {
                        t[q] += dt[q];
                        // default_step_adjuster::increase_step
                        if (err[q] < 0.5) {
                            constexpr int P = Tab::STEPPER_ORDER;
                            double floor_ = 1.0;
#pragma unroll
                            for (int k = 0; k < P; ++k) floor_ *= 0.2; // 5^-P
                            // err <= 5^-P: the growth factor is exactly 0.9 * 5 (pow(5^-P, -1/P) == 5 in glibc as well)
                            dt[q] *= (err[q] <= floor_) ? 4.5 : 9.0 / 10.0 * inv_root<P>(err[q]);
                        }
                        act[q] = va_less_with_sign(t[q], tf, dt[q]);
                    } else {
                        t[q] = a.ti + (double)nck[q] * dt[q]; // detail/runge_kutta.hpp:64
                        act[q] = va_less_eq_with_sign(t[q] + dt[q], tf, dt[q]);
                    }
                    fresh[q] = true;
                    g0[q] = gl[q];
                    K[q][0] = Kl[q];
                }
            }
        }
        // close the trajectories: final time, status, x(tf)
        int Tq[NQ];
        bool failedq[NQ];
        double x_tfq[NQ], t_finalq[NQ];
        int st_all = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            Tq[q] = nck[q];
            if (tid == 0 && ex[q]) sp[q][-HDR] = t[q]; // header of block T carries the final time
            if (!isfinite(x[q])) status[q] |= VA_TRAJ_NONFINITE;
            st_all |= status[q] << (8 * q);
        }
        fence_proxy_async(); // generic-proxy slab writes -> visible to the TMA reads of the reverse sweep
        traj_sync();
        st_all = traj_or(st_all);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            status[q] = (st_all >> (8 * q)) & 0xff;
            failedq[q] = status[q] & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS);
            x_tfq[q] = x[q];
            t_finalq[q] = t[q];
            if (ex[q]) {
                if (writer && own < n) a.x_final[bq[q] * n + own] = failedq[q] ? nan("") : x[q];
                if (tid == 0) {
                    if (a.n_accept) a.n_accept[bq[q]] = Tq[q];
                    if (a.n_reject) a.n_reject[bq[q]] = rejects[q];
                    if (a.status) a.status[bq[q]] = status[q];
                }
            }
        }

#pragma unroll 1
        for (int q = 0; q < NQ; ++q) {
        if (!ex[q]) continue;
        const int64_t b = bq[q];
        const double *pb = a.params + b * npar;
        double *const slab = a.slab + (gslot * NQ + q) * a.slab_stride;
        const int T = Tq[q];
        const bool failed = failedq[q];
        const double x_tf = x_tfq[q], t_final = t_finalq[q];
        // ================================ reverse sweep =====================================
        // transposed tile: TG rows FG(r), 4 columns 4hi + k (held permuted: register k <- column R^k); the owned
        // component stays `own`. A is re-read (L2 hit).
        double Ab[TG][4];
        if (a.n_out > 0) {
#pragma unroll
            for (int r = 0; r < TG; ++r) {
                const int row = FG(r);
                if (EXACT) {
                    const double2 *src = reinterpret_cast<const double2 *>(pb + NP + row * NP + 4 * hi);
                    const double2 v0 = __ldg(src), v1 = __ldg(src + 1);
                    Ab[r][0] = v0.x; Ab[r][1] = v0.y; Ab[r][2] = v1.x; Ab[r][3] = v1.y;
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int col = 4 * hi + c;
                        {
                            const bool in = row < n && col < n;
                            const double v = __ldg(pb + n + (in ? row * n + col : 0));
                            Ab[r][c] = in ? v : 0.0;
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < TG; ++r) {
                cswap(sw_lo, Ab[r][0], Ab[r][1]);
                cswap(sw_lo, Ab[r][2], Ab[r][3]);
                cswap(sw_hi, Ab[r][0], Ab[r][2]);
                cswap(sw_hi, Ab[r][1], Ab[r][3]);
            }
        }

        for (int o = 0; o < a.n_out; ++o) {
            double *lam_io = a.lambda + (b * a.n_out + o) * n;
            double *mu_o = a.mu + (a.reduce == VA_REDUCE_SUM ? (int64_t)o : (b * a.n_out + o)) * npar;
            if (failed) {
                if (writer && own < n) lam_io[own] = nan("");
                if (a.reduce == VA_REDUCE_NONE)
                    for (int k = tid; k < npar; k += NTT) mu_o[k] = nan("");
                continue;
            }
            double lam;
            if (a.objective == VA_OBJ_SUM) lam = (own < n) ? 1.0 : 0.0;
            else if (a.objective == VA_OBJ_HALF_NORM2) lam = x_tf;
            else lam = (own < n) ? lam_io[own < n ? own : 0] : 0.0;

            if (a.reduce == VA_REDUCE_NONE) {
#pragma unroll
                for (int r = 0; r < TG; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) Abar[r][c] = 0.0;
                rbar = 0.0;
            }

            // stream the step blocks back, newest first; block `step` -> buffer (T-1-step)&1, fetched one step ahead
            if (o > 0) traj_sync(); // the previous seed's last reads of xg[] are done before it is refilled
            if (T > 0 && tid == 0) {
                mbar_expect_tx(&mbar[0], BLK * 8);
                bulk_g2s(xg[0], slab + (int64_t)(T - 1) * BLK, BLK * 8, &mbar[0]);
            }
            double t_hi = t_final;
            for (int step = T - 1; step >= 0; --step) {
                const int bufi = (T - 1 - step) & 1;
                mbar_wait(&mbar[bufi], (mbar_parity >> bufi) & 1);
                mbar_parity ^= 1u << bufi;
                const double *blk = xg[bufi];
                const double t_lo = blk[0];
                const double dt_s = t_hi - t_lo; // StateStorage::GetDt: difference of the stored times
                t_hi = t_lo;
                double W[SADJ + 1];
                W[0] = lam;
#pragma unroll
                for (int m = 1; m <= SADJ; ++m) W[m] = Tab::b(m - 1) != 0.0 ? (a.coef.b[m - 1] * dt_s) * lam : 0.0;
                // v = w_m o X_{m-1} for the stage about to be processed; later stages get it from the previous one
                double v = W[SADJ] * blk[HDR + (SADJ - 1) * NP + own];
#pragma unroll
                for (int m = SADJ; m >= 1; --m) {
                    double *vb = xs[0][m & 1];
                    if (writer) vb[own] = v;
                    traj_sync();
                    if (m == SADJ && step > 0 && tid == 0) {
                        // every thread is past its reads of the other buffer (previous step): refill it
                        mbar_expect_tx(&mbar[bufi ^ 1], BLK * 8);
                        bulk_g2s(xg[bufi ^ 1], slab + (int64_t)(step - 1) * BLK, BLK * 8, &mbar[bufi ^ 1]);
                    }
                    const double2 *vv2 = reinterpret_cast<const double2 *>(vb) + g;
                    const double2 *xx2 = reinterpret_cast<const double2 *>(blk + HDR + (m - 1) * NP + 4 * hi);
                    double vr[TG];
#pragma unroll
                    for (int j = 0; j < TG / 2; ++j) {
                        const double2 vq = vv2[LG * j];
                        vr[2 * j] = vq.x;
                        vr[2 * j + 1] = vq.y;
                    }
                    const double2 x01 = xx2[0], x23 = xx2[1];
                    const double xc[4] = {x01.x, x01.y, x23.x, x23.y};
                    double s[4], u[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) { s[c] = Ab[0][c] * vr[0]; u[c] = Ab[TG / 2][c] * vr[TG / 2]; }
#pragma unroll
                    for (int r = 1; r < TG / 2; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) { s[c] = fma(Ab[r][c], vr[r], s[c]); u[c] = fma(Ab[TG / 2 + r][c], vr[TG / 2 + r], u[c]); }
#pragma unroll
                    for (int c = 0; c < 4; ++c) s[c] += u[c];
                    // gx = (A^T v)_own + w_m g_{m-1}. The next stage's v = (w_{m-1} + gx a dt) X_{m-2} is arranged as
                    // fma(sum, c1, c2) with c1, c2 known before the reduction returns: one DFMA on the critical path.
                    const double wg = W[m] * blk[HDR + (SADJ + m - 1) * NP + own];
                    double c1 = 0.0, c2 = 0.0;
                    if (m > 1) {
                        const double Xn = blk[HDR + (m - 2) * NP + own];
                        if (Tab::a(m - 1, m - 2) != 0.0) c1 = (a.coef.a[m - 1][m - 2] * dt_s) * Xn;
                        c2 = fma(wg, c1, W[m - 1] * Xn);
                    }
                    // first exchange round of the reduction is issued BEFORE the rank-1 gradient update: the 32 independent
                    // Abar DFMAs (not on the critical path) then execute while the shuffles are in flight
                    double k0 = s[0], k1 = s[1];
                    const double e2 = shfl_xor_d(s[2], LG / 2), e3 = shfl_xor_d(s[3], LG / 2);
#pragma unroll
                    for (int r = 0; r < TG; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) Abar[r][c] = fma(vr[r], xc[c], Abar[r][c]);
                    k0 += e2;
                    k1 += e3;
                    double sum;
                    if (LG == 8) {
                        const double a1 = shfl_xor_d(k0, 1), a2 = shfl_xor_d(k1, 2), a3 = shfl_xor_d(k1, 3);
                        sum = (k0 + a1) + (a2 + a3);
                    } else {
                        k0 += shfl_xor_d(k1, LG / 4);
#pragma unroll
                        for (int d = LG / 8; d >= 1; d >>= 1) k0 += shfl_xor_d(k0, d);
                        sum = k0;
                    }
                    const double v_next = fma(sum, c1, c2);
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
            if (writer && own < n) lam_io[own] = lam;
            if (a.reduce == VA_REDUCE_NONE) {
                if (writer && own < n) mu_o[own] = rbar;
#pragma unroll
                for (int r = 0; r < TG; ++r) {
                    const int row = FG(r);
                    if (EXACT) {
                        double2 *dst = reinterpret_cast<double2 *>(mu_o + NP + row * NP + 4 * hi);
                        dst[0] = make_double2(Abar[r][0], Abar[r][1]);
                        dst[1] = make_double2(Abar[r][2], Abar[r][3]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int col = 4 * hi + c;
                            if (row < n && col < n) mu_o[n + row * n + col] = Abar[r][c];
                        }
                    }
                }
            }
        }
        traj_sync(); // slab and shared buffers are reused by the next trajectory
        } // q: reverse sweeps, one trajectory after the other
    }

    if (a.reduce == VA_REDUCE_SUM) {
        double *part = a.partial + gslot * npar;
        if (writer && own < n) part[own] = rbar;
#pragma unroll
        for (int r = 0; r < TG; ++r) {
            const int row = FG(r);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = 4 * hi + c;
                if (row < n && col < n) part[n + row * n + col] = Abar[r][c];
            }
        }
    }
}

constexpr int kLG = VA_GLV_LG; // lanes per reduction group, chosen at build time (csrc/Makefile)

int np_of(int n) { return n <= 16 ? 16 : n <= 32 ? 32 : 64; }

template <class Tab, bool ADAPTIVE, int NP>
cudaError_t launch_np(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.n == NP) k_glv_wide<Tab, ADAPTIVE, true, kLG, NP><<<a.grid, 128, 0, st>>>(a);
    else k_glv_wide<Tab, ADAPTIVE, false, kLG, NP><<<a.grid, 128, 0, st>>>(a);
    return cudaGetLastError();
}

template <class Tab, bool ADAPTIVE>
cudaError_t launch(const VaGlvWideArgs &a_in, cudaStream_t st)
{
    VaGlvWideArgs a = a_in;
    // tableau values travel in the kernel arguments (constant bank): DFMA takes them as c[bank][offset] operands,
    // while the compile-time copy in Tab:: only decides which terms exist
    for (int m = 0; m < Tab::S; ++m) {
        for (int j = 0; j < m && j < 6; ++j) a.coef.a[m][j] = Tab::a(m, j);
        a.coef.b[m] = Tab::b(m);
        a.coef.db[m] = Tab::db(m);
    }
    switch (np_of(a.n)) {
    case 16: return launch_np<Tab, ADAPTIVE, 16>(a, st);
    case 32: return launch_np<Tab, ADAPTIVE, 32>(a, st);
    default: return launch_np<Tab, ADAPTIVE, 64>(a, st);
    }
}

template <class Tab, bool ADAPTIVE, int NP>
cudaError_t occupancy_np(int n, int *ctas_per_sm)
{
    if (n == NP) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_glv_wide<Tab, ADAPTIVE, true, kLG, NP>, 128, 0);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_glv_wide<Tab, ADAPTIVE, false, kLG, NP>, 128, 0);
}

template <class Tab, bool ADAPTIVE>
cudaError_t occupancy(int n, int *ctas_per_sm)
{
    switch (np_of(n)) {
    case 16: return occupancy_np<Tab, ADAPTIVE, 16>(n, ctas_per_sm);
    case 32: return occupancy_np<Tab, ADAPTIVE, 32>(n, ctas_per_sm);
    default: return occupancy_np<Tab, ADAPTIVE, 64>(n, ctas_per_sm);
    }
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

bool va_glv_wide_supported(int n, int stepper, int adaptive)
{
    if (n < 1 || n > NPMAX) return false;
    if (stepper == VA_RK_RK4) return !adaptive;
    if (stepper == VA_RK_CK54 || stepper == VA_RK_DOPRI5) return adaptive != 0;
    return false;
}

int va_glv_wide_padded(int n) { return np_of(n); }

int va_glv_wide_pair() { return VA_GLV_PAIR; }

int va_glv_wide_block_doubles(int n, int stepper) { return HDR + 2 * sadj_of(stepper) * np_of(n); }

int64_t va_glv_wide_slab_doubles(int n, int stepper, int cap)
{
    return (int64_t)(cap + 1) * va_glv_wide_block_doubles(n, stepper); // block T holds only the final time in its header
}

cudaError_t va_glv_wide_config(int n, int stepper, int device, int *grid, int *ctas_per_sm, int *threads, int *slots_per_cta)
{
    int sms = 0;
    cudaError_t err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err != cudaSuccess) return err;
    int occ = 0;
    switch (stepper) {
    case VA_RK_RK4: err = occupancy<TabRK4, false>(n, &occ); break;
    case VA_RK_CK54: err = occupancy<TabCK54, true>(n, &occ); break;
    case VA_RK_DOPRI5: err = occupancy<TabDOPRI5, true>(n, &occ); break;
    default: return cudaErrorInvalidValue;
    }
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    if (const char *env = getenv("VA_GLV_CTAS_PER_SM")) { // experiment knob: fewer resident CTAs than the occupancy limit
        const int v = atoi(env);
        if (v >= 1 && v < occ) occ = v;
    }
    *ctas_per_sm = occ;
    *grid = sms * occ;
    *threads = 128;
    *slots_per_cta = 128 / ((np_of(n) / 4) * kLG);
    return cudaSuccess;
}

cudaError_t va_glv_wide_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    switch (a.stepper) {
    case VA_RK_RK4: return launch<TabRK4, false>(a, st);
    case VA_RK_CK54: return launch<TabCK54, true>(a, st);
    case VA_RK_DOPRI5: return launch<TabDOPRI5, true>(a, st);
    }
    return cudaErrorInvalidValue;
}
