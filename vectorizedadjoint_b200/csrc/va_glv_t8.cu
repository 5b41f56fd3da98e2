// va_glv_t8.cu -- GLV forward + discrete-adjoint kernel for 33..64 species, second generation of the headline path
//                  (GLV N = 64, one million parameter sets): 64 threads per trajectory, 64-entry register tiles, three phases.
//
// f_i = x_i (r_i + (A x)_i), parameters [r, A row-major] (reference examples/GeneralizedLotkaVolterra/main.cpp:105-119).
// Same algorithm as va_glv_wide.cu (reference lib/include/detail/runge_kutta.hpp:76-118 forward + odeint controlled
// stepper; detail/backpropagation.hpp:83-158, 231-254 reverse); what changed is the thread <-> data map, chosen from
// the ncu profile of the first generation (profiles/r01_glv64_ncu_full_ve.txt: FP64 pipe 48 % busy of which only 65 % were
// the algorithmic DFMAs, 8 warps per SM, every stage a barrier -> loads -> DFMA -> exchange chain):
//   * 64 threads = one trajectory (four trajectories per 256-thread CTA, one CTA per SM). The matrix is cut into tiles of
//     64 entries per thread (128 registers): 4 rows x 16 columns by default (VA_T8_RT). 16 loaded vector operands feed
//     64 DFMA, the four partial sums of a row are combined over 4 lanes in ONE round of 3 independent double shuffles
//     (select-free: the tile is held permuted), after which EVERY thread owns exactly one vector component: the stage
//     recurrences are no longer computed redundantly (they were 2x redundant). Measured: 8x8 tiles / 8 lanes / 7
//     shuffles 5.65 M gradients/s, 4x16 / 4 lanes / 3 shuffles 5.94 M, 2x32 / 2 lanes 5.75 M.
//   * the gradient accumulator Abar (another 128 registers) cannot be live next to the matrix, so the reverse sweep is
//     split: phase 2 propagates lambda through the stored steps with A^T in registers and writes the seeds
//     v_m = w_m o X_m to the step block in the slab; phase 3 streams the blocks once more and accumulates
//     Abar += v_m X_m^T -- 64 independent DFMA chains per thread, no reductions, no barriers inside a step: pure FP64
//     throughput that fills the pipe while other trajectories of the SM sit in their latency-bound phases.
//   * <= 255 registers per thread, 8 warps per SM. (The register file is partitioned per SM sub-partition, so warps per
//     SM come in multiples of four: the next step, 12 warps, leaves 168 registers, serialises the operand loads and
//     measured slower, 5.04 M.)
// Step blocks: [8-double header (t_n) | X_0..X_{s-1} | g_0..g_{s-1} | v_1..v_s] in a private slab per CTA (reused for
// every trajectory -> L2 resident), streamed back by TMA bulk copies (cp.async.bulk + mbarrier), NB buffers deep. With a
// single seed per trajectory the v section aliases the g section (g_m is dead once v_m exists).
//
// The accumulation order of every sum follows the reference (newest step first, stages s..1); matrix-vector products
// use FMA and tree reductions, so values differ from the scalar reference by round-off (parity section).
#include <cstdlib>

#include "va_glv_common.cuh"
#include "va_tma.cuh"

#ifndef VA_T8_MINB
#define VA_T8_MINB 4 // resident CTAs per SM the register allocation is sized for
#endif
#ifndef VA_T8_SLOTS
#define VA_T8_SLOTS 4 // trajectories ("slots") per CTA: 4 -> one 256-thread CTA per SM; 1 -> 64-thread CTAs, VA_T8_MINB per SM
#endif
#ifndef VA_T8_RT
#define VA_T8_RT 4 // rows of the matrix tile a thread holds in phases 1 and 2: 8 (8x8 tile, sums over 8 lanes) or 4 (4x16, 4 lanes)
#endif
#ifndef VA_T8_NC
#define VA_T8_NC 1 // independent accumulator chains per tile row in a product (column pairs are dealt round-robin)
#endif
#ifndef VA_T8_NB
#define VA_T8_NB 4 // step-block buffers per slot (2 / 3 / 4: 5.92 / 5.91 / 5.94 M gradients/s)
#endif
#ifndef VA_T8_MAP
#define VA_T8_MAP 0 // warp -> trajectory map inside a CTA (see the kernel)
#endif
#ifndef VA_T8_STAGE
#define VA_T8_STAGE 1 // one seed per trajectory: the parameter set [r, A] is brought into shared memory by one TMA bulk copy (see the kernel)
#endif
#ifndef VA_T8_P3
#define VA_T8_P3 0 // phase 3 (Abar += v x^T): 0 = DFMA on 8x8 register tiles, 1 = FP64 tensor instructions (DMMA m8n8k4), see the kernel
#endif

namespace {

constexpr int NP = 64;  // padded species count
constexpr int NTT = 64; // threads per trajectory
constexpr int SLOTS = VA_T8_SLOTS;
constexpr int NT = NTT * SLOTS; // threads per CTA
constexpr int MINB = SLOTS == 1 ? VA_T8_MINB : 1;

constexpr int HDR = 8;  // doubles in a step-block header (hdr[0] = t_n)
constexpr int NB = VA_T8_P3 ? 4 : VA_T8_NB;
constexpr int NB_STAGED = 3;          // step-block buffers per slot when the parameter sets are staged (shared memory: 4 x 3 x 6.2 KB + 4 x 32.5 KB)
constexpr int PA_DOUBLES = NP * NP + NP; // shared-memory copy of one parameter set
// DMMA variant of phase 3: a step's 2 x SADJ vectors sit in shared memory with a stride of 68 doubles (544 B = 32 B more than a
// multiple of 128), so that the fragment loads -- lane l reads element (l >> 2) of vector (l & 3): four vectors, 32 bytes each
// per half-warp -- touch every bank once; with the natural stride of 512 B they were 4-way bank conflicts
constexpr int VS3 = 68;
constexpr int RT = VA_T8_RT; // tile rows per thread = lanes per reduction group
constexpr int CT = 64 / RT; // tile columns per thread
constexpr int NC = VA_T8_NC;
static_assert((RT == 8 || RT == 4 || RT == 2) && (NC == 1 || NC == 2 || NC == 4), "tile geometry");

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// doubles per shared-memory step buffer: a step block, or (DMMA variant of phase 3) its 2 x SADJ vectors at the padded stride
// (+ 8: consecutive buffers are 64 B off the 128 B grid, so that a fragment group that spans two buffers stays conflict-free)
__host__ __device__ constexpr int buf_doubles(int blk, int sadj) { return VA_T8_P3 && 2 * sadj * VS3 + 8 > blk ? 2 * sadj * VS3 + 8 : blk; }
#if VA_T8_P3
// D(8x8) += A(8x4) B(4x8) in FP64 on the tensor path: lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2], D[l >> 2][2 (l & 3) + {0, 1}]
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

// mbarrier + TMA bulk copy + L2 eviction-policy helpers: va_tma.cuh (one copy for all kernels).
// L2 residency control. The slabs (592 x ~130 KB of step blocks in flight per GPU) are written, read twice and then
// overwritten; the parameters (35 GB per pass) stream through once. Without hints the stream evicts the slabs (ncu:
// 280 KB of DRAM traffic per trajectory against 35 KB of compulsory bytes), so slab accesses carry evict_last and the
// last read of the parameters evict_first.
using va_tma::smem_u32;
using va_tma::mbar_init;
using va_tma::mbar_expect_tx;
using va_tma::bulk_g2s;
using va_tma::policy_evict_last;
using va_tma::policy_evict_first;
using va_tma::mbar_wait;
using va_tma::fence_proxy_async;
#if defined(VA_T8_FENCE_ALL) && VA_T8_FENCE_ALL
__device__ __forceinline__ void fence_proxy_async_global() { va_tma::fence_proxy_async(); } // experiment: the all-spaces form
#else
using va_tma::fence_proxy_async_global;
#endif
using va_tma::ldg_hint;
__device__ __forceinline__ void st_hint(double *p, double v, uint64_t policy) { va_tma::st_hint_relaxed(p, v, policy); }
// A step block is dead once the gradient accumulation has copied it to shared memory, but its L2 lines are dirty: when the next
// trajectory of the slot does not overwrite them before they are evicted, each costs a DRAM write of data nobody will read.
// discard.L2 drops a line without the write-back. Measured (2^20 sets): DRAM traffic 184 -> 128 GB per launch, 5.95 -> 5.88 M
// gradients/s whether the discards sit inside the accumulation loop or behind it -- the kernel is bound by the FP64 pipe, not by
// DRAM, so the engine only asks for it under VA_T8_DISCARD=1.
__device__ __forceinline__ void discard_line(const double *p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }

// max over the warp of non-negative doubles (or NaN, which orders above everything): two 32-bit redux operations on
// the bit pattern instead of five double shuffles + DMNMX
__device__ __forceinline__ double warp_max_nonneg(double e)
{
    const unsigned hi = (unsigned)__double2hiint(e), lo = (unsigned)__double2loint(e);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}

// e^(-1/P), the controller's step-size root (odeint default_step_adjuster; pow() in the reference). Every thread evaluates
// it on the critical path between two steps, so the dependent chain is kept short: float seed (relative error ~1e-6), two
// Newton steps on y^-P = e (error -> ~3e-12 -> below one ulp), y^P by squaring. e is clamped so that e y^P cannot overflow.
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

// STAGE (one seed per trajectory; round 2, second session): the parameter set of a trajectory is copied to shared memory by ONE TMA
// bulk copy, issued as soon as the previous trajectory of the slot has taken its second (transposed) matrix tile, i.e. a whole
// reverse sweep ahead of its use. Both register tiles are then cut from shared memory. Before, each tile was gathered from
// global memory / L2 by per-lane loads that touch 32 (row-major tile) and 8 (transposed tile) different 128-byte lines per warp
// instruction -- 1500 L1 tag requests per warp and trajectory, a tenth of all load/store-unit work of this kernel, whose LSU is as
// busy as its FP64 pipe -- and the parameters had to survive in L2 from the forward sweep to the reverse sweep. Costs one of the
// four step-block buffers per slot (shared memory).
template <class Tab, bool ADAPTIVE, bool EXACT, bool STAGE>
__global__ void __launch_bounds__(NT, MINB) k_glv_t8(const __grid_constant__ VaGlvWideArgs a)
{
    constexpr int NB = STAGE ? NB_STAGED : ::NB; // step-block buffers per slot in THIS instantiation
    constexpr int S = Tab::S, SADJ = Tab::SADJ;
    constexpr int SE = Tab::FSAL ? S - 1 : S; // stages evaluated through an intermediate state
    extern __shared__ __align__(128) double xg_all[]; // per slot: NB step-block buffers of a.blk_doubles
    __shared__ __align__(16) double xs_all[SLOTS][2][NP]; // operand vector of the current matrix-vector product (double buffered)
    __shared__ double red_all[SLOTS][2];
    __shared__ int st_all[SLOTS][2];
    __shared__ __align__(8) uint64_t mbar_all[SLOTS][NB];
    __shared__ __align__(8) uint64_t pbar_all[SLOTS]; // STAGE: the slot's parameter set has landed

    // Warp w of the CTA runs on SM sub-partition w % 4 (tools/microbench/smsp_map.cu). Slot s takes the warp pair
    // {2 s, 2 s + 1}: the two warps of a trajectory sit on different sub-partitions, and every sub-partition hosts one warp of
    // two different trajectories. (Tried and dropped, three times: making those two warps take turns on the FP64 pipe --
    // with an atomic lock, with advisory flags, and with flags whose stores and loads are data-dependent on the burst so
    // that ptxas cannot move them: 4.82 M, 4.60 M and 4.53 M gradients/s against 5.7 M without. ncu shows bursts stretched
    // from ~155 to ~280 cycles by the neighbour, but serialising them costs more than the overlap.)
    const int wc = threadIdx.x >> 5;
#if VA_T8_MAP == 1
    // experiment: every warp of a trajectory shares its sub-partition with a DIFFERENT other trajectory
    // (warps 0,1 -> slot 0; 2,3 -> slot 1; 4,6 -> slot 2; 5,7 -> slot 3)
    const int slot = SLOTS == 1 ? 0 : (wc < 4 ? (wc >> 1) : 2 + (wc & 1));
    const int warp = wc < 4 ? (wc & 1) : ((wc >> 1) & 1); // warp inside the trajectory
#else
    const int slot = SLOTS == 1 ? 0 : ((wc >> 2) << 1) | ((wc >> 1) & 1);
    const int warp = wc & 1;                       // warp inside the trajectory
#endif
    const int tid = warp * 32 + (threadIdx.x & 31); // thread inside the trajectory
    const int g = tid & (RT - 1); // lane inside the reduction group (RT lanes)
    const int hi = tid / RT;      // reduction group
    const int own = tid;          // vector component this thread owns after a reduction (RT hi + g)
    const int g8 = tid & 7, h8 = tid >> 3; // phase 3 keeps an 8x8 accumulator tile whatever RT is
    const int n = a.n;
    const int npar = n * n + n;
    const int cap = a.cap;
    const int blk = a.blk_doubles;
    const int voff = blk - SADJ * NP;                        // v section of a step block
    const uint32_t xg_bytes = (HDR + 2 * SADJ * NP) * 8;     // header, X and g
    const bool vsep = voff != HDR + SADJ * NP;               // v has its own section (several seeds per trajectory)
    const bool drop_blocks = (a.flags & VA_GLV_FLAG_DISCARD) != 0; // the slabs will not be read after this call (batch > one wave of slots)
    const int64_t gslot = (int64_t)blockIdx.x * SLOTS + slot; // global slot: owns one slab and one partial-sum row
    double *const slab = a.slab + gslot * a.slab_stride;
    const int bstr = buf_doubles(blk, SADJ); // doubles per shared-memory step buffer
    double *const xg = xg_all + (size_t)slot * NB * bstr;
    double(*xs)[NP] = xs_all[slot];
    double *red = red_all[slot];
    uint64_t *mbar = mbar_all[slot];
    uint64_t *const pbar = &pbar_all[slot];
    // STAGE: this slot's copy of the current parameter set, behind the step-block buffers of all slots
    double *const pa = xg_all + (size_t)SLOTS * NB * bstr + (size_t)slot * PA_DOUBLES;
    const int64_t bstride = (int64_t)gridDim.x * SLOTS; // trajectories between two of this slot
    auto fetch_params = [&](int64_t bb) { // one thread of the slot, after every thread's last read of `pa`
        if (bb < a.B) {
            mbar_expect_tx(pbar, (uint32_t)npar * 8u);
            bulk_g2s(pa, a.params + bb * npar, (uint32_t)npar * 8u, pbar, policy_evict_first());
        }
    };
    const double tf = a.tf;
    const uint64_t keep = policy_evict_last();

    // barrier over the 64 threads of one trajectory
    auto slot_sync = [&]() {
        if (SLOTS == 1) __syncthreads();
        else asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");
    };
    auto slot_or = [&](int v) -> int {
        v = __reduce_or_sync(0xffffffffu, v);
        if ((tid & 31) == 0) st_all[slot][warp] = v;
        slot_sync();
        return st_all[slot][0] | st_all[slot][1];
    };

    // g-direction entries of a tile, as double2 offsets PO(j), j = 0..3: the 8 lanes of a group read 8 consecutive double2
    // (one conflict-free LDS.128, broadcast to the 4 groups of the warp). FG(e) = 2 PO(e>>1) + (e&1); FH is the same
    // pattern along the group index (phase 3 rows).
    auto PO = [&](int j) { return g + RT * j; };
    auto FG = [&](int e) { return 2 * PO(e >> 1) + (e & 1); };
    auto PO8 = [&](int j) { return g8 + 8 * j; };
    auto FG8 = [&](int e) { return 2 * PO8(e >> 1) + (e & 1); };
    auto FH = [&](int e) { return 2 * h8 + (e & 1) + 16 * (e >> 1); };

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NB; ++i) mbar_init(&mbar[i], 1);
        mbar_init(pbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t mbar_parity = 0; // bit i: parity of the next completion of mbar[i]; bit 16: of pbar

    // summed mode: this slot's partial-sum row; every thread zeroes exactly the entries it later adds to
    double *const part = a.partial + gslot * npar;
    if (a.reduce == VA_REDUCE_SUM && a.n_out > 0) {
        if (own < n) part[own] = 0.0;
#if VA_T8_P3
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J < 8; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int row = 32 * warp + 8 * I + ((tid & 31) >> 2), col = 8 * J + 2 * (tid & 3) + e;
                    if (row < n && col < n) part[n + row * n + col] = 0.0;
                }
#else
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int row = FH(k), col = FG8(c);
                if (row < n && col < n) part[n + row * n + col] = 0.0;
            }
#endif
    }
    __syncthreads();
    if (STAGE && tid == 0) fetch_params(gslot); // the slot's first parameter set

    // y_own = sum_c M[k][c] xin[FG(c)] summed over the group; M is held permuted (register row k <-> tile row k ^ g): lane
    // g ^ j keeps the partial sum of lane g's row in register j, so the group sum is ONE round of 7 independent double
    // shuffles and an add tree (recursive halving needs as many shuffles in three dependent rounds: measured 2.5 % slower).
    // `after_sync` runs right behind the barrier, `extra` between the DFMAs and the exchange (work that does not depend on
    // the result, off the critical path).
    // Besides the sum the product returns y = c1 * sum + c0 (c1, c0 are set by `extra`): the value the NEXT product needs.
    // With four partial sums it is built as fma(c1, t2 + t3, fma(c1, t1, fma(c1, t0, c0))) -- two dependent operations behind
    // the last shuffle instead of three.
    auto matvec = [&](const double(&M)[RT][CT], double X, int p, auto &&after_sync, auto &&extra, const double &c1, const double &c0,
                      double &y) -> double {
        xs[p][own] = X;
        slot_sync();
        after_sync();
        const double2 *xv = reinterpret_cast<const double2 *>(xs[p]);
        double sc[NC][RT];
#pragma unroll
        for (int j = 0; j < CT / 2; ++j) {
            const double2 v = xv[PO(j)];
            const int ch = j % NC;
            if (j < NC) {
#pragma unroll
                for (int k = 0; k < RT; ++k) sc[ch][k] = M[k][2 * j] * v.x;
            } else {
#pragma unroll
                for (int k = 0; k < RT; ++k) sc[ch][k] = fma(M[k][2 * j], v.x, sc[ch][k]);
            }
#pragma unroll
            for (int k = 0; k < RT; ++k) sc[ch][k] = fma(M[k][2 * j + 1], v.y, sc[ch][k]);
        }
        double s[RT];
#pragma unroll
        for (int k = 0; k < RT; ++k)
            s[k] = NC == 1 ? sc[0][k] : NC == 2 ? sc[0][k] + sc[1 % NC][k] : (sc[0][k] + sc[1 % NC][k]) + (sc[2 % NC][k] + sc[3 % NC][k]);
        extra();
        double t[RT];
        t[0] = s[0];
#pragma unroll
        for (int j = 1; j < RT; ++j) t[j] = shx(s[j], j);
        if (RT == 4) {
            const double q = t[2 % RT] + t[3 % RT];
            y = fma(c1, q, fma(c1, t[1], fma(c1, t[0], c0)));
            return (t[0] + t[1]) + q;
        }
        const double sum = RT == 2 ? t[0] + t[1]
                                   : ((t[0] + t[1]) + (t[2 % RT] + t[3 % RT])) + ((t[4 % RT] + t[5 % RT]) + (t[6 % RT] + t[7 % RT]));
        y = fma(c1, sum, c0);
        return sum;
    };
    auto nop = [] {};
    const double zero = 0.0;
    double ydummy;

    for (int64_t b = gslot; b < a.B; b += (int64_t)gridDim.x * SLOTS) {
        const double *pb = a.params + b * npar;
        double M[RT][CT];
        // ================================ phase 1: forward sweep =====================================
        if (STAGE) { // the parameter set was requested a reverse sweep ago
            mbar_wait(pbar, (mbar_parity >> 16) & 1);
            mbar_parity ^= 1u << 16;
        }
        // tile rows RT hi + (k ^ g), columns FG(c)
#pragma unroll
        for (int k = 0; k < RT; ++k) {
            const int row = RT * hi + (k ^ g);
            if (EXACT) {
                const double2 *src = reinterpret_cast<const double2 *>((STAGE ? pa : pb) + NP + row * NP);
#pragma unroll
                for (int j = 0; j < CT / 2; ++j) {
                    const double2 v = STAGE ? src[PO(j)] : __ldg(src + PO(j));
                    M[k][2 * j] = v.x;
                    M[k][2 * j + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    const int col = FG(c);
                    {
                        // padded entries: the load goes to a valid address and is discarded (the compiler may turn a guarded
                        // load into load + select, which must not leave the parameter block)
                        const bool in = row < n && col < n;
                        const int at = n + (in ? row * n + col : 0);
                        const double v = STAGE ? pa[at] : __ldg(pb + at);
                        M[k][c] = in ? v : 0.0;
                    }
                }
            }
        }
        double r_own = 0.0, x = 0.0;
        {
            const int oi = own < n ? own : 0; // padded lanes read a valid address and discard it
            const double rv = STAGE ? pa[oi] : __ldg(pb + oi), xv0 = __ldg(a.x0 + b * n + oi);
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
            // X = new solution. f(xnew) is evaluated now: it is the FSAL stage of dopri5, and for the other steppers the
            // first slope of the next step (speculative: discarded if the step is rejected).
            // default_error_checker::error, max norm over species: each warp's maximum crosses to the other warp together
            // with the operands of this product (no barrier of its own) unless the error needs the FSAL slope.
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
        fence_proxy_async_global(); // generic-proxy slab writes -> visible to the TMA reads of the reverse sweep
        status = slot_or(status);
        const bool failed = status & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS);
        const double x_tf = x, t_final = t;
        if (own < n && !a.skip_forward) a.x_final[b * n + own] = failed ? nan("") : x;
        if (tid == 0 && !a.skip_forward) {
            if (a.n_accept) a.n_accept[b] = T;
            if (a.n_reject) a.n_reject[b] = rejects;
            if (a.status) a.status[b] = status;
        }
        if (a.n_out <= 0 || (STAGE && failed)) {
            // no reverse sweep will read the staged parameter set (the barrier inside slot_or is behind every thread's reads of it)
            if (STAGE && tid == 0) fetch_params(b + bstride);
            if (a.n_out <= 0) continue;
        }

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
            // stream the step blocks back, newest first: iteration it <-> step T-1-it, buffer it % NB, NB-1 blocks ahead. The first
            // requests go out BEFORE the tile is cut, so that they travel meanwhile: every thread is past its reads of the buffers --
            // the barrier inside slot_or (first seed; the previous trajectory ended behind one as well), or the one here
            auto issue2 = [&](int it) {
                if (it < T) {
                    const int bi = it % NB;
                    mbar_expect_tx(&mbar[bi], xg_bytes);
                    bulk_g2s(xg + bi * bstr, slab + (int64_t)(T - 1 - it) * blk, xg_bytes, &mbar[bi], keep);
                }
            };
            if (o > 0) slot_sync(); // the gradient accumulation of the previous seed read the buffers
            if (tid == 0)
                for (int it = 0; it < NB - 1; ++it) issue2(it);
            // transposed tile: M[k][c] = A[FG(c)][RT hi + (k ^ g)]; the owned component stays `own`. A is re-read (L2 hit),
            // once per seed: the tile must not stay live across phase 3, whose accumulator needs its registers.
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const int row = FG(c);
#pragma unroll
                for (int k = 0; k < RT; ++k) {
                    const int col = RT * hi + (k ^ g);
                    if (EXACT) M[k][c] = STAGE ? pa[NP + row * NP + col] : ldg_hint(pb + NP + row * NP + col, drop);
                    else {
                        const bool in = row < n && col < n;
                        const int at = n + (in ? row * n + col : 0);
                        const double v = STAGE ? pa[at] : ldg_hint(pb + at, drop);
                        M[k][c] = in ? v : 0.0;
                    }
                }
            }
            double lam;
            if (a.objective == VA_OBJ_SUM) lam = (own < n) ? 1.0 : 0.0;
            else if (a.objective == VA_OBJ_HALF_NORM2) lam = x_tf;
            else lam = (own < n) ? lam_io[own < n ? own : 0] : 0.0;
            double rbar = 0.0;

            if (STAGE) {
                slot_sync(); // every thread has cut its tile from the staged parameter set
                if (tid == 0 && o == a.n_out - 1) fetch_params(b + bstride); // the slot's next one travels while this reverse sweep runs
            }
            double t_hi = t_final;
            for (int it = 0; it < T; ++it) {
                const int step = T - 1 - it, bi = it % NB;
                mbar_wait(&mbar[bi], (mbar_parity >> bi) & 1);
                mbar_parity ^= 1u << bi;
                const double *bs = xg + bi * bstr;
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
            if (own < n) lam_io[own] = lam;

            // ================================ phase 3: gradient accumulation =====================================
            // Abar[i][j] += v_m[i] X_{m-1}[j] over all steps and stages; accumulator tile rows FH(k), columns FG(c)
            fence_proxy_async_global(); // the v sections were written through the generic proxy
            slot_sync();
#if VA_T8_P3
            // ---- FP64 tensor instructions. Abar(64 x 64) += V(64 x K) X(K x 64)^T with K = SADJ T (stage, step) pairs. Warp w of the
            // trajectory owns rows 32 w .. 32 w + 31: 4 x 8 accumulator tiles of 8 x 8 (64 registers per lane, as before), and per
            // group of four (stage, step) pairs 4 A fragments (v) + 8 B fragments (X) feed 32 DMMAs = 8192 FMAs: 12 LDS.64 where the
            // DFMA form needs 32 LDS.128 for the same work. DMMA shares the DFMA pipe on B200 (tools/microbench/fp64_mma.cu), so the
            // gain is load/store-unit relief for the trajectories that sit in their latency-bound phases on the same SM.
            // Two steps (2 SADJ = 12 stages = 3 groups of four when SADJ = 6) are consumed per iteration: buffers {0,1} / {2,3}.
            static_assert(NB == 4, "phase 3 works on pairs of step buffers");
            auto issue3 = [&](int it) { // step T-1-it -> buffer it % NB, vector by vector at the padded stride
                if (it < T) {
                    const int bi = it % NB;
                    const double *src = slab + (int64_t)(T - 1 - it) * blk;
                    double *dst = xg + bi * bstr;
                    mbar_expect_tx(&mbar[bi], 2 * SADJ * NP * 8);
#pragma unroll 1
                    for (int m = 0; m < SADJ; ++m) {
                        bulk_g2s(dst + m * VS3, src + HDR + m * NP, NP * 8, &mbar[bi], keep);
                        bulk_g2s(dst + (SADJ + m) * VS3, src + voff + m * NP, NP * 8, &mbar[bi], keep);
                    }
                }
            };
            if (tid == 0) { issue3(0); issue3(1); }
            double C[4][8][2];
#pragma unroll
            for (int I = 0; I < 4; ++I)
#pragma unroll
                for (int J = 0; J < 8; ++J) C[I][J][0] = C[I][J][1] = 0.0;
            const int fr = (tid & 31) >> 2, fk = tid & 3; // fragment row / k index of this lane
            constexpr int NG = (2 * SADJ + 3) / 4;        // groups of four stages per pair of steps
            for (int it = 0; it < T; it += 2) {
                const int b0 = it % NB; // buffers b0, b0 + 1 hold steps it, it + 1
                slot_sync();            // every thread is done with the previous pair: its buffers can be refilled
                if (tid == 0) { issue3(it + 2); issue3(it + 3); }
                mbar_wait(&mbar[b0], (mbar_parity >> b0) & 1);
                mbar_parity ^= 1u << b0;
                const bool two = it + 1 < T;
                if (two) {
                    mbar_wait(&mbar[b0 + 1], (mbar_parity >> (b0 + 1)) & 1);
                    mbar_parity ^= 1u << (b0 + 1);
                }
                const double *pb2 = xg + b0 * bstr;
#pragma unroll
                for (int gq = 0; gq < NG; ++gq) {
                    const int sg = 4 * gq + fk;                 // stage inside the pair handled by this lane's k index
                    const int st = sg / SADJ, m = sg - st * SADJ; // step inside the pair, stage
                    const bool valid = sg < 2 * SADJ && (st == 0 || two);
                    const double *xv = pb2 + (valid ? st * bstr + m * VS3 : 0);
                    double af[4], bf[8];
#pragma unroll
                    for (int I = 0; I < 4; ++I) {
                        const double v = xv[SADJ * VS3 + 32 * warp + 8 * I + fr];
                        af[I] = valid ? v : 0.0;
                    }
#pragma unroll
                    for (int J = 0; J < 8; ++J) {
                        const double v = xv[8 * J + fr];
                        bf[J] = valid ? v : 0.0;
                    }
#pragma unroll
                    for (int I = 0; I < 4; ++I)
#pragma unroll
                        for (int J = 0; J < 8; ++J) dmma(C[I][J][0], C[I][J][1], af[I], bf[J]);
                }
            }
            if (own < n) {
                if (a.reduce == VA_REDUCE_NONE) mu_o[own] = rbar;
                else atomicAdd(part + own, rbar);
            }
#pragma unroll
            for (int I = 0; I < 4; ++I) {
                const int row = 32 * warp + 8 * I + fr;
#pragma unroll
                for (int J = 0; J < 8; ++J) {
                    const int col = 8 * J + 2 * fk;
                    if (a.reduce == VA_REDUCE_NONE) {
                        if (EXACT) *reinterpret_cast<double2 *>(mu_o + NP + row * NP + col) = make_double2(C[I][J][0], C[I][J][1]);
                        else {
                            if (row < n && col < n) mu_o[n + row * n + col] = C[I][J][0];
                            if (row < n && col + 1 < n) mu_o[n + row * n + col + 1] = C[I][J][1];
                        }
                    } else {
                        // summed objective: fire-and-forget FP64 reductions into this slot's partial-sum row (one writer per address)
                        if (row < n && col < n) atomicAdd(part + n + row * n + col, C[I][J][0]);
                        if (row < n && col + 1 < n) atomicAdd(part + n + row * n + col + 1, C[I][J][1]);
                    }
                }
            }
        }
#else
            auto issue3 = [&](int it) {
                if (it < T) {
                    const int bi = it % NB;
                    const double *src = slab + (int64_t)(T - 1 - it) * blk;
                    if (vsep) {
                        mbar_expect_tx(&mbar[bi], (HDR + 2 * SADJ * NP) * 8);
                        bulk_g2s(xg + bi * bstr, src, (HDR + SADJ * NP) * 8, &mbar[bi], keep);
                        bulk_g2s(xg + bi * bstr + voff, src + voff, SADJ * NP * 8, &mbar[bi], keep);
                    } else {
                        mbar_expect_tx(&mbar[bi], xg_bytes);
                        bulk_g2s(xg + bi * bstr, src, xg_bytes, &mbar[bi], keep);
                    }
                }
            };
            if (tid == 0)
                for (int it = 0; it < NB - 1; ++it) issue3(it);
            double Ab[8][8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int c = 0; c < 8; ++c) Ab[k][c] = 0.0;
            for (int it = 0; it < T; ++it) {
                const int bi = it % NB;
                // every thread is done with iteration it-1: its buffer can be refilled. (Measured alternative: a per-buffer
                // "released" mbarrier with one arrival per warp instead of this 64-thread barrier, so that the two warps of the
                // slot need not meet once per step -- 5.83 M against 5.95 M gradients/s: the warps drifting apart costs more than
                // the barrier, tools/gpu_r2_p3sync.sh.)
                slot_sync();
                if (tid == 0) issue3(it + NB - 1);
                mbar_wait(&mbar[bi], (mbar_parity >> bi) & 1);
                mbar_parity ^= 1u << bi;
                const double *bs = xg + bi * bstr;
#pragma unroll
                for (int m = SADJ; m >= 1; --m) {
                    const double2 *vv = reinterpret_cast<const double2 *>(bs + voff + (m - 1) * NP) + h8;
                    const double2 *xx = reinterpret_cast<const double2 *>(bs + HDR + (m - 1) * NP);
                    double vr[8], xc[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double2 p = vv[8 * j], q = xx[PO8(j)];
                        vr[2 * j] = p.x; vr[2 * j + 1] = p.y;
                        xc[2 * j] = q.x; xc[2 * j + 1] = q.y;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)
#pragma unroll
                        for (int c = 0; c < 8; ++c) Ab[k][c] = fma(vr[k], xc[c], Ab[k][c]);
                }
            }
            if (drop_blocks && o == a.n_out - 1) {
                // last use of this trajectory's step blocks: drop the 128-byte lines that lie completely inside blocks 0..T-1 (a block
                // is 6208 bytes, so the slab's first and last touched lines may be shared with the header of block T: left alone)
                const uintptr_t lo = (reinterpret_cast<uintptr_t>(slab) + 127) & ~(uintptr_t)127;
                const uintptr_t hi_ = (reinterpret_cast<uintptr_t>(slab + (int64_t)T * blk)) & ~(uintptr_t)127;
                for (uintptr_t p = lo + (uintptr_t)tid * 128; p < hi_; p += (uintptr_t)NTT * 128) discard_line(reinterpret_cast<const double *>(p));
            }
            if (a.reduce == VA_REDUCE_NONE) {
                if (own < n) mu_o[own] = rbar;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int row = FH(k);
                    if (EXACT) {
                        double2 *dst = reinterpret_cast<double2 *>(mu_o + NP + row * NP);
#pragma unroll
                        for (int j = 0; j < 4; ++j) dst[PO8(j)] = make_double2(Ab[k][2 * j], Ab[k][2 * j + 1]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const int col = FG8(c);
                            if (row < n && col < n) mu_o[n + row * n + col] = Ab[k][c];
                        }
                    }
                }
            } else {
                // summed objective: fire-and-forget FP64 reductions into this CTA's private partial-sum row (one writer
                // per address, program order -> deterministic)
                if (own < n) atomicAdd(part + own, rbar);
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int row = FH(k), col = FG8(c);
                        if (row < n && col < n) atomicAdd(part + n + row * n + col, Ab[k][c]);
                    }
            }
        }
#endif
        slot_sync(); // slab and shared buffers are reused by the next trajectory
    }
}

// Staging needs room for SLOTS parameter sets next to the step-block buffers, one seed per trajectory, and 16-byte aligned
// parameter sets (the bulk copy's requirement; cudaMalloc'ed and page-locked buffers are, a caller's sub-array may not be)
bool stage_ok(int blk_doubles, int sadj, int n_out, const void *params)
{
    if (!VA_T8_STAGE || VA_T8_P3 || SLOTS != 4 || n_out > 1) return false;
    if (const char *env = getenv("VA_T8_NO_STAGE"))
        if (atoi(env) != 0) return false;
    if (reinterpret_cast<uintptr_t>(params) & 15) return false;
    return (size_t)SLOTS * (NB_STAGED * buf_doubles(blk_doubles, sadj) + PA_DOUBLES) * 8 <= 220 * 1024;
}
size_t smem_bytes(int blk_doubles, int sadj, bool stage)
{
    return (size_t)SLOTS * ((stage ? NB_STAGED : NB) * buf_doubles(blk_doubles, sadj) + (stage ? PA_DOUBLES : 0)) * 8;
}

template <class Tab, bool ADAPTIVE, bool EXACT, bool STAGE>
cudaError_t launch_k(const VaGlvWideArgs &a, cudaStream_t st)
{
    const size_t smem = smem_bytes(a.blk_doubles, Tab::SADJ, STAGE);
    cudaError_t e = cudaFuncSetAttribute(k_glv_t8<Tab, ADAPTIVE, EXACT, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_glv_t8<Tab, ADAPTIVE, EXACT, STAGE><<<a.grid, NT, smem, st>>>(a);
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
    const bool stage = stage_ok(a.blk_doubles, Tab::SADJ, a.n_out, a.params);
    if (a.n == NP) return stage ? launch_k<Tab, ADAPTIVE, true, true>(a, st) : launch_k<Tab, ADAPTIVE, true, false>(a, st);
    return stage ? launch_k<Tab, ADAPTIVE, false, true>(a, st) : launch_k<Tab, ADAPTIVE, false, false>(a, st);
}

template <class Tab, bool ADAPTIVE>
cudaError_t occupancy(int n, size_t smem, bool stage, int *ctas_per_sm)
{
    auto occ = [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kernel, NT, smem);
    };
    if (n == NP) return stage ? occ(k_glv_t8<Tab, ADAPTIVE, true, true>) : occ(k_glv_t8<Tab, ADAPTIVE, true, false>);
    return stage ? occ(k_glv_t8<Tab, ADAPTIVE, false, true>) : occ(k_glv_t8<Tab, ADAPTIVE, false, false>);
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

bool va_glv_t8_supported(int n, int stepper, int adaptive)
{
    if (n <= 32 || n > NP) return false; // smaller systems: va_glv_wide.cu (warp per trajectory)
    if (stepper == VA_RK_RK4) return !adaptive;
    if (stepper == VA_RK_CK54 || stepper == VA_RK_DOPRI5) return adaptive != 0;
    return false;
}

// step block: [8-double header | X_0..X_{s-1} | g_0..g_{s-1} | v_1..v_s]; with one seed per trajectory v aliases g
int va_glv_t8_block_doubles(int stepper, int n_out) { return HDR + (n_out > 1 ? 3 : 2) * sadj_of(stepper) * NP; }
int va_glv_t8_header_doubles() { return HDR; }

cudaError_t va_glv_t8_config(int n, int stepper, int n_out, int device, int *grid, int *ctas_per_sm, int *threads, int *slots_per_cta)
{
    int sms = 0;
    cudaError_t err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err != cudaSuccess) return err;
    // the staged form is what the launcher picks for one seed per trajectory (an unaligned parameter pointer falls back at launch time
    // to the form with the smaller shared-memory footprint, so its occupancy is at least this)
    const bool stage = stage_ok(va_glv_t8_block_doubles(stepper, n_out), sadj_of(stepper), n_out, nullptr);
    const size_t smem = smem_bytes(va_glv_t8_block_doubles(stepper, n_out), sadj_of(stepper), stage);
    int occ = 0;
    switch (stepper) {
    case VA_RK_RK4: err = occupancy<TabRK4, false>(n, smem, stage, &occ); break;
    case VA_RK_CK54: err = occupancy<TabCK54, true>(n, smem, stage, &occ); break;
    case VA_RK_DOPRI5: err = occupancy<TabDOPRI5, true>(n, smem, stage, &occ); break;
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
    *threads = NT;
    *slots_per_cta = SLOTS;
    return cudaSuccess;
}

cudaError_t va_glv_t8_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    switch (a.stepper) {
    case VA_RK_RK4: return launch<TabRK4, false>(a, st);
    case VA_RK_CK54: return launch<TabCK54, true>(a, st);
    case VA_RK_DOPRI5: return launch<TabDOPRI5, true>(a, st);
    }
    return cudaErrorInvalidValue;
}
