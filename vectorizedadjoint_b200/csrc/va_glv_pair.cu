// va_glv_pair.cu -- Generalized Lotka-Volterra with 256 species (BASELINE config 5) with the 512 KB interaction matrix
// held ON CHIP by a thread-block cluster of two CTAs (two SMs) per trajectory: CTA `rank` owns rows
// [128 rank, 128 rank + 128) of A -- 64 of them in shared memory (128 KB, loaded once per trajectory by TMA bulk copies),
// 64 in registers (128 registers per thread) -- so no matrix byte is re-read from L2/HBM during a sweep. The ring-streamed
// kernel (va_glv_ring.cu) is bound by the L2 -> SM path (11 TB/s of matrix chunks); this one exchanges 2 KB per product
// through distributed shared memory instead.
//
// Both CTAs of a pair run the same program on the same values (thread i of either CTA owns component i of every vector;
// state, stage slopes and stage adjoints live in registers; the controller's decisions are bitwise identical), and meet
// once per matrix-vector product:
//   g = r + A X   : a CTA computes its 128 rows (row dot products: lanes stride over the columns, transposing shuffle
//                   reduction), stores them into its own AND the partner's result vector (st through map_shared_rank),
//                   then one cluster barrier;
//   A^T v         : a CTA sums over its 128 rows for all 256 columns (thread j owns column j: no reduction), stores the
//                   partial sum into the partner's buffer, one cluster barrier, then own + partner's (commutative: both CTAs
//                   get the same bits).
// Result and partial-sum buffers are double-buffered by product parity, so a fast CTA can start the next product while the
// partner still reads the last one.
//
// Same algorithm and phases as va_glv_ring.cu (reference lib/include/detail/runge_kutta.hpp:76-118 forward sweep with
// odeint's controlled stepper, detail/backpropagation.hpp:83-158, 231-254 reverse sweep; store-stages policy: every accepted
// step leaves a block [8-double header (t_n) | X_0.. | g_0.. | v_0..] in the CTA's slab): 1. forward sweep, 2. state adjoint,
// 3. gradient accumulation Abar = sum_k v_k X_k^T as a matrix product with 8 x 8 accumulators per thread -- each CTA of the
// pair takes two of the four 64-column passes.
#include <cooperative_groups.h>

#include "va_glv_common.cuh"
#include "va_tma.cuh"

namespace cg = cooperative_groups;

namespace {
using namespace va_tma;

constexpr int NT = 256;        // threads per CTA
constexpr int N = 256;         // species: thread i owns component i
constexpr int CR = 64;         // matrix rows per CTA held in registers (own rows [SR, HR)), 128 registers per thread
// CL = CTAs per cluster (2 or 4): HR = N / CL matrix rows per CTA, of which SR = HR - CR in shared memory (own rows [0, SR))
constexpr int PCOLS = 64;      // columns of Abar per accumulation pass
constexpr int SADJ_MAX = 6;
constexpr int P3_STAGES = 4;   // phase 3: step blocks (v and X operands) in flight
constexpr int P3_STAGE_DOUBLES = SADJ_MAX * (N + PCOLS);
constexpr uint32_t TMA_PIECE = 32768;
static_assert(N == NT && CR == 64, "geometry");

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// sums of 8 values over the 32 lanes with 4 + 2 + 1 + 1 + 1 exchanges; lane l ends up with the total of value
// 4 bit4(l) + 2 bit3(l) + bit2(l) (all four lanes of a quad hold it)
__device__ __forceinline__ double transpose_sum8(double (&s)[8], int lane)
{
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = b4 ? s[i] : s[i + 4], keep = b4 ? s[i + 4] : s[i];
        s[i] = keep + shx(send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = b3 ? s[i] : s[i + 2], keep = b3 ? s[i + 2] : s[i];
        s[i] = keep + shx(send, 8);
    }
    const double send = b2 ? s[0] : s[1], keep = b2 ? s[1] : s[0];
    double r = keep + shx(send, 4);
    r += shx(r, 2);
    r += shx(r, 1);
    return r;
}

// ---- distributed shared memory: remote stores that signal the receiver's mbarrier (no cluster-wide fences) --------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
    return r;
}
// 8-byte store into the partner's shared memory; its completion is counted (8 bytes) on the partner's mbarrier
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar)
{
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "l"(__double_as_longlong(v)), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Exchange protocol of one product q (both CTAs run it in step): thread 0 arms the local barrier xbar[q & 1] with the bytes
// the partner will deliver, every thread computes its share and st.async-es it into the partner's buffer [q & 1], then waits
// for the local barrier's phase (q >> 1) & 1. Buffers, barriers and the input vector xs are double-buffered by q & 1: the
// partner delivers product q + 2 only after it has received this CTA's product q + 1, which this CTA sends after the CTA
// barrier that follows all its reads of product q.
template <int CL>
struct Pair {
    static constexpr int HR = N / CL, SR = HR - CR;
    static_assert((CL == 2 || CL == 4) && SR >= 0 && SR % 8 == 0 && (SR * N * 8) % TMA_PIECE == 0, "geometry");
    unsigned rank;      // this CTA's rank in the cluster
    const double *sc;   // [SR][N] shared-memory rows (own rows 0..SR-1)
    double *xs;         // [2][N] product input (stage state / seed vector)
    double *gb;         // [2][N] product results (row phase)
    double *yp;         // [2][CL][N] partial sums by source rank (column phase)
    uint64_t *xbar;     // [2] exchange barriers
    uint32_t gb_remote[CL - 1], yp_remote[CL - 1], xbar_remote[CL - 1]; // shared::cluster addresses of peer (rank + 1 + d) % CL's buffers
    uint32_t q;         // products done
    __device__ __forceinline__ double *xin() const { return xs + (q & 1u) * N; }
};

// returns g_tid = r_tid + (A xin)_tid. P.xin() must be visible to all threads of this CTA (all CTAs hold the same vector).
template <int CL>
__device__ __forceinline__ double product_rows(Pair<CL> &P, const double *rr, const double (&creg)[CR], int tid)
{
    constexpr int HR = Pair<CL>::HR, SR = Pair<CL>::SR;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t pp = P.q & 1u, par = (P.q >> 1) & 1u;
    if (tid == 0) mbar_expect_tx(&P.xbar[pp], (CL - 1) * HR * 8); // the peers' rows
    const double *xin = P.xin();
    double xr[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double2 t = *reinterpret_cast<const double2 *>(xin + 64 * k + 2 * lane);
        xr[2 * k] = t.x;
        xr[2 * k + 1] = t.y;
    }
    double *mine = P.gb + pp * N;
    const int sub = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    double s[8];
    // register rows: warp w holds own rows SR + 8 w + r, lane l their columns {64k + 2l, 64k + 2l + 1}
#pragma unroll
    for (int r = 0; r < 8; ++r) { // two chains per row
        double acc0 = creg[r * 8] * xr[0], acc1 = creg[r * 8 + 4] * xr[4];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            acc0 = fma(creg[r * 8 + k], xr[k], acc0);
            acc1 = fma(creg[r * 8 + 4 + k], xr[4 + k], acc1);
        }
        s[r] = acc0 + acc1;
    }
    const double sum_reg = transpose_sum8(s, lane);
    double sum_sm = 0.0;
    if (SR > 0) {
        // shared-memory rows: warp w takes own rows (SR / 8) w + r. (No store or asm statement between the two parts: the
        // compiler hoists these loads above the register rows' arithmetic.)
        constexpr int RW = SR / 8; // 8 at CL = 2
        static_assert(SR == 0 || RW == 8, "shared-memory rows: 8 per warp");
        const double *rowp = P.sc + (size_t)(8 * warp) * N + 2 * lane;
#pragma unroll
        for (int r0 = 0; r0 < 8; r0 += 4) { // four rows (16 LDS.128) in flight, two chains per row
            double2 t[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) t[r][k] = *reinterpret_cast<const double2 *>(rowp + (r0 + r) * N + 64 * k);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                double acc0 = t[r][0].x * xr[0], acc1 = t[r][2].x * xr[4];
                acc0 = fma(t[r][0].y, xr[1], acc0);
                acc1 = fma(t[r][2].y, xr[5], acc1);
                acc0 = fma(t[r][1].x, xr[2], acc0);
                acc1 = fma(t[r][3].x, xr[6], acc1);
                acc0 = fma(t[r][1].y, xr[3], acc0);
                acc1 = fma(t[r][3].y, xr[7], acc1);
                s[r0 + r] = acc0 + acc1;
            }
        }
        sum_sm = transpose_sum8(s, lane);
    }
    if ((lane & 3) == 0) {
        const int row_reg = (int)P.rank * HR + SR + 8 * warp + sub;
        const double g_reg = rr[row_reg] + sum_reg;
        mine[row_reg] = g_reg;
#pragma unroll
        for (int d = 0; d < CL - 1; ++d) st_async_f64(P.gb_remote[d] + (pp * N + row_reg) * 8u, g_reg, P.xbar_remote[d] + pp * 8u);
        if (SR > 0) {
            const int row_sm = (int)P.rank * HR + 8 * warp + sub;
            const double g_sm = rr[row_sm] + sum_sm;
            mine[row_sm] = g_sm;
#pragma unroll
            for (int d = 0; d < CL - 1; ++d) st_async_f64(P.gb_remote[d] + (pp * N + row_sm) * 8u, g_sm, P.xbar_remote[d] + pp * 8u);
        }
    }
    __syncthreads();                      // the own rows are visible to the CTA
    mbar_wait_or_trap(&P.xbar[pp], par);  // the peers' rows have landed
    const double g = mine[tid];
    ++P.q;
    return g;
}

// returns (A^T v)_tid for v = P.xin(), visible to all threads of this CTA. Same register/shared-memory row layout as product_rows
// ("transposed row product"): warp w forms, for the 8 columns of each lane, the partial sums over its 16 rows (8 in registers, 8
// in shared memory) -- the v operands are 16 warp-uniform values instead of the 128 every thread would need as the owner of a
// whole column (an LDS.128 costs four LSU wavefronts even when all lanes read the same address: the column-owner version spent
// 2/3 of its LSU time on those broadcasts and was LSU-bound at 384 wavefronts per warp and product; this one needs 176) -- then
// the 8 warps' partial rows are summed through shared memory (wpart: [8][N] doubles, free between products).
template <int CL>
__device__ __forceinline__ double product_cols(Pair<CL> &P, const double (&creg)[CR], double *wpart, int tid)
{
    constexpr int HR = Pair<CL>::HR, SR = Pair<CL>::SR;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t pp = P.q & 1u, par = (P.q >> 1) & 1u;
    if (tid == 0) mbar_expect_tx(&P.xbar[pp], (CL - 1) * N * 8); // the peers' partial sums for all columns
    const double *vo = P.xin() + (int)P.rank * HR; // v of the own rows
    double pc[8];                                  // pc[2k + e]: column 64 k + 2 lane + e
    {
        double vr[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 t = *reinterpret_cast<const double2 *>(vo + SR + 8 * warp + 2 * i);
            vr[2 * i] = t.x;
            vr[2 * i + 1] = t.y;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) pc[k] = creg[k] * vr[0];
#pragma unroll
        for (int r = 1; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) pc[k] = fma(creg[r * 8 + k], vr[r], pc[k]);
    }
    if (SR > 0) {
        double vs[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 t = *reinterpret_cast<const double2 *>(vo + 8 * warp + 2 * i);
            vs[2 * i] = t.x;
            vs[2 * i + 1] = t.y;
        }
        const double *rowp = P.sc + (size_t)(8 * warp) * N + 2 * lane;
#pragma unroll
        for (int r0 = 0; r0 < 8; r0 += 4) { // four rows (16 LDS.128) in flight
            double2 t[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) t[r][k] = *reinterpret_cast<const double2 *>(rowp + (r0 + r) * N + 64 * k);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    pc[2 * k] = fma(t[r][k].x, vs[r0 + r], pc[2 * k]);
                    pc[2 * k + 1] = fma(t[r][k].y, vs[r0 + r], pc[2 * k + 1]);
                }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<double2 *>(wpart + warp * N + 64 * k + 2 * lane) = make_double2(pc[2 * k], pc[2 * k + 1]);
    __syncthreads();
    double part = wpart[tid];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) part += wpart[w * N + tid];
#pragma unroll
    for (int d = 0; d < CL - 1; ++d) st_async_f64(P.yp_remote[d] + ((pp * CL + P.rank) * N + tid) * 8u, part, P.xbar_remote[d] + pp * 8u);
    mbar_wait_or_trap(&P.xbar[pp], par);
    double y;
    if (CL == 2) {
        y = part + P.yp[(pp * CL + (P.rank ^ 1u)) * N + tid]; // IEEE addition is commutative: both CTAs get the same bits
    } else {
        // sum in rank order, the own partial sum in its place: every CTA of the cluster gets the same bits
        y = 0.0;
#pragma unroll
        for (int r = 0; r < CL; ++r) {
            const double term = (unsigned)r == P.rank ? part : P.yp[(pp * CL + r) * N + tid];
            y = r == 0 ? term : y + term;
        }
    }
    ++P.q;
    return y;
}

// register rows of the CTA's matrix part, row layout: warp w holds own rows SR + 8 w + r, lane l their columns {64k + 2l, 64k + 2l + 1}
// (always inlined: the register array must never be addressed through a pointer)
template <bool EXACT, int SR>
__device__ __forceinline__ void load_rows_f(double (&creg)[CR], const double *Ac, const double *pb, int n, int row0, int warp, int lane)
{
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (EXACT) {
                const double2 t = __ldg(reinterpret_cast<const double2 *>(Ac + (size_t)(warp * 8 + r) * N + 64 * k + 2 * lane));
                creg[r * 8 + 2 * k] = t.x;
                creg[r * 8 + 2 * k + 1] = t.y;
            } else {
                const int row = row0 + SR + warp * 8 + r, col = 64 * k + 2 * lane;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool in = row < n && col + e < n;
                    const double v = __ldg(pb + n + (in ? (size_t)row * n + col + e : 0)); // padded entries read a valid address and discard it
                    creg[r * 8 + 2 * k + e] = in ? v : 0.0;
                }
            }
        }
}
// EXACT: a.n == N. Otherwise 64 < a.n < N species are padded with inert ones (x = r = 0, zero matrix rows and columns): the
// parameter blocks keep their row stride a.n in global memory, the kernel's vectors and step blocks are N wide.
// SEG: recompute policy (north_star item 4; the reference's policy, detail/backpropagation.hpp:24-64). The forward sweep keeps
// only (t_n, x_n) per accepted step in the CTA's state store; the reverse sweep walks the trajectory in segments of a.seg_len
// steps, newest first: it re-integrates a segment from the stored states with dt = t_{n+1} - t_n (StateStorage::GetDt,
// StateStorage.hpp:22), leaving the usual stage blocks in a slab of seg_len blocks, then runs phases 2 and 3 over that
// segment. Six more products per step (18 instead of 12); checkpoint memory 8 (N + 8) B per step instead of 36.9 KB.
template <class Tab, bool ADAPTIVE, int CL, bool EXACT, bool SEG>
__global__ void __launch_bounds__(NT, 1) k_glv_pair(const __grid_constant__ VaGlvWideArgs a)
{
    constexpr int HR = Pair<CL>::HR, SR = Pair<CL>::SR;
    constexpr int S = Tab::S, SADJ = Tab::SADJ;
    constexpr int SE = Tab::FSAL ? S - 1 : S;
    constexpr int BLK = 8 + 3 * SADJ * N; // [header | X_0.. | g_0.. | v_0..]
    constexpr int OFF_X = 8, OFF_G = 8 + SADJ * N, OFF_V = 8 + 2 * SADJ * N;
    static_assert(SADJ <= SADJ_MAX, "staging buffers");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sc = reinterpret_cast<double *>(smem_raw); // [SR][N]
    double *xs = sc + (size_t)SR * N;                   // [2][N] stage state / seed vector handed to a product
    double *gb = xs + 2 * N;                            // [2][N]
    double *yp = gb + 2 * N;                            // [2][CL][N]
    double *rr = yp + 2 * CL * N;                       // growth rates r
    double *p3buf = rr + N;                             // [P3_STAGES][SADJ x N of v | SADJ x PCOLS of X]  phase 3 operands
    double *red = p3buf + P3_STAGES * P3_STAGE_DOUBLES; // [8] error-norm partials
    uint64_t *bar = reinterpret_cast<uint64_t *>(red + 8); // [0]: matrix rows landed (TMA); [1], [2]: exchange; [3..]: phase 3 ring
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = EXACT ? N : a.n;
    const int npar = n * n + n;
    const bool live = tid < n; // this thread's component exists

    Pair<CL> P;
    P.rank = cg::this_cluster().block_rank();
    P.sc = sc;
    P.xs = xs;
    P.gb = gb;
    P.yp = yp;
    P.xbar = bar + 1;
#pragma unroll
    for (int d = 0; d < CL - 1; ++d) {
        const uint32_t peer = (P.rank + 1u + (uint32_t)d) % CL;
        P.gb_remote[d] = mapa_u32(smem_u32(gb), peer);
        P.yp_remote[d] = mapa_u32(smem_u32(yp), peer);
        P.xbar_remote[d] = mapa_u32(smem_u32(bar + 1), peer);
    }
    P.q = 0;
    const unsigned rank = P.rank;
    const int64_t pair_id = blockIdx.x / CL, n_pairs = gridDim.x / CL;
    uint32_t bar_parity = 0;
    uint32_t p3q = 0; // phase 3 ring: step blocks consumed since kernel start (stage = p3q % P3_STAGES, parity = (p3q / P3_STAGES) & 1)
    const uint64_t pol_slab = policy_evict_normal();
    const uint64_t pol = policy_evict_first(); // the matrix is read once per sweep: do not let it displace the slabs in L2
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_init(bar + 2, 1);
#pragma unroll 1
        for (int k = 0; k < P3_STAGES; ++k) mbar_init(bar + 3 + k, 1);
        mbar_init_fence();
    }
    cluster_barrier(); // barriers initialised; the partner has started (its shared memory may be written from here on)
    double *const slab = a.slab + (int64_t)blockIdx.x * a.slab_stride;
    constexpr int XB = 8 + N; // state-store entry: [8-double header (t_n) | x_n]
    double *const xstore = SEG ? a.xstore + (int64_t)blockIdx.x * a.xstore_stride : nullptr;
    // SPARSE checkpoints (a.sparse, SURVEY section 8 f3: "checkpoint scheduling beyond store-all"): the store keeps t_n of every
    // accepted step but x_n only of the first step of each segment (every seg_len-th step); the segment re-integration below
    // then carries the state from step to step itself. 8 B + 8 N / seg_len B per step instead of 8 (N + 8) B.
    //   layout: [t_0 .. t_cap+1 | pad to 8 | x of step 0 | x of step seg_len | ...]
    const bool sparse = SEG && a.sparse;
    const int64_t xs_states = ((int64_t)a.cap + 2 + 7) / 8 * 8;
    auto ck_time = [&](int nn) -> double * { return sparse ? xstore + nn : xstore + (int64_t)nn * XB; };
    auto ck_state = [&](int nn) -> double * { return sparse ? xstore + xs_states + (int64_t)(nn / a.seg_len) * N : xstore + (int64_t)nn * XB + 8; };
    double creg[CR];
    bool row_init = false; // summed mode: this pair's partial-sum row has been written

    for (int64_t b = pair_id; b < a.B; b += n_pairs) {
        const double *pb = a.params + b * npar;
        const double *Aown = pb + n + (size_t)rank * HR * n; // own rows; [0, SR) -> shared memory, [SR, HR) -> registers
        const double *Ac = Aown + (size_t)SR * n;
        const int row0 = (int)rank * HR; // global index of own row 0
        // ------------------------------------------ forward sweep ------------------------------------------------------
        if (SR > 0 && !EXACT) {
            // padded shared-memory rows: thread j fills column j (coalesced), zeros outside the n x n matrix
            double *scw = sc;
#pragma unroll 8
            for (int i = 0; i < SR; ++i) {
                const bool in = live && row0 + i < n;
                const double v = __ldg(pb + n + (in ? (size_t)(row0 + i) * n + tid : 0)); // out-of-range lanes read a valid address and discard it
                scw[(size_t)i * N + tid] = in ? v : 0.0;
            }
        }
        if (SR > 0 && EXACT && tid == 0) {
            mbar_expect_tx(bar, (uint32_t)(SR * N * 8));
#pragma unroll 1
            for (uint32_t off = 0; off < (uint32_t)(SR * N * 8); off += TMA_PIECE)
                bulk_g2s(reinterpret_cast<unsigned char *>(sc) + off, reinterpret_cast<const unsigned char *>(Aown) + off, TMA_PIECE, bar, pol);
        }
        auto load_rows = [&]() { load_rows_f<EXACT, SR>(creg, Ac, pb, n, row0, warp, lane); };
        load_rows();
        rr[tid] = live ? __ldg(pb + tid) : 0.0;
        double x = live ? a.x0[b * n + tid] : 0.0;
        P.xin()[tid] = x;
        if (SR > 0 && EXACT) {
            mbar_wait_or_trap(bar, bar_parity);
            bar_parity ^= 1u;
        }
        __syncthreads();
        double t = a.ti, dt = a.dt0;
        const double tf = a.tf;
        int nck = 0, rejects = 0, status = 0, trials = 0;
        double K[S];
        double g0 = product_rows(P, rr, creg, tid);
        K[0] = x * g0;
        bool active = ADAPTIVE ? va_less_with_sign(t, tf, dt) : va_less_eq_with_sign(t + dt, tf, dt);
        bool fresh = true;
        while (active) {
            double *blk = slab + (int64_t)nck * BLK;
            if (fresh) {
                if (nck >= a.cap) { status |= VA_TRAJ_CKPT_OVERFLOW; break; }
                if (SEG) {
                    if (!sparse || nck % a.seg_len == 0) ck_state(nck)[tid] = x;
                    if (tid == 0) *ck_time(nck) = t;
                } else {
                    blk[OFF_X + tid] = x;
                    blk[OFF_G + tid] = g0;
                    if (tid == 0) blk[0] = t;
                }
                if (ADAPTIVE && va_less_with_sign(tf, t + dt, dt)) dt = tf - t;
                trials = 0;
                fresh = false;
            }
#pragma unroll
            for (int m = 1; m < SE; ++m) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < m; ++j)
                    if (Tab::a(m, j) != 0.0) acc = fma(Tab::a(m, j), K[j], acc);
                const double xm = fma(dt, acc, x);
                P.xin()[tid] = xm;
                if (!SEG && m < SADJ) blk[OFF_X + m * N + tid] = xm;
                __syncthreads();
                const double gm = product_rows(P, rr, creg, tid);
                K[m] = xm * gm;
                if (!SEG && m < SADJ) blk[OFF_G + m * N + tid] = gm;
            }
            double xnew;
            {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < SE; ++j)
                    if (Tab::b(j) != 0.0) acc = fma(Tab::b(j), K[j], acc);
                xnew = fma(dt, acc, x);
            }
            double gnew = 0.0;
            if (Tab::FSAL) {
                P.xin()[tid] = xnew;
                __syncthreads();
                gnew = product_rows(P, rr, creg, tid);
                K[S - 1] = xnew * gnew;
            }
            bool accept = true;
            double err = 0.0;
            if (ADAPTIVE) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < S; ++j)
                    if (Tab::db(j) != 0.0) acc = fma(Tab::db(j), K[j], acc);
                double e = fabs(dt * acc) / (a.eps_abs + a.eps_rel * (fabs(x) + fabs(dt) * fabs(K[0])));
                if (!EXACT && !live) e = 0.0; // inert species (0 / 0 when eps_abs = 0)
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, d));
                __syncthreads(); // previous readers of red are done
                if (lane == 0) red[warp] = e;
                __syncthreads();
#pragma unroll
                for (int w = 0; w < NT / 32; ++w) err = fmax(err, red[w]);
                accept = !(err > 1.0);
            }
            if (!accept) {
                dt *= fmax(0.9 * inv_root<(Tab::ERROR_ORDER > 1 ? Tab::ERROR_ORDER - 1 : 1)>(err), 0.2);
                ++rejects;
                if (++trials >= 500) { status |= VA_TRAJ_NO_PROGRESS; break; }
            } else {
                x = xnew;
                ++nck;
                if (ADAPTIVE) {
                    t += dt;
                    if (err < 0.5) {
                        constexpr int PO = Tab::STEPPER_ORDER;
                        double floor_ = 1.0;
#pragma unroll
                        for (int k = 0; k < PO; ++k) floor_ *= 0.2;
                        dt *= (err <= floor_) ? 4.5 : 9.0 / 10.0 * inv_root<PO>(err);
                    }
                    active = va_less_with_sign(t, tf, dt);
                } else {
                    t = a.ti + (double)nck * dt;
                    active = va_less_eq_with_sign(t + dt, tf, dt);
                }
                fresh = true;
                if (Tab::FSAL) {
                    g0 = gnew;
                    K[0] = K[S - 1];
                } else if (active) {
                    P.xin()[tid] = x;
                    __syncthreads();
                    g0 = product_rows(P, rr, creg, tid);
                    K[0] = x * g0;
                }
            }
        }
        const int T = nck;
        if (tid == 0) {
            if (SEG) *ck_time(T) = t; // entry T carries the final time
            else slab[(int64_t)T * BLK] = t;
        }
        if (!isfinite(x)) status |= VA_TRAJ_NONFINITE;
        status = __syncthreads_or(status); // also: every slab store of the forward sweep is visible to the CTA
        const bool failed = status & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS);
        if (rank == 0) {
            if (live) a.x_final[b * n + tid] = failed ? nan("") : x;
            if (tid == 0) {
                if (a.n_accept) a.n_accept[b] = T;
                if (a.n_reject) a.n_reject[b] = rejects;
                if (a.status) a.status[b] = status;
            }
        }
        const double t_final = t;

        // ------------------------------------------ reverse sweep ------------------------------------------------------
        for (int o = 0; o < a.n_out; ++o) {
            double *lam_io = a.lambda + (b * a.n_out + o) * n;
            const bool sum_mode = a.reduce == VA_REDUCE_SUM;
            double *gbar = sum_mode ? a.partial + pair_id * npar : a.mu + (b * a.n_out + o) * npar;
            const bool overwrite = !sum_mode || !row_init; // first use of this accumulator row
            if (failed) {
                if (rank == 0) {
                    if (live) lam_io[tid] = nan("");
                    if (!sum_mode)
                        for (int k = tid; k < npar; k += NT) gbar[k] = nan("");
                }
                continue;
            }
            double lam = !live ? 0.0 : a.objective == VA_OBJ_SUM ? 1.0 : a.objective == VA_OBJ_HALF_NORM2 ? x : lam_io[tid];
            double rbar = 0.0;
            double t_hi = t_final;
            // the trajectory is walked in segments [s0, s1) of steps, newest first (store-stages policy: one segment, all steps)
            int s1 = T;
            bool first_seg = true;
            do {
            // segment boundaries are multiples of seg_len counted from step 0 (the newest segment may be shorter), so that every
            // segment starts on a step whose state the sparse store holds
            const int s0 = SEG && s1 > 0 ? (s1 - 1) / a.seg_len * a.seg_len : 0; // s1 == 0: a trajectory without steps (ti == tf)
            const int Tseg = s1 - s0;
            // The register rows are reloaded at the top of EVERY segment (and seed), although they never change: an unconditional
            // reload ends their live range before phase 3, whose 128 accumulator registers cannot coexist with them (a conditional or
            // missing reload makes ptxas spill a third of the rows for the whole kernel).
            load_rows();
            if (SEG) {
                // ---- re-integration of the segment from the stored states ----
                double xcarry = 0.0; // sparse store: x_{nn} as re-integrated from the segment's first state
#pragma unroll 1
                for (int nn = s0; nn < s1; ++nn) {
                    const double tn = *ck_time(nn), tn1 = *ck_time(nn + 1);
                    const double xn = (!sparse || nn == s0) ? ck_state(nn)[tid] : xcarry;
                    const double dts = tn1 - tn;
                    double *blk = slab + (int64_t)(nn - s0) * BLK;
                    if (tid == 0) blk[0] = tn;
                    double Kr[SADJ];
#pragma unroll
                    for (int m = 0; m < SADJ; ++m) {
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < m; ++j)
                            if (Tab::a(m, j) != 0.0) acc = fma(Tab::a(m, j), Kr[j], acc);
                        const double xm = m == 0 ? xn : fma(dts, acc, xn);
                        P.xin()[tid] = xm;
                        blk[OFF_X + m * N + tid] = xm;
                        __syncthreads();
                        const double gm = product_rows(P, rr, creg, tid);
                        Kr[m] = xm * gm;
                        blk[OFF_G + m * N + tid] = gm;
                    }
                    if (sparse) { // x_{nn+1} = x_nn + dt sum_j b_j K_j, the forward sweep's own update
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < SADJ; ++j)
                            if (Tab::b(j) != 0.0) acc = fma(Tab::b(j), Kr[j], acc);
                        xcarry = fma(dts, acc, xn);
                    }
                }
                __syncthreads();
            }
            // ---- phase 2: state adjoint (transposed products with the same row layout) ----
#pragma unroll 1
            for (int step = s1 - 1; step >= s0; --step) {
                double *blk = slab + (int64_t)(step - s0) * BLK;
                const double t_lo = blk[0];
                const double dt_s = t_hi - t_lo;
                t_hi = t_lo;
                double Xr[SADJ], Gr[SADJ], W[SADJ + 1];
#pragma unroll
                for (int m = 0; m < SADJ; ++m) {
                    Xr[m] = blk[OFF_X + m * N + tid];
                    Gr[m] = blk[OFF_G + m * N + tid];
                }
                W[0] = lam;
#pragma unroll
                for (int m = 1; m <= SADJ; ++m) W[m] = (Tab::b(m - 1) * dt_s) * lam;
#pragma unroll
                for (int m = SADJ; m >= 1; --m) {
                    const double v = W[m] * Xr[m - 1];
                    P.xin()[tid] = v;
                    blk[OFF_V + (m - 1) * N + tid] = v;
                    rbar += v;
                    __syncthreads();
                    const double atv = product_cols(P, creg, p3buf, tid);
                    const double gx = fma(W[m], Gr[m - 1], atv);
                    W[0] += gx;
#pragma unroll
                    for (int k = 1; k < m; ++k)
                        if (Tab::a(m - 1, k - 1) != 0.0) W[k] = fma(gx * Tab::a(m - 1, k - 1), dt_s, W[k]);
                }
                lam = W[0];
            }
            // every v block of this CTA's slab is written (generic proxy); the bulk copies below read them through the async proxy
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();
            if (SEG) {
                // phase 3 needs every register for its accumulators: park the three values that outlive it (each thread reads back
                // what it wrote; xs and yp are idle until the next product, which comes after the reload)
                xs[tid] = lam;
                xs[N + tid] = rbar;
                yp[tid] = t_hi;
            }

            // ---- phase 3: Abar = sum_k v_k X_k^T; this CTA takes columns [128 rank, 128 rank + 128) in two 64-column passes ----
            // thread (ty, tx): rows 8 ty + r, columns cb + 16 c + 2 tx + e  (r < 8, c < 4, e < 2). The operands of a step
            // (v_0..v_{s-1}: 2 KB each, X_0..X_{s-1}: 512 B of the pass's columns each) are brought in by TMA bulk copies,
            // P3_STAGES steps ahead.
            const int ty = tid >> 3, tx = tid & 7;
            auto p3_issue = [&](int step, int cb, uint32_t qi) { // thread 0: request the operands of `step` into stage qi % P3_STAGES
                const int st = (int)(qi % P3_STAGES);
                double *dstV = p3buf + (size_t)st * P3_STAGE_DOUBLES, *dstX = dstV + SADJ * N;
                const double *blk = slab + (int64_t)step * BLK;
                mbar_expect_tx(bar + 3 + st, (uint32_t)(SADJ * (N + PCOLS) * 8));
#pragma unroll 1
                for (int m = 0; m < SADJ; ++m) {
                    bulk_g2s(dstV + m * N, blk + OFF_V + m * N, N * 8, bar + 3 + st, pol_slab);
                    bulk_g2s(dstX + m * PCOLS, blk + OFF_X + m * N + cb, PCOLS * 8, bar + 3 + st, pol_slab);
                }
            };
#pragma unroll 1
            for (int cb = (int)rank * HR; cb < (int)rank * HR + HR && cb < n; cb += PCOLS) {
                double acc[8][8];
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;
                if (tid == 0) {
#pragma unroll 1
                    for (int k = 0; k < P3_STAGES && k < Tseg; ++k) p3_issue(k, cb, p3q + (uint32_t)k);
                }
#pragma unroll 1
                for (int step = 0; step < Tseg; ++step) {
                    const int st = (int)(p3q % P3_STAGES);
                    mbar_wait_or_trap(bar + 3 + st, (p3q / P3_STAGES) & 1u);
                    const double *Vs = p3buf + (size_t)st * P3_STAGE_DOUBLES, *Xs = Vs + SADJ * N;
#pragma unroll
                    for (int m = 0; m < SADJ; ++m) {
                        double v8[8], x8[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double2 tv = *reinterpret_cast<const double2 *>(Vs + m * N + 8 * ty + 2 * q);
                            v8[2 * q] = tv.x;
                            v8[2 * q + 1] = tv.y;
                            const double2 tx2 = *reinterpret_cast<const double2 *>(Xs + m * PCOLS + 16 * q + 2 * tx);
                            x8[2 * q] = tx2.x;
                            x8[2 * q + 1] = tx2.y;
                        }
#pragma unroll
                        for (int r = 0; r < 8; ++r)
#pragma unroll
                            for (int c = 0; c < 8; ++c) acc[r][c] = fma(v8[r], x8[c], acc[r][c]);
                    }
                    __syncthreads(); // every thread is done with this stage
                    if (tid == 0 && step + P3_STAGES < Tseg) p3_issue(step + P3_STAGES, cb, p3q + (uint32_t)P3_STAGES);
                    ++p3q;
                }
                const bool ow = overwrite && first_seg;
                // Abar tile out: plain stores on first use of the row, fire-and-forget reductions afterwards (one writer per
                // address, program order: deterministic)
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int row = 8 * ty + r, col = cb + 16 * q + 2 * tx;
                        double *dst = gbar + n + (size_t)row * n + col;
                        if (EXACT) {
                            if (ow) {
                                *reinterpret_cast<double2 *>(dst) = make_double2(acc[r][2 * q], acc[r][2 * q + 1]);
                            } else {
                                atomicAdd(dst, acc[r][2 * q]);
                                atomicAdd(dst + 1, acc[r][2 * q + 1]);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                if (row < n && col + e < n) {
                                    if (ow) dst[e] = acc[r][2 * q + e];
                                    else atomicAdd(dst + e, acc[r][2 * q + e]);
                                }
                        }
                    }
            }
            __syncthreads(); // the slab is free for the next segment
            if (SEG) {
                lam = xs[tid];
                rbar = xs[N + tid];
                t_hi = yp[tid];
            }
            first_seg = false;
            s1 = s0;
            } while (s1 > 0);
            if (rank == 0 && live) {
                lam_io[tid] = lam;
                if (overwrite) gbar[tid] = rbar;
                else atomicAdd(gbar + tid, rbar);
            }
            row_init = true;
        }
        __syncthreads(); // every reader of the shared-memory rows is done before the next matrix lands
    }
    if (a.reduce == VA_REDUCE_SUM && a.n_out > 0 && !row_init && rank == 0)
        for (int k = tid; k < npar; k += NT) a.partial[pair_id * npar + k] = 0.0; // every trajectory of this pair failed
    cluster_barrier(); // neither CTA leaves while the other could still address its shared memory
}

template <class Tab, bool ADAPTIVE, int CL>
size_t smem_bytes()
{
    return (size_t)Pair<CL>::SR * N * 8 + (size_t)((5 + 2 * CL) * N + P3_STAGES * P3_STAGE_DOUBLES + 8) * 8 + (3 + P3_STAGES) * 8 + 64;
}

// launch configuration with the cluster dimension attribute (the kernel carries no compile-time cluster size)
template <class Tab, bool ADAPTIVE, int CL>
cudaError_t configure(cudaLaunchConfig_t &cfg, cudaLaunchAttribute &at, int grid, cudaStream_t st)
{
    const size_t smem = smem_bytes<Tab, ADAPTIVE, CL>();
    cudaError_t e = cudaFuncSetAttribute(k_glv_pair<Tab, ADAPTIVE, CL, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_glv_pair<Tab, ADAPTIVE, CL, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_glv_pair<Tab, ADAPTIVE, CL, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_glv_pair<Tab, ADAPTIVE, CL, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = CL;
    at.val.clusterDim.y = 1;
    at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    return cudaSuccess;
}
template <class Tab, bool ADAPTIVE, int CL>
cudaError_t launch2(const VaGlvWideArgs &a, cudaStream_t st)
{
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at;
    cudaError_t e = configure<Tab, ADAPTIVE, CL>(cfg, at, a.grid, st);
    if (e != cudaSuccess) return e;
    if (a.recompute) {
        if (!a.xstore || a.seg_len < 1) return cudaErrorInvalidValue;
        return a.n == N ? cudaLaunchKernelEx(&cfg, k_glv_pair<Tab, ADAPTIVE, CL, true, true>, a)
                        : cudaLaunchKernelEx(&cfg, k_glv_pair<Tab, ADAPTIVE, CL, false, true>, a);
    }
    return a.n == N ? cudaLaunchKernelEx(&cfg, k_glv_pair<Tab, ADAPTIVE, CL, true, false>, a)
                    : cudaLaunchKernelEx(&cfg, k_glv_pair<Tab, ADAPTIVE, CL, false, false>, a);
}
template <class Tab, bool ADAPTIVE>
cudaError_t launch(const VaGlvWideArgs &a, cudaStream_t st)
{
    return a.cluster == 4 ? launch2<Tab, ADAPTIVE, 4>(a, st) : launch2<Tab, ADAPTIVE, 2>(a, st);
}
template <class Tab, bool ADAPTIVE, int CL>
cudaError_t max_clusters2(int sm_count, int *n)
{
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at;
    cudaError_t e = configure<Tab, ADAPTIVE, CL>(cfg, at, sm_count / CL * CL, nullptr);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveClusters(n, k_glv_pair<Tab, ADAPTIVE, CL, true, false>, &cfg);
}
template <class Tab, bool ADAPTIVE>
cudaError_t max_clusters(int cl, int sm_count, int *n)
{
    return cl == 4 ? max_clusters2<Tab, ADAPTIVE, 4>(sm_count, n) : max_clusters2<Tab, ADAPTIVE, 2>(sm_count, n);
}

} // namespace

bool va_glv_pair_supported(int n, int stepper, int adaptive)
{
    if (n <= 64 || n > N) return false; // <= 64 species: the register kernels (va_glv_t8.cu, va_glv_wide.cu)
    if (stepper == VA_RK_RK4) return !adaptive;
    if (stepper == VA_RK_CK54 || stepper == VA_RK_DOPRI5) return adaptive != 0;
    return false;
}

int va_glv_pair_block_doubles(int stepper)
{
    const int sadj = stepper == VA_RK_RK4 ? TabRK4::SADJ : stepper == VA_RK_CK54 ? TabCK54::SADJ : TabDOPRI5::SADJ;
    return 8 + 3 * sadj * N;
}

// co-resident clusters of `cluster` CTAs the device can hold (the persistent grid is cluster x this)
cudaError_t va_glv_pair_max_clusters(int stepper, int cluster, int sm_count, int *n)
{
    switch (stepper) {
    case VA_RK_RK4: return max_clusters<TabRK4, false>(cluster, sm_count, n);
    case VA_RK_CK54: return max_clusters<TabCK54, true>(cluster, sm_count, n);
    case VA_RK_DOPRI5: return max_clusters<TabDOPRI5, true>(cluster, sm_count, n);
    }
    return cudaErrorInvalidValue;
}

// a.cluster = CTAs per trajectory (2 or 4), a.grid a multiple of it: CTAs [c k, c k + c) form cluster k (one slab per CTA,
// partial-sum row k)
cudaError_t va_glv_pair_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    if ((a.cluster != 2 && a.cluster != 4) || a.grid % a.cluster) return cudaErrorInvalidValue;
    switch (a.stepper) {
    case VA_RK_RK4: return launch<TabRK4, false>(a, st);
    case VA_RK_CK54: return launch<TabCK54, true>(a, st);
    case VA_RK_DOPRI5: return launch<TabDOPRI5, true>(a, st);
    }
    return cudaErrorInvalidValue;
}
