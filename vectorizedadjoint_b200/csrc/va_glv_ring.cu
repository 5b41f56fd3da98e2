// va_glv_ring.cu -- Generalized Lotka-Volterra with 256 species (BASELINE config 5): one persistent 256-thread CTA per
// SM integrates trajectory after trajectory; thread i owns component i of every vector (state, stage slopes, stage
// adjoints live in registers), and the 512 KB interaction matrix -- too large for one SM -- is streamed through a
// shared-memory ring of 32 KB row chunks filled by TMA bulk copies (cp.async.bulk + mbarrier), RING chunks in flight per
// SM. The last CR rows of the matrix stay in registers for a whole sweep (25 % of the matrix at CR = 64), so each product
// streams (N - CR) N 8 bytes.
//
// Same algorithm as the other GLV families (reference lib/include/detail/runge_kutta.hpp:76-118 forward sweep with
// odeint's controlled stepper, detail/backpropagation.hpp:83-158, 231-254 reverse sweep), three phases per trajectory:
//   1. forward sweep: g = r + A X per stage (row dot products: a warp takes two rows of a chunk, lanes stride over the
//      columns, transposing shuffle reduction); every accepted step leaves a block
//      [8-double header (t_n) | X_0..X_{s-1} | g_0..g_{s-1} | v_0..v_{s-1}] in the CTA's slab (store-stages policy);
//   2. state adjoint: per stage v_m = w_m o X_{m-1}, A^T v_m (thread j owns column j of every chunk row: no reduction),
//      stage-adjoint recurrences; v_m goes to the slab, rbar accumulates in a register;
//   3. gradient accumulation Abar = sum over steps and stages of v_m X_{m-1}^T: a [256 x 6T] x [6T x 256] matrix product
//      with an 8 x 8 accumulator tile per thread, four 64-column passes, operands staged through shared memory. Unlike the
//      streamed family (va_glv_stream.cu) Abar is never read-modify-written per stage: it is written once.
// Bytes streamed per accepted step: 12 products x (N - CR) N 8 B = 4.7 MB at CR = 64 (va_glv_stream.cu: 12.6 MB); the
// kernel is bound by the L2 -> SM path (matrices of all resident CTAs: 148 x 512 KB = 76 MB, L2-resident under an
// evict_last policy) or by HBM when they spill.
#include "va_glv_common.cuh"
#include "va_tma.cuh"

#ifndef VA_RING_STAGES
#define VA_RING_STAGES 5
#endif

namespace {
using namespace va_tma;

constexpr int NT = 256;                    // threads per CTA
constexpr int N = 256;                     // species: thread i owns component i
constexpr int CH_ROWS = 16;                // matrix rows per ring chunk
constexpr int CH_DOUBLES = CH_ROWS * N;    // 4096 doubles = 32 KB
constexpr uint32_t CH_BYTES = CH_DOUBLES * 8;
constexpr int RING = VA_RING_STAGES;       // chunks in flight per SM
constexpr int PCOLS = 64;                  // columns of Abar per accumulation pass
static_assert(N == NT, "thread i owns component i");

struct Ring {
    double *buf;     // [RING][CH_DOUBLES]
    uint64_t *full;  // [RING] "chunk landed" barriers
    const double *A; // first streamed row of the current matrix
    uint64_t pol;    // L2 policy of the matrix stream
    int stage;       // stage of the next chunk to consume
    uint32_t parity; // its barrier phase
    int cnext;       // next chunk index to request (thread 0 only)
};

template <int NCH>
__device__ __forceinline__ void ring_issue(Ring &R, int stage)
{
    mbar_expect_tx(&R.full[stage], CH_BYTES);
    bulk_g2s(R.buf + (size_t)stage * CH_DOUBLES, R.A + (size_t)R.cnext * CH_DOUBLES, CH_BYTES, &R.full[stage], R.pol);
    R.cnext = (R.cnext + 1 == NCH) ? 0 : R.cnext + 1;
}
// nothing in flight -> RING chunks in flight, starting with chunk 0 of matrix A
template <int NCH>
__device__ __forceinline__ void ring_prime(Ring &R, const double *A, int tid)
{
    R.A = A;
    if (tid == 0) {
        R.cnext = 0;
        int s = R.stage;
#pragma unroll 1
        for (int k = 0; k < RING; ++k) {
            ring_issue<NCH>(R, s);
            s = (s + 1 == RING) ? 0 : s + 1;
        }
    }
}
__device__ __forceinline__ const double *ring_acquire(Ring &R)
{
    mbar_wait(&R.full[R.stage], R.parity);
    return R.buf + (size_t)R.stage * CH_DOUBLES;
}
// every thread is done with the current chunk: refill its stage with the chunk RING positions ahead
template <int NCH>
__device__ __forceinline__ void ring_release(Ring &R, int tid)
{
    __syncthreads();
    if (tid == 0) ring_issue<NCH>(R, R.stage);
    if (++R.stage == RING) { R.stage = 0; R.parity ^= 1u; }
}
// consume and discard what is in flight (end of a sweep: the ring memory is reused / the matrix changes)
__device__ __forceinline__ void ring_drain(Ring &R)
{
#pragma unroll 1
    for (int k = 0; k < RING; ++k) {
        mbar_wait(&R.full[R.stage], R.parity);
        if (++R.stage == RING) { R.stage = 0; R.parity ^= 1u; }
    }
    __syncthreads();
}

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// g_out[i] = r_i + (A xin)_i for all rows. Row dot products: lane l holds columns {64k + 2l, 64k + 2l + 1}, k < 4
// (conflict-free LDS.128 of a chunk row); the cached rows (row layout: warp w holds rows N-CR + w CR/8 + r) come first,
// under the latency of the first chunk. xin must be visible to all threads; g_out is visible to all on return.
template <int CR, int NCH>
__device__ __forceinline__ void matvec_rows(Ring &R, const double *xin, const double *rr, double *g_out, const double (&creg)[CR ? CR : 1],
                                            int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    double xr[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double2 t = *reinterpret_cast<const double2 *>(xin + 64 * k + 2 * lane);
        xr[2 * k] = t.x;
        xr[2 * k + 1] = t.y;
    }
    if (CR > 0) {
        constexpr int RW = CR / 8; // cached rows per warp (8 at CR = 64)
        static_assert(CR == 0 || RW == 8, "cached-row reduction is written for 8 rows per warp");
        double s[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = fma(creg[(r * 8 + k) % (CR ? CR : 1)], xr[k], acc);
            s[r] = acc;
        }
        // transposing butterfly: 8 row sums over 32 lanes with 4 + 2 + 1 + 1 + 1 exchanges
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
        {
            const double send = b2 ? s[0] : s[1], keep = b2 ? s[1] : s[0];
            s[0] = keep + shx(send, 4);
        }
        s[0] += shx(s[0], 2);
        s[0] += shx(s[0], 1);
        if ((lane & 3) == 0) {
            const int row = N - CR + warp * RW + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            g_out[row] = rr[row] + s[0];
        }
    }
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
        const double *rowp = ring_acquire(R) + (2 * warp) * N + 2 * lane;
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double2 a0 = *reinterpret_cast<const double2 *>(rowp + 64 * k);
            const double2 a1 = *reinterpret_cast<const double2 *>(rowp + N + 64 * k);
            acc0 = fma(a0.x, xr[2 * k], acc0);
            acc1 = fma(a1.x, xr[2 * k], acc1);
            acc0 = fma(a0.y, xr[2 * k + 1], acc0);
            acc1 = fma(a1.y, xr[2 * k + 1], acc1);
        }
        const bool hi = lane & 16;
        double keep = (hi ? acc1 : acc0) + shx(hi ? acc0 : acc1, 16);
        keep += shx(keep, 8);
        keep += shx(keep, 4);
        keep += shx(keep, 2);
        keep += shx(keep, 1);
        if ((lane & 15) == 0) {
            const int row = c * CH_ROWS + 2 * warp + (lane >> 4);
            g_out[row] = rr[row] + keep;
        }
        ring_release<NCH>(R, tid);
    }
    if (NCH == 0) __syncthreads();
}

// (A^T v)_j for j = tid: thread j owns column j of every row (column layout of the cached rows: creg[i] = A[N-CR+i][j]).
// v must be visible to all threads.
template <int CR, int NCH>
__device__ __forceinline__ double matvec_cols(Ring &R, const double *v, const double (&creg)[CR ? CR : 1], int tid)
{
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (CR > 0) {
#pragma unroll
        for (int i = 0; i < CR; i += 2) {
            const double2 t = *reinterpret_cast<const double2 *>(v + N - CR + i);
            acc[i & 3] = fma(creg[i % (CR ? CR : 1)], t.x, acc[i & 3]);
            acc[(i + 1) & 3] = fma(creg[(i + 1) % (CR ? CR : 1)], t.y, acc[(i + 1) & 3]);
        }
    }
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
        const double *col = ring_acquire(R) + tid;
        const double *vc = v + c * CH_ROWS;
#pragma unroll
        for (int i = 0; i < CH_ROWS; i += 2) {
            const double2 t = *reinterpret_cast<const double2 *>(vc + i);
            acc[i & 3] = fma(col[i * N], t.x, acc[i & 3]);
            acc[(i + 1) & 3] = fma(col[(i + 1) * N], t.y, acc[(i + 1) & 3]);
        }
        ring_release<NCH>(R, tid);
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

template <class Tab, bool ADAPTIVE, int CR>
__global__ void __launch_bounds__(NT, 1) k_glv_ring(const __grid_constant__ VaGlvWideArgs a)
{
    constexpr int S = Tab::S, SADJ = Tab::SADJ;
    constexpr int SE = Tab::FSAL ? S - 1 : S;
    constexpr int NCH = (N - CR) / CH_ROWS;
    constexpr int BLK = 8 + 3 * SADJ * N; // [header | X_0.. | g_0.. | v_0..]
    constexpr int OFF_X = 8, OFF_G = 8 + SADJ * N, OFF_V = 8 + 2 * SADJ * N;
    static_assert((N - CR) % CH_ROWS == 0 && NCH >= 1, "streamed part must be whole chunks");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ringbuf = reinterpret_cast<double *>(smem_raw);
    double *xs = ringbuf + (size_t)RING * CH_DOUBLES; // stage state / seed vector handed to a product
    double *gout = xs + N;                             // product result
    double *rr = gout + N;                             // growth rates r
    double *red = rr + N;                              // [8] error-norm partials
    uint64_t *bars = reinterpret_cast<uint64_t *>(red + 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int npar = N * N + N;

    Ring R;
    R.buf = ringbuf;
    R.full = bars;
    R.stage = 0;
    R.parity = 0;
    R.cnext = 0;
    R.A = nullptr;
    R.pol = (a.flags & 2) ? policy_evict_last() : policy_evict_normal(); // keep the matrices of the resident CTAs in L2
    if (tid == 0) {
#pragma unroll 1
        for (int s = 0; s < RING; ++s) mbar_init(&bars[s], 1);
        mbar_init_fence();
    }
    __syncthreads();
    double *const slab = a.slab + (int64_t)blockIdx.x * a.slab_stride;
    double creg[CR ? CR : 1];
    creg[0] = 0.0;
    bool row_init = false; // summed mode: this CTA's partial-sum row has been written

    for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
        const double *pb = a.params + b * npar;
        const double *A = pb + N;
        const double *Ac = A + (size_t)(N - CR) * N; // cached rows
        // ------------------------------------------ forward sweep ------------------------------------------------------
        ring_prime<NCH>(R, A, tid);
        if (CR > 0) { // row layout: 8 rows per warp, 8 columns per lane and row
#pragma unroll
            for (int r = 0; r < CR / 8; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 t = __ldg(reinterpret_cast<const double2 *>(Ac + (size_t)(warp * (CR / 8) + r) * N + 64 * k + 2 * lane));
                    creg[(r * 8 + 2 * k) % (CR ? CR : 1)] = t.x;
                    creg[(r * 8 + 2 * k + 1) % (CR ? CR : 1)] = t.y;
                }
        }
        rr[tid] = __ldg(pb + tid);
        double x = a.x0[b * N + tid];
        xs[tid] = x;
        __syncthreads();
        double t = a.ti, dt = a.dt0;
        const double tf = a.tf;
        int nck = 0, rejects = 0, status = 0, trials = 0;
        double K[S];
        matvec_rows<CR, NCH>(R, xs, rr, gout, creg, tid);
        double g0 = gout[tid];
        K[0] = x * g0;
        bool active = ADAPTIVE ? va_less_with_sign(t, tf, dt) : va_less_eq_with_sign(t + dt, tf, dt);
        bool fresh = true;
        while (active) {
            double *blk = slab + (int64_t)nck * BLK;
            if (fresh) {
                if (nck >= a.cap) { status |= VA_TRAJ_CKPT_OVERFLOW; break; }
                blk[OFF_X + tid] = x;
                blk[OFF_G + tid] = g0;
                if (tid == 0) blk[0] = t;
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
                xs[tid] = xm;
                if (m < SADJ) blk[OFF_X + m * N + tid] = xm;
                __syncthreads();
                matvec_rows<CR, NCH>(R, xs, rr, gout, creg, tid);
                const double gm = gout[tid];
                K[m] = xm * gm;
                if (m < SADJ) blk[OFF_G + m * N + tid] = gm;
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
                xs[tid] = xnew;
                __syncthreads();
                matvec_rows<CR, NCH>(R, xs, rr, gout, creg, tid);
                gnew = gout[tid];
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
                        constexpr int P = Tab::STEPPER_ORDER;
                        double floor_ = 1.0;
#pragma unroll
                        for (int k = 0; k < P; ++k) floor_ *= 0.2;
                        dt *= (err <= floor_) ? 4.5 : 9.0 / 10.0 * inv_root<P>(err);
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
                    xs[tid] = x;
                    __syncthreads();
                    matvec_rows<CR, NCH>(R, xs, rr, gout, creg, tid);
                    g0 = gout[tid];
                    K[0] = x * g0;
                }
            }
        }
        ring_drain(R);
        const int T = nck;
        if (tid == 0) slab[(int64_t)T * BLK] = t;
        if (!isfinite(x)) status |= VA_TRAJ_NONFINITE;
        status = __syncthreads_or(status);
        const bool failed = status & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS);
        a.x_final[b * N + tid] = failed ? nan("") : x;
        if (tid == 0) {
            if (a.n_accept) a.n_accept[b] = T;
            if (a.n_reject) a.n_reject[b] = rejects;
            if (a.status) a.status[b] = status;
        }
        const double t_final = t;

        // ------------------------------------------ reverse sweep ------------------------------------------------------
        for (int o = 0; o < a.n_out; ++o) {
            double *lam_io = a.lambda + (b * a.n_out + o) * N;
            const bool sum_mode = a.reduce == VA_REDUCE_SUM;
            double *gbar = sum_mode ? a.partial + (int64_t)blockIdx.x * npar : a.mu + (b * a.n_out + o) * npar;
            const bool overwrite = !sum_mode || !row_init; // first use of this accumulator row
            if (failed) {
                lam_io[tid] = nan("");
                if (!sum_mode)
                    for (int k = tid; k < npar; k += NT) gbar[k] = nan("");
                continue;
            }
            // ---- phase 2: state adjoint ----
            ring_prime<NCH>(R, A, tid);
            if (CR > 0) { // column layout
#pragma unroll
                for (int i = 0; i < CR; ++i) creg[i % (CR ? CR : 1)] = __ldg(Ac + (size_t)i * N + tid);
            }
            double lam = a.objective == VA_OBJ_SUM ? 1.0 : a.objective == VA_OBJ_HALF_NORM2 ? x : lam_io[tid];
            double rbar = 0.0;
            double t_hi = t_final;
#pragma unroll 1
            for (int step = T - 1; step >= 0; --step) {
                double *blk = slab + (int64_t)step * BLK;
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
                    xs[tid] = v;
                    blk[OFF_V + (m - 1) * N + tid] = v;
                    rbar += v;
                    __syncthreads();
                    const double atv = matvec_cols<CR, NCH>(R, xs, creg, tid);
                    const double gx = fma(W[m], Gr[m - 1], atv);
                    W[0] += gx;
#pragma unroll
                    for (int k = 1; k < m; ++k)
                        if (Tab::a(m - 1, k - 1) != 0.0) W[k] = fma(gx * Tab::a(m - 1, k - 1), dt_s, W[k]);
                }
                lam = W[0];
            }
            lam_io[tid] = lam;
            if (overwrite) gbar[tid] = rbar;
            else gbar[tid] += rbar;
            ring_drain(R); // ends with a CTA barrier: every v block is written, the ring memory is free

            // ---- phase 3: Abar = sum_k v_k X_k^T, 64 columns per pass, 8 x 8 accumulators per thread ----
            // thread (ty, tx): rows 8 ty + r, columns cb + 16 c + 2 tx + e  (r < 8, c < 4, e < 2)
            double *Vs = ringbuf;            // [SADJ][N]
            double *Xs = ringbuf + SADJ * N; // [SADJ][PCOLS]
            const int ty = tid >> 3, tx = tid & 7;
#pragma unroll 1
            for (int cb = 0; cb < N; cb += PCOLS) {
                double acc[8][8];
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;
                double vreg[SADJ], xreg[SADJ];
#pragma unroll
                for (int m = 0; m < SADJ; ++m) { vreg[m] = 0.0; xreg[m] = 0.0; }
                if (T > 0) {
#pragma unroll
                    for (int m = 0; m < SADJ; ++m) {
                        vreg[m] = slab[OFF_V + m * N + tid];
                        if (tid < PCOLS) xreg[m] = slab[OFF_X + m * N + cb + tid];
                    }
                }
#pragma unroll 1
                for (int step = 0; step < T; ++step) {
                    __syncthreads(); // the previous step's operands have been consumed
#pragma unroll
                    for (int m = 0; m < SADJ; ++m) {
                        Vs[m * N + tid] = vreg[m];
                        if (tid < PCOLS) Xs[m * PCOLS + tid] = xreg[m];
                    }
                    __syncthreads();
                    if (step + 1 < T) {
                        const double *nb = slab + (int64_t)(step + 1) * BLK;
#pragma unroll
                        for (int m = 0; m < SADJ; ++m) {
                            vreg[m] = nb[OFF_V + m * N + tid];
                            if (tid < PCOLS) xreg[m] = nb[OFF_X + m * N + cb + tid];
                        }
                    }
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
                }
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        double2 *dst = reinterpret_cast<double2 *>(gbar + N + (size_t)(8 * ty + r) * N + cb + 16 * q + 2 * tx);
                        double2 o2 = make_double2(acc[r][2 * q], acc[r][2 * q + 1]);
                        if (!overwrite) {
                            const double2 old = *dst;
                            o2.x += old.x;
                            o2.y += old.y;
                        }
                        *dst = o2;
                    }
            }
            row_init = true;
            __syncthreads(); // the ring memory goes back to the matrix stream
        }
    }
    if (a.reduce == VA_REDUCE_SUM && a.n_out > 0 && !row_init)
        for (int k = tid; k < npar; k += NT) a.partial[(int64_t)blockIdx.x * npar + k] = 0.0; // every trajectory of this CTA failed
}

template <class Tab, bool ADAPTIVE, int CR>
cudaError_t launch2(const VaGlvWideArgs &a, cudaStream_t st, size_t smem)
{
    cudaError_t e = cudaFuncSetAttribute(k_glv_ring<Tab, ADAPTIVE, CR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_glv_ring<Tab, ADAPTIVE, CR><<<a.grid, NT, smem, st>>>(a);
    return cudaGetLastError();
}
template <class Tab, bool ADAPTIVE>
cudaError_t launch(const VaGlvWideArgs &a, cudaStream_t st, size_t smem)
{
    return (a.flags & 4) ? launch2<Tab, ADAPTIVE, 0>(a, st, smem) : launch2<Tab, ADAPTIVE, 64>(a, st, smem);
}

} // namespace

size_t va_glv_ring_smem() { return (size_t)RING * CH_BYTES + (size_t)(3 * N + 8) * 8 + RING * 8 + 64; }

bool va_glv_ring_supported(int n, int stepper, int adaptive)
{
    if (n != N) return false;
    if (stepper == VA_RK_RK4) return !adaptive;
    if (stepper == VA_RK_CK54 || stepper == VA_RK_DOPRI5) return adaptive != 0;
    return false;
}

int va_glv_ring_block_doubles(int stepper)
{
    const int sadj = stepper == VA_RK_RK4 ? TabRK4::SADJ : stepper == VA_RK_CK54 ? TabCK54::SADJ : TabDOPRI5::SADJ;
    return 8 + 3 * sadj * N;
}

// a.flags: bit 1 = evict_last policy on the matrix stream, bit 2 = no register-cached rows (CR = 0)
cudaError_t va_glv_ring_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const size_t smem = va_glv_ring_smem();
    switch (a.stepper) {
    case VA_RK_RK4: return launch<TabRK4, false>(a, st, smem);
    case VA_RK_CK54: return launch<TabCK54, true>(a, st, smem);
    case VA_RK_DOPRI5: return launch<TabDOPRI5, true>(a, st, smem);
    }
    return cudaErrorInvalidValue;
}
