// va_glv_oct.cu -- Generalized Lotka-Volterra with up to 16 species (BASELINE config 3: N = 16, one million parameter sets),
// second generation: EIGHT lanes per trajectory, four trajectories per warp, the matrix A AND the gradient accumulator Abar
// resident in registers for the whole life of a trajectory, recompute checkpoint policy. No block-level barrier anywhere.
//
// Why (ncu, profiles/r02/glv16_quad_ncu_full_before.txt): the four-lane kernel (va_glv_quad.cu) keeps ONE 16 x 16 matrix per
// quad in registers (128 registers per lane), so the reverse sweep is split into phases that swap A^T and Abar through the same
// registers and hand the stage states X_m, slopes g_m and seeds v_m from phase to phase through a 32 KB slab per trajectory.
// 9472 quads in flight x 32 KB = 300 MB do not fit L2: 108 KB of DRAM traffic per trajectory against 2.6 KB of compulsory
// bytes, 4.5 TB/s = 69 % of HBM bandwidth -- the kernel is bound by its own checkpoints, the FP64 pipe is 35 % busy.
// Here a lane holds a 4 x 8 tile of A and the same tile of Abar (64 + 64 registers), so that
//   * the forward sweep, the stage recompute of the reverse sweep and the transposed product A^T v all use the ONE resident
//     copy of A (the transposed product is a row combination of the same tile, reduce-scattered over the four lanes that share
//     a column block), and the outer products Abar += v_m X_{m-1}^T are accumulated in the same pass: nothing is handed
//     between phases, so nothing but (t_n, x_n) is stored -- the reference's own policy (lib/include/StateStorage.hpp:7-8,
//     detail/backpropagation.hpp:24-64: stages recomputed from the stored state with dt = t_{n+1} - t_n);
//   * checkpoints are 192 B per accepted step (4 KB per trajectory, 19 MB for all 4736 trajectories in flight: L2-resident);
//     DRAM traffic is the compulsory parameter read and result write;
//   * price: six more matrix-vector products per accepted step (24 instead of 18 product-equivalents).
// Thread <-> data map (lane o = 2 q + h inside the trajectory, q = 0..3, h = 0..1):
//   owned components          8h + 2q, 8h + 2q + 1           (every vector recurrence is carried for these two only)
//   tile rows    kk = 0,1:    the owned components           kk = 2,3: the h-partner's components 8(1-h) + 2q, +1
//   tile columns cc = 0..7:   8h + (cc ^ 2q)                 (held permuted so that both reductions are select-free)
//   y = A x:     32 DFMAs over the lane's 8 columns, then y_own[k] = y[k] + shfl_xor(y[k + 2], 1)
//   z = A^T v:   32 DFMAs over the lane's 4 rows, then a reduce-scatter over q: 4 + 2 shuffles leave z_own in registers 0, 1
//   Abar += v x^T: 32 DFMAs, operands v (own + partner's by shuffle) and the same 8 x entries the row product loads.
// Same algorithm as the other GLV kernels (reference lib/include/detail/runge_kutta.hpp:76-118 forward sweep with odeint's
// controlled stepper; detail/backpropagation.hpp:83-158, 231-254 reverse sweep). The four trajectories of a warp run in lock
// step to the longest of the four; lanes of finished trajectories compute with zero seeds and discard.
#include "va_glv_common.cuh"

#ifndef VA_OCT_WARPS
#define VA_OCT_WARPS 8
#endif

namespace {

constexpr int NP = 16;                 // padded species count
constexpr int HDR = 8;                 // doubles in a checkpoint header (hdr[0] = t_n)
constexpr int BLK = HDR + NP;          // doubles per checkpoint: [8-double header | x_n]
constexpr int NT = 32 * VA_OCT_WARPS;  // threads per CTA
constexpr int TPC = NT / 8;            // trajectories per CTA
constexpr unsigned FULL = 0xffffffffu;
constexpr int NBUF = 8;                // per-trajectory stage-state buffers in shared memory (one per stage, <= 7 stages)
constexpr int TSTRIDE = NBUF * NP;     // doubles of shared memory per trajectory

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(FULL, v, m); }

template <class Tab, bool ADAPTIVE, bool EXACT>
__global__ void __launch_bounds__(NT, 1) k_glv_oct(const __grid_constant__ VaGlvWideArgs a)
{
    constexpr int S = Tab::S, SADJ = Tab::SADJ;
    constexpr int SE = Tab::FSAL ? S - 1 : S;
    static_assert(S <= NBUF, "stage buffers");
    extern __shared__ __align__(128) double sm_all[];
    const int lane = threadIdx.x & 31, o = lane & 7, h = o & 1, q = o >> 1;
    const int tc = threadIdx.x >> 3; // trajectory slot inside the CTA
    double *const xs = sm_all + (size_t)tc * TSTRIDE;
    const int n = EXACT ? NP : a.n;
    const int npar = n * n + n;
    const int64_t slot = (int64_t)blockIdx.x * TPC + tc, nslots = (int64_t)gridDim.x * TPC;
    double *const slab = a.slab + slot * a.slab_stride;
    double *const part = a.partial + slot * npar; // summed mode: this slot's partial-sum row
    const double tf = a.tf;
    const int own0 = 8 * h + 2 * q; // first owned component (the second is own0 + 1)
    const bool live[2] = {own0 < n, own0 + 1 < n};
    int rowi[4], coli[8];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) rowi[kk] = (kk < 2 ? 8 * h : 8 * (1 - h)) + 2 * q + (kk & 1);
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) coli[cc] = 8 * h + (cc ^ (2 * q));
    bool row_init = false;

    // own pair -> stage buffer m; one warp barrier, then every lane may read the whole 16-vector
    auto put = [&](int m, const double(&v)[2]) {
        *reinterpret_cast<double2 *>(xs + m * NP + own0) = make_double2(v[0], v[1]);
        __syncwarp();
    };
    // the 8 entries of stage vector m under the lane's columns: 4 LDS.128; the 8 lanes of a trajectory read its 128 bytes
    // exactly once per instruction (conflict-free, one quarter-warp per trajectory)
    auto cols = [&](int m, double(&xc)[8]) {
        const double2 *s2 = reinterpret_cast<const double2 *>(xs + m * NP) + 4 * h;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2 v = s2[j ^ q];
            xc[2 * j] = v.x;
            xc[2 * j + 1] = v.y;
        }
    };
    // (A x)_own for the stage vector in buffer m
    auto rowprod = [&](const double(&A)[4][8], const double(&xc)[8], double(&y)[2]) {
        double s[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            double acc = A[kk][0] * xc[0];
#pragma unroll
            for (int cc = 1; cc < 8; ++cc) acc = fma(A[kk][cc], xc[cc], acc);
            s[kk] = acc;
        }
        y[0] = s[0] + shx(s[2], 1);
        y[1] = s[1] + shx(s[3], 1);
    };

    for (int64_t it = 0;; ++it) {
        const int64_t b = slot + it * nslots;
        const bool has = b < a.B;
        if (!__any_sync(FULL, has)) break;
        const int64_t bs = has ? b : 0; // idle slots read trajectory 0 and discard
        const double *pb = a.params + bs * npar;
        double A[4][8];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            if (EXACT) {
                const double2 *src = reinterpret_cast<const double2 *>(pb + NP + rowi[kk] * NP) + 4 * h;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double2 v = __ldg(src + (j ^ q));
                    A[kk][2 * j] = v.x;
                    A[kk][2 * j + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const bool in = rowi[kk] < n && coli[cc] < n;
                    const double v = __ldg(pb + n + (in ? rowi[kk] * n + coli[cc] : 0)); // padded entries read a valid address and discard it
                    A[kk][cc] = in ? v : 0.0;
                }
            }
        }
        double r[2], x[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int oi = live[k] ? own0 + k : 0;
            const double rv = __ldg(pb + oi), xv = __ldg(a.x0 + bs * n + oi);
            r[k] = live[k] ? rv : 0.0;
            x[k] = live[k] ? xv : 0.0;
        }
        // ================================ forward sweep =====================================
        double t = a.ti, dt = a.dt0, K[S][2];
        int nck = 0, rejects = 0, status = 0, trials = 0;
        bool act = has && !a.skip_forward && (ADAPTIVE ? va_less_with_sign(t, tf, dt) : va_less_eq_with_sign(t + dt, tf, dt));
        bool fresh = true;
        K[0][0] = K[0][1] = 0.0;
        if (a.skip_forward && has) {
            // split API (va_forward_batch then va_adjoint_batch): the checkpoints of this very trajectory are still in the slab
            nck = a.n_accept[b];
            status = a.status[b];
            t = slab[(int64_t)nck * BLK];
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (live[k]) x[k] = a.x_final[b * n + own0 + k];
        }
        while (__any_sync(FULL, act)) {
            // first slope of a step, f(x_n): after every acceptance (dopri5: only for the very first step, afterwards the FSAL
            // slope is reused). All lanes of the warp run the product; only trajectories that need it keep the result.
            const bool need0 = act && fresh && (!Tab::FSAL || nck == 0);
            if (__any_sync(FULL, need0)) {
                double xc[8], gg[2];
                put(0, x);
                cols(0, xc);
                rowprod(A, xc, gg);
                if (need0) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) K[0][k] = x[k] * (r[k] + gg[k]);
                }
            }
            if (act && fresh) {
                if (nck >= a.cap) {
                    status |= VA_TRAJ_CKPT_OVERFLOW;
                    act = false;
                } else {
                    double *blkp = slab + (int64_t)nck * BLK;
                    *reinterpret_cast<double2 *>(blkp + HDR + own0) = make_double2(x[0], x[1]);
                    if (o == 0) blkp[0] = t;
                    if (ADAPTIVE && va_less_with_sign(tf, t + dt, dt)) dt = tf - t;
                    trials = 0;
                    fresh = false;
                }
            }
#pragma unroll
            for (int m = 1; m < SE; ++m) {
                double xm[2], gm[2], xc[8];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < m; ++j)
                        if (Tab::a(m, j) != 0.0) acc = fma(Tab::a(m, j), K[j][k], acc);
                    xm[k] = fma(dt, acc, x[k]);
                }
                put(m, xm);
                cols(m, xc);
                rowprod(A, xc, gm);
#pragma unroll
                for (int k = 0; k < 2; ++k) K[m][k] = xm[k] * (r[k] + gm[k]);
            }
            double xn[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < SE; ++j)
                    if (Tab::b(j) != 0.0) acc = fma(Tab::b(j), K[j][k], acc);
                xn[k] = fma(dt, acc, x[k]);
            }
            if (Tab::FSAL) {
                double xc[8], gn[2];
                put(S - 1, xn);
                cols(S - 1, xc);
                rowprod(A, xc, gn);
#pragma unroll
                for (int k = 0; k < 2; ++k) K[S - 1][k] = xn[k] * (r[k] + gn[k]);
            }
            double err = 0.0;
            if (ADAPTIVE) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < S; ++j)
                        if (Tab::db(j) != 0.0) acc = fma(Tab::db(j), K[j][k], acc);
                    const double e = fabs(dt * acc) / (a.eps_abs + a.eps_rel * (fabs(x[k]) + fabs(dt) * fabs(K[0][k])));
                    if (live[k]) err = fmax(err, e);
                }
                err = fmax(err, shx(err, 1));
                err = fmax(err, shx(err, 2));
                err = fmax(err, shx(err, 4));
            }
            if (act) {
                const bool accept = !ADAPTIVE || !(err > 1.0);
                if (!accept) {
                    dt *= fmax(0.9 * inv_root<(Tab::ERROR_ORDER > 1 ? Tab::ERROR_ORDER - 1 : 1)>(err), 0.2);
                    ++rejects;
                    if (++trials >= 500) {
                        status |= VA_TRAJ_NO_PROGRESS;
                        act = false;
                    }
                } else {
                    x[0] = xn[0];
                    x[1] = xn[1];
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
                        act = va_less_with_sign(t, tf, dt);
                    } else {
                        t = a.ti + (double)nck * dt;
                        act = va_less_eq_with_sign(t + dt, tf, dt);
                    }
                    fresh = true;
                    if (Tab::FSAL) {
                        K[0][0] = K[S - 1][0];
                        K[0][1] = K[S - 1][1];
                    }
                }
            }
        }
        // close the trajectory: final time, status, x(tf)
        const int T = nck;
        if (has && o == 0 && !a.skip_forward) slab[(int64_t)T * BLK] = t; // header of checkpoint T carries the final time
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (live[k] && !isfinite(x[k])) status |= VA_TRAJ_NONFINITE;
        status |= __shfl_xor_sync(FULL, status, 1);
        status |= __shfl_xor_sync(FULL, status, 2);
        status |= __shfl_xor_sync(FULL, status, 4);
        const bool failed = status & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS);
        const double t_final = t;
        if (has && !a.skip_forward) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (live[k]) a.x_final[b * n + own0 + k] = failed ? nan("") : x[k];
            if (o == 0) {
                if (a.n_accept) a.n_accept[b] = T;
                if (a.n_reject) a.n_reject[b] = rejects;
                if (a.status) a.status[b] = status;
            }
        }
        const bool ok = has && !failed;
        const int Tw = __reduce_max_sync(FULL, ok ? T : 0); // steps the warp walks in the reverse sweep
        __syncwarp(); // the trajectory's checkpoint stores are ordered before the loads below

        // ================================ reverse sweep, one pass per cost function =====================================
        for (int oc = 0; oc < a.n_out; ++oc) {
            double *lam_io = a.lambda + (bs * a.n_out + oc) * n;
            const bool sum_mode = a.reduce == VA_REDUCE_SUM;
            double *gbar = sum_mode ? part : a.mu + (bs * a.n_out + oc) * npar;
            const bool overwrite = !sum_mode || !row_init; // first use of this accumulator row
            if (has && failed) {
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (live[k]) lam_io[own0 + k] = nan("");
                if (!sum_mode)
                    for (int k = o; k < npar; k += 8) gbar[k] = nan("");
            }
            double Ab[4][8];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) Ab[kk][cc] = 0.0;
            double rbar[2] = {0.0, 0.0}, lam[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double seed = (ok && live[k] && a.objective == VA_OBJ_SEED) ? lam_io[own0 + k] : 0.0;
                lam[k] = !live[k] ? 0.0 : a.objective == VA_OBJ_SUM ? 1.0 : a.objective == VA_OBJ_HALF_NORM2 ? x[k] : seed;
            }
            double t_hi = t_final;
            // checkpoint of the first step to process; the next one is fetched while the current one is being worked on
            const double *ckp = slab + (int64_t)(ok && T > 0 ? T - 1 : 0) * BLK;
            // (idle slots and failed trajectories have no checkpoints to read: their lanes carry zeros through the warp's steps)
            double2 xn_next = ok ? *reinterpret_cast<const double2 *>(ckp + HDR + own0) : make_double2(0.0, 0.0);
            double tn_next = ok ? ckp[0] : 0.0;
#pragma unroll 1
            for (int s = 0; s < Tw; ++s) {
                const int step = T - 1 - s;
                const bool a2 = ok && step >= 0;
                const double xn0 = xn_next.x, xn1 = xn_next.y, t_lo = tn_next;
                {
                    const double *nx = slab + (int64_t)(ok && step >= 1 ? step - 1 : 0) * BLK;
                    if (ok) {
                        xn_next = *reinterpret_cast<const double2 *>(nx + HDR + own0);
                        tn_next = nx[0];
                    }
                }
                const double dt_s = t_hi - t_lo; // StateStorage::GetDt: difference of the stored times
                // ---- stage recompute from x_n (detail/backpropagation.hpp:24-64): X_m -> shared buffer m, g_m, K_m in registers
                double g[SADJ][2], Kr[SADJ][2];
                {
                    double Xo[2] = {xn0, xn1};
#pragma unroll
                    for (int m = 0; m < SADJ; ++m) {
                        double xc[8], y[2];
                        put(m, Xo);
                        cols(m, xc);
                        rowprod(A, xc, y);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            g[m][k] = r[k] + y[k];
                            Kr[m][k] = Xo[k] * g[m][k];
                        }
                        if (m + 1 < SADJ) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                double acc = 0.0;
#pragma unroll
                                for (int j = 0; j <= m; ++j)
                                    if (Tab::a(m + 1, j) != 0.0) acc = fma(Tab::a(m + 1, j), Kr[j][k], acc);
                                Xo[k] = fma(dt_s, acc, k == 0 ? xn0 : xn1);
                            }
                        }
                    }
                }
                // ---- one-step adjoint (detail/backpropagation.hpp:83-158) with the gradient accumulated on the fly
                double W[SADJ + 1][2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const double l = a2 ? lam[k] : 0.0; // finished trajectories carry zero seeds through the warp's remaining steps
                    W[0][k] = l;
#pragma unroll
                    for (int m = 1; m <= SADJ; ++m) W[m][k] = (Tab::b(m - 1) * dt_s) * l;
                }
#pragma unroll
                for (int m = SADJ; m >= 1; --m) {
                    double xc[8], vv[4], p[8];
                    cols(m - 1, xc);
                    const double2 xo = *reinterpret_cast<const double2 *>(xs + (m - 1) * NP + own0);
                    vv[0] = W[m][0] * xo.x;
                    vv[1] = W[m][1] * xo.y;
                    vv[2] = shx(vv[0], 1);
                    vv[3] = shx(vv[1], 1);
                    // z = A^T v over the lane's rows, and Abar += v X_{m-1}^T on the same operands
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        double acc = A[0][cc] * vv[0];
#pragma unroll
                        for (int kk = 1; kk < 4; ++kk) acc = fma(A[kk][cc], vv[kk], acc);
                        p[cc] = acc;
                    }
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc) Ab[kk][cc] = fma(vv[kk], xc[cc], Ab[kk][cc]);
                    rbar[0] += vv[0];
                    rbar[1] += vv[1];
                    // reduce-scatter over q: registers 0..3 then 0..1 are the ones this lane keeps (columns held permuted)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) p[cc] += shx(p[cc + 4], 4);
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) p[cc] += shx(p[cc + 2], 2);
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const double gx = fma(W[m][k], g[m - 1][k], p[k]);
                        W[0][k] += gx;
#pragma unroll
                        for (int j = 1; j < m; ++j)
                            if (Tab::a(m - 1, j - 1) != 0.0) W[j][k] = fma(gx * Tab::a(m - 1, j - 1), dt_s, W[j][k]);
                    }
                }
                if (a2) {
                    t_hi = t_lo;
                    lam[0] = W[0][0];
                    lam[1] = W[0][1];
                }
                __syncwarp(); // the stage buffers are rewritten by the next step
            }
            if (ok) {
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (live[k]) lam_io[own0 + k] = lam[k];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (!live[k]) continue;
                    if (overwrite) gbar[own0 + k] = rbar[k];
                    else atomicAdd(gbar + own0 + k, rbar[k]);
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if (rowi[kk] >= n) continue;
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        if (!EXACT && coli[cc] >= n) continue;
                        double *dst = gbar + n + rowi[kk] * n + coli[cc];
                        if (overwrite) *dst = Ab[kk][cc];
                        else atomicAdd(dst, Ab[kk][cc]);
                    }
                }
                row_init = true;
            }
        }
    }
    if (a.reduce == VA_REDUCE_SUM && a.n_out > 0 && !row_init)
        for (int k = o; k < npar; k += 8) part[k] = 0.0; // this slot integrated nothing (or only failed trajectories)
}

template <class Tab, bool ADAPTIVE>
cudaError_t launch(const VaGlvWideArgs &a, cudaStream_t st, size_t smem)
{
    cudaError_t e = cudaFuncSetAttribute(k_glv_oct<Tab, ADAPTIVE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_glv_oct<Tab, ADAPTIVE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (a.n == NP) k_glv_oct<Tab, ADAPTIVE, true><<<a.grid, NT, smem, st>>>(a);
    else k_glv_oct<Tab, ADAPTIVE, false><<<a.grid, NT, smem, st>>>(a);
    return cudaGetLastError();
}

} // namespace

bool va_glv_oct_supported(int n, int stepper, int adaptive)
{
    if (n < 1 || n > NP) return false;
    if (stepper == VA_RK_RK4) return !adaptive;
    if (stepper == VA_RK_CK54 || stepper == VA_RK_DOPRI5) return adaptive != 0;
    return false;
}

int va_glv_oct_block_doubles() { return BLK; }
int va_glv_oct_slots_per_cta() { return TPC; }
int va_glv_oct_threads() { return NT; }

cudaError_t va_glv_oct_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const size_t smem = (size_t)TPC * TSTRIDE * 8;
    switch (a.stepper) {
    case VA_RK_RK4: return launch<TabRK4, false>(a, st, smem);
    case VA_RK_CK54: return launch<TabCK54, true>(a, st, smem);
    case VA_RK_DOPRI5: return launch<TabDOPRI5, true>(a, st, smem);
    }
    return cudaErrorInvalidValue;
}
