// va_glv_quad.cu -- Generalized Lotka-Volterra with up to 16 species (BASELINE config 3: N = 16, one million parameter sets):
// FOUR lanes per trajectory, eight trajectories per warp, no block-level barrier anywhere.
//
// The first-generation kernel (va_glv_wide.cu) gives a warp to each 16-species trajectory: 8 matrix entries per thread, so
// every stage is an exchange (barrier, operand loads, 3-level reduction) around 8 DFMAs -- 18 % of the FP64 peak, and not
// occupancy-limited (profiles/README.md). Here lane q of a quad holds rows 4q..4q+3 of the current phase's matrix COMPLETE
// (4 x 16 entries, 128 registers) and owns components 4q..4q+3 of every vector: a matrix-vector product is 64 DFMAs per
// lane with no reduction at all; the only exchange is the all-gather of the 16-vector through 128 bytes of shared memory
// behind one __syncwarp (double-buffered, so one warp barrier per product).
//
// Same algorithm and three phases as va_glv_t8.cu (reference lib/include/detail/runge_kutta.hpp:76-118 forward sweep with
// odeint's controlled stepper; detail/backpropagation.hpp:83-158, 231-254 reverse sweep; store-stages policy):
//   1. forward sweep with A rows in registers; every accepted step leaves a block
//      [8-double header (t_n) | X_0..X_{s-1} | g_0..g_{s-1} (| v_0..v_{s-1})] in the quad's slab (v aliases g with one seed);
//   2. state adjoint with A^T rows in registers; the step blocks come back through cp.async into a per-quad shared-memory
//      double buffer one step ahead; v_m = w_m o X_{m-1} is written to the slab;
//   3. gradient accumulation Abar += v_m X_{m-1}^T with the Abar rows in registers (64 independent DFMA chains per lane),
//      blocks again one step ahead through cp.async.
// The eight trajectories of a warp run in lock step: every loop runs to the longest of the eight, lanes of finished
// trajectories compute and discard (accepted-step counts of neighbouring parameter sets differ by a few steps at most).
#include "va_glv_common.cuh"

#ifndef VA_QUAD_WARPS
#define VA_QUAD_WARPS 8
#endif

namespace {

constexpr int NP = 16;               // padded species count
constexpr int HDR = 8;               // doubles in a step-block header (hdr[0] = t_n)
constexpr int NT = 32 * VA_QUAD_WARPS; // threads per CTA
constexpr int QPC = NT / 4;          // trajectories (quads) per CTA
constexpr unsigned FULL = 0xffffffffu;
constexpr int SADJ_MAX = 6;
constexpr int BUF = HDR + 2 * SADJ_MAX * NP; // doubles per shared-memory block buffer (header, X section, g or v section)
// shared memory per quad; 16 bytes more than a multiple of 128: the eight quads of a warp broadcast-read eight different
// 16-byte bank groups (one wavefront per LDS.128; with a multiple of 128 every operand load was an 8-way bank conflict and
// the LSU data path 86 % busy)
constexpr int QSTRIDE = 2 * NP + 2 * BUF + 2;
static_assert((QSTRIDE * 8) % 128 == 16, "bank-group skew between the quads of a warp");

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(FULL, v, m); }
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NLEFT>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NLEFT) : "memory"); }

template <class Tab, bool ADAPTIVE, bool EXACT>
__global__ void __launch_bounds__(NT, 1) k_glv_quad(const __grid_constant__ VaGlvWideArgs a)
{
    constexpr int S = Tab::S, SADJ = Tab::SADJ;
    constexpr int SE = Tab::FSAL ? S - 1 : S;
    static_assert(SADJ <= SADJ_MAX, "block buffers");
    extern __shared__ __align__(128) double sm_all[]; // per quad: xs[2][NP] exchange buffers, bb[2][BUF] block buffers
    const int lane = threadIdx.x & 31, q = lane & 3;
    const int qc = threadIdx.x >> 2; // quad inside the CTA
    double *const xs = sm_all + (size_t)qc * QSTRIDE;
    double *const bb = xs + 2 * NP;
    const int n = EXACT ? NP : a.n;
    const int npar = n * n + n;
    const int blk = a.blk_doubles;
    const int voff = blk - SADJ * NP; // v section of a step block (== the g section with a single seed per trajectory)
    const int64_t slot = (int64_t)blockIdx.x * QPC + qc, nslots = (int64_t)gridDim.x * QPC;
    double *const slab = a.slab + slot * a.slab_stride;
    double *const part = a.partial + slot * npar; // summed mode: this quad's partial-sum row
    const double tf = a.tf;
    bool live[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) live[k] = 4 * q + k < n;
    bool row_init = false;
    int p = 0; // exchange buffer parity

    // all-gather of a 16-vector inside the quad: every lane contributes its 4 components; one warp barrier per product
    // (a buffer is rewritten two products later, behind the barrier of the product in between)
    auto put = [&](const double(&v)[4]) {
        double2 *d = reinterpret_cast<double2 *>(xs + p * NP + 4 * q);
        d[0] = make_double2(v[0], v[1]);
        d[1] = make_double2(v[2], v[3]);
        __syncwarp();
    };
    // y_k = sum_c M[k][c] vec[c] for the lane's four rows, vec = the vector just put
    auto dot = [&](const double(&M)[4][NP], double(&y)[4]) {
        const double2 *s2 = reinterpret_cast<const double2 *>(xs + p * NP);
        double z[4]; // second chain per row: the dependent chains are 8 DFMAs long instead of 16
#pragma unroll
        for (int c = 0; c < NP / 2; c += 2) {
            const double2 v = s2[c], w = s2[c + 1];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                y[k] = c == 0 ? M[k][0] * v.x : fma(M[k][2 * c], v.x, y[k]);
                z[k] = c == 0 ? M[k][2] * w.x : fma(M[k][2 * c + 2], w.x, z[k]);
                y[k] = fma(M[k][2 * c + 1], v.y, y[k]);
                z[k] = fma(M[k][2 * c + 3], w.y, z[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] += z[k];
        p ^= 1;
    };
    // one step block (header + X section, and the g or v section) -> shared-memory buffer `which`, 16 bytes per cp.async
    auto fetch_block = [&](const double *blkp, int second_off, int which) {
        double *dst = bb + which * BUF;
        constexpr int C1 = (HDR + SADJ * NP) / 2, C2 = SADJ * NP / 2; // 16-byte pieces
        for (int i = q; i < C1; i += 4) cp_async16(dst + 2 * i, blkp + 2 * i);
        for (int i = q; i < C2; i += 4) cp_async16(dst + HDR + SADJ * NP + 2 * i, blkp + second_off + 2 * i);
        cp_async_commit();
    };

    for (int64_t it = 0;; ++it) {
        const int64_t b = slot + it * nslots;
        const bool has = b < a.B;
        if (!__any_sync(FULL, has)) break;
        const int64_t bs = has ? b : 0; // idle quads read trajectory 0 and discard
        const double *pb = a.params + bs * npar;
        double M[4][NP];
        // ================================ phase 1: forward sweep (rows of A) =====================================
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int row = 4 * q + k;
            if (EXACT) {
                const double2 *src = reinterpret_cast<const double2 *>(pb + NP + row * NP);
#pragma unroll
                for (int c = 0; c < NP / 2; ++c) {
                    const double2 v = __ldg(src + c);
                    M[k][2 * c] = v.x;
                    M[k][2 * c + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int c = 0; c < NP; ++c) {
                    const bool in = row < n && c < n;
                    const double v = __ldg(pb + n + (in ? row * n + c : 0)); // padded entries read a valid address and discard it
                    M[k][c] = in ? v : 0.0;
                }
            }
        }
        double r[4], x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int oi = live[k] ? 4 * q + k : 0;
            const double rv = __ldg(pb + oi), xv = __ldg(a.x0 + bs * n + oi);
            r[k] = live[k] ? rv : 0.0;
            x[k] = live[k] ? xv : 0.0;
        }
        double t = a.ti, dt = a.dt0, K[S][4], g0[4];
        int nck = 0, rejects = 0, status = 0, trials = 0;
        bool act = has && (ADAPTIVE ? va_less_with_sign(t, tf, dt) : va_less_eq_with_sign(t + dt, tf, dt));
        bool fresh = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) { g0[k] = 0.0; K[0][k] = 0.0; }
        while (__any_sync(FULL, act)) {
            // first slope of a step, f(x_n): needed after every acceptance (dopri5: only for the very first step, afterwards the
            // FSAL slope is reused). All lanes of the warp run the product; only quads that need it keep the result.
            const bool need0 = act && fresh && (!Tab::FSAL || nck == 0);
            if (__any_sync(FULL, need0)) {
                double gg[4];
                put(x);
                dot(M, gg);
                if (need0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        g0[k] = r[k] + gg[k];
                        K[0][k] = x[k] * g0[k];
                    }
                }
            }
            double *blkp = slab + (int64_t)(act ? nck : 0) * blk;
            if (act && fresh) {
                if (nck >= a.cap) {
                    status |= VA_TRAJ_CKPT_OVERFLOW;
                    act = false;
                } else {
                    double2 *dx = reinterpret_cast<double2 *>(blkp + HDR + 4 * q), *dg = reinterpret_cast<double2 *>(blkp + HDR + SADJ * NP + 4 * q);
                    dx[0] = make_double2(x[0], x[1]);
                    dx[1] = make_double2(x[2], x[3]);
                    dg[0] = make_double2(g0[0], g0[1]);
                    dg[1] = make_double2(g0[2], g0[3]);
                    if (q == 0) blkp[0] = t;
                    if (ADAPTIVE && va_less_with_sign(tf, t + dt, dt)) dt = tf - t;
                    trials = 0;
                    fresh = false;
                }
            }
#pragma unroll
            for (int m = 1; m < SE; ++m) {
                double xm[4], gm[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < m; ++j)
                        if (Tab::a(m, j) != 0.0) acc = fma(Tab::a(m, j), K[j][k], acc);
                    xm[k] = fma(dt, acc, x[k]);
                }
                put(xm);
                dot(M, gm);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    gm[k] += r[k];
                    K[m][k] = xm[k] * gm[k];
                }
                if (act && m < SADJ) {
                    double2 *dx = reinterpret_cast<double2 *>(blkp + HDR + m * NP + 4 * q),
                            *dg = reinterpret_cast<double2 *>(blkp + HDR + (SADJ + m) * NP + 4 * q);
                    dx[0] = make_double2(xm[0], xm[1]);
                    dx[1] = make_double2(xm[2], xm[3]);
                    dg[0] = make_double2(gm[0], gm[1]);
                    dg[1] = make_double2(gm[2], gm[3]);
                }
            }
            double xn[4], gn[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < SE; ++j)
                    if (Tab::b(j) != 0.0) acc = fma(Tab::b(j), K[j][k], acc);
                xn[k] = fma(dt, acc, x[k]);
                gn[k] = 0.0;
            }
            if (Tab::FSAL) {
                put(xn);
                dot(M, gn);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    gn[k] += r[k];
                    K[S - 1][k] = xn[k] * gn[k];
                }
            }
            double err = 0.0;
            if (ADAPTIVE) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < S; ++j)
                        if (Tab::db(j) != 0.0) acc = fma(Tab::db(j), K[j][k], acc);
                    const double e = fabs(dt * acc) / (a.eps_abs + a.eps_rel * (fabs(x[k]) + fabs(dt) * fabs(K[0][k])));
                    if (live[k]) err = fmax(err, e);
                }
                err = fmax(err, shx(err, 1));
                err = fmax(err, shx(err, 2));
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
#pragma unroll
                    for (int k = 0; k < 4; ++k) x[k] = xn[k];
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
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            g0[k] = gn[k];
                            K[0][k] = K[S - 1][k];
                        }
                    }
                }
            }
        }
        // close the trajectory: final time, status, x(tf)
        const int T = nck;
        if (has && q == 0) slab[(int64_t)T * blk] = t; // header of block T carries the final time
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (live[k] && !isfinite(x[k])) status |= VA_TRAJ_NONFINITE;
        status |= __shfl_xor_sync(FULL, status, 1);
        status |= __shfl_xor_sync(FULL, status, 2);
        const bool failed = status & (VA_TRAJ_CKPT_OVERFLOW | VA_TRAJ_NO_PROGRESS);
        const double t_final = t;
        if (has) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (live[k]) a.x_final[b * n + 4 * q + k] = failed ? nan("") : x[k];
            if (q == 0) {
                if (a.n_accept) a.n_accept[b] = T;
                if (a.n_reject) a.n_reject[b] = rejects;
                if (a.status) a.status[b] = status;
            }
        }
        const bool ok = has && !failed;
        const int Tw = __reduce_max_sync(FULL, ok ? T : 0); // steps the warp walks in the reverse phases
        __syncwarp(); // the quad's slab stores are ordered before the block fetches below

        for (int o = 0; o < a.n_out; ++o) {
            double *lam_io = a.lambda + (bs * a.n_out + o) * n;
            const bool sum_mode = a.reduce == VA_REDUCE_SUM;
            double *gbar = sum_mode ? part : a.mu + (bs * a.n_out + o) * npar;
            const bool overwrite = !sum_mode || !row_init; // first use of this accumulator row
            if (has && failed) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (live[k]) lam_io[4 * q + k] = nan("");
                if (!sum_mode)
                    for (int k = q; k < npar; k += 4) gbar[k] = nan("");
            }
            // ================================ phase 2: adjoint of the state (rows of A^T) =====================================
#pragma unroll
            for (int c = 0; c < NP; ++c) {
                if (EXACT) {
                    const double2 *src = reinterpret_cast<const double2 *>(pb + NP + c * NP + 4 * q);
                    const double2 v0 = __ldg(src), v1 = __ldg(src + 1);
                    M[0][c] = v0.x;
                    M[1][c] = v0.y;
                    M[2][c] = v1.x;
                    M[3][c] = v1.y;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool in = c < n && live[k];
                        const double v = __ldg(pb + n + (in ? c * n + 4 * q + k : 0));
                        M[k][c] = in ? v : 0.0;
                    }
                }
            }
            double W[SADJ + 1][4], lam[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double seed = (ok && live[k] && a.objective == VA_OBJ_SEED) ? lam_io[4 * q + k] : 0.0;
                lam[k] = !live[k] ? 0.0 : a.objective == VA_OBJ_SUM ? 1.0 : a.objective == VA_OBJ_HALF_NORM2 ? x[k] : seed;
            }
            double t_hi = t_final;
            if (Tw > 0) fetch_block(slab + (int64_t)(ok && T > 0 ? T - 1 : 0) * blk, HDR + SADJ * NP, 0);
#pragma unroll 1
            for (int s = 0; s < Tw; ++s) {
                const int step = T - 1 - s;
                const bool a2 = ok && step >= 0;
                if (s + 1 < Tw) {
                    fetch_block(slab + (int64_t)(ok && step >= 1 ? step - 1 : 0) * blk, HDR + SADJ * NP, (s + 1) & 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncwarp();
                const double *cur = bb + (s & 1) * BUF;
                double *blkp = slab + (int64_t)(a2 ? step : 0) * blk;
                const double t_lo = cur[0];
                const double dt_s = t_hi - t_lo;
#pragma unroll
                for (int k = 0; k < 4; ++k) W[0][k] = lam[k];
#pragma unroll
                for (int m = 1; m <= SADJ; ++m)
#pragma unroll
                    for (int k = 0; k < 4; ++k) W[m][k] = (Tab::b(m - 1) * dt_s) * lam[k];
#pragma unroll
                for (int m = SADJ; m >= 1; --m) {
                    const double2 *px = reinterpret_cast<const double2 *>(cur + HDR + (m - 1) * NP + 4 * q),
                                  *pg = reinterpret_cast<const double2 *>(cur + HDR + (SADJ + m - 1) * NP + 4 * q);
                    const double2 x0v = px[0], x1v = px[1], g0v = pg[0], g1v = pg[1];
                    const double X[4] = {x0v.x, x0v.y, x1v.x, x1v.y}, G[4] = {g0v.x, g0v.y, g1v.x, g1v.y};
                    double v[4], atv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = W[m][k] * X[k];
                    put(v);
                    if (a2) {
                        double2 *dv = reinterpret_cast<double2 *>(blkp + voff + (m - 1) * NP + 4 * q);
                        dv[0] = make_double2(v[0], v[1]);
                        dv[1] = make_double2(v[2], v[3]);
                    }
                    dot(M, atv);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double gx = fma(W[m][k], G[k], atv[k]);
                        W[0][k] += gx;
#pragma unroll
                        for (int j = 1; j < m; ++j)
                            if (Tab::a(m - 1, j - 1) != 0.0) W[j][k] = fma(gx * Tab::a(m - 1, j - 1), dt_s, W[j][k]);
                    }
                }
                if (a2) { // quads whose trajectory has no such step keep their lambda while the warp finishes longer ones
                    t_hi = t_lo;
#pragma unroll
                    for (int k = 0; k < 4; ++k) lam[k] = W[0][k];
                }
            }
            if (ok) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (live[k]) lam_io[4 * q + k] = lam[k];
            }
            __syncwarp(); // every v of the quad is written before the fetches of phase 3

            // ================================ phase 3: gradient accumulation (rows of Abar) =====================================
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int c = 0; c < NP; ++c) M[k][c] = 0.0;
            double rbar[4] = {0.0, 0.0, 0.0, 0.0};
            if (Tw > 0) fetch_block(slab, voff, 0);
#pragma unroll 1
            for (int s = 0; s < Tw; ++s) {
                const bool a3 = ok && s < T;
                if (s + 1 < Tw) {
                    fetch_block(slab + (int64_t)(ok && s + 1 < T ? s + 1 : 0) * blk, voff, (s + 1) & 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncwarp();
                const double *cur = bb + (s & 1) * BUF;
                if (a3) {
#pragma unroll
                    for (int m = 0; m < SADJ; ++m) {
                        const double2 *pv = reinterpret_cast<const double2 *>(cur + HDR + (SADJ + m) * NP + 4 * q);
                        const double2 v0 = pv[0], v1 = pv[1];
                        const double v[4] = {v0.x, v0.y, v1.x, v1.y};
                        const double2 *pxv = reinterpret_cast<const double2 *>(cur + HDR + m * NP);
#pragma unroll
                        for (int c = 0; c < NP / 2; ++c) {
                            const double2 xv = pxv[c];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                M[k][2 * c] = fma(v[k], xv.x, M[k][2 * c]);
                                M[k][2 * c + 1] = fma(v[k], xv.y, M[k][2 * c + 1]);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) rbar[k] += v[k];
                    }
                }
                __syncwarp(); // the buffer is refilled two iterations later, behind this barrier
            }
            if (ok) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int row = 4 * q + k;
                    if (!live[k]) continue;
                    if (overwrite) gbar[row] = rbar[k];
                    else atomicAdd(gbar + row, rbar[k]);
#pragma unroll
                    for (int c = 0; c < NP; ++c) {
                        if (!EXACT && c >= n) continue;
                        double *dst = gbar + n + row * n + c;
                        if (overwrite) *dst = M[k][c];
                        else atomicAdd(dst, M[k][c]);
                    }
                }
                row_init = true;
            }
        }
    }
    if (a.reduce == VA_REDUCE_SUM && a.n_out > 0 && !row_init)
        for (int k = q; k < npar; k += 4) part[k] = 0.0; // this quad integrated nothing (or only failed trajectories)
}

template <class Tab, bool ADAPTIVE>
cudaError_t launch(const VaGlvWideArgs &a, cudaStream_t st, size_t smem)
{
    cudaError_t e = cudaFuncSetAttribute(k_glv_quad<Tab, ADAPTIVE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_glv_quad<Tab, ADAPTIVE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (a.n == NP) k_glv_quad<Tab, ADAPTIVE, true><<<a.grid, NT, smem, st>>>(a);
    else k_glv_quad<Tab, ADAPTIVE, false><<<a.grid, NT, smem, st>>>(a);
    return cudaGetLastError();
}

} // namespace

bool va_glv_quad_supported(int n, int stepper, int adaptive)
{
    if (n < 1 || n > NP) return false;
    if (stepper == VA_RK_RK4) return !adaptive;
    if (stepper == VA_RK_CK54 || stepper == VA_RK_DOPRI5) return adaptive != 0;
    return false;
}

int va_glv_quad_block_doubles(int stepper, int n_out)
{
    const int sadj = stepper == VA_RK_RK4 ? TabRK4::SADJ : stepper == VA_RK_CK54 ? TabCK54::SADJ : TabDOPRI5::SADJ;
    return HDR + (n_out > 1 ? 3 : 2) * sadj * NP;
}

int va_glv_quad_slots_per_cta() { return QPC; }
int va_glv_quad_threads() { return NT; }

cudaError_t va_glv_quad_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st)
{
    if (a.B <= 0) return cudaSuccess;
    const size_t smem = (size_t)QPC * QSTRIDE * 8;
    switch (a.stepper) {
    case VA_RK_RK4: return launch<TabRK4, false>(a, st, smem);
    case VA_RK_CK54: return launch<TabCK54, true>(a, st, smem);
    case VA_RK_DOPRI5: return launch<TabDOPRI5, true>(a, st, smem);
    }
    return cudaErrorInvalidValue;
}
