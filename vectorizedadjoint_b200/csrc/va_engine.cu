// va_engine.cu -- the C-ABI (include/va_engine.h) on top of the kernel families.
//
// Host-side role of the reference's Driver (lib/include/Driver.hpp:15-79): it owns the checkpoint storage
// (StateStorage -> device arena / slabs), the tableau (ButcherTable -> VaTableau) and the RHS/VJP functors
// (AadData -> CUDA device functors), and it borrows the caller's lambda / mu buffers for the duration of a call.
//
// Memory plan on a 180 GB B200
//   device-resident calls (mem == VA_MEM_DEVICE): inputs/outputs are the caller's HBM buffers, nothing is staged.
//     wide GLV family  : one persistent launch over the whole batch; checkpoints in per-CTA slabs (L2-resident).
//     scalar family    : the batch is cut into arena-sized waves; forward and reverse kernels alternate per wave.
//   host calls (mem == VA_MEM_HOST): a 3-stream pipeline (H2D | compute | D2H) over double-buffered chunks, so PCIe
//     transfers of chunk c+1 / c-1 overlap the kernels of chunk c.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "va_engine_impl.h"

namespace {
thread_local std::string g_last_error;
} // namespace

int va_fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}
const std::string &va_tls_error() { return g_last_error; }

namespace {
int fail(int code, const std::string &msg) { return va_fail(code, msg); }
} // namespace

namespace {

// The persisting-L2 carve-out is DEVICE state: it shrinks the L2 every other kernel on the GPU sees (measured: the 16-species
// kernel of a later engine ran 2.4x slower with an 83 MB carve-out left behind by a destroyed 64-species engine). Engines that
// want it take a per-device reference; the last one to go restores the limit it found and drops the persisting lines.
std::mutex g_persist_mutex;
int g_persist_refs[64] = {0};
size_t g_persist_prev[64] = {0};
void persist_l2_acquire(va_engine *e, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_persist_mutex);
    const int d = e->device & 63;
    if (g_persist_refs[d]++ == 0) {
        cudaDeviceGetLimit(&g_persist_prev[d], cudaLimitPersistingL2CacheSize);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes);
    }
    e->holds_persist_l2 = true;
}
void persist_l2_release(va_engine *e)
{
    if (!e->holds_persist_l2) return;
    std::lock_guard<std::mutex> lk(g_persist_mutex);
    const int d = e->device & 63;
    if (--g_persist_refs[d] == 0) {
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, g_persist_prev[d]);
    }
    e->holds_persist_l2 = false;
}

int ck_layout_of(const va_engine *e) { return e->desc.adaptive ? 1 : 0; } // see va_scalar_kernels.cuh: ck_store

bool is_glv(const va_engine *e) { return e->family == FAM_GLV_WIDE || e->family == FAM_GLV_STREAM; }

// doubles per checkpoint record: layout 1 (adaptive) pads (t, x) to a multiple of four (VA_CK_REC in va_scalar_kernels.cuh)
int ck_rec_doubles(const va_engine *e) { return e->desc.adaptive ? ((e->desc.n_state + 1 + 3) & ~3) : e->desc.n_state + 1; }
int64_t per_traj_arena_bytes(const va_engine *e) { return (int64_t)(e->cap + 1) * ck_rec_doubles(e) * 8; }

int ensure_workspace(va_engine *e, int64_t B)
{
    if (is_glv(e)) {
        const size_t need = (size_t)e->grid * e->tpc * e->pair * e->slab_stride * 8;
        const void *slab_before = e->slab.p;
        if (int rc = e->slab.ensure(need)) return rc;
        // a fresh slab is cleared once: idle slots (fewer trajectories than slots, lock-step warps) prefetch block 0 of slabs that
        // no trajectory has written yet and discard it -- harmless, but it should not be indeterminate memory (initcheck-clean)
        if (e->slab.p != slab_before) {
            VA_CUDA(cudaMemsetAsync(e->slab.p, 0, e->slab.bytes, e->s_comp));
            VA_CUDA(cudaStreamSynchronize(e->s_comp)); // once per allocation; the kernels may run on a caller's stream
        }
        if (int rc = e->partial.ensure((size_t)e->grid * e->tpc * e->desc.n_par * 8)) return rc;
        if (e->pair_seg)
            if (int rc = e->xstore.ensure((size_t)e->grid * e->xstore_stride * 8)) return rc;
        e->workspace_bytes = (int64_t)(e->slab.bytes + e->partial.bytes + e->xstore.bytes);
        e->chunk_traj = B;
        return VA_OK;
    }
    // scalar: arena for min(B, budget) trajectories
    if (e->arena_traj >= B) {
        e->chunk_traj = B;
        return VA_OK;
    }
    size_t free_b = 0, total_b = 0;
    VA_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const double frac = e->desc.workspace_fraction > 0 ? e->desc.workspace_fraction : 0.5;
    const int64_t have = (int64_t)(e->ck_t.bytes + e->ck_x.bytes);
    const int64_t budget = (int64_t)(frac * (double)(free_b + have));
    const int64_t per = per_traj_arena_bytes(e);
    int64_t traj = std::min<int64_t>(B, std::max<int64_t>(1, budget / per));
    if (traj < B) traj = std::max<int64_t>(128, traj / 128 * 128); // whole CTAs per wave
    if (traj > e->arena_traj) {
        if (ck_layout_of(e) == 1) {
            if (int rc = e->ck_t.ensure((size_t)traj * (e->cap + 1) * ck_rec_doubles(e) * 8)) return rc;
        } else {
            if (int rc = e->ck_t.ensure((size_t)traj * (e->cap + 1) * 8)) return rc;
            if (int rc = e->ck_x.ensure((size_t)traj * (e->cap + 1) * e->desc.n_state * 8)) return rc;
        }
        e->arena_traj = traj;
    }
    e->workspace_bytes = (int64_t)(e->ck_t.bytes + e->ck_x.bytes);
    e->chunk_traj = std::min<int64_t>(B, e->arena_traj);
    return VA_OK;
}

// thread-per-trajectory kernels: ahead-of-time instantiations for the built-in systems, NVRTC module for recorded ones
int scalar_forward(va_engine *e, const VaScalarArgs &a_in, cudaStream_t st)
{
    if (int rc = e->work_counter.ensure(16)) return rc;
    VA_CUDA(cudaMemsetAsync(e->work_counter.p, 0, 16, st));
    VaScalarArgs a = a_in;
    a.ck_layout = ck_layout_of(e);
    a.work_counter = e->work_counter.as<unsigned long long>();
    a.grid_limit = e->sm_count * 16; // 16 x 128 threads = the most an SM can hold; surplus CTAs just find the queue empty
    if (e->family == FAM_TAPE) {
        if (int cr = va_jit_launch(e->jit, 0, a, a.B, st)) return fail(VA_E_CUDA, "cuLaunchKernel(forward) failed: " + std::to_string(cr));
        return VA_OK;
    }
    VA_CUDA(va_scalar_forward(a, st));
    return VA_OK;
}
int scalar_adjoint(va_engine *e, const VaScalarArgs &a_in, cudaStream_t st)
{
    if (int rc = e->work_counter.ensure(16)) return rc;
    VA_CUDA(cudaMemsetAsync(e->work_counter.as<unsigned long long>() + 1, 0, 8, st));
    VaScalarArgs a = a_in;
    a.ck_layout = ck_layout_of(e);
    a.work_counter = e->work_counter.as<unsigned long long>();
    a.grid_limit = e->sm_count * 16; // 16 x 128 threads = the most an SM can hold; surplus CTAs just find the queue empty
    if (e->family == FAM_TAPE) {
        if (int cr = va_jit_launch(e->jit, 1, a, a.B * a.n_out, st)) return fail(VA_E_CUDA, "cuLaunchKernel(adjoint) failed: " + std::to_string(cr));
        return VA_OK;
    }
    VA_CUDA(va_scalar_adjoint(a, st));
    return VA_OK;
}

struct DevArgs { // all pointers on the device
    int64_t B;
    const double *x0, *params;
    double ti, tf, dt0;
    int objective, reduce;
    double *x_final, *lambda, *mu; // mu: [B][nout][npar] or [nout][npar]
    int32_t *n_accept, *n_reject, *status;
    bool mu_accumulate; // VA_REDUCE_SUM: add to mu instead of overwriting (chunked host pipeline)
    bool forward_only;
    bool skip_forward = false; // split API: the slabs still hold the forward sweep of exactly these trajectories
};

// Runs forward (+ adjoint) for device-resident buffers on stream st.
int run_device(va_engine *e, const DevArgs &d, cudaStream_t st)
{
    const int n = e->desc.n_state, npar = e->desc.n_par, nout = e->desc.n_out;
    if (d.B <= 0) return VA_OK;
    if (int rc = ensure_workspace(e, d.B)) return rc;
    int32_t *acc = d.n_accept, *rej = d.n_reject, *sta = d.status;
    if (!is_glv(e)) { // the reverse kernel needs them
        if (!acc) { if (int rc = e->own_accept.ensure((size_t)d.B * 4)) return rc; acc = e->own_accept.as<int32_t>(); }
        if (!rej) { if (int rc = e->own_reject.ensure((size_t)d.B * 4)) return rc; rej = e->own_reject.as<int32_t>(); }
        if (!sta) { if (int rc = e->own_status.ensure((size_t)d.B * 4)) return rc; sta = e->own_status.as<int32_t>(); }
    }
    const bool sum = d.reduce == VA_REDUCE_SUM;

    if (is_glv(e)) {
        VaGlvWideArgs a;
        std::memset(&a, 0, sizeof(a));
        a.n = n; a.stepper = e->desc.stepper; a.adaptive = e->desc.adaptive; a.n_out = d.forward_only ? 0 : nout;
        a.objective = d.objective;
        a.eps_abs = e->desc.eps_abs; a.eps_rel = e->desc.eps_rel; a.ti = d.ti; a.tf = d.tf; a.dt0 = d.dt0;
        a.B = d.B; a.cap = e->cap; a.x0 = d.x0; a.params = d.params; a.x_final = d.x_final; a.lambda = d.lambda;
        a.n_accept = acc; a.n_reject = rej; a.status = sta;
        a.slab = e->slab.as<double>(); a.slab_stride = e->slab_stride; a.partial = e->partial.as<double>();
        a.grid = (int)std::min<int64_t>(e->grid, (d.B + e->tpc - 1) / e->tpc);
        if (e->pairk) a.grid = e->pair_cl * (int)std::min<int64_t>(e->grid / e->pair_cl, d.B); // pair_cl CTAs per trajectory
        a.cluster = e->pair_cl;
        if (e->pair_seg) {
            a.xstore = e->xstore.as<double>();
            a.xstore_stride = e->xstore_stride;
            a.seg_len = e->pair_seg_len;
            a.sparse = e->pair_sparse ? 1 : 0;
        }
        a.recompute = e->desc.ckpt_policy == VA_CKPT_RECOMPUTE || e->desc.ckpt_policy == VA_CKPT_SPARSE;
        a.blk_doubles = e->glv_blk;
        a.skip_forward = d.skip_forward ? 1 : 0;
        // VA_T8_DISCARD=1 (measured trade, off by default): with more trajectories than slots every slab is reused inside this launch
        // and serves no later call, so the headline kernel may drop its dead step blocks from L2 instead of letting them be written
        // back -- DRAM traffic 184 -> 128 GB per 2^20 sets (5.0 x -> 3.5 x the compulsory bytes) at 1.1 % fewer gradients/s
        if (e->family == FAM_GLV_WIDE && e->t8 && !e->t8s && d.B > (int64_t)e->grid * e->tpc && getenv("VA_T8_DISCARD") && atoi(getenv("VA_T8_DISCARD")) != 0)
            a.flags |= VA_GLV_FLAG_DISCARD;
        const bool native_sum = sum && nout == 1 && !d.forward_only;
        if (sum && !native_sum && !d.forward_only) {
            // several cost functions per trajectory: per-trajectory gradients into a scratch buffer, then a row reduction
            if (int rc = e->mu_tmp.ensure((size_t)d.B * nout * npar * 8)) return rc;
            a.reduce = VA_REDUCE_NONE;
            a.mu = e->mu_tmp.as<double>();
        } else {
            a.reduce = native_sum ? VA_REDUCE_SUM : VA_REDUCE_NONE;
            a.mu = d.mu;
        }
        if (e->family == FAM_GLV_WIDE && e->oct) VA_CUDA(va_glv_oct_forward_adjoint(a, st));
        else if (e->family == FAM_GLV_WIDE && e->quad) VA_CUDA(va_glv_quad_forward_adjoint(a, st));
        else if (e->family == FAM_GLV_WIDE && e->t8s) VA_CUDA(va_glv_t8s_forward_adjoint(a, st));
        else if (e->family == FAM_GLV_WIDE && e->t8) VA_CUDA(va_glv_t8_forward_adjoint(a, st));
        else if (e->family == FAM_GLV_WIDE) VA_CUDA(va_glv_wide_forward_adjoint(a, st));
        else if (e->pairk) VA_CUDA(va_glv_pair_forward_adjoint(a, st));
        else if (e->ring) {
            a.flags = e->ring_flags;
            VA_CUDA(va_glv_ring_forward_adjoint(a, st));
        } else VA_CUDA(va_glv_stream_forward_adjoint(a, st));
        ++e->launches;
        if (native_sum) {
            VA_CUDA(va_reduce_rows(e->partial.as<double>(), e->pairk ? a.grid / e->pair_cl : (int64_t)a.grid * e->tpc, npar, npar, d.mu,
                                   d.mu_accumulate ? 1 : 0, st));
            ++e->launches;
        } else if (sum && !d.forward_only) {
            VA_CUDA(va_reduce_rows(e->mu_tmp.as<double>(), d.B, (int64_t)nout * npar, (int64_t)nout * npar, d.mu,
                                   d.mu_accumulate ? 1 : 0, st));
            ++e->launches;
        }
        return VA_OK;
    }

    // scalar family: waves of arena_traj trajectories
    if (sum && !d.forward_only) {
        if (int rc = e->mu_tmp.ensure((size_t)std::min<int64_t>(d.B, e->arena_traj) * nout * npar * 8)) return rc;
    }
    bool first_wave = true;
    for (int64_t b0 = 0; b0 < d.B; b0 += e->arena_traj) {
        const int64_t Bw = std::min<int64_t>(e->arena_traj, d.B - b0);
        VaScalarArgs a;
        std::memset(&a, 0, sizeof(a));
        a.system = e->desc.system; a.stepper = e->desc.stepper; a.adaptive = e->desc.adaptive; a.n_out = nout;
        a.tab = e->tab; a.eps_abs = e->desc.eps_abs; a.eps_rel = e->desc.eps_rel; a.ti = d.ti; a.tf = d.tf; a.dt0 = d.dt0;
        a.B = Bw; a.arena_stride = e->arena_traj; a.cap = e->cap; a.objective = d.objective;
        a.x0 = d.x0 + b0 * n; a.params = d.params + b0 * npar; a.x_final = d.x_final + b0 * n;
        a.lambda = d.lambda ? d.lambda + b0 * nout * n : nullptr;
        a.mu = sum ? e->mu_tmp.as<double>() : (d.mu ? d.mu + b0 * nout * npar : nullptr);
        a.n_accept = acc + b0; a.n_reject = rej + b0; a.status = sta + b0;
        a.ck_t = e->ck_t.as<double>(); a.ck_x = e->ck_x.as<double>();
        if (int rc = scalar_forward(e, a, st)) return rc;
        ++e->launches;
        if (d.forward_only) continue;
        if (int rc = scalar_adjoint(e, a, st)) return rc;
        ++e->launches;
        if (sum) {
            VA_CUDA(va_reduce_rows(e->mu_tmp.as<double>(), Bw, (int64_t)nout * npar, (int64_t)nout * npar, d.mu,
                                   (d.mu_accumulate || !first_wave) ? 1 : 0, st));
            ++e->launches;
        }
        first_wave = false;
    }
    return VA_OK;
}

int check_args(const va_engine *e, const va_batch_args *a, bool need_adjoint)
{
    if (!e || !a) return fail(VA_E_INVALID, "null engine or args");
    if (a->batch < 0) return fail(VA_E_INVALID, "negative batch");
    if (!a->x0 || !a->params || !a->x_final) return fail(VA_E_INVALID, "x0, params and x_final are required");
    if (need_adjoint && (!a->lambda || !a->mu)) return fail(VA_E_INVALID, "lambda and mu are required (call setCostGradients first)");
    if (a->objective < VA_OBJ_SEED || a->objective > VA_OBJ_HALF_NORM2) return fail(VA_E_INVALID, "bad objective");
    if (a->reduce != VA_REDUCE_NONE && a->reduce != VA_REDUCE_SUM) return fail(VA_E_INVALID, "bad reduce");
    if (a->mem != VA_MEM_HOST && a->mem != VA_MEM_DEVICE) return fail(VA_E_INVALID, "bad mem");
    if (a->dt0 == 0.0 || !(std::isfinite(a->ti) && std::isfinite(a->tf) && std::isfinite(a->dt0)))
        return fail(VA_E_INVALID, "ti, tf, dt0 must be finite and dt0 non-zero");
    return VA_OK;
}

// Host buffers: chunked 3-stream pipeline.
int run_host(va_engine *e, const va_batch_args *a, bool forward_only)
{
    const int n = e->desc.n_state, npar = e->desc.n_par, nout = e->desc.n_out;
    const int64_t B = a->batch;
    const bool sum = a->reduce == VA_REDUCE_SUM;
    const int64_t in_bytes = 8LL * (n + npar) + (a->objective == VA_OBJ_SEED ? 8LL * nout * n : 0);
    const int64_t out_bytes = 8LL * n + 8LL * nout * n + (sum ? 0 : 8LL * nout * npar) + 12;
    // Chunk size of the pipeline. The copies are the bottleneck (PCIe: 55.6 GB/s against 200+ GB/s of kernel appetite), so what
    // matters is how soon the first kernel starts and how little is left to compute after the last copy: 512 MB chunks leave
    // 9 ms + 3 ms outside the copy stream (2 GB chunks: 36 ms + 10 ms of a 655 ms pass). VA_HOST_CHUNK_MB overrides.
    const char *chunk_env = getenv("VA_HOST_CHUNK_MB");
    const int64_t slot_budget = (chunk_env && atoll(chunk_env) > 0 ? atoll(chunk_env) : 512LL) << 20;
    int64_t Bc = std::max<int64_t>(1, slot_budget / (in_bytes + out_bytes));
    if (is_glv(e) && Bc > (int64_t)e->grid * e->tpc) Bc = Bc / (e->grid * e->tpc) * (e->grid * e->tpc); // whole waves
    Bc = std::min(Bc, B);
    if (B == 0) return VA_OK;
    for (int s = 0; s < 2; ++s) {
        if (int rc = e->st_x0[s].ensure((size_t)Bc * n * 8)) return rc;
        if (int rc = e->st_par[s].ensure((size_t)Bc * npar * 8)) return rc;
        if (int rc = e->st_xf[s].ensure((size_t)Bc * n * 8)) return rc;
        if (int rc = e->st_lam[s].ensure((size_t)Bc * nout * n * 8)) return rc;
        if (!sum && !forward_only)
            if (int rc = e->st_mu[s].ensure((size_t)Bc * nout * npar * 8)) return rc;
        if (int rc = e->st_acc[s].ensure((size_t)Bc * 4)) return rc;
        if (int rc = e->st_rej[s].ensure((size_t)Bc * 4)) return rc;
        if (int rc = e->st_sta[s].ensure((size_t)Bc * 4)) return rc;
    }
    if (sum && !forward_only) {
        if (int rc = e->st_musum.ensure((size_t)nout * npar * 8)) return rc;
        VA_CUDA(cudaMemsetAsync(e->st_musum.p, 0, (size_t)nout * npar * 8, e->s_comp));
    }
    VA_CUDA(cudaEventRecord(e->ev_t0, e->s_comp));
    int64_t c = 0;
    for (int64_t b0 = 0; b0 < B; b0 += Bc, ++c) {
        const int s = (int)(c & 1);
        const int64_t Bn = std::min(Bc, B - b0);
        // inputs of this slot are free once the kernels of chunk c-2 are done
        if (c >= 2) {
            VA_CUDA(cudaStreamWaitEvent(e->s_in, e->ev_comp[s], 0));
            VA_CUDA(cudaStreamWaitEvent(e->s_in, e->ev_out[s], 0)); // st_lam[s] is seeds in AND dJ/dx(t0) out: drain chunk c-2 first
        }
        VA_CUDA(cudaMemcpyAsync(e->st_x0[s].p, a->x0 + b0 * n, (size_t)Bn * n * 8, cudaMemcpyHostToDevice, e->s_in));
        VA_CUDA(cudaMemcpyAsync(e->st_par[s].p, a->params + b0 * npar, (size_t)Bn * npar * 8, cudaMemcpyHostToDevice, e->s_in));
        if (!forward_only && a->objective == VA_OBJ_SEED)
            VA_CUDA(cudaMemcpyAsync(e->st_lam[s].p, a->lambda + b0 * nout * n, (size_t)Bn * nout * n * 8, cudaMemcpyHostToDevice, e->s_in));
        VA_CUDA(cudaEventRecord(e->ev_in[s], e->s_in));
        // compute: needs the inputs, and the outputs of this slot drained (chunk c-2)
        VA_CUDA(cudaStreamWaitEvent(e->s_comp, e->ev_in[s], 0));
        if (c >= 2) VA_CUDA(cudaStreamWaitEvent(e->s_comp, e->ev_out[s], 0));
        DevArgs d;
        d.B = Bn; d.x0 = e->st_x0[s].as<double>(); d.params = e->st_par[s].as<double>();
        d.ti = a->ti; d.tf = a->tf; d.dt0 = a->dt0; d.objective = a->objective; d.reduce = a->reduce;
        d.x_final = e->st_xf[s].as<double>(); d.lambda = e->st_lam[s].as<double>();
        d.mu = sum ? e->st_musum.as<double>() : e->st_mu[s].as<double>();
        d.n_accept = e->st_acc[s].as<int32_t>(); d.n_reject = e->st_rej[s].as<int32_t>(); d.status = e->st_sta[s].as<int32_t>();
        d.mu_accumulate = true; d.forward_only = forward_only;
        if (int rc = run_device(e, d, e->s_comp)) return rc;
        VA_CUDA(cudaEventRecord(e->ev_comp[s], e->s_comp));
        // outputs
        VA_CUDA(cudaStreamWaitEvent(e->s_out, e->ev_comp[s], 0));
        VA_CUDA(cudaMemcpyAsync(a->x_final + b0 * n, e->st_xf[s].p, (size_t)Bn * n * 8, cudaMemcpyDeviceToHost, e->s_out));
        if (!forward_only) {
            VA_CUDA(cudaMemcpyAsync(a->lambda + b0 * nout * n, e->st_lam[s].p, (size_t)Bn * nout * n * 8, cudaMemcpyDeviceToHost, e->s_out));
            if (!sum)
                VA_CUDA(cudaMemcpyAsync(a->mu + b0 * nout * npar, e->st_mu[s].p, (size_t)Bn * nout * npar * 8, cudaMemcpyDeviceToHost, e->s_out));
        }
        if (a->n_accept) VA_CUDA(cudaMemcpyAsync(a->n_accept + b0, e->st_acc[s].p, (size_t)Bn * 4, cudaMemcpyDeviceToHost, e->s_out));
        if (a->n_reject) VA_CUDA(cudaMemcpyAsync(a->n_reject + b0, e->st_rej[s].p, (size_t)Bn * 4, cudaMemcpyDeviceToHost, e->s_out));
        if (a->status) VA_CUDA(cudaMemcpyAsync(a->status + b0, e->st_sta[s].p, (size_t)Bn * 4, cudaMemcpyDeviceToHost, e->s_out));
        VA_CUDA(cudaEventRecord(e->ev_out[s], e->s_out));
    }
    if (sum && !forward_only && e->comm) // several GPUs: the one collective of the path, on the sum of this GPU's chunks
        if (int rc = va_comm_allreduce_sum(e, e->st_musum.as<double>(), (int64_t)nout * npar, e->s_comp)) return rc;
    VA_CUDA(cudaEventRecord(e->ev_t1, e->s_comp));
    if (sum && !forward_only && e->mu_host_writer)
        VA_CUDA(cudaMemcpyAsync(a->mu, e->st_musum.p, (size_t)nout * npar * 8, cudaMemcpyDeviceToHost, e->s_comp));
    VA_CUDA(cudaStreamSynchronize(e->s_in));
    VA_CUDA(cudaStreamSynchronize(e->s_comp));
    VA_CUDA(cudaStreamSynchronize(e->s_out));
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1);
    e->last_ms = ms;
    return VA_OK;
}

int run_call(va_engine *e, const va_batch_args *a, bool forward_only)
{
    VA_CUDA(cudaSetDevice(e->device));
    // a fused call reuses the checkpoint arena / slabs: a split-API session recorded earlier on this engine is gone
    if (is_glv(e)) { e->se_slab_valid = false; e->se_ck_b = -1; }
    else e->se_B = 0;
    if (a->mem == VA_MEM_HOST) return run_host(e, a, forward_only);
    cudaStream_t st = a->stream ? static_cast<cudaStream_t>(a->stream) : e->s_comp;
    DevArgs d;
    d.B = a->batch; d.x0 = a->x0; d.params = a->params; d.ti = a->ti; d.tf = a->tf; d.dt0 = a->dt0;
    d.objective = a->objective; d.reduce = a->reduce; d.x_final = a->x_final; d.lambda = a->lambda; d.mu = a->mu;
    d.n_accept = a->n_accept; d.n_reject = a->n_reject; d.status = a->status; d.mu_accumulate = false;
    d.forward_only = forward_only;
    VA_CUDA(cudaEventRecord(e->ev_t0, st));
    if (int rc = run_device(e, d, st)) return rc;
    if (a->reduce == VA_REDUCE_SUM && !forward_only && e->comm)
        if (int rc = va_comm_allreduce_sum(e, d.mu, (int64_t)e->desc.n_out * e->desc.n_par, st)) return rc;
    VA_CUDA(cudaEventRecord(e->ev_t1, st));
    if (!a->stream) {
        VA_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1);
        e->last_ms = ms;
    }
    return VA_OK;
}

} // namespace

int va_single_create(const va_engine_desc *desc, va_engine **out)
{
    if (!desc || !out) return fail(VA_E_INVALID, "null argument");
    *out = nullptr;
    if (desc->n_state < 1 || desc->n_par < 1 || desc->n_out < 1) return fail(VA_E_INVALID, "n_state, n_par, n_out must be >= 1");
    VaTableau tab;
    if (va_tableau_host(desc->stepper, &tab) != 0)
        return fail(VA_E_UNSUPPORTED, "This stepper is not supported yet!"); // ButcherTable.hpp:121-124, 247-250 (no terminate here)
    if (desc->adaptive && !tab.has_error)
        return fail(VA_E_INVALID, "a controlled stepper needs an error stepper (cash_karp54, dopri5, fehlberg78)");
    if (desc->adaptive && !(desc->eps_abs >= 0 && desc->eps_rel >= 0 && desc->eps_abs + desc->eps_rel > 0))
        return fail(VA_E_INVALID, "tolerances must be non-negative and not both zero");
    int family;
    switch (desc->system) {
    case VA_SYS_HARMONIC:
    case VA_SYS_VANDERPOL:
        if (desc->n_state != 2 || desc->n_par != 1) return fail(VA_E_INVALID, "harmonic / van der pol: n_state = 2, n_par = 1");
        family = FAM_SCALAR;
        break;
    case VA_SYS_GLV:
        if (desc->n_par != desc->n_state * desc->n_state + desc->n_state) return fail(VA_E_INVALID, "GLV: n_par must be N*N + N");
        if (va_glv_wide_supported(desc->n_state, desc->stepper, desc->adaptive) && !getenv("VA_GLV_FORCE_STREAM") &&
            ((desc->ckpt_policy != VA_CKPT_RECOMPUTE && desc->ckpt_policy != VA_CKPT_SPARSE) ||
             (va_glv_oct_supported(desc->n_state, desc->stepper, desc->adaptive) && !getenv("VA_GLV_NO_OCT") && !getenv("VA_GLV_V1"))))
            family = FAM_GLV_WIDE; // N <= 64: matrix in registers (store-stages policy; up to 16 species also recompute)
        else if (va_glv_stream_supported(desc->n_state, desc->stepper, desc->adaptive))
            family = FAM_GLV_STREAM; // any N: matrix streamed from L2/HBM
        else
            return fail(VA_E_UNSUPPORTED, "GLV: this species count / stepper does not fit the streamed kernel's shared memory (euler, rk4 fixed step; cash_karp54, dopri5, fehlberg78 fixed step or controlled)");
        break;
    case VA_SYS_TAPE:
        if (!desc->tape_cuda_src) return fail(VA_E_INVALID, "VA_SYS_TAPE needs tape_cuda_src (va::Tape::cuda_source(\"VaUserSys\"))");
        // thread-per-trajectory kernels around the generated functor: state, stage slopes and stage adjoints are per-lane arrays
        // (registers, spilling to local memory when wide), parameters beyond VA_REG_PARAMS are read in place. The bound below
        // is the per-lane local-memory footprint ((2 s + 6) n doubles), not a limit of the tape.
        if (desc->n_state > 64) return fail(VA_E_UNSUPPORTED, "recorded systems run on the thread-per-trajectory kernels: n_state <= 64");
        family = FAM_TAPE;
        break;
    default:
        return fail(VA_E_UNSUPPORTED, "system kind not available in this build");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(VA_E_CUDA, "no CUDA device: this engine has no CPU fallback");
    }
    if (desc->device < 0 || desc->device >= ndev) return fail(VA_E_INVALID, "bad device ordinal");
    va_engine *e = new (std::nothrow) va_engine();
    if (!e) return fail(VA_E_NOMEM, "out of host memory");
    e->desc = *desc;
    e->tab = tab;
    e->family = family;
    e->device = desc->device;
    e->cap = desc->max_steps > 0 ? desc->max_steps : (family == FAM_GLV_WIDE || family == FAM_GLV_STREAM ? 256 : 2048);
    auto bail = [&](int code, const std::string &m) { va_single_destroy(e); return fail(code, m); };
    if (cudaSetDevice(e->device) != cudaSuccess) return bail(VA_E_CUDA, "cudaSetDevice failed");
    cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, e->device);
    bool ok = cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&e->ev_t0) == cudaSuccess && cudaEventCreate(&e->ev_t1) == cudaSuccess;
    for (int s = 0; s < 2 && ok; ++s)
        ok = cudaEventCreateWithFlags(&e->ev_in[s], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&e->ev_comp[s], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&e->ev_out[s], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) return bail(VA_E_CUDA, "stream/event creation failed");
    if (family == FAM_GLV_STREAM) {
        e->ctas_per_sm = 2;
        e->grid = e->sm_count * e->ctas_per_sm;
        e->threads = 256;
        e->tpc = 1;
        // checkpoint policy (north_star item 4): STORE_STAGES unless the caller asks for RECOMPUTE, or (AUTO) the slabs of
        // all resident CTAs at the requested capacity would take more than the workspace share of free HBM
        // 65..256 species: the cluster kernel (va_glv_pair.cu, matrix on chip) under either policy; VA_GLV_NO_PAIR selects the
        // ring-streamed kernel (256 species, store-stages only), VA_GLV_NO_RING the plain streamed kernel
        e->pairk = va_glv_pair_supported(desc->n_state, desc->stepper, desc->adaptive) && !getenv("VA_GLV_NO_PAIR") &&
                   !getenv("VA_GLV_NO_RING") && e->sm_count >= 2;
        int policy = desc->ckpt_policy;
        if (policy == VA_CKPT_AUTO) {
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            const double frac = desc->workspace_fraction > 0 ? desc->workspace_fraction : 0.5;
            const double need = e->pairk ? (double)e->sm_count * (e->cap + 1) * va_glv_pair_block_doubles(desc->stepper) * 8.0
                                         : (double)e->grid * (e->cap + 1) * va_glv_stream_block_doubles(desc->n_state, desc->stepper, 0) * 8.0;
            policy = need > frac * (double)free_b ? VA_CKPT_RECOMPUTE : VA_CKPT_STORE_STAGES;
            // ... and when even one state per accepted step is too much (2.1 KB x capacity x 148 CTAs), every L-th state only
            if (policy == VA_CKPT_RECOMPUTE && e->pairk &&
                (double)e->sm_count * (e->cap + 1) * (8 + 256) * 8.0 > frac * (double)free_b)
                policy = VA_CKPT_SPARSE;
        }
        if (policy == VA_CKPT_SPARSE && !e->pairk) policy = VA_CKPT_RECOMPUTE; // only the cluster kernel thins its state store
        e->desc.ckpt_policy = policy;
        e->slab_stride = (int64_t)(e->cap + 1) * va_glv_stream_block_doubles(desc->n_state, desc->stepper, policy == VA_CKPT_RECOMPUTE);
        e->ring = !e->pairk && policy == VA_CKPT_STORE_STAGES && va_glv_ring_supported(desc->n_state, desc->stepper, desc->adaptive) &&
                  !getenv("VA_GLV_NO_RING");
        if (e->pairk) {
            e->ctas_per_sm = 1;
            e->pair_cl = getenv("VA_GLV_CLUSTER") && atoi(getenv("VA_GLV_CLUSTER")) == 4 ? 4 : 2;
            int nclusters = 0;
            cudaError_t ce = va_glv_pair_max_clusters(desc->stepper, e->pair_cl, e->sm_count, &nclusters);
            if (ce != cudaSuccess || nclusters < 1)
                return bail(VA_E_CUDA, std::string("cluster occupancy query failed: ") + cudaGetErrorString(ce));
            e->grid = e->pair_cl * std::min(nclusters, e->sm_count / e->pair_cl);
            if (getenv("VA_DEBUG")) fprintf(stderr, "va: k_glv_pair: %d clusters of %d CTAs\n", e->grid / e->pair_cl, e->pair_cl);
            e->glv_blk = va_glv_pair_block_doubles(desc->stepper);
            e->slab_stride = (int64_t)(e->cap + 1) * e->glv_blk;
            e->pair_seg = policy == VA_CKPT_RECOMPUTE || policy == VA_CKPT_SPARSE;
            e->pair_sparse = policy == VA_CKPT_SPARSE;
            if (e->pair_seg) {
                // recompute policy: (t_n, x_n) of every accepted step in a per-CTA state store, stage blocks only for one segment
                // (16 steps x 36.9 KB x 148 CTAs = 87 MB: L2-resident); VA_PAIR_SEG overrides the segment length
                const int sl = getenv("VA_PAIR_SEG") ? atoi(getenv("VA_PAIR_SEG")) : 16;
                e->pair_seg_len = sl < 1 ? 1 : sl;
                e->slab_stride = (int64_t)(e->pair_seg_len + 1) * e->glv_blk;
                e->xstore_stride = (int64_t)(e->cap + 1) * (8 + 256);
                if (e->pair_sparse) // [t_0 .. t_cap+1 | pad | one state per segment]  (va_glv_pair.cu)
                    e->xstore_stride = ((int64_t)e->cap + 2 + 7) / 8 * 8 + ((int64_t)e->cap / e->pair_seg_len + 1) * 256;
            }
        }
        if (e->ring) {
            e->ctas_per_sm = 1;
            e->grid = e->sm_count;
            e->glv_blk = va_glv_ring_block_doubles(desc->stepper);
            e->slab_stride = (int64_t)(e->cap + 1) * e->glv_blk;
            // the matrices of all resident CTAs (148 x 512 KB) are re-read for every product: keep them in L2 (evict_last class,
            // persisting carve-out at the device maximum); VA_RING_FLAGS overrides (bit 1: evict_last, bit 2: no cached rows)
            e->ring_flags = getenv("VA_RING_FLAGS") ? atoi(getenv("VA_RING_FLAGS")) : 2;
            if (e->ring_flags & 2) {
                int max_persist = 0;
                if (cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, e->device) == cudaSuccess && max_persist > 0)
                    persist_l2_acquire(e, (size_t)max_persist);
            }
        }
    } else if (family == FAM_GLV_WIDE) {
        // 33..64 species: second-generation kernel (va_glv_t8.cu); VA_GLV_V1 keeps the first generation for cross-checks
        e->t8 = va_glv_t8_supported(desc->n_state, desc->stepper, desc->adaptive) && !getenv("VA_GLV_V1");
        // up to 16 species: quad kernel (va_glv_quad.cu); VA_GLV_NO_QUAD keeps the first-generation kernel for cross-checks
        e->quad = va_glv_quad_supported(desc->n_state, desc->stepper, desc->adaptive) && !getenv("VA_GLV_NO_QUAD") && !getenv("VA_GLV_V1");
        // up to 16 species, second generation (va_glv_oct.cu): the default unless the caller insists on the store-stages policy;
        // VA_GLV_NO_OCT keeps the four-lane store-stages kernel for cross-checks
        e->oct = va_glv_oct_supported(desc->n_state, desc->stepper, desc->adaptive) && desc->ckpt_policy != VA_CKPT_STORE_STAGES &&
                 !getenv("VA_GLV_NO_OCT") && !getenv("VA_GLV_V1");
        if (e->oct) e->quad = false;
        cudaError_t ce;
        if (e->oct) {
            ce = cudaSuccess;
            e->ctas_per_sm = 1;
            e->grid = e->sm_count;
            e->threads = va_glv_oct_threads();
            e->tpc = va_glv_oct_slots_per_cta();
            e->pair = 1;
            e->glv_blk = va_glv_oct_block_doubles();
            e->slab_stride = (int64_t)(e->cap + 1) * e->glv_blk;
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                const double frac = desc->workspace_fraction > 0 ? desc->workspace_fraction : 0.5;
                while (e->grid > 1 && (double)e->grid * e->tpc * e->slab_stride * 8.0 > frac * (double)free_b) e->grid = (e->grid + 1) / 2;
            }
        } else if (e->quad) {
            ce = cudaSuccess;
            e->ctas_per_sm = 1;
            e->grid = e->sm_count;
            e->threads = va_glv_quad_threads();
            e->tpc = va_glv_quad_slots_per_cta();
            e->pair = 1;
            e->glv_blk = va_glv_quad_block_doubles(desc->stepper, desc->n_out);
            e->slab_stride = (int64_t)(e->cap + 1) * e->glv_blk;
            // 64 slabs per CTA: a large step capacity (max_steps) must not exhaust HBM -- run fewer CTAs instead
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                const double frac = desc->workspace_fraction > 0 ? desc->workspace_fraction : 0.5;
                while (e->grid > 1 && (double)e->grid * e->tpc * e->slab_stride * 8.0 > frac * (double)free_b) e->grid = (e->grid + 1) / 2;
            }
        } else if (e->t8) {
            // third generation (va_glv_t8s.cu); VA_GLV_T8S=1 selects it
            e->t8s = va_glv_t8s_supported(desc->n_state, desc->stepper, desc->adaptive) && getenv("VA_GLV_T8S") && atoi(getenv("VA_GLV_T8S")) != 0;
            if (e->t8s && va_glv_t8s_config(desc->n_state, desc->stepper, desc->n_out, e->device, &e->grid, &e->ctas_per_sm, &e->threads, &e->tpc) != cudaSuccess) {
                cudaGetLastError();
                e->t8s = false; // step blocks too large for its shared-memory rings (several seeds per trajectory on a long tableau)
            }
            ce = e->t8s ? cudaSuccess
                        : va_glv_t8_config(desc->n_state, desc->stepper, desc->n_out, e->device, &e->grid, &e->ctas_per_sm, &e->threads, &e->tpc);
            e->pair = 1;
            e->glv_blk = va_glv_t8_block_doubles(desc->stepper, desc->n_out);
            e->glv_hdr = va_glv_t8_header_doubles();
            // the kernel marks its checkpoint slabs evict_last: give that class the largest L2 share the device allows
            int max_persist = 0;
            if (cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, e->device) == cudaSuccess && max_persist > 0)
                persist_l2_acquire(e, (size_t)max_persist);
            if (getenv("VA_DEBUG")) fprintf(stderr, "va: persisting L2 limit %d bytes\n", max_persist);
            e->slab_stride = e->t8s ? va_glv_t8s_slab_doubles(desc->stepper, desc->n_out, e->cap) : (int64_t)(e->cap + 1) * e->glv_blk;
        } else {
            ce = va_glv_wide_config(desc->n_state, desc->stepper, e->device, &e->grid, &e->ctas_per_sm, &e->threads, &e->tpc);
            e->pair = va_glv_wide_pair();
            e->glv_blk = va_glv_wide_block_doubles(desc->n_state, desc->stepper);
            e->slab_stride = va_glv_wide_slab_doubles(desc->n_state, desc->stepper, e->cap);
        }
        if (ce != cudaSuccess) return bail(VA_E_CUDA, std::string("kernel configuration failed: ") + cudaGetErrorString(ce));
        // test knob: fewer persistent CTAs, so that a small batch (compute-sanitizer runs) gives every slot several trajectories
        if (const char *env = getenv("VA_GLV_MAX_CTAS"))
            if (atoi(env) >= 1) e->grid = std::min(e->grid, atoi(env));
        e->desc.ckpt_policy = e->oct ? VA_CKPT_RECOMPUTE : VA_CKPT_STORE_STAGES;
    } else {
        e->threads = 128;
        e->desc.ckpt_policy = VA_CKPT_RECOMPUTE;
    }
    if (family == FAM_TAPE) {
        cudaFree(0); // make the primary context current for the driver API
        std::string log;
        const int rc = va_jit_compile(desc->tape_cuda_src, tab.s, tab.fsal, tab.s_adj, &e->jit, log);
        if (rc != 0) return bail(VA_E_NVRTC, "tape -> CUDA compilation failed: " + log);
        e->desc.tape_cuda_src = nullptr; // not kept
    }
    *out = e;
    return VA_OK;
}

void va_single_destroy(va_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    if (!e->comm_owned_by_head) va_comm_release(e);
    persist_l2_release(e);
    DevBuf *bufs[] = {&e->slab, &e->partial, &e->xstore, &e->ck_t, &e->ck_x, &e->work_counter, &e->own_accept, &e->own_reject, &e->own_status, &e->mu_tmp,
                      &e->st_musum, &e->se_x0, &e->se_par, &e->se_xf, &e->se_lam, &e->se_mu, &e->se_acc, &e->se_rej, &e->se_sta};
    for (DevBuf *b : bufs) b->release();
    for (int s = 0; s < 2; ++s) {
        DevBuf *sb[] = {&e->st_x0[s], &e->st_par[s], &e->st_xf[s], &e->st_lam[s], &e->st_mu[s], &e->st_acc[s], &e->st_rej[s], &e->st_sta[s]};
        for (DevBuf *b : sb) b->release();
        if (e->ev_in[s]) cudaEventDestroy(e->ev_in[s]);
        if (e->ev_comp[s]) cudaEventDestroy(e->ev_comp[s]);
        if (e->ev_out[s]) cudaEventDestroy(e->ev_out[s]);
    }
    if (e->ev_t0) cudaEventDestroy(e->ev_t0);
    if (e->ev_t1) cudaEventDestroy(e->ev_t1);
    va_jit_destroy(e->jit);
    if (e->s_comp) cudaStreamDestroy(e->s_comp);
    if (e->s_in) cudaStreamDestroy(e->s_in);
    if (e->s_out) cudaStreamDestroy(e->s_out);
    delete e;
}

static int single_get_info(va_engine *e, va_engine_info *info)
{
    std::memset(info, 0, sizeof(*info));
    info->api_version = VA_API_VERSION;
    info->device = e->device;
    info->sm_count = e->sm_count;
    info->kernel_family = e->family;
    info->ckpt_policy = e->desc.ckpt_policy;
    info->max_steps = e->cap;
    info->ctas_per_sm = e->ctas_per_sm;
    info->threads_per_cta = e->threads;
    info->workspace_bytes = e->workspace_bytes;
    info->chunk_trajectories = e->chunk_traj;
    info->kernel_launches = e->launches;
    info->last_kernel_ms = e->last_ms;
    const char *kn = e->family == FAM_SCALAR ? "k_scalar" : e->family == FAM_TAPE ? "jit" : e->family == FAM_GLV_WIDE ? (e->oct ? "k_glv_oct" : e->quad ? "k_glv_quad" : e->t8s ? "k_glv_t8s" : e->t8 ? "k_glv_t8" : "k_glv_wide")
                     : e->pairk ? "k_glv_pair" : e->ring ? "k_glv_ring" : "k_glv_stream";
    std::snprintf(info->kernel_name, sizeof(info->kernel_name), "%s", kn);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, e->device) == cudaSuccess) std::snprintf(info->device_name, sizeof(info->device_name), "%s", prop.name);
    info->n_devices = 1;
    info->comm_world = e->comm ? e->comm_world : 0;
    info->comm_rank = e->comm_rank;
    info->nccl_version = e->comm ? va_nccl_version() : 0;
    info->collectives = e->collectives;
    return VA_OK;
}

int va_single_forward_adjoint(va_engine *e, const va_batch_args *a)
{
    if (int rc = check_args(e, a, true)) return rc;
    if (a->mem == VA_MEM_DEVICE) {
        // the fused call hands the caller's device arrays to the kernels, which read and write them with 16-byte vector accesses
        // and bulk copies: a misaligned pointer would be a sticky "misaligned address" fault inside the kernel. (Host arrays, and
        // the arrays of the split API, are copied into the engine's own buffers: no constraint.)
        const void *ptrs[] = {a->x0, a->params, a->x_final, a->lambda, a->mu};
        for (const void *q : ptrs)
            if (reinterpret_cast<uintptr_t>(q) & 15)
                return fail(VA_E_INVALID, "device arrays must be 16-byte aligned (cudaMalloc / torch allocations are; pass an aligned copy of a sub-array)");
    }
    return run_call(e, a, false);
}

// Split API. The forward call keeps a device-side session (inputs, x(tf), step counts, checkpoints) so that
// va_adjoint_batch / va_get_checkpoints can serve Driver::GetT/GetTime/GetState and adjointSolve afterwards.
int va_single_forward(va_engine *e, const va_batch_args *a)
{
    if (int rc = check_args(e, a, false)) return rc;
    VA_CUDA(cudaSetDevice(e->device));
    const int n = e->desc.n_state, npar = e->desc.n_par;
    const int64_t B = a->batch;
    // GLV path: the slabs keep the checkpoints of the first wave of trajectories only (va_get_checkpoints serves those);
    // va_adjoint_batch does not need them, it re-integrates from the recorded inputs (deterministic, identical steps)
    if (int rc = e->se_x0.ensure((size_t)B * n * 8)) return rc;
    if (int rc = e->se_par.ensure((size_t)B * npar * 8)) return rc;
    if (int rc = e->se_xf.ensure((size_t)B * n * 8)) return rc;
    if (int rc = e->se_acc.ensure((size_t)B * 4)) return rc;
    if (int rc = e->se_rej.ensure((size_t)B * 4)) return rc;
    if (int rc = e->se_sta.ensure((size_t)B * 4)) return rc;
    const cudaMemcpyKind in_kind = a->mem == VA_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const cudaMemcpyKind out_kind = a->mem == VA_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    VA_CUDA(cudaMemcpyAsync(e->se_x0.p, a->x0, (size_t)B * n * 8, in_kind, e->s_comp));
    VA_CUDA(cudaMemcpyAsync(e->se_par.p, a->params, (size_t)B * npar * 8, in_kind, e->s_comp));
    if (!is_glv(e)) {
        if (int rc = ensure_workspace(e, B)) return rc;
        if (e->arena_traj < B) return fail(VA_E_NOMEM, "checkpoint arena too small for a split forward/adjoint of this batch; use va_forward_adjoint_batch");
    }
    DevArgs d;
    d.B = B; d.x0 = e->se_x0.as<double>(); d.params = e->se_par.as<double>(); d.ti = a->ti; d.tf = a->tf; d.dt0 = a->dt0;
    d.objective = VA_OBJ_SUM; d.reduce = VA_REDUCE_NONE; d.x_final = e->se_xf.as<double>(); d.lambda = nullptr; d.mu = nullptr;
    d.n_accept = e->se_acc.as<int32_t>(); d.n_reject = e->se_rej.as<int32_t>(); d.status = e->se_sta.as<int32_t>();
    d.mu_accumulate = false; d.forward_only = true;
    VA_CUDA(cudaEventRecord(e->ev_t0, e->s_comp));
    if (int rc = run_device(e, d, e->s_comp)) return rc;
    VA_CUDA(cudaEventRecord(e->ev_t1, e->s_comp));
    VA_CUDA(cudaMemcpyAsync(a->x_final, e->se_xf.p, (size_t)B * n * 8, out_kind, e->s_comp));
    if (a->n_accept) VA_CUDA(cudaMemcpyAsync(a->n_accept, e->se_acc.p, (size_t)B * 4, out_kind, e->s_comp));
    if (a->n_reject) VA_CUDA(cudaMemcpyAsync(a->n_reject, e->se_rej.p, (size_t)B * 4, out_kind, e->s_comp));
    if (a->status) VA_CUDA(cudaMemcpyAsync(a->status, e->se_sta.p, (size_t)B * 4, out_kind, e->s_comp));
    e->se_accept_host.resize((size_t)B);
    e->se_xf_host.resize((size_t)B * n);
    VA_CUDA(cudaMemcpyAsync(e->se_accept_host.data(), e->se_acc.p, (size_t)B * 4, cudaMemcpyDeviceToHost, e->s_comp));
    VA_CUDA(cudaMemcpyAsync(e->se_xf_host.data(), e->se_xf.p, (size_t)B * n * 8, cudaMemcpyDeviceToHost, e->s_comp));
    VA_CUDA(cudaStreamSynchronize(e->s_comp));
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1);
    e->last_ms = ms;
    e->se_B = B; e->se_ti = a->ti; e->se_tf = a->tf; e->se_dt0 = a->dt0;
    e->se_slab_valid = is_glv(e); e->se_ck_b = -1;
    e->se_blocks_intact = true;
    return VA_OK;
}

int va_single_adjoint(va_engine *e, const va_batch_args *a)
{
    if (!e || !a) return fail(VA_E_INVALID, "null engine or args");
    if (e->se_B <= 0) return fail(VA_E_STATE, "va_adjoint_batch needs a preceding va_forward_batch (runge_kutta) on this engine");
    if (a->batch != e->se_B) return fail(VA_E_INVALID, "batch differs from the preceding va_forward_batch");
    if (!a->lambda || !a->mu) return fail(VA_E_INVALID, "Must call setCostGradients() first!"); // backpropagation.hpp:22-26
    if (a->reduce != VA_REDUCE_NONE && a->reduce != VA_REDUCE_SUM) return fail(VA_E_INVALID, "bad reduce");
    if (a->objective < VA_OBJ_SEED || a->objective > VA_OBJ_HALF_NORM2) return fail(VA_E_INVALID, "bad objective");
    if (a->mem != VA_MEM_HOST && a->mem != VA_MEM_DEVICE) return fail(VA_E_INVALID, "bad mem");
    VA_CUDA(cudaSetDevice(e->device));
    const int n = e->desc.n_state, npar = e->desc.n_par, nout = e->desc.n_out;
    const int64_t B = e->se_B;
    const bool sum = a->reduce == VA_REDUCE_SUM;
    const size_t mu_elems = sum ? (size_t)nout * npar : (size_t)B * nout * npar;
    if (int rc = e->se_lam.ensure((size_t)B * nout * n * 8)) return rc;
    if (int rc = e->se_mu.ensure(mu_elems * 8)) return rc;
    const cudaMemcpyKind in_kind = a->mem == VA_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const cudaMemcpyKind out_kind = a->mem == VA_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (a->objective == VA_OBJ_SEED)
        VA_CUDA(cudaMemcpyAsync(e->se_lam.p, a->lambda, (size_t)B * nout * n * 8, in_kind, e->s_comp));
    VA_CUDA(cudaEventRecord(e->ev_t0, e->s_comp));
    if (!is_glv(e)) {
        // reverse kernel over the checkpoints of the session
        VaScalarArgs s;
        std::memset(&s, 0, sizeof(s));
        s.system = e->desc.system; s.stepper = e->desc.stepper; s.adaptive = e->desc.adaptive; s.n_out = nout; s.tab = e->tab;
        s.B = B; s.arena_stride = e->arena_traj; s.cap = e->cap; s.objective = a->objective;
        s.params = e->se_par.as<double>(); s.x_final = e->se_xf.as<double>(); s.lambda = e->se_lam.as<double>();
        s.n_accept = e->se_acc.as<int32_t>(); s.n_reject = e->se_rej.as<int32_t>(); s.status = e->se_sta.as<int32_t>();
        s.ck_t = e->ck_t.as<double>(); s.ck_x = e->ck_x.as<double>();
        if (sum) {
            if (int rc = e->mu_tmp.ensure((size_t)B * nout * npar * 8)) return rc;
            s.mu = e->mu_tmp.as<double>();
        } else {
            s.mu = e->se_mu.as<double>();
        }
        if (int rc = scalar_adjoint(e, s, e->s_comp)) return rc;
        ++e->launches;
        if (sum) {
            VA_CUDA(va_reduce_rows(e->mu_tmp.as<double>(), B, (int64_t)nout * npar, (int64_t)nout * npar, e->se_mu.as<double>(), 0, e->s_comp));
            ++e->launches;
        }
    } else {
        // GLV path: the fused kernel re-integrates (deterministic, identical checkpoints) and sweeps back
        DevArgs d;
        d.B = B; d.x0 = e->se_x0.as<double>(); d.params = e->se_par.as<double>(); d.ti = e->se_ti; d.tf = e->se_tf; d.dt0 = e->se_dt0;
        d.objective = a->objective; d.reduce = a->reduce; d.x_final = e->se_xf.as<double>(); d.lambda = e->se_lam.as<double>();
        d.mu = e->se_mu.as<double>(); d.n_accept = e->se_acc.as<int32_t>(); d.n_reject = e->se_rej.as<int32_t>();
        d.status = e->se_sta.as<int32_t>(); d.mu_accumulate = false; d.forward_only = false;
        // runge_kutta then adjointSolve on a batch that fitted one wave of slots (always so for the one-trajectory Driver): the
        // step blocks / checkpoints of the forward call are still in the slabs, slot b = trajectory b -- sweep back only.
        // Otherwise the fused kernel re-integrates (deterministic, identical steps).
        const int64_t slots = e->pairk ? e->grid / e->pair_cl : (int64_t)e->grid * e->tpc;
        d.skip_forward = e->se_slab_valid && e->se_blocks_intact && B <= slots && e->family == FAM_GLV_WIDE && (e->oct || e->t8) &&
                         !getenv("VA_SPLIT_REINTEGRATE");
        if (int rc = run_device(e, d, e->s_comp)) return rc;
        if (!d.skip_forward) { e->se_ck_b = -1; e->se_slab_valid = B <= slots; } // re-integrated: the same first wave is back in the slabs
        // the store-stages kernels write the seeds v_m over the slopes g_m during the reverse sweep (one seed per trajectory: the two
        // sections alias): the blocks serve GetTime / GetState afterwards, but not another reverse sweep
        if (!e->oct) e->se_blocks_intact = false;
    }
    if (sum && e->comm)
        if (int rc = va_comm_allreduce_sum(e, e->se_mu.as<double>(), (int64_t)nout * npar, e->s_comp)) return rc;
    VA_CUDA(cudaEventRecord(e->ev_t1, e->s_comp));
    VA_CUDA(cudaMemcpyAsync(a->lambda, e->se_lam.p, (size_t)B * nout * n * 8, out_kind, e->s_comp));
    if (!sum || e->mu_host_writer || a->mem == VA_MEM_DEVICE)
        VA_CUDA(cudaMemcpyAsync(a->mu, e->se_mu.p, mu_elems * 8, out_kind, e->s_comp));
    VA_CUDA(cudaStreamSynchronize(e->s_comp));
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1);
    e->last_ms = ms;
    return VA_OK;
}

int va_single_get_checkpoints(va_engine *e, int64_t b, int32_t capacity, double *t, double *x, int32_t *count)
{
    if (!e || !count) return fail(VA_E_INVALID, "null argument");
    if (e->se_B <= 0) return fail(VA_E_STATE, "no forward sweep recorded on this engine");
    if (b < 0 || b >= e->se_B) return fail(VA_E_INVALID, "trajectory index out of range");
    // GLV: a slot's slab holds the LAST trajectory it integrated. Trajectory b sits in slot b only while the batch fitted one
    // wave of slots and nothing ran since; otherwise it is re-integrated alone (forward only, deterministic) into slot 0.
    int64_t slot = b;
    if (is_glv(e)) {
        const int64_t slots = e->pairk ? e->grid / e->pair_cl : (int64_t)e->grid * e->tpc;
        if (!(e->se_slab_valid && e->se_B <= slots)) {
            VA_CUDA(cudaSetDevice(e->device));
            if (e->se_ck_b != b) {
                const int n_ = e->desc.n_state, npar_ = e->desc.n_par;
                if (int rc = e->se_scratch.ensure((size_t)n_ * 8 + 64)) return rc;
                DevArgs d;
                d.B = 1; d.x0 = e->se_x0.as<double>() + b * n_; d.params = e->se_par.as<double>() + b * npar_;
                d.ti = e->se_ti; d.tf = e->se_tf; d.dt0 = e->se_dt0; d.objective = VA_OBJ_SUM; d.reduce = VA_REDUCE_NONE;
                d.x_final = e->se_scratch.as<double>(); d.lambda = nullptr; d.mu = nullptr;
                int32_t *ints = reinterpret_cast<int32_t *>(e->se_scratch.as<double>() + n_);
                d.n_accept = ints; d.n_reject = ints + 1; d.status = ints + 2; d.mu_accumulate = false; d.forward_only = true;
                if (int rc = run_device(e, d, e->s_comp)) return rc;
                VA_CUDA(cudaStreamSynchronize(e->s_comp));
                e->se_slab_valid = false;
                e->se_ck_b = b;
            }
            slot = 0;
        }
    }
    const int n = e->desc.n_state;
    const int T = e->se_accept_host[(size_t)b];
    *count = T + 1;
    if (!t && !x) return VA_OK;
    if (capacity < T + 1) return fail(VA_E_INVALID, "capacity too small");
    VA_CUDA(cudaSetDevice(e->device));
    if (!is_glv(e)) {
        if (ck_layout_of(e) == 1) { // per-trajectory records {t, x}
            const double *rec = e->ck_t.as<double>() + b * (int64_t)(e->cap + 1) * ck_rec_doubles(e);
            const size_t pitch = (size_t)ck_rec_doubles(e) * 8;
            if (t) VA_CUDA(cudaMemcpy2D(t, 8, rec, pitch, 8, (size_t)T + 1, cudaMemcpyDeviceToHost));
            if (x) VA_CUDA(cudaMemcpy2D(x, (size_t)n * 8, rec + 1, pitch, (size_t)n * 8, (size_t)T + 1, cudaMemcpyDeviceToHost));
        } else {
            const size_t pitch = (size_t)e->arena_traj * 8;
            if (t) VA_CUDA(cudaMemcpy2D(t, 8, e->ck_t.as<double>() + b, pitch, 8, (size_t)T + 1, cudaMemcpyDeviceToHost));
            if (x) VA_CUDA(cudaMemcpy2D(x, 8, e->ck_x.as<double>() + b, pitch, 8, (size_t)(T + 1) * n, cudaMemcpyDeviceToHost));
        }
    } else {
        // slab of CTA b: one block per accepted step, header[0] = t_n, then the stage states; stage 0 is x_n. Block T
        // carries the final time only; x_T is x(tf).
        // slab 0 of the slot (cluster-pair kernel: CTA 2 * slot of the pair)
        if (e->pair_sparse) {
            // sparse store: every time is there, the states only at segment starts
            const double *tb = e->xstore.as<double>() + slot * e->pair_cl * e->xstore_stride;
            if (t) VA_CUDA(cudaMemcpy(t, tb, ((size_t)T + 1) * 8, cudaMemcpyDeviceToHost));
            if (x) return fail(VA_E_UNSUPPORTED, "VA_CKPT_SPARSE keeps the state of every L-th accepted step only: request the times (x = NULL), "
                                                "or use VA_CKPT_RECOMPUTE / VA_CKPT_STORE_STAGES when Driver::GetState is needed");
            return VA_OK;
        }
        const double *base = e->pair_seg ? e->xstore.as<double>() + slot * e->pair_cl * e->xstore_stride
                                         : e->slab.as<double>() + slot * (e->pairk ? e->pair_cl : e->pair) * e->slab_stride;
        const size_t pitch = e->pair_seg ? (size_t)(8 + 256) * 8 : (size_t)(e->family == FAM_GLV_WIDE || e->ring || e->pairk ? e->glv_blk
                                                                : va_glv_stream_block_doubles(n, e->desc.stepper, e->desc.ckpt_policy == VA_CKPT_RECOMPUTE)) * 8;
        if (t) VA_CUDA(cudaMemcpy2D(t, 8, base, pitch, 8, (size_t)T + 1, cudaMemcpyDeviceToHost));
        if (x) {
            if (T > 0) VA_CUDA(cudaMemcpy2D(x, (size_t)n * 8, base + e->glv_hdr, pitch, (size_t)n * 8, (size_t)T, cudaMemcpyDeviceToHost));
            std::memcpy(x + (size_t)T * n, e->se_xf_host.data() + (size_t)b * n, (size_t)n * 8);
        }
    }
    return VA_OK;
}

extern "C" {

const char *va_last_error(void) { return g_last_error.c_str(); }

int va_engine_create(const va_engine_desc *desc, va_engine **out)
{
    if (!desc || !out) return fail(VA_E_INVALID, "null argument");
    *out = nullptr;
    if (desc->n_devices < 0 || desc->n_devices > 64) return fail(VA_E_INVALID, "n_devices out of range");
    if (desc->n_devices > 0 && !desc->devices) return fail(VA_E_INVALID, "n_devices > 0 needs the devices array");
    if (desc->n_devices > 1) return va_multi_create(desc, out);
    va_engine_desc d = *desc;
    if (desc->n_devices == 1) d.device = desc->devices[0];
    d.n_devices = 0;
    d.devices = nullptr;
    return va_single_create(&d, out);
}

void va_engine_destroy(va_engine *e)
{
    if (!e) return;
    if (!e->members.empty()) va_multi_destroy(e);
    else va_single_destroy(e);
}

int va_engine_get_info(va_engine *e, va_engine_info *info)
{
    if (!e || !info) return fail(VA_E_INVALID, "null argument");
    if (e->members.empty()) return single_get_info(e, info);
    // multi-device engine: the first member describes the kernel; counters are summed, the time is the slowest GPU's
    single_get_info(e->members[0], info);
    info->n_devices = (int32_t)e->members.size();
    info->comm_rank = 0;
    for (size_t g = 1; g < e->members.size(); ++g) {
        va_engine_info ig;
        single_get_info(e->members[g], &ig);
        info->kernel_launches += ig.kernel_launches;
        info->workspace_bytes += ig.workspace_bytes;
        info->chunk_trajectories += ig.chunk_trajectories;
        info->collectives += ig.collectives;
        if (ig.last_kernel_ms > info->last_kernel_ms) info->last_kernel_ms = ig.last_kernel_ms;
    }
    return VA_OK;
}

int va_forward_adjoint_batch(va_engine *e, const va_batch_args *a)
{
    if (e && !e->members.empty()) return va_multi_call(e, 0, a);
    return va_single_forward_adjoint(e, a);
}
int va_forward_batch(va_engine *e, const va_batch_args *a)
{
    if (e && !e->members.empty()) return va_multi_call(e, 1, a);
    return va_single_forward(e, a);
}
int va_adjoint_batch(va_engine *e, const va_batch_args *a)
{
    if (e && !e->members.empty()) return va_multi_call(e, 2, a);
    return va_single_adjoint(e, a);
}
int va_forward_adjoint_batch_sharded(va_engine *e, int32_t n_shards, const va_batch_args *shards)
{
    if (!e || !shards) return fail(VA_E_INVALID, "null engine or shards");
    if (e->members.empty()) {
        if (n_shards != 1) return fail(VA_E_INVALID, "a single-device engine takes exactly one shard");
        return va_single_forward_adjoint(e, shards);
    }
    return va_multi_call_sharded(e, n_shards, shards);
}
int va_get_checkpoints(va_engine *e, int64_t b, int32_t capacity, double *t, double *x, int32_t *count)
{
    if (e && !e->members.empty()) return va_multi_get_checkpoints(e, b, capacity, t, x, count);
    return va_single_get_checkpoints(e, b, capacity, t, x, count);
}

int va_tape_compile_check(const char *tape_cuda_src, int32_t stepper, char *log, int32_t log_capacity)
{
    if (!tape_cuda_src) return fail(VA_E_INVALID, "null source");
    VaTableau tab;
    if (va_tableau_host(stepper, &tab) != 0) return fail(VA_E_UNSUPPORTED, "This stepper is not supported yet!");
    std::string l;
    const int rc = va_jit_compile(tape_cuda_src, tab.s, tab.fsal, tab.s_adj, nullptr, l);
    if (log && log_capacity > 0) std::snprintf(log, (size_t)log_capacity, "%s", l.c_str());
    if (rc != 0) return fail(VA_E_NVRTC, "tape -> CUDA compilation failed: " + l);
    return VA_OK;
}

int va_synth_batch_device(int32_t system, int32_t n_state, uint64_t seed, int64_t b0, int64_t B, double *params_dev, double *x0_dev,
                          void *stream)
{
    if (!params_dev || B < 0) return fail(VA_E_INVALID, "bad argument");
    VA_CUDA(va_synth_launch(system, n_state, seed, b0, B, params_dev, x0_dev, static_cast<cudaStream_t>(stream)));
    return VA_OK;
}

} // extern "C"
