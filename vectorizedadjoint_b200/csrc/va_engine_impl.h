// va_engine_impl.h -- the engine object behind the opaque `va_engine` of include/va_engine.h, shared by va_engine.cu (one
// GPU: workspaces, the host pipeline, the split-API session) and va_multi.cu (several GPUs: worker threads, NCCL).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "va_common.cuh"
#include "va_jit.h"

int va_fail(int code, const std::string &msg);   // sets the calling thread's va_last_error() text, returns code
const std::string &va_tls_error();

#define VA_CUDA(call)                                                                                       \
    do {                                                                                                    \
        cudaError_t err__ = (call);                                                                         \
        if (err__ != cudaSuccess)                                                                           \
            return va_fail(err__ == cudaErrorMemoryAllocation ? VA_E_NOMEM : VA_E_CUDA,                     \
                           std::string(#call) + ": " + cudaGetErrorString(err__));                          \
    } while (0)

enum Family { FAM_SCALAR = 0, FAM_GLV_WIDE = 1, FAM_GLV_STREAM = 2, FAM_TAPE = 3 }; // TAPE: scalar kernels compiled at run time for a recorded system

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); } // `delete engine` frees every buffer (the engine's device is current in va_engine_destroy)
    int ensure(size_t need)
    {
        if (need <= bytes) return VA_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) { cudaGetLastError(); return va_fail(VA_E_NOMEM, "cudaMalloc(" + std::to_string(need) + " B) failed"); }
        bytes = need;
        return VA_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct VaWorker; // va_multi.cu: one host thread bound to one GPU of a multi-device engine

struct va_engine {
    va_engine_desc desc;
    VaTableau tab;
    int family = FAM_SCALAR;
    VaJitModule *jit = nullptr;
    int device = 0, sm_count = 0;
    int cap = 0;
    cudaStream_t s_comp = nullptr, s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    // wide family
    int grid = 0, ctas_per_sm = 0, threads = 0, tpc = 1, pair = 1; // tpc: slots per CTA; pair: slabs per slot
    bool t8 = false;  // FAM_GLV_WIDE served by va_glv_t8.cu (33..64 species)
    bool t8s = false; // ... by its third generation va_glv_t8s.cu (sweep / accumulate warps); implies t8
    bool quad = false; // FAM_GLV_WIDE served by va_glv_quad.cu (up to 16 species, store-stages policy)
    bool oct = false;  // FAM_GLV_WIDE served by va_glv_oct.cu (up to 16 species, recompute policy: the default there)
    int glv_blk = 0;  // doubles per step block of the register-kernel slab
    int glv_hdr = 8;  // doubles in front of the stage states of a step block (header[0] = t_n)
    bool ring = false; // FAM_GLV_STREAM served by va_glv_ring.cu (256 species, store-stages policy)
    int ring_flags = 0;
    bool pairk = false; // FAM_GLV_STREAM served by va_glv_pair.cu (256 species, matrix on chip in a cluster of pair_cl CTAs)
    int pair_cl = 2;
    bool pair_seg = false; // ... under the recompute policy: per-CTA state store + segment re-integration
    int pair_seg_len = 16;
    bool pair_sparse = false; // ... with the sparse state store (VA_CKPT_SPARSE): t_n of every step, x_n of every pair_seg_len-th
    int64_t xstore_stride = 0;
    DevBuf xstore;
    int64_t slab_stride = 0;
    DevBuf slab, partial;
    // scalar family
    DevBuf ck_t, ck_x, work_counter; // work_counter: 2 x u64, dynamic trajectory / work-item fetch of the persistent grids
    int64_t arena_traj = 0;
    // per-trajectory bookkeeping when the caller passes NULL
    DevBuf own_accept, own_reject, own_status, mu_tmp;
    // host-mode staging, two slots
    DevBuf st_x0[2], st_par[2], st_xf[2], st_lam[2], st_mu[2], st_acc[2], st_rej[2], st_sta[2], st_musum;
    // split API session (va_forward_batch -> va_adjoint_batch / va_get_checkpoints)
    DevBuf se_x0, se_par, se_xf, se_lam, se_mu, se_acc, se_rej, se_sta, se_scratch;
    int64_t se_B = 0;
    double se_ti = 0, se_tf = 0, se_dt0 = 0;
    std::vector<int32_t> se_accept_host;
    std::vector<double> se_xf_host;
    bool se_slab_valid = false; // GLV: the slabs still hold the first wave of the session's forward sweep
    bool se_blocks_intact = false; // ... and no reverse sweep has written into those step blocks yet (they can be swept back once more)
    int64_t se_ck_b = -1;       // GLV: trajectory whose checkpoints were re-integrated into slot 0 for va_get_checkpoints
    int64_t launches = 0;
    double last_ms = 0.0;
    int64_t workspace_bytes = 0, chunk_traj = 0;
    // several GPUs (va_multi.cu). A multi-device engine is a head (no kernels of its own) over one member engine per GPU.
    std::vector<va_engine *> members; // head only
    std::vector<VaWorker *> workers;  // head only, one per member
    void *comm = nullptr;             // ncclComm_t of a member / of a single-device engine after va_engine_comm_init
    int comm_world = 0, comm_rank = 0;
    bool comm_owned_by_head = false;
    int64_t collectives = 0;
    bool holds_persist_l2 = false;    // this engine holds a reference on its device's persisting-L2 carve-out
    bool mu_host_writer = true;       // host pipeline, VA_REDUCE_SUM: this engine copies the (all-reduced) sum to the caller's mu
};

// single-device entry points (va_engine.cu), used directly and by the workers of a multi-device engine
int va_single_create(const va_engine_desc *desc, va_engine **out);
void va_single_destroy(va_engine *e);
int va_single_forward_adjoint(va_engine *e, const va_batch_args *a);
int va_single_forward(va_engine *e, const va_batch_args *a);
int va_single_adjoint(va_engine *e, const va_batch_args *a);
int va_single_get_checkpoints(va_engine *e, int64_t b, int32_t capacity, double *t, double *x, int32_t *count);

// several GPUs (va_multi.cu)
int va_multi_create(const va_engine_desc *desc, va_engine **out);
void va_multi_destroy(va_engine *head);
int va_multi_call(va_engine *head, int which, const va_batch_args *a); // which: 0 forward_adjoint, 1 forward, 2 adjoint (contiguous split of a host batch)
int va_multi_call_sharded(va_engine *head, int32_t n, const va_batch_args *shards);
int va_multi_get_checkpoints(va_engine *head, int64_t b, int32_t capacity, double *t, double *x, int32_t *count);
// sum over the ranks of e's communicator, in place, on stream st (ncclAllReduce, ncclDouble, ncclSum)
int va_comm_allreduce_sum(va_engine *e, double *buf, int64_t count, cudaStream_t st);
void va_comm_release(va_engine *e);
int va_nccl_version();


