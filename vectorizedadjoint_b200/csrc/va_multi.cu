// va_multi.cu -- several GPUs below the C line (include/va_engine.h, "Several GPUs").
//
// The reference integrates one parameter set on one CPU thread; its AAD workspace carries the note "needs to spawn several
// workspaces to allow multi-threading" (reference lib/include/AadData.hpp:32) and that is all the parallelism it has. Here the
// batch of parameter sets is the parallel axis and it shards over GPUs without any exchange during integration:
//   * multi-device engine (one process): a head engine over one member engine per GPU, each member driven by its own host
//     worker thread (its CUDA calls, its three streams, its host<->device pipeline), shards = contiguous ranges of the batch;
//   * one process per GPU: a single-device engine attached to a communicator created from a broadcast id.
// The only collective on the path: with VA_REDUCE_SUM each GPU reduces its shard to one [n_out][n_par] vector and ONE
// ncclAllReduce(ncclDouble, ncclSum) over NVLink combines them, enqueued on the compute stream inside the call.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already loaded when there is one, the system's otherwise),
// so single-GPU use never loads it and the library has no link-time dependency on it.
#include <dlfcn.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>

#include "va_engine_impl.h"

// ---- NCCL, bound at run time ------------------------------------------------------------------------------------------
namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.0)
struct NcclUniqueId { char internal[VA_COMM_ID_BYTES]; };
typedef void *NcclComm;
enum { kNcclSuccess = 0, kNcclDouble = 8, kNcclSum = 0 };

struct Nccl {
    void *handle = nullptr;
    int (*GetVersion)(int *) = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(NcclComm *, int, const int *) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};

Nccl &nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("VA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm) continue;
            n.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (n.handle) break;
        }
        if (!n.handle) {
            n.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?");
            return;
        }
        auto sym = [&](const char *s) {
            void *p = dlsym(n.handle, s);
            if (!p && n.error.empty()) n.error = std::string("libnccl lacks ") + s;
            return p;
        };
        n.GetVersion = reinterpret_cast<decltype(n.GetVersion)>(sym("ncclGetVersion"));
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
        n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return n;
}

int nccl_ready()
{
    Nccl &n = nccl();
    if (!n.error.empty()) return va_fail(VA_E_UNSUPPORTED, n.error);
    return VA_OK;
}

int nccl_fail(const char *what, int rc)
{
    Nccl &n = nccl();
    return va_fail(VA_E_CUDA, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
}

} // namespace

int va_nccl_version()
{
    Nccl &n = nccl();
    int v = 0;
    if (n.error.empty() && n.GetVersion) n.GetVersion(&v);
    return v;
}

int va_comm_allreduce_sum(va_engine *e, double *buf, int64_t count, cudaStream_t st)
{
    if (!e->comm) return VA_OK;
    if (int rc = nccl_ready()) return rc;
    const int r = nccl().AllReduce(buf, buf, (size_t)count, kNcclDouble, kNcclSum, static_cast<NcclComm>(e->comm), st);
    if (r != kNcclSuccess) return nccl_fail("ncclAllReduce", r);
    ++e->collectives;
    return VA_OK;
}

void va_comm_release(va_engine *e)
{
    if (e->comm && nccl().CommDestroy) nccl().CommDestroy(static_cast<NcclComm>(e->comm));
    e->comm = nullptr;
    e->comm_world = 0;
    e->comm_rank = 0;
}

// ---- worker threads: one per GPU of a multi-device engine ----------------------------------------------------------------
struct VaWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
    std::string err;

    explicit VaWorker(int device)
    {
        th = std::thread([this, device] {
            cudaSetDevice(device);
            std::unique_lock<std::mutex> lk(m);
            for (;;) {
                cv.wait(lk, [this] { return has_job || quit; });
                if (quit) return;
                std::function<int()> j = std::move(job);
                has_job = false;
                lk.unlock();
                const int r = j();
                std::string msg = r ? va_tls_error() : std::string();
                lk.lock();
                rc = r;
                err = std::move(msg);
                done = true;
                cv.notify_all();
            }
        });
    }
    void submit(std::function<int()> j)
    {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(j);
        has_job = true;
        done = false;
        cv.notify_all();
    }
    int wait()
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [this] { return done; });
        return rc;
    }
    ~VaWorker()
    {
        {
            std::lock_guard<std::mutex> lk(m);
            quit = true;
            cv.notify_all();
        }
        if (th.joinable()) th.join();
    }
};

namespace {

// run one job per member, all concurrently; the first failure's code and message are handed to the calling thread
int run_all(va_engine *head, const std::function<int(int)> &job_of)
{
    const int G = (int)head->members.size();
    for (int g = 0; g < G; ++g) head->workers[(size_t)g]->submit([&job_of, g] { return job_of(g); });
    int rc = VA_OK;
    std::string msg;
    for (int g = 0; g < G; ++g) {
        const int r = head->workers[(size_t)g]->wait();
        if (r != VA_OK && rc == VA_OK) {
            rc = r;
            msg = "device " + std::to_string(head->members[(size_t)g]->device) + ": " + head->workers[(size_t)g]->err;
        }
    }
    return rc == VA_OK ? VA_OK : va_fail(rc, msg);
}

} // namespace

extern "C" void va_shard_range(int64_t batch, int32_t g, int32_t world, int64_t *b0, int64_t *count)
{
    if (world < 1) world = 1;
    const int64_t base = batch / world, rem = batch % world;
    if (count) *count = base + (g < rem ? 1 : 0);
    if (b0) *b0 = g * base + std::min<int64_t>(g, rem);
}

int va_multi_create(const va_engine_desc *desc, va_engine **out)
{
    const int G = desc->n_devices;
    for (int g = 0; g < G; ++g)
        for (int h = 0; h < g; ++h)
            if (desc->devices[g] == desc->devices[h]) return va_fail(VA_E_INVALID, "devices must be distinct");
    if (int rc = nccl_ready()) return rc;
    va_engine *head = new (std::nothrow) va_engine();
    if (!head) return va_fail(VA_E_NOMEM, "out of host memory");
    head->desc = *desc;
    head->desc.devices = nullptr;
    head->desc.tape_cuda_src = nullptr;
    head->device = desc->devices[0];
    for (int g = 0; g < G; ++g) {
        va_engine_desc d = *desc;
        d.device = desc->devices[g];
        d.n_devices = 0;
        d.devices = nullptr;
        va_engine *m = nullptr;
        const int rc = va_single_create(&d, &m);
        if (rc != VA_OK) {
            const std::string msg = "device " + std::to_string(d.device) + ": " + va_tls_error();
            va_multi_destroy(head);
            return va_fail(rc, msg);
        }
        m->comm_owned_by_head = true;
        m->mu_host_writer = g == 0;
        head->members.push_back(m);
    }
    // one communicator per GPU, all created by this thread (ncclCommInitAll), used afterwards one per worker thread
    std::vector<NcclComm> comms((size_t)G, nullptr);
    const int r = nccl().CommInitAll(comms.data(), G, desc->devices);
    if (r != kNcclSuccess) {
        va_multi_destroy(head);
        return nccl_fail("ncclCommInitAll", r);
    }
    for (int g = 0; g < G; ++g) {
        head->members[(size_t)g]->comm = comms[(size_t)g];
        head->members[(size_t)g]->comm_world = G;
        head->members[(size_t)g]->comm_rank = g;
        head->workers.push_back(new VaWorker(desc->devices[g]));
    }
    head->comm_world = G;
    *out = head;
    return VA_OK;
}

void va_multi_destroy(va_engine *head)
{
    if (!head) return;
    for (VaWorker *w : head->workers) delete w;
    head->workers.clear();
    for (va_engine *m : head->members) {
        cudaSetDevice(m->device);
        cudaDeviceSynchronize();
    }
    for (va_engine *m : head->members) {
        cudaSetDevice(m->device);
        va_comm_release(m);
    }
    for (va_engine *m : head->members) va_single_destroy(m);
    head->members.clear();
    delete head;
}

// contiguous split of a HOST batch (which: 0 fused, 1 forward, 2 adjoint of the split API)
int va_multi_call(va_engine *head, int which, const va_batch_args *a)
{
    if (!a) return va_fail(VA_E_INVALID, "null args");
    if (a->mem != VA_MEM_HOST)
        return va_fail(VA_E_INVALID, "a multi-device engine splits HOST batches; device-resident shards go through va_forward_adjoint_batch_sharded");
    if (a->batch < 0) return va_fail(VA_E_INVALID, "negative batch");
    const int G = (int)head->members.size();
    const int n = head->desc.n_state, npar = head->desc.n_par, nout = head->desc.n_out;
    const bool sum = a->reduce == VA_REDUCE_SUM;
    std::vector<va_batch_args> sh((size_t)G, *a);
    for (int g = 0; g < G; ++g) {
        int64_t b0 = 0, cnt = 0;
        va_shard_range(a->batch, g, G, &b0, &cnt);
        va_batch_args &s = sh[(size_t)g];
        s.batch = cnt;
        if (a->x0) s.x0 = a->x0 + b0 * n;
        if (a->params) s.params = a->params + b0 * npar;
        if (a->x_final) s.x_final = a->x_final + b0 * n;
        if (a->lambda) s.lambda = a->lambda + b0 * nout * n;
        if (a->mu && !sum) s.mu = a->mu + b0 * nout * npar;
        if (a->n_accept) s.n_accept = a->n_accept + b0;
        if (a->n_reject) s.n_reject = a->n_reject + b0;
        if (a->status) s.status = a->status + b0;
        s.stream = nullptr;
    }
    if (which == 2) { // the split must be the one of the forward call
        int64_t total = 0;
        for (va_engine *m : head->members) total += m->se_B;
        if (total <= 0) return va_fail(VA_E_STATE, "va_adjoint_batch needs a preceding va_forward_batch (runge_kutta) on this engine");
        if (total != a->batch) return va_fail(VA_E_INVALID, "batch differs from the preceding va_forward_batch");
    }
    // a summed call is collective over ALL members: a shard may be empty (batch < G) but must still join the all-reduce
    return run_all(head, [&](int g) {
        va_engine *m = head->members[(size_t)g];
        const va_batch_args *s = &sh[(size_t)g];
        if (s->batch == 0 && which != 1) {
            if (sum) {
                if (int rc = m->st_musum.ensure((size_t)nout * npar * 8)) return rc;
                if (cudaMemsetAsync(m->st_musum.p, 0, (size_t)nout * npar * 8, m->s_comp) != cudaSuccess) return va_fail(VA_E_CUDA, "cudaMemsetAsync failed");
                if (int rc = va_comm_allreduce_sum(m, m->st_musum.as<double>(), (int64_t)nout * npar, m->s_comp)) return rc;
                if (m->mu_host_writer && a->mu &&
                    cudaMemcpyAsync(a->mu, m->st_musum.p, (size_t)nout * npar * 8, cudaMemcpyDeviceToHost, m->s_comp) != cudaSuccess)
                    return va_fail(VA_E_CUDA, "cudaMemcpyAsync failed");
                if (cudaStreamSynchronize(m->s_comp) != cudaSuccess) return va_fail(VA_E_CUDA, "cudaStreamSynchronize failed");
            }
            m->se_B = 0;
            return (int)VA_OK;
        }
        if (s->batch == 0) { m->se_B = 0; return (int)VA_OK; }
        return which == 0 ? va_single_forward_adjoint(m, s) : which == 1 ? va_single_forward(m, s) : va_single_adjoint(m, s);
    });
}

int va_multi_call_sharded(va_engine *head, int32_t n_shards, const va_batch_args *shards)
{
    const int G = (int)head->members.size();
    if (n_shards != G) return va_fail(VA_E_INVALID, "n_shards must equal the engine's n_devices");
    for (int g = 1; g < G; ++g)
        if (shards[g].ti != shards[0].ti || shards[g].tf != shards[0].tf || shards[g].dt0 != shards[0].dt0 ||
            shards[g].objective != shards[0].objective || shards[g].reduce != shards[0].reduce || shards[g].mem != shards[0].mem)
            return va_fail(VA_E_INVALID, "ti, tf, dt0, objective, reduce and mem must agree across shards");
    const bool sum = shards[0].reduce == VA_REDUCE_SUM;
    if (sum)
        for (int g = 0; g < G; ++g)
            if (shards[g].batch <= 0) return va_fail(VA_E_INVALID, "with VA_REDUCE_SUM every shard must hold at least one parameter set");
    // caller-sharded host buffers: every member writes its own mu (each receives the all-reduced sum)
    std::vector<bool> writer((size_t)G);
    for (int g = 0; g < G; ++g) { writer[(size_t)g] = head->members[(size_t)g]->mu_host_writer; head->members[(size_t)g]->mu_host_writer = true; }
    const int rc = run_all(head, [&](int g) {
        if (shards[g].batch <= 0) return (int)VA_OK;
        return va_single_forward_adjoint(head->members[(size_t)g], &shards[g]);
    });
    for (int g = 0; g < G; ++g) head->members[(size_t)g]->mu_host_writer = writer[(size_t)g];
    return rc;
}

int va_multi_get_checkpoints(va_engine *head, int64_t b, int32_t capacity, double *t, double *x, int32_t *count)
{
    if (b < 0) return va_fail(VA_E_INVALID, "trajectory index out of range");
    int64_t off = 0;
    for (size_t g = 0; g < head->members.size(); ++g) {
        va_engine *m = head->members[g];
        if (b < off + m->se_B) {
            const int64_t bl = b - off;
            return run_all(head, [&, g, bl](int gg) {
                return (size_t)gg == g ? va_single_get_checkpoints(m, bl, capacity, t, x, count) : (int)VA_OK;
            });
        }
        off += m->se_B;
    }
    return va_fail(off == 0 ? VA_E_STATE : VA_E_INVALID, off == 0 ? "no forward sweep recorded on this engine" : "trajectory index out of range");
}

// ---- one process per GPU ---------------------------------------------------------------------------------------------------
extern "C" int va_comm_unique_id(void *id, int32_t id_bytes)
{
    if (!id || id_bytes < VA_COMM_ID_BYTES) return va_fail(VA_E_INVALID, "id buffer must hold VA_COMM_ID_BYTES");
    if (int rc = nccl_ready()) return rc;
    NcclUniqueId u;
    const int r = nccl().GetUniqueId(&u);
    if (r != kNcclSuccess) return nccl_fail("ncclGetUniqueId", r);
    std::memcpy(id, &u, VA_COMM_ID_BYTES);
    return VA_OK;
}

extern "C" int va_engine_comm_init(va_engine *e, const void *id, int32_t id_bytes, int32_t rank, int32_t world)
{
    if (!e || !id || id_bytes < VA_COMM_ID_BYTES) return va_fail(VA_E_INVALID, "null engine or id");
    if (!e->members.empty()) return va_fail(VA_E_INVALID, "a multi-device engine already owns its communicator");
    if (world < 1 || rank < 0 || rank >= world) return va_fail(VA_E_INVALID, "bad rank / world");
    if (e->comm) return va_fail(VA_E_STATE, "the engine is already attached to a communicator");
    if (int rc = nccl_ready()) return rc;
    VA_CUDA(cudaSetDevice(e->device));
    NcclUniqueId u;
    std::memcpy(&u, id, VA_COMM_ID_BYTES);
    NcclComm c = nullptr;
    const int r = nccl().CommInitRank(&c, world, u, rank);
    if (r != kNcclSuccess) return nccl_fail("ncclCommInitRank", r);
    e->comm = c;
    e->comm_world = world;
    e->comm_rank = rank;
    return VA_OK;
}

// ---- page-locked host buffers ------------------------------------------------------------------------------------------------
namespace {

std::mutex g_host_mutex;
struct HostAlloc { size_t bytes; int kind; }; // kind 0: cudaHostAlloc, 1: mmap + mbind + cudaHostRegister
std::map<void *, HostAlloc> g_host_allocs;

int numa_nodes_online()
{
    int n = 0;
    for (int k = 0; k < 64; ++k) {
        const std::string p = "/sys/devices/system/node/node" + std::to_string(k);
        if (access(p.c_str(), F_OK) == 0) ++n;
    }
    return n;
}

int numa_node_of_device(int device)
{
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
    std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
    int node = -1;
    if (f) f >> node;
    return node;
}

} // namespace

extern "C" int va_host_alloc(void **ptr, int64_t bytes, int32_t flags, int32_t device)
{
    if (!ptr || bytes <= 0) return va_fail(VA_E_INVALID, "bad argument");
    *ptr = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return va_fail(VA_E_CUDA, "no CUDA device"); }
    if (device < 0 || device >= ndev) return va_fail(VA_E_INVALID, "bad device ordinal");
    VA_CUDA(cudaSetDevice(device));
    const int node = (flags & VA_HOST_NUMA_LOCAL) && numa_nodes_online() > 1 ? numa_node_of_device(device) : -1;
    if (node >= 0 && !(flags & VA_HOST_WRITE_COMBINED)) {
        // bind the pages to the GPU's node, then page-lock them (write-combined memory can only come from cudaHostAlloc)
        const long pg = sysconf(_SC_PAGESIZE);
        const size_t len = ((size_t)bytes + (size_t)pg - 1) / (size_t)pg * (size_t)pg;
        void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            unsigned long mask[16] = {0};
            mask[node / 64] |= 1UL << (node % 64);
            // mbind(addr, len, MPOL_BIND = 2, nodemask, maxnode, 0)
            if (syscall(SYS_mbind, p, len, 2, mask, (unsigned long)(sizeof(mask) * 8), 0U) == 0 &&
                cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(g_host_mutex);
                g_host_allocs[p] = HostAlloc{len, 1};
                *ptr = p;
                return VA_OK;
            }
            cudaGetLastError();
            munmap(p, len);
        }
        // fall through to the plain allocation
    }
    void *p = nullptr;
    unsigned f = cudaHostAllocPortable | ((flags & VA_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : 0);
    cudaError_t ce = cudaHostAlloc(&p, (size_t)bytes, f);
    if (ce != cudaSuccess) { cudaGetLastError(); return va_fail(VA_E_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(ce)); }
    std::lock_guard<std::mutex> lk(g_host_mutex);
    g_host_allocs[p] = HostAlloc{(size_t)bytes, 0};
    *ptr = p;
    return VA_OK;
}

extern "C" int va_host_free(void *ptr)
{
    if (!ptr) return VA_OK;
    HostAlloc h;
    {
        std::lock_guard<std::mutex> lk(g_host_mutex);
        auto it = g_host_allocs.find(ptr);
        if (it == g_host_allocs.end()) return va_fail(VA_E_INVALID, "not a va_host_alloc pointer");
        h = it->second;
        g_host_allocs.erase(it);
    }
    if (h.kind == 1) {
        cudaHostUnregister(ptr);
        munmap(ptr, h.bytes);
    } else {
        cudaFreeHost(ptr);
    }
    return VA_OK;
}

// ---- host -> device copy ceiling ---------------------------------------------------------------------------------------------
extern "C" int va_measure_h2d_copy(const int32_t *devices, int32_t n, int64_t bytes, int32_t reps, int32_t flags, const void *host,
                                   double *gbytes_per_s, double *aggregate)
{
    if (!devices || n < 1 || n > 64 || bytes <= 0) return va_fail(VA_E_INVALID, "bad argument");
    if (reps < 1) reps = 3;
    struct Per { void *dev = nullptr; void *hst = nullptr; bool own = false; cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr; double best_ms = 1e30; int rc = 0; };
    std::vector<Per> per((size_t)n);
    int rc = VA_OK;
    for (int g = 0; g < n && rc == VA_OK; ++g) {
        Per &p = per[(size_t)g];
        if (cudaSetDevice(devices[g]) != cudaSuccess) { rc = va_fail(VA_E_CUDA, "cudaSetDevice failed"); break; }
        if (cudaMalloc(&p.dev, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); rc = va_fail(VA_E_NOMEM, "cudaMalloc failed"); break; }
        if (host) p.hst = const_cast<char *>(static_cast<const char *>(host)) + (size_t)g * (size_t)bytes;
        else {
            rc = va_host_alloc(&p.hst, bytes, flags, devices[g]);
            if (rc != VA_OK) break;
            p.own = true;
            std::memset(p.hst, 1, (size_t)bytes); // touch the pages (first touch happens on this thread's node unless bound)
        }
        cudaStreamCreateWithFlags(&p.st, cudaStreamNonBlocking);
        cudaEventCreate(&p.e0);
        cudaEventCreate(&p.e1);
    }
    double best_wall = 1e30;
    if (rc == VA_OK) {
        for (int rep = 0; rep <= reps; ++rep) { // rep 0 = warm-up
            for (int g = 0; g < n; ++g) { cudaSetDevice(devices[g]); cudaStreamSynchronize(per[(size_t)g].st); }
            const auto t0 = std::chrono::steady_clock::now();
            for (int g = 0; g < n; ++g) {
                Per &p = per[(size_t)g];
                cudaSetDevice(devices[g]);
                cudaEventRecord(p.e0, p.st);
                cudaMemcpyAsync(p.dev, p.hst, (size_t)bytes, cudaMemcpyHostToDevice, p.st);
                cudaEventRecord(p.e1, p.st);
            }
            for (int g = 0; g < n; ++g) {
                cudaSetDevice(devices[g]);
                if (cudaStreamSynchronize(per[(size_t)g].st) != cudaSuccess) rc = va_fail(VA_E_CUDA, "copy failed");
            }
            const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (rep == 0 || rc != VA_OK) continue;
            best_wall = std::min(best_wall, wall);
            for (int g = 0; g < n; ++g) {
                float ms = 0;
                cudaEventElapsedTime(&ms, per[(size_t)g].e0, per[(size_t)g].e1);
                per[(size_t)g].best_ms = std::min(per[(size_t)g].best_ms, (double)ms);
            }
        }
    }
    for (int g = 0; g < n; ++g) {
        Per &p = per[(size_t)g];
        cudaSetDevice(devices[g]);
        if (gbytes_per_s && rc == VA_OK) gbytes_per_s[g] = (double)bytes / (p.best_ms * 1e-3) / 1e9;
        if (p.e0) cudaEventDestroy(p.e0);
        if (p.e1) cudaEventDestroy(p.e1);
        if (p.st) cudaStreamDestroy(p.st);
        if (p.dev) cudaFree(p.dev);
        if (p.own && p.hst) va_host_free(p.hst);
    }
    if (aggregate && rc == VA_OK) *aggregate = (double)bytes * n / (best_wall * 1e-3) / 1e9;
    return rc;
}
