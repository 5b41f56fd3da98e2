/*
 * va_engine.h -- C-ABI of the B200-native discrete-adjoint engine (libva_engine.so).
 *
 * This is the drop-in boundary for the ONE hot path of RuiMartins1996/VectorizedAdjoint: the forward explicit-RK
 * sweep with checkpointing and the backward discrete-adjoint sweep, batched over parameter sets. Plain C types only:
 * pointers, sizes, status codes. No C++ / torch types cross this line. Every entry point returns 0 on success and a
 * negative VA_E_* code on failure (never throws, never aborts); va_last_error() gives the message for the calling thread.
 *
 * What each entry point replaces in the reference (paths relative to the reference repository):
 *
 *   va_engine_create         Driver::Driver(Nin,Nout,Npar)                      lib/include/Driver.hpp:39-42
 *                            + constructDriverButcherTableau(driver, stepper)   lib/include/Driver.hpp:87-93
 *                              (ButcherTable: lib/include/ButcherTable.hpp:11-264)
 *                            + recordDriverRHSFunction(driver, system)          lib/include/Driver.hpp:95-100
 *                              (AadData::Record: lib/include/AadData.hpp:124-171 -- the AADC JIT is replaced by CUDA
 *                               device functors: built-in HARMONIC / VANDERPOL / GLV, or a recorded TAPE)
 *   va_forward_batch         vectorizedadjoint::runge_kutta(...)                lib/include/runge_kutta.hpp:47-59
 *                              (loops: lib/include/detail/runge_kutta.hpp:38-72 fixed, :76-118 adaptive;
 *                               checkpoint store: lib/include/StateStorage.hpp:4-42)
 *   va_adjoint_batch         setCostGradients + adjointSolve                    lib/include/Driver.hpp:103-114,
 *                                                                               lib/include/backpropagation.hpp:18-43
 *                              (recursion: lib/include/detail/backpropagation.hpp:24-158, 231-348;
 *                               VJP: lib/include/AadData.hpp:291-330)
 *   va_forward_adjoint_batch the two calls above fused (the benchmarked call; checkpoints never leave the GPU)
 *   va_get_checkpoints       Driver::GetT / GetTime / GetDt / GetState          lib/include/Driver.hpp:53-66
 *   va_engine_destroy        delete_driver_handle(void*)                        lib/include/Driver.hpp:81-85
 *   va_forward_adjoint_batch_sharded, va_comm_unique_id, va_engine_comm_init
 *                            nothing: the reference is single-threaded and single-device; its own note that the AAD
 *                            workspace "needs to spawn several workspaces to allow multi-threading" is the closest it
 *                            gets                                               lib/include/AadData.hpp:32
 *
 * Several GPUs. Parameter sets are independent, so a batch shards contiguously (set b of B goes to device
 * g = the shard whose range [b0_g, b0_g + count_g) holds b; counts differ by at most one) and nothing is exchanged during
 * integration. Two ways to use G GPUs, both below this C line:
 *   (1) one process, one engine: va_engine_desc.n_devices = G, .devices = ordinals. Every batch call shards over the
 *       listed GPUs (one host worker thread and one set of streams per GPU). With VA_REDUCE_SUM the per-GPU sums are
 *       combined by ONE ncclAllReduce(ncclDouble, ncclSum) inside the call (communicator from ncclCommInitAll).
 *   (2) one process per GPU (torchrun / MPI): every rank creates a single-device engine and attaches it to a communicator
 *       with va_comm_unique_id (rank 0, then broadcast by the launcher's own means) + va_engine_comm_init (all ranks).
 *       From then on a VA_REDUCE_SUM call on that engine is collective: it ends in the same ncclAllReduce over the ranks.
 *
 * The reference has no batch axis (one Driver = one trajectory; its SIMD lanes carry adjoint seeds). Here the batch of
 * parameter sets is the parallel axis; B = 1 reproduces the reference call for call.
 *
 * Layouts (host or device, chosen per call by va_batch_args.mem): array-of-parameter-sets, exactly what a caller of
 * the reference holds in its std::vector<double>s:
 *     x0[b*n_state + i], params[b*n_par + k], lambda[(b*n_out + o)*n_state + i], mu[(b*n_out + o)*n_par + k].
 * With reduce == VA_REDUCE_SUM, mu is [n_out][n_par] = sum over b. GLV parameter order is the reference's:
 * params = [r_0..r_{N-1}, A_00, A_01, ... A_{N-1,N-1}] (examples/GeneralizedLotkaVolterra/main.cpp:114).
 */
#ifndef VA_ENGINE_H
#define VA_ENGINE_H

#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define VA_API_VERSION 2

typedef struct va_engine va_engine;

enum va_system { VA_SYS_HARMONIC = 0, VA_SYS_VANDERPOL = 1, VA_SYS_GLV = 2, VA_SYS_TAPE = 3 };
enum va_stepper { VA_RK_EULER = 0, VA_RK_RK4 = 1, VA_RK_CK54 = 2, VA_RK_DOPRI5 = 3, VA_RK_RKF78 = 4 };
enum va_objective {
    VA_OBJ_SEED = 0,       /* lambda holds dJ/dx(tf) on entry (setCostGradients semantics)             */
    VA_OBJ_SUM = 1,        /* J = sum_i x_i(tf): seed = 1, computed on the device                         */
    VA_OBJ_HALF_NORM2 = 2  /* J = |x(tf)|^2 / 2: seed = x(tf) (HarmonicOscillator example), on the device */
};
enum va_reduce { VA_REDUCE_NONE = 0, VA_REDUCE_SUM = 1 };
enum va_mem { VA_MEM_HOST = 0, VA_MEM_DEVICE = 1 };
enum va_ckpt_policy {
    VA_CKPT_AUTO = 0,
    VA_CKPT_RECOMPUTE = 1,   /* store accepted (t_n, x_n); recompute the stages in the reverse sweep (reference policy) */
    VA_CKPT_STORE_STAGES = 2,/* additionally store the stage states/slopes of accepted steps; no recompute            */
    VA_CKPT_SPARSE = 3       /* store t_n of every accepted step but x_n only of every L-th; the reverse sweep re-integrates
                                each segment of L steps from its first state (checkpoint memory 8 + 8 N / L bytes per step).
                                GLV 65..256 species (cluster kernel); elsewhere it means VA_CKPT_RECOMPUTE               */
};
/* per-trajectory status word (bit mask) */
enum va_traj_status { VA_TRAJ_OK = 0, VA_TRAJ_CKPT_OVERFLOW = 1, VA_TRAJ_NO_PROGRESS = 2, VA_TRAJ_NONFINITE = 4 };
/* call status */
enum va_error {
    VA_OK = 0, VA_E_INVALID = -1, VA_E_UNSUPPORTED = -2, VA_E_CUDA = -3, VA_E_NOMEM = -4, VA_E_STATE = -5, VA_E_NVRTC = -6
};

typedef struct va_engine_desc {
    int32_t system;      /* va_system                                                                          */
    int32_t n_state;     /* Nin                                                                                */
    int32_t n_par;       /* Npar (GLV: n_state^2 + n_state)                                                    */
    int32_t n_out;       /* Nout: cost functions (adjoint seeds) per trajectory                                */
    int32_t stepper;     /* va_stepper                                                                         */
    int32_t adaptive;    /* 0: fixed step (stepper_tag loop), 1: controlled (make_controlled<stepper>(abs,rel)) */
    double eps_abs, eps_rel;
    int32_t device;      /* CUDA ordinal (ignored when n_devices > 0)                                          */
    int32_t max_steps;   /* checkpoint capacity per trajectory (accepted steps); 0 = default                   */
    int32_t ckpt_policy; /* va_ckpt_policy                                                                     */
    int32_t n_devices;   /* 0: one GPU, `device`; G >= 1: every batch is sharded over devices[0..G)            */
    double workspace_fraction; /* share of free HBM the checkpoint arena may take (0 = default 0.5)            */
    const char *tape_cuda_src; /* VA_SYS_TAPE: CUDA source of the rhs/vjp device functors (see va_tape.h)      */
    const int32_t *devices;    /* [n_devices] distinct CUDA ordinals (read during va_engine_create only)       */
} va_engine_desc;

typedef struct va_batch_args {
    int64_t batch;        /* B parameter sets                                                                   */
    const double *x0;     /* [B][n_state]                                                                       */
    const double *params; /* [B][n_par]                                                                         */
    double ti, tf, dt0;
    int32_t objective;    /* va_objective                                                                       */
    int32_t reduce;       /* va_reduce                                                                          */
    int32_t mem;          /* va_mem: where x0/params/x_final/lambda/mu/n_accept/... live. VA_MEM_DEVICE arrays    */
                          /* of va_forward_adjoint_batch must be 16-byte aligned (VA_E_INVALID otherwise)        */
    int32_t reserved0;
    double *x_final;      /* [B][n_state] out: x(tf)                                                            */
    double *lambda;       /* [B][n_out][n_state] in (VA_OBJ_SEED): dJ/dx(tf); out: dJ/dx(ti)                    */
    double *mu;           /* out, OVERWRITTEN: [B][n_out][n_par] or, with VA_REDUCE_SUM, [n_out][n_par]         */
    int32_t *n_accept;    /* [B] accepted steps (optional, may be NULL)                                         */
    int32_t *n_reject;    /* [B] rejected attempts (optional)                                                   */
    int32_t *status;      /* [B] va_traj_status bits (optional)                                                 */
    void *stream;         /* cudaStream_t to run on when mem == VA_MEM_DEVICE (NULL = the engine's own stream)  */
} va_batch_args;

typedef struct va_engine_info {
    int32_t api_version, device, sm_count, kernel_family; /* family: 0 scalar, 1 glv-wide, 2 glv-generic, 3 tape */
    int32_t ckpt_policy, max_steps, ctas_per_sm, threads_per_cta;
    int64_t workspace_bytes, chunk_trajectories;
    int64_t kernel_launches; /* launches of this library's kernels since creation                                */
    double last_kernel_ms;   /* device time of the last fused/forward/adjoint call (CUDA events on its stream)   */
    char device_name[64];
    char kernel_name[32];    /* the kernel that serves this engine: k_scalar, k_glv_wide, k_glv_t8, k_glv_stream, k_glv_ring,
                                k_glv_pair, or jit (run-time compiled thread-per-trajectory kernels of a recorded system)  */
    int32_t n_devices;       /* GPUs this engine shards over (1 for a single-device engine)                      */
    int32_t comm_world;      /* ranks of the attached communicator (n_devices, or va_engine_comm_init's world; 0 = none) */
    int32_t comm_rank;       /* this engine's rank in it (multi-device engine: 0)                                */
    int32_t nccl_version;    /* ncclGetVersion() of the library in use, 0 when no communicator is attached       */
    int64_t collectives;     /* ncclAllReduce calls issued since creation (all devices of the engine)            */
} va_engine_info;

#ifndef __CUDACC_RTC__ /* the enums and structs above are also seen by run-time compiled device code */
int va_engine_create(const va_engine_desc *desc, va_engine **out);
void va_engine_destroy(va_engine *e);
int va_engine_get_info(va_engine *e, va_engine_info *info);
const char *va_last_error(void);

int va_forward_batch(va_engine *e, const va_batch_args *args);
int va_adjoint_batch(va_engine *e, const va_batch_args *args);
int va_forward_adjoint_batch(va_engine *e, const va_batch_args *args);

/* Multi-device engine (n_devices = G), caller-sharded form: shards[g] describes the part of the batch that lives on
 * devices[g] (its own batch count and buffers; VA_MEM_DEVICE buffers must be on that GPU, shards[g].stream a stream of that
 * GPU or NULL). All G shards run concurrently; ti/tf/dt0/objective/reduce/mem must agree. With VA_REDUCE_SUM every
 * shards[g].mu receives the sum over ALL shards (one ncclAllReduce). va_forward_adjoint_batch on a multi-device engine
 * is this call with the contiguous split of a VA_MEM_HOST batch (mu then is one host buffer). */
int va_forward_adjoint_batch_sharded(va_engine *e, int32_t n_shards, const va_batch_args *shards);
/* shard g of `world`: first parameter set and count of the contiguous split used by the multi-device calls */
void va_shard_range(int64_t batch, int32_t g, int32_t world, int64_t *b0, int64_t *count);

/* One process per GPU. va_comm_unique_id: rank 0 obtains an opaque id (VA_COMM_ID_BYTES) and hands it to the other ranks
 * by the launcher's means; va_engine_comm_init: collective over all `world` ranks, attaches a single-device engine to the
 * communicator. Afterwards VA_REDUCE_SUM calls on the engine are collective and return the sum over all ranks. */
#define VA_COMM_ID_BYTES 128
int va_comm_unique_id(void *id, int32_t id_bytes);
int va_engine_comm_init(va_engine *e, const void *id, int32_t id_bytes, int32_t rank, int32_t world);

/* Checkpoints of trajectory b of the last va_forward_batch call: count = T+1 entries (t_n, x_n[n_state]).
 * Pass t = x = NULL to query the count only. */
int va_get_checkpoints(va_engine *e, int64_t b, int32_t capacity, double *t, double *x, int32_t *count);

/* VA_SYS_TAPE support: compile (NVRTC, sm_100a) the thread-per-trajectory kernels for the recorded system whose device
 * functors are given as CUDA source (va::Tape::cuda_source("VaUserSys"), replacing AadData::Record + the AADC JIT,
 * lib/include/AadData.hpp:124-171). Compile only, no device needed; log receives the compiler output. */
int va_tape_compile_check(const char *tape_cuda_src, int32_t stepper, char *log, int32_t log_capacity);

/* Synthetic, seeded parameter sets generated on the device (bench inputs; bit-identical to the host generator used
 * by the tests). params_dev [B][n_par], x0_dev [B][n_state]; b0 = global index of the first set (sharding). */
int va_synth_batch_device(int32_t system, int32_t n_state, uint64_t seed, int64_t b0, int64_t B, double *params_dev,
                          double *x0_dev, void *stream);

/* Microbenchmarks for the roofline denominators, measured on the device the call runs on. */
int va_measure_fp64_peak(int32_t device, double *tflops);
int va_measure_hbm_copy(int32_t device, double *gbytes_per_s);
/* Host -> device copy ceiling (the denominator of the end-to-end number): `bytes` per device from page-locked host memory to
 * each of devices[0..n), all devices copying CONCURRENTLY, best of `reps`; gbytes_per_s[g] = rate of device g, the aggregate
 * is their sum over the common wall time, returned in *aggregate. `host` may be NULL (the call allocates its own buffers with
 * va_host_alloc(flags)) or a caller buffer of n * bytes (e.g. the bench's own input buffer: measures THAT memory). */
int va_measure_h2d_copy(const int32_t *devices, int32_t n, int64_t bytes, int32_t reps, int32_t flags, const void *host,
                        double *gbytes_per_s, double *aggregate);

/* Page-locked host buffers for VA_MEM_HOST calls. flags: VA_HOST_WRITE_COMBINED for input-only buffers (parameters, x0:
 * the CPU only writes them; uncached, so the GPU's reads need no cache snoops), VA_HOST_NUMA_LOCAL binds the pages to the
 * NUMA node of `device` when the host exposes more than one node. Plain malloc'ed memory also works in every call (the
 * copies then go through the driver's staging and are slower). */
enum va_host_flags { VA_HOST_DEFAULT = 0, VA_HOST_WRITE_COMBINED = 1, VA_HOST_NUMA_LOCAL = 2 };
int va_host_alloc(void **ptr, int64_t bytes, int32_t flags, int32_t device);
int va_host_free(void *ptr);
#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif
