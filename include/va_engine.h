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

#define VA_API_VERSION 1

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
    VA_CKPT_STORE_STAGES = 2 /* additionally store the stage states/slopes of accepted steps; no recompute            */
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
    int32_t device;      /* CUDA ordinal                                                                       */
    int32_t max_steps;   /* checkpoint capacity per trajectory (accepted steps); 0 = default                   */
    int32_t ckpt_policy; /* va_ckpt_policy                                                                     */
    int32_t reserved0;
    double workspace_fraction; /* share of free HBM the checkpoint arena may take (0 = default 0.5)            */
    const char *tape_cuda_src; /* VA_SYS_TAPE: CUDA source of the rhs/vjp device functors (see va_tape.h)      */
} va_engine_desc;

typedef struct va_batch_args {
    int64_t batch;        /* B parameter sets                                                                   */
    const double *x0;     /* [B][n_state]                                                                       */
    const double *params; /* [B][n_par]                                                                         */
    double ti, tf, dt0;
    int32_t objective;    /* va_objective                                                                       */
    int32_t reduce;       /* va_reduce                                                                          */
    int32_t mem;          /* va_mem: where x0/params/x_final/lambda/mu/n_accept/... live                        */
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
} va_engine_info;

#ifndef __CUDACC_RTC__ /* the enums and structs above are also seen by run-time compiled device code */
int va_engine_create(const va_engine_desc *desc, va_engine **out);
void va_engine_destroy(va_engine *e);
int va_engine_get_info(va_engine *e, va_engine_info *info);
const char *va_last_error(void);

int va_forward_batch(va_engine *e, const va_batch_args *args);
int va_adjoint_batch(va_engine *e, const va_batch_args *args);
int va_forward_adjoint_batch(va_engine *e, const va_batch_args *args);

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
#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif
