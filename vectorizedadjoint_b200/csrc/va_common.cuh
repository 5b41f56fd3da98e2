// va_common.cuh -- types shared by the engine (va_engine.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "va_types.h"

int va_tableau_host(int kind, VaTableau *tb); // va_tableau.cpp
cudaError_t va_scalar_forward(const VaScalarArgs &a, cudaStream_t st);
cudaError_t va_scalar_adjoint(const VaScalarArgs &a, cudaStream_t st);

// CTA-per-trajectory GLV family (va_glv_wide.cu). One persistent CTA integrates trajectory after trajectory, forward
// then backward, with its private checkpoint slab (L2-resident for typical step counts).
struct VaGlvWideArgs {
    int n;                // species (<= 64 for the NP=64 kernel; padded with inert species)
    int stepper, adaptive, n_out, objective, reduce;
    double eps_abs, eps_rel, ti, tf, dt0;
    int64_t B;
    int cap;              // accepted-step capacity of the slab
    const double *x0, *params;
    double *x_final, *lambda, *mu;
    int32_t *n_accept, *n_reject, *status;
    double *slab;         // (grid * slots per CTA) * slab_stride doubles
    int64_t slab_stride;
    double *partial;      // VA_REDUCE_SUM: [grid * slots per CTA][n_par] per-CTA partial sums (reduced by va_reduce_rows)
    int grid;
    int blk_doubles;      // va_glv_t8.cu: doubles per step block (the v section is separate when n_out > 1)
    int recompute;        // streamed family: 1 = keep only (t_n, x_n) and recompute the stages in the reverse sweep
    struct { double a[7][6], b[7], db[7]; } coef; // tableau values, filled by the launcher
    // Fields added after the headline kernel was tuned go HERE, behind everything va_glv_t8.cu reads: inserting them above
    // moved the constant-bank offsets of `coef` and cost that kernel 3 registers, 104 bytes of spills and 4 % (5.92 -> 5.66 M).
    int cluster;          // cluster kernel (va_glv_pair.cu): CTAs per trajectory (2 or 4)
    int flags;            // ring kernel: bit 1 = evict_last policy on the matrix stream, bit 2 = no register-cached rows
    // cluster kernel, recompute policy (a.recompute != 0): per-CTA state store of (cap + 1) entries [8-double header (t_n) | x_n]
    // and the segment length (steps re-integrated at a time; the slab then holds seg_len blocks)
    double *xstore;
    int64_t xstore_stride;
    int seg_len;
    int sparse;           // cluster kernel, recompute policy: keep t_n of every step but x_n only of every seg_len-th (VA_CKPT_SPARSE)
    int skip_forward;     // va_glv_oct.cu, va_glv_t8.cu: the slabs still hold the forward sweep of exactly these trajectories (split API,
                          // runge_kutta then adjointSolve): take T, status and x(tf) from n_accept / status / x_final and sweep back only
};
// VaGlvWideArgs::flags, va_glv_t8.cu: the step blocks are dead once the gradient accumulation has read them (the batch is larger than
// one wave of slots, so no later call can use the slabs): drop their L2 lines instead of letting them be written back
#define VA_GLV_FLAG_DISCARD 0x100
bool va_glv_wide_supported(int n, int stepper, int adaptive);
int64_t va_glv_wide_slab_doubles(int n, int stepper, int cap);
int va_glv_wide_pair();                            // trajectories a slot integrates forward together (slabs per slot)
int va_glv_wide_padded(int n);                      // padded species count the kernel runs with (16, 32 or 64)
int va_glv_wide_block_doubles(int n, int stepper); // step block: [8-double header (t_n) | X_0..X_{s-1} | g_0..g_{s-1}], padded width
cudaError_t va_glv_wide_config(int n, int stepper, int device, int *grid, int *ctas_per_sm, int *threads, int *slots_per_cta);
cudaError_t va_glv_wide_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st);

// second-generation register kernel for 33..64 species (va_glv_t8.cu): 64 threads per trajectory, 8x8 tiles, three phases
bool va_glv_t8_supported(int n, int stepper, int adaptive);
int va_glv_t8_block_doubles(int stepper, int n_out);
int va_glv_t8_header_doubles();
cudaError_t va_glv_t8_config(int n, int stepper, int n_out, int device, int *grid, int *ctas_per_sm, int *threads, int *slots_per_cta);
cudaError_t va_glv_t8_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st);

// third generation for 33..64 species (va_glv_t8s.cu): sweeps and gradient accumulation on different warps (setmaxnreg register
// split, job queue between them); a slot's slab has two halves of (cap + 1) step blocks
bool va_glv_t8s_supported(int n, int stepper, int adaptive);
int va_glv_t8s_block_doubles(int stepper, int n_out);
int64_t va_glv_t8s_slab_doubles(int stepper, int n_out, int cap);
cudaError_t va_glv_t8s_config(int n, int stepper, int n_out, int device, int *grid, int *ctas_per_sm, int *threads, int *slots_per_cta);
cudaError_t va_glv_t8s_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st);

// quad kernel for up to 16 species (va_glv_quad.cu): four lanes per trajectory, eight trajectories per warp, three phases
bool va_glv_quad_supported(int n, int stepper, int adaptive);
int va_glv_quad_block_doubles(int stepper, int n_out);
int va_glv_quad_slots_per_cta();
int va_glv_quad_threads();
cudaError_t va_glv_quad_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st);

// second generation for up to 16 species (va_glv_oct.cu): eight lanes per trajectory, A and Abar both resident, recompute policy
// (only (t_n, x_n) is stored; checkpoints [8-double header | x_n] per accepted step)
bool va_glv_oct_supported(int n, int stepper, int adaptive);
int va_glv_oct_block_doubles();
int va_glv_oct_slots_per_cta();
int va_glv_oct_threads();
cudaError_t va_glv_oct_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st);

// streamed-matrix GLV family (va_glv_stream.cu): any N, one 256-thread CTA per trajectory, same argument block
bool va_glv_stream_supported(int n, int stepper, int adaptive);
int va_glv_stream_block_doubles(int n, int stepper, int recompute);
cudaError_t va_glv_stream_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st);

// ring-streamed GLV kernel for 256 species (va_glv_ring.cu): one 256-thread CTA per SM, matrix streamed through a TMA ring,
// store-stages policy; a.flags: bit 1 = evict_last matrix stream, bit 2 = no register-cached rows
bool va_glv_ring_supported(int n, int stepper, int adaptive);
int va_glv_ring_block_doubles(int stepper);
size_t va_glv_ring_smem();
cudaError_t va_glv_ring_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st);

// cluster GLV kernel for 256 species (va_glv_pair.cu): a.cluster = 2 or 4 CTAs (one thread-block cluster) per trajectory hold
// the matrix on chip (registers, + shared memory at 2) and exchange product parts through distributed shared memory; a.grid
// must be a multiple of a.cluster
bool va_glv_pair_supported(int n, int stepper, int adaptive);
int va_glv_pair_block_doubles(int stepper);
cudaError_t va_glv_pair_max_clusters(int stepper, int cluster, int sm_count, int *n);
cudaError_t va_glv_pair_forward_adjoint(const VaGlvWideArgs &a, cudaStream_t st);

// out[k] (+)= sum_{g<G} in[g*stride + k], deterministic order
cudaError_t va_reduce_rows(const double *in, int64_t G, int64_t stride, int64_t n, double *out, int accumulate, cudaStream_t st);

// synthetic inputs / microbenchmarks (va_util.cu)
cudaError_t va_synth_launch(int system, int n, uint64_t seed, int64_t b0, int64_t B, double *params, double *x0, cudaStream_t st);
