"""Small runs for compute-sanitizer (memcheck / racecheck / initcheck / synccheck) over every kernel family.
usage: compute-sanitizer --tool memcheck python tools/sanitize.py
Round 1: all four tools report 0 errors (memcheck found, and the fix removed, out-of-bounds speculative loads of padded
GLV tile entries for species counts that are not 16/32/64). The N = 256 kernels (cluster pair, TMA ring) were added in
the third session, as was the quad kernel for up to 16 species. Round 2: up to 16 species now run on the eight-lane kernel
(va_glv_oct.cu; N = 16, 10, 5 below), the quad kernel is reached through the store-stages policy; added: split API with the
sweep-back-only adjoint, the streamed family with euler / fehlberg78 / fixed-step error steppers, the sparse checkpoint
policy on the cluster kernel, a recorded (NVRTC) system, and -- with two GPUs -- the multi-device engine."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import vectorizedadjoint_b200 as va

# round 2, second session: the warp-specialised 33..64-species kernel (setmaxnreg, job queue, two slab halves) and the L2 discards
# of the second generation, on two persistent CTAs so that every slot integrates several trajectories
for env, N, stepper, adaptive, tol, n_out, B in [({"VA_GLV_T8S": "1"}, 64, va.RK_CK54, True, 1e-6, 1, 29), ({"VA_GLV_T8S": "1"}, 50, va.RK_DOPRI5, True, 1e-5, 2, 19),
                                                 ({"VA_GLV_T8S": "1"}, 40, va.RK_RK4, False, 0.0, 1, 21), ({"VA_T8_DISCARD": "1"}, 64, va.RK_CK54, True, 1e-6, 1, 29),
                                                 ({"VA_T8_DISCARD": "1"}, 50, va.RK_CK54, True, 1e-6, 2, 19)]:
    os.environ.update(env)
    os.environ["VA_GLV_MAX_CTAS"] = "2"
    p = oracle.synth_params(oracle.SYS_GLV, N, 8, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    tf, dt0 = (10.0, 1e-3) if adaptive else (0.2, 0.01)
    seeds = np.random.default_rng(0).standard_normal((B, n_out, N))
    with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol, n_out=n_out) as e:
        r = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SEED, seeds=seeds)
        s = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SEED, seeds=seeds, reduce=va.REDUCE_SUM)
        info = e.info()
    for k in list(env) + ["VA_GLV_MAX_CTAS"]:
        os.environ.pop(k)
    assert (r["status"] == 0).all()
    np.testing.assert_allclose(s["mu"], r["mu"].sum(axis=0), rtol=1e-10, atol=1e-300)
    print(N, info["kernel_name"], env, "ctas", info["grid"] if "grid" in info else "?", "steps", r["n_accept"].tolist(), flush=True)
if os.environ.get("VA_SANITIZE_ONLY") == "t8s":
    print("sanitize: t8s / discard section only")
    sys.exit(0)
for N, stepper, adaptive, tol, n_out, B in [(64, va.RK_CK54, True, 1e-6, 1, 9), (50, va.RK_DOPRI5, True, 1e-5, 2, 5), (64, va.RK_RK4, False, 0.0, 1, 3),
                                            (16, va.RK_CK54, True, 1e-6, 1, 9), (10, va.RK_CK54, True, 1e-6, 2, 7), (5, va.RK_DOPRI5, True, 1e-6, 1, 3),
                                            (20, va.RK_RK4, False, 0.0, 1, 3), (33, va.RK_CK54, True, 1e-5, 1, 3), (100, va.RK_CK54, True, 1e-5, 1, 2),
                                            (129, va.RK_CK54, True, 1e-5, 2, 2), (300, va.RK_CK54, True, 1e-5, 1, 1)]:
    p = oracle.synth_params(oracle.SYS_GLV, N, 5, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    tf, dt0 = (10.0, 1e-3) if adaptive else (0.2, 0.01)
    seeds = np.random.default_rng(0).standard_normal((B, n_out, N))
    with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol, n_out=n_out) as e:
        r = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SEED, seeds=seeds)
        s = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SEED, seeds=seeds, reduce=va.REDUCE_SUM)
        info = e.info()
        if n_out == 1:  # split API: forward, then the adjoint over the blocks / checkpoints the forward call left
            e.forward(x0, p, 0.0, tf, dt0)
            e.checkpoints(B - 1)
            sa = e.adjoint(objective=va.OBJ_SEED, seeds=seeds)
            assert np.array_equal(sa["mu"], r["mu"])
    assert (r["status"] == 0).all()
    print(N, "family", info["kernel_family"], info["kernel_name"], "threads", info["threads_per_cta"], "steps", r["n_accept"].tolist(), flush=True)
# round 2: the four-lane store-stages kernel, every reference tableau on the streamed family, sparse checkpoints
for N, stepper, adaptive, tol, policy, ms in [(16, va.RK_CK54, True, 1e-6, va.CKPT_STORE_STAGES, 0), (12, va.RK_EULER, False, 0.0, va.CKPT_AUTO, 64),
                                              (40, va.RK_RKF78, True, 1e-6, va.CKPT_AUTO, 0), (16, va.RK_CK54, False, 0.0, va.CKPT_RECOMPUTE, 64),
                                              (256, va.RK_CK54, True, 1e-5, va.CKPT_SPARSE, 0), (130, va.RK_DOPRI5, True, 1e-5, va.CKPT_SPARSE, 0)]:
    B = 3
    p = oracle.synth_params(oracle.SYS_GLV, N, 6, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    tf, dt0 = (10.0, 1e-3) if adaptive else (0.3, 0.01)
    os.environ["VA_PAIR_SEG"] = "4"
    with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol, ckpt_policy=policy, max_steps=ms) as e:
        r = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM)
        s = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
        info = e.info()
    os.environ.pop("VA_PAIR_SEG")
    assert (r["status"] == 0).all()
    print(N, info["kernel_name"], "stepper", stepper, "policy", info["ckpt_policy"], "steps", r["n_accept"].tolist(), flush=True)
# 256 species: cluster-pair kernel (DSMEM exchange, TMA loads), ring-streamed kernel, plain streamed kernel
for env, policy, B in [({}, va.CKPT_STORE_STAGES, 3), ({"VA_PAIR_SEG": "3"}, va.CKPT_RECOMPUTE, 2), ({"VA_GLV_NO_PAIR": "1"}, va.CKPT_AUTO, 2),
                       ({"VA_GLV_NO_RING": "1"}, va.CKPT_AUTO, 1)]:
    os.environ.update(env)  # recompute on the cluster kernel = state store + segment re-integration (3 steps per segment here)
    p = oracle.synth_params(oracle.SYS_GLV, 256, 5, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, 256, p)
    seeds = np.random.default_rng(0).standard_normal((B, 2, 256))
    with va.Engine(va.SYS_GLV, 256, va.RK_CK54, True, 1e-5, 1e-5, n_out=2, ckpt_policy=policy) as e:
        r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds)
        s = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds, reduce=va.REDUCE_SUM)
        info = e.info()
    assert (r["status"] == 0).all()
    print(256, info["kernel_name"], "policy", info["ckpt_policy"], "steps", r["n_accept"].tolist(), flush=True)
# thread-per-trajectory family: Van der Pol (all adaptive steppers), harmonic oscillator (fixed step), waves + summed mode
for system, stepper, adaptive, tol, tf, dt0, B in [(va.SYS_VANDERPOL, va.RK_DOPRI5, True, 1e-6, 0.5, 1e-3, 300), (va.SYS_VANDERPOL, va.RK_RKF78, True, 1e-6, 0.5, 1e-3, 70),
                                                   (va.SYS_VANDERPOL, va.RK_CK54, True, 1e-5, 0.5, 1e-3, 33), (va.SYS_HARMONIC, va.RK_RK4, False, 0.0, 1.0, 0.01, 129)]:
    osys = oracle.SYS_VANDERPOL if system == va.SYS_VANDERPOL else oracle.SYS_HARMONIC
    p = oracle.synth_params(osys, 2, 7, 0, B)
    x0 = oracle.synth_x0(osys, 2, p)
    with va.Engine(system, 2, stepper, adaptive, tol, tol, max_steps=1024) as e:
        r = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM)
        s = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
        f = e.forward(x0, p, 0.0, tf, dt0)
        a = e.adjoint(objective=va.OBJ_SUM)
    assert (r["status"] == 0).all()
    print("system", system, "stepper", stepper, "steps", int(r["n_accept"].min()), "..", int(r["n_accept"].max()), flush=True)
# recorded system (tape -> CUDA -> NVRTC): the generated functor inside the thread-per-trajectory kernels
SRC = """struct VaUserSys { static constexpr int N = 2, NPAR = 2;
  __device__ static void rhs(const double *x, const double *p, double t, double *dx) { dx[0] = x[1]; dx[1] = -p[0] * sin(x[0]) - p[1] * x[1]; }
  __device__ static void vjp(const double *x, const double *p, double t, const double *w, double *gx, double *gp) {
    gx[0] = -p[0] * cos(x[0]) * w[1]; gx[1] = w[0] - p[1] * w[1]; gp[0] += -sin(x[0]) * w[1]; gp[1] += -x[1] * w[1]; } };"""
with va.Engine(va.SYS_TAPE, 2, va.RK_DOPRI5, True, 1e-7, 1e-7, n_par=2, max_steps=512, tape_cuda_src=SRC) as e:
    r = e.forward_adjoint(np.tile([0.4, -0.2], (200, 1)), np.tile([1.3, 0.15], (200, 1)) * (1 + 0.1 * np.random.default_rng(1).random((200, 2))), 0.0, 2.0, 0.01,
                          objective=va.OBJ_SUM)
    assert (r["status"] == 0).all()
    print("tape: steps", int(r["n_accept"].min()), "..", int(r["n_accept"].max()), flush=True)
# several GPUs: worker threads, NCCL all-reduce inside the call
import torch
if torch.cuda.device_count() >= 2:
    p = oracle.synth_params(oracle.SYS_GLV, 64, 5, 0, 11)
    x0 = oracle.synth_x0(oracle.SYS_GLV, 64, p)
    with va.Engine(va.SYS_GLV, 64, va.RK_CK54, True, 1e-6, 1e-6, devices=[0, 1]) as e:
        r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
    print("multi-device: ok", flush=True)
print("done")
