"""Small runs for compute-sanitizer (memcheck / racecheck / initcheck / synccheck) over every kernel family.
usage: compute-sanitizer --tool memcheck python tools/sanitize.py
Round 1: all four tools report 0 errors (memcheck found, and the fix removed, out-of-bounds speculative loads of padded
GLV tile entries for species counts that are not 16/32/64). The N = 256 kernels (cluster pair, TMA ring) were added in
the third session, as was the quad kernel for up to 16 species (N = 16, 10, 5 below run on it)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import vectorizedadjoint_b200 as va

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
    assert (r["status"] == 0).all()
    print(N, "family", info["kernel_family"], info["kernel_name"], "threads", info["threads_per_cta"], "steps", r["n_accept"].tolist(), flush=True)
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
print("done")
