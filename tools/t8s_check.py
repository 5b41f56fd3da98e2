"""GPU check of the warp-specialised N <= 64 kernel (va_glv_t8s.cu) against the second generation (va_glv_t8.cu): the two form
every sum in the same order, so per-trajectory results must be bit-identical. Cases: fewer trajectories than slots, many per slot
(both halves of a slab in use, the accumulate warps' queue under load), padded species counts, two seeds (separate v section,
synchronous hand-over), summed mode, rk4 fixed step, dopri5, the split API and the checkpoints of a first-wave trajectory."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorizedadjoint_b200 as va  # noqa: E402
import oracle  # noqa: E402  (input generator only)


def run(env, N, B, stepper, adaptive, tol, tf, dt0, n_out, seed):
    for k in ("VA_GLV_T8S",):
        os.environ.pop(k, None)
    os.environ.update(env)
    p = oracle.synth_params(oracle.SYS_GLV, N, seed, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    seeds = np.random.default_rng(seed).standard_normal((B, n_out, N))
    with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol, n_out=n_out) as e:
        name = e.info()["kernel_name"]
        t0 = time.time()
        r = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SEED, seeds=seeds)
        s = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
        f = e.forward(x0, p, 0.0, tf, dt0)
        ck = e.checkpoints(min(B - 1, 2))
        a = e.adjoint(objective=va.OBJ_SEED, seeds=seeds)
        dtm = time.time() - t0
    return name, r, s, f, ck, a, dtm


cases = [(64, 5, 2, True, 1e-8, 10.0, 1e-3, 1), (64, 3000, 2, True, 1e-8, 10.0, 1e-3, 1), (50, 700, 2, True, 1e-8, 10.0, 1e-3, 2),
         (64, 1300, 3, True, 1e-6, 10.0, 1e-3, 1), (40, 1500, 1, False, 0.0, 0.5, 0.01, 1), (64, 600, 2, True, 1e-8, 10.0, 1e-3, 2),
         (33, 1, 2, True, 1e-8, 10.0, 1e-3, 1), (64, 20000, 2, True, 1e-8, 10.0, 1e-3, 1)]
ok = True
for c in cases:
    N, B, stepper, adaptive, tol, tf, dt0, n_out = c
    base = run({}, N, B, stepper, adaptive, tol, tf, dt0, n_out, 77)
    spec = run({"VA_GLV_T8S": "1"}, N, B, stepper, adaptive, tol, tf, dt0, n_out, 77)
    assert base[0] == "k_glv_t8" and spec[0] == "k_glv_t8s", (base[0], spec[0])
    same = True
    for k in ("x_final", "lam", "mu", "n_accept", "n_reject", "status"):
        same &= np.array_equal(base[1][k], spec[1][k], equal_nan=True)
    same &= np.array_equal(base[2]["mu"], spec[2]["mu"]) and np.array_equal(base[2]["lam"], spec[2]["lam"])
    same &= np.array_equal(base[3]["x_final"], spec[3]["x_final"])
    same &= np.array_equal(base[4][0], spec[4][0]) and np.array_equal(base[4][1], spec[4][1])
    same &= np.array_equal(base[5]["mu"], spec[5]["mu"]) and np.array_equal(base[5]["lam"], spec[5]["lam"])
    same &= np.array_equal(base[5]["mu"], base[1]["mu"])  # split API == fused call
    dmax = float(np.nanmax(np.abs(base[1]["mu"] - spec[1]["mu"])))
    print("T8SCHECK", c, "bit-identical" if same else f"DIFFERENT (max |d mu| {dmax:.3e})", f"{base[6]:.2f}s {spec[6]:.2f}s", flush=True)
    ok &= bool(same)
print("T8SCHECK", "ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
