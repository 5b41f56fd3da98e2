#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE ITSELF. Run in the build container only
(needs /root/reference): ``python tests/golden/make_goldens.py``.

Sources of truth, strongest first:
  1. ``printed``  -- stdout of the reference's own prebuilt example binaries
                     (/root/reference/examples/*/debug/*), 6 significant digits.
  2. ``ref17``    -- the UNMODIFIED reference headers (lib/include) + AADC, compiled against the Boost stand-in
                     (oracle/Makefile -> oracle/_ref/libva_ref.so) and driven through the reference's public API;
                     full double precision. Gate: the same build reproduces (1) digit for digit
                     (``shim_examples`` below stores its stdout next to the prebuilt binaries').
The fixtures are what the CPU oracle (oracle/va_oracle.c) and the CUDA path are tested against on machines where
/root/reference does not exist.
"""
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference"
AADC = f"{REF}/aadc/lib/build_full_g++_opt_nosvml_avx2"
OUT = os.path.dirname(os.path.abspath(__file__))
VDP_TOLS = ["1e-3", "1e-4", "1e-5", "1e-6", "1e-7", "1e-8", "1e-9", "1e-10", "1e-12"]


def run(cmd, cwd=None, prebuilt=True):
    env = dict(os.environ)
    if prebuilt:
        env["LD_LIBRARY_PATH"] = AADC
    return subprocess.run(cmd, cwd=cwd, env=env, check=True, capture_output=True, text=True).stdout


def parse_ho(txt):
    g = lambda pat: re.search(pat, txt).group(1)
    return dict(steps=int(g(r"Number of steps: (\d+)")), x=g(r"Solution: r = \[(.*)\]"),
                dEdmr=g(r"dEdmr:(\S+)"), dEdmv=g(r"dEdmv:(\S+)"), dEdmu=g(r"dEdmu:(\S+)"))


def parse_vdp(txt):
    g = lambda pat: re.search(pat, txt).group(1)
    return dict(steps=int(g(r"Number of steps: (\d+)")), x=g(r"Solution: x = \[(.*)\]"),
                mu00=g(r"mu\[0\]\[0\] = (\S+)"), mu10=g(r"mu\[1\]\[0\] = (\S+)"))


def main():
    oracle.build(ref=True)
    gold = {"generated_from": "reference @ /root/reference (prebuilt example binaries + unmodified lib/include + AADC)",
            "printed": {}, "shim_examples": {}, "ref17": {}}

    # 1. the reference's own prebuilt binaries
    gold["printed"]["harmonic"] = parse_ho(run([f"{REF}/examples/HarmonicOscillator/debug/harmonic"]))
    gold["printed"]["vanderpol"] = {t: parse_vdp(run([f"{REF}/examples/VanDerPol/debug/vanderpol", t])) for t in VDP_TOLS}
    # gate for the shim: the unmodified example sources built against it print the same thing
    gold["shim_examples"]["harmonic"] = parse_ho(run([f"{ROOT}/oracle/_ref/harmonic"], prebuilt=False))
    gold["shim_examples"]["vanderpol"] = {t: parse_vdp(run([f"{ROOT}/oracle/_ref/vanderpol", t], prebuilt=False))
                                         for t in VDP_TOLS}
    assert gold["shim_examples"] == gold["printed"], "Boost stand-in does not reproduce the prebuilt reference binaries"

    # 2. full-precision values through the reference's public API
    r17 = gold["ref17"]
    res = oracle.reference_forward_adjoint(oracle.SYS_HARMONIC, 2, oracle.RK_RK4, 0, 0, [[0.0, 1.0]], [[0.151]], 0.0, 10.0, 0.01,
                                           objective=oracle.OBJ_HALF_NORM2)
    r17["harmonic_rk4"] = dict(steps=int(res["n_accept"][0]), x_final=res["x_final"][0].tolist(),
                               lam=res["lam"][0, 0].tolist(), mu=res["mu"][0, 0].tolist())
    mu0 = 1e3
    x0 = [2.0, -2.0 / 3.0 + 10.0 / (81.0 * mu0) - 292.0 / (2187.0 * mu0 * mu0)]
    for name, st in (("rkf78", oracle.RK_RKF78), ("ck54", oracle.RK_CK54)):
        for tol in VDP_TOLS if name == "rkf78" else ["1e-5", "1e-8"]:
            res = oracle.reference_forward_adjoint(oracle.SYS_VANDERPOL, 2, st, float(tol), float(tol), [x0], [[mu0]], 0.0, 0.5,
                                                   1e-3, objective=oracle.OBJ_SEED, seeds=[[[1, 0], [0, 1]]], nout=2)
            r17[f"vanderpol_{name}_{tol}"] = dict(steps=int(res["n_accept"][0]), x_final=res["x_final"][0].tolist(),
                                                  lam=res["lam"][0].tolist(), mu=res["mu"][0].tolist())
    for N in (5, 10):
        al = np.loadtxt(f"{REF}/examples/GeneralizedLotkaVolterra/data/N{N}/alphasfile_cpp.csv", delimiter=",").ravel()
        np.save(os.path.join(OUT, f"glv_data_N{N}_alphas.npy"), al)  # input data of the reference example (not source code)
        for tol in ("1e-5", "1e-8"):
            res = oracle.reference_forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, float(tol), float(tol), [[0.1] * N], [al],
                                                   0.0, 10.0, 1e-3, objective=oracle.OBJ_SUM)
            r17[f"glv_N{N}_ck54_{tol}"] = dict(steps=int(res["n_accept"][0]), x_final=res["x_final"][0].tolist(),
                                               lam=res["lam"][0, 0].tolist(), mu=res["mu"][0, 0].tolist())
    with open(os.path.join(OUT, "reference_goldens.json"), "w") as f:
        json.dump(gold, f, indent=1)

    # 3. synthetic batches (seed 1234, the bench generator) through the reference: arrays -> npz
    arrays = {}
    for N, B in ((16, 8), (64, 4)):
        p = oracle.synth_params(oracle.SYS_GLV, N, 1234, 0, B)
        x0s = oracle.synth_x0(oracle.SYS_GLV, N, p)
        for tol in ("1e-5", "1e-8"):
            res = oracle.reference_forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, float(tol), float(tol), x0s, p, 0.0, 10.0,
                                                   1e-3, objective=oracle.OBJ_SUM, threads=4)
            k = f"glv_N{N}_ck54_{tol}"
            arrays[k + "_steps"] = res["n_accept"]
            arrays[k + "_x_final"] = res["x_final"]
            arrays[k + "_lam"] = res["lam"][:, 0]
            arrays[k + "_mu"] = res["mu"][:, 0]
    B = 32
    p = oracle.synth_params(oracle.SYS_VANDERPOL, 2, 1234, 0, B)
    x0s = oracle.synth_x0(oracle.SYS_VANDERPOL, 2, p)
    for name, st in (("rkf78", oracle.RK_RKF78), ("ck54", oracle.RK_CK54)):
        res = oracle.reference_forward_adjoint(oracle.SYS_VANDERPOL, 2, st, 1e-8, 1e-8, x0s, p, 0.0, 0.5, 1e-3,
                                               objective=oracle.OBJ_SEED, seeds=np.tile([[1.0, 0.0]], (B, 1)), threads=4)
        k = f"vdp_sweep_{name}_1e-8"
        arrays[k + "_steps"] = res["n_accept"]
        arrays[k + "_x_final"] = res["x_final"]
        arrays[k + "_lam"] = res["lam"][:, 0]
        arrays[k + "_mu"] = res["mu"][:, 0]
    arrays["vdp_sweep_params"] = p
    p = oracle.synth_params(oracle.SYS_HARMONIC, 2, 1234, 0, B)
    res = oracle.reference_forward_adjoint(oracle.SYS_HARMONIC, 2, oracle.RK_RK4, 0, 0, oracle.synth_x0(oracle.SYS_HARMONIC, 2, p), p,
                                           0.0, 10.0, 0.01, objective=oracle.OBJ_HALF_NORM2, threads=4)
    arrays["ho_sweep_params"] = p
    arrays["ho_sweep_steps"] = res["n_accept"]
    arrays["ho_sweep_x_final"] = res["x_final"]
    arrays["ho_sweep_lam"] = res["lam"][:, 0]
    arrays["ho_sweep_mu"] = res["mu"][:, 0]
    # 4. GLV through the steppers the first fixture set did not cover: euler (fixed), fehlberg78 (controlled), and error
    #    steppers used un-controlled (the reference's stepper_tag loop takes them through tag inheritance)
    N, B = 16, 4
    p = oracle.synth_params(oracle.SYS_GLV, N, 777, 0, B)
    x0s = oracle.synth_x0(oracle.SYS_GLV, N, p)
    for name, st, tol, tf, dt in (("euler_fixed", oracle.RK_EULER, 0.0, 1.0, 0.01), ("rkf78_1e-8", oracle.RK_RKF78, 1e-8, 10.0, 1e-3),
                                  ("ck54_fixed", oracle.RK_CK54_FIXED, 0.0, 1.0, 0.02), ("rkf78_fixed", oracle.RK_RKF78_FIXED, 0.0, 1.0, 0.02)):
        res = oracle.reference_forward_adjoint(oracle.SYS_GLV, N, st, tol, tol, x0s, p, 0.0, tf, dt, objective=oracle.OBJ_SUM)
        k = f"glv_N{N}_{name}"
        arrays[k + "_steps"] = res["n_accept"]
        arrays[k + "_x_final"] = res["x_final"]
        arrays[k + "_lam"] = res["lam"][:, 0]
        arrays[k + "_mu"] = res["mu"][:, 0]
    # 5. GLV N = 256 (BASELINE config 5), two parameter sets; the 65792-entry gradients are stored as the growth-rate block,
    #    every 16th matrix entry, and the sum (526 KB each in full)
    N, B = 256, 2
    p = oracle.synth_params(oracle.SYS_GLV, N, 1234, 0, B)
    x0s = oracle.synth_x0(oracle.SYS_GLV, N, p)
    res = oracle.reference_forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, 1e-8, 1e-8, x0s, p, 0.0, 10.0, 1e-3, objective=oracle.OBJ_SUM, threads=2)
    k = "glv_N256_ck54_1e-8"
    arrays[k + "_steps"] = res["n_accept"]
    arrays[k + "_x_final"] = res["x_final"]
    arrays[k + "_lam"] = res["lam"][:, 0]
    arrays[k + "_mu_r"] = res["mu"][:, 0, :N]
    arrays[k + "_mu_A_every16"] = res["mu"][:, 0, N::16]
    arrays[k + "_mu_sum"] = res["mu"][:, 0].sum(axis=1)
    # 6. the recorded (tape path) example systems of vectorizedadjoint_b200/examples/tape_systems.hpp through the reference's
    #    AADC recording. Gradients only for the AUTONOMOUS variants: the reference's reverse sweep evaluates every stage at t_n
    #    (detail/backpropagation.hpp:48,127), which is wrong for an explicitly time-dependent right-hand side; for those the
    #    forward sweep (odeint handles t correctly) is pinned and the gradient is checked by finite differences in the tests.
    rng = np.random.default_rng(20261018)
    B = 16
    pend_p = np.array([1.3, 0.15, 0.8]) * (1.0 + 0.2 * rng.uniform(-1, 1, (B, 3)))
    pend_x0 = np.array([0.4, -0.2]) + 0.1 * rng.uniform(-1, 1, (B, 2))
    sw_p = np.array([1.1, 0.4, 0.7]) * (1.0 + 0.2 * rng.uniform(-1, 1, (B, 3)))
    sw_x0 = np.array([0.6, 0.3]) + 0.1 * rng.uniform(-1, 1, (B, 2))
    pend_p[0], pend_x0[0], sw_p[0], sw_x0[0] = [1.3, 0.15, 0.8], [0.4, -0.2], [1.1, 0.4, 0.7], [0.6, 0.3]  # the examples' own inputs
    seeds = rng.standard_normal((B, 2, 2))
    arrays.update(tape_pendulum_params=pend_p, tape_pendulum_x0=pend_x0, tape_switched_params=sw_p, tape_switched_x0=sw_x0, tape_seeds=seeds)
    for sysname, sid, sid_auto, pp, xx, tf in (("pendulum", oracle.SYS_PENDULUM, oracle.SYS_PENDULUM_AUTONOMOUS, pend_p, pend_x0, 2.0),
                                               ("switched", oracle.SYS_SWITCHED, oracle.SYS_SWITCHED_AUTONOMOUS, sw_p, sw_x0, 3.0)):
        for stname, st, tol, dt in (("rk4", oracle.RK_RK4, 0.0, 0.01), ("ck54_1e-8", oracle.RK_CK54, 1e-8, 0.01), ("rkf78_1e-8", oracle.RK_RKF78, 1e-8, 0.01)):
            res = oracle.reference_forward_adjoint(sid_auto, 2, st, tol, tol, xx, pp, 0.0, tf, dt, objective=oracle.OBJ_SEED, seeds=seeds, nout=2)
            k = f"tape_{sysname}_autonomous_{stname}"
            arrays[k + "_steps"] = res["n_accept"]
            arrays[k + "_x_final"] = res["x_final"]
            arrays[k + "_lam"] = res["lam"]
            arrays[k + "_mu"] = res["mu"]
            res = oracle.reference_forward_adjoint(sid, 2, st, tol, tol, xx, pp, 0.0, tf, dt, objective=oracle.OBJ_SEED, seeds=seeds, nout=2)
            k = f"tape_{sysname}_{stname}"
            arrays[k + "_steps"] = res["n_accept"]
            arrays[k + "_x_final"] = res["x_final"]  # forward sweep only (see above)
    # 7. a recorded system wider than the per-lane register budget (16 species, 272 parameters): harvested Lotka-Volterra
    N, B = 16, 8
    p = oracle.synth_params(oracle.SYS_GLV, N, 4242, 0, B)
    x0s = oracle.synth_x0(oracle.SYS_GLV, N, p)
    res = oracle.reference_forward_adjoint(oracle.SYS_HARVESTED_GLV, N, oracle.RK_CK54, 1e-8, 1e-8, x0s, p, 0.0, 10.0, 1e-3, objective=oracle.OBJ_SUM)
    k = "tape_harvested_glv16_ck54_1e-8"
    arrays[k + "_steps"] = res["n_accept"]
    arrays[k + "_x_final"] = res["x_final"]
    arrays[k + "_lam"] = res["lam"][:, 0]
    arrays[k + "_mu"] = res["mu"][:, 0]
    np.savez_compressed(os.path.join(OUT, "reference_synth.npz"), **arrays)
    print("wrote", os.path.join(OUT, "reference_goldens.json"), "and reference_synth.npz")


if __name__ == "__main__":
    main()
