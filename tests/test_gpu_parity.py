"""GPU parity tests: the CUDA path (through the C-ABI, libva_engine.so) against the CPU oracle on identical seeded
inputs, against the committed reference goldens, and through size-independent properties.

Tolerances (BASELINE.json north_star): accepted-step counts equal per trajectory except rounding-induced
accept/reject flips; final states and adjoint gradients within 1e-8 relative in FP64.
"""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-8


@pytest.fixture(scope="module")
def va():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import vectorizedadjoint_b200 as va
    va.lib()  # fails loudly when the extension is missing
    return va


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), floor) if floor else np.abs(b).max() if b.size else 1.0
    return np.abs(a - b).max() / scale if not floor else (np.abs(a - b) / scale).max()


def assert_close(a, b, rtol=RTOL, what=""):
    """max-norm relative agreement per array row (a gradient is judged as a vector, not entry by entry)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    a2, b2 = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    scale = np.abs(b2).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    err = (np.abs(a2 - b2) / scale).max()
    assert err <= rtol, f"{what}: relative error {err:.3e} > {rtol:.1e}"


def assert_flips_bounded(r, o, same, tol):
    """north_star: 'accepted-step counts must match per trajectory except documented rounding-induced accept/reject flips'.
    A flipped trajectory took a different (equally valid) step sequence: it is a different discretisation of the same ODE, so it
    must still agree with the oracle's to the order of the integration tolerance -- bounded here at 10 * tol relative (never
    silently dropped from the comparison)."""
    flipped = ~same
    if not flipped.any():
        return
    bound = 10.0 * max(tol, 1e-12)
    assert np.abs(r["n_accept"][flipped] - o["n_accept"][flipped]).max() <= 2, "a flip changes the step count by one or two"
    assert_close(r["x_final"][flipped], o["x_final"][flipped], rtol=bound, what="x(tf), flipped trajectories")
    assert_close(r["lam"][flipped, 0], o["lam"][flipped], rtol=bound, what="lambda, flipped trajectories")
    assert_close(r["mu"][flipped, 0], o["mu"][flipped], rtol=bound, what="mu, flipped trajectories")


# ---------------------------------------------------------------------------------------------------------------------
# small systems, thread per trajectory
# ---------------------------------------------------------------------------------------------------------------------

def test_harmonic_golden_and_sweep(va, goldens, synth_goldens):
    g = goldens["ref17"]["harmonic_rk4"]
    with va.Engine(va.SYS_HARMONIC, 2, va.RK_RK4, False, max_steps=1024) as e:
        r = e.forward_adjoint([[0.0, 1.0]], [[0.151]], 0.0, 10.0, 0.01, objective=va.OBJ_HALF_NORM2)
        assert r["n_accept"][0] == g["steps"] == 1000 and r["status"][0] == 0
        np.testing.assert_allclose(r["x_final"][0], g["x_final"], rtol=1e-14)
        np.testing.assert_allclose(r["lam"][0, 0], g["lam"], rtol=1e-12)
        np.testing.assert_allclose(r["mu"][0, 0], g["mu"], rtol=1e-12)
        # seeded sweep vs the reference goldens and vs the oracle
        p = synth_goldens["ho_sweep_params"]
        r = e.forward_adjoint(oracle.synth_x0(oracle.SYS_HARMONIC, 2, p), p, 0.0, 10.0, 0.01, objective=va.OBJ_HALF_NORM2)
        np.testing.assert_array_equal(r["n_accept"], synth_goldens["ho_sweep_steps"])
        np.testing.assert_allclose(r["x_final"], synth_goldens["ho_sweep_x_final"], rtol=1e-13)
        np.testing.assert_allclose(r["lam"][:, 0], synth_goldens["ho_sweep_lam"], rtol=1e-11)
        np.testing.assert_allclose(r["mu"][:, 0], synth_goldens["ho_sweep_mu"], rtol=1e-11)
        o = oracle.forward_adjoint(oracle.SYS_HARMONIC, 2, oracle.RK_RK4, False, 0, 0, oracle.synth_x0(oracle.SYS_HARMONIC, 2, p), p,
                                   0.0, 10.0, 0.01, objective=oracle.OBJ_HALF_NORM2)
        # same operation order, no FMA contraction: the forward sweep is bit-identical
        np.testing.assert_array_equal(r["x_final"], o["x_final"])


@pytest.mark.parametrize("tol", ["1e-3", "1e-4", "1e-5", "1e-6", "1e-7", "1e-8", "1e-9", "1e-10", "1e-12"])
def test_vanderpol_reference_step_counts(va, goldens, tol):
    """mu = 1e3, RKF78: the nine accepted-step counts printed by the reference binary, and its gradients."""
    mu0 = 1e3
    x0 = [2.0, -2.0 / 3.0 + 10.0 / (81.0 * mu0) - 292.0 / (2187.0 * mu0 * mu0)]
    g = goldens["ref17"][f"vanderpol_rkf78_{tol}"]
    with va.Engine(va.SYS_VANDERPOL, 2, va.RK_RKF78, True, float(tol), float(tol), n_out=2, max_steps=2048) as e:
        r = e.forward_adjoint([x0], [[mu0]], 0.0, 0.5, 1e-3, objective=va.OBJ_SEED, seeds=[[[1, 0], [0, 1]]])
    assert r["n_accept"][0] == g["steps"] == goldens["printed"]["vanderpol"][tol]["steps"]
    assert_close(r["x_final"], [g["x_final"]], what="x(tf)")
    assert_close(r["lam"][0], g["lam"], what="lambda")
    assert_close(r["mu"][0], g["mu"], what="mu")


@pytest.mark.parametrize("stepper,name", [(4, "rkf78"), (2, "ck54"), (3, "dopri5")])
@pytest.mark.parametrize("tol", [1e-5, 1e-8])
def test_vanderpol_sweep_vs_oracle(va, stepper, name, tol):
    B = 512
    p = oracle.synth_params(oracle.SYS_VANDERPOL, 2, 1234, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_VANDERPOL, 2, p)
    seeds = np.tile([[1.0, 0.0]], (B, 1))
    o = oracle.forward_adjoint(oracle.SYS_VANDERPOL, 2, stepper, True, tol, tol, x0, p, 0.0, 0.5, 1e-3, objective=oracle.OBJ_SEED,
                               seeds=seeds, threads=8)
    with va.Engine(va.SYS_VANDERPOL, 2, stepper, True, tol, tol, max_steps=2048) as e:
        r = e.forward_adjoint(x0, p, 0.0, 0.5, 1e-3, objective=va.OBJ_SEED, seeds=seeds)
    assert (r["status"] == 0).all()
    # Same operation order, no FMA contraction, glibc-exact pow() in the controller (csrc/va_pow.h): the forward sweep
    # is bit-identical to the oracle -- every accept/reject decision, every step size, the final state.
    np.testing.assert_array_equal(r["n_accept"], o["n_accept"])
    np.testing.assert_array_equal(r["n_reject"], o["n_reject"])
    np.testing.assert_array_equal(r["x_final"], o["x_final"])
    assert_close(r["lam"][:, 0], o["lam"], rtol=1e-12, what="lambda")
    assert_close(r["mu"][:, 0], o["mu"], rtol=1e-12, what="mu")


def test_vanderpol_sweep_vs_reference_goldens(va, synth_goldens):
    p = synth_goldens["vdp_sweep_params"]
    B = p.shape[0]
    x0 = oracle.synth_x0(oracle.SYS_VANDERPOL, 2, p)
    for name, st in (("rkf78", va.RK_RKF78), ("ck54", va.RK_CK54)):
        with va.Engine(va.SYS_VANDERPOL, 2, st, True, 1e-8, 1e-8, max_steps=2048) as e:
            r = e.forward_adjoint(x0, p, 0.0, 0.5, 1e-3, objective=va.OBJ_SEED, seeds=np.tile([[1.0, 0.0]], (B, 1)))
        k = f"vdp_sweep_{name}_1e-8"
        np.testing.assert_array_equal(r["n_accept"], synth_goldens[k + "_steps"])
        assert_close(r["x_final"], synth_goldens[k + "_x_final"], what="x(tf)")
        assert_close(r["lam"][:, 0], synth_goldens[k + "_lam"], what="lambda")
        assert_close(r["mu"][:, 0], synth_goldens[k + "_mu"], what="mu")


def test_scalar_waves_split_api_and_edge_cases(va):
    B = 1000
    p = oracle.synth_params(oracle.SYS_VANDERPOL, 2, 99, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_VANDERPOL, 2, p)
    with va.Engine(va.SYS_VANDERPOL, 2, va.RK_DOPRI5, True, 1e-6, 1e-6, max_steps=1024) as e:
        full = e.forward_adjoint(x0, p, 0.0, 0.5, 1e-3, objective=va.OBJ_SUM)
        # reduce = sum equals the sum of the per-trajectory gradients
        s = e.forward_adjoint(x0, p, 0.0, 0.5, 1e-3, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
        np.testing.assert_allclose(s["mu"], full["mu"].sum(axis=0), rtol=1e-12)
        # split API: forward, checkpoints, adjoint
        f = e.forward(x0, p, 0.0, 0.5, 1e-3)
        np.testing.assert_array_equal(f["x_final"], full["x_final"])
        np.testing.assert_array_equal(f["n_accept"], full["n_accept"])
        t, x = e.checkpoints(7)
        assert len(t) == f["n_accept"][7] + 1 and t[0] == 0.0 and abs(t[-1] - 0.5) < 1e-15
        np.testing.assert_array_equal(x[0], x0[7])
        np.testing.assert_array_equal(x[-1], f["x_final"][7])
        assert (np.diff(t) > 0).all()
        a = e.adjoint(objective=va.OBJ_SUM)
        np.testing.assert_array_equal(a["lam"], full["lam"])
        np.testing.assert_array_equal(a["mu"], full["mu"])
        # empty batch
        r0 = e.forward_adjoint(np.zeros((0, 2)), np.zeros((0, 1)), 0.0, 0.5, 1e-3)
        assert r0["x_final"].shape == (0, 2)
    # a small arena forces several waves; results must not change
    with va.Engine(va.SYS_VANDERPOL, 2, va.RK_DOPRI5, True, 1e-6, 1e-6, max_steps=1024, workspace_fraction=1e-6) as e:
        w = e.forward_adjoint(x0, p, 0.0, 0.5, 1e-3, objective=va.OBJ_SUM)
        assert e.info()["chunk_trajectories"] < B
        for k in ("x_final", "lam", "mu", "n_accept"):
            np.testing.assert_array_equal(w[k], full[k])
    # checkpoint overflow is reported per trajectory, the batch is not aborted
    with va.Engine(va.SYS_VANDERPOL, 2, va.RK_DOPRI5, True, 1e-6, 1e-6, max_steps=20) as e:
        r = e.forward_adjoint(x0, p, 0.0, 0.5, 1e-3, objective=va.OBJ_SUM)
        over = full["n_accept"] > 20
        assert over.any() and (~over).any()
        assert ((r["status"] & va.TRAJ_CKPT_OVERFLOW) != 0)[over].all() and (r["status"][~over] == 0).all()
        np.testing.assert_array_equal(r["mu"][~over], full["mu"][~over])
        assert np.isnan(r["mu"][over]).all()


def test_device_buffers_match_host_buffers(va):
    import torch
    B = 4096
    p = oracle.synth_params(oracle.SYS_VANDERPOL, 2, 5, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_VANDERPOL, 2, p)
    with va.Engine(va.SYS_VANDERPOL, 2, va.RK_CK54, True, 1e-6, 1e-6, max_steps=1024) as e:
        h = e.forward_adjoint(x0, p, 0.0, 0.5, 1e-3, objective=va.OBJ_SUM)
        dev = torch.device("cuda:0")
        tx0, tp = torch.from_numpy(x0).to(dev), torch.from_numpy(p).to(dev)
        xf, lam, mu = torch.empty(B, 2, dtype=torch.float64, device=dev), torch.empty(B, 1, 2, dtype=torch.float64, device=dev), \
            torch.empty(B, 1, 1, dtype=torch.float64, device=dev)
        na = torch.empty(B, dtype=torch.int32, device=dev)
        e.call("va_forward_adjoint_batch", B, tx0, tp, 0.0, 0.5, 1e-3, xf, lam, mu, va.OBJ_SUM, va.REDUCE_NONE, na)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(xf.cpu().numpy(), h["x_final"])
        np.testing.assert_array_equal(lam.cpu().numpy(), h["lam"])
        np.testing.assert_array_equal(mu.cpu().numpy(), h["mu"])
        np.testing.assert_array_equal(na.cpu().numpy(), h["n_accept"])
        # device generator == host generator, bit for bit
        gp, gx = torch.empty_like(tp), torch.empty_like(tx0)
        va.synth_batch_device(va.SYS_VANDERPOL, 2, 5, 0, B, gp, gx)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(gp.cpu().numpy(), p)
        np.testing.assert_array_equal(gx.cpu().numpy(), x0)


# ---------------------------------------------------------------------------------------------------------------------
# Generalized Lotka-Volterra, CTA per trajectory
# ---------------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("N", [5, 10])
@pytest.mark.parametrize("tol", ["1e-5", "1e-8"])
def test_glv_reference_example_data(va, goldens, N, tol):
    """The reference's own parameter files (data/N5, data/N10) through the 64-wide kernel (inert padding species)."""
    al = np.load(os.path.join(GOLD, f"glv_data_N{N}_alphas.npy"))
    g = goldens["ref17"][f"glv_N{N}_ck54_{tol}"]
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, float(tol), float(tol)) as e:
        r = e.forward_adjoint([[0.1] * N], [al], 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
    assert r["status"][0] == 0 and r["n_accept"][0] == g["steps"]
    assert_close(r["x_final"], [g["x_final"]], what="x(tf)")
    assert_close(r["lam"][0], [g["lam"]], what="lambda")
    assert_close(r["mu"][0], [g["mu"]], what="mu")


@pytest.mark.parametrize("N,B", [(16, 8), (64, 4)])
@pytest.mark.parametrize("tol", ["1e-5", "1e-8"])
def test_glv_synthetic_vs_reference_goldens(va, synth_goldens, N, B, tol):
    p = oracle.synth_params(oracle.SYS_GLV, N, 1234, 0, B)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, float(tol), float(tol)) as e:
        r = e.forward_adjoint(oracle.synth_x0(oracle.SYS_GLV, N, p), p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
    k = f"glv_N{N}_ck54_{tol}"
    np.testing.assert_array_equal(r["n_accept"], synth_goldens[k + "_steps"])
    assert_close(r["x_final"], synth_goldens[k + "_x_final"], what="x(tf)")
    assert_close(r["lam"][:, 0], synth_goldens[k + "_lam"], what="lambda")
    assert_close(r["mu"][:, 0], synth_goldens[k + "_mu"], what="mu")


@pytest.mark.parametrize("N,stepper,adaptive,tol,tf,dt0", [
    (64, 2, True, 1e-8, 10.0, 1e-3), (64, 2, True, 1e-5, 10.0, 1e-3), (64, 3, True, 1e-8, 10.0, 1e-3),
    (64, 1, False, 0.0, 0.5, 0.01), (16, 2, True, 1e-8, 10.0, 1e-3), (33, 2, True, 1e-6, 10.0, 1e-3), (1, 2, True, 1e-8, 10.0, 1e-3)])
def test_glv_batch_vs_oracle(va, N, stepper, adaptive, tol, tf, dt0):
    """More trajectories than resident CTAs: every persistent CTA walks several trajectories."""
    B = 700 if N == 64 else 300
    p = oracle.synth_params(oracle.SYS_GLV, N, 4321, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, stepper, adaptive, tol, tol, x0, p, 0.0, tf, dt0, objective=oracle.OBJ_SUM, threads=8)
    with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol) as e:
        r = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM)
        s = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
    assert (r["status"] == 0).all()
    same = (r["n_accept"] == o["n_accept"]) & (r["n_reject"] == o["n_reject"])
    assert same.mean() >= 0.99, f"{(~same).sum()} of {B} trajectories differ in accepted steps"
    assert_close(r["x_final"][same], o["x_final"][same], what="x(tf)")
    assert_close(r["lam"][same, 0], o["lam"][same], what="lambda")
    assert_close(r["mu"][same, 0], o["mu"][same], what="mu")
    assert_flips_bounded(r, o, same, tol)
    # summed-objective mode (per-CTA register accumulation + deterministic reduction) == sum of per-trajectory gradients
    assert_close(s["mu"], r["mu"][:, 0].sum(axis=0, keepdims=True), rtol=1e-11, what="mu sum")
    np.testing.assert_array_equal(s["lam"], r["lam"])


def test_glv_half_norm_objective_seeds_and_split_api(va):
    N, B = 64, 37
    p = oracle.synth_params(oracle.SYS_GLV, N, 77, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, x0, p, 0.0, 10.0, 1e-3, objective=oracle.OBJ_HALF_NORM2)
    rng = np.random.default_rng(0)
    seeds = rng.standard_normal((B, 2, N))
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
        r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_HALF_NORM2)
        assert_close(r["lam"][:, 0], o["lam"], what="lambda")
        assert_close(r["mu"][:, 0], o["mu"], what="mu")
        f = e.forward(x0, p, 0.0, 10.0, 1e-3)
        np.testing.assert_array_equal(f["x_final"], r["x_final"])
        t, x = e.checkpoints(B - 1)
        assert len(t) == f["n_accept"][B - 1] + 1 and t[0] == 0.0 and abs(t[-1] - 10.0) < 1e-14
        np.testing.assert_array_equal(x[0], x0[B - 1])
        np.testing.assert_array_equal(x[-1], f["x_final"][B - 1])
        a = e.adjoint(objective=va.OBJ_HALF_NORM2)  # B <= resident slots: sweeps back over the step blocks the forward call left
        np.testing.assert_array_equal(a["mu"], r["mu"])
        np.testing.assert_array_equal(a["lam"], r["lam"])
        a2 = e.adjoint(objective=va.OBJ_HALF_NORM2)  # and again (the slabs are still valid)
        np.testing.assert_array_equal(a2["mu"], r["mu"])
        # more parameter sets than resident slots: the slabs cannot hold them all, the adjoint call re-integrates
        pb = oracle.synth_params(oracle.SYS_GLV, N, 78, 0, 700)
        xb = oracle.synth_x0(oracle.SYS_GLV, N, pb)
        rb = e.forward_adjoint(xb, pb, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
        fb = e.forward(xb, pb, 0.0, 10.0, 1e-3)
        ab = e.adjoint(objective=va.OBJ_SUM)
        np.testing.assert_array_equal(fb["x_final"], rb["x_final"])
        np.testing.assert_array_equal(ab["mu"], rb["mu"])
        tb, xcb = e.checkpoints(699)  # not in its slot any more: re-integrated on demand
        assert len(tb) == fb["n_accept"][699] + 1
        np.testing.assert_array_equal(xcb[0], xb[699])
        np.testing.assert_array_equal(xcb[-1], fb["x_final"][699])
    # two explicit seeds per trajectory (the reference's SIMD axis): each must match a single-seed oracle sweep
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8, n_out=2) as e:
        r2 = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds)
        s2 = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds, reduce=va.REDUCE_SUM)
    for k in range(2):
        ok = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, x0, p, 0.0, 10.0, 1e-3, objective=oracle.OBJ_SEED,
                                    seeds=seeds[:, k])
        assert_close(r2["lam"][:, k], ok["lam"], what=f"lambda seed {k}")
        assert_close(r2["mu"][:, k], ok["mu"], what=f"mu seed {k}")
    assert_close(s2["mu"], r2["mu"].sum(axis=0), rtol=1e-11, what="mu sum, 2 seeds")


def test_glv_finite_difference_cross_check(va):
    """Central finite differences of J = sum x_i(tf) w.r.t. a few parameters (fixed-step RK4 so that J is smooth)."""
    N = 64
    p = oracle.synth_params(oracle.SYS_GLV, N, 11, 0, 1)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    ks = [0, 17, 63, 64, 64 + 65, 64 + 64 * 10 + 3, 64 + 64 * 63 + 63]
    h = 1e-6
    pp = np.repeat(p, 2 * len(ks), axis=0)
    for m, k in enumerate(ks):
        pp[2 * m, k] += h
        pp[2 * m + 1, k] -= h
    with va.Engine(va.SYS_GLV, N, va.RK_RK4, False, max_steps=256) as e:
        base = e.forward_adjoint(x0, p, 0.0, 1.0, 0.01, objective=va.OBJ_SUM)
        pert = e.forward_adjoint(np.repeat(x0, 2 * len(ks), axis=0), pp, 0.0, 1.0, 0.01, objective=va.OBJ_SUM)
    J = pert["x_final"].sum(axis=1)
    for m, k in enumerate(ks):
        fd = (J[2 * m] - J[2 * m + 1]) / (2 * h)
        assert abs(fd - base["mu"][0, 0, k]) <= 1e-7 * abs(fd) + 5e-9, (k, fd, base["mu"][0, 0, k])  # FD noise ~ eps/h


def test_glv64_flip_census_against_the_oracle(va):
    """4096 seeded parameter sets of the headline workload against the oracle, EVERY trajectory compared: same accept/reject
    sequence -> 1e-8 relative (observed ~1e-12); flipped -> 10 * tol. The flip count is printed (pytest -s) and bounded."""
    N, B, tol = 64, 4096, 1e-8
    p = oracle.synth_params(oracle.SYS_GLV, N, 1234, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, tol, tol, x0, p, 0.0, 10.0, 1e-3, objective=oracle.OBJ_SUM,
                               threads=os.cpu_count() or 1)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, tol, tol) as e:
        r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
    assert (r["status"] == 0).all()
    same = (r["n_accept"] == o["n_accept"]) & (r["n_reject"] == o["n_reject"])
    print(f"flip census N=64 tol=1e-8: {(~same).sum()} of {B} trajectories took a different accept/reject sequence")
    assert (~same).mean() <= 0.002
    assert_close(r["x_final"][same], o["x_final"][same], what="x(tf)")
    assert_close(r["lam"][same, 0], o["lam"][same], what="lambda")
    assert_close(r["mu"][same, 0], o["mu"][same], what="mu")
    assert_flips_bounded(r, o, same, tol)


@pytest.mark.parametrize("N", [16, 64, 100])
@pytest.mark.parametrize("stepper,adaptive,tol,tf,dt0,max_steps", [
    (0, False, 0.0, 1.0, 0.01, 128),      # euler, fixed step (reference ButcherTable.hpp:50-65)
    (4, True, 1e-8, 10.0, 1e-3, 0),       # fehlberg78 controlled, 13 stages (ButcherTable.hpp:191-246)
    (4, False, 0.0, 1.0, 0.02, 64),       # error steppers used un-controlled: the stepper_tag loop (detail/runge_kutta.hpp:38-72)
    (2, False, 0.0, 1.0, 0.02, 64),
    (3, False, 0.0, 1.0, 0.02, 64),
    (4, True, 1e-5, 10.0, 1e-3, 0)])
def test_glv_every_reference_tableau_on_every_size(va, N, stepper, adaptive, tol, tf, dt0, max_steps):
    """The reference runs any of its tableaux on any system; so does the GLV path (streamed-matrix family for the steppers the
    register kernels do not specialise)."""
    B = 40
    p = oracle.synth_params(oracle.SYS_GLV, N, 777, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, stepper, adaptive, tol, tol, x0, p, 0.0, tf, dt0, objective=oracle.OBJ_SUM, threads=8)
    for policy in (va.CKPT_STORE_STAGES, va.CKPT_RECOMPUTE):
        with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol, max_steps=max_steps, ckpt_policy=policy) as e:
            r = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM)
            s = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
            assert e.info()["kernel_name"] == "k_glv_stream"
        assert (r["status"] == 0).all()
        same = (r["n_accept"] == o["n_accept"]) & (r["n_reject"] == o["n_reject"])
        assert same.all() if not adaptive else same.mean() >= 0.95
        assert_close(r["x_final"][same], o["x_final"][same], what="x(tf)")
        assert_close(r["lam"][same, 0], o["lam"][same], what="lambda")
        assert_close(r["mu"][same, 0], o["mu"][same], what="mu")
        assert_flips_bounded(r, o, same, tol)
        assert_close(s["mu"], r["mu"][:, 0].sum(axis=0, keepdims=True), rtol=1e-11, what="mu sum")


def test_glv_full_size_properties(va):
    """BASELINE size class (N = 64, tol 1e-8) on device-generated inputs: properties that need no oracle at scale.
    Linearity of the adjoint in its seed, determinism, step-count statistics, and a sampled oracle comparison."""
    import torch
    N, B = 64, 32768
    npar = N * N + N
    dev = torch.device("cuda:0")
    p = torch.empty(B, npar, dtype=torch.float64, device=dev)
    x0 = torch.empty(B, N, dtype=torch.float64, device=dev)
    va.synth_batch_device(va.SYS_GLV, N, 1234, 0, B, p, x0)
    mk = lambda *s: torch.empty(*s, dtype=torch.float64, device=dev)
    xf, lam1, mu1 = mk(B, N), torch.ones(B, 1, N, dtype=torch.float64, device=dev), mk(B, 1, npar)
    na = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
        e.call("va_forward_adjoint_batch", B, x0, p, 0.0, 10.0, 1e-3, xf, lam1, mu1, va.OBJ_SEED, va.REDUCE_NONE, na, None, st)
        lam3, mu3 = torch.full((B, 1, N), 3.0, dtype=torch.float64, device=dev), mk(B, 1, npar)
        xf2 = mk(B, N)
        e.call("va_forward_adjoint_batch", B, x0, p, 0.0, 10.0, 1e-3, xf2, lam3, mu3, va.OBJ_SEED, va.REDUCE_NONE)
        musum = mk(1, npar)
        lam_s = torch.ones(B, 1, N, dtype=torch.float64, device=dev)
        e.call("va_forward_adjoint_batch", B, x0, p, 0.0, 10.0, 1e-3, xf2, lam_s, musum, va.OBJ_SEED, va.REDUCE_SUM)
        torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0
    assert torch.equal(xf, xf2)  # deterministic
    assert torch.isfinite(mu1).all()
    # the reverse sweep is linear in the seed
    assert float((mu3 - 3.0 * mu1).abs().max() / mu1.abs().max()) < 1e-13
    assert float((lam3 - 3.0 * lam1).abs().max() / lam1.abs().max()) < 1e-13
    # summed objective == sum of per-trajectory gradients
    ref_sum = mu1[:, 0].sum(dim=0)
    assert float((musum[0] - ref_sum).abs().max() / ref_sum.abs().max()) < 1e-11
    steps = na.cpu().numpy()
    assert 15 <= steps.min() and steps.max() <= 40, (steps.min(), steps.max())
    # sampled comparison with the oracle on the very same parameter sets (generator is bit-identical host/device)
    idx = np.array([0, 1, 4095, 12345, B - 1])
    ph = p[idx].cpu().numpy()
    np.testing.assert_array_equal(ph, np.concatenate([oracle.synth_params(oracle.SYS_GLV, N, 1234, int(i), 1) for i in idx]))
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, x0[idx].cpu().numpy(), ph, 0.0, 10.0, 1e-3,
                               objective=oracle.OBJ_SUM)
    np.testing.assert_array_equal(steps[idx], o["n_accept"])
    assert_close(xf[idx].cpu().numpy(), o["x_final"], what="x(tf)")
    assert_close(mu1[idx, 0].cpu().numpy(), o["mu"], what="mu")
    assert_close(lam1[idx, 0].cpu().numpy(), o["lam"], what="lambda")


@pytest.mark.parametrize("N,B,stepper,adaptive,tol,tf,dt0", [(100, 6, 2, True, 1e-8, 10.0, 1e-3), (256, 3, 2, True, 1e-8, 10.0, 1e-3),
                                                            (256, 2, 3, True, 1e-6, 10.0, 1e-3), (80, 4, 1, False, 0.0, 0.3, 0.01),
                                                            (65, 3, 3, True, 1e-6, 10.0, 1e-3), (129, 3, 2, True, 1e-8, 10.0, 1e-3),
                                                            (255, 2, 2, True, 1e-8, 10.0, 1e-3), (200, 3, 1, False, 0.0, 0.3, 0.01),
                                                            (300, 2, 2, True, 1e-6, 10.0, 1e-3)])
def test_glv_large_species_counts_streamed_matrix(va, N, B, stepper, adaptive, tol, tf, dt0):
    """N > 64 (BASELINE config 5 uses N = 256): the matrix no longer fits one SM's registers. 65..256 species run on the
    cluster kernel (two SMs hold the matrix; fewer than 256 species are padded with inert ones, odd N included), larger
    systems on the streamed-matrix kernel."""
    p = oracle.synth_params(oracle.SYS_GLV, N, 2024, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, stepper, adaptive, tol, tol, x0, p, 0.0, tf, dt0, objective=oracle.OBJ_SUM, threads=8)
    with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol) as e:
        assert e.info()["kernel_family"] == 2 and e.info()["kernel_name"] == ("k_glv_pair" if N <= 256 else "k_glv_stream")
        r = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM)
        s = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
    assert (r["status"] == 0).all()
    np.testing.assert_array_equal(r["n_accept"], o["n_accept"])
    assert_close(r["x_final"], o["x_final"], what="x(tf)")
    assert_close(r["lam"][:, 0], o["lam"], what="lambda")
    assert_close(r["mu"][:, 0], o["mu"], what="mu")
    assert_close(s["mu"], r["mu"][:, 0].sum(axis=0, keepdims=True), rtol=1e-11, what="mu sum")


@pytest.mark.parametrize("N,B,stepper,adaptive,tol,tf,dt0", [(100, 5, 2, True, 1e-8, 10.0, 1e-3), (256, 2, 3, True, 1e-6, 10.0, 1e-3),
                                                            (64, 6, 1, False, 0.0, 0.3, 0.01)])
def test_glv_checkpoint_policy_recompute_matches_store_stages(va, N, B, stepper, adaptive, tol, tf, dt0):
    """north_star item 4: the reverse sweep either reads stored stages or recomputes them from the stored (t_n, x_n)
    (the reference's policy, detail/backpropagation.hpp:24-64). Both must give the reference's gradients; the recompute
    policy keeps 8(N+8) B per step instead of 8(12N+8) B."""
    p = oracle.synth_params(oracle.SYS_GLV, N, 77, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, stepper, adaptive, tol, tol, x0, p, 0.0, tf, dt0, objective=oracle.OBJ_SUM, threads=8)
    res = {}
    for pol in (va.CKPT_STORE_STAGES, va.CKPT_RECOMPUTE):
        with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol, ckpt_policy=pol) as e:
            info = e.info()
            assert info["ckpt_policy"] == pol
            if pol == va.CKPT_RECOMPUTE:
                assert info["kernel_family"] == 2 and info["kernel_name"] == ("k_glv_pair" if 64 < N <= 256 else "k_glv_stream")
            res[pol] = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM)
            e.forward(x0, p, 0.0, tf, dt0)
            t, x = e.checkpoints(1)
        r = res[pol]
        assert (r["status"] == 0).all()
        np.testing.assert_array_equal(r["n_accept"], o["n_accept"])
        assert_close(r["x_final"], o["x_final"], what="x(tf)")
        assert_close(r["lam"][:, 0], o["lam"], what="lambda")
        assert_close(r["mu"][:, 0], o["mu"], what="mu")
        assert len(t) == r["n_accept"][1] + 1 and t[0] == 0.0
        np.testing.assert_array_equal(x[0], x0[1])
    assert_close(res[va.CKPT_RECOMPUTE]["mu"], res[va.CKPT_STORE_STAGES]["mu"], rtol=1e-11, what="mu, recompute vs store")


@pytest.mark.parametrize("seg", ["1", "4", "16", "64"])
def test_glv256_recompute_policy_segments_on_the_cluster_kernel(va, monkeypatch, seg):
    """Recompute policy on the cluster kernel: the forward sweep keeps (t_n, x_n) only, the reverse sweep re-integrates segments
    of VA_PAIR_SEG steps (newest first) and sweeps back through each. Segment lengths that divide the step count, do not divide
    it, and exceed it; two seeds, summed mode, both objectives; against the store-stages policy, the oracle and the checkpoints."""
    N, B = 256, 4
    p = oracle.synth_params(oracle.SYS_GLV, N, 8080, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    seeds = np.random.default_rng(11).standard_normal((B, 2, N))
    monkeypatch.setenv("VA_PAIR_SEG", seg)
    res = {}
    for pol in (va.CKPT_RECOMPUTE, va.CKPT_STORE_STAGES):
        with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8, n_out=2, ckpt_policy=pol) as e:
            info = e.info()
            assert info["kernel_name"] == "k_glv_pair" and info["ckpt_policy"] == pol
            r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds)
            s = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds, reduce=va.REDUCE_SUM)
            z = e.forward_adjoint(x0, p, 2.0, 2.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds)
            e.forward(x0, p, 0.0, 10.0, 1e-3)
            t, x = e.checkpoints(3)
        with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8, ckpt_policy=pol) as e:
            h = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_HALF_NORM2)
            hs = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_HALF_NORM2, reduce=va.REDUCE_SUM)
        assert (r["status"] == 0).all() and (z["n_accept"] == 0).all()
        np.testing.assert_array_equal(z["lam"], seeds)
        assert (z["mu"] == 0).all()
        assert_close(s["mu"], r["mu"].sum(axis=0), rtol=1e-11, what="mu sum")
        assert_close(hs["mu"], h["mu"].sum(axis=0), rtol=1e-11, what="mu sum (native summed mode)")
        assert len(t) == r["n_accept"][3] + 1 and t[0] == 0.0 and abs(t[-1] - 10.0) < 1e-12
        np.testing.assert_array_equal(x[0], x0[3])
        res[pol] = (r, h, t, x)
    (a, ha, ta, xa), (b, hb, tb, xb) = res[va.CKPT_RECOMPUTE], res[va.CKPT_STORE_STAGES]
    np.testing.assert_array_equal(a["n_accept"], b["n_accept"])
    np.testing.assert_array_equal(a["x_final"], b["x_final"])  # the forward sweeps are the same code
    np.testing.assert_array_equal(ta, tb)
    np.testing.assert_array_equal(xa, xb)
    assert_close(a["lam"].reshape(B * 2, -1), b["lam"].reshape(B * 2, -1), rtol=1e-10, what="lambda")
    assert_close(a["mu"].reshape(B * 2, -1), b["mu"].reshape(B * 2, -1), rtol=1e-10, what="mu")
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, x0, p, 0.0, 10.0, 1e-3,
                               objective=oracle.OBJ_HALF_NORM2, threads=8)
    np.testing.assert_array_equal(ha["n_accept"], o["n_accept"])
    assert_close(ha["lam"][:, 0], o["lam"], what="lambda")
    assert_close(ha["mu"][:, 0], o["mu"], what="mu")


def test_glv_auto_policy_long_horizon_uses_the_cluster_kernel(va):
    """VA_CKPT_AUTO with a step capacity whose stage blocks would not fit the workspace share of HBM: the engine chooses the
    recompute policy and stays on the cluster kernel (state store + segment re-integration); results as with stored stages."""
    N, B = 200, 3
    p = oracle.synth_params(oracle.SYS_GLV, N, 12, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-12, 1e-12, max_steps=1000, workspace_fraction=0.01) as e:
        info = e.info()  # 148 x 1001 x 36.9 KB = 5.5 GB of stage blocks > 1 % of HBM
        assert info["kernel_name"] == "k_glv_pair" and info["ckpt_policy"] == va.CKPT_RECOMPUTE
        r = e.forward_adjoint(x0, p, 0.0, 1000.0, 1e-3, objective=va.OBJ_SUM)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-12, 1e-12, max_steps=1000, ckpt_policy=va.CKPT_STORE_STAGES) as e:
        q = e.forward_adjoint(x0, p, 0.0, 1000.0, 1e-3, objective=va.OBJ_SUM)
    assert (r["status"] == 0).all() and r["n_accept"].min() > 100
    np.testing.assert_array_equal(r["n_accept"], q["n_accept"])
    assert_close(r["mu"][:, 0], q["mu"][:, 0], rtol=1e-9, what="mu")
    assert_close(r["lam"][:, 0], q["lam"][:, 0], rtol=1e-9, what="lambda")


@pytest.mark.parametrize("seg", ["1", "5", "16", "64"])
@pytest.mark.parametrize("N,tol,tf", [(256, 1e-8, 10.0), (100, 1e-12, 300.0)])
def test_glv_sparse_checkpoints_on_the_cluster_kernel(va, monkeypatch, seg, N, tol, tf):
    """VA_CKPT_SPARSE ('checkpoint scheduling beyond store-all'; the reference stores every accepted state,
    lib/include/StateStorage.hpp:7-8): t_n of every accepted step, x_n of every L-th only; each segment is re-integrated from its
    first state. Against the store-stages policy and the oracle, for segment lengths 1 (= dense), 5, 16 and longer than the
    trajectory; a short horizon (~21 steps) and a long one (~150 steps); two seeds and the summed mode."""
    B = 3
    p = oracle.synth_params(oracle.SYS_GLV, N, 515, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    seeds = np.random.default_rng(5).standard_normal((B, 2, N))
    monkeypatch.setenv("VA_PAIR_SEG", seg)
    res = {}
    for pol in (va.CKPT_SPARSE, va.CKPT_STORE_STAGES):
        with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, tol, tol, n_out=2, max_steps=512, ckpt_policy=pol) as e:
            info = e.info()
            assert info["kernel_name"] == "k_glv_pair" and info["ckpt_policy"] == pol
            r = e.forward_adjoint(x0, p, 0.0, tf, 1e-3, objective=va.OBJ_SEED, seeds=seeds)
            s = e.forward_adjoint(x0, p, 0.0, tf, 1e-3, objective=va.OBJ_SEED, seeds=seeds, reduce=va.REDUCE_SUM)
            f = e.forward(x0, p, 0.0, tf, 1e-3)
            if pol == va.CKPT_SPARSE:
                with pytest.raises(va.EngineError, match="every L-th"):
                    e.checkpoints(1)  # the intermediate states are not there: the call says so instead of returning something else
            ws = info["workspace_bytes"] or e.info()["workspace_bytes"]
        assert (r["status"] == 0).all()
        assert_close(s["mu"], r["mu"].sum(axis=0), rtol=1e-11, what="mu sum")
        res[pol] = (r, f, ws)
    (a, fa, wa), (b, fb, wb) = res[va.CKPT_SPARSE], res[va.CKPT_STORE_STAGES]
    np.testing.assert_array_equal(a["n_accept"], b["n_accept"])
    np.testing.assert_array_equal(a["x_final"], b["x_final"])
    assert_close(a["lam"].reshape(B * 2, -1), b["lam"].reshape(B * 2, -1), rtol=1e-9, what="lambda")
    assert_close(a["mu"].reshape(B * 2, -1), b["mu"].reshape(B * 2, -1), rtol=1e-9, what="mu")
    assert wa < wb / 4  # what it is for: checkpoint memory (segment slab + times + every L-th state against 513 stage blocks per CTA)
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, tol, tol, x0, p, 0.0, tf, 1e-3, objective=oracle.OBJ_SEED,
                               seeds=seeds[:, 0], threads=8)
    np.testing.assert_array_equal(a["n_accept"], o["n_accept"])
    assert_close(a["lam"][:, 0], o["lam"], what="lambda vs oracle")
    assert_close(a["mu"][:, 0], o["mu"], what="mu vs oracle")


def test_glv_auto_policy_picks_sparse_checkpoints_for_very_long_horizons(va):
    """VA_CKPT_AUTO: stage blocks too large -> recompute; one state per step still too large -> sparse."""
    with va.Engine(va.SYS_GLV, 256, va.RK_CK54, True, 1e-8, 1e-8, max_steps=2_000_000) as e:
        info = e.info()  # 148 x 2e6 x 2.1 KB = 625 GB of states: more than the GPU has
        assert info["kernel_name"] == "k_glv_pair" and info["ckpt_policy"] == va.CKPT_SPARSE
        p = oracle.synth_params(oracle.SYS_GLV, 256, 3, 0, 2)
        x0 = oracle.synth_x0(oracle.SYS_GLV, 256, p)
        r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
        assert e.info()["workspace_bytes"] < 60e9
    o = oracle.forward_adjoint(oracle.SYS_GLV, 256, oracle.RK_CK54, True, 1e-8, 1e-8, x0, p, 0.0, 10.0, 1e-3, objective=oracle.OBJ_SUM, threads=2)
    np.testing.assert_array_equal(r["n_accept"], o["n_accept"])
    assert_close(r["mu"][:, 0], o["mu"], what="mu")


@pytest.mark.parametrize("N,B", [(64, 5), (50, 3), (64, 1)])
def test_glv_register_kernel_generations_agree(va, monkeypatch, N, B):
    """33..64 species run on va_glv_t8.cu (64 threads per trajectory, 8x8 tiles, three phases); VA_GLV_V1 selects the
    first-generation kernel (va_glv_wide.cu). Independent thread/data maps, same algorithm: cross-check them, with fewer
    trajectories than slots in a CTA (idle slots), a non-padded and a padded species count, two seeds and the summed mode."""
    p = oracle.synth_params(oracle.SYS_GLV, N, 909, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    seeds = np.random.default_rng(1).standard_normal((B, 2, N))
    res = []
    for v1 in (False, True):
        if v1:
            monkeypatch.setenv("VA_GLV_V1", "1")
        with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8, n_out=2) as e:
            info = e.info()
            assert info["kernel_family"] == 1 and info["threads_per_cta"] == (128 if v1 else 256)
            r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds)
            s = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds, reduce=va.REDUCE_SUM)
            z = e.forward_adjoint(x0, p, 2.0, 2.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds)  # ti == tf: no step at all
        assert (r["status"] == 0).all() and (z["n_accept"] == 0).all()
        np.testing.assert_array_equal(z["x_final"], x0)
        np.testing.assert_array_equal(z["lam"], seeds)
        assert (z["mu"] == 0).all()
        assert_close(s["mu"], r["mu"].sum(axis=0), rtol=1e-11, what="mu sum")
        res.append(r)
    a, b = res
    np.testing.assert_array_equal(a["n_accept"], b["n_accept"])
    assert_close(a["x_final"], b["x_final"], rtol=1e-12, what="x(tf)")
    assert_close(a["lam"].reshape(B * 2, -1), b["lam"].reshape(B * 2, -1), rtol=1e-11, what="lambda")
    assert_close(a["mu"].reshape(B * 2, -1), b["mu"].reshape(B * 2, -1), rtol=1e-11, what="mu")


@pytest.mark.parametrize("N,B,stepper,adaptive,tol,tf,dt0,n_out,max_steps", [(64, 1900, 2, True, 1e-8, 10.0, 1e-3, 1, 0), (50, 700, 3, True, 1e-6, 10.0, 1e-3, 2, 0),
                                                                            (40, 900, 1, False, 0.0, 0.5, 0.01, 1, 0), (64, 3, 2, True, 1e-8, 10.0, 1e-3, 1, 0),
                                                                            (64, 1500, 2, True, 1e-8, 10.0, 1e-3, 1, 20)])
def test_glv_warp_specialised_kernel_returns_the_same_bits(va, monkeypatch, N, B, stepper, adaptive, tol, tf, dt0, n_out, max_steps):
    """va_glv_t8s.cu (VA_GLV_T8S=1; kept as a measured negative result) runs the sweeps and the gradient
    accumulation on different warps of the CTA (setmaxnreg register split, job queue, two slab halves per slot) but forms every sum
    in the order of va_glv_t8.cu: per-trajectory results, the summed gradient, the split API and the checkpoints must be
    bit-identical -- with several trajectories per slot (both slab halves, the queue under load), padded species counts, two seeds
    (synchronous hand-over), fixed-step rk4, dopri5, fewer trajectories than slots, and a checkpoint capacity that about half of the
    trajectories overflow (failed trajectories post no accumulation job; their outputs are NaN in both kernels)."""
    p = oracle.synth_params(oracle.SYS_GLV, N, 4242, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    seeds = np.random.default_rng(5).standard_normal((B, n_out, N))
    res = []
    for spec in (False, True):
        if spec:
            monkeypatch.setenv("VA_GLV_T8S", "1")
        with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol, n_out=n_out, max_steps=max_steps) as e:
            assert e.info()["kernel_name"] == ("k_glv_t8s" if spec else "k_glv_t8")
            r = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SEED, seeds=seeds)
            s = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
            ck = a = None
            if not max_steps:
                e.forward(x0, p, 0.0, tf, dt0)
                ck = e.checkpoints(min(B - 1, 2))
                a = e.adjoint(objective=va.OBJ_SEED, seeds=seeds)
        if max_steps:
            assert 0 < (r["status"] != 0).sum() < B
        else:
            assert (r["status"] == 0).all()
        res.append((r, s, ck, a))
    (r0, s0, ck0, a0), (r1, s1, ck1, a1) = res
    for k in ("x_final", "lam", "mu", "n_accept", "n_reject", "status"):
        np.testing.assert_array_equal(r0[k], r1[k], err_msg=k)
    np.testing.assert_array_equal(s0["mu"], s1["mu"])
    if max_steps:
        return
    np.testing.assert_array_equal(ck0[0], ck1[0])
    np.testing.assert_array_equal(ck0[1], ck1[1])
    np.testing.assert_array_equal(a0["mu"], a1["mu"])
    np.testing.assert_array_equal(a1["mu"], r1["mu"])  # split API == fused call
    # and against the oracle on a few sets (the second-generation kernel is checked against it everywhere else)
    nb = min(B, 3)
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, stepper, adaptive, tol, tol, x0[:nb], p[:nb], 0.0, tf, dt0, objective=oracle.OBJ_SEED,
                               seeds=seeds[:nb, 0], threads=2)
    np.testing.assert_array_equal(r1["n_accept"][:nb], o["n_accept"])
    assert_close(r1["mu"][:nb, 0], o["mu"], what="mu vs oracle")


@pytest.mark.parametrize("N", [64, 50])
def test_glv_parameter_staging_returns_the_same_bits_and_misaligned_device_arrays_are_refused(va, monkeypatch, N):
    """With one seed per trajectory va_glv_t8.cu brings each parameter set into shared memory with one TMA bulk copy and cuts both
    register tiles from there; VA_T8_NO_STAGE=1 keeps the per-lane global loads. Same arithmetic: identical bits (several
    trajectories per slot, padded species count). Both forms -- like every kernel here -- use 16-byte accesses on the caller's
    device arrays, so the fused call refuses a device array that starts on an odd multiple of 8 bytes with VA_E_INVALID instead
    of faulting inside the kernel (found by this test: it used to be a sticky 'misaligned address' error)."""
    import torch
    B = 2100
    npar = N * N + N
    dev = torch.device("cuda:0")
    raw = torch.empty(B * npar + 1, dtype=torch.float64, device=dev)
    params = raw[:-1].view(B, npar)
    x0 = torch.empty(B, N, dtype=torch.float64, device=dev)
    va.synth_batch_device(va.SYS_GLV, N, 77, 0, B, params, x0)
    out = []
    for nostage in (False, True):
        if nostage:
            monkeypatch.setenv("VA_T8_NO_STAGE", "1")
        xf, lam, mu = (torch.empty(B, N, dtype=torch.float64, device=dev), torch.ones(B, 1, N, dtype=torch.float64, device=dev),
                       torch.empty(B, 1, npar, dtype=torch.float64, device=dev))
        musum = torch.empty(1, npar, dtype=torch.float64, device=dev)
        lam_s = torch.ones(B, 1, N, dtype=torch.float64, device=dev)
        na = torch.empty(B, dtype=torch.int32, device=dev)
        with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
            assert e.info()["kernel_name"] == "k_glv_t8"
            e.call("va_forward_adjoint_batch", B, x0, params, 0.0, 10.0, 1e-3, xf, lam, mu, va.OBJ_SEED, va.REDUCE_NONE, na)
            e.call("va_forward_adjoint_batch", B, x0, params, 0.0, 10.0, 1e-3, xf, lam_s, musum, va.OBJ_SEED, va.REDUCE_SUM)
            torch.cuda.synchronize()
            if not nostage:
                keep = params.clone()
                odd = raw[1:].view(B, npar)  # 8 bytes further: not 16-byte aligned
                odd.copy_(keep)
                assert odd.data_ptr() % 16 == 8
                with pytest.raises(va.EngineError, match="16-byte aligned"):
                    e.call("va_forward_adjoint_batch", B, x0, odd, 0.0, 10.0, 1e-3, xf.clone(), lam.clone(), mu.clone(), va.OBJ_SEED, va.REDUCE_NONE)
                params = keep
        out.append((xf, lam, mu, musum, na))
    for a, b in zip(*out):
        assert torch.equal(a, b)
    assert int(out[0][4].min()) > 5 and torch.isfinite(out[0][2]).all()


def test_glv64_dead_step_blocks_dropped_from_l2_change_nothing(va, monkeypatch):
    """VA_T8_DISCARD=1: with more trajectories than slots the headline kernel drops the L2 lines of a trajectory's step blocks once its
    gradient accumulation has read them (discard.global.L2; -30 % DRAM traffic, -1.1 % rate, hence opt-in). The blocks are dead by
    then: every result must be bit-identical, in both gradient modes and with two seeds (blocks reused by the second seed)."""
    N, B = 64, 2500
    p = oracle.synth_params(oracle.SYS_GLV, N, 99, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    seeds = np.random.default_rng(6).standard_normal((B, 2, N))
    res = []
    for drop in (False, True):
        if drop:
            monkeypatch.setenv("VA_T8_DISCARD", "1")
        out = []
        for n_out in (1, 2):
            with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8, n_out=n_out) as e:
                assert e.info()["kernel_name"] == "k_glv_t8"
                out.append(e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds[:, :n_out]))
                out.append(e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM))
        res.append(out)
    for a, b in zip(*res):
        for k in ("x_final", "lam", "mu", "n_accept"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.parametrize("N,B,stepper,adaptive,tol,tf,dt0", [(16, 300, 2, True, 1e-8, 10.0, 1e-3), (16, 40, 3, True, 1e-6, 10.0, 1e-3),
                                                            (10, 37, 2, True, 1e-8, 10.0, 1e-3), (13, 9, 3, True, 1e-7, 10.0, 1e-3),
                                                            (5, 33, 2, True, 1e-5, 10.0, 1e-3), (16, 21, 1, False, 0.0, 0.3, 0.01),
                                                            (3, 5, 2, True, 1e-8, 3.0, 1e-3)])
def test_glv_quad_kernel_agrees_with_first_generation_and_oracle(va, monkeypatch, N, B, stepper, adaptive, tol, tf, dt0):
    """Up to 16 species run on va_glv_oct.cu by default (eight lanes per trajectory, A and Abar resident, recompute policy);
    the store-stages policy selects va_glv_quad.cu (four lanes per trajectory, eight trajectories per warp in lock step, three
    phases), VA_GLV_NO_QUAD the first-generation kernel (va_glv_wide.cu, a warp per trajectory). Three independent thread/data
    maps (and two checkpoint policies) of the same algorithm: cross-check them and the oracle, with trajectories of different
    step counts sharing a warp, idle lanes, padded species counts, two seeds, the summed mode and ti == tf."""
    p = oracle.synth_params(oracle.SYS_GLV, N, 606, 0, B)
    p[::3, :N] *= 3.0  # spread the growth rates so that neighbouring trajectories take different numbers of steps
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    seeds = np.random.default_rng(3).standard_normal((B, 2, N))
    res = []
    for kernel in ("k_glv_oct", "k_glv_quad", "k_glv_wide"):
        policy = va.CKPT_AUTO if kernel == "k_glv_oct" else va.CKPT_STORE_STAGES
        if kernel == "k_glv_wide":
            monkeypatch.setenv("VA_GLV_NO_QUAD", "1")
        with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol, n_out=2, ckpt_policy=policy) as e:
            assert e.info()["kernel_name"] == kernel
            r = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SEED, seeds=seeds)
            s = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_SEED, seeds=seeds, reduce=va.REDUCE_SUM)
            z = e.forward_adjoint(x0, p, 2.0, 2.0, dt0, objective=va.OBJ_SEED, seeds=seeds)  # ti == tf: no step at all
            e.forward(x0[:50], p[:50], 0.0, tf, dt0)
            t, x = e.checkpoints(min(B, 50) - 1)
            sa = e.adjoint(objective=va.OBJ_SEED, seeds=seeds[:50])  # split API (runge_kutta, then adjointSolve)
            np.testing.assert_array_equal(sa["mu"], r["mu"][:50])
            np.testing.assert_array_equal(sa["lam"], r["lam"][:50])
        with va.Engine(va.SYS_GLV, N, stepper, adaptive, tol, tol, ckpt_policy=policy) as e:
            h = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_HALF_NORM2)
            hs = e.forward_adjoint(x0, p, 0.0, tf, dt0, objective=va.OBJ_HALF_NORM2, reduce=va.REDUCE_SUM)
        assert (r["status"] == 0).all() and (z["n_accept"] == 0).all()
        np.testing.assert_array_equal(z["x_final"], x0)
        np.testing.assert_array_equal(z["lam"], seeds)
        assert (z["mu"] == 0).all()
        assert_close(s["mu"], r["mu"].sum(axis=0), rtol=1e-11, what="mu sum")
        assert_close(hs["mu"], h["mu"].sum(axis=0), rtol=1e-11, what="mu sum (native summed mode)")
        assert len(t) == r["n_accept"][min(B, 50) - 1] + 1 and t[0] == 0.0
        np.testing.assert_array_equal(x[0], x0[min(B, 50) - 1])
        res.append((r, h))
    (c, hc), (a, ha), (b, hb) = res
    for other in (a, c):
        np.testing.assert_array_equal(other["n_accept"], b["n_accept"])
        assert_close(other["x_final"], b["x_final"], rtol=1e-12, what="x(tf)")
        assert_close(other["lam"].reshape(B * 2, -1), b["lam"].reshape(B * 2, -1), rtol=1e-10, what="lambda")
        assert_close(other["mu"].reshape(B * 2, -1), b["mu"].reshape(B * 2, -1), rtol=1e-10, what="mu")
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, stepper, adaptive, tol, tol, x0, p, 0.0, tf, dt0, objective=oracle.OBJ_HALF_NORM2, threads=8)
    np.testing.assert_array_equal(ha["n_accept"], o["n_accept"])
    assert_close(ha["x_final"], o["x_final"], what="x(tf)")
    assert_close(ha["lam"][:, 0], o["lam"], what="lambda")
    assert_close(ha["mu"][:, 0], o["mu"], what="mu")
    assert_close(hb["mu"][:, 0], o["mu"], what="mu (first generation)")
    np.testing.assert_array_equal(hc["n_accept"], o["n_accept"])
    assert_close(hc["x_final"], o["x_final"], what="x(tf) (eight-lane kernel)")
    assert_close(hc["lam"][:, 0], o["lam"], what="lambda (eight-lane kernel)")
    assert_close(hc["mu"][:, 0], o["mu"], what="mu (eight-lane kernel)")


def test_glv_quad_large_step_capacity_shrinks_the_grid(va):
    """64 checkpoint slabs per CTA: a large max_steps must not exhaust HBM; the engine runs fewer CTAs instead, results unchanged."""
    N, B = 16, 200
    p = oracle.synth_params(oracle.SYS_GLV, N, 17, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    out = []
    for cap, frac in ((0, 0.0), (2000, 0.05)):  # 9472 slabs x 2001 blocks x 1600 B = 30 GB > 5 % of HBM -> fewer CTAs
        with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8, max_steps=cap, workspace_fraction=frac, ckpt_policy=va.CKPT_STORE_STAGES) as e:
            info = e.info()
            assert info["kernel_name"] == "k_glv_quad"
            r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
            assert e.info()["workspace_bytes"] < 10e9
        out.append(r)
    np.testing.assert_array_equal(out[0]["mu"], out[1]["mu"])
    np.testing.assert_array_equal(out[0]["x_final"], out[1]["x_final"])


@pytest.mark.parametrize("kernel", ["k_glv_oct", "k_glv_quad"])
def test_glv16_several_trajectories_per_quad(va, kernel):
    """More parameter sets than resident slots (148 x 32 / 148 x 64 on a B200): every slot integrates several trajectories and, in
    summed mode, keeps adding to its partial-sum row; replicated parameter sets must give identical rows wherever they run."""
    N, B = 16, 30000
    base = oracle.synth_params(oracle.SYS_GLV, N, 99, 0, 7)
    p = base[np.arange(B) % 7]
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8,
                   ckpt_policy=va.CKPT_AUTO if kernel == "k_glv_oct" else va.CKPT_STORE_STAGES) as e:
        assert e.info()["kernel_name"] == kernel
        r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
        s = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
    assert (r["status"] == 0).all()
    for k in range(7):
        np.testing.assert_array_equal(r["mu"][k::7], np.broadcast_to(r["mu"][k], r["mu"][k::7].shape))
        np.testing.assert_array_equal(r["lam"][k::7], np.broadcast_to(r["lam"][k], r["lam"][k::7].shape))
        np.testing.assert_array_equal(r["x_final"][k::7], np.broadcast_to(r["x_final"][k], r["x_final"][k::7].shape))
    assert_close(s["mu"], r["mu"][:, 0].sum(axis=0, keepdims=True), rtol=1e-11, what="mu sum")
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, x0[:7], p[:7], 0.0, 10.0, 1e-3, objective=oracle.OBJ_SUM)
    np.testing.assert_array_equal(r["n_accept"][:7], o["n_accept"])
    assert_close(r["mu"][:7, 0], o["mu"], what="mu")


def test_glv_streamed_family_agrees_with_register_family(va, monkeypatch):
    """The two GLV kernel families are independent implementations: cross-check them at N = 64."""
    N, B = 64, 40
    p = oracle.synth_params(oracle.SYS_GLV, N, 5150, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
        assert e.info()["kernel_family"] == 1
        w = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_HALF_NORM2)
    monkeypatch.setenv("VA_GLV_FORCE_STREAM", "1")
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
        assert e.info()["kernel_family"] == 2
        s = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_HALF_NORM2)
        f = e.forward(x0, p, 0.0, 10.0, 1e-3)
        t, x = e.checkpoints(3)
    np.testing.assert_array_equal(w["n_accept"], s["n_accept"])
    assert_close(w["x_final"], s["x_final"], rtol=1e-12, what="x(tf)")
    assert_close(w["lam"][:, 0], s["lam"][:, 0], rtol=1e-11, what="lambda")
    assert_close(w["mu"][:, 0], s["mu"][:, 0], rtol=1e-11, what="mu")
    assert len(t) == f["n_accept"][3] + 1 and t[0] == 0.0
    np.testing.assert_array_equal(x[0], x0[3])


@pytest.mark.parametrize("kernel", ["k_glv_pair:2", "k_glv_pair:4", "k_glv_ring:2", "k_glv_ring:0", "k_glv_ring:6"])
def test_glv256_onchip_and_ring_kernels_agree_with_streamed_kernel_and_oracle(va, monkeypatch, kernel):
    """256 species, store-stages policy. Three independent thread/data maps of the same algorithm:
    k_glv_pair (va_glv_pair.cu, default: two CTAs of a cluster hold the matrix on chip and exchange product halves through
    distributed shared memory), k_glv_ring (va_glv_ring.cu, VA_GLV_NO_PAIR: TMA ring of matrix chunks, 64 rows cached in
    registers; VA_RING_FLAGS bit 1 = evict_last matrix stream, bit 2 = no register-cached rows) and k_glv_stream
    (va_glv_stream.cu, VA_GLV_NO_RING). A few sets against the oracle, two seeds per trajectory, summed mode, the
    J = |x|^2/2 objective, ti == tf, checkpoints, and the kernels against each other."""
    N, B = 256, 5
    p = oracle.synth_params(oracle.SYS_GLV, N, 4242, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    seeds = np.random.default_rng(7).standard_normal((B, 2, N))
    name, _, flags = kernel.partition(":")
    if name == "k_glv_ring":
        monkeypatch.setenv("VA_GLV_NO_PAIR", "1")
        monkeypatch.setenv("VA_RING_FLAGS", flags)
    else:
        monkeypatch.setenv("VA_GLV_CLUSTER", flags)  # CTAs per trajectory: 2 (registers + shared memory) or 4 (registers only)
    res = []
    for fast in (True, False):
        if not fast:
            monkeypatch.setenv("VA_GLV_NO_RING", "1")
        with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8, n_out=2) as e:
            info = e.info()
            assert info["kernel_family"] == 2 and info["kernel_name"] == (name if fast else "k_glv_stream")
            r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds)
            s = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds, reduce=va.REDUCE_SUM)
            z = e.forward_adjoint(x0, p, 2.0, 2.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds)  # ti == tf: no step at all
            e.forward(x0, p, 0.0, 10.0, 1e-3)
            t, x = e.checkpoints(2)
        with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
            h = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_HALF_NORM2)
            hs = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_HALF_NORM2, reduce=va.REDUCE_SUM)
        assert (r["status"] == 0).all() and (z["n_accept"] == 0).all()
        np.testing.assert_array_equal(z["x_final"], x0)
        np.testing.assert_array_equal(z["lam"], seeds)
        assert (z["mu"] == 0).all()
        assert_close(s["mu"], r["mu"].sum(axis=0), rtol=1e-11, what="mu sum")
        assert_close(hs["mu"], h["mu"].sum(axis=0), rtol=1e-11, what="mu sum (native summed mode)")
        assert len(t) == r["n_accept"][2] + 1 and t[0] == 0.0
        np.testing.assert_array_equal(x[0], x0[2])
        res.append((r, h))
    (a, ha), (b, hb) = res
    np.testing.assert_array_equal(a["n_accept"], b["n_accept"])
    assert_close(a["x_final"], b["x_final"], rtol=1e-12, what="x(tf)")
    assert_close(a["lam"].reshape(B * 2, -1), b["lam"].reshape(B * 2, -1), rtol=1e-10, what="lambda")
    assert_close(a["mu"].reshape(B * 2, -1), b["mu"].reshape(B * 2, -1), rtol=1e-10, what="mu")
    assert_close(ha["mu"][:, 0], hb["mu"][:, 0], rtol=1e-10, what="mu, half-norm objective")
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, x0, p, 0.0, 10.0, 1e-3,
                               objective=oracle.OBJ_HALF_NORM2, threads=8)
    np.testing.assert_array_equal(ha["n_accept"], o["n_accept"])
    assert_close(ha["x_final"], o["x_final"], what="x(tf)")
    assert_close(ha["lam"][:, 0], o["lam"], what="lambda")
    assert_close(ha["mu"][:, 0], o["mu"], what="mu")


def test_glv256_against_the_reference_fixture(va, synth_goldens):
    """BASELINE config 5 size: two parameter sets run through the UNMODIFIED reference + AADC (tests/golden/make_goldens.py
    section 5), compared with the cluster-pair kernel directly (not via the C port)."""
    N, B = 256, 2
    g, k = synth_goldens, "glv_N256_ck54_1e-8"
    p = oracle.synth_params(oracle.SYS_GLV, N, 1234, 0, B)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
        r = e.forward_adjoint(oracle.synth_x0(oracle.SYS_GLV, N, p), p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
        assert e.info()["kernel_name"] == "k_glv_pair"
    np.testing.assert_array_equal(r["n_accept"], g[k + "_steps"])
    assert_close(r["x_final"], g[k + "_x_final"], what="x(tf)")
    assert_close(r["lam"][:, 0], g[k + "_lam"], what="lambda")
    assert_close(r["mu"][:, 0, :N], g[k + "_mu_r"], what="mu, growth rates")
    assert_close(r["mu"][:, 0, N::16], g[k + "_mu_A_every16"], what="mu, every 16th matrix entry")
    np.testing.assert_allclose(r["mu"][:, 0].sum(axis=1), g[k + "_mu_sum"], rtol=1e-9)


@pytest.mark.parametrize("kernel", ["k_glv_pair:2", "k_glv_pair:4", "k_glv_ring"])
def test_glv256_many_waves_summed_mode(va, monkeypatch, kernel):
    """More trajectories than resident CTAs / CTA pairs (each integrates several trajectories and keeps adding to its
    partial-sum row): the summed gradient equals the sum of the per-trajectory gradients, and a replicated parameter set
    gives identical rows."""
    kernel, _, cl = kernel.partition(":")
    if kernel == "k_glv_ring":
        monkeypatch.setenv("VA_GLV_NO_PAIR", "1")
    else:
        monkeypatch.setenv("VA_GLV_CLUSTER", cl)
    N, B = 256, 400
    base = oracle.synth_params(oracle.SYS_GLV, N, 99, 0, 3)
    p = base[np.arange(B) % 3]
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
        assert e.info()["kernel_name"] == kernel
        r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
        s = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
    assert (r["status"] == 0).all()
    for k in range(3):
        np.testing.assert_array_equal(r["mu"][k::3], np.broadcast_to(r["mu"][k], r["mu"][k::3].shape))
        np.testing.assert_array_equal(r["x_final"][k::3], np.broadcast_to(r["x_final"][k], r["x_final"][k::3].shape))
    assert_close(s["mu"], r["mu"][:, 0].sum(axis=0, keepdims=True), rtol=1e-11, what="mu sum")


def test_glv256_full_size_properties_and_finite_differences(va):
    """BASELINE config 5 size class (N = 256, tol 1e-8) on device-generated inputs through the cluster kernel: determinism,
    linearity of the reverse sweep in its seed, summed mode, step-count statistics, a sampled oracle comparison (the generator
    is bit-identical host/device), and central finite differences of J = sum x_i(tf) (fixed-step RK4, so J is smooth)."""
    import torch
    N, B = 256, 1500
    npar = N * N + N
    dev = torch.device("cuda:0")
    p = torch.empty(B, npar, dtype=torch.float64, device=dev)
    x0 = torch.empty(B, N, dtype=torch.float64, device=dev)
    va.synth_batch_device(va.SYS_GLV, N, 1234, 0, B, p, x0)
    mk = lambda *s: torch.empty(*s, dtype=torch.float64, device=dev)
    xf, lam1, mu1 = mk(B, N), torch.ones(B, 1, N, dtype=torch.float64, device=dev), mk(B, 1, npar)
    na = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
        assert e.info()["kernel_name"] == "k_glv_pair"
        e.call("va_forward_adjoint_batch", B, x0, p, 0.0, 10.0, 1e-3, xf, lam1, mu1, va.OBJ_SEED, va.REDUCE_NONE, na, None, st)
        lam3, mu3 = torch.full((B, 1, N), -2.5, dtype=torch.float64, device=dev), mk(B, 1, npar)
        xf2 = mk(B, N)
        e.call("va_forward_adjoint_batch", B, x0, p, 0.0, 10.0, 1e-3, xf2, lam3, mu3, va.OBJ_SEED, va.REDUCE_NONE)
        musum = mk(1, npar)
        lam_s = torch.ones(B, 1, N, dtype=torch.float64, device=dev)
        e.call("va_forward_adjoint_batch", B, x0, p, 0.0, 10.0, 1e-3, xf2, lam_s, musum, va.OBJ_SEED, va.REDUCE_SUM)
        mu1b = mk(B, 1, npar)
        lam1b = torch.ones(B, 1, N, dtype=torch.float64, device=dev)
        e.call("va_forward_adjoint_batch", B, x0, p, 0.0, 10.0, 1e-3, xf2, lam1b, mu1b, va.OBJ_SEED, va.REDUCE_NONE)
        torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0
    assert torch.equal(xf, xf2) and torch.equal(mu1, mu1b) and torch.equal(lam1, lam1b)  # deterministic, bit for bit
    assert torch.isfinite(mu1).all()
    assert float((mu3 + 2.5 * mu1).abs().max() / mu1.abs().max()) < 1e-13
    assert float((lam3 + 2.5 * lam1).abs().max() / lam1.abs().max()) < 1e-13
    ref_sum = mu1[:, 0].sum(dim=0)
    assert float((musum[0] - ref_sum).abs().max() / ref_sum.abs().max()) < 1e-11
    steps = na.cpu().numpy()
    assert 15 <= steps.min() and steps.max() <= 40, (steps.min(), steps.max())
    idx = np.array([0, 73, 74, 777, B - 1])  # 74 = second trajectory of pair 0
    ph = p[idx].cpu().numpy()
    np.testing.assert_array_equal(ph, np.concatenate([oracle.synth_params(oracle.SYS_GLV, N, 1234, int(i), 1) for i in idx]))
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, x0[idx].cpu().numpy(), ph, 0.0, 10.0, 1e-3,
                               objective=oracle.OBJ_SUM, threads=8)
    np.testing.assert_array_equal(steps[idx], o["n_accept"])
    assert_close(xf[idx].cpu().numpy(), o["x_final"], what="x(tf)")
    assert_close(mu1[idx, 0].cpu().numpy(), o["mu"], what="mu")
    assert_close(lam1[idx, 0].cpu().numpy(), o["lam"], what="lambda")
    # finite differences, fixed-step RK4 on the same kernel
    p1 = ph[:1]
    x1 = x0[idx[:1]].cpu().numpy()
    ks = [0, 100, 255, 256, 256 + 257, 256 + 256 * 100 + 3, 256 + 256 * 255 + 255]
    h = 1e-6
    pp = np.repeat(p1, 2 * len(ks), axis=0)
    for m, k in enumerate(ks):
        pp[2 * m, k] += h
        pp[2 * m + 1, k] -= h
    with va.Engine(va.SYS_GLV, N, va.RK_RK4, False, max_steps=256) as e:
        assert e.info()["kernel_name"] == "k_glv_pair"
        base = e.forward_adjoint(x1, p1, 0.0, 1.0, 0.01, objective=va.OBJ_SUM)
        pert = e.forward_adjoint(np.repeat(x1, 2 * len(ks), axis=0), pp, 0.0, 1.0, 0.01, objective=va.OBJ_SUM)
    J = pert["x_final"].sum(axis=1)
    for m, k in enumerate(ks):
        fd = (J[2 * m] - J[2 * m + 1]) / (2 * h)
        assert abs(fd - base["mu"][0, 0, k]) <= 1e-7 * abs(fd) + 5e-9, (k, fd, base["mu"][0, 0, k])


def test_backward_in_time_integration(va):
    """dt0 < 0 (tf < ti): odeint's less_with_sign logic is sign-aware (reference lib/include/detail/runge_kutta.hpp:93,98)."""
    B = 64
    p = np.minimum(oracle.synth_params(oracle.SYS_VANDERPOL, 2, 3, 0, B), 4.0)  # backward in time the stiff cases blow up
    x0 = oracle.synth_x0(oracle.SYS_VANDERPOL, 2, p)
    o = oracle.forward_adjoint(oracle.SYS_VANDERPOL, 2, oracle.RK_CK54, True, 1e-7, 1e-7, x0, p, 0.3, 0.0, -1e-3, objective=oracle.OBJ_SUM)
    assert np.isfinite(o["x_final"]).all()
    with va.Engine(va.SYS_VANDERPOL, 2, va.RK_CK54, True, 1e-7, 1e-7, max_steps=4096) as e:
        r = e.forward_adjoint(x0, p, 0.3, 0.0, -1e-3, objective=va.OBJ_SUM)
    assert (r["status"] == 0).all()
    np.testing.assert_array_equal(r["n_accept"], o["n_accept"])
    np.testing.assert_array_equal(r["x_final"], o["x_final"])
    assert_close(r["mu"][:, 0], o["mu"], rtol=1e-11, what="mu")
    N = 16
    pg = oracle.synth_params(oracle.SYS_GLV, N, 3, 0, 8)
    xg = oracle.synth_x0(oracle.SYS_GLV, N, pg)
    og = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, xg, pg, 0.2, 0.0, -1e-3, objective=oracle.OBJ_SUM)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
        rg = e.forward_adjoint(xg, pg, 0.2, 0.0, -1e-3, objective=va.OBJ_SUM)
    np.testing.assert_array_equal(rg["n_accept"], og["n_accept"])
    assert_close(rg["x_final"], og["x_final"], what="x(tf)")
    assert_close(rg["mu"][:, 0], og["mu"], what="mu")


def test_no_progress_status_and_error_paths(va):
    """A hopeless tolerance makes the controller reject 500 times: per-trajectory status, batch not aborted (reference:
    odeint::no_progress_error from failed_step_checker, lib/include/detail/runge_kutta.hpp:85-86,106)."""
    with pytest.raises(va.EngineError):
        va.Engine(va.SYS_VANDERPOL, 2, 99, True, 1e-6, 1e-6)  # unknown stepper ("This ... stepper is not supported yet!")
    with pytest.raises(va.EngineError):
        va.Engine(va.SYS_VANDERPOL, 2, va.RK_RK4, True, 1e-6, 1e-6)  # rk4 has no error estimate: cannot be controlled
    with pytest.raises(va.EngineError):
        va.Engine(va.SYS_GLV, 8, va.RK_CK54, True, 1e-6, 1e-6, n_par=5)  # wrong parameter count
    p = np.array([[1000.0], [1.0]])
    x0 = oracle.synth_x0(oracle.SYS_VANDERPOL, 2, p)
    with va.Engine(va.SYS_VANDERPOL, 2, va.RK_CK54, True, 1e-300, 0.0, max_steps=64) as e:
        r = e.forward_adjoint(x0, p, 0.0, 0.5, 1e-3, objective=va.OBJ_SUM)
    o = oracle.forward_adjoint(oracle.SYS_VANDERPOL, 2, oracle.RK_CK54, True, 1e-300, 0.0, x0, p, 0.0, 0.5, 1e-3, objective=oracle.OBJ_SUM)
    np.testing.assert_array_equal(r["status"] & va.TRAJ_NO_PROGRESS != 0, o["status"] == 2)
    assert (r["status"] != 0).all() and np.isnan(r["mu"]).all()
