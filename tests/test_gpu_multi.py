"""GPU, two or more devices: the multi-GPU entry points of the C-ABI (include/va_engine.h, "Several GPUs") against the
single-GPU result of the same library (itself checked against the oracle in test_gpu_parity.py).

The reference has no parallel path at all (reference lib/include/AadData.hpp:32 is a TODO about it), so the statement to prove
is: sharding a batch over G GPUs inside the call changes nothing per parameter set (x(tf), dJ/dx0, dJ/dalpha, step counts:
bit-identical) and the all-reduced summed gradient equals the one-GPU sum to summation-order round-off (1e-12 relative).
Skipped on a box with one GPU; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import oracle  # noqa: E402  (checker only)
import vectorizedadjoint_b200 as va  # noqa: E402

SEED = 1234


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")


def _glv(n, B):
    p = oracle.synth_params(oracle.SYS_GLV, n, SEED, 0, B)
    return oracle.synth_x0(oracle.SYS_GLV, n, p), p


@needs2
@pytest.mark.parametrize("n,B", [(64, 1501), (16, 3001), (256, 9)])
def test_multi_device_engine_matches_single_gpu(n, B):
    x0, p = _glv(n, B)
    G = min(_ngpu(), 4)
    with va.Engine(va.SYS_GLV, n, va.RK_CK54, True, 1e-8, 1e-8, device=0) as e1:
        one = e1.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
        one_sum = e1.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
    with va.Engine(va.SYS_GLV, n, va.RK_CK54, True, 1e-8, 1e-8, devices=list(range(G))) as eg:
        many = eg.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM)
        many_sum = eg.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
        info = eg.info()
    assert info["n_devices"] == G and info["comm_world"] == G and info["collectives"] == G and info["nccl_version"] > 0
    assert (many["status"] == 0).all()
    for k in ("n_accept", "n_reject", "x_final", "lam", "mu"):
        assert np.array_equal(one[k], many[k]), k  # per parameter set: bit-identical, wherever it ran
    assert np.array_equal(one_sum["x_final"], many_sum["x_final"]) and np.array_equal(one_sum["n_accept"], many_sum["n_accept"])
    ref = one["mu"].sum(axis=0)
    scale = np.abs(ref).max()
    assert np.abs(many_sum["mu"] - one_sum["mu"]).max() <= 1e-12 * scale   # all-reduced sum == one-GPU sum
    assert np.abs(many_sum["mu"] - ref).max() <= 1e-12 * scale            # == sum of the per-set gradients


@needs2
def test_multi_device_engine_small_and_empty_shards():
    """Fewer parameter sets than GPUs: empty shards still join the collective."""
    G = min(_ngpu(), 4)
    x0, p = _glv(64, 1)
    with va.Engine(va.SYS_GLV, 64, va.RK_CK54, True, 1e-8, 1e-8, device=0) as e1:
        one = e1.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, reduce=va.REDUCE_SUM)
    with va.Engine(va.SYS_GLV, 64, va.RK_CK54, True, 1e-8, 1e-8, devices=list(range(G))) as eg:
        many = eg.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, reduce=va.REDUCE_SUM)
        empty = eg.forward_adjoint(x0[:0], p[:0], 0.0, 10.0, 1e-3, reduce=va.REDUCE_SUM)
    assert np.array_equal(one["x_final"], many["x_final"])
    assert np.abs(many["mu"] - one["mu"]).max() <= 1e-13 * np.abs(one["mu"]).max()
    assert (empty["mu"] == 0).all()


@needs2
def test_sharded_device_resident_call():
    """va_forward_adjoint_batch_sharded: every GPU holds its shard in HBM; each shard's mu receives the sum over all shards."""
    n, B, G = 64, 2400, 2
    x0, p = _glv(n, B)
    with va.Engine(va.SYS_GLV, n, va.RK_CK54, True, 1e-8, 1e-8, device=0) as e1:
        one = e1.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, reduce=va.REDUCE_SUM)
    shards, keep = [], []
    for g in range(G):
        b0, cnt = va.shard_range(B, g, G)
        dev = torch.device("cuda", g)
        t = dict(x0=torch.from_numpy(x0[b0:b0 + cnt]).to(dev), params=torch.from_numpy(p[b0:b0 + cnt]).to(dev),
                 x_final=torch.empty(cnt, n, dtype=torch.float64, device=dev), lam=torch.empty(cnt, 1, n, dtype=torch.float64, device=dev),
                 mu=torch.empty(1, n * n + n, dtype=torch.float64, device=dev), n_accept=torch.empty(cnt, dtype=torch.int32, device=dev),
                 n_reject=torch.empty(cnt, dtype=torch.int32, device=dev), status=torch.empty(cnt, dtype=torch.int32, device=dev))
        keep.append(t)
        shards.append(dict(B=cnt, ti=0.0, tf=10.0, dt0=1e-3, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM, **t))
    with va.Engine(va.SYS_GLV, n, va.RK_CK54, True, 1e-8, 1e-8, devices=[0, 1]) as eg:
        eg.call_sharded(shards)
        for g in range(G):
            torch.cuda.synchronize(g)
    xf = np.concatenate([t["x_final"].cpu().numpy() for t in keep])
    assert np.array_equal(xf, one["x_final"])
    for t in keep:
        assert np.abs(t["mu"].cpu().numpy() - one["mu"]).max() <= 1e-12 * np.abs(one["mu"]).max()
    assert torch.equal(keep[0]["mu"].cpu(), keep[1]["mu"].cpu())  # an all-reduce leaves the same bits everywhere


@needs2
def test_multi_device_split_api_and_checkpoints():
    """runge_kutta / adjointSolve / GetState as separate calls on a multi-device engine (thread-per-trajectory family)."""
    B = 1001
    pv = oracle.synth_params(oracle.SYS_VANDERPOL, 2, SEED, 0, B)
    xv = oracle.synth_x0(oracle.SYS_VANDERPOL, 2, pv)
    kw = dict(max_steps=2048)
    with va.Engine(va.SYS_VANDERPOL, 2, va.RK_DOPRI5, True, 1e-8, 1e-8, device=0, **kw) as e1:
        f1 = e1.forward(xv, pv, 0.0, 0.5, 1e-3)
        a1 = e1.adjoint(objective=va.OBJ_SUM)
        t1, x1 = e1.checkpoints(B - 1)
    with va.Engine(va.SYS_VANDERPOL, 2, va.RK_DOPRI5, True, 1e-8, 1e-8, devices=[0, 1], **kw) as eg:
        fg = eg.forward(xv, pv, 0.0, 0.5, 1e-3)
        ag = eg.adjoint(objective=va.OBJ_SUM)
        tg, xg = eg.checkpoints(B - 1)   # lives on the second GPU
        sg = eg.adjoint(objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
    assert np.array_equal(f1["x_final"], fg["x_final"]) and np.array_equal(f1["n_accept"], fg["n_accept"])
    assert np.array_equal(a1["lam"], ag["lam"]) and np.array_equal(a1["mu"], ag["mu"])
    assert np.array_equal(t1, tg) and np.array_equal(x1, xg)
    ref = a1["mu"].sum(axis=0)
    assert np.abs(sg["mu"] - ref).max() <= 1e-12 * np.abs(ref).max()


def _rank_main(rank, world, comm_id, n, B, q):
    import numpy as np
    import oracle
    import vectorizedadjoint_b200 as va
    b0, cnt = va.shard_range(B, rank, world)
    p = oracle.synth_params(oracle.SYS_GLV, n, SEED, b0, cnt)
    x0 = oracle.synth_x0(oracle.SYS_GLV, n, p)
    with va.Engine(va.SYS_GLV, n, va.RK_CK54, True, 1e-8, 1e-8, device=rank) as e:
        e.comm_init(comm_id, rank, world)
        r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
        info = e.info()
    q.put((rank, r["mu"], r["x_final"], r["n_accept"], info["comm_world"], info["comm_rank"], info["collectives"]))


@needs2
def test_one_process_per_gpu_allreduce_inside_the_call():
    """torchrun-style: one process per GPU, each with a single-device engine attached to a communicator
    (va_comm_unique_id + va_engine_comm_init); the summed gradient every rank gets is the one-GPU sum."""
    import torch.multiprocessing as mp
    n, B, world = 64, 1203, 2
    comm_id = va.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, comm_id, n, B, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    x0, p = _glv(n, B)
    with va.Engine(va.SYS_GLV, n, va.RK_CK54, True, 1e-8, 1e-8, device=0) as e1:
        one = e1.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SUM, reduce=va.REDUCE_SUM)
    assert np.array_equal(np.concatenate([g[2] for g in got]), one["x_final"])
    assert np.array_equal(np.concatenate([g[3] for g in got]), one["n_accept"])
    assert np.array_equal(got[0][1], got[1][1])
    assert np.abs(got[0][1] - one["mu"]).max() <= 1e-12 * np.abs(one["mu"]).max()
    assert [g[4] for g in got] == [world] * world and [g[5] for g in got] == list(range(world)) and all(g[6] == 1 for g in got)


@pytest.mark.skipif(_ngpu() < 1, reason="needs a GPU")
def test_host_buffers_and_h2d_ceiling():
    """va_host_alloc / va_measure_h2d_copy: page-locked buffers (plain and write-combined) feed the host pipeline with the
    same results as pageable numpy memory; the copy ceiling is a positive number per device."""
    n, B = 64, 700
    x0, p = _glv(n, B)
    devs = list(range(min(_ngpu(), 2)))
    per, agg = va.measure_h2d_copy(devs, nbytes=256 << 20, reps=2)
    assert len(per) == len(devs) and all(v > 1.0 for v in per) and agg > 1.0
    with va.Engine(va.SYS_GLV, n, va.RK_CK54, True, 1e-8, 1e-8, device=0) as e:
        ref = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, reduce=va.REDUCE_SUM)
        for flags in (va.HOST_DEFAULT, va.HOST_WRITE_COMBINED, va.HOST_NUMA_LOCAL):
            hp = va.host_alloc((B, n * n + n), flags=flags)
            hx = va.host_alloc((B, n), flags=flags)
            hp.array[:] = p
            hx.array[:] = x0
            xf, lam, mu = np.zeros((B, n)), np.zeros((B, 1, n)), np.zeros((1, n * n + n))
            e.call("va_forward_adjoint_batch", B, hx.array, hp.array, 0.0, 10.0, 1e-3, xf, lam, mu, va.OBJ_SUM, va.REDUCE_SUM)
            assert np.array_equal(xf, ref["x_final"]) and np.array_equal(mu, ref["mu"])
            hp.free()
            hx.free()
