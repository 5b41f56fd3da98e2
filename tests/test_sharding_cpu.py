"""CPU, world_size 2 over gloo: the multi-GPU decomposition of the path. Each rank takes a contiguous shard of the
seeded parameter sets, produces its summed-objective gradient (here with the CPU oracle standing in for the kernel --
what is under test is the host-side sharding and the single all-reduce), and the all-reduced result must equal the
unsharded sum."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import vectorizedadjoint_b200 as va

N, B, SEED = 8, 13, 1234  # odd batch: uneven shards


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b0, cnt = va.shard_range(B, rank, world)
    p = oracle.synth_params(oracle.SYS_GLV, N, SEED, b0, cnt)  # the generator is indexed by the GLOBAL set number
    r = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, oracle.synth_x0(oracle.SYS_GLV, N, p), p, 0.0, 10.0, 1e-3,
                               objective=oracle.OBJ_SUM)
    mu = torch.from_numpy(r["mu"].sum(axis=0))
    dist.all_reduce(mu)  # the only collective on the path
    cnts = torch.tensor([cnt, b0])
    gathered = [torch.zeros_like(cnts) for _ in range(world)]
    dist.all_gather(gathered, cnts)
    if rank == 0:
        out.put((mu.numpy(), [g.tolist() for g in gathered]))
    dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    for batch in (0, 1, 7, 8, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [va.shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == batch
            for (b0, c), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_c_abi_shard_range_is_the_same_split():
    """va_shard_range (what a multi-device engine uses inside the call) == shard_range (what one-process-per-GPU launchers use)."""
    import ctypes
    L = va.lib()
    for batch in (0, 1, 7, 13, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 3, 4, 8):
            for g in range(world):
                b0, cnt = ctypes.c_int64(), ctypes.c_int64()
                L.va_shard_range(batch, g, world, ctypes.byref(b0), ctypes.byref(cnt))
                assert (b0.value, cnt.value) == va.shard_range(batch, g, world)


def test_multi_device_engine_needs_a_gpu_and_validates_its_device_list():
    """No GPU here: creation must fail loudly (no CPU path), and a malformed device list is rejected before anything else."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(va.EngineError):
        va.Engine(va.SYS_GLV, 8, va.RK_CK54, True, 1e-8, 1e-8, devices=[0, 1])
    with pytest.raises(va.EngineError, match="distinct"):
        va.Engine(va.SYS_GLV, 8, va.RK_CK54, True, 1e-8, 1e-8, devices=[0, 0])
    assert len(va.comm_unique_id()) == va.COMM_ID_BYTES  # NCCL is bound at run time and present in the image


def test_two_rank_summed_gradient_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mu, spans = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert spans == [[7, 0], [6, 7]]
    p = oracle.synth_params(oracle.SYS_GLV, N, SEED, 0, B)
    r = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, oracle.synth_x0(oracle.SYS_GLV, N, p), p, 0.0, 10.0, 1e-3,
                               objective=oracle.OBJ_SUM)
    np.testing.assert_allclose(mu, r["mu"].sum(axis=0), rtol=1e-13)
