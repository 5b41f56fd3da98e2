"""vectorizedadjoint_b200 -- B200-native batched forward + discrete-adjoint engine for the VectorizedAdjoint hot path.

This module is a thin ctypes binding of the C-ABI in ``include/va_engine.h`` (``libva_engine.so``, hand-written
sm_100a CUDA). It mirrors the reference's call sequence (reference ``lib/include``):

    Driver(Nin, Nout, Npar) + constructDriverButcherTableau + recordDriverRHSFunction  ->  Engine(...)
    runge_kutta(stepper, system, x0, alphas, ti, tf, dt, driver)                        ->  Engine.forward(...)
    setCostGradients(driver, lambda, mu); adjointSolve(driver, alphas)                  ->  Engine.adjoint(...)
    (both, checkpoints never leaving the GPU)                                           ->  Engine.forward_adjoint(...)

with one extra axis: a batch of parameter sets. There is NO CPU fallback: importing works without a GPU (so the
symbols can be inspected), but creating an Engine without the CUDA library or without a device raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

__all__ = ["Engine", "EngineError", "lib", "build", "SYS_HARMONIC", "SYS_VANDERPOL", "SYS_GLV", "RK_EULER", "RK_RK4", "RK_CK54",
           "RK_DOPRI5", "RK_RKF78", "OBJ_SEED", "OBJ_SUM", "OBJ_HALF_NORM2", "REDUCE_NONE", "REDUCE_SUM", "synth_batch_device",
           "measure_fp64_peak", "measure_hbm_copy", "measure_h2d_copy", "host_alloc", "HostBuffer", "comm_unique_id", "npar_of", "shard_range"]

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VA_ENGINE_LIB") or os.path.join(HERE, "libva_engine.so")  # VA_ENGINE_LIB: experiment builds

SYS_HARMONIC, SYS_VANDERPOL, SYS_GLV, SYS_TAPE = 0, 1, 2, 3
RK_EULER, RK_RK4, RK_CK54, RK_DOPRI5, RK_RKF78 = 0, 1, 2, 3, 4
OBJ_SEED, OBJ_SUM, OBJ_HALF_NORM2 = 0, 1, 2
REDUCE_NONE, REDUCE_SUM = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
CKPT_AUTO, CKPT_RECOMPUTE, CKPT_STORE_STAGES, CKPT_SPARSE = 0, 1, 2, 3
TRAJ_OK, TRAJ_CKPT_OVERFLOW, TRAJ_NO_PROGRESS, TRAJ_NONFINITE = 0, 1, 2, 4

HOST_DEFAULT, HOST_WRITE_COMBINED, HOST_NUMA_LOCAL = 0, 1, 2
COMM_ID_BYTES = 128

EXPORTS = ["va_engine_create", "va_engine_destroy", "va_engine_get_info", "va_last_error", "va_forward_batch", "va_adjoint_batch",
           "va_forward_adjoint_batch", "va_forward_adjoint_batch_sharded", "va_shard_range", "va_comm_unique_id", "va_engine_comm_init",
           "va_get_checkpoints", "va_tape_compile_check", "va_synth_batch_device", "va_measure_fp64_peak", "va_measure_hbm_copy",
           "va_measure_h2d_copy", "va_host_alloc", "va_host_free"]


class EngineError(RuntimeError):
    pass


class _Desc(ctypes.Structure):
    _fields_ = [("system", ctypes.c_int32), ("n_state", ctypes.c_int32), ("n_par", ctypes.c_int32), ("n_out", ctypes.c_int32),
                ("stepper", ctypes.c_int32), ("adaptive", ctypes.c_int32), ("eps_abs", ctypes.c_double), ("eps_rel", ctypes.c_double),
                ("device", ctypes.c_int32), ("max_steps", ctypes.c_int32), ("ckpt_policy", ctypes.c_int32), ("n_devices", ctypes.c_int32),
                ("workspace_fraction", ctypes.c_double), ("tape_cuda_src", ctypes.c_char_p), ("devices", ctypes.POINTER(ctypes.c_int32))]


class _Args(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int64), ("x0", ctypes.c_void_p), ("params", ctypes.c_void_p), ("ti", ctypes.c_double),
                ("tf", ctypes.c_double), ("dt0", ctypes.c_double), ("objective", ctypes.c_int32), ("reduce", ctypes.c_int32),
                ("mem", ctypes.c_int32), ("reserved0", ctypes.c_int32), ("x_final", ctypes.c_void_p), ("lambda_", ctypes.c_void_p),
                ("mu", ctypes.c_void_p), ("n_accept", ctypes.c_void_p), ("n_reject", ctypes.c_void_p), ("status", ctypes.c_void_p),
                ("stream", ctypes.c_void_p)]


class _Info(ctypes.Structure):
    _fields_ = [("api_version", ctypes.c_int32), ("device", ctypes.c_int32), ("sm_count", ctypes.c_int32), ("kernel_family", ctypes.c_int32),
                ("ckpt_policy", ctypes.c_int32), ("max_steps", ctypes.c_int32), ("ctas_per_sm", ctypes.c_int32),
                ("threads_per_cta", ctypes.c_int32), ("workspace_bytes", ctypes.c_int64), ("chunk_trajectories", ctypes.c_int64),
                ("kernel_launches", ctypes.c_int64), ("last_kernel_ms", ctypes.c_double), ("device_name", ctypes.c_char * 64),
                ("kernel_name", ctypes.c_char * 32), ("n_devices", ctypes.c_int32), ("comm_world", ctypes.c_int32),
                ("comm_rank", ctypes.c_int32), ("nccl_version", ctypes.c_int32), ("collectives", ctypes.c_int64)]


def build(verbose: bool = False) -> str:
    """Compile libva_engine.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library. Raises (loudly) when it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the CUDA extension is the only implementation; there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.va_engine_create.argtypes = [ctypes.POINTER(_Desc), ctypes.POINTER(ctypes.c_void_p)]
        L.va_engine_destroy.argtypes = [ctypes.c_void_p]
        L.va_engine_destroy.restype = None
        L.va_engine_get_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Info)]
        L.va_last_error.restype = ctypes.c_char_p
        for name in ("va_forward_batch", "va_adjoint_batch", "va_forward_adjoint_batch"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.POINTER(_Args)]
        L.va_get_checkpoints.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.POINTER(ctypes.c_int32)]
        L.va_synth_batch_device.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.va_forward_adjoint_batch_sharded.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(_Args)]
        L.va_shard_range.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
        L.va_shard_range.restype = None
        L.va_comm_unique_id.argtypes = [ctypes.c_void_p, ctypes.c_int32]
        L.va_engine_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
        L.va_measure_h2d_copy.argtypes = [ctypes.POINTER(ctypes.c_int32), ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.va_host_alloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]
        L.va_host_free.argtypes = [ctypes.c_void_p]
        L.va_measure_fp64_peak.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]
        L.va_measure_hbm_copy.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise EngineError(f"{what} failed ({rc}): {lib().va_last_error().decode()}")


def shard_range(batch: int, rank: int, world: int):
    """Contiguous shard [b0, b0 + count) of a batch of parameter sets for `rank` of `world` (the only multi-GPU
    decomposition on this path: independent trajectories, no data-path collective). Same arithmetic as the C-ABI's
    va_shard_range (kept in Python too so that host-side logic can be tested without the library)."""
    base, rem = divmod(batch, world)
    count = base + (1 if rank < rem else 0)
    b0 = rank * base + min(rank, rem)
    return b0, count


def npar_of(system: int, n: int) -> int:
    return n * n + n if system == SYS_GLV else 1


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return ctypes.c_void_p(x.data_ptr())
    return ctypes.c_void_p(x.ctypes.data)


class Engine:
    """One engine = one ODE system + stepper + tolerances on one GPU (the reference's Driver, batched)."""

    def __init__(self, system: int, n_state: int, stepper: int, adaptive: bool, eps_abs: float = 0.0, eps_rel: float = 0.0,
                 n_out: int = 1, device: int = 0, max_steps: int = 0, n_par: int | None = None, workspace_fraction: float = 0.0,
                 ckpt_policy: int = 0, devices=None, tape_cuda_src: str | None = None):
        """devices=[g0, g1, ...]: a multi-device engine -- every batch call shards over those GPUs inside the C-ABI."""
        self._h = ctypes.c_void_p()
        self.system, self.n, self.n_out = system, n_state, n_out
        self.npar = npar_of(system, n_state) if n_par is None else n_par
        self.devices = list(devices) if devices else [device]
        dev_arr = (ctypes.c_int32 * len(self.devices))(*self.devices) if devices else None
        d = _Desc(system, n_state, self.npar, n_out, stepper, int(adaptive), eps_abs, eps_rel, self.devices[0], max_steps, ckpt_policy,
                  len(self.devices) if devices else 0, workspace_fraction, tape_cuda_src.encode() if tape_cuda_src else None, dev_arr)
        _check(lib().va_engine_create(ctypes.byref(d), ctypes.byref(self._h)), "va_engine_create")

    def comm_init(self, comm_id: bytes, rank: int, world: int):
        """One process per GPU: attach this (single-device) engine to the communicator identified by `comm_id`
        (comm_unique_id() of rank 0, handed around by the launcher). Collective over all ranks. Afterwards REDUCE_SUM
        calls end in one ncclAllReduce over the ranks, inside the C-ABI call."""
        buf = ctypes.create_string_buffer(bytes(comm_id), COMM_ID_BYTES)
        _check(lib().va_engine_comm_init(self._h, buf, COMM_ID_BYTES, rank, world), "va_engine_comm_init")

    def call_sharded(self, shards):
        """Multi-device engine, caller-sharded: `shards` = one dict per device with the keyword arguments of call()
        (B, x0, params, ti, tf, dt0, x_final, lam, mu, objective, reduce, n_accept, n_reject, status, stream)."""
        arr = (_Args * len(shards))()
        for k, sh in enumerate(shards):
            arr[k] = self._args(**sh)
        _check(lib().va_forward_adjoint_batch_sharded(self._h, len(shards), arr), "va_forward_adjoint_batch_sharded")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().va_engine_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def info(self) -> dict:
        i = _Info()
        _check(lib().va_engine_get_info(self._h, ctypes.byref(i)), "va_engine_get_info")
        d = {k: getattr(i, k) for k, _ in _Info._fields_}
        d["device_name"] = i.device_name.decode()
        d["kernel_name"] = i.kernel_name.decode()
        return d

    # ---- raw call: caller-provided buffers (numpy = host memory, torch.cuda = device memory) ----------------------
    def call(self, which: str, B: int, x0, params, ti, tf, dt0, x_final, lam, mu, objective=OBJ_SUM, reduce=REDUCE_NONE,
             n_accept=None, n_reject=None, status=None, stream=None):
        a = self._args(B, x0, params, ti, tf, dt0, x_final, lam, mu, objective, reduce, n_accept, n_reject, status, stream)
        _check(getattr(lib(), which)(self._h, ctypes.byref(a)), which)

    @staticmethod
    def _args(B, x0, params, ti, tf, dt0, x_final, lam, mu, objective=OBJ_SUM, reduce=REDUCE_NONE, n_accept=None, n_reject=None,
              status=None, stream=None):
        bufs = [b for b in (x0, params, x_final, lam, mu, n_accept, n_reject, status) if b is not None]
        on_dev = [_is_torch(b) and b.is_cuda for b in bufs]
        if any(on_dev) and not all(on_dev):
            raise EngineError("mixing host and device buffers in one call")
        for b in bufs:
            if _is_torch(b):
                assert b.is_contiguous()
            else:
                assert b.flags["C_CONTIGUOUS"]
        return _Args(B, _ptr(x0), _ptr(params), ti, tf, dt0, objective, reduce, MEM_DEVICE if all(on_dev) and bufs else MEM_HOST, 0,
                     _ptr(x_final), _ptr(lam), _ptr(mu), _ptr(n_accept), _ptr(n_reject), _ptr(status),
                     ctypes.c_void_p(stream) if stream else None)

    # ---- convenience (numpy in / numpy out) -------------------------------------------------------------------------
    def forward_adjoint(self, x0, params, ti, tf, dt0, objective=OBJ_SUM, seeds=None, reduce=REDUCE_NONE):
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, self.n)
        B = x0.shape[0]
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(B, self.npar)
        xf = np.zeros((B, self.n))
        lam = np.zeros((B, self.n_out, self.n)) if seeds is None else np.array(seeds, dtype=np.float64).reshape(B, self.n_out, self.n)
        mu = np.zeros((self.n_out, self.npar)) if reduce == REDUCE_SUM else np.zeros((B, self.n_out, self.npar))
        na, nr, st = (np.zeros(B, np.int32) for _ in range(3))
        self.call("va_forward_adjoint_batch", B, x0, params, ti, tf, dt0, xf, lam, mu, objective, reduce, na, nr, st)
        return dict(x_final=xf, lam=lam, mu=mu, n_accept=na, n_reject=nr, status=st)

    def forward(self, x0, params, ti, tf, dt0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, self.n)
        B = x0.shape[0]
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(B, self.npar)
        xf = np.zeros((B, self.n))
        na, nr, st = (np.zeros(B, np.int32) for _ in range(3))
        self.call("va_forward_batch", B, x0, params, ti, tf, dt0, xf, None, None, OBJ_SUM, REDUCE_NONE, na, nr, st)
        self._B = B
        return dict(x_final=xf, n_accept=na, n_reject=nr, status=st)

    def adjoint(self, objective=OBJ_SEED, seeds=None, reduce=REDUCE_NONE):
        B = self._B
        lam = np.zeros((B, self.n_out, self.n)) if seeds is None else np.array(seeds, dtype=np.float64).reshape(B, self.n_out, self.n)
        mu = np.zeros((self.n_out, self.npar)) if reduce == REDUCE_SUM else np.zeros((B, self.n_out, self.npar))
        a = _Args(B, None, None, 0.0, 0.0, 0.0, objective, reduce, MEM_HOST, 0, None, _ptr(lam), _ptr(mu), None, None, None, None)
        _check(lib().va_adjoint_batch(self._h, ctypes.byref(a)), "va_adjoint_batch")
        return dict(lam=lam, mu=mu)

    def checkpoints(self, b: int):
        cnt = ctypes.c_int32()
        _check(lib().va_get_checkpoints(self._h, b, 0, None, None, ctypes.byref(cnt)), "va_get_checkpoints")
        t = np.zeros(cnt.value)
        x = np.zeros((cnt.value, self.n))
        _check(lib().va_get_checkpoints(self._h, b, cnt.value, _ptr(t), _ptr(x), ctypes.byref(cnt)), "va_get_checkpoints")
        return t, x


def synth_batch_device(system: int, n: int, seed: int, b0: int, B: int, params, x0=None, stream=None):
    """Fill device tensors with the seeded synthetic parameter sets (bit-identical to the host generator in oracle/)."""
    _check(lib().va_synth_batch_device(system, n, seed, b0, B, _ptr(params), _ptr(x0), ctypes.c_void_p(stream) if stream else None),
           "va_synth_batch_device")


def measure_fp64_peak(device: int = 0) -> float:
    v = ctypes.c_double()
    _check(lib().va_measure_fp64_peak(device, ctypes.byref(v)), "va_measure_fp64_peak")
    return v.value


def measure_hbm_copy(device: int = 0) -> float:
    v = ctypes.c_double()
    _check(lib().va_measure_hbm_copy(device, ctypes.byref(v)), "va_measure_hbm_copy")
    return v.value


def comm_unique_id() -> bytes:
    """Opaque id of a new communicator (rank 0 calls this, the launcher broadcasts it, every rank passes it to
    Engine.comm_init)."""
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    _check(lib().va_comm_unique_id(buf, COMM_ID_BYTES), "va_comm_unique_id")
    return buf.raw


class HostBuffer:
    """Page-locked host memory from va_host_alloc, viewed as a numpy array (flags: HOST_WRITE_COMBINED for input-only
    buffers, HOST_NUMA_LOCAL to bind the pages to the GPU's NUMA node)."""

    def __init__(self, shape, dtype=np.float64, flags: int = HOST_DEFAULT, device: int = 0):
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = ctypes.c_void_p()
        _check(lib().va_host_alloc(ctypes.byref(self._p), max(self.nbytes, 1), flags, device), "va_host_alloc")
        raw = (ctypes.c_char * max(self.nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    @property
    def ptr(self) -> int:
        return self._p.value

    def free(self):
        if self._p:
            self.array = None
            lib().va_host_free(self._p)
            self._p = ctypes.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def host_alloc(shape, dtype=np.float64, flags: int = HOST_DEFAULT, device: int = 0) -> HostBuffer:
    return HostBuffer(shape, dtype, flags, device)


def measure_h2d_copy(devices, nbytes: int = 1 << 30, reps: int = 3, flags: int = HOST_DEFAULT, host_ptr: int | None = None):
    """Host->device copy ceiling with all `devices` copying concurrently: (per-device GB/s list, aggregate GB/s)."""
    devs = (ctypes.c_int32 * len(devices))(*devices)
    per = (ctypes.c_double * len(devices))()
    agg = ctypes.c_double()
    _check(lib().va_measure_h2d_copy(devs, len(devices), nbytes, reps, flags, ctypes.c_void_p(host_ptr) if host_ptr else None, per,
                                     ctypes.byref(agg)), "va_measure_h2d_copy")
    return list(per), agg.value
