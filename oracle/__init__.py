"""ctypes front-end of the parity oracle -- TEST INFRASTRUCTURE ONLY.

Two back-ends, both CPU:

* ``port``      -- oracle/libva_oracle.so, the plain-C restatement (oracle/va_oracle.c) of the reference hot path
                   (reference lib/include/detail/runge_kutta.hpp:38-118, detail/backpropagation.hpp:24-348).
* ``reference`` -- oracle/_ref/libva_ref.so, the UNMODIFIED reference headers + AADC driven through the reference's
                   public API (oracle/ref_driver.cpp). Built in the container from /root/reference by oracle/Makefile;
                   travels to the GPU box as a prebuilt, git-ignored artefact.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may import this
module. The product package (vectorizedadjoint_b200) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_c_dp = ctypes.POINTER(ctypes.c_double)
_c_ip = ctypes.POINTER(ctypes.c_int32)

SYS_HARMONIC, SYS_VANDERPOL, SYS_GLV = 0, 1, 2
# reference back-end only (recorded example systems of vectorizedadjoint_b200/examples/tape_systems.hpp; n = 2, npar = 3)
SYS_PENDULUM, SYS_SWITCHED, SYS_PENDULUM_AUTONOMOUS, SYS_SWITCHED_AUTONOMOUS, SYS_HARVESTED_GLV = 3, 4, 5, 6, 7
RK_CK54_FIXED, RK_RKF78_FIXED = 12, 14  # reference back-end: error steppers through the fixed-step loop
RK_EULER, RK_RK4, RK_CK54, RK_DOPRI5, RK_RKF78 = 0, 1, 2, 3, 4
OBJ_SEED, OBJ_SUM, OBJ_HALF_NORM2 = 0, 1, 2


def _dp(a):
    return a.ctypes.data_as(_c_dp)


def _ip(a):
    return a.ctypes.data_as(_c_ip)


def build(ref: bool = True) -> None:
    """(Re)build the oracle; the reference leg is rebuilt only when /root/reference is present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        path = os.path.join(HERE, "libva_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = ctypes.CDLL(path)
        L.vo_forward_adjoint_batch.restype = ctypes.c_int
        L.vo_forward_adjoint_batch.argtypes = (
            [ctypes.c_int] * 5 + [ctypes.c_double] * 2 + [ctypes.c_long, _c_dp, _c_dp] + [ctypes.c_double] * 3
            + [ctypes.c_int, ctypes.c_long, _c_dp, _c_dp, _c_dp, _c_ip, _c_ip, _c_ip, ctypes.c_int])
        L.vo_synth_params.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_long, ctypes.c_long, _c_dp]
        L.vo_synth_x0.argtypes = [ctypes.c_int, ctypes.c_int, _c_dp, ctypes.c_long, _c_dp]
        L.vo_u01.restype = ctypes.c_double
        L.vo_u01.argtypes = [ctypes.c_uint64] * 4
        _port = L
    return _port


def reference_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libva_ref.so"))


def ref_lib():
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libva_ref.so")
        L = ctypes.CDLL(path)
        L.va_ref_forward_adjoint_batch.restype = ctypes.c_int
        L.va_ref_forward_adjoint_batch.argtypes = (
            [ctypes.c_int] * 5 + [ctypes.c_double] * 2 + [ctypes.c_long, _c_dp, _c_dp] + [ctypes.c_double] * 3
            + [ctypes.c_int, _c_dp, _c_dp, _c_dp, _c_ip, ctypes.c_int])
        L.va_ref_describe.restype = ctypes.c_char_p
        _ref = L
    return _ref


def npar_of(sys: int, n: int) -> int:
    return n * n + n if sys in (SYS_GLV, SYS_HARVESTED_GLV) else 3 if sys >= SYS_PENDULUM else 1


def synth_params(sys: int, n: int, seed: int, b0: int, B: int) -> np.ndarray:
    p = np.empty((B, npar_of(sys, n)), dtype=np.float64)
    port_lib().vo_synth_params(sys, n, seed, b0, B, _dp(p))
    return p


def synth_x0(sys: int, n: int, p: np.ndarray) -> np.ndarray:
    B = p.shape[0]
    x0 = np.empty((B, n), dtype=np.float64)
    port_lib().vo_synth_x0(sys, n, _dp(np.ascontiguousarray(p)), B, _dp(x0))
    return x0


def forward_adjoint(sys, n, stepper, adaptive, eps_abs, eps_rel, x0, p, ti, tf, dt0, objective=OBJ_SUM, seeds=None,
                    ck_cap=8192, threads=1):
    """Plain-C oracle. Returns dict(x_final[B,n], lam[B,n], mu[B,npar], n_accept, n_reject, status)."""
    npar = npar_of(sys, n)
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, n)
    B = x0.shape[0]
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(B, npar)
    xf = np.zeros((B, n))
    lam = np.zeros((B, n)) if seeds is None else np.array(seeds, dtype=np.float64).reshape(B, n).copy()
    mu = np.zeros((B, npar))
    na, nr, st = (np.zeros(B, np.int32) for _ in range(3))
    rc = port_lib().vo_forward_adjoint_batch(sys, n, npar, stepper, int(adaptive), eps_abs, eps_rel, B, _dp(x0), _dp(p),
                                             ti, tf, dt0, objective, ck_cap, _dp(xf), _dp(lam), _dp(mu), _ip(na), _ip(nr),
                                             _ip(st), threads)
    if rc != 0:
        raise RuntimeError(f"vo_forward_adjoint_batch failed: {rc}")
    return dict(x_final=xf, lam=lam, mu=mu, n_accept=na, n_reject=nr, status=st)


def reference_forward_adjoint(sys, n, stepper, eps_abs, eps_rel, x0, p, ti, tf, dt0, objective=OBJ_SUM, seeds=None, nout=1,
                              threads=1):
    """Unmodified reference (lib/include + AADC). lam[B,nout,n], mu[B,nout,npar]."""
    npar = npar_of(sys, n)
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, n)
    B = x0.shape[0]
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(B, npar)
    xf = np.zeros((B, n))
    lam = np.zeros((B, nout, n)) if seeds is None else np.array(seeds, dtype=np.float64).reshape(B, nout, n).copy()
    mu = np.zeros((B, nout, npar))
    na = np.zeros(B, np.int32)
    rc = ref_lib().va_ref_forward_adjoint_batch(sys, n, npar, nout, stepper, eps_abs, eps_rel, B, _dp(x0), _dp(p), ti, tf,
                                                dt0, objective, _dp(xf), _dp(lam), _dp(mu), _ip(na), threads)
    if rc != 0:
        raise RuntimeError(f"va_ref_forward_adjoint_batch failed: {rc}")
    return dict(x_final=xf, lam=lam, mu=mu, n_accept=na)
