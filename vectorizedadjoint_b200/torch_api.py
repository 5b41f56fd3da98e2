"""PyTorch front-end over the C-ABI: a differentiable batched ODE solve.

The reference lists a Python API as future work (README.md:55) and keeps a `void*` driver handle for it
(lib/include/Driver.hpp:81-85, runge_kutta.hpp:32-44). Here the handle is `va_engine*`:

    solver = OdeSolver(va.SYS_GLV, 64, va.RK_CK54, True, 1e-8, 1e-8, ti=0.0, tf=10.0, dt0=1e-3)
    x_tf = solver(x0, params)          # x0 [B, N], params [B, Npar]: float64 CUDA tensors, either may require grad
    loss = cost(x_tf); loss.backward() # x0.grad = dJ/dx(t0), params.grad = dJ/dalpha  (discrete adjoint of the RK scheme)

forward  = `va_forward_batch`   (reference runge_kutta(), lib/include/runge_kutta.hpp:47-59): forward sweep + checkpoints
backward = `va_adjoint_batch`   (reference adjointSolve(), lib/include/backpropagation.hpp:18-43) seeded with dJ/dx(tf),
           the seed the reference's users write into `lambda` before the call (examples/*/main.cpp).
PyTorch is plumbing here (device memory, the autograd graph); every number comes from the CUDA engine. There is no CPU
path: tensors must live on the engine's GPU.
"""
from __future__ import annotations

import torch

from . import OBJ_SEED, REDUCE_NONE, Engine, EngineError


class _Solve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, solver: "OdeSolver", x0: torch.Tensor, params: torch.Tensor):
        x_final = solver._forward(x0, params)
        ctx.solver = solver
        ctx.session = solver._session
        ctx.save_for_backward(x0, params)
        return x_final

    @staticmethod
    def backward(ctx, grad_x_final: torch.Tensor):
        solver: OdeSolver = ctx.solver
        x0, params = ctx.saved_tensors
        if solver._session != ctx.session:
            # another solve ran on this engine since: its checkpoints are gone, integrate again (deterministic)
            solver._forward(x0, params)
            ctx.session = solver._session
        lam, mu = solver._adjoint(grad_x_final)
        return None, (lam if ctx.needs_input_grad[1] else None), (mu if ctx.needs_input_grad[2] else None)


class OdeSolver:
    """x(tf) = solve(x0, params) for a batch of parameter sets, differentiable with respect to x0 and params.

    One OdeSolver owns one engine (one ODE system + stepper + tolerances on one GPU). `n_accept`, `n_reject` and `status`
    of the last solve are kept as int32 CUDA tensors; a trajectory whose status is non-zero (checkpoint store overflow,
    500 consecutive rejections, non-finite state) yields NaN results and NaN gradients, as in the C-ABI.
    """

    def __init__(self, system: int, n_state: int, stepper: int, adaptive: bool, eps_abs: float = 0.0, eps_rel: float = 0.0, *,
                 ti: float, tf: float, dt0: float, device: int = 0, max_steps: int = 0, n_par: int | None = None, ckpt_policy: int = 0):
        if not torch.cuda.is_available():
            raise EngineError("OdeSolver needs a CUDA device: this engine has no CPU path")
        self.engine = Engine(system, n_state, stepper, adaptive, eps_abs, eps_rel, n_out=1, device=device, max_steps=max_steps,
                             n_par=n_par, ckpt_policy=ckpt_policy)
        self.device = torch.device("cuda", device)
        self.ti, self.tf, self.dt0 = float(ti), float(tf), float(dt0)
        self.n, self.npar = self.engine.n, self.engine.npar
        self._session = 0
        self._B = 0
        self.n_accept = self.n_reject = self.status = None

    def close(self):
        self.engine.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, t: torch.Tensor, cols: int, what: str) -> torch.Tensor:
        if t.device != self.device or t.dtype != torch.float64:
            raise EngineError(f"{what}: expected a float64 tensor on {self.device}, got {t.dtype} on {t.device}")
        if t.dim() != 2 or t.shape[1] != cols:
            raise EngineError(f"{what}: expected shape [B, {cols}], got {tuple(t.shape)}")
        return t.detach().contiguous()

    def _forward(self, x0: torch.Tensor, params: torch.Tensor) -> torch.Tensor:
        x0c, pc = self._check(x0, self.n, "x0"), self._check(params, self.npar, "params")
        B = x0c.shape[0]
        if pc.shape[0] != B:
            raise EngineError("x0 and params differ in batch size")
        x_final = torch.empty_like(x0c)
        self.n_accept, self.n_reject, self.status = (torch.empty(B, dtype=torch.int32, device=self.device) for _ in range(3))
        torch.cuda.current_stream(self.device).synchronize()  # inputs may still be in flight on torch's stream
        self.engine.call("va_forward_batch", B, x0c, pc, self.ti, self.tf, self.dt0, x_final, None, None, OBJ_SEED, REDUCE_NONE,
                         self.n_accept, self.n_reject, self.status)
        self._session += 1
        self._B = B
        return x_final

    def _adjoint(self, seed: torch.Tensor):
        B = self._B
        lam = seed.detach().to(torch.float64).contiguous().reshape(B, 1, self.n).clone()  # overwritten with dJ/dx(t0)
        mu = torch.empty(B, 1, self.npar, dtype=torch.float64, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        self.engine.call("va_adjoint_batch", B, None, None, 0.0, 0.0, 0.0, None, lam, mu, OBJ_SEED, REDUCE_NONE)
        return lam.reshape(B, self.n), mu.reshape(B, self.npar)

    def __call__(self, x0: torch.Tensor, params: torch.Tensor) -> torch.Tensor:
        return _Solve.apply(self, x0, params)
