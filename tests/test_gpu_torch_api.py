"""PyTorch front-end (vectorizedadjoint_b200/torch_api.py): autograd through the batched solve gives the engine's discrete
adjoint. Checked against the C-ABI call with the same seed, against the CPU oracle, and against central finite differences."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def va():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import vectorizedadjoint_b200 as va
    va.lib()
    return va


def test_autograd_matches_engine_and_oracle_glv(va):
    import torch
    from vectorizedadjoint_b200.torch_api import OdeSolver
    N, B = 64, 700  # more trajectories than one wave of resident slots (592 on a B200)
    p = oracle.synth_params(oracle.SYS_GLV, N, 31, 0, B)
    x0 = oracle.synth_x0(oracle.SYS_GLV, N, p)
    dev = torch.device("cuda", 0)
    xt = torch.tensor(x0, device=dev, requires_grad=True)
    pt = torch.tensor(p, device=dev, requires_grad=True)
    w = torch.linspace(0.5, 1.5, N, dtype=torch.float64, device=dev)
    with OdeSolver(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8, ti=0.0, tf=10.0, dt0=1e-3) as solve:
        xf = solve(xt, pt)
        loss = 0.5 * (w * xf * xf).sum()  # dJ/dx(tf) = w x(tf)
        loss.backward()
        assert (solve.status == 0).all()
        steps = solve.n_accept.cpu().numpy()
    seeds = (w * xf.detach()).cpu().numpy().reshape(B, 1, N)
    with va.Engine(va.SYS_GLV, N, va.RK_CK54, True, 1e-8, 1e-8) as e:
        r = e.forward_adjoint(x0, p, 0.0, 10.0, 1e-3, objective=va.OBJ_SEED, seeds=seeds)
    np.testing.assert_array_equal(steps, r["n_accept"])
    np.testing.assert_array_equal(xf.detach().cpu().numpy(), r["x_final"])
    np.testing.assert_array_equal(xt.grad.cpu().numpy(), r["lam"][:, 0])
    np.testing.assert_array_equal(pt.grad.cpu().numpy(), r["mu"][:, 0])
    idx = [0, 1, B - 1]
    o = oracle.forward_adjoint(oracle.SYS_GLV, N, oracle.RK_CK54, True, 1e-8, 1e-8, x0[idx], p[idx], 0.0, 10.0, 1e-3,
                               objective=oracle.OBJ_SEED, seeds=seeds[idx, 0])
    scale = np.abs(o["mu"]).max(axis=1, keepdims=True)
    assert (np.abs(pt.grad.cpu().numpy()[idx] - o["mu"]) / scale).max() < 1e-8


def test_autograd_finite_differences_vanderpol_and_interleaved_solves(va):
    import torch
    from vectorizedadjoint_b200.torch_api import OdeSolver
    dev = torch.device("cuda", 0)
    mu = torch.tensor([[0.5], [2.0], [7.5]], dtype=torch.float64, device=dev, requires_grad=True)
    x0 = torch.tensor([[2.0, 0.0]] * 3, dtype=torch.float64, device=dev, requires_grad=True)
    with OdeSolver(va.SYS_VANDERPOL, 2, va.RK_DOPRI5, True, 1e-11, 1e-11, ti=0.0, tf=0.5, dt0=1e-3, max_steps=4096) as solve:
        xf = solve(x0, mu)
        other = solve(x0.detach() * 1.01, mu.detach())  # a second solve on the same engine before backward: forces the re-integration
        (xf[:, 0] * xf[:, 1]).sum().backward()
        g_mu, g_x0 = mu.grad.clone(), x0.grad.clone()

        def J(xv, mv):
            y = solve(xv, mv)
            return (y[:, 0] * y[:, 1])

        h = 1e-6
        fd_mu = (J(x0.detach(), mu.detach() + h) - J(x0.detach(), mu.detach() - h)) / (2 * h)
        e0 = torch.tensor([[1.0, 0.0]], dtype=torch.float64, device=dev)
        fd_x0 = (J(x0.detach() + h * e0, mu.detach()) - J(x0.detach() - h * e0, mu.detach())) / (2 * h)
    assert other.shape == xf.shape
    np.testing.assert_allclose(g_mu[:, 0].cpu().numpy(), fd_mu.cpu().numpy(), rtol=2e-6)
    np.testing.assert_allclose(g_x0[:, 0].cpu().numpy(), fd_x0.cpu().numpy(), rtol=2e-6)


def test_rejects_wrong_device_dtype_shape(va):
    import torch
    from vectorizedadjoint_b200.torch_api import OdeSolver
    dev = torch.device("cuda", 0)
    with OdeSolver(va.SYS_HARMONIC, 2, va.RK_RK4, False, ti=0.0, tf=1.0, dt0=0.01, max_steps=128) as solve:
        good = torch.zeros(4, 2, dtype=torch.float64, device=dev)
        par = torch.full((4, 1), 0.151, dtype=torch.float64, device=dev)
        with pytest.raises(va.EngineError):
            solve(good.cpu(), par)
        with pytest.raises(va.EngineError):
            solve(good.float(), par)
        with pytest.raises(va.EngineError):
            solve(good, par[:3])
        assert solve(good, par).shape == (4, 2)
