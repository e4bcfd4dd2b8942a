"""CPU tests of the gradients-through-the-sampler path (no GPU needed):
  (a) the oracle's restated ``odeint_adjoint`` against back-propagation through a fixed-step RK4 integrator (an independent
      route to the same gradients) -- this is what pins the restatement, upstream torchdiffeq being absent;
  (b) the host logic of flowfusion_b200/adjoint.py (augmented system, flat adaptive driver, default adjoint norm, autograd
      hook) with the fused kernels replaced by their torch-CPU models, against the golden vectors that
      oracle/make_golden_adjoint.py took from the UNMODIFIED reference in training mode."""
import pytest
import torch

from conftest import load_golden
from kernel_model import patched_engine
from test_training import cpu_train_step, _close


def test_restated_odeint_adjoint_matches_backprop_through_rk4():
    from oracle import port  # noqa: F401  (puts oracle/torchdiffeq on the path)
    import torchdiffeq as tde
    torch.manual_seed(0)

    class Fn(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = torch.nn.Sequential(torch.nn.Linear(3, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))

        def forward(self, t, y):
            return self.net(y) * torch.cos(t)

    f = Fn()
    y0 = torch.randn(20, 3, requires_grad=True)
    y = tde.odeint_adjoint(f, y0, torch.tensor([0.0, 1.0]), rtol=1e-6, atol=1e-7)
    (y[-1] ** 2).sum().backward()
    ga, gy = [p.grad.clone() for p in f.parameters()], y0.grad.clone()
    f.zero_grad(); y0.grad = None
    yy, n = y0, 400
    h = 1.0 / n
    for i in range(n):
        tt = torch.tensor(i * h)
        k1 = f(tt, yy); k2 = f(tt + h / 2, yy + h / 2 * k1); k3 = f(tt + h / 2, yy + h / 2 * k2); k4 = f(tt + h, yy + h * k3)
        yy = yy + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    (yy ** 2).sum().backward()
    for a, p in zip(ga, f.parameters()):
        assert float((a - p.grad).abs().max() / p.grad.abs().max()) < 2e-5
    assert float((gy - y0.grad).abs().max() / y0.grad.abs().max()) < 2e-5
    # a tuple state and a descending time span go through the same code
    y = tde.odeint_adjoint(f, y0, torch.tensor([1.0, 0.0]), rtol=1e-6, atol=1e-7)
    assert y.shape == (2, 20, 3) and y.requires_grad


@pytest.fixture
def cpu_kernels(monkeypatch):
    from flowfusion_b200 import training
    monkeypatch.setattr(training, "train_step", cpu_train_step)
    with patched_engine():
        yield


GRAD_TOL = 2e-3      # of max(1, |grad|_inf): two adaptive solves whose step sequences may differ by rounding


@pytest.mark.parametrize("name", ["adjoint_vp_pfode", "adjoint_ve_sigma_pfode"])
def test_score_sampler_adjoint_host_logic(cpu_kernels, name):
    import flowfusion_b200.diffusion as D
    meta, sd, ins, outs = load_golden(name)
    sde = {"vp": D.VPSDE, "ve": D.VESDE}[meta["sde"]]()
    sm = D.ScoreModel(D.MLP(**meta["ctor"]), sde, no_sigma=meta["no_sigma"]).train()
    sm.load_state_dict(sd)
    base = ins["base"].clone().requires_grad_(True)
    opts = None if meta["step_t"] is None else {"step_t": torch.tensor([meta["step_t"]])}
    x, aux = sm.sample_ode_from_base(base, ins.get("cond"), atol=meta["tol"], rtol=meta["tol"], options=opts)
    assert aux == [] and x.requires_grad
    assert _close(outs["x"], x.detach(), 1e-4)
    (x * ins["w"]).sum().backward()
    assert _close(outs["grad/base"], base.grad, GRAD_TOL)
    for k, p in sm.named_parameters():
        if p.requires_grad:
            assert _close(outs["grad/" + k], p.grad, GRAD_TOL), k
    st = sm.last_adjoint_stats
    assert abs(st.accepted - meta["stats_backward"]["accepted"]) <= 2
    # eval mode: plain odeint, no graph (`diffusion.py:630-639`)
    sm.eval()
    x2, _ = sm.sample_ode_from_base(ins["base"], ins.get("cond"), atol=meta["tol"], rtol=meta["tol"], options=opts)
    assert not x2.requires_grad and _close(outs["x"], x2, 1e-4)
    # the log-likelihood paths still refuse training mode
    sm.train()
    with pytest.raises(NotImplementedError):
        sm.log_prob(ins["base"], ins.get("cond"))


def test_flow_sampler_adjoint_host_logic(cpu_kernels):
    import flowfusion_b200.flow as F
    meta, sd, ins, outs = load_golden("adjoint_flow_sample")
    m = F.ODEFlow(**meta["ctor"]).train()
    m.load_state_dict(sd)
    xT = ins["xT"].clone().requires_grad_(True)
    x = m.sample(xT, gradients=True)
    assert _close(outs["x"], x.detach(), 1e-4)
    (x * ins["w"]).sum().backward()
    assert _close(outs["grad/xT"], xT.grad, GRAD_TOL)
    for k, p in m.named_parameters():
        assert _close(outs["grad/" + k], p.grad, GRAD_TOL), k
    with pytest.raises(NotImplementedError):
        m.solve_ode_forward(ins["xT"], adjoint=True)


def test_conditional_flow_sampler_adjoint_host_logic(cpu_kernels):
    """`flow.py:775-785`: the conditional rides in the ODE state, takes part in the norms and gets its own gradient."""
    import flowfusion_b200.flow as F
    meta, sd, ins, outs = load_golden("adjoint_cflow_sample")
    m = F.ConditionalODEFlow(**meta["ctor"]).train()
    m.load_state_dict(sd)
    xT = ins["xT"].clone().requires_grad_(True)
    c = ins["cond"].clone().requires_grad_(True)
    x = m.sample(xT, c, gradients=True)
    assert _close(outs["x"], x.detach(), 1e-4)
    (x * ins["w"]).sum().backward()
    assert _close(outs["grad/xT"], xT.grad, GRAD_TOL)
    assert _close(outs["grad/cond"], c.grad, GRAD_TOL)
    for k, p in m.named_parameters():
        assert _close(outs["grad/" + k], p.grad, GRAD_TOL), k
    with pytest.raises(NotImplementedError):
        m.log_prob(ins["xT"], ins["cond"], adjoint=True)
