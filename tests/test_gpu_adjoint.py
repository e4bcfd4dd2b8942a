"""GPU tests of gradients through the sampler (flowfusion_b200/adjoint.py on the fused vector-Jacobian kernel of
csrc/ffb_train.cu; run with `pytest -m gpu`): ``ScoreModel.sample_ode_from_base`` in training mode and
``ODEFlow.sample(gradients=True)`` against the golden vectors of the UNMODIFIED reference on the restated ``odeint_adjoint``."""
import pytest
import torch

from conftest import load_golden
from test_training import _close, cpu_train_step

pytestmark = pytest.mark.gpu

GRAD_TOL = 2e-3      # of max(1, |grad|_inf): two adaptive solves (FP32 CPU / GPU) whose step sequences may differ by rounding


@pytest.mark.parametrize("name", ["adjoint_vp_pfode", "adjoint_ve_sigma_pfode"])
def test_score_sampler_adjoint_matches_reference_golden(cuda_dev, name):
    import flowfusion_b200.diffusion as D
    meta, sd, ins, outs = load_golden(name)
    sde = {"vp": D.VPSDE, "ve": D.VESDE}[meta["sde"]]()
    sm = D.ScoreModel(D.MLP(**meta["ctor"]), sde, no_sigma=meta["no_sigma"]).train()
    sm.load_state_dict(sd)
    sm.to(cuda_dev)
    base = ins["base"].to(cuda_dev).requires_grad_(True)
    cond = ins["cond"].to(cuda_dev) if "cond" in ins else None
    opts = None if meta["step_t"] is None else {"step_t": torch.tensor([meta["step_t"]])}
    x, aux = sm.sample_ode_from_base(base, cond, atol=meta["tol"], rtol=meta["tol"], options=opts)
    assert aux == [] and x.requires_grad and x.is_cuda
    assert _close(outs["x"], x.detach().cpu(), 1e-4)
    assert (sm.last_stats.accepted, sm.last_stats.rejected) == (meta["stats_forward"]["accepted"], meta["stats_forward"]["rejected"])
    (x * ins["w"].to(cuda_dev)).sum().backward()
    assert _close(outs["grad/base"], base.grad.cpu(), GRAD_TOL)
    for k, p in sm.named_parameters():
        if p.requires_grad:
            assert _close(outs["grad/" + k], p.grad.cpu(), GRAD_TOL), k
    assert sm.model.W.grad is None
    assert abs(sm.last_adjoint_stats.accepted - meta["stats_backward"]["accepted"]) <= 3


def test_flow_sampler_adjoint_matches_reference_golden(cuda_dev):
    import flowfusion_b200.flow as F
    meta, sd, ins, outs = load_golden("adjoint_flow_sample")
    m = F.ODEFlow(**meta["ctor"]).train()
    m.load_state_dict(sd)
    m.to(cuda_dev)
    xT = ins["xT"].to(cuda_dev).requires_grad_(True)
    x = m.sample(xT, gradients=True)
    assert _close(outs["x"], x.detach().cpu(), 1e-4)
    (x * ins["w"].to(cuda_dev)).sum().backward()
    assert _close(outs["grad/xT"], xT.grad.cpu(), GRAD_TOL)
    for k, p in m.named_parameters():
        assert _close(outs["grad/" + k], p.grad.cpu(), GRAD_TOL), k


def test_conditional_flow_sampler_adjoint_matches_reference_golden(cuda_dev):
    import flowfusion_b200.flow as F
    meta, sd, ins, outs = load_golden("adjoint_cflow_sample")
    m = F.ConditionalODEFlow(**meta["ctor"]).train()
    m.load_state_dict(sd)
    m.to(cuda_dev)
    xT = ins["xT"].to(cuda_dev).requires_grad_(True)
    c = ins["cond"].to(cuda_dev).requires_grad_(True)
    x = m.sample(xT, c, gradients=True)
    assert _close(outs["x"], x.detach().cpu(), 1e-4)
    (x * ins["w"].to(cuda_dev)).sum().backward()
    assert _close(outs["grad/xT"], xT.grad.cpu(), GRAD_TOL)
    assert _close(outs["grad/cond"], c.grad.cpu(), GRAD_TOL)
    for k, p in m.named_parameters():
        assert _close(outs["grad/" + k], p.grad.cpu(), GRAD_TOL), k


def test_vjp_mode_of_the_training_kernel(cuda_dev):
    """cot != NULL: grad_w / grad_b / grad_x are scale * cot^T d net / d (W, b, X), `out` is the network output."""
    import copy
    from flowfusion_b200 import training
    torch.manual_seed(4)
    lin = [torch.nn.Linear(28, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 16)]
    x = torch.randn(333, 28); cot = torch.randn(333, 16)
    rl, rg, rx, ro = cpu_train_step(lin, 0, x, None, None, -0.7, want_grad_x=True, cot=cot, want_out=True)
    gl = [copy.deepcopy(l).to(cuda_dev) for l in lin]
    flat = torch.empty(sum(training.param_sizes(gl)), device=cuda_dev)
    loss, grads, gx, out = training.train_step(gl, 0, x.to(cuda_dev), None, None, -0.7, want_grad_x=True, cot=cot.to(cuda_dev),
                                               want_out=True, grad_flat=flat)
    assert _close(ro, out.cpu(), 1e-5) and _close(rx, gx.cpu(), 5e-5)
    assert abs(float(loss) - float(rl)) <= 1e-5 * max(1.0, abs(float(rl)))
    for a, b in zip(rg, grads):
        assert _close(a, b.cpu(), 5e-5)
    assert grads[0].data_ptr() == flat.data_ptr()          # the gradients live in the caller's flat vector


def test_cfg2_network_adjoint_vs_oracle(cuda_dev):
    """The cfg2 architecture (4 x 128, 16-D, 4 conditionals): dL/d(base, weights) of a 64-row training-mode solve against the
    oracle port driven through the restated odeint_adjoint."""
    import flowfusion_b200.diffusion as D
    from oracle import port
    import torchdiffeq as tde
    torch.manual_seed(1234)
    sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 4), D.VPSDE(), no_sigma=True).train()
    base = torch.randn(64, 16, generator=torch.Generator().manual_seed(2))
    cond = torch.randn(64, 4, generator=torch.Generator().manual_seed(3))
    w = torch.randn(64, 16, generator=torch.Generator().manual_seed(4))
    opts = {"step_t": torch.tensor([1e-3])}
    # oracle: the port's drift as an nn.Module over cloned weights
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)

    class Drift(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in M["P"]["net"]["w"]])
            self.b = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in M["P"]["net"]["b"]])

        def forward(self, t, y):
            M["P"]["net"]["w"], M["P"]["net"]["b"] = list(self.w), list(self.b)
            return port.ode_drift(M, t, y, cond)

    dr = Drift()
    b0 = base.clone().requires_grad_(True)
    xr = tde.odeint_adjoint(dr, b0, torch.tensor([1.0, 1e-3]), rtol=1e-5, atol=1e-5, options=opts)[-1]
    (xr * w).sum().backward()
    sm.to(cuda_dev)
    bg = base.to(cuda_dev).requires_grad_(True)
    x, _ = sm.sample_ode_from_base(bg, cond.to(cuda_dev), atol=1e-5, rtol=1e-5, options=opts)
    (x * w.to(cuda_dev)).sum().backward()
    assert _close(xr.detach(), x.detach().cpu(), 1e-4)
    assert _close(b0.grad, bg.grad.cpu(), GRAD_TOL)
    for i, lin in enumerate(sm.model.NN):
        assert _close(dr.w[i].grad, lin.weight.grad.cpu(), GRAD_TOL), i
        assert _close(dr.b[i].grad, lin.bias.grad.cpu(), GRAD_TOL), i
