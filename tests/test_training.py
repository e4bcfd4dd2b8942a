"""CPU tests of the training-side losses (no GPU needed):
  (a) the oracle's restatement (oracle/port.py: dsm_loss / lpsm_loss / fm_loss) against the golden vectors that
      oracle/make_golden_train.py took from the UNMODIFIED reference (loss and every parameter gradient);
  (b) the host logic of flowfusion_b200/training.py -- the reference's prologue, the (alpha, beta, scale) form of each loss,
      the autograd hook -- with the fused CUDA call replaced by a torch-CPU model of it, against the same golden vectors."""
import pytest
import torch

from conftest import load_golden

ACT = {None: None, "Tanh": torch.tanh, "GELU": torch.nn.functional.gelu}
ACT_MOD = {None: None, "Tanh": torch.nn.Tanh, "GELU": torch.nn.GELU}
SCORE_CASES = ["train_dsm_vp", "train_lpsm_ve", "train_dsm_subvp_tanh"]
FLOW_CASES = ["train_fm_flow", "train_fm_cflow_gelu"]


def _close(ref, got, tol):
    return float((ref - got).abs().max()) <= tol * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("name", SCORE_CASES + FLOW_CASES)
def test_oracle_losses_match_reference_golden(name):
    from oracle import port
    meta, sd, ins, outs = load_golden(name)
    act = ACT[meta.get("activation")]
    if meta["case"] == "score_loss":
        M = port.score_model_from_state_dict(sd, port.make_sde(meta["sde"]), meta["no_sigma"], act=act)
        fn = port.dsm_loss if meta["loss"] == "dsm" else port.lpsm_loss
        loss, grads = port.loss_and_grads(fn, M["P"]["net"], M, ins["x"], ins["z"], ins["t"], ins.get("cond"))
        keys = [f"grad/model.NN.{i}.{w}" for i in range(len(grads) // 2) for w in ("weight", "bias")]
    else:
        Fl = port.flow_from_state_dict(sd, act=act)
        loss, grads = port.loss_and_grads(port.fm_loss, Fl["net"], Fl, ins["x"], ins["xT"], ins["t"], ins.get("cond"))
        keys = [f"grad/layers.{2 * i}.{w}" for i in range(len(grads) // 2) for w in ("weight", "bias")]
    assert _close(outs["loss"], loss, 1e-6)
    for k, g in zip(keys, grads):
        assert _close(outs[k], g, 2e-5), k


def cpu_train_step(linears, activation, x_in, alpha, beta, scale, want_grad_x=False, cot=None, want_out=False, grad_flat=None):
    """torch-CPU model of ffb_train_step (csrc/ffb_train.cu): same arguments, same returns as training.train_step."""
    act = [torch.nn.functional.silu, torch.tanh, torch.relu, torch.nn.functional.softplus, torch.nn.functional.gelu][activation]
    params = []
    for lin in linears:
        params += [lin.weight.detach().clone().requires_grad_(True), lin.bias.detach().clone().requires_grad_(True)]
    h = x_in.detach().clone().requires_grad_(want_grad_x)
    x0 = h
    with torch.enable_grad():
        for i in range(len(linears)):
            h = torch.nn.functional.linear(h, params[2 * i], params[2 * i + 1])
            if i < len(linears) - 1:
                h = act(h)
        if cot is not None:
            loss = scale * torch.sum(cot.detach() * h)
        else:
            a = 1.0 if alpha is None else alpha.reshape(-1, 1)
            loss = scale * torch.sum((a * h + beta) ** 2)
        grads = torch.autograd.grad(loss, params + ([x0] if want_grad_x else []))
    pg = list(grads[: len(params)])
    if grad_flat is not None:
        pos = 0
        for i, g in enumerate(pg):
            grad_flat[pos: pos + g.numel()].copy_(g.reshape(-1))
            pg[i] = grad_flat[pos: pos + g.numel()].view_as(g)
            pos += g.numel()
    res = (loss.detach().double(), pg, (grads[-1] if want_grad_x else None))
    return res + (h.detach(),) if want_out else res


@pytest.fixture
def cpu_kernel(monkeypatch):
    from flowfusion_b200 import training, engine
    monkeypatch.setattr(training, "train_step", cpu_train_step)
    monkeypatch.setattr(engine, "require_cuda", lambda *a, **k: None)


@pytest.mark.parametrize("name", SCORE_CASES)
def test_score_losses_host_logic(cpu_kernel, name):
    import flowfusion_b200.diffusion as D
    meta, sd, ins, outs = load_golden(name)
    kw = {} if meta["activation"] is None else {"activation": ACT_MOD[meta["activation"]]()}
    sde = {"vp": D.VPSDE, "ve": D.VESDE, "subvp": D.SUBVPSDE}[meta["sde"]]()
    sm = D.ScoreModel(D.MLP(**meta["ctor"], **kw), sde, no_sigma=meta["no_sigma"]).train()
    sm.load_state_dict(sd)
    fn = D.denoising_score_matching if meta["loss"] == "dsm" else D.log_prob_score_matching
    loss = fn(sm, ins["x"], conditional=ins.get("cond"), z=ins["z"], t=ins["t"])
    assert loss.dim() == 0 and loss.requires_grad
    assert _close(outs["loss"], loss.detach(), 1e-5)
    loss.backward()                                   # what an optimiser step sees
    for k, p in sm.named_parameters():
        if p.requires_grad:
            assert _close(outs["grad/" + k], p.grad, 5e-5), k
    assert sm.model.W.grad is None                    # the embedding frequencies are not trained (diffusion.py:73-76)
    if meta["loss"] == "dsm":                         # ScoreModel.loss_fn is the same loss (diffusion.py:240-256)
        torch.manual_seed(0)
        assert sm.loss_fn(ins["x"], conditional=ins.get("cond")).dim() == 0


@pytest.mark.parametrize("name", FLOW_CASES)
def test_flow_matching_loss_host_logic(cpu_kernel, name):
    import flowfusion_b200.flow as F
    meta, sd, ins, outs = load_golden(name)
    kw = {} if meta.get("activation") is None else {"activation": ACT_MOD[meta["activation"]]}
    cls = F.ConditionalODEFlow if meta["case"] == "cflow_loss" else F.ODEFlow
    m = cls(**meta["ctor"], **kw).train()
    m.load_state_dict(sd)                              # shift / scale buffers come with the state dict
    args = (ins["x"],) + ((ins["cond"],) if "cond" in ins else ())
    loss = m.flow_matching_loss(*args, xT=ins["xT"], t=ins["t"])
    assert _close(outs["loss"], loss.detach(), 1e-5)
    loss.backward()
    seen = 0
    for k, p in m.named_parameters():
        assert _close(outs["grad/" + k], p.grad, 5e-5), k
        seen += 1
    assert seen == 2 * (len(meta["ctor"]["hidden_units"]) + 1)
    xt, v = m.compute_linear_velocity_field(ins["x"], ins["xT"], ins["t"])
    assert xt.shape == ins["x"].shape and torch.allclose(v, ins["xT"] - (ins["x"] - m.target_shift) / m.target_scale)
    with pytest.raises(TypeError):                     # the reference's two signatures: (x) and (x, conditional)
        m.flow_matching_loss(ins["x"]) if "cond" in ins else m.flow_matching_loss(ins["x"], ins["x"][:, :1])


def test_loss_scales_through_autograd(cpu_kernel):
    """The fused loss is an ordinary graph node: (3 * loss).backward() gives 3 x the gradients."""
    import flowfusion_b200.flow as F
    torch.manual_seed(3)
    m = F.ODEFlow(3, [16, 16]).train()
    x = torch.randn(50, 3); xT = torch.randn(50, 3); t = torch.rand(50, 1)
    m.flow_matching_loss(x, xT=xT, t=t).backward()
    g1 = [p.grad.clone() for p in m.parameters()]
    m.zero_grad()
    (3.0 * m.flow_matching_loss(x, xT=xT, t=t)).backward()
    for a, b in zip(g1, m.parameters()):
        assert torch.allclose(3.0 * a, b.grad, rtol=1e-6, atol=1e-7)
