"""GPU tests of the fused training step (csrc/ffb_train.cu; run with `pytest -m gpu`): loss and every parameter gradient against
the golden vectors of the UNMODIFIED reference (oracle/make_golden_train.py) and the CPU oracle, through the reference's own
entry points (denoising_score_matching, log_prob_score_matching, ScoreModel.loss_fn, flow_matching_loss) and autograd."""
import copy

import pytest
import torch

from conftest import load_golden
from test_training import ACT, ACT_MOD, SCORE_CASES, FLOW_CASES, _close, cpu_train_step

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-5        # relative
GRAD_TOL = 5e-5        # of max(1, |grad|_inf): FP32 sums over the batch in a different order


@pytest.mark.parametrize("name", SCORE_CASES)
def test_score_losses_match_reference_golden(cuda_dev, name):
    import flowfusion_b200.diffusion as D
    meta, sd, ins, outs = load_golden(name)
    kw = {} if meta["activation"] is None else {"activation": ACT_MOD[meta["activation"]]()}
    sde = {"vp": D.VPSDE, "ve": D.VESDE, "subvp": D.SUBVPSDE}[meta["sde"]]()
    sm = D.ScoreModel(D.MLP(**meta["ctor"], **kw), sde, no_sigma=meta["no_sigma"]).train()
    sm.load_state_dict(sd)
    sm.to(cuda_dev)
    g = {k: v.to(cuda_dev) for k, v in ins.items()}
    fn = D.denoising_score_matching if meta["loss"] == "dsm" else D.log_prob_score_matching
    loss = fn(sm, g["x"], conditional=g.get("cond"), z=g["z"], t=g["t"])
    assert loss.dim() == 0 and loss.is_cuda and loss.requires_grad
    assert _close(outs["loss"], loss.detach().cpu(), LOSS_TOL)
    loss.backward()
    for k, p in sm.named_parameters():
        if p.requires_grad:
            assert _close(outs["grad/" + k], p.grad.cpu(), GRAD_TOL), k
    # same call, same bits (no atomics anywhere)
    sm.zero_grad()
    loss2 = fn(sm, g["x"], conditional=g.get("cond"), z=g["z"], t=g["t"])
    loss2.backward()
    assert torch.equal(loss, loss2)
    # the reference's internal draws: finite, right scale
    lf = sm.loss_fn(g["x"], conditional=g.get("cond")) if meta["loss"] == "dsm" else fn(sm, g["x"], conditional=g.get("cond"))
    assert torch.isfinite(lf) and float(lf.detach()) > 0


@pytest.mark.parametrize("name", FLOW_CASES)
def test_flow_matching_loss_matches_reference_golden(cuda_dev, name):
    import flowfusion_b200.flow as F
    meta, sd, ins, outs = load_golden(name)
    kw = {} if meta.get("activation") is None else {"activation": ACT_MOD[meta["activation"]]}
    cls = F.ConditionalODEFlow if meta["case"] == "cflow_loss" else F.ODEFlow
    m = cls(**meta["ctor"], **kw).train()
    m.load_state_dict(sd)
    m.to(cuda_dev)
    g = {k: v.to(cuda_dev) for k, v in ins.items()}
    args = (g["x"],) + ((g["cond"],) if "cond" in g else ())
    loss = m.flow_matching_loss(*args, xT=g["xT"], t=g["t"])
    assert _close(outs["loss"], loss.detach().cpu(), LOSS_TOL)
    loss.backward()
    for k, p in m.named_parameters():
        assert _close(outs["grad/" + k], p.grad.cpu(), GRAD_TOL), k


@pytest.mark.parametrize("units,B,act", [([256] * 3, 1000, 0), ([128] * 8, 77, 0), ([48], 1, 1), ([200, 72], 4097, 3), ([], 300, 0)])
def test_train_step_shapes_vs_cpu_model(cuda_dev, units, B, act):
    """The raw fused call on wide / deep / single-layer networks and ragged batches, d loss / d X included, against autograd."""
    from flowfusion_b200 import training
    torch.manual_seed(len(units) + B)
    dims = [21] + units + [9]
    lin = [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
    x = torch.randn(B, 21); alpha = torch.rand(B) + 0.5; beta = torch.randn(B, 9)
    rl, rg, rx = cpu_train_step(lin, act, x, alpha, beta, 0.37, want_grad_x=True)
    gl = [copy.deepcopy(l).to(cuda_dev) for l in lin]
    loss, grads, gx = training.train_step(gl, act, x.to(cuda_dev), alpha.to(cuda_dev), beta.to(cuda_dev), 0.37, want_grad_x=True)
    assert abs(float(loss) - float(rl)) <= LOSS_TOL * max(1.0, abs(float(rl)))
    for a, b in zip(rg, grads):
        assert _close(a, b.cpu(), GRAD_TOL)
    assert _close(rx, gx.cpu(), GRAD_TOL)
    # alpha = None means 1
    rl1, rg1, _ = cpu_train_step(lin, act, x, None, beta, 1.0)
    loss1, grads1, gx1 = training.train_step(gl, act, x.to(cuda_dev), None, beta.to(cuda_dev), 1.0)
    assert gx1 is None and abs(float(loss1) - float(rl1)) <= LOSS_TOL * max(1.0, abs(float(rl1)))
    assert _close(rg1[0], grads1[0].cpu(), GRAD_TOL)


def test_training_loop_follows_the_oracle(cuda_dev):
    """Five SGD steps on the GPU (fused loss + autograd + torch.optim.SGD) against the same steps on the CPU oracle."""
    import flowfusion_b200.flow as F
    from oracle import port
    torch.manual_seed(9)
    m = F.ODEFlow(6, [64, 64], target_shift=torch.zeros(6) + 0.1, target_scale=torch.ones(6) * 1.3).train()
    Fl = port.flow_from_state_dict(m.state_dict())
    m.to(cuda_dev)
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    gen = torch.Generator().manual_seed(5)
    losses_gpu, losses_cpu = [], []
    for step in range(5):
        x = torch.randn(512, 6, generator=gen); xT = torch.randn(512, 6, generator=gen); t = torch.rand(512, 1, generator=gen)
        opt.zero_grad()
        loss = m.flow_matching_loss(x.to(cuda_dev), xT=xT.to(cuda_dev), t=t.to(cuda_dev))
        loss.backward()
        opt.step()
        losses_gpu.append(float(loss.detach()))
        pl, pg = port.loss_and_grads(port.fm_loss, Fl["net"], Fl, x, xT, t)
        Fl["net"]["w"] = [w - 0.05 * pg[2 * i] for i, w in enumerate(Fl["net"]["w"])]
        Fl["net"]["b"] = [b - 0.05 * pg[2 * i + 1] for i, b in enumerate(Fl["net"]["b"])]
        losses_cpu.append(float(pl))
    assert max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(losses_gpu, losses_cpu)) < 1e-4
    lin = [l for l in m.layers if isinstance(l, torch.nn.Linear)]
    for i, l in enumerate(lin):
        assert _close(Fl["net"]["w"][i], l.weight.detach().cpu(), 1e-4)
    # the trained weights are what the sampling kernels see next (re-packed on change)
    m.eval()
    xs = m.sample(torch.randn(64, 6, generator=gen).to(cuda_dev))
    assert torch.isfinite(xs).all()


def test_training_limits(cuda_dev):
    from flowfusion_b200 import training, _lib
    lin = [torch.nn.Linear(8, 512), torch.nn.Linear(512, 512), torch.nn.Linear(512, 4)]
    with pytest.raises(_lib.FFBError):                 # no CPU path
        training.train_step(lin, 0, torch.zeros(10, 8), None, torch.zeros(10, 4), 1.0)
    lin = [l.to(cuda_dev) for l in lin]
    with pytest.raises(_lib.FFBError, match="shared memory"):
        training.train_step(lin, 0, torch.zeros(10, 8, device=cuda_dev), None, torch.zeros(10, 4, device=cuda_dev), 1.0)


@pytest.mark.parametrize("D_,C_,units,act_cls,act_fn", [(32, 0, [128] * 4, torch.nn.SiLU, None), (3, 2, [48, 40], torch.nn.Tanh, torch.tanh),
                                                       (1, 0, [16], torch.nn.Softplus, torch.nn.functional.softplus)])
def test_hamiltonian_leapfrog_fused_gradient(cuda_dev, D_, C_, units, act_cls, act_fn):
    """BASELINE.json configs[4] / north_star: leapfrog whose dH/dq, dH/dp is a fused forward+backward MLP kernel (extension; the
    oracle is autograd on the CPU): trajectories, energies, time reversibility and O(dt^2) energy drift."""
    import flowfusion_b200.symplectic as Sy
    from oracle import port
    torch.manual_seed(70 + D_)
    m = Sy.HamiltonianMLP(D_, C_, units, activation=act_cls())
    net = port.net_from_state_dict(m.state_dict(), "net.", 2, act=act_fn)
    B = 203                                           # ragged: 6 passes of 32 rows + 11
    z0 = torch.randn(B, 2 * D_, generator=torch.Generator().manual_seed(1))
    c = torch.randn(B, C_, generator=torch.Generator().manual_seed(2)) if C_ else None
    zr, h0r, h1r = port.hamiltonian_leapfrog(net, z0, c, n_steps=20, dt=0.02)
    m.to(cuda_dev)
    cg = c.to(cuda_dev) if c is not None else None
    z, h = m.leapfrog(z0.to(cuda_dev), cg, num_steps=20, dt=0.02, return_energy=True)
    assert _close(zr, z.cpu(), 1e-4)
    assert _close(h0r, h[:, 0].cpu(), 1e-5) and _close(h1r, h[:, 1].cpu(), 1e-4)
    assert _close(h0r, m.energy(z0.to(cuda_dev), cg).cpu(), 1e-5)
    # time reversibility
    zb = m.leapfrog(z, cg, num_steps=20, dt=-0.02)             # kick-drift-kick is symmetric: a negative step retraces it
    assert _close(z0, zb.cpu(), 2e-4)
    # symplectic integrator: the energy error of a step halves ~4x when dt halves (2nd order), and stays bounded
    _, ha = m.leapfrog(z0.to(cuda_dev), cg, num_steps=10, dt=0.04, return_energy=True)
    _, hb = m.leapfrog(z0.to(cuda_dev), cg, num_steps=20, dt=0.02, return_energy=True)
    ea, eb = (ha[:, 1] - ha[:, 0]).abs().mean(), (hb[:, 1] - hb[:, 0]).abs().mean()
    assert float(eb) < 0.5 * float(ea) + 1e-5
    # zero steps: the state comes back untouched
    assert torch.equal(m.leapfrog(z0.to(cuda_dev), cg, num_steps=0).cpu(), z0)


def test_per_sample_times_forward_only(cuda_dev):
    """`MLP.forward` / `ScoreModel.score` / `SymplecticMLP.forward` with ONE TIME PER SAMPLE (`diffusion.py:104-113`,
    `symplectic.py:101-114`): the forward-only mode of the fused kernel, against the oracle; a 512-wide network fits there."""
    import flowfusion_b200.diffusion as D
    import flowfusion_b200.symplectic as Sy
    from oracle import port
    torch.manual_seed(12)
    for units in ([128, 128], [512, 512, 512]):
        sm = D.ScoreModel(D.MLP(6, 2, 8, units), D.VPSDE(), no_sigma=False).eval()
        g = torch.Generator().manual_seed(1)
        x, c, t = torch.randn(301, 6, generator=g), torch.randn(301, 2, generator=g), torch.rand(301, generator=g) * 0.9 + 0.05
        M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), False)
        ref_net, ref_score = port.score_net(M["P"], t, x, c), port.score(M, t, x, c)
        sm.to(cuda_dev)
        out = sm.model(t.to(cuda_dev), x.to(cuda_dev), conditional=c.to(cuda_dev))
        assert _close(ref_net, out.cpu(), 2e-5)
        sc = sm.score(t.to(cuda_dev), x.to(cuda_dev), conditional=c.to(cuda_dev))
        assert _close(ref_score, sc.cpu(), 2e-5)
    net = Sy.SymplecticMLP(4, 1, 8, [64, 64])
    z = torch.randn(77, 8); c = torch.randn(77, 1); t = torch.rand(77)
    Syp = port.symplectic_from_state_dict({"model." + k: v for k, v in net.state_dict().items()} |
                                          {"shift": torch.zeros(4), "scale": torch.ones(4), "conditional_shift": torch.zeros(1),
                                           "conditional_scale": torch.ones(1)})
    q, p = z[:, :4], z[:, 4:]
    temb = port.fourier_features(t, Syp["W"], torch.tensor(3.141592653589793))
    ref = torch.cat([port._mlp(Syp["net_q"], torch.cat([p, c, temb], 1)), -port._mlp(Syp["net_p"], torch.cat([q, c, temb], 1))], 1)
    net.to(cuda_dev)
    v = net(t.to(cuda_dev), z.to(cuda_dev), c.to(cuda_dev))
    assert _close(ref, v.cpu(), 2e-5)
