"""Host logic on CPU: the package's classes, scalar programs and dopri5 controller driven
through the torch-CPU kernel model (tests/kernel_model.py) must reproduce the golden vectors
that the UNMODIFIED reference produced (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_row_err
from kernel_model import patched_engine

import flowfusion_b200.diffusion as D
import flowfusion_b200.flow as F
import flowfusion_b200.symplectic as Sy

TOL = 2e-5   # the model and the oracle differ only by FP32 summation order


def check_stats(stats, meta_stats):
    assert (stats.accepted, stats.rejected) == (meta_stats["accepted"], meta_stats["rejected"])
    assert stats.nfe == meta_stats["nfe"]


def test_cfg1_flow_sample():
    meta, sd, ins, outs = load_golden("cfg1_flow_sample")
    m = F.ODEFlow(**meta["ctor"]).eval()
    m.load_state_dict(sd)
    with patched_engine():
        x = m.sample(ins["xT"])
    assert rel_row_err(outs["x"], x) < 1e-4
    # rtol 1e-7 is below FP32 resolution: the step count is rounding noise, do not pin it


def test_cfg3_flow_logprob():
    meta, sd, ins, outs = load_golden("cfg3_flow_logprob")
    m = F.ODEFlow(**meta["ctor"], target_shift=sd["target_shift"], target_scale=sd["target_scale"]).eval()
    m.load_state_dict(sd)
    with patched_engine():
        lp = m.log_prob(ins["x"])
    assert lp.shape == outs["log_prob"].shape
    assert float((lp - outs["log_prob"]).abs().max()) < 1e-3
    check_stats(m.last_stats, meta["stats"])


def test_conditional_flow():
    meta, sd, ins, outs = load_golden("cflow_sample_logprob")
    m = F.ConditionalODEFlow(**meta["ctor"]).eval()
    m.load_state_dict(sd)
    with patched_engine():
        x = m.sample(ins["xT"], ins["cond"])
        assert rel_row_err(outs["x"], x) < 1e-4
        lp = m.log_prob(outs["x"], ins["cond"], atol=1e-6, rtol=1e-6)
    assert float((lp - outs["log_prob"]).abs().max()) < 1e-3
    check_stats(m.last_stats, meta["stats_logprob"])


def _score_model(meta, sd):
    net = D.MLP(**meta["ctor"])
    sde = {"vp": D.VPSDE, "ve": D.VESDE, "subvp": D.SUBVPSDE}[meta["sde"]]()
    sm = D.ScoreModel(net, sde, no_sigma=meta["no_sigma"]).eval()
    sm.load_state_dict(sd)
    return sm


def test_cfg2_pfode_all_methods():
    meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
    sm = _score_model(meta, sd)
    with patched_engine():
        x, aux = sm.sample_ode_from_base(ins["base"], ins["cond"], atol=1e-5, rtol=1e-5,
                                         options={"step_t": torch.tensor([1e-3])})
        assert aux == []
        assert rel_row_err(outs["x_dopri5"], x) < 1e-4
        check_stats(sm.last_stats, meta["stats"])
        x4, _ = sm.sample_ode_from_base(ins["base"], ins["cond"], method="rk4", options={"step_size": 1 / 64})
        assert rel_row_err(outs["x_rk4"], x4) < TOL
        xe, _ = sm.sample_ode_from_base(ins["base"], ins["cond"], method="euler", options={"step_size": 1 / 128})
        assert rel_row_err(outs["x_euler"], xe) < TOL


@pytest.mark.parametrize("kind", ["ve", "subvp", "vp"])
def test_sigma_pfode(kind):
    meta, sd, ins, outs = load_golden(f"{kind}_sigma_pfode")
    sm = _score_model(meta, sd)
    opts = None if meta["call"]["step_t"] is None else {"step_t": torch.tensor([meta["call"]["step_t"]])}
    with patched_engine():
        x, _ = sm.sample_ode_from_base(ins["base"], atol=1e-5, rtol=1e-5, options=opts)
    assert rel_row_err(outs["x_dopri5"], x) < 1e-4
    check_stats(sm.last_stats, meta["stats"])


def test_score_logprob_exact_and_hutch():
    meta, sd, ins, outs = load_golden("score_logprob_vp")
    sm = _score_model(meta, sd)
    with patched_engine():
        lp = sm.log_prob(ins["x0"], ins["cond"])
        assert lp.shape == outs["lp_exact"].shape == (ins["x0"].shape[0], 1)
        assert float((lp - outs["lp_exact"]).abs().max()) < 1e-3
        check_stats(sm.last_stats, meta["stats"])
        sm.hutch = True
        lph = sm.log_prob(ins["x0"], ins["cond"], probes=ins["probes"])
        assert float((lph - outs["lp_hutch"]).abs().max()) < 1e-3
        check_stats(sm.last_stats, meta["stats_hutch"])


def test_score_logprob_ve():
    meta, sd, ins, outs = load_golden("score_logprob_ve")
    sm = _score_model(meta, sd)
    with patched_engine():
        lp = sm.log_prob(ins["x0"])
    assert float((lp - outs["lp_exact"]).abs().max()) < 1e-3
    check_stats(sm.last_stats, meta["stats"])


def _replay_em_noise(seed, B, D, steps, scale=1.0):
    torch.manual_seed(seed)
    x0 = torch.distributions.Normal(torch.zeros(D), scale).sample([B])
    dw = torch.stack([torch.randn_like(x0) for _ in range(steps)])
    return x0, dw


def test_cfg4_euler_maruyama():
    meta, sd, ins, outs = load_golden("cfg4_vp_em")
    sm = _score_model(meta, sd)
    for run in meta["runs"]:
        x0, dw = _replay_em_noise(run["seed"], run["B"], 32, run["steps"])
        with patched_engine():
            x = sm.sample_sde((run["B"], 32), steps=run["steps"], x0=x0, noise=dw)
        assert rel_row_err(outs[f"x_{run['steps']}"], x) < TOL


def test_em_stops_at_first_nan_like_the_reference(capsys):
    """`diffusion.py:560-563`: the first step that leaves a NaN anywhere in x stops the WHOLE batch and that step's x_mean
    is returned.  A NaN planted in the caller's noise at (step 5, row 7) makes the step deterministic."""
    from oracle import port
    torch.manual_seed(11)
    sm = D.ScoreModel(D.MLP(5, 0, 8, [32, 32]), D.VPSDE(), no_sigma=True).eval()
    x0 = torch.randn(40, 5, generator=torch.Generator().manual_seed(1))
    dw = torch.randn(20, 40, 5, generator=torch.Generator().manual_seed(2))
    dw[5, 7, 2] = float("nan")
    ref = port.sample_sde(port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True), x0, dw)
    assert torch.isfinite(ref).all()                                     # x_mean of step 5: the NaN has not entered it yet
    with patched_engine():
        x = sm.sample_sde((40, 5), steps=20, x0=x0, noise=dw)
    assert rel_row_err(ref, x) < TOL
    assert sm.stopped_at_step == 5 and not sm.check_stability()
    assert "Diffusion is not stable, NaN were produced. Stopped sampling." in capsys.readouterr().out
    with patched_engine():
        with pytest.raises(UnboundLocalError):                           # T < epsilon: the reference returns an unbound x_mean
            D.ScoreModel(D.MLP(5, 0, 8, [32]), D.VPSDE(T=0.5, epsilon=0.9), no_sigma=True).eval().sample_sde((4, 5), steps=3, x0=x0[:4], noise=dw[:3, :4])


def test_em_ve_conditional():
    meta, sd, ins, outs = load_golden("ve_em_cond")
    sm = _score_model(meta, sd)
    with patched_engine():
        x = sm.sample_sde((96, 3), conditional=ins["cond"], steps=50, x0=ins["x0"], noise=ins["dw"])
    assert rel_row_err(outs["x"], x) < TOL


@pytest.mark.parametrize("name", ["cfg5_symplectic", "symplectic_cond"])
def test_symplectic(name):
    meta, sd, ins, outs = load_golden(name)
    net = Sy.SymplecticMLP(**meta["ctor"])
    m = Sy.SymplecticFlowModel(net, sd["shift"], sd["scale"], sd["conditional_shift"], sd["conditional_scale"]).eval()
    m.load_state_dict(sd)
    cond = ins.get("cond")
    D_ = meta["ctor"]["n_data_dims"]
    with patched_engine():
        x = m.sample((ins["z0"].shape[0], D_), conditional=cond, num_steps=meta["num_steps"], z0=ins["z0"])
        assert rel_row_err(outs["x_sample"], x) < TOL
        lp = m.log_prob(ins["x"], conditional=cond, p0=ins["p0"])
    assert lp.shape == outs["log_prob"].shape
    assert float((lp - outs["log_prob"]).abs().max()) < 1e-3
    check_stats(m.last_stats, meta["stats_logprob"])


@pytest.mark.parametrize("act_cls,act_fn", [(torch.nn.Tanh, torch.tanh), (torch.nn.ReLU, torch.relu),
                                            (torch.nn.Softplus, torch.nn.functional.softplus),
                                            (torch.nn.GELU, torch.nn.functional.gelu)])
def test_non_silu_activations_reach_the_kernels(act_cls, act_fn):
    """The activation code travels from the model object to the packed network; unknown ones are refused."""
    from oracle import port
    torch.manual_seed(3)
    m = F.ODEFlow(4, [24, 24], activation=act_cls).eval()
    x = torch.randn(40, 4, generator=torch.Generator().manual_seed(2))
    ref = port.flow_log_prob(port.flow_from_state_dict(m.state_dict(), act=act_fn), x)
    with patched_engine():
        lp = m.log_prob(x)
        assert m._net().act is {torch.nn.Tanh: torch.tanh, torch.nn.ReLU: torch.relu}.get(act_cls, m._net().act)
    assert float((lp - ref).abs().max()) < 1e-3
    with pytest.raises(NotImplementedError):
        with patched_engine():
            F.ODEFlow(4, [8], activation=torch.nn.Sigmoid).eval().log_prob(x)
    with pytest.raises(NotImplementedError):
        with patched_engine():
            D.ScoreModel(D.MLP(4, 0, 4, [8], activation=torch.nn.Softplus(beta=2.0)), D.VESDE()).eval().sample_ode_from_base(x)


@pytest.mark.parametrize("method", ["bosh3", "adaptive_heun", "fehlberg2"])
def test_other_adaptive_methods_vs_oracle(method):
    """`method=` is a pass-through argument of the reference's entry points (`diffusion.py:572-573`, `flow.py:313-314`): the
    host controller with the method's tableau / order against the restated torchdiffeq, identical accept / reject
    sequences: PF-ODE sampling (VE with the sigma division), exact-trace log-prob of a flow, conditional flow (the
    conditional rides in the error norm)."""
    from oracle import port
    tol = dict(bosh3=1e-4, adaptive_heun=1e-3, fehlberg2=1e-4)[method]
    torch.manual_seed(21)
    sm = D.ScoreModel(D.MLP(5, 2, 8, [32, 32]), D.VESDE(), no_sigma=False).eval()
    base = torch.randn(60, 5, generator=torch.Generator().manual_seed(1)); cond = torch.randn(60, 2, generator=torch.Generator().manual_seed(2))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("ve"), False)
    ref = port.sample_ode_from_base(M, base, cond, tol, tol, method=method)[0]
    rs = port.last_stats()
    with patched_engine():
        x, _ = sm.sample_ode_from_base(base, cond, atol=tol, rtol=tol, method=method)
    assert rel_row_err(ref, x) < TOL
    assert (sm.last_stats.accepted, sm.last_stats.rejected, sm.last_stats.nfe) == (rs.accepted, rs.rejected, rs.nfe)
    assert sm.last_stats.accept_history == rs.accept_history and sm.last_stats.method == method
    torch.manual_seed(22)
    fl = F.ConditionalODEFlow(4, 2, [32, 32]).eval()
    xs = torch.randn(50, 4, generator=torch.Generator().manual_seed(3)); c = torch.randn(50, 2, generator=torch.Generator().manual_seed(4))
    ref_lp = port.flow_log_prob(port.flow_from_state_dict(fl.state_dict()), xs, c, atol=tol, rtol=tol, method=method)
    rs = port.last_stats()
    with patched_engine():
        lp = fl.log_prob(xs, c, atol=tol, rtol=tol, method=method)
    assert float((lp - ref_lp).abs().max()) < 1e-4
    assert (fl.last_stats.accepted, fl.last_stats.rejected) == (rs.accepted, rs.rejected)


@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_fixed_grid_options_perturb_and_grid_constructor(method):
    """options={'perturb': True} and options={'grid_constructor': ...} of torchdiffeq's fixed-grid solvers (pass-through
    `options=` of the reference's entry points) against the restated torchdiffeq."""
    from oracle import port
    torch.manual_seed(31)
    sm = D.ScoreModel(D.MLP(4, 0, 8, [32]), D.VPSDE(), no_sigma=True).eval()
    base = torch.randn(30, 4, generator=torch.Generator().manual_seed(1))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    gc = lambda func, y0, t: torch.tensor([float(t[0]), -0.8, -0.55, -0.3, -0.07, float(t[-1])])      # noqa: E731  (reverse time: t is negated)
    for opts in ({"step_size": 0.125, "perturb": True}, {"grid_constructor": gc}, {"grid_constructor": gc, "perturb": True}):
        ref = port.sample_ode_from_base(M, base, None, method=method, options=opts)[0]
        with patched_engine():
            x, _ = sm.sample_ode_from_base(base, method=method, options=opts)
        assert rel_row_err(ref, x) < TOL, opts
    plain = port.sample_ode_from_base(M, base, None, method=method, options={"step_size": 0.125})[0]
    pert = port.sample_ode_from_base(M, base, None, method=method, options={"step_size": 0.125, "perturb": True})[0]
    assert not torch.equal(plain, pert)                      # the option does change the evaluation times


def test_dopri5_jump_t():
    """options={'jump_t': ...}: steps land on the discontinuity and f is evaluated again just after it (one extra NFE per
    jump), against the restated torchdiffeq: identical step sequence and NFE."""
    from oracle import port
    torch.manual_seed(33)
    sm = D.ScoreModel(D.MLP(4, 1, 8, [32, 32]), D.VPSDE(), no_sigma=True).eval()
    base = torch.randn(40, 4, generator=torch.Generator().manual_seed(1)); cond = torch.randn(40, 1, generator=torch.Generator().manual_seed(2))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    opts = {"step_t": torch.tensor([1e-3]), "jump_t": torch.tensor([0.6, 0.25])}
    ref = port.sample_ode_from_base(M, base, cond, 1e-5, 1e-5, options=opts)[0]
    rs = port.last_stats()
    plain = port.sample_ode_from_base(M, base, cond, 1e-5, 1e-5, options={"step_t": torch.tensor([1e-3])})[0]
    ps = port.last_stats()
    assert rs.nfe != ps.nfe and not torch.equal(ref, plain)
    with patched_engine():
        x, _ = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options=opts)
    assert rel_row_err(ref, x) < TOL
    assert (sm.last_stats.accepted, sm.last_stats.rejected, sm.last_stats.nfe) == (rs.accepted, rs.rejected, rs.nfe)
    assert sm.last_stats.controller == "host"
