"""The CPU oracle (oracle/port.py on the restated torchdiffeq, oracle/torchdiffeq) against
  (a) the golden vectors the UNMODIFIED reference produced (tests/golden, oracle/make_golden.py),
  (b) the live reference, when /root/reference is present (build container only),
  (c) known-answer tests of the restated solver the reference never had.
No GPU, no product code: this file pins the checker itself."""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_row_err
from oracle import loader, port

import os, sys  # noqa: E401
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torchdiffeq as tde  # noqa: E402  (the restatement in oracle/torchdiffeq)

TIGHT = 2e-6   # port and reference run the same FP32 ops in the same order


def stats_match(meta_stats):
    s = port.last_stats()
    assert (s.nfe, s.accepted, s.rejected) == (meta_stats["nfe"], meta_stats["accepted"], meta_stats["rejected"])


# ---------------------------------------------------------------------------------------------
# (a) golden vectors
# ---------------------------------------------------------------------------------------------
def test_golden_cfg1_flow_sample():
    meta, sd, ins, outs = load_golden("cfg1_flow_sample")
    x = port.flow_sample(port.flow_from_state_dict(sd), ins["xT"])
    assert rel_row_err(outs["x"], x) < TIGHT
    stats_match(meta["stats"])


def test_golden_cfg3_flow_logprob():
    meta, sd, ins, outs = load_golden("cfg3_flow_logprob")
    lp = port.flow_log_prob(port.flow_from_state_dict(sd), ins["x"])
    assert float((lp - outs["log_prob"]).abs().max()) < 1e-4
    stats_match(meta["stats"])


def test_golden_conditional_flow():
    meta, sd, ins, outs = load_golden("cflow_sample_logprob")
    Fl = port.flow_from_state_dict(sd)
    assert rel_row_err(outs["x"], port.flow_sample(Fl, ins["xT"], ins["cond"])) < TIGHT
    stats_match(meta["stats_sample"])
    lp = port.flow_log_prob(Fl, outs["x"], ins["cond"], atol=1e-6, rtol=1e-6)
    assert float((lp - outs["log_prob"]).abs().max()) < 1e-4
    stats_match(meta["stats_logprob"])


def test_golden_cfg2_pfode_all_methods():
    meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
    M = port.score_model_from_state_dict(sd, port.make_sde("vp"), True)
    x = port.sample_ode_from_base(M, ins["base"], ins["cond"], 1e-5, 1e-5, options={"step_t": torch.tensor([1e-3])})[0]
    assert rel_row_err(outs["x_dopri5"], x) < TIGHT
    stats_match(meta["stats"])
    x4 = port.sample_ode_from_base(M, ins["base"], ins["cond"], method="rk4", options={"step_size": 1 / 64})[0]
    assert rel_row_err(outs["x_rk4"], x4) < TIGHT
    xe = port.sample_ode_from_base(M, ins["base"], ins["cond"], method="euler", options={"step_size": 1 / 128})[0]
    assert rel_row_err(outs["x_euler"], xe) < TIGHT


@pytest.mark.parametrize("kind", ["ve", "subvp", "vp"])
def test_golden_sigma_pfode(kind):
    meta, sd, ins, outs = load_golden(f"{kind}_sigma_pfode")
    M = port.score_model_from_state_dict(sd, port.make_sde(kind), False)
    opts = None if meta["call"]["step_t"] is None else {"step_t": torch.tensor([meta["call"]["step_t"]])}
    x = port.sample_ode_from_base(M, ins["base"], None, 1e-5, 1e-5, options=opts)[0]
    assert rel_row_err(outs["x_dopri5"], x) < TIGHT
    stats_match(meta["stats"])


def test_golden_score_logprob_exact_and_hutch():
    meta, sd, ins, outs = load_golden("score_logprob_vp")
    M = port.score_model_from_state_dict(sd, port.make_sde("vp"), True)
    lp = port.score_log_prob(M, ins["x0"], ins["cond"])
    assert lp.shape == outs["lp_exact"].shape
    assert float((lp - outs["lp_exact"]).abs().max()) < 1e-4
    stats_match(meta["stats"])
    lph = port.score_log_prob(M, ins["x0"], ins["cond"], probes=ins["probes"])
    assert float((lph - outs["lp_hutch"]).abs().max()) < 1e-4
    stats_match(meta["stats_hutch"])


def test_golden_score_logprob_ve():
    meta, sd, ins, outs = load_golden("score_logprob_ve")
    M = port.score_model_from_state_dict(sd, port.make_sde("ve"), False)
    lp = port.score_log_prob(M, ins["x0"])
    assert float((lp - outs["lp_exact"]).abs().max()) < 1e-4
    stats_match(meta["stats"])


def test_golden_euler_maruyama():
    meta, sd, ins, outs = load_golden("cfg4_vp_em")
    M = port.score_model_from_state_dict(sd, port.make_sde("vp"), True)
    run = meta["runs"][0]                                   # 100 steps; the 1000-step run is covered on the GPU
    torch.manual_seed(run["seed"])
    x0 = torch.distributions.Normal(torch.zeros(32), 1.0).sample([run["B"]])
    dw = torch.stack([torch.randn_like(x0) for _ in range(run["steps"])])
    assert rel_row_err(outs[f"x_{run['steps']}"], port.sample_sde(M, x0, dw)) < TIGHT
    meta, sd, ins, outs = load_golden("ve_em_cond")
    M = port.score_model_from_state_dict(sd, port.make_sde("ve"), False)
    assert rel_row_err(outs["x"], port.sample_sde(M, ins["x0"], ins["dw"], ins["cond"])) < TIGHT


@pytest.mark.parametrize("name", ["cfg5_symplectic", "symplectic_cond"])
def test_golden_symplectic(name):
    meta, sd, ins, outs = load_golden(name)
    Sy = port.symplectic_from_state_dict(sd)
    cond = ins.get("cond")
    assert rel_row_err(outs["x_sample"], port.symplectic_sample(Sy, ins["z0"], cond, meta["num_steps"])) < TIGHT
    lp = port.symplectic_log_prob(Sy, ins["x"], ins["p0"], cond)
    assert float((lp - outs["log_prob"]).abs().max()) < 1e-4
    stats_match(meta["stats_logprob"])


# ---------------------------------------------------------------------------------------------
# (b) the live reference (build container only; the GPU box has no /root/reference)
# ---------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not loader.reference_available(), reason="reference checkout not present")


@needs_ref
def test_reference_flow_sample_and_logprob_fresh_inputs():
    D, F, S = loader.load_reference()
    torch.manual_seed(321)
    m = F.ODEFlow(3, [32, 48]).eval()
    x = torch.randn(40, 3, generator=torch.Generator().manual_seed(5))
    Fl = port.flow_from_state_dict(m.state_dict())
    assert rel_row_err(m.sample(x).detach(), port.flow_sample(Fl, x)) < TIGHT
    assert float((m.log_prob(x).detach() - port.flow_log_prob(Fl, x)).abs().max()) < 1e-5


@needs_ref
def test_reference_pfode_and_em_fresh_inputs():
    D, F, S = loader.load_reference()
    torch.manual_seed(654)
    sm = D.ScoreModel(D.MLP(4, 1, 4, [24, 24]), D.VESDE(), no_sigma=False).eval()
    base = torch.randn(30, 4, generator=torch.Generator().manual_seed(6))
    cond = torch.randn(30, 1, generator=torch.Generator().manual_seed(7))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("ve"), False)
    ref = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5)[0].detach()
    assert rel_row_err(ref, port.sample_ode_from_base(M, base, cond, 1e-5, 1e-5)[0]) < TIGHT
    ref_stats = port.last_stats()
    assert ref_stats.nfe == 2 + 6 * (ref_stats.accepted + ref_stats.rejected)


# ---------------------------------------------------------------------------------------------
# (c) known-answer tests of the restated solver
# ---------------------------------------------------------------------------------------------
def test_solver_linear_ode_matches_exp():
    lam = torch.tensor([-1.0, -0.3, 0.5, 2.0], dtype=torch.float64)
    y0 = torch.ones(4, dtype=torch.float64)
    y = tde.odeint(lambda t, y: lam * y, y0, torch.tensor([0.0, 1.5], dtype=torch.float64), rtol=1e-10, atol=1e-12)[-1]
    assert torch.allclose(y, torch.exp(1.5 * lam), rtol=1e-8)
    # descending time integrates -t with -f (T2)
    yb = tde.odeint(lambda t, y: lam * y, y, torch.tensor([1.5, 0.0], dtype=torch.float64), rtol=1e-10, atol=1e-12)[-1]
    assert torch.allclose(yb, y0, rtol=1e-7)


def test_solver_rk4_is_three_eighths_rule_and_fourth_order():
    f = lambda t, y: torch.stack([y[1], -y[0]])             # noqa: E731  harmonic oscillator
    y0 = torch.tensor([1.0, 0.0], dtype=torch.float64)
    errs = []
    for n in (16, 32):
        y = tde.odeint(f, y0, torch.tensor([0.0, 1.0], dtype=torch.float64), method="rk4", options={"step_size": 1.0 / n})[-1]
        errs.append(float((y - torch.tensor([math.cos(1.0), -math.sin(1.0)], dtype=torch.float64)).abs().max()))
    assert 12.0 < errs[0] / errs[1] < 20.0                  # halving h cuts the error ~16x
    # one step of the 3/8 rule by hand (T13)
    h = 0.25
    k1 = f(0, y0); k2 = f(0, y0 + h * k1 / 3); k3 = f(0, y0 + h * (k2 - k1 / 3)); k4 = f(0, y0 + h * (k1 - k2 + k3))
    y1 = y0 + h * (k1 + 3 * (k2 + k3) + k4) / 8
    got = tde.odeint(f, y0, torch.tensor([0.0, 0.25], dtype=torch.float64), method="rk4")[-1]
    assert torch.allclose(got, y1, rtol=1e-14)


def test_solver_counts_overshoot_and_step_t():
    calls = []

    def f(t, y):
        calls.append(float(t))
        return -y
    y0 = torch.ones(8)
    tde.odeint(f, y0, torch.tensor([0.0, 1.0]), rtol=1e-4, atol=1e-4)
    s = tde.last_stats()
    assert s.nfe == len(calls) == 2 + 6 * (s.accepted + s.rejected)                # T6 + T8
    assert max(calls) > 1.0                                                         # T11: no clipping at t_end
    calls.clear()
    tde.odeint(f, y0, torch.tensor([0.0, 1.0]), rtol=1e-4, atol=1e-4, options={"step_t": torch.tensor([1.0])})
    assert max(calls) <= 1.0                                                        # T12: lands on the grid point


def test_solver_mixed_norm_for_tuple_states():
    # (x, c) with dc/dt = 0: the tuple norm is the max of per-component RMS norms (T5)
    f = lambda t, s: (-s[0], torch.zeros_like(s[1]))        # noqa: E731
    x0, c0 = torch.ones(5, 2), 100.0 * torch.ones(5, 3)
    xa = tde.odeint(f, (x0, c0), torch.tensor([0.0, 1.0]), rtol=1e-5, atol=1e-5)
    sa = tde.last_stats()
    assert torch.allclose(xa[1][-1], c0)
    assert abs(float(xa[0][-1][0, 0]) - math.exp(-1.0)) < 1e-4
    assert sa.accepted >= 1 and sa.rejected >= 0


def test_solver_agrees_with_scipy_rk45_on_solutions():
    """Independent cross-check of the tableau / controller: scipy's RK45 is the same Dormand-Prince pair
    (its error weights are 3/2 of torchdiffeq's, so step counts differ; solutions must agree)."""
    from scipy.integrate import solve_ivp

    def rhs(t, y):
        return np.array([y[1], (1.0 - y[0] ** 2) * y[1] - y[0] + math.sin(t)])     # forced van der Pol
    ref = solve_ivp(rhs, (0.0, 3.0), [1.0, 0.5], method="RK45", rtol=1e-10, atol=1e-12).y[:, -1]
    f = lambda t, y: torch.stack([y[1], (1.0 - y[0] ** 2) * y[1] - y[0] + torch.sin(t)])   # noqa: E731
    got = tde.odeint(f, torch.tensor([1.0, 0.5], dtype=torch.float64), torch.tensor([0.0, 3.0], dtype=torch.float64),
                     rtol=1e-10, atol=1e-12)[-1]
    assert np.allclose(got.numpy(), ref, rtol=1e-7, atol=1e-9)


def test_other_adaptive_tableaus_order_and_scipy_rk23():
    """bosh3 / adaptive_heun / fehlberg2 (restated from torchdiffeq 0.2.x, parity unpinned like the driver): the step
    solution has the tableau's order (one forced step of size h: error ~ h^(order+1)), the embedded error estimate has the
    order of the lower-order solution, bosh3's accepted-step count tracks scipy's RK23 (the same Bogacki-Shampine pair)."""
    import math
    from scipy.integrate import solve_ivp
    from oracle.torchdiffeq import _solver as so
    f = lambda t, y: torch.stack([y[1], -y[0]])                     # noqa: E731
    y0 = torch.tensor([1.0, 0.0], dtype=torch.float64)
    exact = lambda h: torch.tensor([math.cos(h), -math.sin(h)], dtype=torch.float64)      # noqa: E731
    for method, order in (("bosh3", 3), ("adaptive_heun", 2), ("fehlberg2", 2), ("dopri5", 5)):
        errs, ests = [], []
        for h in (0.2, 0.1):
            st = so.SolveStats()
            sol = so._Dopri5(so._Wrapped(f, None, False, st), y0, 1e-9, 1e-9, so._rms, st, tableau=method)
            t0, dt = torch.tensor(0.0, dtype=torch.float64), torch.tensor(h, dtype=torch.float64)
            y1, f1, err, k = sol._rk_step(y0, f(t0, y0), t0, dt, t0 + dt)
            errs.append(float((y1 - exact(h)).abs().max())); ests.append(float(err.abs().max()))
        assert order + 0.6 < math.log2(errs[0] / errs[1]) < order + 1.6, (method, errs)      # local error ~ h^(order+1)
        assert math.log2(ests[0] / ests[1]) > order - 0.5, (method, ests)                   # estimate ~ h^order or higher
    g = lambda t, y: torch.stack([y[1], -y[0] - 0.1 * y[1] + torch.sin(2 * t)])            # noqa: E731
    for tol in (1e-4, 1e-6):
        tde.odeint(g, y0, torch.tensor([0.0, 5.0], dtype=torch.float64), rtol=tol, atol=tol, method="bosh3")
        n_ours = tde.last_stats().accepted
        r = solve_ivp(lambda t, y: [y[1], -y[0] - 0.1 * y[1] + math.sin(2 * t)], (0, 5), [1.0, 0.0], rtol=tol, atol=tol, method="RK23")
        assert abs(n_ours - (r.t.size - 1)) <= 0.1 * n_ours + 2, (tol, n_ours, r.t.size - 1)


@needs_ref
def test_reference_population_wrappers_and_accelerate():
    """Row a10: the population-level wrappers (affine glue, quirks Q6-Q8) -- live reference vs port, and
    flowfusion_b200.accelerate() building a twin with identical weights (construction only: no GPU here)."""
    import flowfusion_b200 as ffb
    D, F, S = loader.load_reference()
    torch.manual_seed(77)
    shift, scale = torch.linspace(-1, 1, 5), torch.linspace(0.5, 2.0, 5)
    cshift, cscale = torch.tensor([0.3, -0.2]), torch.tensor([1.5, 0.7])
    pm = D.PopulationModelDiffusionConditional(D.MLP(5, 2, 6, [32, 48]), D.VESDE(), shift, scale, cshift, cscale).eval()
    base = torch.randn(25, 5, generator=torch.Generator().manual_seed(1)) * 10.0     # VE prior scale
    cond = torch.randn(25, 2, generator=torch.Generator().manual_seed(2))
    M = port.score_model_from_state_dict(pm.score_model.state_dict(), port.make_sde("ve"), False)
    ref = pm.forward(base, cond).detach()
    got = port.population_forward(M, base, shift, scale, cond, cshift, cscale)
    assert rel_row_err(ref, got) < TIGHT
    x = ref[:10]
    ref_lp = pm.log_prob(x, cond[:10]).detach()
    got_lp = port.population_log_prob(M, x, shift, scale, cond[:10], cshift, cscale)
    assert ref_lp.shape == got_lp.shape == (10, 1)
    assert float((ref_lp - got_lp).abs().max()) < 1e-4
    twin = ffb.accelerate(pm, device="cpu")
    assert type(twin).__name__ == "PopulationModelDiffusionConditional"
    sd_ref, sd_twin = pm.state_dict(), twin.state_dict()
    assert set(sd_ref) == set(sd_twin)
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_twin[k]), k
    for obj in (F.ODEFlow(3, [16, 16]), F.ConditionalODEFlow(3, 2, [16]), S.SymplecticFlowModel(
            S.SymplecticMLP(4, 1, 4, [16]), torch.zeros(4), torch.ones(4), torch.zeros(1), torch.ones(1))):
        tw = ffb.accelerate(obj.eval(), device="cpu")
        assert set(obj.state_dict()) == set(tw.state_dict())
        for k, v in obj.state_dict().items():
            assert torch.equal(v, tw.state_dict()[k]), k


@needs_ref
@pytest.mark.parametrize("act_cls", [torch.nn.Tanh, torch.nn.GELU, torch.nn.ReLU, torch.nn.Softplus])
def test_accelerate_keeps_activation(act_cls):
    """accelerate() must rebuild the twin with the reference object's hidden activation (`flow.py:41, 478`,
    `symplectic.py:25`, `diffusion.py:38`): same module types in the same order, same state_dict."""
    import flowfusion_b200 as ffb
    D, F, S = loader.load_reference()
    torch.manual_seed(3)
    objs = [F.ODEFlow(3, [16, 16], activation=act_cls), F.ConditionalODEFlow(3, 2, [16, 8], activation=act_cls),
            S.SymplecticMLP(4, 1, 4, [16, 16], activation=act_cls()),
            S.SymplecticFlowModel(S.SymplecticMLP(4, 1, 4, [16], activation=act_cls()), torch.zeros(4), torch.ones(4),
                                  torch.zeros(1), torch.ones(1)),
            D.ScoreModel(D.MLP(3, 0, 4, [16, 16], activation=act_cls()), D.VPSDE(), no_sigma=True)]
    seqs = [lambda o: o.layers, lambda o: o.layers, lambda o: list(o.mlp_q_dynamics) + list(o.mlp_p_dynamics),
            lambda o: list(o.model.mlp_q_dynamics) + list(o.model.mlp_p_dynamics), lambda o: [o.model.activation]]
    for obj, seq in zip(objs, seqs):
        tw = ffb.accelerate(obj.eval(), device="cpu")
        assert [type(m).__name__ for m in seq(obj)] == [type(m).__name__ for m in seq(tw)], type(obj).__name__
        assert any(isinstance(m, act_cls) for m in seq(tw))
        for k, v in obj.state_dict().items():
            assert torch.equal(v, tw.state_dict()[k]), k


@needs_ref
@pytest.mark.parametrize("act_cls,act_fn", [(torch.nn.Tanh, torch.tanh), (torch.nn.GELU, torch.nn.functional.gelu)])
def test_reference_non_silu_activations(act_cls, act_fn):
    """`activation=` is a public constructor argument of every reference model: the port follows it."""
    D, F, S = loader.load_reference()
    torch.manual_seed(5)
    m = F.ODEFlow(3, [24, 24], activation=act_cls).eval()
    x = torch.randn(30, 3, generator=torch.Generator().manual_seed(1))
    Fl = port.flow_from_state_dict(m.state_dict(), act=act_fn)
    assert rel_row_err(m.sample(x).detach(), port.flow_sample(Fl, x)) < TIGHT
    assert float((m.log_prob(x).detach() - port.flow_log_prob(Fl, x)).abs().max()) < 1e-5
    sm = D.ScoreModel(D.MLP(3, 0, 4, [24], activation=act_cls()), D.VESDE(), no_sigma=True).eval()
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("ve"), True, act=act_fn)
    assert rel_row_err(sm.sample_ode_from_base(x, atol=1e-5, rtol=1e-5)[0].detach(),
                       port.sample_ode_from_base(M, x, None, 1e-5, 1e-5)[0]) < TIGHT
