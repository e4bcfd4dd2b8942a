"""GPU tests of the wide engine (csrc/ffb_engine_wide.cuh; run with `pytest -m gpu`): networks the tensor-core engines do not
hold -- hidden widths above 128 (the reference takes any `units=[...]`, diffusion.py:32-40, flow.py:37-44) and more than 8
Linear layers -- against the CPU oracle (oracle/port.py) on seeded inputs, and, forced with FFB_ENGINE=wide (engine 5), the
reference's golden vectors of every BASELINE config against the same bar as the tensor-core engines."""
import pytest
import torch

from conftest import load_golden, rel_row_err
from test_gpu_engines import engine, gen, _mods
from test_gpu_parity import _score_model, check_stats, _replay_em_noise

pytestmark = pytest.mark.gpu

SAMPLE_TOL = 1e-4
LP_TOL = 1e-3
WIDE = 5


# ---------------------------------------------------------------------------------------------
# the golden vectors of the unmodified reference, on the wide engine
# ---------------------------------------------------------------------------------------------
def test_wide_engine_on_golden_sampling(cuda_dev):
    with engine(WIDE):
        meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
        sm = _score_model(meta, sd, cuda_dev)
        base, cond = ins["base"].to(cuda_dev), ins["cond"].to(cuda_dev)
        x, _ = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options={"step_t": torch.tensor([1e-3])})
        assert rel_row_err(outs["x_dopri5"], x) < SAMPLE_TOL
        check_stats(sm.last_stats, meta["stats"])
        x4, _ = sm.sample_ode_from_base(base, cond, method="rk4", options={"step_size": 1 / 64})
        assert rel_row_err(outs["x_rk4"], x4) < SAMPLE_TOL
        # Euler-Maruyama (cfg4) with the reference's noise
        meta, sd, ins, outs = load_golden("cfg4_vp_em")
        sm = _score_model(meta, sd, cuda_dev)
        run = meta["runs"][0]
        x0, dw = _replay_em_noise(run["seed"], run["B"], 32, run["steps"])
        x = sm.sample_sde((run["B"], 32), steps=run["steps"], x0=x0.to(cuda_dev), noise=dw.to(cuda_dev))
        assert rel_row_err(outs[f"x_{run['steps']}"], x) < SAMPLE_TOL
        # sigma division (VE) and the flow sampler
        meta, sd, ins, outs = load_golden("ve_sigma_pfode")
        sm = _score_model(meta, sd, cuda_dev)
        opts = None if meta["call"]["step_t"] is None else {"step_t": torch.tensor([meta["call"]["step_t"]])}
        x, _ = sm.sample_ode_from_base(ins["base"].to(cuda_dev), atol=1e-5, rtol=1e-5, options=opts)
        assert rel_row_err(outs["x_dopri5"], x) < SAMPLE_TOL
        check_stats(sm.last_stats, meta["stats"])


def test_wide_engine_on_golden_logprob(cuda_dev):
    D, F, Sy = _mods()
    with engine(WIDE):
        # exact trace, D = 16: a sample is 6 row groups (primal + 3 tangents each)
        meta, sd, ins, outs = load_golden("cfg3_flow_logprob")
        m = F.ODEFlow(**meta["ctor"], target_shift=sd["target_shift"], target_scale=sd["target_scale"]).eval()
        m.load_state_dict(sd)
        lp = m.to(cuda_dev).log_prob(ins["x"].to(cuda_dev))
        assert float((lp.cpu() - outs["log_prob"]).abs().max()) < LP_TOL
        check_stats(m.last_stats, meta["stats"])
        # conditional score model: exact trace and Hutchinson (2 samples x (primal, tangent) per row group)
        meta, sd, ins, outs = load_golden("score_logprob_vp")
        sm = _score_model(meta, sd, cuda_dev)
        x0, cond = ins["x0"].to(cuda_dev), ins["cond"].to(cuda_dev)
        lp = sm.log_prob(x0, cond)
        assert float((lp.cpu() - outs["lp_exact"]).abs().max()) < LP_TOL
        check_stats(sm.last_stats, meta["stats"])
        sm.hutch = True
        lph = sm.log_prob(x0, cond, probes=ins["probes"].to(cuda_dev))
        assert float((lph.cpu() - outs["lp_hutch"]).abs().max()) < LP_TOL
        check_stats(sm.last_stats, meta["stats_hutch"])
        # conditional flow with a non-SiLU activation (reference golden)
        meta, sd, ins, outs = load_golden("cflow_sample_logprob")
        m = F.ConditionalODEFlow(**meta["ctor"]).eval()
        m.load_state_dict(sd)
        m.to(cuda_dev)
        x = m.sample(ins["xT"].to(cuda_dev), ins["cond"].to(cuda_dev))
        assert rel_row_err(outs["x"], x) < SAMPLE_TOL
        lp = m.log_prob(outs["x"].to(cuda_dev), ins["cond"].to(cuda_dev), atol=1e-6, rtol=1e-6)
        assert float((lp.cpu() - outs["log_prob"]).abs().max()) < LP_TOL


@pytest.mark.parametrize("name", ["cfg5_symplectic", "symplectic_cond"])
def test_wide_engine_on_golden_symplectic(cuda_dev, name):
    """The two-network (q, p) field: Euler sampling and the dopri5 log-prob of the reference's golden vectors."""
    import test_gpu_parity as P
    with engine(WIDE):
        P.test_symplectic(cuda_dev, name)


# ---------------------------------------------------------------------------------------------
# networks only the wide engine holds, against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("units", [[256] * 3, [512] * 2, [64] * 10, [200, 384, 72]])
def test_wide_networks_pfode_sampling(cuda_dev, units):
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(500 + len(units))
    sm = D.ScoreModel(D.MLP(16, 4, 8, units), D.VPSDE(), no_sigma=True).eval()
    B = 333                                              # ragged: 2 full tiles + 77 rows
    base = torch.randn(B, 16, generator=gen(1)); cond = torch.randn(B, 4, generator=gen(2))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    opts = {"step_t": torch.tensor([1e-3])}
    ref = port.sample_ode_from_base(M, base, cond, 1e-5, 1e-5, options=opts)[0]
    rs = port.last_stats()
    ref4 = port.sample_ode_from_base(M, base, cond, method="rk4", options={"step_size": 1 / 16})[0]
    sm.to(cuda_dev)
    x, _ = sm.sample_ode_from_base(base.to(cuda_dev), cond.to(cuda_dev), atol=1e-5, rtol=1e-5, options=opts)
    assert rel_row_err(ref, x) < SAMPLE_TOL
    assert (sm.last_stats.accepted, sm.last_stats.rejected) == (rs.accepted, rs.rejected)
    x4, _ = sm.sample_ode_from_base(base.to(cuda_dev), cond.to(cuda_dev), method="rk4", options={"step_size": 1 / 16})
    assert rel_row_err(ref4, x4) < SAMPLE_TOL
    # Euler-Maruyama with caller noise
    x0 = torch.randn(B, 16, generator=gen(3)); dw = torch.randn(25, B, 16, generator=gen(4))
    refs = port.sample_sde(M, x0, dw, cond)
    xs = sm.sample_sde((B, 16), conditional=cond.to(cuda_dev), steps=25, x0=x0.to(cuda_dev), noise=dw.to(cuda_dev))
    assert rel_row_err(refs, xs) < SAMPLE_TOL


@pytest.mark.parametrize("D_,units,act_cls,act_fn", [(16, [256] * 3, torch.nn.SiLU, None), (8, [512] * 2, torch.nn.SiLU, None),
                                                     (5, [160, 320], torch.nn.Tanh, torch.tanh), (40, [256, 256], torch.nn.SiLU, None)])
def test_wide_networks_flow_logprob_exact(cuda_dev, D_, units, act_cls, act_fn):
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(600 + D_)
    fl = F.ODEFlow(D_, units, activation=act_cls).eval()
    B = 45
    xs = torch.randn(B, D_, generator=gen(5))
    kw = {} if act_fn is None else {"act": act_fn}
    pm = port.flow_from_state_dict(fl.state_dict(), **kw)
    ref = port.flow_log_prob(pm, xs)
    rs = port.last_stats()
    ref_rk = port.flow_log_prob(pm, xs, method="rk4", options={"step_size": 0.25})
    fl.to(cuda_dev)
    lp = fl.log_prob(xs.to(cuda_dev))
    assert float((lp.cpu() - ref).abs().max()) < LP_TOL
    assert (fl.last_stats.accepted, fl.last_stats.rejected) == (rs.accepted, rs.rejected)
    lp_rk = fl.log_prob(xs.to(cuda_dev), method="rk4", options={"step_size": 0.25})
    assert float((lp_rk.cpu() - ref_rk).abs().max()) < LP_TOL
    # sampling from the same flow
    xT = torch.randn(B, D_, generator=gen(6))
    if D_ <= 8:                                          # rtol 1e-7 (torchdiffeq's defaults): keep the oracle run short
        x = fl.sample(xT.to(cuda_dev))
        assert rel_row_err(port.flow_sample(pm, xT), x) < SAMPLE_TOL


def test_wide_network_score_logprob_hutchinson(cuda_dev):
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(77)
    sm = D.ScoreModel(D.MLP(12, 3, 8, [256, 256]), D.VPSDE(), no_sigma=True, hutchinson=True).eval()
    B = 150
    x0 = torch.randn(B, 12, generator=gen(7)); cond = torch.randn(B, 3, generator=gen(8))
    e = torch.sign(torch.randn(B, 12, generator=gen(9)))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    ref = port.score_log_prob(M, x0, cond, probes=e)
    rs = port.last_stats()
    sm.to(cuda_dev)
    lp = sm.log_prob(x0.to(cuda_dev), cond.to(cuda_dev), probes=e.to(cuda_dev))
    assert float((lp.cpu() - ref).abs().max()) < LP_TOL
    assert (sm.last_stats.accepted, sm.last_stats.rejected) == (rs.accepted, rs.rejected)


def test_wide_limits(cuda_dev):
    """include/ffb200.h: 16 Linear layers, hidden widths up to 512; beyond that NotImplementedError, nothing is launched."""
    D, F, Sy = _mods()
    with pytest.raises(NotImplementedError):
        D.ScoreModel(D.MLP(4, 0, 8, [32] * 16), D.VESDE()).eval().to(cuda_dev).sample_ode_from_base(torch.zeros(4, 4, device=cuda_dev))
    with pytest.raises(NotImplementedError):
        D.ScoreModel(D.MLP(4, 0, 8, [513]), D.VESDE()).eval().to(cuda_dev).sample_ode_from_base(torch.zeros(4, 4, device=cuda_dev))
    # partition invariance of a wide solve: tiles do not talk to each other
    torch.manual_seed(5)
    sm = D.ScoreModel(D.MLP(6, 0, 8, [256, 192]), D.VESDE(), no_sigma=True).eval().to(cuda_dev)
    base = torch.randn(700, 6, generator=gen(1)).to(cuda_dev)
    xa, _ = sm.sample_ode_from_base(base, method="rk4", options={"step_size": 0.125})
    xb = torch.cat([sm.sample_ode_from_base(base[:129], method="rk4", options={"step_size": 0.125})[0],
                    sm.sample_ode_from_base(base[129:], method="rk4", options={"step_size": 0.125})[0]])
    assert torch.equal(xa, xb)


def test_new_kernels_repeat_bit_for_bit(cuda_dev):
    """Stand-in for racecheck (compute-sanitizer is closed on the GPU pool, profiles/r02_sanitizer_unavailable.txt): the wide
    engine's dopri5 attempts, the fused training step (loss, gradients, d/dX) and the Hamiltonian leapfrog, 8 runs each on the
    same inputs, must agree bit for bit -- a race on the weight ring's mbarriers or on a warp's activation rows would not."""
    import copy
    D, F, Sy = _mods()
    from flowfusion_b200 import training
    torch.manual_seed(3)
    sm = D.ScoreModel(D.MLP(16, 4, 8, [256, 256]), D.VPSDE(), no_sigma=True).eval().to(cuda_dev)
    base = torch.randn(1500, 16, generator=gen(1)).to(cuda_dev); cond = torch.randn(1500, 4, generator=gen(2)).to(cuda_dev)
    fl = F.ODEFlow(8, [192, 256]).eval().to(cuda_dev)
    xs = torch.randn(200, 8, generator=gen(3)).to(cuda_dev)
    lin = [torch.nn.Linear(21, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 9)]
    lin = [copy.deepcopy(l).to(cuda_dev) for l in lin]
    x = torch.randn(1000, 21, generator=gen(4)).to(cuda_dev); beta = torch.randn(1000, 9, generator=gen(5)).to(cuda_dev)
    ham = Sy.HamiltonianMLP(8, 0, [128, 128]).to(cuda_dev)
    z0 = torch.randn(777, 16, generator=gen(6)).to(cuda_dev)
    ref = None
    for _ in range(8):
        a, _ = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options={"step_t": torch.tensor([1e-3])})
        b = fl.log_prob(xs)
        loss, grads, gx = training.train_step(lin, 0, x, None, beta, 0.01, want_grad_x=True)
        z = ham.leapfrog(z0, num_steps=7, dt=0.03)
        cur = [a, b, loss, gx, z] + grads
        if ref is None:
            ref = [t.clone() for t in cur]
        else:
            for r, c in zip(ref, cur):
                assert torch.equal(r, c)
