"""GPU parity tests (run on the B200 box with `pytest -m gpu`): the CUDA path, called through
the C ABI, against (i) the golden vectors produced by the UNMODIFIED reference and (ii) the CPU
oracle (oracle/port.py) on fresh seeded inputs.  Tolerances (BASELINE.json north_star):
samples <= 1e-4 relative (per-row inf-norm / max(1, |x|_inf)), log-prob <= 1e-3 nat, identical
accepted/rejected dopri5 step counts (except rtol 1e-7 runs, where FP32 rounding noise decides)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_row_err

pytestmark = pytest.mark.gpu

SAMPLE_TOL = 1e-4
LP_TOL = 1e-3


def _mods():
    import flowfusion_b200.diffusion as D
    import flowfusion_b200.flow as F
    import flowfusion_b200.symplectic as Sy
    return D, F, Sy


def check_stats(stats, meta_stats):
    assert (stats.accepted, stats.rejected) == (meta_stats["accepted"], meta_stats["rejected"]), \
        (stats.accepted, stats.rejected, stats.ratio_history)


def test_library_loaded(cuda_dev):
    from flowfusion_b200 import _lib, engine
    lib = _lib.load()
    assert lib.ffb_abi_version() == _lib.ABI_VERSION
    info = engine.device_info()
    assert info["cc_major"] == 10, info


def test_cfg1_flow_sample(cuda_dev):
    D, F, Sy = _mods()
    meta, sd, ins, outs = load_golden("cfg1_flow_sample")
    m = F.ODEFlow(**meta["ctor"]).eval()
    m.load_state_dict(sd)
    x = m.to(cuda_dev).sample(ins["xT"].to(cuda_dev))
    assert rel_row_err(outs["x"], x) < SAMPLE_TOL


def test_cfg3_flow_logprob(cuda_dev):
    D, F, Sy = _mods()
    meta, sd, ins, outs = load_golden("cfg3_flow_logprob")
    m = F.ODEFlow(**meta["ctor"], target_shift=sd["target_shift"], target_scale=sd["target_scale"]).eval()
    m.load_state_dict(sd)
    lp = m.to(cuda_dev).log_prob(ins["x"].to(cuda_dev))
    assert lp.shape == outs["log_prob"].shape
    assert float((lp.cpu() - outs["log_prob"]).abs().max()) < LP_TOL
    check_stats(m.last_stats, meta["stats"])


def test_conditional_flow(cuda_dev):
    D, F, Sy = _mods()
    meta, sd, ins, outs = load_golden("cflow_sample_logprob")
    m = F.ConditionalODEFlow(**meta["ctor"]).eval()
    m.load_state_dict(sd)
    m.to(cuda_dev)
    x = m.sample(ins["xT"].to(cuda_dev), ins["cond"].to(cuda_dev))
    assert rel_row_err(outs["x"], x) < SAMPLE_TOL
    lp = m.log_prob(outs["x"].to(cuda_dev), ins["cond"].to(cuda_dev), atol=1e-6, rtol=1e-6)
    assert float((lp.cpu() - outs["log_prob"]).abs().max()) < LP_TOL
    check_stats(m.last_stats, meta["stats_logprob"])


def _score_model(meta, sd, dev):
    D, F, Sy = _mods()
    net = D.MLP(**meta["ctor"])
    sde = {"vp": D.VPSDE, "ve": D.VESDE, "subvp": D.SUBVPSDE}[meta["sde"]]()
    sm = D.ScoreModel(net, sde, no_sigma=meta["no_sigma"]).eval()
    sm.load_state_dict(sd)
    return sm.to(dev)


def test_cfg2_pfode_all_methods(cuda_dev):
    meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
    sm = _score_model(meta, sd, cuda_dev)
    base, cond = ins["base"].to(cuda_dev), ins["cond"].to(cuda_dev)
    x, aux = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options={"step_t": torch.tensor([1e-3])})
    assert aux == []
    assert rel_row_err(outs["x_dopri5"], x) < SAMPLE_TOL
    check_stats(sm.last_stats, meta["stats"])
    x4, _ = sm.sample_ode_from_base(base, cond, method="rk4", options={"step_size": 1 / 64})
    assert rel_row_err(outs["x_rk4"], x4) < SAMPLE_TOL
    xe, _ = sm.sample_ode_from_base(base, cond, method="euler", options={"step_size": 1 / 128})
    assert rel_row_err(outs["x_euler"], xe) < SAMPLE_TOL


@pytest.mark.parametrize("kind", ["ve", "subvp", "vp"])
def test_sigma_pfode(cuda_dev, kind):
    meta, sd, ins, outs = load_golden(f"{kind}_sigma_pfode")
    sm = _score_model(meta, sd, cuda_dev)
    opts = None if meta["call"]["step_t"] is None else {"step_t": torch.tensor([meta["call"]["step_t"]])}
    x, _ = sm.sample_ode_from_base(ins["base"].to(cuda_dev), atol=1e-5, rtol=1e-5, options=opts)
    assert rel_row_err(outs["x_dopri5"], x) < SAMPLE_TOL
    check_stats(sm.last_stats, meta["stats"])


def test_score_logprob_exact_and_hutch(cuda_dev):
    meta, sd, ins, outs = load_golden("score_logprob_vp")
    sm = _score_model(meta, sd, cuda_dev)
    x0, cond = ins["x0"].to(cuda_dev), ins["cond"].to(cuda_dev)
    lp = sm.log_prob(x0, cond)
    assert lp.shape == (x0.shape[0], 1)
    assert float((lp.cpu() - outs["lp_exact"]).abs().max()) < LP_TOL
    check_stats(sm.last_stats, meta["stats"])
    sm.hutch = True
    lph = sm.log_prob(x0, cond, probes=ins["probes"].to(cuda_dev))
    assert float((lph.cpu() - outs["lp_hutch"]).abs().max()) < LP_TOL
    check_stats(sm.last_stats, meta["stats_hutch"])


def test_score_logprob_ve(cuda_dev):
    meta, sd, ins, outs = load_golden("score_logprob_ve")
    sm = _score_model(meta, sd, cuda_dev)
    lp = sm.log_prob(ins["x0"].to(cuda_dev))
    assert float((lp.cpu() - outs["lp_exact"]).abs().max()) < LP_TOL
    check_stats(sm.last_stats, meta["stats"])


def _replay_em_noise(seed, B, D, steps, scale=1.0):
    torch.manual_seed(seed)
    x0 = torch.distributions.Normal(torch.zeros(D), scale).sample([B])
    dw = torch.stack([torch.randn_like(x0) for _ in range(steps)])
    return x0, dw


def test_cfg4_euler_maruyama(cuda_dev):
    meta, sd, ins, outs = load_golden("cfg4_vp_em")
    sm = _score_model(meta, sd, cuda_dev)
    for run in meta["runs"]:
        x0, dw = _replay_em_noise(run["seed"], run["B"], 32, run["steps"])
        x = sm.sample_sde((run["B"], 32), steps=run["steps"], x0=x0.to(cuda_dev), noise=dw.to(cuda_dev))
        assert rel_row_err(outs[f"x_{run['steps']}"], x) < SAMPLE_TOL
        assert sm.check_stability()


def test_em_stops_at_first_nan_like_the_reference(cuda_dev, capsys):
    """`diffusion.py:560-563` on the GPU: a NaN planted in the caller's noise at (step 5, row 7) stops the whole batch (3
    tiles) after step 5 and that step's x_mean comes back; same with the in-kernel engines' status words."""
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(11)
    sm = D.ScoreModel(D.MLP(5, 0, 8, [32, 32]), D.VPSDE(), no_sigma=True).eval()
    x0 = torch.randn(300, 5, generator=torch.Generator().manual_seed(1))
    dw = torch.randn(20, 300, 5, generator=torch.Generator().manual_seed(2))
    dw[5, 7, 2] = float("nan")
    ref = port.sample_sde(port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True), x0, dw)
    assert torch.isfinite(ref).all()
    sm.to(cuda_dev)
    x = sm.sample_sde((300, 5), steps=20, x0=x0.to(cuda_dev), noise=dw.to(cuda_dev))
    assert rel_row_err(ref, x) < SAMPLE_TOL
    assert sm.stopped_at_step == 5 and not sm.check_stability()
    assert "Diffusion is not stable, NaN were produced. Stopped sampling." in capsys.readouterr().out
    x = sm.sample_sde((300, 5), steps=20, x0=x0.to(cuda_dev), noise=torch.nan_to_num(dw).to(cuda_dev))
    assert sm.stopped_at_step is None and sm.check_stability() and torch.isfinite(x).all()


def test_em_ve_conditional(cuda_dev):
    meta, sd, ins, outs = load_golden("ve_em_cond")
    sm = _score_model(meta, sd, cuda_dev)
    x = sm.sample_sde((96, 3), conditional=ins["cond"].to(cuda_dev), steps=50, x0=ins["x0"].to(cuda_dev),
                      noise=ins["dw"].to(cuda_dev))
    assert rel_row_err(outs["x"], x) < SAMPLE_TOL


@pytest.mark.parametrize("name", ["cfg5_symplectic", "symplectic_cond"])
def test_symplectic(cuda_dev, name):
    D, F, Sy = _mods()
    meta, sd, ins, outs = load_golden(name)
    net = Sy.SymplecticMLP(**meta["ctor"])
    m = Sy.SymplecticFlowModel(net, sd["shift"], sd["scale"], sd["conditional_shift"], sd["conditional_scale"]).eval()
    m.load_state_dict(sd)
    m.to(cuda_dev)
    cond = ins["cond"].to(cuda_dev) if "cond" in ins else None
    D_ = meta["ctor"]["n_data_dims"]
    x = m.sample((ins["z0"].shape[0], D_), conditional=cond, num_steps=meta["num_steps"], z0=ins["z0"].to(cuda_dev))
    assert rel_row_err(outs["x_sample"], x) < SAMPLE_TOL
    lp = m.log_prob(ins["x"].to(cuda_dev), conditional=cond, p0=ins["p0"].to(cuda_dev))
    assert float((lp.cpu() - outs["log_prob"]).abs().max()) < LP_TOL
    check_stats(m.last_stats, meta["stats_logprob"])


# ---------------------------------------------------------------------------------------------
# kernel-level checks against the CPU oracle on fresh inputs: ragged / tiny / empty batches
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [1, 127, 128, 129, 1000])
def test_field_eval_ragged_batches(cuda_dev, B):
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(7)
    net = D.MLP(5, 3, 6, [40, 72, 100])            # widths that need padding to 64 / 128
    sm = D.ScoreModel(net, D.SUBVPSDE(), no_sigma=False).eval()
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("subvp"), False)
    x = torch.randn(B, 5, generator=torch.Generator().manual_seed(B))
    c = torch.randn(B, 3, generator=torch.Generator().manual_seed(B + 1))
    t = torch.tensor(0.37)
    ref_f, ref_d = port.score_field(M, t, (x,), c, prob=True)
    sm.to(cuda_dev)
    sm.prob, sm.conditional = True, c.to(cuda_dev)
    f, d = sm.forward(t, (x.to(cuda_dev),))
    assert f.shape == (B, 5) and d.shape == (B, 1)
    assert rel_row_err(ref_f, f) < 1e-5
    assert float((d.cpu() - ref_d).abs().max()) < 1e-4 * max(1.0, float(ref_d.abs().max()))
    sm.prob = False
    f2 = sm.forward(t, (x.to(cuda_dev),))
    assert torch.equal(f2, f) or rel_row_err(f, f2) < 1e-6


def test_empty_batch(cuda_dev):
    D, F, Sy = _mods()
    torch.manual_seed(0)
    m = F.ODEFlow(3, [32, 32]).eval().to(cuda_dev)
    out = m.dynamics(torch.tensor(0.5), (torch.zeros(0, 3, device=cuda_dev),))
    assert out.shape == (0, 3)


def test_flow_divergence_matches_autograd(cuda_dev):
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(3)
    m = F.ODEFlow(16, [128] * 4).eval()
    Fl = port.flow_from_state_dict(m.state_dict())
    x = torch.randn(333, 16, generator=torch.Generator().manual_seed(5))
    t = torch.tensor(0.8)
    v_ref, d_ref = port.flow_velocity_and_divergence(Fl, t, x)
    v, d = m.to(cuda_dev).dynamics_with_jacobian(t, (x.to(cuda_dev), torch.zeros(333, 1, device=cuda_dev)))
    assert rel_row_err(v_ref, v) < 1e-5
    assert float((d.cpu() - d_ref).abs().max()) < 2e-5 * max(1.0, float(d_ref.abs().max()))


def test_deterministic_and_partition_invariant(cuda_dev):
    """Same input -> bit-identical output; rows do not depend on which tile they land in."""
    meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
    sm = _score_model(meta, sd, cuda_dev)
    base, cond = ins["base"].to(cuda_dev), ins["cond"].to(cuda_dev)
    a, _ = sm.sample_ode_from_base(base, cond, method="rk4", options={"step_size": 1 / 16})
    b, _ = sm.sample_ode_from_base(base, cond, method="rk4", options={"step_size": 1 / 16})
    assert torch.equal(a, b)
    c, _ = sm.sample_ode_from_base(base[37:300], cond[37:300], method="rk4", options={"step_size": 1 / 16})
    assert torch.equal(a[37:300], c)


def test_gaussian_logprob(cuda_dev):
    from flowfusion_b200 import engine as E
    x = torch.randn(1000, 16, generator=torch.Generator().manual_seed(1))
    add = torch.randn(1000, generator=torch.Generator().manual_seed(2))
    for sigma in (1.0, 10.0):
        ref = torch.distributions.Normal(0.0, sigma).log_prob(x).sum(1) + add
        got = E.gaussian_logprob(x.to(cuda_dev), add.to(cuda_dev), sigma)
        assert float((got.cpu() - ref).abs().max()) < 2e-5 * float(ref.abs().max())


def test_philox_normals_are_standard(cuda_dev):
    from flowfusion_b200 import engine as E
    z = torch.cat([E.philox_normal(200_000, 32, seed=5, offset=0, step=s, device=cuda_dev).reshape(-1) for s in range(3)])
    assert abs(float(z.mean())) < 2e-3 and abs(float(z.var()) - 1.0) < 3e-3
    assert abs(float((z ** 3).mean())) < 1e-2 and abs(float((z ** 4).mean()) - 3.0) < 3e-2
    # different steps / rows / seeds decorrelate; the same key reproduces
    a = E.philox_normal(1000, 32, 5, 0, 0, device=cuda_dev)
    b = E.philox_normal(1000, 32, 5, 0, 1, device=cuda_dev)
    a2 = E.philox_normal(500, 32, 5, 0, 0, row_offset=500, device=cuda_dev)
    assert torch.equal(a[500:], a2) and not torch.equal(a, b)
    assert abs(float((a * b).mean())) < 2e-2


def test_em_throughput_mode_statistics(cuda_dev):
    """In-kernel Philox noise: same seed reproduces, and the sample moments match the
    caller-supplied-noise path (zero score network => closed-form Gaussian)."""
    D, F, Sy = _mods()
    torch.manual_seed(11)
    net = D.MLP(4, 0, 4, [16])
    for p in net.NN.parameters():
        torch.nn.init.zeros_(p)
    sm = D.ScoreModel(net, D.VPSDE(), no_sigma=True).eval().to(cuda_dev)
    B = 100_000
    x0 = torch.randn(B, 4, device=cuda_dev)
    a = sm.sample_sde((B, 4), steps=50, x0=x0, seed=123)
    b = sm.sample_sde((B, 4), steps=50, x0=x0, seed=123)
    c = sm.sample_sde((B, 4), steps=50, x0=x0, seed=124)
    assert torch.equal(a, b) and not torch.equal(a, c)
    noise = torch.randn(50, B, 4, device=cuda_dev)
    d = sm.sample_sde((B, 4), steps=50, x0=x0, noise=noise)
    assert abs(float(a.var()) - float(d.var())) < 0.03 * float(d.var())
    assert abs(float(a.mean())) < 0.05 * float(d.std())


def test_leapfrog_reversible_and_second_order(cuda_dev):
    """Leapfrog extension (no reference oracle): 2nd-order convergence towards the dopri5
    solution of the same Hamiltonian field, and exact-in-FP32 structure q-only output."""
    D, F, Sy = _mods()
    torch.manual_seed(21)
    net = Sy.SymplecticMLP(4, 0, 4, [32, 32])
    net.W.mul_(1.0 / 16.0)        # slow time dependence so that 1/16 steps are in the asymptotic regime
    m = Sy.SymplecticFlowModel(net, torch.zeros(4), torch.ones(4), torch.zeros(0), torch.ones(0)).eval().to(cuda_dev)
    z0 = torch.randn(512, 8, device=cuda_dev)
    errs = []
    ref = m.sample((512, 4), num_steps=1024, z0=z0, method="leapfrog")
    for n in (8, 16, 32):
        q = m.sample((512, 4), num_steps=n, z0=z0, method="leapfrog")
        errs.append(float((q - ref).abs().max()))
    assert errs[0] / errs[1] > 3.0 and errs[1] / errs[2] > 3.0, errs     # ~4x per halving
    e_euler = float((m.sample((512, 4), num_steps=32, z0=z0) - ref).abs().max())
    assert errs[2] < e_euler
