"""The device-side dopri5 controller on a B200: the control kernel between two attempt kernels must take the
steps of the host loop (and therefore of the reference, via the golden vectors), the attempt kernels that read
their step from the controller block must produce what the launch-argument kernels produce, and the device
evaluation of the scalar programs must agree with the host programs to the last place or two."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_row_err

pytestmark = pytest.mark.gpu


def _mods():
    import flowfusion_b200.diffusion as D
    import flowfusion_b200.flow as F
    import flowfusion_b200.symplectic as Sy
    return D, F, Sy


def _both(fn):
    from flowfusion_b200 import solver as S
    with S.controller("host"):
        a, sa = fn()
    with S.controller("device"):
        b, sb = fn()
    torch.cuda.synchronize()
    assert (sa.controller, sb.controller) == ("host", "device")
    assert (sa.accepted, sa.rejected, sa.nfe) == (sb.accepted, sb.rejected, sb.nfe), (sa.ratio_history, sb.ratio_history)
    assert sa.accept_history == sb.accept_history
    # step sizes: identical unless a sqrt / sin / cos of the scalar program differs in the last place (the error
    # estimate is a difference of nearly equal terms, so such a change shows at ~1e-5 in dt)
    assert np.allclose(sa.dt_history, sb.dt_history, rtol=5e-3, atol=0)
    return a, b


def test_device_program_rows_match_host_programs(cuda_dev):
    from test_controller import _programs, _twin_rows, _ulps
    from flowfusion_b200 import _lib as L
    rng = np.random.default_rng(2)
    times = np.concatenate([rng.uniform(1e-5, 1.0, 4000), [1e-5, 1e-3, 0.5, 1.0]]).astype(np.float32)
    for name, prog in _programs():
        want = prog(times)
        want[:, L.MAX_TFEAT + 3] = 1.0
        got = _twin_rows(prog.spec, times, on_device=1)
        tf = slice(0, L.MAX_TFEAT)
        if prog.spec.time_features == L.PROG_RAW_T:
            assert np.array_equal(got[:, tf], want[:, tf])
        else:
            assert np.abs(got[:, tf] - want[:, tf]).max() <= 2 ** -22, name
        for col, tol in ((L.MAX_TFEAT + 0, 0), (L.MAX_TFEAT + 1, 0 if prog.spec.sde == L.SDE_NONE else 8),
                         (L.MAX_TFEAT + 2, 32 if prog.spec.use_sigma else None), (L.MAX_TFEAT + 3, 0)):
            if tol is None:        # sigma is read by the kernels only when use_sigma is set
                continue
            assert _ulps(got[:, col], want[:, col]).max() <= tol, (name, col)


def test_cfg2_pfode_device_vs_host_controller(cuda_dev):
    D, F, Sy = _mods()
    meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
    sm = D.ScoreModel(D.MLP(**meta["ctor"]), D.VPSDE(), no_sigma=True).eval()
    sm.load_state_dict(sd)
    sm.to(cuda_dev)
    base, cond = ins["base"].to(cuda_dev), ins["cond"].to(cuda_dev)

    def run():
        x, _ = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options={"step_t": torch.tensor([1e-3])})
        return x.cpu(), sm.last_stats
    a, b = _both(run)
    assert rel_row_err(a, b) < 1e-5
    assert rel_row_err(outs["x_dopri5"], b) < 1e-4
    assert (sm.last_stats.accepted, sm.last_stats.rejected) == (meta["stats"]["accepted"], meta["stats"]["rejected"])


@pytest.mark.parametrize("kind", ["ve", "subvp", "vp"])
def test_sigma_fields_device_controller(cuda_dev, kind):
    D, F, Sy = _mods()
    meta, sd, ins, outs = load_golden(f"{kind}_sigma_pfode")
    sde = {"vp": D.VPSDE, "ve": D.VESDE, "subvp": D.SUBVPSDE}[meta["sde"]]()
    sm = D.ScoreModel(D.MLP(**meta["ctor"]), sde, no_sigma=meta["no_sigma"]).eval()
    sm.load_state_dict(sd)
    sm.to(cuda_dev)
    opts = None if meta["call"]["step_t"] is None else {"step_t": torch.tensor([meta["call"]["step_t"]])}
    base = ins["base"].to(cuda_dev)

    def run():
        x, _ = sm.sample_ode_from_base(base, atol=1e-5, rtol=1e-5, options=opts)
        return x.cpu(), sm.last_stats
    if kind == "ve":
        a, b = _both(run)
        assert rel_row_err(a, b) < 1e-5
    else:           # sigma(t) = 1 - exp(-small): these fields keep the host loop (no device program)
        from flowfusion_b200 import solver as S
        with S.controller("device"):
            b, st = run()
        assert st.controller == "host"
    assert rel_row_err(outs["x_dopri5"], b) < 1e-4
    assert (sm.last_stats.accepted, sm.last_stats.rejected) == (meta["stats"]["accepted"], meta["stats"]["rejected"])


def test_logprob_paths_device_controller(cuda_dev):
    D, F, Sy = _mods()
    meta, sd, ins, outs = load_golden("cfg3_flow_logprob")
    m = F.ODEFlow(**meta["ctor"], target_shift=sd["target_shift"], target_scale=sd["target_scale"]).eval()
    m.load_state_dict(sd)
    m.to(cuda_dev)
    x = ins["x"].to(cuda_dev)
    a, b = _both(lambda: (m.log_prob(x).cpu(), m.last_stats))
    assert float((a - b).abs().max()) < 1e-5
    assert float((b - outs["log_prob"]).abs().max()) < 1e-3
    assert (m.last_stats.accepted, m.last_stats.rejected) == (meta["stats"]["accepted"], meta["stats"]["rejected"])
    # Hutchinson estimator + score field + conditional
    meta, sd, ins, outs = load_golden("score_logprob_vp")
    sm = D.ScoreModel(D.MLP(**meta["ctor"]), D.VPSDE(), no_sigma=meta["no_sigma"]).eval()
    sm.load_state_dict(sd)
    sm.to(cuda_dev)
    sm.hutch = True
    x0, cond, probes = ins["x0"].to(cuda_dev), ins["cond"].to(cuda_dev), ins["probes"].to(cuda_dev)
    a, b = _both(lambda: (sm.log_prob(x0, cond, probes=probes).cpu(), sm.last_stats))
    assert float((a - b).abs().max()) < 2e-4
    assert float((b - outs["lp_hutch"]).abs().max()) < 1e-3
    # conditional carried in the ODE state (mixed norm with a zero-error component)
    meta, sd, ins, outs = load_golden("cflow_sample_logprob")
    mc = F.ConditionalODEFlow(**meta["ctor"]).eval()
    mc.load_state_dict(sd)
    mc.to(cuda_dev)
    xs, c = outs["x"].to(cuda_dev), ins["cond"].to(cuda_dev)
    a, b = _both(lambda: (mc.log_prob(xs, c, atol=1e-6, rtol=1e-6).cpu(), mc.last_stats))
    assert float((a - b).abs().max()) < 1e-5
    assert (mc.last_stats.accepted, mc.last_stats.rejected) == (meta["stats_logprob"]["accepted"], meta["stats_logprob"]["rejected"])


def test_symplectic_two_network_field_device_controller(cuda_dev):
    D, F, Sy = _mods()
    meta, sd, ins, outs = load_golden("cfg5_symplectic")
    net = Sy.SymplecticMLP(**meta["ctor"])
    m = Sy.SymplecticFlowModel(net, sd["shift"], sd["scale"], sd["conditional_shift"], sd["conditional_scale"]).eval()
    m.load_state_dict(sd)
    m.to(cuda_dev)
    x, p0 = ins["x"].to(cuda_dev), ins["p0"].to(cuda_dev)
    a, b = _both(lambda: (m.log_prob(x, conditional=None, p0=p0).cpu(), m.last_stats))
    assert float((a - b).abs().max()) < 2e-4        # |log p| ~ 45: a few FP32 ulps
    assert float((b - outs["log_prob"]).abs().max()) < 1e-3


def test_rejections_failures_and_ragged_batches(cuda_dev):
    """A solve with rejected attempts (the buffer roles must not flip), the error exits, a ragged batch and the
    fall-back to the host loop on the debug engine."""
    D, F, Sy = _mods()
    from flowfusion_b200 import _lib as L
    from flowfusion_b200 import solver as S
    torch.manual_seed(11)
    m = F.ODEFlow(3, [64, 64]).eval()
    with torch.no_grad():
        for lin in m.velocity:
            if isinstance(lin, torch.nn.Linear):
                lin.weight.mul_(2.5)                 # a stiffer field: the first trial step is rejected
    m.to(cuda_dev)
    x = torch.randn(1000 + 37, 3, generator=torch.Generator().manual_seed(3)).to(cuda_dev)
    run = lambda: (m._integrate(x, None, None, 1.0, 0.0, 1e-6, 1e-6, "dopri5", {"first_step": 0.5}, L.DIV_NONE)[0].cpu(), m.last_stats)
    a, b = _both(run)
    assert m.last_stats.rejected >= 1
    assert rel_row_err(a, b) < 2e-6
    for opts, needle in (({"max_num_steps": 2}, "max_num_steps exceeded (2>=2)"), ({"first_step": 1e-20}, "underflow in dt")):
        with S.controller("device"), pytest.raises(S.SolverError) as e:
            m._integrate(x, None, None, 1.0, 0.0, 1e-6, 1e-6, "dopri5", opts, L.DIV_NONE)
        assert needle in str(e.value)
    bad = x.clone()
    bad[5, 1] = float("nan")
    with S.controller("device"), pytest.raises(S.SolverError) as e:
        m._integrate(bad, None, None, 1.0, 0.0, 1e-6, 1e-6, "dopri5", {"first_step": 0.1}, L.DIV_NONE)
    assert "non-finite" in str(e.value)      # (without first_step the NaN reaches dt first: "underflow in dt nan", as in torchdiffeq)
    lib = L.load()
    lib.ffb_set_engine(0)                            # FP32 FFMA2 debug engine: no controller block support
    try:
        with S.controller("device"):
            y, _ = m._integrate(x, None, None, 1.0, 0.0, 1e-6, 1e-6, "dopri5", {"first_step": 0.5}, L.DIV_NONE)
        assert m.last_stats.controller == "host"
        assert rel_row_err(a, y.cpu()) < 1e-4
    finally:
        lib.ffb_set_engine(1)


def test_large_batch_device_controller_is_deterministic(cuda_dev):
    """1 M rows (the bench shape): two runs are bit-identical and late no-op launches leave the result alone."""
    D, F, Sy = _mods()
    torch.manual_seed(1234)
    sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 4), D.VPSDE(), no_sigma=True).eval().to(cuda_dev)
    B = 1 << 20
    base = torch.randn(B, 16, generator=torch.Generator().manual_seed(2)).to(cuda_dev)
    cond = torch.randn(B, 4, generator=torch.Generator().manual_seed(3)).to(cuda_dev)
    opts = {"step_t": torch.tensor([1e-3])}
    x1, _ = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options=opts)
    s1 = sm.last_stats
    x2, _ = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options=opts)
    torch.cuda.synchronize()
    assert s1.controller == "device" and torch.equal(x1, x2)
    assert (s1.accepted, s1.rejected) == (sm.last_stats.accepted, sm.last_stats.rejected)
    assert bool(torch.isfinite(x1).all())
