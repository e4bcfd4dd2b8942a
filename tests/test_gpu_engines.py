"""GPU tests of the generic code paths of the tensor-core engines (run with `pytest -m gpu`):
state / conditional widths that span several 32-column chunks, tangent tiles with few samples, networks
with one Linear layer, ragged tiles -- each against the CPU oracle (oracle/port.py) on fresh seeded inputs
and against the FP32 FFMA2 engine of the same library; plus size-independent properties at the full
BASELINE sizes (determinism, partition invariance, linearity of the Euler-Maruyama noise path)."""
import numpy as np
import pytest
import torch

from conftest import rel_row_err

pytestmark = pytest.mark.gpu

SAMPLE_TOL = 1e-4
LP_TOL = 1e-3


def _mods():
    import flowfusion_b200.diffusion as D
    import flowfusion_b200.flow as F
    import flowfusion_b200.symplectic as Sy
    return D, F, Sy


class engine:
    """with engine(0): FP32 FFMA2 kernels; engine(1): tensor cores (default); engine(2): tile engine only."""

    def __init__(self, e):
        self.e = e

    def __enter__(self):
        from flowfusion_b200 import _lib
        self.lib = _lib.load()
        self.prev = self.lib.ffb_get_engine()
        self.lib.ffb_set_engine(self.e)

    def __exit__(self, *exc):
        self.lib.ffb_set_engine(self.prev)


def gen(seed):
    return torch.Generator().manual_seed(seed)


# ---------------------------------------------------------------------------------------------
# wide states: several operand chunks, several owned column blocks per thread, slots in global scratch
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D_,C_,units", [(40, 3, [96, 128]), (70, 0, [128]), (33, 37, [64, 64, 64]), (3, 0, [])])
def test_wide_pfode_dopri5_and_fixed(cuda_dev, D_, C_, units):
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(100 + D_)
    sm = D.ScoreModel(D.MLP(D_, C_, 8, units), D.VESDE(), no_sigma=False).eval()
    B = 300
    base = torch.randn(B, D_, generator=gen(1)); cond = torch.randn(B, C_, generator=gen(2)) if C_ else None
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("ve"), False)
    ref = port.sample_ode_from_base(M, base, cond, 1e-5, 1e-5)[0]
    rs = port.last_stats()
    ref4 = port.sample_ode_from_base(M, base, cond, method="rk4", options={"step_size": 1 / 8})[0]
    refm = port.sample_ode_from_base(M, base, cond, method="midpoint", options={"step_size": 1 / 8})[0]
    sm.to(cuda_dev)
    cb = None if cond is None else cond.to(cuda_dev)
    x, _ = sm.sample_ode_from_base(base.to(cuda_dev), cb, atol=1e-5, rtol=1e-5)
    assert rel_row_err(ref, x) < SAMPLE_TOL
    assert (sm.last_stats.accepted, sm.last_stats.rejected) == (rs.accepted, rs.rejected)
    x4, _ = sm.sample_ode_from_base(base.to(cuda_dev), cb, method="rk4", options={"step_size": 1 / 8})
    assert rel_row_err(ref4, x4) < SAMPLE_TOL
    xm, _ = sm.sample_ode_from_base(base.to(cuda_dev), cb, method="midpoint", options={"step_size": 1 / 8})
    assert rel_row_err(refm, xm) < SAMPLE_TOL
    with engine(0):                                      # the FP32 FFMA2 kernels agree with the 3xTF32 ones
        xf, _ = sm.sample_ode_from_base(base.to(cuda_dev), cb, method="rk4", options={"step_size": 1 / 8})
    assert rel_row_err(xf, x4) < 2e-5


def test_wide_euler_maruyama(cuda_dev):
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(5)
    sm = D.ScoreModel(D.MLP(48, 5, 8, [128, 96]), D.VPSDE(), no_sigma=True).eval()
    B, steps = 200, 40
    x0 = torch.randn(B, 48, generator=gen(3)); dw = torch.randn(steps, B, 48, generator=gen(4))
    cond = torch.randn(B, 5, generator=gen(5))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    ref = port.sample_sde(M, x0, dw, cond)
    sm.to(cuda_dev)
    x = sm.sample_sde((B, 48), conditional=cond.to(cuda_dev), steps=steps, x0=x0.to(cuda_dev), noise=dw.to(cuda_dev))
    assert rel_row_err(ref, x) < SAMPLE_TOL


# ---------------------------------------------------------------------------------------------
# tangent tiles: few samples per tile (large D), many (Hutchinson), conditional flows
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D_,hidden", [(40, [64, 128]), (7, [96]), (31, [128, 128]), (2, [64, 64, 64])])
def test_flow_logprob_exact_various_widths(cuda_dev, D_, hidden):
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(200 + D_)
    m = F.ODEFlow(D_, hidden).eval()
    x = torch.randn(97, D_, generator=gen(6))
    ref = port.flow_log_prob(port.flow_from_state_dict(m.state_dict()), x)
    rs = port.last_stats()
    m.to(cuda_dev)
    lp = m.log_prob(x.to(cuda_dev))
    assert float((lp.cpu() - ref).abs().max()) < LP_TOL
    assert (m.last_stats.accepted, m.last_stats.rejected) == (rs.accepted, rs.rejected)
    with engine(2):                                      # whole-layer hand-off tile engine: same answer
        lp2 = m.log_prob(x.to(cuda_dev))
    assert float((lp2 - lp).abs().max()) < 2e-4


def test_conditional_flow_logprob_wide_cond(cuda_dev):
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(9)
    m = F.ConditionalODEFlow(12, 20, [128, 64]).eval()
    x = torch.randn(150, 12, generator=gen(7)); c = torch.randn(150, 20, generator=gen(8))
    ref = port.flow_log_prob(port.flow_from_state_dict(m.state_dict()), x, c)
    rs = port.last_stats()
    m.to(cuda_dev)
    lp = m.log_prob(x.to(cuda_dev), c.to(cuda_dev))
    assert float((lp.cpu() - ref).abs().max()) < LP_TOL
    assert (m.last_stats.accepted, m.last_stats.rejected) == (rs.accepted, rs.rejected)


def test_score_logprob_hutchinson_wide(cuda_dev):
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(11)
    sm = D.ScoreModel(D.MLP(36, 2, 8, [128, 128]), D.VPSDE(), no_sigma=True, hutchinson=True).eval()
    x0 = torch.randn(130, 36, generator=gen(9)); cond = torch.randn(130, 2, generator=gen(10))
    e = torch.sign(torch.randn(130, 36, generator=gen(11)))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    ref = port.score_log_prob(M, x0, cond, probes=e)
    rs = port.last_stats()
    sm.to(cuda_dev)
    lp = sm.log_prob(x0.to(cuda_dev), cond.to(cuda_dev), probes=e.to(cuda_dev))
    assert float((lp.cpu() - ref).abs().max()) < LP_TOL
    assert (sm.last_stats.accepted, sm.last_stats.rejected) == (rs.accepted, rs.rejected)


# ---------------------------------------------------------------------------------------------
# the full BASELINE sizes: against the CPU oracle (the port integrates 1 M cfg2 rows in ~30 s on the box's host cores)
# and through size-independent properties
# ---------------------------------------------------------------------------------------------
def _hist_rel_dev(a, b):
    """Largest deviation of the accepted-or-attempted step END TIMES (cumulative dt), relative to the whole interval.
    (The last step is whatever is left of the interval, so its own relative deviation is not a meaningful number.)"""
    assert len(a) == len(b), (len(a), len(b))
    ta, tb, dev = 0.0, 0.0, 0.0
    for x, y in zip(a, b):
        ta += x; tb += y
        dev = max(dev, abs(ta - tb))
    return dev / max(abs(tb), 1e-30)


def test_cfg2_full_size_vs_oracle(cuda_dev):
    """BASELINE configs[1] at its full size (1 M rows, dopri5 1e-5, the bench workload) against oracle/port.py on the same
    inputs: samples <= 1e-4 relative per row, identical (accepted, rejected), the whole dt history.  This is where the
    FP64-partials-vs-torch-sum question of the global error norm (16 M elements) is decided."""
    D, F, Sy = _mods()
    from oracle import port
    from flowfusion_b200 import solver
    torch.manual_seed(1234)
    sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 4), D.VPSDE(), no_sigma=True).eval()
    B = 1_000_000
    base = torch.randn(B, 16, generator=gen(2)); cond = torch.randn(B, 4, generator=gen(3))
    opts = {"step_t": torch.tensor([1e-3])}
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    ref = port.sample_ode_from_base(M, base, cond, 1e-5, 1e-5, options=opts)[0]
    rs = port.last_stats()
    sm.to(cuda_dev)
    for ctl in ("device", "host"):
        with solver.controller(ctl):
            x, _ = sm.sample_ode_from_base(base.to(cuda_dev), cond.to(cuda_dev), atol=1e-5, rtol=1e-5, options=opts)
        st = sm.last_stats
        err = rel_row_err(ref, x)
        dev_dt = _hist_rel_dev(list(st.dt_history), list(rs.dt_history))
        print(f"cfg2 1M rows [{ctl}]: sample err {err:.2e}, steps {st.accepted}/{st.rejected} (cpu {rs.accepted}/{rs.rejected}), "
              f"max step-time deviation {dev_dt:.2e}")
        assert err < SAMPLE_TOL
        assert (st.accepted, st.rejected) == (rs.accepted, rs.rejected)
        assert list(st.accept_history) == list(rs.accept_history)
        # The step sizes follow ratio^(-1/5), and the error estimate sum_j c_err_j k_j is a ~1e-5 cancellation of O(1) stage
        # derivatives: its 3xTF32 and FP32-FMA roundings differ by ~1e-3 relative, so the step end times agree to a few
        # 1e-4 of the interval (measured on B200 at 1 M rows), not to FP64 resolution.  The accept / reject sequence is identical.
        assert dev_dt < 1e-3


def test_cfg3_100k_vs_oracle(cuda_dev):
    """BASELINE configs[2] architecture on 100 k points (exact divergence trace) against oracle/port.py: log-prob <= 1e-3
    nat, identical (accepted, rejected), dt history."""
    D, F, Sy = _mods()
    from oracle import port
    from flowfusion_b200 import solver
    torch.manual_seed(1234)
    fl = F.ODEFlow(16, [128] * 4).eval()
    B = 100_000
    xs = torch.randn(B, 16, generator=gen(4))
    ref = port.flow_log_prob(port.flow_from_state_dict(fl.state_dict()), xs)
    rs = port.last_stats()
    fl.to(cuda_dev)
    for ctl in ("device", "host"):
        with solver.controller(ctl):
            lp = fl.log_prob(xs.to(cuda_dev))
        st = fl.last_stats
        err = float((lp.cpu() - ref).abs().max())
        dev_dt = _hist_rel_dev(list(st.dt_history), list(rs.dt_history))
        print(f"cfg3 100k rows [{ctl}]: log-prob err {err:.2e} nat, steps {st.accepted}/{st.rejected} (cpu {rs.accepted}/{rs.rejected}), "
              f"max step-time deviation {dev_dt:.2e}")
        assert err < LP_TOL
        assert (st.accepted, st.rejected) == (rs.accepted, rs.rejected)
        # With these weights the local error of the first steps is ~1e-9, far below the FP32 resolution of the state: the
        # error ratio (measured 1e-6, 1.5e-5, 8e-2 against the port's 1.4e-6, 1.8e-5, 7e-2) is rounding noise on both sides, and
        # 0.9 ratio^(-1/5) turns a 35 % difference of it into a 6 % longer step.  The step end times therefore only agree to a few per cent (measured 3 %); what is pinned is the
        # accept / reject sequence and the result.
        print("   error ratios gpu", [f"{r:.2e}" for r in st.ratio_history], "cpu", [f"{r:.2e}" for r in rs.ratio_history])
        assert dev_dt < 0.25



def test_cfg2_full_size_partition_invariance(cuda_dev):
    """1M rows, fixed grid: any row's result is independent of the batch it is integrated in."""
    D, F, Sy = _mods()
    torch.manual_seed(1234)
    sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 4), D.VPSDE(), no_sigma=True).eval().to(cuda_dev)
    B = 1_000_000
    base = torch.randn(B, 16, generator=gen(2)).to(cuda_dev); cond = torch.randn(B, 4, generator=gen(3)).to(cuda_dev)
    opt = {"step_size": 1 / 4}
    full, _ = sm.sample_ode_from_base(base, cond, method="rk4", options=opt)
    assert torch.isfinite(full).all()
    for lo, hi in ((0, 1000), (499_937, 500_070), (B - 777, B)):
        part, _ = sm.sample_ode_from_base(base[lo:hi], cond[lo:hi], method="rk4", options=opt)
        assert torch.equal(full[lo:hi], part)
    again, _ = sm.sample_ode_from_base(base, cond, method="rk4", options=opt)
    assert torch.equal(full, again)


def test_cfg4_em_noise_linearity_full_width(cuda_dev):
    """Euler-Maruyama with a zero-weight network is linear in (x0, noise): x_T = A x0 + sum_k B_k dw_k.
    Checks the in-kernel noise path / step table at cfg4's width without an oracle."""
    D, F, Sy = _mods()
    torch.manual_seed(3)
    net = D.MLP(32, 0, 8, [128] * 4)
    with torch.no_grad():
        for p in net.NN.parameters():
            p.zero_()
    sm = D.ScoreModel(net, D.VPSDE(), no_sigma=True).eval().to(cuda_dev)
    B, steps = 4096, 50
    x0a = torch.randn(B, 32, generator=gen(1)).to(cuda_dev); x0b = torch.randn(B, 32, generator=gen(2)).to(cuda_dev)
    dwa = torch.randn(steps, B, 32, generator=gen(3)).to(cuda_dev); dwb = torch.randn(steps, B, 32, generator=gen(4)).to(cuda_dev)
    f = lambda x0, dw: sm.sample_sde((B, 32), steps=steps, x0=x0, noise=dw)      # noqa: E731
    lhs = f(x0a + x0b, dwa + dwb)
    rhs = f(x0a, dwa) + f(x0b, dwb)
    assert rel_row_err(rhs, lhs) < 1e-5


def test_cfg3_logprob_partition_invariance_fixed_grid(cuda_dev):
    """Exact-trace log-prob on a fixed grid (no global coupling): rows are independent of their tile."""
    D, F, Sy = _mods()
    torch.manual_seed(1234)
    m = F.ODEFlow(16, [128] * 4).eval().to(cuda_dev)
    x = torch.randn(20_000, 16, generator=gen(4)).to(cuda_dev)
    lp = m.log_prob(x, method="rk4", options={"step_size": 0.25})
    assert torch.isfinite(lp).all()
    lp2 = m.log_prob(x[123:4000], method="rk4", options={"step_size": 0.25})
    assert torch.equal(lp[123:4000], lp2)


def test_population_wrappers_row_a10(cuda_dev):
    """PopulationModelDiffusionConditional.forward / .log_prob / .sample_sde (quirks Q6-Q8) vs the oracle port."""
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(77)
    shift, scale = torch.linspace(-1, 1, 5), torch.linspace(0.5, 2.0, 5)
    cshift, cscale = torch.tensor([0.3, -0.2]), torch.tensor([1.5, 0.7])
    pm = D.PopulationModelDiffusionConditional(D.MLP(5, 2, 6, [32, 48]), D.VESDE(), shift, scale, cshift, cscale).eval()
    base = torch.randn(200, 5, generator=gen(1)) * 10.0
    cond = torch.randn(200, 2, generator=gen(2))
    M = port.score_model_from_state_dict(pm.score_model.state_dict(), port.make_sde("ve"), False)
    ref = port.population_forward(M, base, shift, scale, cond, cshift, cscale)
    ref_lp = port.population_log_prob(M, ref[:64], shift, scale, cond[:64], cshift, cscale)
    pm.to(cuda_dev)
    x = pm.forward(base.to(cuda_dev), cond.to(cuda_dev))
    assert rel_row_err(ref, x) < SAMPLE_TOL
    lp = pm.log_prob(ref[:64].to(cuda_dev), cond[:64].to(cuda_dev))
    assert lp.shape == (64, 1)
    # random-init VE score with sigma division: |log p| ~ 60..440 nats, one FP32 ulp there is 3e-5 nats, so the
    # 1e-3 nat tolerance gets a relative part (2e-5 |log p| = under one ulp per unit of |log p| / 1.5)
    assert bool(((lp.cpu() - ref_lp).abs() < LP_TOL + 2e-5 * ref_lp.abs()).all())
    # Q6: the wrapper ignores `steps` (always 100)
    x0 = torch.randn(50, 5, generator=gen(3)).to(cuda_dev) * 10.0
    dw = torch.randn(100, 50, 5, generator=gen(4)).to(cuda_dev)
    a = pm.sample_sde((50, 5), conditional=cond[:50].to(cuda_dev), steps=7, x0=x0, noise=dw)
    b = pm.sample_sde((50, 5), conditional=cond[:50].to(cuda_dev), steps=100, x0=x0, noise=dw)
    assert torch.equal(a, b)


def test_maximum_sizes(cuda_dev):
    """The limits of the tensor-core engines: 128 state columns, 8 Linear layers of width 128 (wider / deeper networks run on
    the wide engine, tests/test_gpu_wide.py), and the error paths of include/ffb200.h's limits."""
    D, F, Sy = _mods()
    from oracle import port
    from flowfusion_b200 import _lib
    torch.manual_seed(21)
    sm = D.ScoreModel(D.MLP(128, 0, 8, [128] * 7), D.VESDE(), no_sigma=True).eval()
    base = torch.randn(150, 128, generator=gen(1))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("ve"), True)
    ref = port.sample_ode_from_base(M, base, None, method="rk4", options={"step_size": 0.25})[0]
    refd = port.sample_ode_from_base(M, base, None, 1e-4, 1e-4)[0]
    rs = port.last_stats()
    sm.to(cuda_dev)
    x, _ = sm.sample_ode_from_base(base.to(cuda_dev), method="rk4", options={"step_size": 0.25})
    assert rel_row_err(ref, x) < SAMPLE_TOL
    xd, _ = sm.sample_ode_from_base(base.to(cuda_dev), atol=1e-4, rtol=1e-4)
    assert rel_row_err(refd, xd) < SAMPLE_TOL
    assert (sm.last_stats.accepted, sm.last_stats.rejected) == (rs.accepted, rs.rejected)
    # one layer too many / one column too wide: refused with NotImplementedError, nothing is launched
    with pytest.raises(NotImplementedError):
        D.ScoreModel(D.MLP(4, 0, 8, [32] * 16), D.VESDE()).eval().to(cuda_dev).sample_ode_from_base(torch.zeros(4, 4, device=cuda_dev))
    with pytest.raises(NotImplementedError):
        D.ScoreModel(D.MLP(4, 0, 8, [513]), D.VESDE()).eval().to(cuda_dev).sample_ode_from_base(torch.zeros(4, 4, device=cuda_dev))
    # a CPU tensor is refused: there is no CPU path
    with pytest.raises(_lib.FFBError):
        sm.sample_ode_from_base(base)


@pytest.mark.parametrize("act_cls,act_fn", [(torch.nn.Tanh, torch.tanh), (torch.nn.ReLU, torch.relu),
                                            (torch.nn.Softplus, torch.nn.functional.softplus),
                                            (torch.nn.GELU, torch.nn.functional.gelu)])
def test_non_silu_activations(cuda_dev, act_cls, act_fn):
    """`activation=` is a pass-through constructor argument of every reference model."""
    D, F, Sy = _mods()
    from oracle import port
    from flowfusion_b200 import _lib
    torch.manual_seed(31)
    sm = D.ScoreModel(D.MLP(16, 4, 8, [128, 96], activation=act_cls()), D.VPSDE(), no_sigma=True).eval()
    base = torch.randn(300, 16, generator=gen(1)); cond = torch.randn(300, 4, generator=gen(2))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True, act=act_fn)
    opts = {"step_t": torch.tensor([1e-3])}
    ref = port.sample_ode_from_base(M, base, cond, 1e-5, 1e-5, options=opts)[0]
    rs = port.last_stats()
    ref4 = port.sample_ode_from_base(M, base, cond, method="rk4", options={"step_size": 1 / 8})[0]
    sm.to(cuda_dev)
    x, _ = sm.sample_ode_from_base(base.to(cuda_dev), cond.to(cuda_dev), atol=1e-5, rtol=1e-5, options=opts)
    assert rel_row_err(ref, x) < SAMPLE_TOL
    assert (sm.last_stats.accepted, sm.last_stats.rejected) == (rs.accepted, rs.rejected)
    x4, _ = sm.sample_ode_from_base(base.to(cuda_dev), cond.to(cuda_dev), method="rk4", options={"step_size": 1 / 8})
    assert rel_row_err(ref4, x4) < SAMPLE_TOL
    # exact-trace log-prob: the tangent rows use the activation's derivative
    torch.manual_seed(32)
    fl = F.ODEFlow(9, [64, 128], activation=act_cls).eval()
    xs = torch.randn(90, 9, generator=gen(3))
    ref_lp = port.flow_log_prob(port.flow_from_state_dict(fl.state_dict(), act=act_fn), xs)
    rs = port.last_stats()
    fl.to(cuda_dev)
    lp = fl.log_prob(xs.to(cuda_dev))
    assert float((lp.cpu() - ref_lp).abs().max()) < LP_TOL
    if act_cls is not torch.nn.ReLU:      # ReLU's kinks make the step sequence sensitive to FP32 summation order
        assert (fl.last_stats.accepted, fl.last_stats.rejected) == (rs.accepted, rs.rejected)
    # fixed grid with the divergence (tangent-row engine), vs the port
    ref_rk = port.flow_log_prob(port.flow_from_state_dict(fl.state_dict(), act=act_fn), xs, method="rk4",
                                options={"step_size": 0.25})
    lp_rk = fl.log_prob(xs.to(cuda_dev), method="rk4", options={"step_size": 0.25})
    assert float((lp_rk.cpu() - ref_rk).abs().max()) < LP_TOL
    # the SiLU-only engines refuse the network instead of silently using SiLU
    with engine(0):
        with pytest.raises(_lib.FFBError):
            fl.log_prob(xs.to(cuda_dev))


@pytest.mark.parametrize("method,opts", [("euler", {"step_size": 1 / 16}), ("midpoint", {"step_size": 1 / 8}), ("rk4", {"step_size": 1 / 4})])
def test_fixed_grid_logprob_on_tangent_engine(cuda_dev, method, opts):
    """Fixed-grid solves of (x, log-det): k_fixed_rrt vs the port, vs the tile engine, exact and Hutchinson."""
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(41)
    fl = F.ConditionalODEFlow(10, 3, [96, 128]).eval()
    xs = torch.randn(77, 10, generator=gen(1)); c = torch.randn(77, 3, generator=gen(2))
    ref = port.flow_log_prob(port.flow_from_state_dict(fl.state_dict()), xs, c, method=method, options=opts)
    fl.to(cuda_dev)
    lp = fl.log_prob(xs.to(cuda_dev), c.to(cuda_dev), method=method, options=opts)
    assert float((lp.cpu() - ref).abs().max()) < LP_TOL
    with engine(2):
        lp2 = fl.log_prob(xs.to(cuda_dev), c.to(cuda_dev), method=method, options=opts)
    assert float((lp2 - lp).abs().max()) < 2e-4
    torch.manual_seed(42)
    sm = D.ScoreModel(D.MLP(6, 0, 8, [64, 64]), D.VPSDE(), no_sigma=True, hutchinson=True).eval()
    x0 = torch.randn(50, 6, generator=gen(3)); e = torch.sign(torch.randn(50, 6, generator=gen(4)))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    refh = port.score_log_prob(M, x0, None, method=method, options=opts, probes=e)
    sm.to(cuda_dev)
    lph = sm.log_prob(x0.to(cuda_dev), method=method, options=opts, probes=e.to(cuda_dev))
    assert float((lph.cpu() - refh).abs().max()) < LP_TOL


# ---------------------------------------------------------------------------------------------
# dual-tile engine (csrc/ffb_engine_rd.cuh): forced with engine(4) so that small batches run on it too
# ---------------------------------------------------------------------------------------------
def test_dual_tile_engine_golden_and_bit_identical_to_single_tile(cuda_dev):
    """The dual-tile dopri5 kernel against the reference's golden vectors (cfg2: same accepted / rejected steps), against
    the single-tile engine BIT FOR BIT (same MMAs in the same order; ragged batch, odd tile count -> dummy tiles, clusters),
    with both controllers, and for a non-SiLU network, a two-call (symplectic) field and an unconditional VE model."""
    from conftest import load_golden
    from flowfusion_b200 import solver
    D, F, Sy = _mods()
    meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
    sm = D.ScoreModel(D.MLP(**meta["ctor"]), D.VPSDE(), no_sigma=True).eval()
    sm.load_state_dict(sd)
    sm.to(cuda_dev)
    opts = {"step_t": torch.tensor([1e-3])}
    base, cond = ins["base"].to(cuda_dev), ins["cond"].to(cuda_dev)
    with engine(4):
        x, _ = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options=opts)
        st = sm.last_stats
    assert rel_row_err(outs["x_dopri5"], x) < SAMPLE_TOL
    assert (st.accepted, st.rejected) == (meta["stats"]["accepted"], meta["stats"]["rejected"])
    # bit-identical to the single-tile engine, every controller, ragged sizes (1, 3 and 19 tiles: the last CTA / cluster
    # runs dummy tiles), larger than one round of the 148 SMs (600 tiles)
    for B in (100, 300, 2400, 76_800 + 77):
        b = torch.randn(B, 16, generator=gen(B), device="cpu").to(cuda_dev)
        c = torch.randn(B, 4, generator=gen(B + 1), device="cpu").to(cuda_dev)
        for ctl in ("host", "device"):
            with solver.controller(ctl):
                with engine(3):
                    x3, _ = sm.sample_ode_from_base(b, c, atol=1e-5, rtol=1e-5, options=opts)
                    s3 = sm.last_stats
                with engine(4):
                    x4, _ = sm.sample_ode_from_base(b, c, atol=1e-5, rtol=1e-5, options=opts)
                    s4 = sm.last_stats
            assert torch.equal(x3, x4), (B, ctl, float((x3 - x4).abs().max()))
            assert (s3.accepted, s3.rejected) == (s4.accepted, s4.rejected)
            assert s3.dt_history == s4.dt_history
    # non-SiLU activation (run-time dispatched epilogue), unconditional VE with the sigma division
    torch.manual_seed(41)
    sv = D.ScoreModel(D.MLP(7, 0, 8, [96, 64], activation=torch.nn.Tanh()), D.VESDE(), no_sigma=False).eval().to(cuda_dev)
    bv = (torch.randn(700, 7, generator=gen(5)) * 10.0).to(cuda_dev)
    with engine(3):
        x3, _ = sv.sample_ode_from_base(bv, atol=1e-5, rtol=1e-5)
    with engine(4):
        x4, _ = sv.sample_ode_from_base(bv, atol=1e-5, rtol=1e-5)
    assert torch.equal(x3, x4)
    # two-network field (symplectic log_prob: dopri5 on the (q | p) state, 8-D phase space)
    torch.manual_seed(42)
    sf = Sy.SymplecticFlowModel(Sy.SymplecticMLP(4, 0, 8, [64, 64]), torch.zeros(4), torch.ones(4), torch.zeros(0), torch.ones(0)).eval().to(cuda_dev)
    xq = torch.randn(500, 4, generator=gen(6)).to(cuda_dev)
    p0 = torch.randn(500, 4, generator=gen(7)).to(cuda_dev)
    outs_ = []
    for e in (3, 4):
        with engine(e):
            lp = sf.log_prob(xq, p0=p0)
            outs_.append((lp.clone(), sf.last_stats.accepted, sf.last_stats.rejected))
    assert torch.equal(outs_[0][0], outs_[1][0]) and outs_[0][1:] == outs_[1][1:]


# ---------------------------------------------------------------------------------------------
# accelerate(): the documented drop-in entry point must keep the reference object's hidden activation
# ---------------------------------------------------------------------------------------------
def _reference_like_flow(meta, sd, act_cls, conditional):
    """A stand-in with the attribute surface of the reference's ODEFlow / ConditionalODEFlow (`flow.py:37-86, 473-551`):
    `/root/reference` does not exist on the GPU box, the golden vectors below were produced there by the real class."""
    c = meta["ctor"]
    D_, hidden = c["target_dimension"], c["hidden_units"]
    Cn = c.get("conditional_dimension", 0)

    class _Flow(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.target_dimension = D_
            if conditional:
                self.conditional_dimension = Cn
            self.layers = torch.nn.ModuleList()
            arch = [D_ + 1 + Cn] + list(hidden) + [D_]
            for i in range(len(arch) - 2):
                self.layers.append(torch.nn.Linear(arch[i], arch[i + 1]))
                self.layers.append(act_cls())
            self.layers.append(torch.nn.Linear(arch[-2], arch[-1]))
            self.velocity = torch.nn.Sequential(*self.layers)
            for k, v in sd.items():
                if k not in self.state_dict():
                    self.register_buffer(k, torch.zeros_like(v))

    _Flow.__name__ = "ConditionalODEFlow" if conditional else "ODEFlow"
    m = _Flow().eval()
    m.load_state_dict(sd)
    return m


@pytest.mark.parametrize("name,act_cls,conditional", [("flow_tanh", torch.nn.Tanh, False), ("cflow_gelu", torch.nn.GELU, True)])
def test_accelerate_non_silu_flow_matches_reference_golden(cuda_dev, name, act_cls, conditional):
    """`flowfusion_b200.accelerate(ref_obj)` on a Tanh / GELU flow, then sample and log_prob against golden vectors
    produced by the UNMODIFIED reference (oracle/make_golden_act.py)."""
    from conftest import load_golden
    import flowfusion_b200 as ffb
    meta, sd, ins, outs = load_golden(name)
    ref_like = _reference_like_flow(meta, sd, act_cls, conditional)
    tw = ffb.accelerate(ref_like, device=cuda_dev)
    assert [type(m).__name__ for m in tw.layers] == [type(m).__name__ for m in ref_like.layers]
    args = (ins["cond"].to(cuda_dev),) if conditional else ()
    x = tw.sample(ins["xT"].to(cuda_dev), *args)
    assert rel_row_err(outs["x"], x) < SAMPLE_TOL
    lp = tw.log_prob(outs["x"].to(cuda_dev), *args, atol=1e-6, rtol=1e-6)
    assert float((lp.cpu() - outs["log_prob"]).abs().max()) < LP_TOL
    st = meta["stats_logprob"]
    assert (tw.last_stats.accepted, tw.last_stats.rejected) == (st["accepted"], st["rejected"])


# ---------------------------------------------------------------------------------------------
# more than one device in the process (these need 2 GPUs: `gpurun --gpus 2`)
# ---------------------------------------------------------------------------------------------
def test_model_on_a_device_that_is_not_current(cuda_dev):
    """The reference works with a model on cuda:1 while cuda:0 is current; the C ABI launches on the current device with
    raw pointers, so every entry point switches to the device of its state (engine.on_device)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    D, F, Sy = _mods()
    from flowfusion_b200 import _lib
    torch.manual_seed(5)
    sm = D.ScoreModel(D.MLP(6, 2, 8, [64, 64]), D.VPSDE(), no_sigma=True).eval()
    base = torch.randn(300, 6, generator=gen(1)); cond = torch.randn(300, 2, generator=gen(2))
    opts = {"step_t": torch.tensor([1e-3])}
    outs_ = []
    for d in ("cuda:0", "cuda:1"):
        sm.to(d)
        torch.cuda.set_device(0)                                    # cuda:0 stays current
        x, _ = sm.sample_ode_from_base(base.to(d), cond.to(d), atol=1e-5, rtol=1e-5, options=opts)
        x4, _ = sm.sample_ode_from_base(base.to(d), cond.to(d), method="rk4", options={"step_size": 0.25})
        xs = sm.sample_sde((300, 6), cond.to(d), steps=20, x0=base.to(d), noise=torch.randn(20, 300, 6, generator=gen(3)).to(d))
        lp = sm.log_prob(base.to(d), cond.to(d))
        assert x.device == torch.device(d)
        outs_.append([t.cpu() for t in (x, x4, xs, lp)])
    for a, b in zip(*outs_):
        assert torch.equal(a, b)
    with pytest.raises(_lib.FFBError):                              # weights on cuda:1, state on cuda:0
        sm.sample_ode_from_base(base.to("cuda:0"), cond.to("cuda:0"), atol=1e-5, rtol=1e-5, options=opts)


# ---------------------------------------------------------------------------------------------
# repeatability stress (compute-sanitizer is closed on the GPU pool, profiles/r02_sanitizer_unavailable.txt): a data race in
# the mbarrier / tensor-memory hand-offs of any engine shows as a bit difference between runs on the same inputs
# ---------------------------------------------------------------------------------------------
def test_repeatability_stress(cuda_dev):
    D, F, Sy = _mods()
    from flowfusion_b200 import solver
    torch.manual_seed(1234)
    sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 4), D.VPSDE(), no_sigma=True).eval().to(cuda_dev)
    B = 148 * 128 * 2 + 333                                       # three rounds of the dual-tile engine, ragged last tile
    base = torch.randn(B, 16, generator=gen(2)).to(cuda_dev); cond = torch.randn(B, 4, generator=gen(3)).to(cuda_dev)
    opts = {"step_t": torch.tensor([1e-3])}
    runs = {}
    for rep in range(12):
        for e, ctl in ((4, "device"), (3, "device"), (4, "host")):
            with engine(e), solver.controller(ctl):
                x, _ = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options=opts)
            runs.setdefault("dopri5_" + ctl, x.clone())         # (the two controllers differ in the last place of sqrt / sin / cos)
            assert torch.equal(runs["dopri5_" + ctl], x), (rep, e, ctl)
        x4, _ = sm.sample_ode_from_base(base, cond, method="rk4", options={"step_size": 0.25})
        runs.setdefault("rk4", x4.clone())
        assert torch.equal(runs["rk4"], x4), rep
        xe = sm.sample_sde((B, 16), cond, steps=8, x0=base, seed=7)
        runs.setdefault("em", xe.clone())
        assert torch.equal(runs["em"], xe), rep
    torch.manual_seed(1234)
    fl = F.ODEFlow(16, [128] * 4).eval().to(cuda_dev)
    xs = torch.randn(148 * 7 + 5, 16, generator=gen(4)).to(cuda_dev)
    sh = D.ScoreModel(D.MLP(16, 0, 8, [128] * 2), D.VPSDE(), no_sigma=True, hutchinson=True).eval().to(cuda_dev)
    probes = torch.sign(torch.randn(xs.shape[0], 16, generator=gen(5))).to(cuda_dev)
    for rep in range(12):
        lp = fl.log_prob(xs)
        runs.setdefault("lp", lp.clone())
        assert torch.equal(runs["lp"], lp), rep
        lh = sh.log_prob(xs, probes=probes)
        runs.setdefault("lh", lh.clone())
        assert torch.equal(runs["lh"], lh), rep


# ---------------------------------------------------------------------------------------------
# the other adaptive methods of torchdiffeq (method= is a pass-through argument of the reference's entry points)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["bosh3", "adaptive_heun", "fehlberg2"])
def test_other_adaptive_methods(cuda_dev, method):
    D, F, Sy = _mods()
    from oracle import port
    tol = dict(bosh3=1e-4, adaptive_heun=1e-3, fehlberg2=1e-4)[method]
    torch.manual_seed(21)
    sm = D.ScoreModel(D.MLP(16, 4, 8, [128, 128]), D.VPSDE(), no_sigma=True).eval()
    base = torch.randn(300, 16, generator=gen(1)); cond = torch.randn(300, 4, generator=gen(2))
    opts = {"step_t": torch.tensor([1e-3])}
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    ref = port.sample_ode_from_base(M, base, cond, tol, tol, method=method, options=opts)[0]
    rs = port.last_stats()
    sm.to(cuda_dev)
    x, _ = sm.sample_ode_from_base(base.to(cuda_dev), cond.to(cuda_dev), atol=tol, rtol=tol, method=method, options=opts)
    assert rel_row_err(ref, x) < SAMPLE_TOL
    assert (sm.last_stats.accepted, sm.last_stats.rejected, sm.last_stats.nfe) == (rs.accepted, rs.rejected, rs.nfe)
    # exact-trace log-prob (tangent-row engine, the log-det column is part of the error norm) and a conditional flow
    torch.manual_seed(22)
    fl = F.ConditionalODEFlow(6, 3, [64, 96]).eval()
    xs = torch.randn(150, 6, generator=gen(3)); c = torch.randn(150, 3, generator=gen(4))
    ref_lp = port.flow_log_prob(port.flow_from_state_dict(fl.state_dict()), xs, c, atol=tol, rtol=tol, method=method)
    rs = port.last_stats()
    fl.to(cuda_dev)
    lp = fl.log_prob(xs.to(cuda_dev), c.to(cuda_dev), atol=tol, rtol=tol, method=method)
    assert float((lp.cpu() - ref_lp).abs().max()) < LP_TOL
    assert (fl.last_stats.accepted, fl.last_stats.rejected) == (rs.accepted, rs.rejected)
    with pytest.raises(NotImplementedError):
        sm.sample_ode_from_base(base.to(cuda_dev), cond.to(cuda_dev), method="dopri8")


def test_solver_options_perturb_grid_constructor_jump_t(cuda_dev):
    """Pass-through `options=` of the reference's entry points beyond step_size / step_t: fixed grids with perturb and a
    grid_constructor, dopri5 with jump_t (host controller: f is re-evaluated after each discontinuity)."""
    D, F, Sy = _mods()
    from oracle import port
    torch.manual_seed(31)
    sm = D.ScoreModel(D.MLP(8, 2, 8, [64, 64]), D.VPSDE(), no_sigma=True).eval()
    base = torch.randn(200, 8, generator=gen(1)); cond = torch.randn(200, 2, generator=gen(2))
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    gc = lambda func, y0, t: torch.tensor([float(t[0]), -0.8, -0.55, -0.3, -0.07, float(t[-1])])      # noqa: E731
    refs = {}
    for method in ("euler", "midpoint", "rk4"):
        for i, opts in enumerate(({"step_size": 0.125, "perturb": True}, {"grid_constructor": gc, "perturb": True})):
            refs[(method, i)] = (opts, port.sample_ode_from_base(M, base, cond, method=method, options=opts)[0])
    jopts = {"step_t": torch.tensor([1e-3]), "jump_t": torch.tensor([0.6, 0.25])}
    jref = port.sample_ode_from_base(M, base, cond, 1e-5, 1e-5, options=jopts)[0]
    js = port.last_stats()
    sm.to(cuda_dev)
    b, c = base.to(cuda_dev), cond.to(cuda_dev)
    for (method, i), (opts, ref) in refs.items():
        x, _ = sm.sample_ode_from_base(b, c, method=method, options=opts)
        assert rel_row_err(ref, x) < SAMPLE_TOL, (method, i)
    x, _ = sm.sample_ode_from_base(b, c, atol=1e-5, rtol=1e-5, options=jopts)
    assert rel_row_err(jref, x) < SAMPLE_TOL
    assert (sm.last_stats.accepted, sm.last_stats.rejected, sm.last_stats.nfe) == (js.accepted, js.rejected, js.nfe)
