"""The dopri5 controller without a GPU: csrc/ffb_control.cuh compiles for the host as well
(ffb_dopri5_control_host, ffb_time_program_rows), so the statements the control kernel executes between two
attempts are checked here against the Python loop of flowfusion_b200/solver.py (itself pinned by the golden
vectors of the unmodified reference) and against the scalar programs of the model classes."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_row_err
from kernel_model import patched_engine

import flowfusion_b200.diffusion as D
import flowfusion_b200.flow as F
import flowfusion_b200.symplectic as Sy
from flowfusion_b200 import _lib as L
from flowfusion_b200 import solver as S


def _ulps(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def _twin_rows(spec, times, sign=1.0, on_device=0):
    lib = L.load()
    times = np.ascontiguousarray(times, np.float32)
    out = (L.EvalScalars * len(times))()
    L.check(lib.ffb_time_program_rows(C.byref(spec), times.ctypes.data, len(times), float(sign), out, on_device),
            "ffb_time_program_rows")
    return np.frombuffer(bytes(out), np.float32).reshape(len(times), L.EV_FLOATS).copy()


def _programs():
    torch.manual_seed(5)
    net = D.MLP(6, 2, 8, [16])
    out = []
    for name, sde, no_sigma in (("vp", D.VPSDE(), True), ("ve_sigma", D.VESDE(), False), ("ve", D.VESDE(), True),
                                ("vp_b", D.VPSDE(beta_min=0.3, beta_max=11.0), True)):
        out.append((name, D.ScoreModel(net, sde, no_sigma=no_sigma)._program()))
        out.append((name + "_sde", D.ScoreModel(net, sde, no_sigma=no_sigma)._program(sde_mode=True)))
    out.append(("flow", F._raw_time_program))
    out.append(("symplectic", Sy.SymplecticMLP(4, 0, 8, [16])._program()))
    return out


@pytest.mark.parametrize("name,prog", _programs(), ids=[n for n, _ in _programs()])
def test_program_twin_matches_the_host_program(name, prog):
    """Products, sums and divisions are rounded exactly as the eager FP32 ops of the host program; sqrt / sin / cos /
    exp / pow come from another math library and may differ in the last place (PyTorch's vectorised CPU sqrt is
    itself not always correctly rounded: sqrt(fl32(10.129405)) comes out one ulp low), so g^2 may differ by a
    few ulps."""
    assert prog.spec is not None
    rng = np.random.default_rng(1)
    times = np.concatenate([rng.uniform(1e-5, 1.0, 2000), [1e-5, 1e-3, 0.5, 1.0, np.nextafter(np.float32(1), 0)]]).astype(np.float32)
    want = prog(times)
    got = _twin_rows(prog.spec, times)
    want[:, L.MAX_TFEAT + 3] = 1.0
    tf = slice(0, L.MAX_TFEAT)
    if prog.spec.time_features == L.PROG_RAW_T:
        assert np.array_equal(got[:, tf], want[:, tf])
    else:
        # |features| <= 1: compare absolutely (a last-place difference of the argument reduction near a zero is
        # many ulps of a tiny value)
        assert np.abs(got[:, tf] - want[:, tf]).max() <= 2 ** -23
    for col, tol in ((L.MAX_TFEAT + 0, 0), (L.MAX_TFEAT + 1, 0 if prog.spec.sde == L.SDE_NONE else 8),
                     (L.MAX_TFEAT + 2, 16 if prog.spec.use_sigma else None), (L.MAX_TFEAT + 3, 0)):
        if tol is None:        # sigma is read by the kernels only when use_sigma is set
            continue
        assert _ulps(got[:, col], want[:, col]).max() <= tol, (name, col, _ulps(got[:, col], want[:, col]).max())


def test_unknown_or_cancellation_prone_sde_has_no_device_program():
    class MySDE(D.VPSDE):
        pass
    torch.manual_seed(0)
    net = D.MLP(2, 0, 4, [8])
    assert D.ScoreModel(net, MySDE(), no_sigma=True)._program().spec is None
    # 1 - exp(-small) near t = epsilon: one ulp of exp() is ~3e-4 of sigma(t) / g(t)^2 -> host program only
    assert D.ScoreModel(net, D.VPSDE(), no_sigma=False)._program().spec is None
    assert D.ScoreModel(net, D.SUBVPSDE(), no_sigma=True)._program().spec is None
    # the twin still restates them (ffb_control.cuh), to the accuracy the cancellation allows
    sm = D.ScoreModel(net, D.SUBVPSDE(), no_sigma=False)
    prog = sm._program()
    spec = D._device_program(D.VESDE(), net.W.detach(), net.pi.detach(), True, False)
    spec.sde, spec.T, spec.beta_min, spec.beta_diff = L.SDE_SUBVP, 1.0, np.float32(0.1), np.float32(19.9)
    spec.half_beta_diff, spec.m2_beta_min = np.float32(9.95), np.float32(-0.2)
    times = np.linspace(1e-3, 1.0, 500).astype(np.float32)
    want, got = prog(times), _twin_rows(spec, times)
    for col in (L.MAX_TFEAT + 0, L.MAX_TFEAT + 1, L.MAX_TFEAT + 2):
        assert np.allclose(got[:, col], want[:, col], rtol=2e-3, atol=0)


def _run_both(fn, exact=True):
    with S.controller("host"), patched_engine():
        a, sa = fn()
    with S.controller("device"), patched_engine():
        b, sb = fn()
    assert (sa.controller, sb.controller) == ("host", "device")
    assert (sa.accepted, sa.rejected, sa.nfe) == (sb.accepted, sb.rejected, sb.nfe)
    assert sa.accept_history == sb.accept_history
    if exact:
        assert sa.dt_history == sb.dt_history, "the two controllers must take bit-identical step sizes"
        assert sa.ratio_history == sb.ratio_history
    else:     # score fields: g(t) goes through a square root that the two math libraries round differently; the error
        #       estimate is a difference of nearly equal terms, so last-place changes of g^2 show at ~1e-5 in dt
        assert np.allclose(sa.dt_history, sb.dt_history, rtol=5e-3, atol=0)
        assert np.allclose(sa.ratio_history, sb.ratio_history, rtol=2e-2, atol=0)
    return a, b


def test_device_controller_takes_the_host_loops_steps_pfode():
    meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
    sm = D.ScoreModel(D.MLP(**meta["ctor"]), D.VPSDE(), no_sigma=True).eval()
    sm.load_state_dict(sd)

    def run():
        x, _ = sm.sample_ode_from_base(ins["base"], ins["cond"], atol=1e-5, rtol=1e-5,
                                       options={"step_t": torch.tensor([1e-3])})
        return x, sm.last_stats
    a, b = _run_both(run, exact=False)
    assert rel_row_err(a, b) < 1e-5 and rel_row_err(outs["x_dopri5"], b) < 1e-4


@pytest.mark.parametrize("kind", ["ve", "subvp", "vp"])
def test_device_controller_sigma_fields(kind):
    meta, sd, ins, outs = load_golden(f"{kind}_sigma_pfode")
    sde = {"vp": D.VPSDE, "ve": D.VESDE, "subvp": D.SUBVPSDE}[meta["sde"]]()
    sm = D.ScoreModel(D.MLP(**meta["ctor"]), sde, no_sigma=meta["no_sigma"]).eval()
    sm.load_state_dict(sd)
    opts = None if meta["call"]["step_t"] is None else {"step_t": torch.tensor([meta["call"]["step_t"]])}

    def run():
        x, _ = sm.sample_ode_from_base(ins["base"], atol=1e-5, rtol=1e-5, options=opts)
        return x, sm.last_stats
    if kind == "ve":
        a, b = _run_both(run, exact=False)
        assert rel_row_err(a, b) < 1e-5
    else:           # sigma(t) = 1 - exp(-small): these fields keep the host loop (no device program)
        with S.controller("device"), patched_engine():
            b, st = run()
        assert st.controller == "host"
    assert rel_row_err(outs["x_dopri5"], b) < 1e-4


def test_device_controller_logprob_and_conditional_state():
    meta, sd, ins, outs = load_golden("cfg3_flow_logprob")
    m = F.ODEFlow(**meta["ctor"], target_shift=sd["target_shift"], target_scale=sd["target_scale"]).eval()
    m.load_state_dict(sd)
    a, b = _run_both(lambda: (m.log_prob(ins["x"]), m.last_stats))
    assert float((a - b).abs().max()) < 1e-5 and float((b - outs["log_prob"]).abs().max()) < 1e-3
    # the conditional rides in the ODE state with zero derivative: it enters the mixed norm (flow.py:857-861)
    meta, sd, ins, outs = load_golden("cflow_sample_logprob")
    mc = F.ConditionalODEFlow(**meta["ctor"]).eval()
    mc.load_state_dict(sd)
    a, b = _run_both(lambda: (mc.log_prob(outs["x"], ins["cond"], atol=1e-6, rtol=1e-6), mc.last_stats))
    assert float((a - b).abs().max()) < 1e-5


def test_device_controller_symplectic_logprob():
    meta, sd, ins, outs = load_golden("cfg5_symplectic")
    net = Sy.SymplecticMLP(**meta["ctor"])
    m = Sy.SymplecticFlowModel(net, sd["shift"], sd["scale"], sd["conditional_shift"], sd["conditional_scale"]).eval()
    m.load_state_dict(sd)
    a, b = _run_both(lambda: (m.log_prob(ins["x"], conditional=None, p0=ins["p0"]), m.last_stats), exact=False)   # sin / cos features
    assert float((a - b).abs().max()) < 2e-4        # |log p| ~ 45: a few FP32 ulps


def test_controllers_fail_alike():
    """max_num_steps, a step below FP64 resolution and min_step / max_step behave as in the host loop."""
    meta, sd, ins, outs = load_golden("cfg1_flow_sample")
    m = F.ODEFlow(**meta["ctor"]).eval()
    m.load_state_dict(sd)
    x = ins["xT"][:64]
    msgs = {}
    for mode in ("host", "device"):
        with S.controller(mode), patched_engine():
            with pytest.raises(S.SolverError) as e:
                m._integrate(x, None, None, 1.0, 0.0, 1e-6, 1e-6, "dopri5", {"max_num_steps": 2}, L.DIV_NONE)
            msgs[mode] = [str(e.value)]
            with pytest.raises(S.SolverError) as e:
                m._integrate(x, None, None, 1.0, 0.0, 1e-6, 1e-6, "dopri5", {"first_step": 1e-20}, L.DIV_NONE)
            msgs[mode].append(str(e.value))
            y, _ = m._integrate(x, None, None, 1.0, 0.0, 1e-6, 1e-6, "dopri5", {"min_step": 0.05, "max_step": 0.2}, L.DIV_NONE)
            msgs[mode].append((m.last_stats.accepted, m.last_stats.rejected, tuple(m.last_stats.dt_history)))
            msgs[mode].append(y)
    assert msgs["host"][:3] == msgs["device"][:3]
    assert "max_num_steps exceeded (2>=2)" in msgs["host"][0] and "underflow in dt" in msgs["host"][1]
    assert torch.equal(msgs["host"][3], msgs["device"][3])


def test_control_twin_rejects_bad_arguments():
    lib = L.load()
    p, c = L.CtlParams(), L.Ctl()
    assert lib.ffb_dopri5_control_host(None, None, C.byref(c), 0) == -1
    assert lib.ffb_dopri5_control(None, None, None, 0, None, 0, None) == -1
    p.n_grid = L.CTL_MAX_GRID + 1
    assert lib.ffb_dopri5_control_host(C.byref(p), None, C.byref(c), 0) == -1
    assert b"step_t" in lib.ffb_last_error()
    p.n_grid = 0
    assert lib.ffb_dopri5_control_host(C.byref(p), None, C.byref(c), 1) == -1       # sums are required after an attempt
    assert lib.ffb_dopri5_ctl_supported(None) == 0
