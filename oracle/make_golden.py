"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` from the UNMODIFIED reference.

Run in the build container (needs ``/root/reference``):

    python oracle/make_golden.py            # writes tests/golden/*.npz, prints port-vs-reference diffs

For each case it (1) builds the reference object under ``torch.manual_seed(1234)``,
(2) calls the reference entry point on seeded inputs (on the restated ``torchdiffeq``),
(3) replays any random draws the reference makes internally so they can be handed to the
port / the CUDA path, (4) checks that ``oracle/port.py`` reproduces the reference output,
and (5) stores state_dict + inputs + outputs + solver statistics.

Each .npz holds: ``meta`` (JSON), ``sd/<key>`` (state_dict), ``in/<name>``, ``out/<name>``.
"""
from __future__ import annotations

import io
import json
import os
import sys
import contextlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.loader import load_reference  # noqa: E402
from oracle import port                   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WSEED = 1234


def gen(seed):
    return torch.Generator().manual_seed(seed)


def quiet(fn, *a, **k):
    with contextlib.redirect_stderr(io.StringIO()):     # tqdm bars
        return fn(*a, **k)


def stats_dict():
    s = port.last_stats()
    return dict(nfe=s.nfe, accepted=s.accepted, rejected=s.rejected, first_step=s.first_step,
                dt_history=s.dt_history, accept_history=[bool(a) for a in s.accept_history])


def save(name, meta, sd, ins, outs):
    arrays = {"meta": np.array(json.dumps(meta))}
    for k, v in sd.items():
        arrays["sd/" + k] = v.detach().cpu().numpy()
    for k, v in ins.items():
        arrays["in/" + k] = v.detach().cpu().numpy()
    for k, v in outs.items():
        arrays["out/" + k] = v.detach().cpu().numpy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {os.path.relpath(path, ROOT)}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def report(name, what, ref, got, tol):
    err = (ref - got).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    print(f"  {name}: {what}: max|ref-port| = {err:.3e} (scale {scale:.2e})")
    assert err <= tol * scale, (name, what, err)


def main():
    D, F, S = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())

    # ---------------------------------------------------------------- cfg1: flow sampling
    print("cfg1 flow sample (torchdiffeq defaults)")
    torch.manual_seed(WSEED)
    m = F.ODEFlow(2, [64, 64, 64]).eval()
    xT = torch.randn(512, 2, generator=gen(1))
    ref = m.sample(xT)
    st = stats_dict()
    got = port.flow_sample(port.flow_from_state_dict(m.state_dict()), xT)
    report("cfg1", "sample", ref, got, 1e-6)
    save("cfg1_flow_sample", dict(case="flow_sample", ctor=dict(target_dimension=2, hidden_units=[64, 64, 64]),
                                  stats=st), m.state_dict(), {"xT": xT}, {"x": ref})

    # ---------------------------------------------------------------- cfg3: flow exact log-prob
    print("cfg3 flow log_prob (exact trace)")
    torch.manual_seed(WSEED)
    m = F.ODEFlow(16, [128] * 4, target_shift=torch.linspace(-1, 1, 16),
                  target_scale=torch.linspace(0.5, 2.0, 16)).eval()
    x = torch.randn(256, 16, generator=gen(4)) * 1.5
    ref = m.log_prob(x).detach()
    st = stats_dict()
    got = port.flow_log_prob(port.flow_from_state_dict(m.state_dict()), x)
    report("cfg3", "log_prob", ref, got, 2e-6)
    save("cfg3_flow_logprob", dict(case="flow_log_prob", ctor=dict(target_dimension=16, hidden_units=[128] * 4),
                                   call=dict(atol=1e-5, rtol=1e-5), stats=st),
         m.state_dict(), {"x": x}, {"log_prob": ref})

    # ---------------------------------------------------------------- conditional flow: sample + log-prob
    print("conditional flow sample / log_prob")
    torch.manual_seed(WSEED)
    m = F.ConditionalODEFlow(6, 3, [64, 96], conditional_shift=torch.tensor([0.5, -0.5, 0.0]),
                             conditional_scale=torch.tensor([2.0, 0.5, 1.0]),
                             target_shift=torch.linspace(-1, 1, 6), target_scale=torch.linspace(0.5, 2.0, 6)).eval()
    xT = torch.randn(300, 6, generator=gen(11)); c = torch.randn(300, 3, generator=gen(12))
    ref_s = m.sample(xT, c)
    st_s = stats_dict()
    Fl = port.flow_from_state_dict(m.state_dict())
    report("cflow", "sample", ref_s, port.flow_sample(Fl, xT, c), 1e-6)
    ref_l = m.log_prob(ref_s.detach(), c, atol=1e-6, rtol=1e-6).detach()
    st_l = stats_dict()
    report("cflow", "log_prob", ref_l, port.flow_log_prob(Fl, ref_s.detach(), c, atol=1e-6, rtol=1e-6), 2e-6)
    save("cflow_sample_logprob",
         dict(case="cflow", ctor=dict(target_dimension=6, conditional_dimension=3, hidden_units=[64, 96]),
              call=dict(atol=1e-6, rtol=1e-6), stats_sample=st_s, stats_logprob=st_l),
         m.state_dict(), {"xT": xT, "cond": c}, {"x": ref_s.detach(), "log_prob": ref_l})

    # ---------------------------------------------------------------- cfg2: VP PF-ODE sampling
    print("cfg2 VP probability-flow ODE sampling (dopri5 + rk4)")
    torch.manual_seed(WSEED)
    net = D.MLP(16, 4, 8, [128] * 4); sde = D.VPSDE()
    sm = D.ScoreModel(net, sde, no_sigma=True).eval()
    base = torch.randn(512, 16, generator=gen(2)); cond = torch.randn(512, 4, generator=gen(3))
    opts = {"step_t": torch.tensor([1e-3])}
    ref = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options=opts)[0].detach()
    st = stats_dict()
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    report("cfg2", "dopri5", ref, port.sample_ode_from_base(M, base, cond, 1e-5, 1e-5, options=opts)[0], 1e-6)
    ref4 = sm.sample_ode_from_base(base, cond, method="rk4", options={"step_size": 1 / 64})[0].detach()
    report("cfg2", "rk4", ref4, port.sample_ode_from_base(M, base, cond, method="rk4", options={"step_size": 1 / 64})[0], 1e-6)
    refe = sm.sample_ode_from_base(base, cond, method="euler", options={"step_size": 1 / 128})[0].detach()
    save("cfg2_vp_pfode", dict(case="score_pfode", sde="vp", no_sigma=True,
                               ctor=dict(n_dimensions=16, n_conditionals=4, embedding_dimensions=8, units=[128] * 4),
                               call=dict(atol=1e-5, rtol=1e-5, step_t=1e-3, rk4_step=1 / 64, euler_step=1 / 128), stats=st),
         sm.state_dict(), {"base": base, "cond": cond}, {"x_dopri5": ref, "x_rk4": ref4, "x_euler": refe})

    # ---------------------------------------------------------------- VE / subVP PF-ODE with sigma division
    for kind, cls, seedo in (("ve", D.VESDE, 20), ("subvp", D.SUBVPSDE, 30), ("vp", D.VPSDE, 40)):
        print(f"{kind} PF-ODE sampling, no_sigma=False, unconditional")
        torch.manual_seed(WSEED + seedo)
        net = D.MLP(5, 0, 6, [48, 80]); sde = cls()
        sm = D.ScoreModel(net, sde, no_sigma=False).eval()
        base = torch.randn(200, 5, generator=gen(seedo))
        opts = None if kind == "ve" else {"step_t": torch.tensor([1e-3])}
        ref = sm.sample_ode_from_base(base, atol=1e-5, rtol=1e-5, options=opts)[0].detach()
        st = stats_dict()
        M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde(kind), False)
        report(kind, "dopri5", ref, port.sample_ode_from_base(M, base, None, 1e-5, 1e-5, options=opts)[0], 1e-6)
        save(f"{kind}_sigma_pfode", dict(case="score_pfode", sde=kind, no_sigma=False,
                                         ctor=dict(n_dimensions=5, n_conditionals=0, embedding_dimensions=6, units=[48, 80]),
                                         call=dict(atol=1e-5, rtol=1e-5, step_t=None if kind == "ve" else 1e-3), stats=st),
             sm.state_dict(), {"base": base}, {"x_dopri5": ref})

    # ---------------------------------------------------------------- score-model log-prob: exact + Hutchinson
    print("score-model log_prob, exact trace and Hutchinson (VP, conditional)")
    torch.manual_seed(WSEED)
    net = D.MLP(8, 2, 8, [64, 64, 64]); sde = D.VPSDE()
    sm = D.ScoreModel(net, sde, no_sigma=True).eval()
    x0 = torch.randn(128, 8, generator=gen(5)); cond = torch.randn(128, 2, generator=gen(6))
    ref = sm.log_prob(x0, cond).detach()
    st = stats_dict()
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    report("score-lp", "exact", ref, port.score_log_prob(M, x0, cond), 1e-5)
    sm.hutch = True
    torch.manual_seed(77)
    refh = sm.log_prob(x0, cond).detach()
    sth = stats_dict()
    torch.manual_seed(77)
    e = torch.sign(torch.randn(x0.shape))                   # replay of `diffusion.py:701`
    report("score-lp", "hutch", refh, port.score_log_prob(M, x0, cond, probes=e), 1e-5)
    save("score_logprob_vp", dict(case="score_logprob", sde="vp", no_sigma=True,
                                  ctor=dict(n_dimensions=8, n_conditionals=2, embedding_dimensions=8, units=[64, 64, 64]),
                                  call=dict(atol=1e-4, rtol=1e-4, min_step=1e-6), stats=st, stats_hutch=sth),
         sm.state_dict(), {"x0": x0, "cond": cond, "probes": e}, {"lp_exact": ref, "lp_hutch": refh})

    print("score-model log_prob, exact, VE with sigma, unconditional")
    torch.manual_seed(WSEED + 1)
    net = D.MLP(4, 0, 4, [32, 32]); sde = D.VESDE()
    sm = D.ScoreModel(net, sde, no_sigma=False).eval()
    x0 = torch.randn(100, 4, generator=gen(8))
    ref = sm.log_prob(x0).detach()
    st = stats_dict()
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("ve"), False)
    report("score-lp-ve", "exact", ref, port.score_log_prob(M, x0), 1e-5)
    save("score_logprob_ve", dict(case="score_logprob", sde="ve", no_sigma=False,
                                  ctor=dict(n_dimensions=4, n_conditionals=0, embedding_dimensions=4, units=[32, 32]),
                                  call=dict(atol=1e-4, rtol=1e-4, min_step=1e-6), stats=st),
         sm.state_dict(), {"x0": x0}, {"lp_exact": ref})

    # ---------------------------------------------------------------- cfg4: Euler-Maruyama
    print("cfg4 reverse SDE Euler-Maruyama (VP, 32-D), 100 and 1000 steps")
    torch.manual_seed(WSEED)
    net = D.MLP(32, 0, 8, [128] * 4); sde = D.VPSDE()
    sm = D.ScoreModel(net, sde, no_sigma=True).eval()
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("vp"), True)
    outs = {}
    for steps, B in ((100, 256), (1000, 64)):
        torch.manual_seed(600 + steps)
        ref = quiet(sm.sample_sde, (B, 32), steps=steps)
        torch.manual_seed(600 + steps)                      # replay the reference's draws
        x0 = sde.prior([32]).sample([B])                    # `diffusion.py:532-536`
        dw = torch.stack([torch.randn_like(x0) for _ in range(steps)])   # `:554`
        report("cfg4", f"{steps} steps", ref, port.sample_sde(M, x0, dw), 1e-6)
        outs[f"x_{steps}"] = ref
    save("cfg4_vp_em", dict(case="score_sde", sde="vp", no_sigma=True,
                            ctor=dict(n_dimensions=32, n_conditionals=0, embedding_dimensions=8, units=[128] * 4),
                            runs=[dict(steps=100, B=256, seed=700), dict(steps=1000, B=64, seed=1600)],
                            note="noise = replay of torch.manual_seed(seed): prior.sample([B]) then randn_like per step"),
         sm.state_dict(), {}, outs)

    print("EM with sigma division + conditional (VE), 50 steps")
    torch.manual_seed(WSEED + 2)
    net = D.MLP(3, 2, 4, [40]); sde = D.VESDE()
    sm = D.ScoreModel(net, sde, no_sigma=False).eval()
    cond = torch.randn(96, 2, generator=gen(9))
    torch.manual_seed(801)
    ref = quiet(sm.sample_sde, (96, 3), conditional=cond, steps=50)
    torch.manual_seed(801)
    x0 = sde.prior([3]).sample([96]); dw = torch.stack([torch.randn_like(x0) for _ in range(50)])
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde("ve"), False)
    report("em-ve", "50 steps", ref, port.sample_sde(M, x0, dw, cond), 1e-6)
    save("ve_em_cond", dict(case="score_sde", sde="ve", no_sigma=False,
                            ctor=dict(n_dimensions=3, n_conditionals=2, embedding_dimensions=4, units=[40])),
         sm.state_dict(), {"x0": x0, "dw": dw, "cond": cond}, {"x": ref})

    # ---------------------------------------------------------------- cfg5: symplectic
    print("cfg5 symplectic Euler sample + dopri5 log_prob")
    torch.manual_seed(WSEED)
    net = S.SymplecticMLP(32, 0, 8, [128] * 4)
    sf = S.SymplecticFlowModel(net, torch.zeros(32), torch.ones(32), torch.zeros(0), torch.ones(0)).eval()
    torch.manual_seed(900)
    ref = quiet(sf.sample, (256, 32), num_steps=100)
    torch.manual_seed(900)
    z0 = torch.randn(256, 64)                               # `symplectic.py:186`
    Sy = port.symplectic_from_state_dict(sf.state_dict())
    report("cfg5", "sample", ref, port.symplectic_sample(Sy, z0, None, 100), 1e-6)
    xq = torch.randn(256, 32, generator=gen(10))
    torch.manual_seed(901)
    refl = sf.log_prob(xq)
    stl = stats_dict()
    torch.manual_seed(901)
    p0 = torch.randn_like(xq)                               # `symplectic.py:228`
    report("cfg5", "log_prob", refl, port.symplectic_log_prob(Sy, xq, p0), 2e-6)
    save("cfg5_symplectic", dict(case="symplectic", ctor=dict(n_data_dims=32, n_conditionals=0, embedding_dimensions=8, units=[128] * 4),
                                 num_steps=100, stats_logprob=stl),
         sf.state_dict(), {"z0": z0, "x": xq, "p0": p0}, {"x_sample": ref, "log_prob": refl})

    print("symplectic, conditional, non-trivial shift/scale")
    torch.manual_seed(WSEED + 3)
    net = S.SymplecticMLP(4, 2, 6, [48, 48])
    sf = S.SymplecticFlowModel(net, torch.linspace(-1, 1, 4), torch.linspace(0.5, 2, 4),
                               torch.tensor([0.3, -0.2]), torch.tensor([1.5, 0.7])).eval()
    cond = torch.randn(120, 2, generator=gen(13))
    torch.manual_seed(902)
    ref = quiet(sf.sample, (120, 4), conditional=cond, num_steps=7)
    torch.manual_seed(902)
    z0 = torch.randn(120, 8)
    Sy = port.symplectic_from_state_dict(sf.state_dict())
    report("symp-c", "sample", ref, port.symplectic_sample(Sy, z0, cond, 7), 1e-6)
    torch.manual_seed(903)
    refl = sf.log_prob(ref, cond)
    stl = stats_dict()
    torch.manual_seed(903)
    p0 = torch.randn_like(ref)
    report("symp-c", "log_prob", refl, port.symplectic_log_prob(Sy, ref, p0, cond), 2e-6)
    save("symplectic_cond", dict(case="symplectic", ctor=dict(n_data_dims=4, n_conditionals=2, embedding_dimensions=6, units=[48, 48]),
                                 num_steps=7, stats_logprob=stl),
         sf.state_dict(), {"z0": z0, "cond": cond, "x": ref, "p0": p0}, {"x_sample": ref, "log_prob": refl})
    print("done")


if __name__ == "__main__":
    main()
