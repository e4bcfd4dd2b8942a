"""TEST INFRASTRUCTURE ONLY -- golden vectors for non-SiLU hidden activations, from the UNMODIFIED reference.

    python oracle/make_golden_act.py        # writes tests/golden/flow_tanh.npz, tests/golden/cflow_gelu.npz

`activation=` is a pass-through constructor argument of every reference model (`flow.py:41, 478`,
`symplectic.py:25`, `diffusion.py:38`).  The GPU tests feed stand-ins with the reference's attribute surface to
``flowfusion_b200.accelerate`` and compare sample / log_prob with these files (tests/test_gpu_engines.py).
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.loader import load_reference  # noqa: E402
from oracle import port                   # noqa: E402
from oracle.make_golden import gen, stats_dict, save, report, WSEED, OUT  # noqa: E402


def main():
    D, F, S = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())

    print("ODEFlow(activation=Tanh): sample / log_prob")
    torch.manual_seed(WSEED)
    m = F.ODEFlow(5, [48, 64], activation=torch.nn.Tanh, target_shift=torch.linspace(-1, 1, 5),
                  target_scale=torch.linspace(0.5, 2.0, 5)).eval()
    xT = torch.randn(200, 5, generator=gen(21))
    ref_s = m.sample(xT).detach()
    st_s = stats_dict()
    Fl = port.flow_from_state_dict(m.state_dict(), act=torch.tanh)
    report("flow_tanh", "sample", ref_s, port.flow_sample(Fl, xT), 1e-6)
    ref_l = m.log_prob(ref_s, atol=1e-6, rtol=1e-6).detach()
    st_l = stats_dict()
    report("flow_tanh", "log_prob", ref_l, port.flow_log_prob(Fl, ref_s, atol=1e-6, rtol=1e-6), 2e-6)
    save("flow_tanh", dict(case="flow_act", activation="Tanh", ctor=dict(target_dimension=5, hidden_units=[48, 64]),
                           call=dict(atol=1e-6, rtol=1e-6), stats_sample=st_s, stats_logprob=st_l),
         m.state_dict(), {"xT": xT}, {"x": ref_s, "log_prob": ref_l})

    print("ConditionalODEFlow(activation=GELU): sample / log_prob")
    torch.manual_seed(WSEED)
    m = F.ConditionalODEFlow(4, 2, [64, 32], activation=torch.nn.GELU, conditional_shift=torch.tensor([0.5, -0.5]),
                             conditional_scale=torch.tensor([2.0, 0.5])).eval()
    xT = torch.randn(150, 4, generator=gen(22)); c = torch.randn(150, 2, generator=gen(23))
    ref_s = m.sample(xT, c).detach()
    st_s = stats_dict()
    Fl = port.flow_from_state_dict(m.state_dict(), act=torch.nn.functional.gelu)
    report("cflow_gelu", "sample", ref_s, port.flow_sample(Fl, xT, c), 1e-6)
    ref_l = m.log_prob(ref_s, c, atol=1e-6, rtol=1e-6).detach()
    st_l = stats_dict()
    report("cflow_gelu", "log_prob", ref_l, port.flow_log_prob(Fl, ref_s, c, atol=1e-6, rtol=1e-6), 2e-6)
    save("cflow_gelu", dict(case="cflow_act", activation="GELU",
                            ctor=dict(target_dimension=4, conditional_dimension=2, hidden_units=[64, 32]),
                            call=dict(atol=1e-6, rtol=1e-6), stats_sample=st_s, stats_logprob=st_l),
         m.state_dict(), {"xT": xT, "cond": c}, {"x": ref_s, "log_prob": ref_l})
    print("done")


if __name__ == "__main__":
    main()
