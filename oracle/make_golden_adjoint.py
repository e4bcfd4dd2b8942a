"""TEST INFRASTRUCTURE ONLY -- golden vectors for gradients through the sampler, from the
UNMODIFIED reference running on the restated ``torchdiffeq.odeint_adjoint`` (oracle/torchdiffeq/_solver.py; parity of that
third-party piece is unpinned like the forward driver's, see its header -- it is pinned by tests/test_adjoint.py against
back-propagation through a fixed-step integrator).

    python oracle/make_golden_adjoint.py    # writes tests/golden/adjoint_*.npz

Cases: ``ScoreModel.sample_ode_from_base`` in training mode (`diffusion.py:620-629`) for a conditional VP model (no_sigma) and a
VE model with sigma division, and ``ODEFlow.sample(gradients=True)`` (`flow.py:286-295`) and ``ConditionalODEFlow.sample(gradients=True)`` (`:775-785`).  Stored: inputs, samples, the
cotangent dL/dx used (L = sum(w * x) for a fixed random w), dL/d(base samples) and dL/d(every weight and bias), and the
step counts of the backward solve.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.loader import load_reference  # noqa: E402
from oracle import port                   # noqa: E402
from oracle.make_golden import gen, save, stats_dict, WSEED, OUT  # noqa: E402


def main():
    D, F, S = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    cases = [("adjoint_vp_pfode", "vp", True, 6, 2, [48, 48], 96, {"step_t": torch.tensor([1e-3])}, 1e-5),
             ("adjoint_ve_sigma_pfode", "ve", False, 4, 0, [40], 70, None, 1e-4)]
    for name, kind, no_sigma, Dn, Cn, units, B, options, tol in cases:
        print(name)
        torch.manual_seed(WSEED)
        sde = {"vp": D.VPSDE, "ve": D.VESDE}[kind]()
        sm = D.ScoreModel(D.MLP(Dn, Cn, 8, units), sde, no_sigma=no_sigma).train()
        base = torch.randn(B, Dn, generator=gen(41)).requires_grad_(True)
        cond = torch.randn(B, Cn, generator=gen(42)) if Cn else None
        w = torch.randn(B, Dn, generator=gen(43))
        x, _ = sm.sample_ode_from_base(base, cond, atol=tol, rtol=tol, options=options)
        fwd = stats_dict()
        (x * w).sum().backward()
        bwd = stats_dict()
        print(f"  forward {fwd['accepted']}/{fwd['rejected']}  backward {bwd['accepted']}/{bwd['rejected']}")
        ins = {"base": base.detach(), "w": w}
        if cond is not None:
            ins["cond"] = cond
        outs = {"x": x.detach(), "grad/base": base.grad}
        outs.update({"grad/" + k: p.grad for k, p in sm.named_parameters() if p.requires_grad})
        save(name, dict(case="score_adjoint", sde=kind, no_sigma=no_sigma, tol=tol, step_t=None if options is None else 1e-3,
                        ctor=dict(n_dimensions=Dn, n_conditionals=Cn, embedding_dimensions=8, units=units),
                        stats_forward=fwd, stats_backward=bwd), sm.state_dict(), ins, outs)

    print("adjoint_flow_sample")
    torch.manual_seed(WSEED)
    m = F.ODEFlow(3, [32, 32], target_shift=torch.tensor([0.5, -0.5, 0.0]), target_scale=torch.tensor([2.0, 0.5, 1.0])).train()
    xT = torch.randn(60, 3, generator=gen(44)).requires_grad_(True)
    w = torch.randn(60, 3, generator=gen(45))
    x = m.sample(xT, gradients=True)
    fwd = stats_dict()
    (x * w).sum().backward()
    bwd = stats_dict()
    print(f"  forward {fwd['accepted']}/{fwd['rejected']}  backward {bwd['accepted']}/{bwd['rejected']}")
    outs = {"x": x.detach(), "grad/xT": xT.grad}
    outs.update({"grad/" + k: p.grad for k, p in m.named_parameters()})
    save("adjoint_flow_sample", dict(case="flow_adjoint", ctor=dict(target_dimension=3, hidden_units=[32, 32]),
                                     stats_forward=fwd, stats_backward=bwd), m.state_dict(), {"xT": xT.detach(), "w": w}, outs)

    print("adjoint_cflow_sample")
    torch.manual_seed(WSEED)
    m = F.ConditionalODEFlow(3, 2, [32, 24], conditional_shift=torch.tensor([0.5, -0.5]),
                             conditional_scale=torch.tensor([2.0, 0.5])).train()
    xT = torch.randn(50, 3, generator=gen(46)).requires_grad_(True)
    c = torch.randn(50, 2, generator=gen(47)).requires_grad_(True)
    w = torch.randn(50, 3, generator=gen(48))
    x = m.sample(xT, c, gradients=True)
    fwd = stats_dict()
    (x * w).sum().backward()
    bwd = stats_dict()
    print(f"  forward {fwd['accepted']}/{fwd['rejected']}  backward {bwd['accepted']}/{bwd['rejected']}")
    outs = {"x": x.detach(), "grad/xT": xT.grad, "grad/cond": c.grad}
    outs.update({"grad/" + k: p.grad for k, p in m.named_parameters()})
    save("adjoint_cflow_sample", dict(case="cflow_adjoint", ctor=dict(target_dimension=3, conditional_dimension=2, hidden_units=[32, 24]),
                                      stats_forward=fwd, stats_backward=bwd), m.state_dict(),
         {"xT": xT.detach(), "cond": c.detach(), "w": w}, outs)
    print("done")


if __name__ == "__main__":
    main()
