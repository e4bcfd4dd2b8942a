"""TEST INFRASTRUCTURE ONLY -- golden vectors for the training-side losses, from the
UNMODIFIED reference.

    python oracle/make_golden_train.py      # writes tests/golden/train_*.npz

For every case the reference's own loss function (`diffusion.py:1369-1463`, `flow.py:226-256`, `:716-747`) is called on
seeded data with its internal draws replayed (``torch.manual_seed(s)`` right before the call: the function draws
``randn_like(x)`` then ``rand(batch[, 1])``, and the script draws the same two tensors after re-seeding), the loss is
back-propagated with autograd, and loss + every parameter gradient are stored with the inputs and the draws.  The port's
restatement (oracle/port.py: dsm_loss / lpsm_loss / fm_loss + loss_and_grads) is asserted to reproduce them.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.loader import load_reference  # noqa: E402
from oracle import port                   # noqa: E402
from oracle.make_golden import gen, save, report, WSEED, OUT  # noqa: E402


def grads_of(module, loss):
    names, params = zip(*[(n, p) for n, p in module.named_parameters() if p.requires_grad])
    return dict(zip(names, torch.autograd.grad(loss, params)))


def replay(seed, x, t_shape):
    torch.manual_seed(seed)
    z = torch.randn_like(x)
    u = torch.rand(*t_shape)
    return z, u


def main():
    D, F, S = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())

    cases = [("train_dsm_vp", "vp", True, "dsm", 16, 4, [128] * 4, 700, None),          # the cfg2 network
             ("train_lpsm_ve", "ve", False, "lpsm", 5, 0, [64, 48], 333, None),
             ("train_dsm_subvp_tanh", "subvp", False, "dsm", 3, 2, [40, 24, 32], 130, torch.nn.Tanh)]
    for name, kind, no_sigma, which, Dn, Cn, units, B, act in cases:
        print(name)
        torch.manual_seed(WSEED)
        kw = {} if act is None else {"activation": act()}
        net = D.MLP(Dn, Cn, 8, units, **kw)
        sde = {"vp": D.VPSDE, "ve": D.VESDE, "subvp": D.SUBVPSDE}[kind]()
        sm = D.ScoreModel(net, sde, no_sigma=no_sigma).train()
        x = torch.randn(B, Dn, generator=gen(31)) * 1.5 + 0.3
        cond = torch.randn(B, Cn, generator=gen(32)) if Cn else None
        fn = D.denoising_score_matching if which == "dsm" else D.log_prob_score_matching
        torch.manual_seed(77)
        loss = fn(sm, x, conditional=cond)
        g = grads_of(sm, loss)
        z, u = replay(77, x, (B,))
        t = u * (sde.T - sde.epsilon) + sde.epsilon                       # `diffusion.py:1395-1398`
        M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde(kind), no_sigma,
                                             act=None if act is None else torch.tanh)
        pl, pg = port.loss_and_grads(port.dsm_loss if which == "dsm" else port.lpsm_loss, M["P"]["net"], M, x, z, t, cond)
        report(name, "loss", loss.detach(), pl, 1e-6)
        for i in range(len(units) + 1):
            report(name, f"dW{i}", g[f"model.NN.{i}.weight"], pg[2 * i], 2e-5)
            report(name, f"db{i}", g[f"model.NN.{i}.bias"], pg[2 * i + 1], 2e-5)
        ins = {"x": x, "z": z, "t": t}
        if cond is not None:
            ins["cond"] = cond
        outs = {"loss": loss.detach()}
        outs.update({"grad/" + k: v for k, v in g.items()})
        save(name, dict(case="score_loss", loss=which, sde=kind, no_sigma=no_sigma,
                        activation=None if act is None else act.__name__,
                        ctor=dict(n_dimensions=Dn, n_conditionals=Cn, embedding_dimensions=8, units=units)),
             sm.state_dict(), ins, outs)

    print("train_fm_flow")
    torch.manual_seed(WSEED)
    m = F.ODEFlow(16, [128] * 4, target_shift=torch.linspace(-1, 1, 16), target_scale=torch.linspace(0.5, 2.0, 16)).train()
    x = torch.randn(600, 16, generator=gen(33)) * 2.0
    torch.manual_seed(78)
    loss = m.flow_matching_loss(x)
    g = grads_of(m, loss)
    xT, t = replay(78, x, (600, 1))
    Fl = port.flow_from_state_dict(m.state_dict())
    pl, pg = port.loss_and_grads(port.fm_loss, Fl["net"], Fl, x, xT, t)
    report("train_fm_flow", "loss", loss.detach(), pl, 1e-6)
    for i in range(5):
        report("train_fm_flow", f"dW{i}", g[f"layers.{2 * i}.weight"], pg[2 * i], 2e-5)
    outs = {"loss": loss.detach()}
    outs.update({"grad/" + k: v for k, v in g.items()})
    save("train_fm_flow", dict(case="flow_loss", ctor=dict(target_dimension=16, hidden_units=[128] * 4)),
         m.state_dict(), {"x": x, "xT": xT, "t": t}, outs)

    print("train_fm_cflow_gelu")
    torch.manual_seed(WSEED)
    m = F.ConditionalODEFlow(4, 2, [64, 32], activation=torch.nn.GELU, conditional_shift=torch.tensor([0.5, -0.5]),
                             conditional_scale=torch.tensor([2.0, 0.5])).train()
    x = torch.randn(257, 4, generator=gen(34)); c = torch.randn(257, 2, generator=gen(35))
    torch.manual_seed(79)
    loss = m.flow_matching_loss(x, c)
    g = grads_of(m, loss)
    xT, t = replay(79, x, (257, 1))
    Fl = port.flow_from_state_dict(m.state_dict(), act=torch.nn.functional.gelu)
    pl, pg = port.loss_and_grads(port.fm_loss, Fl["net"], Fl, x, xT, t, c)
    report("train_fm_cflow_gelu", "loss", loss.detach(), pl, 1e-6)
    for i in range(3):
        report("train_fm_cflow_gelu", f"dW{i}", g[f"layers.{2 * i}.weight"], pg[2 * i], 2e-5)
    outs = {"loss": loss.detach()}
    outs.update({"grad/" + k: v for k, v in g.items()})
    save("train_fm_cflow_gelu", dict(case="cflow_loss", activation="GELU",
                                     ctor=dict(target_dimension=4, conditional_dimension=2, hidden_units=[64, 32])),
         m.state_dict(), {"x": x, "cond": c, "xT": xT, "t": t}, outs)
    print("done")


if __name__ == "__main__":
    main()
