"""TEST INFRASTRUCTURE ONLY -- CPU oracle: a self-contained restatement of the reference's
sampling / density-evaluation path (Cosmo-Pop/flowfusion), in plain FP32 PyTorch-on-CPU.

Why it exists: the reference is Python and lives only in the build container
(``/root/reference``); GPU boxes have neither it nor ``torchdiffeq``.  This port carries the
same arithmetic (same op order where FP32 rounding is observable) on top of the restated
solver in ``oracle/torchdiffeq``, and is pinned against the UNMODIFIED reference by
``oracle/make_golden.py`` / ``oracle/make_golden_trace.py`` (golden vectors in ``tests/golden``; the latter covers the
Hutch++ / XTrace estimators, which this port reproduces bit for bit) and by
``tests/test_oracle.py`` (golden vectors always; the live reference whenever ``/root/reference`` is present).
PARITY NOTE: the field definitions are pinned by the real reference; the ODE driver is the
restated third-party ``torchdiffeq`` (parity unpinned upstream -- see its docstring).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm
may import this module.  The product path never does.

Every function cites the reference lines it follows (paths relative to /root/reference/).
Weights are passed as plain dicts of tensors (see ``net_from_state_dict``) so the oracle does
not depend on any nn.Module class of the reference or of the product.
"""
from __future__ import annotations

import math
import os
import sys
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:          # make the restated `torchdiffeq` importable
    sys.path.insert(0, _HERE)
import torchdiffeq as _tde         # noqa: E402  (oracle/torchdiffeq)

silu = torch.nn.functional.silu


# ------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------
def net_from_state_dict(sd, prefix, stride=1, act=None):
    """Collect ``{prefix}{i*stride}.weight|bias`` into ``{"w": [...], "b": [...]}``.

    ``stride=1`` for ``diffusion.MLP.NN`` (`diffusion.py:67-72`), ``stride=2`` for the
    Sequentials of ``flow.py:63-74`` and ``symplectic.py:71-78`` (Linear, act, Linear, ...).
    """
    w, b, i = [], [], 0
    while f"{prefix}{i * stride}.weight" in sd:
        w.append(sd[f"{prefix}{i * stride}.weight"].detach().float().cpu())
        b.append(sd[f"{prefix}{i * stride}.bias"].detach().float().cpu())
        i += 1
    if not w:
        raise KeyError(f"no layers under {prefix!r}")
    return {"w": w, "b": b, "act": act or silu}      # `act`: the activation callable the model was built with


def _mlp(net, h):
    """Linear -> act -> ... -> Linear  (`diffusion.py:116-119`, `flow.py:118`, `symplectic.py:120`; act defaults to SiLU)."""
    n = len(net["w"])
    act = net.get("act", silu)
    for i in range(n - 1):
        h = act(torch.nn.functional.linear(h, net["w"][i], net["b"][i]))
    return torch.nn.functional.linear(h, net["w"][-1], net["b"][-1])


# ------------------------------------------------------------------------------------
# diffusion.MLP  (`diffusion.py:82-121`)
# ------------------------------------------------------------------------------------
def fourier_features(t, W, pi):
    """`diffusion.py:109-110`: ((t*W)*2)*pi in FP32, then [sin | cos]."""
    proj = t[:, None] * W[None, :] * 2 * pi
    return torch.cat([torch.sin(proj), torch.cos(proj)], dim=1)


def score_net(P, t, x, cond=None):
    """`diffusion.py:100-121`.  P = {"net", "W", "pi"}; input is cat[temb, x, cond]."""
    if cond is not None:
        x = torch.cat([x, cond], dim=1)
    if t.size() == torch.Size([]):
        t = t * torch.ones(x.shape[:-1])
    h = torch.cat([fourier_features(t, P["W"], P["pi"]), x], dim=1)
    return _mlp(P["net"], h)


# ------------------------------------------------------------------------------------
# SDE closed forms (`diffusion.py:818-1366`).  sde = {"kind": "vp"|"subvp"|"ve", ...}
# ------------------------------------------------------------------------------------
def make_sde(kind, **kw):
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)  # noqa: E731
    if kind == "ve":                                       # `diffusion.py:833-850`
        d = dict(sigma_min=1e-2, sigma_max=10.0, T=1.0, epsilon=1e-5); d.update(kw)
        return {"kind": "ve", **{k: f32(v) for k, v in d.items()}}
    d = dict(beta_min=0.1, beta_max=20, T=1.0, epsilon=1e-3); d.update(kw)   # `:1022-1045`
    return {"kind": kind, "beta_min": d["beta_min"], "beta_max": d["beta_max"], "T": d["T"],
            "epsilon": f32(d["epsilon"])}


def sde_beta(s, t):                                        # `:1061`, `:1237`
    return s["beta_min"] + (s["beta_max"] - s["beta_min"]) * (t / s["T"])


def sde_sigma(s, t):
    if s["kind"] == "ve":                                  # `:866`
        return s["sigma_min"] * (s["sigma_max"] / s["sigma_min"]) ** (t / s["T"])
    lc = 0.5 * (s["beta_max"] - s["beta_min"]) * t ** 2 / s["T"] + s["beta_min"] * t
    if s["kind"] == "vp":                                  # `:1152-1156`
        return torch.sqrt(1.0 - torch.exp(-lc))
    return 1.0 - torch.exp(-lc)                            # subVP `:1337-1340`


def sde_diffusion(s, t, x):
    ones = [1] * (x.dim() - 1)
    if s["kind"] == "ve":                                  # `:884-887`
        return sde_sigma(s, t).view(-1, *ones) * torch.sqrt(
            2 * (torch.log(s["sigma_max"]) - torch.log(s["sigma_min"])) / s["T"])
    if s["kind"] == "vp":                                  # `:1111-1112`
        return torch.sqrt(sde_beta(s, t)).view(-1, *ones)
    return torch.sqrt(sde_beta(s, t) * (1.0 - torch.exp(                      # `:1288-1297`
        -2 * s["beta_min"] * t - (s["beta_max"] - s["beta_min"]) * t ** 2 / s["T"]))).view(-1, *ones)


def sde_drift(s, t, x):
    if s["kind"] == "ve":                                  # `:905`
        return torch.zeros_like(x)
    return -0.5 * sde_beta(s, t).view(-1, *[1] * (x.dim() - 1)) * x           # `:1131`, `:1316`


def sde_prior_logprob(s, xT):
    """`diffusion.py:814` with `:1003` (VE: N(0, sigma_max)) / `:1093` (VP, subVP: N(0,1))."""
    scale = s["sigma_max"] if s["kind"] == "ve" else 1.0
    return torch.distributions.Normal(torch.zeros(xT.shape), scale).log_prob(xT)


# ------------------------------------------------------------------------------------
# diffusion.ScoreModel
# ------------------------------------------------------------------------------------
def score(M, t, x, cond=None):
    """`diffusion.py:233-238`.  M = {"P", "sde", "no_sigma"}."""
    out = score_net(M["P"], t, x, cond)
    if M["no_sigma"]:
        return out
    return out / sde_sigma(M["sde"], t).view(-1, *[1] * (x.dim() - 1))


def ode_drift(M, t, x, cond=None):
    """`diffusion.py:276-279`: f - 0.5 g^2 score."""
    f = sde_drift(M["sde"], t, x)
    g = sde_diffusion(M["sde"], t, x)
    return f - 0.5 * g ** 2 * score(M, t, x, cond)


def _batched_vjp(M, t, x, cond):
    """`diffusion.py:359-374`: v (n, B, D) -> J^T v, J = d x_dot / d x, through torch.func.vjp + vmap."""
    x_dot, vjp_fn = torch.func.vjp(lambda x_: ode_drift(M, t, x_, cond), x)
    return x_dot, torch.func.vmap(lambda v: vjp_fn(v.reshape_as(x))[0].reshape(x.shape[0], -1))


def hutchpp_divergence(M, t, x, cond, S, G):
    """`diffusion.py:336-400`.  S (r, B, D), G (m, B, D): the probes drawn once per solve (`:703-711`)."""
    m = G.shape[0]
    x_dot, bvjp = _batched_vjp(M, t, x, cond)
    Y = bvjp(S).permute(1, 2, 0).detach()                    # (B, D, r)              `:367-369`
    Q, _ = torch.linalg.qr(Y, mode="reduced")                #                         `:372`
    AQ = bvjp(Q.permute(2, 0, 1)).permute(1, 2, 0).detach()  #                         `:376-379`
    trace_lr = torch.einsum("bdk,bdk->b", Q, AQ)             #                         `:380`
    Gp = G.permute(1, 2, 0)
    U = Gp - torch.einsum("bdk,bkm->bdm", Q, torch.einsum("bdk,bdm->bkm", Q, Gp))     # `:383-387`
    AU = bvjp(U.permute(2, 0, 1)).permute(1, 2, 0).detach()  #                         `:389-392`
    trace_res = torch.einsum("bdm,bdm->b", U, AU)
    return x_dot, trace_lr + trace_res / float(m)            #                         `:395`


def xtrace_divergence(M, t, x, cond, O):
    """`diffusion.py:402-481`.  O (m, B, D), m <= D."""
    x_dot, bvjp = _batched_vjp(M, t, x, cond)
    Y = bvjp(O).permute(1, 2, 0).detach()                    # (B, D, m)               `:433-435`
    Q, R = torch.linalg.qr(Y, mode="reduced")                #                          `:438`
    k = Q.shape[2]
    AQ = bvjp(Q.permute(2, 0, 1)).permute(1, 2, 0).detach()  #                          `:443-445`
    H = torch.einsum("bdi,bdj->bij", Q, AQ)                  #                          `:447`
    W = torch.einsum("bdk,mbd->bkm", Q, O)                   #                          `:449`
    T = torch.einsum("bdk,mbd->bkm", AQ, O)                  #                          `:451`
    St = torch.linalg.solve_triangular(R, torch.eye(k), upper=True)                   # `:453`
    St = St / torch.linalg.vector_norm(St, dim=-1, keepdim=True)                      # `:455`
    S = St.permute(0, 2, 1)
    trace_H = torch.diagonal(H, 0, 1, 2).sum(dim=-1)         #                          `:459`
    X = W - torch.sum(S * W, dim=1, keepdim=True) * S        #                          `:463`
    SHS = torch.sum(S * torch.einsum("bim,bmk->bik", H, S), dim=1)
    XHX = torch.sum(X * torch.einsum("bim,bmk->bik", H, X), dim=1)
    WS = torch.sum(W * S, dim=1)
    SR = torch.sum(S * R, dim=1)
    TX = torch.sum(T * X, dim=1)
    ests = trace_H[:, None] - SHS + WS * SR - TX + XHX       #                          `:475`
    return x_dot, torch.mean(ests, dim=1)                    #                          `:477`


def score_field(M, t, states, cond=None, prob=False, probes=None):
    """`diffusion.py:313-508` -- exact trace (`:483-503`), Hutchinson (`:327-334`), Hutch++ (`:336-400`) or
    XTrace (`:402-481`).  ``probes``: None (exact), a (B, D) tensor (Hutchinson), ("hutchpp", S, G) or ("xtrace", O)."""
    x = states[0]
    B = x.shape[0]
    if prob and isinstance(probes, tuple):
        with torch.set_grad_enabled(True):
            xg = x.detach().requires_grad_(True)
            if probes[0] == "hutchpp":
                x_dot, div = hutchpp_divergence(M, t, xg, cond, probes[1], probes[2])
            else:
                x_dot, div = xtrace_divergence(M, t, xg, cond, probes[1])
        return x_dot.detach(), div.detach().view(B, 1)
    if not prob:
        # the reference returns the bare tensor (`:508`), which torchdiffeq then re-flattens
        # row by row, and keeps autograd on; a detached 1-tuple gives bit-identical
        # numbers without those costs (so this port is a FASTER-than-reference CPU baseline)
        with torch.no_grad():
            return (ode_drift(M, t, x, cond),)
    with torch.set_grad_enabled(True):
        x = x.detach().requires_grad_(True)
        x_dot = ode_drift(M, t, x, cond)
        if probes is not None:                             # Hutchinson
            div = torch.sum(torch.autograd.grad(x_dot, x, probes, create_graph=False,
                                                retain_graph=True)[0] * probes, dim=1)
        else:
            def one(xs, cs):
                def f(xi):
                    ci = cs.unsqueeze(0) if cs is not None else None
                    return ode_drift(M, t, xi.unsqueeze(0), ci).squeeze(0)
                return torch.trace(torch.func.jacrev(f)(xs))
            dims = (0, 0) if cond is not None else (0, None)
            div = torch.vmap(one, in_dims=dims)(x, cond)
    return x_dot.detach(), div.detach().view(B, 1)


def sample_sde(M, x0, dw_unit, cond=None):
    """Euler-Maruyama, `diffusion.py:529-563`, with the prior draw ``x0`` (`:532-536`) and the
    per-step unit normals ``dw_unit[step]`` (`:554`) supplied by the caller.  Returns x_mean."""
    s = M["sde"]
    steps, batch = dw_unit.shape[0], x0.shape[0]
    x = x0.clone()
    dt = -(s["T"] - s["epsilon"]) / steps                  # `:539`
    t = torch.ones(batch) * s["T"]                         # `:540`
    x_mean = x
    for i in range(steps):
        if t[0] < s["epsilon"]:                            # `:548-551`
            break
        g = sde_diffusion(s, t, x)
        f = sde_drift(s, t, x) - g ** 2 * score(M, t, x, cond)      # `:553`
        dw = dw_unit[i] * (-dt) ** (1.0 / 2.0)             # `:554-556`
        x_mean = x + f * dt
        x = x_mean + g * dw
        t = t + dt                                         # `:559` (in-place there)
        if torch.any(torch.isnan(x)):                      # `:560-562`
            break
    return x_mean


def sample_ode_from_base(M, base, cond=None, atol=1e-4, rtol=1e-4, method="dopri5", options=None):
    """`diffusion.py:604-640`: PF-ODE from t=1.0 to epsilon.  Returns ``(x, [])``."""
    s = M["sde"]
    z = base * s["sigma_max"] if s["kind"] == "ve" else base
    times = torch.tensor([1.0, float(s["epsilon"])])
    times[1] = s["epsilon"]
    with torch.no_grad():
        sol = _tde.odeint(lambda t, st: score_field(M, t, st, cond, prob=False),
                          (z,), times, method=method, atol=atol, rtol=rtol, options=options)
    return sol[0][1, ...], []


def solve_odes_forward(M, x0, cond=None, atol=1e-5, rtol=1e-5, method="dopri5", options=None,
                       probes=None):
    """`diffusion.py:696-754`: (x, delta log p) from epsilon to 1.0; ``probes`` = fixed
    Rademacher vectors for the Hutchinson estimator (`:700-701`), None = exact trace."""
    s = M["sde"]
    dlp = torch.zeros(x0.shape[0], 1)
    times = torch.tensor([0.0, 1.0])
    times[0] = s["epsilon"]
    sol = _tde.odeint(lambda t, st: score_field(M, t, st, cond, prob=True, probes=probes),
                      (x0, dlp), times, method=method, atol=atol, rtol=rtol, options=options)
    return sol[0][1, ...], sol[1][1, ...]


def score_log_prob(M, x0, cond=None, atol=1e-4, rtol=1e-4, method="dopri5",
                   options={"min_step": 1e-6}, probes=None):
    """`diffusion.py:806-815` -> (B, 1)."""
    xT, lp = solve_odes_forward(M, x0, cond, atol, rtol, method, options, probes)
    return lp + torch.sum(sde_prior_logprob(M["sde"], xT), dim=1, keepdim=True)


# population wrappers (`diffusion.py:1556-1640`, `:1754-1848`) -- affine glue only
def population_forward(M, base, shift, scale, cond=None, cshift=None, cscale=None,
                       method="dopri5", options=None):
    if cond is not None:
        cond = (cond - cshift) / cscale
    return sample_ode_from_base(M, base, cond, atol=1e-5, rtol=1e-5, method=method,
                                options=options)[0] * scale + shift


def population_log_prob(M, x, shift, scale, cond=None, cshift=None, cscale=None, atol=1e-5,
                        rtol=1e-5, options=None, probes=None):
    if cond is not None:
        cond = (cond - cshift) / cscale
    xT, lp = solve_odes_forward(M, (x - shift) / scale, cond, atol=atol, rtol=rtol,
                                options=options, probes=probes)
    return lp + torch.sum(sde_prior_logprob(M["sde"], xT), 1, keepdim=True)   # no -sum log scale (Q7)


# ------------------------------------------------------------------------------------
# flow.ODEFlow / ConditionalODEFlow.  Fl = {"net", "shift", "scale", ["cshift", "cscale"]}
# ------------------------------------------------------------------------------------
def flow_velocity(Fl, t, x, cond=None):
    """`flow.py:109-120` / `:577-596`: input cat[x, t, (c - cshift)/cscale]."""
    cols = [x, t.view(-1, 1).expand(x.shape[0], 1)]
    if cond is not None:
        cols.append((cond - Fl["cshift"]) / Fl["cscale"])
    return _mlp(Fl["net"], torch.cat(cols, dim=1))


def flow_velocity_and_divergence(Fl, t, x, cond=None):
    """`flow.py:146-166` / `:627-652`: D reverse sweeps, one per output column."""
    with torch.set_grad_enabled(True):
        x = x.detach().requires_grad_(True)
        v = flow_velocity(Fl, t, x, cond)
        div = torch.zeros(x.shape[0], 1)
        for i in range(x.shape[-1]):
            div = div + torch.autograd.grad(v[:, i].sum(), x, retain_graph=True)[0][:, i].unsqueeze(1)
    return v.detach(), div.detach()


def flow_sample(Fl, xT, cond=None):
    """`flow.py:282-306` / `:775-799`: t 1 -> 0 with torchdiffeq DEFAULTS (rtol 1e-7, atol 1e-9)."""
    times = torch.tensor([1.0, 0.0])
    with torch.no_grad():
        if cond is None:
            sol = _tde.odeint(lambda t, st: (flow_velocity(Fl, t, st[0]),), (xT,), times)
        else:
            sol = _tde.odeint(lambda t, st: (flow_velocity(Fl, t, st[0], st[1]), torch.zeros_like(st[1])),
                              (xT, cond), times)
    return sol[0][-1] * Fl["scale"] + Fl["shift"]


def flow_solve_forward(Fl, x, cond=None, atol=1e-5, rtol=1e-5, method="dopri5", options=None):
    """`flow.py:348-384` / `:845-883`: state (x[, cond], logJ), t 0 -> 1."""
    lj = torch.zeros(x.shape[0], 1)
    times = torch.tensor([0.0, 1.0])
    if cond is None:
        def f(t, st):
            return flow_velocity_and_divergence(Fl, t, st[0])
        sol = _tde.odeint(f, (x, lj), times, method=method, atol=atol, rtol=rtol, options=options)
    else:
        def f(t, st):
            v, d = flow_velocity_and_divergence(Fl, t, st[0], st[1])
            return v, torch.zeros_like(st[1]), d
        sol = _tde.odeint(f, (x, cond, lj), times, method=method, atol=atol, rtol=rtol, options=options)
    return sol[0][1, ...], sol[-1][1, ...]


def flow_log_prob(Fl, x, cond=None, atol=1e-5, rtol=1e-5, method="dopri5", options=None):
    """`flow.py:421-438` / `:923-941` -> (B,)."""
    x = (x - Fl["shift"]) / Fl["scale"]
    xT, lj = flow_solve_forward(Fl, x, cond, atol, rtol, method, options)
    twopi = torch.tensor(2.0 * 3.14159265358979323846)
    lp = torch.sum(-0.5 * xT ** 2 - 0.5 * torch.log(twopi), dim=1)
    return lp + lj.squeeze(1) - torch.sum(torch.log(Fl["scale"]))


# ------------------------------------------------------------------------------------
# symplectic.  Sy = {"net_q", "net_p", "W", "shift", "scale", "cshift", "cscale"}
# ------------------------------------------------------------------------------------
def symplectic_field(Sy, t, state, cond=None):
    """`symplectic.py:99-123`: v_q = mlp_q(cat[p, c, temb]), v_p = -mlp_p(cat[q, c, temb])."""
    q, p = torch.chunk(state, 2, dim=-1)
    if t.dim() == 0:
        t = t.expand(q.shape[0])
    proj = t[:, None] * Sy["W"][None, :] * 2 * math.pi
    temb = torch.cat([torch.sin(proj), torch.cos(proj)], dim=1)
    mid = [cond] if cond is not None else []
    v_q = _mlp(Sy["net_q"], torch.cat([p, *mid, temb], dim=1))
    v_p = -_mlp(Sy["net_p"], torch.cat([q, *mid, temb], dim=1))
    return torch.cat([v_q, v_p], dim=-1)


def symplectic_sample(Sy, z0, cond=None, num_steps=1):
    """`symplectic.py:186-201`: forward Euler on linspace(1, 0); ``z0`` replaces the draw `:186`."""
    x = z0.clone()
    if cond is not None:
        cond = (cond - Sy["cshift"]) / Sy["cscale"]
    ts = torch.linspace(1.0, 0.0, num_steps + 1)
    for i in range(num_steps):
        dt = ts[i + 1] - ts[i]
        x = x + symplectic_field(Sy, ts[i].expand(x.shape[0]), x, cond) * dt
    q, _ = torch.chunk(x, 2, dim=-1)
    return q * Sy["scale"] + Sy["shift"]


def symplectic_log_prob(Sy, x, p0, cond=None, atol=1e-5, rtol=1e-5):
    """`symplectic.py:224-254`; ``p0`` replaces the draw at `:228` -> (B,)."""
    q0 = (x - Sy["shift"]) / Sy["scale"]
    if cond is not None:
        cond = (cond - Sy["cshift"]) / Sy["cscale"]
    init = torch.cat([q0, p0], dim=-1)

    def f(t, state):
        return symplectic_field(Sy, torch.full((state.shape[0],), float(t)), state, cond)

    z1 = _tde.odeint(f, init, torch.tensor([0.0, 1.0]), atol=atol, rtol=rtol)[-1]
    n01 = torch.distributions.Normal(0, 1)
    return n01.log_prob(z1).sum(dim=-1) - n01.log_prob(p0).sum(dim=-1) - torch.sum(torch.log(Sy["scale"]))


def hamiltonian_leapfrog(net, z0, cond=None, n_steps=1, dt=0.01):
    """Oracle of the scalar-Hamiltonian leapfrog EXTENSION (flowfusion_b200.symplectic.HamiltonianMLP; the reference has no
    such model): H = MLP(cat[q, p, c]); p -= dt/2 dH/dq; q += dt dH/dp; p -= dt/2 dH/dq with autograd gradients.
    -> (z, H at the start, H at the end)."""
    D = z0.shape[1] // 2

    def H_and_grad(q, p):
        with torch.set_grad_enabled(True):
            z = torch.cat([q, p], dim=1).detach().requires_grad_(True)
            h = _mlp(net, torch.cat([z, cond], dim=1) if cond is not None else z)[:, 0]
            g = torch.autograd.grad(h.sum(), z)[0]
        return h.detach(), g[:, :D], g[:, D:]

    q, p = z0[:, :D].clone(), z0[:, D:].clone()
    h0 = H_and_grad(q, p)[0]
    for _ in range(n_steps):
        p = p - 0.5 * dt * H_and_grad(q, p)[1]
        q = q + dt * H_and_grad(q, p)[2]
        p = p - 0.5 * dt * H_and_grad(q, p)[1]
    return torch.cat([q, p], dim=1), h0, H_and_grad(q, p)[0]


# ------------------------------------------------------------------------------------
# training losses with the draws passed in; `*_and_grads` add d loss / d (weights, biases)
# ------------------------------------------------------------------------------------
def sde_marginal_scalars(s, t):
    """`diffusion.py:907-924` (VE), `:1133-1156` (VP), `:1318-1342` (subVP): (nu, eta) of p[x(t) | x(0)]."""
    if s["kind"] == "ve":
        return torch.ones_like(t), sde_sigma(s, t)
    lc = 0.5 * (s["beta_max"] - s["beta_min"]) * t ** 2 / s["T"] + s["beta_min"] * t
    return torch.exp(-0.5 * lc), sde_sigma(s, t)


def dsm_loss(M, x, z, t, cond=None):
    """`diffusion.py:1369-1414` with the draws z (`:1392`) and t (`:1395-1398`) given."""
    nu, eta = sde_marginal_scalars(M["sde"], t)
    mean, sigma = nu.view(-1, 1) * x, eta.view(-1, 1)
    return torch.sum((z + sigma * score(M, t, mean + sigma * z, cond)) ** 2) / x.shape[0]


def lpsm_loss(M, x, z, t, cond=None):
    """`diffusion.py:1417-1463` (likelihood weighting)."""
    g = sde_diffusion(M["sde"], t, x)
    nu, eta = sde_marginal_scalars(M["sde"], t)
    mean, sigma = nu.view(-1, 1) * x, eta.view(-1, 1)
    return torch.sum(((g / sigma) * z + g * score(M, t, mean + sigma * z, cond)) ** 2) / x.shape[0]


def fm_loss(Fl, x, xT, t, cond=None):
    """`flow.py:226-256` / `:716-747` with the draws xT and t (B, 1) given."""
    x0 = (x - Fl["shift"]) / Fl["scale"]                  # `flow.py:220`
    xt, v_hat = (1 - t) * x0 + t * xT, xT - x0
    return torch.mean((flow_velocity(Fl, t, xt, cond) - v_hat) ** 2)


def loss_and_grads(loss_fn, net, *args, **kw):
    """(loss, [dW_0, db_0, dW_1, db_1, ...]) by autograd over the weight dict ``net`` = {"w": [...], "b": [...]}."""
    saved = (net["w"], net["b"])
    net["w"] = [w.clone().requires_grad_(True) for w in saved[0]]
    net["b"] = [b.clone().requires_grad_(True) for b in saved[1]]
    try:
        with torch.set_grad_enabled(True):
            loss = loss_fn(*args, **kw)
            params = [p for wb in zip(net["w"], net["b"]) for p in wb]
            grads = torch.autograd.grad(loss, params)
    finally:
        net["w"], net["b"] = saved
    return loss.detach(), [g.detach() for g in grads]


# ------------------------------------------------------------------------------------
# builders from state_dicts that use the reference's key layout
# ------------------------------------------------------------------------------------
def score_model_from_state_dict(sd, sde, no_sigma, prefix="model.", act=None):
    P = {"net": net_from_state_dict(sd, prefix + "NN.", act=act), "W": sd[prefix + "W"].detach().float().cpu(),
         "pi": sd[prefix + "pi"].detach().float().cpu()}
    return {"P": P, "sde": sde, "no_sigma": bool(no_sigma)}


def flow_from_state_dict(sd, act=None):
    Fl = {"net": net_from_state_dict(sd, "velocity.", 2, act=act),
          "shift": sd["target_shift"].float().cpu(), "scale": sd["target_scale"].float().cpu()}
    if "conditional_shift" in sd:
        Fl["cshift"] = sd["conditional_shift"].float().cpu()
        Fl["cscale"] = sd["conditional_scale"].float().cpu()
    return Fl


def symplectic_from_state_dict(sd, act=None):
    return {"net_q": net_from_state_dict(sd, "model.mlp_q_dynamics.", 2, act=act),
            "net_p": net_from_state_dict(sd, "model.mlp_p_dynamics.", 2, act=act),
            "W": sd["model.W"].float().cpu(), "shift": sd["shift"].float().cpu(),
            "scale": sd["scale"].float().cpu(), "cshift": sd["conditional_shift"].float().cpu(),
            "cscale": sd["conditional_scale"].float().cpu()}


def last_stats():
    return _tde.last_stats()
