"""Training-side fused losses: the reference's denoising score matching, likelihood-weighted
score matching (`diffusion.py:1369-1463`) and flow matching (`flow.py:191-256`, `:679-747`) on one fused CUDA call.

Each loss is ``scale * sum((alpha_b * net(X) + beta)**2)`` for per-row ``alpha`` and per-element ``beta`` that do not
depend on the weights, so ``ffb_train_step`` (csrc/ffb_train.cu) computes the loss AND its gradient with respect to every
weight and bias in one call (forward, residual, backward sweep, weight gradients).  The returned loss is an ordinary 0-d
tensor attached to the parameters through ``_FusedAffineMSE``: ``loss.backward()`` / an optimiser work as with the
reference.  The noise-perturbed inputs are elementwise device-tensor prologue, exactly the reference's statements; the
draws (``z``/``xT``, ``t``) can be passed in for comparison runs, like ``sample_sde(noise=...)``.

There is no CPU path: a model on the CPU raises ``FFBError``."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib as L
from . import engine as E


def param_sizes(linears: Sequence[torch.nn.Linear]) -> List[int]:
    """Element counts of [W_0, b_0, W_1, b_1, ...]."""
    return [n for lin in linears for n in (lin.weight.numel(), lin.bias.numel())]


def train_step(linears: Sequence[torch.nn.Linear], activation: int, x_in: torch.Tensor, alpha: Optional[torch.Tensor],
               beta: Optional[torch.Tensor], scale: float, want_grad_x: bool = False, cot: Optional[torch.Tensor] = None,
               want_out: bool = False, grad_flat: Optional[torch.Tensor] = None):
    """One fused call -> (loss (0-d float64 device tensor), [dW_0, db_0, dW_1, db_1, ...], dX or None[, net(X)]).

    ``cot`` switches to the vector-Jacobian mode (loss = scale * sum(cot * net(X)); alpha / beta unused).  ``grad_flat``: a
    float32 device vector of sum(param_sizes) elements to write the gradients into (the returned list are views of it)."""
    lib = L.load()
    E.require_cuda(x_in, "x")
    dev = x_in.device
    if len(linears) > L.MAX_LAYERS:
        raise NotImplementedError(f"at most {L.MAX_LAYERS} Linear layers are supported")
    d = L.NetDesc()
    d.n_layers = len(linears)
    d.in_features = linears[0].in_features
    keep = []
    for i, lin in enumerate(linears):
        if lin.out_features > L.MAX_WIDTH or lin.in_features > L.MAX_WIDTH:
            raise NotImplementedError(f"layer widths above {L.MAX_WIDTH} are not supported")
        if lin.weight.device != dev:
            raise L.FFBError(f"the inputs are on {dev} but the model's weights are on {lin.weight.device}")
        w, b = E._dev_f32(lin.weight, dev), E._dev_f32(lin.bias, dev)
        keep += [w, b]
        d.widths[i], d.weight[i], d.bias[i] = lin.out_features, w.data_ptr(), b.data_ptr()
    d.x_dim, d.activation = linears[0].in_features, int(activation)
    B = x_in.shape[0]
    x_in = E._dev_f32(x_in, dev)
    tgt = E._dev_f32(beta if cot is None else cot, dev)
    if x_in.shape != (B, linears[0].in_features) or tgt.shape != (B, linears[-1].out_features):
        raise ValueError("train_step: x_in must be (B, in_features) and beta / cot (B, out_features)")
    if alpha is not None:
        alpha = E._dev_f32(alpha.reshape(-1), dev)
        if alpha.shape[0] != B:
            raise ValueError("train_step: alpha must have one entry per row")
    a = L.TrainArgs()
    a.batch, a.x_in, a.alpha = B, x_in.data_ptr(), (alpha.data_ptr() if alpha is not None else None)
    if cot is None:
        a.beta = tgt.data_ptr()
    else:
        a.cot = tgt.data_ptr()
    a.scale = float(scale)
    out = torch.empty(B, linears[-1].out_features, device=dev) if want_out else None
    a.out = out.data_ptr() if out is not None else None
    sizes = param_sizes(linears)
    if grad_flat is None:
        grad_flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
    elif grad_flat.numel() != sum(sizes) or grad_flat.dtype != torch.float32 or not grad_flat.is_contiguous() or grad_flat.device != dev:
        raise ValueError("train_step: grad_flat must be a contiguous float32 device vector of sum(param_sizes) elements")
    grads: List[torch.Tensor] = [g.view_as(k) for g, k in zip(torch.split(grad_flat, sizes), keep)]
    for i in range(len(linears)):
        a.grad_w[i], a.grad_b[i] = grads[2 * i].data_ptr(), grads[2 * i + 1].data_ptr()
    gx = torch.empty_like(x_in) if want_grad_x else None
    a.grad_x = gx.data_ptr() if gx is not None else None
    loss = torch.empty((), dtype=torch.float64, device=dev)
    a.loss = loss.data_ptr()
    nbytes = int(lib.ffb_train_work_bytes(C.byref(d), B, int(want_grad_x)))
    if nbytes == 0:
        raise L.FFBError("ffb_train_work_bytes: " + lib.ffb_last_error().decode())
    work = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
    a.work = work.data_ptr()
    with E.on_device(dev):
        L.check(lib.ffb_train_step(C.byref(d), C.byref(a), E._stream(dev)), "ffb_train_step")
    # `keep`, the inputs and `work` are allocated and consumed on the current stream: the caching allocator reuses their
    # memory in stream order, so nothing has to be recorded
    if want_out:
        return loss, grads, gx, out
    return loss, grads, gx


def mlp_forward(linears: Sequence[torch.nn.Linear], activation: int, x_in: torch.Tensor) -> torch.Tensor:
    """net(x_in) for per-row inputs (B, in_features) on the forward-only mode of ``ffb_train_step`` (two launches): what the
    reference's ``MLP.forward`` / ``score`` do when every sample has its own time (`diffusion.py:82-121`)."""
    lib = L.load()
    E.require_cuda(x_in, "x")
    dev = x_in.device
    if len(linears) > L.MAX_LAYERS:
        raise NotImplementedError(f"at most {L.MAX_LAYERS} Linear layers are supported")
    d = L.NetDesc()
    d.n_layers, d.in_features = len(linears), linears[0].in_features
    keep = []
    for i, lin in enumerate(linears):
        if lin.out_features > L.MAX_WIDTH or lin.in_features > L.MAX_WIDTH:
            raise NotImplementedError(f"layer widths above {L.MAX_WIDTH} are not supported")
        w, b = E._dev_f32(lin.weight, dev), E._dev_f32(lin.bias, dev)
        keep += [w, b]
        d.widths[i], d.weight[i], d.bias[i] = lin.out_features, w.data_ptr(), b.data_ptr()
    d.x_dim, d.activation = linears[0].in_features, int(activation)
    x_in = E._dev_f32(x_in, dev)
    if x_in.dim() != 2 or x_in.shape[1] != linears[0].in_features:
        raise ValueError("mlp_forward: x_in must be (B, in_features)")
    a = L.TrainArgs()
    out = torch.empty(x_in.shape[0], linears[-1].out_features, device=dev)
    nbytes = int(lib.ffb_train_work_bytes(C.byref(d), x_in.shape[0], 2))
    if nbytes == 0:
        raise L.FFBError("ffb_train_work_bytes: " + lib.ffb_last_error().decode())
    work = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
    a.batch, a.x_in, a.out, a.work, a.scale = x_in.shape[0], x_in.data_ptr(), out.data_ptr(), work.data_ptr(), 1.0
    with E.on_device(dev):
        L.check(lib.ffb_train_step(C.byref(d), C.byref(a), E._stream(dev)), "ffb_train_step")
    return out


class _FusedAffineMSE(torch.autograd.Function):
    """loss = scale * sum((alpha * net(x_in) + beta)**2) with the parameter gradients computed by the same call."""

    @staticmethod
    def forward(ctx, x_in, alpha, beta, scale, activation, linears, *params):
        want_gx = bool(x_in.requires_grad)
        sizes = param_sizes(linears)
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=x_in.device)
        loss, grads, gx = train_step(linears, activation, x_in, alpha, beta, scale, want_grad_x=want_gx, grad_flat=flat)
        ctx.save_for_backward(flat, *([gx] if gx is not None else []))
        ctx.has_gx = gx is not None
        ctx.shapes = [g.shape for g in grads]
        ctx.sizes = sizes
        return loss.to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        scaled = saved[0] * g                      # ONE launch for every parameter: the gradients live in one flat vector
        grads = tuple(v.view(shp) for v, shp in zip(torch.split(scaled, ctx.sizes), ctx.shapes))
        gx = saved[1] * g if ctx.has_gx else None
        return (gx, None, None, None, None, None) + grads


def fused_affine_mse(linears: Sequence[torch.nn.Linear], activation: int, x_in, alpha, beta, scale: float) -> torch.Tensor:
    """The loss as a 0-d tensor whose ``backward()`` fills ``.grad`` of every weight and bias of ``linears``."""
    params = []
    for lin in linears:
        params += [lin.weight, lin.bias]
    return _FusedAffineMSE.apply(x_in, alpha, beta, float(scale), int(activation), list(linears), *params)


# ---- score models (`diffusion.py:1369-1463`) -------------------------------------------------------------------------
def _score_inputs(score_model, x, conditional, z, t):
    """The reference's prologue: draws, marginal of the forward SDE, perturbed inputs, the network's input rows."""
    E.require_cuda(x, "x")
    sde, net = score_model.sde, score_model.model
    batch = x.shape[0]
    if z is None:
        z = torch.randn_like(x)                                                     # `:1392`
    if t is None:
        t = torch.rand(batch, device=x.device) * (sde.T - sde.epsilon) + sde.epsilon    # `:1395-1398`
    mean, sigma = sde.marginal_prob(t, x)                                           # `:1401`
    xt = mean + sigma * z
    proj = t[:, None] * net.W[None, :] * 2 * net.pi                                 # `diffusion.py:109-110`
    cols = [torch.sin(proj), torch.cos(proj), xt]
    if conditional is not None:
        cols.append(conditional)                                                    # `:101-102` (x | conditional)
    return z, t, sigma.reshape(-1), torch.cat(cols, dim=1)


def denoising_score_matching(score_model, x, conditional=None, *, z=None, t=None):
    """`diffusion.py:1369-1414`: sum((z + sigma * score(t, mean + sigma z))**2) / batch."""
    z, t, sigma, x_in = _score_inputs(score_model, x, conditional, z, t)
    # score = net / sigma unless no_sigma (`:236`): sigma * score = net, or sigma * net
    alpha = sigma if score_model.no_sigma else None
    return fused_affine_mse(list(score_model.model.NN), E.activation_code(score_model.model.activation), x_in, alpha, z,
                            1.0 / x.shape[0])


def log_prob_score_matching(score_model, x, conditional=None, *, z=None, t=None):
    """`diffusion.py:1417-1463`: sum(((g / sigma) z + g * score)**2) / batch."""
    z, t, sigma, x_in = _score_inputs(score_model, x, conditional, z, t)
    g = score_model.sde.diffusion(t, x).reshape(x.shape[0], -1)[:, 0]              # `:1446`, one value per row
    alpha = g if score_model.no_sigma else g / sigma
    beta = (g / sigma)[:, None] * z
    return fused_affine_mse(list(score_model.model.NN), E.activation_code(score_model.model.activation), x_in, alpha, beta,
                            1.0 / x.shape[0])


# ---- flows (`flow.py:191-256`, `:679-747`) -------------------------------------------------------------------------
def compute_linear_velocity_field(flow, x0, xT, t):
    """`flow.py:191-224`: (x_t, v_hat) of the straight path between the normalised data and the base sample."""
    x0 = (x0 - flow.target_shift) / flow.target_scale
    return (1 - t) * x0 + t * xT, xT - x0


def flow_matching_loss(flow, x, conditional=None, *, xT=None, t=None):
    """`flow.py:226-256` / `:716-747`: mean((v(x_t, t[, c]) - (xT - x0))**2)."""
    E.require_cuda(x, "x")
    if xT is None:
        xT = torch.randn_like(x)
    if t is None:
        t = torch.rand(x.shape[0], 1, device=x.device)
    xt, v_hat = compute_linear_velocity_field(flow, x, xT, t)
    cols = [xt, t.view(-1, 1).expand(x.shape[0], 1)]                                # `flow.py:112-115`
    if conditional is not None:
        cols.append((conditional - flow.conditional_shift) / flow.conditional_scale)    # `flow.py:580-586`
    linears = [m for m in flow.layers if isinstance(m, torch.nn.Linear)]
    return fused_affine_mse(linears, E.activation_of(flow.layers), torch.cat(cols, dim=1), None, -v_hat, 1.0 / v_hat.numel())
