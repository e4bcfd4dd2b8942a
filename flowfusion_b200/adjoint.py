"""Gradients through the sampler: ``torchdiffeq.odeint_adjoint`` for the reference's sampling
entry points -- ``ScoreModel.sample_ode_from_base`` in training mode (`diffusion.py:620-629`) and
``ODEFlow.sample(gradients=True)`` (`flow.py:286-295`) / ``ConditionalODEFlow.sample(..., gradients=True)`` (`:775-785`).

Forward: the ordinary fused solve (no graph), as ``odeint_adjoint`` does.  Backward: ONE more adaptive solve, from the end
time back to the start time, of the augmented system of torchdiffeq's ``OdeintAdjointMethod.backward``

    d/dt (y, adj_y, adj_params) = ( f(t, y),  -adj_y^T df/dy,  -adj_y^T df/dparams )

started from (y(t_end), dL/dy(t_end), 0), with torchdiffeq's default adjoint norm (max over y, adj_y and every parameter
tensor of the component's RMS), the forward solve's rtol / atol / method and its options minus ``norm``.

Where the work runs: every evaluation of the augmented system is ONE fused call of the training kernels in their
vector-Jacobian mode (csrc/ffb_train.cu: forward, backward sweep, weight gradients: ``training.train_step(cot=adj_y)``);
the Runge-Kutta stage algebra and the error norms of the flat augmented state (2 B D + #parameters floats, exactly
torchdiffeq's flattened tuple) are device-tensor operations driven by ``FlatAdaptiveRK`` below -- the same controller
statements as ``solver._dopri5`` (float64 time bookkeeping, float32 stage times, one-ulp perturbation of the stages with
alpha = 1, ``step_t`` landing, dense output at the end time).  Gradients with respect to ``t``, adjoints of the
log-likelihood paths (the divergence needs second derivatives of the network) are not implemented and raise
``NotImplementedError``.  ``ConditionalODEFlow.sample(gradients=True)`` (`flow.py:775-785`) carries the conditional in the
ODE state as the reference does: it takes part in the norms and receives its own gradient."""
from __future__ import annotations

import bisect
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import engine as E
from . import solver as S
from . import training as T

f32, f64 = np.float32, np.float64


def _rms(x: torch.Tensor) -> torch.Tensor:
    return x.abs().pow(2).mean().sqrt()


class FlatAdaptiveRK:
    """torchdiffeq's adaptive Runge-Kutta driver (rk_common.RKAdaptiveStepsizeODESolver) on ONE flat float32 device vector
    whose error norm is the max over ``segments`` of the segment's RMS (torchdiffeq's mixed norm over the tuple the vector
    was flattened from).  ``func(t32, y) -> dy/dt`` takes the user time as a float32 and the flat state."""

    def __init__(self, func: Callable[[np.float32, torch.Tensor], torch.Tensor], segments: Sequence[Tuple[int, int]],
                 rtol: float, atol: float, method: str = "dopri5", options: Optional[dict] = None):
        if method not in S.TABLEAUS:
            raise NotImplementedError(f"adjoint method {method!r} is not implemented ({', '.join(S.TABLEAUS)})")
        self.func, self.segments, self.tab = func, [s for s in segments if s[1] > s[0]], S.TABLEAUS[method]
        self.rtol, self.atol = f32(rtol), f32(atol)
        opts = dict(options or {})
        for k in ("norm", "dtype"):
            opts.pop(k, None)
        self.min_step = f64(opts.pop("min_step", 0))
        self.max_step = f64(opts.pop("max_step", np.inf))
        self.first_step = opts.pop("first_step", None)
        self.step_t = opts.pop("step_t", None)
        if opts.pop("jump_t", None) is not None:
            raise NotImplementedError("jump_t is not implemented for the adjoint solve")
        self.safety, self.ifactor, self.dfactor = f64(opts.pop("safety", 0.9)), f64(opts.pop("ifactor", 10.0)), f64(opts.pop("dfactor", 0.2))
        self.max_num_steps = int(opts.pop("max_num_steps", 2 ** 31 - 1))
        if opts:
            raise TypeError(f"unsupported adjoint options: {sorted(opts)}")
        self.stats = S.SolveStats(method=method, controller="host")

    def norm(self, v: torch.Tensor) -> float:
        return float(torch.stack([_rms(v[a:b]) for a, b in self.segments]).max())

    def integrate(self, y0: torch.Tensor, t0: float, t1: float) -> torch.Tensor:
        tab, st = self.tab, self.stats
        reverse = t0 > t1
        ts, te = (f64(-t0), f64(-t1)) if reverse else (f64(t0), f64(t1))
        if not te > ts:
            raise ValueError("t must be strictly increasing or decreasing")

        def f(solver_t32, y):                      # T2: descending spans integrate -t with -f
            st.nfe += 1
            out = self.func(f32(-solver_t32) if reverse else f32(solver_t32), y)
            return -out if reverse else out

        rtol, atol = float(self.rtol), float(self.atol)
        t0_32 = f32(ts)
        f0 = f(t0_32, y0)
        if self.first_step is None:                # Hairer's heuristic, torchdiffeq `_select_initial_step`
            scale = atol + y0.abs() * rtol
            d0, d1 = abs(self.norm(y0 / scale)), abs(self.norm(f0 / scale))
            h0 = f32(1e-6) if (d0 < 1e-5 or d1 < 1e-5) else f32(0.01) * f32(d0) / f32(d1)
            h0 = abs(h0)
            f1 = f(t0_32 + h0, y0 + float(h0) * f0)
            d2 = abs(f32(self.norm((f1 - f0) / scale)) / h0)
            if d1 <= 1e-15 and d2 <= 1e-15:
                h1 = max(f32(1e-6), h0 * f32(1e-3))
            else:
                h1 = (f32(0.01) / f32(max(d1, d2))) ** f32(1.0 / float(tab.order))
            dt = f64(min(f32(100) * h0, abs(h1)))
        else:
            dt = f64(self.first_step)
        st.first_step = float(dt)
        grid: List[float] = []
        if self.step_t is not None:
            g = np.atleast_1d(np.asarray(torch.as_tensor(self.step_t, dtype=torch.float64).cpu().numpy(), f64))
            g = -g if reverse else g
            grid = sorted(float(v) for v in g if v >= ts)
        grid_idx = min(bisect.bisect(grid, float(ts)), len(grid) - 1) if grid else 0

        nst = len(tab.alpha)
        k = torch.empty(y0.numel(), nst + 1, dtype=torch.float32, device=y0.device)
        dev = y0.device
        as_dev = lambda v: torch.from_numpy(np.ascontiguousarray(v, f32)).to(dev)      # noqa: E731
        t, y, n_steps = ts, y0, 0
        while True:
            if not n_steps < self.max_num_steps:
                raise S.SolverError("max_num_steps exceeded ({}>={})".format(n_steps, self.max_num_steps))
            if not (t + dt > t):
                raise S.SolverError("underflow in dt {}".format(float(dt)))
            t1s, dts, on_grid = t + dt, dt, False
            if grid:
                nxt = f64(grid[grid_idx])
                on_grid = bool(t < nxt < t + dt)
                if on_grid:
                    t1s, dts = nxt, nxt - t
            t0_32, dt_32, t1_32 = f32(t), f32(dts), f32(t1s)
            k[:, 0] = f0
            yi = y
            for i in range(nst):
                ti = np.nextafter(t1_32, t1_32 - f32(1)) if tab.alpha[i] == 1.0 else t0_32 + tab.alpha[i] * dt_32
                yi = y + k[:, : i + 1].matmul(as_dev(tab.beta[i] * dt_32))
                k[:, i + 1] = f(ti, yi)
            if not tab.fsal:
                yi = y + k.matmul(as_dev(dt_32 * tab.c_sol))
            y1, f1 = yi, k[:, nst].clone()
            err = k.matmul(as_dev(dt_32 * tab.c_err))
            n_steps += 1
            if not bool(torch.isfinite(y).all()):
                raise S.SolverError("non-finite values in state `y`")
            ratio = abs(self.norm(err / (atol + rtol * torch.max(y.abs(), y1.abs()))))
            accept = bool(ratio <= 1)
            if dts > self.max_step:
                accept = False
            if dts <= self.min_step:
                accept = True
            st.dt_history.append(float(dts)); st.accept_history.append(accept); st.ratio_history.append(float(ratio))
            final = accept and not (te > t1s)
            if accept:
                st.accepted += 1
                if final:                          # dense output at the end time (torchdiffeq `_interp_fit` / `_interp_evaluate`)
                    y_mid = y + k.matmul(as_dev(dt_32 * tab.c_mid))
                    dtf = float(dt_32)
                    a = 2 * dtf * (f1 - f0) - 8 * (y1 + y) + 16 * y_mid
                    b = dtf * (5 * f0 - 3 * f1) + 18 * y + 14 * y1 - 32 * y_mid
                    c = dtf * (f1 - 4 * f0) - 11 * y - 5 * y1 + 16 * y_mid
                    d = dtf * f0
                    x = float(f32((te - t) / (t1s - t)))
                    total = y + x * d
                    xp = x
                    for coef in (c, b, a):
                        xp = float(f32(xp) * f32(x))
                        total = total + xp * coef
                    return total
                if on_grid and grid_idx != len(grid) - 1:
                    grid_idx += 1
                t, y, f0 = t1s, y1, f1
            else:
                st.rejected += 1
            if ratio == 0:
                nxt_dt = dts * self.ifactor
            else:
                dfac = f64(1.0) if ratio < 1 else self.dfactor
                with np.errstate(all="ignore"):
                    nxt_dt = dts * np.minimum(self.ifactor, np.maximum(self.safety / f64(ratio) ** (f64(1.0) / f64(tab.order)), dfac))
            dt = f64(np.clip(nxt_dt, self.min_step, self.max_step)) if not np.isnan(nxt_dt) else f64(np.nan)


# ---------------------------------------------------------------------------------------------------------------------
# the reference's fields as (network input rows, output transform): f(t, y) = lin(t) * y - k(t) * net(X(t, y))
# ---------------------------------------------------------------------------------------------------------------------
class _ScoreField:
    """PF-ODE drift of a ScoreModel (`diffusion.py:258-279`): f = a(t) x - c(t) score, score = net [/ sigma(t)]."""

    def __init__(self, sm, conditional):
        m = sm.model
        self.linears, self.act = list(m.NN), E.activation_code(m.activation)
        self.prog = sm._program()
        self.emb, self.D = m.embedding_dimensions, m.n_dimensions
        self.use_sigma, self.has_drift = not sm.no_sigma, sm._field().has_drift
        self.cond = conditional
        self.params = [p for lin in self.linears for p in (lin.weight, lin.bias)]

    def rows(self, t32, y):
        row = self.prog(np.array([t32], f32))[0]
        tf = torch.from_numpy(row[: self.emb].copy()).to(y.device)
        cols = [tf.expand(y.shape[0], self.emb), y] + ([self.cond] if self.cond is not None else [])
        lin = float(row[L.MAX_TFEAT + 0]) if self.has_drift else 0.0
        kk = float(f32(row[L.MAX_TFEAT + 1]) / f32(row[L.MAX_TFEAT + 2])) if self.use_sigma else float(row[L.MAX_TFEAT + 1])
        return torch.cat(cols, dim=1), lin, kk, self.emb


class _FlowField:
    """Velocity of an unconditional flow (`flow.py:89-120`): f = net(cat[x, t])."""

    def __init__(self, flow):
        self.linears = [l for l in flow.layers if isinstance(l, torch.nn.Linear)]
        self.act = E.activation_of(flow.layers)
        self.D = flow.target_dimension
        self.params = [p for lin in self.linears for p in (lin.weight, lin.bias)]

    def rows(self, t32, y):
        tcol = torch.full((y.shape[0], 1), float(t32), dtype=torch.float32, device=y.device)
        return torch.cat([y, tcol], dim=1), 0.0, -1.0, 0


class _CondFlowField:
    """Velocity of a conditional flow (`flow.py:553-596`): f = net(cat[x, t, (c - cshift) / cscale]); the conditional is part of
    the ODE state with zero derivative, so it takes part in the norms and collects its own adjoint."""

    def __init__(self, flow):
        self.linears = [l for l in flow.layers if isinstance(l, torch.nn.Linear)]
        self.act = E.activation_of(flow.layers)
        self.D, self.C = flow.target_dimension, flow.conditional_dimension
        self.cshift, self.cscale = flow.conditional_shift, flow.conditional_scale
        self.params = [p for lin in self.linears for p in (lin.weight, lin.bias)]

    def rows(self, t32, y, c):
        tcol = torch.full((y.shape[0], 1), float(t32), dtype=torch.float32, device=y.device)
        return torch.cat([y, tcol, (c - self.cshift) / self.cscale], dim=1), 0.0, -1.0, 0

    def static_vjp(self, gx):
        """adj_x^T df/dc from the kernel's d/dX: the conditional's columns, through the normalisation."""
        return gx[:, self.D + 1: self.D + 1 + self.C] / self.cscale


def adjoint_backward(field, y_end: torch.Tensor, grad_end: torch.Tensor, t_start: float, t_end: float, rtol, atol,
                     method="dopri5", options=None, static: Optional[torch.Tensor] = None):
    """-> (dL/dy(t_start), dL/dstatic or None, [dL/dparam ...], SolveStats): the backward pass of ``odeint_adjoint`` for a solve
    that ran from ``t_start`` to ``t_end`` and produced ``y_end``; ``grad_end`` = dL/dy(t_end).  ``static``: (B, C) state columns
    with zero derivative that the field reads (the conditional of a conditional flow)."""
    E.require_cuda(y_end, "state")
    B, D = y_end.shape
    Cs = 0 if static is None else static.shape[1]
    nx, nc = B * D, B * Cs
    n = nx + nc                                   # flattened forward state: [x | static], as torchdiffeq flattens the tuple
    sizes = T.param_sizes(field.linears)
    P = sum(sizes)
    segs = [(0, nx), (nx, n), (n, n + nx), (n + nx, 2 * n)]
    pos = 2 * n
    for s in sizes:
        segs.append((pos, pos + s))
        pos += s

    def aug(t32, v):
        y, adj = v[:nx].view(B, D), v[n: n + nx].view(B, D)
        if Cs:
            x_in, lin, kk, x_col = field.rows(t32, y, v[nx:n].view(B, Cs))
        else:
            x_in, lin, kk, x_col = field.rows(t32, y)
        out = torch.empty_like(v)
        # one fused call: net(X), (-kk adj)^T d net / d (W, b) straight into the flat derivative, (-kk adj)^T d net / d X
        _, _, gx, o = T.train_step(field.linears, field.act, x_in, None, None, -kk, want_grad_x=True, cot=adj, want_out=True,
                                   grad_flat=out[2 * n:])
        out[2 * n:].neg_()                                              # d adj_params / dt = -adj^T df/dparams
        vjp_y = gx[:, x_col: x_col + D]
        out[:nx].view(B, D).copy_(lin * y - kk * o if lin != 0.0 else -kk * o)          # f
        out[n: n + nx].view(B, D).copy_(-(lin * adj + vjp_y) if lin != 0.0 else -vjp_y)  # d adj_y / dt = -adj^T df/dy
        if Cs:
            out[nx:n].zero_()                                           # the conditional does not move (`flow.py:591-596`)
            out[n + nx: 2 * n].view(B, Cs).copy_(-field.static_vjp(gx))
        return out

    drv = FlatAdaptiveRK(aug, segs, rtol, atol, method or "dopri5", options)
    dev = y_end.device
    parts = [y_end.reshape(-1).float()] + ([static.reshape(-1).float()] if Cs else []) + [grad_end.reshape(-1).float()]
    parts += [torch.zeros(nc + P, dtype=torch.float32, device=dev)]     # dL/dstatic(t_end) = 0, adj_params = 0
    v0 = torch.cat(parts)
    with torch.no_grad(), E.on_device(dev):
        v1 = drv.integrate(v0, float(t_end), float(t_start))
    grads = [g.view_as(p) for g, p in zip(torch.split(v1[2 * n:], sizes), field.params)]
    gstatic = v1[n + nx: 2 * n].view(B, Cs) if Cs else None
    return v1[n: n + nx].view(B, D), gstatic, grads, drv.stats


class _AdjointSolve(torch.autograd.Function):
    """forward: the model's own fused solve; backward: ``adjoint_backward``."""

    @staticmethod
    def forward(ctx, y0, static, owner, field, solve, t_start, t_end, rtol, atol, method, options, *params):
        with torch.no_grad():
            y1 = solve(y0.detach())
        ctx.cfg = (owner, field, t_start, t_end, rtol, atol, method, options)
        ctx.has_static = static is not None
        ctx.save_for_backward(y1, *([static.detach()] if static is not None else []))
        return y1.clone()

    @staticmethod
    def backward(ctx, grad_y1):
        owner, field, t_start, t_end, rtol, atol, method, options = ctx.cfg
        y1 = ctx.saved_tensors[0]
        static = ctx.saved_tensors[1] if ctx.has_static else None
        adj_opts = {k: v for k, v in options.items() if k != "norm"} if options is not None else {}
        gy0, gstatic, gparams, stats = adjoint_backward(field, y1, grad_y1.contiguous(), t_start, t_end, rtol, atol, method,
                                                        adj_opts, static=static)
        owner.last_adjoint_stats = stats
        return (gy0, gstatic, None, None, None, None, None, None, None, None, None) + tuple(gparams)


def solve_with_adjoint(owner, field, solve, y0, t_start, t_end, rtol, atol, method, options, static=None):
    """``odeint_adjoint`` for one of the reference's sampling solves: ``solve(y0) -> y(t_end)`` is the model's forward solve
    (no graph); the result is attached to ``y0``, to ``static`` (a conditional carried in the ODE state) and to the network's
    weights and biases."""
    return _AdjointSolve.apply(y0, static, owner, field, solve, float(t_start), float(t_end), rtol, atol, method, options,
                               *field.params)
