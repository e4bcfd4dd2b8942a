"""Restated torchdiffeq 0.2.x solver semantics (TEST INFRASTRUCTURE -- see package docstring).

Semantics restated (numbering follows SURVEY.md section 8c, T1..T14):

T1  tuple states are flattened to one 1-D tensor; the user function's return value is
    re-flattened by *iterating* over it (so a bare (B, D) tensor is iterated row by row).
T2  descending ``t`` => integrate ``-t`` with ``-func(-t, y)``; ``step_t`` is negated too.
T3  adaptive solvers keep time in float64, state in ``y0.dtype``; the user function always
    sees ``t`` cast to the state dtype.
T4  stages with alpha == 1 are evaluated one ulp (of the state dtype) before ``t1``.
T5  the error norm is RMS for tensor states and, for tuple states, the MAX over the tuple
    components of each component's RMS ("mixed" norm).
T6  Hairer initial step with order 4.
T7  Dormand-Prince-Shampine tableau, error weights, and mid-point weights.
T8  stage inputs are ``y0 + K[:, :i+1] @ (beta_i * dt)`` with dt cast to the state dtype.
T9  accept iff error ratio <= 1, then max_step / min_step overrides.
T10 step-size controller of order 5 (safety .9, ifactor 10, dfactor .2).
T11 no end clipping; the result at ``t_end`` is a 4th-order interpolant.
T12 ``step_t`` grid points shorten an attempt so that it lands on them.
T13 fixed-grid euler / midpoint / rk4 (3/8 rule) with ``step_size`` grids, time in t.dtype.
T14 defaults rtol=1e-7, atol=1e-9, method dopri5.
"""
from __future__ import annotations

import bisect
import dataclasses
import math
import warnings
from typing import Callable, List, Optional, Sequence

import torch


# --------------------------------------------------------------------------------------
# statistics hook (not part of upstream; lets tests compare step counts with the GPU path)
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class SolveStats:
    method: str = ""
    nfe: int = 0
    accepted: int = 0
    rejected: int = 0
    dt_history: List[float] = dataclasses.field(default_factory=list)      # attempted dt
    t_history: List[float] = dataclasses.field(default_factory=list)       # t0 of each attempt
    accept_history: List[bool] = dataclasses.field(default_factory=list)
    ratio_history: List[float] = dataclasses.field(default_factory=list)
    first_step: Optional[float] = None


_LAST: List[SolveStats] = [SolveStats()]


def last_stats() -> SolveStats:
    """Statistics of the most recent ``odeint`` call in this process."""
    return _LAST[0]


# --------------------------------------------------------------------------------------
# norms (T5)
# --------------------------------------------------------------------------------------
def _rms(x: torch.Tensor) -> torch.Tensor:
    return x.abs().pow(2).mean().sqrt()


def _split(flat: torch.Tensor, lead: Sequence[int], shapes) -> tuple:
    out, pos = [], 0
    for shp in shapes:
        n = int(math.prod(shp))
        out.append(flat[..., pos:pos + n].reshape(tuple(lead) + tuple(shp)))
        pos += n
    return tuple(out)


def _mixed_rms(parts) -> torch.Tensor:
    if len(parts) == 0:
        return 0.0
    return max(_rms(p) for p in parts)


# --------------------------------------------------------------------------------------
# function wrappers (T1, T2, T3, T4)
# --------------------------------------------------------------------------------------
_NONE, _PREV, _NEXT = 0, 1, 2


class _Wrapped:
    """user func -> flat func with reverse-time and perturb handling, counting NFE."""

    def __init__(self, func: Callable, shapes, reverse: bool, stats: SolveStats):
        self.func, self.shapes, self.reverse, self.stats = func, shapes, reverse, stats

    def __call__(self, t: torch.Tensor, y: torch.Tensor, perturb: int = _NONE) -> torch.Tensor:
        self.stats.nfe += 1
        t = t.to(y.abs().dtype)                     # T3: cast BEFORE perturbing
        if perturb == _NEXT:
            t = torch.nextafter(t, t + 1)
        elif perturb == _PREV:
            t = torch.nextafter(t, t - 1)           # T4
        if self.reverse:                            # T2
            t = -t
        if self.shapes is None:
            out = self.func(t, y)
        else:                                       # T1
            f = self.func(t, _split(y, (), self.shapes))
            out = torch.cat([f_.reshape(-1) for f_ in f])
        return -out if self.reverse else out


# --------------------------------------------------------------------------------------
# Dormand-Prince 5(4) (T7)
# --------------------------------------------------------------------------------------
_DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_DP_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_DP_C_ERR = [
    35 / 384 - 1951 / 21600,
    0,
    500 / 1113 - 22642 / 50085,
    125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400,
    11 / 84 - 649 / 6300,
    -1.0 / 60.0,
]
_DP_C_MID = [
    6025192743 / 30085553152 / 2,
    0,
    51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2,
    11237099 / 235043384 / 2,
]


# The other adaptive Runge-Kutta solvers of torchdiffeq 0.2.x (bosh3.py, adaptive_heun.py, fehlberg2.py): same driver
# (rk_common.RKAdaptiveStepsizeODESolver), other tableau and order.  c_sol of adaptive_heun / fehlberg2 is NOT the last beta
# row: y1 = y0 + k @ (dt c_sol) is then formed separately, while f1 is still taken from the last stage (rk_common.py).
_TABLEAUS = {
    "dopri5": dict(order=5, alpha=_DP_ALPHA, beta=_DP_BETA, c_sol=_DP_C_SOL, c_err=_DP_C_ERR, c_mid=_DP_C_MID),
    "bosh3": dict(order=3, alpha=[1 / 2, 3 / 4, 1.0], beta=[[1 / 2], [0.0, 3 / 4], [2 / 9, 1 / 3, 4 / 9]],
                  c_sol=[2 / 9, 1 / 3, 4 / 9, 0.0], c_err=[2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8],
                  c_mid=[0.0, 0.5, 0.0, 0.0]),
    "adaptive_heun": dict(order=2, alpha=[1.0], beta=[[1.0]], c_sol=[0.5, 0.5], c_err=[0.5, -0.5], c_mid=[0.5, 0.0]),
    "fehlberg2": dict(order=2, alpha=[1 / 2, 1.0], beta=[[1 / 2], [1 / 256, 255 / 256]], c_sol=[1 / 512, 255 / 256, 1 / 512],
                      c_err=[-1 / 512, 0.0, 1 / 512], c_mid=[0.0, 0.5, 0.0]),
}


class _Dopri5:
    """The adaptive explicit Runge-Kutta driver; `tableau` picks the method (default: Dormand-Prince 5(4))."""

    def __init__(self, func, y0, rtol, atol, norm, stats, tableau="dopri5", min_step=0, max_step=float("inf"),
                 first_step=None, step_t=None, jump_t=None, safety=0.9, ifactor=10.0,
                 dfactor=0.2, max_num_steps=2 ** 31 - 1, dtype=torch.float64, _replay=None, **unused):
        for k in unused:
            warnings.warn(f"dopri5: unexpected option {k!r}")
        # TEST INFRASTRUCTURE (not a torchdiffeq option): `_replay = (dt_history, accept_history)` replays a recorded step
        # sequence instead of running the controller.  bench.py uses it to check a slice of one rank's shard of a
        # multi-GPU solve against this oracle: the step sizes of that solve came from the GLOBAL error norm.
        self.replay = _replay
        tdt = torch.promote_types(dtype, y0.dtype)
        dev = y0.device
        as64 = lambda v: torch.as_tensor(v, dtype=tdt, device=dev)  # noqa: E731
        self.func, self.y0, self.norm, self.stats = func, y0, norm, stats
        self.rtol, self.atol = as64(rtol), as64(atol)
        self.min_step, self.max_step = as64(min_step), as64(max_step)
        self.first_step = None if first_step is None else as64(first_step)
        self.safety, self.ifactor, self.dfactor = as64(safety), as64(ifactor), as64(dfactor)
        self.max_num_steps = int(max_num_steps)
        self.tdtype = tdt
        self.step_t = None if step_t is None else as64(step_t)
        self.jump_t = None if jump_t is None else as64(jump_t)          # discontinuities of f: land on them, then re-evaluate f
        sd = dict(dtype=y0.dtype, device=dev)
        t64 = lambda v: torch.tensor(v, dtype=torch.float64)  # noqa: E731
        tb = _TABLEAUS[tableau]
        self.order = tb["order"]
        self.alpha = t64(tb["alpha"]).to(**sd)
        self.beta = [t64(b).to(**sd) for b in tb["beta"]]
        self.c_sol = t64(tb["c_sol"]).to(**sd)
        self.c_err = t64(tb["c_err"]).to(**sd)
        self.c_mid = t64(tb["c_mid"]).to(**sd)
        self.fsal = bool(self.c_sol[-1] == 0 and len(self.c_sol) - 1 == len(self.beta[-1])
                         and (self.c_sol[:-1] == self.beta[-1]).all())

    # ---- T6 ----
    def _initial_step(self, t0, y0, f0):
        dtype, dev = y0.dtype, y0.device
        t0 = t0.to(dtype)
        scale = self.atol + torch.abs(y0) * self.rtol
        d0 = self.norm(y0 / scale).abs()
        d1 = self.norm(f0 / scale).abs()
        if d0 < 1e-5 or d1 < 1e-5:
            h0 = torch.tensor(1e-6, dtype=dtype, device=dev)
        else:
            h0 = 0.01 * d0 / d1
        h0 = h0.abs()
        y1 = y0 + h0 * f0
        f1 = self.func(t0 + h0, y1)
        d2 = torch.abs(self.norm((f1 - f0) / scale) / h0)
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = torch.max(torch.tensor(1e-6, dtype=dtype, device=dev), h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** (1.0 / float(self.order - 1 + 1))
        h1 = h1.abs()
        return torch.min(100 * h0, h1).to(self.tdtype)

    def _before(self, t):
        f0 = self.func(t[0], self.y0)
        dt = self._initial_step(t[0], self.y0, f0) if self.first_step is None else self.first_step
        self.stats.first_step = float(dt)
        # state of the integrator: (y, f, t_start_of_last_step, t, dt, interpolant)
        self.y, self.f, self.t_prev, self.t, self.dt = self.y0, f0, t[0], t[0], dt
        self.interp = [self.y0] * 5
        if self.step_t is None:
            self.grid = torch.tensor([], dtype=self.tdtype, device=self.y0.device)
        else:
            g = self.step_t[self.step_t >= t[0]]
            self.grid = torch.sort(g).values.to(self.tdtype)
        self.grid_idx = min(bisect.bisect(self.grid.tolist(), t[0]), len(self.grid) - 1)
        if self.jump_t is None:
            self.jumps = torch.tensor([], dtype=self.tdtype, device=self.y0.device)
        else:
            j = self.jump_t[self.jump_t >= t[0]]
            self.jumps = torch.sort(j).values.to(self.tdtype)
        self.jump_idx = min(bisect.bisect(self.jumps.tolist(), t[0]), len(self.jumps) - 1)

    # ---- T8 ----
    def _rk_step(self, y0, f0, t0, dt, t1):
        sd = y0.dtype
        t0, dt, t1 = t0.to(sd), dt.to(sd), t1.to(sd)
        k = torch.empty(*f0.shape, len(self.alpha) + 1, dtype=sd, device=y0.device)
        k[..., 0] = f0
        yi = y0
        for i, (a_i, b_i) in enumerate(zip(self.alpha, self.beta)):
            if a_i == 1.0:
                ti, perturb = t1, _PREV
            else:
                ti, perturb = t0 + a_i * dt, _NONE
            yi = y0 + k[..., : i + 1].matmul(b_i * dt).view_as(f0)
            k[..., i + 1] = self.func(ti, yi, perturb)
        # FSAL tableaus: c_sol equals the last beta row, so the last stage input is y1; otherwise y1 is formed from c_sol
        # -- and f1 is the last stage's derivative either way (rk_common._runge_kutta_step)
        if not self.fsal:
            yi = y0 + k.matmul(dt * self.c_sol).view_as(f0)
        y1, f1 = yi, k[..., -1]
        err = k.matmul(dt * self.c_err)
        return y1, f1, err, k

    # ---- T9, T10, T12 ----
    def _attempt(self):
        y0, f0, t0, dt = self.y, self.f, self.t, self.dt
        t1 = t0 + dt
        assert t0 + dt > t0, "underflow in dt {}".format(dt.item())
        assert torch.isfinite(y0).all(), "non-finite values in state `y`: {}".format(y0)
        on_grid = False
        if len(self.grid):
            nxt = self.grid[self.grid_idx]
            on_grid = bool(t0 < nxt < t0 + dt)
            if on_grid:
                t1 = nxt
                dt = t1 - t0
        on_jump = False
        if len(self.jumps):
            nxt = self.jumps[self.jump_idx]
            on_jump = bool(t0 < nxt < t0 + dt)
            if on_jump:
                on_grid = False
                t1 = nxt
                dt = t1 - t0
        if self.replay is not None:              # recorded (already clipped) step size of this attempt
            i = len(self.stats.dt_history)
            dt = torch.as_tensor(self.replay[0][i], dtype=self.tdtype, device=y0.device)
            t1 = self.grid[self.grid_idx] if on_grid else t0 + dt
        y1, f1, err, k = self._rk_step(y0, f0, t0, dt, t1)
        tol = self.atol + self.rtol * torch.max(y0.abs(), y1.abs())
        ratio = self.norm(err / tol).abs()
        accept = bool(ratio <= 1)
        if dt > self.max_step:
            accept = False
        if dt <= self.min_step:
            accept = True
        if self.replay is not None:
            accept = bool(self.replay[1][len(self.stats.dt_history)])
        st = self.stats
        st.dt_history.append(float(dt)); st.t_history.append(float(t0))
        st.accept_history.append(accept); st.ratio_history.append(float(ratio))
        if accept:
            st.accepted += 1
            self.interp = self._fit(y0, y1, k, dt)
            if on_grid and self.grid_idx != len(self.grid) - 1:
                self.grid_idx += 1
            if on_jump:
                if self.jump_idx != len(self.jumps) - 1:
                    self.jump_idx += 1
                # just past a discontinuity of f: take f from the side we are now on (rk_common.py)
                f1 = self.func(t1, y1, _NEXT)
            self.y, self.f, self.t_prev, self.t = y1, f1, t0, t1
        else:
            st.rejected += 1
            self.t_prev = t0                     # upstream stores (t0, t0) on rejection
        # controller
        if ratio == 0:
            nxt_dt = dt * self.ifactor
        else:
            dfac = torch.ones((), dtype=dt.dtype, device=dt.device) if ratio < 1 else self.dfactor
            r = ratio.type_as(dt)
            expo = torch.tensor(self.order, dtype=dt.dtype, device=dt.device).reciprocal()
            nxt_dt = dt * torch.min(self.ifactor, torch.max(self.safety / r ** expo, dfac))
        self.dt = nxt_dt.clamp(self.min_step, self.max_step)

    # ---- T11 ----
    def _fit(self, y0, y1, k, dt):
        dt = dt.type_as(y0)
        y_mid = y0 + k.matmul(dt * self.c_mid).view_as(y0)
        f0, f1 = k[..., 0], k[..., -1]
        a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
        b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
        c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
        d = dt * f0
        return [y0, d, c, b, a]

    def _evaluate(self, t):
        t0, t1 = self.t_prev, self.t
        assert (t0 <= t) & (t <= t1), "invalid interpolation, fails `t0 <= t <= t1`: {}, {}, {}".format(t0, t, t1)
        x = ((t - t0) / (t1 - t0)).to(self.interp[0].dtype)
        total = self.interp[0] + x * self.interp[1]
        xp = x
        for coeff in self.interp[2:]:
            xp = xp * x
            total = total + xp * coeff
        return total

    def integrate(self, t):
        sol = torch.empty(len(t), *self.y0.shape, dtype=self.y0.dtype, device=self.y0.device)
        sol[0] = self.y0
        t = t.to(self.tdtype)
        self._before(t)
        for i in range(1, len(t)):
            n = 0
            while t[i] > self.t:
                assert n < self.max_num_steps, "max_num_steps exceeded ({}>={})".format(n, self.max_num_steps)
                self._attempt()
                n += 1
            sol[i] = self._evaluate(t[i])
        return sol


# --------------------------------------------------------------------------------------
# fixed-grid drivers (T13)
# --------------------------------------------------------------------------------------
class _FixedGrid:
    def __init__(self, kind, func, y0, stats, step_size=None, grid_constructor=None,
                 interp="linear", perturb=False, **unused):
        for k in ("atol", "rtol", "norm"):
            unused.pop(k, None)
        for k in unused:
            warnings.warn(f"{kind}: unexpected option {k!r}")
        if interp != "linear":
            raise NotImplementedError("only linear interpolation is restated")
        if step_size is not None and grid_constructor is not None:
            raise ValueError("step_size and grid_constructor are mutually exclusive arguments.")
        self.kind, self.func, self.y0, self.stats = kind, func, y0, stats
        self.step_size, self.grid_constructor, self.perturb = step_size, grid_constructor, perturb

    def _grid(self, t):
        if self.step_size is None:
            return t if self.grid_constructor is None else self.grid_constructor(self.func, self.y0, t)
        h = self.step_size
        n = torch.ceil((t[-1] - t[0]) / h + 1).item()
        g = torch.arange(0, n, dtype=t.dtype, device=t.device) * h + t[0]
        g[-1] = t[-1]
        return g

    def _increment(self, t0, dt, t1, y0):
        f = self.func
        first = _NEXT if self.perturb else _NONE
        last = _PREV if self.perturb else _NONE
        k1 = f(t0, y0, first)
        if self.kind == "euler":
            return dt * k1
        if self.kind == "midpoint":
            half = 0.5 * dt
            return dt * f(t0 + half, y0 + k1 * half)
        # rk4: the 3/8 rule
        third, two_thirds = 1 / 3, 2 / 3
        k2 = f(t0 + dt * third, y0 + dt * k1 * third)
        k3 = f(t0 + dt * two_thirds, y0 + dt * (k2 - k1 * third))
        k4 = f(t1, y0 + dt * (k1 - k2 + k3), last)
        return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125

    def integrate(self, t):
        grid = self._grid(t)
        assert grid[0] == t[0] and grid[-1] == t[-1]
        sol = torch.empty(len(t), *self.y0.shape, dtype=self.y0.dtype, device=self.y0.device)
        sol[0] = self.y0
        j, y0 = 1, self.y0
        for t0, t1 in zip(grid[:-1], grid[1:]):
            dt = t1 - t0
            self.stats.accepted += 1
            self.stats.dt_history.append(float(dt)); self.stats.t_history.append(float(t0))
            y1 = y0 + self._increment(t0, dt, t1, y0)
            while j < len(t) and t1 >= t[j]:
                if t[j] == t0:
                    sol[j] = y0
                elif t[j] == t1:
                    sol[j] = y1
                else:
                    sol[j] = y0 + ((t[j] - t0) / (t1 - t0)) * (y1 - y0)
                j += 1
            y0 = y1
        return sol


_FIXED = ("euler", "midpoint", "rk4")


# --------------------------------------------------------------------------------------
# public entry points
# --------------------------------------------------------------------------------------
def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    """Restated ``torchdiffeq.odeint`` (T14 defaults).  Returns a tensor of shape
    ``(len(t), *y0.shape)`` or, for tuple ``y0``, a tuple of such tensors."""
    if event_fn is not None:
        raise NotImplementedError("event handling is outside the restated path")
    stats = SolveStats()
    _LAST[0] = stats
    shapes = None
    if not isinstance(y0, torch.Tensor):                      # T1
        shapes = [y_.shape for y_ in y0]
        y0 = torch.cat([y_.reshape(-1) for y_ in y0])
    if not torch.is_floating_point(y0):
        raise TypeError("`y0` must be a floating point Tensor but is a {}".format(y0.type()))
    options = {} if options is None else dict(options)
    method = "dopri5" if method is None else method
    stats.method = method
    if shapes is not None:
        user_norm = options.get("norm", _mixed_rms)
        options["norm"] = lambda flat: user_norm(_split(flat, (), shapes))   # T5
    else:
        options.setdefault("norm", _rms)
    if not isinstance(t, torch.Tensor):
        raise TypeError("t must be a torch.Tensor")
    if t.ndim != 1:
        raise ValueError("t must be one dimensional")
    reverse = bool(len(t) > 1 and t[0] > t[1])                # T2
    if reverse:
        t = -t
        if options.get("step_t") is not None:
            options["step_t"] = -torch.as_tensor(options["step_t"])
        if options.get("jump_t") is not None:
            options["jump_t"] = -torch.as_tensor(options["jump_t"])
    if not (t[1:] > t[:-1]).all():
        raise ValueError("t must be strictly increasing or decreasing")
    if t.device != y0.device:
        warnings.warn("t is not on the same device as y0. Coercing to y0.device.")
        t = t.to(y0.device)
    f = _Wrapped(func, shapes, reverse, stats)
    if method in _TABLEAUS:
        solver = _Dopri5(f, y0, rtol, atol, stats=stats, tableau=method, **options)
    elif method in _FIXED:
        solver = _FixedGrid(method, f, y0, stats, rtol=rtol, atol=atol, **options)
    else:
        raise ValueError('Invalid method "{}" (restated: dopri5, bosh3, adaptive_heun, fehlberg2, euler, midpoint, rk4)'.format(method))
    sol = solver.integrate(t)
    if shapes is not None:
        sol = _split(sol, (len(t),), shapes)
    return sol


# --------------------------------------------------------------------------------------
# odeint_adjoint: torchdiffeq/_impl/adjoint.py restated.  Forward = odeint without a graph; backward =
# ONE more solve, backwards over each output interval, of the augmented system
#     d/dt (vjp_t, y, adj_y, adj_params) = (0, f, -adj_y^T df/dy, -adj_y^T df/dparams)
# started from (0, y(t_end), dL/dy(t_end), 0), with torchdiffeq's default adjoint norm
#     max(|vjp_t|, state_norm(y), state_norm(adj_y), mixed_rms(adj_params))
# and the forward solve's rtol / atol / method / options (minus `norm`) unless adjoint_* say otherwise.  Gradients with
# respect to t are not restated (the reference never asks for them).  PARITY UNPINNED like the driver above.
# --------------------------------------------------------------------------------------
class _OdeintAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shapes, func, y0, t, rtol, atol, method, options, adj_rtol, adj_atol, adj_method, adj_options, *params):
        ctx.cfg = (shapes, func, adj_rtol, adj_atol, adj_method, adj_options)
        with torch.no_grad():
            sol = odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
        ctx.save_for_backward(t, sol, *params)
        return sol

    @staticmethod
    def backward(ctx, grad_sol):
        shapes, func, rtol, atol, method, options = ctx.cfg
        t, sol, *params = ctx.saved_tensors
        params = tuple(params)
        with torch.no_grad():
            aug = [torch.zeros((), dtype=sol.dtype), sol[-1], grad_sol[-1]] + [torch.zeros_like(p) for p in params]

            def augmented_dynamics(tt, y_aug):
                y, adj_y = y_aug[1], y_aug[2]
                with torch.enable_grad():
                    t_ = tt.detach()
                    y = y.detach().requires_grad_(True)
                    f = func(t_, y)
                    vjp_y, *vjp_p = torch.autograd.grad(f, (y,) + params, -adj_y, allow_unused=True, retain_graph=True)
                vjp_y = torch.zeros_like(y) if vjp_y is None else vjp_y
                vjp_p = [torch.zeros_like(p) if v is None else v for p, v in zip(params, vjp_p)]
                return (torch.zeros_like(t_), f.detach(), vjp_y, *vjp_p)

            for i in range(len(t) - 1, 0, -1):
                out = odeint(augmented_dynamics, tuple(aug), t[i - 1:i + 1].flip(0), rtol=rtol, atol=atol, method=method,
                             options=options)
                aug = [a[1] for a in out]
                aug[1] = sol[i - 1]
                aug[2] = aug[2] + grad_sol[i - 1]
        return (None, None, aug[2], None, None, None, None, None, None, None, None, None, *aug[3:])


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None, adjoint_rtol=None,
                   adjoint_atol=None, adjoint_method=None, adjoint_options=None, adjoint_params=None):
    """Restated ``torchdiffeq.odeint_adjoint`` (no events, no gradients with respect to ``t``)."""
    if event_fn is not None:
        raise NotImplementedError("event handling is outside the restated path")
    if adjoint_params is None and not isinstance(func, torch.nn.Module):
        raise ValueError("func must be an instance of nn.Module to specify the adjoint parameters; alternatively they can be "
                         "specified explicitly via the `adjoint_params` argument.")
    adjoint_rtol = rtol if adjoint_rtol is None else adjoint_rtol
    adjoint_atol = atol if adjoint_atol is None else adjoint_atol
    adjoint_method = method if adjoint_method is None else adjoint_method
    if adjoint_options is None:
        adjoint_options = {k: v for k, v in options.items() if k != "norm"} if options is not None else {}
    else:
        adjoint_options = dict(adjoint_options)
    params = tuple(func.parameters()) if adjoint_params is None else tuple(adjoint_params)
    params = tuple(p for p in params if p.requires_grad)
    shapes = None
    flat_func = func
    if not isinstance(y0, torch.Tensor):                      # T1: the adjoint works on the flattened state
        shapes = [y_.shape for y_ in y0]
        y0 = torch.cat([y_.reshape(-1) for y_ in y0])

        def flat_func(tt, y, _f=func, _s=shapes):
            out = _f(tt, _split(y, (), _s))
            return torch.cat([o.reshape(-1) for o in out])
    state_norm = (options or {}).get("norm")
    if state_norm is None:
        state_norm = _rms if shapes is None else (lambda flat: _mixed_rms(_split(flat, (), shapes)))
    elif shapes is not None:
        user_norm = state_norm
        state_norm = lambda flat: user_norm(_split(flat, (), shapes))      # noqa: E731
    if "norm" not in adjoint_options:
        def default_adjoint_norm(parts):
            tt, y, adj_y, *adj_params = parts
            return max(tt.abs(), state_norm(y), state_norm(adj_y), _mixed_rms(adj_params))
        adjoint_options["norm"] = default_adjoint_norm
    fwd_options = None
    if options is not None:
        fwd_options = dict(options)
        if shapes is not None and "norm" in fwd_options:
            fwd_options["norm"] = state_norm
    if shapes is not None and (fwd_options is None or "norm" not in fwd_options):
        fwd_options = dict(fwd_options or {})
        fwd_options["norm"] = state_norm
    sol = _OdeintAdjoint.apply(shapes, flat_func, y0, t, rtol, atol, method, fwd_options, adjoint_rtol, adjoint_atol,
                               adjoint_method, adjoint_options, *params)
    if shapes is not None:
        sol = _split(sol, (len(t),), shapes)
    return sol
