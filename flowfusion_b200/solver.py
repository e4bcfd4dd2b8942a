"""Host-side solver drivers: the control plane of ``torchdiffeq.odeint`` re-built for a
batch-sharded, kernel-backed state.

The arithmetic on the state (stage evaluations, error partial sums, dense output) happens in the
CUDA kernels behind a *backend* object (``engine.CudaBackend``); this module only does what
torchdiffeq does on the host -- time bookkeeping in float64, the Hairer initial step, the
accept/reject test and the step-size controller -- plus the one collective the path needs:
a SUM all-reduce of a handful of float64 partial sums per attempted step, so every rank takes
bit-identical decisions.

Semantics follow torchdiffeq 0.2.x (call sites: reference `diffusion.py:631-639, 744-752`,
`flow.py:299-303, 371-382, 792-796, 869-880`, `symplectic.py:237`):
  * state dtype float32, time in float64; stage times are formed in float32 (t0 + alpha*dt);
  * stages with alpha == 1 are evaluated one float32 ulp before t1;
  * descending time spans integrate -t with -f;
  * error norm: RMS for tensor states, max over tuple components of per-component RMS;
  * no clipping at t_end: the result is the 4th-order dense output of the last accepted step;
  * ``step_t`` grid points shorten an attempt so that it lands on them.
"""
from __future__ import annotations

import bisect
import contextlib
import dataclasses
import os
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L

f32, f64 = np.float32, np.float64

# Dormand-Prince(-Shampine) tableau, float64 then rounded to the state dtype like torchdiffeq does
_ALPHA = np.array([1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0])
_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_C_ERR = np.array([35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
                   -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0])
_C_MID = np.array([6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2,
                   -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
                   -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2])
ALPHA32 = _ALPHA.astype(f32)
BETA32 = np.zeros((6, 6), f32)
for _i, _row in enumerate(_BETA):
    BETA32[_i, : len(_row)] = np.array(_row, f64).astype(f32)
C_ERR32 = _C_ERR.astype(f32)
C_MID32 = _C_MID.astype(f32)


@dataclasses.dataclass
class Tableau:
    """An adaptive explicit Runge-Kutta method as torchdiffeq 0.2.x defines it (rk_common._ButcherTableau + the mid-point
    weights of its interpolant), coefficients rounded to float32 like the solver state.  ``fsal``: c_sol equals the last
    beta row, i.e. the last stage's input is y1."""
    name: str
    order: int
    alpha: np.ndarray
    beta: list
    c_sol: np.ndarray
    c_err: np.ndarray
    c_mid: np.ndarray

    @property
    def fsal(self) -> bool:
        return bool(self.c_sol[-1] == 0 and len(self.beta[-1]) == len(self.c_sol) - 1
                    and np.array_equal(self.c_sol[:-1], self.beta[-1]))


def _tab(name, order, alpha, beta, c_sol, c_err, c_mid) -> Tableau:
    r = lambda v: np.array(v, f64).astype(f32)        # noqa: E731
    return Tableau(name, order, r(alpha), [r(b) for b in beta], r(c_sol), r(c_err), r(c_mid))


# torchdiffeq/_impl/{dopri5,bosh3,adaptive_heun,fehlberg2}.py
TABLEAUS = {
    "dopri5": _tab("dopri5", 5, _ALPHA, _BETA, list(_BETA[-1]) + [0.0], _C_ERR, _C_MID),
    "bosh3": _tab("bosh3", 3, [1 / 2, 3 / 4, 1.0], [[1 / 2], [0.0, 3 / 4], [2 / 9, 1 / 3, 4 / 9]], [2 / 9, 1 / 3, 4 / 9, 0.0],
                  [2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8], [0.0, 0.5, 0.0, 0.0]),
    "adaptive_heun": _tab("adaptive_heun", 2, [1.0], [[1.0]], [0.5, 0.5], [0.5, -0.5], [0.5, 0.0]),
    "fehlberg2": _tab("fehlberg2", 2, [1 / 2, 1.0], [[1 / 2], [1 / 256, 255 / 256]], [1 / 512, 255 / 256, 1 / 512],
                      [-1 / 512, 0.0, 1 / 512], [0.0, 0.5, 0.0]),
}
ADAPTIVE_METHODS = tuple(TABLEAUS)


@dataclasses.dataclass
class SolveStats:
    method: str = ""
    nfe: int = 0
    accepted: int = 0
    rejected: int = 0
    first_step: Optional[float] = None
    dt_history: List[float] = dataclasses.field(default_factory=list)
    accept_history: List[bool] = dataclasses.field(default_factory=list)
    ratio_history: List[float] = dataclasses.field(default_factory=list)
    launches: int = 0
    controller: str = "host"
    history_truncated: bool = False     # device controller: only the first CTL_HIST attempts are in the *_history lists


# Where torchdiffeq's accept / step-size loop runs: "device" = csrc/ffb_control.cuh between two attempt kernels
# (the host enqueues ahead and follows through pinned memory), "host" = the Python loop below (one device->host
# read per attempt, ~0.2 ms during which the GPU idles).  Both evaluate the same statements.  "auto" (default)
# takes the device controller whenever the backend and the scalar program support it, except for attempts so
# long that the host round trip no longer matters: the attempt kernels that read their step from the controller
# block are 1-2 % slower than the launch-argument ones (measured: cfg2 4.29 -> 4.33 ms, cfg3 318 -> 325 ms).
_CONTROLLER = os.environ.get("FFB_CONTROLLER", "auto")
_CTL_AHEAD = 2        # attempts enqueued beyond the last controller turn whose outcome the host has seen
_CTL_AUTO_MAX_ATTEMPT_MS = 10.0


def set_controller(mode: str):
    global _CONTROLLER
    if mode not in ("auto", "device", "host"):
        raise ValueError("controller must be 'auto', 'device' or 'host'")
    _CONTROLLER = mode


def get_controller() -> str:
    return _CONTROLLER


@contextlib.contextmanager
def controller(mode: str):
    prev = _CONTROLLER
    set_controller(mode)
    try:
        yield
    finally:
        set_controller(prev)


class SolverError(AssertionError):
    """Raised where torchdiffeq raises AssertionError (dt underflow, non-finite state, ...)."""


def _allreduce(t: torch.Tensor, group):
    if group is not None:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=group)
    return t


def _rms32(sumsq: float, n: int) -> np.float32:
    if n == 0:
        return f32(np.nan)          # torch: mean of an empty tensor
    with np.errstate(all="ignore"):
        return f32(np.sqrt(f64(sumsq) / f64(n)))


def _mixed(values: Sequence[np.float32]) -> np.float32:
    # torchdiffeq's mixed norm is Python's max() over the tuple components, in tuple order
    return max(values)


# ----------------------------------------------------------------------------------------------
# adaptive Dormand-Prince 5(4)
# ----------------------------------------------------------------------------------------------
def adaptive(method: str, backend, program, t0, t1, rtol, atol, options=None, group=None) -> SolveStats:
    """``torchdiffeq.odeint(..., method=method)`` for the adaptive Runge-Kutta methods of ``TABLEAUS``: dopri5 runs on the fused
    attempt kernel (with the device-side controller when it applies), bosh3 / adaptive_heun / fehlberg2 evaluation at a
    time (``backend.attempt_rk``) under the same host controller with the method's own order."""
    return dopri5(backend, program, t0, t1, rtol, atol, options, group, method=method)


def dopri5(backend, program: Callable[[np.ndarray], np.ndarray], t0: float, t1: float, rtol: float,
           atol: float, options: Optional[dict] = None, group=None, method: str = "dopri5") -> SolveStats:
    """Integrate the backend's state from ``t0`` to ``t1`` (float32-representable floats).

    ``program(times32)`` maps float32 *user* times, shape (n,), to the (n, EV_FLOATS) rows of
    host-computed evaluation scalars (time features, SDE coefficients).  The result is left in
    the backend (``backend.output()``).  Every launch of the solve is made with the backend's device current
    (the C ABI launches on the current device with raw pointers)."""
    dev = getattr(backend, "dev", None)
    if dev is None or torch.device(dev).type != "cuda":          # the CPU model of the kernels (tests/kernel_model.py)
        return _dopri5(backend, program, t0, t1, rtol, atol, options, group, TABLEAUS[method])
    with torch.cuda.device(dev):
        return _dopri5(backend, program, t0, t1, rtol, atol, options, group, TABLEAUS[method])


def _dopri5(backend, program, t0, t1, rtol, atol, options=None, group=None, tab: Tableau = None) -> SolveStats:
    tab = TABLEAUS["dopri5"] if tab is None else tab
    fused = tab.name == "dopri5"                   # the fused 6-evaluation attempt kernel; other tableaus go stage by stage
    if not fused and getattr(backend, "attempt_rk", None) is None:
        raise NotImplementedError(f"method {tab.name!r} is not available with this divergence estimator (dopri5 is)")
    opts = dict(options or {})
    for k in ("norm", "dtype"):
        opts.pop(k, None)
    min_step = f64(opts.pop("min_step", 0))
    max_step = f64(opts.pop("max_step", np.inf))
    first_step = opts.pop("first_step", None)
    step_t = opts.pop("step_t", None)
    jump_t = opts.pop("jump_t", None)
    safety, ifactor, dfactor = f64(opts.pop("safety", 0.9)), f64(opts.pop("ifactor", 10.0)), f64(opts.pop("dfactor", 0.2))
    max_num_steps = int(opts.pop("max_num_steps", 2 ** 31 - 1))
    if opts:
        raise TypeError(f"unsupported dopri5 options: {sorted(opts)}")

    st = SolveStats(method=tab.name)
    reverse = t0 > t1
    sgn = -1.0 if reverse else 1.0
    ts, te = (f64(-t0), f64(-t1)) if reverse else (f64(t0), f64(t1))
    if not te > ts:
        raise ValueError("t must be strictly increasing or decreasing")
    rtol32, atol32 = f32(rtol), f32(atol)

    def rows(solver_times32: np.ndarray) -> np.ndarray:
        user = (-solver_times32 if reverse else solver_times32).astype(f32)
        ev = program(user)
        ev[:, L.MAX_TFEAT + 3] = sgn
        return ev

    counts = backend.global_counts(group)          # dict: x, lp, cond element counts (global)

    def norm(sx, slp, sc):
        comps = [_rms32(sx, counts["x"])]
        if counts.get("cond"):
            comps.append(_rms32(sc, counts["cond"]))
        if counts.get("lp"):
            comps.append(_rms32(slp, counts["lp"]))
        return _mixed(comps)

    # ---- f0 and the initial step (Hairer; torchdiffeq `_select_initial_step`, order 4) --------
    t0_32 = f32(ts)
    s = _allreduce(backend.eval0(rows(np.array([t0_32], f32))[0], atol32, rtol32), group).cpu().numpy()
    st.nfe += 1
    st.launches += 2
    if first_step is None:
        with np.errstate(all="ignore"):
            d0 = abs(norm(s[L.P_X_Y], 0.0, s[L.P_C_Y]))
            d1 = abs(norm(s[L.P_X_F], s[L.P_LP_F], 0.0))
            if d0 < 1e-5 or d1 < 1e-5:
                h0 = f32(1e-6)
            else:
                h0 = f32(0.01) * d0 / d1
            h0 = abs(h0)
            s = _allreduce(backend.eval1(h0, rows(np.array([t0_32 + h0], f32))[0], atol32, rtol32), group).cpu().numpy()
            st.nfe += 1
            st.launches += 2
            d2 = abs(norm(s[L.P_X_DF], s[L.P_LP_DF], 0.0) / h0)
            if d1 <= 1e-15 and d2 <= 1e-15:
                h1 = max(f32(1e-6), h0 * f32(1e-3))
            else:
                h1 = (f32(0.01) / max(d1, d2)) ** f32(1.0 / float(tab.order))     # _select_initial_step(order - 1): 1 / (order - 1 + 1)
            h1 = abs(h1)
            dt = f64(min(f32(100) * h0, h1))
    else:
        dt = f64(first_step)
    st.first_step = float(dt)

    grid: List[float] = []
    if step_t is not None:
        g = np.atleast_1d(np.asarray(torch.as_tensor(step_t, dtype=torch.float64).cpu().numpy(), f64))
        if reverse:
            g = -g
        grid = sorted(float(v) for v in g if v >= ts)
    grid_idx = min(bisect.bisect(grid, float(ts)), len(grid) - 1) if grid else 0
    # options['jump_t']: discontinuities of f.  A step lands on the next one exactly (like step_t) and, once accepted, f is
    # evaluated again one ulp AFTER it (rk_common.py: `f1 = func(t_next, y_next, perturb=Perturb.NEXT)`): host loop only.
    jumps: List[float] = []
    if jump_t is not None:
        g = np.atleast_1d(np.asarray(torch.as_tensor(jump_t, dtype=torch.float64).cpu().numpy(), f64))
        if reverse:
            g = -g
        jumps = sorted(float(v) for v in g if v >= ts)
    jump_idx = min(bisect.bisect(jumps, float(ts)), len(jumps) - 1) if jumps else 0

    spec = getattr(program, "spec", None)
    if (fused and not jumps and _CONTROLLER != "host" and spec is not None and len(grid) <= L.CTL_MAX_GRID and counts["x"] > 0
            and getattr(backend, "ctl_supported", lambda: False)()
            and (_CONTROLLER == "device" or backend.ctl_attempt_ms_estimate(_rows_per_rank(counts, backend, group))
                 < _CTL_AUTO_MAX_ATTEMPT_MS)):
        # every term of this condition is the same on all ranks of a sharded solve (global counts, the field, the grid,
        # the process-wide controller mode): the ranks must agree, their sequences of collectives differ otherwise
        p = L.CtlParams()
        p.t_end, p.min_step, p.max_step = float(te), float(min_step), float(max_step)
        p.safety, p.ifactor, p.dfactor = float(safety), float(ifactor), float(dfactor)
        p.n_x, p.n_lp, p.n_cond = int(counts["x"]), int(counts.get("lp") or 0), int(counts.get("cond") or 0)
        p.reverse, p.max_num_steps, p.n_grid = int(reverse), min(max_num_steps, 2 ** 31 - 1), len(grid)
        for i, v in enumerate(grid):
            p.grid[i] = v
        for i in range(6):
            p.alpha[i] = float(ALPHA32[i])
            for j in range(6):
                p.beta[i][j] = float(BETA32[i, j])
        for j in range(7):
            p.c_err[j], p.c_mid[j] = float(C_ERR32[j]), float(C_MID32[j])
        p.prog = spec
        return _dopri5_device(backend, p, st, float(ts), float(dt), grid_idx, atol32, rtol32, group)

    t = ts
    n_steps = 0
    done = False
    while te > t:
        if not n_steps < max_num_steps:
            raise SolverError("max_num_steps exceeded ({}>={})".format(n_steps, max_num_steps))
        with np.errstate(all="ignore"):
            if not (t + dt > t):
                raise SolverError("underflow in dt {}".format(float(dt)))
            t1s = t + dt
            dts = dt
            on_grid = False
            if grid:
                nxt = f64(grid[grid_idx])
                on_grid = bool(t < nxt < t + dt)
                if on_grid:
                    t1s = nxt
                    dts = t1s - t
            on_jump = False
            if jumps:
                nxt = f64(jumps[jump_idx])
                on_jump = bool(t < nxt < t + dt)
                if on_jump:
                    on_grid = False
                    t1s = nxt
                    dts = t1s - t
            t0_32, dt_32, t1_32 = f32(t), f32(dts), f32(t1s)
            nst = len(tab.alpha)
            times = np.empty(nst, f32)
            for i in range(nst):
                if tab.alpha[i] == 1.0:
                    times[i] = np.nextafter(t1_32, t1_32 - f32(1))
                else:
                    times[i] = t0_32 + tab.alpha[i] * dt_32
            ev = rows(times)
            final = not (te > t1s)
            x_interp = f32((te - t) / (t1s - t)) if final else f32(0)
        if fused:
            cb = BETA32 * dt_32
            ce = dt_32 * C_ERR32
            cm = dt_32 * C_MID32
            s = _allreduce(backend.attempt(ev, cb, ce, cm, dt_32, atol32, rtol32, final, x_interp), group).cpu().numpy()
        else:
            s = _allreduce(backend.attempt_rk(tab, ev, dt_32, atol32, rtol32, final, x_interp), group).cpu().numpy()
        st.nfe += nst
        st.launches += 2
        n_steps += 1
        if s[L.P_NONFINITE] > 0:
            raise SolverError("non-finite values in state `y`")
        with np.errstate(all="ignore"):
            ratio = abs(norm(s[L.P_X_ERR], s[L.P_LP_ERR], 0.0))
            accept = bool(ratio <= 1)
            if dts > max_step:
                accept = False
            if dts <= min_step:
                accept = True
            st.dt_history.append(float(dts)); st.accept_history.append(accept); st.ratio_history.append(float(ratio))
            if accept:
                st.accepted += 1
                backend.accept()
                if on_grid and grid_idx != len(grid) - 1:
                    grid_idx += 1
                if on_jump:
                    if jump_idx != len(jumps) - 1:
                        jump_idx += 1
                    backend.refresh_f(rows(np.array([np.nextafter(t1_32, t1_32 + f32(1))], f32))[0])
                    st.nfe += 1
                    st.launches += 1
                t = t1s
                done = final
            else:
                st.rejected += 1
            # controller (order 5)
            if ratio == 0:
                nxt_dt = dts * ifactor
            else:
                dfac = f64(1.0) if ratio < 1 else dfactor
                r = f64(ratio)
                nxt_dt = dts * np.minimum(ifactor, np.maximum(safety / r ** (f64(1.0) / f64(tab.order)), dfac))
            dt = f64(np.clip(nxt_dt, min_step, max_step)) if not np.isnan(nxt_dt) else f64(np.nan)
    assert done, "internal: integration loop ended without a final step"
    return st


def _rows_per_rank(counts, backend, group) -> float:
    """Global batch / world size: what the "auto" controller choice is based on, identical on every rank."""
    world = torch.distributed.get_world_size(group) if group is not None else 1
    return counts["x"] / max(1, getattr(backend, "D", 1)) / world


def _dopri5_device(backend, params, st: SolveStats, ts: float, dt: float, grid_idx: int, atol32, rtol32, group) -> SolveStats:
    """The adaptive loop with the controller on the device: per attempt the host enqueues
    attempt -> tile reduction -> (all-reduce over ranks) -> control and follows the controller's turns through
    pinned host memory, _CTL_AHEAD attempts behind what it has enqueued.  The outcome of turn k is the same on
    every rank (bit-identical sums), so every rank leaves the loop after the same number of collectives."""
    st.controller = "device"
    backend.ctl_begin(params, ts, dt, grid_idx, atol32, rtol32)
    st.launches += 1
    k = 0
    while True:
        if group is None:                       # tile reduction folded into the controller's launch
            backend.ctl_attempt(reduce=False)
            backend.ctl_control(reduce=True)
            st.launches += 2
        else:
            _allreduce(backend.ctl_attempt(reduce=True), group)
            backend.ctl_control(reduce=False)
            st.launches += 3
        k += 1
        # stay _CTL_AHEAD attempts ahead of the last turn whose outcome is known: the device always has work queued
        # and at most _CTL_AHEAD launches are wasted (they return at once) after the solve has ended
        if k > _CTL_AHEAD and backend.ctl_wait_turn(k - _CTL_AHEAD) != L.CTL_RUNNING:
            break
    c = backend.ctl_finish()
    n = int(c.n_attempts)
    st.nfe += 6 * n
    st.accepted, st.rejected = int(c.n_accepted), int(c.n_rejected)
    st.history_truncated = n > L.CTL_HIST
    for i in range(min(n, L.CTL_HIST)):
        st.dt_history.append(float(c.hist_dt[i]))
        st.accept_history.append(bool(c.hist_accept[i]))
        st.ratio_history.append(float(c.hist_ratio[i]))
    if c.done == L.CTL_NONFINITE:
        raise SolverError("non-finite values in state `y`")
    if c.done == L.CTL_DT_UNDERFLOW:
        raise SolverError("underflow in dt {}".format(float(c.dt_next)))
    if c.done == L.CTL_MAX_STEPS:
        raise SolverError("max_num_steps exceeded ({}>={})".format(n, int(params.max_num_steps)))
    assert c.done == L.CTL_FINISHED, "internal: device controller ended in state {}".format(int(c.done))
    return st


# ----------------------------------------------------------------------------------------------
# fixed grids (torchdiffeq FixedGridODESolver): time stays in float32
# ----------------------------------------------------------------------------------------------
def fixed_grid(t0: float, t1: float, step_size: Optional[float], grid_constructor=None, y0=None) -> torch.Tensor:
    """Solver-time grid (ascending) as torchdiffeq's FixedGridODESolver builds it: from ``options['step_size']``, from
    ``options['grid_constructor'](func, y0, t)`` (called with the solver's ascending ``t``; ``func`` is None here -- the
    field lives in the kernels), or just ``t`` (one step)."""
    t = torch.tensor([t0, t1], dtype=torch.float32)
    if t[0] > t[1]:
        t = -t
    if step_size is not None and grid_constructor is not None:
        raise ValueError("step_size and grid_constructor are mutually exclusive arguments.")
    if grid_constructor is not None:
        g = torch.as_tensor(grid_constructor(None, y0, t), dtype=torch.float32).detach().cpu().reshape(-1)
        assert g[0] == t[0] and g[-1] == t[-1]          # torchdiffeq's own assertion
        if not bool((g[1:] > g[:-1]).all()):
            raise ValueError("grid_constructor must return a strictly increasing grid")
        return g
    if step_size is None:
        return t
    n = torch.ceil((t[-1] - t[0]) / step_size + 1).item()
    g = torch.arange(0, n, dtype=t.dtype) * step_size + t[0]
    g[-1] = t[-1]
    return g


def fixed_eval_times(method: str, grid: torch.Tensor, perturb: bool = False):
    """(dt, eval_times[nsteps, nev]) in float32, op order of torchdiffeq's step functions.  ``perturb`` (options['perturb']):
    the first evaluation of a step is made one ulp AFTER t0 and rk4's last one one ulp BEFORE t1 (Perturb.NEXT / PREV)."""
    a, b = grid[:-1], grid[1:]
    dt = b - a
    first = torch.nextafter(a, a + 1) if perturb else a
    if method == "euler":
        times = first[:, None]
    elif method == "midpoint":
        times = torch.stack([first, a + 0.5 * dt], dim=1)
    elif method == "rk4":
        last = torch.nextafter(b, b - 1) if perturb else b
        times = torch.stack([first, a + dt * (1 / 3), a + dt * (2 / 3), last], dim=1)
    else:
        raise ValueError(method)
    return dt, times
