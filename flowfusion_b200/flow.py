"""Drop-in surface of ``flowfusion.flow`` (flow-matching ODE sampling and exact log-likelihood).

``ODEFlow`` / ``ConditionalODEFlow`` keep the reference's constructor signatures, attributes
and ``state_dict`` keys (every Linear appears under ``layers.*`` and ``velocity.*``,
`flow.py:63-74`).  The hot path runs in ``libffb200.so``:

* ``sample``            `flow.py:259-306, 750-799`   dopri5 with torchdiffeq's DEFAULT tolerances
                                                       (rtol 1e-7, atol 1e-9), t: 1 -> 0
* ``solve_ode_forward`` `flow.py:308-384, 801-883`   state (x[, cond], logJ), t: 0 -> 1; the exact
                                                       divergence is D forward-mode tangent rows
                                                       carried through the same fused MLP tile
* ``log_prob``          `flow.py:386-438, 885-941`   + base density and -sum(log scale)

Training-side members (``flow_matching_loss``, ``compute_linear_velocity_field``, ``gradients=True``
/ ``adjoint=True``) are out of scope and raise ``NotImplementedError``.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import dist as _dist
from . import engine as E
from . import solver as S
from .diffusion import _IMPLEMENTED, _METHODS, _eval_once, _solve_fixed


def _raw_time_program(times32: np.ndarray) -> np.ndarray:
    """Flows feed t itself as one input column (`flow.py:112-115`)."""
    rows = np.zeros((times32.shape[0], L.EV_FLOATS), np.float32)
    rows[:, 0] = times32
    rows[:, L.MAX_TFEAT + 3] = 1.0
    return rows


def _raw_time_spec():
    spec = L.TimeProgram()
    spec.time_features, spec.sde, spec.T = L.PROG_RAW_T, L.SDE_NONE, 1.0
    return spec


_raw_time_program.spec = _raw_time_spec()       # the same program for the device-side dopri5 controller


class _FlowBase(nn.Module):
    def _build(self, target_dimension, conditional_dimension, hidden_units, activation, target_shift, target_scale):
        self.target_dimension = target_dimension
        self.layers = nn.ModuleList()
        arch = [target_dimension + 1 + conditional_dimension] + list(hidden_units) + [target_dimension]
        for i in range(len(arch) - 2):
            self.layers.append(nn.Linear(arch[i], arch[i + 1]))
            self.layers.append(activation())
        self.layers.append(nn.Linear(arch[-2], arch[-1]))
        self.velocity = nn.Sequential(*self.layers)
        self.register_buffer("twopi", torch.tensor(2.0 * 3.14159265358979323846))
        self.register_buffer("target_shift", target_shift if target_shift is not None else torch.zeros(target_dimension))
        self.register_buffer("target_scale", target_scale if target_scale is not None else torch.ones(target_dimension))
        self._packed = None
        self.process_group = None
        self.last_stats = None

    def _cdim(self):
        return getattr(self, "conditional_dimension", 0)

    def _net(self) -> E.PackedNet:
        lin = [m for m in self.layers if isinstance(m, nn.Linear)]
        act = E.activation_of(self.layers)
        key = E.weights_fingerprint(lin, act)
        if self._packed is None or self._packed[0] != key:
            D, Cn = self.target_dimension, self._cdim()
            dev = lin[0].weight.device
            E.require_cuda_device(dev)
            self._packed = (key, E.PackedNet(lin, x_col=0, x_dim=D, c_col=D + 1, c_dim=Cn, t_col=D, t_dim=1, device=dev,
                                             activation=act))
        return self._packed[1]

    def _field(self, div_mode=L.DIV_NONE):
        return E.FieldSpec([self._net()], self.target_dimension, self._cdim(), kind=L.FIELD_NET, div_mode=div_mode)

    def _group(self):
        return self.process_group if self.process_group is not None else _dist.current_group()

    def _row(self, t):
        tt = torch.as_tensor(t, dtype=torch.float32).detach().cpu().reshape(-1)[:1].numpy()
        return _raw_time_program(tt)[0]

    def _integrate(self, y0, cond_net, cond_state, t0, t1, atol, rtol, method, options, div_mode):
        field = self._field(div_mode)
        with_lp = div_mode != L.DIV_NONE
        method = "dopri5" if method is None else method
        if method in S.ADAPTIVE_METHODS:
            be = E.CudaBackend(field, y0, cond=cond_net, with_lp=with_lp, cond_in_state=cond_state is not None,
                               cond_state=cond_state)
            self.last_stats = S.adaptive(method, be, _raw_time_program, t0, t1, rtol, atol, options, group=self._group())
            return be.output()
        if method in _METHODS:
            return _solve_fixed(field, _raw_time_program, method, y0, cond_net, None, t0, t1, options, with_lp)
        raise NotImplementedError(f"method {method!r} is not implemented ({_IMPLEMENTED})")

    def _base_logprob(self, xT, log_jacobian):
        lp = E.gaussian_logprob(xT, log_jacobian, 1.0)
        return lp - torch.sum(torch.log(self.target_scale))

    # training side ---------------------------------------------------------------------------
    def compute_linear_velocity_field(self, x0, xT, t):
        """`flow.py:191-224` / `:679-714`."""
        from . import training
        return training.compute_linear_velocity_field(self, x0, xT, t)

    def flow_matching_loss(self, x, conditional=None, **draws):
        """`flow.py:226-256` (x) / `:716-747` (x, conditional): fused forward + backward (csrc/ffb_train.cu); ``xT=`` /
        ``t=`` replay the draws."""
        from . import training
        if (conditional is None) != (self._cdim() == 0):
            raise TypeError("flow_matching_loss: a conditional flow takes (x, conditional), an unconditional one (x)")
        return training.flow_matching_loss(self, x, conditional, **draws)


class ODEFlow(_FlowBase):
    """Unconditional flow-matching model (`flow.py:9-438`)."""

    def __init__(self, target_dimension: int = 1, hidden_units: List[int] = [128, 128],
                 activation: nn.Module = nn.SiLU, target_shift: Optional[torch.Tensor] = None,
                 target_scale: Optional[torch.Tensor] = None):
        super().__init__()
        self._build(target_dimension, 0, hidden_units, activation, target_shift, target_scale)

    def dynamics(self, t: torch.Tensor, states: Tuple[torch.Tensor]):
        """dx/dt = velocity(cat[x, t]) (`flow.py:89-120`), one kernel evaluation."""
        return _eval_once(self._field(), self._row(t), states[0], None)[0]

    def dynamics_with_jacobian(self, t, states):
        """(dx/dt, divergence) (`flow.py:122-166`)."""
        f, d = _eval_once(self._field(L.DIV_EXACT), self._row(t), states[0], None)
        return f, d.view(-1, 1)

    def forward(self, t, states):
        return self.dynamics(t, states)

    def sample(self, xT: torch.Tensor, gradients: bool = False):
        """`flow.py:259-306`: integrate 1 -> 0 with torchdiffeq's default tolerances.  ``gradients=True`` is the reference's
        ``odeint_adjoint`` branch (`:286-295`): the samples are attached to ``xT`` and the weights (adjoint.py)."""
        E.require_cuda(xT, "xT")
        if gradients:
            from . import adjoint

            def solve(y):
                return self._integrate(y, None, None, 1.0, 0.0, 1e-9, 1e-7, None, None, L.DIV_NONE)[0]
            x = adjoint.solve_with_adjoint(self, adjoint._FlowField(self), solve, xT, 1.0, 0.0, 1e-7, 1e-9, None, None)
            return x * self.target_scale + self.target_shift
        with torch.no_grad():
            x, _ = self._integrate(xT, None, None, 1.0, 0.0, 1e-9, 1e-7, None, None, L.DIV_NONE)
            return x * self.target_scale + self.target_shift

    def solve_ode_forward(self, x, atol=1e-5, rtol=1e-5, method="dopri5", options=None, adjoint=False):
        """`flow.py:308-384` -> (x(T), log-Jacobian (B, 1))."""
        if adjoint:
            raise NotImplementedError("adjoint=True of the log-likelihood solve is not implemented (it needs second derivatives of the network); sample(gradients=True) is")
        E.require_cuda(x, "x")
        with torch.no_grad():
            xT, lj = self._integrate(x, None, None, 0.0, 1.0, atol, rtol, method, options, L.DIV_EXACT)
        return xT, lj.view(-1, 1)

    def log_prob(self, x, atol=1e-5, rtol=1e-5, method="dopri5", options=None, adjoint=False):
        """`flow.py:386-438` -> (B,)."""
        x = (x - self.target_shift) / self.target_scale
        xT, lj = self.solve_ode_forward(x, atol, rtol, method, options, adjoint)
        return self._base_logprob(xT, lj.reshape(-1))


class ConditionalODEFlow(_FlowBase):
    """Conditional flow-matching model (`flow.py:441-941`); the conditional rides along in the
    ODE state with zero derivative, so it takes part in torchdiffeq's initial-step norm."""

    def __init__(self, target_dimension: int = 1, conditional_dimension: int = 1,
                 hidden_units: List[int] = [128, 128], activation: nn.Module = nn.SiLU,
                 target_shift: Optional[torch.Tensor] = None, target_scale: Optional[torch.Tensor] = None,
                 conditional_shift: Optional[torch.Tensor] = None, conditional_scale: Optional[torch.Tensor] = None):
        super().__init__()
        self.conditional_dimension = conditional_dimension
        self._build(target_dimension, conditional_dimension, hidden_units, activation, target_shift, target_scale)
        self.register_buffer("conditional_shift", conditional_shift if conditional_shift is not None
                             else torch.zeros(conditional_dimension))
        self.register_buffer("conditional_scale", conditional_scale if conditional_scale is not None
                             else torch.ones(conditional_dimension))

    def _norm_cond(self, conditional):
        return (conditional - self.conditional_shift) / self.conditional_scale      # `flow.py:580`

    def dynamics(self, t, states):
        """`flow.py:553-596` -> (dx/dt, zeros_like(conditional))."""
        x, conditional = states
        f = _eval_once(self._field(), self._row(t), x, self._norm_cond(conditional))[0]
        return f, torch.zeros_like(conditional)

    def dynamics_with_jacobian(self, t, states):
        """`flow.py:598-652`."""
        x, conditional, _ = states
        f, d = _eval_once(self._field(L.DIV_EXACT), self._row(t), x, self._norm_cond(conditional))
        return f, torch.zeros_like(conditional), d.view(-1, 1)

    def forward(self, t, states):
        return self.dynamics(t, states)

    def sample(self, xT, conditional, gradients: bool = False):
        """`flow.py:750-799`."""
        E.require_cuda(xT, "xT")
        if gradients:                              # `flow.py:775-785`: odeint_adjoint over the state (xT, conditional)
            from . import adjoint

            def solve(y):
                return self._integrate(y, self._norm_cond(conditional.detach()), conditional.detach(), 1.0, 0.0, 1e-9, 1e-7,
                                       None, None, L.DIV_NONE)[0]
            x = adjoint.solve_with_adjoint(self, adjoint._CondFlowField(self), solve, xT, 1.0, 0.0, 1e-7, 1e-9, None, None,
                                           static=conditional)
            return x * self.target_scale + self.target_shift
        with torch.no_grad():
            x, _ = self._integrate(xT, self._norm_cond(conditional), conditional, 1.0, 0.0, 1e-9, 1e-7, None, None,
                                   L.DIV_NONE)
            return x * self.target_scale + self.target_shift

    def solve_ode_forward(self, x, conditional, atol=1e-5, rtol=1e-5, method="dopri5", options=None, adjoint=False):
        """`flow.py:801-883`."""
        if adjoint:
            raise NotImplementedError("adjoint=True of the log-likelihood solve is not implemented (it needs second derivatives of the network); sample(gradients=True) is")
        E.require_cuda(x, "x")
        with torch.no_grad():
            xT, lj = self._integrate(x, self._norm_cond(conditional), conditional, 0.0, 1.0, atol, rtol, method,
                                     options, L.DIV_EXACT)
        return xT, lj.view(-1, 1)

    def log_prob(self, x, conditional, atol=1e-5, rtol=1e-5, method="dopri5", options=None, adjoint=False):
        """`flow.py:885-941` -> (B,)."""
        x = (x - self.target_shift) / self.target_scale
        xT, lj = self.solve_ode_forward(x, conditional, atol, rtol, method, options, adjoint)
        return self._base_logprob(xT, lj.reshape(-1))
