"""Drop-in surface of ``flowfusion.diffusion`` for the sampling / density-evaluation path.

Same class names, constructor signatures, attribute names and ``state_dict`` keys as the
reference (`flowfusion/diffusion.py`), so weights trained with the reference load unchanged.
Every entry point of the hot path runs in the sm_100a kernels of ``libffb200.so``:

==============================  =============================  ==================================
this module                     reference                      kernel
==============================  =============================  ==================================
``ScoreModel.sample_sde``       `diffusion.py:510-563`         ``ffb_integrate_fixed`` (EM)
``.sample_ode_from_base``       `diffusion.py:566-640`         ``ffb_dopri5_attempt`` / fixed grid
``.solve_odes_forward``         `diffusion.py:642-754`         same, with tangent rows (trace)
``.log_prob``                   `diffusion.py:756-815`         + ``ffb_gaussian_logprob``
``.forward(t, states)``         `diffusion.py:281-508`         ``ffb_field_eval``
==============================  =============================  ==================================

Out of scope here (training side): the score-matching losses, adjoint
back-propagation (``training=True``).  It raises ``NotImplementedError``.
Tensors must live on a CUDA device; there is no CPU path.
"""
from __future__ import annotations

import copy

import numpy as np
import torch
from torch.distributions import Normal

from . import _lib as L
from . import dist as _dist
from . import engine as E
from . import solver as S


# ----------------------------------------------------------------------------------------------
# score network (`diffusion.py:9-121`)
# ----------------------------------------------------------------------------------------------
class MLP(torch.nn.Module):
    """Score network: input ``cat[sin(2 pi t W), cos(2 pi t W), x, conditional]``."""

    def __init__(self, n_dimensions=2, n_conditionals=1, embedding_dimensions=8, units=[128],
                 activation=torch.nn.SiLU(), sigma_initialization=16):
        super().__init__()
        self.n_dimensions = n_dimensions
        self.n_conditionals = n_conditionals
        self.architecture = [n_dimensions + n_conditionals + embedding_dimensions] + list(units) + [n_dimensions]
        self.n_layers = len(self.architecture) - 1
        # parameter creation order matches the reference so torch.manual_seed gives equal weights
        self.NN = torch.nn.ModuleList(
            [torch.nn.Linear(self.architecture[i], self.architecture[i + 1]) for i in range(self.n_layers)])
        self.W = torch.nn.Parameter(torch.randn(embedding_dimensions // 2) * sigma_initialization,
                                    requires_grad=False)
        self.activation = activation
        self.register_buffer("pi", torch.tensor(np.pi, dtype=torch.float32))
        self._packed = None

    # -- kernel plumbing ----------------------------------------------------------------------
    @property
    def embedding_dimensions(self):
        return 2 * self.W.shape[0]

    def _net(self) -> E.PackedNet:
        act = E.activation_code(self.activation)
        lin = list(self.NN)
        key = E.weights_fingerprint(lin, act)
        if self._packed is None or self._packed[0] != key:
            emb, D, Cn = self.embedding_dimensions, self.n_dimensions, self.n_conditionals
            if emb + D + Cn != lin[0].in_features:
                raise ValueError("embedding_dimensions must be even (reference layout sin|cos)")
            dev = lin[0].weight.device
            E.require_cuda_device(dev)
            self._packed = (key, E.PackedNet(lin, x_col=emb, x_dim=D, c_col=emb + D, c_dim=Cn, t_col=0,
                                             t_dim=emb, device=dev, activation=act))
        return self._packed[1]

    def _host_embedding(self):
        """(W, pi) on the host; a device->host copy stalls the stream, so it is repeated only when W was modified."""
        key = (self.W.data_ptr(), self.W._version, self.pi.data_ptr(), self.pi._version, str(self.W.device))
        cached = getattr(self, "_emb_host", None)
        if cached is None or cached[0] != key:
            cached = (key, self.W.detach().cpu(), self.pi.detach().cpu())
            self._emb_host = cached
        return cached[1], cached[2]

    def _time_features(self, t32: torch.Tensor, W_cpu=None, pi_cpu=None) -> torch.Tensor:
        """(n,) float32 CPU times -> (n, emb) features, op order of `diffusion.py:109-110`.
        Solver loops pass host copies of W and pi (a device->host copy per call would stall the stream)."""
        W = self.W.detach().cpu() if W_cpu is None else W_cpu
        pi = self.pi.detach().cpu() if pi_cpu is None else pi_cpu
        proj = t32[:, None] * W[None, :] * 2 * pi
        return torch.cat([torch.sin(proj), torch.cos(proj)], dim=1)

    def forward(self, t, x, conditional=None):
        """Network output at a batch-uniform time (`diffusion.py:82-121`), on the GPU kernels."""
        E.require_cuda(x, "x")
        tt = torch.as_tensor(t, dtype=torch.float32).detach()
        if tt.dim() > 0:
            if not bool((tt == tt.reshape(-1)[0]).all()):
                # one time per sample (`diffusion.py:104-113`): the network's input rows are built as the reference builds
                # them and evaluated by the forward-only mode of the fused training kernel
                from . import training
                tt = tt.to(x.device).reshape(-1)
                proj = tt[:, None] * self.W[None, :] * 2 * self.pi
                cols = [torch.sin(proj), torch.cos(proj), x] + ([conditional] if conditional is not None else [])
                return training.mlp_forward(list(self.NN), E.activation_code(self.activation), torch.cat(cols, dim=1))
            tt = tt.reshape(-1)[0]
        field = E.FieldSpec([self._net()], self.n_dimensions, self.n_conditionals)
        rows = np.zeros((1, L.EV_FLOATS), np.float32)
        rows[0, : self.embedding_dimensions] = self._time_features(tt.cpu().reshape(1)).numpy()[0]
        rows[0, L.MAX_TFEAT + 3] = 1.0
        return _eval_once(field, rows[0], x, conditional)[0]


def _eval_once(field, ev_row, x, cond, probes=None):
    be = E.CudaBackend(field, x, cond=cond, probes=probes, with_lp=field.div_mode != L.DIV_NONE)
    return be.single_eval(ev_row)


# ----------------------------------------------------------------------------------------------
# SDEs (`diffusion.py:818-1366`): closed forms, same buffers / attributes as the reference
# ----------------------------------------------------------------------------------------------
def _bcast(v, x):
    return v.view(-1, *[1] * (x.dim() - 1))


class VESDE(torch.nn.Module):
    """Variance-exploding SDE (`diffusion.py:818-1003`)."""

    def __init__(self, sigma_min=1e-2, sigma_max=10.0, T=1.0, epsilon=1e-5):
        super().__init__()
        for name, v in (("T", T), ("epsilon", epsilon), ("sigma_min", sigma_min), ("sigma_max", sigma_max)):
            self.register_buffer(name, torch.tensor(v, dtype=torch.float32))

    def sigma(self, t):
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** (t / self.T)

    def _g(self, t):
        return self.sigma(t) * torch.sqrt(2 * (torch.log(self.sigma_max) - torch.log(self.sigma_min)) / self.T)

    def diffusion(self, t, x):
        return _bcast(self.sigma(t), x) * torch.sqrt(
            2 * (torch.log(self.sigma_max) - torch.log(self.sigma_min)) / self.T)

    def drift(self, t, x):
        return torch.zeros_like(x)

    def _drift_coeff(self, t):
        return None                       # no linear drift (`:905`)

    def marginal_prob_scalars(self, t):
        return torch.ones_like(t), self.sigma(t)

    def marginal_prob(self, t, x):
        m, s = self.marginal_prob_scalars(t)
        return _bcast(m, x) * x, _bcast(s, x)

    def sample_marginal(self, t, x0):
        m, s = self.marginal_prob_scalars(t)
        return _bcast(m, x0) * x0 + _bcast(s, x0) * torch.randn_like(x0)

    def prior(self, shape, mu=None):
        if mu is None:
            mu = torch.zeros(shape).to(self.T.device)
        else:
            assert mu.shape == shape
        return Normal(loc=mu, scale=self.sigma_max)

    def _prior_sigma(self):
        return float(self.sigma_max)


class _BetaSDE(torch.nn.Module):
    def __init__(self, beta_min=0.1, beta_max=20, T=1.0, epsilon=1e-3):
        super().__init__()
        self.beta_min, self.beta_max, self.T = beta_min, beta_max, T
        self.register_buffer("epsilon", torch.tensor(epsilon, dtype=torch.float32))

    def beta(self, t):
        return self.beta_min + (self.beta_max - self.beta_min) * (t / self.T)

    def _log_coeff(self, t):
        return 0.5 * (self.beta_max - self.beta_min) * t ** 2 / self.T + self.beta_min * t

    def sigma(self, t):
        return self.marginal_prob_scalars(t)[1]

    def prior(self, shape):
        return Normal(loc=torch.zeros(shape).to(self.epsilon.device), scale=1.0)

    def diffusion(self, t, x):
        return _bcast(self._g(t), x)

    def drift(self, t, x):
        return -0.5 * _bcast(self.beta(t), x) * x

    def _drift_coeff(self, t):
        return -0.5 * self.beta(t)

    def marginal_prob(self, t, x):
        m, s = self.marginal_prob_scalars(t)
        return _bcast(m, x) * x, _bcast(s, x)

    def _prior_sigma(self):
        return 1.0


class VPSDE(_BetaSDE):
    """Variance-preserving SDE (`diffusion.py:1006-1180`)."""

    def _g(self, t):
        return torch.sqrt(self.beta(t))

    def marginal_prob_scalars(self, t):
        lc = self._log_coeff(t)
        return torch.exp(-0.5 * lc), torch.sqrt(1.0 - torch.exp(-lc))


class SUBVPSDE(_BetaSDE):
    """Sub-VP SDE (`diffusion.py:1183-1366`)."""

    def _g(self, t):
        return torch.sqrt(self.beta(t) * (1.0 - torch.exp(
            -2 * self.beta_min * t - (self.beta_max - self.beta_min) * t ** 2 / self.T)))

    def marginal_prob_scalars(self, t):
        lc = self._log_coeff(t)
        return torch.exp(-0.5 * lc), 1.0 - torch.exp(-lc)


def fourier_program_spec(W_cpu: torch.Tensor, pi32: float):
    """ffb_time_program of `sin|cos(((t*W)*2)*pi)` features, or None when there are too many frequencies."""
    n = int(W_cpu.numel())
    if n > L.MAX_FREQ:
        return None
    spec = L.TimeProgram()
    spec.time_features, spec.n_freq, spec.pi = L.PROG_FOURIER, n, float(np.float32(pi32))
    for k in range(n):
        spec.W[k] = float(W_cpu.reshape(-1)[k])
    spec.sde, spec.T = L.SDE_NONE, 1.0
    return spec


def _device_program(sde, W_cpu, pi_cpu, use_sigma, sde_mode):
    """The same scalars as ``program`` for the device-side dopri5 controller (csrc/ffb_control.cuh), with every
    Python scalar rounded to FP32 where the reference's eager ops would round it; None for an SDE class the
    controller does not restate (the host loop is used then)."""
    spec = fourier_program_spec(W_cpu.float(), float(pi_cpu))
    if spec is None:
        return None
    f32 = lambda v: float(np.float32(v))
    spec.use_sigma, spec.sde_mode = int(use_sigma), int(sde_mode)
    # sigma(t) of VP / subVP and g(t) of subVP are 1 - exp(-small) near t = epsilon: one ulp of exp() is ~3e-4 of the
    # result there, so these stay on the host program, which uses the reference's own math library
    if type(sde) is SUBVPSDE or (type(sde) is VPSDE and use_sigma):
        return None
    if type(sde) in (VPSDE, SUBVPSDE):
        spec.sde = L.SDE_VP if type(sde) is VPSDE else L.SDE_SUBVP
        bmin, bmax = float(sde.beta_min), float(sde.beta_max)
        spec.T, spec.beta_min, spec.beta_diff = f32(sde.T), f32(bmin), f32(bmax - bmin)
        spec.half_beta_diff, spec.m2_beta_min = f32(0.5 * (bmax - bmin)), f32(-2 * bmin)
    elif type(sde) is VESDE:
        spec.sde = L.SDE_VE
        spec.T, spec.sigma_min = float(sde.T), float(sde.sigma_min)
        spec.sigma_ratio = float(sde.sigma_max / sde.sigma_min)
        spec.ve_gfac = float(torch.sqrt(2 * (torch.log(sde.sigma_max) - torch.log(sde.sigma_min)) / sde.T))
    else:
        return None
    return spec


# ----------------------------------------------------------------------------------------------
# ScoreModel (`diffusion.py:124-815`)
# ----------------------------------------------------------------------------------------------
_METHODS = {"euler": L.M_EULER, "midpoint": L.M_MIDPOINT, "rk4": L.M_RK4}
_IMPLEMENTED = "dopri5, bosh3, adaptive_heun, fehlberg2, rk4, euler, midpoint"


class ScoreModel(torch.nn.Module):
    def __init__(self, model=None, sde=None, conditional=None, no_sigma=False, hutchinson=False, hutchpp=False,
                 hpp_rank=1, hpp_vecs=1, xtrace=False, xt_vecs=1):
        super().__init__()
        self.model = model
        self.sde = sde
        self.conditional = conditional
        self.no_sigma = no_sigma
        self.prob = False
        self.hutch = hutchinson
        self.hutchpp = hutchpp
        self.hpp_rank = hpp_rank
        self.hpp_vector = hpp_vecs
        self.xtrace = xtrace
        self.xt_vector = xt_vecs
        self.process_group = None          # set (or use dist.use_group) to shard the batch over ranks
        self.last_stats = None

    # -- field / program --------------------------------------------------------------------
    def _field(self, div_mode=L.DIV_NONE):
        m = self.model
        return E.FieldSpec([m._net()], m.n_dimensions, m.n_conditionals, kind=L.FIELD_SCORE,
                           use_sigma=not self.no_sigma, has_drift=not isinstance(self.sde, VESDE),
                           div_mode=div_mode)

    def _program(self, sde_mode=False):
        """times32 -> rows of host scalars, in the reference's FP32 op order
        (`diffusion.py:276-278` for the PF-ODE, `:552-553` for the reverse SDE)."""
        sde = copy.deepcopy(self.sde).cpu()
        model, emb = self.model, self.model.embedding_dimensions
        W_cpu, pi_cpu = model._host_embedding()     # host copies, refreshed only when W changes
        use_sigma = not self.no_sigma

        def program(times32: np.ndarray) -> np.ndarray:
            t = torch.from_numpy(np.ascontiguousarray(times32, np.float32))
            rows = np.zeros((t.shape[0], L.EV_FLOATS), np.float32)
            rows[:, :emb] = model._time_features(t, W_cpu, pi_cpu).numpy()
            g = sde._g(t)
            a = sde._drift_coeff(t)
            rows[:, L.MAX_TFEAT + 0] = 0.0 if a is None else a.numpy()
            rows[:, L.MAX_TFEAT + 1] = (g ** 2 if sde_mode else 0.5 * g ** 2).numpy()
            rows[:, L.MAX_TFEAT + 2] = sde.sigma(t).numpy() if use_sigma else 1.0    # unused by the kernels otherwise
            rows[:, L.MAX_TFEAT + 3] = 1.0
            return rows

        program.g = lambda t: sde._g(t)
        program.spec = _device_program(sde, W_cpu, pi_cpu, use_sigma, sde_mode)
        return program

    def _group(self):
        return self.process_group if self.process_group is not None else _dist.current_group()

    def _check_supported(self):
        if self.training:
            raise NotImplementedError("training=True selects odeint_adjoint: gradients through the solver are implemented for "
                                      "sample_ode_from_base only (the log-likelihood paths would need second derivatives "
                                      "of the network) -- call .eval() first")

    # -- reference API ------------------------------------------------------------------------
    def score(self, t, x, conditional=None):
        """`diffusion.py:215-238`."""
        out = self.model(t, x, conditional=conditional)
        if self.no_sigma:
            return out
        tt = torch.as_tensor(t, dtype=torch.float32, device=x.device)
        return out / _bcast(self.sde.sigma(tt).reshape(-1), x)

    def loss_fn(self, x, conditional=None):
        """`diffusion.py:240-256`: the denoising score matching loss, fused forward + backward (training.py)."""
        return denoising_score_matching(self, x, conditional=conditional)

    def ode_drift(self, t, x, conditional=None):
        """`diffusion.py:258-279`, one kernel evaluation."""
        tt = torch.as_tensor(t, dtype=torch.float32).detach().cpu().reshape(-1)[:1]
        row = self._program()(tt.numpy())[0]
        return _eval_once(self._field(), row, x, conditional)[0]

    def forward(self, t, states):
        """`diffusion.py:281-508`: dx/dt, and d(log p)/dt when ``self.prob`` (exact or Hutchinson)."""
        self._check_supported()
        x = states[0]
        tt = torch.as_tensor(t, dtype=torch.float32).detach().cpu().reshape(-1)[:1]
        row = self._program()(tt.numpy())[0]
        if not self.prob:
            return _eval_once(self._field(), row, x, self.conditional)[0]
        est = None if self.hutch else self._estimator(x, stored=True)
        if est is not None:                 # Hutch++ / XTrace (`:336-481`): staged evaluation
            be = E.StagedBackend(self._field(L.DIV_EXACT), x, est, cond=self.conditional)
            f, d = be.single_eval(row)
            return f, d.view(-1, 1)
        div = L.DIV_HUTCH if self.hutch else L.DIV_EXACT
        f, d = _eval_once(self._field(div), row, x, self.conditional, probes=self.e if self.hutch else None)
        return f, d.view(-1, 1)

    def _estimator(self, x, stored=False, probes=None):
        """The Hutch++ / XTrace estimator of this solve with its probes, or None (exact / Hutchinson).  Flag priority
        and probe shapes follow `diffusion.py:327, 336, 402` and `:703-721`; ``stored=True`` reuses the probes of the
        last solve when their shapes fit (`:346-354`, `:413-418`), else draws fresh ones."""
        if self.hutch or not (self.hutchpp or self.xtrace):
            return None
        B, D = x.shape[0], int(np.prod(x.shape[1:]))
        rs = lambda n: torch.sign(torch.randn(n, B, D, device=x.device, dtype=x.dtype))     # noqa: E731
        if self.hutchpp:
            r, m = int(min(self.hpp_rank, D)), int(max(1, self.hpp_vector))
            if probes is not None:
                self.S, self.G = probes
            elif not stored or getattr(self, "S", None) is None or self.S.shape != (r, B, D) \
                    or getattr(self, "G", None) is None or self.G.shape != (m, B, D):
                self.S, self.G = rs(r), rs(m)
            if tuple(self.S.shape) != (r, B, D) or tuple(self.G.shape) != (m, B, D):
                raise ValueError("Hutch++ probes must have shapes (min(hpp_rank, D), B, D) and (max(1, hpp_vecs), B, D)")
            return E.TraceEstimator(L.TRACE_HUTCHPP, self.S, self.G)
        m = int(min(max(1, self.xt_vector), D))      # `:410`; solve_odes_forward's own draw (`:718`) omits the min
        if probes is not None:
            self.O = probes
        elif not stored or getattr(self, "O", None) is None or self.O.shape != (m, B, D):
            self.O = rs(m)
        if tuple(self.O.shape) != (m, B, D):
            raise ValueError("XTrace probes must have shape (min(max(1, xt_vecs), D), B, D)")
        return E.TraceEstimator(L.TRACE_XTRACE, self.O)

    @torch.no_grad()
    def sample_sde(self, shape, conditional=None, steps=100, *, x0=None, noise=None, seed=None):
        """Reverse-SDE Euler-Maruyama from t=T to epsilon (`diffusion.py:510-563`); returns x_mean.

        Extensions (keyword-only): ``x0`` (B, D) prior draw and ``noise`` (steps, B, D) unit normals
        for bit-comparable runs; otherwise the prior is drawn with torch on the device and the
        per-step noise comes from the in-kernel Philox stream seeded by ``seed`` (default: drawn
        from torch's generator)."""
        batch, *dims = shape
        dev = next(self.model.parameters()).device
        sde = copy.deepcopy(self.sde).cpu()
        if x0 is None:
            x0 = torch.randn(batch, *dims, device=dev) * sde._prior_sigma()
        E.require_cuda(x0, "x0")
        # time grid exactly as the reference accumulates it: FP32 `t += dt` (`:539-559`)
        dt = -(sde.T - sde.epsilon) / steps
        dt32 = np.float32(float(dt))
        eps32 = np.float32(float(sde.epsilon))
        ts = np.empty(steps, np.float32)
        t = np.float32(float(torch.ones(1) * sde.T))
        n_eff = 0
        for k in range(steps):
            if t < eps32:                                  # `:548-551`
                break
            ts[k] = t
            t = np.float32(t + dt32)
            n_eff += 1
        ts = ts[:n_eff]
        prog = self._program(sde_mode=True)
        ev = prog(ts)
        ev[:, L.MAX_TFEAT + 3] = 1.0
        step_table = np.zeros((n_eff, L.STEP_STRIDE), np.float32)
        step_table[:, 0] = dt32
        step_table[:, 1] = prog.g(torch.from_numpy(ts)).numpy()
        step_table[:, 2] = float((-torch.as_tensor(dt, dtype=torch.float32)) ** (1.0 / 2.0))
        philox = None
        if noise is None:
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            philox = (seed, 0)
        else:
            noise = noise[:n_eff]
        if n_eff == 0:
            # the reference's loop breaks before the first assignment of x_mean and `return x_mean` raises (`:548-563`)
            raise UnboundLocalError("cannot access local variable 'x_mean' where it is not associated with a value")
        group = self._group()
        rank_off = _dist.row_offset(batch, group)
        field = self._field()

        def run(n):
            return E.run_fixed(field, L.M_EM, x0.reshape(batch, -1), step_table[:n], ev[:n].reshape(n, 1, L.EV_FLOATS),
                               cond=conditional, noise=None if noise is None else noise[:n], philox=philox,
                               row_offset=rank_off)

        x, _, status = run(n_eff)
        # `diffusion.py:560-563`: the reference stops the WHOLE batch after the first step that left a NaN anywhere in x and
        # returns that step's x_mean.  The kernel integrates every tile to the end and records the first such step; the
        # (rare) unstable case is then integrated again up to that step -- the noise is a function of (seed, row, step)
        # or the caller's tensor, so the second pass reproduces the first.
        flags, first_nan = (int(v) for v in status.cpu())
        if group is not None:                         # "anywhere in x" is global: every rank stops at the same step
            t = torch.tensor([first_nan if flags & L.ST_NAN_SAMPLE else 2 ** 31 - 1], dtype=torch.int64, device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=group)
            first_nan = int(t.item())
            flags |= L.ST_NAN_SAMPLE if first_nan < 2 ** 31 - 1 else 0
        self._status = status
        self.stopped_at_step = None
        if flags & L.ST_NAN_SAMPLE:
            print("Diffusion is not stable, NaN were produced. Stopped sampling.")
            self.stopped_at_step = first_nan
            if first_nan + 1 < n_eff:
                x, _, _ = run(first_nan + 1)
        return x.reshape(batch, *dims)

    def check_stability(self):
        """True when the last ``sample_sde`` produced no NaN (it prints the reference's message itself, `diffusion.py:560-562`)."""
        return not (int(self._status[0].item()) & L.ST_NAN_SAMPLE)

    def sample_ode_from_base(self, base_samples, conditional=None, atol=1e-4, rtol=1e-4, method="dopri5",
                             options=None):
        """Probability-flow ODE from t=1.0 to epsilon (`diffusion.py:566-640`).  Returns ``(x, [])``.  In training mode the
        reference solves with ``odeint_adjoint`` (`:620-629`): the samples are then attached to ``base_samples`` and to the
        network's weights, and their backward pass is the adjoint solve of ``flowfusion_b200/adjoint.py``."""
        E.require_cuda(base_samples, "base_samples")
        z = base_samples * self.sde.sigma_max if hasattr(self.sde, "sigma_max") else base_samples
        self.prob = False
        self.conditional = conditional
        if self.training:
            from . import adjoint
            if (method or "dopri5") not in S.ADAPTIVE_METHODS:
                raise NotImplementedError(f"training=True (odeint_adjoint) is implemented for {', '.join(S.ADAPTIVE_METHODS)}")
            if self._group() is not None:
                raise NotImplementedError("training=True (odeint_adjoint) is not implemented for a sharded batch")
            eps = float(self.sde.epsilon)

            def solve(y):
                return self._solve(y, conditional, 1.0, eps, atol, rtol, method, options, L.DIV_NONE, None)[0]
            x = adjoint.solve_with_adjoint(self, adjoint._ScoreField(self, conditional), solve, z, 1.0, eps, rtol, atol,
                                           method, options)
            return x, []
        x, _ = self._solve(z, conditional, 1.0, float(self.sde.epsilon), atol, rtol, method, options,
                           L.DIV_NONE, None)
        return x, []

    @torch.no_grad()
    def solve_odes_forward(self, x0_samples, conditional=None, atol=1e-5, rtol=1e-5, method="dopri5",
                           options=None, *, probes=None):
        """(x(T), delta log p) from epsilon to 1.0 (`diffusion.py:642-754`)."""
        self._check_supported()
        E.require_cuda(x0_samples, "x0_samples")
        self.prob = True
        if self.hutch:
            self.e = probes if probes is not None else torch.sign(torch.randn(x0_samples.shape)).to(x0_samples.device)
        self.conditional = conditional
        est = self._estimator(x0_samples, probes=probes)
        if est is not None:
            x, lp = self._solve_staged(x0_samples, conditional, est, float(self.sde.epsilon), 1.0, atol, rtol, method, options)
            return x, lp.view(-1, 1)
        x, lp = self._solve(x0_samples, conditional, float(self.sde.epsilon), 1.0, atol, rtol, method, options,
                            L.DIV_HUTCH if self.hutch else L.DIV_EXACT, self.e if self.hutch else None)
        return x, lp.view(-1, 1)

    def _solve_staged(self, y0, cond, est, t0, t1, atol, rtol, method, options):
        """Hutch++ / XTrace solves: evaluation-at-a-time dopri5 (engine.StagedBackend), host controller."""
        method = "dopri5" if method is None else method
        if method in _METHODS:
            be = E.StagedBackend(self._field(L.DIV_EXACT), y0, est, cond=cond)
            dt, ev = _fixed_tables(self._program(), method, t0, t1, options, y0)
            return E.staged_fixed(be, method, dt, ev)
        if method != "dopri5":
            raise NotImplementedError(f"method {method!r} is not implemented with Hutch++ / XTrace (dopri5, rk4, euler, midpoint)")
        be = E.StagedBackend(self._field(L.DIV_EXACT), y0, est, cond=cond)
        self.last_stats = S.dopri5(be, self._program(), t0, t1, rtol, atol, options, group=self._group())
        return be.output()

    @torch.no_grad()
    def log_prob(self, x0_samples, conditional=None, atol=1e-4, rtol=1e-4, method="dopri5",
                 options={"min_step": 1e-6}, *, probes=None):
        """`diffusion.py:756-815` -> (B, 1)."""
        xT, lp = self.solve_odes_forward(x0_samples, conditional=conditional, atol=atol, rtol=rtol, method=method,
                                         options=options, probes=probes)
        sig = copy.deepcopy(self.sde).cpu()._prior_sigma()
        return E.gaussian_logprob(xT, lp.reshape(-1), sig).view(-1, 1)

    # -- solver dispatch ------------------------------------------------------------------------
    def _solve(self, y0, cond, t0, t1, atol, rtol, method, options, div_mode, probes):
        field = self._field(div_mode)
        prog = self._program()
        with_lp = div_mode != L.DIV_NONE
        method = "dopri5" if method is None else method
        if method in S.ADAPTIVE_METHODS:          # dopri5 (fused attempt kernel), bosh3, adaptive_heun, fehlberg2 (stage by stage)
            be = E.CudaBackend(field, y0, cond=cond, probes=probes, with_lp=with_lp)
            self.last_stats = S.adaptive(method, be, prog, t0, t1, rtol, atol, options, group=self._group())
            return be.output()
        if method in _METHODS:
            return _solve_fixed(field, prog, method, y0, cond, probes, t0, t1, options, with_lp)
        raise NotImplementedError(f"method {method!r} is not implemented ({_IMPLEMENTED})")


def _fixed_tables(prog, method, t0, t1, options, y0=None):
    """Step sizes (n,) and evaluation scalars (n, evaluations per step, EV_FLOATS) of a torchdiffeq fixed grid
    (FixedGridODESolver: options step_size | grid_constructor, perturb; the solve returns the state AT t1, a grid point,
    so the interpolation mode never comes into play)."""
    opts = dict(options or {})
    for k in ("norm", "min_step", "max_step", "interp"):
        opts.pop(k, None)
    h = opts.pop("step_size", None)
    perturb = bool(opts.pop("perturb", False))
    reverse = t0 > t1
    grid = S.fixed_grid(t0, t1, h, opts.pop("grid_constructor", None), y0)
    dt, times = S.fixed_eval_times(method, grid, perturb)
    n, nev = times.shape
    user = (-times if reverse else times).reshape(-1).numpy().astype(np.float32)
    ev = prog(user).reshape(n, nev, L.EV_FLOATS)
    ev[:, :, L.MAX_TFEAT + 3] = -1.0 if reverse else 1.0
    return dt.numpy().astype(np.float32), ev


def _solve_fixed(field, prog, method, y0, cond, probes, t0, t1, options, with_lp):
    dt, ev = _fixed_tables(prog, method, t0, t1, options, y0)
    step_table = np.zeros((dt.shape[0], L.STEP_STRIDE), np.float32)
    step_table[:, 0] = dt
    step_table[:, 3] = np.float32(0.5) * dt
    x, lp, _ = E.run_fixed(field, _METHODS[method], y0, step_table, ev, cond=cond, probes=probes,
                           want_lp=with_lp)
    return x, lp


# ----------------------------------------------------------------------------------------------
# population-level wrappers (`diffusion.py:1466-1848`): affine (un)normalisation only
# ----------------------------------------------------------------------------------------------
class PopulationModelDiffusion(torch.nn.Module):
    def __init__(self, model=None, sde=None, shift=None, scale=None, method="dopri5", no_sigma=False,
                 hutchinson=False, options=None):
        super().__init__()
        self.model = model
        self.sde = sde
        self.score_model = ScoreModel(model=self.model, sde=self.sde, hutchinson=hutchinson, no_sigma=no_sigma)
        self.register_buffer("shift", shift if shift is not None else torch.zeros(model.n_dimensions, dtype=torch.float32))
        self.register_buffer("scale", scale if scale is not None else torch.ones(model.n_dimensions, dtype=torch.float32))
        self.method = method
        self.options = options

    def forward(self, base_samples):
        """`diffusion.py:1556-1585` (atol = rtol = 1e-5)."""
        x = self.score_model.sample_ode_from_base(base_samples, method=self.method, atol=1e-5, rtol=1e-5,
                                                  options=self.options)[0]
        return x * self.scale + self.shift

    def sample_sde(self, shape, steps=100, **kw):
        """`diffusion.py:1587-1608`: the reference ignores ``steps`` and always uses 100 (quirk Q6)."""
        return self.score_model.sample_sde(shape, steps=100, **kw) * self.scale + self.shift

    def log_prob(self, x, atol=1e-5, rtol=1e-5, **kw):
        """`diffusion.py:1610-1640`: no -sum(log scale) term (quirk Q7), ``method`` not forwarded (Q8)."""
        sm = self.score_model
        xT, lp = sm.solve_odes_forward((x - self.shift) / self.scale, atol=atol, rtol=rtol, options=self.options, **kw)
        sig = copy.deepcopy(self.sde).cpu()._prior_sigma()
        return E.gaussian_logprob(xT, lp.reshape(-1), sig).view(-1, 1)


class PopulationModelDiffusionConditional(torch.nn.Module):
    def __init__(self, model=None, sde=None, shift=None, scale=None, conditional_shift=None,
                 conditional_scale=None, no_sigma=False, method="dopri5", options=None):
        super().__init__()
        self.model = model
        self.sde = sde
        self.score_model = ScoreModel(model=self.model, sde=self.sde, no_sigma=no_sigma)
        f32 = torch.float32
        self.register_buffer("shift", shift if shift is not None else torch.zeros(model.n_dimensions, dtype=f32))
        self.register_buffer("scale", scale if scale is not None else torch.ones(model.n_dimensions, dtype=f32))
        self.register_buffer("conditional_shift", conditional_shift if conditional_shift is not None
                             else torch.zeros(model.n_conditionals, dtype=f32))
        self.register_buffer("conditional_scale", conditional_scale if conditional_scale is not None
                             else torch.ones(model.n_conditionals, dtype=f32))
        self.options = options
        self.method = method

    def _norm_cond(self, conditional):
        return (conditional - self.conditional_shift) / self.conditional_scale

    def forward(self, base_samples, conditional=None):
        """`diffusion.py:1754-1784`."""
        x = self.score_model.sample_ode_from_base(base_samples, conditional=self._norm_cond(conditional),
                                                  method=self.method, atol=1e-5, rtol=1e-5, options=self.options)[0]
        return x * self.scale + self.shift

    def sample_sde(self, shape, conditional=None, steps=100, **kw):
        """`diffusion.py:1786-1814` (``steps`` ignored, quirk Q6)."""
        return self.score_model.sample_sde(shape, conditional=self._norm_cond(conditional), steps=100, **kw) \
            * self.scale + self.shift

    def log_prob(self, x, conditional=None, atol=1e-5, rtol=1e-5, **kw):
        """`diffusion.py:1816-1848`."""
        xT, lp = self.score_model.solve_odes_forward((x - self.shift) / self.scale,
                                                     conditional=self._norm_cond(conditional), atol=atol, rtol=rtol,
                                                     options=self.options, **kw)
        sig = copy.deepcopy(self.sde).cpu()._prior_sigma()
        return E.gaussian_logprob(xT, lp.reshape(-1), sig).view(-1, 1)


def denoising_score_matching(score_model, x, conditional=None, **draws):
    """`diffusion.py:1369-1414` on the fused training kernels (csrc/ffb_train.cu); ``z=`` / ``t=`` replay the draws."""
    from . import training
    return training.denoising_score_matching(score_model, x, conditional, **draws)


def log_prob_score_matching(score_model, x, conditional=None, **draws):
    """`diffusion.py:1417-1463` (likelihood weighting) on the fused training kernels."""
    from . import training
    return training.log_prob_score_matching(score_model, x, conditional, **draws)
