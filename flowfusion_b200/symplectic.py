"""Drop-in surface of ``flowfusion.symplectic`` (volume-preserving Hamiltonian flow).

``SymplecticMLP`` holds two networks that output dq/dt = mlp_q([p, c, temb]) and
dp/dt = -mlp_p([q, c, temb]) (`symplectic.py:80-123`); the field is divergence-free by
construction, so ``log_prob`` needs no trace.  In the kernels this is a two-network field
acting on the (q | p) column blocks of a 2D-wide state.

* ``SymplecticFlowModel.sample``    `symplectic.py:165-201`  forward Euler on ``linspace(1, 0)``
  (reference behaviour, ``method="euler"``) or the kick-drift-kick **leapfrog** extension
  (``method="leapfrog"``; the reference has no leapfrog, SURVEY H5 -- validated by reversibility,
  volume preservation and 2nd-order convergence instead of an oracle).
* ``SymplecticFlowModel.log_prob``  `symplectic.py:203-254`  dopri5 on the plain (B, 2D) state.
* ``HamiltonianMLP`` (extension, BASELINE.json north_star "the leapfrog integrator's dH/dq, dH/dp is a fused
  forward+backward MLP kernel"): a scalar Hamiltonian network whose leapfrog runs in ONE launch, the gradient being the
  fused forward + backward sweep of csrc/ffb_train.cu.  The reference has no such model (its networks output the
  derivatives); the oracle is autograd on the CPU (oracle/port.py: hamiltonian_leapfrog).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import dist as _dist
from . import engine as E
from . import solver as S


class SymplecticMLP(nn.Module):
    def __init__(self, n_data_dims, n_conditionals, embedding_dimensions, units, activation=nn.SiLU()):
        super().__init__()
        in_dim = n_data_dims + n_conditionals + embedding_dimensions
        self.mlp_q_dynamics = self._create_mlp(in_dim, n_data_dims, units, activation)
        self.mlp_p_dynamics = self._create_mlp(in_dim, n_data_dims, units, activation)
        self.register_buffer("W", torch.randn(embedding_dimensions // 2) * 16.0)
        self._dims = (n_data_dims, n_conditionals, embedding_dimensions)
        self._packed = None

    def _create_mlp(self, input_dim, output_dim, units, activation):
        layers, cur = [], input_dim
        for u in units:
            layers += [nn.Linear(cur, u), activation]
            cur = u
        layers.append(nn.Linear(cur, output_dim))
        return nn.Sequential(*layers)

    # -- kernel plumbing --------------------------------------------------------------------
    def _nets(self):
        D, Cn, emb = self._dims
        lq = [m for m in self.mlp_q_dynamics if isinstance(m, nn.Linear)]
        lp = [m for m in self.mlp_p_dynamics if isinstance(m, nn.Linear)]
        act = E.activation_of(list(self.mlp_q_dynamics) + list(self.mlp_p_dynamics))
        key = E.weights_fingerprint(lq + lp, act)
        if self._packed is None or self._packed[0] != key:
            dev = lq[0].weight.device
            E.require_cuda_device(dev)
            mk = lambda lin: E.PackedNet(lin, x_col=0, x_dim=D, c_col=D, c_dim=Cn, t_col=D + Cn, t_dim=emb, device=dev,  # noqa: E731
                                         activation=act)
            self._packed = (key, (mk(lq), mk(lp)))
        return self._packed[1]

    def _field(self):
        D, Cn, _ = self._dims
        nq, npn = self._nets()
        # call 0: dq/dt = +mlp_q(p) ; call 1: dp/dt = -mlp_p(q)
        return E.FieldSpec([nq, npn], 2 * D, Cn, kind=L.FIELD_NET, in_off=(D, 0), out_off=(0, D), out_sign=(1.0, -1.0))

    def _program(self):
        W = self.W.detach().cpu().float()
        emb = self._dims[2]

        def program(times32: np.ndarray) -> np.ndarray:
            t = torch.from_numpy(np.ascontiguousarray(times32, np.float32))
            proj = t[:, None] * W[None, :] * 2 * math.pi                   # `symplectic.py:103`
            rows = np.zeros((t.shape[0], L.EV_FLOATS), np.float32)
            rows[:, :emb] = torch.cat([torch.sin(proj), torch.cos(proj)], dim=1).numpy()
            rows[:, L.MAX_TFEAT + 3] = 1.0
            return rows

        from .diffusion import fourier_program_spec
        program.spec = fourier_program_spec(W, math.pi)      # device-side dopri5 controller
        return program

    def forward(self, t, state, conditional):
        """`symplectic.py:80-123` at a batch-uniform time -> cat[dq/dt, dp/dt]."""
        E.require_cuda(state, "state")
        tt = torch.as_tensor(t, dtype=torch.float32).detach()
        if tt.dim() > 0:
            if not bool((tt == tt.reshape(-1)[0]).all()):
                # one time per sample (`symplectic.py:101-114`): the two networks on explicit input rows
                from . import training
                D = self._dims[0]
                tt = tt.to(state.device).reshape(-1)
                q, p = state[:, :D], state[:, D:]
                proj = tt[:, None] * self.W[None, :] * 2 * math.pi
                temb = torch.cat([torch.sin(proj), torch.cos(proj)], dim=1)
                mid = [conditional] if conditional is not None else []
                lq = [m for m in self.mlp_q_dynamics if isinstance(m, nn.Linear)]
                lp = [m for m in self.mlp_p_dynamics if isinstance(m, nn.Linear)]
                act = E.activation_of(list(self.mlp_q_dynamics) + list(self.mlp_p_dynamics))
                v_q = training.mlp_forward(lq, act, torch.cat([p] + mid + [temb], dim=1))
                v_p = -training.mlp_forward(lp, act, torch.cat([q] + mid + [temb], dim=1))
                return torch.cat([v_q, v_p], dim=-1)
            tt = tt.reshape(-1)[0]
        row = self._program()(tt.cpu().reshape(1).numpy())[0]
        return E.CudaBackend(self._field(), state, cond=conditional).single_eval(row)[0]


class SymplecticFlowModel(nn.Module):
    def __init__(self, model, shift, scale, conditional_shift, conditional_scale):
        super().__init__()
        self.model = model
        self.register_buffer("shift", shift)
        self.register_buffer("scale", scale)
        self.register_buffer("conditional_shift", conditional_shift)
        self.register_buffer("conditional_scale", conditional_scale)
        self.process_group = None
        self.last_stats = None

    def _norm_cond(self, conditional):
        if conditional is None:
            return None
        return (conditional - self.conditional_shift) / self.conditional_scale

    @torch.no_grad()
    def sample(self, shape, conditional=None, num_steps=1, *, z0=None, method="euler"):
        """`symplectic.py:165-201`.  ``z0`` (B, 2D) replaces the internal draw `:186`."""
        dev = next(self.model.parameters()).device
        B, D = shape[0], shape[1]
        x = torch.randn(B, 2 * D, device=dev) if z0 is None else z0
        E.require_cuda(x, "z0")
        cond = self._norm_cond(conditional)
        ts = torch.linspace(1.0, 0.0, num_steps + 1)                        # `:191`
        dt = ts[1:] - ts[:-1]                                               # `:195`
        prog = self.model._program()
        step_table = np.zeros((num_steps, L.STEP_STRIDE), np.float32)
        step_table[:, 0] = dt.numpy()
        if method == "euler":
            ev = prog(ts[:-1].numpy()).reshape(num_steps, 1, L.EV_FLOATS)
            meth = L.M_EULER
        elif method == "leapfrog":
            half = 0.5 * dt
            step_table[:, 3] = half.numpy()
            times = torch.stack([ts[:-1], ts[:-1] + half, ts[1:]], dim=1).reshape(-1)
            ev = prog(times.numpy()).reshape(num_steps, 3, L.EV_FLOATS)
            meth = L.M_LEAPFROG
        else:
            raise NotImplementedError(f"sample method {method!r} (euler, leapfrog)")
        out, _, _ = E.run_fixed(self.model._field(), meth, x, step_table, ev, cond=cond)
        q0 = out[:, :D]
        return q0 * self.scale + self.shift

    @torch.no_grad()
    def log_prob(self, x, conditional=None, atol=1e-5, rtol=1e-5, *, p0=None):
        """`symplectic.py:203-254` -> (B,).  ``p0`` replaces the random momenta drawn at `:228`."""
        E.require_cuda(x, "x")
        q0 = (x - self.shift) / self.scale
        cond = self._norm_cond(conditional)
        p0 = torch.randn_like(q0) if p0 is None else p0
        init = torch.cat([q0, p0], dim=-1)
        be = E.CudaBackend(self.model._field(), init, cond=cond)
        group = self.process_group if self.process_group is not None else _dist.current_group()
        self.last_stats = S.dopri5(be, self.model._program(), 0.0, 1.0, rtol, atol, None, group=group)
        z1, _ = be.output()
        log_p_z1 = E.gaussian_logprob(z1, None, 1.0)
        log_p_p0 = E.gaussian_logprob(p0, None, 1.0)
        return log_p_z1 - log_p_p0 - torch.sum(torch.log(self.scale))


class HamiltonianMLP(nn.Module):
    """Scalar Hamiltonian ``H(q, p[, c]) = MLP(cat[q, p, c])`` and its kick-drift-kick leapfrog flow on the fused
    forward+backward kernel (``ffb_hamiltonian_leapfrog``).  Extension beyond the reference (see the module docstring)."""

    def __init__(self, n_data_dims, n_conditionals=0, units=(128, 128), activation=nn.SiLU()):
        super().__init__()
        self.n_data_dims, self.n_conditionals = n_data_dims, n_conditionals
        layers, cur = [], 2 * n_data_dims + n_conditionals
        for u in units:
            layers += [nn.Linear(cur, u), activation]
            cur = u
        layers.append(nn.Linear(cur, 1))
        self.net = nn.Sequential(*layers)

    @torch.no_grad()
    def leapfrog(self, z0, conditional=None, num_steps=100, dt=0.01, return_energy=False):
        """z0 (B, 2D) = [q | p] -> z after ``num_steps`` steps of p -= dt/2 dH/dq; q += dt dH/dp; p -= dt/2 dH/dq.
        ``return_energy``: also (B, 2) = H at the start and at the end."""
        import ctypes as C
        E.require_cuda(z0, "z0")
        lib, dev, D, Cn = L.load(), z0.device, self.n_data_dims, self.n_conditionals
        lin = [m for m in self.net if isinstance(m, nn.Linear)]
        if z0.shape[1] != 2 * D or (Cn > 0) != (conditional is not None):
            raise ValueError("leapfrog: z0 must be (B, 2 * n_data_dims) and the conditional given iff n_conditionals > 0")
        d = L.NetDesc()
        d.n_layers, d.in_features = len(lin), lin[0].in_features
        keep = []
        for i, m in enumerate(lin):
            w, b = E._dev_f32(m.weight, dev), E._dev_f32(m.bias, dev)
            keep += [w, b]
            d.widths[i], d.weight[i], d.bias[i] = m.out_features, w.data_ptr(), b.data_ptr()
        d.x_dim, d.activation = lin[0].in_features, E.activation_of(list(self.net))
        a = L.HamiltonianArgs()
        z0 = E._dev_f32(z0, dev)
        cond = E._dev_f32(conditional, dev) if conditional is not None else None
        out = torch.empty_like(z0)
        h = torch.empty(z0.shape[0], 2, device=dev) if return_energy else None
        nbytes = int(lib.ffb_train_work_bytes(C.byref(d), 0, 1))
        if nbytes == 0:
            raise L.FFBError("ffb_train_work_bytes: " + lib.ffb_last_error().decode())
        work = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
        a.batch, a.dim, a.cond_dim = z0.shape[0], D, Cn
        a.z0, a.cond, a.z_out = z0.data_ptr(), (cond.data_ptr() if cond is not None else None), out.data_ptr()
        a.h_out, a.n_steps, a.dt, a.work = (h.data_ptr() if h is not None else None), int(num_steps), float(dt), work.data_ptr()
        with E.on_device(dev):
            L.check(lib.ffb_hamiltonian_leapfrog(C.byref(d), C.byref(a), E._stream(dev)), "ffb_hamiltonian_leapfrog")
        return (out, h) if return_energy else out

    def energy(self, z, conditional=None):
        """H(q, p[, c]) for every row, (B,)."""
        return self.leapfrog(z, conditional, num_steps=0, return_energy=True)[1][:, 0]
