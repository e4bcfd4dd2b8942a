"""TEST-ONLY torch-CPU model of what the CUDA kernels compute.

It implements the backend interface of ``flowfusion_b200.engine`` (``CudaBackend``,
``run_fixed``, ``gaussian_logprob``, ``PackedNet``) in plain FP32 PyTorch so that the HOST logic
of the package -- the dopri5 controller, the per-evaluation scalar programs, the fixed-grid time
tables, the class-level glue, the batch sharding and its all-reduce -- can be exercised on a
machine without a GPU and compared with the CPU oracle.  It is never imported by the package.
"""
from __future__ import annotations

import contextlib
import ctypes as C

import numpy as np
import torch

from flowfusion_b200 import _lib as L
from flowfusion_b200 import engine as E

silu = torch.nn.functional.silu


_ACTS = {0: silu, 1: torch.tanh, 2: torch.relu, 3: torch.nn.functional.softplus, 4: torch.nn.functional.gelu}


class FakePackedNet:
    def __init__(self, linears, x_col, x_dim, c_col, c_dim, t_col, t_dim, device, activation=0):
        self.act = _ACTS[activation]
        self.lin = [(l.weight.detach().float().cpu().clone(), l.bias.detach().float().cpu().clone()) for l in linears]
        self.x_col, self.x_dim, self.c_col, self.c_dim, self.t_col, self.t_dim = x_col, x_dim, c_col, c_dim, t_col, t_dim
        self.handle = C.c_void_p(0)
        self.dims = [linears[0].in_features] + [l.out_features for l in linears]
        self.flops = sum(2 * a * b for a, b in zip(self.dims[:-1], self.dims[1:]))

    def __call__(self, x, cond, tfeat):
        B = x.shape[0]
        h = torch.zeros(B, self.dims[0])
        h[:, self.x_col:self.x_col + self.x_dim] = x
        if self.c_dim:
            h[:, self.c_col:self.c_col + self.c_dim] = cond
        h[:, self.t_col:self.t_col + self.t_dim] = torch.as_tensor(tfeat[: self.t_dim])[None, :]
        for i, (w, b) in enumerate(self.lin):
            h = torch.nn.functional.linear(h, w, b)
            if i < len(self.lin) - 1:
                h = self.act(h)
        return h


def field_eval(field, ev_row, y, cond, probes):
    """(f, dlp) for the whole batch; mirrors ffb_engine.cuh::eval_field."""
    tfeat, a, c, sigma, sign = ev_row[: L.MAX_TFEAT], ev_row[L.MAX_TFEAT], ev_row[L.MAX_TFEAT + 1], \
        ev_row[L.MAX_TFEAT + 2], ev_row[L.MAX_TFEAT + 3]
    a, c, sigma, sign = (torch.tensor(v, dtype=torch.float32) for v in (a, c, sigma, sign))

    def fwd(yy):
        out = torch.zeros_like(yy)
        for i, net in enumerate(field.nets):
            xin = yy[:, field.in_off[i]: field.in_off[i] + net.x_dim]
            o = net(xin, cond, tfeat)
            if field.kind == L.FIELD_SCORE:
                s = o / sigma if field.use_sigma else o
                lin = a * yy[:, field.out_off[i]: field.out_off[i] + o.shape[1]] if field.has_drift else 0.0
                o = lin - c * s
            out[:, field.out_off[i]: field.out_off[i] + o.shape[1]] = o * (sign * field.out_sign[i])
        return out

    if field.div_mode == L.DIV_NONE:
        with torch.no_grad():
            return fwd(y), None
    with torch.enable_grad():
        yy = y.detach().clone().requires_grad_(True)
        f = fwd(yy)
        if field.div_mode == L.DIV_HUTCH:
            div = (torch.autograd.grad(f, yy, probes, retain_graph=False)[0] * probes).sum(1)
        else:
            div = torch.zeros(y.shape[0])
            for i in range(y.shape[1]):
                div = div + torch.autograd.grad(f[:, i].sum(), yy, retain_graph=True)[0][:, i]
    return f.detach(), div.detach()


class FakeBackend:
    def __init__(self, field, y0, cond=None, probes=None, with_lp=False, cond_in_state=False, cond_state=None):
        self.field, self.with_lp, self.cond_in_state = field, with_lp, cond_in_state
        self.y = y0.detach().float().clone()
        self.B, self.D = self.y.shape
        self.cond = None if cond is None else cond.detach().float()
        self.cond_state = None if cond_state is None else cond_state.detach().float()
        self.probes = None if probes is None else probes.detach().float()
        self.lp = torch.zeros(self.B) if with_lp else None
        self.f = self.dlp = None

    def _fe(self, ev_row, y):
        return field_eval(self.field, ev_row, y, self.cond, self.probes)

    def single_eval(self, ev_row):
        return self._fe(ev_row, self.y)

    def global_counts(self, group):
        n = torch.tensor([self.B], dtype=torch.int64)
        if group is not None:
            torch.distributed.all_reduce(n, group=group)
        Bt = int(n)
        out = {"x": Bt * self.D}
        if self.with_lp:
            out["lp"] = Bt
        if self.cond_in_state and self.cond is not None:
            out["cond"] = Bt * self.cond.shape[1]
        return out

    @staticmethod
    def _ss(x):
        return float((x.double() ** 2).sum())

    def eval0(self, ev_row, atol, rtol):
        atol, rtol = torch.tensor(float(atol)), torch.tensor(float(rtol))
        self.f, self.dlp = self._fe(ev_row, self.y)
        s = torch.zeros(L.NPART, dtype=torch.float64)
        sc = atol + self.y.abs() * rtol
        s[L.P_X_Y], s[L.P_X_F] = self._ss(self.y / sc), self._ss(self.f / sc)
        if self.with_lp:
            s[L.P_LP_F] = self._ss(self.dlp / atol)
        if self.cond_in_state and self.cond is not None:
            cs = self.cond_state if self.cond_state is not None else self.cond
            s[L.P_C_Y] = self._ss(cs / (atol + cs.abs() * rtol))
        return s

    def eval1(self, h0, ev_row, atol, rtol):
        atol, rtol, h0 = torch.tensor(float(atol)), torch.tensor(float(rtol)), torch.tensor(float(h0))
        f1, dlp1 = self._fe(ev_row, self.y + h0 * self.f)
        s = torch.zeros(L.NPART, dtype=torch.float64)
        sc = atol + self.y.abs() * rtol
        s[L.P_X_DF] = self._ss((f1 - self.f) / sc)
        if self.with_lp:
            s[L.P_LP_DF] = self._ss((dlp1 - self.dlp) / atol)
        return s

    def attempt(self, ev, cb, ce, cm, dt32, atol, rtol, final, x_interp):
        atol, rtol = torch.tensor(float(atol)), torch.tensor(float(rtol))
        dt = torch.tensor(float(dt32))
        k = [self.f]
        kl = [self.dlp]
        cb, ce, cm = torch.from_numpy(np.asarray(cb, np.float32)), torch.from_numpy(np.asarray(ce, np.float32)), \
            torch.from_numpy(np.asarray(cm, np.float32))
        yi = self.y
        for i in range(6):
            yi = self.y + sum(k[j] * cb[i, j] for j in range(i + 1))
            f, d = self._fe(ev[i], yi)
            k.append(f); kl.append(d)
        y1, f1 = yi, k[6]
        err = sum(k[j] * ce[j] for j in range(7))
        tol = atol + rtol * torch.max(self.y.abs(), y1.abs())
        s = torch.zeros(L.NPART, dtype=torch.float64)
        s[L.P_X_ERR] = self._ss(err / tol)
        s[L.P_NONFINITE] = float((~torch.isfinite(self.y)).sum())
        self._next = (y1, f1)
        if final:
            ymid = self.y + sum(k[j] * cm[j] for j in range(7))
            self.y_out = _dense(self.y, y1, ymid, k[0], k[6], dt, x_interp)
        if self.with_lp:
            l1 = self.lp + sum(kl[j] * cb[5, j] for j in range(6))
            errl = sum(kl[j] * ce[j] for j in range(7))
            toll = atol + rtol * torch.max(self.lp.abs(), l1.abs())
            s[L.P_LP_ERR] = self._ss(errl / toll)
            self._next_lp = (l1, kl[6])
            if final:
                lmid = self.lp + sum(kl[j] * cm[j] for j in range(7))
                self.lp_out = _dense(self.lp, l1, lmid, kl[0], kl[6], dt, x_interp)
        return s

    def attempt_rk(self, tab, ev, dt32, atol, rtol, final, x_interp):
        """Model of CudaBackend.attempt_rk: any adaptive tableau, stage by stage (torchdiffeq rk_common._runge_kutta_step)."""
        atol, rtol = torch.tensor(float(atol)), torch.tensor(float(rtol))
        dt = torch.tensor(float(dt32))
        n_k = len(tab.alpha) + 1
        T = lambda v: torch.from_numpy(np.asarray(v, np.float32))        # noqa: E731
        k, kl = [self.f], [self.dlp]
        yi = self.y
        for i in range(n_k - 1):
            b = T(tab.beta[i]) * dt
            yi = self.y + sum(k[j] * b[j] for j in range(i + 1))
            f, d = self._fe(ev[i], yi)
            k.append(f); kl.append(d)
        cs, ce, cm = T(tab.c_sol) * dt, T(tab.c_err) * dt, T(tab.c_mid) * dt
        y1 = yi if tab.fsal else self.y + sum(k[j] * cs[j] for j in range(n_k))
        f1 = k[-1]
        err = sum(k[j] * ce[j] for j in range(n_k))
        tol = atol + rtol * torch.max(self.y.abs(), y1.abs())
        s = torch.zeros(L.NPART, dtype=torch.float64)
        s[L.P_X_ERR] = self._ss(err / tol)
        s[L.P_NONFINITE] = float((~torch.isfinite(self.y)).sum())
        self._next = (y1, f1)
        if final:
            ymid = self.y + sum(k[j] * cm[j] for j in range(n_k))
            self.y_out = _dense(self.y, y1, ymid, k[0], f1, dt, x_interp)
        if self.with_lp:
            l1 = self.lp + sum(kl[j] * cs[j] for j in range(min(n_k, 6)))
            errl = sum(kl[j] * ce[j] for j in range(n_k))
            toll = atol + rtol * torch.max(self.lp.abs(), l1.abs())
            s[L.P_LP_ERR] = self._ss(errl / toll)
            self._next_lp = (l1, kl[-1])
            if final:
                lmid = self.lp + sum(kl[j] * cm[j] for j in range(n_k))
                self.lp_out = _dense(self.lp, l1, lmid, kl[0], kl[-1], dt, x_interp)
        return s

    def refresh_f(self, ev_row):
        self.f, self.dlp = self._fe(ev_row, self.y)

    def accept(self):
        self.y, self.f = self._next
        if self.with_lp:
            self.lp, self.dlp = self._next_lp

    def output(self):
        return self.y_out, (self.lp_out if self.with_lp else None)

    # -- device-side controller: the CPU twin of the control kernel (ffb_dopri5_control_host) drives this model ----
    def ctl_supported(self):
        return True                    # a property of the field, not of this rank's shard (engine.CudaBackend)

    def ctl_attempt_ms_estimate(self, rows=None):
        return 0.0

    def ctl_begin(self, params, t, dt_next, grid_idx, atol, rtol):
        self._lib = L.load()
        self._p, self._c = params, L.Ctl()
        self._c.t, self._c.dt_next, self._c.grid_idx = float(t), float(dt_next), int(grid_idx)
        self._tol = (atol, rtol)
        self._done_at = []
        self._sums = torch.zeros(L.NPART, dtype=torch.float64)
        L.check(self._lib.ffb_dopri5_control_host(C.byref(self._p), None, C.byref(self._c), 0), "control_host")

    def ctl_attempt(self, reduce=True):
        c = self._c
        if c.done == L.CTL_RUNNING:           # the CUDA attempt kernel returns at once otherwise
            ev = np.frombuffer(bytes(c.ev), np.float32).reshape(6, L.EV_FLOATS).copy()
            cb = np.frombuffer(bytes(c.cb), np.float32).reshape(6, 6).copy()
            ce, cm = np.frombuffer(bytes(c.ce), np.float32).copy(), np.frombuffer(bytes(c.cm), np.float32).copy()
            self._sums = self.attempt(ev, cb, ce, cm, np.float32(c.dt), self._tol[0], self._tol[1], bool(c.final),
                                      np.float32(c.x_interp))
        return self._sums

    def ctl_control(self, reduce=False):
        cur = self._c.cur
        sums = np.ascontiguousarray(self._sums.numpy(), np.float64)
        L.check(self._lib.ffb_dopri5_control_host(C.byref(self._p), sums.ctypes.data, C.byref(self._c), 1), "control_host")
        if self._c.cur != cur:
            self.accept()
        self._done_at.append(int(self._c.done))
        assert self._c.n_turns == len(self._done_at)

    def ctl_wait_turn(self, k):
        return int(self._done_at[k - 1])

    def ctl_finish(self):
        return self._c


class FakeStagedBackend(FakeBackend):
    """Model of engine.StagedBackend: the field from the torch model, the network Jacobian by autograd, and the
    Hutch++ / XTrace algebra from the C twin of the estimator kernel (ffb_trace_estimate_host)."""

    def __init__(self, field, y0, estimator, cond=None):
        super().__init__(field, y0, cond=cond, probes=None, with_lp=True)
        self.est = estimator

    def ctl_supported(self):
        return False

    def _fe(self, ev_row, y):
        field, net = self.field, self.field.nets[0]
        f, _ = field_eval(field.with_div(L.DIV_NONE) if hasattr(field, "with_div") else field, ev_row, y, self.cond, None)
        tfeat = ev_row[: L.MAX_TFEAT]

        with torch.enable_grad():
            yy = y.detach().clone().requires_grad_(True)
            o = net(yy, self.cond, tfeat)
            J = torch.stack([torch.autograd.grad(o[:, n].sum(), yy, retain_graph=True)[0] for n in range(o.shape[1])], dim=1)
        J = J.detach()                                                            # [b][n][j]
        jac = np.ascontiguousarray(J.permute(0, 2, 1).numpy(), np.float32)        # [b][j][n] = d net_n / d x_j
        S = np.ascontiguousarray(self.est.S.detach().float().numpy())
        G = None if self.est.G is None else np.ascontiguousarray(self.est.G.detach().float().numpy())
        out = np.zeros(self.B, np.float32)
        a = L.TraceArgs()
        a.batch, a.dim, a.kind, a.rank, a.nvec = self.B, self.D, self.est.kind, self.est.rank, self.est.nvec
        a.jac, a.S, a.G = jac.ctypes.data, S.ctypes.data, None if G is None else G.ctypes.data
        a.score, a.use_sigma, a.has_drift = int(field.kind == L.FIELD_SCORE), int(field.use_sigma), int(field.has_drift)
        a.a, a.c, a.sigma, a.sign = (float(ev_row[L.MAX_TFEAT + i]) for i in range(4))
        a.dlp = out.ctypes.data
        L.check(L.load().ffb_trace_estimate_host(C.byref(a)), "ffb_trace_estimate_host")
        return f, torch.from_numpy(out)

    def single_eval(self, ev_row):
        return self._fe(ev_row, self.y)

    def state(self):
        return self.y, self.lp

    def feval(self, y, ev_row):
        return self._fe(ev_row, y)

    def combine(self, y0, ks, coefs):
        acc = ks[0] * torch.tensor(float(coefs[0]))
        for k, c in zip(ks[1:], coefs[1:]):
            acc = acc + k * torch.tensor(float(c))
        return y0 + acc


def _dense(y0, y1, ymid, f0, f1, dt, x):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * ymid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * ymid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * ymid
    d = dt * f0
    x = torch.tensor(float(x))
    total = y0 + x * d
    xp = x
    for coeff in (c, b, a):
        xp = xp * x
        total = total + xp * coeff
    return total


def fake_run_fixed(field, method, x0, step_table, ev_table, cond=None, probes=None, lp0=None, noise=None,
                   philox=None, row_offset=0, want_lp=False):
    x = x0.detach().float().clone()
    lp = torch.zeros(x.shape[0]) if want_lp else None
    third = torch.tensor(1 / 3, dtype=torch.float32)
    st = torch.from_numpy(np.asarray(step_table, np.float32))
    x_mean = x
    status = torch.tensor([0, 2 ** 31 - 1], dtype=torch.int32)      # like the kernel: flags, first NaN step of EM
    D = field.state_dim

    def F(ev_row, y):
        return field_eval(field, ev_row, y, cond, probes)

    def call(c, ev_row, y):     # one network of a 2-call field
        sub = E.FieldSpec([field.nets[c]], field.state_dim, field.cond_dim, field.kind, field.use_sigma,
                          field.has_drift, L.DIV_NONE, (field.in_off[c],) * 2, (field.out_off[c],) * 2,
                          (field.out_sign[c],) * 2)
        return field_eval(sub, ev_row, y, cond, None)[0]

    kick = None
    for n in range(st.shape[0]):
        dt = st[n, 0]
        ev = ev_table[n]
        if method == L.M_EULER:
            k1, d1 = F(ev[0], x)
            x = x + dt * k1
            if want_lp:
                lp = lp + dt * d1
        elif method == L.M_MIDPOINT:
            k1, d1 = F(ev[0], x)
            k2, d2 = F(ev[1], x + k1 * st[n, 3])
            x = x + dt * k2
            if want_lp:
                lp = lp + dt * d2
        elif method == L.M_RK4:
            k1, d1 = F(ev[0], x)
            k2, d2 = F(ev[1], x + dt * k1 * third)
            k3, d3 = F(ev[2], x + dt * (k2 - k1 * third))
            k4, d4 = F(ev[3], x + dt * (k1 - k2 + k3))
            x = x + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
            if want_lp:
                lp = lp + (d1 + 3 * (d2 + d3) + d4) * dt * 0.125
        elif method == L.M_EM:
            f, _ = F(ev[0], x)
            x_mean = x + f * dt
            x = x_mean + st[n, 1] * (noise[n] * st[n, 2])
            if torch.isnan(x).any():          # the kernel keeps integrating and records the first such step
                status[0] |= L.ST_NAN_SAMPLE
                status[1] = min(int(status[1]), n)
        elif method == L.M_LEAPFROG:
            half = st[n, 3]
            if kick is None:
                kick = call(1, ev[0], x)
            x = x + half * kick
            x = x + dt * call(0, ev[1], x)
            kick = call(1, ev[2], x)
            x = x + half * kick
    return (x_mean if method == L.M_EM else x), lp, status


def fake_gaussian_logprob(x, add, sigma=1.0):
    out = torch.distributions.Normal(0.0, float(sigma)).log_prob(x.float()).sum(dim=1)
    return out if add is None else out + add


@contextlib.contextmanager
def patched_engine():
    """Route the package's device plumbing to the torch-CPU model (tests only)."""
    saved = {k: getattr(E, k) for k in ("CudaBackend", "StagedBackend", "run_fixed", "gaussian_logprob", "PackedNet", "require_cuda",
                                        "require_cuda_device")}
    E.CudaBackend = FakeBackend
    E.StagedBackend = FakeStagedBackend
    E.run_fixed = fake_run_fixed
    E.gaussian_logprob = fake_gaussian_logprob
    E.PackedNet = FakePackedNet
    E.require_cuda = lambda t, what: None
    E.require_cuda_device = lambda dev: None
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(E, k, v)
