"""Device-side plumbing between the model objects and ``libffb200.so``.

``PackedNet``    packs the ``nn.Linear`` weights of one MLP into the kernels' image (once).
``FieldSpec``    describes a vector field (1-2 networks + the reference's drift formula).
``CudaBackend``  owns the device buffers of one adaptive solve and launches the kernels.
``run_fixed``    launches the whole-trajectory fixed-grid kernel.

PyTorch is used for device memory, streams and (in ``solver.py``) the NCCL all-reduce only.
Nothing here computes on the CPU and nothing falls back: without the CUDA library or a CUDA
tensor these classes raise.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import threading
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L


class _Profiler:
    """Optional CUDA-event timing of the integrator launches (bench.py's roofline leg).
    Events are recorded on torch's current stream, which is the stream the kernels run on."""

    def __init__(self):
        self.enabled = False
        self.records = []          # (name, start_event, end_event, rows)

    def reset(self, enabled=True):
        self.enabled, self.records = enabled, []

    def summary(self):
        """name -> (launches, total_ms, total_rows); call after torch.cuda.synchronize()."""
        out = {}
        for name, e0, e1, rows in self.records:
            n, ms, r = out.get(name, (0, 0.0, 0))
            out[name] = (n + 1, ms + e0.elapsed_time(e1), r + rows)
        return out


profiler = _Profiler()


class _timed:
    def __init__(self, name, rows):
        self.name, self.rows = name, rows

    def __enter__(self):
        if profiler.enabled:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if profiler.enabled:
            self.e1.record()
            profiler.records.append((self.name, self.e0, self.e1, self.rows))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device=None):
    """The current torch stream of ``device`` (default: the current device) as a ``cudaStream_t``."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def on_device(device):
    """Context manager: make ``device`` the current CUDA device.  The C ABI launches on the current device with raw
    pointers, so every entry point runs under the device of its state tensor (a model on cuda:1 works while
    cuda:0 is current, as it does in the reference)."""
    device = torch.device(device)
    if device.type != "cuda":          # the torch-CPU model of the kernels (tests/kernel_model.py): nothing to select
        return contextlib.nullcontext()
    return torch.cuda.device(device)


def _dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def require_cuda(t: torch.Tensor, what: str):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise L.FFBError(f"{what} must be a CUDA tensor: flowfusion_b200 has no CPU path "
                         "(move the model and its inputs to a B200 with .to('cuda'))")


def require_cuda_device(dev):
    if torch.device(dev).type != "cuda":
        raise L.FFBError("the model must be on a CUDA device (flowfusion_b200 has no CPU path)")


def activation_code(act) -> int:
    """FFB_ACT_* of an activation module (instance or class), or NotImplementedError."""
    nn = torch.nn
    m = act() if isinstance(act, type) else act
    if isinstance(m, nn.SiLU):
        return L.ACT_SILU
    if isinstance(m, nn.Tanh):
        return L.ACT_TANH
    if isinstance(m, nn.ReLU):
        return L.ACT_RELU
    if isinstance(m, nn.Softplus) and m.beta == 1 and m.threshold == 20:
        return L.ACT_SOFTPLUS
    if isinstance(m, nn.GELU) and getattr(m, "approximate", "none") == "none":
        return L.ACT_GELU
    raise NotImplementedError(f"activation {type(m).__name__} is not implemented in the CUDA kernels "
                              "(SiLU, Tanh, ReLU, Softplus(beta=1, threshold=20), GELU(erf) are)")


def activation_of(modules) -> int:
    """The one activation code used between the Linear layers of ``modules`` (default SiLU when there is none)."""
    codes = {activation_code(m) for m in modules if not isinstance(m, torch.nn.Linear)}
    if len(codes) > 1:
        raise NotImplementedError("mixed hidden-layer activations are not implemented in the CUDA kernels")
    return codes.pop() if codes else L.ACT_SILU


class PackedNet:
    """One MLP, packed for the tile engine.  ``linears``: the nn.Linear modules in order."""

    def __init__(self, linears: Sequence[torch.nn.Linear], x_col, x_dim, c_col, c_dim, t_col, t_dim, device,
                 activation: int = 0):
        lib = L.load()
        if len(linears) > L.MAX_LAYERS:
            raise NotImplementedError(f"at most {L.MAX_LAYERS} Linear layers are supported")
        if linears[-1].out_features > 128:
            raise NotImplementedError("the output layer is at most 128 wide (the ODE state has at most 128 columns)")
        d = L.NetDesc()
        d.n_layers = len(linears)
        d.in_features = linears[0].in_features
        self._keep = []
        for i, lin in enumerate(linears):
            if lin.out_features > L.MAX_WIDTH:
                raise NotImplementedError(f"layer widths above {L.MAX_WIDTH} are not supported (widths above 128 and networks "
                                          "deeper than 8 Linear layers run on the wide FP32 engine, csrc/ffb_engine_wide.cuh)")
            w, b = _dev_f32(lin.weight, device), _dev_f32(lin.bias, device)
            self._keep += [w, b]
            d.widths[i] = lin.out_features
            d.weight[i] = w.data_ptr()
            d.bias[i] = b.data_ptr()
        d.x_col, d.x_dim, d.c_col, d.c_dim, d.t_col, d.t_dim = x_col, x_dim, c_col, c_dim, t_col, t_dim
        d.activation = int(activation)
        self.activation = int(activation)
        h = C.c_void_p()
        with torch.cuda.device(device):
            L.check(lib.ffb_net_create(C.byref(d), _stream(), C.byref(h)), "ffb_net_create")
            torch.cuda.current_stream().synchronize()      # packing kernels read the torch tensors
        self._keep = []
        self.handle = h
        self.flops = int(lib.ffb_net_flops(h))
        self.dims = [linears[0].in_features] + [l.out_features for l in linears]
        self.x_dim, self.c_dim, self.t_dim = x_dim, c_dim, t_dim
        self.device = torch.device(device)
        self._lib = lib

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self._lib.ffb_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def weights_fingerprint(linears: Sequence[torch.nn.Linear], activation: int = 0):
    """Cheap change detector so trained / reloaded weights (or a swapped activation) are re-packed."""
    return (activation,) + tuple((l.weight.data_ptr(), l.weight._version, l.bias.data_ptr(), l.bias._version, str(l.weight.device))
                 for l in linears)


class FieldSpec:
    def __init__(self, nets: List[PackedNet], state_dim, cond_dim, kind=L.FIELD_NET, use_sigma=False,
                 has_drift=False, div_mode=L.DIV_NONE, in_off=(0, 0), out_off=(0, 0), out_sign=(1.0, 1.0)):
        self.nets, self.state_dim, self.cond_dim = nets, state_dim, cond_dim
        self.kind, self.use_sigma, self.has_drift, self.div_mode = kind, use_sigma, has_drift, div_mode
        self.in_off, self.out_off, self.out_sign = in_off, out_off, out_sign
        f = L.Field()
        f.n_calls = len(nets)
        for i, n in enumerate(nets):
            f.net[i] = n.handle.value
            f.in_off[i], f.out_off[i], f.out_sign[i] = in_off[i], out_off[i], out_sign[i]
        f.state_dim, f.cond_dim, f.kind = state_dim, cond_dim, kind
        f.use_sigma, f.has_drift, f.div_mode = int(use_sigma), int(has_drift), div_mode
        self.c = f
        self.device = getattr(nets[0], "device", None)
        if any(getattr(n, "device", self.device) != self.device for n in nets):
            raise L.FFBError("the networks of a field must live on one CUDA device")

    def check_device(self, t: torch.Tensor, what="state"):
        """The packed weights are raw pointers of ONE device: the state must live there too."""
        if self.device is not None and t.device != self.device:
            raise L.FFBError(f"{what} is on {t.device} but the model's weights are on {self.device}: move them to one device")

    def with_div(self, div_mode):
        return FieldSpec(self.nets, self.state_dim, self.cond_dim, self.kind, self.use_sigma, self.has_drift,
                         div_mode, self.in_off, self.out_off, self.out_sign)

    @property
    def ntan(self):
        return {L.DIV_NONE: 0, L.DIV_EXACT: self.nets[0].x_dim, L.DIV_HUTCH: 1}[self.div_mode]

    def num_tiles(self, batch):
        return int(L.load().ffb_num_tiles(C.byref(self.c), batch))

    def scratch(self, device):
        n = int(L.load().ffb_scratch_bytes(C.byref(self.c)))
        return torch.empty(n // 4, dtype=torch.float32, device=device)

    def flops_per_eval(self):
        """Algorithmic FLOPs (2*MAC over Linear layers) of one evaluation for one trajectory,
        with the forward-mode minimum for tangents."""
        total = sum(n.flops for n in self.nets)
        if self.div_mode != L.DIV_NONE:
            d = self.nets[0].dims
            hidden = sum(2 * d[i] * d[i + 1] for i in range(1, len(d) - 2))
            if self.div_mode == L.DIV_EXACT:      # layer 0 = column gather, last layer = one output column
                total += self.nets[0].x_dim * (hidden + 2 * d[-2])
            else:
                total += 2 * self.nets[0].x_dim * d[1] + hidden + 2 * d[-2] * d[-1]
        return total


def ev_rows_to_struct(dst_array, rows: np.ndarray):
    """Copy (n, EV_FLOATS) float32 rows into a ctypes array of EvalScalars."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    C.memmove(C.addressof(dst_array), rows.ctypes.data, rows.nbytes)


_ctl_tls = threading.local()


def _ctl_pinned():
    """(notify slots, controller-block staging) in pinned host memory, one pair per host thread."""
    bufs = getattr(_ctl_tls, "bufs", None)
    if bufs is None:
        bufs = (torch.zeros(L.CTL_NOTIFY_SLOTS, dtype=torch.int32).pin_memory(),
                torch.zeros(C.sizeof(L.Ctl), dtype=torch.uint8).pin_memory())
        _ctl_tls.bufs = bufs
    return bufs


class CudaBackend:
    """Device state of ONE adaptive solve (this rank's shard of the batch)."""

    def __init__(self, field: FieldSpec, y0: torch.Tensor, cond=None, probes=None, with_lp=False,
                 cond_in_state=False, cond_state=None):
        require_cuda(y0, "state")
        field.check_device(y0)
        self.lib = L.load()
        self.field = field
        dev = y0.device
        self.dev = dev
        B, D = y0.shape
        assert D == field.state_dim
        self.B, self.D = B, D
        self.with_lp = with_lp
        self.cond_in_state = cond_in_state
        self.y = [_dev_f32(y0, dev).clone(), torch.empty(B, D, device=dev)]
        self.f = [torch.empty(B, D, device=dev), torch.empty(B, D, device=dev)]
        self.y_out = torch.empty(B, D, device=dev)
        if with_lp:
            self.lp = [torch.zeros(B, device=dev), torch.empty(B, device=dev)]
            self.dlp = [torch.empty(B, device=dev), torch.empty(B, device=dev)]
            self.lp_out = torch.empty(B, device=dev)
        else:
            self.lp = self.dlp = [None, None]
            self.lp_out = None
        self.cond = None if cond is None else _dev_f32(cond, dev)
        self.probes = None if probes is None else _dev_f32(probes, dev)
        self.cond_state = None if cond_state is None else _dev_f32(cond_state, dev)
        self.ntiles = max(field.num_tiles(B), 1)
        self.partials = torch.zeros(self.ntiles, L.NPART, dtype=torch.float64, device=dev)
        self.sums = torch.zeros(L.NPART, dtype=torch.float64, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.scratch = field.scratch(dev)
        self.cur = 0
        self.dargs = L.Dopri5Args()
        self._dargs_static = False
        self._stream = _stream(dev)        # torch.cuda.current_stream() costs ~20 us per call: once per solve
        self.eargs = L.EvalArgs()

    # -- counts for the RMS norms (global over ranks) -------------------------------------------
    def global_counts(self, group):
        if group is None:
            Bt = self.B                              # no device round trip for a single rank
        else:
            n = torch.tensor([self.B], dtype=torch.int64, device=self.dev)
            torch.distributed.all_reduce(n, group=group)
            Bt = int(n.item())
        out = {"x": Bt * self.D}
        if self.with_lp:
            out["lp"] = Bt
        if self.cond_in_state and self.cond is not None:
            out["cond"] = Bt * self.cond.shape[1]
        return out

    def _reduce(self):
        L.check(self.lib.ffb_reduce_partials(_ptr(self.partials), self.ntiles, _ptr(self.sums), self._stream),
                "ffb_reduce_partials")
        return self.sums

    def _eval(self, ev_row, atol, rtol, norms, h=None):
        a = self.eargs
        c = self.cur
        a.batch = self.B
        a.y = _ptr(self.y[c])
        a.fbase = _ptr(self.f[c]) if norms == 2 else None
        a.dlpbase = _ptr(self.dlp[c]) if (norms == 2 and self.with_lp) else None
        a.h = float(h) if h is not None else 0.0
        a.cond, a.probes, a.cond_state = _ptr(self.cond), _ptr(self.probes), _ptr(self.cond_state)
        a.f = _ptr(self.f[c]) if norms == 1 else None
        a.dlp = _ptr(self.dlp[c]) if (norms == 1 and self.with_lp) else None
        ev_rows_to_struct(a.ev, ev_row[None, :])
        a.atol, a.rtol, a.norms = float(atol), float(rtol), norms
        a.cond_in_state = int(self.cond_in_state)
        a.partials, a.status, a.scratch = _ptr(self.partials), _ptr(self.status), _ptr(self.scratch)
        if self.B:
            with _timed("field_eval", self.B):
                L.check(self.lib.ffb_field_eval(C.byref(self.field.c), C.byref(a), self._stream), "ffb_field_eval")
        return self._reduce()

    def eval0(self, ev_row, atol, rtol):
        return self._eval(ev_row, atol, rtol, 1)

    def single_eval(self, ev_row):
        """One field evaluation at the current state -> (f, dlp or None)."""
        with on_device(self.dev):
            self._eval(ev_row, 1.0, 0.0, 1)
        return self.f[self.cur], self.dlp[self.cur]

    def eval1(self, h0, ev_row, atol, rtol):
        return self._eval(ev_row, atol, rtol, 2, h=h0)

    def attempt(self, ev, cb, ce, cm, dt32, atol, rtol, final, x_interp):
        # The host sits between two launches here (the GPU is idle until the next attempt is enqueued),
        # so the argument block is filled with as few ctypes stores as possible.
        a = self.dargs
        c, n = self.cur, 1 - self.cur
        if not self._dargs_static:
            a.batch = self.B
            a.cond, a.probes = _ptr(self.cond), _ptr(self.probes)
            a.y_out, a.lp_out = _ptr(self.y_out), _ptr(self.lp_out)
            a.partials, a.status, a.scratch = _ptr(self.partials), _ptr(self.status), _ptr(self.scratch)
            self._dargs_ptrs = [(_ptr(self.y[i]), _ptr(self.f[i]), _ptr(self.lp[i]), _ptr(self.dlp[i])) for i in (0, 1)]
            self._dargs_static = True
        a.y0, a.f0, a.lp0, a.dlp0 = self._dargs_ptrs[c]
        a.y1, a.f1, a.lp1, a.dlp1 = self._dargs_ptrs[n]
        ev_rows_to_struct(a.ev, ev)
        C.memmove(C.addressof(a.cb), np.ascontiguousarray(cb, np.float32).ctypes.data, 144)
        C.memmove(C.addressof(a.ce), np.ascontiguousarray(ce, np.float32).ctypes.data, 28)
        C.memmove(C.addressof(a.cm), np.ascontiguousarray(cm, np.float32).ctypes.data, 28)
        a.dt, a.atol, a.rtol, a.x_interp, a.final = float(dt32), float(atol), float(rtol), float(x_interp), int(final)
        if self.B:
            with _timed("dopri5_attempt", self.B):
                L.check(self.lib.ffb_dopri5_attempt(C.byref(self.field.c), C.byref(a), self._stream), "ffb_dopri5_attempt")
        return self._reduce()

    # -- other adaptive Runge-Kutta tableaus (bosh3, adaptive_heun, fehlberg2): evaluation-at-a-time attempts -------------
    def _rk_buffers(self, n_k):
        if getattr(self, "_rk_n", 0) < n_k:
            dev, B, D = self.dev, self.B, self.D
            self._rk_k = [torch.empty(B, D, device=dev) for _ in range(max(n_k - 2, 0))]          # k2 .. k_{n-1}
            self._rk_d = [torch.empty(B, device=dev) for _ in range(max(n_k - 2, 0))] if self.with_lp else []
            self._rk_ystage = torch.empty(B, D, device=dev)
            self._rk_partials = torch.zeros(L.STAGED_BLOCKS, L.NPART, dtype=torch.float64, device=dev)
            self._rk_sums = torch.zeros(L.NPART, dtype=torch.float64, device=dev)
            self._rk_cargs, self._rk_fargs = L.RkCombineArgs(), L.RkFinishArgs()
            self._rk_n = n_k

    def _feval_into(self, y, ev_row, f_out, dlp_out):
        """f(y) -> f_out (and the divergence -> dlp_out) with one ``ffb_field_eval`` launch."""
        a = self.eargs
        a.batch, a.y, a.fbase, a.dlpbase, a.h = self.B, _ptr(y), None, None, 0.0
        a.cond, a.probes, a.cond_state = _ptr(self.cond), _ptr(self.probes), _ptr(self.cond_state)
        a.f, a.dlp, a.jac = _ptr(f_out), _ptr(dlp_out), None
        ev_rows_to_struct(a.ev, ev_row[None, :])
        a.atol, a.rtol, a.norms, a.cond_in_state = 1.0, 0.0, 0, int(self.cond_in_state)
        a.partials, a.status, a.scratch = _ptr(self.partials), _ptr(self.status), _ptr(self.scratch)
        with _timed("field_eval", self.B):
            L.check(self.lib.ffb_field_eval(C.byref(self.field.c), C.byref(a), self._stream), "ffb_field_eval")

    def _combine_into(self, y0, ks, coefs, out):
        ca = self._rk_cargs
        ca.n, ca.n_terms, ca.y0, ca.out = y0.numel(), len(ks), _ptr(y0), _ptr(out)
        for j in range(7):
            ca.k[j] = ks[j].data_ptr() if j < len(ks) else None
            ca.coef[j] = float(coefs[j]) if j < len(ks) else 0.0
        L.check(self.lib.ffb_rk_combine(C.byref(ca), self._stream), "ffb_rk_combine")

    def attempt_rk(self, tab, ev, dt32, atol, rtol, final, x_interp):
        """One attempted step of the adaptive Runge-Kutta method ``tab`` (solver.Tableau): per stage ``ffb_rk_combine``
        (stage input, torchdiffeq's ``y0 + k[..., :i+1] @ (beta_i dt)``) + ``ffb_field_eval``, then ``ffb_rk_finish``
        (error sums, log-det column, dense output).  Same partial sums as ``attempt``."""
        n_k = len(tab.alpha) + 1
        self._rk_buffers(n_k)
        c, n = self.cur, 1 - self.cur
        ks = [self.f[c]] + self._rk_k[: n_k - 2] + [self.f[n]]               # k1 .. k_n ; f1 = the last one
        dks = ([self.dlp[c]] + self._rk_d[: n_k - 2] + [self.dlp[n]]) if self.with_lp else [None] * n_k
        dt = np.float32(dt32)
        if self.B:
            for i in range(n_k - 1):
                last = (i == n_k - 2)
                out = self.y[n] if (last and tab.fsal) else self._rk_ystage
                self._combine_into(self.y[c], ks[: i + 1], [np.float32(b) * dt for b in tab.beta[i]], out)
                self._feval_into(out, ev[i], ks[i + 1], dks[i + 1])
            if not tab.fsal:                                                  # y1 = y0 + k @ (dt c_sol)
                self._combine_into(self.y[c], ks, [np.float32(cs) * dt for cs in tab.c_sol], self.y[n])
            fa = self._rk_fargs
            fa.batch, fa.dim, fa.final, fa.n_k = self.B, self.D, int(final), n_k
            fa.y0, fa.y1 = _ptr(self.y[c]), _ptr(self.y[n])
            fa.lp0, fa.lp1 = (_ptr(self.lp[c]), _ptr(self.lp[n])) if self.with_lp else (None, None)
            for j in range(7):
                fa.k[j] = ks[j].data_ptr() if j < n_k else None
                fa.dlp[j] = dks[j].data_ptr() if (j < n_k and self.with_lp) else None
                fa.ce[j] = float(np.float32(tab.c_err[j]) * dt) if j < n_k else 0.0
                fa.cm[j] = float(np.float32(tab.c_mid[j]) * dt) if j < n_k else 0.0
            for j in range(6):
                fa.cl[j] = float(np.float32(tab.c_sol[j]) * dt) if j < n_k else 0.0
            fa.dt, fa.atol, fa.rtol, fa.x_interp = float(dt), float(atol), float(rtol), float(x_interp)
            fa.y_out, fa.lp_out, fa.partials = _ptr(self.y_out), _ptr(self.lp_out), _ptr(self._rk_partials)
            L.check(self.lib.ffb_rk_finish(C.byref(fa), self._stream), "ffb_rk_finish")
        L.check(self.lib.ffb_reduce_partials(_ptr(self._rk_partials), L.STAGED_BLOCKS, _ptr(self._rk_sums), self._stream),
                "ffb_reduce_partials")
        return self._rk_sums

    def refresh_f(self, ev_row):
        """f (and the divergence) of the CURRENT state again, at the evaluation described by ``ev_row``: torchdiffeq's
        re-evaluation just after a ``jump_t`` discontinuity."""
        if self.B:
            self._feval_into(self.y[self.cur], ev_row, self.f[self.cur], self.dlp[self.cur])

    def accept(self):
        self.cur = 1 - self.cur

    # -- device-side controller (csrc/ffb_control.cuh): the host only enqueues and polls ----------------
    def ctl_supported(self):
        """A property of the FIELD (not of this rank's shard): every rank of a sharded solve must take the same controller,
        or their sequences of collectives differ.  A rank with an empty shard runs the same loop with no-op attempts."""
        return bool(self.lib.ffb_dopri5_ctl_supported(C.byref(self.field.c)))

    def ctl_attempt_ms_estimate(self, rows=None):
        """Rough duration of one attempted step (6 evaluations at ~140 TFLOP/s), for solver.py's "auto" choice.
        ``rows``: rows per rank to assume (solver.py passes the GLOBAL batch / world size so that all ranks agree)."""
        rows = self.B if rows is None else rows
        return 6.0 * rows * self.field.flops_per_eval() / 140e12 * 1e3

    def ctl_begin(self, params: "L.CtlParams", t: float, dt_next: float, grid_idx: int, atol, rtol):
        """Upload the controller block and let the device prepare the first attempt."""
        self._ctl_params = params
        host = L.Ctl()
        host.t, host.dt_next, host.grid_idx = float(t), float(dt_next), int(grid_idx)
        # pinned host memory is device-accessible (UVA): the control kernel stores its progress there.  The two pinned
        # buffers are allocated once per host thread (cudaHostAlloc is slow and serialises against running copies);
        # a solve owns them until ctl_finish() has synchronised, so the next solve of this thread may reuse them.
        notify, staging = _ctl_pinned()
        self._ctl_notify_np = notify.numpy()
        self._ctl_notify_np[:] = 0
        host.notify = notify.data_ptr()
        staging.numpy()[:] = np.frombuffer(bytes(host), np.uint8)
        self.ctl_dev = torch.empty(C.sizeof(L.Ctl), dtype=torch.uint8, device=self.dev)
        self.ctl_dev.copy_(staging, non_blocking=True)
        self._ctl_ptr = C.c_void_p(self.ctl_dev.data_ptr())
        L.check(self.lib.ffb_dopri5_control(C.byref(params), None, None, 0, self._ctl_ptr, 0, self._stream), "ffb_dopri5_control")
        a = self.dargs
        c, n = self.cur, 1 - self.cur
        a.batch = self.B
        a.cond, a.probes = _ptr(self.cond), _ptr(self.probes)
        a.y_out, a.lp_out = _ptr(self.y_out), _ptr(self.lp_out)
        a.partials, a.status, a.scratch = _ptr(self.partials), _ptr(self.status), _ptr(self.scratch)
        a.y0, a.f0, a.lp0, a.dlp0 = _ptr(self.y[c]), _ptr(self.f[c]), _ptr(self.lp[c]), _ptr(self.dlp[c])
        a.y1, a.f1, a.lp1, a.dlp1 = _ptr(self.y[n]), _ptr(self.f[n]), _ptr(self.lp[n]), _ptr(self.dlp[n])
        a.atol, a.rtol, a.final = float(atol), float(rtol), 0
        a.ctl = self._ctl_ptr
        self._dargs_static = False                       # a later host-driven solve refills the block
        self._ctl_first_record = len(profiler.records)
        self._ctl_sums_ptr = _ptr(self.sums)

    def ctl_attempt(self, reduce=True):
        """attempt (+ tile reduction when the sums go through an all-reduce first); the attempt returns at once when
        the solve has already finished"""
        if self.B:                      # a rank with an empty shard only takes part in the reductions
            with _timed("dopri5_attempt", self.B):
                L.check(self.lib.ffb_dopri5_attempt(C.byref(self.field.c), C.byref(self.dargs), self._stream), "ffb_dopri5_attempt")
        return self._reduce() if reduce else self.sums

    def ctl_control(self, reduce=False):
        """controller turn; reduce=True folds the tile reduction into the same launch (no all-reduce in between)"""
        L.check(self.lib.ffb_dopri5_control(C.byref(self._ctl_params), self._ctl_sums_ptr,
                                            _ptr(self.partials) if reduce else None, self.ntiles, self._ctl_ptr, 1,
                                            self._stream), "ffb_dopri5_control")

    def ctl_wait_turn(self, k: int) -> int:
        """Block until the controller has taken turn k (1-based); returns its `done` code as of that turn."""
        arr, i, want = self._ctl_notify_np, (k - 1) % L.CTL_NOTIFY_SLOTS, k & 0xFFFFFF
        spins = 0
        while True:
            v = int(arr[i]) & 0xFFFFFFFF
            if (v >> 8) == want:
                d = v & 0xFF
                return d - 256 if d > 127 else d
            spins += 1
            if spins % 4096 == 0 and torch.cuda.current_stream(self.dev).query():
                v = int(arr[i]) & 0xFFFFFFFF          # the stream has drained: the turn must have been announced
                if (v >> 8) != want:
                    raise L.FFBError("dopri5 controller: turn %d was never announced (kernel fault?)" % k)

    def ctl_finish(self) -> "L.Ctl":
        """Synchronise and fetch the controller block (statistics, final state of the loop)."""
        host = L.Ctl.from_buffer_copy(self.ctl_dev.cpu().numpy().tobytes())
        self.dargs.ctl = None
        if host.cur:
            self.cur = 1 - self.cur
        if profiler.enabled:       # launches enqueued after the solve had finished did no work: not attempts
            keep, seen = [], 0
            for idx, rec in enumerate(profiler.records):
                if idx >= self._ctl_first_record and rec[0] == "dopri5_attempt":
                    seen += 1
                    if seen > host.n_attempts:
                        continue
                keep.append(rec)
            profiler.records[:] = keep
        return host

    def output(self):
        return self.y_out, self.lp_out


class TraceEstimator:
    """Which stochastic trace estimator a staged solve uses, with its fixed probes (reference layout (n, B, D),
    `diffusion.py:703-721`).  kind: L.TRACE_HUTCHPP (S, G) or L.TRACE_XTRACE (S = O)."""

    def __init__(self, kind: int, S: torch.Tensor, G: Optional[torch.Tensor] = None):
        self.kind, self.S, self.G = kind, S, G
        if S.dim() != 3 or (G is not None and G.dim() != 3):
            raise ValueError("probes must have shape (n_vectors, batch, dim)")
        self.rank = int(S.shape[0])
        self.nvec = int(G.shape[0]) if G is not None else 0
        D = int(S.shape[2])
        if D > L.TRACE_MAX_DIM or self.rank > L.TRACE_MAX_RANK:
            raise NotImplementedError(f"Hutch++ / XTrace hold D <= {L.TRACE_MAX_DIM} (a sample and its tangents share one tile) and rank <= "
                                      f"{L.TRACE_MAX_RANK} (got D = {D}, rank = {self.rank})")
        if self.rank > D:
            raise ValueError("rank must not exceed the state dimension")


class StagedBackend(CudaBackend):
    """Adaptive solve with a Hutch++ / XTrace divergence (`diffusion.py:336-481`): every evaluation is
    ``ffb_field_eval`` (tangent-row engine, full Jacobian out) + ``ffb_trace_estimate``; an attempt is six of them
    between ``ffb_rk_combine`` launches plus ``ffb_rk_finish`` (csrc/ffb_staged.cu).  Same interface and the same
    partial sums as ``CudaBackend``, so ``solver.dopri5``'s host controller drives it unchanged."""

    def __init__(self, field: FieldSpec, y0: torch.Tensor, estimator: TraceEstimator, cond=None):
        if field.div_mode != L.DIV_EXACT or len(field.nets) != 1:
            raise ValueError("staged solves run on a one-network field with div_mode DIV_EXACT")
        super().__init__(field, y0, cond=cond, probes=None, with_lp=True)
        dev, B, D = self.dev, self.B, self.D
        if tuple(estimator.S.shape[1:]) != (B, D) or (estimator.G is not None and tuple(estimator.G.shape[1:]) != (B, D)):
            raise ValueError("probe shapes do not match the state")
        self.est = estimator
        self.S = _dev_f32(estimator.S, dev)
        self.G = None if estimator.G is None else _dev_f32(estimator.G, dev)
        self.jac = torch.empty(B, D, D, device=dev)
        self.ks = [torch.empty(B, D, device=dev) for _ in range(5)]       # k2..k6
        self.dks = [torch.empty(B, device=dev) for _ in range(5)]
        self.ystage = torch.empty(B, D, device=dev)
        self.partials2 = torch.zeros(L.STAGED_BLOCKS, L.NPART, dtype=torch.float64, device=dev)
        self.sums2 = torch.zeros(L.NPART, dtype=torch.float64, device=dev)
        self.targs, self.cargs, self.fargs = L.TraceArgs(), L.RkCombineArgs(), L.RkFinishArgs()
        t = self.targs
        t.batch, t.dim, t.kind, t.rank, t.nvec = B, D, estimator.kind, estimator.rank, estimator.nvec
        t.jac, t.S, t.G = _ptr(self.jac), _ptr(self.S), _ptr(self.G)
        t.score, t.use_sigma, t.has_drift = int(field.kind == L.FIELD_SCORE), int(field.use_sigma), int(field.has_drift)
        t.partials = _ptr(self.partials2)

    def ctl_supported(self):
        return False                      # the device controller drives the fused attempt kernel only

    def _reduce2(self):
        L.check(self.lib.ffb_reduce_partials(_ptr(self.partials2), L.STAGED_BLOCKS, _ptr(self.sums2), self._stream),
                "ffb_reduce_partials")
        return self.sums2

    def _field_and_trace(self, y, ev_row, f_out, dlp_out, *, fbase=None, h=0.0, norms=0, atol=1.0, rtol=0.0, dlpbase=None):
        a = self.eargs
        a.batch = self.B
        a.y, a.fbase, a.dlpbase, a.h = _ptr(y), _ptr(fbase), _ptr(dlpbase), float(h)
        a.cond, a.probes, a.cond_state = _ptr(self.cond), None, None
        a.f, a.dlp, a.jac = _ptr(f_out), _ptr(dlp_out), _ptr(self.jac)
        ev_rows_to_struct(a.ev, ev_row[None, :])
        a.atol, a.rtol, a.norms, a.cond_in_state = float(atol), float(rtol), int(norms), 0
        a.partials, a.status, a.scratch = _ptr(self.partials), _ptr(self.status), _ptr(self.scratch)
        t = self.targs
        t.a, t.c = float(ev_row[L.MAX_TFEAT + 0]), float(ev_row[L.MAX_TFEAT + 1])
        t.sigma, t.sign = float(ev_row[L.MAX_TFEAT + 2]), float(ev_row[L.MAX_TFEAT + 3])
        t.dlp, t.norms, t.atol, t.dlpbase = _ptr(dlp_out), int(norms), float(atol), _ptr(dlpbase)
        if self.B:
            with _timed("field_eval_jac", self.B):
                L.check(self.lib.ffb_field_eval(C.byref(self.field.c), C.byref(a), self._stream), "ffb_field_eval")
            # the estimator overwrites the exact trace the evaluation left in dlp_out
            with _timed("trace_estimate", self.B):
                L.check(self.lib.ffb_trace_estimate(C.byref(t), self._stream), "ffb_trace_estimate")

    def _eval(self, ev_row, atol, rtol, norms, h=None):
        c = self.cur
        if norms == 1:
            self._field_and_trace(self.y[c], ev_row, self.f[c], self.dlp[c], norms=1, atol=atol, rtol=rtol)
        else:       # f(y + h f0) against f0: only the norms are kept (k7 scratch buffers take the outputs)
            self._field_and_trace(self.y[c], ev_row, self.f[1 - c], self.dlp[1 - c], fbase=self.f[c], h=h, norms=2,
                                  atol=atol, rtol=rtol, dlpbase=self.dlp[c])
        sums = self._reduce()
        sums2 = self._reduce2()
        sums[L.P_LP_F:L.P_LP_DF + 1] = sums2[L.P_LP_F:L.P_LP_DF + 1]      # log-det norms come from the estimator
        return sums

    def attempt(self, ev, cb, ce, cm, dt32, atol, rtol, final, x_interp):
        c, n = self.cur, 1 - self.cur
        ks = [self.f[c]] + self.ks + [self.f[n]]            # k1 .. k7 (FSAL: k7 = f1)
        dks = [self.dlp[c]] + self.dks + [self.dlp[n]]
        cb = np.ascontiguousarray(cb, np.float32)
        ca = self.cargs
        ca.n, ca.y0 = self.B * self.D, _ptr(self.y[c])
        for i in range(6):
            out = self.y[n] if i == 5 else self.ystage          # the 7th stage's input is y1
            ca.n_terms, ca.out = i + 1, _ptr(out)
            for j in range(7):
                ca.k[j] = ks[j].data_ptr() if j <= i else None
                ca.coef[j] = float(cb[i, j]) if j <= i else 0.0
            if self.B:
                L.check(self.lib.ffb_rk_combine(C.byref(ca), self._stream), "ffb_rk_combine")
                self._field_and_trace(out, ev[i], ks[i + 1], dks[i + 1])
        fa = self.fargs
        fa.batch, fa.dim, fa.final = self.B, self.D, int(final)
        fa.y0, fa.y1, fa.lp0, fa.lp1 = _ptr(self.y[c]), _ptr(self.y[n]), _ptr(self.lp[c]), _ptr(self.lp[n])
        for j in range(7):
            fa.k[j], fa.dlp[j] = ks[j].data_ptr(), dks[j].data_ptr()
            fa.ce[j], fa.cm[j] = float(ce[j]), float(cm[j])
        for j in range(6):
            fa.cl[j] = float(cb[5, j])
        fa.dt, fa.atol, fa.rtol, fa.x_interp = float(dt32), float(atol), float(rtol), float(x_interp)
        fa.y_out, fa.lp_out, fa.partials = _ptr(self.y_out), _ptr(self.lp_out), _ptr(self.partials2)
        if self.B:
            L.check(self.lib.ffb_rk_finish(C.byref(fa), self._stream), "ffb_rk_finish")
        return self._reduce2()

    def single_eval(self, ev_row):
        self._field_and_trace(self.y[self.cur], ev_row, self.f[self.cur], self.dlp[self.cur])
        return self.f[self.cur], self.dlp[self.cur]

    def refresh_f(self, ev_row):
        if self.B:
            self._field_and_trace(self.y[self.cur], ev_row, self.f[self.cur], self.dlp[self.cur])

    attempt_rk = None          # other tableaus are not offered with Hutch++ / XTrace (solver._dopri5 checks hasattr / None)

    # -- primitives of the fixed-grid driver (staged_fixed) ------------------------------------------------
    def state(self):
        return self.y[self.cur], self.lp[self.cur]

    def feval(self, y, ev_row):
        """(f, d log p / dt) at ``y`` into fresh buffers."""
        f, d = torch.empty(self.B, self.D, device=self.dev), torch.empty(self.B, device=self.dev)
        self._field_and_trace(y, ev_row, f, d)
        return f, d

    def combine(self, y0, ks, coefs):
        """y0 + sum_j coefs[j] * ks[j]  (ffb_rk_combine)."""
        out = torch.empty_like(y0)
        ca = self.cargs
        ca.n, ca.n_terms, ca.y0, ca.out = y0.numel(), len(ks), _ptr(y0), _ptr(out)
        for j in range(7):
            ca.k[j] = ks[j].data_ptr() if j < len(ks) else None
            ca.coef[j] = float(coefs[j]) if j < len(ks) else 0.0
        if y0.numel():
            L.check(self.lib.ffb_rk_combine(C.byref(ca), self._stream), "ffb_rk_combine")
        return out


def staged_fixed(be, method: str, dts: np.ndarray, ev: np.ndarray):
    """torchdiffeq's fixed-grid steppers (euler, midpoint, rk4 = 3/8 rule) on a staged backend: one
    ``feval`` per stage, stage inputs and the step itself through ``combine``.  ``dts`` (n,) float32 step sizes,
    ``ev`` (n, evaluations per step, EV_FLOATS).  Returns (x, log-det column) at the end of the grid."""
    third, half, eighth = np.float32(1.0 / 3.0), np.float32(0.5), np.float32(0.125)
    dev = getattr(be, "dev", None)
    if dev is None or torch.device(dev).type != "cuda":      # the CPU model of the kernels (tests/kernel_model.py)
        return _staged_fixed(be, method, dts, ev, third, half, eighth)
    with on_device(dev):
        return _staged_fixed(be, method, dts, ev, third, half, eighth)


def _staged_fixed(be, method, dts, ev, third, half, eighth):
    y, lp = be.state()
    for s in range(int(dts.shape[0])):
        dt = np.float32(dts[s])
        k1, d1 = be.feval(y, ev[s, 0])
        if method == "euler":
            ks, ds, cf = [k1], [d1], [dt]
        elif method == "midpoint":
            k2, d2 = be.feval(be.combine(y, [k1], [half * dt]), ev[s, 1])
            ks, ds, cf = [k2], [d2], [dt]
        elif method == "rk4":
            k2, d2 = be.feval(be.combine(y, [k1], [dt * third]), ev[s, 1])
            k3, d3 = be.feval(be.combine(y, [k1, k2], [-(dt * third), dt]), ev[s, 2])
            k4, d4 = be.feval(be.combine(y, [k1, k2, k3], [dt, -dt, dt]), ev[s, 3])
            e = dt * eighth
            ks, ds, cf = [k1, k2, k3, k4], [d1, d2, d3, d4], [e, np.float32(3) * e, np.float32(3) * e, e]
        else:
            raise ValueError(method)
        y, lp = be.combine(y, ks, cf), be.combine(lp, ds, cf)
    return y, lp


def run_fixed(field: FieldSpec, method: int, x0: torch.Tensor, step_table: np.ndarray, ev_table: np.ndarray,
              cond=None, probes=None, lp0=None, noise=None, philox=None, row_offset=0, want_lp=False):
    """Whole fixed-grid trajectory in one kernel.  ``step_table`` (nsteps, 8) and ``ev_table``
    (nsteps, nev, EV_FLOATS) are host float32 arrays computed in the reference's op order."""
    require_cuda(x0, "state")
    field.check_device(x0)
    lib = L.load()
    dev = x0.device
    x0 = _dev_f32(x0, dev)
    B, D = x0.shape
    nsteps = int(step_table.shape[0])
    st = torch.from_numpy(np.ascontiguousarray(step_table, np.float32)).to(dev)
    ev = torch.from_numpy(np.ascontiguousarray(ev_table, np.float32).reshape(-1)).to(dev)
    x_out = torch.empty_like(x0)
    lp_out = torch.empty(B, device=dev) if want_lp else None
    status = torch.tensor([0, 2 ** 31 - 1], dtype=torch.int32, device=dev)      # [FFB_ST_* bits, first NaN step of EM]
    scratch = field.scratch(dev)
    a = L.FixedArgs()
    a.batch, a.method, a.nsteps = B, method, nsteps
    lp0 = None if lp0 is None else _dev_f32(lp0, dev)      # keep a reference until the launch
    a.x0, a.lp0 = _ptr(x0), _ptr(lp0)
    cond = None if cond is None else _dev_f32(cond, dev)
    probes = None if probes is None else _dev_f32(probes, dev)
    noise = None if noise is None else _dev_f32(noise, dev)
    a.cond, a.probes, a.noise = _ptr(cond), _ptr(probes), _ptr(noise)
    seed, offset = philox if philox is not None else (0, 0)
    a.philox_seed, a.philox_offset, a.row_offset = int(seed), int(offset), int(row_offset)
    a.x_out, a.lp_out = _ptr(x_out), _ptr(lp_out)
    a.step_table, a.ev_table = _ptr(st), _ptr(ev)
    a.status, a.scratch = _ptr(status), _ptr(scratch)
    if B and nsteps:
        with _timed("integrate_fixed", B):
            with on_device(dev):
                L.check(lib.ffb_integrate_fixed(C.byref(field.c), C.byref(a), _stream(dev)), "ffb_integrate_fixed")
    elif B:
        x_out.copy_(x0)
    return x_out, lp_out, status


def gaussian_logprob(x: torch.Tensor, add: Optional[torch.Tensor], sigma: float = 1.0) -> torch.Tensor:
    """sum_d log N(x_d; 0, sigma) + add   (flow.py:434, diffusion.py:814, symplectic.py:240-243)."""
    require_cuda(x, "x")
    x = _dev_f32(x, x.device)
    B, D = x.shape
    out = torch.empty(B, device=x.device)
    if B == 0:
        return out
    add_c = None if add is None else _dev_f32(add, x.device)
    with on_device(x.device):
        L.check(L.load().ffb_gaussian_logprob(_ptr(x), _ptr(add_c), _ptr(out), B, D, float(sigma), _stream(x.device)),
                "ffb_gaussian_logprob")
    return out


def philox_normal(batch, dim, seed, offset, step, row_offset=0, device="cuda"):
    out = torch.empty(batch, dim, device=device)
    with on_device(out.device):
        L.check(L.load().ffb_philox_normal(_ptr(out), batch, dim, int(seed), int(offset), int(step), int(row_offset),
                                           _stream(out.device)), "ffb_philox_normal")
    return out


def ffma_peak_tflops(iters=4096):
    v = C.c_float(0)
    L.check(L.load().ffb_ffma_peak(iters, C.byref(v), _stream()), "ffb_ffma_peak")
    return float(v.value)


def device_info():
    vals = [C.c_int32(0) for _ in range(5)]
    L.check(L.load().ffb_device_info(*[C.byref(v) for v in vals]), "ffb_device_info")
    return dict(zip(("sm_count", "smem_optin", "cc_major", "cc_minor", "clock_khz"), (v.value for v in vals)))
