"""ctypes binding of ``libffb200.so`` (C ABI declared in ``include/ffb200.h``).

There is NO CPU fallback: if the shared library is missing or no CUDA device is present,
every entry point raises.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("FFB_LIB") or os.path.join(_HERE, "libffb200.so")
SOURCES = [os.path.join(_HERE, "csrc", n) for n in ("ffb_kernels.cu", "ffb_staged.cu", "ffb_rd.cu", "ffb_wide.cu", "ffb_wide2.cu", "ffb_train.cu")]
HEADERS = [os.path.join(_HERE, "csrc", h) for h in ("ffb_common.cuh", "ffb_engine.cuh", "ffb_engine_tc.cuh", "ffb_engine_rr.cuh",
                                                     "ffb_kernels_rr.cuh", "ffb_engine_rrt.cuh", "ffb_control.cuh", "ffb_engine_rd.cuh",
                                                     "ffb_kernels_rd.cuh", "ffb_rd.h", "ffb_kernels_generic.cuh", "ffb_engine_wide.cuh", "ffb_wide.h")] + \
          [os.path.join(ROOT, "include", "ffb200.h")]
OBJ_DIR = os.path.join(ROOT, "build", "obj")

ABI_VERSION = 5
MAX_LAYERS, MAX_TFEAT, NPART, STEP_STRIDE, TILE_ROWS = 16, 32, 16, 8, 128
MAX_WIDTH = 512
FIELD_NET, FIELD_SCORE = 0, 1
DIV_NONE, DIV_EXACT, DIV_HUTCH = 0, 1, 2
M_EULER, M_MIDPOINT, M_RK4, M_EM, M_LEAPFROG = 0, 1, 2, 3, 4
ACT_SILU, ACT_TANH, ACT_RELU, ACT_SOFTPLUS, ACT_GELU = 0, 1, 2, 3, 4
ST_NONFINITE_STATE, ST_NAN_SAMPLE = 1, 2
# indices into the per-tile partial sums
P_X_Y, P_X_F, P_X_DF, P_X_ERR, P_LP_Y, P_LP_F, P_LP_DF, P_LP_ERR, P_C_Y, P_NONFINITE = range(10)

c_float_p = C.POINTER(C.c_float)


class NetDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("in_features", C.c_int32), ("widths", C.c_int32 * MAX_LAYERS),
                ("weight", C.c_void_p * MAX_LAYERS), ("bias", C.c_void_p * MAX_LAYERS),
                ("x_col", C.c_int32), ("x_dim", C.c_int32), ("c_col", C.c_int32), ("c_dim", C.c_int32),
                ("t_col", C.c_int32), ("t_dim", C.c_int32), ("activation", C.c_int32)]


class Field(C.Structure):
    _fields_ = [("n_calls", C.c_int32), ("net", C.c_void_p * 2), ("in_off", C.c_int32 * 2),
                ("out_off", C.c_int32 * 2), ("out_sign", C.c_float * 2), ("state_dim", C.c_int32),
                ("cond_dim", C.c_int32), ("kind", C.c_int32), ("use_sigma", C.c_int32),
                ("has_drift", C.c_int32), ("div_mode", C.c_int32)]


class EvalScalars(C.Structure):
    _fields_ = [("tfeat", C.c_float * MAX_TFEAT), ("a", C.c_float), ("c", C.c_float),
                ("sigma", C.c_float), ("sign", C.c_float)]


EV_FLOATS = MAX_TFEAT + 4


class EvalArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("y", C.c_void_p), ("fbase", C.c_void_p), ("dlpbase", C.c_void_p),
                ("h", C.c_float), ("cond", C.c_void_p), ("cond_state", C.c_void_p), ("probes", C.c_void_p),
                ("f", C.c_void_p),
                ("dlp", C.c_void_p), ("ev", EvalScalars), ("atol", C.c_float), ("rtol", C.c_float),
                ("norms", C.c_int32), ("cond_in_state", C.c_int32), ("partials", C.c_void_p),
                ("status", C.c_void_p), ("scratch", C.c_void_p), ("jac", C.c_void_p)]


class Dopri5Args(C.Structure):
    _fields_ = [("batch", C.c_int64), ("y0", C.c_void_p), ("f0", C.c_void_p), ("lp0", C.c_void_p),
                ("dlp0", C.c_void_p), ("cond", C.c_void_p), ("probes", C.c_void_p), ("y1", C.c_void_p),
                ("f1", C.c_void_p), ("lp1", C.c_void_p), ("dlp1", C.c_void_p), ("y_out", C.c_void_p),
                ("lp_out", C.c_void_p), ("ev", EvalScalars * 6), ("cb", (C.c_float * 6) * 6),
                ("ce", C.c_float * 7), ("cm", C.c_float * 7), ("dt", C.c_float), ("atol", C.c_float),
                ("rtol", C.c_float), ("x_interp", C.c_float), ("final", C.c_int32),
                ("partials", C.c_void_p), ("status", C.c_void_p), ("scratch", C.c_void_p),
                ("ctl", C.c_void_p)]


# ---- device-side dopri5 controller (include/ffb200.h, csrc/ffb_control.cuh) ---------------------------
PROG_RAW_T, PROG_FOURIER = 0, 1
SDE_NONE, SDE_VP, SDE_VE, SDE_SUBVP = 0, 1, 2, 3
MAX_FREQ = MAX_TFEAT // 2
CTL_MAX_GRID, CTL_HIST, CTL_NOTIFY_SLOTS = 16, 256, 1024
CTL_RUNNING, CTL_FINISHED, CTL_NONFINITE, CTL_DT_UNDERFLOW, CTL_MAX_STEPS = 0, 1, -1, -2, -3


class TimeProgram(C.Structure):
    _fields_ = [("time_features", C.c_int32), ("n_freq", C.c_int32), ("W", C.c_float * MAX_FREQ), ("pi", C.c_float),
                ("sde", C.c_int32), ("use_sigma", C.c_int32), ("sde_mode", C.c_int32), ("T", C.c_float),
                ("beta_min", C.c_float), ("beta_diff", C.c_float), ("half_beta_diff", C.c_float),
                ("m2_beta_min", C.c_float), ("sigma_min", C.c_float), ("sigma_ratio", C.c_float),
                ("ve_gfac", C.c_float)]


class CtlParams(C.Structure):
    _fields_ = [("t_end", C.c_double), ("min_step", C.c_double), ("max_step", C.c_double), ("safety", C.c_double),
                ("ifactor", C.c_double), ("dfactor", C.c_double), ("n_x", C.c_int64), ("n_lp", C.c_int64),
                ("n_cond", C.c_int64), ("reverse", C.c_int32), ("max_num_steps", C.c_int32), ("n_grid", C.c_int32),
                ("_pad", C.c_int32), ("grid", C.c_double * CTL_MAX_GRID), ("alpha", C.c_float * 6),
                ("beta", (C.c_float * 6) * 6), ("c_err", C.c_float * 7), ("c_mid", C.c_float * 7),
                ("prog", TimeProgram)]


class Ctl(C.Structure):
    _fields_ = [("ev", EvalScalars * 6), ("cb", (C.c_float * 6) * 6), ("ce", C.c_float * 7), ("cm", C.c_float * 7),
                ("dt", C.c_float), ("x_interp", C.c_float), ("final", C.c_int32), ("cur", C.c_int32),
                ("done", C.c_int32), ("grid_idx", C.c_int32), ("t", C.c_double), ("dt_next", C.c_double),
                ("cur_t1", C.c_double), ("cur_dt", C.c_double), ("cur_on_grid", C.c_int32),
                ("n_attempts", C.c_int32), ("n_accepted", C.c_int32), ("n_rejected", C.c_int32),
                ("n_turns", C.c_int32), ("_pad", C.c_int32), ("notify", C.c_void_p),
                ("hist_dt", C.c_double * CTL_HIST), ("hist_ratio", C.c_float * CTL_HIST),
                ("hist_accept", C.c_uint8 * CTL_HIST)]


class FixedArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("method", C.c_int32), ("nsteps", C.c_int32), ("x0", C.c_void_p),
                ("lp0", C.c_void_p), ("cond", C.c_void_p), ("probes", C.c_void_p), ("noise", C.c_void_p),
                ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64), ("row_offset", C.c_int64),
                ("x_out", C.c_void_p), ("lp_out", C.c_void_p), ("step_table", C.c_void_p),
                ("ev_table", C.c_void_p), ("status", C.c_void_p), ("scratch", C.c_void_p)]


# ---- staged solves (Hutch++ / XTrace): include/ffb200.h, csrc/ffb_staged.cu ----------------------------
TRACE_HUTCHPP, TRACE_XTRACE = 1, 2
TRACE_MAX_DIM, TRACE_MAX_RANK, STAGED_BLOCKS = 124, 8, 1024


class TraceArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("dim", C.c_int32), ("kind", C.c_int32), ("rank", C.c_int32),
                ("nvec", C.c_int32), ("jac", C.c_void_p), ("S", C.c_void_p), ("G", C.c_void_p),
                ("score", C.c_int32), ("use_sigma", C.c_int32), ("has_drift", C.c_int32), ("a", C.c_float),
                ("c", C.c_float), ("sigma", C.c_float), ("sign", C.c_float), ("dlp", C.c_void_p),
                ("norms", C.c_int32), ("atol", C.c_float), ("dlpbase", C.c_void_p), ("partials", C.c_void_p)]


class RkCombineArgs(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_terms", C.c_int32), ("y0", C.c_void_p), ("k", C.c_void_p * 7),
                ("coef", C.c_float * 7), ("out", C.c_void_p)]


class RkFinishArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("dim", C.c_int32), ("final", C.c_int32), ("n_k", C.c_int32), ("_pad", C.c_int32),
                ("y0", C.c_void_p),
                ("y1", C.c_void_p), ("k", C.c_void_p * 7), ("lp0", C.c_void_p), ("dlp", C.c_void_p * 7),
                ("lp1", C.c_void_p), ("cl", C.c_float * 6), ("ce", C.c_float * 7), ("cm", C.c_float * 7),
                ("dt", C.c_float), ("atol", C.c_float), ("rtol", C.c_float), ("x_interp", C.c_float),
                ("y_out", C.c_void_p), ("lp_out", C.c_void_p), ("partials", C.c_void_p)]


# ---- fused training step (include/ffb200.h, csrc/ffb_train.cu) ------------------------------------------
class TrainArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("x_in", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p),
                ("scale", C.c_float), ("_pad", C.c_int32), ("grad_w", C.c_void_p * MAX_LAYERS),
                ("grad_b", C.c_void_p * MAX_LAYERS), ("grad_x", C.c_void_p), ("loss", C.c_void_p), ("work", C.c_void_p),
                ("cot", C.c_void_p), ("out", C.c_void_p)]


class HamiltonianArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("dim", C.c_int32), ("cond_dim", C.c_int32), ("z0", C.c_void_p), ("cond", C.c_void_p),
                ("z_out", C.c_void_p), ("h_out", C.c_void_p), ("n_steps", C.c_int32), ("dt", C.c_float), ("work", C.c_void_p)]


# every symbol include/ffb200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ffb_abi_version": (C.c_int, []),
    "ffb_last_error": (C.c_char_p, []),
    "ffb_device_info": (C.c_int, [C.POINTER(C.c_int32)] * 5),
    "ffb_net_create": (C.c_int, [C.POINTER(NetDesc), C.c_void_p, C.POINTER(C.c_void_p)]),
    "ffb_net_destroy": (None, [C.c_void_p]),
    "ffb_net_flops": (C.c_int64, [C.c_void_p]),
    "ffb_num_tiles": (C.c_int64, [C.POINTER(Field), C.c_int64]),
    "ffb_scratch_bytes": (C.c_size_t, [C.POINTER(Field)]),
    "ffb_field_eval": (C.c_int, [C.POINTER(Field), C.POINTER(EvalArgs), C.c_void_p]),
    "ffb_dopri5_attempt": (C.c_int, [C.POINTER(Field), C.POINTER(Dopri5Args), C.c_void_p]),
    "ffb_integrate_fixed": (C.c_int, [C.POINTER(Field), C.POINTER(FixedArgs), C.c_void_p]),
    "ffb_dopri5_ctl_supported": (C.c_int, [C.POINTER(Field)]),
    "ffb_dopri5_control": (C.c_int, [C.POINTER(CtlParams), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "ffb_dopri5_control_host": (C.c_int, [C.POINTER(CtlParams), C.c_void_p, C.POINTER(Ctl), C.c_int32]),
    "ffb_time_program_rows": (C.c_int, [C.POINTER(TimeProgram), C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_int32]),
    "ffb_trace_estimate": (C.c_int, [C.POINTER(TraceArgs), C.c_void_p]),
    "ffb_trace_estimate_host": (C.c_int, [C.POINTER(TraceArgs)]),
    "ffb_rk_combine": (C.c_int, [C.POINTER(RkCombineArgs), C.c_void_p]),
    "ffb_rk_finish": (C.c_int, [C.POINTER(RkFinishArgs), C.c_void_p]),
    "ffb_train_work_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.c_int64, C.c_int32]),
    "ffb_train_step": (C.c_int, [C.POINTER(NetDesc), C.POINTER(TrainArgs), C.c_void_p]),
    "ffb_hamiltonian_leapfrog": (C.c_int, [C.POINTER(NetDesc), C.POINTER(HamiltonianArgs), C.c_void_p]),
    "ffb_reduce_partials": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ffb_gaussian_logprob": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p]),
    "ffb_philox_normal": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, C.c_int64, C.c_void_p]),
    "ffb_ffma_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_float), C.c_void_p]),
    "ffb_set_engine": (C.c_int, [C.c_int]),
    "ffb_debug_trace": (C.c_int, [C.c_void_p]),
    "ffb_debug_trace_rd": (C.c_int, [C.c_void_p]),
    "ffb_get_engine": (C.c_int, []),
    "ffb_launch_count": (C.c_int64, []),
}

_lib = None


class FFBError(RuntimeError):
    pass


_NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(_HERE, "csrc"), "-Xcompiler", "-fPIC"]
# -split-compile 0 (one ptxas job per kernel, in parallel) cuts the build time but the tensor-core kernels it produces are
# 7 % slower: only the translation units without tcgen05 kernels take it
_SPLIT_COMPILE = ("ffb_staged.cu", "ffb_train.cu", "ffb_wide.cu")      # (not ffb_wide2.cu: the two-half wide kernels lose 15 % with it; the one-half ones gain 8 %)


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")


def nvcc_commands(out=LIB_PATH):
    """One compile command per translation unit (run in parallel) and the link command."""
    compiles = [["nvcc"] + _NVCC_FLAGS + (["-split-compile", "0"] if os.path.basename(src) in _SPLIT_COMPILE else []) +
                ["-c", src, "-o", _obj(src)] for src in SOURCES]
    link = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + [_obj(src) for src in SOURCES]
    return compiles, link


def nvcc_command(out=LIB_PATH):
    """The single-command equivalent (documentation, scripts/gpu_ab.sh)."""
    return ["nvcc"] + _NVCC_FLAGS + ["-shared", "-o", out] + SOURCES


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    mt = os.path.getmtime(target)
    return any(os.path.getmtime(p) > mt for p in deps)


def needs_build():
    return _stale(LIB_PATH, SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile libffb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU): the translation units are
    compiled in parallel into build/obj (only the stale ones), then linked."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    compiles, link = nvcc_commands()
    procs = []
    for src, cmd in zip(SOURCES, compiles):
        if force or _stale(_obj(src), [src] + HEADERS):
            if verbose:
                print(" ".join(cmd))
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    errors = []
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            errors.append(" ".join(cmd) + "\n" + out)
    if errors:
        raise FFBError("nvcc failed:\n" + "\n".join(errors))
    if verbose:
        print(" ".join(link))
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise FFBError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def load():
    """dlopen the library and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise FFBError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.ffb_abi_version() != ABI_VERSION:
        raise FFBError("libffb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        raise FFBError(f"{what}: {load().ffb_last_error().decode()} (status {rc})")


def launch_count():
    return int(load().ffb_launch_count())
