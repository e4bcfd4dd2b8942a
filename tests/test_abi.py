"""The C-ABI boundary without a GPU: libffb200.so builds for sm_100a, loads, and exports exactly the
symbols include/ffb200.h declares; the ctypes mirrors in flowfusion_b200/_lib.py have the sizes the
header's structs have; argument errors come back as status codes, not crashes.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

from flowfusion_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ffb200.h")


@pytest.fixture(scope="module")
def lib():
    L.build(force=False)          # nvcc cross-compiles for sm_100a without a GPU (about a minute when stale)
    return L.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ffb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ffb200.h but not exported by libffb200.so"
        assert n in L.SYMBOLS, f"{n} declared in include/ffb200.h but not bound in _lib.SYMBOLS"
    for n in L.SYMBOLS:
        assert n in names, f"{n} bound in _lib.SYMBOLS but not declared in include/ffb200.h"


def test_exports_are_plain_c(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for n in declared_functions():
        assert n in exported                      # unmangled => extern "C"


def test_abi_version_and_error_paths(lib):
    assert lib.ffb_abi_version() == L.ABI_VERSION == 5
    # null arguments are reported through the status code + ffb_last_error, no CUDA call is made
    assert lib.ffb_net_create(None, None, None) == -1
    assert b"null" in lib.ffb_last_error()
    assert lib.ffb_reduce_partials(None, 0, None, None) == -1
    assert lib.ffb_gaussian_logprob(None, None, None, 0, 0, 1.0, None) == -1
    assert lib.ffb_num_tiles(None, 10) == -1
    lib.ffb_net_destroy(None)                     # must be a no-op
    assert lib.ffb_trace_estimate(None, None) == -1 and lib.ffb_rk_combine(None, None) == -1
    assert lib.ffb_rk_finish(None, None) == -1


def test_struct_sizes_match_the_header(lib, tmp_path):
    """Compile a 10-line C program against include/ffb200.h and compare sizeof() with the ctypes mirrors."""
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "ffb200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(ffb_net_desc),sizeof(ffb_field),sizeof(ffb_eval_scalars),sizeof(ffb_eval_args),"
                   "sizeof(ffb_dopri5_args),sizeof(ffb_fixed_args),sizeof(ffb_time_program),"
                   "sizeof(ffb_dopri5_ctl_params),sizeof(ffb_dopri5_ctl),sizeof(ffb_trace_args),sizeof(ffb_rk_combine_args),"
                   "sizeof(ffb_rk_finish_args),sizeof(ffb_train_args),sizeof(ffb_hamiltonian_args));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(t) for t in (L.NetDesc, L.Field, L.EvalScalars, L.EvalArgs, L.Dopri5Args, L.FixedArgs,
                                  L.TimeProgram, L.CtlParams, L.Ctl, L.TraceArgs, L.RkCombineArgs, L.RkFinishArgs, L.TrainArgs, L.HamiltonianArgs)]
    assert got == want


def test_library_holds_sm100a_tensor_core_code(lib):
    sass = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper()
    for needle in ("UTCHMMA", "UBLKCP", "SYNCS", "FFMA2"):       # tcgen05.mma, bulk copy, mbarrier, packed FP32 FMA
        assert needle in sass, needle


def test_net_create_refuses_shapes_beyond_the_header_limits(lib):
    """include/ffb200.h: FFB_MAX_LAYERS = 16 Linear layers, FFB_MAX_WIDTH = 512 columns.  The checks run before any CUDA call."""
    d = L.NetDesc()
    d.n_layers, d.in_features, d.x_dim, d.c_dim, d.t_dim = 2, 5, 4, 0, 1
    d.widths[0], d.widths[1] = L.MAX_WIDTH + 1, 4
    h = C.c_void_p()
    assert lib.ffb_net_create(C.byref(d), None, C.byref(h)) == -1 and b"1..512" in lib.ffb_last_error()
    d.widths[0] = 64
    d.n_layers = L.MAX_LAYERS + 1
    assert lib.ffb_net_create(C.byref(d), None, C.byref(h)) == -1 and b"1..16" in lib.ffb_last_error()
    d.n_layers = 2
    d.widths[1] = 200                              # the output layer feeds the ODE state: at most 128 columns
    assert lib.ffb_net_create(C.byref(d), None, C.byref(h)) == -1 and b"output layer" in lib.ffb_last_error()
