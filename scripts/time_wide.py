"""Timing of the wide engine (csrc/ffb_engine_wide.cuh): fixed-grid rk4 PF-ODE sampling (one launch, every evaluation on-chip)
of score networks wider than 128 / deeper than 8 layers, as a fraction of the measured FFMA2 peak; the cfg2 network on the wide
engine (FFB_ENGINE=wide) beside the tensor-core engine for scale.  Prints one JSON object.
usage: python scripts/time_wide.py [rows]"""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, '.')
import flowfusion_b200.diffusion as D
from flowfusion_b200 import _lib

lib = _lib.load()
dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 4
steps = 8
peak = C.c_float()
with torch.cuda.device(dev):
    lib.ffb_ffma_peak(20000, C.byref(peak), None)
out = {"rows": B, "rk4_steps": steps, "fp32_ffma2_peak_tflops": peak.value, "runs": []}
for units, eng in [([128] * 4, 1), ([128] * 4, 5), ([256] * 4, 1), ([512] * 4, 1), ([512] * 2, 1), ([128] * 12, 1)]:
    torch.manual_seed(1234)
    sm = D.ScoreModel(D.MLP(16, 4, 8, units), D.VPSDE(), no_sigma=True).eval().to(dev)
    base, cond = torch.randn(B, 16, device=dev), torch.randn(B, 4, device=dev)
    opt = {'step_size': (1.0 - 1e-3) / steps}
    prev = lib.ffb_get_engine()
    lib.ffb_set_engine(eng)
    sm.sample_ode_from_base(base, cond, method='rk4', options=opt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        sm.sample_ode_from_base(base, cond, method='rk4', options=opt)
    e1.record(); torch.cuda.synchronize()
    lib.ffb_set_engine(prev)
    ms = e0.elapsed_time(e1) / 3
    dims = [28] + units + [16]
    flop = sum(2 * dims[i] * dims[i + 1] for i in range(len(dims) - 1))
    tf = B * 4 * steps * flop / (ms * 1e-3) / 1e12
    out["runs"].append({"units": units, "engine": "wide" if (eng == 5 or max(units) > 128 or len(units) > 7) else "tensor",
                        "ms": ms, "row_evals_per_s": B * 4 * steps / (ms * 1e-3), "tflops": tf, "frac_of_ffma2_peak": tf / peak.value})
print(json.dumps(out))
