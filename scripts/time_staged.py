"""Timing of the staged Hutch++ / XTrace path next to the fused exact / Hutchinson log-prob on the same model
(cfg3-like score model: 16-D, 4x128).  Prints per-kernel CUDA-event times from engine.profiler."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowfusion_b200.diffusion as D
from flowfusion_b200 import engine as E

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
torch.manual_seed(1234)
x = torch.randn(B, 16, device=dev) 
out = {}
only = os.environ.get("FFB_STAGED_ONLY")
for name, flags in (("exact", {}), ("hutchinson", dict(hutchinson=True)), ("hutchpp_r1_m1", dict(hutchpp=True)),
                    ("hutchpp_r4_m4", dict(hutchpp=True, hpp_rank=4, hpp_vecs=4)), ("xtrace_m1", dict(xtrace=True)),
                    ("xtrace_m4", dict(xtrace=True, xt_vecs=4))):
    if only and name != only:
        continue
    torch.manual_seed(1234)
    sm = D.ScoreModel(D.MLP(16, 0, 8, [128] * 4), D.VPSDE(), no_sigma=True, **flags).eval().to(dev)
    for it in range(2):
        E.profiler.reset(it == 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lp = sm.log_prob(x, atol=1e-4, rtol=1e-4)
        e1.record()
        torch.cuda.synchronize()
    st = sm.last_stats
    summ = {k: dict(launches=v[0], ms=round(v[1], 3)) for k, v in E.profiler.summary().items()}
    out[name] = dict(rows=B, ms=round(e0.elapsed_time(e1), 2), nfe=st.nfe, accepted=st.accepted, rejected=st.rejected,
                     evals_per_s=round(B / (e0.elapsed_time(e1) * 1e-3)), kernels=summ,
                     finite=bool(torch.isfinite(lp).all()))
    print(name, json.dumps(out[name]), flush=True)
E.profiler.reset(False)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/staged_timing.json", "w"), indent=1)
