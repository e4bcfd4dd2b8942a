"""Where a dopri5 solve spends its time outside the attempt kernels: per-launch CUDA events of the profiler
(engine._timed) give every attempt's duration and the gap to the next one (tile reduction + controller + launch
latencies, or the host round trip with FFB_CONTROLLER=host).  Usage: python scripts/ctl_timeline.py [cfg2] [rows]"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import flowfusion_b200.diffusion as D, flowfusion_b200.flow as F, flowfusion_b200.symplectic as Sy
from flowfusion_b200 import engine, solver

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[name]["B"]
dev = torch.device("cuda:0")
model = bench.make_model(name, types.SimpleNamespace(D=D, F=F, Sy=Sy)).to(dev)
inp = {k: v.to(dev) for k, v in bench.make_inputs(name, B, 0).items()}
for mode in ("host", "device"):
    with solver.controller(mode):
        for _ in range(2):
            bench.run_gpu(name, model, inp)
        torch.cuda.synchronize()
        engine.profiler.reset(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        bench.run_gpu(name, model, inp)
        s1.record()
        torch.cuda.synchronize()
        recs = list(engine.profiler.records)
        engine.profiler.reset(False)
    total = s0.elapsed_time(s1)
    print(f"== {name} B={B} controller={mode}: solve {total:.3f} ms, {len(recs)} timed launches, steps {model.last_stats.accepted}/{model.last_stats.rejected}")
    print(f"   start -> first launch {s0.elapsed_time(recs[0][1]):.3f} ms;  last launch end -> end {recs[-1][2].elapsed_time(s1):.3f} ms")
    ksum = gsum = 0.0
    for i, (nm, e0, e1, rows) in enumerate(recs):
        d = e0.elapsed_time(e1)
        g = recs[i][2].elapsed_time(recs[i + 1][1]) if i + 1 < len(recs) else 0.0
        ksum += d; gsum += g
        print(f"   {i:2d} {nm:16s} {d:8.3f} ms   gap to next {g:7.3f} ms")
    print(f"   kernels {ksum:.3f} ms, gaps {gsum:.3f} ms")
