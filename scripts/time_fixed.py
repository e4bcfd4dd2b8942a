"""Debug/ablation timing: cfg2 network, fixed-grid rk4 (one launch), cycles per evaluation per tile.
usage: FFB_LIB=... python scripts/time_fixed.py [tiles_per_sm] [steps]"""
import sys, torch
sys.path.insert(0, '.')
import flowfusion_b200.diffusion as D
from flowfusion_b200 import _lib
_lib.load()
torch.manual_seed(1234)
dev = torch.device('cuda:0')
sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 4), D.VPSDE(), no_sigma=True).eval().to(dev)
tps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
B = 148 * 128 * tps
base, cond = torch.randn(B, 16, device=dev), torch.randn(B, 4, device=dev)
opt = {'step_size': (1.0 - 1e-3) / steps}
sm.sample_ode_from_base(base, cond, method='rk4', options=opt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    sm.sample_ode_from_base(base, cond, method='rk4', options=opt)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
nev = 4 * steps * tps
print(f"{ms:.3f} ms  {ms * 1e-3 * 1.965e9 / nev:.0f} cycles/eval-tile @1.965GHz  {B / ms / 1e3 * 4 * steps:.2f} M row-evals/s  frac {B * 4 * steps * 109568 / (ms * 1e-3) / 276.8e12:.3f}")
