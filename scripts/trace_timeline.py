"""Debug: per-layer hand-off timeline of CTA 0 of the tensor-core engine (clock64 deltas)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import flowfusion_b200.diffusion as D
from flowfusion_b200 import _lib
lib = _lib.load()
lib.ffb_debug_trace.argtypes = [C.c_void_p]
torch.manual_seed(1234)
dev = torch.device('cuda:0')
sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 4), D.VPSDE(), no_sigma=True).eval().to(dev)
B = 148 * 128 * 2
base, cond = torch.randn(B, 16, device=dev), torch.randn(B, 4, device=dev)
sm.sample_ode_from_base(base, cond, method='euler', options={'step_size': 0.5})   # warm up
buf = torch.zeros(2 * 4096, dtype=torch.int64, device=dev)
lib.ffb_debug_trace(C.c_void_p(buf.data_ptr()))
sm.sample_ode_from_base(base, cond, method='rk4', options={'step_size': 0.5})
torch.cuda.synchronize()
lib.ffb_debug_trace(C.c_void_p(0))
ev = buf.cpu().view(-1, 2)
ev = ev[ev[:, 0] > 0]
ev = ev[ev[:, 0].argsort()]
t0 = int(ev[0, 0])
names = {0: 'A_READY_SIGNALED', 1: 'mma: saw a_ready', 2: 'mma: committed', 3: 'epi: saw d_ready', 4: 'epi: done', 5: 'last-layer epilogue done (outb)', 6: 'field transform done (eval end)', 7: 'eval begin', 8: 'beff + layer-0 input written'}
prev = t0
for t, tag in ev[20:75].tolist():
    print(f"{t - t0:8d} (+{t - prev:6d})  L{tag % 100}  {names[tag // 100]}")
    prev = t
