"""Debug: per-tile start times (SM cycles and nanoseconds) of CTA 0 over a whole 1 M-row dopri5 attempt
(library built with -DFFB_TRACE -DFFB_TRACE_ROUNDS)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import flowfusion_b200.diffusion as D
from flowfusion_b200 import _lib
lib = _lib.load()
lib.ffb_debug_trace_rd.argtypes = [C.c_void_p]
torch.manual_seed(1234)
dev = torch.device('cuda:0')
sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 4), D.VPSDE(), no_sigma=True).eval().to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
base, cond = torch.randn(B, 16, device=dev), torch.randn(B, 4, device=dev)
opts = {'step_t': torch.tensor([1e-3])}
for _ in range(2):
    sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options=opts)   # warm up (clocks, L2)
CAP = 2048
buf = torch.zeros(5 * CAP * 2, dtype=torch.int64, device=dev)
lib.ffb_debug_trace_rd(C.c_void_p(buf.data_ptr()))
try:
    sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options={'step_t': torch.tensor([1e-3]), 'first_step': 0.05, 'max_num_steps': 1})
except Exception as e:
    print('stopped after one attempt:', e)
torch.cuda.synchronize()
lib.ffb_debug_trace_rd(C.c_void_p(0))
ev = buf.cpu().view(5, CAP, 2)
for role in (0, 2):
    rows = [(t, tag) for t, tag in ev[role].tolist() if t > 0]
    cyc = [t for t, tag in rows if tag == 900]
    ns = [t for t, tag in rows if tag == 950]
    print(f"role {role}: {len(cyc)} tiles")
    for i in range(1, min(len(cyc), len(ns))):
        dc, dn = cyc[i] - cyc[i - 1], ns[i] - ns[i - 1]
        print(f"  tile {i:3d}: {dc:8d} cycles  {dn:8d} ns  -> {dc / max(dn, 1) * 1000:7.1f} MHz   {dc / 6:8.0f} cycles per evaluation")
