"""Debug: hand-off timeline of CTA 0 of the dual-tile engine (library built with -DFFB_TRACE, FFB_LIB=build/lib_rd_trace.so).
usage: python scripts/trace_rd.py [dopri5|rk4] [first_row last_row]"""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import flowfusion_b200.diffusion as D
from flowfusion_b200 import _lib
lib = _lib.load()
lib.ffb_debug_trace_rd.argtypes = [C.c_void_p]
torch.manual_seed(1234)
dev = torch.device('cuda:0')
sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 4), D.VPSDE(), no_sigma=True).eval().to(dev)
B = 148 * 128 * 4
base, cond = torch.randn(B, 16, device=dev), torch.randn(B, 4, device=dev)
meth = sys.argv[1] if len(sys.argv) > 1 else 'dopri5'
sm.sample_ode_from_base(base, cond, method='euler', options={'step_size': 0.5})   # warm up
CAP = 2048
buf = torch.zeros(5 * CAP * 2, dtype=torch.int64, device=dev)
lib.ffb_debug_trace_rd(C.c_void_p(buf.data_ptr()))
if meth == 'dopri5':
    try:
        sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options={'step_t': torch.tensor([1e-3]), 'first_step': 0.05, 'max_num_steps': 1})
    except Exception as e:
        print('stopped after one attempt:', e)
else:
    sm.sample_ode_from_base(base, cond, method='rk4', options={'step_size': 0.5})
torch.cuda.synchronize()
lib.ffb_debug_trace_rd(C.c_void_p(0))
ev = buf.cpu().view(5, CAP, 2)
rows = []
for role in range(5):
    for t, tag in ev[role].tolist():
        if t > 0:
            rows.append((t, role, tag))
rows.sort()
t0 = rows[0][0]
RN = {0: 'G0 ', 1: 'MM0', 2: 'G1 ', 3: 'LD ', 4: 'MM1'}
def name(tag):
    if tag == 10: return 'eval: stage input final (qbar passed)'
    if tag == 11: return 'eval: layer-0 operand signalled'
    if tag == 12: return 'eval: last-layer output consumed'
    if tag == 13: return 'eval: end (qbar passed)'
    if 100 <= tag < 200: return f'L{(tag-100)//10} g{(tag-100)%10}: a_ready seen'
    if 300 <= tag < 400: return f'   k0={16*(tag-300)}: weights landed'
    if 400 <= tag < 500: return f'   k0={16*(tag-400)}: MMAs issued + commit'
    if 500 <= tag < 600: return f'L{(tag-500)//10} g{(tag-500)%10}: stage free -> copy issued'
    if tag == 290: return 'last layer: saw d_ready'
    if 200 <= tag < 290:
        l, r = (tag - 200) // 10, (tag - 200) % 10
        return f'L{l} saw d_ready' if r == 0 else f'L{l} epilogue done, a signalled'
    return str(tag)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (400, 900)
prev = {}
for t, role, tag in rows[lo:hi]:
    d = '' if role not in prev else f'+{t - prev[role]:6d}'
    prev[role] = t
    print(f"{t - t0:9d} {RN[role]} {d:>8s}  {'    ' * (0 if role in (0, 2) else 1)}{name(tag)}")
