"""Debug: hand-off timeline of CTA 0 of the row-resident engine (library built with -DFFB_TRACE)."""
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
meth = sys.argv[1] if len(sys.argv) > 1 else 'rk4'
import flowfusion_b200.flow as F
sm.sample_ode_from_base(base, cond, method='euler', options={'step_size': 0.5})   # warm up
CAP = 2048
buf = torch.zeros(3 * CAP * 2, dtype=torch.int64, device=dev)
lib.ffb_debug_trace(C.c_void_p(buf.data_ptr()))
if meth == 'cfg3':
    torch.manual_seed(1234)
    fl = F.ODEFlow(16, [128] * 4).eval().to(dev)
    xs = torch.randn(148 * 7 * 2, 16, device=dev)
    fl.log_prob(xs)
elif meth == 'dopri5':
    try:
        sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options={'step_t': torch.tensor([1e-3]), 'first_step': 0.05, 'max_num_steps': 1})
    except Exception as e:
        print('stopped after one attempt:', e)
else:
    sm.sample_ode_from_base(base, cond, method='rk4', options={'step_size': 0.5})
torch.cuda.synchronize()
lib.ffb_debug_trace(C.c_void_p(0))
ev = buf.cpu().view(3, CAP, 2)
rows = []
for role in range(3):
    for t, tag in ev[role].tolist():
        if t > 0:
            rows.append((t, role, tag))
rows.sort()
t0 = rows[0][0]
def name(tag):
    if 20 <= tag < 30: return {20:'last: ld issued',21:'last: wait_ld done',22:'last: transform stored',23:'buildA: split done',24:'buildA: st issued',25:'signal: wait_st done',26:'kernel: algebra begin',27:'kernel: algebra end',28:'eval: trace summed'}[tag]
    if tag == 10: return 'eval: ycur final (qbar passed)'
    if tag == 11: return 'eval: layer-0 operand handed over'
    if tag == 12: return 'eval: last-layer output consumed'
    if tag == 13: return 'eval: end (qbar passed)'
    if 100 <= tag < 190: return f'MMA  L{(tag-100)//10} chunk {(tag-100)%10} operands ready -> issue'
    if 190 <= tag < 200: return f'MMA  L{tag-190} all issued, commit d_ready'
    if 300 <= tag < 310: return f'MMA  chunk {tag-300} MMAs issued'
    if tag == 290: return 'epi  last layer: saw d_ready'
    if 200 <= tag < 290:
        l, r = (tag - 200) // 10, (tag - 200) % 10
        return f'epi  L{l} saw d_ready' if r == 0 else f'epi  L{l} chunk {r-1} signalled'
    return str(tag)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (150, 330)
prev = {0: None, 1: None, 2: None}
for t, role, tag in rows[lo:hi]:
    d = '' if prev[role] is None else f'+{t - prev[role]:6d}'
    prev[role] = t
    print(f"{t - t0:9d} r{role} {d:>8s}  {name(tag)}")
