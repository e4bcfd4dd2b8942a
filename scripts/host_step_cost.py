"""Host cost per dopri5 step: solve a tiny batch (GPU time negligible) and divide by the number of launches."""
import sys, time, types, torch, cProfile, pstats
sys.path.insert(0, '.')
import bench
import flowfusion_b200.diffusion as D, flowfusion_b200.flow as F, flowfusion_b200.symplectic as Sy
from flowfusion_b200 import _lib
_lib.load()
dev = torch.device('cuda:0')
model = bench.make_model('cfg2', types.SimpleNamespace(D=D, F=F, Sy=Sy)).to(dev)
inp = {k: v.to(dev) for k, v in bench.make_inputs('cfg2', 2048).items()}
for _ in range(3):
    bench.run_gpu('cfg2', model, inp)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 20
for _ in range(n):
    bench.run_gpu('cfg2', model, inp)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
st = model.last_stats
print(f"solve {dt*1e3:.3f} ms, {st.accepted + st.rejected} attempts + 2 evals -> {dt*1e6/(st.accepted + st.rejected + 2):.0f} us per step")
pr = cProfile.Profile(); pr.enable()
for _ in range(n):
    bench.run_gpu('cfg2', model, inp)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumtime').print_stats(28)
