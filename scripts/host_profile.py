"""Where does the host spend time inside one cfg2 solve?  (cProfile around 3 solves after warm-up)"""
import cProfile, pstats, sys, types, torch
sys.path.insert(0, '.')
import bench
import flowfusion_b200.diffusion as D, flowfusion_b200.flow as F, flowfusion_b200.symplectic as Sy
from flowfusion_b200 import _lib
_lib.load()
dev = torch.device('cuda:0')
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
model = bench.make_model(name, types.SimpleNamespace(D=D, F=F, Sy=Sy)).to(dev)
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[name]['B']
inp = {k: v.to(dev) for k, v in bench.make_inputs(name, B).items()}
for _ in range(2):
    bench.run_gpu(name, model, inp)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    bench.run_gpu(name, model, inp)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
