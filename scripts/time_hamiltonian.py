"""Timing of the fused Hamiltonian leapfrog (csrc/ffb_train.cu: k_ham_leapfrog): BASELINE.json configs[4] shape -- 64-D phase space
(D = 32), scalar Hamiltonian MLP 64 -> 128 x 4 -> 1 -- kick-drift-kick steps with the fused forward+backward gradient, as a fraction
of the measured FFMA2 peak, beside the reference's direct-output symplectic networks on the tensor-core engine for scale.
Prints one JSON object.    usage: python scripts/time_hamiltonian.py [rows] [steps]"""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, '.')
import flowfusion_b200.symplectic as Sy
from flowfusion_b200 import _lib

lib = _lib.load()
dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 32 * 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 25
peak = C.c_float()
with torch.cuda.device(dev):
    lib.ffb_ffma_peak(20000, C.byref(peak), None)
torch.manual_seed(1234)
m = Sy.HamiltonianMLP(32, 0, [128] * 4).to(dev)
z0 = torch.randn(B, 64, device=dev)


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timed(lambda: m.leapfrog(z0, num_steps=steps, dt=0.01))
mac = 64 * 128 + 3 * 128 * 128 + 128
flop = B * (3 * steps + 1) * 4 * mac                    # forward + backward sweep to the inputs, 2 FLOP per MAC each
out = {"rows": B, "steps": steps, "gradient_evaluations_per_step": 3, "ms": ms, "trajectory_steps_per_s": B * steps / (ms * 1e-3),
       "tflops": flop / (ms * 1e-3) / 1e12, "fp32_ffma2_peak_tflops": peak.value, "frac_of_ffma2_peak": flop / (ms * 1e-3) / 1e12 / peak.value}
net = Sy.SymplecticMLP(32, 0, 8, [128] * 4)
sm = Sy.SymplecticFlowModel(net, torch.zeros(32), torch.ones(32), torch.zeros(1), torch.ones(1)).to(dev)
ms2 = timed(lambda: sm.sample((B, 32), num_steps=steps, z0=z0, method="leapfrog"))
out["reference_style_direct_output_nets_leapfrog"] = {"ms": ms2, "trajectory_steps_per_s": B * steps / (ms2 * 1e-3)}
print(json.dumps(out))
