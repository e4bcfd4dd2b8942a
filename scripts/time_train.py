"""Timing of the fused training step (csrc/ffb_train.cu): flow-matching loss + backward of the cfg3 network (ODEFlow(16, [128]*4))
and denoising score matching of the cfg2 network, per batch size, against the same loss through torch autograd (eager, FP32, TF32 off)
on the same GPU.  Prints one JSON object.
usage: python scripts/time_train.py"""
import json
import sys

import torch

sys.path.insert(0, '.')
import flowfusion_b200.diffusion as D
import flowfusion_b200.flow as F
from flowfusion_b200 import _lib

_lib.load()
dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"runs": []}
for B in (1024, 8192, 65536):
    torch.manual_seed(1234)
    m = F.ODEFlow(16, [128] * 4).train().to(dev)
    x = torch.randn(B, 16, device=dev)

    def fused():
        m.zero_grad(set_to_none=True)
        m.flow_matching_loss(x).backward()

    def eager():
        m.zero_grad(set_to_none=True)
        xT = torch.randn_like(x); t = torch.rand(B, 1, device=dev)
        x0 = (x - m.target_shift) / m.target_scale
        xt = (1 - t) * x0 + t * xT
        v = m.velocity(torch.cat([xt, t], dim=1))
        torch.mean((v - (xT - x0)) ** 2).backward()

    n0 = _lib.launch_count()
    fused()
    launches = _lib.launch_count() - n0
    tf, te = timed(fused), timed(eager)
    flop = 3 * 2 * B * (17 * 128 + 3 * 128 * 128 + 128 * 16)       # forward + dX sweep + dW
    out["runs"].append({"loss": "flow_matching cfg3 net", "batch": B, "fused_ms": tf, "torch_autograd_ms": te, "speedup": te / tf,
                        "own_launches_per_step": launches, "fused_tflops": flop / (tf * 1e-3) / 1e12})
    sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 4), D.VPSDE(), no_sigma=True).train().to(dev)
    c = torch.randn(B, 4, device=dev)

    def fused_dsm():
        sm.zero_grad(set_to_none=True)
        sm.loss_fn(x, conditional=c).backward()

    out["runs"].append({"loss": "denoising score matching cfg2 net", "batch": B, "fused_ms": timed(fused_dsm)})
print(json.dumps(out))
