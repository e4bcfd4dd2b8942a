"""Small-shape pass over every kernel family for compute-sanitizer (scripts/gpu_sanitize.sh): the single-tile and dual-tile
engines (dopri5 attempt with both controllers, fixed grids, Euler-Maruyama), the tangent-row engine (exact trace and
Hutchinson), the whole-layer tile engine, the FFMA engine and the staged Hutch++ path.  Shapes are tiny on purpose: the
tools slow kernels down by 10-1000x."""
import sys
import torch
sys.path.insert(0, ".")
import flowfusion_b200.diffusion as D
import flowfusion_b200.flow as F
from flowfusion_b200 import _lib, solver

which = sys.argv[1] if len(sys.argv) > 1 else "all"
lib = _lib.load()
dev = torch.device("cuda:0")
g = lambda s: torch.Generator().manual_seed(s)      # noqa: E731
torch.manual_seed(1234)
sm = D.ScoreModel(D.MLP(16, 4, 8, [128] * 2), D.VPSDE(), no_sigma=True).eval().to(dev)
base, cond = torch.randn(300, 16, generator=g(2)).to(dev), torch.randn(300, 4, generator=g(3)).to(dev)
opts = {"step_t": torch.tensor([1e-3]), "first_step": 0.2}


def section(name, fn):
    if which in ("all", name):
        fn()
        torch.cuda.synchronize()
        print("ok", name, flush=True)


def dopri5(engine, ctl):
    def run():
        lib.ffb_set_engine(engine)
        with solver.controller(ctl):
            sm.sample_ode_from_base(base, cond, atol=1e-3, rtol=1e-3, options=opts)
        lib.ffb_set_engine(1)
    return run


section("rr_host", dopri5(3, "host"))
section("rr_device", dopri5(3, "device"))
section("rd_host", dopri5(4, "host"))
section("rd_device", dopri5(4, "device"))
section("rr_rk4", lambda: sm.sample_ode_from_base(base, cond, method="rk4", options={"step_size": 0.5}))
section("rr_em", lambda: sm.sample_sde((300, 16), cond, steps=3, x0=base, seed=1))
fl = F.ODEFlow(6, [64, 64]).eval().to(dev)
xs = torch.randn(40, 6, generator=g(4)).to(dev)
section("rrt_exact", lambda: fl.log_prob(xs, atol=1e-3, rtol=1e-3))
section("rrt_fixed", lambda: fl.log_prob(xs, method="euler", options={"step_size": 0.5}))
sh = D.ScoreModel(D.MLP(6, 0, 8, [64]), D.VPSDE(), no_sigma=True, hutchinson=True).eval().to(dev)
section("rrt_hutch", lambda: sh.log_prob(xs, atol=1e-3, rtol=1e-3))


def tc_tile():
    lib.ffb_set_engine(2)
    sm.sample_ode_from_base(base, cond, atol=1e-3, rtol=1e-3, options=opts)
    lib.ffb_set_engine(1)


def ffma():
    lib.ffb_set_engine(0)
    sm.sample_ode_from_base(base, cond, atol=1e-3, rtol=1e-3, options=opts)
    lib.ffb_set_engine(1)


section("tc_tile", tc_tile)
section("ffma", ffma)
hp = D.ScoreModel(D.MLP(6, 0, 8, [64]), D.VPSDE(), no_sigma=True, hutchpp=True, hpp_rank=2, hpp_vecs=2).eval().to(dev)
section("staged", lambda: hp.log_prob(xs, atol=1e-2, rtol=1e-2))
print("done", flush=True)
