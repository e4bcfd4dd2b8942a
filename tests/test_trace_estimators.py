"""Hutch++ / XTrace divergence estimators (reference `diffusion.py:336-481`).

CPU (`-m "not gpu"`):
  * the oracle port vs the golden vectors the UNMODIFIED reference produced (oracle/make_golden_trace.py);
  * the C twin of the estimator kernel (ffb_trace_estimate_host: the same __host__ __device__ statements the
    kernel runs) vs the oracle on seeded Jacobians, all three SDEs, ranks 1..8;
  * the package's host logic (flag priority, probe handling, staged dopri5 attempts) through the torch-CPU kernel
    model + that C twin vs the golden vectors, with identical dopri5 step counts.
GPU (`-m gpu`): the CUDA path through the C ABI vs the golden vectors and vs the oracle, the Jacobian the tangent
engine writes vs autograd, ragged / empty batches, rank limits.

Probes are drawn with full column rank per sample: with equal or opposite Rademacher columns the trailing columns
of the thin QR are rounding noise and the reference itself is not reproducible (XTrace returns NaN).
Tolerance: log-prob <= 1e-3 nat, estimator values <= 1e-4 relative to the batch scale.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden
from kernel_model import patched_engine

import flowfusion_b200.diffusion as D
from flowfusion_b200 import _lib as L
from oracle import port

SDES = {"vp": D.VPSDE, "ve": D.VESDE, "subvp": D.SUBVPSDE}


def full_rank_probes(n, B, Dn, seed):
    g = torch.Generator().manual_seed(seed)
    P = torch.sign(torch.randn(n, B, Dn, generator=g))
    for _ in range(64):
        bad = torch.linalg.svdvals(P.permute(1, 2, 0))[:, -1] < 0.5
        if not bad.any():
            return P
        P[:, bad] = torch.sign(torch.randn(n, int(bad.sum()), Dn, generator=g))
    raise AssertionError("no full-rank probes")


def _model(meta, sd, **flags):
    sm = D.ScoreModel(D.MLP(**meta["ctor"]), SDES[meta["sde"]](), no_sigma=meta["no_sigma"], **flags).eval()
    sm.load_state_dict(sd)
    return sm


# ------------------------------------------------------------------------------------------------------------
# CPU
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["score_logprob_hpp_xt_vp", "score_logprob_hpp_xt_ve"])
def test_port_matches_reference_golden(name):
    meta, sd, ins, outs = load_golden(name)
    M = port.score_model_from_state_dict(sd, port.make_sde(meta["sde"]), meta["no_sigma"])
    cond = ins.get("cond")
    lp = port.score_log_prob(M, ins["x0"], cond, probes=("hutchpp", ins["S"], ins["G"]))
    assert float((lp - outs["lp_hpp"]).abs().max()) < 1e-4
    s = port.last_stats()
    assert (s.accepted, s.rejected) == (meta["stats_hpp"]["accepted"], meta["stats_hpp"]["rejected"])
    lp = port.score_log_prob(M, ins["x0"], cond, probes=("xtrace", ins["O"]))
    assert float((lp - outs["lp_xt"]).abs().max()) < 1e-4


def _twin(jac, kind, S, G, score, use_sigma, has_drift, a, c, sigma, sign=1.0):
    B, Dn = jac.shape[0], jac.shape[1]
    jn = np.ascontiguousarray(jac.numpy(), np.float32)
    Sn = np.ascontiguousarray(S.numpy(), np.float32)
    Gn = None if G is None else np.ascontiguousarray(G.numpy(), np.float32)
    out = np.zeros(B, np.float32)
    t = L.TraceArgs()
    t.batch, t.dim, t.kind, t.rank, t.nvec = B, Dn, kind, S.shape[0], 0 if G is None else G.shape[0]
    t.jac, t.S, t.G = jn.ctypes.data, Sn.ctypes.data, None if Gn is None else Gn.ctypes.data
    t.score, t.use_sigma, t.has_drift = int(score), int(use_sigma), int(has_drift)
    t.a, t.c, t.sigma, t.sign = a, c, sigma, sign
    t.dlp = out.ctypes.data
    L.check(L.load().ffb_trace_estimate_host(C.byref(t)), "ffb_trace_estimate_host")
    return torch.from_numpy(out)


def _port_case(kind, no_sigma, Dn, Cn, seed):
    torch.manual_seed(seed)
    net = D.MLP(Dn, Cn, 8, [48, 48])
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 2:
                p.mul_(2.0)             # a Jacobian that is not dominated by the drift term
    sm = D.ScoreModel(net, SDES[kind](), no_sigma=no_sigma)
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde(kind), no_sigma)
    B = 40
    x = torch.randn(B, Dn)
    cond = torch.randn(B, Cn) if Cn else None
    t = torch.tensor(0.37)

    def netf(xi, ci):
        return port.score_net(M["P"], t, xi[None], None if ci is None else ci[None])[0]

    J = torch.vmap(torch.func.jacrev(netf), in_dims=(0, 0 if Cn else None))(x, cond)       # [b][n][j]
    tt = t * torch.ones(1)
    g = port.sde_diffusion(M["sde"], tt, x[:1]).reshape(-1)[0]
    scal = dict(a=0.0 if kind == "ve" else float(-0.5 * port.sde_beta(M["sde"], tt)[0]), c=float(0.5 * g ** 2),
                sigma=float(port.sde_sigma(M["sde"], tt)[0]))
    return M, t, x, cond, J.permute(0, 2, 1).contiguous(), scal


@pytest.mark.parametrize("kind,no_sigma,Dn,Cn", [("vp", True, 8, 2), ("ve", False, 12, 0), ("subvp", False, 16, 0),
                                                 ("vp", False, 32, 1), ("ve", True, 5, 0), ("vp", True, 48, 0), ("ve", False, 100, 2)])
def test_c_twin_of_the_kernel_matches_the_oracle(kind, no_sigma, Dn, Cn):
    M, t, x, cond, jac, scal = _port_case(kind, no_sigma, Dn, Cn, 11)
    B = x.shape[0]
    for r, m in ((1, 1), (3, 2), (min(Dn, 8), 5)):
        S, G = full_rank_probes(r, B, Dn, 100 + r), torch.sign(torch.randn(m, B, Dn))
        ref = port.score_field(M, t, (x,), cond, True, ("hutchpp", S, G))[1].reshape(-1)
        got = _twin(jac, L.TRACE_HUTCHPP, S, G, True, not no_sigma, kind != "ve", **scal)
        assert float((ref - got).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), ("hutch++", r, m)
        if r < Dn:      # r = D = m makes R's conditioning arbitrary
            ref = port.score_field(M, t, (x,), cond, True, ("xtrace", S))[1].reshape(-1)
            got = _twin(jac, L.TRACE_XTRACE, S, None, True, not no_sigma, kind != "ve", **scal)
            assert float((ref - got).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max())), ("xtrace", r)
    # reversed time negates the estimate; a plain network field (no score transform) is J_net itself
    S, G = full_rank_probes(2, B, Dn, 7), torch.sign(torch.randn(1, B, Dn))
    pos = _twin(jac, L.TRACE_HUTCHPP, S, G, True, not no_sigma, kind != "ve", **scal)
    neg = _twin(jac, L.TRACE_HUTCHPP, S, G, True, not no_sigma, kind != "ve", sign=-1.0, **scal)
    assert torch.equal(pos, -neg)
    full = _twin(jac, L.TRACE_HUTCHPP, full_rank_probes(min(Dn, 8), B, Dn, 9), G, False, False, False, 0.0, 0.0, 1.0)
    if Dn <= 8:         # Q spans everything: Hutch++ is the exact trace
        assert float((full - torch.diagonal(jac, 0, 1, 2).sum(-1)).abs().max()) < 1e-4


def test_c_twin_rejects_bad_arguments():
    lib = L.load()
    t = L.TraceArgs()
    assert lib.ffb_trace_estimate_host(C.byref(t)) != 0
    buf = np.zeros(64, np.float32)
    t.batch, t.dim, t.kind, t.rank, t.nvec = 1, 4, L.TRACE_HUTCHPP, 5, 1
    t.jac = t.S = t.G = t.dlp = buf.ctypes.data
    assert lib.ffb_trace_estimate_host(C.byref(t)) != 0 and b"rank" in lib.ffb_last_error()
    t.rank, t.dim = 1, L.TRACE_MAX_DIM + 1
    assert lib.ffb_trace_estimate_host(C.byref(t)) != 0
    t.dim, t.nvec = 4, 0
    assert lib.ffb_trace_estimate_host(C.byref(t)) != 0          # Hutch++ needs G
    t.kind = L.TRACE_XTRACE
    assert lib.ffb_trace_estimate_host(C.byref(t)) == 0


@pytest.mark.parametrize("name", ["score_logprob_hpp_xt_vp", "score_logprob_hpp_xt_ve"])
def test_host_logic_reproduces_reference_golden(name):
    meta, sd, ins, outs = load_golden(name)
    fl = meta["flags"]
    cond = ins.get("cond")
    with patched_engine():
        sm = _model(meta, sd, hutchpp=True, hpp_rank=fl["hpp_rank"], hpp_vecs=fl["hpp_vecs"])
        lp = sm.log_prob(ins["x0"], cond, probes=(ins["S"], ins["G"]))
        assert lp.shape == outs["lp_hpp"].shape
        assert float((lp - outs["lp_hpp"]).abs().max()) < 1e-3
        assert (sm.last_stats.accepted, sm.last_stats.rejected, sm.last_stats.nfe) == \
            (meta["stats_hpp"]["accepted"], meta["stats_hpp"]["rejected"], meta["stats_hpp"]["nfe"])
        # forward(t, states) reuses the stored probes of the solve (`diffusion.py:346-354`)
        sm.prob, sm.conditional = True, cond
        f, d = sm.forward(torch.tensor(0.5), (ins["x0"], torch.zeros(ins["x0"].shape[0], 1)))
        M = port.score_model_from_state_dict(sd, port.make_sde(meta["sde"]), meta["no_sigma"])
        rf, rd = port.score_field(M, torch.tensor(0.5), (ins["x0"],), cond, True, ("hutchpp", ins["S"], ins["G"]))
        assert float((f - rf).abs().max()) < 1e-4 and float((d - rd).abs().max()) < 1e-3
        sm = _model(meta, sd, xtrace=True, xt_vecs=fl["xt_vecs"])
        lp = sm.log_prob(ins["x0"], cond, probes=ins["O"])
        assert float((lp - outs["lp_xt"]).abs().max()) < 1e-3
        assert (sm.last_stats.accepted, sm.last_stats.rejected) == (meta["stats_xt"]["accepted"], meta["stats_xt"]["rejected"])


@pytest.mark.parametrize("method,h", [("euler", 1 / 64), ("midpoint", 1 / 32), ("rk4", 1 / 16)])
def test_host_logic_fixed_grid_methods_vs_oracle(method, h):
    meta, sd, ins, outs = load_golden("score_logprob_hpp_xt_vp")
    fl, cond = meta["flags"], ins["cond"]
    M = port.score_model_from_state_dict(sd, port.make_sde(meta["sde"]), meta["no_sigma"])
    opts = {"step_size": h}
    rx, rl = port.solve_odes_forward(M, ins["x0"], cond, method=method, options=opts, probes=("hutchpp", ins["S"], ins["G"]))
    with patched_engine():
        sm = _model(meta, sd, hutchpp=True, hpp_rank=fl["hpp_rank"], hpp_vecs=fl["hpp_vecs"])
        x, lp = sm.solve_odes_forward(ins["x0"], cond, method=method, options=opts, probes=(ins["S"], ins["G"]))
    assert lp.shape == rl.shape
    assert float((x - rx).abs().max()) <= 1e-4 * max(1.0, float(rx.abs().max()))
    assert float((lp - rl).abs().max()) < 1e-3


def test_flag_priority_probe_shapes_and_refusals():
    meta, sd, ins, outs = load_golden("score_logprob_hpp_xt_ve")
    x0 = ins["x0"]
    B, Dn = x0.shape
    with patched_engine():
        # hutchinson wins over hutchpp, hutchpp over xtrace (`diffusion.py:327, 336, 402`)
        sm = _model(meta, sd, hutchinson=True, hutchpp=True, xtrace=True)
        assert sm._estimator(x0) is None
        sm = _model(meta, sd, hutchpp=True, xtrace=True, hpp_rank=100, hpp_vecs=0)
        est = sm._estimator(x0)
        assert est.kind == L.TRACE_HUTCHPP and sm.S.shape == (Dn, B, Dn) and sm.G.shape == (1, B, Dn)   # r = min(rank, D), m = max(1, vecs)
        assert set(sm.S.unique().tolist()) <= {-1.0, 1.0}
        sm = _model(meta, sd, xtrace=True, xt_vecs=3)
        assert sm._estimator(x0).kind == L.TRACE_XTRACE and sm.O.shape == (3, B, Dn)
        O = sm.O
        assert sm._estimator(x0, stored=True).S is O                 # forward() keeps the solve's probes
        assert sm._estimator(x0[:5], stored=True).S is not O         # ... unless the shape changed (`:417`)
        with pytest.raises(ValueError):
            sm.solve_odes_forward(x0, probes=torch.ones(2, B, Dn))
        with pytest.raises(NotImplementedError):
            sm.solve_odes_forward(x0, method="bosh3")
        sm = _model(meta, sd, hutchpp=True, hpp_rank=1)
        sm.train()
        with pytest.raises(NotImplementedError):
            sm.solve_odes_forward(x0)
    wide = D.ScoreModel(D.MLP(126, 0, 4, [32]), D.VPSDE(), hutchpp=True).eval()     # a sample + its tangents must share one tile
    with patched_engine(), pytest.raises(NotImplementedError):
        wide.solve_odes_forward(torch.zeros(4, 126))


# ------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["score_logprob_hpp_xt_vp", "score_logprob_hpp_xt_ve"])
def test_gpu_logprob_matches_reference_golden(name, cuda_dev):
    meta, sd, ins, outs = load_golden(name)
    fl = meta["flags"]
    cond = None if "cond" not in ins else ins["cond"].to(cuda_dev)
    x0 = ins["x0"].to(cuda_dev)
    n0 = L.launch_count()
    sm = _model(meta, sd, hutchpp=True, hpp_rank=fl["hpp_rank"], hpp_vecs=fl["hpp_vecs"]).to(cuda_dev)
    lp = sm.log_prob(x0, cond, probes=(ins["S"].to(cuda_dev), ins["G"].to(cuda_dev)))
    assert lp.shape == outs["lp_hpp"].shape and L.launch_count() > n0
    assert float((lp.cpu() - outs["lp_hpp"]).abs().max()) < 1e-3
    assert (sm.last_stats.accepted, sm.last_stats.rejected, sm.last_stats.nfe) == \
        (meta["stats_hpp"]["accepted"], meta["stats_hpp"]["rejected"], meta["stats_hpp"]["nfe"])
    sm = _model(meta, sd, xtrace=True, xt_vecs=fl["xt_vecs"]).to(cuda_dev)
    lp = sm.log_prob(x0, cond, probes=ins["O"].to(cuda_dev))
    assert float((lp.cpu() - outs["lp_xt"]).abs().max()) < 1e-3
    assert (sm.last_stats.accepted, sm.last_stats.rejected) == (meta["stats_xt"]["accepted"], meta["stats_xt"]["rejected"])


@pytest.mark.gpu
@pytest.mark.parametrize("kind,no_sigma,Dn,Cn,B", [("vp", True, 16, 4, 1000), ("subvp", False, 8, 0, 129), ("ve", False, 32, 2, 77),
                                                   ("vp", True, 5, 0, 64), ("vp", True, 48, 0, 40), ("ve", False, 100, 2, 20),
                                                   ("vp", True, 124, 0, 9)])
def test_gpu_jacobian_and_estimators_vs_oracle(kind, no_sigma, Dn, Cn, B, cuda_dev):
    """One evaluation: the Jacobian the tangent engine writes vs autograd, both estimators vs the oracle."""
    from flowfusion_b200 import engine as E
    torch.manual_seed(5)
    net = D.MLP(Dn, Cn, 8, [128, 128])
    x = torch.randn(B, Dn)
    cond = torch.randn(B, Cn) if Cn else None
    r = min(3, Dn)
    S, G = full_rank_probes(r, B, Dn, 31), torch.sign(torch.randn(2, B, Dn))
    sm = D.ScoreModel(net, SDES[kind](), no_sigma=no_sigma, hutchpp=True, hpp_rank=r, hpp_vecs=2).eval()
    M = port.score_model_from_state_dict(sm.state_dict(), port.make_sde(kind), no_sigma)
    t = torch.tensor(0.61)
    rf, rd = port.score_field(M, t, (x,), cond, True, ("hutchpp", S, G))
    rx = port.score_field(M, t, (x,), cond, True, ("xtrace", S))[1]
    sm.to(cuda_dev)
    sm.prob, sm.conditional = True, None if cond is None else cond.to(cuda_dev)
    sm.S, sm.G = S.to(cuda_dev), G.to(cuda_dev)
    f, d = sm.forward(t, (x.to(cuda_dev), torch.zeros(B, 1, device=cuda_dev)))
    scale = max(1.0, float(rd.abs().max()))
    assert float((f.cpu() - rf).abs().max()) <= 1e-4 * max(1.0, float(rf.abs().max()))
    assert float((d.cpu() - rd).abs().max()) <= 2e-4 * scale
    sm.hutchpp, sm.xtrace, sm.xt_vector, sm.O = False, True, r, S.to(cuda_dev)
    _, dx = sm.forward(t, (x.to(cuda_dev), torch.zeros(B, 1, device=cuda_dev)))
    assert float((dx.cpu() - rx).abs().max()) <= 3e-4 * scale
    # the Jacobian buffer itself
    be = E.StagedBackend(sm._field(L.DIV_EXACT), x.to(cuda_dev), E.TraceEstimator(L.TRACE_XTRACE, S.to(cuda_dev)),
                         cond=sm.conditional)
    row = sm._program()(np.array([0.61], np.float32))[0]
    be.single_eval(row)

    def netf(xi, ci):
        return port.score_net(M["P"], t, xi[None], None if ci is None else ci[None])[0]

    J = torch.vmap(torch.func.jacrev(netf), in_dims=(0, 0 if Cn else None))(x, cond)       # [b][n][j]
    got = be.jac.cpu()
    assert float((got - J.permute(0, 2, 1)).abs().max()) <= 1e-4 * max(1.0, float(J.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("method,h", [("euler", 1 / 64), ("midpoint", 1 / 32), ("rk4", 1 / 16)])
def test_gpu_fixed_grid_methods_vs_oracle(method, h, cuda_dev):
    meta, sd, ins, outs = load_golden("score_logprob_hpp_xt_vp")
    fl, cond = meta["flags"], ins["cond"]
    M = port.score_model_from_state_dict(sd, port.make_sde(meta["sde"]), meta["no_sigma"])
    opts = {"step_size": h}
    rx, rl = port.solve_odes_forward(M, ins["x0"], cond, method=method, options=opts, probes=("xtrace", ins["O"]))
    sm = _model(meta, sd, xtrace=True, xt_vecs=fl["xt_vecs"]).to(cuda_dev)
    x, lp = sm.solve_odes_forward(ins["x0"].to(cuda_dev), cond.to(cuda_dev), method=method, options=opts, probes=ins["O"].to(cuda_dev))
    assert float((x.cpu() - rx).abs().max()) <= 1e-4 * max(1.0, float(rx.abs().max()))
    assert float((lp.cpu() - rl).abs().max()) < 1e-3


@pytest.mark.gpu
def test_gpu_staged_solve_ragged_empty_and_partition_invariance(cuda_dev):
    torch.manual_seed(3)
    sm = D.ScoreModel(D.MLP(16, 0, 8, [128] * 2), D.VPSDE(), no_sigma=True, hutchpp=True, hpp_rank=2, hpp_vecs=2).eval().to(cuda_dev)
    B = 300
    x = torch.randn(B, 16, device=cuda_dev)
    S, G = full_rank_probes(2, B, 16, 41).to(cuda_dev), torch.sign(torch.randn(2, B, 16)).to(cuda_dev)
    lp = sm.log_prob(x, probes=(S, G), options={"first_step": 0.05, "min_step": 1e-6})
    assert lp.shape == (B, 1) and torch.isfinite(lp).all()
    # a sample's result does not depend on the batch it rides in once the step sequence is pinned (fixed grid)
    opts = {"step_size": 0.125}
    for method in ("rk4", "midpoint"):
        a = sm.log_prob(x[:130], probes=(S[:, :130].contiguous(), G[:, :130].contiguous()), method=method, options=opts)
        b = sm.log_prob(x, probes=(S, G), method=method, options=opts)
        assert torch.equal(a, b[:130])
    e = sm.log_prob(x[:0], probes=(S[:, :0].contiguous(), G[:, :0].contiguous()), method="euler", options={"step_size": 0.25})
    assert e.shape == (0, 1)
    with pytest.raises(NotImplementedError):
        D.ScoreModel(D.MLP(16, 0, 8, [64]), D.VPSDE(), xtrace=True, xt_vecs=9).eval().to(cuda_dev).log_prob(x)
