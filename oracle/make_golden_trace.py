"""TEST INFRASTRUCTURE ONLY -- golden vectors for the Hutch++ / XTrace log-prob paths
(`diffusion.py:336-481`, probes drawn at `:703-721`) from the UNMODIFIED reference.

    python oracle/make_golden_trace.py      # needs /root/reference; writes tests/golden/score_logprob_hpp_xt*.npz

Same recipe as oracle/make_golden.py: build the reference object under a fixed seed, let the reference draw its
probes under ``torch.manual_seed`` and replay that draw, check that ``oracle/port.py`` reproduces the reference,
store state_dict + inputs + probes + outputs + solver statistics.

Probe seeds are chosen so that every sample's probe matrix has full column rank: with rank-deficient Rademacher
probes (two equal or opposite columns, frequent at small D) the thin QR's trailing columns are set by rounding
noise and the REFERENCE's own output is not reproducible across BLAS builds (XTrace returns NaN there).
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.loader import load_reference          # noqa: E402
from oracle import port                           # noqa: E402
from oracle.make_golden import gen, report, save, stats_dict, WSEED   # noqa: E402


def full_rank_seed(shapes, start):
    """First seed >= start whose replayed draws all have per-sample smallest singular value > 0.5."""
    seed = start
    while True:
        torch.manual_seed(seed)
        draws = [torch.sign(torch.randn(*s)) for s in shapes]
        if all(torch.linalg.svdvals(d.permute(1, 2, 0)).min() > 0.5 for d in draws[:1]):
            return seed, draws
        seed += 1


def main():
    D, F, S = load_reference()
    torch.set_num_threads(os.cpu_count())
    cases = [
        ("score_logprob_hpp_xt_vp", "vp", True, dict(n_dimensions=16, n_conditionals=2, embedding_dimensions=8, units=[64, 64, 64]), 96,
         dict(hpp_rank=3, hpp_vecs=2, xt_vecs=3)),
        ("score_logprob_hpp_xt_ve", "ve", False, dict(n_dimensions=8, n_conditionals=0, embedding_dimensions=4, units=[32, 32]), 64,
         dict(hpp_rank=1, hpp_vecs=1, xt_vecs=1)),       # the reference's default ranks
    ]
    for name, kind, no_sigma, ctor, B, flags in cases:
        print(name)
        torch.manual_seed(WSEED + 7)
        net = D.MLP(**ctor)
        sde = {"vp": D.VPSDE, "ve": D.VESDE}[kind]()
        Dn, Cn = ctor["n_dimensions"], ctor["n_conditionals"]
        x0 = torch.randn(B, Dn, generator=gen(21)) * (0.5 if kind == "ve" else 1.0)
        cond = torch.randn(B, Cn, generator=gen(22)) if Cn else None
        M = port.score_model_from_state_dict(D.ScoreModel(net, sde, no_sigma=no_sigma).state_dict(), port.make_sde(kind), no_sigma)
        ins, outs, meta_stats = {"x0": x0}, {}, {}
        if cond is not None:
            ins["cond"] = cond
        # Hutch++: S then G are drawn in this order (`:710-711`)
        r, m = min(flags["hpp_rank"], Dn), max(1, flags["hpp_vecs"])
        seed, (Sp, Gp) = full_rank_seed([(r, B, Dn), (m, B, Dn)], 100)
        sm = D.ScoreModel(net, sde, no_sigma=no_sigma, hutchpp=True, hpp_rank=flags["hpp_rank"], hpp_vecs=flags["hpp_vecs"]).eval()
        torch.manual_seed(seed)
        ref = sm.log_prob(x0, cond).detach()
        meta_stats["stats_hpp"] = stats_dict()
        assert torch.equal(sm.S, Sp) and torch.equal(sm.G, Gp)
        report(name, "hutch++", ref, port.score_log_prob(M, x0, cond, probes=("hutchpp", Sp, Gp)), 1e-5)
        ins["S"], ins["G"] = Sp, Gp
        outs["lp_hpp"] = ref
        # XTrace: O (`:721`)
        mx = min(max(1, flags["xt_vecs"]), Dn)
        seed, (Op,) = full_rank_seed([(mx, B, Dn)], 200)
        sm = D.ScoreModel(net, sde, no_sigma=no_sigma, xtrace=True, xt_vecs=flags["xt_vecs"]).eval()
        torch.manual_seed(seed)
        ref = sm.log_prob(x0, cond).detach()
        meta_stats["stats_xt"] = stats_dict()
        assert torch.equal(sm.O, Op)
        report(name, "xtrace", ref, port.score_log_prob(M, x0, cond, probes=("xtrace", Op)), 1e-5)
        ins["O"] = Op
        outs["lp_xt"] = ref
        save(name, dict(case="score_logprob_trace", sde=kind, no_sigma=no_sigma, ctor=ctor, flags=flags,
                        call=dict(atol=1e-4, rtol=1e-4, min_step=1e-6), **meta_stats),
             sm.state_dict(), ins, outs)


if __name__ == "__main__":
    main()
