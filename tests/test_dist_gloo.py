"""The N > 1 path on CPU: two gloo ranks, each integrating its own block of rows through the package's
host logic (tests/kernel_model.py stands in for the kernels).  Reference-exact dopri5 all-reduces its
float64 partial sums, so both ranks must take exactly the steps the single-process golden run took and
the gathered result must match the golden output."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)                       # noqa: E702
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import load_golden, rel_row_err
        from kernel_model import patched_engine
        import flowfusion_b200.diffusion as D
        import flowfusion_b200.flow as F
        from flowfusion_b200 import dist as fd
        torch.set_num_threads(2)
        out = {}
        # ---- PF-ODE sampling, dopri5 (global RMS norm => identical steps on every rank) ----------------
        meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
        sm = D.ScoreModel(D.MLP(**meta["ctor"]), D.VPSDE(), no_sigma=True).eval()
        sm.load_state_dict(sd)
        base, cond = fd.shard_rows(ins["base"], rank, world), fd.shard_rows(ins["cond"], rank, world)
        with patched_engine(), fd.use_group(td.group.WORLD):
            x, _ = sm.sample_ode_from_base(base, cond, atol=1e-5, rtol=1e-5, options={"step_t": torch.tensor([1e-3])})
            full = fd.gather_rows(x)
        out["pf_err"] = rel_row_err(outs["x_dopri5"], full)
        out["pf_steps"] = (sm.last_stats.accepted, sm.last_stats.rejected, sm.last_stats.nfe)
        out["pf_want"] = (meta["stats"]["accepted"], meta["stats"]["rejected"], meta["stats"]["nfe"])
        out["pf_dt"] = list(sm.last_stats.dt_history)
        # ---- exact-trace log-prob: the log-det column is part of the all-reduced norm -----------------------
        meta, sd, ins, outs = load_golden("cfg3_flow_logprob")
        m = F.ODEFlow(**meta["ctor"], target_shift=sd["target_shift"], target_scale=sd["target_scale"]).eval()
        m.load_state_dict(sd)
        with patched_engine(), fd.use_group(td.group.WORLD):
            lp = m.log_prob(fd.shard_rows(ins["x"], rank, world))
            lp_full = fd.gather_rows(lp)
        out["lp_err"] = float((lp_full - outs["log_prob"]).abs().max())
        out["lp_steps"] = (m.last_stats.accepted, m.last_stats.rejected)
        out["lp_want"] = (meta["stats"]["accepted"], meta["stats"]["rejected"])
        # ---- Hutch++ log-prob (staged attempts): probes are sharded with the rows ------------------------------
        meta, sd, ins, outs = load_golden("score_logprob_hpp_xt_vp")
        fl = meta["flags"]
        sm = D.ScoreModel(D.MLP(**meta["ctor"]), D.VPSDE(), no_sigma=True, hutchpp=True, hpp_rank=fl["hpp_rank"],
                          hpp_vecs=fl["hpp_vecs"]).eval()
        sm.load_state_dict(sd)
        from flowfusion_b200.dist import shard_bounds
        lo, hi = shard_bounds(ins["x0"].shape[0], rank, world)
        with patched_engine(), fd.use_group(td.group.WORLD):
            lp = sm.log_prob(ins["x0"][lo:hi], ins["cond"][lo:hi],
                             probes=(ins["S"][:, lo:hi].contiguous(), ins["G"][:, lo:hi].contiguous()))
            lp_full = fd.gather_rows(lp)
        out["hpp_err"] = float((lp_full - outs["lp_hpp"]).abs().max())
        out["hpp_steps"] = (sm.last_stats.accepted, sm.last_stats.rejected)
        out["hpp_want"] = (meta["stats_hpp"]["accepted"], meta["stats_hpp"]["rejected"])
        # ---- an EMPTY shard on rank 1, every controller: all ranks must run the same sequence of collectives ----------
        # (the device-vs-host controller choice is made from global quantities only; a rank without rows runs the same
        # loop with no-op attempts), and a collective issued right after the solve must pair up correctly
        from flowfusion_b200 import solver
        meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
        sm = D.ScoreModel(D.MLP(**meta["ctor"]), D.VPSDE(), no_sigma=True).eval()
        sm.load_state_dict(sd)
        nb = ins["base"].shape[0]
        lo, hi = (0, nb) if rank == 0 else (nb, nb)
        for mode in ("device", "auto", "host"):
            with patched_engine(), fd.use_group(td.group.WORLD), solver.controller(mode):
                x, _ = sm.sample_ode_from_base(ins["base"][lo:hi], ins["cond"][lo:hi], atol=1e-5, rtol=1e-5,
                                               options={"step_t": torch.tensor([1e-3])})
                full = fd.gather_rows(x)
            out["empty_" + mode] = (rel_row_err(outs["x_dopri5"], full), sm.last_stats.accepted, sm.last_stats.rejected,
                                    sm.last_stats.controller)
        # ---- uneven shards, an empty shard, and the row offsets Philox streams are keyed on -----------------
        t = torch.arange(7, dtype=torch.float32)[:, None]
        mine = fd.shard_rows(t, rank, world)
        out["offset"] = fd.row_offset(mine.shape[0], td.group.WORLD)
        out["gather_ok"] = bool(torch.equal(fd.gather_rows(mine), t))
        empty = t[:0] if rank == 1 else t
        out["gather_empty_ok"] = bool(torch.equal(fd.gather_rows(empty), t))
        q.put((rank, out))
    finally:
        td.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sharded_solves_match_single_process_golden():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=540) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        o = res[r]
        assert o["pf_err"] < 1e-4 and o["lp_err"] < 1e-3
        assert o["pf_steps"] == o["pf_want"], "sharded dopri5 must take the reference's steps"
        assert o["lp_steps"] == o["lp_want"]
        assert o["hpp_err"] < 1e-3 and o["hpp_steps"] == o["hpp_want"]
        assert o["gather_ok"] and o["gather_empty_ok"]
        for mode in ("device", "auto", "host"):
            err, acc, rej, ctl = o["empty_" + mode]
            assert err < 1e-4 and (acc, rej) == o["pf_want"][:2], (mode, o["empty_" + mode])
            assert ctl == res[0]["empty_" + mode][3], "every rank must pick the same controller"
        assert o["empty_device"][3] == "device" and o["empty_host"][3] == "host"
    assert res[0]["pf_dt"] == res[1]["pf_dt"], "every rank must take bit-identical step sizes"
    assert (res[0]["offset"], res[1]["offset"]) == (0, 4)       # 7 rows over 2 ranks: 4 + 3


def test_shard_bounds_cover_every_row_once():
    from flowfusion_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
