"""The multi-GPU data plane on hardware (needs 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`):
two NCCL ranks, one per GPU, integrate their blocks of the golden cfg2 batch.  The all-reduced FP64 partial sums must make
both ranks take exactly the steps of the single-process reference run (bit-identical dt history on both ranks), the
gathered samples must match the reference's golden output, and an empty or uneven shard must not unbalance the
collectives -- with the host controller, the device controller and the automatic choice between them."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)                       # noqa: E702
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from conftest import load_golden, rel_row_err
        import flowfusion_b200.diffusion as D
        from flowfusion_b200 import dist as fd, solver
        dev = torch.device("cuda", rank)
        meta, sd, ins, outs = load_golden("cfg2_vp_pfode")
        sm = D.ScoreModel(D.MLP(**meta["ctor"]), D.VPSDE(), no_sigma=True).eval()
        sm.load_state_dict(sd)
        sm.to(dev)
        nb = ins["base"].shape[0]
        opts = {"step_t": torch.tensor([1e-3])}
        out = {}
        splits = {"even": fd.shard_bounds(nb, rank, world), "uneven": (0, 37) if rank == 0 else (37, nb),
                  "empty": (0, nb) if rank == 0 else (nb, nb)}
        for sname, (lo, hi) in splits.items():
            for mode in ("host", "device", "auto"):
                with fd.use_group(td.group.WORLD), solver.controller(mode):
                    x, _ = sm.sample_ode_from_base(ins["base"][lo:hi].to(dev), ins["cond"][lo:hi].to(dev), atol=1e-5,
                                                   rtol=1e-5, options=opts)
                    full = fd.gather_rows(x)                     # a collective right after the solve: must pair up
                st = sm.last_stats
                out[(sname, mode)] = (rel_row_err(outs["x_dopri5"], full), st.accepted, st.rejected, list(st.dt_history),
                                      st.controller)
        out["want"] = (meta["stats"]["accepted"], meta["stats"]["rejected"])
        q.put((rank, out))
    finally:
        td.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_sharded_dopri5_matches_golden_and_ranks_agree():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
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
    want = res[0]["want"]
    for key in res[0]:
        if key == "want":
            continue
        a, b = res[0][key], res[1][key]
        assert a[0] < 1e-4 and b[0] < 1e-4, (key, a[0], b[0])
        assert (a[1], a[2]) == want and (b[1], b[2]) == want, (key, a[1:3], b[1:3], want)
        assert a[3] == b[3], f"{key}: the ranks must take bit-identical step sizes"
        assert a[4] == b[4], f"{key}: the ranks must pick the same controller"
    # the step sequence does not depend on how the rows are split (the sums are FP64 and all-reduced)
    for mode in ("host", "device"):
        assert res[0][("even", mode)][3] == res[0][("empty", mode)][3] or \
            max(abs(x - y) / abs(y) for x, y in zip(res[0][("even", mode)][3], res[0][("empty", mode)][3])) < 1e-9
