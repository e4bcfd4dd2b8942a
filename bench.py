#!/usr/bin/env python
"""bench.py -- headline benchmark of the flowfusion hot path on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

A "step" is one complete pass of the hot path over one batch of synthetic inputs.  Default
workload = BASELINE.json configs[1] ("cfg2"): 16-D conditional VP-SDE
diffusion (4x128 MLP), probability-flow ODE sampling of 1M samples with dopri5
(atol = rtol = 1e-5, options={'step_t': [eps]}) through ``ScoreModel.sample_ode_from_base``.
Other workloads (cfg3 exact-trace log_prob, cfg4 Euler-Maruyama, cfg5 symplectic, cfg1) are
selectable for the profiles; they are parity-test cases, not the bench line.

N > 1 (launched by torchrun, one rank per GPU): weak scaling -- every rank integrates its own
batch of the same size; dopri5 all-reduces its error-norm partial sums (NCCL) so all ranks take
the same steps.  Time = max over ranks of the CUDA-event time of exactly K steps.

``--impl reference`` times the CPU oracle port (oracle/port.py: the reference's algorithm on the
restated torchdiffeq) on the box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WSEED = 1234


# ----------------------------------------------------------------------------------------------
# workloads: build(module namespace) -> model; inputs(B) -> dict of CPU tensors
# ----------------------------------------------------------------------------------------------
def _gen(seed):
    return torch.Generator().manual_seed(seed)


WORKLOADS = {
    "cfg1": dict(desc="ODEFlow(2,[64]*3).sample, dopri5 torchdiffeq defaults (rtol 1e-7, atol 1e-9)", B=10_000,
                 cpu_B=10_000, metric="samples/s", unit="samples/s"),
    "cfg2": dict(desc="MLP(16,4,8,[128]*4)+VPSDE no_sigma: PF-ODE sampling, dopri5 atol=rtol=1e-5, step_t=[eps]",
                 B=1_000_000, cpu_B=400_000, metric="samples/s", unit="samples/s"),
    "cfg3": dict(desc="ODEFlow(16,[128]*4).log_prob, exact divergence trace, dopri5 atol=rtol=1e-5", B=4_000_000,
                 cpu_B=20_000, metric="log_prob evals/s", unit="evals/s"),
    "cfg4": dict(B_strong=10_000_000,
                 desc="MLP(32,0,8,[128]*4)+VPSDE no_sigma: reverse-SDE Euler-Maruyama, 1000 steps, in-kernel Philox",
                 B=1_250_000, cpu_B=20_000, metric="samples/s", unit="samples/s"),
    "cfg5": dict(desc="SymplecticMLP(32,0,8,[128]*4): forward-Euler sampling, 100 steps, 64-D phase space",
                 B=4_000_000, cpu_B=200_000, metric="samples/s", unit="samples/s"),
    # not a BASELINE config: the cfg2 workload with a 256-wide network, which runs on the wide engine (FP32 pipe,
    # csrc/ffb_engine_wide.cuh) -- read `roofline.frac_of_ffma2_peak`, not `frac` (that one is against the tensor peak)
    "wide256": dict(desc="MLP(16,4,8,[256]*4)+VPSDE no_sigma: PF-ODE sampling, dopri5 atol=rtol=1e-5, step_t=[eps]; wide engine (FP32 pipe)",
                    B=250_000, cpu_B=100_000, metric="samples/s", unit="samples/s"),
}


def make_model(name, ns):
    """ns: object with .D/.F/.Sy modules exposing the reference class names."""
    torch.manual_seed(WSEED)
    if name == "cfg1":
        return ns.F.ODEFlow(2, [64, 64, 64]).eval()
    if name == "cfg2":
        return ns.D.ScoreModel(ns.D.MLP(16, 4, 8, [128] * 4), ns.D.VPSDE(), no_sigma=True).eval()
    if name == "wide256":
        return ns.D.ScoreModel(ns.D.MLP(16, 4, 8, [256] * 4), ns.D.VPSDE(), no_sigma=True).eval()
    if name == "cfg3":
        return ns.F.ODEFlow(16, [128] * 4).eval()
    if name == "cfg4":
        return ns.D.ScoreModel(ns.D.MLP(32, 0, 8, [128] * 4), ns.D.VPSDE(), no_sigma=True).eval()
    if name == "cfg5":
        net = ns.Sy.SymplecticMLP(32, 0, 8, [128] * 4)
        return ns.Sy.SymplecticFlowModel(net, torch.zeros(32), torch.ones(32), torch.zeros(0), torch.ones(0)).eval()
    raise KeyError(name)


def make_inputs(name, B, rank=0):
    s = 1000 * rank
    if name == "cfg1":
        return {"xT": torch.randn(B, 2, generator=_gen(1 + s))}
    if name in ("cfg2", "wide256"):
        return {"base": torch.randn(B, 16, generator=_gen(2 + s)), "cond": torch.randn(B, 4, generator=_gen(3 + s))}
    if name == "cfg3":
        return {"x": torch.randn(B, 16, generator=_gen(4 + s))}
    if name == "cfg4":
        return {"x0": torch.randn(B, 32, generator=_gen(6 + s))}
    if name == "cfg5":
        return {"z0": torch.randn(B, 64, generator=_gen(8 + s))}
    raise KeyError(name)


def run_gpu(name, model, inp):
    """One step through the package's public API; returns the result tensor (on device)."""
    if name == "cfg1":
        return model.sample(inp["xT"])
    if name in ("cfg2", "wide256"):
        return model.sample_ode_from_base(inp["base"], inp["cond"], atol=1e-5, rtol=1e-5,
                                          options={"step_t": torch.tensor([1e-3])})[0]
    if name == "cfg3":
        return model.log_prob(inp["x"])
    if name == "cfg4":
        return model.sample_sde(tuple(inp["x0"].shape), steps=1000, x0=inp["x0"], seed=5)
    if name == "cfg5":
        return model.sample((inp["z0"].shape[0], 32), num_steps=100, z0=inp["z0"])
    raise KeyError(name)


def run_cpu_port(name, model, inp):
    """The same call on the CPU oracle port (weights taken from the model's state_dict)."""
    from oracle import port
    sd = model.state_dict()
    if name == "cfg1":
        return port.flow_sample(port.flow_from_state_dict(sd), inp["xT"])
    if name in ("cfg2", "wide256"):
        M = port.score_model_from_state_dict(sd, port.make_sde("vp"), True)
        return port.sample_ode_from_base(M, inp["base"], inp["cond"], 1e-5, 1e-5, options={"step_t": torch.tensor([1e-3])})[0]
    if name == "cfg3":
        return port.flow_log_prob(port.flow_from_state_dict(sd), inp["x"])
    if name == "cfg4":
        M = port.score_model_from_state_dict(sd, port.make_sde("vp"), True)
        dw = torch.randn(1000, *inp["x0"].shape, generator=_gen(7))
        return port.sample_sde(M, inp["x0"], dw)
    if name == "cfg5":
        return port.symplectic_sample(port.symplectic_from_state_dict(sd), inp["z0"], None, 100)
    raise KeyError(name)


def nfe_of(name, model):
    if name == "cfg4":
        return 1000
    if name == "cfg5":
        return 100
    return model.last_stats.nfe if hasattr(model, "last_stats") and model.last_stats else None


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            try:
                pw.append(float(r[2]))
            except Exception:
                pass
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "sm_mhz_min": min(sm), "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def cpu_baseline(name, cpu_B):
    import types
    import flowfusion_b200.diffusion as D, flowfusion_b200.flow as F, flowfusion_b200.symplectic as Sy  # noqa: E401
    torch.set_num_threads(os.cpu_count())
    model = make_model(name, types.SimpleNamespace(D=D, F=F, Sy=Sy))
    inp = make_inputs(name, cpu_B)
    t0 = time.perf_counter()
    run_cpu_port(name, model, inp)
    dt = time.perf_counter() - t0
    return {"value": cpu_B / dt, "unit": WORKLOADS[name]["unit"], "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{cpu_B} rows of the same workload, one call of the oracle port (oracle/port.py), {dt:.1f} s"}


def reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import types
    import flowfusion_b200.diffusion as D, flowfusion_b200.flow as F, flowfusion_b200.symplectic as Sy  # noqa: E401
    name = args.workload
    w = WORKLOADS[name]
    torch.set_num_threads(os.cpu_count())
    model = make_model(name, types.SimpleNamespace(D=D, F=F, Sy=Sy))
    B = max(1000, min(w["cpu_B"], int(w["cpu_B"] * args.ref_scale)))
    inp = make_inputs(name, B)
    for _ in range(args.warmup):
        run_cpu_port(name, model, inp)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_cpu_port(name, model, inp)
    dt = (time.perf_counter() - t0) / args.steps
    val = B / dt
    line = {"impl": "reference", "metric": w["metric"], "value": val, "unit": w["unit"], "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{name}: {w['desc']}", "rows_per_step": B,
                       "note": "CPU oracle port of the reference algorithm on the restated torchdiffeq; bounded sample"},
            "cpu_baseline": {"value": val, "unit": w["unit"], "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{B} rows per step"},
            "e2e": {"value": val, "unit": w["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore its pinned-host allocations, first touch) to the CPUs of the NUMA node the GPU hangs
    off: with 8 ranks on one box the host <-> device copies of the e2e leg otherwise cross the socket interconnect."""
    try:
        bus = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus                      # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return {"numa_node": node, "note": "the kernel reports no NUMA affinity for this GPU: nothing to bind"}
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"[:160]}
    return None


def tf32_gemm_tflops(dev):
    """Measured dense TF32 GEMM rate (cuBLAS through torch.matmul, 8192^3, best of 5): the tensor pipe's own ceiling for
    one of the three products of a 3xTF32 MAC."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev)
        best = 0.0
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        return best
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def shard_check(name, model, inp, out, rows=4096):
    """After the timed region: the first `rows` rows of this rank's result against the CPU oracle (oracle/port.py) driven
    with the step sequence the GPU solve recorded -- at N > 1 those steps came from the all-reduced global error norm, so
    this is the data-plane check of the sharded solve.  Returns the largest per-row relative deviation."""
    from oracle import port
    st = getattr(model, "last_stats", None)
    n = min(rows, out.shape[0])
    if n == 0:
        return None
    sub = {k: v[:n].cpu() for k, v in inp.items()}
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    replay = None if st is None else (list(st.dt_history), list(st.accept_history))
    if name in ("cfg2", "wide256"):
        M = port.score_model_from_state_dict(sd, port.make_sde("vp"), True)
        ref = port.sample_ode_from_base(M, sub["base"], sub["cond"], 1e-5, 1e-5,
                                        options={"step_t": torch.tensor([1e-3]), "_replay": replay})[0]
    elif name == "cfg4":
        return None                                   # in-kernel Philox noise: covered by the golden-vector tests
    elif name == "cfg5":
        ref = port.symplectic_sample(port.symplectic_from_state_dict(sd), sub["z0"], None, 100)
    else:
        return None                                   # cfg1 / cfg3: the solve's first step size comes from a global norm too
    got = out[:n].float().cpu()
    num = (ref - got).abs().reshape(n, -1).amax(dim=1)
    den = ref.abs().reshape(n, -1).amax(dim=1).clamp(min=1.0)
    return {"rows": n, "max_rel_err": float((num / den).max()), "oracle": "oracle/port.py replaying the recorded step sequence"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="rows per GPU (weak) / in total (strong); default: the workload's size")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank integrates its own batch; strong: ONE batch of the workload's size is sharded over the ranks")
    ap.add_argument("--no-shard-check", action="store_true", help="skip the oracle check of a 4096-row slice after the timed region")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-scale", type=float, default=0.5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import types
    import torch.distributed as td
    import flowfusion_b200.diffusion as D, flowfusion_b200.flow as F, flowfusion_b200.symplectic as Sy  # noqa: E401
    from flowfusion_b200 import _lib, engine, dist as fdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
        group = td.group.WORLD
    _lib.load()
    name = args.workload
    w = WORKLOADS[name]
    numa = bind_to_gpu_numa_node(local) if world > 1 else None        # before the pinned buffers are allocated
    if args.scaling == "strong":          # ONE batch of the BASELINE.json size (cfg3: 4 M points, cfg4: 10 M samples) over all ranks
        lo, hi = fdist.shard_bounds(args.batch or w.get("B_strong", w["B"]), rank, world)
        B = hi - lo
    else:
        B = args.batch or w["B"]
    model = make_model(name, types.SimpleNamespace(D=D, F=F, Sy=Sy)).to(dev)
    host = {k: v.pin_memory() for k, v in make_inputs(name, B, rank).items()}
    inp = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    ctx = fdist.use_group(group)
    with ctx:
        for _ in range(args.warmup):
            out = run_gpu(name, model, inp)
        barrier()
        # ---- timed region: exactly K steps, device-resident inputs -------------------------------
        engine.profiler.reset(True)
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            barrier()
            e0.record()
            for _ in range(args.steps):
                out = run_gpu(name, model, inp)
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - n0
        prof = engine.profiler.summary()
        engine.profiler.reset(False)
        nfe = nfe_of(name, model)
        stats = getattr(model, "last_stats", None)
        # ---- end-to-end: host buffers in, host result out, copies inside the timed region --------
        # Every step copies its inputs from pinned host memory and its result back to pinned host memory inside
        # the timed region.  The copies run on a second stream: while step i integrates, the inputs of step i+1 are
        # uploaded into the other device buffer and the result of step i-1 is downloaded (double buffering), so only
        # the first upload and the last download are exposed.
        e2e_s = 0.0
        if not args.no_e2e:
            copy_stream = torch.cuda.Stream(device=dev)
            main_stream = torch.cuda.current_stream()
            dbuf = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host.items()} for _ in range(2)]
            res_host = [torch.empty(out.shape, dtype=out.dtype).pin_memory() for _ in range(2)]
            landed = [torch.cuda.Event(), torch.cuda.Event()]

            def upload(i):
                with torch.cuda.stream(copy_stream):
                    for k, v in host.items():
                        dbuf[i % 2][k].copy_(v, non_blocking=True)
                    landed[i % 2].record(copy_stream)

            def e2e_pass(steps):
                """`steps` end-to-end steps; returns wall seconds from the first upload to the last download."""
                barrier()
                t0 = time.perf_counter()
                upload(0)
                for i in range(steps):
                    if i + 1 < steps:
                        upload(i + 1)                      # overlaps with the integration of step i
                    main_stream.wait_event(landed[i % 2])
                    o = run_gpu(name, model, dbuf[i % 2])
                    ready = torch.cuda.Event()
                    ready.record(main_stream)
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(ready)
                        res_host[i % 2].copy_(o, non_blocking=True)     # overlaps with the integration of step i+1
                        o.record_stream(copy_stream)
                    del o
                copy_stream.synchronize()
                torch.cuda.synchronize()
                barrier()
                return time.perf_counter() - t0

            e2e_pass(min(args.warmup, 2))      # untimed: the two-stream pattern warms the caching allocator's pools
            e2e_s = e2e_pass(args.steps)
    tmax = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    rows_all = torch.tensor([B], dtype=torch.int64, device=dev)
    steps_identical = None
    if world > 1:
        td.all_reduce(tmax, op=td.ReduceOp.MAX)
        td.all_reduce(rows_all, op=td.ReduceOp.SUM)
        if stats is not None and stats.dt_history:
            # every rank must have taken bit-identical step sizes and accept decisions (reference-exact dopri5: the
            # FP64 partial sums of the error norm are all-reduced): gather the histories and compare them on rank 0
            h = torch.zeros(2 * 256 + 2, dtype=torch.float64, device=dev)
            n = min(len(stats.dt_history), 256)
            h[0], h[1] = len(stats.dt_history), stats.accepted * 65536 + stats.rejected
            h[2:2 + n] = torch.tensor(stats.dt_history[:n], dtype=torch.float64)
            h[258:258 + n] = torch.tensor([float(a) for a in stats.accept_history[:n]], dtype=torch.float64)
            allh = [torch.empty_like(h) for _ in range(world)]
            td.all_gather(allh, h)
            steps_identical = all(torch.equal(allh[0], x) for x in allh[1:])
    ms, e2e_ms = float(tmax[0]), float(tmax[1])
    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return
    check = None if args.no_shard_check else shard_check(name, model, inp, out)
    total_rows = int(rows_all.item()) * args.steps
    value = total_rows / (ms * 1e-3)
    peaks, peak_src = measured_peaks()
    # ---- roofline of the dominant kernel: algorithmic FLOPs per launch / mean launch time ---------
    kname = max(prof, key=lambda k: prof[k][1]) if prof else None
    roof = None
    if kname:
        n_l, k_ms, k_rows = prof[kname]
        field_flops = {"cfg1": 17024, "cfg2": 109568, "cfg3": 1683712, "cfg4": 116736, "cfg5": 233472,
                       "wide256": 2 * (28 * 256 + 3 * 256 * 256 + 256 * 16)}[name]
        evals_per_launch = {"dopri5_attempt": 6, "field_eval": 1, "integrate_fixed": nfe or 1}[kname]
        flop_per_launch = field_flops * evals_per_launch * (k_rows / n_l)
        achieved = flop_per_launch / (k_ms / n_l * 1e-3) / 1e12
        ffma_peak = engine.ffma_peak_tflops()
        tensor_fp32_equiv = peaks["bf16_tflops"] / 6.0
        tf32_gemm = tf32_gemm_tflops(dev)
        clk_now = clk.summary().get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
        # the tensor pipe's own rate: one kind::tf32 M128 N128 K8 MMA per 64 cycles per SM (csrc/tc_rate.cu,
        # profiles/r02_tcgen05_rate_sustained.txt), three products per FP32-faithful MAC
        sms = engine.device_info()["sm_count"]
        instr_rate = 128 * 128 * 8 * 2 / 64.0 * sms * clk_now * 1e6 / 3.0 / 1e12
        # DRAM bytes of the dominant kernel from the committed ncu --set full capture, scaled to this launch's rows
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.isfile(tpath):
            t = json.load(open(tpath)).get(name)
            if t and t["kernel"].startswith({"dopri5_attempt": "k_dopri5", "integrate_fixed": "k_fixed", "field_eval": "k_field_eval"}[kname]):
                traffic = t["dram_bytes_per_launch"] / t["rows"] * (k_rows / n_l)
        if name == "wide256":      # the wide engine runs on the FP32 pipe: its roofline is the measured FFMA2 peak
            tensor_fp32_equiv, peak_src = ffma_peak, "measured in this run by ffb_ffma_peak (FP32 FFMA2 pipe: the wide engine's bound)"
        roof = {"bound": "fp32" if name == "wide256" else "tensor", "achieved": achieved, "peak": tensor_fp32_equiv, "unit": "TFLOP/s",
                "frac": achieved / tensor_fp32_equiv, "traffic": traffic,
                "kernel": kname, "launches": n_l, "avg_launch_ms": k_ms / n_l, "share_of_step": k_ms / ms,
                "peak_source": (peak_src if name == "wide256" else
                                f"MEASURED_PEAKS.json bf16_tflops ({peak_src}) / 6 = 3xTF32 FP32-equivalent tensor peak"),
                "fp32_ffma2_peak_measured": ffma_peak, "frac_of_ffma2_peak": achieved / ffma_peak,
                "tf32_gemm_tflops_measured": tf32_gemm, "frac_of_tf32_gemm_over_3": achieved / (tf32_gemm / 3.0),
                "tcgen05_3xtf32_instruction_rate_tflops": instr_rate, "frac_of_instruction_rate": achieved / instr_rate,
                "flop_per_launch": flop_per_launch}
    line = {"metric": w["metric"], "value": value, "unit": w["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{name}: {w['desc']}", "rows_per_gpu": B, "global_rows": int(rows_all.item()), "nfe": nfe,
                       "steps_identical_across_ranks": steps_identical, "shard_check": check, "numa": numa,
                       "dopri5_steps": None if stats is None else [stats.accepted, stats.rejected],
                       "dopri5_controller": getattr(stats, "controller", None),
                       "l2": "working set (state + derivative ping-pong buffers) exceeds the 126 MB L2; no flush needed",
                       "weights": "random init, torch.manual_seed(1234), reference construction order"},
            "clocks": clk.summary(), "gpu_launches": launches, "roofline": roof}
    if not args.no_e2e:
        hb = sum(v.numel() * 4 for v in host.values())
        line["e2e"] = {"value": total_rows / (e2e_ms * 1e-3), "unit": w["unit"], "h2d_bytes_per_step": hb,
                       "d2h_bytes_per_step": int(out.numel() * 4), "ms_per_step": e2e_ms / args.steps,
                       "copies": "pinned host <-> device on a second stream, double-buffered against the integration"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(name, w["cpu_B"])
    print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
