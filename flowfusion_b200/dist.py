"""Batch sharding over the GPUs of one box.

The path shards along the batch of trajectories only: every rank holds the (tiny) weights
and a contiguous block of rows.  Fixed-step integrators need no communication.  dopri5 in
reference-exact mode all-reduces a few float64 partial sums per attempted step so that every
rank takes the same accept/reject decision and the same next step size (``solver.dopri5``).
One process per GPU, ``torch.distributed`` with the NCCL backend (gloo in the CPU tests).
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch
import torch.distributed as td

_GROUP = [None]


def current_group():
    """Process group used by model entry points when the model has no ``process_group``."""
    return _GROUP[0]


@contextlib.contextmanager
def use_group(group):
    """``with dist.use_group(td.group.WORLD): model.log_prob(x_shard)``"""
    prev = _GROUP[0]
    _GROUP[0] = group
    try:
        yield
    finally:
        _GROUP[0] = prev


def shard_bounds(n_rows: int, rank: int, world: int):
    """Contiguous row block of ``rank``: the first ``n_rows % world`` ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(t: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None):
    rank = td.get_rank() if rank is None else rank
    world = td.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def row_offset(local_rows: int, group) -> int:
    """Global index of this rank's first row (exclusive prefix sum of the shard sizes)."""
    if group is None:
        return 0
    world, rank = td.get_world_size(group), td.get_rank(group)
    dev = "cuda" if td.get_backend(group) == "nccl" else "cpu"
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = local_rows
    td.all_reduce(sizes, group=group)
    return int(sizes[:rank].sum().item())


def gather_rows(t: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather row blocks of (possibly) different sizes back into one tensor, in rank order."""
    group = group if group is not None else td.group.WORLD
    world = td.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    td.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    td.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
