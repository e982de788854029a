"""Multi-GPU plumbing: independent video streams are the unit of parallelism.

Stream `s` lives on rank `s mod world` for the whole run; there is no collective on the frame path.
`torch.distributed` (NCCL on the GPU box, gloo in CPU tests) is used only for the start barrier and for
reducing per-rank timings / gathering per-rank frame counts at the end of a run.
"""
from __future__ import annotations

from typing import List, Sequence


def streams_for_rank(n_streams: int, rank: int, world: int) -> List[int]:
    """Static partition: stream s -> rank s % world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return [s for s in range(n_streams) if s % world == rank]


def owner_of_stream(stream: int, world: int) -> int:
    return stream % world


def barrier(dist=None) -> None:
    if dist is not None and dist.is_initialized():
        dist.barrier()


def reduce_max(values: Sequence[float], dist=None, device=None) -> List[float]:
    """Element-wise MAX over ranks (timings are reported as the slowest rank's)."""
    if dist is None or not dist.is_initialized():
        return [float(v) for v in values]
    import torch

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def gather_counts(value: int, dist=None, device=None) -> List[int]:
    """Every rank's frame count, in rank order (so throughput is computed from a common window)."""
    if dist is None or not dist.is_initialized():
        return [int(value)]
    import torch

    mine = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [int(o.item()) for o in out]
