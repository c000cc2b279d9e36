"""Frame-batch sharding.

Frames are independent -- no pass of any hook reads another frame -- so the batch dimension is split
into contiguous chunks, one per GPU, LUTs / weights are replicated (<= 3.7 MB), and outputs stay on
the GPU that produced them.  There is NO data-path collective.  Two usage modes:

* one process, several devices: ``prescale(frames, hook, devices=[0, 1, ...])``;
* one process per GPU (torchrun): every rank calls ``rank_slice()`` on the global batch and runs
  ``prescale`` on its own slice; ``max_over_ranks()`` is the only communication (timing metadata).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, stop) per rank; the first n % world ranks get one extra frame."""
    if world < 1:
        raise ValueError("world must be >= 1")
    base, extra = divmod(max(n, 0), world)
    out, start = [], 0
    for r in range(world):
        stop = start + base + (1 if r < extra else 0)
        out.append((start, stop))
        start = stop
    return out


def rank_slice(n: int, rank: int, world: int) -> slice:
    a, b = shard_bounds(n, world)[rank]
    return slice(a, b)


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """MAX all-reduce of a scalar (step time) over the default process group; identity if not initialised."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def prescale_sharded(frames: torch.Tensor, hook, output_size, devices: Sequence, lut_precision: str, is_yuv: bool, **io_kwargs):
    """Split the batch dimension over ``devices``; returns one output tensor per device (None for empty shards)."""
    from .api import prescale

    devs = [torch.device("cuda", d) if isinstance(d, int) else torch.device(d) for d in devices]
    if frames.dim() < 3:
        raise ValueError("sharding needs a batch dimension")
    n = frames.shape[0]
    outs = []
    for (a, b), dev in zip(shard_bounds(n, len(devs)), devs):
        if b <= a:
            outs.append(None)
            continue
        part = frames[a:b]
        if part.device != dev:
            part = part.to(dev, non_blocking=True)
        with torch.cuda.device(dev):
            outs.append(prescale(part, hook, output_size, None, lut_precision, False, is_yuv, **io_kwargs))
    return outs
