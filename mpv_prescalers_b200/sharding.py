"""Frame-batch sharding.

Frames are independent -- no pass of any hook reads another frame -- so the batch dimension is split
into contiguous chunks, one per GPU, LUTs / weights are replicated (<= 3.7 MB), and outputs stay on
the GPU that produced them.  There is NO data-path collective.  Two usage modes:

* one process, several devices: ``prescale(frames, hook, devices=[0, 1, ...])``;
* one process per GPU (torchrun): every rank calls ``rank_slice()`` on the global batch and runs
  ``prescale`` on its own slice; ``max_over_ranks()`` is the only communication (timing metadata).

A single large frame can instead be split by ROWS (``prescale(..., devices=[...], split="rows")``): GPU g gets
the source rows of its band plus ``ROW_HALO`` rows either side, fetched from the owning GPU by peer-to-peer copies
(NVLink on a B200 box), and the output rows that the halo produced are dropped, so the band seams see real
neighbouring rows while clamp-to-edge still applies at the true image borders only.
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


# Source rows a band needs beyond its own so that no kept output row is influenced by the artificial band edge:
# ravu-lite / 3x reach r-1 <= 3 rows, the fused ravu chain <= 2r = 8 (int11 near the seam is itself computed from
# clamped rows and is tapped +-r rows away), NNEDI3 double_y 3 rows + double_x 4 rows of the W x 2H image.
ROW_HALO = 12


def row_bands(h: int, parts: int, halo: int = ROW_HALO) -> List[Tuple[int, int, int, int]]:
    """(row0, row1, src0, src1) per band: the band owns source rows [row0, row1) and reads [src0, src1)."""
    out = []
    for a, b in shard_bounds(h, parts):
        out.append((a, b, max(a - halo, 0), min(b + halo, h)) if b > a else (a, b, a, b))
    return out


def prescale_rowsplit(frames: torch.Tensor, hook, output_size, devices: Sequence, lut_precision: str, is_yuv: bool,
                      gather: bool = True, **io_kwargs):
    """Split the ROW dimension of the frames over ``devices`` (a device may be listed more than once).

    Returns the full result on ``devices[0]`` (``gather=True``: the bands are concatenated there, peer-to-peer) or
    the list of per-device output bands.  Bit-identical to the unsplit result."""
    from .api import plan, prescale
    from .hookfile import HookError, HookFile, find_hook

    hk = hook if isinstance(hook, HookFile) else HookFile.parse(find_hook(hook))
    v = hk.variant
    if v.family == "ravu-zoom":
        raise HookError("row split is not available for ravu-zoom (its band geometry depends on the ratio); shard frames instead")
    if frames.dim() < 2:
        raise ValueError("row split needs [.., H, W] frames")
    devs = [torch.device("cuda", d) if isinstance(d, int) else torch.device(d) for d in devices]
    h, w = frames.shape[-2], frames.shape[-1]
    pl = plan(hk, (h, w), output_size, is_yuv)
    if not pl.applied:
        return prescale(frames, hk, output_size, None, lut_precision, False, is_yuv, **io_kwargs)
    sy = pl.out_size[0] // h if v.family != "nnedi3" else (2 if pl.double_y else 1)
    bands = list(zip(row_bands(h, len(devs)), devs))
    # 1. every band's source rows go to its GPU FIRST: a peer-to-peer copy is queued on the stream of the GPU that owns
    #    `frames`, and would otherwise wait behind the kernels of the bands launched there before it
    parts = []
    for (a, b, s0, s1), dev in bands:
        if b <= a:
            parts.append(None)
            continue
        part = frames[..., s0:s1, :]
        parts.append(part if part.device == dev else part.to(dev, non_blocking=True))
    # 2. the kernels of all bands run concurrently (launches are asynchronous, one GPU each)
    outs = []
    for ((a, b, s0, s1), dev), part in zip(bands, parts):
        if part is None:
            outs.append(None)
            continue
        # OUTPUT for the band keeps the frame's scale ratio, so that the //!WHEN expressions (ratios of HOOKED and
        # OUTPUT sizes) decide as they did for the whole frame; a band that decides differently is refused below
        osz = None if output_size is None else (max(1, round(output_size[0] * (s1 - s0) / h)), output_size[1])
        with torch.cuda.device(dev):
            full = prescale(part, hk, osz, None, lut_precision, False, is_yuv, **io_kwargs)
        if not getattr(full, "applied", True) or full.shape[-2] != (s1 - s0) * sy:
            raise HookError("row split: a band took a different WHEN decision than the whole frame; use frame sharding")
        keep = full[..., (a - s0) * sy:(b - s0) * sy, :]
        keep.offset, keep.applied, keep.plan = pl.offset, True, pl
        outs.append(keep)
    if not gather:
        return outs
    live = [o for o in outs if o is not None]
    res = torch.cat([o.to(devs[0], non_blocking=True) for o in live], dim=-2)
    res.offset, res.applied, res.plan = pl.offset, True, pl
    return res
