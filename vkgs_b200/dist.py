"""Multi-GPU host logic.  The reference is single-GPU; the path shards two ways, both with the
scene replicated per GPU and no collective on the data path except gathering finished pixels:

  by view   cameras dealt to ranks round-robin (deal_views) or in contiguous blocks (shard_views); every rank runs the
            whole frame pipeline per view;
            images are gathered to one rank (NCCL on GPUs, gloo in the CPU tests).
  by band   one view, rank g bins and blends only rows [edge[g], edge[g+1]) (VKGSB_OPT_BAND_Y0/Y1); the cull and the
            depth sort are replicated so every band sees the same global order; bands concatenate to the frame.

Everything here is plumbing around torch.distributed; it never touches pixels' values.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_views(n_views: int, rank: int, world: int) -> range:
    """Contiguous block of view indices for `rank` (blocks differ by at most one view)."""
    base, extra = divmod(n_views, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def deal_views(n_views: int, rank: int, world: int) -> range:
    """Views rank, rank + world, rank + 2 world, ...: neighbouring views of an orbit cost about the same, so dealing them
    round-robin gives every rank the same mix (contiguous blocks leave the rank with the heaviest arc last)."""
    return range(rank, n_views, world)


def band_edges(height: int, world: int, align: int = 16) -> List[int]:
    """Row boundaries of `world` horizontal bands, tile-aligned where possible, covering [0, height)."""
    tiles = (height + align - 1) // align
    edges = [min(height, ((tiles * g) // world) * align) for g in range(world)] + [height]
    return edges


def balanced_band_edges(row_load, world: int, min_rows: int = 8) -> List[int]:
    """Row boundaries of `world` bands carrying about equal load: `row_load[y]` = splat centres on image row y of a
    previous frame (Renderer.row_histogram()).  A constant per row is added for the per-pixel cost of an empty band,
    every band keeps at least `min_rows` rows, the bands cover [0, height)."""
    import numpy as np
    load = np.asarray(row_load, np.float64)
    h = len(load)
    load = load + max(load.sum(), 1.0) / h * 0.05
    cum = np.concatenate([[0.0], np.cumsum(load)])
    edges = [0]
    for g in range(1, world):
        y = int(np.searchsorted(cum, cum[-1] * g / world))
        y = max(y, edges[-1] + min_rows)
        y = min(y, h - (world - g) * min_rows)
        edges.append(y)
    edges.append(h)
    return edges


def rebalance_band_edges(edges, band_ms, min_rows: int = 8) -> List[int]:
    """One feedback step of the band partition: `band_ms[g]` = measured time of band g = rows [edges[g], edges[g + 1]).
    Every row of a band is charged an equal share of the band's time and the rows are re-cut into bands of equal charge
    (the splat-centre histogram does not see what a band costs apart from its splats: the cull over the frustum, the
    per-pixel work of the blend)."""
    import numpy as np
    edges = [int(e) for e in edges]
    world, h = len(edges) - 1, edges[-1]
    cost = np.zeros(h, np.float64)
    for g in range(world):
        rows = max(edges[g + 1] - edges[g], 1)
        cost[edges[g]:edges[g + 1]] = max(float(band_ms[g]), 0.0) / rows
    if not np.isfinite(cost).all() or cost.sum() <= 0.0:
        return edges
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    out = [0]
    for g in range(1, world):
        y = int(np.searchsorted(cum, cum[-1] * g / world))
        y = max(y, out[-1] + min_rows)
        y = min(y, h - (world - g) * min_rows)
        out.append(y)
    out.append(h)
    return out


def gather_images(local, dst: int = 0, group=None):
    """Gather equally shaped uint8 image tensors [B,H,W,4] to rank `dst` (NCCL on GPUs, gloo in the CPU tests).
    Returns the list on dst, else None.  On one node bench.py prefers rendering straight into rank dst's memory
    (vkgsb_shared_*, no collective); this is the portable path."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return [local]
    outs = [torch.empty_like(local) for _ in range(world)] if dist.get_rank(group) == dst else None
    dist.gather(local, outs, dst=dst, group=group)
    return outs


def assemble_views(gathered: Sequence, n_views: int) -> "np.ndarray":
    """Inverse of shard_views: per-rank [b_r,H,W,4] blocks -> [n_views,H,W,4] in view order."""
    import torch
    world = len(gathered)
    parts = []
    for r in range(world):
        k = len(shard_views(n_views, r, world))
        parts.append(gathered[r][:k])
    return torch.cat(parts, dim=0)


def assemble_bands(gathered: Sequence, edges: Sequence[int]):
    """Per-rank full-size frames whose own band rows are valid -> one frame."""
    import torch
    out = torch.empty_like(gathered[0])
    for g, img in enumerate(gathered):
        out[..., edges[g]:edges[g + 1], :, :] = img[..., edges[g]:edges[g + 1], :, :]
    return out


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
