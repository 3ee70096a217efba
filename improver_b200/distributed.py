"""Multi-GPU decomposition (one process per GPU, torch.distributed).

The hot path shards naturally: every (realization, threshold, lead-time) 2-D
slice is independent (nbhood.py:458-462), so ranks take contiguous ranges of
the flattened slice index and NO collective is needed on the data path.  For
the fused realization mean the unit of sharding is the lead time (all members
of a lead time stay on one rank, so the mean needs no reduction).

For a single grid too large for one GPU the alternative is y-band tiling:
each rank owns a band of rows of every slice and exchanges `cells` halo rows
with its two neighbours (NCCL send/recv over NVLink; gloo in the CPU tests)
before running the same kernels on the extended band.  A periodic x axis
(`wrap_x`) never crosses ranks under y-banding.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `n_items` for `rank` (sizes differ by at most 1)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("invalid rank / world_size")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def band_range(ny: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Rows [y0, y1) of the y-band owned by `rank`."""
    return shard_range(ny, world_size, rank)


def exchange_halo_rows(band: torch.Tensor, cells: int, group=None) -> torch.Tensor:
    """Return `band` (..., rows, nx) extended by up to `cells` rows from each
    neighbouring rank.  Rank 0 gets no top halo and the last rank no bottom halo
    (the domain-edge rule of the kernels applies there).  One batched
    isend/irecv per neighbour; works for CUDA tensors over NCCL and CPU tensors
    over gloo."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = band.shape[-2]
    if rows < cells and world > 1:
        raise ValueError(f"band of {rows} rows is thinner than the halo of {cells} rows")
    ops = []
    top = bottom = None
    send_top = send_bottom = None
    if rank > 0:
        top = torch.empty(band.shape[:-2] + (cells, band.shape[-1]), dtype=band.dtype, device=band.device)
        send_top = band[..., :cells, :].contiguous()
        ops.append(dist.P2POp(dist.isend, send_top, rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, top, rank - 1, group))
    if rank < world - 1:
        bottom = torch.empty(band.shape[:-2] + (cells, band.shape[-1]), dtype=band.dtype, device=band.device)
        send_bottom = band[..., rows - cells:, :].contiguous()
        ops.append(dist.P2POp(dist.isend, send_bottom, rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, bottom, rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    parts = [p for p in (top, band, bottom) if p is not None]
    return torch.cat(parts, dim=-2) if len(parts) > 1 else band


def crop_own_rows(extended: torch.Tensor, cells: int, group=None) -> torch.Tensor:
    """Inverse of exchange_halo_rows on a result computed from the extended band."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo = cells if rank > 0 else 0
    hi = extended.shape[-2] - (cells if rank < world - 1 else 0)
    return extended[..., lo:hi, :]


def neighbourhood_y_tiled(band: torch.Tensor, cells: int, mask_band: Optional[torch.Tensor] = None,
                          group=None, **kwargs):
    """Neighbourhood of a y-tiled grid: halo exchange, kernels on the extended
    band, crop.  `band` is this rank's (..., rows, nx) CUDA tensor.  Equal to
    the untiled result except for the reference's ones-trimming edge behaviour,
    whose bounding boxes are global (not reproduced in tiled mode).  The launch uses the
    kernels with exact window sums (`engine.exact_square_sums`): their result does not depend
    on the row a band starts at, so the bands reproduce the untiled exact-sum result bit for
    bit (the float32 running sums of the default tile kernel agree with it to ~1e-7)."""
    from . import engine

    ext = exchange_halo_rows(band, cells, group)
    m_ext = exchange_halo_rows(mask_band, cells, group) if mask_band is not None else None
    with engine.exact_square_sums():
        out, omask, status = engine.neighbourhood(ext, cells, mask2d=m_ext, **kwargs)
    out = crop_own_rows(out, cells, group)
    if omask is not None:
        omask = crop_own_rows(omask, cells, group)
    return out, omask, status


def neighbourhood_sharded(data: torch.Tensor, cells: int, gather: bool = False, group=None, **kwargs):
    """Slice-sharded neighbourhood: `data` (S, ny, nx) is the FULL batch on every
    rank (or any view of it); each rank computes its contiguous share.  With
    gather=True the shares are all-gathered (the only collective, off the
    compute path)."""
    from . import engine

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    s0, s1 = shard_range(data.shape[0], world, rank)
    out, omask, status = engine.neighbourhood(data[s0:s1].contiguous(), cells, **kwargs)
    if not gather:
        return out, omask, status, (s0, s1)
    sizes = [shard_range(data.shape[0], world, r) for r in range(world)]
    pieces: List[torch.Tensor] = [
        torch.empty((b - a,) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device) for a, b in sizes]
    dist.all_gather(pieces, out, group=group) if len({b - a for a, b in sizes}) == 1 else \
        _all_gather_uneven(pieces, out, group)
    return torch.cat(pieces, dim=0), omask, status, (s0, s1)


def _all_gather_uneven(pieces, mine, group):
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    for r in range(world):
        if r == rank:
            pieces[r].copy_(mine)
        dist.broadcast(pieces[r], src=r, group=group)
