"""Multi-GPU decomposition (one process per GPU, torch.distributed).

The hot path shards naturally: every (realization, threshold, lead-time) 2-D
slice is independent (nbhood.py:458-462), so ranks take contiguous ranges of
the flattened slice index and NO collective is needed on the data path.  For
the fused realization mean the unit of sharding is the lead time (all members
of a lead time stay on one rank, so the mean needs no reduction).

For a single grid too large for one GPU the alternative is y-band tiling:
each rank owns a band of rows of every slice and exchanges `cells` halo rows
with its two neighbours before running the same kernels on the extended band.
On CUDA tensors the exchange is ONE native call (`nbh_halo_exchange`: pack
kernel, ncclSend / ncclRecv pair per neighbour over NVLink, unpack kernel) on
the communicator torch.distributed already owns; CPU tensors (the gloo tests)
use batched isend / irecv.  The per-slice properties of the reference that
belong to the WHOLE slice — nanmin / nanmax for the clip (nbhood.py:252-256,
312) and the bounding boxes of the trimming (nbhood.py:341-409) — become a
10-scalar-per-slice min / max all-reduce.  A periodic x axis (`wrap_x`) never
crosses ranks under y-banding.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def _world(group=None) -> int:
    """World size; 1 when torch.distributed is not initialised (single-process use)."""
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _rank(group=None) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


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


def _nccl_comm_ptr(group, device) -> Optional[int]:
    """The raw ncclComm_t of `group` on `device`, or None when the group is not NCCL-backed."""
    try:
        pg = group if group is not None else dist.distributed_c10d._get_default_group()
        backend = pg._get_backend(torch.device(device))
        ptr = backend._comm_ptr()
        return int(ptr) if ptr else None
    except Exception:
        return None


def exchange_halo_rows(band: torch.Tensor, cells: int, group=None) -> torch.Tensor:
    """Return `band` (..., rows, nx) extended by `cells` rows from each neighbouring rank.
    Rank 0 gets no top halo and the last rank no bottom halo (the domain-edge rule of the kernels
    applies there).  CUDA float32 tensors over NCCL go through the native one-call exchange."""
    world = _world(group)
    rank = _rank(group)
    rows = band.shape[-2]
    if world == 1:
        return band
    if rows < cells:
        raise ValueError(f"band of {rows} rows is thinner than the halo of {cells} rows")
    if band.is_cuda and band.dtype == torch.float32:
        comm = _nccl_comm_ptr(group, band.device)
        if comm is not None:
            return _exchange_native(band, cells, rank, world, comm)
    ops = []
    top = bottom = None
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


def _exchange_native(band: torch.Tensor, cells: int, rank: int, world: int, comm: int) -> torch.Tensor:
    from . import _lib, engine

    lib = _lib.load()
    band = band.contiguous()
    rows, nx = int(band.shape[-2]), int(band.shape[-1])
    lead = tuple(band.shape[:-2])
    n_slices = 1
    for s in lead:
        n_slices *= int(s)
    top = cells if rank > 0 else 0
    bottom = cells if rank < world - 1 else 0
    ext = torch.empty(lead + (top + rows + bottom, nx), dtype=torch.float32, device=band.device)
    nbytes = lib.nbh_halo_exchange_workspace_bytes(n_slices, nx, cells)
    ws = engine._workspace(band.device, nbytes)
    with torch.cuda.device(band.device):
        _lib.check(lib.nbh_halo_exchange(C.c_void_p(comm), rank, world, band.data_ptr(), ext.data_ptr(), n_slices,
                                         rows, nx, cells, ws.data_ptr(), ws.numel(), engine._stream_ptr()))
    return ext


def crop_own_rows(extended: torch.Tensor, cells: int, group=None) -> torch.Tensor:
    """Inverse of exchange_halo_rows on a result computed from the extended band."""
    world = _world(group)
    rank = _rank(group)
    lo = cells if rank > 0 else 0
    hi = extended.shape[-2] - (cells if rank < world - 1 else 0)
    return extended[..., lo:hi, :]


def allreduce_slice_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Combine int32 (slices, 10) statistics of the bands into those of the whole slices: even
    columns are minima, odd columns maxima (include/improver_b200.h NbhSliceStats)."""
    if _world(group) == 1:
        return stats
    mins = stats[:, 0::2].contiguous()
    maxs = stats[:, 1::2].contiguous()
    dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    out = torch.empty_like(stats)
    out[:, 0::2] = mins
    out[:, 1::2] = maxs
    return out


def neighbourhood_y_tiled(band: torch.Tensor, cells: int, ny_global: int, y_offset: int,
                          mask_band: Optional[torch.Tensor] = None, group=None, **kwargs):
    """Neighbourhood of a y-tiled grid.  `band` is this rank's (..., rows, nx) CUDA tensor holding
    rows [y_offset, y_offset + rows) of a grid of `ny_global` rows (ranks ordered top to bottom).
    Halo exchange, statistics of the whole slices (all-reduce), kernels on the extended band with
    its position in the grid, crop.  Bit-identical to the untiled launch of the same kernels
    (`exact=True`: window sums that do not depend on where a band starts), including the
    reference's ones-trimming behaviour whose bounding boxes are global."""
    from . import engine

    world = _world(group)
    rank = _rank(group)
    ext = exchange_halo_rows(band, cells, group)
    m_ext = None
    if mask_band is not None:
        m_ext = exchange_halo_rows((mask_band != 0).to(torch.float32), cells, group).to(torch.uint8)
    top = cells if rank > 0 else 0
    bottom = cells if rank < world - 1 else 0
    pos = engine.Band(y_offset - top, ny_global, top, bottom)
    method = kwargs.get("method", "square")
    stats = None
    if method == "square" and not kwargs.get("sum_only", False):
        local = engine.slice_stats(ext, cells, mask2d=m_ext, mask_nd=kwargs.get("mask_nd"),
                                   thresholds=kwargs.get("thresholds"),
                                   collapse_members=kwargs.get("collapse_members", False), band=pos)
        stats = allreduce_slice_stats(local, group)
    if method == "square":
        out, omask, status = engine.neighbourhood(ext, cells, mask2d=m_ext, exact=True, band=pos, ext_stats=stats,
                                                  **kwargs)
    else:
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

    world = _world(group)
    rank = _rank(group)
    s0, s1 = shard_range(data.shape[0], world, rank)
    if s1 > s0:
        out, omask, status = engine.neighbourhood(data[s0:s1].contiguous(), cells, **kwargs)
    else:
        # more ranks than slices: this rank has no share; one slice gives the trailing shape
        out, omask, status = engine.neighbourhood(data[:1].contiguous(), cells, **kwargs)
        out = out[:0]
        omask = None if omask is None else omask[:0]
    if not gather:
        return out, omask, status, (s0, s1)
    sizes = [shard_range(data.shape[0], world, r) for r in range(world)]
    pieces: List[torch.Tensor] = [
        torch.empty((b - a,) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device) for a, b in sizes]
    dist.all_gather(pieces, out, group=group) if len({b - a for a, b in sizes}) == 1 else \
        _all_gather_uneven(pieces, out, group)
    return torch.cat(pieces, dim=0), omask, status, (s0, s1)


def percentiles_sharded(data: torch.Tensor, cells: int, percentiles, gather: bool = False, group=None):
    """GeneratePercentilesFromANeighbourhood over ranks: the (realization) slices are independent
    (nbhood.py:692-705), so each rank takes a contiguous share of the leading axis.  Returns
    (out (share or all, P, ny, nx), status, (s0, s1))."""
    from . import engine

    world = _world(group)
    rank = _rank(group)
    s0, s1 = shard_range(data.shape[0], world, rank)
    if s1 > s0:
        out, status = engine.neighbourhood_percentiles(data[s0:s1].contiguous(), cells, percentiles)
    else:
        out = torch.empty((0, len(list(percentiles))) + tuple(data.shape[1:]), dtype=torch.float32, device=data.device)
        status = torch.zeros(2, dtype=torch.int32, device=data.device)
    if not gather:
        return out, status, (s0, s1)
    sizes = [shard_range(data.shape[0], world, r) for r in range(world)]
    pieces = [torch.empty((b - a,) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device) for a, b in sizes]
    _all_gather_uneven(pieces, out, group)
    return torch.cat(pieces, dim=0), status, (s0, s1)


def _all_gather_uneven(pieces, mine, group):
    world = _world(group)
    rank = _rank(group)
    for r in range(world):
        if r == rank:
            pieces[r].copy_(mine)
        if pieces[r].numel():
            dist.broadcast(pieces[r], src=r, group=group)
