"""Tensor-level entry points of the B200 engine.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed;
all arithmetic happens in libimprover_b200.so through the C ABI
(include/improver_b200.h).  Inputs are CUDA float32 tensors laid out
(..., y, x); nothing here falls back to the CPU.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import NbhDesc, NbhThreshold

MAX_THRESHOLDS_PER_LAUNCH = 8
NAN_MESSAGE = "Error: NaN detected in input cube data"   # nbhood.py:168, threshold.py:665

def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (improver_b200 has no CPU path)")


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Scratch for one native call, taken from PyTorch's caching allocator: blocks are reused in
    stream order, so concurrent streams never share one, and a call captured into a CUDA graph
    gets its block from the graph's private pool (it stays valid for every replay)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def make_thresholds(thresholds: Sequence[Tuple[float, float, float, str]]):
    """(value, lower bound, upper bound, comparison operator) -> NbhThreshold array."""
    arr = (NbhThreshold * len(thresholds))()
    for i, (val, lo, hi, op) in enumerate(thresholds):
        arr[i].value, arr[i].lo, arr[i].hi = float(val), float(lo), float(hi)
        arr[i].op = _lib.OPS[op] if isinstance(op, str) else int(op)
    return arr


def check_status(status: torch.Tensor) -> int:
    """Synchronise on the status words of a launch; raise the reference's
    ValueError if an unmasked NaN was read.  Returns the number of slices the
    trimming fix-up touched."""
    nan_flag, fixed = status.tolist()
    if nan_flag:
        raise ValueError(NAN_MESSAGE)
    return fixed


class Band:
    """Position of a y-band inside its grid (NbhDesc.y_offset / ny_global / halo_top /
    halo_bottom): the launch sees rows [y_offset, y_offset + rows) of `ny_global`, of which the
    first `halo_top` and the last `halo_bottom` came from the neighbouring bands."""

    def __init__(self, y_offset: int, ny_global: int, halo_top: int = 0, halo_bottom: int = 0):
        self.y_offset, self.ny_global = int(y_offset), int(ny_global)
        self.halo_top, self.halo_bottom = int(halo_top), int(halo_bottom)


def _fill_desc(data, cells, flags, n_planes, n_members, ny, nx, plane_stride, member_stride, thr_arr, T,
               mask2d, mask_nd, plane_cells=None, band: Optional[Band] = None, ext_stats=None):
    desc = NbhDesc()
    desc.n_planes, desc.n_members, desc.ny, desc.nx = n_planes, n_members, ny, nx
    desc.cells = int(cells)
    desc.plane_stride, desc.member_stride = plane_stride, member_stride
    desc.flags = flags
    desc.n_thresholds = T
    desc.thresholds = thr_arr if T else None
    desc.mask2d = mask2d.data_ptr() if mask2d is not None else None
    desc.mask_nd = mask_nd.data_ptr() if mask_nd is not None else None
    if plane_cells is not None:
        desc.plane_cells = plane_cells.ctypes.data_as(C.POINTER(C.c_int32))
    if band is not None:
        desc.y_offset, desc.ny_global = band.y_offset, band.ny_global
        desc.halo_top, desc.halo_bottom = band.halo_top, band.halo_bottom
    desc.ext_stats = ext_stats.data_ptr() if ext_stats is not None else None
    return desc


def neighbourhood(
    data: torch.Tensor,
    cells: int,
    method: str = "square",
    mask2d: Optional[torch.Tensor] = None,
    mask_nd: Optional[torch.Tensor] = None,
    weighted_mode: bool = False,
    sum_only: bool = False,
    thresholds=None,
    collapse_members: bool = False,
    wrap_x: bool = False,
    want_mask: bool = False,
    out: Optional[torch.Tensor] = None,
    exact: bool = False,
    land_sea: bool = False,
    plane_cells: Optional[Sequence[int]] = None,
    band: Optional[Band] = None,
    ext_stats: Optional[torch.Tensor] = None,
):
    """Neighbourhood sum/mean of every 2-D slice of `data`.

    data            float32 CUDA (..., ny, nx); with collapse_members the
                    LEADING axis holds the members (realizations) that are
                    averaged after neighbourhooding.
    mask2d          uint8/bool (ny, nx), non-zero = valid (mask_cube semantics).
    mask_nd         uint8/bool like data, non-zero = masked (np.ma semantics).
    thresholds      optional sequence of (value, lo, hi, op): the fuzzy
                    threshold is applied to each value first and a threshold
                    axis is inserted before (ny, nx).
    exact           single-member square launches: kernels whose window sums do not depend
                    on where a tile / band starts (NBH_EXACT_SUMS).
    land_sea        mask2d is a land mask; every point is normalised over the points of its
                    own surface type (NBH_LAND_SEA: land pass + sea pass + sum in one launch).
    plane_cells     one radius (grid cells) per plane instead of `cells` (lead-time radii).
    band, ext_stats y-band tiling: position of this band in its grid and the int32 (slices, 10)
                    statistics of the whole slices (slice_stats on every band + all-reduce).
    Returns (out, out_mask or None, status).  status is an int32[2] CUDA
    tensor; pass it to check_status() (this is the only synchronisation).
    """
    lib = _lib.load()
    _require_cuda(data, "data")
    if data.dtype != torch.float32:
        raise TypeError("data must be float32")
    if data.dim() < 2:
        raise ValueError("data must have at least two dimensions (y, x)")
    # a member axis with a padded stride is taken as it is (every member block contiguous):
    # the device layout of a fused cube is free, see pipeline.member_padded_empty
    padded_members = (collapse_members and data.dim() >= 3 and not data.is_contiguous()
                      and data[0].is_contiguous() and data.stride(0) >= data[0].numel())
    if not padded_members:
        data = data.contiguous()
    ny, nx = int(data.shape[-2]), int(data.shape[-1])
    lead = tuple(int(s) for s in data.shape[:-2])
    if collapse_members:
        if not lead:
            raise ValueError("collapse_members needs a leading member axis")
        n_members, plane_shape = lead[0], lead[1:]
    else:
        n_members, plane_shape = 1, lead
    n_planes = int(np.prod(plane_shape)) if plane_shape else 1
    slice_elems = ny * nx
    member_stride = (int(data.stride(0)) if padded_members else n_planes * slice_elems) if collapse_members else 0
    if n_planes == 0 or slice_elems == 0:
        raise ValueError("empty input")

    thr_list = list(thresholds) if thresholds is not None else []
    T = len(thr_list)
    out_shape = plane_shape + ((T,) if T else ()) + (ny, nx)
    if out is None:
        out = torch.empty(out_shape, dtype=torch.float32, device=data.device)
    elif tuple(out.shape) != out_shape or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError(f"out must be a contiguous float32 tensor of shape {out_shape}")
    out_mask = torch.empty(out_shape, dtype=torch.uint8, device=data.device) if want_mask else None
    status = torch.empty(2, dtype=torch.int32, device=data.device)

    if mask2d is not None:
        _require_cuda(mask2d, "mask2d")
        if tuple(mask2d.shape) != (ny, nx):
            raise ValueError("mask2d must have shape (ny, nx)")
        mask2d = (mask2d != 0).to(torch.uint8).contiguous() if mask2d.dtype != torch.uint8 \
            else mask2d.contiguous()
    if mask_nd is not None:
        _require_cuda(mask_nd, "mask_nd")
        if collapse_members:
            raise ValueError("mask_nd cannot be combined with collapse_members")
        if tuple(mask_nd.shape) != tuple(data.shape):
            raise ValueError("mask_nd must have the shape of data")
        mask_nd = (mask_nd != 0).to(torch.uint8).contiguous() if mask_nd.dtype != torch.uint8 \
            else mask_nd.contiguous()

    flags = 0
    if sum_only:
        flags |= _lib.SUM_ONLY
    if wrap_x:
        flags |= _lib.WRAP_X
    if weighted_mode:
        flags |= _lib.WEIGHTED
    if exact:
        flags |= _lib.EXACT_SUMS
    if land_sea:
        if mask2d is None:
            raise ValueError("land_sea needs the land mask as mask2d")
        flags |= _lib.LAND_SEA
    pc = None
    if plane_cells is not None:
        pc = np.ascontiguousarray(np.asarray(list(plane_cells), dtype=np.int32))
        if pc.shape != (n_planes,):
            raise ValueError(f"plane_cells must hold one radius per plane ({n_planes})")
    if ext_stats is not None:
        _require_cuda(ext_stats, "ext_stats")
        if ext_stats.dtype != torch.int32 or not ext_stats.is_contiguous() or \
                ext_stats.numel() != n_planes * n_members * max(len(list(thresholds or [])), 1) * 10:
            raise ValueError("ext_stats must be a contiguous int32 (slices, 10) tensor")
    if method == "square":
        fn = lib.nbh_square_f32
    elif method == "circular":
        fn = lib.nbh_circular_f32
    else:
        raise ValueError("{} is not a valid neighbourhood_method.".format(method))

    stream = _stream_ptr()
    # launch groups: runs of at most MAX_THRESHOLDS_PER_LAUNCH thresholds of one evaluation kind
    # (deterministic comparison / fuzzy: the kernels are specialised on it, threshold.py:438-455)
    groups = []
    for i, th in enumerate(thr_list):
        det = float(th[1]) == float(th[2])
        if groups and groups[-1][2] == det and len(groups[-1][1]) < MAX_THRESHOLDS_PER_LAUNCH:
            groups[-1][1].append(th)
        else:
            groups.append([i, [th], det])
    with torch.cuda.device(data.device):
        if len(groups) <= 1:
            thr_arr = make_thresholds(thr_list) if T else None
            desc = _fill_desc(data, cells, flags, n_planes, n_members, ny, nx, slice_elems, member_stride,
                              thr_arr, T, mask2d, mask_nd, pc, band, ext_stats)
            ws = _workspace(data.device, lib.nbh_workspace_bytes(C.byref(desc)))
            _lib.check(fn(C.byref(desc), data.data_ptr(), out.data_ptr(),
                          out_mask.data_ptr() if want_mask else None, status.data_ptr(),
                          ws.data_ptr(), ws.numel(), stream))
        else:
            # the groups write into staging tensors that are scattered into the threshold axis
            acc_status = torch.zeros(2, dtype=torch.int32, device=data.device)
            for t0, chunk, _ in groups:
                sub_stats = None
                if ext_stats is not None:
                    sub_stats = ext_stats.view(-1, T, 10)[:, t0:t0 + len(chunk)].contiguous()
                sub, sub_mask, st = neighbourhood(
                    data, cells, method, mask2d, mask_nd, weighted_mode, sum_only, chunk,
                    collapse_members, wrap_x, want_mask, None, exact, land_sea, plane_cells, band, sub_stats)
                out[..., t0:t0 + len(chunk), :, :] = sub
                if want_mask:
                    out_mask[..., t0:t0 + len(chunk), :, :] = sub_mask
                acc_status = torch.maximum(acc_status, st)
            status = acc_status
    return out, out_mask, status


def slice_stats(data: torch.Tensor, cells: int, mask2d: Optional[torch.Tensor] = None,
                mask_nd: Optional[torch.Tensor] = None, thresholds=None, collapse_members: bool = False,
                band: Optional[Band] = None) -> torch.Tensor:
    """Per-slice statistics int32 (slices, 10) of `data` as neighbourhood() would see it (after
    thresholding and masking): bounding boxes of the values != 0 and != 1 (rows in grid
    coordinates when `band` is given) and nanmin / nanmax keys.  Columns 0, 2, 4, 6, 8 combine
    across bands with MIN, columns 1, 3, 5, 7, 9 with MAX (nbhood.py:252-256, 353-382)."""
    lib = _lib.load()
    _require_cuda(data, "data")
    if data.dtype != torch.float32 or data.dim() < 2:
        raise TypeError("data must be float32 with (y, x) axes")
    data = data.contiguous()
    ny, nx = int(data.shape[-2]), int(data.shape[-1])
    lead = tuple(int(s) for s in data.shape[:-2])
    if collapse_members:
        n_members, plane_shape = lead[0], lead[1:]
    else:
        n_members, plane_shape = 1, lead
    n_planes = int(np.prod(plane_shape)) if plane_shape else 1
    thr_list = list(thresholds) if thresholds is not None else []
    T = len(thr_list)
    if T > MAX_THRESHOLDS_PER_LAUNCH:
        raise ValueError(f"at most {MAX_THRESHOLDS_PER_LAUNCH} thresholds per call")
    if mask2d is not None:
        mask2d = (mask2d != 0).to(torch.uint8).contiguous()
    if mask_nd is not None:
        mask_nd = (mask_nd != 0).to(torch.uint8).contiguous()
    thr_arr = make_thresholds(thr_list) if T else None
    desc = _fill_desc(data, cells, 0, n_planes, n_members, ny, nx, ny * nx,
                      n_planes * ny * nx if collapse_members else 0, thr_arr, T, mask2d, mask_nd, None, band)
    stats = torch.empty((n_planes * n_members * max(T, 1), 10), dtype=torch.int32, device=data.device)
    with torch.cuda.device(data.device):
        _lib.check(lib.nbh_slice_stats_f32(C.byref(desc), data.data_ptr(), stats.data_ptr(), _stream_ptr()))
    return stats


def neighbourhood_land_sea(data: torch.Tensor, cells: int, land_mask: torch.Tensor, method: str = "square",
                           weighted_mode: bool = False, sum_only: bool = False, thresholds=None,
                           collapse_members: bool = False, plane_cells=None):
    """Land points neighbourhooded over land points only, sea points over sea points only, both in
    one launch (the land pass + sea pass + filled(0) sum of cli/nbhood_land_and_sea.py:148-184).
    land_mask uint8/bool (ny, nx), non-zero = land.  Returns (out, status)."""
    out, _, status = neighbourhood(data, cells, method=method, mask2d=land_mask, weighted_mode=weighted_mode,
                                   sum_only=sum_only, thresholds=thresholds, collapse_members=collapse_members,
                                   land_sea=True, plane_cells=plane_cells)
    return out, status


def threshold(data: torch.Tensor, thresholds, mask_nd: Optional[torch.Tensor] = None,
              collapse_leading: bool = False):
    """Truth values of `data` against each threshold (threshold.py:643-713).

    Returns (out float32 (T, ...), contrib int32 (...), status).  With
    collapse_leading the leading axis is averaged over its unmasked members."""
    lib = _lib.load()
    _require_cuda(data, "data")
    if data.dtype != torch.float32:
        raise TypeError("data must be float32")
    data = data.contiguous()
    if collapse_leading:
        n_collapse, pshape = int(data.shape[0]), tuple(data.shape[1:])
    else:
        n_collapse, pshape = 1, tuple(data.shape)
    n_points = int(np.prod(pshape)) if pshape else 1
    thr_list = list(thresholds)
    T = len(thr_list)
    out = torch.empty((T,) + pshape, dtype=torch.float32, device=data.device)
    contrib = torch.empty(pshape, dtype=torch.int32, device=data.device)
    status = torch.zeros(2, dtype=torch.int32, device=data.device)
    if mask_nd is not None:
        _require_cuda(mask_nd, "mask_nd")
        mask_nd = (mask_nd != 0).to(torch.uint8).contiguous() if mask_nd.dtype != torch.uint8 \
            else mask_nd.contiguous()
    stream = _stream_ptr()
    with torch.cuda.device(data.device):
        for t0 in range(0, T, MAX_THRESHOLDS_PER_LAUNCH):
            chunk = thr_list[t0:t0 + MAX_THRESHOLDS_PER_LAUNCH]
            arr = make_thresholds(chunk)
            st = torch.empty(2, dtype=torch.int32, device=data.device)
            _lib.check(lib.nbh_threshold_f32(
                data.data_ptr(), mask_nd.data_ptr() if mask_nd is not None else None,
                out[t0:t0 + len(chunk)].data_ptr(), contrib.data_ptr(), arr, len(chunk),
                n_collapse, n_points, st.data_ptr(), stream))
            status = torch.maximum(status, st)
    return out, contrib, status


def collapse_mean(data: torch.Tensor) -> torch.Tensor:
    """Mean over the leading axis (collapse_realizations, cube_manipulation.py:93-131)."""
    lib = _lib.load()
    _require_cuda(data, "data")
    if data.dtype != torch.float32 or data.dim() < 2:
        raise TypeError("data must be float32 with a leading axis to collapse")
    data = data.contiguous()
    out = torch.empty(tuple(data.shape[1:]), dtype=torch.float32, device=data.device)
    with torch.cuda.device(data.device):
        _lib.check(lib.nbh_collapse_mean_f32(data.data_ptr(), out.data_ptr(), int(data.shape[0]),
                                             int(out.numel()), _stream_ptr()))
    return out


def boxsum_f64(data: torch.Tensor, box_y: int, box_x: int, cumsum: bool = True) -> torch.Tensor:
    """neighbourhood_tools.py:129-211 on float64 CUDA data (..., rows, cols) -> (..., rows - box_y,
    cols - box_x)."""
    lib = _lib.load()
    _require_cuda(data, "data")
    if data.dtype != torch.float64 or data.dim() < 2:
        raise TypeError("data must be float64 with two trailing axes")
    data = data.contiguous()
    H, W = int(data.shape[-2]), int(data.shape[-1])
    lead = tuple(data.shape[:-2])
    n = int(np.prod(lead)) if lead else 1
    out = torch.empty(lead + (H - box_y, W - box_x), dtype=torch.float64, device=data.device)
    with torch.cuda.device(data.device):
        _lib.check(lib.nbh_boxsum_f64(data.data_ptr(), out.data_ptr(), n, H, W, int(box_y), int(box_x),
                                      1 if cumsum else 0, _stream_ptr()))
    return out


def linear_map(data: torch.Tensor, a: float, b: float) -> torch.Tensor:
    """float32(a * double(data) + b): cf_units' conversion of a float32 array (threshold.py:430-431)."""
    lib = _lib.load()
    _require_cuda(data, "data")
    if data.dtype != torch.float32:
        raise TypeError("data must be float32")
    data = data.contiguous()
    out = torch.empty_like(data)
    with torch.cuda.device(data.device):
        _lib.check(lib.nbh_linear_map_f32(data.data_ptr(), out.data_ptr(), float(a), float(b), int(data.numel()),
                                          _stream_ptr()))
    return out


def zone_collapse(values: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """NaN-aware weighted mean over the zone axis (use_nbhood.py:158-192): values (..., Z, ny, nx),
    weights (Z, ny, nx) -> (..., ny, nx)."""
    lib = _lib.load()
    _require_cuda(values, "values")
    _require_cuda(weights, "weights")
    values, weights = values.contiguous(), weights.to(torch.float32).contiguous()
    nz, ny, nx = (int(s) for s in values.shape[-3:])
    if tuple(weights.shape) != (nz, ny, nx):
        raise ValueError("weights must have the (zone, y, x) shape of the values")
    lead = tuple(values.shape[:-3])
    n_lead = int(np.prod(lead)) if lead else 1
    out = torch.empty(lead + (ny, nx), dtype=torch.float32, device=values.device)
    with torch.cuda.device(values.device):
        _lib.check(lib.nbh_zone_collapse_f32(values.data_ptr(), weights.data_ptr(), out.data_ptr(), n_lead, nz,
                                             ny * nx, _stream_ptr()))
    return out


def threshold_vicinity(data: torch.Tensor, thresholds, vicinity_cells: Sequence[int],
                       mask_nd: Optional[torch.Tensor] = None, landmask: Optional[torch.Tensor] = None,
                       collapse_leading: bool = False):
    """Maximum-in-vicinity truth values (threshold.py:473-518, spatial.py:740-855).

    data            float32 CUDA (..., ny, nx); with collapse_leading the leading axis is
                    averaged over its unmasked members after the vicinity maximum.
    vicinity_cells  radii in grid cells (square maximum of side 2 r + 1, edges replicated).
    landmask        uint8/bool (ny, nx), non-zero = land: land and sea are spread separately.
    Returns (out float32 (V, T, ..., ny, nx), contrib int32 (..., ny, nx), status)."""
    lib = _lib.load()
    _require_cuda(data, "data")
    if data.dtype != torch.float32:
        raise TypeError("data must be float32")
    if data.dim() < 2 + (1 if collapse_leading else 0):
        raise ValueError("data must have (y, x) axes" + (" after the collapsed axis" if collapse_leading else ""))
    data = data.contiguous()
    if collapse_leading:
        n_collapse, pshape = int(data.shape[0]), tuple(int(s) for s in data.shape[1:])
    else:
        n_collapse, pshape = 1, tuple(int(s) for s in data.shape)
    ny, nx = pshape[-2], pshape[-1]
    n_planes = int(np.prod(pshape[:-2])) if len(pshape) > 2 else 1
    thr_list = list(thresholds)
    cells = np.asarray([int(c) for c in vicinity_cells], dtype=np.int32)
    if cells.size == 0 or not thr_list:
        raise ValueError("at least one threshold and one vicinity radius are needed")
    out = torch.empty((len(cells), len(thr_list)) + pshape, dtype=torch.float32, device=data.device)
    contrib = torch.empty(pshape, dtype=torch.int32, device=data.device)
    status = torch.zeros(2, dtype=torch.int32, device=data.device)
    if mask_nd is not None:
        _require_cuda(mask_nd, "mask_nd")
        mask_nd = (mask_nd != 0).to(torch.uint8).contiguous()
    if landmask is not None:
        _require_cuda(landmask, "landmask")
        if tuple(landmask.shape) != (ny, nx):
            raise ValueError("landmask must have the (y, x) shape of the data")
        landmask = (landmask != 0).to(torch.uint8).contiguous()
    with torch.cuda.device(data.device):
        # the native call takes any number of thresholds; build the array in chunks the
        # helper supports and pass them in one go
        arr = make_thresholds(thr_list)
        _lib.check(lib.nbh_threshold_vicinity_f32(
            data.data_ptr(), mask_nd.data_ptr() if mask_nd is not None else None,
            landmask.data_ptr() if landmask is not None else None, out.data_ptr(), contrib.data_ptr(),
            arr, len(thr_list), cells.ctypes.data_as(C.POINTER(C.c_int32)), len(cells), n_collapse, n_planes,
            ny, nx, status.data_ptr(), _stream_ptr()))
    return out, contrib, status


def extreme_within_vicinity(data: torch.Tensor, vicinity_cells: Sequence[int], operator: str = "max",
                            mask_nd: Optional[torch.Tensor] = None, landmask: Optional[torch.Tensor] = None):
    """maximum_within_vicinity / minimum_within_vicinity (spatial.py:805-903) of every 2-D
    slice.  data float32 CUDA (..., ny, nx) -> (V, ..., ny, nx); masked points hold 0."""
    lib = _lib.load()
    _require_cuda(data, "data")
    if data.dtype != torch.float32 or data.dim() < 2:
        raise TypeError("data must be float32 with (y, x) axes")
    if operator not in ("max", "min"):
        raise ValueError("operator must be 'max' or 'min'")
    data = data.contiguous()
    ny, nx = int(data.shape[-2]), int(data.shape[-1])
    n_planes = int(np.prod(data.shape[:-2])) if data.dim() > 2 else 1
    cells = np.asarray([int(c) for c in vicinity_cells], dtype=np.int32)
    out = torch.empty((len(cells),) + tuple(data.shape), dtype=torch.float32, device=data.device)
    status = torch.zeros(2, dtype=torch.int32, device=data.device)
    if mask_nd is not None:
        mask_nd = (mask_nd != 0).to(torch.uint8).contiguous()
    if landmask is not None:
        landmask = (landmask != 0).to(torch.uint8).contiguous()
    with torch.cuda.device(data.device):
        _lib.check(lib.nbh_vicinity_f32(
            data.data_ptr(), mask_nd.data_ptr() if mask_nd is not None else None,
            landmask.data_ptr() if landmask is not None else None, out.data_ptr(),
            cells.ctypes.data_as(C.POINTER(C.c_int32)), len(cells), 0 if operator == "max" else 1,
            n_planes, ny, nx, status.data_ptr(), _stream_ptr()))
    return out, status


def recursive_filter(data: torch.Tensor, coeff_x: torch.Tensor, coeff_y: torch.Tensor, iterations: int,
                     edge_width: int = 15, mask: Optional[torch.Tensor] = None, mask_zeros: bool = False,
                     max_slices_per_call: Optional[int] = None):
    """The per-slice body of RecursiveFilter.process (recursive_filter.py:403-436) for every
    (y, x) slice of `data` float32 CUDA (..., ny, nx): symmetric halo of 2 * edge_width,
    `iterations` x (x forward, x backward, y forward, y backward), halo removed.  coeff_x
    (ny, nx - 1), coeff_y (ny - 1, nx) float32.  `mask` (non-zero = masked) is either one
    (ny, nx) plane or has the shape of `data`; coefficients next to masked points are zeroed."""
    lib = _lib.load()
    for t, name in ((data, "data"), (coeff_x, "coeff_x"), (coeff_y, "coeff_y")):
        _require_cuda(t, name)
        if t.dtype != torch.float32:
            raise TypeError(f"{name} must be float32")
    if data.dim() < 2:
        raise TypeError("data must have (y, x) axes")
    data = data.contiguous()
    ny, nx = int(data.shape[-2]), int(data.shape[-1])
    if tuple(coeff_x.shape) != (ny, nx - 1) or tuple(coeff_y.shape) != (ny - 1, nx):
        raise ValueError(f"coefficient shapes {tuple(coeff_x.shape)}, {tuple(coeff_y.shape)} do not match "
                         f"a ({ny}, {nx}) slice")
    n_slices = int(np.prod(data.shape[:-2])) if data.dim() > 2 else 1
    per_slice = 0
    if mask is not None:
        _require_cuda(mask, "mask")
        if tuple(mask.shape) == tuple(data.shape) and data.dim() > 2:
            per_slice = 1
        elif tuple(mask.shape) != (ny, nx):
            raise ValueError("mask must be (ny, nx) or have the shape of data")
        mask = (mask != 0).to(torch.uint8).contiguous()
    coeff_x, coeff_y = coeff_x.contiguous(), coeff_y.contiguous()
    out = torch.empty_like(data)
    pad = 2 * int(edge_width)
    with torch.cuda.device(data.device):
        # slices per launch group: bounded by the grid (65535) and by a workspace budget (padded data
        # planes + per-slice flag planes) of half the free device memory, at most 16 GiB
        coeff_flag = 1 if (per_slice or mask_zeros) else 0
        one = lib.nbh_recursive_filter_workspace_bytes(1, ny, nx, pad, coeff_flag)
        two = lib.nbh_recursive_filter_workspace_bytes(2, ny, nx, pad, coeff_flag)
        budget = min(16 << 30, torch.cuda.mem_get_info(data.device)[0] // 2)
        max_call = int(max(1, min(65535, (budget - one) // max(1, two - one) + 1)))
        if max_slices_per_call is not None:
            max_call = max(1, min(max_call, int(max_slices_per_call)))
        flat_in, flat_out = data.view(n_slices, ny, nx), out.view(n_slices, ny, nx)
        flat_mask = mask.view(n_slices, ny, nx) if per_slice else mask
        for s0 in range(0, n_slices, max_call):
            n = min(max_call, n_slices - s0)
            nbytes = lib.nbh_recursive_filter_workspace_bytes(n, ny, nx, pad, coeff_flag)
            ws = _workspace(data.device, nbytes)
            m = None if flat_mask is None else (flat_mask[s0:s0 + n] if per_slice else flat_mask)
            _lib.check(lib.nbh_recursive_filter_f32(
                flat_in[s0:s0 + n].data_ptr(), coeff_x.data_ptr(), coeff_y.data_ptr(),
                m.data_ptr() if m is not None else None, per_slice, 1 if mask_zeros else 0,
                flat_out[s0:s0 + n].data_ptr(), n, ny, nx, pad, int(iterations), ws.data_ptr(), ws.numel(),
                _stream_ptr()))
    return out


def recurse(grid: torch.Tensor, coeffs: torch.Tensor, axis: int, forward: bool = True) -> torch.Tensor:
    """RecursiveFilter._recurse_forward / _recurse_backward (recursive_filter.py:55-153) in place on
    float32 CUDA (..., ny, nx); axis counts the spatial axes (0 = y, 1 = x)."""
    lib = _lib.load()
    _require_cuda(grid, "grid")
    _require_cuda(coeffs, "coeffs")
    if grid.dtype != torch.float32 or coeffs.dtype != torch.float32 or not grid.is_contiguous():
        raise TypeError("grid and coeffs must be float32, grid contiguous")
    ny, nx = int(grid.shape[-2]), int(grid.shape[-1])
    want = (ny - 1, nx) if axis == 0 else (ny, nx - 1)
    if tuple(coeffs.shape) != want:
        raise ValueError(f"coefficients must have shape {want}")
    n_slices = int(np.prod(grid.shape[:-2])) if grid.dim() > 2 else 1
    with torch.cuda.device(grid.device):
        _lib.check(lib.nbh_recurse_f32(grid.data_ptr(), coeffs.contiguous().data_ptr(), n_slices, ny, nx, int(axis),
                                       1 if forward else 2, _stream_ptr()))
    return grid


def neighbourhood_percentiles(data: torch.Tensor, cells: int, percentiles: Sequence[float]):
    """Percentiles of the circular neighbourhood of every point
    (nbhood.py:649-667).  data float32 (..., ny, nx) -> (..., P, ny, nx)."""
    lib = _lib.load()
    _require_cuda(data, "data")
    if data.dtype != torch.float32:
        raise TypeError("data must be float32")
    data = data.contiguous()
    ny, nx = int(data.shape[-2]), int(data.shape[-1])
    lead = tuple(int(s) for s in data.shape[:-2])
    n_slices = int(np.prod(lead)) if lead else 1
    pcts = np.asarray(list(percentiles), dtype=np.float32)
    out = torch.empty(lead + (len(pcts), ny, nx), dtype=torch.float32, device=data.device)
    status = torch.empty(2, dtype=torch.int32, device=data.device)
    ws = _workspace(data.device, lib.nbh_percentiles_workspace_bytes(n_slices, ny, nx, int(cells)))
    with torch.cuda.device(data.device):
        _lib.check(lib.nbh_percentiles_f32(
            data.data_ptr(), out.data_ptr(), pcts.ctypes.data_as(C.POINTER(C.c_float)), len(pcts),
            n_slices, ny, nx, int(cells), status.data_ptr(), ws.data_ptr(), ws.numel(),
            _stream_ptr()))
    return out, status


def device_info():
    lib = _lib.load()
    sm, smem = C.c_int32(), C.c_int64()
    _lib.check(lib.nbh_device_info(C.byref(sm), C.byref(smem)))
    return sm.value, smem.value
