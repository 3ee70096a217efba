"""Tensor-level entry points of the B200 engine.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed;
all arithmetic happens in libimprover_b200.so through the C ABI
(include/improver_b200.h).  Inputs are CUDA float32 tensors laid out
(..., y, x); nothing here falls back to the CPU.
"""
from __future__ import annotations

import contextlib
import os

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import NbhDesc, NbhThreshold

MAX_THRESHOLDS_PER_LAUNCH = 8
NAN_MESSAGE = "Error: NaN detected in input cube data"   # nbhood.py:168, threshold.py:665

_workspaces = {}


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (improver_b200 has no CPU path)")


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


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


@contextlib.contextmanager
def exact_square_sums():
    """Within this context single-member square launches use the kernels with exact window sums
    (float64 / fixed point: `square_rs_kernel` rows mode and the warp / block kernels) instead of
    the float32 running sums of `square_tile_kernel`.  Their results do not depend on where a tile
    starts, so a y-band of a grid reproduces the rows of the whole grid bit for bit."""
    old = os.environ.get("NBH_SQ_TILE")
    os.environ["NBH_SQ_TILE"] = "0"
    try:
        yield
    finally:
        if old is None:
            del os.environ["NBH_SQ_TILE"]
        else:
            os.environ["NBH_SQ_TILE"] = old


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
    if method == "square":
        fn = lib.nbh_square_f32
    elif method == "circular":
        fn = lib.nbh_circular_f32
    else:
        raise ValueError("{} is not a valid neighbourhood_method.".format(method))

    desc = NbhDesc()
    desc.n_planes, desc.n_members, desc.ny, desc.nx = n_planes, n_members, ny, nx
    desc.cells = int(cells)
    desc.plane_stride, desc.member_stride = slice_elems, member_stride
    desc.flags = flags
    desc.mask2d = mask2d.data_ptr() if mask2d is not None else None
    desc.mask_nd = mask_nd.data_ptr() if mask_nd is not None else None

    stream = _stream_ptr()
    with torch.cuda.device(data.device):
        if T <= MAX_THRESHOLDS_PER_LAUNCH:
            thr_arr = make_thresholds(thr_list) if T else None
            desc.n_thresholds = T
            desc.thresholds = thr_arr if T else None
            ws = _workspace(data.device, lib.nbh_workspace_bytes(C.byref(desc)))
            _lib.check(fn(C.byref(desc), data.data_ptr(), out.data_ptr(),
                          out_mask.data_ptr() if want_mask else None, status.data_ptr(),
                          ws.data_ptr(), ws.numel(), stream))
        else:
            # more thresholds than one launch group: chunks write into a
            # staging tensor that is scattered into the threshold axis
            acc_status = torch.zeros(2, dtype=torch.int32, device=data.device)
            for t0 in range(0, T, MAX_THRESHOLDS_PER_LAUNCH):
                chunk = thr_list[t0:t0 + MAX_THRESHOLDS_PER_LAUNCH]
                sub, sub_mask, st = neighbourhood(
                    data, cells, method, mask2d, mask_nd, weighted_mode, sum_only, chunk,
                    collapse_members, wrap_x, want_mask)
                out[..., t0:t0 + len(chunk), :, :] = sub
                if want_mask:
                    out_mask[..., t0:t0 + len(chunk), :, :] = sub_mask
                acc_status = torch.maximum(acc_status, st)
            status = acc_status
    return out, out_mask, status


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
                     edge_width: int = 15, mask: Optional[torch.Tensor] = None, mask_zeros: bool = False):
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
        max_call = max(1, min(max_call, int(os.environ.get("NBH_RF_MAX_SLICES", max_call))))
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
