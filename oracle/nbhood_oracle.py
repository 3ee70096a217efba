"""CPU oracle for the gridded neighbourhood-processing hot path.

TEST INFRASTRUCTURE ONLY.  This is a numpy/scipy restatement of the reference's
algorithm; it may be imported by `tests/`, by `__graft_entry__.smoke()` and by
`bench.py`'s `cpu_baseline` / `--impl reference` legs, and by nothing else.  The
product path (`improver_b200/`) never imports it and has no CPU fallback.

Parity status: PINNED.  Every function below is checked (tests/test_oracle.py)
against
  * the reference's own known-answer tests for this path, transcribed as
    literal arrays (tests/kat_*.py cite the reference test file:line), and
  * golden vectors produced by running the real reference core from
    /root/reference in the build container (tests/golden/make_golden.py,
    fixtures committed as tests/golden/*.npz).

Each function cites the reference file:line it follows (paths relative to the
reference root).  The same numpy/scipy primitives are used as upstream
(np.pad + cumsum summed-area table, scipy.ndimage.correlate, np.percentile) so
that timing this module is a fair stand-in for the reference's CPU path.
"""
from __future__ import annotations

import operator
from typing import Optional, Sequence, Tuple

import numpy as np
from scipy.ndimage import correlate as _correlate
from scipy.ndimage import maximum_filter as _maximum_filter

F32 = np.float32

# improver/utilities/probability_manipulation.py:15-66
COMPARATORS = {
    ">": (operator.gt, "greater_than"), "gt": (operator.gt, "greater_than"),
    ">=": (operator.ge, "greater_than_or_equal_to"), "ge": (operator.ge, "greater_than_or_equal_to"),
    "<": (operator.lt, "less_than"), "lt": (operator.lt, "less_than"),
    "<=": (operator.le, "less_than_or_equal_to"), "le": (operator.le, "less_than_or_equal_to"),
    "==": (operator.eq, "equal_to"), "eq": (operator.eq, "equal_to"),
}


# --------------------------------------------------------------------------- #
# Thresholding
# --------------------------------------------------------------------------- #
def fuzzy_bounds_from_factor(thresholds: Sequence[float], fuzzy_factor: float):
    """improver/threshold.py:302-315 (`_generate_fuzzy_bounds`)."""
    out = []
    for thr in thresholds:
        lo, hi = thr * fuzzy_factor, thr * (2.0 - fuzzy_factor)
        if thr < 0:
            lo, hi = hi, lo
        out.append((lo, hi))
    return out


def linear_rescale(data, data_range, scale_range, clip=False):
    """improver/utilities/rescale.py:45-69 (numpy weak-scalar promotion kept:
    python-float range ends do not up-cast float32 data)."""
    d0, d1 = data_range
    s0, s1 = scale_range
    if d0 == d1:
        raise ValueError("Cannot rescale a zero input range ({} -> {})".format(d0, d1))
    if s0 == s1:
        raise ValueError("Cannot rescale a zero output range ({} -> {})".format(s0, s1))
    res = ((data - d0) * (s1 - s0) / (d1 - d0)) + s0
    if clip:
        res = np.clip(res, min(s0, s1), max(s0, s1))
    return res


def truth_value(data, threshold: float, bounds: Tuple[float, float], comparison: str = ">"):
    """improver/threshold.py:407-471 (`_calculate_truth_value`, no unit change).

    `data` is float32 ndarray or MaskedArray.  Returns float32 (masked if the
    input is masked and the threshold is fuzzy)."""
    fn, spp = COMPARATORS[comparison]
    thr = np.float64(threshold).astype(data.dtype)
    if bounds[0] == bounds[1]:
        tv = fn(data, thr)
    else:
        tv = np.where(
            data < thr,
            linear_rescale(data, (bounds[0], thr), (0.0, 0.5), clip=True),
            linear_rescale(data, (thr, bounds[1]), (0.5, 1.0), clip=True),
        )
        if np.ma.is_masked(data):
            tv = np.where(data.mask, 0, tv)
            tv = np.ma.masked_array(tv, mask=data.mask)
        if "less_than" in spp:
            tv = 1.0 - tv
    return tv.astype(F32)


def threshold_cube(data, thresholds, bounds, comparison=">", collapse_leading=False):
    """improver/threshold.py:643-713: per-threshold truth values, optionally
    averaged over the leading (realization) axis with masked points excluded.

    data: float32 (S, ...) when collapse_leading else (...).  Returns
    (values float32 (T, ...), invalid-mask bool or None, contribution counts int64)."""
    slices = list(data) if collapse_leading else [data]
    shape = slices[0].shape
    out = np.zeros((len(thresholds),) + shape, dtype=F32)
    contrib = np.zeros(shape, dtype=int)
    for sl in slices:
        if np.isnan(np.ma.getdata(sl)[~np.ma.getmaskarray(sl)]).any():
            raise ValueError("Error: NaN detected in input cube data")
        if np.ma.is_masked(sl):
            unmasked = ~sl.mask
        else:
            unmasked = np.ones(shape, dtype=bool)
        contrib += unmasked
        for i, (thr, bnd) in enumerate(zip(thresholds, bounds)):
            tv = truth_value(sl, thr, bnd, comparison)
            out[i][unmasked] += np.ma.getdata(tv)[unmasked]
    valid = contrib.astype(bool)
    for i in range(len(thresholds)):
        out[i, valid] = np.divide(out[i][valid], contrib[valid])
    inv = None if valid.all() else ~valid
    return out, inv, contrib


# value the reference uses to take masked / other-surface points out of a maximum:
# -1 * netCDF4.default_fillvals["f4"] (improver/utilities/spatial.py:851)
VICINITY_FILL = -9.969209968386869e36


def maximum_within_vicinity(grid, grid_point_radius: int, landmask=None):
    """improver/utilities/spatial.py:740-855 (`operator_within_vicinity` with the
    maximum filter): square maximum of side 2 * radius + 1, edges replicated,
    masked points excluded, land and sea points processed separately when a
    land mask is given.  NaN-free input (the callers raise on NaN first)."""
    width = 2 * int(grid_point_radius) + 1
    processed = grid.copy()
    if np.ma.is_masked(grid):
        unmasked_grid = grid.data.copy()
        unmasked_grid[grid.mask] = VICINITY_FILL
    else:
        unmasked_grid = grid.copy()
    if landmask is not None:
        patch = np.empty_like(grid)
        for match in (True, False):
            matched = unmasked_grid.copy()
            matched[landmask != match] = VICINITY_FILL
            filtered = _maximum_filter(matched, size=width, mode="nearest")
            patch = np.where(landmask == match, filtered, patch)
    else:
        patch = _maximum_filter(unmasked_grid, size=width, mode="nearest")
    if np.ma.is_masked(grid):
        processed.data[~grid.mask] = patch[~grid.mask]
    else:
        processed = patch
    return processed


def threshold_vicinity_cube(data, thresholds, bounds, comparison, vicinity_cells, landmask=None,
                            collapse_leading=False):
    """improver/threshold.py:643-713 with `vicinity` set (:473-518): per slice and
    threshold the truth value is spread with `maximum_within_vicinity` (every 2-D
    slice separately) before it is accumulated over the collapsed axis.

    data: float32 (S, ..., y, x) when collapse_leading else (..., y, x).  Returns
    (values float32 (V, T, ..., y, x), invalid-mask bool or None, counts)."""
    slices = list(data) if collapse_leading else [data]
    shape = slices[0].shape
    out = np.zeros((len(vicinity_cells), len(thresholds)) + shape, dtype=F32)
    contrib = np.zeros(shape, dtype=int)
    lm = None if landmask is None else np.asarray(landmask).astype(bool)
    for sl in slices:
        if np.isnan(np.ma.getdata(sl)[~np.ma.getmaskarray(sl)]).any():
            raise ValueError("Error: NaN detected in input cube data")
        unmasked = ~sl.mask if np.ma.is_masked(sl) else np.ones(shape, dtype=bool)
        contrib += unmasked
        for i, (thr, bnd) in enumerate(zip(thresholds, bounds)):
            tv = truth_value(sl, thr, bnd, comparison)
            for iv, cells in enumerate(vicinity_cells):
                if tv.ndim > 2:
                    maxes = np.zeros(tv.shape, dtype=F32)
                    for idx in np.ndindex(*tv.shape[:-2]):
                        maxes[idx] = maximum_within_vicinity(tv[idx], cells, lm)
                else:
                    maxes = maximum_within_vicinity(tv, cells, lm)
                out[iv][i][unmasked] += np.ma.getdata(maxes)[unmasked]
    valid = contrib.astype(bool)
    for iv in range(len(vicinity_cells)):
        for i in range(len(thresholds)):
            out[iv, i, valid] = np.divide(out[iv, i][valid], contrib[valid])
    inv = None if valid.all() else ~valid
    return out, inv, contrib


# --------------------------------------------------------------------------- #
# Neighbourhood sums
# --------------------------------------------------------------------------- #
def circular_kernel(ranges: int, weighted_mode: bool) -> np.ndarray:
    """improver/nbhood/nbhood.py:55-92."""
    n = int(2 * ranges + 1)
    off = np.arange(-ranges, ranges + 1)
    d2 = (off[:, None] ** 2 + off[None, :] ** 2).astype(float)
    area = ranges * ranges
    kern = np.ones((n, n))
    if weighted_mode:
        kern[:] = (area - d2) / area
        kern[kern < 0.0] = 0.0
    else:
        kern[d2 > area] = 0.0
    return kern


def boxsum(data, boxsize, cumsum=True, **pad_options):
    """improver/utilities/neighbourhood_tools.py:104-211 (summed-area table)."""
    box = np.atleast_1d(boxsize)
    if not issubclass(box.dtype.type, np.integer):
        raise ValueError("The size of the neighbourhood must be of an integer type.")
    if not np.all(box % 2):
        raise ValueError("The size of the neighbourhood must be an odd number.")
    bi, bj = int(box[0]), int(box[-1])
    if pad_options:
        pads = [(0, 0)] * (data.ndim - 2) + [(bi // 2 + 1, bi // 2), (bj // 2 + 1, bj // 2)]
        data = np.pad(data, pads, **pad_options)
    if cumsum:
        data = data.cumsum(-2).cumsum(-1)
    m, n = data.shape[-2] - bi, data.shape[-1] - bj
    return (data[..., bi:bi + m, bj:bj + n] - data[..., :m, bj:bj + n]
            + data[..., :m, :n] - data[..., bi:bi + m, :n])


def _nbhood_sum(data, method, nb_size, kernel, max_extreme=None):
    """improver/nbhood/nbhood.py:319-409 (`_do_nbhood_sum`), including the
    bounding-box trimming of all-zero / all-one margins."""
    full_shape = data.shape
    ny, nx = full_shape
    best = dict(size=data.size, extreme=0, fill=0, y0=0, y1=ny, x0=0, x1=nx)
    half = nb_size // 2
    for extreme, fill in ((0, 0), (1, max_extreme)):
        if fill is None or np.iscomplexobj(data):
            continue
        idx = np.argwhere(data != extreme)
        if idx.size == 0:
            y0 = y1 = x0 = x1 = 0
        else:
            (y0, x0), (y1, x1) = idx.min(0), idx.max(0) + 1
            y0, y1 = max(0, y0 - half), min(ny, y1 + half)
            x0, x1 = max(0, x0 - half), min(nx, x1 + half)
        size = (y1 - y0) * (x1 - x0)
        if size < best["size"]:
            best = dict(size=size, extreme=extreme, fill=fill, y0=y0, y1=y1, x0=x0, x1=x1)
    trimmed = best["size"] != data.size
    if trimmed:
        if isinstance(best["fill"], np.ndarray):
            background = best["fill"].astype(data.dtype)
        else:
            background = np.full(full_shape, best["fill"], dtype=data.dtype)
    if best["size"]:
        sub = data[best["y0"]:best["y1"], best["x0"]:best["x1"]]
        if method == "square":
            sub = boxsum(sub, nb_size, mode="constant", constant_values=best["extreme"])
        else:
            sub = _correlate(sub, kernel, mode="nearest")
    else:
        sub = background
    if sub.shape != full_shape:
        background[best["y0"]:best["y1"], best["x0"]:best["x1"]] = sub
        sub = background
    return sub


def neighbourhood(data, mask=None, method="square", cells=1, weighted_mode=False,
                  sum_only=False, re_mask=True):
    """improver/nbhood/nbhood.py:244-317 (`_calculate_neighbourhood`) for one
    2-D slice.  `cells` is the radius in grid cells.  Returns float32 ndarray
    or MaskedArray (complex64 for complex input)."""
    if method == "circular":
        kernel = circular_kernel(cells, weighted_mode)
        nb_size = max(kernel.shape)
    else:
        kernel = None
        nb_size = 2 * cells + 1
    if not sum_only:
        lo, hi = np.nanmin(data), np.nanmax(data)
    invalid = mask == 0 if mask is not None else np.False_
    if isinstance(data, np.ma.MaskedArray):
        invalid = invalid | data.mask
        data = data.data
    cplx = np.iscomplexobj(data)
    work = np.array(data, dtype=np.complex128 if cplx else np.float64)
    ok = np.ones(work.shape, dtype=np.float32 if method == "circular" else np.int64)
    ok[invalid] = 0
    work[invalid] = 0
    if sum_only:
        total = _nbhood_sum(work, method, nb_size, kernel)
    else:
        area = _nbhood_sum(ok, method, nb_size, kernel)
        total = _nbhood_sum(work, method, nb_size, kernel, max_extreme=area.astype(work.dtype))
        with np.errstate(divide="ignore", invalid="ignore"):
            total = total / area
        total[area == 0] = np.nan
        total = total.clip(lo, hi)
    if re_mask:
        total = np.ma.masked_array(total, invalid, copy=False)
    return total.astype(np.complex64 if cplx else np.float32)


def neighbourhood_cube(data, mask=None, **kw):
    """improver/nbhood/nbhood.py:457-463: every leading index is an independent
    2-D slice sharing the same 2-D mask.  Returns (values, mask-or-None)."""
    lead = data.shape[:-2]
    flat = data.reshape((-1,) + data.shape[-2:])
    vals = np.empty(flat.shape, dtype=F32)
    msk = None
    for i in range(flat.shape[0]):
        res = neighbourhood(flat[i], mask, **kw)
        vals[i] = np.ma.getdata(res)
        if isinstance(res, np.ma.MaskedArray):
            if msk is None:
                msk = np.zeros(flat.shape, dtype=bool)
            msk[i] = np.ma.getmaskarray(res)
    vals = vals.reshape(lead + data.shape[-2:])
    if msk is not None:
        msk = msk.reshape(vals.shape)
    return vals, msk


def land_sea_neighbourhood(data, land_mask, **kw):
    """improver/cli/nbhood_land_and_sea.py:131-184 (no topographic bands): one
    masked pass over land points, one over sea points, `filled(0)` sum."""
    land = (land_mask != 0).astype(np.int8)
    sea = (land_mask == 0).astype(np.int8)
    kw = dict(kw, re_mask=True)
    lv, lm = neighbourhood_cube(data, land, **kw)
    sv, sm = neighbourhood_cube(data, sea, **kw)
    return np.where(lm, 0, lv) + np.where(sm, 0, sv)


def neighbourhood_with_zone_masks(data, masks, weights=None, **kw):
    """improver/nbhood/use_nbhood.py:194-262: every 2-D slice is neighbourhooded
    once per zone mask (re_mask=False); with `weights` (zone, y, x) the zone axis
    is collapsed by a NaN-aware weighted mean (np.ma.average renormalises)."""
    lead = data.shape[:-2]
    flat = data.reshape((-1,) + data.shape[-2:])
    out = np.empty((flat.shape[0], masks.shape[0]) + data.shape[-2:], dtype=F32)
    kw = dict(kw, re_mask=False)
    for i in range(flat.shape[0]):
        for z in range(masks.shape[0]):
            out[i, z] = neighbourhood(flat[i], masks[z], **kw)
    if weights is not None:
        collapsed = np.ma.average(np.ma.masked_invalid(out), axis=1,
                                  weights=np.broadcast_to(weights, out.shape))
        res = np.ma.getdata(collapsed).astype(F32)
        res[np.ma.getmaskarray(collapsed)] = np.nan
        return res.reshape(lead + data.shape[-2:])
    return out.reshape(lead + out.shape[1:])


# --------------------------------------------------------------------------- #
# Neighbourhood percentiles
# --------------------------------------------------------------------------- #
def rolling_window(arr, shape, writeable=False):
    """improver/utilities/neighbourhood_tools.py:13-66."""
    nwin = len(shape)
    if arr.ndim < nwin:
        raise ValueError(
            "Number of dimensions of the input array must be greater than or "
            "equal to the length of the neighbourhood shape used for "
            "constructing rolling window neighbourhoods.")
    new_shape = (*arr.shape[:-nwin],
                 *(a - w + 1 for a, w in zip(arr.shape[-nwin:], shape)), *shape)
    if any(s <= 0 for s in new_shape):
        raise RuntimeError(
            "The calculated shape of the output array view contains a "
            "dimension that is negative or zero. Each dimension of the "
            "neighbourhood shape must be less than or equal to the "
            "corresponding dimension of the input array.")
    return np.lib.stride_tricks.as_strided(
        arr, shape=new_shape, strides=arr.strides + arr.strides[-nwin:], writeable=writeable)


def pad_and_roll(arr, shape, **kwargs):
    """improver/utilities/neighbourhood_tools.py:69-101."""
    writeable = kwargs.pop("writeable", False)
    pads = [(0, 0)] * (arr.ndim - len(shape)) + [(d // 2, d // 2) for d in shape]
    return rolling_window(np.pad(arr, pads, **kwargs), shape, writeable=writeable)


def neighbourhood_percentiles(data2d, cells: int, percentiles) -> np.ndarray:
    """improver/nbhood/nbhood.py:649-667: mean-padded circular windows reduced
    with np.percentile.  Returns float32 (P, y, x)."""
    kernel = circular_kernel(cells, False)
    sel = kernel > 0
    windows = pad_and_roll(data2d, kernel.shape, mode="mean", stat_length=max(kernel.shape) // 2)
    pcts = np.array(percentiles, dtype=F32)
    out = np.empty((len(pcts),) + data2d.shape, dtype=F32)
    for row_windows, row_out in zip(windows, out.swapaxes(0, 1)):
        np.percentile(row_windows[..., sel], pcts, axis=-1, out=row_out, overwrite_input=True)
    return out


def neighbourhood_percentiles_cube(data, cells: int, percentiles) -> np.ndarray:
    """improver/nbhood/nbhood.py:697-726: leading dims are independent; output
    order (leading..., percentile, y, x) — i.e. realization, percentile, y, x."""
    lead = data.shape[:-2]
    flat = data.reshape((-1,) + data.shape[-2:])
    res = np.stack([neighbourhood_percentiles(s, cells, percentiles) for s in flat])
    return res.reshape(lead + res.shape[1:])


# --------------------------------------------------------------------------- #
# Host-side helpers and the fused chain
# --------------------------------------------------------------------------- #
def cells_from_distance(distance_m: float, spacing_m: float) -> int:
    """improver/utilities/spatial.py:120-161 (truncation, positivity checks)."""
    if distance_m <= 0:
        raise ValueError("Please specify a positive distance in metres. "
                         f"{distance_m} is not a valid distance")
    cells = int(distance_m / abs(spacing_m))
    if cells == 0:
        raise ValueError(f"Distance of {distance_m}m gives zero cell extent")
    return cells


def radii_for_lead_times(fp_hours, lead_times, radii):
    """improver/nbhood/nbhood.py:132-149 (`_find_radii`)."""
    return np.interp(fp_hours, lead_times, radii)


def realization_mean(data):
    """improver/utilities/cube_manipulation.py:93-131 (iris MEAN over the
    leading realization axis == np.mean in the data dtype)."""
    return np.mean(data, axis=0, dtype=data.dtype)


def threshold_square_mean(data, thresholds, bounds, comparison, cells,
                          mask=None, method="square", weighted_mode=False):
    """The operational chain the fused kernel replaces: Threshold per
    realization (threshold.py:643-713, no collapse) -> NeighbourhoodProcessing
    per (realization, threshold) slice (nbhood.py:457-463) -> realization mean.

    data float32 (R, L, y, x) -> float32 (L, T, y, x)."""
    R, L = data.shape[:2]
    out = np.empty((L, len(thresholds)) + data.shape[2:], dtype=F32)
    for l in range(L):
        for t, (thr, bnd) in enumerate(zip(thresholds, bounds)):
            members = np.empty((R,) + data.shape[2:], dtype=F32)
            for r in range(R):
                tv = truth_value(data[r, l], thr, bnd, comparison)
                res = neighbourhood(tv, mask, method=method, cells=cells,
                                    weighted_mode=weighted_mode, re_mask=False)
                members[r] = np.ma.getdata(res)
            out[l, t] = realization_mean(members)
    return out
