"""Neighbourhood processing plugins — drop-in surface of
improver/nbhood/nbhood.py backed by the CUDA engine.

Same class names, constructor arguments, error types and messages as the
reference; the per-slice numpy/scipy loops (nbhood.py:457-463, 697-704) are
replaced by one batched launch over all leading dimensions of the cube.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np

from .. import BasePlugin, PostProcessingPlugin
from ..constants import DEFAULT_PERCENTILES
from ..utilities.spatial import check_if_grid_is_equal_area, distance_to_number_of_grid_cells
from . import _as_iterable, radius_by_lead_time


def check_radius_against_distance(cube, radius: float) -> None:
    """nbhood.py:32-52: the radius may not exceed half the domain diagonal."""
    axes = []
    for axis in ["x", "y"]:
        coord = cube.coord(axis=axis).copy()
        coord.convert_units("metres")
        axes.append((max(coord.points) - min(coord.points)))
    max_allowed = np.sqrt(axes[0] ** 2 + axes[1] ** 2) * 0.5
    if radius > max_allowed:
        raise ValueError(f"Distance of {radius}m exceeds max domain " f"distance of {max_allowed}m")


def circular_kernel(ranges: int, weighted_mode: bool) -> np.ndarray:
    """nbhood.py:55-92: disc of radius `ranges` cells (ones) or the
    (r^2 - d^2) / r^2 weighting; host-side description of what the CUDA kernel
    evaluates analytically."""
    ranges = int(ranges)
    off = np.arange(-ranges, ranges + 1)
    d2 = (off[:, None] ** 2 + off[None, :] ** 2).astype(float)
    area = ranges * ranges
    kernel = np.ones((2 * ranges + 1, 2 * ranges + 1))
    if weighted_mode:
        kernel[:] = (area - d2) / area
        kernel[kernel < 0.0] = 0.0
    else:
        kernel[d2 > area] = 0.0
    return kernel


def _forecast_period_hours(cube) -> np.ndarray:
    """improver/metadata/forecast_times.py:25 `forecast_period_coord` + hours."""
    coord = cube.coord("forecast_period").copy()
    coord.convert_units("hours")
    return np.asarray(coord.points)


def _yx_last(cube):
    """Axes permutation that brings (y, x) last, and its inverse."""
    ydim = cube.coord_dims(cube.coord(axis="y"))[0]
    xdim = cube.coord_dims(cube.coord(axis="x"))[0]
    ndim = len(cube.shape)
    perm = [d for d in range(ndim) if d not in (ydim, xdim)] + [ydim, xdim]
    inv = np.argsort(perm)
    return perm, inv


def _to_device(array, dtype=None):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(array))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda(non_blocking=True)


class BaseNeighbourhoodProcessing(PostProcessingPlugin):
    """nbhood.py:95-175: radii / lead-time handling and the NaN guard."""

    def __init__(self, radii: Union[float, List[float]], lead_times: Optional[List] = None) -> None:
        if isinstance(radii, list):
            self.radii = [float(x) for x in radii]
        else:
            self.radius = float(radii)
        self.lead_times = lead_times
        if self.lead_times is not None:
            if len(radii) != len(lead_times):
                msg = ("There is a mismatch in the number of radii "
                       "and the number of lead times. "
                       "Unable to continue due to mismatch.")
                raise ValueError(msg)

    def _find_radii(self, cube_lead_times=None):
        """nbhood.py:132-149."""
        return np.interp(cube_lead_times, self.lead_times, self.radii)

    def process(self, cube):
        """nbhood.py:151-175.  Unmasked NaNs raise ValueError; the check on the
        data itself is done by the kernels (status word), a cheap host check
        keeps the reference's fail-early behaviour for host arrays."""
        data = cube.core_data() if hasattr(cube, "core_data") else cube.data
        if isinstance(data, np.ndarray) and np.isnan(data).any():
            raise ValueError("Error: NaN detected in input cube data")
        if self.lead_times:
            self.radius = self._find_radii(cube_lead_times=_forecast_period_hours(cube))
        return cube


class NeighbourhoodProcessing(BaseNeighbourhoodProcessing):
    """nbhood.py:178-465: square / circular neighbourhood mean or sum with
    mask weighting and re-masking."""

    def __init__(self, neighbourhood_method: str, radii: Union[float, List[float]],
                 lead_times: Optional[List] = None, weighted_mode: bool = False,
                 sum_only: bool = False, re_mask: bool = True) -> None:
        super().__init__(radii, lead_times=lead_times)
        if neighbourhood_method in ["square", "circular"]:
            self.neighbourhood_method = neighbourhood_method
        else:
            msg = "{} is not a valid neighbourhood_method.".format(neighbourhood_method)
            raise ValueError(msg)
        if weighted_mode and neighbourhood_method != "circular":
            msg = ("weighted_mode can only be used if neighbourhood_method is circular."
                   f" weighted_mode provided: {weighted_mode}, "
                   f"neighbourhood_method provided: {neighbourhood_method}.")
            raise ValueError(msg)
        self.weighted_mode = weighted_mode
        self.sum_only = sum_only
        self.re_mask = re_mask

    # -- array level ---------------------------------------------------------
    def _cells(self) -> int:
        if hasattr(self, "kernel") and self.neighbourhood_method == "circular":
            return (max(self.kernel.shape) - 1) // 2
        return (int(self.nb_size) - 1) // 2

    def _calculate_neighbourhood(self, data, mask=None):
        """Array-level entry (nbhood.py:244-317) for (..., y, x) data; `nb_size`
        (and `kernel` for circular) must be set like the reference requires."""
        from .. import engine

        if np.iscomplexobj(data):
            return self._calculate_neighbourhood_complex(data, mask)
        cells = self._cells()
        weighted = self.weighted_mode
        if self.neighbourhood_method == "circular" and hasattr(self, "kernel"):
            # the reference convolves with whatever `kernel` holds; the CUDA path
            # evaluates the two analytic kernels of circular_kernel()
            kernel = np.asarray(self.kernel)
            weighted = bool(np.any((kernel != 0) & (kernel != 1)))
            if not np.allclose(kernel, circular_kernel(cells, weighted)):
                raise NotImplementedError("only kernels produced by circular_kernel() are supported")
        in_mask = None
        raw = data
        if isinstance(data, np.ma.MaskedArray):
            in_mask = np.ma.getmaskarray(data)
            raw = data.data
        dev = _to_device(np.asarray(raw, dtype=np.float32))
        m2 = None
        if mask is not None:
            m2 = _to_device((np.asarray(mask) != 0).astype(np.uint8))
        mnd = _to_device(in_mask.astype(np.uint8)) if in_mask is not None else None
        out, omask, status = engine.neighbourhood(
            dev, cells, method=self.neighbourhood_method, mask2d=m2, mask_nd=mnd,
            weighted_mode=weighted, sum_only=self.sum_only, want_mask=self.re_mask)
        res = out.cpu().numpy()
        engine.check_status(status)
        if self.re_mask:
            res = np.ma.masked_array(res, omask.cpu().numpy().astype(bool), copy=False)
        return res

    def _calculate_neighbourhood_complex(self, data, mask=None):
        """Complex input (wind directions as unit vectors, nbhood.py:277-280): the neighbourhood
        sum is linear, so the real part, the imaginary part and the valid-point count are three
        `sum_only` launches of the float kernels and the mean is their quotient.  The reference
        never trims complex data against ones (nbhood.py:355-357), so no fix-up applies; its
        lexicographic clip to [nanmin, nanmax] is restated with numpy."""
        from .. import engine

        cells = self._cells()
        in_mask = np.ma.getmaskarray(data) if isinstance(data, np.ma.MaskedArray) else None
        raw = np.asarray(np.ma.getdata(data))
        m2 = _to_device((np.asarray(mask) != 0).astype(np.uint8)) if mask is not None else None
        mnd = _to_device(in_mask.astype(np.uint8)) if in_mask is not None else None
        parts = []
        for comp in (raw.real, raw.imag, np.ones(raw.shape, np.float32)):
            out, omask, status = engine.neighbourhood(
                _to_device(np.ascontiguousarray(comp, dtype=np.float32)), cells, method=self.neighbourhood_method,
                mask2d=m2, mask_nd=mnd, weighted_mode=self.weighted_mode, sum_only=True, want_mask=True)
            parts.append(out.cpu().numpy().astype(np.float64))
            engine.check_status(status)
        res = parts[0] + 1j * parts[1]
        if not self.sum_only:
            valid = raw if in_mask is None else raw[~in_mask]
            min_val, max_val = np.nanmin(valid), np.nanmax(valid)
            with np.errstate(divide="ignore", invalid="ignore"):
                res = res / parts[2]
            res[parts[2] == 0] = np.nan
            res = res.clip(min_val, max_val)
        res = res.astype(np.complex64)
        if self.re_mask:
            res = np.ma.masked_array(res, omask.cpu().numpy().astype(bool), copy=False)
        return res

    # -- cube level ------------------------------------------------------------
    def process(self, cube, mask_cube=None):
        """nbhood.py:411-465."""
        super().process(cube)
        check_if_grid_is_equal_area(cube)
        check_radius_against_distance(cube, self.radius)
        grid_cells = distance_to_number_of_grid_cells(cube, self.radius)
        if self.neighbourhood_method == "circular":
            self.kernel = circular_kernel(grid_cells, self.weighted_mode)
            self.nb_size = max(self.kernel.shape)
        else:
            self.nb_size = 2 * grid_cells + 1
        try:
            mask_cube_data = mask_cube.data
        except AttributeError:
            mask_cube_data = None

        perm, inv = _yx_last(cube)
        # deferred (device-resident) data: record the operation (improver_b200.lazy)
        from ..lazy import Deferred, NbhoodOp

        core = cube.core_data() if hasattr(cube, "core_data") else None
        if isinstance(core, Deferred) and list(perm) == list(range(len(perm))):
            return cube.copy(data=NbhoodOp(core, self.neighbourhood_method, grid_cells, self.weighted_mode,
                                           self.sum_only, mask_cube_data, self.re_mask))
        data = cube.data
        masked_in = isinstance(data, np.ma.MaskedArray)
        arr = np.ma.transpose(data, perm) if masked_in else np.transpose(data, perm)
        result = self._calculate_neighbourhood(arr, mask_cube_data)
        if isinstance(result, np.ma.MaskedArray):
            # re_mask: a MaskedArray even without any mask source (all-False mask, nbhood.py:314-317)
            result = np.ma.transpose(result, inv)
        else:
            result = np.transpose(result, inv)
        return cube.copy(data=result)


class GeneratePercentilesFromANeighbourhood(BaseNeighbourhoodProcessing):
    """nbhood.py:468-757: percentiles of a circular neighbourhood."""

    def __init__(self, radii: Union[float, List[float]], lead_times: Optional[List] = None,
                 percentiles: Union[float, List[float]] = DEFAULT_PERCENTILES) -> None:
        super().__init__(radii, lead_times=lead_times)
        self.percentiles = tuple(_as_iterable(percentiles))

    def pad_and_unpad_cube(self, slice_2d, kernel: np.ndarray):
        """nbhood.py:500-669 for one 2-D slice (cube or array): returns the
        (percentile, y, x) array (a cube copy when a cube is given)."""
        from .. import engine

        data = slice_2d.data if hasattr(slice_2d, "data") and not isinstance(slice_2d, np.ndarray) else slice_2d
        cells = (max(kernel.shape) - 1) // 2
        out, status = engine.neighbourhood_percentiles(
            _to_device(np.asarray(data, dtype=np.float32)), cells, self.percentiles)
        res = out.cpu().numpy()
        engine.check_status(status)
        if isinstance(slice_2d, np.ndarray):
            return res
        return self.make_percentile_cube(slice_2d, res)

    def process(self, cube):
        """nbhood.py:671-728: output dims (realization, percentile, other..., y, x)."""
        from .. import engine
        from ..cube import Coord

        super().process(cube)
        if np.ma.is_masked(cube.data):
            msg = ("The use of masked input cubes is not yet implemented in"
                   " the GeneratePercentilesFromANeighbourhood plugin.")
            raise NotImplementedError(msg)
        check_if_grid_is_equal_area(cube)
        grid_cell = distance_to_number_of_grid_cells(cube, self.radius)
        check_radius_against_distance(cube, self.radius)

        perm, _ = _yx_last(cube)
        arr = np.transpose(np.asarray(cube.data, dtype=np.float32), perm)
        out, status = engine.neighbourhood_percentiles(_to_device(arr), grid_cell, self.percentiles)
        res = out.cpu().numpy()               # (lead..., P, y, x)
        engine.check_status(status)
        # dimension order of the reference: realization first, then percentile, then the rest
        lead_dims = [d for d in perm[:-2]]
        names = []
        for d in lead_dims:
            found = [c for c in cube.dim_coords if cube.coord_dims(c) == (d,)]
            names.append(found[0].name() if found else None)
        n_lead = len(lead_dims)
        order = list(range(n_lead))
        if "realization" in names:
            ridx = names.index("realization")
            order = [ridx] + [i for i in order if i != ridx]
            front = 1
        else:
            front = 0
        # res axes: lead(0..n_lead-1), P, y, x -> [realization], P, other lead, y, x
        axes = order[:front] + [n_lead] + order[front:] + [n_lead + 1, n_lead + 2]
        res = np.transpose(res, axes)
        result = cube.copy(data=res) if hasattr(cube, "_dim_coords") else None
        if result is None:
            return res
        # rebuild dimension coordinates for the double
        result._dim_coords = []
        pos = 0
        ordered_lead = [lead_dims[i] for i in order]
        lead_iter = iter(ordered_lead[:front])
        for d in lead_iter:
            for c, cd in cube._dim_coords:
                if cd == d:
                    result._dim_coords.append((c.copy(), pos))
            pos += 1
        pct = Coord(np.array(self.percentiles, dtype=np.float32), "percentile", "%")
        result._dim_coords.append((pct, pos))
        pos += 1
        for d in ordered_lead[front:]:
            for c, cd in cube._dim_coords:
                if cd == d:
                    result._dim_coords.append((c.copy(), pos))
            pos += 1
        for axis in ("y", "x"):
            result._dim_coords.append((cube.coord(axis=axis).copy(), pos))
            pos += 1
        return result

    def make_percentile_cube(self, cube, data=None):
        """nbhood.py:730-757: template cube with a leading percentile dimension."""
        from ..cube import Coord

        ny_nx = tuple(cube.shape)
        if data is None:
            data = np.broadcast_to(np.asarray(cube.data), (len(self.percentiles),) + ny_nx).copy()
        new = cube.copy(data=data)
        if hasattr(new, "_dim_coords"):
            new._dim_coords = [(c, d + 1) for c, d in new._dim_coords]
            new._dim_coords.insert(0, (Coord(np.array(self.percentiles, dtype=np.float32),
                                              "percentile", "%"), 0))
        return new


class MetaNeighbourhood(BasePlugin):
    """nbhood.py:760-894: probabilities / percentiles dispatcher."""

    def __init__(self, neighbourhood_output: str, radii, lead_times=None,
                 neighbourhood_shape: str = "square", degrees_as_complex: bool = False,
                 weighted_mode: bool = False, area_sum: bool = False,
                 percentiles=DEFAULT_PERCENTILES, halo_radius: Optional[float] = None) -> None:
        self._neighbourhood_output = neighbourhood_output
        self._neighbourhood_shape = neighbourhood_shape
        self._radius_or_radii, self._lead_times = radius_by_lead_time(radii, lead_times)
        self._degrees_as_complex = degrees_as_complex
        self._weighted_mode = weighted_mode
        self._area_sum = area_sum
        self._percentiles = percentiles
        self._halo_radius = halo_radius
        if neighbourhood_output == "percentiles":
            if weighted_mode:
                raise RuntimeError("weighted_mode cannot be used with" 'neighbourhood_output="percentiles"')
            if degrees_as_complex:
                raise RuntimeError("Cannot generate percentiles from complex numbers")
        if neighbourhood_shape == "circular":
            if degrees_as_complex:
                raise RuntimeError("Cannot process complex numbers with circular neighbourhoods")

    def process(self, cube, mask=None):
        """nbhood.py:852-894."""
        if self._degrees_as_complex:
            # improver/utilities/complex_conversion.py:10-36 (deg_to_complex, radius 1)
            angle_rad = np.deg2rad(cube.data)
            cube = cube.copy(data=np.cos(angle_rad) + 1j * np.sin(angle_rad))
        if self._neighbourhood_output == "probabilities":
            result = NeighbourhoodProcessing(
                self._neighbourhood_shape, self._radius_or_radii, lead_times=self._lead_times,
                weighted_mode=self._weighted_mode, sum_only=self._area_sum, re_mask=True,
            )(cube, mask_cube=mask)
        elif self._neighbourhood_output == "percentiles":
            result = GeneratePercentilesFromANeighbourhood(
                self._radius_or_radii, lead_times=self._lead_times, percentiles=self._percentiles,
            )(cube)
        else:
            raise ValueError(f"Unknown neighbourhood_output: {self._neighbourhood_output}")
        if self._degrees_as_complex:
            # complex_conversion.py:39-69 (complex_to_deg): angle in [0, 360), float32
            data = result.data
            angle = np.mod(np.float32(np.angle(np.ma.getdata(data), deg=True)), 360).astype(np.float32)
            if isinstance(data, np.ma.MaskedArray):
                angle = np.ma.masked_array(angle, np.ma.getmaskarray(data))
            result = result.copy(data=angle)
        if self._halo_radius is not None:
            result = remove_cube_halo(result, self._halo_radius)
        return result


def remove_cube_halo(cube, halo_radius: float):
    """improver/utilities/pad_spatial.py:222-300: drop `halo_radius` metres of rows / columns from
    every edge of every 2-D slice and trim the spatial coordinates accordingly."""
    wx = distance_to_number_of_grid_cells(cube, halo_radius, axis="x")
    wy = distance_to_number_of_grid_cells(cube, halo_radius, axis="y")
    ydim = cube.coord_dims(cube.coord(axis="y"))[0]
    xdim = cube.coord_dims(cube.coord(axis="x"))[0]
    index = [slice(None)] * len(cube.shape)
    index[ydim] = slice(wy, -wy if wy else None)
    index[xdim] = slice(wx, -wx if wx else None)
    result = cube.copy(data=cube.data[tuple(index)])
    for axis, w in (("y", wy), ("x", wx)):
        coord = result.coord(axis=axis)
        coord.points = coord.points[w:len(coord.points) - w if w else None]
    return result
