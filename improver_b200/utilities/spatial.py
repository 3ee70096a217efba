"""Grid helpers with the semantics of improver/utilities/spatial.py:42-161:
grid spacing from the cube's x / y coordinates, the equal-area check and the
distance -> grid-cell conversion (truncating)."""
from __future__ import annotations

import warnings
from typing import Union

import numpy as np


def calculate_grid_spacing(cube, units: str, axis: str = "x", rtol: float = 1.0e-5) -> float:
    """Mean absolute point spacing of the `axis` coordinate in `units`; raises
    if the points are not equally spaced within `rtol` (spatial.py:69-117)."""
    if rtol < 0:
        raise ValueError(f"rtol must be non-negative, got {rtol}")
    if rtol > 0.01:
        warnings.warn(f"rtol={rtol} exceeds recommended maximum of 0.01."
                      "Large tolerance values may allow non-uniform grids to pass validation.")
    axis_coord = cube.coord(axis=axis).copy()
    axis_coord.convert_units(units)
    steps = np.abs(np.diff(axis_coord.points))
    spacing = np.mean(steps)
    if not np.allclose(steps, spacing, rtol=rtol, atol=0.0):
        raise ValueError("Coordinate {} points are not equally spaced".format(axis_coord.name()))
    return spacing


def check_if_grid_is_equal_area(cube, require_equal_xy_spacing: bool = True) -> None:
    """Both axes equally spaced in metres and (optionally) with the same
    spacing (spatial.py:42-66)."""
    spacing = {ax: calculate_grid_spacing(cube, "metres", axis=ax) for ax in ("x", "y")}
    if require_equal_xy_spacing and not np.isclose(spacing["x"], spacing["y"]):
        raise ValueError("Grid does not have equal spacing in x and y dimensions")


def distance_to_number_of_grid_cells(cube, distance: float, axis: str = "x",
                                     return_int: bool = True) -> Union[float, int]:
    """distance / spacing; truncated to int unless return_int is False
    (spatial.py:120-161)."""
    what = f"Distance of {distance}m"
    if distance <= 0:
        raise ValueError(f"Please specify a positive distance in metres. {what}")
    cells = distance / abs(calculate_grid_spacing(cube, "m", axis=axis))
    if not return_int:
        return cells
    cells = int(np.asarray(cells).item())
    if cells == 0:
        raise ValueError(f"{what} gives zero cell extent")
    return cells


def _vicinity_name(name: str) -> str:
    """`probability_of_X_above_threshold` -> `probability_of_X_in_vicinity_above_threshold`,
    anything else gets the `_in_vicinity` suffix (spatial.py:1009-1023)."""
    for rel in ("_above_threshold", "_below_threshold", "_between_thresholds"):
        if name.startswith("probability_of_") and name.endswith(rel):
            return name[: -len(rel)] + "_in_vicinity" + rel
    return f"{name}_in_vicinity"


class OccurrenceWithinVicinity:
    """Maximum / minimum of a field within square vicinities of every point
    (improver/utilities/spatial.py:1077-1255), backed by the CUDA vicinity kernel
    (nbh_vicinity_f32).  Constructor arguments, validation order and messages follow
    :1105-1176; the "mean" and "std" operators are not part of the CUDA path."""

    def __init__(self, radii=None, grid_point_radii=None, land_mask_cube=None, operator: str = "max",
                 apply_cell_method: bool = True) -> None:
        if radii:
            radii = [float(x) for x in radii]
        if radii and grid_point_radii:
            raise ValueError("Vicinity processing requires that only one of radii or "
                             "grid_point_radii should be set")
        if not radii and not grid_point_radii:
            raise ValueError("Vicinity processing requires that one of radii or "
                             "grid_point_radii should be set to a non-zero value")
        if (radii and any(np.array(radii) < 0)) or (grid_point_radii and any(np.array(grid_point_radii) < 0)):
            raise ValueError("Vicinity processing requires only positive vicinity radii")
        self.radii = radii if radii else grid_point_radii
        self.native_grid_point_radius = False if radii else True
        if land_mask_cube:
            if land_mask_cube.name() != "land_binary_mask":
                raise ValueError(f"Expected land_mask_cube to be called land_binary_mask, "
                                 f"not {land_mask_cube.name()}")
            self.land_mask = np.where(np.asarray(land_mask_cube.data) >= 0.5, True, False)
        else:
            self.land_mask = None
        self.land_mask_cube = land_mask_cube
        self.apply_cell_method = apply_cell_method
        if operator not in ("max", "min", "mean", "std"):
            raise ValueError("Unsupported operator to apply over vicinity.")
        if operator in ("mean", "std"):
            raise NotImplementedError("the mean / std vicinity operators are outside the CUDA hot path")
        self.operator = operator
        self.cell_method = {"max": "maximum", "min": "minimum"}[operator]

    def __call__(self, cube, new_name: str = None):
        return self.process(cube, new_name)

    def process(self, cube, new_name: str = None):
        import torch

        from .. import engine
        from ..cube import Coord

        if self.land_mask is not None and tuple(self.land_mask.shape) != tuple(cube.shape[-2:]):
            raise ValueError("Supplied cube do not have the same spatial coordinates and land mask")
        if not self.native_grid_point_radius:
            cells = [distance_to_number_of_grid_cells(cube, radius) for radius in self.radii]
        else:
            cells = [int(c) for c in self.radii]
        data = cube.data
        masked = np.ma.is_masked(data)
        raw = np.ascontiguousarray(np.ma.getdata(data), dtype=np.float32)
        mnd = torch.from_numpy(np.ascontiguousarray(np.ma.getmaskarray(data)).astype(np.uint8)).cuda() if masked else None
        lm = torch.from_numpy(self.land_mask.astype(np.uint8)).cuda() if self.land_mask is not None else None
        out, status = engine.extreme_within_vicinity(torch.from_numpy(raw).cuda(), cells, self.operator,
                                                     mask_nd=mnd, landmask=lm)
        res = out.cpu().numpy()                                    # (V, ..., y, x)
        engine.check_status(status)
        if masked:
            # masked points keep their value and stay masked (spatial.py:796-799)
            m = np.ma.getmaskarray(data)
            res = np.where(m[None], raw[None], res)
            res = np.ma.masked_array(res, mask=np.broadcast_to(m, res.shape))
        # radius axis follows the leading (non-spatial) dimensions (:1246-1251)
        lead = raw.ndim - 2
        res = np.ma.masked_array(np.moveaxis(np.ma.getdata(res), 0, lead),
                                 mask=np.moveaxis(np.ma.getmaskarray(res), 0, lead)) if masked \
            else np.moveaxis(res, 0, lead)
        squeeze = len(cells) == 1
        if squeeze:
            res = res.squeeze(axis=lead)
        result = cube.copy(data=res.astype(np.float32))
        points = np.array(self.radii, dtype=np.float32)
        if self.native_grid_point_radius:
            vic = Coord(points, "radius_of_vicinity", "1", attributes={
                "comment": "Units of 1 indicate radius of vicinity is defined "
                           "in grid points rather than physical distance"})
        else:
            vic = Coord(points, "radius_of_vicinity", "m")
        if hasattr(result, "_dim_coords"):
            if squeeze:
                result.add_aux_coord(vic)
            else:
                result._dim_coords = ([(c, d) for c, d in result._dim_coords if d < lead] + [(vic, lead)] +
                                      [(c, d + 1) for c, d in result._dim_coords if d >= lead])
        result.rename(new_name if new_name is not None else _vicinity_name(cube.name()))
        if self.apply_cell_method and hasattr(result, "attributes"):
            result.cell_methods = tuple(getattr(result, "cell_methods", ())) + (
                (self.cell_method, ("area",), "vicinity"),)
        return result
