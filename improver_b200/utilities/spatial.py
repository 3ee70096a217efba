"""Grid helpers (improver/utilities/spatial.py:42-161)."""
from __future__ import annotations

import warnings
from typing import Union

import numpy as np


def _points_in(cube, axis: str, units: str):
    coord = cube.coord(axis=axis).copy()
    coord.convert_units(units)
    return coord


def calculate_grid_spacing(cube, units: str, axis: str = "x", rtol: float = 1.0e-5) -> float:
    """Grid spacing of a spatial axis (positive); spatial.py:69-117."""
    if rtol < 0:
        raise ValueError(f"rtol must be non-negative, got {rtol}")
    if rtol > 0.01:
        warnings.warn(f"rtol={rtol} exceeds recommended maximum of 0.01."
                      "Large tolerance values may allow non-uniform grids to pass validation.")
    coord = _points_in(cube, axis, units)
    diffs = np.abs(np.diff(coord.points))
    diffs_mean = np.mean(diffs)
    if not np.allclose(diffs, diffs_mean, rtol=rtol, atol=0.0):
        raise ValueError("Coordinate {} points are not equally spaced".format(coord.name()))
    return diffs_mean


def check_if_grid_is_equal_area(cube, require_equal_xy_spacing: bool = True) -> None:
    """spatial.py:42-66: equally spaced x and y axes (in metres), equal spacing."""
    x_diff = calculate_grid_spacing(cube, "metres", axis="x")
    y_diff = calculate_grid_spacing(cube, "metres", axis="y")
    if require_equal_xy_spacing and not np.isclose(x_diff, y_diff):
        raise ValueError("Grid does not have equal spacing in x and y dimensions")


def distance_to_number_of_grid_cells(cube, distance: float, axis: str = "x",
                                     return_int: bool = True) -> Union[float, int]:
    """spatial.py:120-161: int(distance / spacing) (truncation)."""
    d_error = f"Distance of {distance}m"
    if distance <= 0:
        raise ValueError(f"Please specify a positive distance in metres. {d_error}")
    grid_spacing_metres = calculate_grid_spacing(cube, "m", axis=axis)
    grid_cells = distance / abs(grid_spacing_metres)
    if return_int:
        grid_cells = int(np.asarray(grid_cells).item())
        if grid_cells == 0:
            raise ValueError(f"{d_error} gives zero cell extent")
    return grid_cells
