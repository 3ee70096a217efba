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
