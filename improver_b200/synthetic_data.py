"""Synthetic cubes shaped like improver.synthetic_data.set_up_test_cubes
(improver/synthetic_data/set_up_test_cubes.py:543-844) built on the cube
double: equal-area 2 km grid by default (improver/grids.py:22-40)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .cube import Coord, Cube


def _spatial_coords(ny: int, nx: int, spacing: float, equal_area: bool = True):
    if equal_area:
        y = Coord(np.arange(ny, dtype=np.float32) * spacing - 100000.0, "projection_y_coordinate", "m", "y")
        x = Coord(np.arange(nx, dtype=np.float32) * spacing - 400000.0, "projection_x_coordinate", "m", "x")
    else:
        y = Coord(np.linspace(-20, 20, ny, dtype=np.float32), "latitude", "degrees", "y")
        x = Coord(np.linspace(40, 80, nx, dtype=np.float32), "longitude", "degrees", "x")
    return y, x


def set_up_variable_cube(data, name: str = "air_temperature", units: str = "K",
                         spatial_grid: str = "equalarea", grid_spacing: float = 2000.0,
                         realizations: Optional[Sequence[int]] = None,
                         forecast_period_hours: Optional[float] = 4.0,
                         attributes: Optional[dict] = None) -> Cube:
    """Cube with dims ([realization,] y, x) and a scalar forecast_period."""
    data = np.asarray(data) if not isinstance(data, np.ma.MaskedArray) else data
    ny, nx = data.shape[-2:]
    y, x = _spatial_coords(ny, nx, grid_spacing, spatial_grid == "equalarea")
    dim_coords = []
    lead = data.ndim - 2
    if lead == 1:
        rl = np.arange(data.shape[0], dtype=np.int32) if realizations is None else np.asarray(realizations)
        dim_coords.append((Coord(rl, "realization", "1"), 0))
    elif lead > 1:
        raise ValueError("data must be 2-D or 3-D")
    dim_coords += [(y, lead), (x, lead + 1)]
    scalars = []
    if forecast_period_hours is not None:
        scalars.append(Coord([int(forecast_period_hours * 3600)], "forecast_period", "seconds"))
    return Cube(data, name, units, dim_coords, scalars, attributes)


def set_up_probability_cube(data, thresholds, variable_name: str = "air_temperature",
                            threshold_units: str = "K", spatial_grid: str = "equalarea",
                            grid_spacing: float = 2000.0, forecast_period_hours: Optional[float] = 4.0,
                            attributes: Optional[dict] = None) -> Cube:
    """Cube with dims (threshold, y, x) named probability_of_X_above_threshold."""
    data = np.asarray(data) if not isinstance(data, np.ma.MaskedArray) else data
    ny, nx = data.shape[-2:]
    y, x = _spatial_coords(ny, nx, grid_spacing, spatial_grid == "equalarea")
    thr = Coord(np.asarray(thresholds, dtype=np.float32), variable_name, threshold_units, var_name="threshold",
                attributes={"spp__relative_to_threshold": "greater_than"})
    scalars = []
    if forecast_period_hours is not None:
        scalars.append(Coord([int(forecast_period_hours * 3600)], "forecast_period", "seconds"))
    return Cube(data, f"probability_of_{variable_name}_above_threshold", "1",
                [(thr, 0), (y, 1), (x, 2)], scalars, attributes)
