"""Land / sea neighbourhood processing — the array logic of
improver/cli/nbhood_land_and_sea.py:84-186 without topographic bands: land
points are neighbourhooded with the land mask, sea points with the sea mask
(`re_mask=True`), and the two masked results are recombined with `filled(0)`.
This is what BASELINE.json config 2 ("land-sea mask weighted, re_mask") runs.
"""
from __future__ import annotations

import numpy as np

from .nbhood import NeighbourhoodProcessing


def nbhood_land_and_sea(cube, land_sea_mask, neighbourhood_shape: str = "square", radii=None,
                        lead_times=None, area_sum: bool = False, weighted_mode: bool = False):
    """Returns a cube like `cube` whose land points were smoothed using land
    points only and whose sea points using sea points only
    (cli/nbhood_land_and_sea.py:131-184)."""
    mask = np.asarray(land_sea_mask.data)
    land_only = type("MaskCube", (), {"data": np.where(mask > 0, 1, 0).astype(np.int8)})()
    sea_only = type("MaskCube", (), {"data": np.where(mask > 0, 0, 1).astype(np.int8)})()
    kwargs = dict(lead_times=lead_times, weighted_mode=weighted_mode, sum_only=area_sum, re_mask=True)
    result_land = result_sea = None
    if land_only.data.max() > 0:
        result_land = NeighbourhoodProcessing(neighbourhood_shape, radii, **kwargs)(cube, land_only)
    if sea_only.data.max() > 0:
        result_sea = NeighbourhoodProcessing(neighbourhood_shape, radii, **kwargs)(cube, sea_only)
    if result_land is None:
        return result_sea
    if result_sea is None:
        return result_land
    combined = np.ma.filled(result_land.data, 0) + np.ma.filled(result_sea.data, 0)
    return result_land.copy(data=combined.astype(np.float32))
