"""Land / sea neighbourhood processing — the array logic of
improver/cli/nbhood_land_and_sea.py:84-186 without topographic bands: land
points are neighbourhooded with the land mask, sea points with the sea mask
(`re_mask=True`), and the two masked results are recombined with `filled(0)`.
This is what BASELINE.json config 2 ("land-sea mask weighted, re_mask") runs.
"""
from __future__ import annotations

import numpy as np

from .nbhood import BaseNeighbourhoodProcessing, NeighbourhoodProcessing


def nbhood_land_and_sea(cube, land_sea_mask, neighbourhood_shape: str = "square", radii=None,
                        lead_times=None, area_sum: bool = False, weighted_mode: bool = False):
    """Returns a cube like `cube` whose land points were smoothed using land
    points only and whose sea points using sea points only
    (cli/nbhood_land_and_sea.py:131-184).  Both passes and their recombination are ONE native
    launch (NBH_LAND_SEA): every point is normalised over the points of its own surface type."""
    import torch

    from .. import engine
    from ..utilities.spatial import check_if_grid_is_equal_area, distance_to_number_of_grid_cells
    from .nbhood import _yx_last, check_radius_against_distance

    mask = np.asarray(land_sea_mask.data)
    land = np.where(mask > 0, 1, 0).astype(np.uint8)
    # constructor validation, radius interpolation and the NaN guard of the reference plugin
    plugin = NeighbourhoodProcessing(neighbourhood_shape, radii, lead_times=lead_times, weighted_mode=weighted_mode,
                                     sum_only=area_sum, re_mask=True)
    if land.max() == 0 or land.min() == 1:
        # a single surface type: the plain masked pass (cli/nbhood_land_and_sea.py:148-175)
        only = type("MaskCube", (), {"data": land if land.max() > 0 else 1 - land})()
        return plugin(cube, only)
    BaseNeighbourhoodProcessing.process(plugin, cube)
    check_if_grid_is_equal_area(cube)
    check_radius_against_distance(cube, plugin.radius)
    cells = distance_to_number_of_grid_cells(cube, plugin.radius)
    perm, inv = _yx_last(cube)
    data = cube.data
    if isinstance(data, np.ma.MaskedArray):
        raise NotImplementedError("land / sea neighbourhooding of masked-array input is outside the CUDA hot path")
    arr = np.ascontiguousarray(np.transpose(np.asarray(data, dtype=np.float32), perm))
    out, status = engine.neighbourhood_land_sea(torch.from_numpy(arr).cuda(), cells, torch.from_numpy(land).cuda(),
                                                method=neighbourhood_shape, weighted_mode=weighted_mode,
                                                sum_only=area_sum)
    res = out.cpu().numpy()
    engine.check_status(status)
    return cube.copy(data=np.transpose(res, inv).astype(np.float32))
