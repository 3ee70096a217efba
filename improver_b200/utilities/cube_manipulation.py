"""collapse_realizations — drop-in surface of improver/utilities/cube_manipulation.py:93-131
(the realization mean that closes the Threshold -> neighbourhood -> mean chain)."""
from __future__ import annotations

import numpy as np


def collapse_realizations(cube, method: str = "mean"):
    """Collapse the realization coordinate with the arithmetic mean (the reference maps
    "mean" to iris.analysis.MEAN; other aggregators are outside the hot path).  The realization
    coordinate is removed, every other coordinate and attribute is kept.

    Deferred cubes (improver_b200.lazy.defer) stay deferred: a
    Threshold -> NeighbourhoodProcessing -> collapse_realizations chain is then evaluated by
    the fused kernel when the data is asked for."""
    if method != "mean":
        raise NotImplementedError(f"collapse_realizations(method={method!r}) is outside the CUDA hot path")
    if not cube.coords("realization"):
        raise ValueError("No realization coordinate to collapse")
    dims = cube.coord_dims(cube.coord("realization"))
    if not dims:
        result = cube.copy()                      # scalar realization: nothing to average
        result.remove_coord("realization")
        return result
    axis = dims[0]
    from ..lazy import CollapseOp, Deferred

    core = cube.core_data() if hasattr(cube, "core_data") else None
    if isinstance(core, Deferred):
        data = CollapseOp(core, axis)
    else:
        import torch

        from .. import engine

        host = cube.data
        mask = None
        if isinstance(host, np.ma.MaskedArray):
            mask = np.ma.getmaskarray(host)
            if mask.any() and not (mask == mask.take([0], axis=axis)).all():
                raise NotImplementedError("realization mean of differently masked members")
            host = np.ma.getdata(host)
        dev = torch.from_numpy(np.ascontiguousarray(np.moveaxis(np.asarray(host, dtype=np.float32), axis, 0))).cuda()
        data = engine.collapse_mean(dev).cpu().numpy()
        if mask is not None:
            data = np.ma.masked_array(data, mask.take(0, axis=axis))
    result = cube.copy(data=data)
    if hasattr(result, "_dim_coords"):
        result._dim_coords = [(c, d - 1 if d > axis else d) for c, d in result._dim_coords if d != axis]
    result.remove_coord("realization")
    return result
