"""ApplyNeighbourhoodProcessingWithAMask — drop-in surface of
improver/nbhood/use_nbhood.py:23-262.

The reference loops 2-D slices x topographic-zone masks through
NeighbourhoodProcessing(re_mask=False) and optionally collapses the zone axis
with NaN-aware, renormalised weights.  Here every zone mask is one batched
launch over all slices; the zone collapse is a small elementwise reduction on
the device.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np

from .. import PostProcessingPlugin
from .nbhood import NeighbourhoodProcessing, _yx_last


class ApplyNeighbourhoodProcessingWithAMask(PostProcessingPlugin):
    def __init__(self, coord_for_masking: str, neighbourhood_method: str,
                 radii: Union[float, List[float]], lead_times: Optional[List[float]] = None,
                 collapse_weights=None, weighted_mode: bool = False, sum_only: bool = False) -> None:
        self.coord_for_masking = coord_for_masking
        self.neighbourhood_method = neighbourhood_method
        self.radii = radii
        self.lead_times = lead_times
        self.collapse_weights = collapse_weights
        self.weighted_mode = weighted_mode
        self.sum_only = sum_only
        self.re_mask = False

    @staticmethod
    def _collapse(values: np.ndarray, weights) -> np.ndarray:
        """use_nbhood.py:158-192: weighted mean over the zone axis (axis -3),
        NaN results excluded and the weights renormalised; all-NaN -> NaN.
        Masked weights (e.g. sea points) mask the result.  One native kernel (nbh_zone_collapse_f32)."""
        import torch

        from .. import engine

        w = np.ma.getdata(weights).astype(np.float32)
        wmask = np.ma.getmaskarray(weights)
        v = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float32)).cuda()
        wt = torch.from_numpy(np.ascontiguousarray(np.where(wmask, 0.0, w).astype(np.float32))).cuda()
        res = engine.zone_collapse(v, wt).cpu().numpy()
        if wmask.any():
            all_masked = wmask.all(axis=0)
            res = np.ma.masked_array(res, np.broadcast_to(all_masked, res.shape))
        return res

    def process(self, cube, mask_cube):
        plugin = NeighbourhoodProcessing(self.neighbourhood_method, self.radii, lead_times=self.lead_times,
                                         weighted_mode=self.weighted_mode, sum_only=self.sum_only,
                                         re_mask=self.re_mask)
        masks = np.asarray(mask_cube.data)
        zdim = mask_cube.coord_dims(mask_cube.coord(self.coord_for_masking))
        if zdim:
            masks = np.moveaxis(masks, zdim[0], 0)
        else:
            masks = masks[None]
        perm, inv = _yx_last(cube)
        zone_results = []
        template = None
        for z in range(masks.shape[0]):
            zmask = type("MaskSlice", (), {"data": masks[z]})()
            out = plugin.process(cube, zmask)
            template = out
            zone_results.append(np.transpose(np.ma.getdata(out.data), perm))
        stacked = np.stack(zone_results, axis=-3)                    # (lead..., zone, y, x)
        if self.collapse_weights is not None:
            data = self._collapse(stacked, self.collapse_weights.data)
            data = np.ma.transpose(data, inv) if isinstance(data, np.ma.MaskedArray) else np.transpose(data, inv)
            return template.copy(data=data)
        result = template.copy(data=stacked)
        if hasattr(result, "_dim_coords"):
            n_lead = stacked.ndim - 3
            zcoord = mask_cube.coord(self.coord_for_masking).copy()
            new = []
            for c, d in result._dim_coords:
                pos = perm.index(d)
                new.append((c, pos if pos < n_lead else pos + 1))
            new.append((zcoord, n_lead))
            result._dim_coords = new
        return result
