"""Thresholding plugin — drop-in surface of improver/threshold.py::Threshold
backed by the CUDA threshold kernel (nbh_threshold_f32).

Constructor arguments, validation order, error types and messages follow
improver/threshold.py:47-239; `process` follows :571-755, including maximum-in-
vicinity processing (:473-518, one fused CUDA launch: nbh_threshold_vicinity_f32).
Unit conversion of the data and percentile rebadging are outside the hot path
 and raise NotImplementedError.
"""
from __future__ import annotations

import numbers
from collections.abc import Iterable
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from . import PostProcessingPlugin
from .constants import FLOAT_DTYPE
from .utilities.probability_manipulation import comparison_operator_dict


class Threshold(PostProcessingPlugin):
    """Apply a (fuzzy) threshold truth criterion to a cube."""

    def __init__(self, threshold_values: Optional[Union[float, List[float]]] = None,
                 threshold_config: Optional[Dict[str, Union[List[float], str]]] = None,
                 fuzzy_factor: Optional[float] = None, threshold_units: Optional[str] = None,
                 comparison_operator: str = ">", collapse_coord: Optional[Union[str, List[str]]] = None,
                 collapse_cell_methods: Optional[dict] = None,
                 vicinity: Optional[Union[float, List[float]]] = None,
                 fill_masked: Optional[float] = None) -> None:
        if threshold_config and threshold_values:
            raise ValueError("threshold_config and threshold_values are mutually exclusive "
                             "arguments - please provide one or the other, not both")
        if threshold_config is None and threshold_values is None:
            raise ValueError("One of threshold_config or threshold_values must be provided.")
        thresholds, fuzzy_bounds = self._set_thresholds(threshold_values, threshold_config)
        self.thresholds = [thresholds] if np.isscalar(thresholds) else thresholds
        self.threshold_units = threshold_units
        self.threshold_coord_name = None

        fuzzy_factor_loc = 1.0
        if fuzzy_factor is not None:
            if fuzzy_bounds is not None:
                raise ValueError("Invalid combination of keywords. Cannot specify "
                                 "both a fuzzy_factor and use a threshold_config that "
                                 "specifies bounds.")
            if not 0 < fuzzy_factor < 1:
                raise ValueError("Invalid fuzzy_factor: must be >0 and <1: {}".format(fuzzy_factor))
            if 0 in self.thresholds:
                raise ValueError("Invalid threshold with fuzzy factor: cannot use a "
                                 "multiplicative fuzzy factor with threshold == 0, use "
                                 "the threshold_config approach instead.")
            fuzzy_factor_loc = fuzzy_factor

        if fuzzy_bounds is None:
            self.fuzzy_bounds = self._generate_fuzzy_bounds(fuzzy_factor_loc)
        else:
            self.fuzzy_bounds = [fuzzy_bounds] if isinstance(fuzzy_bounds, tuple) else fuzzy_bounds
            self._check_fuzzy_bounds()

        self.original_units = None
        self.comparison_operator_dict = comparison_operator_dict()
        self.comparison_operator_string = comparison_operator
        self._decode_comparison_operator_string()

        if collapse_coord and not isinstance(collapse_coord, list):
            collapse_coord = [collapse_coord]
        if collapse_coord and not set(collapse_coord).issubset(["percentile", "realization", "time"]):
            raise ValueError("Can only collapse over one or a combination of realization, "
                             "percentile, and time coordinates.")
        if collapse_coord and "percentile" in collapse_coord and "realization" in collapse_coord:
            raise ValueError("Cannot collapse over both percentile and realization coordinates.")
        self.collapse_coord = collapse_coord

        if collapse_cell_methods:
            if not self.collapse_coord:
                raise ValueError("Cannot apply cell methods without collapsing a coordinate.")
            if not set(collapse_cell_methods.keys()).issubset(self.collapse_coord):
                raise ValueError("Cell methods can only be defined for coordinates that are being collapsed.")
        self.collapse_cell_methods = collapse_cell_methods

        self.vicinity = None
        if vicinity:
            if isinstance(vicinity, Iterable):
                self.vicinity = [float(x) for x in vicinity]
            else:
                self.vicinity = [float(vicinity)]
        if fill_masked is not None:
            fill_masked = float(fill_masked)
        self.fill_masked = fill_masked

    # -- configuration helpers (threshold.py:241-333) ------------------------------
    @staticmethod
    def _set_thresholds(threshold_values, threshold_config) -> Tuple[List[float], Optional[List]]:
        if threshold_config:
            thresholds, fuzzy_bounds = [], []
            for key in threshold_config.keys():
                thresholds.append(float(key))
                if threshold_config[key] == "None":
                    fuzzy_bounds = None
                    continue
                fuzzy_bounds.append(tuple(threshold_config[key]))
        else:
            if isinstance(threshold_values, numbers.Number):
                threshold_values = [threshold_values]
            thresholds = [float(x) for x in threshold_values]
            fuzzy_bounds = None
        return thresholds, fuzzy_bounds

    def _generate_fuzzy_bounds(self, fuzzy_factor_loc: float) -> List[Tuple[float, float]]:
        fuzzy_bounds = []
        for thr in self.thresholds:
            lower_thr = thr * fuzzy_factor_loc
            upper_thr = thr * (2.0 - fuzzy_factor_loc)
            if thr < 0:
                lower_thr, upper_thr = upper_thr, lower_thr
            fuzzy_bounds.append((lower_thr, upper_thr))
        return fuzzy_bounds

    def _check_fuzzy_bounds(self) -> None:
        for thr, bounds in zip(self.thresholds, self.fuzzy_bounds):
            if len(bounds) != 2:
                raise ValueError("Invalid bounds for one threshold: {}. Expected 2 floats.".format(bounds))
            if bounds[0] > thr or bounds[1] < thr:
                raise ValueError("Threshold must be within bounds: " "!( {} <= {} <= {} )".format(
                    bounds[0], thr, bounds[1]))

    def _decode_comparison_operator_string(self) -> None:
        try:
            self.comparison_operator = self.comparison_operator_dict[self.comparison_operator_string]
        except KeyError:
            msg = (f'String "{self.comparison_operator_string}" '
                   "does not match any known comparison_operator method")
            raise ValueError(msg)

    # -- array level ---------------------------------------------------------------
    def _engine_thresholds(self):
        op = self.comparison_operator_string
        return [(thr, float(b[0]), float(b[1]), op) for thr, b in zip(self.thresholds, self.fuzzy_bounds)]

    def _calculate_truth_value(self, cube, threshold: float, bounds: Tuple[float, float]) -> np.ndarray:
        """threshold.py:407-471 for one threshold (data stays in its own units)."""
        import torch

        from . import engine

        data = self._in_threshold_units(cube, cube.data)
        masked = isinstance(data, np.ma.MaskedArray)
        raw = np.ascontiguousarray(np.ma.getdata(data), dtype=np.float32)
        mnd = None
        if masked:
            mnd = torch.from_numpy(np.ascontiguousarray(np.ma.getmaskarray(data)).astype(np.uint8)).cuda()
        out, _, status = engine.threshold(torch.from_numpy(raw).cuda(),
                                          [(threshold, float(bounds[0]), float(bounds[1]),
                                            self.comparison_operator_string)], mask_nd=mnd)
        res = out.cpu().numpy()[0]
        engine.check_status(status)
        if masked and bounds[0] != bounds[1]:
            res = np.ma.masked_array(res, mask=np.ma.getmaskarray(data))
        return res.astype(FLOAT_DTYPE)

    def _in_threshold_units(self, cube, data):
        """threshold.py:430-431: the DATA are converted to the threshold units before the comparison
        (the thresholds and fuzzy bounds stay as given).  cf_units converts float32 arrays as
        float32(a * double(x) + b); the same linear map runs on the device (engine.linear_map)."""
        if self.threshold_units is None or str(self.threshold_units) == str(getattr(cube, "units", "")):
            return data
        import torch

        from . import engine
        from .cube import linear_unit_map

        if hasattr(cube, "_dim_coords"):
            a, b = linear_unit_map(str(cube.units), str(self.threshold_units))
        else:                                   # iris cube: let cf_units supply the two constants
            unit = cube.units
            b = float(unit.convert(0.0, self.threshold_units))
            a = float(unit.convert(1.0, self.threshold_units)) - b
        raw = np.ascontiguousarray(np.ma.getdata(data), dtype=np.float32)
        conv = engine.linear_map(torch.from_numpy(raw).cuda(), a, b).cpu().numpy()
        if isinstance(data, np.ma.MaskedArray):
            conv = np.ma.masked_array(conv, np.ma.getmaskarray(data))
        return conv

    # -- cube level --------------------------------------------------------------------
    def process(self, input_cube, landmask=None):
        """threshold.py:571-755."""
        import torch

        from . import engine
        from .cube import Coord

        input_cube_crds = [crd.name() for crd in input_cube.coords()]
        if self.collapse_coord and not set(self.collapse_coord).issubset(input_cube_crds):
            raise ValueError("Cannot collapse over coordinates not present in the input_cube."
                             f"collapse_coords: {self.collapse_coord}\n"
                             f"input_cube_coords: {input_cube_crds}")
        if self.vicinity is None and landmask is not None:
            raise ValueError("Cannot apply land-mask cube without in-vicinity processing")
        if self.collapse_coord and "percentile" in self.collapse_coord:
            raise NotImplementedError("percentile rebadging is outside the CUDA hot path")

        self.original_units = input_cube.units
        self.threshold_coord_name = input_cube.name()
        dim_names = []
        for d in range(len(input_cube.shape)):
            found = [c for c in input_cube.dim_coords if input_cube.coord_dims(c) == (d,)]
            dim_names.append(found[0].name() if found else None)

        # deferred (device-resident) input: record the operation, evaluate with the rest of the chain
        from .lazy import Deferred, ThresholdOp

        core = input_cube.core_data() if hasattr(input_cube, "core_data") else None
        same_units = self.threshold_units is None or str(self.threshold_units) == str(input_cube.units)
        if isinstance(core, Deferred) and self.vicinity is None and not self.collapse_coord and same_units:
            n_in = len(dim_names)
            front = [i for i, n in enumerate(dim_names) if n in ("time", "realization")]
            front.sort(key=lambda i: ["time", "realization"].index(dim_names[i]))
            rest = [i for i in range(n_in) if i not in front]
            axes = [i + 1 for i in front] + [0] + [i + 1 for i in rest]
            full = (len(self.thresholds),) + tuple(input_cube.shape)
            squeeze_axes = tuple(i for i, a in enumerate(axes) if full[a] == 1)
            op = ThresholdOp(core, self._engine_thresholds(), axes, squeeze_axes)
            out_names = [dim_names[i] for i in front] + ["__threshold__"] + [dim_names[i] for i in rest]
            return self._finish_cube(input_cube, op, out_names, squeeze_axes, 0)

        data = self._in_threshold_units(input_cube, input_cube.data)
        if self.fill_masked is not None:
            data = np.ma.filled(data, self.fill_masked)

        # collapse axes first, flattened into one leading axis
        collapse_dims = [dim_names.index(c) for c in (self.collapse_coord or []) if c in dim_names]
        keep_dims = [d for d in range(len(dim_names)) if d not in collapse_dims]
        masked = isinstance(data, np.ma.MaskedArray)
        raw = np.ma.getdata(data)
        if np.isnan(raw[~np.ma.getmaskarray(data)] if masked else raw).any():
            raise ValueError("Error: NaN detected in input cube data")
        perm = collapse_dims + keep_dims
        raw_t = np.ascontiguousarray(np.transpose(raw, perm), dtype=np.float32)
        kept_shape = raw_t.shape[len(collapse_dims):]
        n_collapse = int(np.prod(raw_t.shape[:len(collapse_dims)])) if collapse_dims else 1
        dev = torch.from_numpy(raw_t.reshape((n_collapse,) + kept_shape)).cuda()
        mnd = None
        if masked:
            m_t = np.ascontiguousarray(np.transpose(np.ma.getmaskarray(data), perm))
            mnd = torch.from_numpy(m_t.reshape((n_collapse,) + kept_shape).astype(np.uint8)).cuda()
        n_vic = 0
        if self.vicinity is not None:
            # threshold.py:637-641, :473-518: radii in grid cells, land and sea spread separately
            from .utilities.spatial import distance_to_number_of_grid_cells

            cells = [distance_to_number_of_grid_cells(input_cube, radius) for radius in self.vicinity]
            lm = None
            if landmask is not None:
                lm = torch.from_numpy(np.ascontiguousarray(np.asarray(landmask.data).astype(bool))
                                      .astype(np.uint8)).cuda()
            n_vic = len(cells)
            out, contrib, status = engine.threshold_vicinity(dev, self._engine_thresholds(), cells, mask_nd=mnd,
                                                             landmask=lm, collapse_leading=True)
            res = out.cpu().numpy()                # (V, T, kept...)
            res = np.moveaxis(res, 0, 1)           # (T, V, kept...)
        else:
            out, contrib, status = engine.threshold(dev, self._engine_thresholds(), mask_nd=mnd,
                                                    collapse_leading=True)
            res = out.cpu().numpy()                # (T, kept...)
        contrib = contrib.cpu().numpy()
        engine.check_status(status)
        valid = contrib.astype(bool)
        if not valid.all():
            res = np.ma.masked_array(res, mask=np.broadcast_to(~valid, res.shape))

        # dimension order: time, realization, threshold, rest (threshold.py:744-753), squeezed (:731)
        # with vicinity: ..., threshold, radius_of_vicinity, rest (:744-753)
        kept_names = [dim_names[d] for d in keep_dims]
        front = [i for i, n in enumerate(kept_names) if n in ("time", "realization")]
        front.sort(key=lambda i: ["time", "realization"].index(kept_names[i]))
        rest = [i for i in range(len(kept_names)) if i not in front]
        lead = 2 if n_vic else 1                   # leading axes of `res` that are not cube dims
        axes = [i + lead for i in front] + list(range(lead)) + [i + lead for i in rest]
        res = np.ma.transpose(res, axes) if isinstance(res, np.ma.MaskedArray) else np.transpose(res, axes)
        out_names = ([kept_names[i] for i in front] + ["__threshold__"] + (["__vicinity__"] if n_vic else [])
                     + [kept_names[i] for i in rest])
        squeeze_axes = tuple(i for i, s in enumerate(res.shape) if s == 1)
        res = res.squeeze(axis=squeeze_axes) if squeeze_axes else res
        return self._finish_cube(input_cube, res.astype(FLOAT_DTYPE), out_names, squeeze_axes, n_vic)

    def _finish_cube(self, input_cube, res, out_names, squeeze_axes, n_vic):
        """Metadata of the probability cube (threshold.py:349-405, 716-753)."""
        from .cube import Coord

        result = input_cube.copy(data=res)
        relative = "below" if "less_than" in self.comparison_operator.spp_string else "above"
        if hasattr(result, "_dim_coords"):
            old = {c.name(): c for c in input_cube.dim_coords}
            thr_points = np.array(self.thresholds, dtype=np.float64)
            if self.threshold_units is not None and str(self.threshold_units) != str(input_cube.units):
                # threshold.py:716-718: the coordinate is built in the threshold units and converted back
                from .cube import linear_unit_map
                a, b = linear_unit_map(str(self.threshold_units), str(input_cube.units))
                thr_points = a * thr_points + b
            thr_coord = Coord(thr_points.astype(FLOAT_DTYPE), self.threshold_coord_name,
                              str(input_cube.units), var_name="threshold",
                              attributes={"spp__relative_to_threshold": self.comparison_operator.spp_string})
            vic_coord = None
            if n_vic:
                # spatial.py:1026-1060 create_vicinity_coord
                vic_coord = Coord(np.array(self.vicinity, dtype=FLOAT_DTYPE), "radius_of_vicinity", "m")
            result._dim_coords = []
            pos = 0
            for i, name in enumerate(out_names):
                if name == "__threshold__":
                    coord = thr_coord
                elif name == "__vicinity__":
                    coord = vic_coord
                else:
                    coord = old[name].copy() if name in old else None
                if i in squeeze_axes:
                    if coord is not None:
                        result._scalar_coords.append(coord)
                    continue
                if coord is not None:
                    result._dim_coords.append((coord, pos))
                pos += 1
            for name in (self.collapse_coord or []):
                if name == "realization":
                    result.remove_coord("realization")
            result.rename("probability_of_{}{}_{}_threshold".format(
                self.threshold_coord_name, "_in_vicinity" if n_vic else "", relative))
            result.units = "1"
        return result


#: `north_star` names `BasicThreshold`; the reference snapshot only has `Threshold`
BasicThreshold = Threshold
