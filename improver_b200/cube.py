"""Minimal cube / coordinate doubles.

The reference plugins exchange iris cubes.  iris is not available in this
image, so the plugins in this package are duck-typed: they only use the small
part of the iris API implemented here (`data`, `shape`, `coord(name|axis=)`,
`coords`, `coord_dims`, `copy`, `name`, `attributes`, coordinate `points`,
`units`, `convert_units`, `copy`).  A real iris cube works unchanged; these
classes exist so that the package is usable and testable without iris.
"""
from __future__ import annotations

import copy as _copy
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_LENGTH = {"m": 1.0, "metres": 1.0, "metre": 1.0, "meters": 1.0, "km": 1000.0, "kilometres": 1000.0}
_TIME = {"s": 1.0, "seconds": 1.0, "second": 1.0, "minutes": 60.0, "hours": 3600.0, "hour": 3600.0}


# linear unit conversions of the cube double (value_SI = scale * value + offset), grouped by quantity;
# with real iris cubes cf_units does this (improver/threshold.py:430-431 `cube.convert_units`)
_LINEAR_UNITS = {
    "temperature": {"K": (1.0, 0.0), "kelvin": (1.0, 0.0), "degC": (1.0, 273.15), "celsius": (1.0, 273.15),
                    "degrees_celsius": (1.0, 273.15), "degF": (5.0 / 9.0, 459.67 * 5.0 / 9.0)},
    "speed": {"m s-1": (1.0, 0.0), "m/s": (1.0, 0.0), "km h-1": (1.0 / 3.6, 0.0), "knots": (0.514444444444444, 0.0),
              "mm h-1": (1.0 / 3.6e6, 0.0), "mm s-1": (1e-3, 0.0), "mm/hr": (1.0 / 3.6e6, 0.0)},
    "length": {"m": (1.0, 0.0), "metres": (1.0, 0.0), "km": (1000.0, 0.0), "mm": (1e-3, 0.0), "cm": (1e-2, 0.0),
               "feet": (0.3048, 0.0), "ft": (0.3048, 0.0)},
    "fraction": {"1": (1.0, 0.0), "%": (0.01, 0.0), "percent": (0.01, 0.0)},
    "pressure": {"Pa": (1.0, 0.0), "hPa": (100.0, 0.0), "mbar": (100.0, 0.0), "kPa": (1000.0, 0.0)},
}


def linear_unit_map(unit: str, target: str):
    """(a, b) with value_target = a * value_unit + b, or ValueError for units of different quantities."""
    unit, target = str(unit), str(target)
    if unit == target:
        return 1.0, 0.0
    for table in _LINEAR_UNITS.values():
        if unit in table and target in table:
            s1, o1 = table[unit]
            s2, o2 = table[target]
            return s1 / s2, (o1 - o2) / s2
    raise ValueError(f"Unable to convert from '{unit}' to '{target}'.")


class CoordinateNotFoundError(KeyError):
    pass


def _factor(unit: str, target: str) -> float:
    for table in (_LENGTH, _TIME):
        if unit in table and target in table:
            return table[unit] / table[target]
    if unit == target:
        return 1.0
    raise ValueError(f"Unable to convert from '{unit}' to '{target}'.")


class Coord:
    """1-D coordinate: points, units, optional axis ('x' / 'y')."""

    def __init__(self, points, standard_name: str, units: str = "1", axis: Optional[str] = None,
                 var_name: Optional[str] = None, attributes: Optional[dict] = None):
        self.points = np.atleast_1d(np.asarray(points))
        self._name = standard_name
        self.units = units
        self.axis = axis
        self.var_name = var_name
        self.attributes = dict(attributes or {})

    def name(self) -> str:
        return self._name

    def rename(self, name: str) -> None:
        self._name = name

    def copy(self) -> "Coord":
        return _copy.deepcopy(self)

    def convert_units(self, target: str) -> None:
        target = str(target)
        f = _factor(str(self.units), target)
        if f != 1.0:
            self.points = self.points * f
        self.units = target

    @property
    def shape(self):
        return self.points.shape

    def __repr__(self):
        return f"Coord({self._name!r}, n={self.points.size}, units={self.units!r})"


class Cube:
    """N-D array with named dimension / scalar coordinates."""

    def __init__(self, data, name: str = "unknown", units: str = "1",
                 dim_coords: Sequence[Tuple[Coord, int]] = (), scalar_coords: Sequence[Coord] = (),
                 attributes: Optional[Dict] = None):
        self._data = data
        self._name = name
        self.units = units
        self._dim_coords: List[Tuple[Coord, int]] = [(c, int(d)) for c, d in dim_coords]
        self._scalar_coords: List[Coord] = list(scalar_coords)
        self.attributes = dict(attributes or {})
        self.cell_methods = ()

    # --- data: like iris, `data` realises deferred (device-resident, lazily evaluated) data and
    # `core_data()` hands it out as it is (iris.cube.Cube.core_data / has_lazy_data) ------------
    @property
    def data(self):
        from .lazy import Deferred

        if isinstance(self._data, Deferred):
            self._data = self._data.realise()
        return self._data

    @data.setter
    def data(self, value):
        self._data = value

    def core_data(self):
        return self._data

    def has_lazy_data(self) -> bool:
        from .lazy import Deferred

        return isinstance(self._data, Deferred)

    # --- basic properties -------------------------------------------------
    @property
    def shape(self):
        return tuple(self._data.shape)

    @property
    def ndim(self):
        return len(self._data.shape)

    @property
    def dtype(self):
        return self._data.dtype

    def name(self) -> str:
        return self._name

    def rename(self, name: str) -> None:
        self._name = name

    # --- coordinates ----------------------------------------------------------
    @property
    def dim_coords(self):
        return tuple(c for c, _ in sorted(self._dim_coords, key=lambda cd: cd[1]))

    def coords(self, name: Optional[str] = None, axis: Optional[str] = None, dim_coords=None):
        out = []
        pools = [(c, True) for c, _ in self._dim_coords]
        if not dim_coords:
            pools += [(c, False) for c in self._scalar_coords]
        for c, _is_dim in pools:
            if name is not None and c.name() != name and c.var_name != name:
                continue
            if axis is not None and (c.axis or "").lower() != axis.lower():
                continue
            out.append(c)
        return out

    def coord(self, name: Optional[str] = None, axis: Optional[str] = None, var_name: Optional[str] = None):
        if var_name is not None:
            found = [c for c in self.coords() if c.var_name == var_name]
        else:
            found = self.coords(name=name, axis=axis)
        if len(found) != 1:
            what = name or var_name or f"axis={axis}"
            raise CoordinateNotFoundError(f"Expected to find exactly 1 {what!r} coordinate, "
                                          f"but found {len(found) or 'none'}.")
        return found[0]

    def coord_dims(self, coord) -> Tuple[int, ...]:
        if isinstance(coord, str):
            coord = self.coord(coord)
        for c, d in self._dim_coords:
            if c is coord:
                return (d,)
        return ()

    def add_dim_coord(self, coord: Coord, dim: int) -> None:
        self._dim_coords.append((coord, int(dim)))

    def add_aux_coord(self, coord: Coord) -> None:
        self._scalar_coords.append(coord)

    def remove_coord(self, name: str) -> None:
        self._dim_coords = [(c, d) for c, d in self._dim_coords if c.name() != name]
        self._scalar_coords = [c for c in self._scalar_coords if c.name() != name]

    # --- copying --------------------------------------------------------------
    def copy(self, data=None) -> "Cube":
        new = Cube(self._data.copy() if data is None else data, self._name, self.units,
                   [(c.copy(), d) for c, d in self._dim_coords],
                   [c.copy() for c in self._scalar_coords], _copy.deepcopy(self.attributes))
        new.cell_methods = self.cell_methods
        return new

    def __repr__(self):
        dims = ", ".join(f"{c.name()}: {c.points.size}" for c in self.dim_coords)
        return f"<Cube {self._name} ({dims})>"


def is_masked(data) -> bool:
    return isinstance(data, np.ma.MaskedArray)
