"""CPU legs of bench.py: the hot path on the host, through the REFERENCE'S OWN functions when the
build artefact oracle/_ref is present (kind "reference"), else through the numpy / scipy port in
oracle/nbhood_oracle.py (kind "port").  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Reference call sites restated here (cube plumbing replaced by plain loops):
  * slice loop of NeighbourhoodProcessing.process            improver/nbhood/nbhood.py:457-463
  * Threshold._calculate_truth_value per member and lead     improver/threshold.py:661-707
  * realization mean                                         improver/utilities/cube_manipulation.py:93-131
  * land + sea combination                                   improver/cli/nbhood_land_and_sea.py:148-184
  * GeneratePercentilesFromANeighbourhood.pad_and_unpad_cube improver/nbhood/nbhood.py:649-667
"""
from __future__ import annotations

import numpy as np

from . import nbhood_oracle as port
from . import ref_loader

_ns = None


def kind() -> str:
    return "reference" if ref_loader.available() else "port"


def _ref():
    global _ns
    if _ns is None:
        _ns = ref_loader.load()
    return _ns


class _FakeCube:
    def __init__(self, data):
        self.data = data
        self.dtype = data.dtype


def _nb_plugin(method, cells, weighted=False, sum_only=False, re_mask=True):
    ref = _ref()
    plugin = ref.nb.NeighbourhoodProcessing(method, 2000.0 * cells, weighted_mode=weighted, sum_only=sum_only,
                                            re_mask=re_mask)
    if method == "circular":
        plugin.kernel = ref.nb.circular_kernel(cells, weighted)
        plugin.nb_size = max(plugin.kernel.shape)
    else:
        plugin.nb_size = 2 * cells + 1
    return plugin


def neighbourhood_slices(data, mask=None, method="square", cells=5, weighted=False, sum_only=False,
                         re_mask=False):
    """Every leading index of `data` is an independent 2-D slice (nbhood.py:457-463)."""
    flat = data.reshape((-1,) + data.shape[-2:])
    if kind() == "port":
        out = [np.ma.getdata(port.neighbourhood(s, mask, method=method, cells=cells, weighted_mode=weighted,
                                                sum_only=sum_only, re_mask=re_mask)) for s in flat]
    else:
        plugin = _nb_plugin(method, cells, weighted, sum_only, re_mask)
        out = [np.ma.getdata(plugin._calculate_neighbourhood(s, mask)) for s in flat]
    return np.stack(out).reshape(data.shape)


def land_sea(data, land_mask, method="circular", cells=9):
    """Land pass + sea pass + filled(0) sum (cli/nbhood_land_and_sea.py:148-184)."""
    if kind() == "port":
        return port.land_sea_neighbourhood(data, land_mask, method=method, cells=cells)
    land = (land_mask != 0).astype(np.int8)
    sea = (land_mask == 0).astype(np.int8)
    plugin = _nb_plugin(method, cells, re_mask=True)
    flat = data.reshape((-1,) + data.shape[-2:])
    out = np.empty(flat.shape, dtype=np.float32)
    for i, s in enumerate(flat):
        a = plugin._calculate_neighbourhood(s, land)
        b = plugin._calculate_neighbourhood(s, sea)
        out[i] = np.ma.filled(a, 0) + np.ma.filled(b, 0)
    return out.reshape(data.shape)


def threshold_square_mean(data, thresholds, bounds, comparison, cells, method="square"):
    """Fuzzy threshold per member -> neighbourhood per member -> realization mean.
    data (R, L, y, x) -> (L, T, y, x)."""
    if kind() == "port":
        return port.threshold_square_mean(data, thresholds, bounds, comparison, cells, method=method)
    ref = _ref()
    R, L = data.shape[:2]
    out = np.empty((L, len(thresholds)) + data.shape[-2:], dtype=np.float32)
    nbp = _nb_plugin(method, cells, re_mask=False)
    for k, (thr, bnd) in enumerate(zip(thresholds, bounds)):
        plug = ref.th.Threshold(threshold_config={str(thr): [float(bnd[0]), float(bnd[1])]},
                                comparison_operator=comparison)
        for lead in range(L):
            members = np.empty((R,) + data.shape[-2:], dtype=np.float32)
            for m in range(R):
                tv = plug._calculate_truth_value(_FakeCube(data[m, lead]), plug.thresholds[0], plug.fuzzy_bounds[0])
                members[m] = np.ma.getdata(nbp._calculate_neighbourhood(np.asarray(tv, dtype=np.float32), None))
            out[lead, k] = members.mean(axis=0)
    return out


def percentiles(data2d, cells, pcts):
    """nbhood.py:649-667 for one 2-D slice -> (P, y, x)."""
    if kind() == "port":
        return port.neighbourhood_percentiles(data2d, cells, pcts)
    ref = _ref()
    kernel = ref.nb.circular_kernel(cells, False)
    win = ref.nt.pad_and_roll(data2d, kernel.shape, mode="mean", stat_length=max(kernel.shape) // 2)
    p = np.array(pcts, dtype=np.float32)
    out = np.empty((len(pcts),) + data2d.shape, np.float32)
    for chunk, o in zip(win, out.swapaxes(0, 1)):
        np.percentile(chunk[..., kernel > 0], p, axis=-1, out=o, overwrite_input=True)
    return out
