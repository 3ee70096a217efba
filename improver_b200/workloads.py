"""Synthetic inputs of the BASELINE.json configurations, shared by
bench.py and the full-size parity tests so that both look at the same fields.

Host generators are numpy (seeded, reproducible anywhere); `philox_slices` is the
counter-based device generator of config 5: element i of the global cube only
depends on (seed, i), so any sharding over ranks reproduces the same cube."""
from __future__ import annotations

import numpy as np

UK = (970, 1042)                   # improver/utilities/save.py:95 (y, x) of the UK 2 km grid
GLOBAL = (1920, 2560)              # 10 km global grid, periodic in x
SPACING_UK = 2000.0
# config 2: test_nbhood_land_and_sea.py:53-56 style lead-time dependent radii
CONFIG2_RADII_M = [18000.0, 54000.0, 90000.0, 162000.0]
CONFIG2_LEAD_HOURS = [0, 36, 72, 144]


def land_sea_mask(shape=UK, seed: int = 2, sea_fraction: float = 0.4, smooth: int = 24) -> np.ndarray:
    """int8 (y, x), 1 = land, 0 = sea: white noise smoothed with a (2*smooth+1)^2 box and cut at
    the `sea_fraction` quantile — coasts, islands and lakes at every scale above ~50 cells."""
    rng = np.random.default_rng(seed)
    ny, nx = shape
    noise = rng.standard_normal((ny + 2 * smooth, nx + 2 * smooth))
    c = np.pad(noise, ((1, 0), (1, 0))).cumsum(0).cumsum(1)
    nb = 2 * smooth + 1
    field = c[nb:, nb:] - c[:-nb, nb:] - c[nb:, :-nb] + c[:-nb, :-nb]
    return (field > np.quantile(field, sea_fraction)).astype(np.int8)


def config2_probabilities(n_slices: int, shape=UK, seed: int = 1, structured: bool = False) -> np.ndarray:
    """float32 (n_slices, y, x).  Plain: uniform [0, 1).  Structured: slices alternate between an
    all-0 and an all-1 background with uniform blobs, one of them touching a domain edge (this
    is what exercises the reference's bounding-box trimming, nbhood.py:341-409)."""
    rng = np.random.default_rng(seed)
    ny, nx = shape
    if not structured:
        return rng.random((n_slices, ny, nx), dtype=np.float32)
    out = np.empty((n_slices, ny, nx), dtype=np.float32)
    for i in range(n_slices):
        out[i] = float(i & 1)
        for k in range(4):
            h, w = int(rng.integers(20, ny // 4)), int(rng.integers(20, nx // 4))
            y0 = 0 if k == 0 else int(rng.integers(0, ny - h))
            x0 = int(rng.integers(0, nx - w)) if k != 1 else nx - w
            out[i, y0:y0 + h, x0:x0 + w] = rng.random((h, w), dtype=np.float32)
    return out


def config3_rain(shape, seed: int = 3) -> np.ndarray:
    """gamma(2, 1.5) "rain rate", float32."""
    return np.random.default_rng(seed).gamma(2.0, 1.5, shape).astype(np.float32)


def cells_for_lead_time(fp_hours: float, radii=CONFIG2_RADII_M, lead_times=CONFIG2_LEAD_HOURS,
                        spacing: float = SPACING_UK) -> int:
    """nbhood.py:132-149 + spatial.py:152-156: interpolate the radius, truncate to cells."""
    return int(np.interp(fp_hours, lead_times, radii) / spacing)


# ----------------------------------------------------------------------------- #
# counter-based device generator (config 5)
# ----------------------------------------------------------------------------- #
def _mix32(x):
    """lowbias32 integer hash (public domain, Chris Wellons) on int64 tensors holding uint32."""
    m = 0xFFFFFFFF
    x = x & m
    x = ((x ^ (x >> 16)) * 0x7FEB352D) & m
    x = ((x ^ (x >> 15)) * 0x846CA68B) & m
    return (x ^ (x >> 16)) & m


def counter_uniform(first_slice: int, n_slices: int, shape, seed: int, device, out=None):
    """float32 (n_slices, y, x) uniform [0, 1) where element (s, y, x) is a hash of
    (seed, global slice index s + first_slice, y * nx + x): independent of how the slice axis is
    sharded over ranks.  Pure torch integer arithmetic on `device` (test / bench input only)."""
    import torch

    ny, nx = shape
    res = out if out is not None else torch.empty((n_slices, ny, nx), dtype=torch.float32, device=device)
    pt = torch.arange(ny * nx, dtype=torch.int64, device=device)
    for k in range(n_slices):
        s = first_slice + k
        key = (seed * 0x9E3779B1 + s * 0x85EBCA77) & 0xFFFFFFFF
        h = _mix32(pt ^ key)
        h = _mix32(h + ((s >> 16) & 0xFFFF) + 0x68E31DA4)
        res[k] = ((h >> 8).to(torch.float32) * (1.0 / 16777216.0)).view(ny, nx)
    return res


def counter_uniform_host(first_slice: int, n_slices: int, shape, seed: int) -> np.ndarray:
    """numpy restatement of counter_uniform (bit-identical), for CPU-side checks."""
    ny, nx = shape
    m = np.uint64(0xFFFFFFFF)

    def mix(x):
        x = x & m
        x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x7FEB352D)) & m
        x = ((x ^ (x >> np.uint64(15))) * np.uint64(0x846CA68B)) & m
        return (x ^ (x >> np.uint64(16))) & m

    pt = np.arange(ny * nx, dtype=np.uint64)
    out = np.empty((n_slices, ny, nx), dtype=np.float32)
    for k in range(n_slices):
        s = first_slice + k
        key = np.uint64((seed * 0x9E3779B1 + s * 0x85EBCA77) & 0xFFFFFFFF)
        h = mix(pt ^ key)
        h = mix(h + np.uint64(((s >> 16) & 0xFFFF) + 0x68E31DA4))
        out[k] = ((h >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).reshape(ny, nx)
    return out
