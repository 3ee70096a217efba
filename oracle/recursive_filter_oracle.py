"""CPU oracle for the recursive filter (kernel part).

TEST INFRASTRUCTURE ONLY: imported by `tests/` alone; the product path
(`improver_b200/`) never imports it and has no CPU fallback.

Parity status: PINNED.  tests/test_recursive_filter.py checks this module against
  * the reference's own known answers for the filter (literal arrays of
    improver_tests/nbhood/recursive_filter/test_RecursiveFilter.py:287-360,
    383-448, 467-600), and
  * golden vectors produced by running the real reference functions from
    /root/reference in the build container (tests/golden/make_golden_recursive.py,
    fixture tests/golden/recursive_golden.npz).

Every step is float32, one rounding per numpy operation, exactly as the
reference's row-at-a-time numpy expressions evaluate them.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def recurse_forward(grid: np.ndarray, coeffs: np.ndarray, axis: int) -> np.ndarray:
    """B_i = (1 - c_{i-1}) A_i + c_{i-1} B_{i-1}, i = 1 .. n-1, in place
    (improver/nbhood/recursive_filter.py:55-103)."""
    g = np.moveaxis(grid, axis, 0)
    c = np.moveaxis(coeffs, axis, 0)
    for i in range(1, g.shape[0]):
        g[i] = (1.0 - c[i - 1]) * g[i] + c[i - 1] * g[i - 1]
    return grid


def recurse_backward(grid: np.ndarray, coeffs: np.ndarray, axis: int) -> np.ndarray:
    """B_i = (1 - c_i) A_i + c_i B_{i+1}, i = n-2 .. 0, in place
    (improver/nbhood/recursive_filter.py:105-153)."""
    g = np.moveaxis(grid, axis, 0)
    c = np.moveaxis(coeffs, axis, 0)
    for i in range(g.shape[0] - 2, -1, -1):
        g[i] = (1.0 - c[i]) * g[i] + c[i] * g[i + 1]
    return grid


def zero_masked(coeff_x: np.ndarray, coeff_y: np.ndarray, mask: np.ndarray):
    """Coefficients next to a masked point become 0 (use_mask_boundary=False, invert_mask=False:
    improver/generate_ancillaries/generate_orographic_smoothing_coefficients.py:242-253)."""
    m = np.asarray(mask, bool)
    cx, cy = coeff_x.copy(), coeff_y.copy()
    cx[m[:, :-1] | m[:, 1:]] = 0.0
    cy[m[:-1, :] | m[1:, :]] = 0.0
    return cx, cy


def filter_slice(data: np.ndarray, coeff_x: np.ndarray, coeff_y: np.ndarray, iterations: int,
                 edge_width: int, mask: Optional[np.ndarray] = None) -> np.ndarray:
    """One (y, x) slice: symmetric halo of 2 * edge_width on the data and on the
    coefficients, `iterations` x (x forward, x backward, y forward, y backward), halo removed
    (improver/nbhood/recursive_filter.py:155-202, 262-273, 403-436; utilities/pad_spatial.py:204-208,
    288-292).  `mask` True = masked point (its coefficients are zeroed, :410-422)."""
    pad = 2 * int(edge_width)
    cx, cy = (coeff_x, coeff_y) if mask is None or not np.any(mask) else zero_masked(coeff_x, coeff_y, mask)
    g = np.pad(np.asarray(data), ((pad, pad), (pad, pad)), mode="symmetric")
    px = np.pad(cx, ((pad, pad), (pad, pad)), mode="symmetric")
    py = np.pad(cy, ((pad, pad), (pad, pad)), mode="symmetric")
    for _ in range(iterations):
        recurse_forward(g, px, 1)
        recurse_backward(g, px, 1)
        recurse_forward(g, py, 0)
        recurse_backward(g, py, 0)
    end = -pad if pad else None
    return g[pad:end, pad:end]


def recursive_filter(data, coeff_x: np.ndarray, coeff_y: np.ndarray, iterations: int, edge_width: int = 15,
                     mask_zeros: bool = False):
    """RecursiveFilter.process on a (..., y, x) array or masked array
    (improver/nbhood/recursive_filter.py:369-451): every slice filtered with the coefficients
    zeroed around its own masked points; `mask_zeros` also masks the exact zeros (:395-396)
    and removes that mask again at the end (:443-448)."""
    orig_mask = np.ma.getmaskarray(data) if np.ma.isMaskedArray(data) else None
    raw = np.ma.getdata(data).astype(np.float32, copy=False)
    mask = orig_mask
    if mask_zeros:
        mask = (raw == 0.0) if mask is None else (mask | (raw == 0.0))
    lead = raw.shape[:-2]
    out = np.empty_like(raw)
    any_masked = False
    for idx in np.ndindex(*lead):
        m = None if mask is None else mask[idx]
        any_masked |= m is not None and bool(np.any(m))
        out[idx] = filter_slice(raw[idx], coeff_x, coeff_y, iterations, edge_width, m)
    if mask_zeros:
        return out if orig_mask is None else np.ma.array(out, mask=orig_mask)
    if any_masked:
        return np.ma.MaskedArray(out, mask=mask)
    return out
