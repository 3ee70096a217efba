"""Neighbourhood array tools with the reference's signatures
(improver/utilities/neighbourhood_tools.py:13-211).

`rolling_window` / `pad_and_roll` only build strided *views* (no arithmetic)
and stay on the host.  `boxsum` — the summed-area-table box sum that is the
reference's square-neighbourhood work-horse — runs on the GPU: the padded
array is handed to the sliding-window CUDA kernel (engine.neighbourhood with
sum_only) and the interior is cropped, which is the same quantity as the
reference's four-corner difference of the cumulative sums.
"""
from __future__ import annotations

from typing import Any, Tuple, Union

import numpy as np


_TOO_FEW_DIMS = ("Number of dimensions of the input array must be greater than or "
                 "equal to the length of the neighbourhood shape used for "
                 "constructing rolling window neighbourhoods.")
_WINDOW_TOO_BIG = ("The calculated shape of the output array view contains a "
                   "dimension that is negative or zero. Each dimension of the "
                   "neighbourhood shape must be less than or equal to the "
                   "corresponding dimension of the input array.")


def rolling_window(input_array: np.ndarray, shape: Tuple[int, int], writeable: bool = False) -> np.ndarray:
    """Strided view whose trailing `len(shape)` axes enumerate the window around
    every valid position of the trailing axes of `input_array`
    (neighbourhood_tools.py:13-66).  No data is copied."""
    k = len(shape)
    if input_array.ndim < k:
        raise ValueError(_TOO_FEW_DIMS)
    lead, tail = input_array.shape[:-k], input_array.shape[-k:]
    positions = tuple(n - w + 1 for n, w in zip(tail, shape))
    view_shape = lead + positions + tuple(shape)
    if min(view_shape, default=1) <= 0:
        raise RuntimeError(_WINDOW_TOO_BIG)
    view_strides = input_array.strides + input_array.strides[-k:]
    return np.lib.stride_tricks.as_strided(input_array, view_shape, view_strides, writeable=writeable)


def pad_and_roll(input_array: np.ndarray, shape: Tuple[int, int], **kwargs: Any) -> np.ndarray:
    """np.pad the trailing axes by half a window on each side (kwargs go to
    np.pad) and return the rolling-window view, so that the window axes can be
    reduced to an array of the original shape (neighbourhood_tools.py:69-101)."""
    writeable = kwargs.pop("writeable", False)
    n_lead = input_array.ndim - len(shape)
    widths = [(0, 0)] * n_lead + [(w // 2, w // 2) for w in shape]
    return rolling_window(np.pad(input_array, widths, **kwargs), shape, writeable=writeable)


def pad_boxsum(data: np.ndarray, boxsize: Union[int, Tuple[int, int]], **pad_options: Any) -> np.ndarray:
    """Pad for `boxsum`: half a box after, half a box plus one before, on the
    last two axes (neighbourhood_tools.py:104-126)."""
    box = np.atleast_1d(boxsize)
    half_y, half_x = int(box[0]) // 2, int(box[-1]) // 2
    widths = [(0, 0)] * (data.ndim - 2) + [(half_y + 1, half_y), (half_x + 1, half_x)]
    return np.pad(data, widths, **pad_options)


def boxsum(data: np.ndarray, boxsize: Union[int, Tuple[int, int]], cumsum: bool = True,
           **pad_options: Any) -> np.ndarray:
    """neighbourhood_tools.py:129-211.  Returns float64 like the reference does
    for float input (int64 for integer input)."""
    import torch

    from .. import engine

    boxsize = np.atleast_1d(boxsize)
    if not issubclass(boxsize.dtype.type, np.integer):
        raise ValueError("The size of the neighbourhood must be of an integer type.")
    if not np.all(boxsize % 2):
        raise ValueError("The size of the neighbourhood must be an odd number.")
    bi, bj = int(boxsize[0]), int(boxsize[-1])
    data = np.asarray(data)
    if pad_options:
        data = pad_boxsum(data, boxsize, **pad_options)
    m, n = data.shape[-2] - bi, data.shape[-1] - bj
    integer = issubclass(data.dtype.type, np.integer) or data.dtype == np.bool_
    if m <= 0 or n <= 0:
        return np.zeros(data.shape[:-2] + (max(m, 0), max(n, 0)), dtype=np.int64 if integer else np.float64)
    # float64 on the device like the reference's float64 summed-area table (any odd box, also
    # rectangular); integer input stays exact (sums below 2^53) and comes back as int64
    dev = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64)).cuda()
    res = engine.boxsum_f64(dev, bi, bj, cumsum=cumsum).cpu().numpy()
    if integer:
        return np.rint(res).astype(np.int64)
    return res
