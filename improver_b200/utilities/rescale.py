"""Linear rescaling (improver/utilities/rescale.py:14-69), evaluated on the
GPU with the reference's float32 operation order when a CUDA tensor is given
and with numpy for host scalars / small arrays (it is not on the hot path:
the fuzzy threshold kernels inline the same arithmetic)."""
from __future__ import annotations

import numpy as np


def rescale(data, data_range=None, scale_range=(0.0, 1.0), clip: bool = False):
    try:
        import torch
        is_tensor = isinstance(data, torch.Tensor)
    except ImportError:   # pragma: no cover
        is_tensor = False
    data_left = (data.min() if data_range is None else data_range[0])
    data_right = (data.max() if data_range is None else data_range[1])
    if is_tensor and data_range is None:
        data_left, data_right = float(data_left), float(data_right)
    scale_left, scale_right = scale_range[0], scale_range[1]
    if data_left == data_right:
        raise ValueError("Cannot rescale a zero input range ({} -> {})".format(data_left, data_right))
    if scale_left == scale_right:
        raise ValueError("Cannot rescale a zero output range ({} -> {})".format(scale_left, scale_right))
    result = ((data - data_left) * (scale_right - scale_left) / (data_right - data_left)) + scale_left
    if clip:
        lo, hi = min(scale_left, scale_right), max(scale_left, scale_right)
        result = result.clamp(lo, hi) if is_tensor else np.clip(result, lo, hi)
    return result
