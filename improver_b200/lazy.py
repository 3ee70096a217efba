"""Deferred, device-resident cube data — the analogue of iris' lazy data for this package.

A cube whose data is a `Deferred` is not computed when a plugin is applied to it: the plugin
records what it would do and returns a cube holding another `Deferred`.  The work happens when
the data is asked for (`cube.data`, `np.asarray`, `.realise()`), and by then the whole chain is
known:

    Threshold(...)(cube) -> NeighbourhoodProcessing(...)(cube) -> collapse_realizations(cube)

is recognised and runs as ONE fused kernel launch per lead-time chunk (improver/threshold.py:571-755,
improver/nbhood/nbhood.py:411-465, improver/utilities/cube_manipulation.py:93-131), with the input
streamed from pinned host memory when it lives on the host; anything else runs step by step on the
device without host round trips between the plugins.  `defer(cube)` opts a cube in; cubes holding
plain numpy data keep the eager behaviour.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

FLOAT_DTYPE = np.float32


class Deferred:
    """Array-like placeholder: shape / dtype are known, the values are produced on demand."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(FLOAT_DTYPE)
        self._host = None        # realised numpy array (float32, or MaskedArray)
        self._dev = None         # realised CUDA tensor

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape)) if self.shape else 1

    # -- evaluation ----------------------------------------------------------
    def device(self):
        """CUDA float32 tensor of the values (computed on first use)."""
        if self._dev is None:
            self._dev = self._compute_device()
        return self._dev

    def realise(self):
        """Host numpy array of the values (computed on first use)."""
        if self._host is None:
            self._host = self._compute_host()
        return self._host

    def _compute_device(self):
        raise NotImplementedError

    def _compute_host(self):
        import torch

        dev = self.device()
        torch.cuda.current_stream(dev.device).synchronize()
        return dev.cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        arr = self.realise()
        return arr if dtype is None else np.asarray(arr, dtype=dtype)

    def copy(self):
        return self              # immutable: sharing is safe

    def __repr__(self):
        return f"<{type(self).__name__} shape={self.shape}>"


def _check(status):
    from . import engine

    engine.check_status(status)


class Source(Deferred):
    """Input data: a host numpy array / CPU tensor (preferably pinned) or a CUDA tensor,
    float32, spatial axes last."""

    def __init__(self, data):
        import torch

        if isinstance(data, np.ndarray):
            if isinstance(data, np.ma.MaskedArray):
                raise TypeError("masked arrays are not deferred (use the eager plugins)")
            data = np.ascontiguousarray(data, dtype=np.float32)
            self.host_tensor = torch.from_numpy(data)
            super().__init__(data.shape)
            self._host = data
        elif isinstance(data, torch.Tensor):
            if data.dtype != torch.float32:
                raise TypeError("deferred data must be float32")
            super().__init__(tuple(data.shape))
            if data.is_cuda:
                self.host_tensor = None
                self._dev = data
            else:
                self.host_tensor = data.contiguous()
                self._host = self.host_tensor.numpy()
        else:
            raise TypeError("Source needs a numpy array or a torch tensor")

    @property
    def on_host(self) -> bool:
        return self._dev is None

    def _compute_device(self):
        return self.host_tensor.cuda(non_blocking=True)


class ThresholdOp(Deferred):
    """Threshold.process without vicinity / collapse (threshold.py:643-713, 744-753): truth values
    with a threshold axis, dimensions reordered and squeezed like the reference does."""

    def __init__(self, src: Deferred, thresholds, axes: Sequence[int], squeeze_axes: Sequence[int]):
        self.src = src
        self.thresholds = list(thresholds)
        self.axes = tuple(axes)                   # permutation of (T, *src.shape)
        self.squeeze_axes = tuple(squeeze_axes)   # axes of the permuted array that are dropped
        full = (len(self.thresholds),) + tuple(src.shape)
        permuted = [full[a] for a in self.axes]
        super().__init__([s for i, s in enumerate(permuted) if i not in self.squeeze_axes])

    def layout(self, t):
        """Apply the reference's dimension order to a (T, *src.shape) tensor."""
        t = t.permute(*self.axes)
        for a in sorted(self.squeeze_axes, reverse=True):
            t = t.squeeze(a)
        return t

    def _compute_device(self):
        from . import engine

        out, _, status = engine.threshold(self.src.device(), self.thresholds)
        _check(status)
        return self.layout(out).contiguous()


class NbhoodOp(Deferred):
    """NeighbourhoodProcessing.process (nbhood.py:411-465) on deferred data with (y, x) last."""

    def __init__(self, src: Deferred, method: str, cells: int, weighted: bool, sum_only: bool,
                 mask2d: Optional[np.ndarray], re_mask: bool):
        self.src = src
        self.method, self.cells, self.weighted, self.sum_only = method, int(cells), bool(weighted), bool(sum_only)
        self.mask2d = None if mask2d is None else (np.asarray(mask2d) != 0).astype(np.uint8)
        self.re_mask = bool(re_mask)
        super().__init__(src.shape)

    def _mask_dev(self, device):
        import torch

        return None if self.mask2d is None else torch.from_numpy(self.mask2d).to(device)

    def _compute_device(self):
        from . import engine

        dev = self.src.device()
        out, _, status = engine.neighbourhood(dev, self.cells, method=self.method, mask2d=self._mask_dev(dev.device),
                                              weighted_mode=self.weighted, sum_only=self.sum_only)
        _check(status)
        return out

    def _compute_host(self):
        arr = super()._compute_host()
        if self.re_mask:
            mask = np.zeros(arr.shape, dtype=bool) if self.mask2d is None else \
                np.broadcast_to(self.mask2d == 0, arr.shape).copy()
            arr = np.ma.masked_array(arr, mask)
        return arr


class CollapseOp(Deferred):
    """Mean over one axis (collapse_realizations, cube_manipulation.py:93-131, iris MEAN)."""

    def __init__(self, src: Deferred, axis: int):
        self.src = src
        self.axis = int(axis)
        super().__init__(tuple(s for i, s in enumerate(src.shape) if i != self.axis))

    # -- the fused chain -----------------------------------------------------------
    def _fusable(self):
        """(source, threshold op, nbhood op) when this is mean_realization(nbhood(threshold(source)))
        in a form the fused kernel takes: source (realization, [time,] y, x), thresholded layout
        ([time,] realization, threshold, y, x) with nothing squeezed, means (not sums)."""
        nb = self.src
        if not isinstance(nb, NbhoodOp) or nb.sum_only or nb.weighted:
            return None
        th = nb.src
        if not isinstance(th, ThresholdOp) or not isinstance(th.src, Source):
            return None
        src = th.src
        # a single threshold: the reference squeezes the length-1 threshold axis (threshold.py:731)
        one = len(th.thresholds) == 1
        if src.ndim == 4 and tuple(th.axes) == (2, 1, 0, 3, 4) and self.axis == 1 and \
                tuple(th.squeeze_axes) == ((2,) if one else ()):
            return src, th, nb
        if src.ndim == 3 and tuple(th.axes) == (1, 0, 2, 3) and self.axis == 0 and \
                tuple(th.squeeze_axes) == ((1,) if one else ()):
            return src, th, nb
        return None

    def _drop_threshold_axis(self, arr, th):
        """(..., T, y, x) of the fused kernel -> the squeezed layout when T == 1."""
        return arr.squeeze(-3) if len(th.thresholds) == 1 else arr

    def _compute_device(self):
        from . import engine

        fused = self._fusable()
        if fused is None:
            dev = self.src.device()
            return engine.collapse_mean(dev.movedim(self.axis, 0).contiguous())
        src, th, nb = fused
        dev = src.device()
        out, _, status = engine.neighbourhood(dev, nb.cells, method=nb.method, mask2d=nb._mask_dev(dev.device),
                                              thresholds=th.thresholds, collapse_members=True)
        _check(status)
        return self._drop_threshold_axis(out, th)

    def _compute_host(self):
        fused = self._fusable()
        if fused is not None and fused[0].on_host:
            # host cube: lead-time chunks streamed through pinned buffers, H2D overlapping the kernels
            from . import pipeline

            src, th, nb = fused
            host = src.host_tensor if src.ndim == 4 else src.host_tensor.unsqueeze(1)
            res = pipeline.threshold_neighbourhood_mean_host(
                host, [t[0] for t in th.thresholds], [(t[1], t[2]) for t in th.thresholds], th.thresholds[0][3],
                nb.cells, method=nb.method, mask2d=nb.mask2d, out=self._out_buffer)
            arr = res.numpy()
            return self._drop_threshold_axis(arr if src.ndim == 4 else arr[0], th)
        return super()._compute_host()

    _out_buffer = None      # optional pinned (L, T, ny, nx) CPU tensor receiving the fused host result


def defer(cube):
    """Return a copy of `cube` whose data is deferred (a `Source` over the same buffer: host numpy,
    pinned CPU tensor or CUDA tensor).  Plugins applied to it build a chain that is evaluated —
    fused where possible — when the result's data is asked for."""
    core = cube.core_data() if hasattr(cube, "core_data") else cube.data
    if isinstance(core, Deferred):
        return cube
    return cube.copy(data=Source(core))


def is_deferred(cube) -> bool:
    core = cube.core_data() if hasattr(cube, "core_data") else getattr(cube, "data", None)
    return isinstance(core, Deferred)
