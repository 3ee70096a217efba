"""Recursive filter plugin with the interface of improver/nbhood/recursive_filter.py:22-451,
backed by the CUDA kernels of csrc/nbh_recursive.cu (C-ABI nbh_recursive_filter_f32 /
nbh_recurse_f32).  Constructor arguments, validation order and error messages follow the
reference; all arithmetic runs on the GPU and there is no CPU path."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .. import PostProcessingPlugin
from .nbhood import _to_device, _yx_last


def _device_pass(grid: np.ndarray, smoothing_coefficients: np.ndarray, axis: int, forward: bool) -> np.ndarray:
    from .. import engine

    if np.asarray(grid).dtype != np.float32 or np.asarray(smoothing_coefficients).dtype != np.float32:
        raise TypeError("the CUDA recursive filter works on float32 grids and coefficients")
    dev = _to_device(grid)
    engine.recurse(dev, _to_device(smoothing_coefficients), axis, forward=forward)
    grid[...] = dev.cpu().numpy()
    return grid


class RecursiveFilter(PostProcessingPlugin):
    """Apply a recursive filter to the input cube (recursive_filter.py:22-451)."""

    def __init__(self, iterations: Optional[int] = None, edge_width: int = 15) -> None:
        """recursive_filter.py:27-53."""
        if iterations is not None:
            if iterations < 1:
                raise ValueError("Invalid number of iterations: must be >= 1: {}".format(iterations))
        self.iterations = iterations
        self.edge_width = edge_width
        self.smoothing_coefficient_name_format = "smoothing_coefficient_{}"

    @staticmethod
    def _recurse_forward(grid: np.ndarray, smoothing_coefficients: np.ndarray, axis: int) -> np.ndarray:
        """B_i = (1 - c_{i-1}) A_i + c_{i-1} B_{i-1} along `axis`, in place (:55-103)."""
        return _device_pass(grid, smoothing_coefficients, axis, True)

    @staticmethod
    def _recurse_backward(grid: np.ndarray, smoothing_coefficients: np.ndarray, axis: int) -> np.ndarray:
        """B_i = (1 - c_i) A_i + c_i B_{i+1} along `axis`, in place (:105-153)."""
        return _device_pass(grid, smoothing_coefficients, axis, False)

    @staticmethod
    def _run_recursion(cube, smoothing_coefficients_x, smoothing_coefficients_y, iterations: int):
        """`iterations` sweeps over an already padded 2-D cube, in place (:155-202).  The
        coefficient cubes have the shapes (y, x - 1) and (y - 1, x) of the padded grid."""
        from .. import engine

        (x_index,) = cube.coord_dims(cube.coord(axis="x").name())
        (y_index,) = cube.coord_dims(cube.coord(axis="y").name())
        grid = np.ascontiguousarray(np.ma.getdata(cube.data), dtype=np.float32)
        if (y_index, x_index) == (1, 0):
            grid = np.ascontiguousarray(grid.T)
        dev = _to_device(grid)
        cx = _to_device(np.asarray(smoothing_coefficients_x.data, dtype=np.float32))
        cy = _to_device(np.asarray(smoothing_coefficients_y.data, dtype=np.float32))
        for _ in range(iterations):
            engine.recurse(dev, cx, 1, forward=True)
            engine.recurse(dev, cx, 1, forward=False)
            engine.recurse(dev, cy, 0, forward=True)
            engine.recurse(dev, cy, 0, forward=False)
        out = dev.cpu().numpy()
        cube.data = out.T if (y_index, x_index) == (1, 0) else out
        return cube

    def _validate_coefficients(self, cube, smoothing_coefficients) -> List:
        """Names, the 0.5 ceiling and the coordinates of the two coefficient cubes (:204-271).
        Returns them ordered [x, y]."""
        smoothing_coefficients.sort(key=lambda cell: cell.name())
        axes = ["x", "y"]
        for axis, smoothing_coefficient in zip(axes, smoothing_coefficients):
            expected_name = self.smoothing_coefficient_name_format.format(axis)
            if smoothing_coefficient.name() != expected_name:
                msg = ("The smoothing coefficient cube name {} does not match the "
                       "expected name {}".format(smoothing_coefficient.name(), expected_name))
                raise ValueError(msg)
            if (smoothing_coefficient.data > 0.5).any():
                raise ValueError("All smoothing_coefficient values must be less than 0.5. "
                                 "A large smoothing_coefficient value leads to poor "
                                 "conservation of probabilities")
            for test_axis in axes:
                coefficient_crd = smoothing_coefficient.coord(axis=test_axis)
                points = np.asarray(cube.coord(axis=test_axis).points)
                expected_points = (points[1:] + points[:-1]) / 2 if test_axis == axis else points
                if len(coefficient_crd.points) != len(expected_points) or not np.allclose(
                        coefficient_crd.points, expected_points):
                    msg = (f"The smoothing coefficients {test_axis} dimension does not "
                           "have the expected length or values compared with the cube "
                           "to which smoothing is being applied.\n\nSmoothing "
                           "coefficient cubes must have coordinates that are:\n"
                           "- one element shorter along the dimension being smoothed "
                           f"({axis}) than in the target cube, with points in that "
                           "dimension equal to the mean of each pair of points along "
                           "the dimension in the target cube\n- equal to the points "
                           "in the target cube along the dimension not being smoothed")
                    raise ValueError(msg)
        return smoothing_coefficients

    @staticmethod
    def _spatial_last(cube) -> Tuple[np.ndarray, list, np.ndarray]:
        perm, inv = _yx_last(cube)
        data = cube.data
        arr = np.ma.transpose(data, perm) if np.ma.isMaskedArray(data) else np.transpose(data, perm)
        return arr, perm, inv

    def process(self, cube, smoothing_coefficients, variable_mask: bool = False, mask_zeros: bool = False):
        """Filter every (y, x) slice of the cube (:288-451).  One launch group covers all slices:
        the symmetric halo, the zeroing of coefficients next to masked points (and next to exact
        zeros with `mask_zeros`), the sweeps and the removal of the halo all run on the device."""
        import torch

        from .. import engine

        cube_mask = np.ma.getmaskarray(cube.data) if np.ma.isMaskedArray(cube.data) else None
        coeffs_x, coeffs_y = self._validate_coefficients(cube, smoothing_coefficients)

        arr, perm, inv = self._spatial_last(cube)
        masked = bool(np.ma.is_masked(arr))
        mask = np.ma.getmaskarray(arr) if masked else None
        if not variable_mask and masked:
            flat = mask.reshape((-1,) + mask.shape[-2:])
            if not (flat == flat[0]).all():
                raise ValueError("Input cube contains spatial slices with different masks.")
        raw = np.ma.getdata(arr)
        if raw.dtype != np.float32:
            raise TypeError(f"the CUDA recursive filter works on float32 data, not {raw.dtype}")
        cx = np.ma.getdata(coeffs_x.data)
        cy = np.ma.getdata(coeffs_y.data)
        if cx.dtype != np.float32 or cy.dtype != np.float32:
            raise TypeError("the CUDA recursive filter works on float32 smoothing coefficients")

        dev_mask = None
        if masked:
            # identical masks: one plane serves every slice
            shared = not variable_mask or raw.ndim == 2
            src = mask.reshape((-1,) + mask.shape[-2:])[0] if shared else mask
            dev_mask = torch.from_numpy(np.ascontiguousarray(src).astype(np.uint8)).cuda()
        out = engine.recursive_filter(_to_device(raw), _to_device(cx), _to_device(cy), self.iterations,
                                      self.edge_width, mask=dev_mask, mask_zeros=bool(mask_zeros))
        res = np.transpose(out.cpu().numpy(), inv)
        if mask_zeros:
            if cube_mask is not None:
                res = np.ma.array(res, mask=cube_mask)
        elif masked:
            res = np.ma.MaskedArray(res, mask=cube_mask)
        return cube.copy(data=res)
