"""Recursive filter (kernel part): the oracle against the reference's
known answers (improver_tests/nbhood/recursive_filter/test_RecursiveFilter.py) and the golden
vectors generated from the real reference functions; the CUDA path (C-ABI
nbh_recursive_filter_f32 / nbh_recurse_f32 and the RecursiveFilter plugin) bit-for-bit
against the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import recursive_filter_oracle as rfo

HERE = os.path.dirname(os.path.abspath(__file__))

# test_RecursiveFilter.py:35-48
KAT_DATA = np.array([[0.00, 0.00, 0.10, 0.00, 0.00],
                     [0.00, 0.00, 0.25, 0.00, 0.00],
                     [0.10, 0.25, 0.50, 0.25, 0.10],
                     [0.00, 0.00, 0.25, 0.00, 0.00],
                     [0.00, 0.00, 0.10, 0.00, 0.00]], dtype=np.float32)
KAT_CX = np.full((5, 4), 0.5, dtype=np.float32)
KAT_CY = np.full((4, 5), 0.5, dtype=np.float32)
# test_RecursiveFilter.py:290-298 / 308-316
KAT_FWD_AXIS0 = np.array([[0.0000, 0.00000, 0.100000, 0.00000, 0.0000],
                          [0.0000, 0.00000, 0.175000, 0.00000, 0.0000],
                          [0.0500, 0.12500, 0.337500, 0.12500, 0.0500],
                          [0.0250, 0.06250, 0.293750, 0.06250, 0.0250],
                          [0.0125, 0.03125, 0.196875, 0.03125, 0.0125]])
# test_RecursiveFilter.py:434-443 (x coefficients 0.5, y coefficients 0.25, edge_width 1)
KAT_DIFFERENT = np.array([[0.01320939, 0.02454378, 0.04346254, 0.02469828, 0.01359563],
                          [0.03405095, 0.06060188, 0.09870366, 0.06100013, 0.03504659],
                          [0.0845406, 0.13908109, 0.18816182, 0.14006987, 0.08701254],
                          [0.03405397, 0.06060749, 0.09871361, 0.06100579, 0.03504971],
                          [0.01322224, 0.02456765, 0.04350482, 0.0247223, 0.01360886]], dtype=np.float32)
# test_RecursiveFilter.py:568-578 (mask_zeros)
KAT_MASK_ZEROS = np.array([[0.0, 0.0, 0.17419356, 0.0, 0.0],
                           [0.0, 0.0, 0.21129033, 0.0, 0.0],
                           [0.2, 0.25, 0.22903226, 0.25, 0.2],
                           [0.0, 0.0, 0.21129033, 0.0, 0.0],
                           [0.0, 0.0, 0.17419356, 0.0, 0.0]])


def golden_cases():
    meta = json.load(open(os.path.join(HERE, "golden", "recursive_golden.json")))
    return meta["cases"]


@pytest.fixture(scope="module")
def golden_rf():
    return np.load(os.path.join(HERE, "golden", "recursive_golden.npz"))


def _case_input(z, c):
    n = c["name"]
    data = z[n + "_in"]
    mask = z[n + "_mask"] if c["masked"] else None
    return data, mask, z[n + "_cx"], z[n + "_cy"], z[n + "_out"]


# ---------------------------------------------------------------- oracle (CPU)
def test_oracle_directional_known_answers():
    fwd0 = rfo.recurse_forward(KAT_DATA.copy(), KAT_CY, 0)
    np.testing.assert_array_almost_equal(fwd0, KAT_FWD_AXIS0)
    fwd1 = rfo.recurse_forward(KAT_DATA.copy(), KAT_CX, 1)
    np.testing.assert_array_almost_equal(fwd1, KAT_FWD_AXIS0.T)
    # the field is symmetric, so the backward sweeps are the mirrored forward sweeps
    np.testing.assert_array_almost_equal(rfo.recurse_backward(KAT_DATA.copy(), KAT_CY, 0), KAT_FWD_AXIS0[::-1])
    np.testing.assert_array_almost_equal(rfo.recurse_backward(KAT_DATA.copy(), KAT_CX, 1), KAT_FWD_AXIS0.T[:, ::-1])


def test_oracle_process_known_answers():
    # test_RecursiveFilter.py:383-448: _run_recursion on a ZERO-padded cube (pad_cube_with_halo's
    # default method) with symmetric-padded coefficients, edge_width 1
    def run(cy, iterations):
        g = np.pad(KAT_DATA, 2, mode="constant")
        px, py = np.pad(KAT_CX, 2, mode="symmetric"), np.pad(cy, 2, mode="symmetric")
        for _ in range(iterations):
            for c, axis in ((px, 1), (py, 0)):
                rfo.recurse_forward(g, c, axis)
                rfo.recurse_backward(g, c, axis)
        return g

    assert abs(run(KAT_CY, 1)[4, 4] - 0.12302627) < 1e-7
    assert abs(run(KAT_CY, 3)[4, 4] - 0.034629755) < 1e-7
    np.testing.assert_array_almost_equal(run(KAT_CY * np.float32(0.5), 1)[2:-2, 2:-2], KAT_DIFFERENT)
    # :467-491 (default edge_width 15), :556-580
    assert abs(rfo.recursive_filter(KAT_DATA[None], KAT_CX, KAT_CY, 1)[0, 2, 2] - 0.14994797) < 1e-7
    mask = np.zeros((1, 5, 5), bool)
    mask[0, 3, 2] = True
    res = rfo.recursive_filter(np.ma.MaskedArray(KAT_DATA[None], mask), KAT_CX, KAT_CY, 1)
    assert abs(res.data[0, 2, 2] - 0.184375) < 1e-7
    np.testing.assert_array_equal(res.mask, mask)
    np.testing.assert_array_almost_equal(rfo.recursive_filter(KAT_DATA[None], KAT_CX, KAT_CY, 1, mask_zeros=True)[0],
                                         KAT_MASK_ZEROS)


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_oracle_matches_reference_golden(golden_rf, case):
    data, mask, cx, cy, want = _case_input(golden_rf, case)
    inp = np.ma.masked_array(data, mask) if mask is not None else data
    got = rfo.recursive_filter(inp, cx, cy, case["iterations"], case["edge_width"], case["mask_zeros"])
    np.testing.assert_array_equal(np.ma.getdata(got), want)


def test_golden_fixture_regenerates_from_reference(golden_rf, tmp_path):
    """When the reference tree is present (build container) the committed fixture must be what
    the generator produces today."""
    import subprocess
    import sys

    sys.path.insert(0, os.path.join(HERE, "golden"))
    import _ref_loader

    if not _ref_loader.available():
        pytest.skip("reference tree not present")
    code = ("import runpy, numpy as np, sys; sys.argv=['x'];"
            f"ns = runpy.run_path({os.path.join(HERE, 'golden', 'make_golden_recursive.py')!r});"
            "ns['case'](0, (2, 16, 21), 3, 2, True, False, 102, 'gamma');"
            f"np.save({str(tmp_path / 'o.npy')!r}, ns['arrays']['rf000_out'])")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=os.path.dirname(HERE))
    np.testing.assert_array_equal(np.load(tmp_path / "o.npy"), golden_rf["rf002_out"])


# ---------------------------------------------------------------- host logic (CPU)
def _coefficient_cubes(cube, cx=None, cy=None):
    from improver_b200.cube import Coord, Cube

    y, x = cube.coord(axis="y"), cube.coord(axis="x")
    ny, nx = y.points.size, x.points.size
    cx = np.full((ny, nx - 1), 0.5, np.float32) if cx is None else cx
    cy = np.full((ny - 1, nx), 0.5, np.float32) if cy is None else cy

    def mid(c):
        m = c.copy()
        m.points = ((c.points[1:] + c.points[:-1]) / 2).astype(np.float32)
        return m

    return [Cube(cx, "smoothing_coefficient_x", "1", [(y.copy(), 0), (mid(x), 1)]),
            Cube(cy, "smoothing_coefficient_y", "1", [(mid(y), 0), (x.copy(), 1)])]


def _kat_cube(data=KAT_DATA[None]):
    from improver_b200.synthetic_data import set_up_variable_cube

    return set_up_variable_cube(data, name="precipitation_amount", units="kg m^-2 s^-1")


def test_plugin_validation_errors():
    from improver_b200.nbhood.recursive_filter import RecursiveFilter

    with pytest.raises(ValueError, match="Invalid number of iterations: must be >= 1: -1"):
        RecursiveFilter(iterations=-1)
    plugin = RecursiveFilter(iterations=1)
    assert plugin.edge_width == 15 and plugin.iterations == 1
    cube = _kat_cube()
    good = _coefficient_cubes(cube)
    assert [c.name() for c in plugin._validate_coefficients(cube, good[::-1])] == [
        "smoothing_coefficient_x", "smoothing_coefficient_y"]
    wrong = _coefficient_cubes(cube)
    wrong[0].rename("air_temperature")
    with pytest.raises(ValueError, match="The smoothing coefficient cube name air_temperature does not match"):
        plugin._validate_coefficients(cube, wrong)
    big = _coefficient_cubes(cube, cx=np.full((5, 4), 0.75, np.float32))
    with pytest.raises(ValueError, match="All smoothing_coefficient values must be less than 0.5"):
        plugin._validate_coefficients(cube, big)
    shifted = _coefficient_cubes(cube)
    shifted[0].coord(axis="x").points = shifted[0].coord(axis="x").points + 10
    with pytest.raises(ValueError, match="The smoothing coefficients x dimension does not have the expected length"):
        plugin._validate_coefficients(cube, shifted)
    from improver_b200.cube import Cube

    six = [Cube(np.full((5, 5), 0.5, np.float32), "smoothing_coefficient_x", "1", cube.copy()._dim_coords[1:]),
           good[1]]
    six[0]._dim_coords = [(c, d - 1) for c, d in six[0]._dim_coords]
    with pytest.raises(ValueError, match="The smoothing coefficients x dimension does not have the expected length"):
        plugin._validate_coefficients(cube, six)


def test_plugin_rejects_different_masks_before_any_launch():
    from improver_b200.nbhood.recursive_filter import RecursiveFilter
    from improver_b200.synthetic_data import set_up_variable_cube

    data = np.stack([KAT_DATA, KAT_DATA])
    mask = np.zeros(data.shape, int)
    mask[0, 2, 2] = 1
    mask[1, 2, 3] = 1
    cube = set_up_variable_cube(np.ma.MaskedArray(data, mask=mask))
    with pytest.raises(ValueError, match="Input cube contains spatial slices with different masks."):
        RecursiveFilter(iterations=1)(cube, _coefficient_cubes(cube))


# ---------------------------------------------------------------- CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_cuda_filter_matches_reference_golden(golden_rf, case):
    import torch

    from improver_b200 import engine

    data, mask, cx, cy, want = _case_input(golden_rf, case)
    dev_mask = torch.from_numpy(mask.astype(np.uint8)).cuda() if mask is not None else None
    out = engine.recursive_filter(torch.from_numpy(data).cuda(), torch.from_numpy(cx).cuda(),
                                  torch.from_numpy(cy).cuda(), case["iterations"], case["edge_width"],
                                  mask=dev_mask, mask_zeros=case["mask_zeros"])
    np.testing.assert_array_equal(out.cpu().numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,edge_width,iterations", [((3, 97, 131), 15, 2), ((2, 40, 1042), 4, 1),
                                                         ((1, 970, 61), 15, 1), ((4, 8, 8), 0, 3)])
@pytest.mark.parametrize("masking", ["none", "shared", "per_slice", "zeros"])
def test_cuda_filter_matches_oracle(shape, edge_width, iterations, masking):
    import torch

    from improver_b200 import engine

    rng = np.random.default_rng(7)
    ny, nx = shape[-2:]
    data = rng.random(shape).astype(np.float32)
    data[data < 0.3] = 0.0
    cx = (rng.random((ny, nx - 1)) * 0.5).astype(np.float32)
    cy = (rng.random((ny - 1, nx)) * 0.5).astype(np.float32)
    mask = None
    if masking == "shared":
        mask = np.broadcast_to(rng.random((ny, nx)) < 0.2, shape).copy()
    elif masking == "per_slice":
        mask = rng.random(shape) < 0.2
    inp = np.ma.masked_array(data, mask) if mask is not None else data
    want = np.ma.getdata(rfo.recursive_filter(inp, cx, cy, iterations, edge_width, mask_zeros=masking == "zeros"))
    dev_mask = None
    if mask is not None:
        dev_mask = torch.from_numpy((mask[0] if masking == "shared" else mask).astype(np.uint8)).cuda()
    out = engine.recursive_filter(torch.from_numpy(data).cuda(), torch.from_numpy(cx).cuda(),
                                  torch.from_numpy(cy).cuda(), iterations, edge_width, mask=dev_mask,
                                  mask_zeros=masking == "zeros")
    np.testing.assert_array_equal(out.cpu().numpy(), want)


@pytest.mark.gpu
def test_cuda_directional_sweeps_match_oracle_and_known_answers():
    from improver_b200.nbhood.recursive_filter import RecursiveFilter

    np.testing.assert_array_almost_equal(RecursiveFilter._recurse_forward(KAT_DATA.copy(), KAT_CY, 0), KAT_FWD_AXIS0)
    np.testing.assert_array_almost_equal(RecursiveFilter._recurse_forward(KAT_DATA.copy(), KAT_CX, 1),
                                         KAT_FWD_AXIS0.T)
    rng = np.random.default_rng(11)
    for ny, nx in [(5, 5), (33, 70), (64, 9), (2, 2), (130, 257)]:
        grid = rng.random((ny, nx)).astype(np.float32)
        for axis in (0, 1):
            c = (rng.random((ny - 1, nx) if axis == 0 else (ny, nx - 1)) * 0.5).astype(np.float32)
            np.testing.assert_array_equal(RecursiveFilter._recurse_forward(grid.copy(), c, axis),
                                          rfo.recurse_forward(grid.copy(), c, axis))
            np.testing.assert_array_equal(RecursiveFilter._recurse_backward(grid.copy(), c, axis),
                                          rfo.recurse_backward(grid.copy(), c, axis))


@pytest.mark.gpu
def test_plugin_known_answers_on_gpu():
    """test_RecursiveFilter.py:448-600 through the plugin."""
    from improver_b200.nbhood.recursive_filter import RecursiveFilter
    from improver_b200.synthetic_data import set_up_variable_cube

    cube = _kat_cube()
    result = RecursiveFilter(iterations=1)(cube, smoothing_coefficients=_coefficient_cubes(cube))
    assert result.shape == (1, 5, 5) and result.data.dtype == np.float32
    assert abs(result.data[0][2][2] - 0.14994797) < 1e-7
    # masked input (:479-491)
    mask = np.zeros((1, 5, 5))
    mask[0][3][2] = 1
    mcube = _kat_cube(np.ma.MaskedArray(KAT_DATA[None], mask=mask))
    result = RecursiveFilter(iterations=1)(mcube, smoothing_coefficients=_coefficient_cubes(mcube))
    assert abs(result.data[0][2][2] - 0.184375) < 1e-7
    np.testing.assert_array_equal(result.data.mask, mask)
    # mask_zeros (:523-580)
    result = RecursiveFilter(iterations=1)(cube, _coefficient_cubes(cube), mask_zeros=True)
    np.testing.assert_array_almost_equal(result.data, KAT_MASK_ZEROS[None])
    assert not np.ma.isMaskedArray(result.data)
    np.testing.assert_array_equal(result.data == 0.0, cube.data == 0.0)
    mask = np.zeros((1, 5, 5))
    mask[0, 3, :] = 1
    mcube = _kat_cube(np.ma.MaskedArray(KAT_DATA[None], mask=mask))
    result = RecursiveFilter(iterations=1)(mcube, _coefficient_cubes(mcube), mask_zeros=True)
    np.testing.assert_array_equal(np.ma.getmaskarray(result.data), mask.astype(bool))
    # a different mask per slice with variable_mask (:582-601)
    data = np.stack([KAT_DATA, KAT_DATA])
    mask = np.zeros(data.shape, dtype=int)
    mask[1] = data[1] == 0.0
    pcube = set_up_variable_cube(np.ma.MaskedArray(data, mask=mask))
    result = RecursiveFilter(iterations=1)(pcube, _coefficient_cubes(pcube), variable_mask=True)
    for i, expected in enumerate([0.14994797, 0.22903226]):
        np.testing.assert_array_equal(result.data[i].mask, mask[i])
        assert abs(result.data[i][2][2] - expected) < 1e-7
    # different coefficients in x and y, cube stored (realization, x, y) (:494-521)
    expected = np.array([[0.02554158, 0.05397786, 0.1312837, 0.05397786, 0.02554158],
                         [0.03596632, 0.07334216, 0.1668669, 0.07334216, 0.03596632],
                         [0.05850913, 0.11031596, 0.21073693, 0.11031596, 0.05850913],
                         [0.03596632, 0.07334216, 0.1668669, 0.07334216, 0.03596632],
                         [0.02554158, 0.05397786, 0.1312837, 0.05397786, 0.02554158]], dtype=np.float32)
    tcube = _kat_cube()
    coeffs = _coefficient_cubes(tcube, cy=np.full((4, 5), 0.25, np.float32))
    tcube.data = np.ascontiguousarray(np.swapaxes(tcube.data, 1, 2))
    tcube._dim_coords = [(c, {0: 0, 1: 2, 2: 1}[d]) for c, d in tcube._dim_coords]
    result = RecursiveFilter(iterations=1)(tcube, smoothing_coefficients=coeffs)
    np.testing.assert_array_almost_equal(result.data[0], expected)


@pytest.mark.gpu
def test_run_recursion_known_answers_on_gpu():
    """test_RecursiveFilter.py:383-448: _run_recursion on an already padded cube."""
    from improver_b200.nbhood.recursive_filter import RecursiveFilter
    from improver_b200.synthetic_data import set_up_variable_cube

    pad = 2
    padded = set_up_variable_cube(np.pad(KAT_DATA, pad, mode="constant"))
    px = set_up_variable_cube(np.pad(KAT_CX, pad, mode="symmetric"))
    for cy_scale, iterations, check in [(1.0, 1, 0.12302627), (1.0, 3, 0.034629755), (0.5, 1, None)]:
        py = set_up_variable_cube(np.pad(KAT_CY * np.float32(cy_scale), pad, mode="symmetric"))
        result = RecursiveFilter(edge_width=1)._run_recursion(padded.copy(), px, py, iterations)
        if check is not None:
            assert abs(result.data[4][4] - check) < 1e-7
        else:
            np.testing.assert_array_almost_equal(result.data[2:-2, 2:-2], KAT_DIFFERENT)


@pytest.mark.gpu
def test_cuda_filter_launch_groups_are_equivalent(monkeypatch):
    """Slices are independent: splitting a call into launch groups (workspace budget) must not
    change a bit, with shared and with per-slice masks."""
    import torch

    from improver_b200 import engine

    rng = np.random.default_rng(21)
    data = rng.random((7, 45, 52)).astype(np.float32)
    data[data < 0.3] = 0.0
    mask = rng.random(data.shape) < 0.2
    cx = (rng.random((45, 51)) * 0.5).astype(np.float32)
    cy = (rng.random((44, 52)) * 0.5).astype(np.float32)
    args = (torch.from_numpy(data).cuda(), torch.from_numpy(cx).cuda(), torch.from_numpy(cy).cuda(), 2, 3)
    dev_mask = torch.from_numpy(mask.astype(np.uint8)).cuda()
    whole = [engine.recursive_filter(*args, mask=m, mask_zeros=z).cpu().numpy()
             for m, z in ((None, False), (dev_mask, False), (None, True))]
    split = [engine.recursive_filter(*args, mask=m, mask_zeros=z, max_slices_per_call=3).cpu().numpy()
             for m, z in ((None, False), (dev_mask, False), (None, True))]
    for a, b in zip(whole, split):
        np.testing.assert_array_equal(a, b)
    want = rfo.recursive_filter(np.ma.masked_array(data, mask), cx, cy, 2, 3)
    np.testing.assert_array_equal(split[1], np.ma.getdata(want))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 150, 40), (5, 97, 131), (37, 40, 70)])
def test_cuda_fused_sweeps_match_single_sweeps_and_oracle(shape):
    """The tile-row wavefront (y-forward inside the x-backward sweep, ticket-ordered) is the same
    arithmetic as the four single directional sweeps (nbh_recurse_f32 on the padded planes) and as
    the oracle: bit-identical, with one slice (all tile rows of a slice in one block) and with more
    slices than a block holds."""
    import torch

    from improver_b200 import engine

    rng = np.random.default_rng(33)
    ny, nx = shape[-2:]
    data = rng.random(shape).astype(np.float32)
    data[data < 0.25] = 0.0
    cx = (rng.random((ny, nx - 1)) * 0.5).astype(np.float32)
    cy = (rng.random((ny - 1, nx)) * 0.5).astype(np.float32)
    args = (torch.from_numpy(data).cuda(), torch.from_numpy(cx).cuda(), torch.from_numpy(cy).cuda())
    for mask_zeros in (False, True):
        fused = engine.recursive_filter(*args, 2, 15, mask_zeros=mask_zeros).cpu().numpy()
        again = engine.recursive_filter(*args, 2, 15, mask_zeros=mask_zeros, max_slices_per_call=2).cpu().numpy()
        np.testing.assert_array_equal(fused, again)
        want = rfo.recursive_filter(data, cx, cy, 2, 15, mask_zeros=mask_zeros)
        np.testing.assert_array_equal(fused, np.ma.getdata(want))
