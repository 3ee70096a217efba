"""Threshold-in-vicinity (improver/threshold.py:473-518,
improver/utilities/spatial.py:740-855): the oracle against golden vectors made by
the real reference functions (tests/golden/make_golden_vicinity.py), the CUDA
path against both (bit-exact), and the plugin surface."""
import json
import os

import numpy as np
import pytest

from oracle import nbhood_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "vicinity_golden.json")) as fh:
    CASES = json.load(fh)
ARR = np.load(os.path.join(HERE, "golden", "vicinity_golden.npz"))


def _inputs(case):
    data = ARR[case["name"] + "_in"]
    if case["masked"]:
        data = np.ma.masked_array(data, ARR[case["name"] + "_inmask"])
    land = ARR[case["name"] + "_land"] if case["has_land"] else None
    return data, land


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_golden(case):
    data, land = _inputs(case)
    out, inv, contrib = orc.threshold_vicinity_cube(
        data, case["thresholds"], [tuple(b) for b in case["bounds"]], case["op"], case["cells"],
        landmask=land, collapse_leading=True)
    np.testing.assert_array_equal(contrib, ARR[case["name"] + "_contrib"])
    np.testing.assert_array_equal(out, ARR[case["name"] + "_out"])
    assert (inv is not None) == bool((ARR[case["name"] + "_contrib"] == 0).any())


# improver_tests/threshold/test_Threshold.py:498-560 (vicinity) and :724-765 (land mask):
# (threshold values, vicinity radii in metres on the 2 km grid, data, land mask, expected)
_M = np.ma.masked_array
REFERENCE_KATS = [
    ([0.5], [3000], np.array([[0, 0, 0], [0, 1.0, 0.0], [0, 0, 0]]), None, np.ones((1, 1, 3, 3), np.float32)),
    ([0.5], [3000], np.array([[1.0, 0, 0], [0, 0, 0.0], [0, 0, 0]]), None,
     np.array([[1.0, 1.0, 0], [1.0, 1.0, 0.0], [0, 0, 0]], np.float32)[None, None]),
    ([0.5], [3000], _M([[0, 2.0, 0], [0, 0, 0.0], [0, 0, 0]], mask=[[0, 1, 0], [0, 0, 0], [0, 0, 0]]), None,
     np.zeros((1, 1, 3, 3), np.float32)),
    ([0.5, 1.5], [3000, 5000], np.r_[[1], [0] * 24].reshape((5, 5)).astype(np.float32), None,
     np.stack([np.stack([np.r_[[1] * 2, [0] * 3, [1] * 2, [0] * 18].reshape((5, 5)),
                         np.r_[[1] * 3, [0] * 2, [1] * 3, [0] * 2, [1] * 3, [0] * 12].reshape((5, 5))]),
               np.zeros((2, 5, 5))]).astype(np.float32).transpose(1, 0, 2, 3)),
    ([0.5], [3000], np.array([[0, 0, 1.0], [0, 0, 0], [0, 0, 0]]), np.array([[0, 0, 1], [0, 0, 0], [0, 0, 0]]),
     np.array([[0, 0, 1.0], [0, 0, 0], [0, 0, 0]], np.float32)[None, None]),
    ([0.5], [3000], np.array([[0, 0, 0], [0, 0, 0], [1.0, 0, 0]]), np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]]),
     np.array([[0, 0, 0], [1.0, 0, 0], [1.0, 1.0, 0]], np.float32)[None, None]),
    ([0.5], [5000], np.array([[0, 0, 0], [0, 0, 0], [1.0, 0, 0]]), np.array([[1, 0, 1], [0, 1, 0], [0, 1, 0]]),
     np.array([[0, 1.0, 0], [1.0, 0, 1.0], [1.0, 0, 1.0]], np.float32)[None, None]),
]


@pytest.mark.parametrize("thr,vic,data,land,expected", REFERENCE_KATS)
def test_oracle_reference_kats(thr, vic, data, land, expected):
    """The reference's own known answers, at array level: expected is (V, T, y, x)."""
    d = data.astype(np.float32)
    cells = [int(v / 2000.0) for v in vic]
    out, _, _ = orc.threshold_vicinity_cube(d, thr, [(t, t) for t in thr], ">", cells,
                                            landmask=None if land is None else land.astype(bool))
    np.testing.assert_array_equal(out, expected)


# ----------------------------------------------------------------------------- GPU
torch = pytest.importorskip("torch")


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_matches_reference_golden_bit_exact(case):
    from improver_b200 import engine

    data, land = _inputs(case)
    raw = np.ma.getdata(data)
    mnd = torch.from_numpy(np.ma.getmaskarray(data).astype(np.uint8)).cuda() if case["masked"] else None
    thr = [(t, b[0], b[1], case["op"]) for t, b in zip(case["thresholds"], case["bounds"])]
    out, contrib, status = engine.threshold_vicinity(
        torch.from_numpy(raw).cuda(), thr, case["cells"], mask_nd=mnd,
        landmask=torch.from_numpy(land.astype(np.uint8)).cuda() if land is not None else None,
        collapse_leading=True)
    engine.check_status(status)
    exp = ARR[case["name"] + "_out"]
    cnt = ARR[case["name"] + "_contrib"]
    np.testing.assert_array_equal(contrib.cpu().numpy(), cnt)
    got = out.cpu().numpy()
    valid = np.broadcast_to(cnt > 0, exp.shape)
    np.testing.assert_array_equal(got[valid], exp[valid])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,cells", [((5, 2, 61, 86), [2]), ((1, 3, 40, 130), [1, 5]), ((3, 97, 33), [10])])
def test_cuda_matches_oracle_random(shape, cells):
    from improver_b200 import engine

    rng = np.random.default_rng(9)
    data = rng.gamma(2.0, 1.5, shape).astype(np.float32)
    land = rng.random(shape[-2:]) < 0.5
    thr, bounds = [0.5, 2.0, 4.0], [(0.35, 0.65), (1.4, 2.6), (4.0, 4.0)]
    for lm in (None, land):
        for op in (">", "<="):
            tl = [(t, b[0], b[1], op) for t, b in zip(thr[:2], bounds[:2])]
            exp, _, cnt = orc.threshold_vicinity_cube(data, thr[:2], bounds[:2], op, cells, landmask=lm,
                                                      collapse_leading=True)
            out, contrib, status = engine.threshold_vicinity(
                torch.from_numpy(data).cuda(), tl, cells,
                landmask=torch.from_numpy(lm.astype(np.uint8)).cuda() if lm is not None else None,
                collapse_leading=True)
            engine.check_status(status)
            np.testing.assert_array_equal(out.cpu().numpy(), exp)
            np.testing.assert_array_equal(contrib.cpu().numpy(), cnt)
    # no collapse: every leading slice is thresholded on its own
    exp, _, _ = orc.threshold_vicinity_cube(data, [2.0], [(2.0, 2.0)], ">", cells, landmask=land)
    out, _, status = engine.threshold_vicinity(torch.from_numpy(data).cuda(), [(2.0, 2.0, 2.0, ">")], cells,
                                               landmask=torch.from_numpy(land.astype(np.uint8)).cuda())
    engine.check_status(status)
    np.testing.assert_array_equal(out.cpu().numpy(), exp)


@pytest.mark.gpu
def test_cuda_vicinity_nan_raises():
    from improver_b200 import engine

    data = np.random.default_rng(0).random((2, 20, 30)).astype(np.float32)
    data[1, 3, 4] = np.nan
    _, _, status = engine.threshold_vicinity(torch.from_numpy(data).cuda(), [(0.5, 0.5, 0.5, ">")], [1],
                                             collapse_leading=True)
    with pytest.raises(ValueError, match="NaN detected"):
        engine.check_status(status)


@pytest.mark.gpu
@pytest.mark.parametrize("thr,vic,data,land,expected", REFERENCE_KATS)
def test_plugin_reference_kats(thr, vic, data, land, expected):
    """Threshold(vicinity=...) on the cube double: values, masks, dimension order
    (..., threshold, radius_of_vicinity, y, x squeezed: threshold.py:731, :744-753), name
    (spatial.py:1009-1023) and radius_of_vicinity coordinate (spatial.py:1026-1060)."""
    from improver_b200.synthetic_data import set_up_variable_cube
    from improver_b200.threshold import Threshold

    cube = set_up_variable_cube(data.astype(np.float32), name="precipitation_amount", units="m")
    lm = None if land is None else set_up_variable_cube(land.astype(np.float32), name="land_binary_mask", units="1")
    result = Threshold(threshold_values=thr if len(thr) > 1 else thr[0],
                       vicinity=vic if len(vic) > 1 else vic[0])(cube, lm)
    exp = np.transpose(expected, (1, 0, 2, 3)).squeeze()       # (T, V, y, x), squeezed like the reference
    assert result.data.shape == exp.shape
    np.testing.assert_array_equal(np.ma.getdata(result.data), exp)
    assert result.data.dtype == np.float32
    if isinstance(data, np.ma.MaskedArray):
        np.testing.assert_array_equal(np.ma.getmaskarray(result.data), np.ma.getmaskarray(data))
    assert result.name() == "probability_of_precipitation_amount_in_vicinity_above_threshold"
    rv = result.coord("radius_of_vicinity")
    np.testing.assert_array_equal(rv.points, np.asarray(vic, np.float32))
    assert str(rv.units) == "m"


@pytest.mark.gpu
def test_plugin_vicinity_collapse_and_errors():
    """threshold.py:760-768 KAT (collapse over realization with a land mask, both members
    identical) and the land-mask-without-vicinity error (:619-622)."""
    from improver_b200.synthetic_data import set_up_variable_cube
    from improver_b200.threshold import Threshold

    one = np.array([[0, 0, 0], [0, 0, 0], [1.0, 0, 0]], np.float32)
    land = np.array([[1, 0, 1], [0, 1, 0], [0, 1, 0]], np.float32)
    cube = set_up_variable_cube(np.stack([one, one]), name="precipitation_amount", units="m")
    lm = set_up_variable_cube(land, name="land_binary_mask", units="1")
    result = Threshold(threshold_values=0.5, vicinity=5000, collapse_coord="realization")(cube, lm)
    np.testing.assert_array_equal(result.data, np.array([[0, 1.0, 0], [1.0, 0, 1.0], [1.0, 0, 1.0]], np.float32))
    assert not result.coords("realization")
    with pytest.raises(ValueError, match="Cannot apply land-mask cube without in-vicinity processing"):
        Threshold(threshold_values=0.5)(cube, lm)


@pytest.mark.gpu
@pytest.mark.parametrize("operator", ["max", "min"])
def test_occurrence_within_vicinity_vs_reference_semantics(operator):
    """OccurrenceWithinVicinity (spatial.py:1077-1255) for the max / min operators: oracle
    (operator_within_vicinity restated with scipy's maximum_filter) on every 2-D slice,
    land / sea separation, masked points kept and re-masked, radius axis after the leading dims."""
    from improver_b200.synthetic_data import set_up_variable_cube
    from improver_b200.utilities.spatial import OccurrenceWithinVicinity

    rng = np.random.default_rng(12)
    data = rng.random((3, 21, 34)).astype(np.float32)
    land = (rng.random((21, 34)) < 0.4)
    lm = set_up_variable_cube(land.astype(np.float32), name="land_binary_mask", units="1")
    sign = 1.0 if operator == "max" else -1.0
    for masked in (False, True):
        inp = np.ma.masked_array(data, rng.random(data.shape) < 0.15) if masked else data
        cube = set_up_variable_cube(inp, name="lwe_precipitation_rate", units="m s-1")
        result = OccurrenceWithinVicinity(grid_point_radii=[1, 3], land_mask_cube=lm, operator=operator)(cube)
        assert result.shape == (3, 2, 21, 34)
        assert result.name() == "lwe_precipitation_rate_in_vicinity"
        for iv, cells in enumerate((1, 3)):
            for k in range(3):
                sl = inp[k]
                neg = sign * sl if not masked else np.ma.masked_array(sign * sl.data, sl.mask)
                exp = sign * np.ma.getdata(orc.maximum_within_vicinity(neg, cells, land))
                got = result.data[k, iv]
                if masked:
                    np.testing.assert_array_equal(np.ma.getmaskarray(got), sl.mask)
                    np.testing.assert_array_equal(np.ma.getdata(got)[~sl.mask], exp[~sl.mask])
                    np.testing.assert_array_equal(np.ma.getdata(got)[sl.mask], sl.data[sl.mask])
                else:
                    np.testing.assert_array_equal(got, exp)
    single = OccurrenceWithinVicinity(radii=[4000.0])(set_up_variable_cube(data[0], name="probability_of_rain_above_threshold", units="1"))
    assert single.shape == (21, 34)
    assert single.name() == "probability_of_rain_in_vicinity_above_threshold"
    np.testing.assert_array_equal(single.data, orc.maximum_within_vicinity(data[0], 2))
    with pytest.raises(ValueError, match="only one of radii"):
        OccurrenceWithinVicinity(radii=[2000], grid_point_radii=[1])
    with pytest.raises(ValueError, match="non-zero value"):
        OccurrenceWithinVicinity()
