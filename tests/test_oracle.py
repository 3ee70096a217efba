"""Pin the CPU oracle: golden vectors from the real reference + transcribed
known-answer tests of the reference's own suite (file:line cited per test)."""
import numpy as np
import pytest

from oracle import nbhood_oracle as orc
from tests.golden_cases import load_golden

G = load_golden()


def _ids(cases):
    return [c["name"] for c in cases]


@pytest.mark.parametrize("case", G.of_kind("nbhood"), ids=_ids(G.of_kind("nbhood")))
def test_neighbourhood_matches_reference_golden(case):
    data, mask, inmask = G.nbhood_inputs(case)
    inp = np.ma.masked_array(data, inmask) if inmask is not None else data
    out = orc.neighbourhood(inp, mask, method=case["method"], cells=case["cells"],
                            weighted_mode=case["weighted"], sum_only=case["sum_only"],
                            re_mask=case["re_mask"])
    np.testing.assert_array_equal(np.ma.getdata(out), G.get(case, "out"))
    assert isinstance(out, np.ma.MaskedArray) == case["masked_output"]
    if case["masked_output"]:
        np.testing.assert_array_equal(np.ma.getmaskarray(out), G.get(case, "outmask"))


@pytest.mark.parametrize("case", G.of_kind("truth"), ids=_ids(G.of_kind("truth")))
def test_truth_value_matches_reference_golden(case):
    data = G.get(case, "in")
    m = G.get(case, "inmask")
    inp = np.ma.masked_array(data, m) if m is not None else data
    tv = orc.truth_value(inp, case["threshold"], tuple(case["bounds"]), case["op"])
    assert tv.dtype == np.float32
    np.testing.assert_array_equal(np.ma.getdata(tv), G.get(case, "out"))


@pytest.mark.parametrize("case", G.of_kind("percentile"), ids=_ids(G.of_kind("percentile")))
def test_percentiles_match_reference_golden(case):
    out = orc.neighbourhood_percentiles(G.get(case, "in"), case["cells"], case["percentiles"])
    np.testing.assert_array_equal(out, G.get(case, "out"))


def test_circular_kernels_match_reference_golden():
    for r in G.of_kind("kernels")[0]["radii"]:
        for w in (0, 1):
            np.testing.assert_array_equal(orc.circular_kernel(r, bool(w)),
                                          G.arrays[f"kernel_r{r}_w{w}"])


# ---- transcribed reference KATs ------------------------------------------- #
def test_kat_circular_kernel_basic():
    """improver_tests/nbhood/nbhood/test_circular_kernel.py:22-36."""
    np.testing.assert_array_equal(
        orc.circular_kernel(1, False), [[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    np.testing.assert_array_equal(
        orc.circular_kernel(1, True), [[0, 0, 0], [0, 1, 0], [0, 0, 0]])
    k5 = orc.circular_kernel(5, False)
    assert k5.shape == (11, 11) and k5.sum() == 81


def test_kat_boxsum():
    """improver_tests/utilities/test_neighbourhood_tools.py (boxsum cases):
    5x5 ones, box 3 with zero padding -> corner 4, edge 6, interior 9."""
    res = orc.boxsum(np.ones((5, 5)), 3, mode="constant")
    assert res[0, 0] == 4 and res[0, 2] == 6 and res[2, 2] == 9
    with pytest.raises(ValueError, match="odd number"):
        orc.boxsum(np.ones((5, 5)), 4)
    with pytest.raises(ValueError, match="integer type"):
        orc.boxsum(np.ones((5, 5)), 3.0)


def test_kat_square_basic():
    """improver_tests/nbhood/nbhood/test_NeighbourhoodProcessing.py:95-109
    (`test_basic_square`): single zero at the centre of a 5x5 block of ones,
    3x3 neighbourhood -> 8/9 around it."""
    data = np.ones((5, 5), np.float32)
    data[2, 2] = 0
    out = orc.neighbourhood(data, None, "square", 1, re_mask=False)
    exp = np.ones((5, 5), np.float32)
    exp[1:4, 1:4] = 8 / 9
    np.testing.assert_allclose(out, exp, atol=1e-6)


def test_kat_fuzzy_threshold_values():
    """improver_tests/threshold/test_Threshold.py:284-361 expected values:
    thr 0.5 / fuzzy 0.5 at 0.5 -> 0.5; thr 0.4 -> 0.75;
    thr 0.8 -> 0.125; config {0.6: [0.4, 0.7]} -> 0.25."""
    v = np.array([0.5], np.float32)
    for thr, bnd, exp in [(0.5, (0.25, 0.75), 0.5), (0.4, (0.2, 0.6), 0.75),
                          (0.8, (0.4, 1.2), 0.125), (0.6, (0.4, 0.7), 0.25)]:
        np.testing.assert_allclose(orc.truth_value(v, thr, bnd, ">"), [exp], atol=1e-6)
    assert orc.fuzzy_bounds_from_factor([0.5], 0.5) == [(0.25, 0.75)]


def test_kat_radius_interpolation():
    """improver_tests/nbhood/nbhood/test_BaseNeighbourhoodProcessing.py:27-47."""
    np.testing.assert_allclose(
        orc.radii_for_lead_times([2, 3, 4], [1, 3, 5], [10000, 20000, 30000]),
        [15000, 20000, 25000])


def test_kat_percentile_docstring_example():
    """improver/nbhood/nbhood.py:515-655 worked example: 3x3 ones with a zero
    centre, r=2 circular kernel, 10th percentile -> 0.5 everywhere."""
    data = np.ones((3, 3), np.float32)
    data[1, 1] = 0
    out = orc.neighbourhood_percentiles(data, 2, [10])
    np.testing.assert_allclose(out[0], np.full((3, 3), 0.5), atol=1e-6)


def test_cells_from_distance():
    """improver/utilities/spatial.py:120-161."""
    assert orc.cells_from_distance(10000, 2000) == 5
    assert orc.cells_from_distance(11999, 2000) == 5
    with pytest.raises(ValueError, match="zero cell extent"):
        orc.cells_from_distance(1000, 2000)
    with pytest.raises(ValueError, match="positive distance"):
        orc.cells_from_distance(-1, 2000)
