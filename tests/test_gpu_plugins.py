"""Plugin-surface parity on the GPU: transcribed known-answer tests of the
reference's own suite (file:line cited) plus cube-level runs against the CPU
oracle.  Tolerance 1e-5 on probabilities, masks bit-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from improver_b200.nbhood.nbhood import (  # noqa: E402
    GeneratePercentilesFromANeighbourhood, MetaNeighbourhood, NeighbourhoodProcessing, circular_kernel)
from improver_b200.synthetic_data import set_up_probability_cube, set_up_variable_cube  # noqa: E402
from improver_b200.threshold import Threshold  # noqa: E402
from oracle import nbhood_oracle as orc  # noqa: E402

RADIUS = 2500


def _kat_data():
    data = np.ones((5, 5), dtype=np.float32)
    data[2, 2] = 0
    return data


MASKED_DATA = np.array([[1.0, 1.0, 0.0, 1.0, 1.0], [1.0, 1.0, 1.0, 0.0, 0.0], [1.0, 0.0, 1.0, 0.0, 0.0],
                        [0.0, 0.0, 1.0, 1.0, 0.0], [0.0, 1.0, 1.0, 0.0, 1.0]])
MASK = np.array([[0, 0, 1, 1, 0], [0, 1, 1, 1, 0], [0, 0, 1, 1, 1], [0, 0, 1, 1, 0], [0, 0, 1, 1, 0]])
MASKED_EXPECTED = np.array([[1.0000, 0.666667, 0.600000, 0.500000, 0.50],
                            [1.0000, 0.750000, 0.571429, 0.428571, 0.25],
                            [1.0000, 1.000000, 0.714286, 0.571429, 0.25],
                            [np.nan, 1.000000, 0.666667, 0.571429, 0.25],
                            [np.nan, 1.000000, 0.750000, 0.750000, 0.50]])


def _plugin(method, nb_size=3, kernel=None, **kw):
    p = NeighbourhoodProcessing(method, RADIUS, **kw)
    p.nb_size = nb_size
    if kernel is not None:
        p.kernel = kernel
    return p


def test_kat_basic_square_and_sum():
    """test_NeighbourhoodProcessing.py:95-109, 165-181."""
    exp = np.ones((5, 5))
    exp[1:4, 1:4] = 8 / 9
    res = _plugin("square")._calculate_neighbourhood(_kat_data())
    np.testing.assert_array_almost_equal(res, exp)
    exp_sum = np.array([[4, 6, 6, 6, 4], [6, 8, 8, 8, 6], [6, 8, 8, 8, 6], [6, 8, 8, 8, 6], [4, 6, 6, 6, 4.0]])
    res = _plugin("square", sum_only=True)._calculate_neighbourhood(_kat_data())
    np.testing.assert_array_almost_equal(res, exp_sum)


def test_kat_circular_basic_edge_weighted_sum():
    """test_NeighbourhoodProcessing.py:111-163, 183-198."""
    k1 = circular_kernel(1, False)
    exp = np.ones((5, 5))
    exp[1, 2] = exp[2, 1] = exp[2, 2] = exp[2, 3] = exp[3, 2] = 0.8
    res = _plugin("circular", kernel=k1)._calculate_neighbourhood(_kat_data())
    np.testing.assert_array_almost_equal(res.data, exp)
    # zero in the left column: mode="nearest" repeats it (3/5 instead of 4/5)
    data = np.ones((5, 5), np.float32)
    data[:, :3] = _kat_data()[:, 2:]
    exp_edge = np.ones((5, 5))
    exp_edge[1, 0] = exp_edge[3, 0] = 0.8
    exp_edge[2, 0] = 0.6
    exp_edge[2, 1] = 0.8
    res = _plugin("circular", kernel=k1)._calculate_neighbourhood(data)
    np.testing.assert_array_almost_equal(res.data, exp_edge)
    exp_w = np.array([[1, 1, 1, 1, 1], [1, 0.916667, 0.875, 0.916667, 1], [1, 0.875, 0.833333, 0.875, 1],
                      [1, 0.916667, 0.875, 0.916667, 1], [1, 1, 1, 1, 1.0]])
    res = _plugin("circular", kernel=circular_kernel(2, True))._calculate_neighbourhood(_kat_data())
    np.testing.assert_array_almost_equal(res.data, exp_w)
    exp_s = np.full((5, 5), 5.0)
    exp_s[1, 2] = exp_s[2, 1] = exp_s[2, 2] = exp_s[2, 3] = exp_s[3, 2] = 4.0
    res = _plugin("circular", kernel=k1, sum_only=True)._calculate_neighbourhood(_kat_data())
    np.testing.assert_array_almost_equal(res.data, exp_s)


def test_kat_annulus_trimming():
    """test_NeighbourhoodProcessing.py:200-230 (array-shrinking optimisation)."""
    data = np.ones((10, 10), np.float32)
    data[4:6, 4:6] = 0
    exp = np.ones_like(data)
    exp[4:6, 4:6] = 5 / 9
    exp[3:7:3, 4:6] = 7 / 9
    exp[4:6, 3:7:3] = 7 / 9
    exp[3:7:3, 3:7:3] = 8 / 9
    np.testing.assert_array_almost_equal(_plugin("square")._calculate_neighbourhood(data), exp)
    exp = np.ones_like(data)
    exp[4:6, 4:6] = 0.4
    exp[3:7:3, 4:6] = 0.8
    exp[4:6, 3:7:3] = 0.8
    res = _plugin("circular", kernel=circular_kernel(1, False))._calculate_neighbourhood(data)
    np.testing.assert_array_almost_equal(res, exp)


def test_kat_masked_arrays():
    """test_NeighbourhoodProcessing.py:232-300."""
    inp = np.ma.masked_where(MASK == 0, MASKED_DATA)
    res = _plugin("square")._calculate_neighbourhood(inp)
    np.testing.assert_array_almost_equal(res.data, MASKED_EXPECTED)
    np.testing.assert_array_equal(res.mask, ~MASK.astype(bool))
    exp_c = np.array([[np.nan, 0.5, 0.5, 0.5, 1.0], [1.0, 1.0, 0.6, 0.5, 0.0], [np.nan, 1.0, 0.75, 0.4, 0.0],
                      [np.nan, 1.0, 1.0, 0.5, 0.5], [np.nan, 1.0, 0.75, 0.5, 0.0]])
    res = _plugin("circular", kernel=circular_kernel(1, False))._calculate_neighbourhood(inp)
    np.testing.assert_array_almost_equal(res.data, exp_c)
    np.testing.assert_array_equal(res.mask, ~MASK.astype(bool))
    res = _plugin("square", re_mask=False)._calculate_neighbourhood(inp)
    np.testing.assert_array_almost_equal(res, MASKED_EXPECTED)
    assert not isinstance(res, np.ma.MaskedArray)
    # NaN hidden under the mask must not matter
    with_nan = MASKED_DATA.copy()
    with_nan[0, 0] = np.nan
    res = _plugin("square")._calculate_neighbourhood(np.ma.masked_where(MASK == 0, with_nan))
    np.testing.assert_array_almost_equal(res.data[1:], MASKED_EXPECTED[1:])


def test_kat_external_mask():
    """test_NeighbourhoodProcessing.py:335-370 (external mask, re_mask both ways)."""
    res = _plugin("square")._calculate_neighbourhood(MASKED_DATA.astype(np.float32), MASK)
    np.testing.assert_array_almost_equal(res.data, MASKED_EXPECTED)
    np.testing.assert_array_equal(res.mask, ~MASK.astype(bool))
    res = _plugin("square", re_mask=False)._calculate_neighbourhood(MASKED_DATA.astype(np.float32), MASK)
    np.testing.assert_array_almost_equal(res, MASKED_EXPECTED)


@pytest.mark.parametrize("method", ["square", "circular"])
def test_process_probability_cube_vs_oracle(method):
    """Config 1 of BASELINE.json: 200x200 2 km grid, radius 10 km (5 cells)."""
    rng = np.random.default_rng(0)
    data = rng.random((1, 200, 200)).astype(np.float32)
    cube = set_up_probability_cube(data, [278.0], attributes={"title": "UKV Model Forecast"})
    plugin = NeighbourhoodProcessing(method, 10000)
    result = plugin(cube)
    exp, _ = orc.neighbourhood_cube(data, None, method=method, cells=5)
    np.testing.assert_allclose(np.asarray(result.data), exp, atol=1e-5)
    assert result.shape == cube.shape and result.name() == cube.name()
    assert result.attributes["title"] == "Post-Processed UKV Model Forecast"
    assert plugin.nb_size == 11


def test_process_re_mask_without_mask_source_is_masked_array():
    """nbhood.py:314-317: re_mask wraps the result in a MaskedArray even when nothing is masked."""
    rng = np.random.default_rng(5)
    data = rng.random((2, 30, 36)).astype(np.float32)
    cube = set_up_probability_cube(data, [1.0, 2.0])
    for re_mask in (True, False):
        res = NeighbourhoodProcessing("square", 6000, re_mask=re_mask)(cube).data
        assert isinstance(res, np.ma.MaskedArray) == re_mask
        if re_mask:
            assert not np.ma.getmaskarray(res).any()
        exp, _ = orc.neighbourhood_cube(data, None, method="square", cells=3)
        np.testing.assert_allclose(np.ma.getdata(res), exp, atol=1e-5)


def test_process_lead_time_radii_and_mask_cube():
    """nbhood.py:132-175: radius interpolated at the cube's forecast period;
    land-sea style mask cube with re_mask (config 2 semantics)."""
    rng = np.random.default_rng(1)
    data = rng.random((3, 40, 50)).astype(np.float32)
    mask = (rng.random((40, 50)) < 0.6).astype(np.int8)
    cube = set_up_probability_cube(data, [1.0, 2.0, 3.0], forecast_period_hours=4.0)
    mask_cube = set_up_variable_cube(mask.astype(np.float32), name="land_binary_mask", units="1")
    plugin = NeighbourhoodProcessing("circular", [5600, 9500], lead_times=[3, 5])
    result = plugin(cube, mask_cube)
    assert float(np.asarray(plugin.radius).item()) == pytest.approx(7550.0)     # test_BaseNeighbourhoodProcessing.py:75-84
    exp, emask = orc.neighbourhood_cube(data, mask, method="circular", cells=3)
    assert isinstance(result.data, np.ma.MaskedArray)
    np.testing.assert_array_equal(np.ma.getmaskarray(result.data), emask)
    got = np.ma.getdata(result.data)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    np.testing.assert_allclose(got[~np.isnan(exp)], exp[~np.isnan(exp)], atol=1e-5)
    # land + sea recombination (improver/cli/nbhood_land_and_sea.py:148-184)
    land = NeighbourhoodProcessing("circular", 7550)(cube, mask_cube)
    sea_mask_cube = set_up_variable_cube((mask == 0).astype(np.float32), name="sea", units="1")
    sea = NeighbourhoodProcessing("circular", 7550)(cube, sea_mask_cube)
    combined = land.data.filled(0) + sea.data.filled(0)
    np.testing.assert_allclose(combined, orc.land_sea_neighbourhood(data, mask, method="circular", cells=3),
                               atol=1e-5)


def test_process_errors():
    """nbhood.py:167-168, spatial.py:66,148-159, nbhood.py:49-52."""
    data = np.random.default_rng(2).random((16, 16)).astype(np.float32)
    cube = set_up_variable_cube(data)
    with pytest.raises(ValueError, match="exceeds max domain distance"):
        NeighbourhoodProcessing("square", 50000)(cube)
    with pytest.raises(ValueError, match="gives zero cell extent"):
        NeighbourhoodProcessing("square", 1000)(cube)
    bad = data.copy()
    bad[3, 3] = np.nan
    with pytest.raises(ValueError, match="NaN detected in input cube data"):
        NeighbourhoodProcessing("square", 4000)(set_up_variable_cube(bad))
    with pytest.raises(ValueError):
        NeighbourhoodProcessing("square", 4000)(set_up_variable_cube(data, spatial_grid="latlon"))
    masked = np.ma.masked_array(data, data > 0.9)
    with pytest.raises(NotImplementedError, match="masked input cubes is not yet implemented"):
        GeneratePercentilesFromANeighbourhood(4000)(set_up_variable_cube(masked))


def test_percentiles_process_multi_realization():
    """test_GeneratePercentilesFromANeighbourhood.py:383-448: realizations are a
    batch axis; output dims (realization, percentile, y, x)."""
    rng = np.random.default_rng(4)
    data = rng.random((3, 21, 24)).astype(np.float32)
    cube = set_up_variable_cube(data)
    pcts = (5, 25, 50, 75, 95)
    result = GeneratePercentilesFromANeighbourhood(4000, percentiles=pcts)(cube)
    assert result.shape == (3, 5, 21, 24)
    assert [c.name() for c in result.dim_coords][:2] == ["realization", "percentile"]
    exp = orc.neighbourhood_percentiles_cube(data, 2, pcts)
    np.testing.assert_allclose(result.data, exp, atol=1e-5)
    meta = MetaNeighbourhood("percentiles", 4000, neighbourhood_shape="circular", percentiles=pcts)(cube)
    np.testing.assert_allclose(meta.data, exp, atol=1e-5)


def test_percentiles_docstring_example():
    """nbhood.py:515-655 worked example."""
    data = np.ones((3, 3), np.float32)
    data[1, 1] = 0
    plugin = GeneratePercentilesFromANeighbourhood(4000, percentiles=[10])
    out = plugin.pad_and_unpad_cube(data, circular_kernel(2, False))
    np.testing.assert_allclose(out[0], np.full((3, 3), 0.5), atol=1e-6)


@pytest.mark.parametrize("kwargs,expected", [
    (dict(threshold_values=0.5, fuzzy_factor=0.5), 0.5),
    (dict(threshold_values=0.4, fuzzy_factor=0.5), 0.75),
    (dict(threshold_values=0.8, fuzzy_factor=0.5), 0.125),
    (dict(threshold_config={"0.6": [0.4, 0.7]}), 0.25),
    (dict(threshold_values=0.1), 1.0),
    (dict(threshold_values=0.6), 0.0),
    (dict(threshold_values=0.5, comparison_operator=">="), 1.0),
    (dict(threshold_values=0.5, comparison_operator=">"), 0.0),
    (dict(threshold_values=0.5, comparison_operator="=="), 1.0),
    (dict(threshold_values=0.4, fuzzy_factor=0.5, comparison_operator="<"), 0.25),
])
def test_threshold_expected_values(kwargs, expected):
    """improver_tests/threshold/test_Threshold.py:284-361: a single 0.5 at the
    centre of a zero field."""
    data = np.zeros((5, 5), np.float32)
    data[2, 2] = 0.5
    cube = set_up_variable_cube(data, name="precipitation_amount", units="m")
    result = Threshold(**kwargs)(cube)
    assert result.data.dtype == np.float32
    assert result.data[2, 2] == pytest.approx(expected, abs=1e-6)
    relative = "below" if "<" in kwargs.get("comparison_operator", ">") else "above"
    assert result.name() == f"probability_of_precipitation_amount_{relative}_threshold"
    assert result.units == "1"
    thr = result.coord(var_name="threshold")
    assert thr.points.dtype == np.float32


def test_threshold_collapse_realization_and_masks_vs_oracle():
    """threshold.py:643-713: realization collapse with masked members."""
    rng = np.random.default_rng(5)
    data = rng.gamma(2.0, 1.5, (4, 12, 13)).astype(np.float32)
    m = rng.random(data.shape) < 0.25
    m[:, 3, 3] = True
    cube = set_up_variable_cube(np.ma.masked_array(data, m), name="rainfall_rate", units="mm h-1")
    plugin = Threshold(threshold_values=[1.0, 3.0], fuzzy_factor=0.7, collapse_coord="realization")
    result = plugin(cube)
    exp, inv, _ = orc.threshold_cube(np.ma.masked_array(data, m), [1.0, 3.0], plugin.fuzzy_bounds, ">",
                                     collapse_leading=True)
    assert result.shape == (2, 12, 13)
    np.testing.assert_array_equal(np.ma.getdata(result.data), exp)
    np.testing.assert_array_equal(np.ma.getmaskarray(result.data), np.broadcast_to(inv, exp.shape))
    assert not result.coords("realization")
    with pytest.raises(ValueError, match="NaN detected"):
        bad = data.copy()
        bad[0, 0, 0] = np.nan
        Threshold(threshold_values=1.0)(set_up_variable_cube(bad))
    with pytest.raises(ValueError, match="Cannot collapse over coordinates not present"):
        Threshold(threshold_values=1.0, collapse_coord="realization")(set_up_variable_cube(data[0]))


def test_boxsum_matches_oracle():
    """improver/utilities/neighbourhood_tools.py:129-211 through the CUDA kernel."""
    from improver_b200.utilities.neighbourhood_tools import boxsum

    rng = np.random.default_rng(6)
    data = rng.random((2, 23, 31))
    for box in (3, 7):
        np.testing.assert_allclose(boxsum(data, box), orc.boxsum(data, box), rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(boxsum(data, box, mode="constant", constant_values=0),
                                   orc.boxsum(data, box, mode="constant", constant_values=0), rtol=1e-6,
                                   atol=1e-5)
        np.testing.assert_allclose(boxsum(data, box, mode="edge"), orc.boxsum(data, box, mode="edge"),
                                   rtol=1e-6, atol=1e-5)
    ints = (rng.random((9, 9)) < 0.5).astype(np.int64)
    np.testing.assert_array_equal(boxsum(ints, 3, mode="constant"), orc.boxsum(ints, 3, mode="constant"))
    cum = data.cumsum(-2).cumsum(-1)
    np.testing.assert_allclose(boxsum(cum, 3, cumsum=False), orc.boxsum(cum, 3, cumsum=False))


def test_apply_neighbourhood_with_zone_masks_vs_oracle():
    """improver/nbhood/use_nbhood.py:194-262."""
    from improver_b200.cube import Coord, Cube
    from improver_b200.nbhood.use_nbhood import ApplyNeighbourhoodProcessingWithAMask

    rng = np.random.default_rng(8)
    data = rng.random((2, 30, 36)).astype(np.float32)
    zones = rng.integers(0, 3, (30, 36))
    zones[:8, :8] = 3                                         # "sea": in no zone at all
    masks = np.stack([(zones == z).astype(np.int8) for z in range(3)])
    weights = rng.random((3, 30, 36)).astype(np.float32)
    weights /= weights.sum(axis=0, keepdims=True)
    cube = set_up_probability_cube(data, [1.0, 2.0])
    y, x = cube.coord(axis="y"), cube.coord(axis="x")
    zcoord = Coord(np.arange(3, dtype=np.float32), "topographic_zone", "m")
    mask_cube = Cube(masks, "topography_mask", "1", [(zcoord, 0), (y.copy(), 1), (x.copy(), 2)])
    plugin = ApplyNeighbourhoodProcessingWithAMask("topographic_zone", "square", 4000)
    result = plugin(cube, mask_cube)
    exp = orc.neighbourhood_with_zone_masks(data, masks, method="square", cells=2)
    assert result.shape == (2, 3, 30, 36)
    got = np.asarray(result.data)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    np.testing.assert_allclose(got[~np.isnan(exp)], exp[~np.isnan(exp)], atol=1e-5)
    wcube = Cube(weights, "topographic_zone_weights", "1", [(zcoord.copy(), 0), (y.copy(), 1), (x.copy(), 2)])
    collapsed = ApplyNeighbourhoodProcessingWithAMask("topographic_zone", "square", 4000,
                                                      collapse_weights=wcube)(cube, mask_cube)
    exp_c = orc.neighbourhood_with_zone_masks(data, masks, weights=weights, method="square", cells=2)
    assert collapsed.shape == (2, 30, 36)
    gotc = np.asarray(collapsed.data)
    np.testing.assert_array_equal(np.isnan(gotc), np.isnan(exp_c))
    np.testing.assert_allclose(gotc[~np.isnan(exp_c)], exp_c[~np.isnan(exp_c)], atol=1e-5)


def test_nbhood_land_and_sea_vs_oracle():
    """improver/cli/nbhood_land_and_sea.py:131-184 (config 2 semantics)."""
    from improver_b200.nbhood.land_and_sea import nbhood_land_and_sea

    rng = np.random.default_rng(12)
    data = rng.random((2, 44, 51)).astype(np.float32)
    land = (rng.random((44, 51)) < 0.55).astype(np.int8)
    cube = set_up_probability_cube(data, [1.0, 2.0])
    mask_cube = set_up_variable_cube(land.astype(np.float32), name="land_binary_mask", units="1")
    for shape, cells in (("square", 3), ("circular", 3)):
        result = nbhood_land_and_sea(cube, mask_cube, neighbourhood_shape=shape, radii=6000)
        exp = orc.land_sea_neighbourhood(data, land, method=shape, cells=cells)
        np.testing.assert_allclose(np.asarray(result.data), exp, atol=1e-5)


def test_kat_complex_square():
    """improver_tests/nbhood/nbhood/test_NeighbourhoodProcessing.py:289-331 (test_complex):
    complex input is processed component-wise, output complex64."""
    data = np.ones((5, 5), dtype=complex)
    data[2, 2] = 0
    data[1, 3] = 0.5 + 0.5j
    data[4, 3] = 0.4 + 0.6j
    expected = np.array([
        [1.0 + 0.0j, 1.0 + 0.0j, 0.91666667 + 0.083333333j, 0.91666667 + 0.083333333j, 0.875 + 0.125j],
        [1.0 + 0.0j, 0.88888889 + 0.0j, 0.83333333 + 0.055555556j, 0.83333333 + 0.055555556j, 0.91666667 + 0.083333333j],
        [1.0 + 0.0j, 0.88888889 + 0.0j, 0.83333333 + 0.055555556j, 0.83333333 + 0.055555556j, 0.91666667 + 0.083333333j],
        [1.0 + 0.0j, 0.88888889 + 0.0j, 0.82222222 + 0.066666667j, 0.82222222 + 0.066666667j, 0.9 + 0.1j],
        [1.0 + 0.0j, 1.0 + 0.0j, 0.9 + 0.1j, 0.9 + 0.1j, 0.85 + 0.15j]])
    plugin = NeighbourhoodProcessing("square", 2500)
    plugin.nb_size = 3
    result = plugin._calculate_neighbourhood(data)
    assert result.dtype == np.complex64
    np.testing.assert_array_almost_equal(np.ma.getdata(result), expected)


def test_meta_neighbourhood_degrees_as_complex_and_halo():
    """MetaNeighbourhood(degrees_as_complex=True) (nbhood.py:866-889; complex_conversion.py) against
    the numpy restatement: unit vectors, box mean of the components, angle back in [0, 360); and
    halo_radius removal (nbhood.py:890-893, pad_spatial.py:222-300)."""
    from improver_b200.nbhood.nbhood import MetaNeighbourhood

    rng = np.random.default_rng(17)
    deg = (rng.random((3, 24, 30)) * 360).astype(np.float32)
    cube = set_up_variable_cube(deg, name="wind_from_direction", units="degrees")
    result = MetaNeighbourhood("probabilities", 4000, neighbourhood_shape="square", degrees_as_complex=True)(cube)
    rad = np.deg2rad(deg)
    z = np.cos(rad) + 1j * np.sin(rad)
    exp = np.empty(deg.shape, np.float32)
    for k in range(3):
        re = np.ma.getdata(orc.neighbourhood(z[k].real.astype(np.float64), None, cells=2, sum_only=True, re_mask=False))
        im = np.ma.getdata(orc.neighbourhood(z[k].imag.astype(np.float64), None, cells=2, sum_only=True, re_mask=False))
        exp[k] = np.mod(np.float32(np.angle(re + 1j * im, deg=True)), 360)
    got = np.ma.getdata(result.data)
    diff = np.abs(((got - exp) + 180.0) % 360.0 - 180.0)        # angular distance
    assert float(diff.max()) < 2e-3
    assert got.dtype == np.float32 and got.min() >= 0.0 and got.max() < 360.0

    prob = set_up_probability_cube(rng.random((2, 24, 30)).astype(np.float32), [1.0, 2.0])
    full = MetaNeighbourhood("probabilities", 4000)(prob)
    trimmed = MetaNeighbourhood("probabilities", 4000, halo_radius=6000)(prob)
    assert trimmed.shape == (2, 18, 24)
    np.testing.assert_array_equal(np.ma.getdata(trimmed.data), np.ma.getdata(full.data)[:, 3:-3, 3:-3])
    np.testing.assert_array_equal(trimmed.coord(axis="x").points, full.coord(axis="x").points[3:-3])


def test_threshold_units_converts_the_data_on_the_device():
    """threshold.py:430-431, 716-718: thresholds given in other units than the cube's: the data
    are converted (float32(a * double(x) + b) like cf_units), the threshold coordinate comes back
    in the cube's units."""
    from improver_b200.synthetic_data import set_up_variable_cube
    from improver_b200.threshold import Threshold

    rng = np.random.default_rng(17)
    kelvin = (273.15 + 10.0 * rng.random((3, 12, 14))).astype(np.float32)
    cube = set_up_variable_cube(kelvin, name="air_temperature", units="K")
    res = Threshold(threshold_values=[2.0, 5.0], threshold_units="degC", comparison_operator=">")(cube)
    celsius = (kelvin.astype(np.float64) - 273.15).astype(np.float32)
    exp = np.stack([(celsius > np.float32(t)).astype(np.float32) for t in (2.0, 5.0)], axis=1)
    np.testing.assert_array_equal(np.asarray(res.data), exp)
    np.testing.assert_allclose(res.coord(var_name="threshold").points, [275.15, 278.15], rtol=1e-6)
    assert str(res.coord(var_name="threshold").units) == "K"
    fuzzy = Threshold(threshold_values=[5.0], fuzzy_factor=0.8, threshold_units="degC")(cube)
    from oracle import nbhood_oracle as orc
    want = orc.truth_value(celsius, 5.0, (4.0, 6.0), ">")
    np.testing.assert_array_equal(np.asarray(fuzzy.data), want)
