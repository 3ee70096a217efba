"""Deferred (device-resident) cubes: Threshold -> NeighbourhoodProcessing -> collapse_realizations
on `lazy.defer`-ed cubes gives the same result as the eager plugins and the oracle, fused (host and
device sources) or step by step; metadata is the eager plugins' metadata."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import nbhood_oracle as orc  # noqa: E402


def _cube(data, with_time=True):
    from improver_b200.cube import Coord, Cube

    R = data.shape[0]
    ny, nx = data.shape[-2:]
    y = Coord(np.arange(ny, dtype=np.float32) * 2000.0, "projection_y_coordinate", "m", "y")
    x = Coord(np.arange(nx, dtype=np.float32) * 2000.0, "projection_x_coordinate", "m", "x")
    dims = [(Coord(np.arange(R, dtype=np.int32), "realization", "1"), 0)]
    if with_time:
        dims.append((Coord(np.arange(data.shape[1], dtype=np.int64) * 3600, "time", "seconds"), 1))
    n = len(dims)
    dims += [(y, n), (x, n + 1)]
    return Cube(data, "lwe_precipitation_rate", "mm h-1", dims, [Coord([3600], "forecast_period", "seconds")])


@pytest.mark.parametrize("source", ["host", "pinned", "device"])
@pytest.mark.parametrize("with_time", [True, False])
@pytest.mark.parametrize("thr_values", [[1.0, 2.0], [2.0]])
def test_fused_chain_through_the_plugins(source, with_time, thr_values):
    from improver_b200 import lazy
    from improver_b200.nbhood.nbhood import NeighbourhoodProcessing
    from improver_b200.threshold import Threshold
    from improver_b200.utilities.cube_manipulation import collapse_realizations

    rng = np.random.default_rng(3)
    shape = (5, 4, 60, 82) if with_time else (5, 60, 82)
    data = rng.gamma(2.0, 1.5, shape).astype(np.float32)
    if source == "host":
        held = data
    elif source == "pinned":
        held = torch.from_numpy(data).pin_memory()
    else:
        held = torch.from_numpy(data).cuda()
    cube = _cube(data, with_time).copy(data=lazy.Source(held))
    thr = Threshold(threshold_values=thr_values, fuzzy_factor=0.7, comparison_operator=">")
    prob = thr(cube)
    assert prob.has_lazy_data() and prob.name() == "probability_of_lwe_precipitation_rate_above_threshold"
    nbh = NeighbourhoodProcessing("square", 6000.0)(prob)
    mean = collapse_realizations(nbh)
    assert mean.has_lazy_data() and mean.core_data()._fusable() is not None
    assert not mean.coords("realization")
    got = np.asarray(mean.data)
    d4 = data if with_time else data[:, None]
    bounds = orc.fuzzy_bounds_from_factor(thr_values, 0.7)
    exp = orc.threshold_square_mean(d4, thr_values, bounds, ">", 3)
    exp = exp if with_time else exp[0]
    if len(thr_values) == 1:
        exp = exp.squeeze(-3)                                   # threshold.py:731: length-1 dims are squeezed
    assert got.shape == exp.shape == tuple(mean.shape)
    np.testing.assert_allclose(got, exp, atol=1e-5, rtol=0)
    # the eager plugins on the plain cube agree (and so does the metadata)
    eager = collapse_realizations(NeighbourhoodProcessing("square", 6000.0)(thr(_cube(data, with_time))))
    np.testing.assert_allclose(np.ma.getdata(eager.data), got, atol=1e-5, rtol=0)
    assert [c.name() for c in eager.dim_coords] == [c.name() for c in mean.dim_coords]


def test_unfused_deferred_steps_and_masks():
    """Chains the fused kernel does not take run step by step on the device: threshold alone,
    neighbourhood of raw data with a mask cube (re_mask), circular with a mean afterwards."""
    from improver_b200 import lazy
    from improver_b200.nbhood.nbhood import NeighbourhoodProcessing
    from improver_b200.threshold import Threshold
    from improver_b200.utilities.cube_manipulation import collapse_realizations

    rng = np.random.default_rng(4)
    data = rng.random((3, 40, 50)).astype(np.float32)
    mask = (rng.random((40, 50)) < 0.8).astype(np.int8)
    cube = lazy.defer(_cube(data, with_time=False))
    # threshold alone
    prob = Threshold(threshold_values=[0.5], comparison_operator=">=")(cube)
    exp = (data >= 0.5).astype(np.float32)
    np.testing.assert_array_equal(np.asarray(prob.data), exp)
    # masked neighbourhood of the raw field
    mask_cube = type("M", (), {"data": mask})()
    nb = NeighbourhoodProcessing("circular", 6000.0, re_mask=True)(cube, mask_cube)
    assert nb.has_lazy_data()
    got = nb.data
    assert isinstance(got, np.ma.MaskedArray)
    exp, expm = orc.neighbourhood_cube(data, mask, method="circular", cells=3)
    np.testing.assert_allclose(np.ma.getdata(got), exp, atol=1e-5)
    np.testing.assert_array_equal(np.ma.getmaskarray(got), expm)
    # mean over realizations of a (non thresholded) neighbourhood: collapse kernel
    m = collapse_realizations(NeighbourhoodProcessing("square", 4000.0, re_mask=False)(cube))
    exp2 = np.stack([np.ma.getdata(orc.neighbourhood(s, None, method="square", cells=2, re_mask=False))
                     for s in data]).mean(axis=0)
    np.testing.assert_allclose(np.asarray(m.data), exp2, atol=1e-5)


def test_collapse_realizations_eager_and_zone_collapse_kernels():
    from improver_b200 import engine
    from improver_b200.utilities.cube_manipulation import collapse_realizations

    rng = np.random.default_rng(6)
    data = rng.random((7, 33, 20)).astype(np.float32)
    res = collapse_realizations(_cube(data, with_time=False))
    np.testing.assert_array_equal(np.asarray(res.data), data.mean(axis=0))     # numpy's float32 axis-0 mean
    vals = rng.random((2, 3, 15, 11)).astype(np.float32)
    vals[0, 1, 3, 4] = np.nan
    vals[1, :, 5, 5] = np.nan
    w = rng.random((3, 15, 11)).astype(np.float32)
    got = engine.zone_collapse(torch.from_numpy(vals).cuda(), torch.from_numpy(w).cuda()).cpu().numpy()
    valid = ~np.isnan(vals)
    num = np.where(valid, vals, 0) * w
    den = np.where(valid, w, 0)
    with np.errstate(invalid="ignore"):
        exp = num.sum(axis=1) / den.sum(axis=1)
    assert np.isnan(got[1, 5, 5])
    np.testing.assert_allclose(got[~np.isnan(exp)], exp[~np.isnan(exp)], rtol=1e-6)
