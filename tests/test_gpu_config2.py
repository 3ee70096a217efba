"""BASELINE.json config 2 at full size: circular neighbourhood, lead-time dependent radii
(9 / 27 / 45 / 81 cells), land-sea mask weighted, re_mask, on the UK 2 km grid (970 x 1042).

scipy's correlate costs O((2r+1)^2) per point, so the oracle runs on CROPPED sub-regions:
a crop extended by r cells on every side that is not a domain edge gives, in its inner part,
exactly the values of the full field (the window of an inner point lies inside the crop; where
the crop touches the domain edge the reference's edge replication is the same in both).  One
interior crop and crops at two opposite corners (all four domain edges) are checked per radius.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from improver_b200 import workloads as W  # noqa: E402
from oracle import nbhood_oracle as orc  # noqa: E402

NY, NX = W.UK
ATOL = 1e-5
SIDE = 20          # inner crop side


def _crops(r):
    """(y0, y1, x0, x1) of the inner regions: interior, top-left corner, bottom-right corner,
    and a coastline-rich strip on the left edge."""
    return [(500, 500 + SIDE, 480, 480 + SIDE),
            (0, SIDE, 0, SIDE),
            (NY - SIDE, NY, NX - SIDE, NX),
            (300, 300 + SIDE, 0, SIDE)]


def _extend(y0, y1, x0, x1, r):
    return max(0, y0 - r), min(NY, y1 + r), max(0, x0 - r), min(NX, x1 + r)


def _check_crops(got, data, mask, r, oracle_fn, atol=ATOL):
    """got: numpy (n, NY, NX) full-size engine result; oracle_fn(sub_data, sub_mask) -> values."""
    for (y0, y1, x0, x1) in _crops(r):
        ya, yb, xa, xb = _extend(y0, y1, x0, x1, r)
        sub = np.ascontiguousarray(data[:, ya:yb, xa:xb])
        sub_mask = np.ascontiguousarray(mask[ya:yb, xa:xb]) if mask is not None else None
        exp = oracle_fn(sub, sub_mask)[:, y0 - ya:y1 - ya, x0 - xa:x1 - xa]
        g = got[:, y0:y1, x0:x1]
        np.testing.assert_array_equal(np.isnan(g), np.isnan(exp))
        ok = ~np.isnan(exp)
        np.testing.assert_allclose(g[ok], exp[ok], atol=atol, rtol=0.0,
                                   err_msg=f"crop {(y0, y1, x0, x1)} r={r}")


@pytest.mark.parametrize("cells", [9, 27, 45, 81])
def test_config2_land_sea_circular_fullsize(cells):
    """Land pass + sea pass + filled(0) sum (cli/nbhood_land_and_sea.py:148-184) through the
    engine, against the oracle on crops, plain and structured probabilities."""
    from improver_b200 import engine

    land = W.land_sea_mask()
    n = 2
    for structured in (False, True):
        data = W.config2_probabilities(n, structured=structured)
        dev = torch.from_numpy(data).cuda()
        land_d = torch.from_numpy((land > 0).astype(np.uint8)).cuda()
        sea_d = torch.from_numpy((land == 0).astype(np.uint8)).cuda()
        lv, lm, st = engine.neighbourhood(dev, cells, method="circular", mask2d=land_d, want_mask=True)
        engine.check_status(st)
        sv, sm, st = engine.neighbourhood(dev, cells, method="circular", mask2d=sea_d, want_mask=True)
        engine.check_status(st)
        got = (torch.where(lm != 0, torch.zeros_like(lv), lv) + torch.where(sm != 0, torch.zeros_like(sv), sv))
        got = got.cpu().numpy()
        # re_mask masks are exactly the complements of the two masks
        assert bool((lm[0] == sea_d).all()) and bool((sm[0] == land_d).all())
        _check_crops(got, data, land, cells,
                     lambda d, m: orc.land_sea_neighbourhood(d, m, method="circular", cells=cells))
        # one-pass land+sea entry point: same field, one launch
        both, st = engine.neighbourhood_land_sea(dev, cells, land_d, method="circular")
        engine.check_status(st)
        np.testing.assert_allclose(both.cpu().numpy(), got, atol=2e-6, rtol=0.0)


@pytest.mark.parametrize("cells", [27, 81])
def test_config2_unmasked_and_weighted_fullsize(cells):
    """No mask (edge replication only) at full size; weighted_mode at the radius it supports."""
    from improver_b200 import engine

    data = W.config2_probabilities(1, structured=True)
    dev = torch.from_numpy(data).cuda()
    out, _, st = engine.neighbourhood(dev, cells, method="circular")
    engine.check_status(st)
    _check_crops(out.cpu().numpy(), data, None, cells,
                 lambda d, m: orc.neighbourhood_cube(d, None, method="circular", cells=cells)[0])
    if cells <= 40:
        out, _, st = engine.neighbourhood(dev, cells, method="circular", weighted_mode=True)
        engine.check_status(st)
        _check_crops(out.cpu().numpy(), data, None, cells,
                     lambda d, m: orc.neighbourhood_cube(d, None, method="circular", cells=cells,
                                                         weighted_mode=True)[0])


def test_config2_plugin_lead_time_radii_fullsize():
    """The plugin path: radii interpolated from the cube's forecast period (nbhood.py:132-149),
    land+sea combination, one lead time per call."""
    from improver_b200.nbhood.land_and_sea import nbhood_land_and_sea
    from improver_b200.synthetic_data import set_up_probability_cube, set_up_variable_cube

    land = W.land_sea_mask()
    data = W.config2_probabilities(2)
    for fp_hours in (6, 36):
        cells = W.cells_for_lead_time(fp_hours)
        cube = set_up_probability_cube(data, [0.5, 1.0], forecast_period_hours=fp_hours)
        mask_cube = set_up_variable_cube(land, name="land_binary_mask", units="1", forecast_period_hours=None)
        res = nbhood_land_and_sea(cube, mask_cube, "circular", W.CONFIG2_RADII_M, W.CONFIG2_LEAD_HOURS)
        got = np.asarray(res.data)
        assert got.dtype == np.float32 and got.shape == data.shape
        _check_crops(got, data, land, cells,
                     lambda d, m: orc.land_sea_neighbourhood(d, m, method="circular", cells=cells))


@pytest.mark.parametrize("cells", [5, 27])
def test_square_land_sea_fullsize_vs_oracle_crops(cells):
    """Square method with the land-sea masks at full size (mask-weighted normalisation)."""
    from improver_b200 import engine

    land = W.land_sea_mask()
    data = W.config2_probabilities(2, structured=True)
    dev = torch.from_numpy(data).cuda()
    land_d = torch.from_numpy((land > 0).astype(np.uint8)).cuda()
    both, st = engine.neighbourhood_land_sea(dev, cells, land_d, method="square")
    engine.check_status(st)
    got = both.cpu().numpy()
    # square trimming quirk only concerns all-ones margins next to the domain edge: the crop
    # oracle is valid wherever the crop's own trimming decision equals the full field's, which
    # the structured field guarantees for the interior crop; edge crops use the plain field
    plain = W.config2_probabilities(2)
    devp = torch.from_numpy(plain).cuda()
    bothp, st = engine.neighbourhood_land_sea(devp, cells, land_d, method="square")
    engine.check_status(st)
    _check_crops(bothp.cpu().numpy(), plain, land, cells,
                 lambda d, m: orc.land_sea_neighbourhood(d, m, method="square", cells=cells))
    y0, y1, x0, x1 = _crops(cells)[0]
    ya, yb, xa, xb = _extend(y0, y1, x0, x1, cells)
    sub = np.ascontiguousarray(data[:, ya:yb, xa:xb])
    if not ((sub == 1).all() or (sub == 0).all()):
        exp = orc.land_sea_neighbourhood(sub, land[ya:yb, xa:xb], method="square", cells=cells)
        np.testing.assert_allclose(got[:, y0:y1, x0:x1], exp[:, y0 - ya:y1 - ya, x0 - xa:x1 - xa], atol=ATOL)


@pytest.mark.parametrize("cells,masked", [(5, False), (9, True)])
def test_circular_many_slices_equal_single_slice_launches(cells, masked):
    """More slices than one prefix batch holds (the float64 / fixed-point row prefixes of a batch of
    planes live in a bounded workspace): every batch must be computed, raw and land-sea masked."""
    from improver_b200 import engine

    n = 75
    dev = W.counter_uniform(0, n, W.UK, 9, "cuda")
    land = torch.from_numpy((W.land_sea_mask() > 0).astype(np.uint8)).cuda() if masked else None
    if masked:
        out, st = engine.neighbourhood_land_sea(dev, cells, land, method="circular")
    else:
        out, _, st = engine.neighbourhood(dev, cells, method="circular")
    engine.check_status(st)
    for k in (0, 32, 33, 40, 65, 66, 74):
        if masked:
            one, st = engine.neighbourhood_land_sea(dev[k:k + 1].contiguous(), cells, land, method="circular")
        else:
            one, _, st = engine.neighbourhood(dev[k:k + 1].contiguous(), cells, method="circular")
        engine.check_status(st)
        assert torch.equal(out[k].nan_to_num(-1.0), one[0].nan_to_num(-1.0)), k
