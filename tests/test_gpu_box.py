"""The fixed-point box kernel (single-member square neighbourhood, radius <= 8 cells) and its
guarded follow-ups against the oracle: every radius, grid widths that exercise partial lanes,
inactive windows and x panels, external masks, thresholds, sum_only, the clip of the reference
(nbhood.py:312: results never leave [min, max] of the slice, constant fields are reproduced
bit for bit) and raw data that is not a probability field (float64 re-run)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import nbhood_oracle as orc  # noqa: E402

ATOL = 1e-5


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _oracle(data, mask, cells, sum_only=False):
    return np.stack([np.ma.getdata(orc.neighbourhood(sl, mask, method="square", cells=cells, sum_only=sum_only,
                                                     re_mask=False)) for sl in data])


def _close(got, exp, atol=ATOL):
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    ok = ~np.isnan(exp)
    np.testing.assert_allclose(got[ok], exp[ok], atol=atol, rtol=0.0)


@pytest.mark.parametrize("cells", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("shape", [(3, 37, 250), (2, 21, 16), (2, 40, 1042), (1, 19, 482), (2, 12, 2560)])
def test_box_kernel_vs_oracle(cells, shape):
    from improver_b200 import engine

    if cells * 2 + 1 > min(shape[-2:]):
        pytest.skip("window larger than the grid")
    rng = np.random.default_rng(100 * cells + shape[-1])
    data = rng.random(shape).astype(np.float32)
    data[0, : shape[1] // 2, : shape[2] // 3] = 1.0          # plateaus of exact ones and zeros
    data[-1, shape[1] // 2:, shape[2] // 2:] = 0.0
    mask = (rng.random(shape[-2:]) < 0.75).astype(np.uint8)
    for m in (None, mask):
        md = _dev(m) if m is not None else None
        for sum_only in (False, True):
            out, omask, st = engine.neighbourhood(_dev(data), cells, mask2d=md, sum_only=sum_only, want_mask=True)
            engine.check_status(st)
            exp = _oracle(data, m, cells, sum_only)
            scale = float((2 * cells + 1) ** 2) if sum_only else 1.0
            _close(out.cpu().numpy(), exp, ATOL * scale)
            if m is not None:
                np.testing.assert_array_equal(omask.cpu().numpy()[0].astype(bool), m == 0)
            if not sum_only:
                got = out.cpu().numpy()
                for i in range(shape[0]):
                    ok = ~np.isnan(got[i])
                    assert got[i][ok].min() >= data[i].min() and got[i][ok].max() <= data[i].max()


def test_box_kernel_constant_and_plateau_fields_are_exact():
    """nbhood.py:312 clip: a constant field comes back bit for bit, results stay inside the range
    of the slice (the quantisation of the fixed-point sums is undone by the guarded clamp)."""
    from improver_b200 import engine

    ny, nx = 64, 250
    data = np.empty((4, ny, nx), dtype=np.float32)
    data[0] = 0.3
    data[1] = np.float32(1.0) - np.float32(2.0 ** -24)
    data[2] = 0.0
    data[2, 20:40, 30:90] = 0.3                               # blob surrounded by zeros: exact zeros outside
    data[3] = 0.7
    data[3, 10:30, 100:200] = np.linspace(0.7, 0.9, 100, dtype=np.float32)
    out, _, st = engine.neighbourhood(_dev(data), 5)
    engine.check_status(st)
    got = out.cpu().numpy()
    np.testing.assert_array_equal(got[0], data[0])
    np.testing.assert_array_equal(got[1], data[1])
    assert got[2].min() == 0.0 and (got[2][:10] == 0.0).all() and got[2].max() <= np.float32(0.3)
    assert got[3].min() >= np.float32(0.7) and got[3].max() <= np.float32(0.9)
    _close(got, _oracle(data, None, 5))


def test_box_kernel_non_probability_slices_are_recomputed_in_float64():
    """Raw data outside [0, 1] (temperatures, negative values, +-inf) in some slices of a batch:
    those slices take the float64 kernels, the others keep the fixed-point result."""
    from improver_b200 import engine

    rng = np.random.default_rng(5)
    data = rng.random((5, 50, 250)).astype(np.float32)
    data[1] = 283.15 + 10.0 * data[1]
    data[3] = data[3] - 0.5
    data[4, 7, 9] = 1.0000001
    mask = (rng.random((50, 250)) < 0.8).astype(np.uint8)
    for m in (None, mask):
        out, _, st = engine.neighbourhood(_dev(data), 4, mask2d=_dev(m) if m is not None else None)
        engine.check_status(st)
        exp = _oracle(data, m, 4)
        got = out.cpu().numpy()
        _close(got[[0, 2]], exp[[0, 2]])
        _close(got[[3, 4]], exp[[3, 4]])
        np.testing.assert_allclose(got[1], exp[1], rtol=2e-7, atol=0)
    const = np.full((2, 40, 64), 283.15, dtype=np.float32)
    out, _, st = engine.neighbourhood(_dev(const), 3)
    engine.check_status(st)
    np.testing.assert_array_equal(out.cpu().numpy(), const)


def test_box_kernel_nan_raises_and_thresholds():
    from improver_b200 import engine

    rng = np.random.default_rng(8)
    data = rng.gamma(2.0, 1.5, (3, 45, 482)).astype(np.float32)
    thr = [(1.0, 0.7, 1.3, ">"), (3.0, 3.0, 3.0, ">="), (2.0, 1.4, 2.6, "<")]
    out, _, st = engine.neighbourhood(_dev(data), 5, thresholds=thr)
    engine.check_status(st)
    for k, (v, lo, hi, op) in enumerate(thr):
        tv = orc.truth_value(data, v, (lo, hi), op)
        _close(out[:, k].cpu().numpy(), _oracle(tv, None, 5))
    bad = data.copy()
    bad[1, 20, 300] = np.nan
    _, _, st = engine.neighbourhood(_dev(bad), 5)
    with pytest.raises(ValueError, match="NaN detected"):
        engine.check_status(st)
    _, _, st = engine.neighbourhood(_dev(bad), 5, thresholds=thr[:1])
    with pytest.raises(ValueError, match="NaN detected"):
        engine.check_status(st)


def test_box_kernel_band_origin_independence():
    """Any row range of a slice, run as its own launch with halo rows, reproduces the rows of the
    whole-slice launch bit for bit (exact fixed-point sums)."""
    from improver_b200 import engine

    rng = np.random.default_rng(12)
    data = rng.random((2, 120, 250)).astype(np.float32)
    full, _, st = engine.neighbourhood(_dev(data), 5)
    engine.check_status(st)
    for y0, y1 in ((0, 50), (37, 91), (70, 120)):
        a, b = max(0, y0 - 5), min(120, y1 + 5)
        sub = np.ascontiguousarray(data[:, a:b])
        stats = engine.slice_stats(_dev(data), 5)
        band = engine.Band(a, 120, y0 - a, b - y1)
        part, _, st = engine.neighbourhood(_dev(sub), 5, band=band, ext_stats=stats)
        engine.check_status(st)
        assert torch.equal(part[:, y0 - a:y0 - a + (y1 - y0)], full[:, y0:y1])


@pytest.mark.parametrize("shape,cells,n_thr", [((5, 2, 48, 70), 3, 5), ((18, 1, 60, 1042), 5, 4), ((3, 2, 40, 250), 8, 2),
                                                ((4, 1, 33, 1400), 1, 3), ((6, 3, 24, 16), 2, 6)])
def test_multi_threshold_fused_kernel_vs_oracle(shape, cells, n_thr):
    """square_mt_kernel: several thresholds from one read of the input (members x thresholds),
    with / without an external mask, slices that trigger the ones-trimming fix-up, wide rows
    (x panels) and a remainder group of thresholds."""
    from improver_b200 import engine

    rng = np.random.default_rng(33 + n_thr)
    data = rng.gamma(2.0, 1.5, shape).astype(np.float32)
    data[1, 0] = 50.0                                        # an all-ones member (edge flags)
    data[1, 0, :2, 4:9] = 0.1
    thr = [0.5, 1.0, 2.0, 4.0, 6.0, 8.0][:n_thr]
    bounds = orc.fuzzy_bounds_from_factor(thr, 0.7)
    mask = (rng.random(shape[-2:]) < 0.8).astype(np.uint8)
    for m in (None, mask):
        exp = orc.threshold_square_mean(data, thr, bounds, ">", cells, mask=m)
        out, _, status = engine.neighbourhood(_dev(data), cells, mask2d=_dev(m) if m is not None else None,
                                              thresholds=[(t, b[0], b[1], ">") for t, b in zip(thr, bounds)],
                                              collapse_members=True)
        engine.check_status(status)
        _close(out.cpu().numpy(), exp)
    # deterministic comparison and a "less than" fuzzy threshold in one call (two launch groups)
    mixed = [(1.0, 1.0, 1.0, ">="), (2.0, 2.0, 2.0, ">"), (2.0, 1.4, 2.6, "<"), (4.0, 3.0, 5.0, "<")]
    out, _, status = engine.neighbourhood(_dev(data), cells, thresholds=mixed, collapse_members=True)
    engine.check_status(status)
    for k, (v, lo, hi, op) in enumerate(mixed):
        exp = orc.threshold_square_mean(data, [v], [(lo, hi)], op, cells)
        _close(out[:, k:k + 1].cpu().numpy(), exp)


@pytest.mark.parametrize("cells", [9, 16, 27, 40])
@pytest.mark.parametrize("shape", [(3, 130, 250), (2, 300, 1042), (2, 97, 64)])
def test_two_pass_kernels_vs_oracle(cells, shape):
    """square_v_kernel + square_h_kernel (single-member slices, radius 9 .. 63): plain, masked,
    sum_only, thresholded, a slice that is not a probability field, ones-trimming slices, several
    y segments (300 rows) and the range property of the clip."""
    from improver_b200 import engine

    if 2 * cells + 1 > min(shape[-2:]):
        pytest.skip("window larger than the grid")
    rng = np.random.default_rng(7 * cells + shape[-1])
    data = rng.random(shape).astype(np.float32)
    data[0] = 1.0
    data[0, :3, 5:40] = 0.25                                  # all ones apart from an edge blob (trimming fix-up)
    data[-1, shape[1] // 2:, shape[2] // 2:] = 0.0
    mask = (rng.random(shape[-2:]) < 0.75).astype(np.uint8)
    for m in (None, mask):
        md = _dev(m) if m is not None else None
        for sum_only in (False, True):
            out, omask, st = engine.neighbourhood(_dev(data), cells, mask2d=md, sum_only=sum_only, want_mask=True)
            engine.check_status(st)
            exp = _oracle(data, m, cells, sum_only)
            scale = float((2 * cells + 1) ** 2) if sum_only else 1.0
            _close(out.cpu().numpy(), exp, ATOL * scale)
            if not sum_only:
                got = out.cpu().numpy()
                for i in range(shape[0]):
                    ok = ~np.isnan(got[i])
                    assert got[i][ok].min() >= data[i].min() and got[i][ok].max() <= data[i].max()
    temps = data.copy()
    temps[1] = 280.0 + 5.0 * temps[1]
    out, _, st = engine.neighbourhood(_dev(temps), cells)
    engine.check_status(st)
    exp = _oracle(temps, None, cells)
    _close(out.cpu().numpy()[[0]], exp[[0]])
    np.testing.assert_allclose(out.cpu().numpy()[1], exp[1], rtol=2e-7)
    gam = rng.gamma(2.0, 1.5, shape).astype(np.float32)
    thr = [(2.0, 1.4, 2.6, ">"), (1.0, 1.0, 1.0, ">=")]
    out, _, st = engine.neighbourhood(_dev(gam), cells, thresholds=thr)
    engine.check_status(st)
    for k, (v, lo, hi, op) in enumerate(thr):
        _close(out[:, k].cpu().numpy(), _oracle(orc.truth_value(gam, v, (lo, hi), op), None, cells))
    const = np.full((2,) + shape[-2:], 0.3, dtype=np.float32)
    out, _, st = engine.neighbourhood(_dev(const), cells)
    engine.check_status(st)
    np.testing.assert_array_equal(out.cpu().numpy(), const)


@pytest.mark.parametrize("cells", [3, 5, 12])
def test_periodic_x_two_pass_many_slices(cells):
    """wrap_x launches take the two-pass kernels at every radius (the box kernel has no periodic
    x): many slices (the intermediate plane must be part of the workspace whatever the radius)
    against single-slice launches and against the oracle on the wrap-padded field."""
    from improver_b200 import engine

    rng = np.random.default_rng(17 + cells)
    data = rng.random((40, 96, 160)).astype(np.float32)
    mask = (rng.random((96, 160)) < 0.8).astype(np.uint8)
    dev = _dev(data)
    for m in (None, mask):
        md = _dev(m) if m is not None else None
        out, _, st = engine.neighbourhood(dev, cells, wrap_x=True, mask2d=md)
        engine.check_status(st)
        for k in (0, 17, 39):
            one, _, st = engine.neighbourhood(dev[k:k + 1].contiguous(), cells, wrap_x=True, mask2d=md)
            engine.check_status(st)
            assert torch.equal(out[k].nan_to_num(-1.0), one[0].nan_to_num(-1.0))
        padded = np.pad(data[:3], ((0, 0), (0, 0), (cells, cells)), mode="wrap")
        pm = np.pad(m, ((0, 0), (cells, cells)), mode="wrap") if m is not None else None
        exp, _ = orc.neighbourhood_cube(padded, pm, method="square", cells=cells)
        exp = exp[..., cells:-cells]
        got = out[:3].cpu().numpy()
        inner = slice(cells, -cells)                         # rows away from the top / bottom trimming quirk
        ok = ~np.isnan(exp[:, inner])
        np.testing.assert_allclose(got[:, inner][ok], exp[:, inner][ok], atol=1e-5)
