"""Edge cases of the CUDA path: odd grid widths (scalar loads), tiny grids,
radii larger than the grid, member counts that do not fill a register chunk,
more members than one chunk, more thresholds than one launch group, every
comparison operator, negative thresholds, empty / invalid arguments."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import nbhood_oracle as orc  # noqa: E402


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _close(got, exp, atol=1e-5):
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.shape == exp.shape
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    ok = ~np.isnan(exp)
    np.testing.assert_allclose(got[ok], exp[ok], atol=atol)


@pytest.mark.parametrize("shape", [(7, 9), (1, 33), (33, 1), (5, 101), (64, 67), (3, 3), (130, 258)])
@pytest.mark.parametrize("cells", [1, 4, 9])
@pytest.mark.parametrize("method", ["square", "circular"])
def test_odd_and_tiny_grids(shape, cells, method):
    from improver_b200 import engine

    rng = np.random.default_rng(shape[0] * 1000 + shape[1] + cells)
    data = rng.random((2,) + shape).astype(np.float32)
    data[1] = (data[1] > 0.7).astype(np.float32)
    mask = (rng.random(shape) < 0.75).astype(np.uint8)
    for m in (None, mask):
        out, omask, st = engine.neighbourhood(_dev(data), cells, method=method,
                                              mask2d=_dev(m) if m is not None else None, want_mask=True)
        engine.check_status(st)
        exp, emask = orc.neighbourhood_cube(data, m, method=method, cells=cells)
        _close(out.cpu().numpy(), exp)
        np.testing.assert_array_equal(omask.cpu().numpy().astype(bool), emask)


@pytest.mark.parametrize("n_members", [2, 5, 17, 18, 19, 36, 40])
def test_member_counts(n_members):
    """Partial register chunks (R < 18), exactly one chunk, several chunks (R > 18)."""
    from improver_b200 import engine

    rng = np.random.default_rng(n_members)
    data = rng.gamma(2.0, 1.5, (n_members, 2, 40, 66)).astype(np.float32)
    exp = orc.threshold_square_mean(data, [2.0], [(1.4, 2.6)], ">", 3)
    out, _, st = engine.neighbourhood(_dev(data), 3, thresholds=[(2.0, 1.4, 2.6, ">")], collapse_members=True)
    engine.check_status(st)
    _close(out.cpu().numpy(), exp)
    exp_c = orc.threshold_square_mean(data, [2.0], [(1.4, 2.6)], ">", 3, method="circular")
    out, _, st = engine.neighbourhood(_dev(data), 3, method="circular", thresholds=[(2.0, 1.4, 2.6, ">")],
                                      collapse_members=True)
    engine.check_status(st)
    _close(out.cpu().numpy(), exp_c)


def test_many_thresholds_and_operators():
    """11 thresholds (> one launch group of 8), all comparison operators, a
    negative threshold with swapped bounds, deterministic and fuzzy."""
    from improver_b200 import engine

    rng = np.random.default_rng(77)
    data = (rng.standard_normal((3, 2, 30, 34)) * 2).astype(np.float32)
    thr = [-2.5, -1.0, -0.5, -0.1, 0.1, 0.3, 0.6, 1.0, 1.5, 2.0, 3.0]
    for op in (">", ">=", "<", "<="):
        for fuzzy in (True, False):
            bounds = orc.fuzzy_bounds_from_factor(thr, 0.8) if fuzzy else [(t, t) for t in thr]
            exp = orc.threshold_square_mean(data, thr, bounds, op, 2)
            out, _, st = engine.neighbourhood(_dev(data), 2, collapse_members=True,
                                              thresholds=[(t, b[0], b[1], op) for t, b in zip(thr, bounds)])
            engine.check_status(st)
            assert tuple(out.shape) == (2, 11, 30, 34)
            _close(out.cpu().numpy(), exp)
            tv_exp, _, _ = orc.threshold_cube(data[0], thr, bounds, op)
            tv, _, st = engine.threshold(_dev(data[0]), [(t, b[0], b[1], op) for t, b in zip(thr, bounds)])
            engine.check_status(st)
            np.testing.assert_array_equal(tv.cpu().numpy(), tv_exp)


def test_equality_operator_and_non_contiguous_input():
    from improver_b200 import engine

    data = np.random.default_rng(3).integers(0, 4, (2, 20, 24)).astype(np.float32)
    tv, _, st = engine.threshold(_dev(data), [(2.0, 2.0, 2.0, "==")])
    engine.check_status(st)
    np.testing.assert_array_equal(tv.cpu().numpy()[0], (data == 2.0).astype(np.float32))
    base = _dev(np.random.default_rng(4).random((2, 20, 48)).astype(np.float32))
    view = base[:, :, ::2]                              # non-contiguous view
    out, _, st = engine.neighbourhood(view, 2)
    engine.check_status(st)
    exp, _ = orc.neighbourhood_cube(view.cpu().numpy(), None, method="square", cells=2)
    _close(out.cpu().numpy(), exp)


def test_invalid_arguments_are_rejected_before_launch():
    from improver_b200 import engine
    from improver_b200._lib import EngineError

    data = _dev(np.zeros((2, 8, 8), np.float32))
    with pytest.raises(EngineError, match="not positive"):
        engine.neighbourhood(data, 0)
    with pytest.raises(ValueError, match="not a valid neighbourhood_method"):
        engine.neighbourhood(data, 1, method="oval")
    with pytest.raises(TypeError, match="float32"):
        engine.neighbourhood(data.double(), 1)
    with pytest.raises(TypeError, match="CUDA tensor"):
        engine.neighbourhood(data.cpu(), 1)
    with pytest.raises(EngineError, match="Threshold must be within bounds"):
        engine.neighbourhood(data, 1, thresholds=[(1.0, 2.0, 3.0, ">")])
    with pytest.raises(ValueError, match="mask2d must have shape"):
        engine.neighbourhood(data, 1, mask2d=_dev(np.ones((4, 4), np.uint8)))
    with pytest.raises(EngineError, match="range"):
        engine.neighbourhood_percentiles(data, 1, [101.0])
    with pytest.raises(ValueError, match="empty"):
        engine.neighbourhood(_dev(np.zeros((0, 8, 8), np.float32)), 1)


def test_inf_and_masked_nan_inputs():
    """+-inf are legal data (reference test 'inf' cases); NaNs under the input mask are ignored."""
    from improver_b200 import engine

    data = np.random.default_rng(9).random((1, 16, 18)).astype(np.float32) * 4
    data[0, 3, 3] = np.inf
    data[0, 5, 7] = -np.inf
    bounds = (1.4, 2.6)
    exp = orc.threshold_square_mean(data[None], [2.0], [bounds], ">", 2)
    out, _, st = engine.neighbourhood(_dev(data[None]), 2, thresholds=[(2.0, 1.4, 2.6, ">")],
                                      collapse_members=True)
    engine.check_status(st)
    _close(out.cpu().numpy(), exp)
    m = np.zeros(data.shape, bool)
    m[0, 2, 2] = True
    d2 = data.copy()
    d2[0, 2, 2] = np.nan
    d2[0, 3, 3] = 1.0
    d2[0, 5, 7] = 1.0
    out, omask, st = engine.neighbourhood(_dev(d2), 2, mask_nd=_dev(m.astype(np.uint8)), want_mask=True)
    engine.check_status(st)
    res = orc.neighbourhood(np.ma.masked_array(d2[0], m[0]), None, "square", 2)
    _close(out.cpu().numpy()[0], np.ma.getdata(res))
    np.testing.assert_array_equal(omask.cpu().numpy()[0].astype(bool), np.ma.getmaskarray(res))


@pytest.mark.parametrize("cells", [3, 20])
def test_circular_wrap_x_extension(cells):
    """Periodic x for circular neighbourhoods (config 5 extension; oracle =
    reference semantics on the wrap-padded field)."""
    from improver_b200 import engine

    rng = np.random.default_rng(41 + cells)
    data = rng.random((2, 70, 96)).astype(np.float32)
    mask = (rng.random((70, 96)) < 0.8).astype(np.uint8)
    for m in (None, mask):
        out, _, st = engine.neighbourhood(_dev(data), cells, method="circular", wrap_x=True,
                                          mask2d=_dev(m) if m is not None else None)
        engine.check_status(st)
        padded = np.pad(data, ((0, 0), (0, 0), (cells, cells)), mode="wrap")
        pm = np.pad(m, ((0, 0), (cells, cells)), mode="wrap") if m is not None else None
        exp, _ = orc.neighbourhood_cube(padded, pm, method="circular", cells=cells)
        _close(out.cpu().numpy(), exp[..., cells:-cells])


@pytest.mark.parametrize("shape,cells", [((2, 1, 4, 6), 1), ((11, 2, 9, 64), 4), ((64, 1, 24, 40), 2),
                                         ((3, 7, 5, 1042), 5), ((13, 1, 300, 66), 3), ((2, 150, 12, 34), 2)])
def test_row_stream_kernel_odd_shapes(shape, cells):
    """Row-streaming fused kernel on awkward shapes: grids smaller than the neighbourhood, prime
    member counts (one member per stage), 64 members, more planes than SMs, single-window rows."""
    from improver_b200 import engine

    rng = np.random.default_rng(61)
    data = rng.gamma(2.0, 1.5, shape).astype(np.float32)
    exp = orc.threshold_square_mean(data, [2.0], [(1.4, 2.6)], ">", cells)
    out, _, st = engine.neighbourhood(_dev(data), cells, thresholds=[(2.0, 1.4, 2.6, ">")], collapse_members=True)
    engine.check_status(st)
    np.testing.assert_allclose(out.cpu().numpy(), exp, atol=1e-5)
